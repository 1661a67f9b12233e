// parm_b200 drop-in for ParM's src/trackers.hpp: StateTracker and NeighborList (trackers.hpp:18-31,
// 157-214). The list lives on the device (skin-drift trigger + cell-list build, see
// parm_b200/csrc/nlist.cu); `begin()/end()/get(i)` materialise the reference's `vector<IDPair>`
// (same pair set, same (i asc, j<i asc) order) on demand for host-side consumers.
#include <set>
#include "box.hpp"

#ifndef PARM_B200_TRACKERS_H
#define PARM_B200_TRACKERS_H

class StateTracker {
   public:
    virtual void update(Box &box) = 0;
    virtual bool every_collision() { return false; }
    virtual void update_collision(Box &, AtomID, AtomID, flt, Vec) {}
    virtual ~StateTracker() {}
};

class NeighborList : public StateTracker {
   protected:
    sptr<Box> box;
    flt skin;
    sptr<AtomVec> atoms;
    parm_nlist *nl;
    vector<flt> diameters;  // by AtomVec index, < 0: never add()ed
    bool diam_dirty;
    vector<IDPair> curpairs;  // host copy, valid while pairs_for == which()
    uint pairs_for;

    void flush() {
        if (diam_dirty) {
            parm_b200::check(parm_nlist_set_diameters(nl, diameters.data()));
            diam_dirty = false;
        }
    }
    void fetch_pairs() {
        uint w = which();
        if (pairs_for == w && w != 0) return;
        uint64_t np = 0;
        parm_b200::check(parm_nlist_numpairs(nl, &np));
        vector<uint32_t> a(np ? np : 1), b(np ? np : 1);
        if (np) parm_b200::check(parm_nlist_download_pairs(nl, a.data(), b.data(), np));
        curpairs.clear();
        curpairs.reserve(np);
        AtomVec &av = *atoms;
        for (uint64_t k = 0; k < np; k++) curpairs.push_back(IDPair(av.get_id(a[k]), av.get_id(b[k])));
        pairs_for = w;
    }

   public:
    NeighborList(sptr<Box> box, sptr<AtomVec> atomv, const flt skin)
        : box(box), skin(skin), atoms(atomv), nl(NULL), diameters(atomv->size(), -1.0), diam_dirty(false), pairs_for(0) {
        sptr<OriginBox> ob = boost::dynamic_pointer_cast<OriginBox>(box);
        if (!ob) throw std::runtime_error("parm_b200: NeighborList needs an OriginBox (other boxes are out of scope)");
        ob->attach(atoms->context());
        parm_b200::check(parm_nlist_create(atoms->context(), skin, &nl));
    }
    ~NeighborList() { parm_nlist_destroy(nl); }

    void update(Box &newbox) {
        assert(&newbox == box.get());
        update_list(false);
    }
    bool update_list(bool force = true) {
        flush();
        int rebuilt = 0;
        atoms->device();  // a rebuild re-orders the device arrays; the mirror is refreshed lazily
        parm_b200::check(parm_nlist_update(nl, force ? 1 : 0, &rebuilt));
        return rebuilt != 0;
    }
    AtomVec &vec() { return *atoms; }
    inline uint which() {
        uint32_t u = 0;
        parm_b200::check(parm_nlist_which(nl, &u));
        return u;
    }
    inline uint numpairs() {
        uint64_t np = 0;
        parm_b200::check(parm_nlist_numpairs(nl, &np));
        return (uint)np;
    }
    inline void ignore(AtomID a, AtomID b) {  // trackers.hpp:190-193
        uint32_t ia = a.n(), ib = b.n();
        parm_b200::check(parm_nlist_ignore(nl, &ia, &ib, 1));
    }
    void add(AtomID a, flt diameter) {
        if (a.n() >= diameters.size()) throw std::invalid_argument("NeighborList::add: AtomID is not from this AtomVec");
        if (diameters[a.n()] >= 0)
            throw std::invalid_argument("Cannot add AtomID to SubGroup: it already exists.");  // box.hpp:498-500
        diameters[a.n()] = diameter;
        diam_dirty = true;
    }
    inline uint ignore_size() const {
        uint64_t k = 0;
        parm_b200::check(parm_nlist_ignore_size(nl, &k));
        return (uint)k;
    }
    inline uint size() const {
        uint k = 0;
        for (size_t i = 0; i < diameters.size(); i++) k += diameters[i] >= 0;
        return k;
    }
    inline vector<IDPair>::iterator begin() { fetch_pairs(); return curpairs.begin(); }
    inline vector<IDPair>::iterator end() { fetch_pairs(); return curpairs.end(); }
    inline IDPair get(uint i) { fetch_pairs(); return curpairs[i]; }
    // device-side handles for the other facade classes
    parm_nlist *handle() { flush(); return nl; }
    sptr<AtomVec> atomvec() { return atoms; }
};

// Grid (trackers.hpp:227-309, trackers.cpp:97-219): cells no narrower than any atom; the atoms of a cell and of its
// neighbouring cells are an atom's candidate neighbours. make_grid() bins every atom ON THE DEVICE from the resident
// positions (parm_grid_locs: Grid::get_loc for all atoms at once, 4 bytes per atom come back) and keeps the cells as
// sorted index vectors (the reference's set<AtomID> orders by index as well). The event-driven collections that drive
// time_to_edge / GridIterator step by step are out of scope; all_pairs() / all_pairs(AtomID)
// give the same sets as the reference's iterators.
class Grid {
   public:
    sptr<OriginBox> box;
    sptr<AtomVec> atoms;
    flt minwidth, goalwidth;
    uint widths[NDIM];
    vector<vector<uint> > gridlocs;  // atoms (AtomVec indices, ascending) of every cell
    vector<uint32_t> locs;           // cell of every atom, as of the last make_grid()

    Grid(sptr<OriginBox> box, sptr<AtomVec> atoms, const uint width = 1) : box(box), atoms(atoms), minwidth(-1), goalwidth(-1) {
        for (uint d = 0; d < NDIM; d++) widths[d] = width;
    }
    Grid(sptr<OriginBox> box, sptr<AtomVec> atoms, vector<uint> width) : box(box), atoms(atoms), minwidth(-1), goalwidth(-1) {
        assert(width.size() == NDIM);
        for (uint d = 0; d < NDIM; d++) widths[d] = width[d];
    }
    Grid(sptr<OriginBox> box, sptr<AtomVec> atoms, const flt minwidth, const flt goalwidth)
        : box(box), atoms(atoms), minwidth(minwidth), goalwidth(goalwidth) {
        for (uint d = 0; d < NDIM; d++) widths[d] = 1;
        optimize_widths();
    }
    void optimize_widths() {  // trackers.cpp:168-190
        if (minwidth <= 0) return;
        flt wpa = pow(box->V() * goalwidth / atoms->size(), OVERNDIM);
        if (wpa < minwidth) wpa = minwidth;
        Vec b = box->box_shape();
        bool small = false;
        for (uint d = 0; d < NDIM; d++) {
            widths[d] = (uint)floor(b[d] / wpa);
            small = small || widths[d] < 3;
        }
        if (small)
            for (uint d = 0; d < NDIM; d++) widths[d] = 1;
    }
    uint numcells(uint i) { assert(i < NDIM); return widths[i]; }
    uint numcells() {
        uint k = 1;
        for (uint d = 0; d < NDIM; d++) k *= widths[d];
        return k;
    }
    void make_grid() {
        box->attach(atoms->context());
        uint32_t w[3] = {1, 1, 1};
        for (uint d = 0; d < NDIM; d++) w[d] = widths[d];
        locs.assign(atoms->size() ? atoms->size() : 1, 0);
        parm_b200::check(parm_grid_locs(atoms->device(false), w, locs.data()));
        locs.resize(atoms->size());
        gridlocs.assign(numcells(), vector<uint>());
        for (uint i = 0; i < atoms->size(); i++) gridlocs[locs[i]].push_back(i);
    }
    uint get_loc(Vec v, Vec bsize) {  // trackers.cpp:192-206
        v = vec_mod(v - bsize / 2., bsize) + bsize / 2;
        uint k[3] = {0, 0, 0};
        for (uint d = 0; d < NDIM; d++) {
            k[d] = (uint)floor(v[d] * widths[d] / bsize[d]);
            if (k[d] == widths[d]) k[d] = 0;
        }
        return NDIM == 3 ? (k[2] * widths[1] + k[1]) * widths[0] + k[0] : k[1] * widths[0] + k[0];
    }
    vector<uint> neighbors(uint i) {  // trackers.cpp:125-166: the 3^NDIM cells around (and including) cell i, periodic
        vector<uint> v;
        uint w0 = widths[0], w1 = widths[1];
        uint x = i % w0, y = (i / w0) % w1, z = NDIM == 3 ? i / (w0 * w1) : 0;
        for (int dz = (NDIM == 3 ? -1 : 0); dz <= (NDIM == 3 ? 1 : 0); dz++)
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    uint zz = NDIM == 3 ? (uint)((int)(z + widths[NDIM - 1]) + dz) % widths[NDIM - 1] : 0;
                    uint yy = (uint)((int)(y + w1) + dy) % w1, xx = (uint)((int)(x + w0) + dx) % w0;
                    v.push_back((zz * w1 + yy) * w0 + xx);
                }
        return v;
    }
    vector<AtomID> all_pairs(AtomID a) {
        vector<AtomID> out;
        std::set<uint> seen;
        vector<uint> nb = neighbors(locs[a.n()]);
        for (size_t c = 0; c < nb.size(); c++) {
            if (!seen.insert(nb[c]).second) continue;  // (an axis with fewer than 3 cells lists a cell more than once)
            for (size_t k = 0; k < gridlocs[nb[c]].size(); k++)
                if (gridlocs[nb[c]][k] != a.n()) out.push_back(atoms->get_id(gridlocs[nb[c]][k]));
        }
        return out;
    }
    vector<IDPair> all_pairs() {
        std::set<std::pair<uint, uint> > ps;
        for (uint c = 0; c < gridlocs.size(); c++) {
            vector<uint> nb = neighbors(c);
            for (size_t q = 0; q < nb.size(); q++)
                for (size_t a = 0; a < gridlocs[c].size(); a++)
                    for (size_t b = 0; b < gridlocs[nb[q]].size(); b++)
                        if (gridlocs[nb[q]][b] < gridlocs[c][a]) ps.insert(std::make_pair(gridlocs[c][a], gridlocs[nb[q]][b]));
        }
        vector<IDPair> out;
        for (std::set<std::pair<uint, uint> >::iterator it = ps.begin(); it != ps.end(); ++it)
            out.push_back(IDPair(atoms->get_id(it->first), atoms->get_id(it->second)));
        return out;
    }
};

#endif
