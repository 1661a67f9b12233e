// parm_b200 drop-in for ParM's src/trackers.hpp: StateTracker and NeighborList (trackers.hpp:18-31,
// 157-214). The list lives on the device (skin-drift trigger + cell-list build, see
// parm_b200/csrc/nlist.cu); `begin()/end()/get(i)` materialise the reference's `vector<IDPair>`
// (same pair set, same (i asc, j<i asc) order) on demand for host-side consumers.
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

#endif
