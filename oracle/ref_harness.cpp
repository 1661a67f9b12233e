// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or executed from
// the product path (parm_b200/). Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load the library this
// file is built into.
//
// C-ABI harness around the UNMODIFIED reference classes. It is compiled
// together with /root/reference/src/{vecrand,box,trackers,interaction,
// collection}.cpp (in place, never copied) against oracle/shim/ into
// oracle/_ref/libparm_ref{2,3}d.so by oracle/Makefile. Every physics result
// returned here is computed by the reference's own code:
//   OriginBox::diff                      box.hpp:97-104
//   NeighborList::update_list            trackers.cpp:19-85
//   NListed<A,P>::{set_forces,energy,pressure,stress,...}  interaction.hpp:2102-2291
//   CollectionVerlet::timestep           collection.cpp:442-469
//   CollectionSol::timestep              collection.cpp:265-322
//   Collection::{energy,pressure,temp,...}   collection.cpp:21-142
//
// The only non-reference logic is InjectedNeighborList (large-N pair finder,
// SURVEY 8c): a CPU cell list that evaluates the reference's own predicate
// (box->diff(...).norm() < diam + skin) and emits pairs in the reference's
// (i ascending, j<i ascending) order; tests prove it identical to
// update_list(true) at small N.
#include <cstdint>
#include <cstring>
#include <string>

#include "collection.hpp"
#include "interaction.hpp"

extern "C" {
// the shim's normal_distribution reads from here when non-null (noise injection
// for CollectionSol parity; see oracle/shim/boost/random/normal_distribution.hpp)
const double *parm_oracle_noise = 0;
size_t parm_oracle_noise_len = 0;
size_t parm_oracle_noise_pos = 0;
}

namespace {

struct FastSubGroup : public SubGroup {
    // SubGroup::add is O(size) (std::find duplicate check, box.hpp:495-501);
    // for 1e6 atoms that is 5e11 compares. Skips only the duplicate check.
    void add_fast(AtomID a) { ids.push_back(a); }
};

class InjectedNeighborList : public NeighborList {
   public:
    sptr<OriginBox> obox;
    bool injected;
    InjectedNeighborList(sptr<OriginBox> b, sptr<AtomVec> av, flt skin, bool injected)
        : NeighborList(boost::static_pointer_cast<Box>(b), av, skin), obox(b), injected(injected) {}

    void add_fast(AtomID a, flt diameter) {
        static_cast<FastSubGroup &>(atoms).add_fast(a);
        diameters.push_back(diameter);
        lastlocs.push_back(a->x);
        ignorechanged = true;
    }

    // reference drift rule, trackers.cpp:23-53, then cell-list build
    bool update_list_cells(bool force) {
        if (!injected) return update_list(force);
        if (not force and not ignorechanged) {
            flt bigdist = 0, biggestdist = 0;
            for (uint i = 0; i < atoms.size(); i++) {
                Atom &atm = atoms[i];
                flt curdist = (atm.x - lastlocs[i]).norm();
                if (curdist > biggestdist) {
                    bigdist = biggestdist;
                    biggestdist = curdist;
                } else if (curdist > bigdist) {
                    bigdist = curdist;
                } else
                    continue;
                if (bigdist + biggestdist >= skin) {
                    force = true;
                    break;
                }
            }
            if (not force) return false;
        }
        updatenum++;
        ignorechanged = false;
        curpairs.clear();
        const uint N = atoms.size();
        Vec L = obox->box_shape();
        flt maxd = 0;
        for (uint i = 0; i < N; i++) maxd = max(maxd, diameters[i]);
        flt rc = (maxd + skin) * (1 + 1e-9) + 1e-9;
        int nc[NDIM];
        bool small = false;
        size_t ncell = 1;
        for (uint d = 0; d < NDIM; d++) {
            nc[d] = (int)floor(L[d] / rc);
            if (nc[d] < 3) small = true;
            if (nc[d] > 256) nc[d] = 256;
            ncell *= (size_t)(nc[d] < 1 ? 1 : nc[d]);
        }
        if (small) {  // box too small for a 3^D stencil: reference loop
            for (uint i = 0; i < N; i++) {
                AtomID a1 = atoms.get_id(i);
                lastlocs[i] = a1->x;
                for (uint j = 0; j < i; j++) {
                    AtomID a2 = atoms.get_id(j);
                    if (ignorepairs.has_pair(a1, a2)) continue;
                    flt diam = (diameters[i] + diameters[j]) / 2;
                    if (box->diff(a1->x, a2->x).norm() < (diam + skin)) curpairs.push_back(IDPair(a1, a2));
                }
            }
            return true;
        }
        vector<int> cidx(N * NDIM);
        vector<uint> cellstart(ncell + 1, 0), cellatoms(N);
        vector<size_t> cellof(N);
        for (uint i = 0; i < N; i++) {
            Atom &a = atoms[i];
            lastlocs[i] = a.x;
            size_t c = 0;
            for (uint d = 0; d < NDIM; d++) {
                flt w = a.x[d] - L[d] * floor(a.x[d] / L[d]);
                int k = (int)floor(w / L[d] * nc[d]);
                if (k < 0) k = 0;
                if (k >= nc[d]) k = nc[d] - 1;
                cidx[i * NDIM + d] = k;
                c = c * nc[d] + k;
            }
            cellof[i] = c;
            cellstart[c + 1]++;
        }
        for (size_t c = 0; c < ncell; c++) cellstart[c + 1] += cellstart[c];
        {
            vector<uint> fill(cellstart.begin(), cellstart.end() - 1);
            for (uint i = 0; i < N; i++) cellatoms[fill[cellof[i]]++] = i;  // ascending i inside a cell
        }
        vector<uint> js;
        const bool has_ignored = ignorepairs.size() > 0;  // has_pair() inserts map nodes: skip it when nothing is ignored
        int nst = 1;
        for (uint d = 0; d < NDIM; d++) nst *= 3;
        for (uint i = 0; i < N; i++) {
            AtomID a1 = atoms.get_id(i);
            js.clear();
            for (int s = 0; s < nst; s++) {
                size_t c = 0;
                int t = s;
                int off[NDIM];
                for (int d = NDIM - 1; d >= 0; d--) { off[d] = t % 3 - 1; t /= 3; }
                for (uint d = 0; d < NDIM; d++) {
                    int k = (cidx[i * NDIM + d] + off[d] + nc[d]) % nc[d];
                    c = c * nc[d] + k;
                }
                for (uint q = cellstart[c]; q < cellstart[c + 1]; q++) {
                    uint j = cellatoms[q];
                    if (j >= i) break;  // ascending within a cell
                    AtomID a2 = atoms.get_id(j);
                    if (has_ignored && ignorepairs.has_pair(a1, a2)) continue;
                    flt diam = (diameters[i] + diameters[j]) / 2;
                    if (box->diff(a1->x, a2->x).norm() < (diam + skin)) js.push_back(j);
                }
            }
            std::sort(js.begin(), js.end());
            for (size_t q = 0; q < js.size(); q++) curpairs.push_back(IDPair(a1, atoms.get_id(js[q])));
        }
        return true;
    }
    void update(Box &newbox) {
        assert(&newbox == box.get());
        update_list_cells(false);
    }
    uint index_of(const AtomID &a) { return a.n(); }
};

template <class A, class P>
class FastNListed : public NListed<A, P> {
   public:
    FastNListed(sptr<AtomVec> vec, sptr<NeighborList> nl) : NListed<A, P>(vec, nl) {}
    void add_fast(A atm, InjectedNeighborList *inl) {
        if (inl->injected) {
            inl->add_fast(atm, atm.max_size());
            this->atoms[atm.n()] = atm;
        } else {
            this->add(atm);
        }
    }
};

struct Sys {
    vector<sptr<StateTracker> > stats;  // RsqTracker / ISFTracker / EnergyTracker, in creation order
    sptr<OriginBox> box;
    sptr<AtomVec> atoms;
    vector<sptr<Interaction> > inters;
    vector<sptr<InjectedNeighborList> > nls;
    sptr<Collection> collec;
    std::string err;
};

Vec vec_from(const double *p) {
    Vec v;
    for (uint d = 0; d < NDIM; d++) v[d] = p[d];
    return v;
}

}  // namespace

extern "C" {

int ref_ndim() { return NDIM; }
size_t ref_sizeof_atom() { return sizeof(Atom); }

void *ref_sys_create(uint32_t n, const double *L, const double *x, const double *v, const double *m) {
    Sys *s = new Sys();
    s->box.reset(new OriginBox(vec_from(L)));
    vector<double> masses(m, m + n);
    s->atoms.reset(new AtomVec(masses));
    AtomVec &av = *s->atoms;
    for (uint i = 0; i < n; i++) {
        av[i].x = vec_from(x + (size_t)i * NDIM);
        if (v) av[i].v = vec_from(v + (size_t)i * NDIM);
    }
    return s;
}

void ref_sys_destroy(void *h) { delete static_cast<Sys *>(h); }

}  // extern "C"
// Adds every (member) atom i as A(make(i)) to a new FastNListed<A,P>.
template <class A, class P, class MK>
static void add_all(Sys *s, sptr<NeighborList> nlbase, InjectedNeighborList *nl, const uint8_t *member, MK make) {
    sptr<FastNListed<A, P> > I(new FastNListed<A, P>(s->atoms, nlbase));
    const uint n = s->atoms->size();
    for (uint i = 0; i < n; i++)
        if (!member || member[i]) I->add_fast(make(i), nl);
    s->inters.push_back(I);
}

extern "C" {
// kind and per-atom parameter layout: include/parm_b200.h (parm_inter_set_params_ex). params: n x nper
// row-major; type[i] = indx; eps_table / sig_table: ntypes x ntypes, row t = the `epsilons` / `sigmas`
// vector of atoms with indx t (NULL = that quantity is not indexed, which selects the A struct).
// member: optional n bytes, 0 = atom not added.
// injected: 0 = reference O(N^2) NeighborList, 1 = cell-list InjectedNeighborList
// share_nl: -1 = new NeighborList, else index of an existing list to share
int ref_add_interaction_ex(void *h, int kind, double skin, const double *params, int nper, const uint32_t *type,
                           const double *eps_table, const double *sig_table, int ntypes, const uint8_t *member,
                           int injected, int share_nl) {
    Sys *s = static_cast<Sys *>(h);
    AtomVec &av = *s->atoms;
    // an epsilon table selects the indexed atom structs; IEpsISig* need the sigma table as well
    if ((kind == 1 || kind == 3) && eps_table && !sig_table) return -3;
    sptr<InjectedNeighborList> nl;
    if (share_nl >= 0)
        nl = s->nls[share_nl];
    else {
        nl.reset(new InjectedNeighborList(s->box, s->atoms, skin, injected != 0));
        s->nls.push_back(nl);
    }
    sptr<NeighborList> nlbase = boost::static_pointer_cast<NeighborList>(nl);
    InjectedNeighborList *inl = nl.get();
    auto P = [&](uint i, int q) { return params[(size_t)nper * i + q]; };
    auto T = [&](uint i) { return type ? type[i] : 0u; };
    auto row = [&](const double *tab, uint t) { return vector<flt>(tab + (size_t)t * ntypes, tab + (size_t)(t + 1) * ntypes); };
    try {
        if (kind == 0) {
            add_all<EpsSigAtom, LJRepulsePair>(s, nlbase, inl, member, [&](uint i) { return EpsSigAtom(av.get_id(i), P(i, 0), P(i, 1)); });
        } else if (kind == 1 && !eps_table) {
            add_all<EpsSigExpAtom, RepulsionPair>(s, nlbase, inl, member, [&](uint i) { return EpsSigExpAtom(av.get_id(i), P(i, 0), P(i, 1), P(i, 2)); });
        } else if (kind == 1) {
            add_all<IEpsISigExpAtom, RepulsionPair>(s, nlbase, inl, member, [&](uint i) {
                return IEpsISigExpAtom(av.get_id(i), row(eps_table, T(i)), row(sig_table, T(i)), T(i), P(i, 2)); });
        } else if (kind == 2) {
            add_all<IEpsSigCutAtom, LJAttractRepulsePair>(s, nlbase, inl, member, [&](uint i) {
                return IEpsSigCutAtom(av.get_id(i), row(eps_table, T(i)), T(i), P(i, 1), P(i, 2)); });
        } else if (kind == 3 && !eps_table) {
            add_all<EpsSigCutAtom, LennardJonesCutPair>(s, nlbase, inl, member, [&](uint i) { return EpsSigCutAtom(av.get_id(i), P(i, 0), P(i, 1), P(i, 2)); });
        } else if (kind == 3) {
            add_all<IEpsISigCutAtom, LennardJonesCutPair>(s, nlbase, inl, member, [&](uint i) {
                return IEpsISigCutAtom(av.get_id(i), row(eps_table, T(i)), row(sig_table, T(i)), T(i), P(i, 2)); });
        } else if (kind == 4 && !eps_table) {
            add_all<EpsSigCutAtom, LJAttractCutPair>(s, nlbase, inl, member, [&](uint i) { return EpsSigCutAtom(av.get_id(i), P(i, 0), P(i, 1), P(i, 2)); });
        } else if (kind == 4 && !sig_table) {
            add_all<IEpsSigCutAtom, LJAttractCutPair>(s, nlbase, inl, member, [&](uint i) {
                return IEpsSigCutAtom(av.get_id(i), row(eps_table, T(i)), T(i), P(i, 1), P(i, 2)); });
        } else if (kind == 4) {
            add_all<IEpsISigCutAtom, LJAttractCutPair>(s, nlbase, inl, member, [&](uint i) {
                return IEpsISigCutAtom(av.get_id(i), row(eps_table, T(i)), row(sig_table, T(i)), T(i), P(i, 2)); });
        } else if (kind == 5) {
            add_all<IEpsRepsSigCutAtom, LJAttractFixedRepulsePair>(s, nlbase, inl, member, [&](uint i) {
                return IEpsRepsSigCutAtom(av.get_id(i), row(eps_table, T(i)), P(i, 3), P(i, 1), T(i), P(i, 2)); });
        } else if (kind == 6) {
            add_all<EisMclachlanAtom, EisMclachlanPair>(s, nlbase, inl, member, [&](uint i) { return EisMclachlanAtom(av.get_id(i), P(i, 1), P(i, 0)); });
        } else if (kind == 7) {
            add_all<IEpsRepsSigExpCutAtom, LJishPair>(s, nlbase, inl, member, [&](uint i) {
                return IEpsRepsSigExpCutAtom(av.get_id(i), row(eps_table, T(i)), P(i, 3), P(i, 1), P(i, 4), T(i), P(i, 2)); });
        } else if (kind == 8) {
            add_all<EpsEpsSigSigCutAtom, LJAttractRepulseSigsPair>(s, nlbase, inl, member, [&](uint i) {
                return EpsEpsSigSigCutAtom(av.get_id(i), P(i, 0), P(i, 3), P(i, 1), P(i, 4), P(i, 2)); });
        } else if (kind == 9) {
            add_all<EpsSigExpDragAtom, RepulsionDragPair>(s, nlbase, inl, member, [&](uint i) {
                return EpsSigExpDragAtom(av.get_id(i), P(i, 0), P(i, 1), P(i, 3), P(i, 2)); });
        } else if (kind == 10 || kind == 12) {
            // the LoisOhernAtom ctor stores its arguments unchanged
            if (kind == 10)
                add_all<LoisOhernAtom, LoisOhernPair>(s, nlbase, inl, member, [&](uint i) { return LoisOhernAtom(av.get_id(i), P(i, 0), P(i, 1), P(i, 2), P(i, 3)); });
            else
                add_all<LoisOhernAtom, LoisOhernPairMinCLs>(s, nlbase, inl, member, [&](uint i) { return LoisOhernAtom(av.get_id(i), P(i, 0), P(i, 1), P(i, 2), P(i, 3)); });
        } else if (kind == 11 || kind == 13) {
            // LoisLinAtom(a, eps, sigma, depth, width) stores f = depth/width: pass depth = f*width back in
            auto mk = [&](uint i) {
                LoisLinAtom a(av.get_id(i), P(i, 0), P(i, 1), 0.0, P(i, 3));
                a.f = P(i, 2);
                return a;
            };
            if (kind == 11)
                add_all<LoisLinAtom, LoisLinPair>(s, nlbase, inl, member, mk);
            else
                add_all<LoisLinAtom, LoisLinPairMin>(s, nlbase, inl, member, mk);
        } else
            return -1;
    } catch (std::exception &e) {
        s->err = e.what();
        return -2;
    }
    return (int)s->inters.size() - 1;
}

int ref_add_interaction(void *h, int kind, double skin, const double *params, const uint32_t *type,
                        const double *eps_table, int ntypes, const uint8_t *member, int injected, int share_nl) {
    return ref_add_interaction_ex(h, kind, skin, params, 3, type, kind == 2 ? eps_table : 0, 0, ntypes, member, injected, share_nl);
}

}  // extern "C"
// NListed<A,P>::contacts / overlaps (interaction.hpp:2126-2151) are not virtuals of Interaction: dispatch on the type
template <class A, class P>
static bool try_contacts(Interaction *I, Box &box, unsigned long long *c, unsigned long long *o) {
    NListed<A, P> *n = dynamic_cast<NListed<A, P> *>(I);
    if (!n) return false;
    *c = n->contacts(box);
    *o = n->overlaps(box);
    return true;
}
extern "C" {
int ref_inter_contacts(void *h, int k, unsigned long long *c, unsigned long long *o) {
    Sys *s = static_cast<Sys *>(h);
    Interaction *I = s->inters[k].get();
    Box &b = *s->box;
    return (try_contacts<EpsSigAtom, LJRepulsePair>(I, b, c, o) || try_contacts<EpsSigExpAtom, RepulsionPair>(I, b, c, o) ||
            try_contacts<IEpsISigExpAtom, RepulsionPair>(I, b, c, o) || try_contacts<IEpsSigCutAtom, LJAttractRepulsePair>(I, b, c, o) ||
            try_contacts<EpsSigCutAtom, LennardJonesCutPair>(I, b, c, o) || try_contacts<IEpsISigCutAtom, LennardJonesCutPair>(I, b, c, o) ||
            try_contacts<EpsSigCutAtom, LJAttractCutPair>(I, b, c, o) || try_contacts<IEpsSigCutAtom, LJAttractCutPair>(I, b, c, o) ||
            try_contacts<IEpsISigCutAtom, LJAttractCutPair>(I, b, c, o) ||
            try_contacts<IEpsRepsSigCutAtom, LJAttractFixedRepulsePair>(I, b, c, o) ||
            try_contacts<EisMclachlanAtom, EisMclachlanPair>(I, b, c, o) || try_contacts<IEpsRepsSigExpCutAtom, LJishPair>(I, b, c, o) ||
            try_contacts<EpsEpsSigSigCutAtom, LJAttractRepulseSigsPair>(I, b, c, o) ||
            try_contacts<EpsSigExpDragAtom, RepulsionDragPair>(I, b, c, o) || try_contacts<LoisOhernAtom, LoisOhernPair>(I, b, c, o) ||
            try_contacts<LoisOhernAtom, LoisOhernPairMinCLs>(I, b, c, o) || try_contacts<LoisLinAtom, LoisLinPair>(I, b, c, o) ||
            try_contacts<LoisLinAtom, LoisLinPairMin>(I, b, c, o))
               ? 0
               : -1;
}

// integrator: 0 CollectionVerlet(dt), 1 CollectionSol(dt, damping, T).
// Mirrors LJatoms.cpp:79-83: construct, add_tracker(nl), add_interaction(I).
int ref_make_collection(void *h, int integrator, double dt, double damping, double T) {
    Sys *s = static_cast<Sys *>(h);
    try {
        sptr<Box> b = boost::static_pointer_cast<Box>(s->box);
        sptr<AtomGroup> ag = boost::static_pointer_cast<AtomGroup>(s->atoms);
        if (integrator == 0)
            s->collec.reset(new CollectionVerlet(b, ag, dt));
        else if (integrator == 1)
            s->collec.reset(new CollectionSol(b, ag, dt, damping, T));
        else
            return -1;
        for (size_t k = 0; k < s->nls.size(); k++) s->collec->add_tracker(boost::static_pointer_cast<StateTracker>(s->nls[k]));
        for (size_t k = 0; k < s->inters.size(); k++) s->collec->add_interaction(s->inters[k]);
    } catch (std::exception &e) {
        s->err = e.what();
        return -2;
    }
    return 0;
}

// integrator types and parameter order: include/parm_b200.h (PARM_INTEG_*, parm_integ_create)
int ref_make_collection_ex(void *h, int type, const double *p, int np) {
    Sys *s = static_cast<Sys *>(h);
    if (type == 0) return ref_make_collection(h, 0, p[0], 0, 0);
    if (type == 1) return ref_make_collection(h, 1, p[0], p[1], p[2]);
    try {
        sptr<Box> b = boost::static_pointer_cast<Box>(s->box);
        sptr<AtomGroup> ag = boost::static_pointer_cast<AtomGroup>(s->atoms);
        switch (type) {
            case 2: s->collec.reset(new CollectionDamped(b, ag, p[0], p[1])); break;
            case 3: s->collec.reset(new CollectionSolHT(b, ag, p[0], p[1], p[2])); break;
            case 4: s->collec.reset(new CollectionOverdamped(b, ag, p[0], np > 1 ? p[1] : 1.0)); break;
            case 5: s->collec.reset(new CollectionNoseHoover(b, ag, p[0], p[1], p[2])); break;
            case 6: s->collec.reset(new CollectionGaussianT(b, ag, p[0])); break;
            case 7: s->collec.reset(new CollectionGear3A(b, ag, p[0])); break;
            case 8: s->collec.reset(new CollectionGear4A(b, ag, p[0], (uint)(np > 1 ? p[1] : 1))); break;
            case 9: s->collec.reset(new CollectionGear5A(b, ag, p[0], (uint)(np > 1 ? p[1] : 1))); break;
            case 10: s->collec.reset(new CollectionGear6A(b, ag, p[0], (uint)(np > 1 ? p[1] : 1))); break;
            case 11:  // CollectionNLCG(box, atoms, dt, P0, {}, {}, {}, kappa, kmax, secmax, seceps), collection.hpp:430-437
                s->collec.reset(new CollectionNLCG(s->box, ag, p[0], p[1], vector<sptr<Interaction> >(),
                                                   vector<sptr<StateTracker> >(), vector<sptr<Constraint> >(),
                                                   np > 2 ? p[2] : 10.0, np > 3 ? p[3] : 1000, (uint)(np > 4 ? p[4] : 40),
                                                   np > 5 ? p[5] : 1e-20));
                break;
            default: return -1;
        }
        for (size_t k = 0; k < s->nls.size(); k++) s->collec->add_tracker(boost::static_pointer_cast<StateTracker>(s->nls[k]));
        for (size_t k = 0; k < s->inters.size(); k++) s->collec->add_interaction(s->inters[k]);
    } catch (std::exception &e) {
        s->err = e.what();
        return -2;
    }
    return 0;
}
}  // extern "C"
// CollectionGaussianT::xi is protected and has no getter: read it through a pointer to member
struct GaussianTXi : public CollectionGaussianT {
    static flt get(CollectionGaussianT &c) { return c.*(&GaussianTXi::xi); }
};
extern "C" {
// thermostat state: out[0] = xi, out[1] = lns (CollectionNoseHoover), out[0] = xi (CollectionGaussianT)
void ref_get_scalars(void *h, double *out) {
    out[0] = out[1] = 0;
    Collection *c = static_cast<Sys *>(h)->collec.get();
    if (CollectionNoseHoover *nh = dynamic_cast<CollectionNoseHoover *>(c)) {
        out[0] = nh->get_xi();
        out[1] = nh->get_lns();
    } else if (CollectionGaussianT *gt = dynamic_cast<CollectionGaussianT *>(c)) {
        out[0] = GaussianTXi::get(*gt);
    }
}

// ---- CollectionNLCG: selectors as in include/parm_b200.h (PARM_NLCG_*) ----
static CollectionNLCG *nlcg_of(void *h) { return dynamic_cast<CollectionNLCG *>(static_cast<Sys *>(h)->collec.get()); }
int ref_nlcg_set(void *h, int which, double v) {
    CollectionNLCG *c = nlcg_of(h);
    if (!c) return -1;
    switch (which) {
        case 0: c->set_dt(v); break;
        case 1: c->set_pressure_goal(v); break;
        case 2: c->set_kappa(v); break;
        case 3: c->set_max_alpha(v); break;
        case 4: c->set_max_alpha_fraction(v); break;
        case 5: c->set_max_dx(v); break;
        case 6: c->set_max_step(v); break;
        case 7: c->maxdV = v; break;
        case 8: c->kmax = v; break;
        case 9: c->secmax = (uint)v; break;
        case 10: c->seceps = v; break;
        default: return -1;
    }
    return 0;
}
int ref_nlcg_get(void *h, double *o) {
    CollectionNLCG *c = nlcg_of(h);
    if (!c) return -1;
    o[0] = c->dt; o[1] = c->P0; o[2] = c->kappa; o[3] = c->Knew; o[4] = c->k; o[5] = c->vl; o[6] = c->fl; o[7] = c->al;
    o[8] = c->alpha; o[9] = c->beta; o[10] = c->betaused; o[11] = c->dxsum; o[12] = c->alphavmax; o[13] = c->sec;
    o[14] = c->kmax; o[15] = c->secmax;
    return 0;
}
int ref_nlcg_set_forces(void *h, int caa, int setV) {
    CollectionNLCG *c = nlcg_of(h);
    if (!c) return -1;
    c->set_forces(caa != 0, setV != 0);
    return 0;
}
int ref_nlcg_reset(void *h) { CollectionNLCG *c = nlcg_of(h); if (!c) return -1; c->reset(); return 0; }
int ref_nlcg_descend(void *h) { CollectionNLCG *c = nlcg_of(h); if (!c) return -1; c->descend(); return 0; }
double ref_nlcg_reduce(void *h, int what) {
    CollectionNLCG *c = nlcg_of(h);
    if (!c) return NAN;
    switch (what) {
        case 0: return c->fdotf();
        case 1: return c->fdota();
        case 2: return c->fdotv();
        case 3: return c->vdotv();
        case 4: return c->kinetic_energy();
        case 5: return c->pressure();
        case 6: return c->hamiltonian();
    }
    return NAN;
}
void ref_get_box(void *h, double *L) {
    Vec s = static_cast<Sys *>(h)->box->box_shape();
    for (uint d = 0; d < NDIM; d++) L[d] = s[d];
}
void ref_set_box(void *h, const double *L) { static_cast<Sys *>(h)->box->resize_to(vec_from(L)); }  // box.cpp:21-25

// ---- statistics trackers (constraints.hpp:260-414); each is add_tracker()ed to the collection, which
// calls update_trackers() once (collection.hpp:117-120) ----
static int push_stat(Sys *s, sptr<StateTracker> t) {
    s->stats.push_back(t);
    if (s->collec) s->collec->add_tracker(t);
    return (int)s->stats.size() - 1;
}
int ref_add_rsq_tracker(void *h, const unsigned long long *ns, int nns, int usecom) {
    Sys *s = static_cast<Sys *>(h);
    vector<unsigned long> v(ns, ns + nns);
    return push_stat(s, sptr<StateTracker>(new RsqTracker(boost::static_pointer_cast<AtomGroup>(s->atoms), v, usecom != 0)));
}
int ref_add_isf_tracker(void *h, const double *ks, int nks, const unsigned long long *ns, int nns, int usecom) {
    Sys *s = static_cast<Sys *>(h);
    vector<unsigned long> v(ns, ns + nns);
    vector<flt> kv(ks, ks + nks);
    return push_stat(s, sptr<StateTracker>(new ISFTracker(boost::static_pointer_cast<AtomGroup>(s->atoms), kv, v, usecom != 0)));
}
int ref_add_energy_tracker(void *h, unsigned n_skip) {
    Sys *s = static_cast<Sys *>(h);
    return push_stat(s, sptr<StateTracker>(new EnergyTracker(boost::static_pointer_cast<AtomGroup>(s->atoms), s->inters, n_skip)));
}
void ref_tracker_update(void *h, int t) { Sys *s = static_cast<Sys *>(h); s->stats[t]->update(*s->box); }
void ref_tracker_reset(void *h, int t) {
    StateTracker *p = static_cast<Sys *>(h)->stats[t].get();
    if (RsqTracker *r = dynamic_cast<RsqTracker *>(p)) r->reset();
    else if (ISFTracker *i = dynamic_cast<ISFTracker *>(p)) i->reset();
    else if (EnergyTracker *e = dynamic_cast<EnergyTracker *>(p)) e->reset();
}
void ref_tracker_counts(void *h, int t, unsigned long long *out, int cap) {
    StateTracker *p = static_cast<Sys *>(h)->stats[t].get();
    vector<flt> c;
    if (RsqTracker *r = dynamic_cast<RsqTracker *>(p)) c = r->counts();
    else if (ISFTracker *i = dynamic_cast<ISFTracker *>(p)) c = i->counts();
    for (int k = 0; k < cap && k < (int)c.size(); k++) out[k] = (unsigned long long)c[k];
}
// xyz2(), xyz4() (n x NDIM row-major) and r4() (n) of lag `single`
void ref_rsq_read(void *h, int t, int single, double *xyz2, double *xyz4, double *r4) {
    RsqTracker *r = dynamic_cast<RsqTracker *>(static_cast<Sys *>(h)->stats[t].get());
    Eigen::Matrix<flt, Eigen::Dynamic, NDIM> a = r->xyz2()[single], b = r->xyz4()[single];
    vector<flt> c = r->r4()[single];
    for (int i = 0; i < a.rows(); i++) {
        for (uint j = 0; j < NDIM; j++) {
            if (xyz2) xyz2[(size_t)i * NDIM + j] = a(i, j);
            if (xyz4) xyz4[(size_t)i * NDIM + j] = b(i, j);
        }
        if (r4) r4[i] = c[i];
    }
}
// ISFxyz() of lag `single`: out[nks][n][NDIM][2]
void ref_isf_read(void *h, int t, int single, double *out) {
    ISFTracker *r = dynamic_cast<ISFTracker *>(static_cast<Sys *>(h)->stats[t].get());
    vector<vector<barray<cmplx, NDIM> > > v = r->ISFxyz()[single];
    size_t q = 0;
    for (size_t ki = 0; ki < v.size(); ki++)
        for (size_t i = 0; i < v[ki].size(); i++)
            for (uint j = 0; j < NDIM; j++) {
                out[q++] = v[ki][i][j].real();
                out[q++] = v[ki][i][j].imag();
            }
}
// out[8] = n(), E(), U(), K(), E_squared_mean(), U_squared_mean(), K_squared_mean(), get_U0()
void ref_energy_tracker_read(void *h, int t, double *out) {
    EnergyTracker *e = dynamic_cast<EnergyTracker *>(static_cast<Sys *>(h)->stats[t].get());
    out[0] = e->n();
    out[1] = e->E(); out[2] = e->U(); out[3] = e->K();
    out[4] = e->E_squared_mean(); out[5] = e->U_squared_mean(); out[6] = e->K_squared_mean();
    out[7] = e->get_U0();
}
void ref_energy_tracker_set_U0(void *h, int t, int from_box, double U0) {
    Sys *s = static_cast<Sys *>(h);
    EnergyTracker *e = dynamic_cast<EnergyTracker *>(s->stats[t].get());
    if (from_box) e->set_U0(*s->box);
    else e->set_U0(U0);
}

const char *ref_last_error(void *h) { return static_cast<Sys *>(h)->err.c_str(); }

int ref_update_list(void *h, int nl, int force) { return static_cast<Sys *>(h)->nls[nl]->update_list_cells(force != 0) ? 1 : 0; }
// NeighborList::ignore(AtomID, AtomID), trackers.hpp:190-193
void ref_ignore(void *h, int nl, const uint32_t *a, const uint32_t *b, uint64_t npairs) {
    Sys *s = static_cast<Sys *>(h);
    for (uint64_t k = 0; k < npairs; k++) s->nls[nl]->ignore(s->atoms->get_id(a[k]), s->atoms->get_id(b[k]));
}
uint32_t ref_ignore_size(void *h, int nl) { return static_cast<Sys *>(h)->nls[nl]->ignore_size(); }
uint32_t ref_which(void *h, int nl) { return static_cast<Sys *>(h)->nls[nl]->which(); }
uint32_t ref_numpairs(void *h, int nl) { return static_cast<Sys *>(h)->nls[nl]->numpairs(); }
// pairs in the reference's own order; first() is the later atom (trackers.cpp:59-68)
void ref_get_pairs(void *h, int nl, uint32_t *first, uint32_t *last) {
    NeighborList &L = *static_cast<Sys *>(h)->nls[nl];
    uint k = 0;
    for (vector<IDPair>::iterator it = L.begin(); it != L.end(); ++it, ++k) {
        first[k] = it->first().n();
        last[k] = it->last().n();
    }
}

void ref_set_atoms(void *h, const double *x, const double *v, const double *a, const double *f) {
    AtomVec &av = *static_cast<Sys *>(h)->atoms;
    for (uint i = 0; i < av.size(); i++) {
        if (x) av[i].x = vec_from(x + (size_t)i * NDIM);
        if (v) av[i].v = vec_from(v + (size_t)i * NDIM);
        if (a) av[i].a = vec_from(a + (size_t)i * NDIM);
        if (f) av[i].f = vec_from(f + (size_t)i * NDIM);
    }
}
void ref_get_atoms(void *h, double *x, double *v, double *a, double *f) {
    AtomVec &av = *static_cast<Sys *>(h)->atoms;
    for (uint i = 0; i < av.size(); i++)
        for (uint d = 0; d < NDIM; d++) {
            if (x) x[(size_t)i * NDIM + d] = av[i].x[d];
            if (v) v[(size_t)i * NDIM + d] = av[i].v[d];
            if (a) a[(size_t)i * NDIM + d] = av[i].a[d];
            if (f) f[(size_t)i * NDIM + d] = av[i].f[d];
        }
}

void ref_box_diff(void *h, const double *r1, const double *r2, double *out) {
    Vec d = static_cast<Sys *>(h)->box->diff(vec_from(r1), vec_from(r2));
    for (uint k = 0; k < NDIM; k++) out[k] = d[k];
}
double ref_box_V(void *h) { return static_cast<Sys *>(h)->box->V(); }

// Interaction-level calls (do not reset forces: exactly the virtuals)
void ref_reset_forces(void *h) { static_cast<Sys *>(h)->atoms->reset_forces(); }
void ref_inter_set_forces(void *h, int k) { Sys *s = static_cast<Sys *>(h); s->inters[k]->set_forces(*s->box); }
double ref_inter_set_forces_get_pressure(void *h, int k) { Sys *s = static_cast<Sys *>(h); return s->inters[k]->set_forces_get_pressure(*s->box); }
double ref_inter_energy(void *h, int k) { Sys *s = static_cast<Sys *>(h); return s->inters[k]->energy(*s->box); }
double ref_inter_pressure(void *h, int k) { Sys *s = static_cast<Sys *>(h); return s->inters[k]->pressure(*s->box); }
void ref_inter_stress(void *h, int k, double *out) {
    Sys *s = static_cast<Sys *>(h);
    Matrix m = s->inters[k]->stress(*s->box);
    for (uint i = 0; i < NDIM; i++)
        for (uint j = 0; j < NDIM; j++) out[i * NDIM + j] = m(i, j);
}

// Collection-level
void ref_timestep(void *h, int nsteps) {
    Collection &c = *static_cast<Sys *>(h)->collec;
    for (int i = 0; i < nsteps; i++) c.timestep();
}
void ref_set_forces(void *h, int constraints_and_a) { static_cast<Sys *>(h)->collec->set_forces(constraints_and_a != 0); }
double ref_energy(void *h) { return static_cast<Sys *>(h)->collec->energy(); }
double ref_potential_energy(void *h) { return static_cast<Sys *>(h)->collec->potential_energy(); }
double ref_kinetic_energy(void *h) { return static_cast<Sys *>(h)->collec->kinetic_energy(); }
double ref_temp(void *h, int minuscomv) { return static_cast<Sys *>(h)->collec->temp(minuscomv != 0); }
double ref_pressure(void *h) { return static_cast<Sys *>(h)->collec->pressure(); }
double ref_virial(void *h) { return static_cast<Sys *>(h)->collec->virial(); }
double ref_degrees_of_freedom(void *h) { return static_cast<Sys *>(h)->collec->degrees_of_freedom(); }
void ref_reset_com_velocity(void *h) { static_cast<Sys *>(h)->collec->reset_com_velocity(); }
void ref_scale_velocities(void *h, double s) { static_cast<Sys *>(h)->collec->scale_velocities(s); }
void ref_scale_velocities_to_temp(void *h, double T, int minuscomv) { static_cast<Sys *>(h)->collec->scale_velocities_to_temp(T, minuscomv != 0); }
void ref_scale_velocities_to_energy(void *h, double E) { static_cast<Sys *>(h)->collec->scale_velocities_to_energy(E); }
void ref_com_velocity(void *h, double *out) {
    Vec v = static_cast<Sys *>(h)->atoms->com_velocity();
    for (uint k = 0; k < NDIM; k++) out[k] = v[k];
}
void ref_momentum(void *h, double *out) {
    Vec v = static_cast<Sys *>(h)->atoms->momentum();
    for (uint k = 0; k < NDIM; k++) out[k] = v[k];
}
double ref_mass(void *h) { return static_cast<Sys *>(h)->atoms->mass(); }
double ref_atoms_kinetic_energy(void *h, const double *v0) { return static_cast<Sys *>(h)->atoms->kinetic_energy(vec_from(v0)); }
void ref_add_velocity(void *h, const double *dv) { static_cast<Sys *>(h)->atoms->add_velocity(vec_from(dv)); }

// Noise injection for CollectionSol: subsequent Gaussian draws made by the
// reference come from `draws` (in draw order) instead of the RNG.
void ref_inject_noise(const double *draws, size_t n) {
    parm_oracle_noise = draws;
    parm_oracle_noise_len = n;
    parm_oracle_noise_pos = 0;
}
size_t ref_noise_consumed() { return parm_oracle_noise_pos; }
// Which component receives which draw in `Vec(gauss(), gauss(), gauss())`
// (argument evaluation order is compiler-specific): out[c] = draw index.
void ref_probe_draw_order(int *out) {
    double probe[NDIM];
    for (uint d = 0; d < NDIM; d++) probe[d] = (double)d;
    ref_inject_noise(probe, NDIM);
    Vec v = rand_vec();
    ref_inject_noise(0, 0);
    for (uint d = 0; d < NDIM; d++) out[d] = (int)v[d];
}
void ref_seed(uint32_t n) { seed(n); }

}  // extern "C"
