// parm_b200 drop-in for ParM's src/interaction.hpp -- hot-path subset: the Interaction interface
// (interaction.hpp:59-104), the per-atom parameter structs and pair functors in scope, and
// NListed<A,P> (interaction.hpp:1876-1945). The pair loop runs on the device
// (parm_b200/csrc/force.cu); unsupported (A,P) combinations fail to compile.
#include "trackers.hpp"

#ifndef PARM_B200_INTERACTION_H
#define PARM_B200_INTERACTION_H

#include <typeinfo>

class Interaction {
   public:
    virtual flt energy(Box &box) = 0;
    virtual void set_forces(Box &box) = 0;
    virtual flt set_forces_get_pressure(Box &) {
        std::string s = std::string("set_forces_get_pressure not defined for class ");
        s.append(typeid(*this).name());
        throw std::runtime_error(s);
    }
    virtual flt pressure(Box &box) = 0;
    virtual Matrix stress(Box &) {
        std::string s = std::string("stress not defined for class ");
        s.append(typeid(*this).name());
        throw std::runtime_error(s);
    }
    virtual ~Interaction() {}
};

namespace parm_b200 {
// Interactions whose loop runs on the device expose their C handle to the Collection facade.
class DeviceInteraction {
   public:
    virtual parm_inter *device_handle() = 0;
    virtual ~DeviceInteraction() {}
};
}  // namespace parm_b200

// ---- per-atom parameter structs ------------------------------------------------------
struct EpsSigAtom : public AtomID {  // interaction.hpp:857-865
    flt epsilon, sigma;
    EpsSigAtom() {}
    EpsSigAtom(AtomID a, flt epsilon, flt sigma) : AtomID(a), epsilon(epsilon), sigma(sigma) {}
    EpsSigAtom(AtomID a, EpsSigAtom other) : AtomID(a), epsilon(other.epsilon), sigma(other.sigma) {}
    flt max_size() { return sigma; }
    void pack(flt *p, uint32_t &type, vector<flt> *&eps) { p[0] = epsilon; p[1] = sigma; p[2] = 0; type = 0; eps = NULL; }
};
struct EpsSigCutAtom : public EpsSigAtom {  // interaction.hpp:897-905
    flt sigcut;
    EpsSigCutAtom() {}
    EpsSigCutAtom(AtomID a, flt epsilon, flt sigma, flt cut) : EpsSigAtom(a, epsilon, sigma), sigcut(cut) {}
    EpsSigCutAtom(AtomID a, EpsSigCutAtom other) : EpsSigAtom(a, other), sigcut(other.sigcut) {}
    flt max_size() { return sigma * sigcut; }
    void pack(flt *p, uint32_t &type, vector<flt> *&eps) { p[0] = epsilon; p[1] = sigma; p[2] = sigcut; type = 0; eps = NULL; }
};
struct IEpsSigCutAtom : public AtomID {  // interaction.hpp:989-1018
    vector<flt> epsilons;
    uint indx;
    flt sigma;
    flt sigcut;
    IEpsSigCutAtom() {}
    IEpsSigCutAtom(AtomID a, vector<flt> epsilons, uint indx, flt sigma, flt cut)
        : AtomID(a), epsilons(epsilons), indx(indx), sigma(sigma), sigcut(cut) {}
    IEpsSigCutAtom(AtomID a, IEpsSigCutAtom other)
        : AtomID(a), epsilons(other.epsilons), indx(other.indx), sigma(other.sigma), sigcut(other.sigcut) {}
    flt get_epsilon(IEpsSigCutAtom &other) {
        assert(other.indx < epsilons.size());
        flt myeps = epsilons[other.indx];
        assert(indx < other.epsilons.size());
        assert(other.epsilons[indx] == myeps);
        return myeps;
    }
    flt max_size() { return sigma * sigcut; }
    void pack(flt *p, uint32_t &type, vector<flt> *&eps) { p[0] = 0; p[1] = sigma; p[2] = sigcut; type = indx; eps = &epsilons; }
};
struct EpsSigExpAtom : public AtomID {  // interaction.hpp:1454-1465
    flt eps, sigma, exponent;
    EpsSigExpAtom() {}
    EpsSigExpAtom(AtomID a, flt eps, flt sigma, flt exponent) : AtomID(a), eps(eps), sigma(sigma), exponent(exponent) {}
    EpsSigExpAtom(AtomID a, EpsSigExpAtom other) : AtomID(a), eps(other.eps), sigma(other.sigma), exponent(other.exponent) {}
    flt max_size() { return sigma; }
    void pack(flt *p, uint32_t &type, vector<flt> *&ep) { p[0] = eps; p[1] = sigma; p[2] = exponent; type = 0; ep = NULL; }
};

// ---- pair functors: on the device these are tags selecting the kernel ------------------
struct LJRepulsePair { enum { kind = PARM_PAIR_LJREPULSE }; typedef EpsSigAtom atom_type; };                 // :875-891
typedef LJRepulsePair LJRepulsivePair;  // planned rename, src/namereplacements.txt:181
struct RepulsionPair { enum { kind = PARM_PAIR_REPULSION }; typedef EpsSigExpAtom atom_type; };              // :1528-1566
struct LJAttractRepulsePair { enum { kind = PARM_PAIR_LJATTRACTREPULSE }; typedef IEpsSigCutAtom atom_type; }; // :1251-1299
struct LennardJonesCutPair { enum { kind = PARM_PAIR_LJCUT }; typedef EpsSigCutAtom atom_type; };            // :967-987

namespace parm_b200 {
template <class A, class P>
struct supported_pair { enum { value = 0 }; };
template <> struct supported_pair<EpsSigAtom, LJRepulsePair> { enum { value = 1 }; };
template <> struct supported_pair<EpsSigExpAtom, RepulsionPair> { enum { value = 1 }; };
template <> struct supported_pair<IEpsSigCutAtom, LJAttractRepulsePair> { enum { value = 1 }; };
template <> struct supported_pair<EpsSigCutAtom, LennardJonesCutPair> { enum { value = 1 }; };
}  // namespace parm_b200

template <class A, class P>
class NListed : public Interaction, public parm_b200::DeviceInteraction {
    static_assert(parm_b200::supported_pair<A, P>::value,
                  "parm_b200: this NListed<A,P> combination is outside the hot-path scope (see DESIGN.md)");

   protected:
    vector<A> atoms;  // indexed by AtomVec index, like the reference (interaction.hpp:1893-1897)
    vector<unsigned char> member;
    sptr<AtomVec> atomvec;
    sptr<NeighborList> neighbors;
    parm_inter *inter;
    bool dirty;
    uint ntypes;

    void create() {
        member.assign(atoms.size(), 0);
        inter = NULL;
        dirty = false;
        ntypes = 0;
        parm_b200::check(parm_inter_create(atomvec->context(), neighbors->handle(), (int)P::kind, &inter));
    }
    void flush() {
        if (!dirty) return;
        const size_t n = atoms.size();
        vector<flt> params(3 * (n ? n : 1), 0.0), table;
        vector<uint32_t> types(n ? n : 1, 0);
        uint nt = 0;
        for (size_t i = 0; i < n; i++) {
            if (!member[i]) continue;
            vector<flt> *eps = NULL;
            atoms[i].pack(&params[3 * i], types[i], eps);
            if (eps) nt = max(nt, max((uint)eps->size(), types[i] + 1));
        }
        if (nt) {
            table.assign((size_t)nt * nt, 0.0);
            for (size_t i = 0; i < n; i++) {
                if (!member[i]) continue;
                vector<flt> *eps = NULL;
                uint32_t t;
                flt tmp[3];
                atoms[i].pack(tmp, t, eps);
                for (size_t k = 0; eps && k < eps->size(); k++) table[(size_t)t * nt + k] = (*eps)[k];
            }
        }
        parm_b200::check(parm_inter_set_params(inter, params.data(), types.data(), nt ? table.data() : NULL, (int)nt,
                                               member.data(), 0));
        dirty = false;
    }
    parm_ctx *ready(bool modifies) {
        flush();
        neighbors->handle();
        return atomvec->device(modifies);
    }

   public:
    NListed(sptr<AtomVec> vec, sptr<NeighborList> neighbors)
        : atoms(vec->size()), atomvec(vec), neighbors(neighbors) { create(); }
    NListed(sptr<Box> box, sptr<AtomVec> atomv, const flt skin)
        : atoms(atomv->size()), atomvec(atomv), neighbors(new NeighborList(box, atomv, skin)) { create(); }
    ~NListed() { parm_inter_destroy(inter); }

    inline void add(A atm) {  // interaction.hpp:1906-1910
        neighbors->add(atm, atm.max_size());
        atoms[atm.n()] = atm;
        member[atm.n()] = 1;
        dirty = true;
    }
    A &getatom(uint n) { return atoms[n]; }
    uint size() { return ((uint)(atoms.size())); }
    inline vector<A> &atom_list() { return atoms; }
    inline sptr<NeighborList> neighbor_list() { return neighbors; }
    parm_inter *device_handle() { flush(); return inter; }

    flt energy(Box &) {
        ready(false);
        flt e = 0;
        parm_b200::check(parm_inter_energy(inter, &e));
        return e;
    }
    flt pressure(Box &) {
        ready(false);
        flt p = 0;
        parm_b200::check(parm_inter_pressure(inter, &p));
        return p;
    }
    Matrix stress(Box &) {
        ready(false);
        flt s[NDIM * NDIM];
        parm_b200::check(parm_inter_stress(inter, s));
        Matrix m;
        for (uint i = 0; i < NDIM; i++)
            for (uint j = 0; j < NDIM; j++) m(i, j) = s[i * NDIM + j];
        return m;
    }
    void set_forces(Box &) {
        ready(true);
        parm_b200::check(parm_inter_set_forces(inter, 0, NULL));
    }
    flt set_forces_get_pressure(Box &) {
        ready(true);
        flt p = 0;
        parm_b200::check(parm_inter_set_forces(inter, PARM_WANT_VIRIAL, &p));
        return p;
    }
    Matrix set_forces_get_stress(Box &) {
        ready(true);
        flt s[NDIM * NDIM];
        parm_b200::check(parm_inter_set_forces(inter, PARM_WANT_STRESS, s));
        Matrix m;
        for (uint i = 0; i < NDIM; i++)
            for (uint j = 0; j < NDIM; j++) m(i, j) = s[i * NDIM + j];
        return m;
    }
    unsigned long long contacts(Box &) {
        ready(false);
        uint64_t c = 0, o = 0;
        parm_b200::check(parm_inter_contacts(inter, &c, &o));
        return c;
    }
    unsigned long long overlaps(Box &) {
        ready(false);
        uint64_t c = 0, o = 0;
        parm_b200::check(parm_inter_contacts(inter, &c, &o));
        return o;
    }
};

#endif
