// parm_b200 drop-in for ParM's src/interaction.hpp -- the neighbour-listed pair interactions: the
// Interaction interface (interaction.hpp:59-104), every per-atom parameter struct and pair functor that
// sim.i:621-643 instantiates NListed with, and NListed<A,P> itself (interaction.hpp:1876-1945). The pair
// loop runs on the device (parm_b200/csrc/force.cu); (A,P) combinations the reference has no pair
// constructor for fail to compile here too.
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

// ---- per-atom parameter structs (same members and constructors as the reference) ---------------------
// pack(): the five parameter slots of include/parm_b200.h (parm_inter_set_params_ex), indx, and the
// `epsilons` / `sigmas` vectors for the indexed structs.
struct EpsSigAtom : public AtomID {  // interaction.hpp:857-865
    flt epsilon, sigma;
    EpsSigAtom() {}
    EpsSigAtom(AtomID a, flt epsilon, flt sigma) : AtomID(a), epsilon(epsilon), sigma(sigma) {}
    EpsSigAtom(AtomID a, EpsSigAtom other) : AtomID(a), epsilon(other.epsilon), sigma(other.sigma) {}
    flt max_size() { return sigma; }
    void pack(flt *p, uint32_t &, vector<flt> *&, vector<flt> *&) { p[0] = epsilon; p[1] = sigma; }
};
struct EpsSigCutAtom : public EpsSigAtom {  // interaction.hpp:897-905
    flt sigcut;
    EpsSigCutAtom() {}
    EpsSigCutAtom(AtomID a, flt epsilon, flt sigma, flt cut) : EpsSigAtom(a, epsilon, sigma), sigcut(cut) {}
    EpsSigCutAtom(AtomID a, EpsSigCutAtom other) : EpsSigAtom(a, other), sigcut(other.sigcut) {}
    flt max_size() { return sigma * sigcut; }
    void pack(flt *p, uint32_t &, vector<flt> *&, vector<flt> *&) { p[0] = epsilon; p[1] = sigma; p[2] = sigcut; }
};
struct IEpsISigCutAtom : public AtomID {  // interaction.hpp:911-960
    vector<flt> epsilons, sigmas;
    uint indx;
    flt sigcut;
    IEpsISigCutAtom() {}
    IEpsISigCutAtom(AtomID a, vector<flt> epsilons, vector<flt> sigmas, uint indx, flt cut)
        : AtomID(a), epsilons(epsilons), sigmas(sigmas), indx(indx), sigcut(cut) {
        assert(sigmas.size() == epsilons.size());
    }
    IEpsISigCutAtom(AtomID a, IEpsISigCutAtom other)
        : AtomID(a), epsilons(other.epsilons), sigmas(other.sigmas), indx(other.indx), sigcut(other.sigcut) {}
    flt get_epsilon(IEpsISigCutAtom &other) { assert(other.indx < epsilons.size()); return epsilons[other.indx]; }
    flt get_sigma(IEpsISigCutAtom &other) { assert(other.indx < sigmas.size()); return sigmas[other.indx]; }
    flt max_size() {
        flt sigma = sigmas[0];
        for (uint i = 1; i < sigmas.size(); ++i)
            if (sigma < sigmas[i]) sigma = sigmas[i];
        return sigma * sigcut;
    }
    void pack(flt *p, uint32_t &type, vector<flt> *&eps, vector<flt> *&sig) { p[2] = sigcut; type = indx; eps = &epsilons; sig = &sigmas; }
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
    void pack(flt *p, uint32_t &type, vector<flt> *&eps, vector<flt> *&) { p[1] = sigma; p[2] = sigcut; type = indx; eps = &epsilons; }
};
struct IEpsRepsSigExpCutAtom : public AtomID {  // interaction.hpp:1051-1093
    vector<flt> epsilons;
    flt repeps, sigma;
    flt exponent;
    uint indx;
    flt sigcut;
    IEpsRepsSigExpCutAtom() {}
    IEpsRepsSigExpCutAtom(AtomID a, vector<flt> epsilons, flt repeps, flt sigma, flt n, uint indx, flt cut)
        : AtomID(a), epsilons(epsilons), repeps(repeps), sigma(sigma), exponent(n), indx(indx), sigcut(cut) {
        assert(indx < epsilons.size());
    }
    IEpsRepsSigExpCutAtom(AtomID a, IEpsRepsSigExpCutAtom other)
        : AtomID(a), epsilons(other.epsilons), repeps(other.repeps), sigma(other.sigma), exponent(other.exponent),
          indx(other.indx), sigcut(other.sigcut) {}
    flt get_epsilon(IEpsRepsSigExpCutAtom &other) { assert(other.indx < epsilons.size()); return epsilons[other.indx]; }
    flt get_sigma(IEpsRepsSigExpCutAtom &other) { return (sigma + other.sigma) / 2.0; }
    flt max_size() { return sigma * sigcut; }
    void pack(flt *p, uint32_t &type, vector<flt> *&eps, vector<flt> *&) {
        p[1] = sigma; p[2] = sigcut; p[3] = repeps; p[4] = exponent; type = indx; eps = &epsilons;
    }
};
struct EpsEpsSigSigCutAtom : public AtomID {  // interaction.hpp:1149-1169
    flt eps_r, eps_a, sig_r, sig_a;
    flt sigcut;
    EpsEpsSigSigCutAtom() {}
    EpsEpsSigSigCutAtom(AtomID a, flt eps_r, flt eps_a, flt sigma_r, flt sigma_a, flt cut)
        : AtomID(a), eps_r(eps_r), eps_a(eps_a), sig_r(sigma_r), sig_a(sigma_a), sigcut(cut) {}
    EpsEpsSigSigCutAtom(AtomID a, EpsEpsSigSigCutAtom other)
        : AtomID(a), eps_r(other.eps_r), eps_a(other.eps_a), sig_r(other.sig_r), sig_a(other.sig_a), sigcut(other.sigcut) {}
    flt max_size() { return sig_r + sig_a * (sigcut - 1); }
    void pack(flt *p, uint32_t &, vector<flt> *&, vector<flt> *&) { p[0] = eps_r; p[1] = sig_r; p[2] = sigcut; p[3] = eps_a; p[4] = sig_a; }
};
struct IEpsRepsSigCutAtom : public AtomID {  // interaction.hpp:1307-1341
    vector<flt> epsilons;
    flt repeps, sig;
    uint indx;
    flt sigcut;
    IEpsRepsSigCutAtom() {}
    IEpsRepsSigCutAtom(AtomID a, vector<flt> epsilons, flt repeps, flt sigma, uint indx, flt cut)
        : AtomID(a), epsilons(epsilons), repeps(repeps), sig(sigma), indx(indx), sigcut(cut) {
        assert(indx < epsilons.size());
    }
    IEpsRepsSigCutAtom(AtomID a, IEpsRepsSigCutAtom other)
        : AtomID(a), epsilons(other.epsilons), repeps(other.repeps), sig(other.sig), indx(other.indx), sigcut(other.sigcut) {}
    flt get_epsilon(IEpsRepsSigCutAtom &other) { assert(other.indx < epsilons.size()); return epsilons[other.indx]; }
    flt max_size() { return sig * sigcut; }
    void pack(flt *p, uint32_t &type, vector<flt> *&eps, vector<flt> *&) { p[1] = sig; p[2] = sigcut; p[3] = repeps; type = indx; eps = &epsilons; }
};
struct EisMclachlanAtom : public AtomID {  // interaction.hpp:1415-1423
    flt dist, sigmai;
    EisMclachlanAtom() {}
    EisMclachlanAtom(AtomID a, flt dist, flt sigmai) : AtomID(a), dist(dist), sigmai(sigmai) {}
    EisMclachlanAtom(AtomID a, EisMclachlanAtom other) : AtomID(a), dist(other.dist), sigmai(other.sigmai) {}
    flt max_size() { return dist; }
    void pack(flt *p, uint32_t &, vector<flt> *&, vector<flt> *&) { p[0] = sigmai; p[1] = dist; }
};
struct EpsSigExpAtom : public AtomID {  // interaction.hpp:1454-1465
    flt eps, sigma, exponent;
    EpsSigExpAtom() {}
    EpsSigExpAtom(AtomID a, flt eps, flt sigma, flt exponent) : AtomID(a), eps(eps), sigma(sigma), exponent(exponent) {}
    EpsSigExpAtom(AtomID a, EpsSigExpAtom other) : AtomID(a), eps(other.eps), sigma(other.sigma), exponent(other.exponent) {}
    flt max_size() { return sigma; }
    void pack(flt *p, uint32_t &, vector<flt> *&, vector<flt> *&) { p[0] = eps; p[1] = sigma; p[2] = exponent; }
};
struct IEpsISigExpAtom : public AtomID {  // interaction.hpp:1477-1524
    vector<flt> epsilons, sigmas;
    flt exponent;
    uint indx;
    IEpsISigExpAtom() {}
    IEpsISigExpAtom(AtomID a, vector<flt> epsilons, vector<flt> sigmas, uint indx, flt exponent = 2.5)
        : AtomID(a), epsilons(epsilons), sigmas(sigmas), exponent(exponent), indx(indx) {
        assert(sigmas.size() == epsilons.size());
    }
    flt get_epsilon(IEpsISigExpAtom &other) { assert(other.indx < epsilons.size()); return epsilons[other.indx]; }
    flt get_sigma(IEpsISigExpAtom &other) { assert(other.indx < sigmas.size()); return sigmas[other.indx]; }
    flt max_size() {
        flt sigma = sigmas[0];
        for (uint i = 1; i < sigmas.size(); ++i)
            if (sigma < sigmas[i]) sigma = sigmas[i];
        return sigma;
    }
    void pack(flt *p, uint32_t &type, vector<flt> *&eps, vector<flt> *&sig) { p[2] = exponent; type = indx; eps = &epsilons; sig = &sigmas; }
};
struct EpsSigExpDragAtom : public AtomID {  // interaction.hpp:1598-1611
    flt eps, sigma, exponent, gamma;
    EpsSigExpDragAtom() {}
    EpsSigExpDragAtom(AtomID a, flt eps, flt sigma, flt gamma, flt exponent = 2.5)
        : AtomID(a), eps(eps), sigma(sigma), exponent(exponent), gamma(gamma) {}
    EpsSigExpDragAtom(AtomID a, EpsSigExpDragAtom other)
        : AtomID(a), eps(other.eps), sigma(other.sigma), exponent(other.exponent), gamma(other.gamma) {}
    flt max_size() { return sigma; }
    void pack(flt *p, uint32_t &, vector<flt> *&, vector<flt> *&) { p[0] = eps; p[1] = sigma; p[2] = exponent; p[3] = gamma; }
};
struct LoisOhernAtom : public AtomID {  // interaction.hpp:1679-1691
    flt eps, sigma, C, l;
    LoisOhernAtom() {}
    LoisOhernAtom(AtomID a, flt eps, flt sigma, flt C, flt l) : AtomID(a), eps(eps), sigma(sigma), C(C), l(l) {}
    LoisOhernAtom(AtomID a, LoisOhernAtom other) : AtomID(a), eps(other.eps), sigma(other.sigma), C(other.C), l(other.l) {}
    flt max_size() { return sigma * (1 + C + l); }
    void pack(flt *p, uint32_t &, vector<flt> *&, vector<flt> *&) { p[0] = eps; p[1] = sigma; p[2] = C; p[3] = l; }
};
struct LoisLinAtom : public AtomID {  // interaction.hpp:1764-1780
    flt eps, sigma, f, l;
    LoisLinAtom() {}
    LoisLinAtom(AtomID a, flt eps, flt sigma, flt depth, flt width)
        : AtomID(a), eps(eps), sigma(sigma), f(width > 0 ? depth / width : 0), l(width) {}
    LoisLinAtom(AtomID a, LoisLinAtom other) : AtomID(a), eps(other.eps), sigma(other.sigma), f(other.f), l(other.l) {}
    flt max_size() { return sigma * (1 + l); }
    void pack(flt *p, uint32_t &, vector<flt> *&, vector<flt> *&) { p[0] = eps; p[1] = sigma; p[2] = f; p[3] = l; }
};

// ---- pair functors: on the device these are tags selecting the kernel (csrc/pairs.cuh) -----------------
struct LJRepulsePair { enum { kind = PARM_PAIR_LJREPULSE }; };                         // :875-891
typedef LJRepulsePair LJRepulsivePair;  // planned rename, src/namereplacements.txt:181
struct RepulsionPair { enum { kind = PARM_PAIR_REPULSION }; };                         // :1528-1566
struct LJAttractRepulsePair { enum { kind = PARM_PAIR_LJATTRACTREPULSE }; };           // :1251-1299
struct LennardJonesCutPair { enum { kind = PARM_PAIR_LJCUT }; };                       // :967-987
struct LJAttractCutPair { enum { kind = PARM_PAIR_LJATTRACTCUT }; };                   // :1020-1049
struct LJAttractFixedRepulsePair { enum { kind = PARM_PAIR_LJATTRACTFIXEDREPULSE }; }; // :1343-1413
struct EisMclachlanPair { enum { kind = PARM_PAIR_EISMCLACHLAN }; };                   // :1425-1452
struct LJishPair { enum { kind = PARM_PAIR_LJISH }; };                                 // :1095-1142
struct LJAttractRepulseSigsPair { enum { kind = PARM_PAIR_LJATTRACTREPULSESIGS }; };   // :1171-1244
struct RepulsionDragPair { enum { kind = PARM_PAIR_REPULSIONDRAG }; };                 // :1613-1642
struct LoisOhernPair { enum { kind = PARM_PAIR_LOISOHERN }; };                         // :1693-1744
struct LoisOhernPairMinCLs { enum { kind = PARM_PAIR_LOISOHERNMIN }; };                // :1746-1752
struct LoisLinPair { enum { kind = PARM_PAIR_LOISLIN }; };                             // :1782-1829
struct LoisLinPairMin { enum { kind = PARM_PAIR_LOISLINMIN }; };                       // :1831-1837

namespace parm_b200 {
// the (A, P) combinations for which the reference's P has a constructor P(A, A)
template <class A, class P>
struct supported_pair { enum { value = 0 }; };
#define PARM_B200_PAIR(A, P) template <> struct supported_pair<A, P> { enum { value = 1 }; }
PARM_B200_PAIR(EpsSigAtom, LJRepulsePair);
PARM_B200_PAIR(EpsSigExpAtom, RepulsionPair);
PARM_B200_PAIR(IEpsISigExpAtom, RepulsionPair);
PARM_B200_PAIR(IEpsSigCutAtom, LJAttractRepulsePair);
PARM_B200_PAIR(EpsSigCutAtom, LennardJonesCutPair);
PARM_B200_PAIR(IEpsISigCutAtom, LennardJonesCutPair);
PARM_B200_PAIR(EpsSigCutAtom, LJAttractCutPair);
PARM_B200_PAIR(IEpsSigCutAtom, LJAttractCutPair);
PARM_B200_PAIR(IEpsISigCutAtom, LJAttractCutPair);
PARM_B200_PAIR(IEpsRepsSigCutAtom, LJAttractFixedRepulsePair);
PARM_B200_PAIR(EisMclachlanAtom, EisMclachlanPair);
PARM_B200_PAIR(IEpsRepsSigExpCutAtom, LJishPair);
PARM_B200_PAIR(EpsEpsSigSigCutAtom, LJAttractRepulseSigsPair);
PARM_B200_PAIR(EpsSigExpDragAtom, RepulsionDragPair);
PARM_B200_PAIR(LoisOhernAtom, LoisOhernPair);
PARM_B200_PAIR(LoisOhernAtom, LoisOhernPairMinCLs);
PARM_B200_PAIR(LoisLinAtom, LoisLinPair);
PARM_B200_PAIR(LoisLinAtom, LoisLinPairMin);
#undef PARM_B200_PAIR
}  // namespace parm_b200

template <class A, class P>
class NListed : public Interaction, public parm_b200::DeviceInteraction {
    static_assert(parm_b200::supported_pair<A, P>::value,
                  "parm_b200: the reference has no P(A, A) constructor for this NListed<A,P> combination");

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
        const int NP = PARM_PAIR_MAXPARAMS;
        vector<flt> params(NP * (n ? n : 1), 0.0), etable, stable;
        vector<uint32_t> types(n ? n : 1, 0);
        uint nt = 0;
        bool have_sig = false;
        for (size_t i = 0; i < n; i++) {
            if (!member[i]) continue;
            vector<flt> *eps = NULL, *sig = NULL;
            atoms[i].pack(&params[NP * i], types[i], eps, sig);
            if (eps) nt = max(nt, max((uint)eps->size(), types[i] + 1));
            if (sig) have_sig = true;
        }
        if (nt) {  // row indx of the table = the `epsilons` / `sigmas` vector carried by atoms of that indx
            etable.assign((size_t)nt * nt, 0.0);
            if (have_sig) stable.assign((size_t)nt * nt, 0.0);
            for (size_t i = 0; i < n; i++) {
                if (!member[i]) continue;
                vector<flt> *eps = NULL, *sig = NULL;
                uint32_t t = 0;
                flt tmp[PARM_PAIR_MAXPARAMS];
                atoms[i].pack(tmp, t, eps, sig);
                for (size_t k = 0; eps && k < eps->size(); k++) etable[(size_t)t * nt + k] = (*eps)[k];
                for (size_t k = 0; sig && k < sig->size(); k++) stable[(size_t)t * nt + k] = (*sig)[k];
            }
        }
        parm_b200::check(parm_inter_set_params_ex(inter, params.data(), NP, types.data(), nt ? etable.data() : NULL,
                                                  have_sig ? stable.data() : NULL, (int)nt, member.data(), 0));
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
