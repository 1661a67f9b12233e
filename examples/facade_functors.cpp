// Every NListed<A,P> instantiation of the reference's sim.i:621-643 through the parm_b200 C++ facade, built
// with the reference's own per-atom structs and constructors. tests/test_gpu_facade.py feeds the same numbers
// to the CPU oracle and compares.
//   facade_functors <in.bin> <out.bin>
// in.bin : int32 n, kind, variant (0 "", 1 "I", 2 "II"), ntypes; double L[NDIM], skin;
//          double x[n][NDIM], v[n][NDIM], m[n], params[n][5]; uint32 type[n]; double eps[nt][nt], sig[nt][nt]
// out.bin: double E, virial, stress[NDIM][NDIM]; uint64 contacts, overlaps, numpairs; double f[n][NDIM]
#include <cstdio>
#include <cstdlib>

#include "collection.hpp"
#include "interaction.hpp"

template <class T>
static void rd(FILE *f, T *p, size_t n) {
    if (n && fread(p, sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
}

struct Input {
    uint n;
    int kind, variant, nt;
    vector<double> par, eps, sig;
    vector<uint32_t> type;
    double p(uint i, int q) const { return par[5 * i + q]; }
    vector<flt> erow(uint i) const { return vector<flt>(eps.begin() + type[i] * nt, eps.begin() + (type[i] + 1) * nt); }
    vector<flt> srow(uint i) const { return vector<flt>(sig.begin() + type[i] * nt, sig.begin() + (type[i] + 1) * nt); }
};

struct Result {
    double E, virial;
    Matrix stress;
    unsigned long long contacts, overlaps, numpairs;
};

template <class A, class P, class MK>
static Result run(sptr<OriginBox> box, sptr<AtomVec> atoms, flt skin, const Input &in, MK make) {
    sptr<NListed<A, P> > I(new NListed<A, P>(box, atoms, skin));
    for (uint i = 0; i < in.n; i++) I->add(make(atoms->get_id(i), i));
    I->neighbor_list()->update_list(true);
    Result r;
    r.numpairs = I->neighbor_list()->numpairs();
    r.E = I->energy(*box);
    r.contacts = I->contacts(*box);
    r.overlaps = I->overlaps(*box);
    r.stress = I->stress(*box);
    atoms->reset_forces();
    r.virial = I->set_forces_get_pressure(*box);
    return r;
}

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
    FILE *fi = fopen(argv[1], "rb");
    if (!fi) { perror("in"); return 2; }
    int hdr[4];
    rd(fi, hdr, 4);
    Input in;
    in.n = (uint)hdr[0];
    in.kind = hdr[1];
    in.variant = hdr[2];
    in.nt = hdr[3];
    const uint n = in.n;
    double L[NDIM], skin;
    rd(fi, L, NDIM);
    rd(fi, &skin, 1);
    vector<double> x(n * NDIM), v(n * NDIM), m(n);
    in.par.resize(n * 5);
    in.type.resize(n);
    in.eps.resize(in.nt * in.nt);
    in.sig.resize(in.nt * in.nt);
    rd(fi, x.data(), x.size());
    rd(fi, v.data(), v.size());
    rd(fi, m.data(), m.size());
    rd(fi, in.par.data(), in.par.size());
    rd(fi, in.type.data(), in.type.size());
    rd(fi, in.eps.data(), in.eps.size());
    rd(fi, in.sig.data(), in.sig.size());
    fclose(fi);

    Vec Lv;
    for (uint d = 0; d < NDIM; d++) Lv[d] = L[d];
    sptr<OriginBox> box(new OriginBox(Lv));
    sptr<AtomVec> atomptr(new AtomVec(m));
    AtomVec &atoms = *atomptr;
    for (uint i = 0; i < n; i++)
        for (uint d = 0; d < NDIM; d++) {
            atoms[i].x[d] = x[i * NDIM + d];
            atoms[i].v[d] = v[i * NDIM + d];
        }
    Result r;
    const int k = in.kind, var = in.variant;
    if (k == PARM_PAIR_REPULSION && var == 2)
        r = run<IEpsISigExpAtom, RepulsionPair>(box, atomptr, skin, in, [&](AtomID a, uint i) {
            return IEpsISigExpAtom(a, in.erow(i), in.srow(i), in.type[i], in.p(i, 2)); });
    else if (k == PARM_PAIR_LJCUT && var == 2)
        r = run<IEpsISigCutAtom, LennardJonesCutPair>(box, atomptr, skin, in, [&](AtomID a, uint i) {
            return IEpsISigCutAtom(a, in.erow(i), in.srow(i), in.type[i], in.p(i, 2)); });
    else if (k == PARM_PAIR_LJATTRACTCUT && var == 0)
        r = run<EpsSigCutAtom, LJAttractCutPair>(box, atomptr, skin, in, [&](AtomID a, uint i) {
            return EpsSigCutAtom(a, in.p(i, 0), in.p(i, 1), in.p(i, 2)); });
    else if (k == PARM_PAIR_LJATTRACTCUT && var == 1)
        r = run<IEpsSigCutAtom, LJAttractCutPair>(box, atomptr, skin, in, [&](AtomID a, uint i) {
            return IEpsSigCutAtom(a, in.erow(i), in.type[i], in.p(i, 1), in.p(i, 2)); });
    else if (k == PARM_PAIR_LJATTRACTCUT)
        r = run<IEpsISigCutAtom, LJAttractCutPair>(box, atomptr, skin, in, [&](AtomID a, uint i) {
            return IEpsISigCutAtom(a, in.erow(i), in.srow(i), in.type[i], in.p(i, 2)); });
    else if (k == PARM_PAIR_LJATTRACTFIXEDREPULSE)
        r = run<IEpsRepsSigCutAtom, LJAttractFixedRepulsePair>(box, atomptr, skin, in, [&](AtomID a, uint i) {
            return IEpsRepsSigCutAtom(a, in.erow(i), in.p(i, 3), in.p(i, 1), in.type[i], in.p(i, 2)); });
    else if (k == PARM_PAIR_EISMCLACHLAN)
        r = run<EisMclachlanAtom, EisMclachlanPair>(box, atomptr, skin, in, [&](AtomID a, uint i) {
            return EisMclachlanAtom(a, in.p(i, 1), in.p(i, 0)); });
    else if (k == PARM_PAIR_LJISH)
        r = run<IEpsRepsSigExpCutAtom, LJishPair>(box, atomptr, skin, in, [&](AtomID a, uint i) {
            return IEpsRepsSigExpCutAtom(a, in.erow(i), in.p(i, 3), in.p(i, 1), in.p(i, 4), in.type[i], in.p(i, 2)); });
    else if (k == PARM_PAIR_LJATTRACTREPULSESIGS)
        r = run<EpsEpsSigSigCutAtom, LJAttractRepulseSigsPair>(box, atomptr, skin, in, [&](AtomID a, uint i) {
            return EpsEpsSigSigCutAtom(a, in.p(i, 0), in.p(i, 3), in.p(i, 1), in.p(i, 4), in.p(i, 2)); });
    else if (k == PARM_PAIR_REPULSIONDRAG)
        r = run<EpsSigExpDragAtom, RepulsionDragPair>(box, atomptr, skin, in, [&](AtomID a, uint i) {
            return EpsSigExpDragAtom(a, in.p(i, 0), in.p(i, 1), in.p(i, 3), in.p(i, 2)); });
    else if (k == PARM_PAIR_LOISOHERN)
        r = run<LoisOhernAtom, LoisOhernPair>(box, atomptr, skin, in, [&](AtomID a, uint i) {
            return LoisOhernAtom(a, in.p(i, 0), in.p(i, 1), in.p(i, 2), in.p(i, 3)); });
    else if (k == PARM_PAIR_LOISOHERNMIN)
        r = run<LoisOhernAtom, LoisOhernPairMinCLs>(box, atomptr, skin, in, [&](AtomID a, uint i) {
            return LoisOhernAtom(a, in.p(i, 0), in.p(i, 1), in.p(i, 2), in.p(i, 3)); });
    else if (k == PARM_PAIR_LOISLIN)  // LoisLinAtom(a, eps, sigma, depth, width): depth = f * width
        r = run<LoisLinAtom, LoisLinPair>(box, atomptr, skin, in, [&](AtomID a, uint i) {
            return LoisLinAtom(a, in.p(i, 0), in.p(i, 1), in.p(i, 2) * in.p(i, 3), in.p(i, 3)); });
    else if (k == PARM_PAIR_LOISLINMIN)
        r = run<LoisLinAtom, LoisLinPairMin>(box, atomptr, skin, in, [&](AtomID a, uint i) {
            return LoisLinAtom(a, in.p(i, 0), in.p(i, 1), in.p(i, 2) * in.p(i, 3), in.p(i, 3)); });
    else {
        fprintf(stderr, "facade_functors: kind %d variant %d not handled here\n", k, var);
        return 2;
    }
    vector<double> fo(n * NDIM), st(NDIM * NDIM);
    for (uint i = 0; i < n; i++)
        for (uint d = 0; d < NDIM; d++) fo[i * NDIM + d] = atoms[i].f[d];
    for (uint a = 0; a < NDIM; a++)
        for (uint b = 0; b < NDIM; b++) st[a * NDIM + b] = r.stress(a, b);
    FILE *fo_ = fopen(argv[2], "wb");
    double sc[2] = {r.E, r.virial};
    unsigned long long u[3] = {r.contacts, r.overlaps, r.numpairs};
    fwrite(sc, 8, 2, fo_);
    fwrite(st.data(), 8, st.size(), fo_);
    fwrite(u, 8, 3, fo_);
    fwrite(fo.data(), 8, fo.size(), fo_);
    fclose(fo_);
    printf("facade_functors ok: kind=%d variant=%d n=%u pairs=%llu E=%.12g contacts=%llu\n", k, var, n, u[2], r.E, u[0]);
    return 0;
}
