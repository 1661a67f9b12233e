// Drives the parm_b200 C++ facade exactly the way a ParM user program does (cf. the reference's
// src/bin/LJatoms.cpp:30-104): OriginBox, AtomVec, NListed<...>, NeighborList, CollectionVerlet,
// Atom& access between steps. Inputs and outputs are raw binary so tests/test_gpu_facade.py can
// feed the same numbers to the CPU oracle and compare.
//   facade_run <in.bin> <out.bin>
// in.bin : int32 n, kind, steps; double L[NDIM], skin, dt; double x[n][NDIM], v[n][NDIM], m[n], params[n][3]
// out.bin: double E0, K0, U0, E1, K1, U1, P1, T1; uint32 numpairs0, which1; double x[n][NDIM], v[n][NDIM], f[n][NDIM];
//          then the statistics trackers registered with the collection: double rsq_count[2], r2[2][n] (lags 1 and 5),
//          r4[n] (lag 5), isf_re[n], isf_im[n] (k = 1.5, lag 3), et_n, et_E, et_U, et_K, et_Estd
#include <cstdio>
#include <cstdlib>

#include "collection.hpp"
#include "interaction.hpp"
#include "vecrand.hpp"

template <class T>
static void rd(FILE *f, T *p, size_t n) {
    if (fread(p, sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
}

template <class A, class P>
static sptr<NListed<A, P> > make(sptr<OriginBox> box, sptr<AtomVec> atoms, flt skin) {
    return sptr<NListed<A, P> >(new NListed<A, P>(box, atoms, skin));
}

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
    FILE *fi = fopen(argv[1], "rb");
    if (!fi) { perror("in"); return 2; }
    int hdr[3];
    rd(fi, hdr, 3);
    const uint n = (uint)hdr[0];
    const int kind = hdr[1], steps = hdr[2];
    double L[NDIM], skin, dt;
    rd(fi, L, NDIM);
    rd(fi, &skin, 1);
    rd(fi, &dt, 1);
    vector<double> x(n * NDIM), v(n * NDIM), m(n), par(n * 3);
    rd(fi, x.data(), x.size());
    rd(fi, v.data(), v.size());
    rd(fi, m.data(), m.size());
    rd(fi, par.data(), par.size());
    fclose(fi);

    Vec Lv;
    for (uint d = 0; d < NDIM; d++) Lv[d] = L[d];
    boost::shared_ptr<OriginBox> obox(new OriginBox(Lv));
    boost::shared_ptr<AtomVec> atomptr(new AtomVec(m));
    AtomVec &atoms = *atomptr;
    for (uint i = 0; i < n; i++)
        for (uint d = 0; d < NDIM; d++) {
            atoms[i].x[d] = x[i * NDIM + d];  // Atom& writes, as user code does
            atoms[i].v[d] = v[i * NDIM + d];
        }
    boost::shared_ptr<Interaction> inter;
    boost::shared_ptr<NeighborList> nl;
    if (kind == PARM_PAIR_LJCUT) {
        sptr<NListed<EpsSigCutAtom, LennardJonesCutPair> > I = make<EpsSigCutAtom, LennardJonesCutPair>(obox, atomptr, skin);
        for (uint i = 0; i < n; i++) I->add(EpsSigCutAtom(atoms.get_id(i), par[3 * i], par[3 * i + 1], par[3 * i + 2]));
        inter = I;
        nl = I->neighbor_list();
    } else if (kind == PARM_PAIR_LJREPULSE) {
        sptr<NListed<EpsSigAtom, LJRepulsivePair> > I = make<EpsSigAtom, LJRepulsivePair>(obox, atomptr, skin);
        for (uint i = 0; i < n; i++) I->add(EpsSigAtom(atoms.get_id(i), par[3 * i], par[3 * i + 1]));
        inter = I;
        nl = I->neighbor_list();
    } else if (kind == PARM_PAIR_REPULSION) {
        sptr<NListed<EpsSigExpAtom, RepulsionPair> > I = make<EpsSigExpAtom, RepulsionPair>(obox, atomptr, skin);
        for (uint i = 0; i < n; i++) I->add(EpsSigExpAtom(atoms.get_id(i), par[3 * i], par[3 * i + 1], par[3 * i + 2]));
        inter = I;
        nl = I->neighbor_list();
    } else {
        sptr<NListed<IEpsSigCutAtom, LJAttractRepulsePair> > I = make<IEpsSigCutAtom, LJAttractRepulsePair>(obox, atomptr, skin);
        vector<flt> eps(1, 1.0);
        for (uint i = 0; i < n; i++) I->add(IEpsSigCutAtom(atoms.get_id(i), eps, 0, par[3 * i + 1], par[3 * i + 2]));
        inter = I;
        nl = I->neighbor_list();
    }
    nl->update_list(true);
    uint32_t np0 = nl->numpairs();

    CollectionVerlet collec = CollectionVerlet(boost::static_pointer_cast<Box>(obox), atomptr, dt);
    collec.add_tracker(nl);
    collec.add_interaction(inter);
    // statistics trackers on the device (constraints.hpp:260-414), registered like any StateTracker
    vector<unsigned long> lags;
    lags.push_back(1);
    lags.push_back(5);
    sptr<RsqTracker> rsq(new RsqTracker(atomptr, lags, true));
    sptr<ISFTracker> isf(new ISFTracker(atomptr, vector<flt>(1, 1.5), vector<unsigned long>(1, 3), false));
    sptr<EnergyTracker> et(new EnergyTracker(atomptr, vector<sptr<Interaction> >(1, inter), 2));
    collec.add_tracker(rsq);
    collec.add_tracker(isf);
    collec.add_tracker(et);
    collec.set_forces(true);
    double out[8];
    out[0] = collec.energy();
    out[1] = collec.kinetic_energy();
    out[2] = inter->energy(*obox);
    for (int s = 0; s < steps; s++) collec.timestep();
    out[3] = collec.energy();
    out[4] = collec.kinetic_energy();
    out[5] = inter->energy(*obox);
    out[6] = collec.pressure();
    out[7] = collec.temp();
    uint32_t u[2] = {np0, nl->which()};
    vector<double> xo(n * NDIM), vo(n * NDIM), fo(n * NDIM);
    for (uint i = 0; i < n; i++)
        for (uint d = 0; d < NDIM; d++) {
            xo[i * NDIM + d] = atoms[i].x[d];  // Atom& reads pull the device state lazily
            vo[i * NDIM + d] = atoms[i].v[d];
            fo[i * NDIM + d] = atoms[i].f[d];
        }
    // IDPair iteration (host-side consumers of the list)
    size_t cnt = 0;
    for (vector<IDPair>::iterator it = nl->begin(); it != nl->end(); ++it) cnt += it->first().n() > it->last().n();
    if (cnt != nl->numpairs()) { fprintf(stderr, "pair iteration mismatch\n"); return 3; }
    FILE *fo_ = fopen(argv[2], "wb");
    fwrite(out, 8, 8, fo_);
    fwrite(u, 4, 2, fo_);
    fwrite(xo.data(), 8, xo.size(), fo_);
    fwrite(vo.data(), 8, vo.size(), fo_);
    fwrite(fo.data(), 8, fo.size(), fo_);
    {
        vector<flt> cnt = rsq->counts();
        vector<vector<flt> > r2 = rsq->r2(), r4 = rsq->r4();
        vector<vector<vector<cmplx> > > I = isf->ISFs();
        vector<double> re(n), im(n);
        for (uint i = 0; i < n; i++) { re[i] = I[0][0][i].real(); im[i] = I[0][0][i].imag(); }
        double e5[5] = {(double)et->n(), et->E(), et->U(), et->K(), et->E_std()};
        fwrite(cnt.data(), 8, 2, fo_);
        fwrite(r2[0].data(), 8, n, fo_);
        fwrite(r2[1].data(), 8, n, fo_);
        fwrite(r4[1].data(), 8, n, fo_);
        fwrite(re.data(), 8, n, fo_);
        fwrite(im.data(), 8, n, fo_);
        fwrite(e5, 8, 5, fo_);
    }
    fclose(fo_);
    printf("facade_run ok: n=%u pairs=%u E0=%.12g E1=%.12g which=%u\n", n, np0, out[0], out[3], u[1]);
    return 0;
}
