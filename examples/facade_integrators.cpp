// The extra integrators through the parm_b200 C++ facade (same constructors as the reference's collection.hpp).
//   facade_integrators <in.bin> <out.bin>
// in.bin : int32 n, integ (PARM_INTEG_*), steps, nparams; double L[NDIM], skin, dt, params[nparams];
//          double x[n][NDIM], v[n][NDIM], m[n], sigma[n]
// out.bin: double E, K, U, xi, lns; uint32 which; double x[n][NDIM], v[n][NDIM]
// Interaction: NListed<EpsSigExpAtom, RepulsionPair> with eps 1, exponent 2.5 (Hertzian), per-atom sigma.
#include <cstdio>
#include <cstdlib>

#include "collection.hpp"
#include "interaction.hpp"

template <class T>
static void rd(FILE *f, T *p, size_t n) {
    if (n && fread(p, sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
}

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
    FILE *fi = fopen(argv[1], "rb");
    if (!fi) { perror("in"); return 2; }
    int hdr[4];
    rd(fi, hdr, 4);
    const uint n = (uint)hdr[0];
    const int integ = hdr[1], steps = hdr[2], np = hdr[3];
    double L[NDIM], skin, dt, par[6] = {0, 0, 0, 0, 0, 0};
    rd(fi, L, NDIM);
    rd(fi, &skin, 1);
    rd(fi, &dt, 1);
    rd(fi, par, np);
    vector<double> x(n * NDIM), v(n * NDIM), m(n), sig(n);
    rd(fi, x.data(), x.size());
    rd(fi, v.data(), v.size());
    rd(fi, m.data(), m.size());
    rd(fi, sig.data(), sig.size());
    fclose(fi);

    Vec Lv;
    for (uint d = 0; d < NDIM; d++) Lv[d] = L[d];
    sptr<OriginBox> obox(new OriginBox(Lv));
    sptr<Box> box = boost::static_pointer_cast<Box>(obox);
    sptr<AtomVec> atomptr(new AtomVec(m));
    AtomVec &atoms = *atomptr;
    for (uint i = 0; i < n; i++)
        for (uint d = 0; d < NDIM; d++) {
            atoms[i].x[d] = x[i * NDIM + d];
            atoms[i].v[d] = v[i * NDIM + d];
        }
    sptr<NListed<EpsSigExpAtom, RepulsionPair> > I(new NListed<EpsSigExpAtom, RepulsionPair>(obox, atomptr, skin));
    for (uint i = 0; i < n; i++) I->add(EpsSigExpAtom(atoms.get_id(i), 1.0, sig[i], 2.5));
    sptr<NeighborList> nl = I->neighbor_list();
    nl->update_list(true);

    sptr<Collection> collec;
    CollectionNoseHoover *nh = NULL;
    switch (integ) {
        case PARM_INTEG_DAMPED: collec.reset(new CollectionDamped(box, atomptr, dt, par[0])); break;
        case PARM_INTEG_SOLHT: collec.reset(new CollectionSolHT(box, atomptr, dt, par[0], par[1])); break;
        case PARM_INTEG_OVERDAMPED: collec.reset(new CollectionOverdamped(box, atomptr, dt, par[0])); break;
        case PARM_INTEG_NOSEHOOVER: collec.reset(nh = new CollectionNoseHoover(box, atomptr, dt, par[0], par[1])); break;
        case PARM_INTEG_GAUSSIANT: collec.reset(new CollectionGaussianT(box, atomptr, dt)); break;
        case PARM_INTEG_GEAR3A: collec.reset(new CollectionGear3A(box, atomptr, dt)); break;
        case PARM_INTEG_GEAR4A: collec.reset(new CollectionGear4A(box, atomptr, dt, (uint)par[0])); break;
        case PARM_INTEG_GEAR5A: collec.reset(new CollectionGear5A(box, atomptr, dt, (uint)par[0])); break;
        case PARM_INTEG_GEAR6A: collec.reset(new CollectionGear6A(box, atomptr, dt, (uint)par[0])); break;
        case PARM_INTEG_NLCG: {  // par = (P0, kappa, kmax, secmax, seceps), then packmin.py:73-76
            CollectionNLCG *cg = new CollectionNLCG(obox, atomptr, dt, par[0], vector<sptr<Interaction> >(),
                                                    vector<sptr<StateTracker> >(), vector<sptr<Constraint> >(), par[1], par[2],
                                                    (uint)par[3], par[4]);
            collec.reset(cg);
            cg->set_max_alpha(2.0);
            cg->set_max_dx(10.0);
            cg->set_max_step(1e-3);
            break;
        }
        default: fprintf(stderr, "facade_integrators: type %d not handled here\n", integ); return 2;
    }
    collec->add_tracker(nl);
    collec->add_interaction(I);
    collec->set_forces(true);
    for (int s = 0; s < steps; s++) collec->timestep();
    double out[5] = {collec->energy(), collec->kinetic_energy(), collec->potential_energy(), 0, 0};
    if (nh) {
        out[3] = nh->get_xi();
        out[4] = nh->get_lns();
    }
    if (integ == PARM_INTEG_NLCG) {  // report the box the minimiser arrived at
        out[3] = obox->V();
        out[4] = obox->L();
    }
    uint32_t which = nl->which();
    vector<double> xo(n * NDIM), vo(n * NDIM);
    for (uint i = 0; i < n; i++)
        for (uint d = 0; d < NDIM; d++) {
            xo[i * NDIM + d] = atoms[i].x[d];
            vo[i * NDIM + d] = atoms[i].v[d];
        }
    FILE *fo = fopen(argv[2], "wb");
    fwrite(out, 8, 5, fo);
    fwrite(&which, 4, 1, fo);
    fwrite(xo.data(), 8, xo.size(), fo);
    fwrite(vo.data(), 8, vo.size(), fo);
    fclose(fo);
    printf("facade_integrators ok: integ=%d n=%u E=%.12g which=%u\n", integ, n, out[0], which);
    return 0;
}
