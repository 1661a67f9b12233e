/* parm_b200 -- C ABI of the B200-native ParM hot path.
 *
 * ParM has no FFI/plugin registry: its boundary is the public C++ class API
 * (SURVEY.md 8b). The look-alike C++ headers in parm_b200/include/parm/ keep that
 * API source-compatible and call ONLY the functions declared here; this file is
 * therefore the exact set of entry points a binding (the C++ facade, a SWIG/ctypes
 * module, ...) needs.  Every entry point cites the reference interface it replaces
 * (paths relative to the reference's src/).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; the message is
 *     available from parm_b200_last_error() (thread-local).  Error classes mirror
 *     the reference's exception types (SURVEY 8b "Error convention").
 *   - all pointers are HOST pointers; vectors are NDIM doubles; "n x NDIM" arrays
 *     are addressed with an explicit byte stride so that both ParM's AoS
 *     `struct Atom` (box.hpp:234-249) and plain numpy arrays can be passed.
 *   - atoms are always addressed by their AtomVec index (AtomID::n(), box.hpp:288-298);
 *     the library re-orders them internally (cell order) and hides that.
 *   - a context is bound to one CUDA device and one host thread (the reference is
 *     single-threaded and not re-entrant; same contract).
 *   - there is NO CPU fallback: without a usable CUDA device every call fails
 *     with PARM_ERR_CUDA.
 */
#ifndef PARM_B200_H
#define PARM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PARM_OK 0
#define PARM_ERR_INVALID 1   /* std::invalid_argument in the reference */
#define PARM_ERR_RUNTIME 2   /* std::runtime_error */
#define PARM_ERR_CUDA 3      /* CUDA / NCCL failure (-> std::runtime_error) */
#define PARM_ERR_UNSUPPORTED 4 /* feature outside the hot-path scope */

typedef struct parm_ctx parm_ctx;     /* OriginBox + AtomVec device state */
typedef struct parm_nlist parm_nlist; /* NeighborList */
typedef struct parm_inter parm_inter; /* NListed<A,P> */
typedef struct parm_integ parm_integ;
typedef struct parm_tracker parm_tracker; /* Collection{Verlet,Sol} */

const char *parm_b200_last_error(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
uint64_t parm_b200_launch_count(void);
const char *parm_b200_version(void);
/* roofline denominators measured by this library's own micro-kernels (csrc/probe.cu): fp64 DFMA-chain
 * throughput in GFLOP/s (2 flops per DFMA) and the bandwidth of a 1 GiB -> 1 GiB streaming copy in GB/s
 * (read + write bytes). No reference counterpart; used by bench.py next to MEASURED_PEAKS.json. */
int parm_b200_probe_peaks(int device, double *fp64_gflops, double *copy_gbs);

/* ---- context: AtomVec(N, m) box.hpp:441-479 + OriginBox(L) box.hpp:97-158 ---- */
int parm_ctx_create(int ndim, uint32_t n_atoms, int device, parm_ctx **out);
int parm_ctx_destroy(parm_ctx *ctx);
/* OriginBox::boxsize (box.hpp:99-101); OriginBox::resize_to (box.cpp:21-25) re-calls this */
int parm_set_box(parm_ctx *ctx, const double *L);
int parm_get_box(parm_ctx *ctx, double *L);
/* OriginBox::diff for a batch of point pairs, on the device (box.hpp:103, vec_mod :69-78) */
int parm_box_diff(parm_ctx *ctx, uint32_t npts, const double *r1, const double *r2, double *out);

#define PARM_X 1u
#define PARM_V 2u
#define PARM_A 4u
#define PARM_F 8u
#define PARM_M 16u
#define PARM_ALL 31u
/* Atom fields (box.hpp:234-249) host->device / device->host. Pointers address the
 * field of atom 0; consecutive atoms are stride_vec (x,v,a,f) / stride_m (m) BYTES
 * apart. For ParM's AoS pass &atoms[0].x, ... with both strides = sizeof(Atom).
 * A whole struct-Atom array (mask PARM_ALL, AoS) travels as ONE copy on the context's stream and parm_upload_atoms
 * returns without waiting for it: when that array is page-locked (parm_host_register), the caller must leave it
 * unchanged until the next call that synchronises (parm_sync, parm_download_atoms, any reduction or query). Every
 * other layout is staged and has been consumed when the call returns. */
int parm_upload_atoms(parm_ctx *ctx, unsigned mask, const double *x, const double *v, const double *a,
                      const double *f, const double *m, size_t stride_vec, size_t stride_m);
int parm_download_atoms(parm_ctx *ctx, unsigned mask, double *x, double *v, double *a, double *f, double *m,
                        size_t stride_vec, size_t stride_m);
/* page-lock a host buffer (the AtomVec mirror) so transfers run at PCIe speed */
int parm_host_register(void *ptr, size_t bytes);
int parm_host_unregister(void *ptr);
int parm_sync(parm_ctx *ctx);

/* Asynchronous trajectory frame (the XYZ output of LJatoms.cpp:130-158 and pyparm/xyzfile.py:7-76 without stopping the
 * run): _begin gathers x and/or v (mask of PARM_X | PARM_V) by AtomVec index on the context's stream and starts their
 * device->host copy on a second stream; parm_integ_timestep calls issued afterwards overlap that copy. _wait blocks
 * until the frame has arrived and copies it to x / v (host, n * ndim doubles each, either may be NULL). One frame at a
 * time per context. */
int parm_snapshot_begin(parm_ctx *ctx, unsigned mask);
int parm_snapshot_wait(parm_ctx *ctx, double *x, double *v);

/* Grid::get_loc (trackers.cpp:192-219, class Grid trackers.hpp:227-309) of every atom, computed on the device from the
 * resident positions: loc[i] = cell of AtomVec index i for `widths` (ndim entries) divisions per axis; the facade
 * builds Grid::gridlocs from it. */
int parm_grid_locs(parm_ctx *ctx, const uint32_t *widths, uint32_t *loc);
/* the CUDA stream (cudaStream_t) every kernel of this context is launched on, so a caller
 * can bracket calls with its own CUDA events */
int parm_get_stream(parm_ctx *ctx, void **cuda_stream);
/* per-kernel-class device timing with CUDA events on that stream (bench.py rooflines) */
#define PARM_PROF_INTEG1 0  /* K1 first half step */
#define PARM_PROF_FORCE 1   /* pair force kernel(s) */
#define PARM_PROF_INTEG2 2  /* K3 second half step + drift reduction */
#define PARM_PROF_REBUILD 3 /* bin + sort + permute + neighbour build */
#define PARM_PROF_N 4
int parm_profile_enable(parm_ctx *ctx, int enable);
/* waits for the stream, adds up the recorded intervals per class since the last read */
int parm_profile_read(parm_ctx *ctx, double *ms /*[PARM_PROF_N]*/, uint64_t *intervals /*[PARM_PROF_N]*/);

/* ---- AtomGroup reductions box.cpp:239-260, 401-431; Collection collection.cpp:21-29,116-133 ---- */
#define PARM_RED_MASS 0      /* out[1]      AtomGroup::mass,  skips m<=0||inf */
#define PARM_RED_MOMENTUM 1  /* out[NDIM]   AtomGroup::momentum */
#define PARM_RED_KE 2        /* out[1]      AtomGroup::kinetic_energy(v0), skips m==0||inf */
#define PARM_RED_COM 3       /* out[NDIM]   AtomGroup::com (sum x m / mass) */
#define PARM_RED_NDOF 4      /* out[1]      NDIM * #mobile atoms, Collection::degrees_of_freedom */
#define PARM_RED_COMFORCE 5  /* out[NDIM]   AtomGroup::com_force */
int parm_reduce(parm_ctx *ctx, int what, const double *v0, double *out);
int parm_scale_velocities(parm_ctx *ctx, double scaleby); /* Collection::scale_velocities collection.cpp:21-29 */
int parm_add_velocity(parm_ctx *ctx, const double *dv);   /* AtomGroup::add_velocity box.cpp:413-417 */
int parm_reset_forces(parm_ctx *ctx);                     /* AtomGroup::reset_forces box.cpp:427-431 */

/* ---- NeighborList trackers.hpp:157-214, trackers.cpp:10-85 ---- */
int parm_nlist_create(parm_ctx *ctx, double skin, parm_nlist **out);
int parm_nlist_destroy(parm_nlist *nl);
/* NeighborList::add(AtomID, diameter) for every atom at once: diam[i] < 0 or NaN
 * means atom i was never add()ed. Sets ignorechanged (trackers.hpp:194-201). */
int parm_nlist_set_diameters(parm_nlist *nl, const double *diam);
/* NeighborList::update_list(force) trackers.cpp:19-85: drift rule then pair build. */
int parm_nlist_update(parm_nlist *nl, int force, int *rebuilt);
int parm_nlist_which(parm_nlist *nl, uint32_t *updatenum);  /* which() */
/* ignore(AtomID a, AtomID b) for npairs pairs of AtomVec indices (trackers.hpp:190-193): the pair never enters
 * the list (trackers.cpp:64); sets ignorechanged, so the next update_list() rebuilds. ignore_size(): :204 */
int parm_nlist_ignore(parm_nlist *nl, const uint32_t *a, const uint32_t *b, uint64_t npairs);
int parm_nlist_ignore_size(parm_nlist *nl, uint64_t *n);
int parm_nlist_numpairs(parm_nlist *nl, uint64_t *npairs);  /* numpairs() */
/* curpairs in the reference's order (trackers.cpp:59-68): first = later atom i,
 * last = earlier atom j < i, sorted by (i, j).  cap = capacity of both arrays. */
int parm_nlist_download_pairs(parm_nlist *nl, uint32_t *first, uint32_t *last, uint64_t cap);
/* mean / max neighbours per atom in the device (full) list, for bench rooflines */
int parm_nlist_stats(parm_nlist *nl, double *mean_full_neighbors, uint32_t *max_full_neighbors);
/* cell-tile layout of the last rebuild (the shared-memory staged pair kernel):
 * active = 1 when single-species Lennard-Jones interactions on this list run on the tile kernel;
 * chunks = blocks per launch, max_tile_atoms = largest number of positions one block stages,
 * wide_chunks = chunks whose tile spans more than half the box (per-pair minimum image kept). */
int parm_nlist_tile_stats(parm_nlist *nl, int *active, uint32_t *chunks, uint32_t *max_tile_atoms, uint32_t *wide_chunks);

/* ---- NListed<A,P> interaction.hpp:1876-1945, 2102-2291 ---- */
#define PARM_PAIR_LJREPULSE 0        /* NListed<EpsSigAtom, LJRepulsePair>       :857-891, 119-152 */
#define PARM_PAIR_REPULSION 1        /* NListed<EpsSigExpAtom, RepulsionPair>    :1454-1465, 1528-1550 */
#define PARM_PAIR_LJATTRACTREPULSE 2 /* NListed<IEpsSigCutAtom, LJAttractRepulsePair> :989-1018, 1251-1299 */
#define PARM_PAIR_LJCUT 3            /* NListed<EpsSigCutAtom, LennardJonesCutPair>   :897-905, 967-987, 238-281 */
/* the remaining NListed functors of sim.i:621-643 */
#define PARM_PAIR_LJATTRACTCUT 4           /* LJAttractCutPair :1020-1049 + LJAttractCut :197-236 */
#define PARM_PAIR_LJATTRACTFIXEDREPULSE 5  /* NListed<IEpsRepsSigCutAtom, LJAttractFixedRepulsePair> :1307-1413 */
#define PARM_PAIR_EISMCLACHLAN 6           /* NListed<EisMclachlanAtom, EisMclachlanPair> :1415-1452 */
#define PARM_PAIR_LJISH 7                  /* NListed<IEpsRepsSigExpCutAtom, LJishPair> :1051-1142 */
#define PARM_PAIR_LJATTRACTREPULSESIGS 8   /* NListed<EpsEpsSigSigCutAtom, LJAttractRepulseSigsPair> :1149-1244 */
#define PARM_PAIR_REPULSIONDRAG 9          /* NListed<EpsSigExpDragAtom, RepulsionDragPair> :1598-1642 */
#define PARM_PAIR_LOISOHERN 10             /* NListed<LoisOhernAtom, LoisOhernPair> :1679-1744 */
#define PARM_PAIR_LOISLIN 11               /* NListed<LoisLinAtom, LoisLinPair> :1764-1829 */
#define PARM_PAIR_LOISOHERNMIN 12          /* NListed<LoisOhernAtom, LoisOhernPairMinCLs> :1746-1752 */
#define PARM_PAIR_LOISLINMIN 13            /* NListed<LoisLinAtom, LoisLinPairMin> :1831-1837 */
#define PARM_PAIR_NKINDS 14
#define PARM_PAIR_MAXPARAMS 5
int parm_inter_create(parm_ctx *ctx, parm_nlist *nl, int pair_kind, parm_inter **out);
int parm_inter_destroy(parm_inter *inter);
/* per-atom A structs for all atoms at once. params is n x 3 doubles:
 *   kind 0: (epsilon, sigma, -)   kind 1: (eps, sigma, exponent)
 *   kind 2: (-, sigma, sigcut) + type[i] = IEpsSigCutAtom::indx and the symmetric
 *           ntypes x ntypes table eps_table[t1*ntypes+t2] = epsilons[indx]
 *   kind 3: (epsilon, sigma, sigcut)
 * The library also forwards A::max_size() to the NeighborList (NListed::add,
 * interaction.hpp:1906-1910) when set_diameters != 0. member: optional n bytes. */
int parm_inter_set_params(parm_inter *inter, const double *params, const uint32_t *type, const double *eps_table,
                          int ntypes, const uint8_t *member, int set_diameters);
/* General form: params is n x nper doubles (nper <= PARM_PAIR_MAXPARAMS), per kind
 *   0  LJREPULSE              EpsSigAtom             (epsilon, sigma)
 *   1  REPULSION              EpsSigExpAtom          (eps, sigma, exponent)
 *                             IEpsISigExpAtom        (-, -, exponent) + eps_table + sig_table   [RepulsionII]
 *   2  LJATTRACTREPULSE       IEpsSigCutAtom         (-, sigma, sigcut) + eps_table
 *   3  LJCUT                  EpsSigCutAtom          (epsilon, sigma, sigcut)
 *                             IEpsISigCutAtom        (-, -, sigcut) + eps_table + sig_table     [LJIICut]
 *   4  LJATTRACTCUT           EpsSigCutAtom          (epsilon, sigma, sigcut)                   [LJAttractCut]
 *                             IEpsSigCutAtom         (-, sigma, sigcut) + eps_table             [LJAttractICut]
 *                             IEpsISigCutAtom        (-, -, sigcut) + eps_table + sig_table     [LJAttractIICut]
 *   5  LJATTRACTFIXEDREPULSE  IEpsRepsSigCutAtom     (-, sig, sigcut, repeps) + eps_table
 *   6  EISMCLACHLAN           EisMclachlanAtom       (sigmai, dist)
 *   7  LJISH                  IEpsRepsSigExpCutAtom  (-, sigma, sigcut, repeps, exponent) + eps_table
 *   8  LJATTRACTREPULSESIGS   EpsEpsSigSigCutAtom    (eps_r, sig_r, sigcut, eps_a, sig_a)
 *   9  REPULSIONDRAG          EpsSigExpDragAtom      (eps, sigma, exponent, gamma)
 *   10 LOISOHERN / 12 ..MIN   LoisOhernAtom          (eps, sigma, C, l)
 *   11 LOISLIN / 13 ..MIN     LoisLinAtom            (eps, sigma, f, l)   f = depth/width as the ctor stores it
 * type[i] is the atom's `indx`; eps_table / sig_table are ntypes x ntypes, row t = the `epsilons` / `sigmas`
 * vector carried by atoms of indx t, and must be symmetric (the reference asserts it for IEpsSigCutAtom,
 * :1013-1014; for the other indexed structs an asymmetric table would make the result depend on pair order). */
int parm_inter_set_params_ex(parm_inter *inter, const double *params, int nper, const uint32_t *type,
                             const double *eps_table, const double *sig_table, int ntypes, const uint8_t *member,
                             int set_diameters);
#define PARM_WANT_ENERGY 1u
#define PARM_WANT_VIRIAL 2u
#define PARM_WANT_STRESS 4u
/* set_forces / set_forces_get_pressure / set_forces_get_stress (:2166-2175, 2232-2245,
 * 2264-2277): ADDS pair forces to Atom::f. out (may be NULL if want==0) receives
 * [E][virial][stress NDIM*NDIM row-major] for the requested quantities, in that order. */
int parm_inter_set_forces(parm_inter *inter, unsigned want, double *out);
int parm_inter_energy(parm_inter *inter, double *E);        /* energy(Box&)   :2154-2163 */
int parm_inter_pressure(parm_inter *inter, double *p);      /* pressure(Box&) :2248-2261 (sum r.f) */
int parm_inter_stress(parm_inter *inter, double *stress);   /* stress(Box&)   :2279-2291 */
int parm_inter_contacts(parm_inter *inter, uint64_t *contacts, uint64_t *overlaps); /* :2126-2151 */

/* ---- Collection / CollectionVerlet / CollectionSol collection.hpp:23-130,205-263,360-374 ---- */
int parm_verlet_create(parm_ctx *ctx, double dt, parm_integ **out);
int parm_sol_create(parm_ctx *ctx, double dt, double damping, double T, uint64_t seed, parm_integ **out);
/* the other fixed-box integrators that only need set_forces(). params per type:
 *   PARM_INTEG_DAMPED      CollectionDamped      (dt, damping)        collection.cpp:324-381
 *   PARM_INTEG_SOLHT       CollectionSolHT       (dt, damping, T)     :383-440  (seed: Gaussian stream)
 *   PARM_INTEG_OVERDAMPED  CollectionOverdamped  (dt, gamma)          :471-492
 *   PARM_INTEG_NOSEHOOVER  CollectionNoseHoover  (dt, Q, T)           :1170-1249
 *   PARM_INTEG_GAUSSIANT   CollectionGaussianT   (dt)                 :1251-1299
 *   PARM_INTEG_GEAR3A      CollectionGear3A      (dt)                 :1301-1322
 *   PARM_INTEG_GEAR4A/5A/6A CollectionGear4A/5A/6A (dt, ncorrec)      :1324-1496
 * Single-GPU contexts only (the correctors move atoms between force evaluations). */
#define PARM_INTEG_VERLET 0
#define PARM_INTEG_SOL 1
#define PARM_INTEG_DAMPED 2
#define PARM_INTEG_SOLHT 3
#define PARM_INTEG_OVERDAMPED 4
#define PARM_INTEG_NOSEHOOVER 5
#define PARM_INTEG_GAUSSIANT 6
#define PARM_INTEG_GEAR3A 7
#define PARM_INTEG_GEAR4A 8
#define PARM_INTEG_GEAR5A 9
#define PARM_INTEG_GEAR6A 10
#define PARM_INTEG_NLCG 11
/* CollectionNLCG (collection.hpp:400-474, collection.cpp:494-854): conjugate-gradient minimisation of
 * H = U + P0 V over the atom coordinates and kappa ln V (the packer's minimiser, pyparm/packmin.py). The box
 * of the context changes during timestep(): read it back with parm_get_box. */
int parm_nlcg_create(parm_ctx *ctx, double dt, double P0, double kappa, double kmax, unsigned secmax, double seceps,
                     parm_integ **out);
#define PARM_NLCG_DT 0        /* set_dt (calls reset())            */
#define PARM_NLCG_P0 1        /* set_pressure_goal (calls reset()) */
#define PARM_NLCG_KAPPA 2     /* set_kappa (calls reset())         */
#define PARM_NLCG_ALPHAMAX 3  /* set_max_alpha                     */
#define PARM_NLCG_AFRAC 4     /* set_max_alpha_fraction            */
#define PARM_NLCG_DXMAX 5     /* set_max_dx                        */
#define PARM_NLCG_STEPMAX 6   /* set_max_step                      */
#define PARM_NLCG_MAXDV 7     /* public member maxdV               */
#define PARM_NLCG_KMAX 8
#define PARM_NLCG_SECMAX 9
#define PARM_NLCG_SECEPS 10
int parm_nlcg_set(parm_integ *integ, int which, double value);
/* out[16]: dt, P0, kappa, alphamax, afrac, dxmax, stepmax, maxdV, Knew, k, vl, fl, al, alpha, beta|betaused.. :
 * [0] dt [1] P0 [2] kappa [3] Knew [4] k [5] vl [6] fl [7] al [8] alpha [9] beta [10] betaused [11] dxsum
 * [12] alphavmax [13] sec [14] kmax [15] secmax */
int parm_nlcg_get(parm_integ *integ, double *out16);
int parm_nlcg_set_forces(parm_integ *integ, int constraints_and_a, int setV); /* set_forces(bool, bool) :534-568 */
int parm_nlcg_reset(parm_integ *integ);    /* :525-532 */
int parm_nlcg_descend(parm_integ *integ);  /* :837-854 */
#define PARM_NLCG_FDOTF 0
#define PARM_NLCG_FDOTA 1
#define PARM_NLCG_FDOTV 2
#define PARM_NLCG_VDOTV 3
#define PARM_NLCG_KINETIC 4      /* CollectionNLCG::kinetic_energy :574-588 */
#define PARM_NLCG_PRESSURE 5     /* CollectionNLCG::pressure :590-600       */
#define PARM_NLCG_HAMILTONIAN 6  /* :570-572                                */
int parm_nlcg_reduce(parm_integ *integ, int what, double *out);
int parm_integ_create(parm_ctx *ctx, int type, const double *params, int nparams, uint64_t seed, parm_integ **out);
/* thermostat state: out[0] = xi, out[1] = lns (CollectionNoseHoover::get_xi/get_lns; GaussianT: xi) */
int parm_integ_get_scalars(parm_integ *integ, double *out2);
int parm_integ_reset_bath(parm_integ *integ);        /* CollectionNoseHoover::reset_bath, collection.hpp:589-592 */
int parm_integ_set_param(parm_integ *integ, int which, double value); /* 1: Q (set_Q), 2: T, 3: damping, 4: gamma */
int parm_integ_destroy(parm_integ *integ);

/* ---- statistics trackers: accumulators in device memory, read on demand ---- */
/* RsqTracker(atoms, ns, usecom) constraints.hpp:342-368: per-atom <dx_j^2>, <dx_j^4>, <|dr|^4> over lags of ns[k] steps */
int parm_rsq_create(parm_ctx *ctx, const uint64_t *ns, int nns, int usecom, parm_tracker **out);
/* ISFTracker(atoms, ks, ns, usecom) constraints.hpp:393-414: per-atom, per-axis sums of exp(i k dx_j) */
int parm_isf_create(parm_ctx *ctx, const double *ks, int nks, const uint64_t *ns, int nns, int usecom, parm_tracker **out);
/* EnergyTracker(atoms, interactions, n_skip) constraints.hpp:260-316 */
int parm_energy_tracker_create(parm_ctx *ctx, parm_inter **inters, int ninters, unsigned n_skip, parm_tracker **out);
int parm_tracker_destroy(parm_tracker *t);
int parm_tracker_update(parm_tracker *t);  /* StateTracker::update(Box&) */
int parm_tracker_reset(parm_tracker *t);   /* reset() */
int parm_tracker_counts(parm_tracker *t, uint64_t *counts, int cap);  /* counts() per lag */
/* xyz2(), xyz4() (n x NDIM row-major) and r4() (n) of lag `single`, divided by its count; pointers may be NULL */
int parm_rsq_read(parm_tracker *t, int single, double *xyz2, double *xyz4, double *r4);
/* ISFxyz() of lag `single`: out[nks][n][NDIM][2] = (re, im) */
int parm_isf_read(parm_tracker *t, int single, double *out);
/* out[8] = N, Es, Us, Ks, Esq, Usq, Ksq, U0 (E() = Es/N, E_std() = sqrt(Esq/N - Es*Es/N/N) ...) */
int parm_energy_tracker_read(parm_tracker *t, double *out8);
int parm_energy_tracker_set_u0(parm_tracker *t, int from_box, double U0);  /* set_U0(flt) / set_U0(Box&) */
/* Collection::add_tracker / constructor vector for these trackers (collection.hpp:117-120): they are updated at
 * the end of every timestep(), after the NeighborList, without a host synchronisation */
int parm_integ_add_stat_tracker(parm_integ *integ, parm_tracker *t);
int parm_integ_register_stat_tracker(parm_integ *integ, parm_tracker *t);
/* add_interaction / add_tracker (collection.hpp:113-120): append, then update_trackers() */
int parm_integ_add_interaction(parm_integ *integ, parm_inter *inter);
int parm_integ_add_tracker(parm_integ *integ, parm_nlist *nl);
/* Collection constructor (collection.cpp:3-11): store the vectors as given, no update_trackers() */
int parm_integ_register_interaction(parm_integ *integ, parm_inter *inter);
int parm_integ_register_tracker(parm_integ *integ, parm_nlist *nl);
int parm_integ_initialize(parm_integ *integ);                  /* Collection::initialize collection.cpp:13-19 */
int parm_integ_set_dt(parm_integ *integ, double dt);           /* set_dt; Sol recomputes constants :230-263 */
int parm_integ_set_temperature(parm_integ *integ, double damping, double T); /* CollectionSol::change_temperature */
int parm_integ_set_forces(parm_integ *integ, int constraints_and_a); /* Collection::set_forces :159-179 */
/* nsteps x timestep() (collection.cpp:442-469 / 265-322). Returns when every rebuild decision of the call has been
 * taken: with a NeighborList that means the host has waited for the decision word of every step (an event wait per
 * step, one step behind the enqueued work), so only the tail of the last step may still be running; without a tracker
 * the steps are just enqueued. parm_sync / any download waits for the rest. Small CollectionVerlet systems (one
 * one-species Lennard-Jones interaction, N <= 4096) run the whole call as one persistent kernel and return when it has
 * finished (csrc/small.cu). */
int parm_integ_timestep(parm_integ *integ, int nsteps);
int parm_integ_update_trackers(parm_integ *integ);             /* collection.cpp:45-50 */
int parm_integ_potential_energy(parm_integ *integ, double *E); /* collection.cpp:98-108 */
int parm_integ_virial(parm_integ *integ, double *w);           /* collection.cpp:73-80 */
/* CollectionSol exact-parity hook: use these standard normals instead of the device RNG.
 * z: per step, per mobile atom in AtomVec order, NDIM normals (x1) then NDIM (x2)
 * (BivariateGauss::gen_vecs vecrand.cpp:73-85). len in doubles. NULL/0 restores the RNG. */
int parm_integ_inject_noise(parm_integ *integ, const double *z, size_t len);
int parm_integ_get_sol_constants(parm_integ *integ, double *c /* c0,c1,c2,x11,x21,x22 */);
/* steps / rebuilds executed so far, kernels launched by this integrator */
int parm_integ_stats(parm_integ *integ, uint64_t *steps, uint64_t *rebuilds, uint64_t *launches);

/* ---- slab decomposition over the GPUs of one box (new; the reference is single-core) ----
 * One process per GPU. The slab axis is x: rank r owns the atoms whose wrapped x lies in
 * [r L_x / nranks, (r+1) L_x / nranks) and keeps ghost copies of its neighbours' boundary layers.
 * Atom ids are GLOBAL AtomVec indices; per-atom parameter / diameter arrays passed to
 * parm_inter_set_params / parm_nlist_set_diameters have n_global entries on every rank.
 * Everything else (NeighborList, NListed, Collection* calls) is used unchanged and collectively:
 * every rank makes the same calls in the same order; scalars come back all-reduced. */
int parm_nccl_unique_id(void *id128);  /* rank 0 creates it, the caller broadcasts the 128 bytes */
int parm_ctx_create_sharded(int ndim, uint32_t n_global, uint32_t cap_slots, int device, int rank, int nranks,
                            const void *id128, parm_ctx **out);
/* replace this rank's local atoms: dense row-major arrays (n_local x NDIM), gid = global indices;
 * atoms must lie in, or within one cell layer of, this rank's slab (v, a, f may be NULL = 0) */
int parm_shard_set_atoms(parm_ctx *ctx, uint32_t n_local, const uint32_t *gid, const double *x, const double *v,
                         const double *a, const double *f, const double *m);
/* current local atoms in slot order; gid == NULL only queries n_local */
int parm_shard_get_atoms(parm_ctx *ctx, uint32_t cap, uint32_t *n_local, uint32_t *gid, double *x, double *v, double *a,
                         double *f, double *m);
/* overwrite fields of the local atoms in the order parm_shard_get_atoms returned them (NULL = keep) */
int parm_shard_put_atoms(parm_ctx *ctx, uint32_t n_local, const double *x, const double *v, const double *a, const double *f);
int parm_shard_info(parm_ctx *ctx, uint32_t *out6 /* n_local, ghosts_down, ghosts_up, send_down, send_up, slots */);
/* how the list rebuilds of this rank migrated their atoms so far: out2[0] one-sort rebuilds (leavers packed into
 * fixed-capacity messages, PARM_B200_SHARD_MIGCAP atoms each), out2[1] two-sort rebuilds (the fall-back when more atoms
 * than that left a slab at once anywhere, or with PARM_B200_SHARD_FAST=0). No reference counterpart (the reference is
 * single-process); NeighborList::update_list, trackers.cpp:19-85, is what both paths implement. */
int parm_shard_rebuild_stats(parm_ctx *ctx, uint64_t *out2);

#ifdef __cplusplus
}
#endif
#endif
