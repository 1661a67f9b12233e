// Collection / CollectionVerlet / CollectionSol on the device.
//   CollectionVerlet::timestep   collection.cpp:442-469
//   CollectionSol::timestep      collection.cpp:265-322, set_constants :230-263,
//                                BivariateGauss::{set,gen_vecs} vecrand.cpp:48-85
//   Collection::{initialize,update_trackers,set_forces,potential_energy,virial}
//                                collection.cpp:13-19, 45-50, 159-179, 98-108, 73-80
// Each step is: K1 (first half-kick + drift of positions, fused with the NeighborList skin-drift
// reduction) -> force kernel(s) -> K3 (a = f/m, second half-kick). The streaming kernels
// reproduce the reference's expression order without FMA contraction so that, given the same
// forces, positions and velocities are bit-identical to the CPU path.
#include <math.h>
#include <string.h>

#include <algorithm>

#include <stdlib.h>
#include "drift.cuh"
#include "rng.cuh"
#include "internal.cuh"

#define I_BLOCK 256

static inline unsigned grid_for(const parm_ctx *ctx, uint32_t n, unsigned block, unsigned per_sm = 8) {
    unsigned need = (n + block - 1) / block;
    unsigned cap = (unsigned)ctx->num_sms * per_sm;
    if (need < 1) need = 1;
    return need < cap ? need : cap;
}

// ---- K1: x += v*dt + a*(dt*dt/2); v += a*(dt/2)   (collection.cpp:443-451) -------------
// fused with the NeighborList skin-drift reduction (trackers.cpp:23-53): the displacement only depends on
// x(t+dt), which this kernel produces, so the rebuild decision of the step is known while the forces of the
// step are still being computed (and, sharded, its all-gather overlaps the force kernel).
// PREL: also writes the image-resolved copy of x(t+dt) that the bulk-copy staged pair kernel reads (tile.cu, k_tile_prep)
struct PrelOut {
    const float4 *img;
    double2 *xy;
    double *z;
    double L[3];
};
template <int D, bool DRIFT, bool PREL>
__global__ void __launch_bounds__(I_BLOCK)
k_verlet1(double4 *__restrict__ pos, double *__restrict__ v, const double *__restrict__ a, uint32_t n, uint32_t npad,
          double dt, double hdt2, double hdt, const int *__restrict__ abort_flag, const double *__restrict__ xlast,
          double skin, double *d_top2, unsigned int *counter, NlistFlags *dflags, NlistFlags *hflags, int *d_slot,
          int *h_slot, const PrelOut R, const int step_no) {
    if (abort_flag && *abort_flag) return; // speculative step behind a rebuild request: leave the state alone
    double b1 = 0.0, b2 = 0.0;
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        // every load of the atom is issued before the first use (one memory round trip per atom, not two)
        double4 p = pos[s];
        double vd[D], ad[D], xl[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int d = 0; d < D; d++) {
            vd[d] = v[(size_t)d * npad + s];
            ad[d] = a[(size_t)d * npad + s];
        }
        if (DRIFT) {
            xl[0] = xlast[s];
            xl[1] = xlast[npad + s];
            xl[2] = xlast[2 * (size_t)npad + s];
        }
        float4 im = make_float4(0.f, 0.f, 0.f, 0.f);
        if (PREL) im = R.img[s];
        if (frozen_le(p.w)) { // m <= 0 || isinf(m): v = 0, position untouched
#pragma unroll
            for (int d = 0; d < D; d++) v[(size_t)d * npad + s] = 0.0;
        } else {
            double x[3] = {p.x, p.y, p.z};
#pragma unroll
            for (int d = 0; d < D; d++) {
                x[d] = __dadd_rn(x[d], __dadd_rn(__dmul_rn(vd[d], dt), __dmul_rn(ad[d], hdt2)));
                v[(size_t)d * npad + s] = __dadd_rn(vd[d], __dmul_rn(ad[d], hdt));
            }
            p.x = x[0];
            p.y = x[1];
            if (D == 3) p.z = x[2];
            pos[s] = p;
        }
        if (PREL) {
            R.xy[s] = make_double2(fma(-(double)im.x, R.L[0], p.x), fma(-(double)im.y, R.L[1], p.y));
            R.z[s] = fma(-(double)im.z, R.L[2], p.z);
        }
        // atoms that were never add()ed have NaN lastlocs (set at rebuild): NaN never wins a '>' comparison
        if (DRIFT) top2_push(b1, b2, drift_dist(p, xl[0], xl[1], xl[2]));
    }
    if (DRIFT) drift_finish(b1, b2, skin, d_top2, counter, dflags, hflags, d_slot, h_slot, step_no);
}

// ---- K3: a = f/m; v += a*(dt/2)  (collection.cpp:457-465) -----------------------------------
template <int D>
__global__ void __launch_bounds__(I_BLOCK)
k_verlet2(const double4 *__restrict__ pos, double *__restrict__ v, double *__restrict__ a, const double *__restrict__ f,
          uint32_t n, uint32_t npad, double hdt, const int *__restrict__ abort_flag) {
    if (abort_flag && *abort_flag) return;
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        const double m = pos[s].w;
        double fd[D], vd[D];
#pragma unroll
        for (int d = 0; d < D; d++) { // loads issued before the branch on m
            fd[d] = f[(size_t)d * npad + s];
            vd[d] = v[(size_t)d * npad + s];
        }
        if (frozen_le(m)) {
#pragma unroll
            for (int d = 0; d < D; d++) a[(size_t)d * npad + s] = 0.0;
        } else {
#pragma unroll
            for (int d = 0; d < D; d++) {
                const size_t q = (size_t)d * npad + s;
                const double ad = __ddiv_rn(fd[d], m);
                a[q] = ad;
                v[q] = __dadd_rn(vd[d], __dmul_rn(ad, hdt));
            }
        }
    }
}

// ---- K3 of step s fused with K1 of step s+1 (inside a multi-step timestep(n) call) ----------------------------------
// One pass over the atoms instead of two: reads f, v, x (+m), lastlocs and writes a, v, x (and the image-resolved copy
// of x for the pair kernel): 224 bytes per atom against 328 for K3 followed by K1. `abort_flag` is the guard of step s
// (the decision of step s-1), `next_flag` the decision of step s itself -- taken by K1 of step s, i.e. before this
// kernel started: when it asks for a rebuild only the K3 half runs, the host rebuilds and starts step s+1 with a plain
// K1. Same expressions, in the same order, as k_verlet2 followed by k_verlet1: bit-identical trajectories.
template <int D, bool PREL>
__global__ void __launch_bounds__(I_BLOCK)
k_verlet21(double4 *__restrict__ pos, double *__restrict__ v, double *__restrict__ a, const double *__restrict__ f, uint32_t n,
           uint32_t npad, double dt, double hdt2, double hdt, const int *__restrict__ abort_flag, const int *__restrict__ next_flag,
           const double *__restrict__ xlast, double skin, double *d_top2, unsigned int *counter, NlistFlags *dflags,
           NlistFlags *hflags, int *d_slot, int *h_slot, const PrelOut R, const int step_no) {
    // An aborted step hands the abort on (its own decision slot is what guards the step after it): with several steps
    // queued behind a rebuild request every one of them stays a no-op.
    if (abort_flag && *abort_flag) {
        if (d_slot && blockIdx.x == 0 && threadIdx.x == 0) *d_slot = 1;
        return;
    }
    const bool go = !(*next_flag);
    if (!go && d_slot && blockIdx.x == 0 && threadIdx.x == 0) *d_slot = 1; // K3 half only: step s+1 does not start
    double b1 = 0.0, b2 = 0.0;
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        double4 p = pos[s];
        double fd[D], vd[D], xl[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int d = 0; d < D; d++) {
            fd[d] = f[(size_t)d * npad + s];
            vd[d] = v[(size_t)d * npad + s];
        }
        float4 im = make_float4(0.f, 0.f, 0.f, 0.f);
        if (go) {
            xl[0] = xlast[s];
            xl[1] = xlast[npad + s];
            xl[2] = xlast[2 * (size_t)npad + s];
            if (PREL) im = R.img[s];
        }
        if (frozen_le(p.w)) { // K3: a = 0; K1: v = 0, position untouched
#pragma unroll
            for (int d = 0; d < D; d++) {
                a[(size_t)d * npad + s] = 0.0;
                if (go) v[(size_t)d * npad + s] = 0.0;
            }
        } else {
            double x[3] = {p.x, p.y, p.z};
#pragma unroll
            for (int d = 0; d < D; d++) {
                const size_t q = (size_t)d * npad + s;
                const double ad = __ddiv_rn(fd[d], p.w);                    // K3: a = f / m
                a[q] = ad;
                double vn = __dadd_rn(vd[d], __dmul_rn(ad, hdt));           //     v += a dt/2
                if (go) {
                    x[d] = __dadd_rn(x[d], __dadd_rn(__dmul_rn(vn, dt), __dmul_rn(ad, hdt2))); // K1: x += v dt + a dt^2/2
                    vn = __dadd_rn(vn, __dmul_rn(ad, hdt));                                    //     v += a dt/2
                }
                v[q] = vn;
            }
            if (go) {
                p.x = x[0];
                p.y = x[1];
                if (D == 3) p.z = x[2];
                pos[s] = p;
            }
        }
        if (go) {
            if (PREL) {
                R.xy[s] = make_double2(fma(-(double)im.x, R.L[0], p.x), fma(-(double)im.y, R.L[1], p.y));
                R.z[s] = fma(-(double)im.z, R.L[2], p.z);
            }
            top2_push(b1, b2, drift_dist(p, xl[0], xl[1], xl[2]));
        }
    }
    if (go) drift_finish(b1, b2, skin, d_top2, counter, dflags, hflags, d_slot, h_slot, step_no);
}

struct SolConst {
    double dt, c0, c1dt, c2dtdt, dtc1mc2, dtc2, x11, x21, x22, desT, damping;
};

// ---- Sol K1 (collection.cpp:276-298), fused with the drift reduction like k_verlet1 --------------
template <int D, bool DRIFT>
__global__ void __launch_bounds__(I_BLOCK, 4)
k_sol1(double4 *__restrict__ pos, double *__restrict__ v, const double *__restrict__ a, const uint32_t *__restrict__ order,
       uint32_t n, uint32_t npad, SolConst K, const double *__restrict__ noise, const uint32_t *__restrict__ mobile_rank,
       uint64_t step, uint64_t seed, const int *__restrict__ abort_flag, const double *__restrict__ xlast, double skin,
       double *d_top2, unsigned int *counter, NlistFlags *dflags, NlistFlags *hflags, int *d_slot, int *h_slot) {
    if (abort_flag && *abort_flag) return;
    double b1 = 0.0, b2 = 0.0;
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        // every load of the atom is issued before the ~500 instructions of the noise generation: the memory round trip
        // hides behind them instead of following them (ncu r02_d: 35 % active warps at 80 registers, long-scoreboard bound)
        double4 p = pos[s];
        double vd0[D], ad0[D], xl[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int d = 0; d < D; d++) {
            vd0[d] = v[(size_t)d * npad + s];
            ad0[d] = a[(size_t)d * npad + s];
        }
        if (DRIFT) {
            xl[0] = xlast[s];
            xl[1] = xlast[npad + s];
            xl[2] = xlast[2 * (size_t)npad + s];
        }
        const uint32_t id = K.damping > 0 ? order[s] : 0u;
        if (frozen_le(p.w)) {
#pragma unroll
            for (int d = 0; d < D; d++) v[(size_t)d * npad + s] = 0.0;
        } else {
            const double v0 = __dsqrt_rn(__ddiv_rn(K.desT, p.w));
            const double r0 = __dmul_rn(K.dt, v0);
            double x1[3] = {0, 0, 0}, x2[3] = {0, 0, 0};
            if (K.damping > 0) {
                if (noise) {
                    const double *z = noise + (size_t)mobile_rank[id] * 2 * D;
#pragma unroll
                    for (int d = 0; d < D; d++) {
                        x1[d] = z[d];
                        x2[d] = z[D + d];
                    }
                } else {
                    double za, zb;
                    normal_pair(id, step, 0, seed, x1[0], x1[1]);
                    normal_pair(id, step, 1, seed, x2[0], x2[1]);
                    normal_pair(id, step, 2, seed, za, zb);
                    if (D == 3) {
                        x1[2] = za;
                        x2[2] = zb;
                    }
                }
            }
            double x[3] = {p.x, p.y, p.z};
#pragma unroll
            for (int d = 0; d < D; d++) {
                const size_t q = (size_t)d * npad + s;
                const double vd = vd0[d], ad = ad0[d];
                const double drG = __dmul_rn(x1[d], K.x11);                                           // x1 * x11
                const double dvG = __dadd_rn(__dmul_rn(x1[d], K.x21), __dmul_rn(x2[d], K.x22));       // x1*x21 + x2*x22
                x[d] = __dadd_rn(x[d], __dadd_rn(__dadd_rn(__dmul_rn(vd, K.c1dt), __dmul_rn(ad, K.c2dtdt)), __dmul_rn(drG, r0)));
                v[q] = __dadd_rn(__dadd_rn(__dmul_rn(vd, K.c0), __dmul_rn(ad, K.dtc1mc2)), __dmul_rn(dvG, v0));
            }
            p.x = x[0];
            p.y = x[1];
            if (D == 3) p.z = x[2];
            pos[s] = p;
        }
        if (DRIFT) top2_push(b1, b2, drift_dist(p, xl[0], xl[1], xl[2]));
    }
    if (DRIFT) drift_finish(b1, b2, skin, d_top2, counter, dflags, hflags, d_slot, h_slot);
}

// ---- Sol K3 (collection.cpp:303-318): note the m == 0 (not m <= 0) tests ---------------------
template <int D>
__global__ void __launch_bounds__(I_BLOCK)
k_sol2(const double4 *__restrict__ pos, double *__restrict__ v, double *__restrict__ a, const double *__restrict__ f,
       uint32_t n, uint32_t npad, double dtc2, const int *__restrict__ abort_flag) {
    if (abort_flag && *abort_flag) return;
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        const double m = pos[s].w;
        if (frozen_eq(m)) {
#pragma unroll
            for (int d = 0; d < D; d++) {
                a[(size_t)d * npad + s] = 0.0;
                v[(size_t)d * npad + s] = 0.0;
            }
        } else {
#pragma unroll
            for (int d = 0; d < D; d++) {
                const size_t q = (size_t)d * npad + s;
                const double ad = __ddiv_rn(f[q], m);
                a[q] = ad;
                v[q] = __dadd_rn(v[q], __dmul_rn(ad, dtc2));
            }
        }
    }
}

// ---- a = f/m of Collection::set_forces(true) (collection.cpp:171-178) -----------------------
template <int D>
__global__ void k_accel(const double4 *__restrict__ pos, double *__restrict__ a, const double *__restrict__ f, uint32_t n, uint32_t npad) {
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        const double m = pos[s].w;
        const bool fr = frozen_le(m);
#pragma unroll
        for (int d = 0; d < D; d++) {
            const size_t q = (size_t)d * npad + s;
            a[q] = fr ? 0.0 : __ddiv_rn(f[q], m);
        }
    }
}

// ---- host ---------------------------------------------------------------------------------------
static int sol_set_constants(parm_integ *g) { // collection.cpp:230-263 + vecrand.cpp:48-63
    if (g->force_mag <= 0.0) {
        g->c0 = 1; g->c1 = 1; g->c2 = .5; g->sigmar = 0; g->sigmav = 0; g->corr = 1;
    } else {
        double dampdt = g->force_mag * g->dt;
        g->c0 = exp(-dampdt);
        g->c1 = (-expm1(-dampdt)) / dampdt;
        g->c2 = (1 - g->c1) / dampdt;
        if (dampdt > 1e-4)
            g->sigmar = sqrt((1 / dampdt) * (2 - (-4 * expm1(-dampdt) + expm1(-2 * dampdt)) / dampdt));
        else
            g->sigmar = sqrt(2 * dampdt / 3 - dampdt * dampdt / 2 + 7 * dampdt * dampdt * dampdt / 30);
        g->sigmav = sqrt(-expm1(-2 * dampdt));
        double exdpdt = (-expm1(-dampdt));
        g->corr = exdpdt * exdpdt / dampdt / g->sigmar / g->sigmav;
    }
    if (!(g->sigmar >= 0)) { parm_set_error("BivariateGauss::set: s1 >= 0"); return PARM_ERR_INVALID; }
    if (!(g->sigmav >= 0)) { parm_set_error("BivariateGauss::set: s2 >= 0"); return PARM_ERR_INVALID; }
    if (!(g->corr >= 0)) { parm_set_error("BivariateGauss::set: corr >= 0"); return PARM_ERR_INVALID; }
    if (!(g->corr <= 1)) { parm_set_error("BivariateGauss::set: corr <= 1"); return PARM_ERR_INVALID; }
    g->x11 = g->sigmar;
    g->x21 = g->sigmav * g->corr;
    g->x22 = g->sigmav * sqrt(1 - g->corr * g->corr);
    return 0;
}

extern "C" int parm_verlet_create(parm_ctx *c, double dt, parm_integ **out) {
    if (!c || !out) { parm_set_error("parm_verlet_create: NULL argument"); return PARM_ERR_INVALID; }
    parm_integ *g = new parm_integ();
    g->ctx = c;
    g->type = 0;
    g->dt = dt;
    *out = g;
    return 0;
}

extern "C" int parm_sol_create(parm_ctx *c, double dt, double damping, double T, uint64_t seed, parm_integ **out) {
    if (!c || !out) { parm_set_error("parm_sol_create: NULL argument"); return PARM_ERR_INVALID; }
    *out = 0;
    if (dt <= 0) { parm_set_error("Collection::CollectionSol: dt >= 0"); return PARM_ERR_INVALID; } // collection.cpp:222-224
    parm_integ *g = new parm_integ();
    g->ctx = c;
    g->type = 1;
    g->dt = dt;
    g->damping = damping;
    g->force_mag = damping;
    g->desT = T;
    g->seed = seed;
    int r = sol_set_constants(g);
    if (r) { delete g; return r; }
    *out = g;
    return 0;
}

extern "C" int parm_integ_destroy(parm_integ *g) {
    if (!g) return 0;
    cudaSetDevice(g->ctx->device);
    cudaStreamSynchronize(g->ctx->stream);
    if (g->d_noise) cudaFree(g->d_noise);
    if (g->d_mobile_rank) cudaFree(g->d_mobile_rank);
    if (g->nlcg) parm_nlcg_free(g);
    if (g->small) parm_small_free(g);
    if (g->d_gear) cudaFree(g->d_gear);
    if (g->d_scal) cudaFree(g->d_scal);
    if (g->d_xpart) cudaFree(g->d_xpart);
    if (g->ev_ok) { cudaEventDestroy(g->ev[0]); cudaEventDestroy(g->ev[1]); }
    delete g;
    return 0;
}

extern "C" int parm_integ_update_trackers(parm_integ *g) {
    for (parm_nlist *nl : g->trackers) {
        int rebuilt = 0;
        PTRY(parm_nlist_update(nl, 0, &rebuilt));
        if (rebuilt) g->rebuilds++;
    }
    for (parm_tracker *t : g->stat_trackers) PTRY(parm_tracker_enqueue_update(t, nullptr));
    return 0;
}

extern "C" int parm_integ_register_stat_tracker(parm_integ *g, parm_tracker *t) {
    if (!g || !t) { parm_set_error("parm_integ_register_stat_tracker: NULL argument"); return PARM_ERR_INVALID; }
    g->stat_trackers.push_back(t);
    return 0;
}
extern "C" int parm_integ_add_stat_tracker(parm_integ *g, parm_tracker *t) {
    PTRY(parm_integ_register_stat_tracker(g, t));
    return parm_integ_update_trackers(g); // collection.hpp:117-120
}

extern "C" int parm_integ_register_interaction(parm_integ *g, parm_inter *it) {
    if (!g || !it) { parm_set_error("parm_integ_register_interaction: NULL argument"); return PARM_ERR_INVALID; }
    if (it->ctx != g->ctx) { parm_set_error("interaction belongs to another AtomVec"); return PARM_ERR_INVALID; }
    g->inters.push_back(it);
    return 0;
}
extern "C" int parm_integ_register_tracker(parm_integ *g, parm_nlist *nl) {
    if (!g || !nl) { parm_set_error("parm_integ_register_tracker: NULL argument"); return PARM_ERR_INVALID; }
    if (nl->ctx != g->ctx) { parm_set_error("tracker belongs to another AtomVec"); return PARM_ERR_INVALID; }
    g->trackers.push_back(nl);
    return 0;
}
extern "C" int parm_integ_add_interaction(parm_integ *g, parm_inter *it) {
    PTRY(parm_integ_register_interaction(g, it));
    return parm_integ_update_trackers(g); // collection.hpp:113-116
}
extern "C" int parm_integ_add_tracker(parm_integ *g, parm_nlist *nl) {
    PTRY(parm_integ_register_tracker(g, nl));
    return parm_integ_update_trackers(g); // collection.hpp:117-120
}

static int launch_all_forces(parm_integ *g, const int *abort_flag = nullptr, uint32_t first = 0, uint32_t count = 0xffffffffu) {
    parm_ctx *c = g->ctx;
    // atoms->reset_forces(); for each interaction: set_forces(box)   (collection.cpp:160-166)
    if (g->inters.empty()) return parm_reset_forces(c);
    bool first_inter = true;
    for (parm_inter *it : g->inters) {
        PTRY(parm_inter_launch_forces(it, 0, !first_inter, nullptr, abort_flag, first, count)); // first interaction overwrites f == reset + add
        first_inter = false;
    }
    return 0;
}

int parm_integ_launch_all_forces(parm_integ *g, const int *abort_flag) { return launch_all_forces(g, abort_flag); }

// Collection::set_forces (collection.cpp:159-179), non-virtual part
static int base_set_forces(parm_integ *g, int constraints_and_a) {
    parm_ctx *c = g->ctx;
    CK(cudaSetDevice(c->device));
    PTRY(launch_all_forces(g));
    const uint32_t no = parm_owned(c);
    if (!constraints_and_a || no == 0) return 0;
    if (c->D == 3) k_accel<3><<<grid_for(c, no, 256), 256, 0, c->stream>>>(c->pos, c->a, c->f, no, c->npad);
    else k_accel<2><<<grid_for(c, no, 256), 256, 0, c->stream>>>(c->pos, c->a, c->f, no, c->npad);
    CK_LAUNCH(c);
    return 0;
}

extern "C" int parm_integ_set_forces(parm_integ *g, int constraints_and_a) {
    if (g->type == PARM_INTEG_NLCG) return parm_nlcg_set_forces(g, constraints_and_a, 1); // collection.hpp:449-451
    if (g->type == PARM_INTEG_GAUSSIANT) { // CollectionGaussianT::set_forces(bool) -> set_forces(true, true), collection.hpp:618
        PTRY(base_set_forces(g, 1));
        return parm_integ_extra_after_set_forces(g);
    }
    return base_set_forces(g, constraints_and_a);
}

extern "C" int parm_integ_initialize(parm_integ *g) { // collection.cpp:13-19
    PTRY(parm_integ_update_trackers(g));
    // called from the base-class constructor in the reference: the base set_forces, not a derived override
    PTRY(base_set_forces(g, 1));
    return parm_integ_update_trackers(g);
}

extern "C" int parm_integ_set_dt(parm_integ *g, double dt) {
    g->dt = dt;
    if (g->type == PARM_INTEG_SOL) return sol_set_constants(g);
    return 0;
}
extern "C" int parm_integ_set_temperature(parm_integ *g, double damping, double T) {
    if (g->type != 1) { parm_set_error("change_temperature: not a CollectionSol"); return PARM_ERR_INVALID; }
    g->damping = damping;
    g->force_mag = damping;
    g->desT = T;
    return sol_set_constants(g);
}
extern "C" int parm_integ_get_sol_constants(parm_integ *g, double *cst) {
    cst[0] = g->c0; cst[1] = g->c1; cst[2] = g->c2; cst[3] = g->x11; cst[4] = g->x21; cst[5] = g->x22;
    return 0;
}

extern "C" int parm_integ_inject_noise(parm_integ *g, const double *z, size_t len) {
    parm_ctx *c = g->ctx;
    CK(cudaSetDevice(c->device));
    if (c->sh.on) { parm_set_error("noise injection is a single-GPU parity hook"); return PARM_ERR_UNSUPPORTED; }
    CK(cudaStreamSynchronize(c->stream));
    if (g->d_noise) cudaFree(g->d_noise);
    g->d_noise = 0;
    g->noise_len = g->noise_pos = 0;
    if (!z || !len) return 0;
    // rank of every mobile atom in AtomVec order (the reference draws in atom order, skipping frozen ones)
    std::vector<double> m(c->n);
    PTRY(parm_download_atoms(c, PARM_M, 0, 0, 0, 0, m.data(), (size_t)c->D * 8, 8));
    std::vector<uint32_t> rank(c->npad, 0);
    uint32_t r = 0;
    for (uint32_t i = 0; i < c->n; i++) {
        rank[i] = r;
        if (!(m[i] <= 0 || isinf(m[i]))) r++;
    }
    g->n_mobile = r;
    if (!g->d_mobile_rank) CK(cudaMalloc(&g->d_mobile_rank, (size_t)c->npad * 4));
    CK(cudaMemcpy(g->d_mobile_rank, rank.data(), (size_t)c->npad * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&g->d_noise, len * 8));
    CK(cudaMemcpy(g->d_noise, z, len * 8, cudaMemcpyHostToDevice));
    g->noise_len = len;
    g->noise_step0 = g->steps;
    return 0;
}

// Enqueue the kernels of step number `step` (K1 -> halo exchange -> forces -> K3 + drift). With a
// tracker, the rebuild decision of the step is left in decision slot `slot` (device word + pinned host
// word). abort_flag (device, may be NULL) is the decision word of the PREVIOUS step: when it is set the
// kernels of this step return immediately, so a step can be enqueued before the host has seen whether
// its predecessor asked for a rebuild.
// k1_done: the first half of this step already ran inside the previous step's fused K3+K1 kernel. fuse_next: end this
// step with that fused kernel (K3 of this step + K1 of the next, the latter guarded by THIS step's decision and leaving
// the next step's decision in slot (slot + 1) % 3) instead of a plain K3. Both only for CollectionVerlet (fusable()).
static int enqueue_step_core(parm_integ *g, uint64_t step, const int *abort_flag, int slot, bool k1_done, bool fuse_next, int batch_step);
static bool fusable(const parm_integ *g) {
    const parm_ctx *c = g->ctx;
    const char *e = getenv("PARM_B200_FUSE_K3K1"); // (read per call: the sweeps toggle it inside one process)
    return g->type == 0 && !g->trackers.empty() && g->stat_trackers.empty() && !c->prof_on && (e ? atoi(e) != 0 : true);
}
static int enqueue_step(parm_integ *g, uint64_t step, const int *abort_flag, int slot, bool k1_done = false, bool fuse_next = false,
                        int batch_step = -1) {
    if (g->type >= PARM_INTEG_DAMPED) PTRY(parm_integ_extra_enqueue(g, step, abort_flag, slot));
    else PTRY(enqueue_step_core(g, step, abort_flag, slot, k1_done, fuse_next, batch_step));
    // update_trackers() ends every step: the statistics trackers follow the NeighborList
    for (parm_tracker *t : g->stat_trackers) PTRY(parm_tracker_enqueue_update(t, abort_flag));
    return 0;
}
static int enqueue_step_core(parm_integ *g, uint64_t step, const int *abort_flag, int slot, bool k1_done, bool fuse_next, int batch_step) {
    parm_ctx *c = g->ctx;
    const uint32_t n = parm_owned(c); // ghost copies are never integrated
    parm_nlist *nl = g->trackers.empty() ? nullptr : g->trackers[0];
    // Verlet K1 is a pure stream with a block reduction at its end: 4 blocks per SM (6-7 atoms per thread at 1e6
    // atoms) measured best (0.034 ms; 16 per SM: 0.039). The Langevin K1 generates its noise in the kernel and keeps 16.
    static const unsigned k1_per_sm = getenv("PARM_B200_K1_PER_SM") ? (unsigned)atoi(getenv("PARM_B200_K1_PER_SM")) : 4u;
    const unsigned grid = grid_for(c, n, I_BLOCK, 16);
    const unsigned grid1 = std::min(grid_for(c, n, I_BLOCK, g->type == 0 ? std::max(k1_per_sm, 1u) : 16u), 4096u); // drift_finish: d_top2 holds 4096 block entries
    const double dt = g->dt;
    // sharded: K1 only leaves the local top-2; the global decision is folded after the all-gather
    int *d_slot = nl && !c->sh.on ? nl->d_slot + slot : nullptr;
    int *h_slot = nl && !c->sh.on ? nl->h_slot + slot : nullptr;
#define DRIFTARGS abort_flag, nl ? nl->xlast : nullptr, nl ? nl->skin : 0.0, nl ? nl->d_top2 : nullptr, \
                  nl ? nl->d_counter : nullptr, nl ? nl->d_flags : nullptr, nl ? nl->h_flags : nullptr, d_slot, h_slot
    SolConst K;
    // CollectionVerlet whose interactions all use the tracked list, staged by bulk copies: K1 itself leaves the
    // image-resolved copy of x(t+dt) for the pair kernel (one pass over the positions fewer per step)
    bool k1_prel = g->type == 0 && nl && c->D == 3 && nl->tile.valid && nl->tile.stage_aligned && !g->inters.empty();
    for (parm_inter *it : g->inters) k1_prel = k1_prel && it->nl == nl;
    const char *ef = getenv("PARM_B200_K1_PREL"); // (read per step: the sweeps toggle it inside one process)
    const int fuse_env = ef ? atoi(ef) : 1;
    k1_prel = k1_prel && fuse_env;
    PrelOut R;
    R.img = nl ? nl->tile.img : nullptr;
    R.xy = nl ? nl->tile.prel_xy : nullptr;
    R.z = nl ? nl->tile.prel_z : nullptr;
    for (int d = 0; d < 3; d++) R.L[d] = c->box.L[d];
    PTRY(parm_prof_begin(c, PARM_PROF_INTEG1));
    if (k1_done) {
        // x(t+dt), v(t+dt/2), the drift reduction (and prel) of this step were produced by the previous step's fused kernel
    } else if (g->type == 0) {
        const double hdt2 = dt * dt / 2, hdt = dt / 2;
        if (c->D == 3) {
            if (k1_prel) k_verlet1<3, true, true><<<grid1, I_BLOCK, 0, c->stream>>>(c->pos, c->v, c->a, n, c->npad, dt, hdt2, hdt, DRIFTARGS, R, batch_step);
            else if (nl) k_verlet1<3, true, false><<<grid1, I_BLOCK, 0, c->stream>>>(c->pos, c->v, c->a, n, c->npad, dt, hdt2, hdt, DRIFTARGS, R, batch_step);
            else k_verlet1<3, false, false><<<grid1, I_BLOCK, 0, c->stream>>>(c->pos, c->v, c->a, n, c->npad, dt, hdt2, hdt, DRIFTARGS, R, batch_step);
        } else {
            if (nl) k_verlet1<2, true, false><<<grid1, I_BLOCK, 0, c->stream>>>(c->pos, c->v, c->a, n, c->npad, dt, hdt2, hdt, DRIFTARGS, R, batch_step);
            else k_verlet1<2, false, false><<<grid1, I_BLOCK, 0, c->stream>>>(c->pos, c->v, c->a, n, c->npad, dt, hdt2, hdt, DRIFTARGS, R, batch_step);
        }
    } else {
        K.dt = dt;
        K.c0 = g->c0;
        K.c1dt = g->c1 * dt;
        K.c2dtdt = g->c2 * dt * dt;
        K.dtc1mc2 = dt * (g->c1 - g->c2);
        K.dtc2 = dt * g->c2;
        K.x11 = g->x11;
        K.x21 = g->x21;
        K.x22 = g->x22;
        K.desT = g->desT;
        K.damping = g->damping;
        const double *noise = nullptr;
        if (g->d_noise) {
            const size_t per = (size_t)g->n_mobile * 2 * c->D;
            const size_t off = (size_t)(step - g->noise_step0) * per;
            if (off + per > g->noise_len) { parm_set_error("CollectionSol: injected noise exhausted"); return PARM_ERR_INVALID; }
            noise = g->d_noise + off;
        }
#define S1ARGS c->pos, c->v, c->a, c->order, n, c->npad, K, noise, g->d_mobile_rank, step, g->seed, DRIFTARGS
        if (c->D == 3) {
            if (nl) k_sol1<3, true><<<grid1, I_BLOCK, 0, c->stream>>>(S1ARGS);
            else k_sol1<3, false><<<grid1, I_BLOCK, 0, c->stream>>>(S1ARGS);
        } else {
            if (nl) k_sol1<2, true><<<grid1, I_BLOCK, 0, c->stream>>>(S1ARGS);
            else k_sol1<2, false><<<grid1, I_BLOCK, 0, c->stream>>>(S1ARGS);
        }
#undef S1ARGS
    }
    if (!k1_done) CK_LAUNCH(c);
    PTRY(parm_prof_end(c));

    PTRY(parm_prof_begin(c, PARM_PROF_FORCE));
    if (c->sh.on) {
        // communication stream: ghost positions for x(t+dt) and the global drift decision, overlapped with the
        // forces of the atoms whose neighbours are all owned (everything but the two boundary layers)
        // bulk-copy staged tile kernels read prel: owned slots are refreshed here, behind K1 and ahead of the event the
        // communication stream waits for; the ghost slots behind the exchange on that stream (below)
        std::vector<parm_nlist *> lists;
        for (parm_inter *it : g->inters)
            if (std::find(lists.begin(), lists.end(), it->nl) == lists.end()) lists.push_back(it->nl);
        for (parm_nlist *l : lists)
            if (!(k1_prel && l == nl)) PTRY(parm_tile_prep(l, 0, n, c->stream, abort_flag));
        c->tile_prep_external = true;
        int rcs = parm_shard_step_comm(c, nl, nl ? nl->d_slot + slot : nullptr, nl ? nl->h_slot + slot : nullptr);
        const uint32_t lo = c->sh.s_dn, hi = n - c->sh.s_up;
        if (!rcs && hi > lo) rcs = launch_all_forces(g, abort_flag, lo, hi - lo);
        if (rcs) { c->tile_prep_external = false; return rcs; }
        // the two boundary layers follow the exchange on the communication stream itself (highest priority), so
        // their small grids share the GPU with the interior kernel instead of running as two partial waves after it
        {
            cudaStream_t main_stream = c->stream;
            c->stream = c->sh.comm_stream;
            int rc = 0;
            for (parm_nlist *l : lists)
                if (!rc && c->n > n) rc = parm_tile_prep(l, n, c->n - n, c->stream, abort_flag);
            if (!rc && lo) rc = launch_all_forces(g, abort_flag, 0, lo);
            if (!rc && n > hi) rc = launch_all_forces(g, abort_flag, std::max(hi, lo), n - std::max(hi, lo));
            c->stream = main_stream;
            c->tile_prep_external = false;
            PTRY(rc);
        }
        PTRY(parm_shard_step_join(c)); // records the end of the communication stream's work; the main stream waits for it
    } else {
        c->tile_prep_external = k1_prel;
        const int rcf = launch_all_forces(g, abort_flag);
        c->tile_prep_external = false;
        PTRY(rcf);
    }
    PTRY(parm_prof_end(c));

    PTRY(parm_prof_begin(c, PARM_PROF_INTEG2));
    if (g->type == 0 && fuse_next && nl) {
        // K3 of this step + K1 of the next in one pass; the next step's decision goes to slot (slot + 1) % 3 (sharded:
        // folded from the all-gather by the next step's communication phase, as for a plain K1)
        const int nslot = (slot + 1) % 3;
        int *d_slot2 = !c->sh.on ? nl->d_slot + nslot : nullptr, *h_slot2 = !c->sh.on ? nl->h_slot + nslot : nullptr;
        const double hdt2 = dt * dt / 2, hdt = dt / 2;
#define F21ARGS c->pos, c->v, c->a, c->f, n, c->npad, dt, hdt2, hdt, abort_flag, nl->d_slot + slot, nl->xlast, nl->skin, nl->d_top2, \
                nl->d_counter, nl->d_flags, nl->h_flags, d_slot2, h_slot2, R, batch_step >= 0 ? batch_step + 1 : -1
        if (c->D == 3) {
            if (k1_prel) k_verlet21<3, true><<<grid1, I_BLOCK, 0, c->stream>>>(F21ARGS);
            else k_verlet21<3, false><<<grid1, I_BLOCK, 0, c->stream>>>(F21ARGS);
        } else {
            k_verlet21<2, false><<<grid1, I_BLOCK, 0, c->stream>>>(F21ARGS);
        }
#undef F21ARGS
    } else if (g->type == 0) {
        if (c->D == 3) k_verlet2<3><<<grid, I_BLOCK, 0, c->stream>>>(c->pos, c->v, c->a, c->f, n, c->npad, dt / 2, abort_flag);
        else k_verlet2<2><<<grid, I_BLOCK, 0, c->stream>>>(c->pos, c->v, c->a, c->f, n, c->npad, dt / 2, abort_flag);
    } else {
        if (c->D == 3) k_sol2<3><<<grid, I_BLOCK, 0, c->stream>>>(c->pos, c->v, c->a, c->f, n, c->npad, K.dtc2, abort_flag);
        else k_sol2<2><<<grid, I_BLOCK, 0, c->stream>>>(c->pos, c->v, c->a, c->f, n, c->npad, K.dtc2, abort_flag);
    }
    CK_LAUNCH(c);
    PTRY(parm_prof_end(c));
#undef DRIFTARGS
    return 0;
}

extern "C" int parm_integ_timestep(parm_integ *g, int nsteps) {
    if (!g) { parm_set_error("parm_integ_timestep: NULL integrator"); return PARM_ERR_INVALID; }
    parm_ctx *c = g->ctx;
    CK(cudaSetDevice(c->device));
    if (c->n == 0 && !c->sh.on) { g->steps += nsteps > 0 ? nsteps : 0; return 0; }
    for (size_t k = 1; k < g->trackers.size(); k++)
        if (g->trackers[k] != g->trackers[0]) { parm_set_error("one NeighborList per Collection is supported"); return PARM_ERR_UNSUPPORTED; }
    if (nsteps <= 0) return 0;
    if (g->type == PARM_INTEG_NLCG) { // host-driven secant / conjugate-gradient loop (nlcg.cu)
        for (int s = 0; s < nsteps; s++) {
            PTRY(parm_nlcg_timestep(g));
            g->steps++;
        }
        return 0;
    }
    if (g->type == PARM_INTEG_NOSEHOOVER) { // Collection::degrees_of_freedom(), collection.cpp:116-133
        double out[4];
        PTRY(parm_reduce(c, PARM_RED_NDOF, nullptr, out));
        g->ndof_cached = out[0];
    }
    parm_nlist *nl = g->trackers.empty() ? nullptr : g->trackers[0];
    if (!nl) { // no tracker: nothing to decide, the steps just queue up
        for (int s = 0; s < nsteps; s++) {
            PTRY(enqueue_step(g, g->steps, nullptr, 0));
            g->steps++;
        }
        return 0;
    }
    if (!g->ev_ok) {
        CK(cudaEventCreateWithFlags(&g->ev[0], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&g->ev[1], cudaEventDisableTiming));
        g->ev_ok = true;
    }
    static int speculate = -1;
    if (speculate < 0) { const char *e = getenv("PARM_B200_SPECULATE"); speculate = e ? atoi(e) : 1; }
    // update_trackers() ends every step (collection.cpp:468): NeighborList::update -> update_list(false).
    // The host runs one step ahead of the decisions: step s+1 is enqueued (guarded by the decision word
    // of step s) before the host waits for step s. A rebuild request makes the guarded kernels no-ops;
    // the host then rebuilds and enqueues step s+1 again.
    // Decision slots are used round robin modulo 3: a fused K3+K1 kernel reads the slots of the two previous steps and
    // writes the third. With CollectionVerlet every step but the last of the call ends with that fused kernel.
    const bool fuse = fusable(g);
    // Batched stepping (single GPU, CollectionVerlet): up to `batch` steps are queued without the host looking at a single
    // decision in between. Every step is guarded on the device by its predecessor's decision slot, an aborted step hands
    // the abort on, and the step whose drift rule fired leaves its number in pinned memory: the host waits ONCE per
    // batch, counts the steps that ran, rebuilds if one of them asked for it and queues the rest again. Semantics are
    // those of collection.cpp:442-469 step by step (the rebuild still follows the step that triggered it); what changes
    // is that small systems are no longer bound by one host round trip per step.
    {
        const char *eb = getenv("PARM_B200_STEP_BATCH");
        // (measured, profiles/README.md r02: no gain -- N = 1000 is bound by ~20 us of dependent kernel latency per step,
        // not by the host round trip, and at N = 1e6 the kernels queued behind a trigger cost 4 %: off unless asked for)
        const int batch = eb ? atoi(eb) : 0;
        if (fuse && batch > 1 && !c->sh.on && !nl->ignorechanged && nsteps > 1) {
            int s = 0;
            while (s < nsteps) {
                const int B = std::min(nsteps - s, batch);
                nl->h_flags->trigger = 0xffffffffu;
                for (int j = 0; j < B; j++)
                    PTRY(enqueue_step(g, g->steps + j, j ? nl->d_slot + (j - 1) % 3 : nullptr, j % 3, j > 0, j + 1 < B, j));
                CK(cudaEventRecord(g->ev[0], c->stream));
                CK(cudaEventSynchronize(g->ev[0]));
                const uint32_t trig = nl->h_flags->trigger;
                const int done = trig < (uint32_t)B ? (int)trig + 1 : B;
                g->steps += done;
                s += done;
                if (trig < (uint32_t)B) {
                    PTRY(parm_nlist_rebuild(nl));
                    g->rebuilds++;
                }
                CK(cudaMemsetAsync(nl->d_slot, 0, 4 * sizeof(int), c->stream));
                nl->h_slot[0] = nl->h_slot[1] = nl->h_slot[2] = nl->h_slot[3] = 0;
            }
            return 0;
        }
    }
    // small systems: the whole call as one persistent kernel (small.cu); what it leaves over takes the general path
    if (nsteps >= 2) {
        PTRY(parm_small_run(g, nl, &nsteps));
        if (nsteps <= 0) return 0;
    }
    int cur = 0; // slot of step s
    PTRY(enqueue_step(g, g->steps, nullptr, cur, false, fuse && nsteps > 1));
    CK(cudaEventRecord(g->ev[0], c->stream));
    for (int s = 0; s < nsteps; s++) {
        const int p = s & 1, nxt = (cur + 1) % 3;
        // (per-class event timing counts launches, so it runs without speculation)
        // (statistics trackers keep host-side step counters: their steps are never enqueued speculatively)
        const bool spec = speculate && !c->prof_on && s + 1 < nsteps && !nl->ignorechanged && g->stat_trackers.empty();
        const bool fuse_after_next = fuse && s + 2 < nsteps; // step s+1 ends with a fused kernel as well
        if (spec) {
            // step s+1, guarded by step s's decision; its K1 ran inside step s's fused kernel when there was one
            PTRY(enqueue_step(g, g->steps + 1, nl->d_slot + cur, nxt, fuse, fuse_after_next));
            CK(cudaEventRecord(g->ev[p ^ 1], c->stream));
        }
        CK(cudaEventSynchronize(g->ev[p]));
        const bool rebuild = nl->ignorechanged || nl->h_slot[cur] != 0;
        g->steps++;
        if (rebuild) {
            PTRY(parm_nlist_rebuild(nl)); // synchronises the stream: the guarded kernels of step s+1 have returned
            g->rebuilds++;
            CK(cudaMemsetAsync(nl->d_slot, 0, 4 * sizeof(int), c->stream));
            nl->h_slot[0] = nl->h_slot[1] = nl->h_slot[2] = nl->h_slot[3] = 0;
        }
        if (s + 1 < nsteps && (rebuild || !spec)) {
            // after a rebuild the fused kernel only did its K3 half: step s+1 starts with a plain K1. Without speculation
            // and without a rebuild the K1 half did run.
            PTRY(enqueue_step(g, g->steps, nullptr, nxt, fuse && !rebuild, fuse_after_next));
            CK(cudaEventRecord(g->ev[p ^ 1], c->stream));
        }
        cur = nxt;
    }
    return 0;
}

extern "C" int parm_integ_potential_energy(parm_integ *g, double *E) {
    double tot = 0;
    for (parm_inter *it : g->inters) {
        double e;
        PTRY(parm_inter_energy(it, &e));
        tot += e;
    }
    *E = tot;
    return 0;
}
extern "C" int parm_integ_virial(parm_integ *g, double *w) {
    double tot = 0;
    for (parm_inter *it : g->inters) {
        double p;
        PTRY(parm_inter_pressure(it, &p));
        tot += p;
    }
    *w = tot;
    return 0;
}
extern "C" int parm_integ_stats(parm_integ *g, uint64_t *steps, uint64_t *rebuilds, uint64_t *launches) {
    if (steps) *steps = g->steps;
    if (rebuilds) *rebuilds = g->rebuilds;
    if (launches) *launches = g->ctx->launches;
    return 0;
}
