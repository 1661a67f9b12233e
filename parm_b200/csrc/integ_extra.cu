// The other fixed-box integrators of the reference that only need set_forces():
//   CollectionDamped::timestep      collection.cpp:356-381 (set_constants :342-354)
//   CollectionSolHT::timestep       collection.cpp:401-440
//   CollectionOverdamped::timestep  collection.cpp:471-492
//   CollectionNoseHoover::timestep  collection.cpp:1170-1242, solve_cubic :2304-2353
//   CollectionGaussianT::timestep   collection.cpp:1268-1299, set_xi :1251-1260
//   CollectionGear3A..6A::timestep  collection.cpp:1301-1496
// Each step is a short sequence of streaming kernels around the pair-force kernel(s); thermostat scalars
// (xi, lns) live on the device and are advanced by one-block kernels, so a run of steps queues up without the
// host in the loop, exactly like the Verlet / Sol path (integ.cu): every kernel honours the abort word of the
// previous step's rebuild decision, and the step ends with the NeighborList skin-drift reduction.
// The streaming kernels keep the reference's expression order without FMA contraction.
#include <math.h>
#include <string.h>

#include <algorithm>

#include "drift.cuh"
#include "rng.cuh"
#include "internal.cuh"

#define X_BLOCK 256
#define X_MAXBLOCKS 2048

static inline unsigned xgrid(const parm_ctx *ctx, uint32_t n) {
    unsigned nb = (n + X_BLOCK - 1) / X_BLOCK;
    unsigned cap = (unsigned)ctx->num_sms * 8;
    if (cap > X_MAXBLOCKS) cap = X_MAXBLOCKS;
    return nb < 1 ? 1 : (nb < cap ? nb : cap);
}

__device__ __forceinline__ bool x_frozen(double m) { return m <= 0 || isinf(m); } // collection.cpp:359, :411, :476

#define XLOOP for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x)
#define XGUARD if (abort_flag && *abort_flag) return

// x += v*A + a*B; v = v*C + a*E   (Damped :363-364, SolHT :416-419, Gear3A :1304-1305 with C = 1)
template <int D>
__global__ void __launch_bounds__(X_BLOCK)
k_x_lin1(double4 *__restrict__ pos, double *__restrict__ v, const double *__restrict__ a, uint32_t n, uint32_t npad, double A,
         double B, double C, double E, int check_frozen, const int *__restrict__ abort_flag) {
    XGUARD;
    XLOOP {
        double4 p = pos[s];
        if (check_frozen && x_frozen(p.w)) {
#pragma unroll
            for (int d = 0; d < D; d++) v[(size_t)d * npad + s] = 0.0;
            continue;
        }
        double x[3] = {p.x, p.y, p.z};
#pragma unroll
        for (int d = 0; d < D; d++) {
            const size_t q = (size_t)d * npad + s;
            const double vd = v[q], ad = a[q];
            x[d] = __dadd_rn(x[d], __dadd_rn(__dmul_rn(vd, A), __dmul_rn(ad, B)));
            v[q] = __dadd_rn(__dmul_rn(vd, C), __dmul_rn(ad, E));
        }
        p.x = x[0];
        p.y = x[1];
        if (D == 3) p.z = x[2];
        pos[s] = p;
    }
}

// thermostatted first half (no frozen-atom test in the reference):
//   GaussianT :1271-1273  x += v*dt + a*(dt*dt/2);        v += (a - (v*xi))*(dt/2)
//   NoseHoover :1199-1203 F~ = a - (v*xi); x += v*dt + F~*(dt*dt/2); v += F~*(dt/2)
template <int D, bool NOSE>
__global__ void __launch_bounds__(X_BLOCK)
k_x_thermo1(double4 *__restrict__ pos, double *__restrict__ v, const double *__restrict__ a, uint32_t n, uint32_t npad, double dt,
            double hdt2, double hdt, const IntegScalars *__restrict__ sc, const int *__restrict__ abort_flag) {
    XGUARD;
    const double xi = sc->xi;
    XLOOP {
        double4 p = pos[s];
        double x[3] = {p.x, p.y, p.z};
#pragma unroll
        for (int d = 0; d < D; d++) {
            const size_t q = (size_t)d * npad + s;
            const double vd = v[q], ad = a[q];
            const double ft = __dsub_rn(ad, __dmul_rn(vd, xi));
            x[d] = __dadd_rn(x[d], __dadd_rn(__dmul_rn(vd, dt), __dmul_rn(NOSE ? ft : ad, hdt2)));
            v[q] = __dadd_rn(vd, __dmul_rn(ft, hdt));
        }
        p.x = x[0];
        p.y = x[1];
        if (D == 3) p.z = x[2];
        pos[s] = p;
    }
}

// a = f/m; v += a*coef   (Damped :370-378 with the frozen test of set_forces(true); NoseHoover :1211-1215 and
// GaussianT :1279-1283 without)
template <int D>
__global__ void __launch_bounds__(X_BLOCK)
k_x_accel_kick(const double4 *__restrict__ pos, double *__restrict__ v, double *__restrict__ a, const double *__restrict__ f,
               uint32_t n, uint32_t npad, double coef, int check_frozen, const int *__restrict__ abort_flag) {
    XGUARD;
    XLOOP {
        const double m = pos[s].w;
        if (check_frozen && x_frozen(m)) {
#pragma unroll
            for (int d = 0; d < D; d++) a[(size_t)d * npad + s] = 0.0;
            continue;
        }
#pragma unroll
        for (int d = 0; d < D; d++) {
            const size_t q = (size_t)d * npad + s;
            const double ad = __ddiv_rn(f[q], m);
            a[q] = ad;
            v[q] = __dadd_rn(v[q], __dmul_rn(ad, coef));
        }
    }
}

// SolHT second half :425-438: a = (f + g)/m with g ~ N(0, sigma) per component; v += a*(dt/2)
template <int D>
__global__ void __launch_bounds__(X_BLOCK)
k_x_solht2(const double4 *__restrict__ pos, double *__restrict__ v, double *__restrict__ a, const double *__restrict__ f,
           const uint32_t *__restrict__ order, uint32_t n, uint32_t npad, double hdt, double sigma, const double *__restrict__ noise,
           const uint32_t *__restrict__ mobile_rank, uint64_t step, uint64_t seed, const int *__restrict__ abort_flag) {
    XGUARD;
    XLOOP {
        const double m = pos[s].w;
        if (x_frozen(m)) {
#pragma unroll
            for (int d = 0; d < D; d++) a[(size_t)d * npad + s] = 0.0;
            continue;
        }
        const uint32_t id = order[s];
        double z[4] = {0, 0, 0, 0};
        if (noise) {
            const double *zz = noise + (size_t)mobile_rank[id] * D;
#pragma unroll
            for (int d = 0; d < D; d++) z[d] = zz[d];
        } else {
            normal_pair(id, step, 3, seed, z[0], z[1]);
            if (D == 3) normal_pair(id, step, 4, seed, z[2], z[3]);
        }
#pragma unroll
        for (int d = 0; d < D; d++) {
            const size_t q = (size_t)d * npad + s;
            const double g = __dmul_rn(z[d], sigma);
            const double ad = __ddiv_rn(__dadd_rn(f[q], g), m);
            a[q] = ad;
            v[q] = __dadd_rn(v[q], __dmul_rn(ad, hdt));
        }
    }
}

// Overdamped :475-489: a = f/m; v = a*gamma (frozen: both zero); then x += v*dt for every atom
template <int D>
__global__ void __launch_bounds__(X_BLOCK)
k_x_overdamped(double4 *__restrict__ pos, double *__restrict__ v, double *__restrict__ a, const double *__restrict__ f, uint32_t n,
               uint32_t npad, double dt, double gamma, const int *__restrict__ abort_flag) {
    XGUARD;
    XLOOP {
        double4 p = pos[s];
        const bool fr = x_frozen(p.w);
        double x[3] = {p.x, p.y, p.z};
#pragma unroll
        for (int d = 0; d < D; d++) {
            const size_t q = (size_t)d * npad + s;
            const double ad = fr ? 0.0 : __ddiv_rn(f[q], p.w);
            const double vd = fr ? 0.0 : __dmul_rn(ad, gamma);
            a[q] = ad;
            v[q] = vd;
            x[d] = __dadd_rn(x[d], __dmul_rn(vd, dt));
        }
        p.x = x[0];
        p.y = x[1];
        if (D == 3) p.z = x[2];
        pos[s] = p;
    }
}

// v /= ytov   (NoseHoover :1227-1230, GaussianT :1291-1294)
template <int D>
__global__ void __launch_bounds__(X_BLOCK)
k_x_scale_v(double *__restrict__ v, uint32_t n, uint32_t npad, const IntegScalars *__restrict__ sc, const int *__restrict__ abort_flag) {
    XGUARD;
    const double y = sc->ytov;
    XLOOP {
#pragma unroll
        for (int d = 0; d < D; d++) {
            const size_t q = (size_t)d * npad + s;
            v[q] = __ddiv_rn(v[q], y);
        }
    }
}

// Gear3A corrector :1312-1317: a' = f/m; v += (a' - a)*(dt/2); a = a'
template <int D>
__global__ void __launch_bounds__(X_BLOCK)
k_x_gear3_correct(const double4 *__restrict__ pos, double *__restrict__ v, double *__restrict__ a, const double *__restrict__ f,
                  uint32_t n, uint32_t npad, double hdt, const int *__restrict__ abort_flag) {
    XGUARD;
    XLOOP {
        const double m = pos[s].w;
#pragma unroll
        for (int d = 0; d < D; d++) {
            const size_t q = (size_t)d * npad + s;
            const double an = __ddiv_rn(f[q], m);
            const double corr = __dsub_rn(an, a[q]);
            v[q] = __dadd_rn(v[q], __dmul_rn(corr, hdt));
            a[q] = an;
        }
    }
}

// Gear predictor of order Q (4A :1330-1336, 5A :1369-1377, 6A :1416-1430). Higher derivatives b, c, e (the
// reference's bs, cs, ds) are stored by AtomVec index, like the reference's vectors.
struct GearK {
    double p1, p2, p3, p4, p5;  // dt, dt^2/2, dt^3/6 ... as the reference forms them
    double c0, c1, c3, c4, c5;  // corrector coefficients
};
template <int D, int Q>
__global__ void __launch_bounds__(X_BLOCK)
k_x_gear_predict(double4 *__restrict__ pos, double *__restrict__ v, double *__restrict__ a, double *__restrict__ gear,
                 const uint32_t *__restrict__ order, uint32_t n, uint32_t npad, size_t gstride, GearK K,
                 const int *__restrict__ abort_flag) {
    XGUARD;
    XLOOP {
        double4 p = pos[s];
        const uint32_t id = order[s];
        double x[3] = {p.x, p.y, p.z};
#pragma unroll
        for (int d = 0; d < D; d++) {
            const size_t q = (size_t)d * npad + s;
            double *bp = gear + (size_t)d * gstride + id, *cp = bp + 3 * gstride, *ep = bp + 6 * gstride;
            const double vd = v[q], ad = a[q], b = *bp;
            const double c = Q >= 5 ? *cp : 0.0, e = Q >= 6 ? *ep : 0.0;
            // left-to-right sums, as the reference writes them
            double sx = __dadd_rn(__dadd_rn(__dmul_rn(vd, K.p1), __dmul_rn(ad, K.p2)), __dmul_rn(b, K.p3));
            double sv = __dadd_rn(__dmul_rn(ad, K.p1), __dmul_rn(b, K.p2));
            double sa = __dmul_rn(b, K.p1);
            if (Q >= 5) {
                sx = __dadd_rn(sx, __dmul_rn(c, K.p4));
                sv = __dadd_rn(sv, __dmul_rn(c, K.p3));
                sa = __dadd_rn(sa, __dmul_rn(c, K.p2));
            }
            if (Q >= 6) {
                sx = __dadd_rn(sx, __dmul_rn(e, K.p5));
                sv = __dadd_rn(sv, __dmul_rn(e, K.p4));
                sa = __dadd_rn(sa, __dmul_rn(e, K.p3));
            }
            x[d] = __dadd_rn(x[d], sx);
            v[q] = __dadd_rn(vd, sv);
            a[q] = __dadd_rn(ad, sa);
            if (Q == 5) *bp = __dadd_rn(b, __dmul_rn(c, K.p1));
            if (Q == 6) {
                *bp = __dadd_rn(b, __dadd_rn(__dmul_rn(c, K.p1), __dmul_rn(e, K.p2)));
                *cp = __dadd_rn(c, __dmul_rn(e, K.p1));
            }
        }
        p.x = x[0];
        p.y = x[1];
        if (D == 3) p.z = x[2];
        pos[s] = p;
    }
}
// Gear corrector (4A :1349-1357, 5A :1392-1401, 6A :1453-1464)
template <int D, int Q>
__global__ void __launch_bounds__(X_BLOCK)
k_x_gear_correct(double4 *__restrict__ pos, double *__restrict__ v, double *__restrict__ a, const double *__restrict__ f,
                 double *__restrict__ gear, const uint32_t *__restrict__ order, uint32_t n, uint32_t npad, size_t gstride, GearK K,
                 double dt, const int *__restrict__ abort_flag) {
    XGUARD;
    XLOOP {
        double4 p = pos[s];
        const uint32_t id = order[s];
        double x[3] = {p.x, p.y, p.z};
#pragma unroll
        for (int d = 0; d < D; d++) {
            const size_t q = (size_t)d * npad + s;
            double *bp = gear + (size_t)d * gstride + id, *cp = bp + 3 * gstride, *ep = bp + 6 * gstride;
            const double an = __ddiv_rn(f[q], p.w);
            const double corr = __dsub_rn(an, a[q]);
            x[d] = __dadd_rn(x[d], __dmul_rn(corr, K.c0));
            v[q] = __dadd_rn(v[q], __dmul_rn(corr, K.c1));
            a[q] = an;
            if (Q == 4) *bp = __dadd_rn(*bp, __ddiv_rn(corr, dt)); // bs[n] += correction / dt
            else *bp = __dadd_rn(*bp, __dmul_rn(corr, K.c3));
            if (Q >= 5) *cp = __dadd_rn(*cp, __dmul_rn(corr, K.c4));
            if (Q >= 6) *ep = __dadd_rn(*ep, __dmul_rn(corr, K.c5));
        }
        p.x = x[0];
        p.y = x[1];
        if (D == 3) p.z = x[2];
        pos[s] = p;
    }
}

// ---- thermostat reductions -----------------------------------------------------------------------------------
// what 0: q0 = sum over m != 0, finite of m * v.v   (2 * AtomGroup::kinetic_energy, box.cpp:401-411)
// what 1: q0 = sum f.v, q1 = sum (v.v) * m          (CollectionGaussianT::set_xi :1251-1260, every atom)
template <int D>
__global__ void __launch_bounds__(X_BLOCK)
k_x_reduce(int what, const double4 *__restrict__ pos, const double *__restrict__ v, const double *__restrict__ f, uint32_t n,
           uint32_t npad, double *partials) {
    double q0 = 0.0, q1 = 0.0;
    XLOOP {
        const double m = pos[s].w;
        const double vx = v[s], vy = v[npad + s], vz = D == 3 ? v[2 * (size_t)npad + s] : 0.0;
        const double vv = D == 3 ? __dadd_rn(__dmul_rn(vx, vx), __dadd_rn(__dmul_rn(vy, vy), __dmul_rn(vz, vz)))
                                 : __dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy));
        if (what == 0) {
            if (!(m == 0 || isinf(m))) q0 += __dmul_rn(m, vv);
        } else {
            const double fx = f[s], fy = f[npad + s], fz = D == 3 ? f[2 * (size_t)npad + s] : 0.0;
            q0 += D == 3 ? __dadd_rn(__dmul_rn(fx, vx), __dadd_rn(__dmul_rn(fy, vy), __dmul_rn(fz, vz)))
                         : __dadd_rn(__dmul_rn(fx, vx), __dmul_rn(fy, vy));
            q1 += __dmul_rn(vv, m);
        }
    }
    __shared__ double r0[X_BLOCK / 32], r1[X_BLOCK / 32];
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        q0 += __shfl_xor_sync(0xffffffffu, q0, o);
        q1 += __shfl_xor_sync(0xffffffffu, q1, o);
    }
    if ((threadIdx.x & 31) == 0) {
        r0[threadIdx.x >> 5] = q0;
        r1[threadIdx.x >> 5] = q1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t0 = 0, t1 = 0;
        for (int w = 0; w < X_BLOCK / 32; w++) {
            t0 += r0[w];
            t1 += r1[w];
        }
        partials[2 * blockIdx.x] = t0;
        partials[2 * blockIdx.x + 1] = t1;
    }
}

__device__ double x_fold(const double *partials, unsigned nblocks, int q) { // one block of X_BLOCK threads, deterministic
    __shared__ double red[X_BLOCK / 32];
    double t = 0.0;
    for (unsigned b = threadIdx.x; b < nblocks; b += blockDim.x) t += partials[2 * b + q];
#pragma unroll
    for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
    __syncthreads();
    double tot = 0.0;
    for (int w = 0; w < X_BLOCK / 32; w++) tot += red[w];
    return tot;
}

// solve_cubic, collection.cpp:2304-2353 (Numerical Recipes form; the root closest to `closeto` when three are real)
__device__ double x_solve_cubic(double a1, double a2, double a3, double closeto) {
    const double Q = (a1 * a1 - 3 * a2) / 9;
    const double Q3 = Q * Q * Q;
    const double R = ((2 * a1 * a1 * a1) - (9 * a1 * a2) + 27 * a3) / 54;
    const double R2 = R * R;
    if (Q3 >= R2) {
        const double theta = acos(R / sqrt(Q3));
        const double sqQ = -2 * sqrt(Q);
        const double x1 = sqQ * cos(theta / 3) - (a1 / 3);
        const double x2 = sqQ * cos((theta + (2 * M_PI)) / 3) - (a1 / 3);
        const double x3 = sqQ * cos((theta + (4 * M_PI)) / 3) - (a1 / 3);
        const double d1 = fabs(x1 - closeto), d2 = fabs(x2 - closeto), d3 = fabs(x3 - closeto);
        if (d1 < d2 && d1 < d3) return x1;
        if (d2 < d1 && d2 < d3) return x2;
        return x3;
    }
    const double R2Q3 = cbrt(sqrt(R2 - Q3) + fabs(R));
    const int sgn = (0.0 < R) - (R < 0.0);
    return -(sgn * (R2Q3 + (Q / R2Q3))) - (a1 / 3);
}

// NoseHoover, before the first half: Kt = 2 KE(t); lns += xi*dt + (Kt - ndof*T)*(dt*dt/2/Q)   (:1193-1194, :1207)
__global__ void k_x_nose_pre(const double *partials, unsigned nblocks, IntegScalars *sc, double ndof, double T, double dt, double Q,
                             const int *__restrict__ abort_flag) {
    XGUARD;
    const double Kt = x_fold(partials, nblocks, 0);
    if (threadIdx.x == 0) {
        sc->Kt = Kt;
        sc->lns += sc->xi * dt + (Kt - ndof * T) * (dt * dt / 2 / Q);
    }
}
// NoseHoover, after the second half: solve for xi(t+dt)   (:1217-1226)
__global__ void k_x_nose_post(const double *partials, unsigned nblocks, IntegScalars *sc, double ndof, double T, double dt, double Q,
                              const int *__restrict__ abort_flag) {
    XGUARD;
    const double Ky = x_fold(partials, nblocks, 0);
    if (threadIdx.x == 0) {
        const double xi = sc->xi, Kt = sc->Kt;
        const double z0 = xi + (Kt - 2 * ndof * T) * (dt / 2 / Q);
        const double z1 = Ky * 2 / dt / Q;
        const double nxi = x_solve_cubic(4 / dt - z0, 4 / dt / dt - 4 * z0 / dt, -(z0 * 4 / dt / dt) - z1, xi);
        sc->xi = nxi;
        sc->ytov = 1 + nxi * dt / 2;
    }
}
// GaussianT: z = sum f.v / sum m v.v; timestep (:1286-1290): xi = z / (1 - z*dt/2); set_forces (:1262-1266): xi = z
__global__ void k_x_gauss_post(const double *partials, unsigned nblocks, IntegScalars *sc, double dt, int in_timestep,
                               const int *__restrict__ abort_flag) {
    XGUARD;
    const double num = x_fold(partials, nblocks, 0);
    const double den = x_fold(partials, nblocks, 1);
    if (threadIdx.x == 0) {
        const double z = num / den;
        const double xi = in_timestep ? z / (1 - z * dt / 2) : z;
        sc->xi = xi;
        sc->ytov = 1 + xi * dt / 2;
    }
}

// NeighborList skin-drift reduction (trackers.cpp:23-53) at the end of the step, decision left in the step slot
__global__ void __launch_bounds__(X_BLOCK)
k_x_drift(const double4 *__restrict__ pos, const double *__restrict__ xlast, uint32_t n, uint32_t npad, double skin, double *d_top2,
          unsigned int *counter, NlistFlags *dflags, NlistFlags *hflags, int *d_slot, int *h_slot,
          const int *__restrict__ abort_flag) {
    XGUARD;
    double b1 = 0.0, b2 = 0.0;
    XLOOP top2_push(b1, b2, drift_dist(pos[s], xlast[s], xlast[npad + s], xlast[2 * (size_t)npad + s]));
    drift_finish(b1, b2, skin, d_top2, counter, dflags, hflags, d_slot, h_slot);
}

// ---- host ----------------------------------------------------------------------------------------------------
extern "C" int parm_integ_create(parm_ctx *c, int type, const double *params, int nparams, uint64_t seed, parm_integ **out) {
    if (!c || !out || (nparams > 0 && !params)) { parm_set_error("parm_integ_create: NULL argument"); return PARM_ERR_INVALID; }
    *out = 0;
    auto P = [&](int k, double dflt) { return k < nparams ? params[k] : dflt; };
    if (type == PARM_INTEG_VERLET) return parm_verlet_create(c, P(0, 0.0), out);
    if (type == PARM_INTEG_SOL) return parm_sol_create(c, P(0, 0.0), P(1, 0.0), P(2, 0.0), seed, out);
    if (type < PARM_INTEG_DAMPED || type > PARM_INTEG_GEAR6A) {
        parm_set_error("parm_integ_create: integrator type %d is not implemented (include/parm_b200.h PARM_INTEG_*)", type);
        return PARM_ERR_UNSUPPORTED;
    }
    if (c->sh.on) { parm_set_error("only CollectionVerlet and CollectionSol run on slab-decomposed contexts"); return PARM_ERR_UNSUPPORTED; }
    if (nparams < 1) { parm_set_error("parm_integ_create: dt missing"); return PARM_ERR_INVALID; }
    if (type == PARM_INTEG_DAMPED && !(P(0, 0.0) > 0)) { parm_set_error("Collection::CollectionSol: dt >= 0"); return PARM_ERR_INVALID; } // sic, :334-336
    CK(cudaSetDevice(c->device));
    parm_integ *g = new parm_integ();
    g->ctx = c;
    g->type = type;
    g->dt = P(0, 0.0);
    g->seed = seed;
    g->ncorrec = 1;
    switch (type) {
        case PARM_INTEG_DAMPED: g->damping = P(1, 0.0); break;
        case PARM_INTEG_SOLHT: g->damping = P(1, 0.0); g->desT = P(2, 0.0); break;
        case PARM_INTEG_OVERDAMPED: g->gamma = P(1, 1.0); break;
        case PARM_INTEG_NOSEHOOVER: g->Q = P(1, 1.0); g->desT = P(2, 0.0); break;
        case PARM_INTEG_GEAR4A: case PARM_INTEG_GEAR5A: case PARM_INTEG_GEAR6A: g->ncorrec = (int)P(1, 1.0); break;
        default: break;
    }
    if (type == PARM_INTEG_NOSEHOOVER || type == PARM_INTEG_GAUSSIANT) {
        CK(cudaMalloc(&g->d_scal, sizeof(IntegScalars)));
        IntegScalars z = {0.0, 0.0, 0.0, 1.0};
        CK(cudaMemcpy(g->d_scal, &z, sizeof(z), cudaMemcpyHostToDevice));
        CK(cudaMalloc(&g->d_xpart, 2 * X_MAXBLOCKS * sizeof(double)));
    }
    if (type >= PARM_INTEG_GEAR4A) { // resetbs(): zeros, collection.hpp:645
        const size_t gs = std::max(c->npad, c->nid_pad);
        CK(cudaMalloc(&g->d_gear, 9 * gs * sizeof(double)));
        CK(cudaMemset(g->d_gear, 0, 9 * gs * sizeof(double)));
    }
    *out = g;
    return 0;
}

extern "C" int parm_integ_get_scalars(parm_integ *g, double *out2) {
    if (!g || !out2) { parm_set_error("parm_integ_get_scalars: NULL argument"); return PARM_ERR_INVALID; }
    out2[0] = out2[1] = 0.0;
    if (!g->d_scal) return 0;
    CK(cudaSetDevice(g->ctx->device));
    CK(cudaStreamSynchronize(g->ctx->stream));
    IntegScalars z;
    CK(cudaMemcpy(&z, g->d_scal, sizeof(z), cudaMemcpyDeviceToHost));
    out2[0] = z.xi;
    out2[1] = z.lns;
    return 0;
}
extern "C" int parm_integ_reset_bath(parm_integ *g) {
    if (!g || !g->d_scal) { parm_set_error("reset_bath: not a thermostatted collection"); return PARM_ERR_INVALID; }
    CK(cudaSetDevice(g->ctx->device));
    CK(cudaStreamSynchronize(g->ctx->stream));
    IntegScalars z = {0.0, 0.0, 0.0, 1.0};
    CK(cudaMemcpy(g->d_scal, &z, sizeof(z), cudaMemcpyHostToDevice));
    return 0;
}
extern "C" int parm_integ_set_param(parm_integ *g, int which, double value) {
    if (!g) { parm_set_error("parm_integ_set_param: NULL integrator"); return PARM_ERR_INVALID; }
    switch (which) {
        case 1: g->Q = value; return 0;
        case 2: g->desT = value; return 0;
        case 3: g->damping = value; return 0;
        case 4: g->gamma = value; return 0;
    }
    parm_set_error("parm_integ_set_param: unknown parameter %d", which);
    return PARM_ERR_INVALID;
}

static int thermo_reduce(parm_integ *g, int what, unsigned *nblocks) {
    parm_ctx *c = g->ctx;
    const uint32_t n = parm_owned(c);
    const unsigned nb = xgrid(c, n);
    if (c->D == 3) k_x_reduce<3><<<nb, X_BLOCK, 0, c->stream>>>(what, c->pos, c->v, c->f, n, c->npad, g->d_xpart);
    else k_x_reduce<2><<<nb, X_BLOCK, 0, c->stream>>>(what, c->pos, c->v, c->f, n, c->npad, g->d_xpart);
    CK_LAUNCH(c);
    *nblocks = nb;
    return 0;
}

int parm_integ_extra_after_set_forces(parm_integ *g) { // CollectionGaussianT::set_forces(.., true) -> set_xi()
    parm_ctx *c = g->ctx;
    if (g->type != PARM_INTEG_GAUSSIANT || parm_owned(c) == 0) return 0;
    unsigned nb;
    PTRY(thermo_reduce(g, 1, &nb));
    k_x_gauss_post<<<1, X_BLOCK, 0, c->stream>>>(g->d_xpart, nb, g->d_scal, g->dt, 0, nullptr);
    CK_LAUNCH(c);
    return 0;
}

#define XD(kern, ...)                                                              \
    do {                                                                           \
        if (c->D == 3) kern<3><<<grid, X_BLOCK, 0, c->stream>>>(__VA_ARGS__);      \
        else kern<2><<<grid, X_BLOCK, 0, c->stream>>>(__VA_ARGS__);                \
        CK_LAUNCH(c);                                                              \
    } while (0)
#define XDQ(kern, QQ, ...)                                                         \
    do {                                                                           \
        if (c->D == 3) kern<3, QQ><<<grid, X_BLOCK, 0, c->stream>>>(__VA_ARGS__);  \
        else kern<2, QQ><<<grid, X_BLOCK, 0, c->stream>>>(__VA_ARGS__);            \
        CK_LAUNCH(c);                                                              \
    } while (0)

int parm_integ_extra_enqueue(parm_integ *g, uint64_t step, const int *abort_flag, int slot) {
    parm_ctx *c = g->ctx;
    const uint32_t n = parm_owned(c);
    parm_nlist *nl = g->trackers.empty() ? nullptr : g->trackers[0];
    const unsigned grid = xgrid(c, n);
    const double dt = g->dt;
    const int type = g->type;
    PTRY(parm_prof_begin(c, PARM_PROF_INTEG1));
    if (type == PARM_INTEG_DAMPED) {
        double c0 = 1, c1 = 1, c2 = .5; // set_constants, :342-354
        if (g->damping > 0.0) {
            const double dampdt = g->damping * dt;
            c0 = exp(-dampdt);
            c1 = (-expm1(-dampdt)) / dampdt;
            c2 = (1 - c1) / dampdt;
        }
        XD(k_x_lin1, c->pos, c->v, c->a, n, c->npad, c1 * dt, c2 * dt * dt, c0, dt * (c1 - c2), 1, abort_flag);
        PTRY(parm_prof_end(c));
        PTRY(parm_prof_begin(c, PARM_PROF_FORCE));
        PTRY(parm_integ_launch_all_forces(g, abort_flag));
        PTRY(parm_prof_end(c));
        PTRY(parm_prof_begin(c, PARM_PROF_INTEG2));
        XD(k_x_accel_kick, c->pos, c->v, c->a, c->f, n, c->npad, dt * c2, 1, abort_flag);
    } else if (type == PARM_INTEG_SOLHT) {
        const double keepv = 1 - (g->damping * dt);
        const double xpartfromv = dt - (dt * dt * g->damping / 2);
        XD(k_x_lin1, c->pos, c->v, c->a, n, c->npad, xpartfromv, .5 * dt * dt, keepv, dt / 2, 1, abort_flag);
        PTRY(parm_prof_end(c));
        PTRY(parm_prof_begin(c, PARM_PROF_FORCE));
        PTRY(parm_integ_launch_all_forces(g, abort_flag));
        PTRY(parm_prof_end(c));
        PTRY(parm_prof_begin(c, PARM_PROF_INTEG2));
        const double *noise = nullptr;
        if (g->d_noise) {
            const size_t per = (size_t)g->n_mobile * c->D;
            const size_t off = (size_t)(step - g->noise_step0) * per;
            if (off + per > g->noise_len) { parm_set_error("CollectionSolHT: injected noise exhausted"); return PARM_ERR_INVALID; }
            noise = g->d_noise + off;
        }
        const double sigma = sqrt(2.0 * g->desT * g->damping / dt); // GaussVec sigma, :393, :398
        XD(k_x_solht2, c->pos, c->v, c->a, c->f, c->order, n, c->npad, dt / 2, sigma, noise, g->d_mobile_rank, step, g->seed, abort_flag);
    } else if (type == PARM_INTEG_OVERDAMPED) {
        PTRY(parm_prof_end(c));
        PTRY(parm_prof_begin(c, PARM_PROF_FORCE));
        PTRY(parm_integ_launch_all_forces(g, abort_flag));
        PTRY(parm_prof_end(c));
        PTRY(parm_prof_begin(c, PARM_PROF_INTEG2));
        XD(k_x_overdamped, c->pos, c->v, c->a, c->f, n, c->npad, dt, g->gamma, abort_flag);
    } else if (type == PARM_INTEG_NOSEHOOVER || type == PARM_INTEG_GAUSSIANT) {
        const bool nose = type == PARM_INTEG_NOSEHOOVER;
        unsigned nb = 0;
        double ndof = 0;
        if (nose) {
            ndof = g->ndof_cached; // degrees_of_freedom() (:116-133), read by parm_integ_timestep once per call
            PTRY(thermo_reduce(g, 0, &nb));
            k_x_nose_pre<<<1, X_BLOCK, 0, c->stream>>>(g->d_xpart, nb, g->d_scal, ndof, g->desT, dt, g->Q, abort_flag);
            CK_LAUNCH(c);
            XDQ(k_x_thermo1, true, c->pos, c->v, c->a, n, c->npad, dt, dt * dt / 2, dt / 2, g->d_scal, abort_flag);
        } else {
            XDQ(k_x_thermo1, false, c->pos, c->v, c->a, n, c->npad, dt, dt * dt / 2, dt / 2, g->d_scal, abort_flag);
        }
        PTRY(parm_prof_end(c));
        PTRY(parm_prof_begin(c, PARM_PROF_FORCE));
        PTRY(parm_integ_launch_all_forces(g, abort_flag));
        PTRY(parm_prof_end(c));
        PTRY(parm_prof_begin(c, PARM_PROF_INTEG2));
        XD(k_x_accel_kick, c->pos, c->v, c->a, c->f, n, c->npad, dt / 2, 0, abort_flag);
        PTRY(thermo_reduce(g, nose ? 0 : 1, &nb));
        if (nose) k_x_nose_post<<<1, X_BLOCK, 0, c->stream>>>(g->d_xpart, nb, g->d_scal, ndof, g->desT, dt, g->Q, abort_flag);
        else k_x_gauss_post<<<1, X_BLOCK, 0, c->stream>>>(g->d_xpart, nb, g->d_scal, dt, 1, abort_flag);
        CK_LAUNCH(c);
        XD(k_x_scale_v, c->v, n, c->npad, g->d_scal, abort_flag);
    } else if (type == PARM_INTEG_GEAR3A) {
        XD(k_x_lin1, c->pos, c->v, c->a, n, c->npad, dt, dt * dt / 2, 1.0, dt, 0, abort_flag);
        PTRY(parm_prof_end(c));
        PTRY(parm_prof_begin(c, PARM_PROF_FORCE));
        PTRY(parm_integ_launch_all_forces(g, abort_flag));
        PTRY(parm_prof_end(c));
        PTRY(parm_prof_begin(c, PARM_PROF_INTEG2));
        XD(k_x_gear3_correct, c->pos, c->v, c->a, c->f, n, c->npad, dt / 2, abort_flag);
    } else {
        const size_t gs = std::max(c->npad, c->nid_pad);
        GearK K;
        memset(&K, 0, sizeof(K));
        K.p1 = dt;
        if (type == PARM_INTEG_GEAR4A) {
            K.p2 = dt * dt / 2;
            K.p3 = dt * dt * dt / 6;
            K.c0 = dt * dt / 12;
            K.c1 = 5 * dt / 12;
        } else { // 5A :1366-1370, :1380-1383; 6A :1409-1415, :1434-1437
            K.p2 = dt * dt / 2;
            K.p3 = dt * K.p2 / 3;
            K.p4 = dt * K.p3 / 4;
            K.p5 = dt * K.p4 / 5;
            if (type == PARM_INTEG_GEAR5A) {
                K.c0 = 19 * dt * dt / 240; K.c1 = 3 * dt / 8; K.c3 = 3 / (2 * dt); K.c4 = 1 / (dt * dt);
            } else {
                K.c0 = 3 * dt * dt / 40; K.c1 = 251 * dt / 720; K.c3 = 11 / (6 * dt); K.c4 = 2 / (dt * dt); K.c5 = 1 / (dt * dt * dt);
            }
        }
#define PREDICT(QQ) XDQ(k_x_gear_predict, QQ, c->pos, c->v, c->a, g->d_gear, c->order, n, c->npad, gs, K, abort_flag)
#define CORRECT(QQ) XDQ(k_x_gear_correct, QQ, c->pos, c->v, c->a, c->f, g->d_gear, c->order, n, c->npad, gs, K, dt, abort_flag)
        if (type == PARM_INTEG_GEAR4A) PREDICT(4); else if (type == PARM_INTEG_GEAR5A) PREDICT(5); else PREDICT(6);
        PTRY(parm_prof_end(c));
        PTRY(parm_prof_begin(c, PARM_PROF_FORCE));
        for (int m = 0; m < g->ncorrec; m++) {
            PTRY(parm_integ_launch_all_forces(g, abort_flag));
            if (type == PARM_INTEG_GEAR4A) CORRECT(4); else if (type == PARM_INTEG_GEAR5A) CORRECT(5); else CORRECT(6);
        }
        PTRY(parm_prof_end(c));
        PTRY(parm_prof_begin(c, PARM_PROF_INTEG2));
#undef PREDICT
#undef CORRECT
    }
    if (nl) { // update_trackers() ends every step
        const unsigned g1 = std::min(grid, 4096u);
        k_x_drift<<<g1, X_BLOCK, 0, c->stream>>>(c->pos, nl->xlast, n, c->npad, nl->skin, nl->d_top2, nl->d_counter, nl->d_flags,
                                                 nl->h_flags, nl->d_slot + slot, nl->h_slot + slot, abort_flag);
        CK_LAUNCH(c);
    }
    PTRY(parm_prof_end(c));
    return 0;
}
