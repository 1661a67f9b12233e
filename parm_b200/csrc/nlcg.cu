// CollectionNLCG on the device (collection.hpp:400-474, collection.cpp:494-854): nonlinear conjugate-gradient
// minimisation of H = U + P0 V over the atom coordinates and kappa ln V, with secant line search -- the packer's
// minimiser (pyparm/packmin.py). The scalar logic of timestep() (:682-835) runs on the host exactly as the
// reference writes it; everything O(N) is a kernel: stepx (:602-618), the four dot products (:630-680) in one
// pass, the a/v assignments, and Collection::set_forces_get_pressure (:181-208) through the pair-force kernel
// with its virial reduction. Each secant iteration needs the dot products on the host, so the step is
// latency-bound by one device->host scalar read per iteration, not by bandwidth.
#include <math.h>
#include <string.h>

#include <algorithm>

#include "internal.cuh"

#define N_BLOCK 256
#define N_MAXBLOCKS 1024

static inline unsigned ngrid(const parm_ctx *c, uint32_t n) {
    unsigned nb = (n + N_BLOCK - 1) / N_BLOCK;
    unsigned cap = std::min((unsigned)c->num_sms * 8, (unsigned)N_MAXBLOCKS);
    return nb < 1 ? 1 : std::min(nb, cap);
}
__device__ __forceinline__ bool n_frozen(double m) { return m <= 0 || isinf(m); }
#define NLOOP for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x)

// stepx :606-612: x *= Lfac; x += v*dx   (every atom)
template <int D>
__global__ void __launch_bounds__(N_BLOCK)
k_n_stepx(double4 *__restrict__ pos, const double *__restrict__ v, uint32_t n, uint32_t npad, double Lfac, double dx) {
    NLOOP {
        double4 p = pos[s];
        p.x = __dadd_rn(__dmul_rn(p.x, Lfac), __dmul_rn(v[s], dx));
        p.y = __dadd_rn(__dmul_rn(p.y, Lfac), __dmul_rn(v[npad + s], dx));
        if (D == 3) p.z = __dadd_rn(__dmul_rn(p.z, Lfac), __dmul_rn(v[2 * (size_t)npad + s], dx));
        pos[s] = p;
    }
}

// mode 0  set_forces(constraints_and_a) :556-566   frozen: v = a = 0; else a = f, v = f
// mode 1  timestep :800-808                        frozen: a = 0;     else a = f
// mode 2  timestep :822-830                        frozen: v = 0;     else v = a + v*beta
// mode 3  descend :840-847                         frozen: -;         else v = f, a = f
// mode 4  reset :528-530                           every atom: v = a
// mode 5  set_forces_get_pressure(false) :195-200  frozen: a = 0
template <int D>
__global__ void __launch_bounds__(N_BLOCK)
k_n_assign(int mode, const double4 *__restrict__ pos, double *__restrict__ v, double *__restrict__ a, const double *__restrict__ f,
           uint32_t n, uint32_t npad, double beta) {
    NLOOP {
        const bool fr = n_frozen(pos[s].w);
#pragma unroll
        for (int d = 0; d < D; d++) {
            const size_t q = (size_t)d * npad + s;
            if (mode == 0) {
                const double t = fr ? 0.0 : f[q];
                a[q] = t;
                v[q] = t;
            } else if (mode == 1) {
                a[q] = fr ? 0.0 : f[q];
            } else if (mode == 2) {
                v[q] = fr ? 0.0 : __dadd_rn(a[q], __dmul_rn(v[q], beta));
            } else if (mode == 3) {
                if (!fr) {
                    v[q] = f[q];
                    a[q] = f[q];
                }
            } else if (mode == 4) {
                v[q] = a[q];
            } else {
                if (fr) a[q] = 0.0;
            }
        }
    }
}

// per block: sum over mobile atoms of f.f, f.a, f.v, v.v, and |v + x*Lfac|^2 (kinetic_energy :574-588)
template <int D>
__global__ void __launch_bounds__(N_BLOCK)
k_n_dots(const double4 *__restrict__ pos, const double *__restrict__ v, const double *__restrict__ a, const double *__restrict__ f,
         uint32_t n, uint32_t npad, double Lfac, double *partials) {
    double q[5] = {0, 0, 0, 0, 0};
    NLOOP {
        const double4 p = pos[s];
        if (n_frozen(p.w)) continue;
        const double x[3] = {p.x, p.y, p.z};
        double ff[3], fa[3], fv[3], vv[3], kk[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            if (d < D) {
                const size_t i = (size_t)d * npad + s;
                const double fd = f[i], ad = a[i], vd = v[i];
                ff[d] = __dmul_rn(fd, fd);
                fa[d] = __dmul_rn(fd, ad);
                fv[d] = __dmul_rn(fd, vd);
                vv[d] = __dmul_rn(vd, vd);
                const double w = __dadd_rn(vd, __dmul_rn(x[d], Lfac));
                kk[d] = __dmul_rn(w, w);
            } else {
                ff[d] = fa[d] = fv[d] = vv[d] = kk[d] = 0.0;
            }
        }
        // dot / squaredNorm as e0 + (e1 + e2), like every reduction of the path
        q[0] += D == 3 ? __dadd_rn(ff[0], __dadd_rn(ff[1], ff[2])) : __dadd_rn(ff[0], ff[1]);
        q[1] += D == 3 ? __dadd_rn(fa[0], __dadd_rn(fa[1], fa[2])) : __dadd_rn(fa[0], fa[1]);
        q[2] += D == 3 ? __dadd_rn(fv[0], __dadd_rn(fv[1], fv[2])) : __dadd_rn(fv[0], fv[1]);
        q[3] += D == 3 ? __dadd_rn(vv[0], __dadd_rn(vv[1], vv[2])) : __dadd_rn(vv[0], vv[1]);
        q[4] += D == 3 ? __dadd_rn(kk[0], __dadd_rn(kk[1], kk[2])) : __dadd_rn(kk[0], kk[1]);
    }
    __shared__ double red[5][N_BLOCK / 32];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        double t = q[k];
#pragma unroll
        for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = t;
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        double t = 0;
        for (int w = 0; w < N_BLOCK / 32; w++) t += red[threadIdx.x][w];
        partials[5 * blockIdx.x + threadIdx.x] = t;
    }
}
__global__ void k_n_fold(const double *partials, unsigned nblocks, double *out) {
    __shared__ double red[N_BLOCK / 32];
    for (int k = 0; k < 5; k++) {
        double t = 0;
        for (unsigned b = threadIdx.x; b < nblocks; b += blockDim.x) t += partials[5 * b + k];
#pragma unroll
        for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0;
            for (int w = 0; w < N_BLOCK / 32; w++) tot += red[w];
            out[k] = tot;
        }
        __syncthreads();
    }
}

// ---- host ------------------------------------------------------------------------------------------------------
static double box_V(const parm_ctx *c) { // OriginBox::V, box.hpp:122,127
    return c->D == 3 ? c->box.L[0] * c->box.L[1] * c->box.L[2] : c->box.L[0] * c->box.L[1];
}
static double length_squared(const parm_ctx *c) { // get_length_squared :620-628
    return c->D == 3 ? pow(box_V(c), 2.0 / 3.0) : box_V(c);
}
static void box_resize(parm_ctx *c, double factor) { // OriginBox::resize(factor), box.cpp:3-6: boxsize *= factor
    for (int k = 0; k < c->D; k++) {
        const double l = c->box.L[k] * factor;
        c->box.L[k] = l;
        c->box.invL[k] = 1.0 / l;
        c->box.halfL[k] = l * 0.5;
    }
    for (parm_nlist *nl : c->nlists) parm_tile_invalidate(nl);
}

#define ND(kern, ...)                                                                    \
    do {                                                                                 \
        if (c->D == 3) kern<3><<<ngrid(c, n), N_BLOCK, 0, c->stream>>>(__VA_ARGS__);     \
        else kern<2><<<ngrid(c, n), N_BLOCK, 0, c->stream>>>(__VA_ARGS__);               \
        CK_LAUNCH(c);                                                                    \
    } while (0)

static int assign(parm_integ *g, int mode, double beta = 0.0) {
    parm_ctx *c = g->ctx;
    const uint32_t n = parm_owned(c);
    if (!n) return 0;
    ND(k_n_assign, mode, c->pos, c->v, c->a, c->f, n, c->npad, beta);
    return 0;
}

// out[5]: sums of f.f, f.a, f.v, v.v, |v + x Lfac|^2 over the mobile atoms
static int dots(parm_integ *g, double Lfac, double *out) {
    parm_ctx *c = g->ctx;
    const uint32_t n = parm_owned(c);
    for (int k = 0; k < 5; k++) out[k] = 0.0;
    if (!n) return 0;
    PTRY(parm_ctx_ensure_red(c, 5 * (N_MAXBLOCKS + 2)));
    const unsigned nb = ngrid(c, n);
    if (c->D == 3) k_n_dots<3><<<nb, N_BLOCK, 0, c->stream>>>(c->pos, c->v, c->a, c->f, n, c->npad, Lfac, c->d_red + 8);
    else k_n_dots<2><<<nb, N_BLOCK, 0, c->stream>>>(c->pos, c->v, c->a, c->f, n, c->npad, Lfac, c->d_red + 8);
    CK_LAUNCH(c);
    k_n_fold<<<1, N_BLOCK, 0, c->stream>>>(c->d_red + 8, nb, c->d_red);
    CK_LAUNCH(c);
    CK(cudaMemcpyAsync(c->h_red, c->d_red, 5 * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 5; k++) out[k] = c->h_red[k];
    return 0;
}
static int fdotv(parm_integ *g, double *r) { // :656-667
    double d[5];
    PTRY(dots(g, 0.0, d));
    *r = d[2] / length_squared(g->ctx) + g->nlcg->fl * g->nlcg->vl;
    return 0;
}
static int fdota(parm_integ *g, double *r) { // :643-654
    double d[5];
    PTRY(dots(g, 0.0, d));
    *r = d[1] / length_squared(g->ctx) + g->nlcg->fl * g->nlcg->al;
    return 0;
}
static int vdotv(parm_integ *g, double *r) { // :669-680
    double d[5];
    PTRY(dots(g, 0.0, d));
    *r = d[3] / length_squared(g->ctx) + g->nlcg->vl * g->nlcg->vl;
    return 0;
}

static int stepx(parm_integ *g, double dx) { // :602-618
    parm_ctx *c = g->ctx;
    NlcgState *S = g->nlcg;
    const uint32_t n = parm_owned(c);
    const double Lfac = exp(dx * S->vl / (S->kappa * c->D));
    if (n) ND(k_n_stepx, c->pos, c->v, n, c->npad, Lfac, dx);
    box_resize(c, Lfac);
    return 0;
}

// Collection::set_forces_get_pressure(false) :181-208
static int forces_get_pressure(parm_integ *g, double *p_out) {
    parm_ctx *c = g->ctx;
    PTRY(parm_reset_forces(c));
    double p = 0;
    for (parm_inter *it : g->inters) {
        double w = 0;
        PTRY(parm_inter_set_forces(it, PARM_WANT_VIRIAL, &w));
        p += w;
    }
    PTRY(assign(g, 5));
    *p_out = p;
    return 0;
}

extern "C" int parm_nlcg_set_forces(parm_integ *g, int constraints_and_a, int setV) { // :534-568
    if (!g || !g->nlcg) { parm_set_error("parm_nlcg_set_forces: not a CollectionNLCG"); return PARM_ERR_INVALID; }
    parm_ctx *c = g->ctx;
    NlcgState *S = g->nlcg;
    CK(cudaSetDevice(c->device));
    const double V = box_V(c);
    if (setV) {
        double interacP;
        PTRY(forces_get_pressure(g, &interacP));
        S->fl = ((interacP / c->D) - (S->P0 * V)) / S->kappa;
        if (constraints_and_a) {
            S->al = S->fl;
            S->vl = S->fl;
        }
    } else {
        PTRY(parm_integ_launch_all_forces(g, nullptr)); // Collection::set_forces(false)
    }
    if (constraints_and_a) {
        PTRY(assign(g, 0));
        PTRY(fdota(g, &S->Knew));
    }
    return 0;
}

extern "C" int parm_nlcg_reset(parm_integ *g) { // :525-532
    if (!g || !g->nlcg) { parm_set_error("parm_nlcg_reset: not a CollectionNLCG"); return PARM_ERR_INVALID; }
    g->nlcg->k = 0;
    PTRY(parm_nlcg_set_forces(g, 1, 1));
    PTRY(assign(g, 4));
    g->nlcg->vl = g->nlcg->al;
    return 0;
}

extern "C" int parm_nlcg_descend(parm_integ *g) { // :837-854
    if (!g || !g->nlcg) { parm_set_error("parm_nlcg_descend: not a CollectionNLCG"); return PARM_ERR_INVALID; }
    NlcgState *S = g->nlcg;
    PTRY(parm_nlcg_set_forces(g, 0, 1));
    PTRY(assign(g, 3));
    S->al = S->fl;
    S->vl = S->fl;
    PTRY(stepx(g, g->dt));
    return parm_integ_update_trackers(g);
}

int parm_nlcg_timestep(parm_integ *g) { // :682-835
    NlcgState *S = g->nlcg;
    const double dt = g->dt;
    const int NDIM = g->ctx->D;
    PTRY(stepx(g, dt));
    PTRY(parm_nlcg_set_forces(g, 0, 1)); // sets both Atom.f and fl, but not Atom.a or al
    double t;
    PTRY(fdotv(g, &t));
    double eta0 = -t; // slope at x0 - dt
    double eta;
    PTRY(stepx(g, -dt));
    PTRY(parm_integ_update_trackers(g));
    S->alpha = -dt;
    PTRY(parm_nlcg_set_forces(g, 0, 1));
    S->dxsum = 0;
    double vdv;
    PTRY(vdotv(g, &vdv));

    // Secant stepping: v does not change here; step along v until the force along v gets very small
    for (S->sec = 0; S->sec < S->secmax; S->sec++) {
        PTRY(fdotv(g, &t));
        eta = -t; // slope at x; eta0 is the slope at x - alpha
        double alphafac = -eta / fabs(eta0 - eta);
        if (fabs(eta0 - eta) <= 1e-12 * fabs(eta)) {
            alphafac = S->alphamax > 0 ? S->alphamax : 1.1;
            S->sec = S->sec > 0 ? S->sec * 2 - 1 : 1;
        }
        if (S->alphamax > 0 && alphafac > S->alphamax) alphafac = S->alphamax;
        if (S->alphamax > 0 && alphafac < -S->alphamax) alphafac = -S->alphamax;
        S->alpha = fabs(S->alpha) * alphafac;

        double newdxsum = fabs(S->dxsum + S->alpha);
        if (S->dxmax > 0 && newdxsum > S->dxmax) {
            S->k = 0;
            break;
        }
        double dVoverV = expm1(fabs(S->dxsum + S->alpha) * S->vl / (S->kappa * NDIM));
        if (S->maxdV > 0 && dVoverV > S->maxdV) {
            S->k = 0;
            break;
        }
        S->dxsum += S->alpha;
        PTRY(stepx(g, S->alpha));
        PTRY(parm_nlcg_set_forces(g, 0, 1));
        eta0 = eta;
        if (S->alpha * S->alpha * vdv < S->seceps * S->seceps) break;
        if ((S->sec > 1) && (S->afrac > 0) && (fabs(S->alpha) < fabs(S->dxsum)) && (fabs(S->alpha) / fabs(S->dxsum) < S->afrac)) break;
        if (S->stepmax > 0 && S->dxsum * S->dxsum * vdv > S->stepmax * S->stepmax) {
            S->k = 0;
            break;
        }
    }

    S->alphavmax = sqrt(S->alpha * S->alpha * vdv);

    double Kold = S->Knew;
    double Kmid;
    PTRY(fdota(g, &Kmid));
    PTRY(assign(g, 1));
    S->al = S->fl;
    PTRY(fdota(g, &S->Knew));
    S->beta = (S->Knew - Kmid) / Kold;
    S->betaused = S->beta;
    S->k++;
    if (S->k >= S->kmax || isinf(S->betaused) || isnan(S->betaused) || S->betaused <= 0) {
        S->k = 0;
        S->betaused = 0;
    } else if (S->betaused > 1) {
        S->betaused = 1;
    }
    PTRY(assign(g, 2, S->betaused));
    S->vl = S->al + S->betaused * S->vl;
    return 0;
}

extern "C" int parm_nlcg_create(parm_ctx *c, double dt, double P0, double kappa, double kmax, unsigned secmax, double seceps,
                                parm_integ **out) {
    if (!c || !out) { parm_set_error("parm_nlcg_create: NULL argument"); return PARM_ERR_INVALID; }
    *out = 0;
    if (c->sh.on) { parm_set_error("CollectionNLCG resizes the box: single-GPU contexts only"); return PARM_ERR_UNSUPPORTED; }
    parm_integ *g = new parm_integ();
    g->ctx = c;
    g->type = PARM_INTEG_NLCG;
    g->dt = dt;
    NlcgState *S = new NlcgState();
    memset(S, 0, sizeof(*S));
    S->seceps = seceps; // ctor :494-523
    S->secmax = secmax;
    S->kappa = kappa;
    S->alphamax = 2.0;
    S->afrac = 0;
    S->dxmax = 100;
    S->stepmax = 1e-3;
    S->kmax = kmax;
    S->P0 = P0;
    g->nlcg = S;
    *out = g;
    return 0;
}
void parm_nlcg_free(parm_integ *g) {
    delete g->nlcg;
    g->nlcg = 0;
}

extern "C" int parm_nlcg_set(parm_integ *g, int which, double value) {
    if (!g || !g->nlcg) { parm_set_error("parm_nlcg_set: not a CollectionNLCG"); return PARM_ERR_INVALID; }
    NlcgState *S = g->nlcg;
    switch (which) {
        case PARM_NLCG_DT: g->dt = value; return parm_nlcg_reset(g);       // set_dt :455-458
        case PARM_NLCG_P0: S->P0 = value; return parm_nlcg_reset(g);       // set_pressure_goal :459-462
        case PARM_NLCG_KAPPA: S->kappa = value; return parm_nlcg_reset(g); // set_kappa :464-467
        case PARM_NLCG_ALPHAMAX: S->alphamax = value; return 0;
        case PARM_NLCG_AFRAC: S->afrac = value; return 0;
        case PARM_NLCG_DXMAX: S->dxmax = value; return 0;
        case PARM_NLCG_STEPMAX: S->stepmax = value; return 0;
        case PARM_NLCG_MAXDV: S->maxdV = value; return 0;
        case PARM_NLCG_KMAX: S->kmax = value; return 0;
        case PARM_NLCG_SECMAX: S->secmax = (unsigned)value; return 0;
        case PARM_NLCG_SECEPS: S->seceps = value; return 0;
    }
    parm_set_error("parm_nlcg_set: unknown parameter %d", which);
    return PARM_ERR_INVALID;
}

extern "C" int parm_nlcg_get(parm_integ *g, double *o) {
    if (!g || !g->nlcg || !o) { parm_set_error("parm_nlcg_get: not a CollectionNLCG"); return PARM_ERR_INVALID; }
    const NlcgState *S = g->nlcg;
    o[0] = g->dt; o[1] = S->P0; o[2] = S->kappa; o[3] = S->Knew; o[4] = S->k; o[5] = S->vl; o[6] = S->fl; o[7] = S->al;
    o[8] = S->alpha; o[9] = S->beta; o[10] = S->betaused; o[11] = S->dxsum; o[12] = S->alphavmax; o[13] = S->sec;
    o[14] = S->kmax; o[15] = S->secmax;
    return 0;
}

extern "C" int parm_nlcg_reduce(parm_integ *g, int what, double *out) {
    if (!g || !g->nlcg || !out) { parm_set_error("parm_nlcg_reduce: not a CollectionNLCG"); return PARM_ERR_INVALID; }
    parm_ctx *c = g->ctx;
    NlcgState *S = g->nlcg;
    CK(cudaSetDevice(c->device));
    double d[5];
    switch (what) {
        case PARM_NLCG_FDOTF: // :630-641
            PTRY(dots(g, 0.0, d));
            *out = d[0] / length_squared(c) + S->fl * S->fl;
            return 0;
        case PARM_NLCG_FDOTA: return fdota(g, out);
        case PARM_NLCG_FDOTV: return fdotv(g, out);
        case PARM_NLCG_VDOTV: return vdotv(g, out);
        case PARM_NLCG_KINETIC: // :574-588
            PTRY(dots(g, exp(S->vl / (S->kappa * c->D)), d));
            *out = d[4] / 2.0;
            return 0;
        case PARM_NLCG_PRESSURE: { // :590-600
            double w;
            PTRY(parm_integ_virial(g, &w));
            *out = w / box_V(c) / (double)c->D;
            return 0;
        }
        case PARM_NLCG_HAMILTONIAN: { // :570-572
            double e;
            PTRY(parm_integ_potential_energy(g, &e));
            *out = e + S->P0 * box_V(c);
            return 0;
        }
    }
    parm_set_error("parm_nlcg_reduce: unknown quantity %d", what);
    return PARM_ERR_INVALID;
}
