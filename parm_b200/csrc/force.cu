// NListed<A,P> on the device: the pair force / energy / virial / stress loop
// (interaction.hpp:2102-2291) over the four pair functors in scope:
//   LJRepulsePair        :875-891 + LJRepulsive :119-152
//   RepulsionPair        :1528-1550
//   LJAttractRepulsePair :1251-1299
//   LennardJonesCutPair  :967-987 + LennardJonesCut :238-281
// One thread per atom walks its FULL neighbour row (no Newton's-third-law scatter, hence no
// atomics and a run-to-run deterministic sum); energy/virial/stress are reduced with warp
// shuffles, then per block, then by one folding block (deterministic), and halved because
// every pair is visited from both ends.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <tuple>

#include "internal.cuh"

#define F_BLOCK 128
#define NPART 13 // E, virial, stress[9], contacts, overlaps

enum { MODE_F = 0, MODE_FALL = 1, MODE_OBS = 2 };

template <int KIND>
__device__ __forceinline__ void pair_eval(const PairConst &P, double dsq, bool want_e, double &scal, double &e) {
    scal = 0.0;
    e = 0.0;
    if (KIND == PARM_PAIR_REPULSION) {
        // RepulsionPair::forces / energy, interaction.hpp:1537-1550
        if (dsq > P.sig2) return;
        double R = sqrt(dsq);
        double t = 1.0 - R / P.sig;
        double pm1, p0; // pow(t, n-1), pow(t, n)
        if (P.expo == 2.0) {
            pm1 = t;
            p0 = t * t;
        } else if (P.expo == 2.5) {
            double sq = sqrt(t);
            pm1 = t * sq;
            p0 = t * t * sq;
        } else {
            pm1 = pow(t, P.expo - 1.0);
            p0 = pm1 * t;
        }
        scal = P.eps * pm1 / P.sig / R;
        if (want_e) e = P.eps * p0 / P.expo;
    } else {
        // rsq = dsq/(sig*sig); if (rsq > cut*cut) -> 0; rsix = sigma^6/r^6
        // f = rij * (12 eps rsix (rsix - 1) / dsq)      (:135-151, :259-267, :1289-1298)
        if (dsq > P.rc2) return;
        double w = 1.0 / dsq;
        double s2 = P.sig2 * w;
        double ir6 = s2 * s2 * s2;
        scal = 12.0 * P.eps * ir6 * (ir6 - 1.0) * w;
        if (want_e) {
            double mid = 1.0 - ir6;
            if (KIND == PARM_PAIR_LJCUT)
                e = P.eps * (mid * mid - 1.0) - P.cutE; // :253-258
            else
                e = P.eps * (mid * mid) - P.cutE;       // :126-133 (cutE = 0), :1271-1288
        }
    }
}

// TEAM lanes share one atom: lane t of the team takes row entries t, t+TEAM, ... (the team reads
// TEAM consecutive indices = one or more full 32-byte sectors of the row), U entries per lane are
// in flight at once (index loads first, then the pos[j] gathers, then the arithmetic), and the
// team folds its partial force with xor-shuffles. TEAM*U divides 32 so rows (kmax % 32 == 0)
// are always readable up to the padded end.
// Per-pair mixing on the device for continuously polydisperse systems; same rules as mix() below
// (sqrt(e1 e2) is formed as sqrt(e1) sqrt(e2): identical to ~1 ulp).
template <int KIND>
__device__ __forceinline__ PairConst mix_dev(const double4 &qi, const double4 &qj, const double *__restrict__ tab, int nt,
                                             bool want_e) {
    PairConst P;
    P.sig = (qi.y + qj.y) * 0.5;
    P.sig2 = P.sig * P.sig;
    P.inv_sig2 = 0.0;
    P.cutE = 0.0;
    P.expo = 0.0;
    P.cut2 = 1.0;
    if (KIND == PARM_PAIR_LJREPULSE) {
        P.eps = qi.x * qj.x;
        P.rc2 = P.sig2;
    } else if (KIND == PARM_PAIR_REPULSION) {
        P.eps = qi.x * qj.x;
        P.expo = (qi.z + qj.z) * 0.5;
        P.rc2 = P.sig2;
    } else {
        double cut = fmax(qi.z, qj.z);
        double e;
        bool shifted = true;
        if (KIND == PARM_PAIR_LJATTRACTREPULSE) {
            e = __ldg(tab + (int)qi.w * nt + (int)qj.w);
            if (e <= 0) { // purely repulsive (interaction.hpp:1261-1266)
                cut = 1.0;
                e = fabs(e);
                shifted = false;
            }
        } else {
            e = qi.x * qj.x;
        }
        P.eps = e;
        P.cut2 = cut * cut;
        P.rc2 = P.cut2 * P.sig2;
        if (want_e && shifted) {
            const double ic6 = 1.0 / (P.cut2 * P.cut2 * P.cut2);
            const double mid = 1.0 - ic6;
            P.cutE = KIND == PARM_PAIR_LJCUT ? e * (mid * mid - 1.0) : e * (mid * mid);
        }
    }
    return P;
}

// SPEC: 0 one species (constants in registers), 1 species table in shared memory, 2 per-atom parameters
template <int KIND, int SPEC, int MODE, int TEAM, int U>
__global__ void __launch_bounds__(F_BLOCK)
k_force(const double4 *__restrict__ pos, const uint32_t *__restrict__ nbr, const uint32_t *__restrict__ cnt, uint32_t kmax,
        const uint8_t *__restrict__ spec, const PairConst *__restrict__ table, int nspecies, PairConst P1, double *f,
        uint32_t n, uint32_t npad, BoxDev box, int accumulate, double *partials, const double4 *__restrict__ par,
        const double *__restrict__ eps_tab, int ntypes, const int *__restrict__ abort_flag, uint32_t first) {
    if (abort_flag && *abort_flag) return; // speculatively enqueued step whose predecessor asked for a rebuild
    extern __shared__ PairConst s_table[];
    if (SPEC == 1) {
        for (int q = threadIdx.x; q < nspecies * nspecies; q += blockDim.x) s_table[q] = table[q];
        __syncthreads();
    }
    const uint32_t s = first + (blockIdx.x * blockDim.x + threadIdx.x) / TEAM; // slots [first, n)
    const uint32_t tl = threadIdx.x % TEAM;
    const bool want_obs = MODE != MODE_F;
    double fx = 0, fy = 0, fz = 0;
    double acc[NPART];
    if (want_obs)
#pragma unroll
        for (int q = 0; q < NPART; q++) acc[q] = 0.0;
    const bool valid = s < n;
    const uint32_t sc = valid ? s : 0;
    const uint32_t my = valid ? cnt[sc] : 0;
    const double4 pi = pos[sc];
    const uint32_t *row = nbr + (size_t)sc * kmax;
    const PairConst *prow = SPEC == 1 ? s_table + (int)spec[sc] * nspecies : nullptr;
    double4 qi = make_double4(0, 0, 0, 0);
    if (SPEC == 2) qi = par[sc];
    for (uint32_t k0 = tl; k0 < my; k0 += TEAM * U) {
        uint32_t j[U];
        bool ok[U];
        double4 pj[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t k = k0 + u * TEAM;
            ok[u] = k < my;
            j[u] = ok[u] ? __ldg(row + k) : sc;
        }
#pragma unroll
        for (int u = 0; u < U; u++) pj[u] = ld_pos4(pos + j[u]);
#pragma unroll
        for (int u = 0; u < U; u++) {
            // OriginBox::diff(atom1->x, atom2->x), box.hpp:103
            double dx = min_image_fast(pi.x - pj[u].x, box.L[0], box.invL[0]);
            double dy = min_image_fast(pi.y - pj[u].y, box.L[1], box.invL[1]);
            double dz = min_image_fast(pi.z - pj[u].z, box.L[2], box.invL[2]);
            double dsq = dx * dx + (dy * dy + dz * dz);
            double scal, e;
            if (SPEC == 0) {
                pair_eval<KIND>(P1, dsq, want_obs, scal, e);
            } else if (SPEC == 1) {
                const PairConst &P = prow[__ldg(spec + j[u])];
                pair_eval<KIND>(P, dsq, want_obs, scal, e);
            } else {
                const PairConst P = mix_dev<KIND>(qi, ld_pos4(par + j[u]), eps_tab, ntypes, want_obs);
                pair_eval<KIND>(P, dsq, want_obs, scal, e);
            }
            if (!ok[u]) { // padding lane (j == self, dsq == 0): contributes nothing
                scal = 0.0;
                e = 0.0;
            }
            double gx = dx * scal, gy = dy * scal, gz = dz * scal;
            fx += gx;
            fy += gy;
            fz += gz;
            if (want_obs) {
                acc[0] += e;
                acc[1] += dx * gx + (dy * gy + dz * gz); // r.dot(f), :2241
                acc[2] += dx * gx; acc[3] += dx * gy; acc[4] += dx * gz; // stress += r * f^T, :2274
                acc[5] += dy * gx; acc[6] += dy * gy; acc[7] += dy * gz;
                acc[8] += dz * gx; acc[9] += dz * gy; acc[10] += dz * gz;
                acc[11] += (e != 0.0) ? 1.0 : 0.0; // contacts :2126-2137
                acc[12] += (e > 0.0) ? 1.0 : 0.0;  // overlaps :2140-2151
            }
        }
    }
    if (MODE != MODE_OBS) {
#pragma unroll
        for (int o = TEAM / 2; o; o >>= 1) {
            fx += __shfl_xor_sync(0xffffffffu, fx, o);
            fy += __shfl_xor_sync(0xffffffffu, fy, o);
            fz += __shfl_xor_sync(0xffffffffu, fz, o);
        }
        if (valid && tl == 0) {
            if (accumulate) {
                f[s] += fx;
                f[npad + s] += fy;
                f[2 * (size_t)npad + s] += fz;
            } else {
                f[s] = fx;
                f[npad + s] = fy;
                f[2 * (size_t)npad + s] = fz;
            }
        }
    }
    if (want_obs) {
        __shared__ double red[NPART][F_BLOCK / 32];
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
        for (int q = 0; q < NPART; q++) {
            double x = acc[q];
#pragma unroll
            for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == 0) red[q][w] = x;
        }
        __syncthreads();
        if (threadIdx.x < NPART) {
            double x = 0;
            for (int ww = 0; ww < F_BLOCK / 32; ww++) x += red[threadIdx.x][ww];
            partials[(size_t)blockIdx.x * NPART + threadIdx.x] = x;
        }
    }
}

// folds the per-block partials; every pair was visited from both ends -> * 0.5
__global__ void k_force_fold(const double *__restrict__ partials, uint32_t nblocks, double *out) {
    __shared__ double red[256 / 32];
    for (int q = 0; q < NPART; q++) {
        double x = 0;
        for (uint32_t b = threadIdx.x; b < nblocks; b += blockDim.x) x += partials[(size_t)b * NPART + q];
#pragma unroll
        for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0;
            for (int w = 0; w < 256 / 32; w++) t += red[w];
            out[q] = t * 0.5;
        }
        __syncthreads();
    }
}

#define FARGS pos, nbr, cnt, kmax, spec, table, nsp, P1, f, n, npad, box, acc, partials, par, eps_tab, ntypes, abortf, first
#define FPARAMS const double4 *pos, const uint32_t *nbr, const uint32_t *cnt, uint32_t kmax, const uint8_t *spec, \
                const PairConst *table, int nsp, PairConst P1, double *f, uint32_t n, uint32_t npad, BoxDev box, int acc, \
                double *partials, const double4 *par, const double *eps_tab, int ntypes, const int *abortf, uint32_t first
template <int KIND, int SPEC, int TEAM, int U>
static cudaError_t launch_mode(int mode, dim3 grid, size_t smem, cudaStream_t st, FPARAMS) {
    if (mode == MODE_F) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(k_force<KIND, SPEC, MODE_F, TEAM, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_force<KIND, SPEC, MODE_F, TEAM, U><<<grid, F_BLOCK, smem, st>>>(FARGS);
    } else if (mode == MODE_FALL) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(k_force<KIND, SPEC, MODE_FALL, TEAM, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_force<KIND, SPEC, MODE_FALL, TEAM, U><<<grid, F_BLOCK, smem, st>>>(FARGS);
    } else {
        if (smem > 48 * 1024) cudaFuncSetAttribute(k_force<KIND, SPEC, MODE_OBS, TEAM, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_force<KIND, SPEC, MODE_OBS, TEAM, U><<<grid, F_BLOCK, smem, st>>>(FARGS);
    }
    return cudaGetLastError();
}

template <int KIND, int SPEC>
static cudaError_t launch_team(int team, int mode, dim3 grid, size_t smem, cudaStream_t st, FPARAMS) {
    static int u4 = -1;
    if (u4 < 0) { const char *e = getenv("PARM_B200_U"); u4 = e ? atoi(e) : 2; }
    if (team == 4 && u4 == 4) return launch_mode<KIND, SPEC, 4, 4>(mode, grid, smem, st, FARGS);
    if (team == 4 && u4 == 1) return launch_mode<KIND, SPEC, 4, 1>(mode, grid, smem, st, FARGS);
    if (team == 4) return launch_mode<KIND, SPEC, 4, 2>(mode, grid, smem, st, FARGS);
    if (team == 16) return launch_mode<KIND, SPEC, 16, 2>(mode, grid, smem, st, FARGS);
    return launch_mode<KIND, SPEC, 8, 2>(mode, grid, smem, st, FARGS);
}

template <int KIND>
static cudaError_t launch_kind(int specmode, int team, int mode, uint32_t natoms, size_t smem, cudaStream_t st, FPARAMS) {
    const dim3 grid((unsigned)(((size_t)natoms * team + F_BLOCK - 1) / F_BLOCK));
    if (specmode == 0) return launch_team<KIND, 0>(team, mode, grid, 0, st, FARGS);
    if (specmode == 1) return launch_team<KIND, 1>(team, mode, grid, smem, st, FARGS);
    return launch_team<KIND, 2>(team, mode, grid, 0, st, FARGS);
}
#undef FARGS
#undef FPARAMS

// d_out: device pointer to NPART doubles (E, virial, stress[9], contacts, overlaps) or NULL
static int launch_forces(parm_inter *it, int mode, bool accumulate, double *d_out, const int *abort_flag = nullptr,
                         uint32_t first = 0, uint32_t count = 0xffffffffu) {
    parm_ctx *c = it->ctx;
    parm_nlist *nl = it->nl;
    if (!it->have_params || nl->updatenum == 0) {
        // no atoms add()ed yet, or the list was never built: the reference iterates an empty pair vector
        if (mode != MODE_OBS && !accumulate) CK(cudaMemsetAsync(c->f, 0, 3 * (size_t)c->npad * 8, c->stream));
        if (d_out) CK(cudaMemsetAsync(d_out, 0, NPART * 8, c->stream));
        return 0;
    }
    // lanes per atom: enough entries per lane to keep its loop busy, few enough to fill the last pass
    int team = 4; // measured at n ~ 110 (N=1e6 LJ): TEAM=4 0.368 ms, TEAM=8 0.385 ms, TEAM=16 0.476 ms
    {
        double mean = (double)nl->total_full / (double)(parm_owned(c) ? parm_owned(c) : 1);
        if (mean > 400) team = 8;
        static int forced = -1;
        if (forced < 0) { const char *e = getenv("PARM_B200_TEAM"); forced = e ? atoi(e) : 0; }
        if (forced == 4 || forced == 8 || forced == 16) team = forced;
    }
    uint32_t nown = parm_owned(c); // rows exist for owned atoms only
    if (first > nown) first = nown;
    nown = count == 0xffffffffu ? nown : std::min(nown, first + count); // slots [first, nown)
    const uint32_t nrange = nown - first;
    const uint32_t nblocks = (uint32_t)(((size_t)nrange * team + F_BLOCK - 1) / F_BLOCK);
    if (nrange == 0) {
        if (d_out) CK(cudaMemsetAsync(d_out, 0, NPART * 8, c->stream));
        return 0;
    }
    if (mode != MODE_F && (size_t)nblocks * NPART > it->partial_doubles) {
        if (it->d_partials) cudaFree(it->d_partials);
        it->d_partials = 0;
        it->partial_doubles = (size_t)nblocks * NPART;
        CK(cudaMalloc(&it->d_partials, it->partial_doubles * 8));
    }
    const int specmode = it->generic ? 2 : (it->nspecies == 1 ? 0 : 1);
    size_t smem = specmode == 1 ? (size_t)it->nspecies * it->nspecies * sizeof(PairConst) : 0;
    PairConst P1 = it->h_table[0];
    cudaError_t e;
#define ARGS specmode, team, mode, nrange, smem, c->stream, c->pos, nl->nbr, nl->cnt, nl->kmax, it->d_spec, it->d_table, \
             it->nspecies, P1, c->f, nown, c->npad, c->box, accumulate ? 1 : 0, it->d_partials, it->d_par, it->d_eps_table, \
             it->ntypes, abort_flag, first
    switch (it->kind) {
        case PARM_PAIR_LJREPULSE: e = launch_kind<PARM_PAIR_LJREPULSE>(ARGS); break;
        case PARM_PAIR_REPULSION: e = launch_kind<PARM_PAIR_REPULSION>(ARGS); break;
        case PARM_PAIR_LJATTRACTREPULSE: e = launch_kind<PARM_PAIR_LJATTRACTREPULSE>(ARGS); break;
        default: e = launch_kind<PARM_PAIR_LJCUT>(ARGS); break;
    }
#undef ARGS
    parm_count_launch(c);
    CK(e);
    if (mode != MODE_F) {
        if (!d_out) { parm_set_error("internal: observables requested without an output buffer"); return PARM_ERR_RUNTIME; }
        k_force_fold<<<1, 256, 0, c->stream>>>(it->d_partials, nblocks, d_out);
        CK_LAUNCH(c);
    }
    return 0;
}

int parm_inter_launch_forces(parm_inter *it, unsigned want, bool accumulate, double *d_out, const int *abort_flag,
                             uint32_t first, uint32_t count) {
    return launch_forces(it, want ? MODE_FALL : MODE_F, accumulate, d_out, abort_flag, first, count);
}

// ---- host API ---------------------------------------------------------------------------
extern "C" int parm_inter_create(parm_ctx *c, parm_nlist *nl, int kind, parm_inter **out) {
    if (!c || !nl || !out) { parm_set_error("parm_inter_create: NULL argument"); return PARM_ERR_INVALID; }
    *out = 0;
    if (nl->ctx != c) { parm_set_error("parm_inter_create: NeighborList belongs to another AtomVec"); return PARM_ERR_INVALID; }
    if (kind < 0 || kind > 3) {
        parm_set_error("parm_inter_create: pair type %d is outside the hot-path scope (supported: LJRepulsePair, "
                       "RepulsionPair, LJAttractRepulsePair, LennardJonesCutPair)", kind);
        return PARM_ERR_UNSUPPORTED;
    }
    CK(cudaSetDevice(c->device));
    parm_inter *it = new parm_inter();
    it->ctx = c;
    it->nl = nl;
    it->kind = kind;
    it->h_spec_id.assign(c->nid, 0);
    CK(cudaMalloc(&it->d_spec_id, std::max(c->npad, c->nid_pad)));
    CK(cudaMalloc(&it->d_spec, c->npad));
    CK(cudaMemsetAsync(it->d_spec_id, 0, std::max(c->npad, c->nid_pad), c->stream));
    CK(cudaMemsetAsync(it->d_spec, 0, c->npad, c->stream));
    CK(cudaMalloc(&it->d_table, sizeof(PairConst) * PARM_MAX_SPECIES * PARM_MAX_SPECIES));
    c->inters.push_back(it);
    *out = it;
    return 0;
}

extern "C" int parm_inter_destroy(parm_inter *it) {
    if (!it) return 0;
    parm_ctx *c = it->ctx;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (it->d_spec_id) cudaFree(it->d_spec_id);
    if (it->d_spec) cudaFree(it->d_spec);
    if (it->d_table) cudaFree(it->d_table);
    if (it->d_partials) cudaFree(it->d_partials);
    if (it->d_par_id) cudaFree(it->d_par_id);
    if (it->d_par) cudaFree(it->d_par);
    if (it->d_eps_table) cudaFree(it->d_eps_table);
    c->inters.erase(std::remove(c->inters.begin(), c->inters.end(), it), c->inters.end());
    delete it;
    return 0;
}

__global__ void k_gather_par(const double4 *__restrict__ par_id, const uint32_t *__restrict__ order, uint32_t n, double4 *par) {
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) par[s] = par_id[order[s]];
}
__global__ void k_gather_spec(const uint8_t *__restrict__ spec_id, const uint32_t *__restrict__ order, uint32_t n, uint8_t *spec) {
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) spec[s] = spec_id[order[s]];
}

int parm_inter_regather(parm_inter *it) {
    parm_ctx *c = it->ctx;
    if (!c->n) return 0;
    unsigned nb = (c->n + 255) / 256;
    unsigned cap = (unsigned)c->num_sms * 8;
    if (it->generic) k_gather_par<<<nb < cap ? nb : cap, 256, 0, c->stream>>>(it->d_par_id, c->order, c->n, it->d_par);
    else k_gather_spec<<<nb < cap ? nb : cap, 256, 0, c->stream>>>(it->d_spec_id, c->order, c->n, it->d_spec);
    CK_LAUNCH(c);
    return 0;
}

// Pair constructors of the reference, evaluated once per species pair on the host.
static PairConst mix(int kind, const double *p1, uint32_t t1, const double *p2, uint32_t t2, const double *tab, int nt) {
    PairConst P;
    memset(&P, 0, sizeof(P));
    if (kind == PARM_PAIR_LJREPULSE) { // :878-883
        P.eps = sqrt(p1[0] * p2[0]);
        P.sig = (p1[1] + p2[1]) / 2;
        P.cut2 = 1.0;
    } else if (kind == PARM_PAIR_REPULSION) { // :1531-1536
        P.eps = sqrt(p1[0] * p2[0]);
        P.sig = (p1[1] + p2[1]) / 2.0;
        P.expo = (p1[2] + p2[2]) / 2.0;
    } else if (kind == PARM_PAIR_LJATTRACTREPULSE) { // :1255-1270
        double eps = tab[(size_t)t1 * nt + t2];
        P.sig = (p1[1] + p2[1]) / 2.0;
        double cut = std::max(p1[2], p2[2]);
        if (eps <= 0) {
            cut = 1;
            P.cutE = 0;
            eps = fabs(eps);
        } else {
            double mid = (1 - pow(cut, -6));
            P.cutE = eps * (mid * mid);
        }
        P.eps = eps;
        P.cut2 = cut * cut;
    } else { // :970-974, :247-252
        P.eps = sqrt(p1[0] * p2[0]);
        P.sig = (p1[1] + p2[1]) / 2;
        double cut = std::max(p1[2], p2[2]);
        double rsix = pow(cut, 6);
        double mid = (1 - 1 / rsix);
        P.cutE = P.eps * (mid * mid - 1);
        P.cut2 = cut * cut;
    }
    P.sig2 = P.sig * P.sig;
    P.inv_sig2 = 1.0 / P.sig2;
    P.rc2 = kind == PARM_PAIR_REPULSION ? P.sig2 : P.cut2 * P.sig2;
    return P;
}

extern "C" int parm_inter_set_params(parm_inter *it, const double *params, const uint32_t *type, const double *eps_table,
                                     int ntypes, const uint8_t *member, int set_diameters) {
    if (!it || !params) { parm_set_error("parm_inter_set_params: NULL argument"); return PARM_ERR_INVALID; }
    parm_ctx *c = it->ctx;
    CK(cudaSetDevice(c->device));
    const int kind = it->kind;
    if (kind == PARM_PAIR_LJATTRACTREPULSE) {
        if (!eps_table || ntypes < 1) { parm_set_error("LJAttractRepulsePair needs the epsilon table (IEpsSigCutAtom::epsilons)"); return PARM_ERR_INVALID; }
        for (int a = 0; a < ntypes; a++)
            for (int b = 0; b < ntypes; b++)
                if (eps_table[a * ntypes + b] != eps_table[b * ntypes + a]) { // assert at interaction.hpp:1013-1014
                    parm_set_error("LJAttractRepulsePair: epsilon table must be symmetric");
                    return PARM_ERR_INVALID;
                }
    }
    typedef std::tuple<double, double, double, uint32_t> Key;
    bool too_many = false;
    std::map<Key, int> ids;
    std::vector<Key> keys;
    std::vector<double> diam(c->nid, -1.0);
    for (uint32_t i = 0; i < c->nid; i++) {
        if (member && !member[i]) { it->h_spec_id[i] = 0; continue; }
        const double *p = params + 3 * (size_t)i;
        uint32_t t = (kind == PARM_PAIR_LJATTRACTREPULSE && type) ? type[i] : 0;
        if (kind == PARM_PAIR_LJATTRACTREPULSE && (int)t >= ntypes) { parm_set_error("atom %u: type %u >= ntypes %d", i, t, ntypes); return PARM_ERR_INVALID; }
        Key k(kind == PARM_PAIR_LJATTRACTREPULSE ? 0.0 : p[0], p[1], kind == PARM_PAIR_LJREPULSE ? 0.0 : p[2], t);
        auto f = ids.find(k);
        int id;
        if (f == ids.end()) {
            id = (int)keys.size();
            if (id >= PARM_MAX_SPECIES) {
                too_many = true; // continuous polydispersity: per-atom parameters, mixed per pair on the device
                id = 0;
            } else {
                ids[k] = id;
                keys.push_back(k);
            }
        } else
            id = f->second;
        it->h_spec_id[i] = (uint8_t)id;
        // A::max_size(): sigma (:864, :1464) or sigma*sigcut (:905, :1017)
        diam[i] = (kind == PARM_PAIR_LJREPULSE || kind == PARM_PAIR_REPULSION) ? p[1] : p[1] * p[2];
    }
    it->generic = too_many;
    it->ntypes = ntypes > 0 ? ntypes : 1;
    if (too_many) {
        keys.resize(1);
        it->h_par_id.assign(4 * (size_t)std::max(c->npad, c->nid_pad), 0.0);
        for (uint32_t i = 0; i < c->nid; i++) {
            if (member && !member[i]) continue;
            const double *p = params + 3 * (size_t)i;
            double *q = &it->h_par_id[4 * (size_t)i];
            q[0] = kind == PARM_PAIR_LJATTRACTREPULSE ? 0.0 : sqrt(p[0]);
            q[1] = p[1];
            q[2] = kind == PARM_PAIR_LJREPULSE ? 0.0 : p[2];
            q[3] = (kind == PARM_PAIR_LJATTRACTREPULSE && type) ? (double)type[i] : 0.0;
        }
        const size_t npar = std::max(c->npad, c->nid_pad);
        if (!it->d_par_id) CK(cudaMalloc(&it->d_par_id, npar * sizeof(double4)));
        if (!it->d_par) CK(cudaMalloc(&it->d_par, (size_t)c->npad * sizeof(double4)));
        CK(cudaMemcpyAsync(it->d_par_id, it->h_par_id.data(), npar * sizeof(double4), cudaMemcpyHostToDevice, c->stream));
        if (it->d_eps_table) { cudaFree(it->d_eps_table); it->d_eps_table = 0; }
        CK(cudaMalloc(&it->d_eps_table, sizeof(double) * it->ntypes * it->ntypes));
        if (eps_table) CK(cudaMemcpyAsync(it->d_eps_table, eps_table, sizeof(double) * it->ntypes * it->ntypes, cudaMemcpyHostToDevice, c->stream));
    }
    int S = (int)keys.size();
    if (S == 0) { S = 1; keys.push_back(Key(1.0, 1.0, 1.0, 0)); }
    it->nspecies = S;
    it->h_table.assign((size_t)S * S, PairConst());
    for (int a = 0; a < S; a++)
        for (int b = 0; b < S; b++) {
            double p1[3] = {std::get<0>(keys[a]), std::get<1>(keys[a]), std::get<2>(keys[a])};
            double p2[3] = {std::get<0>(keys[b]), std::get<1>(keys[b]), std::get<2>(keys[b])};
            it->h_table[(size_t)a * S + b] = mix(kind, p1, std::get<3>(keys[a]), p2, std::get<3>(keys[b]), eps_table, ntypes);
        }
    CK(cudaMemcpyAsync(it->d_table, it->h_table.data(), sizeof(PairConst) * S * S, cudaMemcpyHostToDevice, c->stream));
    if (c->nid) CK(cudaMemcpyAsync(it->d_spec_id, it->h_spec_id.data(), c->nid, cudaMemcpyHostToDevice, c->stream));
    PTRY(parm_inter_regather(it));
    CK(cudaStreamSynchronize(c->stream));
    it->have_params = true;
    if (set_diameters) PTRY(parm_nlist_set_diameters(it->nl, diam.data()));
    return 0;
}

static int fetch(parm_inter *it, int mode, bool accumulate, double *host13) {
    parm_ctx *c = it->ctx;
    CK(cudaSetDevice(c->device));
    PTRY(parm_ctx_ensure_red(c, 64));
    PTRY(launch_forces(it, mode, accumulate, c->d_red));
    if (c->sh.on) PTRY(parm_shard_allreduce_sum(c, c->d_red, NPART));
    CK(cudaMemcpyAsync(c->h_red, c->d_red, NPART * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    memcpy(host13, c->h_red, NPART * 8);
    return 0;
}

extern "C" int parm_inter_set_forces(parm_inter *it, unsigned want, double *out) {
    if (!it) { parm_set_error("parm_inter_set_forces: NULL interaction"); return PARM_ERR_INVALID; }
    parm_ctx *c = it->ctx;
    CK(cudaSetDevice(c->device));
    if (!want) return launch_forces(it, MODE_F, true, nullptr);
    if (!out) { parm_set_error("parm_inter_set_forces: out is NULL"); return PARM_ERR_INVALID; }
    double r[NPART];
    PTRY(fetch(it, MODE_FALL, true, r));
    int k = 0;
    const int D = c->D;
    if (want & PARM_WANT_ENERGY) out[k++] = r[0];
    if (want & PARM_WANT_VIRIAL) out[k++] = r[1];
    if (want & PARM_WANT_STRESS)
        for (int a = 0; a < D; a++)
            for (int b = 0; b < D; b++) out[k++] = r[2 + a * 3 + b];
    return 0;
}
extern "C" int parm_inter_energy(parm_inter *it, double *E) {
    double r[NPART];
    PTRY(fetch(it, MODE_OBS, false, r));
    *E = r[0];
    return 0;
}
extern "C" int parm_inter_pressure(parm_inter *it, double *p) {
    double r[NPART];
    PTRY(fetch(it, MODE_OBS, false, r));
    *p = r[1];
    return 0;
}
extern "C" int parm_inter_stress(parm_inter *it, double *st) {
    double r[NPART];
    PTRY(fetch(it, MODE_OBS, false, r));
    const int D = it->ctx->D;
    for (int a = 0; a < D; a++)
        for (int b = 0; b < D; b++) st[a * D + b] = r[2 + a * 3 + b];
    return 0;
}
extern "C" int parm_inter_contacts(parm_inter *it, uint64_t *contacts, uint64_t *overlaps) {
    double r[NPART];
    PTRY(fetch(it, MODE_OBS, false, r));
    if (contacts) *contacts = (uint64_t)llround(r[11]);
    if (overlaps) *overlaps = (uint64_t)llround(r[12]);
    return 0;
}
