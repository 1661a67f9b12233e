// Small systems (BASELINE config 1: LJatoms.cpp, N = 1000): CollectionVerlet::timestep (collection.cpp:442-469) as ONE
// persistent kernel for a whole timestep(n) call.
//
// With a few thousand atoms a step is not bandwidth or arithmetic: it is a chain of dependent kernel latencies (two
// kernels per step, each 5-8 us at N = 1000: launch, flag read, loads, reduction, last-block fold ...). Here the grid
// stays resident for all the steps of the call (cooperative launch: every block is guaranteed to be co-resident):
//   * 16 lanes per atom, 16 atoms per block of 256 threads; the atom's x, v, a, lastlocs live in the registers of lanes
//     0..2 of its team (one component each), its neighbour row (<= 128 entries) in the registers of all 16 lanes --
//     nothing but the new positions is written per step;
//   * per step: K1 in registers (collection.cpp:443-451) -> new coordinates to a ping-pong array + the block's top-2
//     displacement (trackers.cpp:23-53) to a slot -> ONE grid barrier -> every block copies all positions to its shared
//     memory and folds the slots into the step's rebuild decision (identical arithmetic everywhere) -> pair forces from
//     shared memory (NListed::set_forces, interaction.hpp:2154-2175; full rows, fixed order, no atomics) -> K3 in
//     registers (collection.cpp:457-465);
//   * the kernel ends after the step whose drift rule fired (update_trackers(), collection.cpp:468): the host rebuilds
//     the list and launches again for the remaining steps.
// Same expressions in the same order as k_verlet1 / k_verlet2 (integ.cu), same drift rule as drift.cuh; the pair
// arithmetic is the cell-tile kernel's (force_tile.cuh) with the per-pair minimum image of the gather kernel.
// Eligible: CollectionVerlet, 3-D, single GPU, one one-species Lennard-Jones-family interaction on the tracked list,
// rows <= 128 entries, N <= what one co-resident grid covers and one shared-memory copy of the positions holds.
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "drift.cuh"
#include "force_tile.cuh"
#include "internal.cuh"

#define SM_NT 256
#define SM_TEAM 16
#define SM_APB (SM_NT / SM_TEAM) // atoms per block
#define SM_ROWQ 8                // row entries per lane (rows <= SM_TEAM * SM_ROWQ)

struct SmallArgs {
    double4 *pos;
    double *v, *a, *f;
    const double *xlast;
    const uint32_t *nbr, *cnt;
    uint32_t kmax, mask;
    uint32_t n, npad, nS; // nS: stride of the scratch coordinate arrays
    PairConst P1;
    BoxDev box;
    double dt, hdt2, hdt, skin;
    double *P;         // [2][3][nS] coordinates of step parity 0 / 1
    double *slots;     // [2][grid][2] per-block top-2 displacements
    unsigned int *bar; // grid barrier counter (zero at launch)
    int *d_res, *h_res; // [0] steps done, [1] rebuild requested
    NlistFlags *dflags, *hflags;
    int nsteps;
};

struct SmallState { // hung off parm_integ
    double *P, *slots;
    unsigned int *bar;
    int *d_res, *h_res;
    uint32_t nS, grid_cap;
    bool unavailable; // a cooperative launch failed once: this integrator stays on the general path
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// (one kernel for LJRepulsePair :875-891, LJAttractRepulsePair :1271-1298 and LennardJonesCutPair :253-267: their forces
// differ only in the constants of PairConst -- sigma^2, 12 epsilon, cut-off --, exactly as in force_tile.cuh)
__global__ void __launch_bounds__(SM_NT) k_small_steps(const SmallArgs A) {
    extern __shared__ __align__(16) double s_pos[]; // [3][nS]; slot nS - 1 is the far-away sentinel of the row pads
    __shared__ double s_b1[SM_NT / 32], s_b2[SM_NT / 32];
    __shared__ int s_need;
    __shared__ __align__(8) unsigned long long s_mbar;
    const uint32_t tid = threadIdx.x, tl = tid & (SM_TEAM - 1), lane = tid & 31u, w = tid >> 5;
    const uint32_t s = blockIdx.x * SM_APB + tid / SM_TEAM;
    const bool valid = s < A.n;
    const bool comp = valid && tl < 3; // this lane carries coordinate tl of the atom
    const uint32_t G = gridDim.x, nS = A.nS;
    const double nanv = __longlong_as_double(0x7ff8000000000000LL);
    double m = 1.0, xc = 0.0, vc = 0.0, ac = 0.0, xl = nanv, fc = 0.0;
    if (valid) m = A.pos[s].w;
    if (comp) {
        xc = reinterpret_cast<const double *>(A.pos + s)[tl];
        vc = A.v[(size_t)tl * A.npad + s];
        ac = A.a[(size_t)tl * A.npad + s];
        xl = A.xlast[(size_t)tl * A.npad + s];
    }
    const bool frozen = frozen_le(m);
    // the neighbour row stays in registers for the whole call (the list only changes between launches)
    const uint32_t my = valid ? min(A.cnt[s], A.kmax) : 0u;
    uint32_t ent[SM_ROWQ];
#pragma unroll
    for (int q = 0; q < SM_ROWQ; q++) {
        const uint32_t k = tl + (uint32_t)q * SM_TEAM;
        ent[q] = k < my ? (__ldg(A.nbr + (size_t)s * A.kmax + k) & A.mask) : nS - 1u; // pads: the sentinel (exact zeros, no branches)
    }
    const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(&s_mbar);
    if (tid == 0) mbar_init(mbar, 1);
    if (blockIdx.x == 0 && tid < 6) A.P[(size_t)tid * nS + nS - 1] = 1e100; // the sentinel coordinate of both step parities
    const double c12 = 12.0 * A.P1.eps;
    const long long rc2_bits = __double_as_longlong(A.P1.rc2);
    unsigned int target = 0;
    int step = 0, need = 0;
    while (step < A.nsteps) {
        const int buf = step & 1;
        double *Pb = A.P + (size_t)buf * 3 * nS;
        double *Sb = A.slots + (size_t)buf * 2 * G;
        // ---- K1: x += v dt + a dt^2/2; v += a dt/2 (frozen atoms: v = 0), collection.cpp:443-451
        if (comp) {
            if (frozen) {
                vc = 0.0;
            } else {
                xc = __dadd_rn(xc, __dadd_rn(__dmul_rn(vc, A.dt), __dmul_rn(ac, A.hdt2)));
                vc = __dadd_rn(vc, __dmul_rn(ac, A.hdt));
            }
            Pb[(size_t)tl * nS + s] = xc;
        }
        // ---- displacement since the last rebuild: sqrt(e0 + (e1 + e2)), trackers.cpp:27; NaN lastlocs never win
        {
            const double dc = comp ? __dsub_rn(xc, xl) : 0.0;
            const double e0 = __dmul_rn(dc, dc);
            const double e1 = __shfl_sync(0xffffffffu, e0, 1, SM_TEAM), e2 = __shfl_sync(0xffffffffu, e0, 2, SM_TEAM);
            double b1 = 0.0, b2 = 0.0;
            if (valid && tl == 0) top2_push(b1, b2, __dsqrt_rn(__dadd_rn(e0, __dadd_rn(e1, e2))));
            top2_warp(b1, b2);
            if (lane == 0) {
                s_b1[w] = b1;
                s_b2[w] = b2;
            }
            __syncthreads();
            if (w == 0) {
                b1 = lane < SM_NT / 32 ? s_b1[lane] : 0.0;
                b2 = lane < SM_NT / 32 ? s_b2[lane] : 0.0;
                top2_warp(b1, b2);
                if (lane == 0) {
                    Sb[2 * blockIdx.x] = b1;
                    Sb[2 * blockIdx.x + 1] = b2;
                }
            }
        }
        // ---- grid barrier: every block's coordinates and top-2 slot are out (the coordinates are read back through the
        // async proxy -- bulk copies --, so the writers order their generic-proxy stores against it as well)
        asm volatile("fence.proxy.async;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(A.bar) : "memory");
            target += G;
            while (ld_acquire_u32(A.bar) < target) {
            }
            __threadfence();
            // ---- all coordinates to shared memory: three bulk copies (TMA), one round trip whatever N is
            asm volatile("fence.proxy.async;" ::: "memory");
            mbar_arrive_expect_tx(mbar, 3u * nS * 8u);
#pragma unroll
            for (int d = 0; d < 3; d++) bulk_g2s((uint32_t)__cvta_generic_to_shared(s_pos + (size_t)d * nS), Pb + (size_t)d * nS, nS * 8u, mbar);
        }
        __syncthreads();
        // ---- the step's decision from all slots (same arithmetic in every block), while the copies are in flight
        {
            double b1 = 0.0, b2 = 0.0;
            for (uint32_t t = tid; t < G; t += SM_NT) top2_merge(b1, b2, __ldcg(Sb + 2 * t), __ldcg(Sb + 2 * t + 1));
            top2_warp(b1, b2);
            if (lane == 0) {
                s_b1[w] = b1;
                s_b2[w] = b2;
            }
            __syncthreads();
            if (w == 0) {
                b1 = lane < SM_NT / 32 ? s_b1[lane] : 0.0;
                b2 = lane < SM_NT / 32 ? s_b2[lane] : 0.0;
                top2_warp(b1, b2);
                if (lane == 0) s_need = (__dadd_rn(b2, b1) >= A.skin) ? 1 : 0; // bigdist + biggestdist >= skin
            }
            __syncthreads();
            need = s_need;
        }
        mbar_wait(mbar, (uint32_t)step & 1u);
        // ---- pair forces of this atom's row (OriginBox::diff per pair, box.hpp:103)
        double fx = 0.0, fy = 0.0, fz = 0.0;
        {
            const uint32_t si = valid ? s : 0u;
            const double xi = s_pos[si], yi = s_pos[nS + si], zi = s_pos[2 * nS + si];
#pragma unroll
            for (int q = 0; q < SM_ROWQ; q++) {
                const uint32_t j = ent[q];
                {
                    const double dx = min_image_fast(xi - s_pos[j], A.box.L[0], A.box.invL[0]);
                    const double dy = min_image_fast(yi - s_pos[nS + j], A.box.L[1], A.box.invL[1]);
                    const double dz = min_image_fast(zi - s_pos[2 * nS + j], A.box.L[2], A.box.invL[2]);
                    // (pads are recognised by their index: the minimum image folds the far-away sentinel back into the box,
                    // to distance zero when a box edge is a power of two)
                    const double dsq = j == nS - 1u ? 1e200 : dx * dx + (dy * dy + dz * dz);
                    const double wr = rcp_pos(dsq);
                    const double s2 = A.P1.sig2 * wr;
                    const double ir6 = (s2 * s2) * s2;
                    const double b = ir6 * wr;
                    const double t = fma(b, ir6, -b);
                    const double scal = __double_as_longlong(dsq) <= rc2_bits ? t : 0.0;
                    fx = fma(dx, scal, fx);
                    fy = fma(dy, scal, fy);
                    fz = fma(dz, scal, fz);
                }
            }
            fx *= c12;
            fy *= c12;
            fz *= c12;
#pragma unroll
            for (int o = SM_TEAM / 2; o; o >>= 1) {
                fx += __shfl_xor_sync(0xffffffffu, fx, o);
                fy += __shfl_xor_sync(0xffffffffu, fy, o);
                fz += __shfl_xor_sync(0xffffffffu, fz, o);
            }
        }
        // ---- K3: a = f / m; v += a dt/2 (frozen atoms: a = 0), collection.cpp:457-465
        if (comp) {
            fc = tl == 0 ? fx : (tl == 1 ? fy : fz);
            if (frozen) {
                ac = 0.0;
            } else {
                ac = __ddiv_rn(fc, m);
                vc = __dadd_rn(vc, __dmul_rn(ac, A.hdt));
            }
        }
        step++;
        if (need) break; // update_trackers(): the list is rebuilt before the next step (by the host)
    }
    if (comp) {
        reinterpret_cast<double *>(A.pos + s)[tl] = xc;
        A.v[(size_t)tl * A.npad + s] = vc;
        A.a[(size_t)tl * A.npad + s] = ac;
        A.f[(size_t)tl * A.npad + s] = fc;
    }
    if (blockIdx.x == 0 && tid == 0) {
        A.d_res[0] = step;
        A.d_res[1] = need;
        A.h_res[0] = step;
        A.h_res[1] = need;
        A.dflags->need_rebuild = need;
        A.hflags->need_rebuild = need;
    }
}

static cudaError_t small_launch(unsigned grid, size_t smem, cudaStream_t st, SmallArgs &A, int num_sms, bool *fits) {
    cudaError_t e = cudaSuccess;
    if (smem > 40 * 1024) e = cudaFuncSetAttribute(k_small_steps, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_small_steps, SM_NT, smem);
    if (e != cudaSuccess) return e;
    *fits = (unsigned)per_sm * (unsigned)num_sms >= grid;
    if (!*fits) return cudaSuccess;
    void *args[] = {(void *)&A};
    return cudaLaunchCooperativeKernel((const void *)k_small_steps, dim3(grid), dim3(SM_NT), args, smem, st);
}

void parm_small_free(parm_integ *g) {
    SmallState *S = g->small;
    if (!S) return;
    if (S->P) cudaFree(S->P);
    if (S->slots) cudaFree(S->slots);
    if (S->bar) cudaFree(S->bar);
    if (S->d_res) cudaFree(S->d_res);
    if (S->h_res) cudaFreeHost(S->h_res);
    delete S;
    g->small = nullptr;
}

static bool small_eligible(const parm_integ *g, const parm_nlist *nl) {
    const parm_ctx *c = g->ctx;
    const char *e = getenv("PARM_B200_SMALL_PERSIST"); // (read per call)
    const int on = e ? atoi(e) : 1;
    e = getenv("PARM_B200_SMALL_PERSIST_NMAX");
    const int nmax = e ? atoi(e) : 4096;
    if (!on || g->type != 0 || c->sh.on || c->D != 3 || c->prof_on || !nl || nl->ignorechanged || nl->updatenum == 0) return false;
    if (c->n == 0 || c->n > (uint32_t)nmax || !g->stat_trackers.empty() || g->inters.size() != 1 || g->trackers.size() != 1) return false;
    const parm_inter *it = g->inters[0];
    const int kk = PARM_KERNEL_KIND(it->kind);
    if (it->nl != nl || !it->have_params || it->generic || it->nspecies != 1) return false;
    if (!(kk == PARM_PAIR_LJREPULSE || kk == PARM_PAIR_LJATTRACTREPULSE || kk == PARM_PAIR_LJCUT)) return false;
    if (nl->h_flags->maxcnt > SM_TEAM * SM_ROWQ) return false;
    return true;
}

// Runs as many of the `*nsteps` steps as stay eligible; *nsteps holds what is left for the general path (0: all done).
int parm_small_run(parm_integ *g, parm_nlist *nl, int *nsteps) {
    parm_ctx *c = g->ctx;
    while (*nsteps > 0 && small_eligible(g, nl) && !(g->small && g->small->unavailable)) {
        PTRY(parm_nlist_ensure_rows32(nl));
        const uint32_t n = c->n;
        const unsigned grid = (n + SM_APB - 1) / SM_APB;
        const uint32_t nS = (n + 16u) & ~15u; // a multiple of 16 with at least one free slot behind the atoms: the sentinel
        if (!g->small) {
            g->small = new SmallState();
            memset(g->small, 0, sizeof(SmallState));
        }
        SmallState *S = g->small;
        if (nS > S->nS || grid > S->grid_cap) {
            if (S->P) cudaFree(S->P);
            if (S->slots) cudaFree(S->slots);
            S->P = S->slots = nullptr;
            S->nS = nS + 256;
            S->grid_cap = grid + 16;
            CK(cudaMalloc(&S->P, 2 * 3 * (size_t)S->nS * 8));
            CK(cudaMalloc(&S->slots, 2 * 2 * (size_t)S->grid_cap * 8));
        }
        if (!S->bar) {
            CK(cudaMalloc(&S->bar, 4));
            CK(cudaMalloc(&S->d_res, 8));
            CK(cudaHostAlloc(&S->h_res, 8, cudaHostAllocDefault));
        }
        const parm_inter *it = g->inters[0];
        SmallArgs A;
        A.pos = c->pos; A.v = c->v; A.a = c->a; A.f = c->f;
        A.xlast = nl->xlast;
        A.nbr = nl->nbr; A.cnt = nl->cnt;
        A.kmax = nl->kmax; A.mask = PARM_NBR_MASK_OF(nl);
        A.n = n; A.npad = c->npad; A.nS = nS;
        A.P1 = it->h_table[0];
        A.box = c->box;
        A.dt = g->dt; A.hdt2 = g->dt * g->dt / 2; A.hdt = g->dt / 2; A.skin = nl->skin;
        A.P = S->P; A.slots = S->slots; A.bar = S->bar;
        A.d_res = S->d_res; A.h_res = S->h_res;
        A.dflags = nl->d_flags; A.hflags = nl->h_flags;
        A.nsteps = std::min(*nsteps, 1 << 20); // (the barrier counter is 32 bits wide)
        S->h_res[0] = -1;
        S->h_res[1] = 0;
        CK(cudaMemsetAsync(S->bar, 0, 4, c->stream));
        const size_t smem = 3 * (size_t)nS * 8;
        bool fits = false;
        {
            const cudaError_t e = small_launch(grid, smem, c->stream, A, c->num_sms, &fits);
            if (e != cudaSuccess) { // no cooperative launch here (e.g. under a tool that does not support it): general path
                cudaGetLastError();
                S->unavailable = true;
                return 0;
            }
        }
        if (!fits) return 0; // the grid would not be co-resident: general path
        parm_count_launch(c);
        CK(cudaStreamSynchronize(c->stream));
        const int done = S->h_res[0], need = S->h_res[1];
        if (done <= 0 || done > *nsteps) { parm_set_error("small-system kernel returned %d steps of %d", done, *nsteps); return PARM_ERR_RUNTIME; }
        g->steps += (uint64_t)done;
        *nsteps -= done;
        if (need) {
            PTRY(parm_nlist_rebuild(nl));
            g->rebuilds++;
        }
    }
    return 0;
}
