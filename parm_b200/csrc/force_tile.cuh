// Cell-tile variant of the NListed pair kernel (interaction.hpp:2154-2291) for single-species interactions
// with long rows (the LJ configurations of BASELINE.json). One block per chunk (tile.cu):
//   1. stage the chunk's tile: every position of the <= 18 contiguous slot runs is loaded ONCE, coalesced,
//      reduced to min_image(x - origin) and stored in shared memory as (x, y) double2 + z (24 B per atom);
//   2. TEAM lanes per atom walk the atom's 16-bit tile-local row: one vector load brings V entries per lane
//      (requested one pass ahead), every pair costs one LDS.128 + one LDS.64 instead of a 32-byte global
//      gather, the atom's own position comes from the tile too, and there is no per-pair minimum
//      image (OriginBox::diff, box.hpp:103, is applied once per staged atom) unless the tile is wider than
//      half the box (small boxes), in which case the per-pair form runs on top.
// Same FULL rows, same fixed summation order inside a lane, team and block as force_kernel.cuh: no atomics,
// run-to-run deterministic.
#pragma once
#include <algorithm>
#include "force_kernel.cuh"

struct TileForceArgs {
    const double4 *pos;
    const double2 *prel_xy;  // STAGE 1: positions in the image of the last rebuild (tile.cu: k_tile_prep), (x, y)
    const double *prel_z;    //          and z; the runs of a tile are bulk-copied from these two arrays
    const TileChunk *chunks; // already offset to the first chunk of the launch
    const uint16_t *rows16;
    const uint32_t *cnt;
    uint32_t kmax;
    uint32_t cap;            // doubles per coordinate array in shared memory (> max tile size, sentinel included)
    PairConst P1;
    double *f;
    uint32_t npad;
    BoxDev box;
    int accumulate, store;
    double *partials;
    const int *abort_flag;
    uint32_t pers_blocks;    // > 0: persistent double-buffered kernel with this many blocks (bulk-copy staging only)
    uint32_t half_ok;        // rows are in the lane-vector layout of tile.cu (not bank-ordered): a pass whose longest row has
                             // <= TEAM * m entries left holds them all in the first m entries of every lane's vector
    uint32_t pf_dist;        // > 0: every block pulls the descriptor of chunk blockIdx.x + pf_dist into L2
    const uint32_t *chunk_s0; // first slot of every chunk of the launch and the end of the last one (persistent kernel)
};

// 1/x for normal positive x without the division's special-case path: MUFU.RCP64H seed (2^-23) and one
// cubic Newton step, relative error ~2^-52 (the pair distance squared is never zero, subnormal or infinite)
__device__ __forceinline__ double rcp_pos(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    return fma(y, fma(e, e, e), y);
}

// Branch-free form of pair_eval (pairs.cuh) for the Lennard-Jones family, so that the V pairs of a pass can
// be interleaved: LJRepulsePair::forces/energy :875-891 (:126-151), LJAttractRepulsePair :1271-1298,
// LennardJonesCutPair :253-267. c12 = 12 epsilon.
template <int KIND>
__device__ __forceinline__ void lj_eval(const PairConst &P, double c12, double dsq, bool want_e, double &scal, double &e) {
    const bool in = !(dsq > P.rc2);
    const double w = rcp_pos(dsq);
    const double s2 = P.sig2 * w;
    const double ir6 = s2 * s2 * s2;
    const double t = fma(ir6, ir6, -ir6) * (c12 * w); // 12 eps ir6 (ir6 - 1) / dsq in three operations
    scal = in ? t : 0.0;
    e = 0.0;
    if (want_e) {
        const double mid = 1.0 - ir6;
        const double en = KIND == PARM_PAIR_LJCUT ? P.eps * (mid * mid - 1.0) - P.cutE : P.eps * (mid * mid) - P.cutE;
        e = in ? en : 0.0;
    }
}

// One pass of a team's row: V 16-bit entries of this lane as V/2 words (one vector load)
template <int V>
struct RowWords {
    uint32_t w[V / 2];
};
template <int V>
__device__ __forceinline__ RowWords<V> load_row_words(const uint16_t *p) {
    RowWords<V> r;
    if constexpr (V == 8) {
        const uint4 q = __ldg(reinterpret_cast<const uint4 *>(p));
        r.w[0] = q.x; r.w[1] = q.y; r.w[2] = q.z; r.w[3] = q.w;
    } else {
        const uint2 q = __ldg(reinterpret_cast<const uint2 *>(p));
        r.w[0] = q.x; r.w[1] = q.y;
    }
    return r;
}

// V pair evaluations of one lane: entries [0, NV) of the lane's vector of this pass
template <int KIND, int MODE, int V, int NV, bool MI>
__device__ __forceinline__ void tile_pairs(const TileForceArgs &A, const RowWords<V> &q, const double2 *sxy, const double *sz, double xi,
                                           double yi, double zi, double c12, long long rc2_bits, double &fx, double &fy, double &fz,
                                           double (&acc)[NPART], uint32_t sent_eo) {
    constexpr bool want_obs = MODE != MODE_F;
#pragma unroll
    for (int e = 0; e < NV; e++) {
        // the entry is 8 * index: byte offset of z, half the byte offset of (x, y)
        const uint32_t eo = (e & 1) ? (q.w[e >> 1] >> 16) : (q.w[e >> 1] & 0xffffu);
        const double2 pxy = *reinterpret_cast<const double2 *>(reinterpret_cast<const char *>(sxy) + 2u * eo); // one LDS.128
        double dx = xi - pxy.x, dy = yi - pxy.y, dz = zi - *reinterpret_cast<const double *>(reinterpret_cast<const char *>(sz) + eo);
        bool pad = false;
        if (MI) {
            // the minimum image folds the far-away sentinel of the row pads back into the box (to distance ZERO when a box
            // edge is a power of two: 1e100 / L is exact): pads are recognised by their index here, not by their distance
            pad = eo >= sent_eo;
            dx = min_image_fast(dx, A.box.L[0], A.box.invL[0]);
            dy = min_image_fast(dy, A.box.L[1], A.box.invL[1]);
            dz = min_image_fast(dz, A.box.L[2], A.box.invL[2]);
            if (pad) dx = dy = dz = 1e100;
        }
        const double dsq = dx * dx + (dy * dy + dz * dz);
        if (!want_obs) {
            // forces only: 17 operations on the fp64 pipe per pair. The factor 12 epsilon is applied once per atom, the
            // cut-off test is an integer comparison of the bit patterns (both numbers are positive),
            // ir6 (ir6 - 1) / dsq comes out of one fused multiply-add
            const double w = rcp_pos(dsq);
            const double s2 = A.P1.sig2 * w;
            const double ir6 = (s2 * s2) * s2;
            const double b = ir6 * w;
            const double t = fma(b, ir6, -b);
            // (select, then accumulate: predicated DFMAs -- C or PTX -- come out of ptxas as DFMA + 6 FSEL)
            const double scal = __double_as_longlong(dsq) <= rc2_bits ? t : 0.0; // sentinel pads: dsq ~ 1e200, beyond any cutoff
            fx = fma(dx, scal, fx);
            fy = fma(dy, scal, fy);
            fz = fma(dz, scal, fz);
            continue;
        }
        double scal, en;
        lj_eval<KIND>(A.P1, c12, dsq, want_obs, scal, en); // sentinel pads: dsq ~ 1e200, beyond any cutoff
        const double gx = dx * scal, gy = dy * scal, gz = dz * scal;
        fx += gx;
        fy += gy;
        fz += gz;
        if (want_obs) {
            acc[0] += en;
            acc[1] += dx * gx + (dy * gy + dz * gz); // r.dot(f), :2241
            acc[2] += dx * gx; acc[3] += dx * gy; acc[4] += dx * gz; // stress += r * f^T, :2274
            acc[5] += dy * gx; acc[6] += dy * gy; acc[7] += dy * gz;
            acc[8] += dz * gx; acc[9] += dz * gy; acc[10] += dz * gz;
            acc[11] += (en != 0.0) ? 1.0 : 0.0; // contacts :2126-2137
            acc[12] += (en > 0.0) ? 1.0 : 0.0;  // overlaps :2140-2151
        }
    }
}

// The 8 teams of a warp walk their rows in lock step, one pass of TEAM * V entries at a time, for as many passes as the
// warp's LONGEST row has (a team that has run out of entries walks sentinel words: exact zeros). The trip count is
// therefore warp-uniform, and so is the decision that ends the row: when the longest row has at most TEAM * m entries
// left, all of them sit in the first m entries of every lane's vector (entry k of a pass belongs to lane k % TEAM,
// position k / TEAM: tile.cu) and the pass runs m = 2, 4 or 6 pair evaluations per lane instead of 8: the rows of a
// warp are padded to the next multiple of 8 entries of its longest row, not of 32.
// The row words of pass k+1 (or, in the warp's last pass, of the first pass of the team's NEXT atom) and the next atom's
// row length are requested before the arithmetic of pass k starts, so the DRAM latency of the streamed rows hides behind
// ~160 fp64 instructions. my0 / q: length and first pass of the team's first atom, loaded by the caller before the tile
// was staged. sentw: two sentinel entries (8 * ntile in both halves).
template <int KIND, int MODE, int TEAM, int V, bool MI, int NT = TILE_NT>
__device__ __forceinline__ void tile_rows(const TileForceArgs &A, uint32_t na, uint32_t s0, uint32_t own, const double2 *sxy,
                                          const double *sz, double (&acc)[NPART], uint32_t my0, RowWords<V> q, uint32_t sentw) {
    constexpr bool want_obs = MODE != MODE_F;
    constexpr uint32_t NTEAM = NT / TEAM;
    const uint32_t tl = threadIdx.x % TEAM;
    const double c12 = 12.0 * A.P1.eps;
    const long long rc2_bits = __double_as_longlong(A.P1.rc2);
    const uint32_t partial = A.half_ok ? 1u : 0u; // 0: every pass walks all V entries of the lane's vector
    const uint32_t sent_eo = sentw & 0xffffu;     // 8 * ntile: entries from here on are pads (also the class sentinels of bank-ordered rows)
    uint32_t my = my0;
    for (uint32_t a = threadIdx.x / TEAM; a - threadIdx.x / TEAM < na; a += NTEAM) { // every lane of a warp runs the same trips (shuffles below)
        const bool valid = a < na;
        const uint32_t s = s0 + (valid ? a : 0);
        const bool validn = a + NTEAM < na;
        const uint32_t sn = s0 + (validn ? a + NTEAM : 0);
        const uint32_t myn = validn ? min(A.cnt[sn], A.kmax) : 0; // consumed after this atom's passes
        const uint32_t ti = own + (valid ? a : 0); // staged copy of this atom: same min_image(x - origin) as its neighbours
        const double2 pixy = sxy[ti];
        const double xi = pixy.x, yi = pixy.y, zi = sz[ti];
        const uint16_t *row = A.rows16 + (size_t)s * A.kmax + tl * V;
        const uint16_t *rown = A.rows16 + (size_t)sn * A.kmax + tl * V;
        double fx = 0, fy = 0, fz = 0;
        const uint32_t mymax = __reduce_max_sync(0xffffffffu, my);
        if (my == 0) { // no row (or no atom) for this team: whatever was prefetched is not a row
#pragma unroll
            for (int e = 0; e < V / 2; e++) q.w[e] = sentw;
        }
        if (mymax == 0 && validn) q = load_row_words<V>(rown);
        for (uint32_t k0 = 0; k0 < mymax; k0 += TEAM * V) {
            RowWords<V> qn;
#pragma unroll
            for (int e = 0; e < V / 2; e++) qn.w[e] = sentw;
            if (k0 + TEAM * V < my) qn = load_row_words<V>(row + k0 + TEAM * V);
            else if (k0 + TEAM * V >= mymax && validn) qn = load_row_words<V>(rown); // rows are allocated to kmax: safe whatever the next length is
            // entries per lane this pass needs: the longest row's remainder dealt over the TEAM lanes, in steps of two
            const uint32_t rem = mymax - k0;
            if (V == 8 && rem <= 2 * TEAM * partial) tile_pairs<KIND, MODE, V, 2, MI>(A, q, sxy, sz, xi, yi, zi, c12, rc2_bits, fx, fy, fz, acc, sent_eo);
            else if (rem <= V / 2 * TEAM * partial) tile_pairs<KIND, MODE, V, V / 2, MI>(A, q, sxy, sz, xi, yi, zi, c12, rc2_bits, fx, fy, fz, acc, sent_eo);
            else if (V == 8 && rem <= 6 * TEAM * partial) tile_pairs<KIND, MODE, V, 6, MI>(A, q, sxy, sz, xi, yi, zi, c12, rc2_bits, fx, fy, fz, acc, sent_eo);
            else tile_pairs<KIND, MODE, V, V, MI>(A, q, sxy, sz, xi, yi, zi, c12, rc2_bits, fx, fy, fz, acc, sent_eo);
            q = qn;
        }
        my = myn;
        if (!want_obs) {
            fx *= c12;
            fy *= c12;
            fz *= c12;
        }
        if (MODE == MODE_F || A.store) {
#pragma unroll
            for (int o = TEAM / 2; o; o >>= 1) {
                fx += __shfl_xor_sync(0xffffffffu, fx, o);
                fy += __shfl_xor_sync(0xffffffffu, fy, o);
                fz += __shfl_xor_sync(0xffffffffu, fz, o);
            }
            if (valid && tl == 0) {
                double *f = A.f;
                if (A.accumulate) {
                    f[s] += fx;
                    f[A.npad + s] += fy;
                    f[2 * (size_t)A.npad + s] += fz;
                } else {
                    f[s] = fx;
                    f[A.npad + s] = fy;
                    f[2 * (size_t)A.npad + s] = fz;
                }
            }
        }
    }
}

// ---- bulk-copy staging (STAGE 1): mbarrier + cp.async.bulk (the TMA unit's linear copy), sm_90+ PTX ----------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar) : "memory");
}

// STAGE 0: every thread loads positions from pos, reduces them to min_image(x - origin) and stores them (any box, any
// team layout). STAGE 1: the <= 18 runs of the tile arrive as <= 36 cp.async.bulk copies from prel (tile.cu) issued by
// the lanes of warp 0 and counted by one mbarrier; no thread touches the positions, except in the chunks next to a
// periodic face, whose wrapped runs get their image shift (TileChunk::sh) added once they have landed.
// The time between a block's first instruction and the arrival of its tile is dead time for its 8 warps (ncu, round 2:
// a third of the kernel's warp time), so the chain is kept short: the chunk descriptor is read ONCE, as 64 coalesced
// words that go to shared memory -- run table, origin, flags and image shifts all come from there (the shifts used to
// be read from global memory run by run, after the tile had landed) --, the read is issued before the abort word is
// looked at, and every block pulls the descriptor of the chunk a wave of blocks later into L2.
static_assert(sizeof(TileChunk) % 4 == 0 && sizeof(TileChunk) / 4 <= TILE_NT, "chunk descriptor is read as one word per thread");
template <int KIND, int MODE, int TEAM, int V, int STAGE>
__global__ void __launch_bounds__(TILE_NT) k_force_tile(const TileForceArgs A) {
    constexpr uint32_t HW = sizeof(TileChunk) / 4;
    extern __shared__ __align__(16) double s_xyz[];
    __shared__ __align__(16) uint32_t s_hdr[HW];
    __shared__ __align__(8) unsigned long long s_bar;
    const TileChunk *C = A.chunks + blockIdx.x;
    uint32_t hw = 0;
    if (threadIdx.x < HW) hw = __ldg(reinterpret_cast<const uint32_t *>(C) + threadIdx.x);
    const uint32_t s0 = __ldg(&C->s0), na = __ldg(&C->n), ntile = __ldg(&C->ntile); // (the same line: what the row prefetch below needs)
    if (A.abort_flag && *A.abort_flag) return; // speculatively enqueued step whose predecessor asked for a rebuild
    if (threadIdx.x < HW) s_hdr[threadIdx.x] = hw;
    if (A.pf_dist && threadIdx.x < (sizeof(TileChunk) + 127) / 128 && blockIdx.x + A.pf_dist < gridDim.x)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(C + A.pf_dist) + 128 * threadIdx.x));
    const TileChunk *H = reinterpret_cast<const TileChunk *>(s_hdr); // valid after the first block barrier
    double2 *sxy = reinterpret_cast<double2 *>(s_xyz); // (x, y) pairs, then the z array
    double *sz = s_xyz + 2 * (size_t)A.cap;
    // first atom of this team: row length and first pass, in flight while the tile is staged
    uint32_t my0 = 0;
    RowWords<V> q0;
    const uint32_t sentw = (ntile << TILE_IDX_SHIFT) | (ntile << (TILE_IDX_SHIFT + 16));
#pragma unroll
    for (int e = 0; e < V / 2; e++) q0.w[e] = sentw;
    {
        const uint32_t a = threadIdx.x / TEAM;
        if (a < na) {
            const uint32_t s = s0 + a;
            my0 = min(A.cnt[s], A.kmax);
            q0 = load_row_words<V>(A.rows16 + (size_t)s * A.kmax + (threadIdx.x % TEAM) * V);
        }
    }
    // the sentinel every row is padded with (index ntile) and the 16 class sentinels of bank-ordered rows
    // (S .. S + 15, bank_order.cuh): all staged far away
    const uint32_t send = ((ntile + 1u + 15u) & ~15u) + 16u;
    for (uint32_t t = ntile + threadIdx.x; t < send; t += TILE_NT) {
        sxy[t] = make_double2(1e100, 1e100);
        sz[t] = 1e100;
    }
    uint32_t cflags;
    if (STAGE == 1) {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
        if (threadIdx.x == 0) mbar_init(bar, 1);
        __syncthreads();
        cflags = H->flags;
        if (threadIdx.x < 32) {
            if (threadIdx.x == 0) mbar_arrive_expect_tx(bar, ntile * 24u);
            __syncwarp();
            if (threadIdx.x < TILE_MAXSEG) {
                const uint32_t o0 = H->seg_off[threadIdx.x], len = H->seg_off[threadIdx.x + 1] - o0, j0 = H->seg_start[threadIdx.x];
                if (len) { // even start, even length: both copies are 16-byte aligned at both ends
                    bulk_g2s((uint32_t)__cvta_generic_to_shared(sxy + o0), A.prel_xy + j0, len * 16u, bar);
                    bulk_g2s((uint32_t)__cvta_generic_to_shared(sz + o0), A.prel_z + j0, len * 8u, bar);
                }
            }
        }
        mbar_wait(bar, 0);
        if (cflags & 2u) { // runs reached across a periodic face: add their image shift
            for (uint32_t seg = 0; seg < TILE_MAXSEG; seg++) {
                const int sx = H->sh[seg][0], sy = H->sh[seg][1], szz = H->sh[seg][2];
                if (!(sx | sy | szz)) continue;
                const double ax = sx * A.box.L[0], ay = sy * A.box.L[1], az = szz * A.box.L[2];
                for (uint32_t t = H->seg_off[seg] + threadIdx.x; t < H->seg_off[seg + 1]; t += TILE_NT) {
                    double2 p = sxy[t];
                    p.x += ax;
                    p.y += ay;
                    sxy[t] = p;
                    sz[t] += az;
                }
            }
            __syncthreads();
        }
    } else {
        __syncthreads();
        cflags = H->flags;
        const double ox = H->o[0], oy = H->o[1], oz = H->o[2];
        // four positions per thread in flight: the loads of a group are issued before the first is reduced and stored
        // (a run-major variant -- thread t takes element t of every run, no table search -- was slower: 0.292 vs 0.278 ms)
        constexpr int SU = 4;
        uint32_t seg = 0;
        for (uint32_t t0 = threadIdx.x; t0 < ntile; t0 += SU * TILE_NT) {
            double4 p[SU];
#pragma unroll
            for (int u = 0; u < SU; u++) {
                const uint32_t t = t0 + u * TILE_NT;
                if (t < ntile) {
                    while (t >= H->seg_off[seg + 1]) seg++;
                    p[u] = A.pos[H->seg_start[seg] + (t - H->seg_off[seg])];
                }
            }
#pragma unroll
            for (int u = 0; u < SU; u++) {
                const uint32_t t = t0 + u * TILE_NT;
                if (t < ntile) {
                    sxy[t] = make_double2(min_image_fast(p[u].x - ox, A.box.L[0], A.box.invL[0]),
                                          min_image_fast(p[u].y - oy, A.box.L[1], A.box.invL[1]));
                    sz[t] = min_image_fast(p[u].z - oz, A.box.L[2], A.box.invL[2]);
                }
            }
        }
        __syncthreads();
    }
    double acc[NPART];
    if (MODE != MODE_F)
#pragma unroll
        for (int q = 0; q < NPART; q++) acc[q] = 0.0;
    // the chunk's own atoms are consecutive entries of the centre column's run (piece 0 or 1 of stencil column 4)
    const uint32_t st8 = H->seg_start[8], d8 = s0 - st8;
    const uint32_t own = (s0 >= st8 && d8 < H->seg_off[9] - H->seg_off[8]) ? H->seg_off[8] + d8 : H->seg_off[9] + (s0 - H->seg_start[9]);
    if (cflags & 1u) tile_rows<KIND, MODE, TEAM, V, true>(A, na, s0, own, sxy, sz, acc, my0, q0, sentw);
    else tile_rows<KIND, MODE, TEAM, V, false>(A, na, s0, own, sxy, sz, acc, my0, q0, sentw);
    if (MODE != MODE_F) {
        __shared__ double red[NPART][TILE_NT / 32];
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
            for (int ww = 0; ww < TILE_NT / 32; ww++) x += red[threadIdx.x][ww];
            A.partials[(size_t)blockIdx.x * NPART + threadIdx.x] = x;
        }
    }
}

// ---- persistent, double-buffered, warp-specialised variant (forces only, TEAM 4, V 8, bulk-copy staging) ------------
// ncu on the one-block-per-chunk kernel: a quarter of the warp time goes into waiting for the chunk's tile (the block
// reads its run table, then issues the copies, then waits for them) and every block pays its own tail. Here 2 blocks of
// 512 threads stay resident per SM and walk the chunks blockIdx.x, blockIdx.x + gridDim.x, ...; each owns TWO tile
// buffers. Warp 0 is the PRODUCER: it waits until the 15 compute warps have released a buffer (mbarrier `empty`), reads
// the next chunk's run table, issues its bulk copies (mbarrier `full`) and, for the chunks next to a periodic face,
// adds the image shifts once the copies have landed (mbarrier `ready`). The 15 COMPUTE warps (120 teams: chunks hold at
// most TILE_PCH atoms) never meet at a block barrier: each walks the chunks at its own pace, at most one chunk ahead of
// the slowest, with the row length and first pass of its team's next atom requested one chunk ahead and the first slot
// of the chunk after that two chunks ahead. Same rows, same summation order: bit-identical to k_force_tile.
#define TILE_PNT 512
#define TILE_PCH 120 // (TILE_PNT / 32 - 1) compute warps x 8 teams
struct TileMeta { // per buffer, written by the producer before it arms `full`
    uint32_t ntile, flags, own, pad0;
    uint32_t off[TILE_MAXSEG + 1];
    int8_t sh[TILE_MAXSEG][3];
};
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <int KIND>
__global__ void __launch_bounds__(TILE_PNT, 2) k_force_tile_pers(const TileForceArgs A, const uint32_t nchunks) {
    if (A.abort_flag && *A.abort_flag) return; // speculatively enqueued step whose predecessor asked for a rebuild
    constexpr int TEAM = 4, V = 8, NCW = TILE_PNT / 32 - 1;
    extern __shared__ __align__(16) double s_xyz[]; // two buffers of 3 * cap doubles
    __shared__ TileMeta s_meta[2];
    __shared__ __align__(8) unsigned long long s_full[2], s_empty[2], s_ready[2];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t bufdoubles = 3 * A.cap, G = gridDim.x;
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; b++) {
            mbar_init((uint32_t)__cvta_generic_to_shared(&s_full[b]), 1);
            mbar_init((uint32_t)__cvta_generic_to_shared(&s_empty[b]), NCW);
            mbar_init((uint32_t)__cvta_generic_to_shared(&s_ready[b]), 1);
        }
    }
    __syncthreads();
    if (blockIdx.x >= nchunks) return;
    if (warp == 0) {
        // ---------------- producer ----------------
        // run table of the next chunk in registers one chunk ahead (lane t: run t)
        const TileChunk *C = A.chunks + blockIdx.x;
        uint32_t st = 0, o0 = 0, o1 = 0, sh = 0;
        auto load_hdr = [&](const TileChunk *Cx) {
            st = o0 = o1 = sh = 0;
            if (lane < TILE_MAXSEG) {
                st = Cx->seg_start[lane];
                o0 = Cx->seg_off[lane];
                o1 = Cx->seg_off[lane + 1];
                sh = (uint32_t)(uint8_t)Cx->sh[lane][0] | (uint32_t)(uint8_t)Cx->sh[lane][1] << 8 | (uint32_t)(uint8_t)Cx->sh[lane][2] << 16;
            } else if (lane == 18) {
                st = Cx->ntile;
                o0 = Cx->s0;
                o1 = Cx->flags;
            }
        };
        load_hdr(C);
        uint32_t k = 0;
        for (uint32_t c = blockIdx.x; c < nchunks; c += G, k++) {
            const uint32_t b = k & 1u;
            double2 *sxy = reinterpret_cast<double2 *>(s_xyz + (size_t)b * bufdoubles);
            double *sz = s_xyz + (size_t)b * bufdoubles + 2 * (size_t)A.cap;
            const uint32_t full = (uint32_t)__cvta_generic_to_shared(&s_full[b]);
            const uint32_t cst = st, co0 = o0, co1 = o1, csh = sh;
            if (c + G < nchunks) load_hdr(A.chunks + c + G); // consumed in the next iteration
            if (k >= 2) mbar_wait((uint32_t)__cvta_generic_to_shared(&s_empty[b]), ((k >> 1) - 1u) & 1u); // chunk k - 2 released it
            const uint32_t ntile = __shfl_sync(0xffffffffu, cst, 18), s0 = __shfl_sync(0xffffffffu, co0, 18);
            const uint32_t flags = __shfl_sync(0xffffffffu, co1, 18);
            TileMeta &M = s_meta[b];
            if (lane < TILE_MAXSEG) {
                M.off[lane] = co0;
                M.sh[lane][0] = (int8_t)(csh & 0xffu);
                M.sh[lane][1] = (int8_t)((csh >> 8) & 0xffu);
                M.sh[lane][2] = (int8_t)((csh >> 16) & 0xffu);
            }
            // the chunk's own atoms are consecutive entries of the centre column's run (piece 0 or 1 of stencil column 4)
            const uint32_t st8 = __shfl_sync(0xffffffffu, cst, 8), st9 = __shfl_sync(0xffffffffu, cst, 9);
            const uint32_t of8 = __shfl_sync(0xffffffffu, co0, 8), of9 = __shfl_sync(0xffffffffu, co0, 9);
            if (lane == 0) {
                M.off[TILE_MAXSEG] = ntile;
                M.ntile = ntile;
                M.flags = flags;
                const uint32_t d8 = s0 - st8;
                M.own = (s0 >= st8 && d8 < of9 - of8) ? of8 + d8 : of9 + (s0 - st9);
            }
            // the sentinel every row is padded with (index ntile) and the 16 class sentinels of bank-ordered rows
            const uint32_t send = ((ntile + 1u + 15u) & ~15u) + 16u;
            for (uint32_t t = ntile + lane; t < send; t += 32) {
                sxy[t] = make_double2(1e100, 1e100);
                sz[t] = 1e100;
            }
            __syncwarp(); // the lanes' writes are ordered before lane 0's (releasing) arrive
            if (lane == 0) mbar_arrive_expect_tx(full, ntile * 24u);
            __syncwarp();
            if (lane < TILE_MAXSEG && co1 > co0) { // even start, even length: both copies are 16-byte aligned at both ends
                bulk_g2s((uint32_t)__cvta_generic_to_shared(sxy + co0), A.prel_xy + cst, (co1 - co0) * 16u, full);
                bulk_g2s((uint32_t)__cvta_generic_to_shared(sz + co0), A.prel_z + cst, (co1 - co0) * 8u, full);
            }
            if (flags & 2u) { // runs reached across a periodic face: add their image shift once they have landed
                mbar_wait(full, (k >> 1) & 1u);
                for (uint32_t seg = 0; seg < TILE_MAXSEG; seg++) {
                    const uint32_t sg = __shfl_sync(0xffffffffu, csh, seg);
                    if (!sg) continue;
                    const uint32_t t0 = __shfl_sync(0xffffffffu, co0, seg), t1 = __shfl_sync(0xffffffffu, co1, seg);
                    const double ax = (int8_t)(sg & 0xffu) * A.box.L[0], ay = (int8_t)((sg >> 8) & 0xffu) * A.box.L[1];
                    const double az = (int8_t)((sg >> 16) & 0xffu) * A.box.L[2];
                    for (uint32_t t = t0 + lane; t < t1; t += 32) {
                        double2 p = sxy[t];
                        p.x += ax;
                        p.y += ay;
                        sxy[t] = p;
                        sz[t] += az;
                    }
                }
                __syncwarp();
            }
            // `ready` completes one phase per chunk like `full` (the compute warps only look at it for shifted chunks)
            if (lane == 0) mbar_arrive((uint32_t)__cvta_generic_to_shared(&s_ready[b]));
        }
        return;
    }
    // ---------------- compute warps ----------------
    const uint32_t team = (threadIdx.x - 32) / TEAM, tl = threadIdx.x % TEAM;
    auto rows_of = [&](uint32_t s0x, uint32_t nax, uint32_t &my, RowWords<V> &q) {
        my = 0;
#pragma unroll
        for (int e = 0; e < V / 2; e++) q.w[e] = 0;
        if (team < nax) {
            const uint32_t s = s0x + team;
            my = min(A.cnt[s], A.kmax);
            q = load_row_words<V>(A.rows16 + (size_t)s * A.kmax + tl * V);
        }
    };
    uint32_t c = blockIdx.x;
    uint32_t s0 = A.chunk_s0[c], na = A.chunk_s0[c + 1] - s0;
    uint32_t s0n = 0, nan = 0;
    if (c + G < nchunks) { s0n = A.chunk_s0[c + G]; nan = A.chunk_s0[c + G + 1] - s0n; }
    uint32_t my0;
    RowWords<V> q0;
    rows_of(s0, na, my0, q0);
    const double c12 = 12.0 * A.P1.eps;
    const long long rc2_bits = __double_as_longlong(A.P1.rc2);
    for (uint32_t k = 0; c < nchunks; k++, c += G) {
        const uint32_t b = k & 1u, par = (k >> 1) & 1u;
        const uint32_t cn = c + G;
        uint32_t my1;
        RowWords<V> q1;
        rows_of(s0n, nan, my1, q1); // next chunk: consumed at the bottom of this iteration
        uint32_t s0nn = 0, ennn = 0; // first slot and end of the chunk after the next: raw loads, subtracted at the bottom
        if (cn + G < nchunks) { s0nn = A.chunk_s0[cn + G]; ennn = A.chunk_s0[cn + G + 1]; }
        const double2 *sxy = reinterpret_cast<const double2 *>(s_xyz + (size_t)b * bufdoubles);
        const double *sz = s_xyz + (size_t)b * bufdoubles + 2 * (size_t)A.cap;
        mbar_wait((uint32_t)__cvta_generic_to_shared(&s_full[b]), par);
        const TileMeta &M = s_meta[b];
        const uint32_t cflags = M.flags;
        if (cflags & 2u) mbar_wait((uint32_t)__cvta_generic_to_shared(&s_ready[b]), par);
        // one atom per team: the row walk of tile_rows without its next-atom bookkeeping
        const bool valid = team < na;
        const uint32_t s = s0 + (valid ? team : 0);
        const uint32_t ti = M.own + (valid ? team : 0);
        const double2 pixy = sxy[ti];
        const double xi = pixy.x, yi = pixy.y, zi = sz[ti];
        const uint16_t *row = A.rows16 + (size_t)s * A.kmax + tl * V;
        double fx = 0, fy = 0, fz = 0;
        RowWords<V> q = q0;
        const bool mi = (cflags & 1u) != 0;
        for (uint32_t k0 = 0; k0 < my0; k0 += TEAM * V) {
            RowWords<V> qn = q;
            if (k0 + TEAM * V < my0) qn = load_row_words<V>(row + k0 + TEAM * V);
#pragma unroll
            for (int e = 0; e < V; e++) {
                const uint32_t eo = (e & 1) ? (q.w[e >> 1] >> 16) : (q.w[e >> 1] & 0xffffu); // 8 * index
                const double2 pxy = *reinterpret_cast<const double2 *>(reinterpret_cast<const char *>(sxy) + 2u * eo);
                double dx = xi - pxy.x, dy = yi - pxy.y, dz = zi - *reinterpret_cast<const double *>(reinterpret_cast<const char *>(sz) + eo);
                if (mi) { // (uniform per chunk) tile wider than half the box: minimum image per pair as well
                    dx = min_image_fast(dx, A.box.L[0], A.box.invL[0]);
                    dy = min_image_fast(dy, A.box.L[1], A.box.invL[1]);
                    dz = min_image_fast(dz, A.box.L[2], A.box.invL[2]);
                    if (eo >= (M.ntile << TILE_IDX_SHIFT)) dx = dy = dz = 1e100; // pads: see tile_pairs
                }
                const double dsq = dx * dx + (dy * dy + dz * dz);
                const double w = rcp_pos(dsq);
                const double s2 = A.P1.sig2 * w;
                const double ir6 = (s2 * s2) * s2;
                const double bb = ir6 * w;
                const double t = fma(bb, ir6, -bb);
                const double scal = __double_as_longlong(dsq) <= rc2_bits ? t : 0.0; // sentinel pads: dsq ~ 1e200, beyond any cutoff
                fx = fma(dx, scal, fx);
                fy = fma(dy, scal, fy);
                fz = fma(dz, scal, fz);
            }
            q = qn;
        }
        fx *= c12;
        fy *= c12;
        fz *= c12;
#pragma unroll
        for (int o = TEAM / 2; o; o >>= 1) {
            fx += __shfl_xor_sync(0xffffffffu, fx, o);
            fy += __shfl_xor_sync(0xffffffffu, fy, o);
            fz += __shfl_xor_sync(0xffffffffu, fz, o);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive((uint32_t)__cvta_generic_to_shared(&s_empty[b])); // this warp is done with buffer b
        if (valid && tl == 0) {
            double *f = A.f;
            if (A.accumulate) {
                f[s] += fx;
                f[A.npad + s] += fy;
                f[2 * (size_t)A.npad + s] += fz;
            } else {
                f[s] = fx;
                f[A.npad + s] = fy;
                f[2 * (size_t)A.npad + s] = fz;
            }
        }
        s0 = s0n; na = nan; my0 = my1; q0 = q1;
        s0n = s0nn; nan = ennn - s0nn;
    }
}

template <int KIND, int TEAM, int V, int STAGE>
static cudaError_t launch_tile_mode(int mode, unsigned nchunks, size_t smem, cudaStream_t st, const TileForceArgs &A) {
    if (mode == MODE_F) {
        if (smem + 2048 > 48 * 1024) cudaFuncSetAttribute(k_force_tile<KIND, MODE_F, TEAM, V, STAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_force_tile<KIND, MODE_F, TEAM, V, STAGE><<<nchunks, TILE_NT, smem, st>>>(A);
    } else {
        if (smem + 2048 > 48 * 1024) cudaFuncSetAttribute(k_force_tile<KIND, MODE_FALL, TEAM, V, STAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_force_tile<KIND, MODE_FALL, TEAM, V, STAGE><<<nchunks, TILE_NT, smem, st>>>(A);
    }
    return cudaGetLastError();
}

// A.prel_xy != NULL selects bulk-copy staging (TEAM 4, V 8 only: tile.cu plans aligned runs for no other layout)
template <int KIND>
cudaError_t parm_launch_force_tile_kind(int team, int v, int mode, unsigned nchunks, size_t smem, cudaStream_t st, const TileForceArgs &A) {
    if (team == 8) return launch_tile_mode<KIND, 8, 4, 0>(mode, nchunks, smem, st, A);
    if (team == 2) return launch_tile_mode<KIND, 2, 8, 0>(mode, nchunks, smem, st, A);
    if (v == 4) return launch_tile_mode<KIND, 4, 4, 0>(mode, nchunks, smem, st, A);
    if (A.prel_xy && A.pers_blocks && mode == MODE_F) { // persistent kernel: 2 blocks of 512 threads per SM, two tile buffers each
        const size_t smem2 = 2 * smem;
        cudaFuncSetAttribute(k_force_tile_pers<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        k_force_tile_pers<KIND><<<std::min(nchunks, A.pers_blocks), TILE_PNT, smem2, st>>>(A, nchunks);
        return cudaGetLastError();
    }
    // (register budgets other than ptxas's own 64 were measured and lost: 80 / 114 registers (3 / 2 blocks per SM) 0.256 /
    // 0.321 ms against 0.243; 48 registers with 5 blocks of 91-atom chunks 0.250; 40 registers with 6 blocks 0.305)
    if (A.prel_xy) return launch_tile_mode<KIND, 4, 8, 1>(mode, nchunks, smem, st, A);
    return launch_tile_mode<KIND, 4, 8, 0>(mode, nchunks, smem, st, A);
}

#define PARM_INSTANTIATE_FORCE_TILE_KIND(K) \
    template cudaError_t parm_launch_force_tile_kind<K>(int, int, int, unsigned, size_t, cudaStream_t, const TileForceArgs &);
