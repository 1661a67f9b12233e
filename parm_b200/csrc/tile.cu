// Cell-tile staging for the NListed pair loop (interaction.hpp:2154-2291): per-rebuild planning.
//
// After the atoms are sorted by cell index (cells of one (x,y) column are contiguous in slot order), the
// neighbours of any run of consecutive slots inside one column lie in at most 18 contiguous slot runs:
// the 3 x 3 surrounding columns, each over the run's z range +- one cell (two runs where that range wraps).
// k_tile_cols cuts every owned column into chunks of <= ch slots (deterministic prefix scan),
// k_tile_chunks writes each chunk's run table, origin and size, and k_tile_localize rewrites the
// 32-bit neighbour rows as 16-bit indices into the chunk's tile (rows16). The pair kernel
// (force_tile.cuh) then stages the tile once per block in shared memory and never gathers positions
// from global memory or evaluates a per-pair minimum image.
#include <algorithm>
#include <stdlib.h>
#include <string.h>

#include "pairs.cuh"
#include "bank_order.cuh"

// ---- K5a: chunks per owned column, exclusive scan (one block; a few thousand columns at most) ----
__global__ void __launch_bounds__(1024)
k_tile_cols(const uint32_t *__restrict__ cell_start, uint32_t nc2, uint32_t ncol, uint32_t ch, uint32_t *col_slot,
            uint32_t *col_chunk, TileInfo *info) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_running;
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_running = 0;
    __syncthreads();
    for (uint32_t base = 0; base < ncol; base += 1024) {
        const uint32_t q = base + threadIdx.x;
        uint32_t a = 0, nch = 0;
        if (q < ncol) {
            a = cell_start[(size_t)q * nc2];
            const uint32_t b = cell_start[(size_t)(q + 1) * nc2];
            nch = (b - a + ch - 1) / ch;
        }
        uint32_t x = nch; // inclusive warp scan
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= (uint32_t)o) x += y;
        }
        if (lane == 31) s_warp[w] = x;
        __syncthreads();
        if (w == 0) {
            uint32_t t = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, t, o);
                if (lane >= (uint32_t)o) t += y;
            }
            s_warp[lane] = t; // inclusive over warps
        }
        __syncthreads();
        const uint32_t before = s_running + (w ? s_warp[w - 1] : 0) + (x - nch);
        if (q < ncol) {
            col_slot[q] = a;
            col_chunk[q] = before;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_running += s_warp[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        col_slot[ncol] = cell_start[(size_t)ncol * nc2];
        col_chunk[ncol] = s_running;
        info->nchunks = s_running;
    }
}

// ---- K5b: one thread per chunk: run table, origin, size -----------------------------------------
__global__ void __launch_bounds__(128)
k_tile_chunks(const uint32_t *__restrict__ cell_start, const uint32_t *__restrict__ cell_id_sorted,
              const uint32_t *__restrict__ col_slot, const uint32_t *__restrict__ col_chunk, uint32_t ncol, uint32_t ch,
              BoxDev box, GridDev g, StencilDev st, ShardDev sd, double margin, int aligned, TileChunk *chunks, uint32_t *s0arr,
              TileInfo *info) {
    const uint32_t cidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (cidx == col_chunk[ncol]) s0arr[cidx] = col_slot[ncol]; // end of the last chunk
    if (cidx >= col_chunk[ncol]) return;
    uint32_t lo = 0, hi = ncol; // largest q with col_chunk[q] <= cidx (columns without atoms own no chunk)
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (col_chunk[mid] <= cidx) lo = mid; else hi = mid;
    }
    const uint32_t q = lo;
    const uint32_t k = cidx - col_chunk[q];
    // the column's atoms are spread evenly over its chunks (a column of 907 atoms: 8 x 114 instead of 7 x 128 + 11)
    const uint32_t ccnt = col_slot[q + 1] - col_slot[q], cnch = col_chunk[q + 1] - col_chunk[q];
    const uint32_t csz = min(ch, (ccnt + cnch - 1) / cnch);
    const uint32_t a = col_slot[q] + k * csz, b = min(a + csz, col_slot[q + 1]);
    const int nc0 = g.nc[0], nc1 = g.nc[1], nc2 = g.nc[2];
    const int cx = (int)(q / (uint32_t)nc1), cy = (int)(q % (uint32_t)nc1);
    const int zlo = (int)(cell_id_sorted[a] % (uint32_t)nc2), zhi = (int)(cell_id_sorted[b - 1] % (uint32_t)nc2);
    TileChunk C;
    s0arr[cidx] = a;
    C.s0 = a;
    C.n = b - a;
    C.pad = q; // owned column index (the mask-mode localize pass derives the atoms' warp groups from it)
    C.pad2[0] = C.pad2[1] = 0;
    // z pieces (run axis): the chunk's cells +- one, split where the range wraps; the whole column once when
    // the range would overlap itself
    int za[2], zb[2], npiece = 1;
    bool wide = false;
    if (zhi - zlo + 3 > nc2) {
        za[0] = 0; zb[0] = nc2 - 1;
        wide = true;
    } else if (zlo - 1 < 0) {
        za[0] = zlo - 1 + nc2; zb[0] = nc2 - 1;
        za[1] = 0; zb[1] = zhi + 1;
        npiece = 2;
    } else if (zhi + 1 >= nc2) {
        za[0] = zlo - 1; zb[0] = nc2 - 1;
        za[1] = 0; zb[1] = zhi + 1 - nc2;
        npiece = 2;
    } else {
        za[0] = zlo - 1; zb[0] = zhi + 1;
    }
    // positional run table: run (ix * 3 + iy) * 2 + p is piece p of stencil column (cx - 1 + ix, cy - 1 + iy); runs
    // without atoms (and the second piece of an unwrapped range) have length zero
    uint32_t off = 0;
    const uint32_t nseg = TILE_MAXSEG;
    int sidx = 0;
    // origin: centre of the chunk's cells (wrapped frame); min_image(x - o) then picks the right image of every
    // staged atom whatever multiple of L its unwrapped coordinate carries
    double h[3];
    C.o[0] = st.open0 ? sd.lo + (cx + 0.5) * (sd.Ls / sd.nci) : (cx + 0.5) / g.scale[0];
    h[0] = st.open0 ? 0.5 * sd.Ls / sd.nci : 0.5 / g.scale[0];
    C.o[1] = (cy + 0.5) / g.scale[1];
    h[1] = 0.5 / g.scale[1];
    C.o[2] = 0.5 * (zlo + zhi + 1) / g.scale[2];
    h[2] = 0.5 * (zhi - zlo + 1) / g.scale[2];
    bool shifted = false;
    for (int xx = cx - 1; xx <= cx + 1; xx++) {
        int x2 = xx;
        if (st.open0) {
            if (x2 < 0) x2 = nc0 - 1; // halo layer below (numbered last), halo above is layer nc0 - 2 = cx + 1
        } else if (x2 < 0) x2 += nc0;
        else if (x2 >= nc0) x2 -= nc0;
        for (int yy = cy - 1; yy <= cy + 1; yy++) {
            int y2 = yy;
            if (y2 < 0) y2 += nc1; else if (y2 >= nc1) y2 -= nc1;
            const uint32_t cbase = ((uint32_t)x2 * (uint32_t)nc1 + (uint32_t)y2) * (uint32_t)nc2;
            for (int p = 0; p < 2; p++, sidx++) {
                uint32_t jb = 0, je = 0;
                if (p < npiece) {
                    jb = cell_start[cbase + (uint32_t)za[p]];
                    je = cell_start[cbase + (uint32_t)zb[p] + 1];
                }
                C.sh[sidx][0] = C.sh[sidx][1] = C.sh[sidx][2] = 0;
                if (aligned && je > jb) {
                    // bulk-copy staging: 16-byte aligned source and size for the 8-byte z array -> runs start and end on
                    // even slots (at most one stranger at either end, staged but never referenced by a row)
                    jb &= ~1u;
                    je = (je + 1u) & ~1u;
                    // image of the run next to the chunk, from the geometry alone: centre of the run's cells in the frame
                    // of prel (wrapped; slab-relative and continuous along the slab axis of a sharded context)
                    const double cxr = (x2 + 0.5) / g.scale[0], cyr = (y2 + 0.5) / g.scale[1];
                    const double czr = 0.5 * (za[p] + zb[p] + 1) / g.scale[2];
                    const int s0i = st.open0 ? 0 : -(int)rint((cxr - C.o[0]) * box.invL[0]);
                    const int s1i = -(int)rint((cyr - C.o[1]) * box.invL[1]);
                    const int s2i = -(int)rint((czr - C.o[2]) * box.invL[2]);
                    C.sh[sidx][0] = (int8_t)s0i; C.sh[sidx][1] = (int8_t)s1i; C.sh[sidx][2] = (int8_t)s2i;
                    if (s0i | s1i | s2i) shifted = true;
                }
                C.seg_start[sidx] = jb;
                C.seg_off[sidx] = off;
                off += je - jb;
            }
        }
    }
    C.seg_off[TILE_MAXSEG] = off;
    C.nseg = nseg;
    C.ntile = off;
    // staged differences equal the minimum image iff |x_i - o| + r_list stays below L/2 (margin = r_list + 2 skin); with
    // per-run image shifts every atom of a run (whole cells) must lie on the same side: one more cell instead
    for (int d = 0; d < 3; d++) {
        const double cell = d == 0 && st.open0 ? sd.Ls / sd.nci : 1.0 / g.scale[d];
        const double reach = aligned ? fmax(margin, cell * (1.0 + 1e-9)) : margin;
        if (!(h[d] + reach < 0.4999 * box.L[d])) wide = true;
    }
    C.flags = (wide ? 1u : 0u) | (shifted ? 2u : 0u);
    chunks[cidx] = C;
    atomicMax(&info->max_tile, off);
    if (wide) atomicAdd(&info->wide, 1u);
    if (off > 65534u) atomicOr(&info->bad, 1u);
}

// ---- K5c: rows16 ----------------------------------------------------------------------------------
// One block per chunk, one warp per atom (round robin). The build kernel left the stencil column (0..8) each
// neighbour was found in in the top bits of the 32-bit entry; the column's two runs in the chunk's positional
// table turn the slot into the tile-local index without a search. Entry k of the row goes to position
//   (k / (team v)) (team v) + (k % team) v + (k % (team v)) / team
// so that lane t of a team finds entries t, t + team, ... of every pass in ONE vector load while the team as
// a whole still reads consecutive entries (= near-consecutive tile slots, fewer bank conflicts: 0.308 vs
// 0.333 ms for the pair kernel at N = 1e6) at every step. The tail of the last pass is padded with the sentinel
// index ntile (staged far away).
template <int V>
__global__ void __launch_bounds__(TILE_NT)
k_tile_localize(const TileChunk *__restrict__ chunks, const uint32_t *__restrict__ nbr, const uint32_t *__restrict__ cnt,
                uint32_t kmax, int team, uint16_t *__restrict__ rows16, TileInfo *info) {
    __shared__ uint4 s_tab[9]; // per stencil column: first run (start, offset), second run (start, offset)
    __shared__ uint32_t s_ntile;
    const TileChunk *C = chunks + blockIdx.x;
    if (threadIdx.x < 9) {
        const int kc = 2 * threadIdx.x;
        s_tab[threadIdx.x] = make_uint4(C->seg_start[kc], C->seg_off[kc], C->seg_start[kc + 1], C->seg_off[kc + 1]);
    }
    if (threadIdx.x == 0) s_ntile = C->ntile;
    __syncthreads();
    const uint32_t ntile = s_ntile, s0 = C->s0, na = C->n;
    // the thread layout of the pair kernel: `team` lanes per atom, lane tl produces the vector of V entries it will read
    const uint32_t tl = threadIdx.x % (uint32_t)team, per_iter = TILE_NT / (uint32_t)team;
    const uint32_t tv = (uint32_t)team * V;
    bool lost = false;
    for (uint32_t a0 = 0; a0 < na; a0 += per_iter) {
        const uint32_t a = a0 + threadIdx.x / (uint32_t)team;
        if (a >= na) continue;
        const uint32_t s = s0 + a;
        const uint32_t my = min(cnt[s], kmax);
        const uint32_t *row = nbr + (size_t)s * kmax;
        uint16_t *out = rows16 + (size_t)s * kmax;
        for (uint32_t k0 = 0; k0 < my; k0 += tv) {
            uint32_t e[V], loc[V];
#pragma unroll
            for (int q = 0; q < V; q++) {
                const uint32_t k = k0 + (uint32_t)q * (uint32_t)team + tl;
                e[q] = k < my ? __ldg(row + k) : 0xffffffffu;
            }
#pragma unroll
            for (int q = 0; q < V; q++) {
                const uint4 t = s_tab[min(e[q] >> PARM_NBR_SLOT_BITS, 8u)];
                const uint32_t j = e[q] & PARM_NBR_SLOT_MASK;
                const uint32_t l = j - t.x < t.w - t.y ? t.y + (j - t.x) : t.w + (j - t.z);
                const bool pad = e[q] == 0xffffffffu;
                if (!pad && l >= ntile) lost = true; // (not expected: the entry is in neither run of its column)
                loc[q] = (pad || l >= ntile ? ntile : l) << TILE_IDX_SHIFT; // rows hold 8 * index: the pair kernel's byte offset of z
            }
            if (V == 8) {
                uint4 o;
                o.x = loc[0] | (loc[1] << 16); o.y = loc[2] | (loc[3] << 16);
                o.z = loc[4 % V] | (loc[5 % V] << 16); o.w = loc[6 % V] | (loc[7 % V] << 16);
                *reinterpret_cast<uint4 *>(out + k0 + tl * V) = o;
            } else {
                uint2 o;
                o.x = loc[0] | (loc[1] << 16); o.y = loc[2] | (loc[3] << 16);
                *reinterpret_cast<uint2 *>(out + k0 + tl * V) = o;
            }
        }
    }
    if (lost) atomicOr(&info->bad, 2u);
}

// ---- bank-aware row order (experimental, bank_order.cuh) ---------------------------------------------------------
// One block per chunk, the pair kernel's thread layout (4 lanes per atom). A team copies its row to shared memory,
// pre-fills the output with the class sentinels, runs the two lane phases of bank_order.cuh and writes its vectors back.
__global__ void __launch_bounds__(TILE_NT)
k_tile_bank_order(const TileChunk *__restrict__ chunks, const uint32_t *__restrict__ cnt, uint32_t kmax, uint16_t *__restrict__ rows16) {
    extern __shared__ __align__(16) uint16_t s_io[]; // [2][TILE_NT / 4][kmax]: rows in, rows out
    __shared__ uint32_t s_packed[TILE_NT / 4][4], s_nrest[TILE_NT / 4][4];
    __shared__ uint16_t s_rest[TILE_NT / 4][4 * BO_RCAP];
    __shared__ uint32_t s_ovf[TILE_NT / 4];
    const TileChunk *C = chunks + blockIdx.x;
    const uint32_t ntile = C->ntile, s0 = C->s0, na = C->n;
    const uint32_t S = (ntile + 1u + 15u) & ~15u; // first of the 16 class sentinels staged behind the tile
    const uint32_t tl = threadIdx.x & 3u, team = threadIdx.x >> 2;
    uint16_t *in = s_io + (size_t)team * kmax, *out = s_io + (size_t)(TILE_NT / 4 + team) * kmax;
    for (uint32_t a0 = 0; a0 < na; a0 += TILE_NT / 4) {
        const uint32_t a = a0 + team;
        const bool valid = a < na;
        const uint32_t s = s0 + (valid ? a : 0);
        const uint32_t my = valid ? min(cnt[s], kmax) : 0;
        const uint32_t mypad = (my + 31u) & ~31u, G = mypad >> 2, q = a & 3u;
        uint16_t *row = rows16 + (size_t)s * kmax;
        for (uint32_t k0 = 0; k0 < mypad; k0 += 32) {
            {   // rows hold 8 * index (TILE_IDX_SHIFT): the classes are those of the index
                const uint4 rv = *reinterpret_cast<const uint4 *>(row + k0 + tl * 8);
                *reinterpret_cast<uint4 *>(in + k0 + tl * 8) = make_uint4((rv.x >> TILE_IDX_SHIFT) & 0x1fff1fffu, (rv.y >> TILE_IDX_SHIFT) & 0x1fff1fffu,
                                                                          (rv.z >> TILE_IDX_SHIFT) & 0x1fff1fffu, (rv.w >> TILE_IDX_SHIFT) & 0x1fff1fffu);
            }
#pragma unroll
            for (uint32_t e = 0; e < 8; e++) out[k0 + tl * 8 + e] = bo_sentinel(S, (k0 >> 2) + e, q, tl);
        }
        if (tl == 0) s_ovf[team] = 0;
        __syncwarp();
        bool ovf = false;
        uint32_t nrest = 0;
        const uint32_t packed = bo_phase1(in, my, G, q, tl, out, &s_rest[team][tl * BO_RCAP], &nrest, &ovf);
        s_packed[team][tl] = packed;
        s_nrest[team][tl] = nrest;
        if (ovf) s_ovf[team] = 1;
        __syncwarp();
        const bool keep = s_ovf[team] != 0; // a lane ran out of surplus space: the row keeps its build order
        if (!keep) bo_phase2(G, q, tl, s_packed[team], s_rest[team], s_nrest[team], out);
        __syncwarp();
        if (valid && !keep)
            for (uint32_t k0 = 0; k0 < mypad; k0 += 32)
            {
                const uint4 ov = *reinterpret_cast<const uint4 *>(out + k0 + tl * 8);
                *reinterpret_cast<uint4 *>(row + k0 + tl * 8) = make_uint4(ov.x << TILE_IDX_SHIFT, ov.y << TILE_IDX_SHIFT, ov.z << TILE_IDX_SHIFT, ov.w << TILE_IDX_SHIFT);
            }
        __syncwarp();
    }
}

// ---- mask-mode lists: rows16 straight from the build's pass masks -------------------------------------------------
// Same thread layout as k_tile_localize<8> (TEAM = 4 lanes per atom, V = 8). Lane tl of a team takes a CONTIGUOUS quarter
// of the candidate blocks of the atom's warp group: one pass adds up the popcounts of its masks, one exclusive scan over
// the four lanes gives the lane its first row position (rows keep the block order of the classic build), a second pass
// turns every set bit into a tile index through the 9-entry column table. A lane never waits for its team mates between
// blocks, so a warp runs as long as its busiest lane (~40 entries), not as long as the sum of the per-round maxima.
// The row is assembled in shared memory in the permuted lane-vector layout of the pair kernel (row stride padded by 16
// bytes: the eight teams of a warp start in different banks) and leaves as 16-byte stores, padded to whole passes with
// the sentinel index ntile.
__global__ void __launch_bounds__(TILE_NT)
k_tile_localize_masks(const TileChunk *__restrict__ chunks, MaskOut mo, const uint32_t *__restrict__ cell_id_sorted, uint32_t nc2,
                      uint32_t gpc, uint32_t zg, const uint32_t *__restrict__ cnt, uint32_t kmax, uint16_t *__restrict__ rows16,
                      TileInfo *info) {
    extern __shared__ __align__(16) uint16_t s_rows[]; // [TILE_NT / 4][kmax + 8]
    __shared__ uint4 s_tab[9];
    __shared__ uint32_t s_ntile;
    const TileChunk *C = chunks + blockIdx.x;
    if (threadIdx.x < 9) {
        const int kc = 2 * threadIdx.x;
        s_tab[threadIdx.x] = make_uint4(C->seg_start[kc], C->seg_off[kc], C->seg_start[kc + 1], C->seg_off[kc + 1]);
    }
    if (threadIdx.x == 0) s_ntile = C->ntile;
    __syncthreads();
    const uint32_t ntile = s_ntile, s0 = C->s0, na = C->n;
    const uint32_t tl = threadIdx.x & 3u, team = threadIdx.x >> 2;
    const uint32_t stride = kmax + 8u;
    uint16_t *buf = s_rows + (size_t)team * stride;
    bool lost = false;
    for (uint32_t a0 = 0; a0 < na; a0 += TILE_NT / 4) { // every lane of a warp runs the same trips (team shuffles below)
        const uint32_t a = a0 + team;
        const bool valid = a < na;
        const uint32_t s = s0 + (valid ? a : 0);
        const uint32_t my = valid ? min(cnt[s], kmax) : 0;
        const uint32_t mypad = (my + 31u) & ~31u;
        for (uint32_t k = tl * 8u; k < mypad; k += 32u) // sentinel everywhere, entries overwrite it
            *reinterpret_cast<uint4 *>(buf + k) = make_uint4((ntile << TILE_IDX_SHIFT) * 0x10001u, (ntile << TILE_IDX_SHIFT) * 0x10001u,
                                                             (ntile << TILE_IDX_SHIFT) * 0x10001u, (ntile << TILE_IDX_SHIFT) * 0x10001u);
        uint32_t nb = 0, grp = 0;
        if (my) {
            const uint32_t cid = cell_id_sorted[s];
            grp = (cid / nc2) * gpc + (cid % nc2) / zg;
            nb = min(mo.grp_nb[grp], mo.mb_cap);
        }
        const uint32_t nbq = (nb + 3u) >> 2, b0 = tl * nbq, b1 = min(b0 + nbq, nb);
        const uint32_t *mp = mo.masks + (size_t)b0 * mo.npad + s;
        uint32_t tot = 0;
        for (uint32_t b = b0; b < b1; b++, mp += mo.npad) tot += __popc(__ldg(mp));
        uint32_t incl = tot; // inclusive scan over the 4 lanes of the team
        uint32_t y = __shfl_up_sync(0xffffffffu, incl, 1, 4);
        if (tl >= 1) incl += y;
        y = __shfl_up_sync(0xffffffffu, incl, 2, 4);
        if (tl >= 2) incl += y;
        uint32_t k = incl - tot;
        __syncwarp(); // the sentinel fill of the whole row is complete
        mp = mo.masks + (size_t)b0 * mo.npad + s;
        const uint32_t *bp = mo.blk_base + (size_t)grp * mo.mb_cap + b0;
        for (uint32_t b = b0; b < b1; b++, mp += mo.npad, bp++) {
            uint32_t m = __ldg(mp);
            if (!m) continue;
            const uint32_t base = __ldg(bp);
            const uint4 t = s_tab[min(base >> PARM_NBR_SLOT_BITS, 8u)];
            const uint32_t jb = base & PARM_NBR_SLOT_MASK;
            // candidate slot jb + bit -> tile index: both runs of the column in one unsigned comparison each
            const uint32_t l0 = jb - t.x, n0 = t.w - t.y;
            while (m) {
                const uint32_t bit = (uint32_t)__ffs(m) - 1u;
                m &= m - 1u;
                const uint32_t r = l0 + bit;
                const uint32_t l = r < n0 ? t.y + r : t.w + (jb + bit - t.z);
                if (l >= ntile) lost = true; // (not expected: the entry is in neither run of its column)
                if (k < my) {
                    // entry k of the row is read by lane (k & 3) of the team at step (k & 31) >> 2 of pass k / 32
                    const uint32_t r32 = k & 31u;
                    buf[(k & ~31u) + (r32 & 3u) * 8u + (r32 >> 2)] = (uint16_t)((l >= ntile ? ntile : l) << TILE_IDX_SHIFT);
                }
                k++;
            }
        }
        __syncwarp();
        if (valid) {
            uint16_t *out = rows16 + (size_t)s * kmax;
            for (uint32_t k0 = 0; k0 < mypad; k0 += 32)
                *reinterpret_cast<uint4 *>(out + k0 + tl * 8) = *reinterpret_cast<const uint4 *>(buf + k0 + tl * 8);
        }
        __syncwarp();
    }
    if (lost) atomicOr(&info->bad, 2u);
}

// Fast path of the above for groups of at most NB = 48 candidate blocks (nbmax of the build). The four lanes of a team
// load the masks of contiguous quarters of the candidate blocks; the NON-EMPTY ones are compacted into team-shared
// memory together with the exclusive prefix of their popcounts and a ready-made descriptor (what turns a bit of the
// block into a tile index). Then the ROW ENTRIES, not the blocks, are dealt out: lane tl produces entries
// [tl T/4, (tl+1) T/4) of the T-entry row, starting in the middle of whatever block holds its first entry, in ONE flat
// loop over set bits in which moving to the next block costs two shared-memory loads. Every lane of the warp runs the
// same ~28 iterations (the populous blocks of the centre columns no longer make one lane the straggler), and the row
// keeps the block order of the classic build, so the pair kernel's sums are bit-identical to it.
// two shapes: 64 teams per block with tables of 32 blocks, or 32 teams per block with tables of 48 (static shared memory)
template <int LM_NT, int LM_NB>
__global__ void __launch_bounds__(LM_NT)
k_tile_localize_masks_flat(const TileChunk *__restrict__ chunks, MaskOut mo, const uint32_t *__restrict__ cell_id_sorted, uint32_t nc2,
                           uint32_t gpc, uint32_t zg, const uint32_t *__restrict__ cnt, uint32_t kmax, uint16_t *__restrict__ rows16,
                           TileInfo *info) {
    constexpr uint32_t NB = LM_NB, MAXQ = LM_NB / 4, NTEAMS = LM_NT / 4;
    extern __shared__ __align__(16) uint16_t s_rows[]; // [LM_NT / 4][kmax + 8]
    __shared__ uint4 s_tab[9];
    __shared__ uint32_t s_ntile;
    __shared__ uint4 s_desc[NTEAMS][NB + 1];     // per non-empty block: (l0, n0, ty, tw + jz), see below
    __shared__ uint32_t s_mk[NTEAMS][NB + 1];    // (+1: the teams of a warp start in different banks)
    __shared__ uint16_t s_pre[NTEAMS][NB + 2];
    const TileChunk *C = chunks + blockIdx.x;
    if (threadIdx.x < 9) {
        const int kc = 2 * threadIdx.x;
        s_tab[threadIdx.x] = make_uint4(C->seg_start[kc], C->seg_off[kc], C->seg_start[kc + 1], C->seg_off[kc + 1]);
    }
    if (threadIdx.x == 0) s_ntile = C->ntile;
    __syncthreads();
    const uint32_t ntile = s_ntile, s0 = C->s0, na = C->n, col = C->pad, colcell0 = C->pad * nc2;
    const uint32_t tl = threadIdx.x & 3u, team = threadIdx.x >> 2;
    const uint32_t stride = kmax + 8u;
    uint16_t *buf = s_rows + (size_t)team * stride;
    uint32_t lmax = 0;
    for (uint32_t a0 = 0; a0 < na; a0 += LM_NT / 4) { // every lane of a warp runs the same trips (team shuffles below)
        const uint32_t a = a0 + team;
        const bool valid = a < na;
        const uint32_t s = s0 + (valid ? a : 0);
        const uint32_t my = valid ? min(cnt[s], kmax) : 0;
        const uint32_t mypad = (my + 31u) & ~31u;
        // (the row is assembled in NATURAL order; the lane-vector permutation of the pair kernel happens on the way out)
        uint32_t nb = 0, grp = 0;
        if (my) {
            const uint32_t cz = cell_id_sorted[s] - colcell0; // cell along the column (every atom of a chunk is in column `col`)
            grp = col * gpc + (zg == 1u ? cz : cz / zg);
            nb = min(min(mo.grp_nb[grp], mo.mb_cap), NB);
        }
        const uint32_t nbq = (nb + 3u) >> 2, b0 = tl * nbq, b1 = min(b0 + nbq, nb);
        uint32_t tot = 0, nne = 0; // entries / non-empty blocks of this lane's quarter
        uint32_t mk[MAXQ], bs[MAXQ];
#pragma unroll
        for (int q = 0; q < (int)MAXQ; q++) { // all loads in flight before the first use
            const uint32_t bq = b0 + (uint32_t)q;
            mk[q] = bq < b1 ? __ldg(mo.masks + (size_t)bq * mo.npad + s) : 0u;
            bs[q] = bq < b1 ? __ldg(mo.blk_base + (size_t)grp * mo.mb_cap + bq) : 0u;
            tot += __popc(mk[q]);
            nne += mk[q] ? 1u : 0u;
        }
        // inclusive scans over the 4 lanes of the team: entries and non-empty blocks (packed: both fit 16 bits)
        uint32_t pk = tot | nne << 16, incl = pk;
        uint32_t y = __shfl_up_sync(0xffffffffu, incl, 1, 4);
        if (tl >= 1) incl += y;
        y = __shfl_up_sync(0xffffffffu, incl, 2, 4);
        if (tl >= 2) incl += y;
        const uint32_t all = __shfl_sync(0xffffffffu, incl, 3, 4);
        uint32_t run = (incl - pk) & 0xffffu, slot = (incl - pk) >> 16;
        const uint32_t nbc = all >> 16; // non-empty blocks of the row
#pragma unroll
        for (int q = 0; q < (int)MAXQ; q++)
            if (mk[q]) {
                // candidate slot jb + bit -> tile index: r = l0 + bit < n0 ? ty + r : (tw + jz) + bit  (the two runs of the column)
                const uint4 t = s_tab[min(bs[q] >> PARM_NBR_SLOT_BITS, 8u)];
                const uint32_t jb = bs[q] & PARM_NBR_SLOT_MASK;
                s_desc[team][slot] = make_uint4(jb - t.x, t.w - t.y, t.y, t.w + (jb - t.z));
                s_mk[team][slot] = mk[q];
                s_pre[team][slot] = (uint16_t)run;
                run += __popc(mk[q]);
                slot++;
            }
        __syncwarp();
        // this lane's share of the entries and the block its first entry lies in
        const uint32_t T = min(all & 0xffffu, my);
        const uint32_t k0 = (T * tl) >> 2, k1 = (T * (tl + 1u)) >> 2;
        uint32_t k = k0, bq = 0, m = 0;
        uint4 d = make_uint4(0, 0, 0, 0);
        if (k1 > k0) {
            uint32_t lo = 0, hi = nbc; // last block whose prefix is <= k0: it holds entry k0
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (s_pre[team][mid] <= k0) lo = mid; else hi = mid;
            }
            bq = lo;
            m = s_mk[team][bq];
            d = s_desc[team][bq];
            for (uint32_t skip = k0 - s_pre[team][bq]; skip; skip--) m &= m - 1u; // its entries that belong to the lane before
        }
        while (k < k1) {
            if (!m) { // next block (all stored blocks are non-empty; there is one: k < k1 <= T)
                bq++;
                m = s_mk[team][bq];
                d = s_desc[team][bq];
            }
            const uint32_t bit = (uint32_t)__ffs(m) - 1u;
            m &= m - 1u;
            const uint32_t r = d.x + bit;
            const uint32_t l = r < d.y ? d.z + r : d.w + bit;
            lmax = max(lmax, l); // (an index beyond the tile is not expected: the entry would be in neither run of its column)
            buf[k++] = (uint16_t)(l << TILE_IDX_SHIFT); // rows hold 8 * index: the pair kernel's byte offset of z
        }
        __syncwarp();
        if (valid) {
            // lane tl's vector of pass p: entries 32 p + tl, + 4, ..., + 28 of the row (sentinel index past its end)
            uint16_t *out = rows16 + (size_t)s * kmax;
            for (uint32_t kk = 0; kk < mypad; kk += 32) {
                uint32_t w[4];
#pragma unroll
                for (int g = 0; g < 8; g += 2) {
                    const uint32_t ka = kk + tl + 4u * g, kb = ka + 4u;
                    const uint32_t ea = ka < T ? buf[ka] : ntile << TILE_IDX_SHIFT, eb = kb < T ? buf[kb] : ntile << TILE_IDX_SHIFT;
                    w[g >> 1] = ea | eb << 16;
                }
                *reinterpret_cast<uint4 *>(out + kk + tl * 8) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
        __syncwarp();
    }
    if (lmax >= ntile) atomicOr(&info->bad, 2u);
}

// ---- per step: positions in the image every atom had at the last rebuild (bulk-copy staging) ----------------------
// prel = x - img * L is what the pair kernel's cp.async.bulk copies bring into shared memory unchanged: OriginBox::diff
// (box.hpp:103) is resolved once per atom and step here instead of once per staged copy (17.7 per atom) there.
__global__ void __launch_bounds__(256)
k_tile_prep(const double4 *__restrict__ pos, const float4 *__restrict__ img, uint32_t first, uint32_t count, BoxDev box,
            double2 *__restrict__ prel_xy, double *__restrict__ prel_z, const int *abort_flag) {
    if (abort_flag && *abort_flag) return;
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < count; q += gridDim.x * blockDim.x) {
        const uint32_t s = first + q;
        const double4 p = pos[s];
        const float4 g = img[s];
        prel_xy[s] = make_double2(fma(-(double)g.x, box.L[0], p.x), fma(-(double)g.y, box.L[1], p.y));
        prel_z[s] = fma(-(double)g.z, box.L[2], p.z);
    }
}

int parm_tile_prep(parm_nlist *nl, uint32_t first, uint32_t count, cudaStream_t stream, const int *abort_flag) {
    parm_ctx *c = nl->ctx;
    const TileState &t = nl->tile;
    if (!t.valid || !t.stage_aligned || !count) return 0;
    const unsigned nb = std::min<unsigned>((count + 255) / 256, (unsigned)c->num_sms * 16u);
    k_tile_prep<<<nb, 256, 0, stream>>>(c->pos, t.img, first, count, c->box, t.prel_xy, t.prel_z, abort_flag);
    CK_LAUNCH(c);
    return 0;
}

// ---- host side -----------------------------------------------------------------------------------
static bool kind_on_tile(int kind) {
    const int kk = PARM_KERNEL_KIND(kind);
    return kk == PARM_PAIR_LJREPULSE || kk == PARM_PAIR_LJATTRACTREPULSE || kk == PARM_PAIR_LJCUT; // the Lennard-Jones family
}
static bool inter_fits(const parm_inter *it) {
    return it->have_params && !it->generic && it->nspecies == 1 && kind_on_tile(it->kind);
}

static void tile_config(parm_nlist *nl) {
    TileState &t = nl->tile;
    if (t.ch) return;
    const char *e = getenv("PARM_B200_TILE");
    t.enabled = e ? atoi(e) : 1;
    e = getenv("PARM_B200_TILE_MIN_NEIGHBORS");
    t.min_nbrs = e ? atoi(e) : 32;
    e = getenv("PARM_B200_TILE_CH");
    t.ch = e ? atoi(e) : 120; // 15 compute warps x 8 teams of the persistent pair kernel (force_tile.cuh: TILE_PCH)
    if (t.ch < 32) t.ch = 32;
    if (t.ch > 1024) t.ch = 1024;
    e = getenv("PARM_B200_TILE_TEAM");
    t.team = e ? atoi(e) : 4;
    e = getenv("PARM_B200_TILE_V");
    t.v = e ? atoi(e) : 8;
    e = getenv("PARM_B200_TILE_STAGE");
    t.stage = e ? atoi(e) : 1;
    if (!((t.team == 4 && (t.v == 8 || t.v == 4)) || (t.team == 8 && t.v == 4) || (t.team == 2 && t.v == 8))) {
        t.team = 4;
        t.v = 8;
    }
}

void parm_tile_invalidate(parm_nlist *nl) {
    nl->tile.planned = false;
    nl->tile.valid = false;
}

void parm_tile_free(parm_nlist *nl) {
    TileState &t = nl->tile;
    if (t.d_chunks) cudaFree(t.d_chunks);
    if (t.d_s0) cudaFree(t.d_s0);
    t.d_s0 = 0;
    if (t.rows16) cudaFree(t.rows16);
    if (t.d_col) cudaFree(t.d_col);
    if (t.d_info) cudaFree(t.d_info);
    if (t.h_info) cudaFreeHost(t.h_info);
    if (t.h_col) cudaFreeHost(t.h_col);
    if (t.img) cudaFree(t.img);
    if (t.prel_xy) cudaFree(t.prel_xy);
    if (t.prel_z) cudaFree(t.prel_z);
    t.img = 0; t.prel_xy = 0; t.prel_z = 0;
    t.d_chunks = 0; t.rows16 = 0; t.d_col = 0; t.d_info = 0; t.h_info = 0; t.h_col = 0;
}

int parm_tile_plan_enqueue(parm_nlist *nl) {
    parm_ctx *c = nl->ctx;
    TileState &t = nl->tile;
    tile_config(nl);
    parm_tile_invalidate(nl);
    if (!t.enabled || c->D != 3 || nl->smallbox || nl->st.sub != 1) return 0;
    if (!(nl->st.full[1] && nl->st.full[2] && (nl->st.full[0] || nl->st.open0))) return 0;
    bool any = false;
    for (parm_inter *it : c->inters) any = any || (it->nl == nl && inter_fits(it));
    const uint32_t nown = parm_owned(c);
    if (!any || !nown) return 0;
    const uint32_t ncol = (uint32_t)(c->sh.on ? nl->sd.nci : nl->g.nc[0]) * (uint32_t)nl->g.nc[1];
    if (!t.d_info) {
        CK(cudaMalloc(&t.d_info, sizeof(TileInfo)));
        CK(cudaHostAlloc(&t.h_info, sizeof(TileInfo), cudaHostAllocDefault));
    }
    if (ncol + 1 > t.col_cap) {
        if (t.d_col) cudaFree(t.d_col);
        if (t.h_col) cudaFreeHost(t.h_col);
        t.d_col = 0;
        t.h_col = 0;
        t.col_cap = ncol + 1 + ncol / 4;
        CK(cudaMalloc(&t.d_col, 2 * (size_t)t.col_cap * 4));
        CK(cudaHostAlloc(&t.h_col, 2 * (size_t)t.col_cap * 4, cudaHostAllocDefault));
    }
    const uint32_t maxchunks = nown / (uint32_t)t.ch + ncol + 1;
    if (maxchunks > t.chunk_cap) {
        if (t.d_chunks) cudaFree(t.d_chunks);
        if (t.d_s0) cudaFree(t.d_s0);
        t.d_chunks = 0;
        t.d_s0 = 0;
        t.chunk_cap = maxchunks + maxchunks / 4;
        CK(cudaMalloc(&t.d_chunks, (size_t)t.chunk_cap * sizeof(TileChunk)));
        CK(cudaMalloc(&t.d_s0, ((size_t)t.chunk_cap + 1) * 4));
    }
    t.ncol = ncol;
    CK(cudaMemsetAsync(t.d_info, 0, sizeof(TileInfo), c->stream));
    k_tile_cols<<<1, 1024, 0, c->stream>>>(nl->cell_start, (uint32_t)nl->g.nc[2], ncol, (uint32_t)t.ch, t.d_col, t.d_col + t.col_cap, t.d_info);
    CK_LAUNCH(c);
    const double margin = (nl->maxdiam + nl->skin) * (1.0 + 1e-6) + 2.0 * nl->skin;
    // bulk-copy staging subtracts absolute coordinates of the order of the box edge: keep it to boxes where that costs
    // less than 2^-40 (|x| < 4096); larger boxes stage origin-relative positions with the kernel's own threads
    const double lmaxbox = std::max(c->box.L[0], std::max(c->box.L[1], c->box.L[2]));
    t.stage_aligned = t.stage == 1 && lmaxbox < 4096.0 && t.team == 4 && t.v == 8;
    k_tile_chunks<<<(maxchunks + 1 + 127) / 128, 128, 0, c->stream>>>(nl->cell_start, nl->cell_id_sorted, t.d_col, t.d_col + t.col_cap, ncol,
                                                                 (uint32_t)t.ch, c->box, nl->g, nl->st, nl->sd, margin,
                                                                 t.stage_aligned ? 1 : 0, t.d_chunks, t.d_s0, t.d_info);
    CK_LAUNCH(c);
    t.planned = true;
    return 0;
}

int parm_tile_plan_fetch(parm_nlist *nl) {
    parm_ctx *c = nl->ctx;
    TileState &t = nl->tile;
    if (!t.planned) return 0;
    CK(cudaMemcpyAsync(t.h_info, t.d_info, sizeof(TileInfo), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(t.h_col, t.d_col, 2 * (size_t)t.col_cap * 4, cudaMemcpyDeviceToHost, c->stream));
    return 0;
}

// experimental (PARM_B200_TILE_BANKS=1): bank-aware order of every row (bank_order.cuh), after either localize pass
static int tile_bank_order(parm_nlist *nl) {
    parm_ctx *c = nl->ctx;
    TileState &t = nl->tile;
    const char *eb = getenv("PARM_B200_TILE_BANKS"); // read per rebuild: the sweeps toggle it inside one process
    const int banks = eb ? atoi(eb) : 0;
    t.banked = false;
    if (!banks || t.team != 4 || t.v != 8) return 0;
    const size_t smem = 2 * (size_t)(TILE_NT / 4) * nl->kmax * sizeof(uint16_t);
    if (smem > 160 * 1024) return 0;
    if (smem > 32 * 1024) CK(cudaFuncSetAttribute(k_tile_bank_order, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_tile_bank_order<<<t.nchunks, TILE_NT, smem, c->stream>>>(t.d_chunks, nl->cnt, nl->kmax, t.rows16);
    CK_LAUNCH(c);
    t.banked = true;
    return 0;
}

// Called when the rows are final (after NeighborList::ignore, before species packing drops the column tags) and
// the plan summary is on the host.
int parm_tile_localize(parm_nlist *nl) {
    parm_ctx *c = nl->ctx;
    TileState &t = nl->tile;
    t.valid = false;
    if (!t.planned || !nl->tagged) return 0;
    const uint32_t nown = parm_owned(c);
    t.nchunks = t.h_info->nchunks;
    t.max_tile = t.h_info->max_tile;
    if (t.h_info->bad || !t.nchunks || t.max_tile + 1 > TILE_MAX_ATOMS) return 0;
    // short rows (contact-range potentials) are not gather bound: staging a tile per chunk would cost more than it saves
    if (nl->total_full < (uint64_t)t.min_nbrs * nown) return 0;
    const size_t need = (size_t)c->npad * nl->kmax;
    if (need > t.rows16_cap) {
        if (t.rows16) cudaFree(t.rows16);
        t.rows16 = 0;
        t.rows16_cap = 0;
        CK(cudaMalloc(&t.rows16, need * 2));
        t.rows16_cap = need;
    }
    if (t.v == 8) k_tile_localize<8><<<t.nchunks, TILE_NT, 0, c->stream>>>(t.d_chunks, nl->nbr, nl->cnt, nl->kmax, t.team, t.rows16, t.d_info);
    else k_tile_localize<4><<<t.nchunks, TILE_NT, 0, c->stream>>>(t.d_chunks, nl->nbr, nl->cnt, nl->kmax, t.team, t.rows16, t.d_info);
    CK_LAUNCH(c);
    if (int rc = tile_bank_order(nl)) return rc;
    static int check = -1;
    if (check < 0) { const char *e = getenv("PARM_B200_TILE_CHECK"); check = e ? atoi(e) : 0; }
    if (check) { // debugging aid: every row entry must have been found in its chunk's tile
        CK(cudaMemcpyAsync(t.h_info, t.d_info, sizeof(TileInfo), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        if (t.h_info->bad) { parm_set_error("cell-tile plan lost a neighbour (bad = %u)", t.h_info->bad); return PARM_ERR_RUNTIME; }
    }
    t.valid = true;
    return 0;
}

bool parm_tile_all_fit(const parm_nlist *nl) {
    bool any = false;
    for (const parm_inter *it : nl->ctx->inters)
        if (it->nl == nl) {
            if (!inter_fits(it)) return false;
            any = true;
        }
    return any;
}

int parm_tile_rows16_reserve(parm_nlist *nl) {
    parm_ctx *c = nl->ctx;
    TileState &t = nl->tile;
    const size_t need = (size_t)c->npad * nl->kmax;
    if (need > t.rows16_cap) {
        if (t.rows16) cudaFree(t.rows16);
        t.rows16 = 0;
        t.rows16_cap = 0;
        CK(cudaMalloc(&t.rows16, need * 2));
        t.rows16_cap = need;
    }
    return 0;
}

// Mask-mode twin of parm_tile_localize (TEAM = 4, V = 8 only; other layouts fall back to the 32-bit rows).
int parm_tile_localize_masks(parm_nlist *nl) {
    parm_ctx *c = nl->ctx;
    TileState &t = nl->tile;
    t.valid = false;
    if (!t.planned || !nl->mask.active) return 0;
    const uint32_t nown = parm_owned(c);
    t.nchunks = t.h_info->nchunks;
    t.max_tile = t.h_info->max_tile;
    if (t.h_info->bad || !t.nchunks || t.max_tile + 1 > TILE_MAX_ATOMS) return 0;
    if (nl->total_full < (uint64_t)t.min_nbrs * nown) return 0;
    if (t.team != 4 || t.v != 8) return 0;
    const size_t smem = (size_t)(TILE_NT / 4) * (nl->kmax + 8) * sizeof(uint16_t);
    if (smem > 160 * 1024) return 0;
    const size_t need = (size_t)c->npad * nl->kmax;
    if (need > t.rows16_cap) {
        if (t.rows16) cudaFree(t.rows16);
        t.rows16 = 0;
        t.rows16_cap = 0;
        CK(cudaMalloc(&t.rows16, need * 2));
        t.rows16_cap = need;
    }
#define LMARGS t.d_chunks, nl->mask.out, nl->cell_id_sorted, (uint32_t)nl->g.nc[2], nl->mask.gpc, (uint32_t)nl->mask.zg, nl->cnt, nl->kmax, t.rows16, t.d_info
    if (nl->mask.direct) {
        // the build kernel has written rows16 itself
    } else if (nl->h_flags->nbmax <= 32) {
        // (~45 KB of static shared memory on top of the row buffers: always opt in)
        CK(cudaFuncSetAttribute(k_tile_localize_masks_flat<256, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_tile_localize_masks_flat<256, 32><<<t.nchunks, 256, smem, c->stream>>>(LMARGS);
    } else if (nl->h_flags->nbmax <= 48) {
        const size_t smem_f = (size_t)(128 / 4) * (nl->kmax + 8) * sizeof(uint16_t);
        CK(cudaFuncSetAttribute(k_tile_localize_masks_flat<128, 48>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f));
        k_tile_localize_masks_flat<128, 48><<<t.nchunks, 128, smem_f, c->stream>>>(LMARGS);
    } else {
        if (smem + 2048 > 48 * 1024) CK(cudaFuncSetAttribute(k_tile_localize_masks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_tile_localize_masks<<<t.nchunks, TILE_NT, smem, c->stream>>>(LMARGS);
    }
#undef LMARGS
    CK_LAUNCH(c);
    if (int rc = tile_bank_order(nl)) return rc;
    static int check = -1;
    if (check < 0) { const char *e = getenv("PARM_B200_TILE_CHECK"); check = e ? atoi(e) : 0; }
    if (check) {
        CK(cudaMemcpyAsync(t.h_info, t.d_info, sizeof(TileInfo), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        if (t.h_info->bad) { parm_set_error("cell-tile plan lost a neighbour in mask mode (bad = %u)", t.h_info->bad); return PARM_ERR_RUNTIME; }
    }
    t.valid = true;
    return 0;
}

bool parm_tile_usable(const parm_inter *it) {
    const parm_nlist *nl = it->nl;
    return nl->tile.valid && inter_fits(it);
}

// Owned slots [first, end) as a chunk range; false when the range does not start and end on column boundaries.
bool parm_tile_chunk_range(const parm_nlist *nl, uint32_t first, uint32_t end, uint32_t *c0, uint32_t *c1) {
    const TileState &t = nl->tile;
    const uint32_t *slot = t.h_col, *chunk = t.h_col + t.col_cap;
    const uint32_t *a = std::lower_bound(slot, slot + t.ncol + 1, first);
    const uint32_t *b = std::lower_bound(slot, slot + t.ncol + 1, end);
    if (a == slot + t.ncol + 1 || *a != first || b == slot + t.ncol + 1 || *b != end) return false;
    *c0 = chunk[a - slot];
    *c1 = chunk[b - slot];
    return true;
}

extern "C" int parm_nlist_tile_stats(parm_nlist *nl, int *active, uint32_t *chunks, uint32_t *max_tile_atoms, uint32_t *wide_chunks) {
    if (!nl) { parm_set_error("parm_nlist_tile_stats: NULL list"); return PARM_ERR_INVALID; }
    const TileState &t = nl->tile;
    if (active) *active = t.valid ? 1 : 0;
    if (chunks) *chunks = t.valid ? t.nchunks : 0;
    if (max_tile_atoms) *max_tile_atoms = t.valid ? t.max_tile : 0;
    if (wide_chunks) *wide_chunks = t.valid && t.h_info ? t.h_info->wide : 0;
    return 0;
}
