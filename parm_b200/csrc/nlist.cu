// NeighborList on the device: skin-drift trigger + cell-list pair build.
// Replaces NeighborList::update_list (trackers.cpp:19-85), an O(N^2) double loop on the CPU, with
//   cell binning -> radix sort by cell index -> slot re-ordering -> warp-cooperative stencil scan.
// The pair SET equals the reference's bit for bit: a pair is listed iff
//     box->diff(x_i, x_j).norm() < (diam_i + diam_j)/2 + skin            (trackers.cpp:65-66)
// evaluated with the reference's arithmetic (no FMA contraction, e0+(e1+e2), IEEE sqrt, IEEE
// remainder). To keep that off the hot loop, candidates are first classified on wrapped
// coordinates in plain fp64: clearly inside / clearly outside / within a relative band delta of
// the threshold; only the band (a ~1e-9 fraction of the tests) runs the exact predicate.
// The device list is a FULL list (i->j and j->i), row-major: nbr[slot * kmax + k], so the force
// kernel needs no atomics and reads its rows coalesced.
#include <algorithm>
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#include "drift.cuh"
#include "internal.cuh"

static inline unsigned grid_for(const parm_ctx *ctx, uint32_t n, unsigned block, unsigned per_sm = 8) {
    unsigned need = (n + block - 1) / block;
    unsigned cap = (unsigned)ctx->num_sms * per_sm;
    if (need < 1) need = 1;
    return need < cap ? need : cap;
}

// ---- K4a: cell index from the wrapped coordinate; also max |x| (bounds the wrap rounding) ----
// (nearest reference analogue: Grid::get_loc, trackers.cpp:192-219)
__global__ void k_cell_id(const double4 *__restrict__ pos, const uint32_t *__restrict__ src, uint32_t n, BoxDev box,
                          GridDev g, ShardDev sd, uint32_t *cell_id, uint32_t *iota, NlistFlags *flags,
                          uint32_t *cell_count) {
    double xm = 0.0;
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        const uint32_t s = src ? src[q] : q;
        double4 p = pos[s];
        double x[3] = {p.x, p.y, p.z};
        uint32_t c = 0;
#pragma unroll
        for (int d = 0; d < 3; d++) {
            int k;
            if (d == 0 && sd.on) {
                // slab axis, no periodic wrap: interior layers 0..nci-1, then the halo layer above (nci), then
                // the halo layer below (nci+1) -- so a sort by cell id puts owned atoms first, ghosts last
                double rel = shard_rel(x[0], sd);
                if (rel < 0.0) k = sd.nci + 1;
                else if (rel >= sd.Ls) k = sd.nci;
                else {
                    k = (int)floor(rel * sd.nci / sd.Ls);
                    if (k > sd.nci - 1) k = sd.nci - 1;
                }
            } else {
                double w = x[d] - box.L[d] * floor(x[d] * box.invL[d]); // ~[0, L]
                k = (int)floor(w * g.scale[d]);
                if (!(k >= 0)) k = 0; // also catches NaN
                if (k >= g.nc[d]) k = g.nc[d] - 1;
            }
            c = c * (uint32_t)g.nc[d] + (uint32_t)k;
            double ax = fabs(x[d]);
            if (ax > xm && ax < 1e300) xm = ax;
        }
        cell_id[q] = c;
        iota[q] = s;
        atomicAdd(cell_count + c, 1u); // histogram of the one-digit radix (counting) sort; integer, order independent
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) xm = fmax(xm, __shfl_xor_sync(0xffffffffu, xm, o));
    if ((threadIdx.x & 31) == 0 && xm > 0.0) atomicMax(&flags->xmax_bits, (unsigned long long)__double_as_longlong(xm));
}

// relative half-width of the "evaluate exactly" band around the threshold: rounding of the wrapped
// coordinates (~ulp(xmax + L)) relative to the smallest threshold, plus slack for the arithmetic
__device__ __forceinline__ double band_delta(const NlistFlags *flags, double lmax, double thr_min) {
    double xmax = __longlong_as_double((long long)flags->xmax_bits);
    return 1e-9 + 2e-14 * (xmax + lmax) / thr_min;
}

// ---- K4s: radix sort by cell index -------------------------------------------------------------
// The key is the cell index itself, so one counting pass (a radix sort with a single digit as wide as
// the key) is enough: histogram (fused into k_cell_id) -> exclusive scan = cell_start -> every atom
// claims a place in its cell's range -> each atom's final rank inside the cell is the number of cell
// mates that came earlier in the source order, which makes the sort STABLE and hence deterministic
// although the places were claimed with atomics.
#define SCAN_ITEMS 1024
__global__ void __launch_bounds__(256) k_scan_blocks(uint32_t *data, uint32_t n, uint32_t *bsum) {
    __shared__ uint32_t s_warp[8];
    const uint32_t base = blockIdx.x * SCAN_ITEMS + threadIdx.x * 4;
    uint32_t v[4], tot = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        v[k] = base + k < n ? data[base + k] : 0u;
        tot += v[k];
    }
    uint32_t incl = tot;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (int k = 0; k < w; k++) woff += s_warp[k];
    uint32_t run = woff + incl - tot; // exclusive prefix of this thread's 4 items
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (base + k < n) data[base + k] = run;
        run += v[k];
    }
    if (threadIdx.x == 255) bsum[blockIdx.x] = woff + incl;
}
__global__ void __launch_bounds__(1024) k_scan_sums(uint32_t *bsum, uint32_t nb) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (uint32_t b0 = 0; b0 < nb; b0 += 1024) {
        const uint32_t i = b0 + threadIdx.x;
        const uint32_t v = i < nb ? bsum[i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[w] = incl;
        __syncthreads();
        uint32_t woff = 0;
        for (int k = 0; k < w; k++) woff += s_warp[k];
        const uint32_t carry = s_carry;
        if (i < nb) bsum[i] = carry + woff + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + woff + incl;
        __syncthreads();
    }
}
__global__ void k_scan_add(uint32_t *data, uint32_t n, const uint32_t *__restrict__ bsum, uint32_t total_at, uint32_t total) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) data[i] += bsum[i / SCAN_ITEMS];
    if (blockIdx.x == 0 && threadIdx.x == 0) data[total_at] = total;
}
__global__ void k_sort_place(const uint32_t *__restrict__ cell_id, uint32_t n, const uint32_t *__restrict__ cell_start,
                             uint32_t *cell_fill, uint32_t *tmp_q) {
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        const uint32_t c = cell_id[q];
        tmp_q[cell_start[c] + atomicAdd(cell_fill + c, 1u)] = q;
    }
}
__global__ void k_sort_rank(const uint32_t *__restrict__ cell_id, const uint32_t *__restrict__ iota, uint32_t n,
                            const uint32_t *__restrict__ cell_start, const uint32_t *__restrict__ tmp_q, uint32_t *perm,
                            uint32_t *cell_id_sorted) {
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const uint32_t q = tmp_q[t];
        const uint32_t c = cell_id[q];
        const uint32_t lo = cell_start[c], hi = cell_start[c + 1];
        uint32_t rank = 0;
        for (uint32_t k = lo; k < hi; k++) rank += tmp_q[k] < q;
        perm[lo + rank] = iota[q];
        cell_id_sorted[lo + rank] = c;
    }
}

// ---- K4b: apply the sort permutation to every per-slot array -----------------------
__global__ void k_permute(const uint32_t *__restrict__ perm, uint32_t n, uint32_t npad, const double4 *__restrict__ pos,
                          const double *__restrict__ v, const double *__restrict__ a, const double *__restrict__ f,
                          const uint32_t *__restrict__ order, double4 *pos_o, double *v_o, double *a_o, double *f_o,
                          uint32_t *order_o, uint32_t *slot_of, const double *__restrict__ diam_id, double *diam,
                          double *xlast, double4 *pw, BoxDev box, double half_skin, double lmax, double thr_min,
                          const NlistFlags *flags, ShardDev sd, const uint8_t *__restrict__ ghost, uint8_t *ghost_o,
                          float4 *img) {
    const double delta = band_delta(flags, lmax, thr_min);
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        uint32_t o = perm[s];
        double4 p = pos[o];
        pos_o[s] = p;
#pragma unroll
        for (int d = 0; d < 3; d++) {
            size_t so = (size_t)d * npad + s, oo = (size_t)d * npad + o;
            v_o[so] = v[oo];
            a_o[so] = a[oo];
            f_o[so] = f[oo];
        }
        uint32_t id = order[o];
        order_o[s] = id;
        slot_of[id] = s;
        const double dm = diam_id[id];
        diam[s] = dm;
        // wrapped copy for the build (same formula as k_cell_id, so it lies in the atom's cell);
        // w = upper half threshold g_i (1 + delta): g_i + g_j = ((d_i + d_j)/2 + skin)(1 + delta);
        // NaN marks atoms that were never add()ed to the list
        double4 q;
        q.x = sd.on ? shard_rel(p.x, sd) : p.x - box.L[0] * floor(p.x * box.invL[0]);
        q.y = p.y - box.L[1] * floor(p.y * box.invL[1]);
        q.z = p.z - box.L[2] * floor(p.z * box.invL[2]);
        q.w = dm >= 0.0 ? (0.5 * dm + half_skin) * (1.0 + delta) : __longlong_as_double(0x7ff8000000000000LL);
        pw[s] = q;
        // image numbers of the wrapped copy (the slab axis of a sharded context: of the slab-relative frame lo + rel)
        img[s] = make_float4((float)rint((p.x - (sd.on ? q.x + sd.lo : q.x)) * box.invL[0]), (float)rint((p.y - q.y) * box.invL[1]),
                             (float)rint((p.z - q.z) * box.invL[2]), 0.f);
        // lastlocs[i] = a1->x (trackers.cpp:61); ghosts never take part in the drift rule (NaN never wins a >)
        const bool gh = ghost ? ghost[o] != 0 : false;
        if (ghost_o) ghost_o[s] = gh ? 1 : 0;
        const bool skip = gh || !(dm >= 0.0); // ghosts and atoms never add()ed take no part in the drift rule
        const double nanv = __longlong_as_double(0x7ff8000000000000LL);
        xlast[s] = skip ? nanv : p.x;
        xlast[npad + s] = skip ? nanv : p.y;
        xlast[2 * (size_t)npad + s] = skip ? nanv : p.z;
    }
}

// ---- K4b': ghost copies appended behind the (already cell-sorted) owned atoms ------------------
// The sender's boundary layer arrives in the sender's cell order, which is this rank's cell order for its
// halo layer (same y,z grid, stable sorts), so the ghosts only need their per-slot data, not a sort.
__global__ void k_append_ghosts(uint32_t first, uint32_t count, uint32_t npad, const double4 *__restrict__ pos,
                                const uint32_t *__restrict__ order, uint32_t *slot_of, const double *__restrict__ diam_id,
                                double *diam, double *xlast, double4 *pw, uint8_t *ghost, uint32_t *cell_id_sorted,
                                uint32_t *cnt, BoxDev box, GridDev g, ShardDev sd, double half_skin, double lmax,
                                double thr_min, const NlistFlags *flags, float4 *img) {
    const double delta = band_delta(flags, lmax, thr_min);
    const double nanv = __longlong_as_double(0x7ff8000000000000LL);
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < count; q += gridDim.x * blockDim.x) {
        const uint32_t s = first + q;
        const double4 p = pos[s];
        const uint32_t id = order[s];
        const double rel = shard_rel(p.x, sd);
        const int k0 = rel < 0.0 ? sd.nci + 1 : (rel >= sd.Ls ? sd.nci : min((int)floor(rel * sd.nci / sd.Ls), sd.nci - 1));
        double4 q4;
        q4.x = rel;
        q4.y = p.y - box.L[1] * floor(p.y * box.invL[1]);
        q4.z = p.z - box.L[2] * floor(p.z * box.invL[2]);
        int k1 = (int)floor(q4.y * g.scale[1]), k2 = (int)floor(q4.z * g.scale[2]);
        if (!(k1 >= 0)) k1 = 0;
        if (k1 >= g.nc[1]) k1 = g.nc[1] - 1;
        if (!(k2 >= 0)) k2 = 0;
        if (k2 >= g.nc[2]) k2 = g.nc[2] - 1;
        cell_id_sorted[s] = ((uint32_t)k0 * (uint32_t)g.nc[1] + (uint32_t)k1) * (uint32_t)g.nc[2] + (uint32_t)k2;
        const double dm = diam_id[id];
        diam[s] = dm;
        q4.w = dm >= 0.0 ? (0.5 * dm + half_skin) * (1.0 + delta) : nanv;
        pw[s] = q4;
        img[s] = make_float4((float)rint((p.x - (rel + sd.lo)) * box.invL[0]), (float)rint((p.y - q4.y) * box.invL[1]),
                             (float)rint((p.z - q4.z) * box.invL[2]), 0.f);
        xlast[s] = nanv;
        xlast[npad + s] = nanv;
        xlast[2 * (size_t)npad + s] = nanv;
        ghost[s] = 1;
        slot_of[id] = s;
        cnt[s] = 0;
    }
}

// ---- K4c: cell_start[c] = first slot of cell c (sorted ids); cell_start[ncell] = n -----
__global__ void k_cell_start(const uint32_t *__restrict__ cid, uint32_t n, uint32_t ncell, uint32_t *cell_start) {
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s <= n; s += gridDim.x * blockDim.x) {
        // cells in (cid[s-1], cid[s]] start at slot s (the ones strictly between are empty)
        uint32_t lo = s == 0 ? 0u : cid[s - 1] + 1u;
        uint32_t hi = s == n ? ncell : cid[s];
        for (uint32_t c = lo; c <= hi; c++) cell_start[c] = s;
    }
}

// ---- K5: neighbour build ----------------------------------------------------------
// Exact reference predicate (trackers.cpp:65-66) on the UNWRAPPED coordinates.
__device__ __noinline__ bool pair_pred_exact(const double4 *__restrict__ pos, const double *__restrict__ diam,
                                             uint32_t i, uint32_t j, BoxDev box, double skin) {
    const double4 pi = pos[i], pj = pos[j];
    const double di = diam[i], dj = diam[j];
    // box->diff(a1->x, a2->x): remainder(r1 - r2, L) per component (box.hpp:69-72,103)
    double rx = min_image_exact(__dsub_rn(pi.x, pj.x), box.L[0], box.invL[0], box.halfL[0]);
    double ry = min_image_exact(__dsub_rn(pi.y, pj.y), box.L[1], box.invL[1], box.halfL[1]);
    double rz = min_image_exact(__dsub_rn(pi.z, pj.z), box.L[2], box.invL[2], box.halfL[2]);
    // .norm(): sqrt(e0 + (e1 + e2)), no contraction
    double dsq = __dadd_rn(__dmul_rn(rx, rx), __dadd_rn(__dmul_rn(ry, ry), __dmul_rn(rz, rz)));
    // flt diam = (diameters[i] + diameters[j]) / 2;  ... < (diam + skin)
    double thr = __dadd_rn(__dmul_rn(__dadd_rn(di, dj), 0.5), skin);
    return __dsqrt_rn(dsq) < thr;
}

#define BUILD_WARPS 4
// Warp-cooperative build: one warp owns a tile of 32 consecutive slots (atoms of one or two
// neighbouring cells). For every distinct cell in the tile it walks that cell's stencil; the 32
// lanes load 32 consecutive candidates at once (coalesced), then the warp loops over the tile's
// atoms of that cell (their wrapped coordinates are broadcast from shared memory), each lane tests
// its candidate, and the survivors are compacted with ballot + popc into the atom's row.
// UNIFORM: every atom is a list member and all diameters are equal, so the threshold is a constant.
template <bool SMALLBOX, bool UNIFORM>
__global__ void __launch_bounds__(BUILD_WARPS * 32)
k_build(const double4 *__restrict__ pos, const double4 *__restrict__ pw, const double *__restrict__ diam,
        const uint32_t *__restrict__ cid, const uint32_t *__restrict__ cell_start, uint32_t n, BoxDev box, GridDev g,
        StencilDev st, double skin, double lmax, double thr_min, double uthr, uint32_t kmax, uint32_t *__restrict__ nbr,
        uint32_t *__restrict__ cnt, NlistFlags *flags, const uint8_t *__restrict__ ghost) {
    __shared__ double4 s_w[BUILD_WARPS][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t tile = blockIdx.x * BUILD_WARPS + wib;
    const uint32_t base_i = tile * 32u;
    const uint32_t i = base_i + lane;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    double4 wi = make_double4(0, 0, 0, nan);
    uint32_t ci = 0xffffffffu;
    if (i < n) {
        wi = pw[i];
        ci = cid[i];
    }
    // rows are built for list members that this rank owns; ghosts are candidates only
    const bool member = wi.w == wi.w && !(ghost && i < n && ghost[i]);
    s_w[wib][lane] = wi;
    __syncwarp();
    // a test is decided by the fast arithmetic when dsq is outside [lo, 1) * thr_hi^2
    const double delta = band_delta(flags, lmax, thr_min);
    const double lo = ((1.0 - delta) / (1.0 + delta)) * ((1.0 - delta) / (1.0 + delta));
    const double uthr2 = uthr * (1.0 + delta) * (uthr * (1.0 + delta)), uthr2lo = uthr2 * lo;
    const unsigned lt = (1u << lane) - 1u;
    uint32_t count = 0;
    unsigned done = __ballot_sync(0xffffffffu, !member);
    while (done != 0xffffffffu) {
        const int leader = __ffs(~done) - 1;
        const uint32_t cur = __shfl_sync(0xffffffffu, ci, leader);
        const unsigned same = __ballot_sync(0xffffffffu, ci == cur);
        const unsigned act = same & ~done; // tile atoms of this cell that are list members
        done |= same;
        const int cz = (int)(cur % (uint32_t)g.nc[2]);
        const uint32_t t = cur / (uint32_t)g.nc[2];
        const int cy = (int)(t % (uint32_t)g.nc[1]);
        const int cx = (int)(t / (uint32_t)g.nc[1]);
        int x0 = st.full[0] ? cx - st.sub : 0, x1 = st.full[0] ? cx + st.sub : g.nc[0] - 1;
        if (st.open0) { // slab axis of a sharded context: halo layers instead of periodic wrap
            x0 = cx - 1; // layer -1 is the halo below (stored as layer nc-1), layer nc-2 is the halo above
            x1 = cx + 1;
        }
        const int y0 = st.full[1] ? cy - st.sub : 0, y1 = st.full[1] ? cy + st.sub : g.nc[1] - 1;
        const int z0 = st.full[2] ? cz - st.sub : 0, z1 = st.full[2] ? cz + st.sub : g.nc[2] - 1;
        for (int xx = x0; xx <= x1; xx++) {
            int x2 = xx;
            double sx = 0.0; // image shift of the candidates of this cell relative to the tile atoms
            if (st.open0) {
                if (x2 < 0) x2 = g.nc[0] - 1; // halo below; relative coordinates are continuous, no image shift
            } else if (x2 < 0) { x2 += g.nc[0]; sx = -box.L[0]; }
            else if (x2 >= g.nc[0]) { x2 -= g.nc[0]; sx = box.L[0]; }
            for (int yy = y0; yy <= y1; yy++) {
                int y2 = yy;
                double sy = 0.0;
                if (y2 < 0) { y2 += g.nc[1]; sy = -box.L[1]; }
                else if (y2 >= g.nc[1]) { y2 -= g.nc[1]; sy = box.L[1]; }
                const uint32_t rowbase = ((uint32_t)x2 * (uint32_t)g.nc[1] + (uint32_t)y2) * (uint32_t)g.nc[2];
                // cells along z are contiguous in slot order: at most 3 segments (below 0, inside, above nc-1)
                for (int seg = 0; seg < 3; seg++) {
                    int za, zb;
                    double sz = 0.0;
                    if (seg == 0) { za = z0 < 0 ? z0 + g.nc[2] : 1; zb = z0 < 0 ? g.nc[2] - 1 : 0; sz = -box.L[2]; }
                    else if (seg == 1) { za = z0 < 0 ? 0 : z0; zb = z1 >= g.nc[2] ? g.nc[2] - 1 : z1; }
                    else { za = z1 >= g.nc[2] ? 0 : 1; zb = z1 >= g.nc[2] ? z1 - g.nc[2] : 0; sz = box.L[2]; }
                    if (za > zb) continue;
                    const uint32_t jb = cell_start[rowbase + (uint32_t)za], je = cell_start[rowbase + (uint32_t)zb + 1];
                    for (uint32_t jbase = jb; jbase < je; jbase += 32) {
                        const uint32_t j = jbase + lane;
                        double4 wj = make_double4(nan, nan, nan, nan);
                        if (j < je) wj = pw[j];
                        wj.x += sx;
                        wj.y += sy;
                        wj.z += sz;
                        unsigned m = act;
                        while (m) {
                            const int b = __ffs(m) - 1;
                            m &= m - 1;
                            const double4 a = s_w[wib][b];
                            double dx = a.x - wj.x, dy = a.y - wj.y, dz = a.z - wj.z;
                            if (SMALLBOX) { // some axis has too few cells for unique images: fold explicitly
                                if (!st.open0) dx = min_image_fast(dx, box.L[0], box.invL[0]);
                                dy = min_image_fast(dy, box.L[1], box.invL[1]);
                                dz = min_image_fast(dz, box.L[2], box.invL[2]);
                            }
                            const double dsq = fma(dx, dx, fma(dy, dy, dz * dz));
                            double thr2, thr2lo;
                            if (UNIFORM) {
                                thr2 = uthr2;
                                thr2lo = uthr2lo;
                            } else {
                                const double thr = a.w + wj.w; // NaN for non-members / padding lanes
                                thr2 = thr * thr;
                                thr2lo = thr2 * lo;
                            }
                            bool pass = dsq < thr2 && j != base_i + (uint32_t)b; // padding lanes: dsq is NaN
                            if (pass && !(dsq < thr2lo)) pass = pair_pred_exact(pos, diam, base_i + b, j, box, skin);
                            const unsigned bal = __ballot_sync(0xffffffffu, pass);
                            if (bal) {
                                const uint32_t cb = __shfl_sync(0xffffffffu, count, b);
                                if (pass) {
                                    const uint32_t k = cb + __popc(bal & lt);
                                    if (k < kmax) nbr[(size_t)(base_i + b) * kmax + k] = j;
                                }
                                if (lane == b) count += __popc(bal);
                            }
                        }
                    }
                }
            }
        }
    }
    if (i < n) cnt[i] = count;
    // totals: integer atomics are order independent, so the result is deterministic
    unsigned long long wsum = count;
    uint32_t wmax = count;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
        wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    }
    if (lane == 0 && wmax) {
        atomicAdd(&flags->total, wsum);
        atomicMax(&flags->maxcnt, wmax);
    }
}

// ---- K5b: cell-per-warp build (default) ------------------------------------------------------
// One warp owns one cell; lane l owns the cell's l-th atom (cells with more than 32 atoms are walked
// in chunks). For every block of 32 consecutive candidates of the stencil the warp stages the block
// in shared memory (one coalesced load), then every lane tests ITS atom against the 32 candidates
// (broadcast shared-memory reads, no shuffles or ballots) and collects a 32-bit pass mask. Survivors
// are appended to the lane's own row through an 8-entry shared-memory buffer that is written out as
// one full 32-byte sector, so the row stores cost one L2 transaction per 8 neighbours. Same banded
// classification / exact predicate as k_build; ~2.4x fewer instructions per candidate test.
#define CB_WARPS 4
#define CB_DIRECT_BLOCKS 44 // candidate blocks per warp group whose masks fit the warp's shared memory (direct mode)
// packed fp32 pairs (FADD2 / FMUL2 / FFMA2 on sm_100): two candidates per instruction in the fp32 pre-test
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 f2_sub(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 f2_dup(float x) { const unsigned long long b = __float_as_uint(x); return b | (b << 32); }
// RUN2D: in 2-D the cells that are contiguous in slot order run along y (z is a single layer), so y takes the
// role of the run axis. zg: consecutive cells of the run axis handled by one warp (sparse systems have only a
// few atoms per cell; grouping keeps the lanes busy and shares the stencil between the group's cells).
// F32: candidates are first classified in packed fp32 on coordinates relative to the centre of the warp's own
// cells (two candidates per FADD2/FFMA2, pass bits collected with funnel shifts of the sign of
// bits(dsq) - bits(threshold)): clearly inside / clearly outside a relative band beta32 (the rigorous fp32
// rounding bound, x3) around the threshold. A lane that sees a candidate inside that band repeats the block
// with the fp64 test below (which in turn defers to the reference's exact predicate inside its own ~1e-9 band),
// so the pair set stays bit-exact; ~1e-5 of the tests take that path.
// MASK (experimental, internal.cuh MaskState): no rows are written; every lane stores the 32-bit pass mask of each
// candidate block and lane 0 the block's first candidate slot (+ column tag), see tile.cu for the expansion.
template <bool SMALLBOX, bool UNIFORM, bool RUN2D, bool F32, bool MASK = false>
__global__ void __launch_bounds__(CB_WARPS * 32)
k_build_cell(const double4 *__restrict__ pos, const double4 *__restrict__ pw, const double *__restrict__ diam,
             const uint32_t *__restrict__ cell_start, uint32_t ngroups, int zg, uint32_t n, BoxDev box, GridDev g, StencilDev st,
             double skin, double lmax, double thr_min, double uthr, uint32_t kmax, uint32_t *__restrict__ nbr,
             uint32_t *__restrict__ cnt, NlistFlags *flags, const uint8_t *__restrict__ ghost, uint32_t tagcols,
             double cs0, double beta32, MaskOut mo = MaskOut()) {
    __shared__ double4 s_c[CB_WARPS][32];
    __shared__ __align__(8) float s_fx[CB_WARPS][32], s_fy[CB_WARPS][32], s_fz[CB_WARPS][32], s_fw[CB_WARPS][32];
    __shared__ uint32_t s_buf[CB_WARPS][MASK ? 1 : 64][32]; // per lane: ring of 64 pending row entries (a block adds <= 32 to <= 7), flushed 8 (one 32-byte sector) at a time
    // MASK, direct mode: the warp's pass masks [block][lane], the blocks' first slot | column tag, and a per-lane ring of
    // one 32-entry pass of the 16-bit row being assembled
    __shared__ uint32_t s_dm[MASK ? CB_WARPS : 1][MASK ? CB_DIRECT_BLOCKS : 1][32];
    __shared__ uint32_t s_db[MASK ? CB_WARPS : 1][MASK ? CB_DIRECT_BLOCKS : 1];
    __shared__ __align__(16) uint16_t s_dr[MASK ? CB_WARPS : 1][MASK ? 32 : 1][MASK ? 40 : 8]; // row stride 80 bytes: conflict-free 16-byte reads
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t cur = blockIdx.x * CB_WARPS + wib;
    if (cur >= ngroups) return;
    constexpr int RAX = RUN2D ? 1 : 2;                 // physical axis of the run
    const int ncr = g.nc[RAX], ncm = RUN2D ? 1 : g.nc[1];
    const uint32_t gpc = (uint32_t)((ncr + zg - 1) / zg); // groups per column
    const int cr0 = (int)(cur % gpc) * zg, cr1 = min(cr0 + zg, ncr) - 1;
    const uint32_t tcol = cur / gpc;
    const int cy = RUN2D ? 0 : (int)(tcol % (uint32_t)ncm); // mid-axis cell (3-D only)
    const int cx = (int)(tcol / (uint32_t)ncm);
    // cell id of (slow x, mid m, run r)
    auto cell_of = [&](int x, int m, int r) -> uint32_t {
        return RUN2D ? ((uint32_t)x * (uint32_t)g.nc[1] + (uint32_t)r) * (uint32_t)g.nc[2]
                     : ((uint32_t)x * (uint32_t)g.nc[1] + (uint32_t)m) * (uint32_t)g.nc[2] + (uint32_t)r;
    };
    const uint32_t a0 = cell_start[cell_of(cx, cy, cr0)], a1 = cell_start[cell_of(cx, cy, cr1) + 1];
    if (a0 == a1) return;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const double delta = band_delta(flags, lmax, thr_min);
    const double lo = ((1.0 - delta) / (1.0 + delta)) * ((1.0 - delta) / (1.0 + delta));
    const double uthr2 = uthr * (1.0 + delta) * (uthr * (1.0 + delta)), uthr2lo = uthr2 * lo;
    // fp32 pre-test: origin = centre of the warp's own cells (frame of pw), thresholds widened by the band
    const double bt = beta32 + 4.0 * delta;
    const double Ox = ((double)cx + 0.5) * cs0;
    const double Oy = RUN2D ? 0.5 * (double)(cr0 + cr1 + 1) / g.scale[1] : ((double)cy + 0.5) / g.scale[1];
    const double Oz = RUN2D ? 0.0 : 0.5 * (double)(cr0 + cr1 + 1) / g.scale[2];
    const uint32_t u_thi = __float_as_uint(__double2float_ru(uthr2 * (1.0 + bt))), u_tlo = __float_as_uint(__double2float_rd(uthr2lo * (1.0 - bt)));
    const f32x2 c_hi2 = f2_dup(__double2float_ru(1.0 + bt)), c_lo2 = f2_dup(__double2float_rd(lo * (1.0 - bt)));
    int x0 = st.full[0] ? cx - st.sub : 0, x1 = st.full[0] ? cx + st.sub : g.nc[0] - 1;
    if (st.open0) { // slab axis: layer -1 is the halo below (stored as layer nc-1), layer nc-2 the halo above
        x0 = cx - 1;
        x1 = cx + 1;
    }
    const int y0 = RUN2D ? 0 : (st.full[1] ? cy - st.sub : 0), y1 = RUN2D ? 0 : (st.full[1] ? cy + st.sub : g.nc[1] - 1);
    const int z0 = st.full[RAX] ? cr0 - st.sub : 0, z1 = st.full[RAX] ? cr1 + st.sub : ncr - 1; // run-axis range
    unsigned long long wsum_tot = 0;
    uint32_t wmax_tot = 0;
    for (uint32_t chunk = a0; chunk < a1; chunk += 32) {
        const uint32_t i = chunk + lane;
        const bool valid = i < a1;
        double4 wi = make_double4(0, 0, 0, nan);
        if (valid) wi = pw[i];
        const bool member = valid && wi.w == wi.w && !(ghost && ghost[i]);
        if (!__any_sync(0xffffffffu, member)) {
            if (valid) cnt[i] = 0;
            continue;
        }
        const f32x2 xi2 = f2_dup((float)(wi.x - Ox)), yi2 = f2_dup((float)(wi.y - Oy)), zi2 = f2_dup((float)(wi.z - Oz));
        const f32x2 wi2 = f2_dup((float)wi.w);
        uint32_t count = 0, flushed = 0; // entries found / entries already written to the row (a multiple of 8)
        uint32_t nblk = 0;               // MASK: candidate blocks walked so far (the same sequence for every chunk of the group)
        uint32_t *row = nbr + (size_t)(valid ? i : a0) * kmax;
        auto flush8 = [&](uint32_t *r, uint32_t at) {
            const uint32_t h = MASK ? 0u : (at & 56u);
            uint4 u0, u1;
            u0.x = s_buf[wib][h + 0][lane]; u0.y = s_buf[wib][h + 1][lane]; u0.z = s_buf[wib][h + 2][lane]; u0.w = s_buf[wib][h + 3][lane];
            u1.x = s_buf[wib][h + 4][lane]; u1.y = s_buf[wib][h + 5][lane]; u1.z = s_buf[wib][h + 6][lane]; u1.w = s_buf[wib][h + 7][lane];
            uint4 *dst = reinterpret_cast<uint4 *>(r + at);
            dst[0] = u0;
            dst[1] = u1;
        };
        for (int xx = x0; xx <= x1; xx++) {
            int x2 = xx;
            double sx = 0.0;
            if (st.open0) {
                if (x2 < 0) x2 = g.nc[0] - 1; // halo below; relative coordinates are continuous, no image shift
            } else if (x2 < 0) { x2 += g.nc[0]; sx = -box.L[0]; }
            else if (x2 >= g.nc[0]) { x2 -= g.nc[0]; sx = box.L[0]; }
            for (int yy = y0; yy <= y1; yy++) {
                int y2 = yy;
                double sy = 0.0;
                if (y2 < 0) { y2 += g.nc[1]; sy = -box.L[1]; }
                else if (y2 >= g.nc[1]) { y2 -= g.nc[1]; sy = box.L[1]; }
                // stencil column 0..8 in the top bits of the entries (tile.cu turns them into tile-local indices)
                const uint32_t tag = tagcols ? (uint32_t)((xx - x0) * 3 + (yy - y0)) << PARM_NBR_SLOT_BITS : 0u;
                for (int seg = 0; seg < 3; seg++) {
                    int za, zb;
                    double sr = 0.0; // image shift along the run axis
                    if (seg == 0) { za = z0 < 0 ? z0 + ncr : 1; zb = z0 < 0 ? ncr - 1 : 0; sr = -box.L[RAX]; }
                    else if (seg == 1) { za = z0 < 0 ? 0 : z0; zb = z1 >= ncr ? ncr - 1 : z1; }
                    else { za = z1 >= ncr ? 0 : 1; zb = z1 >= ncr ? z1 - ncr : 0; sr = box.L[RAX]; }
                    if (za > zb) continue;
                    const double sz = RUN2D ? 0.0 : sr;
                    const double syy = RUN2D ? sr : sy;
                    const uint32_t jb = cell_start[cell_of(x2, y2, za)], je = cell_start[cell_of(x2, y2, zb) + 1];
                    for (uint32_t jbase = jb; jbase < je; jbase += 32) {
                        {   // stage 32 candidates (image shift applied once per candidate)
                            const uint32_t j = jbase + lane;
                            double4 wj = make_double4(nan, nan, nan, nan);
                            if (j < je) wj = pw[j];
                            wj.x += sx;
                            wj.y += syy;
                            wj.z += sz;
                            __syncwarp();
                            s_c[wib][lane] = wj;
                            if (F32) { // non-members and lanes past the end of the run hold NaN and fail every test
                                const bool cand = wj.w == wj.w;
                                s_fx[wib][lane] = cand ? (float)(wj.x - Ox) : __int_as_float(0x7fffffff);
                                s_fy[wib][lane] = (float)(wj.y - Oy);
                                s_fz[wib][lane] = (float)(wj.z - Oz);
                                s_fw[wib][lane] = (float)wj.w;
                            }
                            __syncwarp();
                        }
                        const int nb = (int)min(32u, je - jbase);
                        const uint32_t iself = i - jbase; // candidate index of this lane's own atom, if in the block
                        uint32_t m = 0;
                        bool near_any = false;
                        // fp64 test, 4 x 8: the inner 8 tests are unrolled (bit position = immediate + c0); lanes past the end
                        // of the run hold NaN and fail every test
                        auto test64 = [&]() {
                            m = 0;
                            for (int c0 = 0; c0 < 32; c0 += 8) {
                                uint32_t m8 = 0;
#pragma unroll
                                for (int k = 0; k < 8; k++) {
                                    const double4 q = s_c[wib][c0 + k];
                                    double dx = wi.x - q.x, dy = wi.y - q.y, dz = wi.z - q.z;
                                    if (SMALLBOX) {
                                        if (!st.open0) dx = min_image_fast(dx, box.L[0], box.invL[0]);
                                        dy = min_image_fast(dy, box.L[1], box.invL[1]);
                                        dz = min_image_fast(dz, box.L[2], box.invL[2]);
                                    }
                                    const double dsq = fma(dx, dx, fma(dy, dy, dz * dz));
                                    double thr2, thr2lo;
                                    if (UNIFORM) {
                                        thr2 = uthr2;
                                        thr2lo = uthr2lo;
                                    } else {
                                        const double thr = wi.w + q.w; // NaN for non-members
                                        thr2 = thr * thr;
                                        thr2lo = thr2 * lo;
                                    }
                                    const bool pass = dsq < thr2;
                                    near_any |= pass && !(dsq < thr2lo);
                                    if (pass) m8 |= 1u << k;
                                }
                                m |= m8 << c0;
                                if (c0 + 8 >= nb) break;
                            }
                        };
                        if (F32) {
                            uint32_t mp = 0, mi = 0; // "possibly inside" / "certainly inside", first candidate in the top bit
                            int done = 0;
                            for (int c0 = 0; c0 < 32; c0 += 8) {
#pragma unroll
                                for (int k = 0; k < 4; k++) {
                                    const int p2 = c0 + 2 * k;
                                    const f32x2 dx = f2_sub(xi2, *reinterpret_cast<const f32x2 *>(&s_fx[wib][p2]));
                                    const f32x2 dy = f2_sub(yi2, *reinterpret_cast<const f32x2 *>(&s_fy[wib][p2]));
                                    const f32x2 dz = f2_sub(zi2, *reinterpret_cast<const f32x2 *>(&s_fz[wib][p2]));
                                    const f32x2 dsq = f2_fma(dx, dx, f2_fma(dy, dy, f2_mul(dz, dz)));
                                    uint32_t h0 = u_thi, h1 = u_thi, l0 = u_tlo, l1 = u_tlo;
                                    if (!UNIFORM) {
                                        const f32x2 t = f2_add(wi2, *reinterpret_cast<const f32x2 *>(&s_fw[wib][p2]));
                                        const f32x2 t2 = f2_mul(t, t);
                                        const f32x2 th = f2_mul(t2, c_hi2), tl = f2_mul(t2, c_lo2);
                                        h0 = (uint32_t)th; h1 = (uint32_t)(th >> 32);
                                        l0 = (uint32_t)tl; l1 = (uint32_t)(tl >> 32);
                                    }
                                    // non-negative floats order like their bit patterns; NaN (0x7fffffff) is never below a threshold
                                    const uint32_t s0 = (uint32_t)dsq, s1 = (uint32_t)(dsq >> 32);
                                    mp = __funnelshift_l(s0 - h0, mp, 1);
                                    mi = __funnelshift_l(s0 - l0, mi, 1);
                                    mp = __funnelshift_l(s1 - h1, mp, 1);
                                    mi = __funnelshift_l(s1 - l1, mi, 1);
                                }
                                done = c0 + 8;
                                if (done >= nb) break;
                            }
                            mp = __brev(mp << (32 - done));
                            mi = __brev(mi << (32 - done));
                            m = mp;
                            const bool near32 = (mp & ~mi) != 0 && member;
                            if (__any_sync(0xffffffffu, near32)) {
                                if (near32) test64(); // some candidate within the fp32 band: this lane repeats the block in fp64
                            }
                        } else {
                            test64();
                        }
                        if (iself < 32u) m &= ~(1u << iself); // never its own neighbour
                        if (__any_sync(0xffffffffu, near_any && member)) {
                            // some test fell inside the band (~1e-9 of them): decide with the reference's exact predicate
                            if (near_any && member) {
                                m = 0;
                                for (int c = 0; c < nb; c++) {
                                    const double4 q = s_c[wib][c];
                                    double dx = wi.x - q.x, dy = wi.y - q.y, dz = wi.z - q.z;
                                    if (SMALLBOX) {
                                        if (!st.open0) dx = min_image_fast(dx, box.L[0], box.invL[0]);
                                        dy = min_image_fast(dy, box.L[1], box.invL[1]);
                                        dz = min_image_fast(dz, box.L[2], box.invL[2]);
                                    }
                                    const double dsq = fma(dx, dx, fma(dy, dy, dz * dz));
                                    const double thr = UNIFORM ? uthr * (1.0 + delta) : wi.w + q.w;
                                    const double thr2 = thr * thr;
                                    bool pass = dsq < thr2 && iself != (uint32_t)c;
                                    if (pass && !(dsq < thr2 * lo)) pass = pair_pred_exact(pos, diam, i, jbase + c, box, skin);
                                    m |= (pass ? 1u : 0u) << c;
                                }
                            }
                        }
                        if (!member) m = 0;
                        // append this lane's survivors to its ring; full groups of 8 leave as one 32-byte sector. The flush
                        // is checked once per candidate block, not per survivor: with 32 lanes some lane would be flushing
                        // in nearly every iteration of the divergent loop
                        const uint32_t jt = jbase | tag;
                        if (MASK) {
                            if (nblk < mo.mb_cap) {
                                if (valid) mo.masks[(size_t)nblk * mo.npad + i] = m;
                                if (lane == 0) mo.blk_base[(size_t)cur * mo.mb_cap + nblk] = jt;
                            }
                            if (mo.direct && nblk < CB_DIRECT_BLOCKS) {
                                s_dm[wib][nblk][lane] = m;
                                if (lane == 0) s_db[wib][nblk] = jt;
                            }
                            count += __popc(m);
                            nblk++;
                            continue;
                        }
                        while (m) {
                            const int bit = __ffs(m) - 1;
                            m &= m - 1;
                            s_buf[wib][count & 63u][lane] = jt + (uint32_t)bit;
                            count++;
                        }
                        while (count - flushed >= 8u) {
                            if (flushed + 8 <= kmax) flush8(row, flushed);
                            flushed += 8;
                        }
                    }
                }
            }
        }
        if (MASK) {
            if (lane == 0) {
                mo.grp_nb[cur] = nblk;
                atomicMax(&flags->nbmax, nblk);
            }
            if (mo.direct) {
                // Expansion, lane = atom: every lane walks the set bits of ITS masks in one flat loop (rows are ~110 +- 10
                // long, so the lanes of a warp finish almost together), turns candidate slots into indices of its chunk's
                // tile through the chunk's run table, assembles one 32-entry pass of the row at a time in the pair
                // kernel's lane-vector layout and writes it out as 64 bytes; the last pass is padded with the sentinel.
                __syncwarp();
                if (nblk > CB_DIRECT_BLOCKS) {
                    if (lane == 0) *mo.direct_fail = 1u;
                } else if (valid) {
                    const uint32_t q = tcol; // owned column (single GPU: column index = cx * nc1 + cy)
                    const uint32_t cs = mo.col_slot[q], cc = mo.col_chunk[q];
                    const uint32_t ccnt = mo.col_slot[q + 1] - cs, cnch = mo.col_chunk[q + 1] - cc;
                    const uint32_t csz = min(mo.ch, (ccnt + cnch - 1) / cnch);
                    const TileChunk *TC = mo.chunks + cc + (i - cs) / csz;
                    const uint32_t ntile = TC->ntile;
                    uint16_t *out = mo.rows16 + (size_t)i * kmax;
                    uint16_t *ring = s_dr[wib][lane];
                    const uint32_t kend = min(count, kmax);
                    uint32_t k = 0, m = 0, ty = 0, tw = 0, l0 = 0, n0 = 0, jz = 0, b = 0;
                    for (;;) {
                        if (!m) {
                            while (b < nblk && !(m = s_dm[wib][b][lane])) b++;
                            if (b >= nblk) break;
                            const uint32_t base = s_db[wib][b++];
                            const uint32_t kc = 2u * min(base >> PARM_NBR_SLOT_BITS, 8u);
                            const uint32_t jb = base & PARM_NBR_SLOT_MASK;
                            const uint32_t tx = TC->seg_start[kc], tz = TC->seg_start[kc + 1];
                            ty = TC->seg_off[kc];
                            tw = TC->seg_off[kc + 1];
                            l0 = jb - tx; n0 = tw - ty; jz = jb - tz;
                        }
                        const uint32_t bit = (uint32_t)__ffs(m) - 1u;
                        m &= m - 1u;
                        const uint32_t r = l0 + bit;
                        const uint32_t l = r < n0 ? ty + r : tw + (jz + bit);
                        if (k < kend) {
                            const uint32_t r32 = k & 31u;
                            ring[((r32 << 3) & 24u) | (r32 >> 2)] = (uint16_t)(min(l, ntile) << TILE_IDX_SHIFT);
                            if (r32 == 31u) { // a whole pass: 64 bytes
                                uint4 *dst = reinterpret_cast<uint4 *>(out + (k & ~31u));
                                const uint4 *src = reinterpret_cast<const uint4 *>(ring);
                                dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
                            }
                        }
                        k++;
                    }
                    if (kend & 31u) { // tail: pad the last pass with the sentinel index
                        for (uint32_t kk = kend & 31u; kk < 32u; kk++) ring[((kk << 3) & 24u) | (kk >> 2)] = (uint16_t)(ntile << TILE_IDX_SHIFT);
                        uint4 *dst = reinterpret_cast<uint4 *>(out + (kend & ~31u));
                        const uint4 *src = reinterpret_cast<const uint4 *>(ring);
                        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
                    }
                }
                __syncwarp();
            }
        } else if (member) { // tail of the row
            for (uint32_t k = flushed; k < count && k < kmax; k++) row[k] = s_buf[wib][k & 63u][lane];
        }
        if (valid) cnt[i] = count;
        wsum_tot += count;
        wmax_tot = max(wmax_tot, count);
    }
    // totals: integer atomics are order independent, so the result is deterministic
    unsigned long long wsum = wsum_tot;
    uint32_t wmax = wmax_tot;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
        wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    }
    if (lane == 0 && wmax) {
        atomicAdd(&flags->total, wsum);
        atomicMax(&flags->maxcnt, wmax);
    }
}

// ---- standalone drift check (update_list(false) outside timestep()) -------------------
__global__ void __launch_bounds__(256)
k_drift(const double4 *__restrict__ pos, const double *__restrict__ xlast, const double *__restrict__ diam, uint32_t n,
        uint32_t npad, double skin, double *d_top2, unsigned int *counter, NlistFlags *dflags, NlistFlags *hflags) {
    double b1 = 0.0, b2 = 0.0;
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        if (!(diam[s] >= 0.0)) continue;
        double4 p = pos[s];
        double d = drift_dist(p, xlast[s], xlast[npad + s], xlast[2 * (size_t)npad + s]);
        top2_push(b1, b2, d);
    }
    drift_finish(b1, b2, skin, d_top2, counter, dflags, hflags);
}

// ---- host side ------------------------------------------------------------------------
extern "C" int parm_nlist_create(parm_ctx *c, double skin, parm_nlist **out) {
    if (!c || !out) { parm_set_error("parm_nlist_create: NULL argument"); return PARM_ERR_INVALID; }
    *out = 0;
    if (!c->nlists.empty()) {
        parm_set_error("parm_nlist_create: one NeighborList per AtomVec context is supported (several NListed "
                       "interactions may share it, as in LJatoms.cpp:46-48); see DESIGN.md out-of-scope");
        return PARM_ERR_UNSUPPORTED;
    }
    if (!(skin >= 0)) { parm_set_error("parm_nlist_create: skin must be >= 0"); return PARM_ERR_INVALID; }
    CK(cudaSetDevice(c->device));
    parm_nlist *nl = new parm_nlist();
    nl->ctx = c;
    nl->skin = skin;
    nl->ignorechanged = true; // trackers.cpp:17
    nl->h_diam.assign(c->nid, -1.0);
    size_t np = c->npad;
    const size_t nidp = std::max(c->npad, c->nid_pad);
    CK(cudaMalloc(&nl->d_diam_id, nidp * 8));
    CK(cudaMalloc(&nl->d_diam, np * 8));
    CK(cudaMalloc(&nl->xlast, 3 * np * 8));
    CK(cudaMalloc(&nl->pw, np * sizeof(double4)));
    CK(cudaMalloc(&nl->tile.img, np * sizeof(float4)));
    CK(cudaMalloc(&nl->tile.prel_xy, np * sizeof(double2)));
    CK(cudaMalloc(&nl->tile.prel_z, np * sizeof(double)));
    CK(cudaMemsetAsync(nl->tile.prel_xy, 0, np * sizeof(double2), c->stream));
    CK(cudaMemsetAsync(nl->tile.prel_z, 0, np * sizeof(double), c->stream));
    nl->cell_sub = 1;
    if (const char *e = getenv("PARM_B200_CELL_SUB")) nl->cell_sub = atoi(e) == 2 ? 2 : 1;
    CK(cudaMemsetAsync(nl->xlast, 0, 3 * np * 8, c->stream));
    CK(cudaMalloc(&nl->cell_id, np * 4));
    CK(cudaMalloc(&nl->cell_id_sorted, np * 4));
    CK(cudaMalloc(&nl->perm, np * 4));
    CK(cudaMalloc(&nl->sort_tmp, np * 4));
    CK(cudaMalloc(&nl->iota, np * 4));
    CK(cudaMalloc(&nl->cnt, np * 4));
    CK(cudaMemsetAsync(nl->cnt, 0, np * 4, c->stream));
    CK(cudaMalloc(&nl->d_top2, 2 * sizeof(double) * 4096));
    CK(cudaMalloc(&nl->d_counter, 4));
    CK(cudaMemsetAsync(nl->d_counter, 0, 4, c->stream));
    CK(cudaMalloc(&nl->d_flags, sizeof(NlistFlags)));
    CK(cudaMemsetAsync(nl->d_flags, 0, sizeof(NlistFlags), c->stream));
    CK(cudaHostAlloc(&nl->h_flags, sizeof(NlistFlags), cudaHostAllocMapped));
    memset(nl->h_flags, 0, sizeof(NlistFlags));
    CK(cudaMalloc(&nl->d_slot, 4 * sizeof(int)));
    CK(cudaMemsetAsync(nl->d_slot, 0, 4 * sizeof(int), c->stream));
    CK(cudaHostAlloc(&nl->h_slot, 4 * sizeof(int), cudaHostAllocMapped));
    nl->h_slot[0] = nl->h_slot[1] = nl->h_slot[2] = nl->h_slot[3] = 0;
    // all atoms start as non-members
    std::vector<double> neg(nidp, -1.0);
    CK(cudaMemcpyAsync(nl->d_diam_id, neg.data(), nidp * 8, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(nl->d_diam, neg.data(), np * 8, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->nlists.push_back(nl);
    *out = nl;
    return 0;
}

extern "C" int parm_nlist_destroy(parm_nlist *nl) {
    if (!nl) return 0;
    parm_ctx *c = nl->ctx;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    void *ptrs[] = {nl->pw, nl->d_diam_id, nl->d_diam, nl->xlast, nl->cell_id, nl->cell_id_sorted, nl->perm, nl->iota,
                    nl->cell_start, nl->cell_fill, nl->scan_sums, nl->sort_tmp, nl->nbr, nl->cnt, nl->d_top2, nl->d_counter, nl->d_flags, nl->d_slot};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (nl->d_excl_start) cudaFree(nl->d_excl_start);
    if (nl->d_excl) cudaFree(nl->d_excl);
    if (nl->h_flags) cudaFreeHost(nl->h_flags);
    if (nl->h_slot) cudaFreeHost(nl->h_slot);
    if (nl->mask.d_fail) cudaFree(nl->mask.d_fail);
    if (nl->mask.h_fail) cudaFreeHost(nl->mask.h_fail);
    parm_tile_free(nl);
    c->nlists.erase(std::remove(c->nlists.begin(), c->nlists.end(), nl), c->nlists.end());
    delete nl;
    return 0;
}

__global__ void k_gather_by_order_d(const double *__restrict__ src_id, const uint32_t *__restrict__ order, uint32_t n, double *dst) {
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) dst[s] = src_id[order[s]];
}

extern "C" int parm_nlist_set_diameters(parm_nlist *nl, const double *diam) {
    if (!nl || !diam) { parm_set_error("parm_nlist_set_diameters: NULL argument"); return PARM_ERR_INVALID; }
    parm_ctx *c = nl->ctx;
    CK(cudaSetDevice(c->device));
    double maxd = 0, mind = INFINITY;
    bool any = false;
    for (uint32_t i = 0; i < c->nid; i++) {
        double d = diam[i];
        if (d >= 0 && !isinf(d)) {
            nl->h_diam[i] = d;
            any = true;
            if (d > maxd) maxd = d;
            if (d < mind) mind = d;
        } else
            nl->h_diam[i] = -1.0;
    }
    nl->have_diam = any;
    nl->maxdiam = maxd;
    nl->uniform = any && mind == maxd;
    if (nl->uniform)
        for (uint32_t i = 0; i < c->nid; i++)
            if (!(nl->h_diam[i] >= 0)) { nl->uniform = false; break; }
    nl->mindiam = any ? mind : 0.0;
    if (c->nid) CK(cudaMemcpyAsync(nl->d_diam_id, nl->h_diam.data(), (size_t)c->nid * 8, cudaMemcpyHostToDevice, c->stream));
    if (c->n) {
        k_gather_by_order_d<<<grid_for(c, c->n, 256), 256, 0, c->stream>>>(nl->d_diam_id, c->order, c->n, nl->d_diam);
        CK_LAUNCH(c);
        CK(cudaStreamSynchronize(c->stream)); // h_diam may be re-written by the caller's next call
    }
    nl->ignorechanged = true; // trackers.hpp:200
    return 0;
}

static int alloc_nbr(parm_nlist *nl, uint32_t kmax) {
    parm_ctx *c = nl->ctx;
    kmax = (kmax + 31u) & ~31u; // rows start on 128-byte boundaries
    size_t need = (size_t)c->npad * kmax;
    if (need > nl->nbr_cap_entries) {
        if (nl->nbr) cudaFree(nl->nbr);
        nl->nbr = 0;
        nl->nbr_cap_entries = 0;
        CK(cudaMalloc(&nl->nbr, need * 4));
        nl->nbr_cap_entries = need;
    }
    nl->kmax = kmax;
    return 0;
}

// ---- rebuild steps (shared by the single-GPU path below and csrc/shard.cu) -------------------
// Cell grid for the current box: cells no smaller than the largest possible pair threshold / sub.
int parm_nlist_prepare_grid(parm_nlist *nl) {
    parm_ctx *c = nl->ctx;
    const uint32_t n = c->n;
    const double rlist = (nl->maxdiam + nl->skin) * (1.0 + 1e-6) + 1e-12;
    GridDev &g = nl->g;
    StencilDev &st = nl->st;
    ShardDev &sd = nl->sd;
    memset(&sd, 0, sizeof(sd));
    uint64_t ncell = 1;
    // identical on every rank of a sharded context (the ranks must agree on the y/z grid)
    const uint64_t cell_cap = std::max<uint64_t>(64, c->sh.on ? 8ull * (c->nid / c->sh.nranks + 1) : 4ull * std::max<uint32_t>(n, 1u));
    const int sub = nl->cell_sub;
    for (int d = 0; d < 3; d++) {
        int k = 1;
        if (d < c->D) {
            double q = floor(c->box.L[d] * sub / rlist);
            k = q < 1 ? 1 : (q > 2048 ? 2048 : (int)q);
        }
        g.nc[d] = k;
    }
    if (c->sh.on) {
        // slab axis: interior layers of width Ls/nci >= r_list plus one halo layer on each side
        int nci = (int)floor(c->sh.Ls / rlist);
        if (nci < 2) {
            parm_set_error("slab decomposition: slab width %g must be at least 2 x (max diameter + skin) = %g", c->sh.Ls, 2 * rlist);
            return PARM_ERR_INVALID;
        }
        if (nci > 2046) nci = 2046;
        sd.on = 1;
        sd.lo = c->sh.lo;
        sd.Ls = c->sh.Ls;
        sd.L = c->box.L[0];
        sd.nci = nci;
        g.nc[0] = nci + 2;
    }
    // keep the cell table within a few entries per atom (never shrink the slab axis)
    for (;;) {
        ncell = (uint64_t)g.nc[0] * g.nc[1] * g.nc[2];
        if (ncell <= cell_cap) break;
        int big = c->sh.on ? 1 : 0;
        for (int d = big + 1; d < 3; d++)
            if (g.nc[d] > g.nc[big]) big = d;
        if (g.nc[big] <= 1) break;
        g.nc[big] = (g.nc[big] + 1) / 2;
    }
    st.sub = c->sh.on ? 1 : sub;
    st.open0 = c->sh.on ? 1 : 0;
    nl->smallbox = false;
    nl->lmax = 0;
    for (int d = 0; d < 3; d++) {
        g.scale[d] = g.nc[d] / c->box.L[d];
        st.full[d] = (g.nc[d] >= 2 * st.sub + 1) ? 1 : 0;
        if (!st.full[d] && d < c->D && !(d == 0 && st.open0)) nl->smallbox = true;
        nl->nc[d] = g.nc[d];
        if (d < c->D) nl->lmax = std::max(nl->lmax, c->box.L[d]);
    }
    nl->ncell = (uint32_t)ncell;
    if (nl->ncell + 2 > nl->cell_start_cap) {
        if (nl->cell_start) cudaFree(nl->cell_start);
        nl->cell_start = 0;
        nl->cell_start_cap = nl->ncell + 2 + nl->ncell / 4;
        CK(cudaMalloc(&nl->cell_start, (size_t)nl->cell_start_cap * 4));
        if (nl->cell_fill) cudaFree(nl->cell_fill);
        if (nl->scan_sums) cudaFree(nl->scan_sums);
        nl->cell_fill = nl->scan_sums = 0;
        CK(cudaMalloc(&nl->cell_fill, (size_t)nl->cell_start_cap * 4));
        CK(cudaMalloc(&nl->scan_sums, ((size_t)nl->cell_start_cap / SCAN_ITEMS + 2) * 4));
    }
    nl->thr_min = nl->mindiam + nl->skin;
    if (!(nl->thr_min > 1e-300)) nl->thr_min = 1e-300;
    return 0;
}

// Bin the atoms in slots d_src[0..nsrc) (NULL: slots 0..nsrc-1), sort them by cell index (stable LSD
// radix sort) and re-order every per-slot array; afterwards the context holds exactly those nsrc atoms
// in slots 0..nsrc-1, cell-sorted, and cell_start is valid.
int parm_nlist_sort_permute(parm_nlist *nl, const uint32_t *d_src, uint32_t nsrc) {
    parm_ctx *c = nl->ctx;
    const uint32_t n = nsrc;
    c->n = nsrc;
    if (n == 0) {
        CK(cudaMemsetAsync(nl->cell_start, 0, (size_t)(nl->ncell + 1) * 4, c->stream));
        return 0;
    }
    // one-digit radix (counting) sort by cell index; cell_start falls out of the scan
    const uint32_t nce = nl->ncell;
    CK(cudaMemsetAsync(nl->cell_start, 0, (size_t)(nce + 1) * 4, c->stream));
    CK(cudaMemsetAsync(nl->cell_fill, 0, (size_t)(nce + 1) * 4, c->stream));
    k_cell_id<<<grid_for(c, n, 256), 256, 0, c->stream>>>(c->pos, d_src, n, c->box, nl->g, nl->sd, nl->cell_id, nl->iota,
                                                          nl->d_flags, nl->cell_start);
    CK_LAUNCH(c);
    const uint32_t nsb = (nce + SCAN_ITEMS - 1) / SCAN_ITEMS;
    k_scan_blocks<<<nsb, 256, 0, c->stream>>>(nl->cell_start, nce, nl->scan_sums);
    CK_LAUNCH(c);
    k_scan_sums<<<1, 1024, 0, c->stream>>>(nl->scan_sums, nsb);
    CK_LAUNCH(c);
    k_scan_add<<<grid_for(c, nce, 256), 256, 0, c->stream>>>(nl->cell_start, nce, nl->scan_sums, nce, n);
    CK_LAUNCH(c);
    k_sort_place<<<grid_for(c, n, 256), 256, 0, c->stream>>>(nl->cell_id, n, nl->cell_start, nl->cell_fill, nl->sort_tmp);
    CK_LAUNCH(c);
    k_sort_rank<<<grid_for(c, n, 256, 16), 256, 0, c->stream>>>(nl->cell_id, nl->iota, n, nl->cell_start, nl->sort_tmp, nl->perm,
                                                                nl->cell_id_sorted);
    CK_LAUNCH(c);
    k_permute<<<grid_for(c, n, 256), 256, 0, c->stream>>>(nl->perm, n, c->npad, c->pos, c->v, c->a, c->f, c->order,
                                                          c->pos_alt, c->v_alt, c->a_alt, c->f_alt, c->order_alt,
                                                          c->slot_of, nl->d_diam_id, nl->d_diam, nl->xlast, nl->pw,
                                                          c->box, 0.5 * nl->skin, nl->lmax, nl->thr_min, nl->d_flags,
                                                          nl->sd, c->ghost, c->ghost_alt, nl->tile.img);
    CK_LAUNCH(c);
    std::swap(c->pos, c->pos_alt);
    std::swap(c->v, c->v_alt);
    std::swap(c->a, c->a_alt);
    std::swap(c->f, c->f_alt);
    std::swap(c->order, c->order_alt);
    std::swap(c->ghost, c->ghost_alt);
    return 0;
}

// Sharded contexts: `count` ghost copies (pos + id already in slots [first, first+count), in the sender's
// cell order) become part of the cell structure without another sort.
int parm_nlist_append_ghosts(parm_nlist *nl, uint32_t first, uint32_t count) {
    parm_ctx *c = nl->ctx;
    c->n = first + count;
    if (count) {
        k_append_ghosts<<<grid_for(c, count, 256), 256, 0, c->stream>>>(first, count, c->npad, c->pos, c->order, c->slot_of,
                                                                        nl->d_diam_id, nl->d_diam, nl->xlast, nl->pw, c->ghost,
                                                                        nl->cell_id_sorted, nl->cnt, c->box, nl->g, nl->sd,
                                                                        0.5 * nl->skin, nl->lmax, nl->thr_min, nl->d_flags, nl->tile.img);
        CK_LAUNCH(c);
    }
    k_cell_start<<<grid_for(c, c->n + 1, 256), 256, 0, c->stream>>>(nl->cell_id_sorted, c->n, nl->ncell, nl->cell_start);
    CK_LAUNCH(c);
    return 0;
}

// NeighborList::ignore (trackers.cpp:64 `if (ignorepairs.has_pair(a1, a2)) continue;`): excluded pairs are
// removed from the finished rows. One warp per slot whose atom has exclusions; the row is compacted in place
// (stable), 32 entries at a time.
__global__ void k_apply_ignore(const uint32_t *__restrict__ order, const uint32_t *__restrict__ excl_start,
                               const uint32_t *__restrict__ excl, uint32_t nown, uint32_t kmax, uint32_t mask, uint32_t *nbr,
                               uint32_t *cnt, NlistFlags *flags) {
    const uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (s >= nown) return;
    const uint32_t id = order[s];
    const uint32_t e0 = excl_start[id], e1 = excl_start[id + 1];
    if (e0 == e1) return;
    const uint32_t my = min(cnt[s], kmax);
    uint32_t *row = nbr + (size_t)s * kmax;
    uint32_t kept = 0;
    for (uint32_t k0 = 0; k0 < my; k0 += 32) {
        const uint32_t k = k0 + lane;
        const bool valid = k < my;
        const uint32_t j = valid ? row[k] : 0; // (tag bits move with the entry)
        bool keep = valid;
        if (valid) {
            const uint32_t jid = order[j & mask];
            for (uint32_t e = e0; e < e1; e++) keep = keep && excl[e] != jid;
        }
        const uint32_t m = __ballot_sync(0xffffffffu, keep);
        __syncwarp();
        if (keep) row[kept + __popc(m & ((1u << lane) - 1))] = j;
        kept += __popc(m);
        __syncwarp();
    }
    if (lane == 0 && kept != my) {
        cnt[s] = kept;
        atomicAdd(&flags->total, (unsigned long long)0 - (unsigned long long)(my - kept));
    }
}

static int apply_ignore(parm_nlist *nl) {
    parm_ctx *c = nl->ctx;
    if (nl->ignored.empty()) return 0;
    if (nl->ignore_dirty) { // CSR by AtomVec index, both directions
        std::vector<uint32_t> start(c->nid + 1, 0), list(2 * nl->ignored.size());
        for (const auto &pr : nl->ignored) { start[pr.first + 1]++; start[pr.second + 1]++; }
        for (uint32_t i = 0; i < c->nid; i++) start[i + 1] += start[i];
        std::vector<uint32_t> fill(start.begin(), start.end() - 1);
        for (const auto &pr : nl->ignored) { list[fill[pr.first]++] = pr.second; list[fill[pr.second]++] = pr.first; }
        if (nl->d_excl) cudaFree(nl->d_excl);
        nl->d_excl = 0;
        if (!nl->d_excl_start) CK(cudaMalloc(&nl->d_excl_start, ((size_t)c->nid + 1) * 4));
        CK(cudaMalloc(&nl->d_excl, list.size() * 4));
        CK(cudaMemcpyAsync(nl->d_excl_start, start.data(), start.size() * 4, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(nl->d_excl, list.data(), list.size() * 4, cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream)); // the host vectors go out of scope
        nl->ignore_dirty = false;
    }
    const uint32_t nown = parm_owned(c);
    if (!nown) return 0;
    k_apply_ignore<<<(unsigned)(((size_t)nown * 32 + 255) / 256), 256, 0, c->stream>>>(c->order, nl->d_excl_start, nl->d_excl, nown,
                                                                                      nl->kmax, PARM_NBR_MASK_OF(nl), nl->nbr, nl->cnt, nl->d_flags);
    CK_LAUNCH(c);
    CK(cudaMemcpyAsync(nl->h_flags, nl->d_flags, sizeof(NlistFlags), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    nl->total_full = nl->h_flags->total;
    return 0;
}

// entry |= species[entry] << 27 for every row entry (once per rebuild, so that the per-step force kernel does not
// gather a species byte per neighbour)
__global__ void k_pack_species(const uint8_t *__restrict__ spec, uint32_t nown, uint32_t kmax, uint32_t *nbr,
                               const uint32_t *__restrict__ cnt) {
    const uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (s >= nown) return;
    const uint32_t my = min(cnt[s], kmax);
    uint32_t *row = nbr + (size_t)s * kmax;
    for (uint32_t k = lane; k < my; k += 32) {
        const uint32_t j = row[k] & PARM_NBR_SLOT_MASK; // drops the stencil-column tag of the build
        row[k] = j | ((uint32_t)spec[j] << PARM_NBR_SLOT_BITS);
    }
}

static int pack_species(parm_nlist *nl) {
    parm_ctx *c = nl->ctx;
    nl->packed = false; // the build just rewrote every entry as a plain slot index
    nl->packed_for = nullptr;
    if (c->npad > PARM_NBR_SLOT_MASK) return 0;
    parm_inter *primary = nullptr;
    for (parm_inter *it : c->inters)
        if (it->nl == nl && it->have_params && !it->generic && it->nspecies > 1) { primary = it; break; }
    const uint32_t nown = parm_owned(c);
    if (!primary || !nown) return 0;
    // short rows (e.g. contact-range repulsion, ~8 neighbours) are dominated by per-atom overhead: the extra pass per
    // rebuild does not pay (measured: WCA, 8 neighbours, 0.340 ms unpacked vs 0.371 ms packed; LJ, 110: 0.509 vs 0.405)
    if (nl->total_full < (uint64_t)PARM_PACK_MIN_NEIGHBORS * nown) return 0;
    k_pack_species<<<(unsigned)(((size_t)nown * 32 + 255) / 256), 256, 0, c->stream>>>(primary->d_spec, nown, nl->kmax, nl->nbr, nl->cnt);
    CK_LAUNCH(c);
    nl->packed = true;
    nl->tagged = false;
    nl->packed_for = primary;
    primary->spec_stale = false;
    return 0;
}

static int finish_rows(parm_nlist *nl) {
    if (nl->mask.active) { // no exclusions, one species: the masks go straight into the 16-bit tile rows
        nl->packed = false;
        nl->packed_for = nullptr;
        return parm_tile_localize_masks(nl);
    }
    PTRY(apply_ignore(nl));
    PTRY(parm_tile_localize(nl)); // 16-bit tile-local rows for the cell-tile pair kernel (tile.cu): needs the column tags
    return pack_species(nl);
}

// ---- mask-mode build (experimental): 32-bit rows on demand ---------------------------------------------------
// One warp per atom: lane b takes candidate block b of the atom's warp group, a warp scan of the pass-mask popcounts
// gives every block its position in the row, and each lane writes `first slot | tag` + bit for its set bits -- the
// same entries, in the same order, as the append loop of the classic build.
__global__ void __launch_bounds__(256)
k_expand_masks(MaskOut mo, const uint32_t *__restrict__ cell_id_sorted, uint32_t nc2, uint32_t gpc, uint32_t zg, uint32_t n,
               uint32_t kmax, const uint32_t *__restrict__ cnt, uint32_t *__restrict__ nbr) {
    const uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (s >= n) return;
    if (cnt[s] == 0) return; // (groups without members never wrote their masks)
    const uint32_t cid = cell_id_sorted[s];
    const uint32_t grp = (cid / nc2) * gpc + (cid % nc2) / zg;
    const uint32_t nb = min(mo.grp_nb[grp], mo.mb_cap);
    uint32_t *row = nbr + (size_t)s * kmax;
    uint32_t pos0 = 0;
    for (uint32_t r = 0; r < nb; r += 32) {
        const uint32_t b = r + lane;
        uint32_t m = b < nb ? mo.masks[(size_t)b * mo.npad + s] : 0u;
        const uint32_t base = b < nb ? mo.blk_base[(size_t)grp * mo.mb_cap + b] : 0u;
        const uint32_t p = __popc(m);
        uint32_t incl = p;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (uint32_t)o) incl += y;
        }
        uint32_t k = pos0 + incl - p;
        while (m) {
            const int bit = __ffs(m) - 1;
            m &= m - 1;
            if (k < kmax) row[k] = base + (uint32_t)bit;
            k++;
        }
        pos0 += __shfl_sync(0xffffffffu, incl, 31);
    }
}

int parm_nlist_ensure_rows32(parm_nlist *nl) {
    MaskState &ms = nl->mask;
    if (!ms.active || ms.rows32_valid) return 0;
    parm_ctx *c = nl->ctx;
    const uint32_t n = c->n;
    if (n) {
        k_expand_masks<<<(unsigned)(((size_t)n * 32 + 255) / 256), 256, 0, c->stream>>>(
            ms.out, nl->cell_id_sorted, (uint32_t)nl->g.nc[2], ms.gpc, (uint32_t)ms.zg, n, nl->kmax, nl->cnt, nl->nbr);
        CK_LAUNCH(c);
        // Slab-decomposed contexts run the force launches of a step on two streams (interior: the context's own stream,
        // boundary layers: the communication stream, integ.cu). Whichever of them asked for the rows first, BOTH wait for
        // the expansion: the host-side flag below is all the second one would look at.
        if (c->sh.on) {
            CK(cudaEventRecord(c->sh.ev_rows, c->stream));
            CK(cudaStreamWaitEvent(c->sh.main_stream, c->sh.ev_rows, 0));
            CK(cudaStreamWaitEvent(c->sh.comm_stream, c->sh.ev_rows, 0));
        }
    }
    ms.rows32_valid = true;
    return 0;
}

static int alloc_masks(parm_nlist *nl, uint32_t mb_cap, uint32_t ngroups) {
    parm_ctx *c = nl->ctx;
    MaskState &ms = nl->mask;
    const size_t need_m = (size_t)mb_cap * c->npad, need_b = (size_t)ngroups * mb_cap;
    if (need_m > ms.masks_cap) {
        if (ms.out.masks) cudaFree(ms.out.masks);
        ms.out.masks = 0;
        ms.masks_cap = 0;
        CK(cudaMalloc(&ms.out.masks, need_m * 4));
        ms.masks_cap = need_m;
    }
    if (need_b > ms.blk_cap) {
        if (ms.out.blk_base) cudaFree(ms.out.blk_base);
        ms.out.blk_base = 0;
        ms.blk_cap = 0;
        CK(cudaMalloc(&ms.out.blk_base, need_b * 4));
        ms.blk_cap = need_b;
    }
    if (ngroups > ms.grp_cap) {
        if (ms.out.grp_nb) cudaFree(ms.out.grp_nb);
        ms.out.grp_nb = 0;
        ms.grp_cap = 0;
        CK(cudaMalloc(&ms.out.grp_nb, (size_t)ngroups * 4));
        ms.grp_cap = ngroups;
    }
    ms.out.mb_cap = mb_cap;
    ms.out.npad = c->npad;
    return 0;
}

// Build the rows for the atoms now in slots 0..n-1, growing the per-atom capacity if a row overflowed.
int parm_nlist_build_rows(parm_nlist *nl) {
    parm_ctx *c = nl->ctx;
    const uint32_t n = c->n;
    for (parm_inter *it : c->inters) PTRY(parm_inter_regather(it));
    parm_tile_invalidate(nl);
    if (n == 0) { // (a sharded rank that owns no atoms) close the profiling record opened by the caller
        nl->total_full = 0;
        nl->maxcnt = 0;
        if (c->prof_on && !c->prof_pending.empty()) PTRY(parm_prof_end(c));
        return 0;
    }
    if (nl->kmax == 0) {
        double vol = 1;
        for (int d = 0; d < c->D; d++) vol *= c->box.L[d];
        double rl = nl->maxdiam + nl->skin;
        double sphere = c->D == 3 ? 4.18879020478639 * rl * rl * rl : 3.14159265358979 * rl * rl;
        double ntot = c->sh.on ? (double)c->nid : (double)n;
        double est = ntot / vol * sphere;
        uint32_t k0 = (uint32_t)std::min<double>(std::max(16.0, 1.25 * est + 12.0), std::max(ntot, 2.0) - 1.0);
        PTRY(alloc_nbr(nl, std::max<uint32_t>(k0, 1u)));
    }
    const unsigned ntiles = (n + 31) / 32;
    const unsigned nblocks = (ntiles + BUILD_WARPS - 1) / BUILD_WARPS;
    const double uthr = nl->maxdiam + nl->skin;
    PTRY(parm_tile_plan_enqueue(nl)); // chunk table of the cell-tile pair kernel: needs the cell structure only
    for (int attempt = 0; attempt < 8; attempt++) {
        uint32_t kmax_launch = nl->kmax;
        CK(cudaMemsetAsync((char *)nl->d_flags + offsetof(NlistFlags, maxcnt), 0,
                           offsetof(NlistFlags, xmax_bits) - offsetof(NlistFlags, maxcnt), c->stream));
#define CARGS c->pos, nl->pw, nl->d_diam, nl->cell_start, ngroups, zg, n, c->box, nl->g, nl->st, nl->skin, nl->lmax, \
              nl->thr_min, uthr, nl->kmax, nl->nbr, nl->cnt, nl->d_flags, c->sh.on ? c->ghost : nullptr, tagcols, cs0, beta32
#define BARGS c->pos, nl->pw, nl->d_diam, nl->cell_id_sorted, nl->cell_start, n, c->box, nl->g, nl->st, nl->skin, nl->lmax, \
              nl->thr_min, uthr, nl->kmax, nl->nbr, nl->cnt, nl->d_flags, c->sh.on ? c->ghost : nullptr
        static int per_cell = -1;
        if (per_cell < 0) { const char *e = getenv("PARM_B200_BUILD_PER_CELL"); per_cell = e ? atoi(e) : 1; }
        // with a tile plan (3 x 3 stencil columns, tile.cu) the entries carry their stencil column in the top bits
        const uint32_t tagcols = per_cell && nl->tile.planned && c->npad <= PARM_NBR_SLOT_MASK ? 1u : 0u;
        nl->tagged = tagcols != 0;
        if (!tagcols) parm_tile_invalidate(nl);
        if (per_cell) {
            // one warp per group of zg consecutive cells along the run axis, ~28 atoms per group
            const bool run2d = c->D == 2;
            const int rax = run2d ? 1 : 2;
            const int ncr = nl->g.nc[rax];
            int zg = 1;
            if (nl->st.full[rax]) {
                const double per_cell_atoms = (double)n / (double)nl->ncell;
                zg = (int)floor(28.0 / std::max(per_cell_atoms, 0.05) + 0.5);
                zg = std::max(1, std::min(zg, std::min(64, ncr - 2 * nl->st.sub)));
            }
            const uint32_t gpc = (uint32_t)((ncr + zg - 1) / zg);
            const uint32_t ngroups = (uint32_t)nl->g.nc[0] * (run2d ? 1u : (uint32_t)nl->g.nc[1]) * gpc;
            const unsigned cblocks = (ngroups + CB_WARPS - 1) / CB_WARPS;
            // fp32 pre-test: |coordinate - warp origin| <= E, so |dsq32 - dsq| <= (6.93 E / thr + 5) 2^-24 thr^2 at the
            // threshold, plus 7 x 2^-24 for a per-pair threshold formed in fp32 and 2 x 2^-24 for its scaling; x3 margin
            const double cs0 = c->sh.on ? nl->sd.Ls / nl->sd.nci : c->box.L[0] / nl->g.nc[0];
            double ext = (nl->st.sub + 1.0) * cs0;
            for (int d = 1; d < c->D; d++) {
                const double cs = c->box.L[d] / nl->g.nc[d];
                ext = std::max(ext, d == rax ? (0.5 * zg + nl->st.sub) * cs : (nl->st.sub + 1.0) * cs);
            }
            double beta32 = 3.0 * ((6.93 * ext / nl->thr_min + 14.0) * 5.9604644775390625e-8);
            // mask mode (default; PARM_B200_BUILD_MASKS=0 selects the classic append build): 3-D tile-planned single-GPU lists without exclusions whose last build
            // had long rows and whose interactions all run on the cell-tile kernel
            const char *em = getenv("PARM_B200_BUILD_MASKS"); // read per rebuild: the sweeps toggle it inside one process
            const int masks_env = em ? atoi(em) : 1;
            nl->mask.enabled = masks_env;
            const char *ems = getenv("PARM_B200_BUILD_MASKS_SHARDED");
            const bool masks_sharded = ems ? atoi(ems) != 0 : true; // slab-decomposed contexts: rows of the owned atoms only, ghosts are candidates
            const bool use_masks = masks_env && tagcols && !run2d && !nl->smallbox && (!c->sh.on || masks_sharded) && nl->ignored.empty() &&
                                   nl->total_full >= (uint64_t)nl->tile.min_nbrs * n && parm_tile_all_fit(nl);
            kmax_launch = nl->kmax;
            nl->mask.direct = false;
            if (use_masks) {
                PTRY(alloc_masks(nl, std::max(nl->mask.out.mb_cap, 48u), ngroups));
                nl->mask.ngroups = ngroups;
                nl->mask.gpc = gpc;
                nl->mask.zg = zg;
                // direct mode: the build warps expand their masks into rows16 themselves (tile.cu's localize pass is skipped)
                const char *ed = getenv("PARM_B200_BUILD_DIRECT");
                MaskOut &mo = nl->mask.out;
                // (measured: +0.6 ms per rebuild at N = 1e6 -- the build kernel runs 16 warps per SM, too few to hide the
                // latencies of the expansion loop; off by default, the separate localize pass is faster)
                mo.direct = (ed ? atoi(ed) != 0 : false) && nl->tile.team == 4 && nl->tile.v == 8 ? 1u : 0u;
                if (mo.direct) {
                    if (!nl->mask.d_fail) {
                        CK(cudaMalloc(&nl->mask.d_fail, 4));
                        CK(cudaHostAlloc(&nl->mask.h_fail, 4, cudaHostAllocDefault));
                    }
                    PTRY(parm_tile_rows16_reserve(nl));
                    CK(cudaMemsetAsync(nl->mask.d_fail, 0, 4, c->stream));
                    mo.chunks = nl->tile.d_chunks;
                    mo.col_slot = nl->tile.d_col;
                    mo.col_chunk = nl->tile.d_col + nl->tile.col_cap;
                    mo.ch = (uint32_t)nl->tile.ch;
                    mo.rows16 = nl->tile.rows16;
                    mo.direct_fail = nl->mask.d_fail;
                    nl->mask.direct = true;
                }
            }
            nl->mask.active = use_masks;
            nl->mask.rows32_valid = !use_masks;
            static int use_f32 = -1;
            if (use_f32 < 0) { const char *e = getenv("PARM_B200_BUILD_F32"); use_f32 = e ? atoi(e) : 1; }
            const bool f32 = use_f32 && !nl->smallbox && beta32 < 1e-4;
            if (!f32) beta32 = 0.0;
#define CLAUNCH(SB, UN, F)                                                                                      \
    do {                                                                                                        \
        if (run2d) k_build_cell<SB, UN, true, F><<<cblocks, CB_WARPS * 32, 0, c->stream>>>(CARGS);                 \
        else k_build_cell<SB, UN, false, F><<<cblocks, CB_WARPS * 32, 0, c->stream>>>(CARGS);                      \
    } while (0)
#define MLAUNCH(UN, F) k_build_cell<false, UN, false, F, true><<<cblocks, CB_WARPS * 32, 0, c->stream>>>(CARGS, nl->mask.out)
            if (use_masks) {
                if (f32) { if (nl->uniform) MLAUNCH(true, true); else MLAUNCH(false, true); }
                else { if (nl->uniform) MLAUNCH(true, false); else MLAUNCH(false, false); }
            } else
            if (nl->smallbox) { if (nl->uniform) CLAUNCH(true, true, false); else CLAUNCH(true, false, false); }
            else if (f32) { if (nl->uniform) CLAUNCH(false, true, true); else CLAUNCH(false, false, true); }
            else { if (nl->uniform) CLAUNCH(false, true, false); else CLAUNCH(false, false, false); }
#undef CLAUNCH
#undef MLAUNCH
        } else if (nl->smallbox) {
            if (nl->uniform) k_build<true, true><<<nblocks, BUILD_WARPS * 32, 0, c->stream>>>(BARGS);
            else k_build<true, false><<<nblocks, BUILD_WARPS * 32, 0, c->stream>>>(BARGS);
        } else {
            if (nl->uniform) k_build<false, true><<<nblocks, BUILD_WARPS * 32, 0, c->stream>>>(BARGS);
            else k_build<false, false><<<nblocks, BUILD_WARPS * 32, 0, c->stream>>>(BARGS);
        }
#undef BARGS
#undef CARGS
        CK_LAUNCH(c);
        CK(cudaMemcpyAsync(nl->h_flags, nl->d_flags, sizeof(NlistFlags), cudaMemcpyDeviceToHost, c->stream));
        if (nl->mask.active && nl->mask.direct) CK(cudaMemcpyAsync(nl->mask.h_fail, nl->mask.d_fail, 4, cudaMemcpyDeviceToHost, c->stream));
        PTRY(parm_tile_plan_fetch(nl));
        CK(cudaStreamSynchronize(c->stream));
        nl->total_full = nl->h_flags->total;
        nl->maxcnt = nl->h_flags->maxcnt;
        if (nl->mask.active) {
            if (nl->h_flags->nbmax > nl->mask.out.mb_cap) { // a group walked more candidate blocks than the mask table holds
                nl->mask.out.mb_cap = nl->h_flags->nbmax + 8;
                continue; // (alloc_masks runs again at the top of the next attempt)
            }
            // rows16 written by the build itself are good unless a group overflowed its shared-memory masks or a row its capacity
            if (nl->mask.direct && (*nl->mask.h_fail || nl->maxcnt > kmax_launch)) nl->mask.direct = false;
            if (nl->maxcnt > nl->kmax) PTRY(alloc_nbr(nl, nl->maxcnt + nl->maxcnt / 8 + 8)); // capacity only: the masks are complete
            PTRY(finish_rows(nl));
            return parm_prof_end(c); // the rebuild record covers sort, build and the tile rows
        }
        if (nl->maxcnt <= nl->kmax) {
            PTRY(finish_rows(nl));
            return parm_prof_end(c);
        }
        uint32_t k2 = nl->maxcnt + nl->maxcnt / 8 + 8;
        PTRY(alloc_nbr(nl, k2));
    }
    parm_set_error("neighbour list capacity did not converge");
    return PARM_ERR_RUNTIME;
}

int parm_nlist_rebuild(parm_nlist *nl) {
    parm_ctx *c = nl->ctx;
    CK(cudaSetDevice(c->device));
    if (!c->box_set) { parm_set_error("NeighborList update before the box was set"); return PARM_ERR_INVALID; }
    nl->updatenum++; // trackers.cpp:56-57
    nl->ignorechanged = false;
    nl->rebuilds++;
    if (c->sh.on) return parm_shard_rebuild(nl);
    if (c->n == 0) { nl->total_full = 0; nl->maxcnt = 0; return 0; }
    PTRY(parm_nlist_prepare_grid(nl));
    PTRY(parm_prof_begin(c, PARM_PROF_REBUILD));
    CK(cudaMemsetAsync(nl->d_flags, 0, sizeof(NlistFlags), c->stream));
    PTRY(parm_nlist_sort_permute(nl, nullptr, c->n));
    return parm_nlist_build_rows(nl);
}

int parm_nlist_drift_check_async(parm_nlist *nl) {
    parm_ctx *c = nl->ctx;
    unsigned nb = grid_for(c, c->n, 256, 4);
    k_drift<<<nb, 256, 0, c->stream>>>(c->pos, nl->xlast, nl->d_diam, c->n, c->npad, nl->skin, nl->d_top2, nl->d_counter,
                                        nl->d_flags, nl->h_flags);
    CK_LAUNCH(c);
    return 0;
}

extern "C" int parm_nlist_update(parm_nlist *nl, int force, int *rebuilt) {
    if (!nl) { parm_set_error("parm_nlist_update: NULL list"); return PARM_ERR_INVALID; }
    parm_ctx *c = nl->ctx;
    CK(cudaSetDevice(c->device));
    if (rebuilt) *rebuilt = 0;
    if (!force && !nl->ignorechanged) {
        if (c->n == 0 && !c->sh.on) return 0;
        PTRY(parm_nlist_drift_check_async(nl));
        if (c->sh.on) {
            bool rb = false;
            PTRY(parm_shard_drift_decision(nl, &rb));
            if (!rb) return 0;
        } else {
            CK(cudaStreamSynchronize(c->stream));
            if (!nl->h_flags->need_rebuild) return 0;
        }
    }
    PTRY(parm_nlist_rebuild(nl));
    if (rebuilt) *rebuilt = 1;
    return 0;
}

extern "C" int parm_nlist_which(parm_nlist *nl, uint32_t *u) { *u = nl->updatenum; return 0; }

extern "C" int parm_nlist_ignore(parm_nlist *nl, const uint32_t *a, const uint32_t *b, uint64_t npairs) {
    if (!nl || (npairs && (!a || !b))) { parm_set_error("parm_nlist_ignore: NULL argument"); return PARM_ERR_INVALID; }
    parm_ctx *c = nl->ctx;
    for (uint64_t k = 0; k < npairs; k++) {
        if (a[k] >= c->nid || b[k] >= c->nid) { parm_set_error("parm_nlist_ignore: atom index out of range"); return PARM_ERR_INVALID; }
        if (a[k] == b[k]) continue; // an atom never pairs with itself
        nl->ignored.insert(std::make_pair(std::max(a[k], b[k]), std::min(a[k], b[k])));
    }
    nl->ignore_dirty = true;
    nl->ignorechanged = true; // trackers.hpp:192
    return 0;
}
extern "C" int parm_nlist_ignore_size(parm_nlist *nl, uint64_t *n) {
    if (!nl || !n) { parm_set_error("parm_nlist_ignore_size: NULL argument"); return PARM_ERR_INVALID; }
    *n = nl->ignored.size();
    return 0;
}

// Host copy of the pair list in the reference's order (trackers.cpp:59-68): first = later atom i,
// last = earlier atom j < i, sorted by (i, j). Sharded contexts return the pairs whose `first`
// atom is local to this rank (the union over ranks is the global list, each pair exactly once).
static int collect_pairs(parm_nlist *nl, std::vector<std::pair<uint32_t, uint32_t> > &out) {
    parm_ctx *c = nl->ctx;
    out.clear();
    const uint32_t n = c->n;
    if (n == 0 || nl->updatenum == 0) return 0;
    std::vector<uint32_t> h_cnt(n), h_order(n), h_nbr((size_t)n * nl->kmax);
    std::vector<uint8_t> h_ghost(n, 0);
    PTRY(parm_nlist_ensure_rows32(nl)); // mask-mode lists expand their 32-bit rows on demand
    CK(cudaMemcpyAsync(h_cnt.data(), nl->cnt, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(h_order.data(), c->order, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(h_nbr.data(), nl->nbr, h_nbr.size() * 4, cudaMemcpyDeviceToHost, c->stream));
    if (c->sh.on) CK(cudaMemcpyAsync(h_ghost.data(), c->ghost, n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    out.reserve(nl->total_full / 2 + 16);
    for (uint32_t s = 0; s < n; s++) {
        if (h_ghost[s]) continue;
        const uint32_t i = h_order[s];
        const uint32_t *row = h_nbr.data() + (size_t)s * nl->kmax;
        for (uint32_t q = 0; q < h_cnt[s]; q++) {
            uint32_t j = h_order[row[q] & PARM_NBR_MASK_OF(nl)];
            if (j < i) out.push_back(std::make_pair(i, j));
        }
    }
    std::sort(out.begin(), out.end());
    if (!c->sh.on && out.size() != nl->total_full / 2) {
        parm_set_error("neighbour list is not symmetric (%llu canonical pairs vs %llu row entries / 2)",
                       (unsigned long long)out.size(), (unsigned long long)(nl->total_full / 2));
        return PARM_ERR_RUNTIME;
    }
    return 0;
}

extern "C" int parm_nlist_numpairs(parm_nlist *nl, uint64_t *np) {
    if (!nl->ctx->sh.on) { *np = nl->total_full / 2; return 0; }
    std::vector<std::pair<uint32_t, uint32_t> > pr;
    CK(cudaSetDevice(nl->ctx->device));
    PTRY(collect_pairs(nl, pr));
    *np = pr.size();
    return 0;
}

extern "C" int parm_nlist_stats(parm_nlist *nl, double *mean_full, uint32_t *max_full) {
    uint64_t members = 0;
    if (nl->ctx->sh.on) members = nl->ctx->sh.n_local;
    else for (double d : nl->h_diam) members += d >= 0;
    if (mean_full) *mean_full = members ? (double)nl->total_full / members : 0.0;
    if (max_full) *max_full = nl->maxcnt;
    return 0;
}

extern "C" int parm_nlist_download_pairs(parm_nlist *nl, uint32_t *first, uint32_t *last, uint64_t cap) {
    CK(cudaSetDevice(nl->ctx->device));
    std::vector<std::pair<uint32_t, uint32_t> > pr;
    PTRY(collect_pairs(nl, pr));
    if (cap < pr.size()) { parm_set_error("parm_nlist_download_pairs: capacity %llu < %llu pairs", (unsigned long long)cap, (unsigned long long)pr.size()); return PARM_ERR_INVALID; }
    for (size_t k = 0; k < pr.size(); k++) {
        first[k] = pr[k].first;
        last[k] = pr[k].second;
    }
    return 0;
}
