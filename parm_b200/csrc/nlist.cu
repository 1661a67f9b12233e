// NeighborList on the device: skin-drift trigger + cell-list pair build.
// Replaces NeighborList::update_list (trackers.cpp:19-85), which is an O(N^2) double
// loop on the CPU, with: cell binning -> radix sort by cell index -> slot re-ordering ->
// per-atom scan of the 3^D cell stencil with the reference's exact predicate
//     box->diff(x_i, x_j).norm() < (diam_i + diam_j)/2 + skin     (trackers.cpp:65-66)
// evaluated bit-for-bit (no FMA contraction, e0+(e1+e2) association, IEEE sqrt), so the
// pair SET equals the reference's. The device list is a FULL list (both i->j and j->i)
// stored in warp tiles nbr[tile][k][lane] so the force kernel needs no atomics.
#include <algorithm>
#include <cub/device/device_radix_sort.cuh>
#include <math.h>
#include <string.h>

#include "drift.cuh"
#include "internal.cuh"

static inline unsigned grid_for(const parm_ctx *ctx, uint32_t n, unsigned block, unsigned per_sm = 8) {
    unsigned need = (n + block - 1) / block;
    unsigned cap = (unsigned)ctx->num_sms * per_sm;
    if (need < 1) need = 1;
    return need < cap ? need : cap;
}

struct GridDev {
    int nc[3];
    double scale[3]; // nc / L
};

// ---- K4a: cell index from the wrapped coordinate --------------------------------
// (nearest reference analogue: Grid::get_loc, trackers.cpp:192-219)
__global__ void k_cell_id(const double4 *__restrict__ pos, uint32_t n, BoxDev box, GridDev g, uint32_t *cell_id,
                          uint32_t *iota) {
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        double4 p = pos[s];
        double x[3] = {p.x, p.y, p.z};
        uint32_t c = 0;
#pragma unroll
        for (int d = 0; d < 3; d++) {
            double w = x[d] - box.L[d] * floor(x[d] * box.invL[d]); // ~[0, L)
            int k = (int)floor(w * g.scale[d]);
            if (!(k >= 0)) k = 0; // also catches NaN
            if (k >= g.nc[d]) k = g.nc[d] - 1;
            c = c * (uint32_t)g.nc[d] + (uint32_t)k;
        }
        cell_id[s] = c;
        iota[s] = s;
    }
}

// ---- K4b: apply the sort permutation to every per-slot array -----------------------
__global__ void k_permute(const uint32_t *__restrict__ perm, uint32_t n, uint32_t npad, const double4 *__restrict__ pos,
                          const double *__restrict__ v, const double *__restrict__ a, const double *__restrict__ f,
                          const uint32_t *__restrict__ order, double4 *pos_o, double *v_o, double *a_o, double *f_o,
                          uint32_t *order_o, uint32_t *slot_of, const double *__restrict__ diam_id, double *diam,
                          double *xlast) {
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
        diam[s] = diam_id[id];
        // lastlocs[i] = a1->x (trackers.cpp:61)
        xlast[s] = p.x;
        xlast[npad + s] = p.y;
        xlast[2 * (size_t)npad + s] = p.z;
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
struct StencilDev {
    int noff[3];
    int off[3][3];
};

// One thread per atom (slot). Lanes of a warp are consecutive slots, i.e. atoms of the same
// or adjacent cells, so the candidate loads pos[j], diam[j] are warp-wide broadcasts; each
// lane appends to its own column of the tile, so appends of a warp coalesce.
__global__ void __launch_bounds__(128)
k_build(const double4 *__restrict__ pos, const double *__restrict__ diam, const uint32_t *__restrict__ cid,
        const uint32_t *__restrict__ cell_start, uint32_t n, BoxDev box, GridDev g, StencilDev st, double skin,
        uint32_t kmax, uint32_t *__restrict__ nbr, uint32_t *__restrict__ cnt, NlistFlags *flags) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t count = 0;
    if (s < n) {
        const double4 pi = pos[s];
        const double di = diam[s];
        if (di >= 0.0) {
            uint32_t c = cid[s];
            int cz = (int)(c % (uint32_t)g.nc[2]);
            uint32_t t = c / (uint32_t)g.nc[2];
            int cy = (int)(t % (uint32_t)g.nc[1]);
            int cx = (int)(t / (uint32_t)g.nc[1]);
            uint32_t *col = nbr + ((size_t)(s >> 5) * kmax) * PARM_TILE + (s & 31u);
            for (int ix = 0; ix < st.noff[0]; ix++) {
                int x2 = cx + st.off[0][ix];
                x2 += x2 < 0 ? g.nc[0] : (x2 >= g.nc[0] ? -g.nc[0] : 0);
                for (int iy = 0; iy < st.noff[1]; iy++) {
                    int y2 = cy + st.off[1][iy];
                    y2 += y2 < 0 ? g.nc[1] : (y2 >= g.nc[1] ? -g.nc[1] : 0);
                    for (int iz = 0; iz < st.noff[2]; iz++) {
                        int z2 = cz + st.off[2][iz];
                        z2 += z2 < 0 ? g.nc[2] : (z2 >= g.nc[2] ? -g.nc[2] : 0);
                        uint32_t c2 = ((uint32_t)x2 * (uint32_t)g.nc[1] + (uint32_t)y2) * (uint32_t)g.nc[2] + (uint32_t)z2;
                        uint32_t jb = cell_start[c2], je = cell_start[c2 + 1];
                        for (uint32_t j = jb; j < je; j++) {
                            const double4 pj = pos[j];
                            const double dj = diam[j];
                            if (j == s || !(dj >= 0.0)) continue;
                            // box->diff(a1->x, a2->x): remainder(r1 - r2, L) per component (box.hpp:69-72,103)
                            double rx = min_image_exact(__dsub_rn(pi.x, pj.x), box.L[0], box.invL[0], box.halfL[0]);
                            double ry = min_image_exact(__dsub_rn(pi.y, pj.y), box.L[1], box.invL[1], box.halfL[1]);
                            double rz = min_image_exact(__dsub_rn(pi.z, pj.z), box.L[2], box.invL[2], box.halfL[2]);
                            // .norm(): sqrt(e0 + (e1 + e2)), no contraction
                            double dsq = __dadd_rn(__dmul_rn(rx, rx), __dadd_rn(__dmul_rn(ry, ry), __dmul_rn(rz, rz)));
                            // flt diam = (diameters[i] + diameters[j]) / 2;  ... < (diam + skin)
                            double thr = __dadd_rn(__dmul_rn(__dadd_rn(di, dj), 0.5), skin);
                            if (__dsqrt_rn(dsq) < thr) {
                                if (count < kmax) col[(size_t)count * PARM_TILE] = j;
                                count++;
                            }
                        }
                    }
                }
            }
        }
        cnt[s] = count;
    }
    // totals: integer atomics are order independent, so the result is deterministic
    unsigned long long wsum = count;
    uint32_t wmax = count;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
        wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    }
    if ((threadIdx.x & 31) == 0 && wmax) {
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
    nl->h_diam.assign(c->n, -1.0);
    size_t np = c->npad;
    CK(cudaMalloc(&nl->d_diam_id, np * 8));
    CK(cudaMalloc(&nl->d_diam, np * 8));
    CK(cudaMalloc(&nl->xlast, 3 * np * 8));
    CK(cudaMemsetAsync(nl->xlast, 0, 3 * np * 8, c->stream));
    CK(cudaMalloc(&nl->cell_id, np * 4));
    CK(cudaMalloc(&nl->cell_id_sorted, np * 4));
    CK(cudaMalloc(&nl->perm, np * 4));
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
    // all atoms start as non-members
    std::vector<double> neg(np, -1.0);
    CK(cudaMemcpyAsync(nl->d_diam_id, neg.data(), np * 8, cudaMemcpyHostToDevice, c->stream));
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
    void *ptrs[] = {nl->d_diam_id, nl->d_diam, nl->xlast, nl->cell_id, nl->cell_id_sorted, nl->perm, nl->iota,
                    nl->cell_start, nl->sort_temp, nl->nbr, nl->cnt, nl->d_top2, nl->d_counter, nl->d_flags};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (nl->h_flags) cudaFreeHost(nl->h_flags);
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
    double maxd = 0;
    bool any = false;
    for (uint32_t i = 0; i < c->n; i++) {
        double d = diam[i];
        if (d >= 0 && !isinf(d)) {
            nl->h_diam[i] = d;
            any = true;
            if (d > maxd) maxd = d;
        } else
            nl->h_diam[i] = -1.0;
    }
    nl->have_diam = any;
    nl->maxdiam = maxd;
    if (c->n) {
        CK(cudaMemcpyAsync(nl->d_diam_id, nl->h_diam.data(), (size_t)c->n * 8, cudaMemcpyHostToDevice, c->stream));
        k_gather_by_order_d<<<grid_for(c, c->n, 256), 256, 0, c->stream>>>(nl->d_diam_id, c->order, c->n, nl->d_diam);
        CK_LAUNCH(c);
        CK(cudaStreamSynchronize(c->stream)); // h_diam may be re-written by the caller's next call
    }
    nl->ignorechanged = true; // trackers.hpp:200
    return 0;
}

static int alloc_nbr(parm_nlist *nl, uint32_t kmax) {
    parm_ctx *c = nl->ctx;
    size_t ntiles = c->npad / PARM_TILE;
    size_t need = ntiles * (size_t)kmax * PARM_TILE;
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

int parm_nlist_rebuild(parm_nlist *nl) {
    parm_ctx *c = nl->ctx;
    CK(cudaSetDevice(c->device));
    if (!c->box_set) { parm_set_error("NeighborList update before the box was set"); return PARM_ERR_INVALID; }
    nl->updatenum++;          // trackers.cpp:56-57
    nl->ignorechanged = false;
    nl->rebuilds++;
    const uint32_t n = c->n;
    if (n == 0) { nl->total_full = 0; nl->maxcnt = 0; return 0; }

    // --- cell grid: cells no smaller than the largest possible pair threshold
    const double rlist = (nl->maxdiam + nl->skin) * (1.0 + 1e-9) + 1e-12;
    GridDev g;
    StencilDev st;
    uint64_t ncell = 1;
    const uint64_t cell_cap = std::max<uint64_t>(64, 4ull * n);
    for (int d = 0; d < 3; d++) {
        int k = 1;
        if (d < c->D) {
            double q = floor(c->box.L[d] / rlist);
            k = q < 1 ? 1 : (q > 1024 ? 1024 : (int)q);
        }
        g.nc[d] = k;
    }
    // keep the cell table within a few entries per atom
    for (;;) {
        ncell = (uint64_t)g.nc[0] * g.nc[1] * g.nc[2];
        if (ncell <= cell_cap) break;
        int big = 0;
        for (int d = 1; d < 3; d++)
            if (g.nc[d] > g.nc[big]) big = d;
        g.nc[big] = (g.nc[big] + 1) / 2;
    }
    for (int d = 0; d < 3; d++) {
        g.scale[d] = g.nc[d] / c->box.L[d];
        int k = g.nc[d];
        if (k >= 3) { st.noff[d] = 3; st.off[d][0] = -1; st.off[d][1] = 0; st.off[d][2] = 1; }
        else if (k == 2) { st.noff[d] = 2; st.off[d][0] = 0; st.off[d][1] = 1; st.off[d][2] = 0; }
        else { st.noff[d] = 1; st.off[d][0] = 0; st.off[d][1] = 0; st.off[d][2] = 0; }
        nl->nc[d] = k;
    }
    nl->ncell = (uint32_t)ncell;
    if (nl->ncell + 2 > nl->cell_start_cap) {
        if (nl->cell_start) cudaFree(nl->cell_start);
        nl->cell_start = 0;
        nl->cell_start_cap = nl->ncell + 2 + nl->ncell / 4;
        CK(cudaMalloc(&nl->cell_start, (size_t)nl->cell_start_cap * 4));
    }

    PTRY(parm_prof_begin(c, PARM_PROF_REBUILD));
    // --- bin, sort by cell index (stable LSD radix sort), re-order every per-slot array
    k_cell_id<<<grid_for(c, n, 256), 256, 0, c->stream>>>(c->pos, n, c->box, g, nl->cell_id, nl->iota);
    CK_LAUNCH(c);
    int end_bit = 1;
    while ((1ull << end_bit) < ncell) end_bit++;
    size_t need = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, need, nl->cell_id, nl->cell_id_sorted, nl->iota, nl->perm, (int)n, 0,
                                        end_bit, c->stream));
    if (need > nl->sort_temp_bytes) {
        if (nl->sort_temp) cudaFree(nl->sort_temp);
        nl->sort_temp = 0;
        CK(cudaMalloc(&nl->sort_temp, need + need / 8 + 256));
        nl->sort_temp_bytes = need + need / 8 + 256;
    }
    size_t tb = nl->sort_temp_bytes;
    CK(cub::DeviceRadixSort::SortPairs(nl->sort_temp, tb, nl->cell_id, nl->cell_id_sorted, nl->iota, nl->perm, (int)n, 0,
                                        end_bit, c->stream));
    parm_count_launch(c, 3);
    k_permute<<<grid_for(c, n, 256), 256, 0, c->stream>>>(nl->perm, n, c->npad, c->pos, c->v, c->a, c->f, c->order,
                                                          c->pos_alt, c->v_alt, c->a_alt, c->f_alt, c->order_alt,
                                                          c->slot_of, nl->d_diam_id, nl->d_diam, nl->xlast);
    CK_LAUNCH(c);
    std::swap(c->pos, c->pos_alt);
    std::swap(c->v, c->v_alt);
    std::swap(c->a, c->a_alt);
    std::swap(c->f, c->f_alt);
    std::swap(c->order, c->order_alt);
    for (parm_inter *it : c->inters) PTRY(parm_inter_regather(it));
    k_cell_start<<<grid_for(c, n + 1, 256), 256, 0, c->stream>>>(nl->cell_id_sorted, n, nl->ncell, nl->cell_start);
    CK_LAUNCH(c);

    // --- build, growing the per-atom capacity if some row overflowed
    if (nl->kmax == 0) {
        double vol = 1;
        for (int d = 0; d < c->D; d++) vol *= c->box.L[d];
        double rl = nl->maxdiam + nl->skin;
        double sphere = c->D == 3 ? 4.18879020478639 * rl * rl * rl : 3.14159265358979 * rl * rl;
        double est = (double)n / vol * sphere;
        uint32_t k0 = (uint32_t)std::min<double>(std::max(16.0, 1.3 * est + 16.0), (double)std::max<uint32_t>(n, 2u) - 1.0);
        PTRY(alloc_nbr(nl, std::max<uint32_t>(k0, 1u)));
    }
    for (int attempt = 0; attempt < 8; attempt++) {
        CK(cudaMemsetAsync(nl->d_flags, 0, sizeof(NlistFlags), c->stream));
        k_build<<<(n + 127) / 128, 128, 0, c->stream>>>(c->pos, nl->d_diam, nl->cell_id_sorted, nl->cell_start, n, c->box, g,
                                                       st, nl->skin, nl->kmax, nl->nbr, nl->cnt, nl->d_flags);
        CK_LAUNCH(c);
        if (attempt == 0) PTRY(parm_prof_end(c));
        CK(cudaMemcpyAsync(nl->h_flags, nl->d_flags, sizeof(NlistFlags), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        nl->total_full = nl->h_flags->total;
        nl->maxcnt = nl->h_flags->maxcnt;
        if (nl->maxcnt <= nl->kmax) return 0;
        uint32_t k2 = nl->maxcnt + nl->maxcnt / 8 + 8;
        PTRY(alloc_nbr(nl, k2));
    }
    parm_set_error("neighbour list capacity did not converge");
    return PARM_ERR_RUNTIME;
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
        if (c->n == 0) return 0;
        PTRY(parm_nlist_drift_check_async(nl));
        CK(cudaStreamSynchronize(c->stream));
        if (!nl->h_flags->need_rebuild) return 0;
    }
    PTRY(parm_nlist_rebuild(nl));
    if (rebuilt) *rebuilt = 1;
    return 0;
}

extern "C" int parm_nlist_which(parm_nlist *nl, uint32_t *u) { *u = nl->updatenum; return 0; }
extern "C" int parm_nlist_numpairs(parm_nlist *nl, uint64_t *np) { *np = nl->total_full / 2; return 0; }

extern "C" int parm_nlist_stats(parm_nlist *nl, double *mean_full, uint32_t *max_full) {
    uint32_t members = 0;
    for (double d : nl->h_diam) members += d >= 0;
    if (mean_full) *mean_full = members ? (double)nl->total_full / members : 0.0;
    if (max_full) *max_full = nl->maxcnt;
    return 0;
}

extern "C" int parm_nlist_download_pairs(parm_nlist *nl, uint32_t *first, uint32_t *last, uint64_t cap) {
    parm_ctx *c = nl->ctx;
    CK(cudaSetDevice(c->device));
    uint64_t np = nl->total_full / 2;
    if (cap < np) { parm_set_error("parm_nlist_download_pairs: capacity %llu < %llu pairs", (unsigned long long)cap, (unsigned long long)np); return PARM_ERR_INVALID; }
    if (np == 0 || nl->updatenum == 0) return 0;
    const uint32_t n = c->n;
    size_t ntiles = c->npad / PARM_TILE;
    std::vector<uint32_t> h_cnt(c->npad), h_order(c->npad), h_slot(c->npad), h_nbr(ntiles * nl->kmax * PARM_TILE);
    CK(cudaMemcpyAsync(h_cnt.data(), nl->cnt, (size_t)c->npad * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(h_order.data(), c->order, (size_t)c->npad * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(h_slot.data(), c->slot_of, (size_t)c->npad * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(h_nbr.data(), nl->nbr, h_nbr.size() * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    uint64_t k = 0;
    std::vector<uint32_t> js;
    for (uint32_t i = 0; i < n; i++) { // reference order: i ascending, j < i ascending (trackers.cpp:59-68)
        uint32_t s = h_slot[i];
        js.clear();
        const uint32_t *col = h_nbr.data() + ((size_t)(s >> 5) * nl->kmax) * PARM_TILE + (s & 31u);
        for (uint32_t q = 0; q < h_cnt[s]; q++) {
            uint32_t j = h_order[col[(size_t)q * PARM_TILE]];
            if (j < i) js.push_back(j);
        }
        std::sort(js.begin(), js.end());
        for (uint32_t j : js) {
            if (k >= np) { parm_set_error("parm_nlist_download_pairs: device list is not symmetric"); return PARM_ERR_RUNTIME; }
            first[k] = i;
            last[k] = j;
            k++;
        }
    }
    if (k != np) { parm_set_error("parm_nlist_download_pairs: device list is not symmetric (%llu vs %llu)", (unsigned long long)k, (unsigned long long)np); return PARM_ERR_RUNTIME; }
    return 0;
}
