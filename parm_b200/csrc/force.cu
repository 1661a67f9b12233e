// NListed<A,P> on the device: host side of the pair force / energy / virial / stress loop
// (interaction.hpp:2102-2291). The pair functors are in pairs.cuh, the kernel in force_kernel.cuh:
//   LJRepulsePair :875-891            RepulsionPair :1528-1550          LJAttractRepulsePair :1251-1299
//   LennardJonesCutPair :967-987      LJAttractCutPair :1020-1049       LJAttractFixedRepulsePair :1343-1413
//   EisMclachlanPair :1425-1452       LJishPair :1095-1142              LJAttractRepulseSigsPair :1171-1244
//   RepulsionDragPair :1613-1642      LoisOhernPair(+MinCLs) :1693-1752 LoisLinPair(+Min) :1782-1837
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <array>
#include <map>

#include "force_tile.cuh"

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

typedef cudaError_t (*force_launcher)(int, int, int, uint32_t, size_t, cudaStream_t, const ForceArgs &);
#define DECL(K) extern template cudaError_t parm_launch_force_kind<K>(int, int, int, uint32_t, size_t, cudaStream_t, const ForceArgs &);
DECL(PARM_PAIR_LJREPULSE) DECL(PARM_PAIR_REPULSION) DECL(PARM_PAIR_LJATTRACTREPULSE) DECL(PARM_PAIR_LJCUT)
DECL(PARM_PAIR_LJATTRACTCUT) DECL(PARM_PAIR_LJATTRACTFIXEDREPULSE) DECL(PARM_PAIR_EISMCLACHLAN) DECL(PARM_PAIR_LJISH)
DECL(PARM_PAIR_LJATTRACTREPULSESIGS) DECL(PARM_PAIR_REPULSIONDRAG) DECL(PARM_PAIR_LOISOHERN) DECL(PARM_PAIR_LOISLIN)
#undef DECL
static const force_launcher k_launchers[PARM_NKERNEL_KINDS] = {
    parm_launch_force_kind<PARM_PAIR_LJREPULSE>, parm_launch_force_kind<PARM_PAIR_REPULSION>,
    parm_launch_force_kind<PARM_PAIR_LJATTRACTREPULSE>, parm_launch_force_kind<PARM_PAIR_LJCUT>,
    parm_launch_force_kind<PARM_PAIR_LJATTRACTCUT>, parm_launch_force_kind<PARM_PAIR_LJATTRACTFIXEDREPULSE>,
    parm_launch_force_kind<PARM_PAIR_EISMCLACHLAN>, parm_launch_force_kind<PARM_PAIR_LJISH>,
    parm_launch_force_kind<PARM_PAIR_LJATTRACTREPULSESIGS>, parm_launch_force_kind<PARM_PAIR_REPULSIONDRAG>,
    parm_launch_force_kind<PARM_PAIR_LOISOHERN>, parm_launch_force_kind<PARM_PAIR_LOISLIN>};

typedef cudaError_t (*tile_launcher)(int, int, int, unsigned, size_t, cudaStream_t, const TileForceArgs &);
#define DECL(K) extern template cudaError_t parm_launch_force_tile_kind<K>(int, int, int, unsigned, size_t, cudaStream_t, const TileForceArgs &);
DECL(PARM_PAIR_LJREPULSE) DECL(PARM_PAIR_LJATTRACTREPULSE) DECL(PARM_PAIR_LJCUT)
#undef DECL
static tile_launcher tile_launcher_of(int kernel_kind) {
    switch (kernel_kind) {
        case PARM_PAIR_LJREPULSE: return parm_launch_force_tile_kind<PARM_PAIR_LJREPULSE>;
        case PARM_PAIR_LJATTRACTREPULSE: return parm_launch_force_tile_kind<PARM_PAIR_LJATTRACTREPULSE>;
        case PARM_PAIR_LJCUT: return parm_launch_force_tile_kind<PARM_PAIR_LJCUT>;
        default: return nullptr;
    }
}

enum { RUN_F = 0, RUN_FALL = 1, RUN_OBS = 2 }; // forces only / forces + observables / observables only

// d_out: device pointer to NPART doubles (E, virial, stress[9], contacts, overlaps) or NULL
static int launch_forces(parm_inter *it, int run, bool accumulate, double *d_out, const int *abort_flag = nullptr,
                         uint32_t first = 0, uint32_t count = 0xffffffffu) {
    parm_ctx *c = it->ctx;
    parm_nlist *nl = it->nl;
    if (!it->have_params || nl->updatenum == 0) {
        // no atoms add()ed yet, or the list was never built: the reference iterates an empty pair vector
        if (run != RUN_OBS && !accumulate) CK(cudaMemsetAsync(c->f, 0, 3 * (size_t)c->npad * 8, c->stream));
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
        if (forced == 4 || forced == 8) team = forced;
    }
    uint32_t nown = parm_owned(c); // rows exist for owned atoms only
    if (first > nown) first = nown;
    nown = count == 0xffffffffu ? nown : std::min(nown, first + count); // slots [first, nown)
    const uint32_t nrange = nown - first;
    uint32_t nblocks = (uint32_t)(((size_t)nrange * team + F_BLOCK - 1) / F_BLOCK);
    if (nrange == 0) {
        if (d_out) CK(cudaMemsetAsync(d_out, 0, NPART * 8, c->stream));
        return 0;
    }
    const int mode = run == RUN_F ? MODE_F : MODE_FALL;
    // cell-tile kernel (force_tile.cuh) when the list carries tile-local rows for this interaction and the slot
    // range is made of whole cell columns; otherwise the gather kernel below
    uint32_t chunk0 = 0, chunk1 = 0;
    tile_launcher tl = parm_tile_usable(it) ? tile_launcher_of(PARM_KERNEL_KIND(it->kind)) : nullptr;
    if (tl && !parm_tile_chunk_range(nl, first, nown, &chunk0, &chunk1)) tl = nullptr;
    if (tl && chunk1 == chunk0) tl = nullptr;
    // Small systems (BASELINE config 1, N = 1000): a step is a chain of dependent kernel latencies, not throughput. The
    // tile kernel would run one block per chunk on a handful of SMs (9 chunks at N = 1000) and four lanes would walk a
    // row of ~100 entries in 13 dependent trips; 16 lanes per atom of the gather kernel spread the same rows over every
    // SM and finish in 4 trips.
    {
        const char *est = getenv("PARM_B200_SMALL_TEAM"); // (read per launch: the tests of the tile kernel's small-box path turn it off)
        const int small_team = est ? atoi(est) : 16;
        const bool one_species = !it->generic && it->nspecies == 1;
        const uint32_t nch = tl ? chunk1 - chunk0 : (nrange + 119u) / 120u;
        // (single-GPU contexts only: the slot ranges of a slab-decomposed step keep the kernels they were validated with)
        if (small_team == 16 && !c->sh.on && one_species && nch < (uint32_t)c->num_sms && nl->total_full >= 16ull * nrange) {
            team = 16;
            tl = nullptr;
            nblocks = (uint32_t)(((size_t)nrange * team + F_BLOCK - 1) / F_BLOCK);
        }
    }
    if (tl) nblocks = chunk1 - chunk0;
    if (mode != MODE_F && (size_t)nblocks * NPART > it->partial_doubles) {
        if (it->d_partials) cudaFree(it->d_partials);
        it->d_partials = 0;
        it->partial_doubles = (size_t)nblocks * NPART;
        CK(cudaMalloc(&it->d_partials, it->partial_doubles * 8));
    }
    if (tl) {
        TileForceArgs T;
        T.pos = c->pos;
        const bool bulk = nl->tile.stage_aligned;
        T.prel_xy = bulk ? nl->tile.prel_xy : nullptr;
        T.prel_z = bulk ? nl->tile.prel_z : nullptr;
        // prel of every slot (ghost copies included) from the current positions, unless the caller has done it
        if (bulk && !c->tile_prep_external) PTRY(parm_tile_prep(nl, 0, c->n, c->stream, abort_flag));
        // two buffers of 3 * cap doubles per block, two blocks per SM: the tile must fit a quarter of the shared memory
        const char *ep = getenv("PARM_B200_TILE_PERS"); // (read per launch: the sweeps toggle it inside one process)
        // (measured at N = 1e6: 0.250 ms against 0.243 ms for the one-block-per-chunk kernel -- two tile buffers per block
        // leave the compute warps at most one chunk of slack; off by default)
        T.cap = ((nl->tile.max_tile + 1 + 15u) & ~15u) + 16u; // tile + the single sentinel, rounded up, + 16 class sentinels
        const bool pers = bulk && (ep ? atoi(ep) != 0 : false) && 2 * 3 * (size_t)T.cap * 8 + 2048 <= 113 * 1024 && nl->tile.ch <= 120;
        T.pers_blocks = pers ? 2u * (uint32_t)c->num_sms : 0u;
        {   // measurement switches (read per launch like the one above)
            const char *eh = getenv("PARM_B200_TILE_HALF"), *ef = getenv("PARM_B200_TILE_PF");
            T.half_ok = (eh ? atoi(eh) != 0 : true) && !nl->tile.banked ? 1u : 0u;
            T.pf_dist = ef ? (uint32_t)atoi(ef) : 4u * (uint32_t)c->num_sms; // one wave of blocks ahead
        }
        T.chunk_s0 = nl->tile.d_s0 + chunk0;
        T.chunks = nl->tile.d_chunks + chunk0;
        T.rows16 = nl->tile.rows16;
        T.cnt = nl->cnt;
        T.kmax = nl->kmax;
        T.cap = ((nl->tile.max_tile + 1 + 15u) & ~15u) + 16u; // tile + the single sentinel, rounded up, + 16 class sentinels
        T.P1 = it->h_table[0];
        T.f = c->f;
        T.npad = c->npad;
        T.box = c->box;
        T.accumulate = accumulate ? 1 : 0;
        T.store = run != RUN_OBS;
        T.partials = it->d_partials;
        T.abort_flag = abort_flag;
        cudaError_t e = tl(nl->tile.team, nl->tile.v, mode, nblocks, 3 * (size_t)T.cap * 8, c->stream, T);
        parm_count_launch(c);
        if (e != cudaSuccess) {
            parm_set_error("cell-tile pair kernel launch failed: %s (chunks %u..%u, tile capacity %u atoms = %zu bytes of shared memory, "
                           "bulk staging %d)", cudaGetErrorString(e), chunk0, chunk1, T.cap, 3 * (size_t)T.cap * 8, (int)bulk);
            return PARM_ERR_CUDA;
        }
        if (mode != MODE_F) {
            if (!d_out) { parm_set_error("internal: observables requested without an output buffer"); return PARM_ERR_RUNTIME; }
            k_force_fold<<<1, 256, 0, c->stream>>>(it->d_partials, nblocks, d_out);
            CK_LAUNCH(c);
        }
        return 0;
    }
    PTRY(parm_nlist_ensure_rows32(nl)); // gather kernel: needs the expanded 32-bit rows (mask-mode lists build them on demand)
    // 0 one species, 1 species table in shared memory, 2 per-atom parameters, 3 two species in registers
    const bool long_rows = nl->total_full >= (uint64_t)PARM_PACK_MIN_NEIGHBORS * (parm_owned(c) ? parm_owned(c) : 1);
    const int specmode = it->generic ? 2 : (it->nspecies == 1 ? 0 : (it->nspecies == 2 && long_rows ? 3 : 1));
    size_t smem = specmode == 1 ? (size_t)it->nspecies * it->nspecies * sizeof(PairConst) : 0;
    ForceArgs A;
    A.pos = c->pos;
    A.nbr = nl->nbr;
    A.cnt = nl->cnt;
    A.kmax = nl->kmax;
    A.spec = it->d_spec;
    A.table = it->d_table;
    A.nspecies = it->nspecies;
    A.P1 = it->h_table[0];
    if (specmode == 3)
        for (int q = 0; q < 4; q++) A.P4[q] = it->h_table[q];
    A.mask = PARM_NBR_MASK_OF(nl);
    A.packed = nl->packed_for == it && !it->spec_stale;
    A.f = c->f;
    A.vel = c->v;
    A.n = nown;
    A.npad = c->npad;
    A.box = c->box;
    A.accumulate = accumulate ? 1 : 0;
    A.store = run != RUN_OBS;
    A.partials = it->d_partials;
    A.par = it->d_par;
    A.eps_tab = it->d_eps_table;
    A.sig_tab = it->d_sig_table;
    A.ntypes = it->ntypes;
    A.minmix = it->minmix ? 1 : 0;
    A.abort_flag = abort_flag;
    A.first = first;
    cudaError_t e = k_launchers[PARM_KERNEL_KIND(it->kind)](specmode, team, mode, nrange, smem, c->stream, A);
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
    return launch_forces(it, want ? RUN_FALL : RUN_F, accumulate, d_out, abort_flag, first, count);
}

// energy / virial / stress / contact counts into device memory, no host round trip (EnergyTracker, trackers.cu)
int parm_inter_launch_obs(parm_inter *it, double *d_out, const int *abort_flag) {
    return launch_forces(it, RUN_OBS, false, d_out, abort_flag);
}

// ---- host API ---------------------------------------------------------------------------
static const char *k_kind_names[PARM_PAIR_NKINDS] = {
    "LJRepulsePair", "RepulsionPair", "LJAttractRepulsePair", "LennardJonesCutPair", "LJAttractCutPair",
    "LJAttractFixedRepulsePair", "EisMclachlanPair", "LJishPair", "LJAttractRepulseSigsPair", "RepulsionDragPair",
    "LoisOhernPair", "LoisLinPair", "LoisOhernPairMinCLs", "LoisLinPairMin"};

extern "C" int parm_inter_create(parm_ctx *c, parm_nlist *nl, int kind, parm_inter **out) {
    if (!c || !nl || !out) { parm_set_error("parm_inter_create: NULL argument"); return PARM_ERR_INVALID; }
    *out = 0;
    if (nl->ctx != c) { parm_set_error("parm_inter_create: NeighborList belongs to another AtomVec"); return PARM_ERR_INVALID; }
    if (kind < 0 || kind >= PARM_PAIR_NKINDS) {
        parm_set_error("parm_inter_create: pair type %d is not an NListed functor of this library (0..%d, include/parm_b200.h)",
                       kind, PARM_PAIR_NKINDS - 1);
        return PARM_ERR_UNSUPPORTED;
    }
    if (kind == PARM_PAIR_REPULSIONDRAG && c->sh.on) {
        parm_set_error("RepulsionDragPair needs neighbour velocities; ghost atoms of a slab decomposition carry none");
        return PARM_ERR_UNSUPPORTED;
    }
    CK(cudaSetDevice(c->device));
    parm_inter *it = new parm_inter();
    it->ctx = c;
    it->nl = nl;
    it->kind = kind;
    it->minmix = kind == PARM_PAIR_LOISOHERNMIN || kind == PARM_PAIR_LOISLINMIN;
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
    if (it->d_sig_table) cudaFree(it->d_sig_table);
    c->inters.erase(std::remove(c->inters.begin(), c->inters.end(), it), c->inters.end());
    if (it->nl->packed_for == it) it->nl->packed_for = nullptr; // (the entries stay packed until the next rebuild)
    delete it;
    return 0;
}

// both halves of the per-atom parameter block: par[s] = par_id[order[s]], par[npad + s] = par_id[nid_stride + order[s]]
__global__ void k_gather_par(const double4 *__restrict__ par_id, size_t id_stride, const uint32_t *__restrict__ order, uint32_t n,
                             double4 *par, size_t slot_stride) {
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        const uint32_t id = order[s];
        par[s] = par_id[id];
        par[slot_stride + s] = par_id[id_stride + id];
    }
}
__global__ void k_gather_spec(const uint8_t *__restrict__ spec_id, const uint32_t *__restrict__ order, uint32_t n, uint8_t *spec) {
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) spec[s] = spec_id[order[s]];
}

int parm_inter_regather(parm_inter *it) {
    parm_ctx *c = it->ctx;
    if (!c->n) return 0;
    unsigned nb = (c->n + 255) / 256;
    unsigned cap = (unsigned)c->num_sms * 8;
    if (it->generic)
        k_gather_par<<<nb < cap ? nb : cap, 256, 0, c->stream>>>(it->d_par_id, std::max(c->npad, c->nid_pad), c->order, c->n, it->d_par, c->npad);
    else k_gather_spec<<<nb < cap ? nb : cap, 256, 0, c->stream>>>(it->d_spec_id, c->order, c->n, it->d_spec);
    CK_LAUNCH(c);
    return 0;
}

// The reference's pair constructors, evaluated once per species pair on the host (pairs.cuh).
static PairConst mix(int kind, const double *p1, int t1, const double *p2, int t2, const double *et, const double *st, int nt) {
    switch (kind) {
        case PARM_PAIR_LJREPULSE: return mix_pair<PARM_PAIR_LJREPULSE>(p1, t1, p2, t2, et, st, nt, false, true);
        case PARM_PAIR_REPULSION: return mix_pair<PARM_PAIR_REPULSION>(p1, t1, p2, t2, et, st, nt, false, true);
        case PARM_PAIR_LJATTRACTREPULSE: return mix_pair<PARM_PAIR_LJATTRACTREPULSE>(p1, t1, p2, t2, et, st, nt, false, true);
        case PARM_PAIR_LJCUT: return mix_pair<PARM_PAIR_LJCUT>(p1, t1, p2, t2, et, st, nt, false, true);
        case PARM_PAIR_LJATTRACTCUT: return mix_pair<PARM_PAIR_LJATTRACTCUT>(p1, t1, p2, t2, et, st, nt, false, true);
        case PARM_PAIR_LJATTRACTFIXEDREPULSE: return mix_pair<PARM_PAIR_LJATTRACTFIXEDREPULSE>(p1, t1, p2, t2, et, st, nt, false, true);
        case PARM_PAIR_EISMCLACHLAN: return mix_pair<PARM_PAIR_EISMCLACHLAN>(p1, t1, p2, t2, et, st, nt, false, true);
        case PARM_PAIR_LJISH: return mix_pair<PARM_PAIR_LJISH>(p1, t1, p2, t2, et, st, nt, false, true);
        case PARM_PAIR_LJATTRACTREPULSESIGS: return mix_pair<PARM_PAIR_LJATTRACTREPULSESIGS>(p1, t1, p2, t2, et, st, nt, false, true);
        case PARM_PAIR_REPULSIONDRAG: return mix_pair<PARM_PAIR_REPULSIONDRAG>(p1, t1, p2, t2, et, st, nt, false, true);
        case PARM_PAIR_LOISOHERN: return mix_pair<PARM_PAIR_LOISOHERN>(p1, t1, p2, t2, et, st, nt, false, true);
        case PARM_PAIR_LOISOHERNMIN: return mix_pair<PARM_PAIR_LOISOHERN>(p1, t1, p2, t2, et, st, nt, true, true);
        case PARM_PAIR_LOISLIN: return mix_pair<PARM_PAIR_LOISLIN>(p1, t1, p2, t2, et, st, nt, false, true);
        default: return mix_pair<PARM_PAIR_LOISLIN>(p1, t1, p2, t2, et, st, nt, true, true);
    }
}

// A::max_size(): what NListed::add hands to NeighborList::add (interaction.hpp:1906-1910)
static double max_size(int kind, const double *p, uint32_t t, const double *sig_table, int nt) {
    double sigma = p[1];
    if (sig_table) { // IEpsISigCutAtom::max_size :953-959, IEpsISigExpAtom::max_size :1517-1523: largest of `sigmas`
        sigma = sig_table[(size_t)t * nt];
        for (int k = 1; k < nt; k++) sigma = std::max(sigma, sig_table[(size_t)t * nt + k]);
    }
    switch (kind) {
        case PARM_PAIR_LJREPULSE:
        case PARM_PAIR_REPULSION:
        case PARM_PAIR_REPULSIONDRAG: return sigma;                       // :864, :1464, :1610
        case PARM_PAIR_EISMCLACHLAN: return p[1];                         // dist :1422
        case PARM_PAIR_LJATTRACTREPULSESIGS: return p[1] + p[4] * (p[2] - 1); // :1168
        case PARM_PAIR_LOISOHERN:
        case PARM_PAIR_LOISOHERNMIN: return p[1] * (1 + p[2] + p[3]);     // :1690
        case PARM_PAIR_LOISLIN:
        case PARM_PAIR_LOISLINMIN: return p[1] * (1 + p[3]);              // :1779
        default: return sigma * p[2];                                     // sigma*sigcut :905, :1017, :1091, :1340
    }
}

extern "C" int parm_inter_set_params_ex(parm_inter *it, const double *params, int nper, const uint32_t *type,
                                        const double *eps_table, const double *sig_table, int ntypes, const uint8_t *member,
                                        int set_diameters) {
    if (!it || !params) { parm_set_error("parm_inter_set_params: NULL argument"); return PARM_ERR_INVALID; }
    parm_ctx *c = it->ctx;
    CK(cudaSetDevice(c->device));
    const int kind = it->kind, kk = PARM_KERNEL_KIND(kind);
    const char *name = k_kind_names[kind];
    if (nper < parm_nparams(kk) || nper > PARM_PAIR_MAXPARAMS) {
        parm_set_error("%s takes %d parameters per atom, got %d", name, parm_nparams(kk), nper);
        return PARM_ERR_INVALID;
    }
    const bool needs_eps_table = kk == PARM_PAIR_LJATTRACTREPULSE || kk == PARM_PAIR_LJATTRACTFIXEDREPULSE || kk == PARM_PAIR_LJISH;
    const bool may_index_eps = needs_eps_table || kk == PARM_PAIR_REPULSION || kk == PARM_PAIR_LJCUT || kk == PARM_PAIR_LJATTRACTCUT;
    const bool may_index_sig = kk == PARM_PAIR_REPULSION || kk == PARM_PAIR_LJCUT || kk == PARM_PAIR_LJATTRACTCUT;
    if (needs_eps_table && (!eps_table || ntypes < 1)) { parm_set_error("%s needs the epsilon table (the atoms' `epsilons` vectors)", name); return PARM_ERR_INVALID; }
    if (eps_table && !may_index_eps) { parm_set_error("%s does not take an epsilon table", name); return PARM_ERR_INVALID; }
    if (sig_table && (!may_index_sig || !eps_table)) { parm_set_error("%s does not take a sigma table (or it came without an epsilon table)", name); return PARM_ERR_INVALID; }
    if ((eps_table || sig_table) && ntypes < 1) { parm_set_error("%s: ntypes must be >= 1 with indexed parameters", name); return PARM_ERR_INVALID; }
    for (int a = 0; a < ntypes && eps_table; a++)
        for (int b = 0; b < ntypes; b++)
            if (eps_table[a * ntypes + b] != eps_table[b * ntypes + a] ||
                (sig_table && sig_table[a * ntypes + b] != sig_table[b * ntypes + a])) { // assert at interaction.hpp:1013-1014
                parm_set_error("%s: epsilon / sigma tables must be symmetric", name);
                return PARM_ERR_INVALID;
            }
    const bool indexed = eps_table != nullptr;
    const unsigned geo = parm_geo_mask(kk, indexed);
    typedef std::array<double, PARM_PAIR_MAXPARAMS + 1> Key;
    bool too_many = false;
    std::map<Key, int> ids;
    std::vector<Key> keys;
    std::vector<double> diam(c->nid, -1.0);
    // fields the functor does not read are zeroed so that they cannot split species
    auto canon = [&](uint32_t i, Key &k) {
        const double *p = params + (size_t)nper * i;
        for (int q = 0; q < PARM_PAIR_MAXPARAMS; q++) k[q] = q < parm_nparams(kk) ? p[q] : 0.0;
        if (indexed && kk != PARM_PAIR_LJATTRACTFIXEDREPULSE && kk != PARM_PAIR_LJISH) k[0] = 0.0; // eps comes from the table
        if (indexed && (kk == PARM_PAIR_LJATTRACTFIXEDREPULSE || kk == PARM_PAIR_LJISH)) k[0] = 0.0;
        if (sig_table) k[1] = 0.0;
        k[PARM_PAIR_MAXPARAMS] = indexed && type ? (double)type[i] : 0.0;
    };
    for (uint32_t i = 0; i < c->nid; i++) {
        if (member && !member[i]) { it->h_spec_id[i] = 0; continue; }
        const uint32_t t = indexed && type ? type[i] : 0;
        if (indexed && (int)t >= ntypes) { parm_set_error("atom %u: type %u >= ntypes %d", i, t, ntypes); return PARM_ERR_INVALID; }
        Key k;
        canon(i, k);
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
        diam[i] = max_size(kind, k.data(), t, sig_table, ntypes);
    }
    it->generic = too_many;
    it->spec_stale = true; // until the next rebuild re-packs the list entries
    it->ntypes = ntypes > 0 ? ntypes : 1;
    if (it->d_eps_table) { cudaFree(it->d_eps_table); it->d_eps_table = 0; }
    if (it->d_sig_table) { cudaFree(it->d_sig_table); it->d_sig_table = 0; }
    if (too_many) {
        keys.resize(1);
        const size_t npar = std::max(c->npad, c->nid_pad);
        it->h_par_id.assign(8 * npar, 0.0);
        for (uint32_t i = 0; i < c->nid; i++) {
            if (member && !member[i]) continue;
            Key k;
            canon(i, k);
            for (int q = 0; q < PARM_PAIR_MAXPARAMS; q++)
                if (geo & (1u << q)) k[q] = sqrt(k[q]); // the device constructor multiplies (pairs.cuh geo_mean)
            double *lo = &it->h_par_id[4 * (size_t)i], *hi = &it->h_par_id[4 * (npar + (size_t)i)];
            lo[0] = k[0]; lo[1] = k[1]; lo[2] = k[2]; lo[3] = k[PARM_PAIR_MAXPARAMS];
            hi[0] = k[3]; hi[1] = k[4];
        }
        if (!it->d_par_id) CK(cudaMalloc(&it->d_par_id, 2 * npar * sizeof(double4)));
        if (!it->d_par) CK(cudaMalloc(&it->d_par, 2 * (size_t)c->npad * sizeof(double4)));
        CK(cudaMemcpyAsync(it->d_par_id, it->h_par_id.data(), 2 * npar * sizeof(double4), cudaMemcpyHostToDevice, c->stream));
        const size_t tb = sizeof(double) * it->ntypes * it->ntypes;
        if (eps_table) {
            CK(cudaMalloc(&it->d_eps_table, tb));
            CK(cudaMemcpyAsync(it->d_eps_table, eps_table, tb, cudaMemcpyHostToDevice, c->stream));
        }
        if (sig_table) {
            CK(cudaMalloc(&it->d_sig_table, tb));
            CK(cudaMemcpyAsync(it->d_sig_table, sig_table, tb, cudaMemcpyHostToDevice, c->stream));
        }
    }
    int S = (int)keys.size();
    if (S == 0) {
        S = 1;
        Key k;
        k.fill(1.0);
        k[PARM_PAIR_MAXPARAMS] = 0.0;
        keys.push_back(k);
    }
    it->nspecies = S;
    it->h_table.assign((size_t)S * S, PairConst());
    for (int a = 0; a < S; a++)
        for (int b = 0; b < S; b++)
            it->h_table[(size_t)a * S + b] = mix(kind, keys[a].data(), (int)keys[a][PARM_PAIR_MAXPARAMS], keys[b].data(),
                                                 (int)keys[b][PARM_PAIR_MAXPARAMS], eps_table, sig_table, ntypes);
    CK(cudaMemcpyAsync(it->d_table, it->h_table.data(), sizeof(PairConst) * S * S, cudaMemcpyHostToDevice, c->stream));
    if (c->nid) CK(cudaMemcpyAsync(it->d_spec_id, it->h_spec_id.data(), c->nid, cudaMemcpyHostToDevice, c->stream));
    PTRY(parm_inter_regather(it));
    CK(cudaStreamSynchronize(c->stream));
    it->have_params = true;
    if (set_diameters) PTRY(parm_nlist_set_diameters(it->nl, diam.data()));
    return 0;
}

extern "C" int parm_inter_set_params(parm_inter *it, const double *params, const uint32_t *type, const double *eps_table,
                                     int ntypes, const uint8_t *member, int set_diameters) {
    if (it && !(it->kind == PARM_PAIR_LJATTRACTREPULSE)) eps_table = nullptr; // the 3-parameter form ignores it elsewhere
    return parm_inter_set_params_ex(it, params, 3, type, eps_table, nullptr, ntypes, member, set_diameters);
}

static int fetch(parm_inter *it, int run, bool accumulate, double *host13) {
    parm_ctx *c = it->ctx;
    CK(cudaSetDevice(c->device));
    PTRY(parm_ctx_ensure_red(c, 64));
    PTRY(launch_forces(it, run, accumulate, c->d_red));
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
    if (!want) return launch_forces(it, RUN_F, true, nullptr);
    if (!out) { parm_set_error("parm_inter_set_forces: out is NULL"); return PARM_ERR_INVALID; }
    double r[NPART];
    PTRY(fetch(it, RUN_FALL, true, r));
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
    PTRY(fetch(it, RUN_OBS, false, r));
    *E = r[0];
    return 0;
}
extern "C" int parm_inter_pressure(parm_inter *it, double *p) {
    double r[NPART];
    PTRY(fetch(it, RUN_OBS, false, r));
    *p = r[1];
    return 0;
}
extern "C" int parm_inter_stress(parm_inter *it, double *st) {
    double r[NPART];
    PTRY(fetch(it, RUN_OBS, false, r));
    const int D = it->ctx->D;
    for (int a = 0; a < D; a++)
        for (int b = 0; b < D; b++) st[a * D + b] = r[2 + a * 3 + b];
    return 0;
}
extern "C" int parm_inter_contacts(parm_inter *it, uint64_t *contacts, uint64_t *overlaps) {
    double r[NPART];
    PTRY(fetch(it, RUN_OBS, false, r));
    if (contacts) *contacts = (uint64_t)llround(r[11]);
    if (overlaps) *overlaps = (uint64_t)llround(r[12]);
    return 0;
}
