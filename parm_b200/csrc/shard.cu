// Slab decomposition across the GPUs of one box: one process per GPU, NCCL over NVLink/NVSwitch.
// New relative to the reference (which is single-core); it distributes exactly the same
// per-step path: each rank owns the atoms whose wrapped x lies in its slab, keeps read-only ghost
// copies of the neighbouring ranks' boundary layers, and
//   * every step, after K1:   sends its two boundary-layer position ranges and receives the two
//                              ghost ranges -- contiguous double4 ranges of the pos array, no
//                              pack/unpack kernels (ncclSend/ncclRecv in one group);
//   * every step, after K3:   all-gathers the per-rank top-2 displacements so that every rank takes
//                              the identical rebuild decision (trackers.cpp:23-53 semantics);
//   * on rebuild:             migrates atoms that left the slab, re-selects ghosts, then runs the
//                              ordinary bin/sort/build over local + ghost atoms.
// Slot layout after a rebuild (cell ids have the slab axis as slowest index with the two halo layers
// numbered last, so a sort by cell id produces it):
//   [owned atoms, lowest layer first ... highest layer last | ghosts from above | ghosts from below].
// Full neighbour rows mean no reverse (force) communication.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "internal.cuh"

namespace {
struct NcclApi {
    void *h;
    decltype(&ncclGetUniqueId) GetUniqueId;
    decltype(&ncclCommInitRank) CommInitRank;
    decltype(&ncclCommDestroy) CommDestroy;
    decltype(&ncclSend) Send;
    decltype(&ncclRecv) Recv;
    decltype(&ncclGroupStart) GroupStart;
    decltype(&ncclGroupEnd) GroupEnd;
    decltype(&ncclAllReduce) AllReduce;
    decltype(&ncclAllGather) AllGather;
    decltype(&ncclGetErrorString) GetErrorString;
};
NcclApi *nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.h ? &api : nullptr;
    tried = true;
    const char *names[] = {getenv("PARM_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        if (!nm) continue;
        api.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (api.h) break;
    }
    if (!api.h) return nullptr;
#define LOAD(sym) api.sym = (decltype(api.sym))dlsym(api.h, "nccl" #sym); if (!api.sym) { api.h = nullptr; return nullptr; }
    LOAD(GetUniqueId) LOAD(CommInitRank) LOAD(CommDestroy) LOAD(Send) LOAD(Recv) LOAD(GroupStart) LOAD(GroupEnd)
    LOAD(AllReduce) LOAD(AllGather) LOAD(GetErrorString)
#undef LOAD
    return &api;
}
}  // namespace

#define NCK(call)                                                                                   \
    do {                                                                                            \
        ncclResult_t r_ = (call);                                                                   \
        if (r_ != ncclSuccess) {                                                                    \
            parm_set_error("NCCL error at %s:%d: %s", __FILE__, __LINE__, nccl_api()->GetErrorString(r_)); \
            return PARM_ERR_CUDA;                                                                   \
        }                                                                                           \
    } while (0)

extern "C" int parm_nccl_unique_id(void *out128) {
    NcclApi *n = nccl_api();
    if (!n) { parm_set_error("libnccl.so.2 could not be loaded (set PARM_B200_NCCL_LIB)"); return PARM_ERR_CUDA; }
    ncclUniqueId id;
    NCK(n->GetUniqueId(&id));
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    memcpy(out128, &id, 128);
    return 0;
}

// ---- migration messages of the one-sort rebuild --------------------------------------------------
// [header 16 B: count, overflow, -, -][pos: cap double4][v: 3 x cap][a: 3 x cap][f: 3 x cap][id: cap u32]; always sent whole
// (ncclSend / ncclRecv need the size on both sides before the counts are known anywhere but on the sender's device)
static inline size_t mig_bytes(uint32_t cap) { return 16 + (size_t)cap * (32 + 9 * 8 + 4); }
struct MigView {
    uint32_t *hdr;
    double4 *pos;
    double *v, *a, *f; // [3][cap]
    uint32_t *id;
};
static __host__ __device__ inline MigView mig_view(char *buf, uint32_t cap) {
    MigView m;
    m.hdr = reinterpret_cast<uint32_t *>(buf);
    m.pos = reinterpret_cast<double4 *>(buf + 16);
    m.v = reinterpret_cast<double *>(buf + 16 + (size_t)cap * 32);
    m.a = m.v + 3 * (size_t)cap;
    m.f = m.a + 3 * (size_t)cap;
    m.id = reinterpret_cast<uint32_t *>(m.f + 3 * (size_t)cap);
    return m;
}

extern "C" int parm_ctx_create_sharded(int ndim, uint32_t n_global, uint32_t cap_slots, int device, int rank, int nranks,
                                       const void *id128, parm_ctx **out) {
    if (nranks < 2 || rank < 0 || rank >= nranks || !id128) { parm_set_error("parm_ctx_create_sharded: bad rank/nranks/id"); return PARM_ERR_INVALID; }
    NcclApi *n = nccl_api();
    if (!n) { parm_set_error("libnccl.so.2 could not be loaded (set PARM_B200_NCCL_LIB)"); return PARM_ERR_CUDA; }
    PTRY(parm_ctx_alloc(ndim, n_global, cap_slots, device, out));
    parm_ctx *c = *out;
    c->n = 0;
    ShardState &sh = c->sh;
    sh.on = true;
    sh.rank = rank;
    sh.nranks = nranks;
    sh.up = (rank + 1) % nranks;
    sh.down = (rank + nranks - 1) % nranks;
    CK(cudaMalloc(&c->ghost, c->npad));
    CK(cudaMalloc(&c->ghost_alt, c->npad));
    CK(cudaMemset(c->ghost, 0, c->npad));
    CK(cudaMemset(c->ghost_alt, 0, c->npad));
    CK(cudaMalloc(&sh.d_counts, 16 * 4));
    CK(cudaMallocHost(&sh.h_counts, 16 * 4));
    CK(cudaMalloc(&sh.d_gather, 2 * 8 * (size_t)nranks));
    CK(cudaMallocHost(&sh.h_gather, 2 * 8 * (size_t)nranks));
    {   // highest priority: the small exchange kernels must get SM slots while the force kernel fills the GPU
        int lo_prio = 0, hi_prio = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
        CK(cudaStreamCreateWithPriority(&sh.comm_stream, cudaStreamNonBlocking, hi_prio));
    }
    {
        const char *e = getenv("PARM_B200_SHARD_FAST");
        sh.mig_fast = e ? atoi(e) : 1;
        e = getenv("PARM_B200_SHARD_MIGCAP");
        sh.mig_cap = e ? (uint32_t)std::max(1, atoi(e)) : 4096u;
        CK(cudaMalloc(&sh.mig_list, 2 * (size_t)sh.mig_cap * 4));
        CK(cudaMalloc(&sh.mig_cnt, 4 * 4));
        for (int d = 0; d < 2; d++) {
            CK(cudaMalloc(&sh.mig_send[d], mig_bytes(sh.mig_cap)));
            CK(cudaMalloc(&sh.mig_recv[d], mig_bytes(sh.mig_cap)));
        }
    }
    CK(cudaEventCreateWithFlags(&sh.ev_k1, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&sh.ev_comm, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&sh.ev_rows, cudaEventDisableTiming));
    sh.main_stream = c->stream;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t comm;
    NCK(n->CommInitRank(&comm, nranks, id, rank));
    sh.comm = comm;
    return 0;
}

// ---- dense host arrays <-> local slots -------------------------------------------------------
template <int D>
__global__ void k_shard_set(uint32_t n, uint32_t npad, const uint32_t *__restrict__ gid, const double *__restrict__ x,
                            const double *__restrict__ v, const double *__restrict__ a, const double *__restrict__ f,
                            const double *__restrict__ m, double4 *pos, double *vo, double *ao, double *fo, uint32_t *order,
                            uint32_t *slot_of, uint8_t *ghost) {
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        double4 p;
        p.x = x[(size_t)s * D];
        p.y = x[(size_t)s * D + 1];
        p.z = D == 3 ? x[(size_t)s * D + 2] : 0.0;
        p.w = m[s];
        pos[s] = p;
        for (int d = 0; d < 3; d++) {
            const size_t q = (size_t)d * npad + s;
            vo[q] = d < D && v ? v[(size_t)s * D + d] : 0.0;
            ao[q] = d < D && a ? a[(size_t)s * D + d] : 0.0;
            fo[q] = d < D && f ? f[(size_t)s * D + d] : 0.0;
        }
        order[s] = gid[s];
        slot_of[gid[s]] = s;
        ghost[s] = 0;
    }
}

template <int D>
__global__ void k_shard_get(uint32_t first, uint32_t n, uint32_t npad, const double4 *__restrict__ pos,
                            const double *__restrict__ vi, const double *__restrict__ ai, const double *__restrict__ fi,
                            const uint32_t *__restrict__ order, uint32_t *gid, double *x, double *v, double *a, double *f,
                            double *m) {
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        const uint32_t s = first + q;
        const double4 p = pos[s];
        x[(size_t)q * D] = p.x;
        x[(size_t)q * D + 1] = p.y;
        if (D == 3) x[(size_t)q * D + 2] = p.z;
        m[q] = p.w;
        for (int d = 0; d < D; d++) {
            const size_t k = (size_t)d * npad + s;
            v[(size_t)q * D + d] = vi[k];
            a[(size_t)q * D + d] = ai[k];
            f[(size_t)q * D + d] = fi[k];
        }
        gid[q] = order[s];
    }
}

static int ensure_stage(parm_ctx *c, size_t bytes) {
    if (bytes <= c->d_stage_bytes) return 0;
    if (c->d_stage) cudaFree(c->d_stage);
    c->d_stage = 0;
    c->d_stage_bytes = 0;
    CK(cudaMalloc(&c->d_stage, bytes));
    c->d_stage_bytes = bytes;
    return 0;
}

extern "C" int parm_shard_set_atoms(parm_ctx *c, uint32_t n_local, const uint32_t *gid, const double *x, const double *v,
                                    const double *a, const double *f, const double *m) {
    if (!c || !c->sh.on) { parm_set_error("parm_shard_set_atoms: not a sharded context"); return PARM_ERR_INVALID; }
    if (n_local > c->npad) { parm_set_error("parm_shard_set_atoms: %u atoms exceed the slot capacity %u", n_local, c->npad); return PARM_ERR_INVALID; }
    if (n_local && (!gid || !x || !m)) { parm_set_error("parm_shard_set_atoms: gid, x and m are required"); return PARM_ERR_INVALID; }
    CK(cudaSetDevice(c->device));
    const int D = c->D;
    const size_t nv = (size_t)n_local * D * 8;
    PTRY(ensure_stage(c, 4 * nv + (size_t)n_local * 12 + 64));
    char *s = (char *)c->d_stage;
    double *dx = (double *)s, *dv = (double *)(s + nv), *da = (double *)(s + 2 * nv), *df = (double *)(s + 3 * nv);
    double *dm = (double *)(s + 4 * nv);
    uint32_t *dg = (uint32_t *)(s + 4 * nv + (size_t)n_local * 8);
    if (n_local) {
        CK(cudaMemcpyAsync(dx, x, nv, cudaMemcpyHostToDevice, c->stream));
        if (v) CK(cudaMemcpyAsync(dv, v, nv, cudaMemcpyHostToDevice, c->stream));
        if (a) CK(cudaMemcpyAsync(da, a, nv, cudaMemcpyHostToDevice, c->stream));
        if (f) CK(cudaMemcpyAsync(df, f, nv, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(dm, m, (size_t)n_local * 8, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(dg, gid, (size_t)n_local * 4, cudaMemcpyHostToDevice, c->stream));
        unsigned grid = std::min<unsigned>((n_local + 255) / 256, (unsigned)c->num_sms * 8);
        if (D == 3) k_shard_set<3><<<grid, 256, 0, c->stream>>>(n_local, c->npad, dg, dx, v ? dv : nullptr, a ? da : nullptr, f ? df : nullptr, dm, c->pos, c->v, c->a, c->f, c->order, c->slot_of, c->ghost);
        else k_shard_set<2><<<grid, 256, 0, c->stream>>>(n_local, c->npad, dg, dx, v ? dv : nullptr, a ? da : nullptr, f ? df : nullptr, dm, c->pos, c->v, c->a, c->f, c->order, c->slot_of, c->ghost);
        CK_LAUNCH(c);
    }
    CK(cudaStreamSynchronize(c->stream));
    c->n = n_local;
    c->sh.n_local = n_local;
    c->sh.g_dn = c->sh.g_up = c->sh.s_dn = c->sh.s_up = 0;
    for (parm_nlist *nl : c->nlists) nl->ignorechanged = true;
    return 0;
}

extern "C" int parm_shard_get_atoms(parm_ctx *c, uint32_t cap, uint32_t *n_local, uint32_t *gid, double *x, double *v,
                                    double *a, double *f, double *m) {
    if (!c || !c->sh.on) { parm_set_error("parm_shard_get_atoms: not a sharded context"); return PARM_ERR_INVALID; }
    CK(cudaSetDevice(c->device));
    const uint32_t n = c->sh.n_local;
    if (n_local) *n_local = n;
    if (!gid) return 0;
    if (cap < n) { parm_set_error("parm_shard_get_atoms: capacity %u < %u local atoms", cap, n); return PARM_ERR_INVALID; }
    if (!n) return 0;
    const int D = c->D;
    const size_t nv = (size_t)n * D * 8;
    PTRY(ensure_stage(c, 4 * nv + (size_t)n * 12 + 64));
    char *s = (char *)c->d_stage;
    double *dx = (double *)s, *dv = (double *)(s + nv), *da = (double *)(s + 2 * nv), *df = (double *)(s + 3 * nv);
    double *dm = (double *)(s + 4 * nv);
    uint32_t *dg = (uint32_t *)(s + 4 * nv + (size_t)n * 8);
    unsigned grid = std::min<unsigned>((n + 255) / 256, (unsigned)c->num_sms * 8);
    if (D == 3) k_shard_get<3><<<grid, 256, 0, c->stream>>>(0, n, c->npad, c->pos, c->v, c->a, c->f, c->order, dg, dx, dv, da, df, dm);
    else k_shard_get<2><<<grid, 256, 0, c->stream>>>(0, n, c->npad, c->pos, c->v, c->a, c->f, c->order, dg, dx, dv, da, df, dm);
    CK_LAUNCH(c);
    if (x) CK(cudaMemcpyAsync(x, dx, nv, cudaMemcpyDeviceToHost, c->stream));
    if (v) CK(cudaMemcpyAsync(v, dv, nv, cudaMemcpyDeviceToHost, c->stream));
    if (a) CK(cudaMemcpyAsync(a, da, nv, cudaMemcpyDeviceToHost, c->stream));
    if (f) CK(cudaMemcpyAsync(f, df, nv, cudaMemcpyDeviceToHost, c->stream));
    if (m) CK(cudaMemcpyAsync(m, dm, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(gid, dg, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

template <int D>
__global__ void k_shard_put(uint32_t first, uint32_t n, uint32_t npad, const double *__restrict__ x, const double *__restrict__ v,
                            const double *__restrict__ a, const double *__restrict__ f, double4 *pos, double *vo, double *ao,
                            double *fo) {
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        const uint32_t s = first + q;
        if (x) {
            double4 p = pos[s];
            p.x = x[(size_t)q * D];
            p.y = x[(size_t)q * D + 1];
            if (D == 3) p.z = x[(size_t)q * D + 2];
            pos[s] = p;
        }
        for (int d = 0; d < D; d++) {
            const size_t k = (size_t)d * npad + s;
            if (v) vo[k] = v[(size_t)q * D + d];
            if (a) ao[k] = a[(size_t)q * D + d];
            if (f) fo[k] = f[(size_t)q * D + d];
        }
    }
}

// Overwrite fields of the local atoms, in the slot order parm_shard_get_atoms returned them.
// Like parm_upload_atoms it does not touch the neighbour list (the reference only looks at positions
// again at the next NeighborList::update).
extern "C" int parm_shard_put_atoms(parm_ctx *c, uint32_t n_local, const double *x, const double *v, const double *a,
                                    const double *f) {
    if (!c || !c->sh.on) { parm_set_error("parm_shard_put_atoms: not a sharded context"); return PARM_ERR_INVALID; }
    if (n_local != c->sh.n_local) { parm_set_error("parm_shard_put_atoms: %u atoms given, %u local", n_local, c->sh.n_local); return PARM_ERR_INVALID; }
    CK(cudaSetDevice(c->device));
    if (!n_local) return 0;
    const int D = c->D;
    const size_t nv = (size_t)n_local * D * 8;
    PTRY(ensure_stage(c, 4 * nv + 64));
    char *s = (char *)c->d_stage;
    double *dx = (double *)s, *dv = (double *)(s + nv), *da = (double *)(s + 2 * nv), *df = (double *)(s + 3 * nv);
    if (x) CK(cudaMemcpyAsync(dx, x, nv, cudaMemcpyHostToDevice, c->stream));
    if (v) CK(cudaMemcpyAsync(dv, v, nv, cudaMemcpyHostToDevice, c->stream));
    if (a) CK(cudaMemcpyAsync(da, a, nv, cudaMemcpyHostToDevice, c->stream));
    if (f) CK(cudaMemcpyAsync(df, f, nv, cudaMemcpyHostToDevice, c->stream));
    unsigned grid = std::min<unsigned>((n_local + 255) / 256, (unsigned)c->num_sms * 8);
    if (D == 3) k_shard_put<3><<<grid, 256, 0, c->stream>>>(0, n_local, c->npad, x ? dx : nullptr, v ? dv : nullptr, a ? da : nullptr, f ? df : nullptr, c->pos, c->v, c->a, c->f);
    else k_shard_put<2><<<grid, 256, 0, c->stream>>>(0, n_local, c->npad, x ? dx : nullptr, v ? dv : nullptr, a ? da : nullptr, f ? df : nullptr, c->pos, c->v, c->a, c->f);
    CK_LAUNCH(c);
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int parm_shard_info(parm_ctx *c, uint32_t *out /*n_local, ghosts_down, ghosts_up, send_down, send_up, slots*/) {
    out[0] = c->sh.n_local;
    out[1] = c->sh.g_dn;
    out[2] = c->sh.g_up;
    out[3] = c->sh.s_dn;
    out[4] = c->sh.s_up;
    out[5] = c->n;
    return 0;
}

extern "C" int parm_shard_rebuild_stats(parm_ctx *c, uint64_t *out2) {
    if (!c || !c->sh.on || !out2) { parm_set_error("parm_shard_rebuild_stats: not a sharded context"); return PARM_ERR_INVALID; }
    out2[0] = c->sh.fast_rebuilds;
    out2[1] = c->sh.slow_rebuilds;
    return 0;
}

int parm_shard_allreduce_sum(parm_ctx *c, double *d_buf, int count) {
    NcclApi *n = nccl_api();
    NCK(n->AllReduce(d_buf, d_buf, (size_t)count, ncclFloat64, ncclSum, (ncclComm_t)c->sh.comm, c->stream));
    return 0;
}

static int halo_exchange_on(parm_ctx *c, cudaStream_t st);
int parm_shard_halo_exchange(parm_ctx *c) { return halo_exchange_on(c, c->stream); }

static int halo_exchange_on(parm_ctx *c, cudaStream_t st) {
    ShardState &sh = c->sh;
    if (!sh.on) return 0;
    NcclApi *n = nccl_api();
    ncclComm_t comm = (ncclComm_t)sh.comm;
    const uint32_t nl = sh.n_local;
    NCK(n->GroupStart());
    // sends: [to down, to up]; receives: [from up, from down] -- with 2 ranks both neighbours are the same
    // peer and NCCL matches same-peer messages in issue order. Every range is contiguous in `pos`.
    if (sh.s_dn) NCK(n->Send(c->pos, (size_t)sh.s_dn * 4, ncclFloat64, sh.down, comm, st));
    if (sh.s_up) NCK(n->Send(c->pos + (nl - sh.s_up), (size_t)sh.s_up * 4, ncclFloat64, sh.up, comm, st));
    if (sh.g_up) NCK(n->Recv(c->pos + nl, (size_t)sh.g_up * 4, ncclFloat64, sh.up, comm, st));
    if (sh.g_dn) NCK(n->Recv(c->pos + nl + sh.g_up, (size_t)sh.g_dn * 4, ncclFloat64, sh.down, comm, st));
    NCK(n->GroupEnd());
    return 0;
}

int parm_shard_drift_decision(parm_nlist *nl, bool *rebuild) {
    parm_ctx *c = nl->ctx;
    ShardState &sh = c->sh;
    NcclApi *n = nccl_api();
    // K3 left this rank's top-2 displacements in d_flags->top2
    NCK(n->AllGather(nl->d_flags->top2, sh.d_gather, 2, ncclFloat64, (ncclComm_t)sh.comm, c->stream));
    CK(cudaMemcpyAsync(sh.h_gather, sh.d_gather, 16 * (size_t)sh.nranks, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    double b1 = 0, b2 = 0;
    for (int k = 0; k < 2 * sh.nranks; k++) {
        double d = sh.h_gather[k];
        if (d > b1) { b2 = b1; b1 = d; }
        else if (d > b2) b2 = d;
    }
    *rebuild = (b2 + b1 >= nl->skin); // bigdist + biggestdist >= skin, identical on every rank
    return 0;
}

// device-side fold of the gathered top-2 values: identical arithmetic on every rank
__global__ void k_fold_decision(const double *__restrict__ gathered, int nvals, double skin, int *d_slot, int *h_slot) {
    if (threadIdx.x || blockIdx.x) return;
    double b1 = 0, b2 = 0;
    for (int k = 0; k < nvals; k++) {
        double d = gathered[k];
        if (d > b1) { b2 = b1; b1 = d; }
        else if (d > b2) b2 = d;
    }
    const int need = (__dadd_rn(b2, b1) >= skin) ? 1 : 0; // bigdist + biggestdist >= skin
    *d_slot = need;
    *h_slot = need;
    __threadfence_system();
}

static int drift_enqueue_on(parm_nlist *nl, int *d_slot, int *h_slot, cudaStream_t st) {
    parm_ctx *c = nl->ctx;
    ShardState &sh = c->sh;
    NcclApi *n = nccl_api();
    NCK(n->AllGather(nl->d_flags->top2, sh.d_gather, 2, ncclFloat64, (ncclComm_t)sh.comm, st));
    k_fold_decision<<<1, 32, 0, st>>>(sh.d_gather, 2 * sh.nranks, nl->skin, d_slot, h_slot);
    CK_LAUNCH(c);
    return 0;
}
int parm_shard_drift_enqueue(parm_nlist *nl, int *d_slot, int *h_slot) { return drift_enqueue_on(nl, d_slot, h_slot, nl->ctx->stream); }

// After K1 (queued on the main stream): exchange the boundary-layer positions and, with a tracker, all-gather
// the per-rank top-2 displacements and fold them into the step's decision word -- all on the communication
// stream, so the main stream can meanwhile compute the forces of the interior atoms.
int parm_shard_step_comm(parm_ctx *c, parm_nlist *nl, int *d_slot, int *h_slot) {
    ShardState &sh = c->sh;
    CK(cudaEventRecord(sh.ev_k1, c->stream));
    CK(cudaStreamWaitEvent(sh.comm_stream, sh.ev_k1, 0));
    PTRY(halo_exchange_on(c, sh.comm_stream));
    if (nl) PTRY(drift_enqueue_on(nl, d_slot, h_slot, sh.comm_stream));
    CK(cudaEventRecord(sh.ev_comm, sh.comm_stream));
    return 0;
}
int parm_shard_step_join(parm_ctx *c) {
    CK(cudaEventRecord(c->sh.ev_comm, c->sh.comm_stream)); // after whatever was queued behind the exchange
    CK(cudaStreamWaitEvent(c->stream, c->sh.ev_comm, 0));
    return 0;
}

__global__ void k_iota2(uint32_t *dst, uint32_t a0, uint32_t na, uint32_t b0, uint32_t nb) {
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < na + nb; q += gridDim.x * blockDim.x)
        dst[q] = q < na ? a0 + q : b0 + (q - na);
}
__global__ void k_fill_u8(uint8_t *dst, uint32_t n, uint8_t val) {
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) dst[q] = val;
}

// counts[0] -> down, counts[1] -> up; returns what the up / down neighbours sent to this rank
static int exchange_counts(parm_ctx *c, uint32_t to_down, uint32_t to_up, uint32_t *from_up, uint32_t *from_down) {
    ShardState &sh = c->sh;
    NcclApi *n = nccl_api();
    ncclComm_t comm = (ncclComm_t)sh.comm;
    sh.h_counts[0] = to_down;
    sh.h_counts[1] = to_up;
    CK(cudaMemcpyAsync(sh.d_counts, sh.h_counts, 8, cudaMemcpyHostToDevice, c->stream));
    NCK(n->GroupStart());
    NCK(n->Send(sh.d_counts + 0, 1, ncclUint32, sh.down, comm, c->stream));
    NCK(n->Send(sh.d_counts + 1, 1, ncclUint32, sh.up, comm, c->stream));
    NCK(n->Recv(sh.d_counts + 2, 1, ncclUint32, sh.up, comm, c->stream));
    NCK(n->Recv(sh.d_counts + 3, 1, ncclUint32, sh.down, comm, c->stream));
    NCK(n->GroupEnd());
    CK(cudaMemcpyAsync(sh.h_counts + 2, sh.d_counts + 2, 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    *from_up = sh.h_counts[2];
    *from_down = sh.h_counts[3];
    return 0;
}

// send slot ranges [dn0, dn0+ndn) -> down and [up0, up0+nup) -> up; receive rup atoms from up into
// [dst, dst+rup) and rdn from down into [dst+rup, ...). full: whole state (migration), else pos + id (ghosts)
static int exchange_ranges(parm_ctx *c, bool full, uint32_t dn0, uint32_t ndn, uint32_t up0, uint32_t nup, uint32_t dst,
                           uint32_t rup, uint32_t rdn) {
    ShardState &sh = c->sh;
    NcclApi *n = nccl_api();
    ncclComm_t comm = (ncclComm_t)sh.comm;
    const size_t np = c->npad;
    NCK(n->GroupStart());
    for (int pass = 0; pass < 4; pass++) {
        // pass 0: send down, 1: send up, 2: recv from up, 3: recv from down (same-peer ordering, see halo exchange)
        const bool snd = pass < 2;
        const uint32_t s0 = pass == 0 ? dn0 : pass == 1 ? up0 : pass == 2 ? dst : dst + rup;
        const uint32_t cnt = pass == 0 ? ndn : pass == 1 ? nup : pass == 2 ? rup : rdn;
        const int peer = (pass == 0 || pass == 3) ? sh.down : sh.up;
        if (!cnt) continue;
        if (snd) {
            NCK(n->Send(c->pos + s0, (size_t)cnt * 4, ncclFloat64, peer, comm, c->stream));
            NCK(n->Send(c->order + s0, cnt, ncclUint32, peer, comm, c->stream));
        } else {
            NCK(n->Recv(c->pos + s0, (size_t)cnt * 4, ncclFloat64, peer, comm, c->stream));
            NCK(n->Recv(c->order + s0, cnt, ncclUint32, peer, comm, c->stream));
        }
        if (full) {
            double *arrs[3] = {c->v, c->a, c->f};
            for (int q = 0; q < 3; q++)
                for (int d = 0; d < 3; d++) {
                    double *p = arrs[q] + (size_t)d * np + s0;
                    if (snd) NCK(n->Send(p, cnt, ncclFloat64, peer, comm, c->stream));
                    else NCK(n->Recv(p, cnt, ncclFloat64, peer, comm, c->stream));
                }
        }
    }
    NCK(n->GroupEnd());
    return 0;
}

// ---- one-sort rebuild -------------------------------------------------------------------------------
// The two-sort path below sorts every owned atom to find the few hundred that left the slab, waits for their number,
// tells the neighbours, waits again, and sorts the new owned set a second time. Here the leavers are FOUND without a
// sort (the binning rule of k_cell_id on the slab axis, nothing else), packed in slot order -- which is what makes the
// order of the arrivals, and with it every later summation order, reproducible -- into fixed-capacity messages whose
// header carries the count, and exchanged before the host knows any number. One host round trip later (counts +
// overflow flag, max over all ranks) the arrivals are appended behind the owned slots and ONE sort bins everything: the
// leavers land in the halo layers behind the new owned set, exactly where the two-sort path left them, and are
// overwritten by the ghosts. The resulting slot order is the two-sort path's (stable sorts, same source order).
__global__ void k_find_leavers(const double4 *__restrict__ pos, uint32_t n, ShardDev sd, uint32_t cap, uint32_t *list, uint32_t *cnt) {
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        const double rel = shard_rel(pos[s].x, sd);
        const int dir = rel < 0.0 ? 0 : (rel >= sd.Ls ? 1 : -1); // k_cell_id: halo layer below / above / an interior layer
        if (dir >= 0) {
            const uint32_t k = atomicAdd(cnt + dir, 1u); // (claim order is arbitrary: k_pack_leavers sorts the slots)
            if (k < cap) list[(size_t)dir * cap + k] = s;
        }
    }
}

// block 0: leavers downwards, block 1: upwards. Ranks every slot among the direction's leavers (a few hundred) and packs
// the full state in ascending slot order.
__global__ void __launch_bounds__(1024) k_pack_leavers(const uint32_t *__restrict__ list, uint32_t *cnt, uint32_t cap, uint32_t npad,
                                                       const double4 *__restrict__ pos, const double *__restrict__ v,
                                                       const double *__restrict__ a, const double *__restrict__ f,
                                                       const uint32_t *__restrict__ order, char *buf_dn, char *buf_up) {
    extern __shared__ uint32_t s_list[];
    const int dir = blockIdx.x;
    const uint32_t nraw = cnt[dir], n = min(nraw, cap);
    const MigView m = mig_view(dir ? buf_up : buf_dn, cap);
    for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) s_list[k] = list[(size_t)dir * cap + k];
    __syncthreads();
    if (threadIdx.x == 0) {
        m.hdr[0] = nraw > cap ? 0u : n;
        m.hdr[1] = nraw > cap ? 1u : 0u;
        m.hdr[2] = m.hdr[3] = 0u;
        if (nraw > cap) atomicMax(cnt + 2, 1u);
    }
    if (nraw > cap) return; // (uniform) every rank falls back to the two-sort path
    for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) {
        const uint32_t s = s_list[k];
        uint32_t rank = 0;
        for (uint32_t j = 0; j < n; j++) rank += s_list[j] < s;
        m.pos[rank] = pos[s];
        for (int d = 0; d < 3; d++) {
            m.v[(size_t)d * cap + rank] = v[(size_t)d * npad + s];
            m.a[(size_t)d * cap + rank] = a[(size_t)d * npad + s];
            m.f[(size_t)d * cap + rank] = f[(size_t)d * npad + s];
        }
        m.id[rank] = order[s];
    }
}

// arrivals behind the owned slots: first what came from the upper neighbour, then from the lower one
__global__ void k_unpack_arrivals(const char *from_up, const char *from_dn, uint32_t cap, uint32_t dst, uint32_t npad, double4 *pos,
                                  double *v, double *a, double *f, uint32_t *order, uint8_t *ghost) {
    const MigView mu = mig_view(const_cast<char *>(from_up), cap), md = mig_view(const_cast<char *>(from_dn), cap);
    const uint32_t ru = mu.hdr[0], rd = md.hdr[0];
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < ru + rd; q += gridDim.x * blockDim.x) {
        const MigView &m = q < ru ? mu : md;
        const uint32_t k = q < ru ? q : q - ru, s = dst + q;
        pos[s] = m.pos[k];
        for (int d = 0; d < 3; d++) {
            v[(size_t)d * npad + s] = m.v[(size_t)d * cap + k];
            a[(size_t)d * npad + s] = m.a[(size_t)d * cap + k];
            f[(size_t)d * npad + s] = m.f[(size_t)d * cap + k];
        }
        order[s] = m.id[k];
        ghost[s] = 0;
    }
}

// boundary-layer sizes of the freshly sorted owned set, straight from the cell table
__global__ void k_ghost_counts(const uint32_t *__restrict__ cell_start, uint32_t plane, uint32_t nci, uint32_t n_local, uint32_t *out) {
    if (threadIdx.x || blockIdx.x) return;
    out[0] = cell_start[plane];                                 // lowest interior layer -> ghosts of the lower neighbour
    out[1] = n_local - cell_start[(size_t)(nci - 1) * plane];   // highest interior layer -> ghosts of the upper neighbour
    out[4] = cell_start[(size_t)nci * plane];                   // atoms binned into interior layers (must be n_local)
}

static int shard_rebuild_two_sorts(parm_nlist *nl);

// 0: done; 1: more leavers than a message holds somewhere -- nothing was changed, take the two-sort path
static int shard_rebuild_one_sort(parm_nlist *nl, bool *fallback) {
    parm_ctx *c = nl->ctx;
    ShardState &sh = c->sh;
    NcclApi *n = nccl_api();
    ncclComm_t comm = (ncclComm_t)sh.comm;
    *fallback = false;
    const uint32_t plane = (uint32_t)nl->g.nc[1] * (uint32_t)nl->g.nc[2];
    const uint32_t nci = (uint32_t)nl->sd.nci;
    const uint32_t cap = sh.mig_cap, n_old = sh.n_local;
    auto grid = [&](uint32_t k) { return std::max(1u, std::min<unsigned>((k + 255) / 256, (unsigned)c->num_sms * 8)); };
    CK(cudaMemsetAsync(sh.mig_cnt, 0, 16, c->stream));
    if (n_old) {
        k_find_leavers<<<grid(n_old), 256, 0, c->stream>>>(c->pos, n_old, nl->sd, cap, sh.mig_list, sh.mig_cnt);
        CK_LAUNCH(c);
    }
    if ((size_t)cap * 4 > 48 * 1024) CK(cudaFuncSetAttribute(k_pack_leavers, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(cap * 4)));
    k_pack_leavers<<<2, 1024, (size_t)cap * 4, c->stream>>>(sh.mig_list, sh.mig_cnt, cap, c->npad, c->pos, c->v, c->a, c->f, c->order,
                                                            sh.mig_send[0], sh.mig_send[1]);
    CK_LAUNCH(c);
    const size_t mb = mig_bytes(cap);
    NCK(n->GroupStart());
    // sends: [to down, to up]; receives: [from up, from down] (same-peer ordering, see the halo exchange)
    NCK(n->Send(sh.mig_send[0], mb, ncclUint8, sh.down, comm, c->stream));
    NCK(n->Send(sh.mig_send[1], mb, ncclUint8, sh.up, comm, c->stream));
    NCK(n->Recv(sh.mig_recv[0], mb, ncclUint8, sh.up, comm, c->stream));
    NCK(n->Recv(sh.mig_recv[1], mb, ncclUint8, sh.down, comm, c->stream));
    NCK(n->GroupEnd());
    NCK(n->AllReduce(sh.mig_cnt + 2, sh.mig_cnt + 2, 1, ncclUint32, ncclMax, comm, c->stream)); // anyone over capacity?
    CK(cudaMemcpyAsync(sh.h_counts, sh.mig_cnt, 12, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(sh.h_counts + 4, sh.mig_recv[0], 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(sh.h_counts + 5, sh.mig_recv[1], 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (sh.h_counts[2]) { *fallback = true; return 0; }
    const uint32_t m_dn = sh.h_counts[0], m_up = sh.h_counts[1], r_up = sh.h_counts[4], r_dn = sh.h_counts[5];
    if ((uint64_t)n_old + r_up + r_dn > c->npad) { parm_set_error("slab rank %d: slot capacity %u too small for migration", sh.rank, c->npad); return PARM_ERR_RUNTIME; }
    if (r_up + r_dn) {
        k_unpack_arrivals<<<grid(r_up + r_dn), 256, 0, c->stream>>>(sh.mig_recv[0], sh.mig_recv[1], cap, n_old, c->npad, c->pos, c->v, c->a,
                                                                   c->f, c->order, c->ghost);
        CK_LAUNCH(c);
    }
    // the one sort: stayers in their old slot order, then the arrivals; the leavers are binned into the halo layers
    const uint32_t n_local = n_old - m_dn - m_up + r_up + r_dn;
    PTRY(parm_nlist_sort_permute(nl, nullptr, n_old + r_up + r_dn));
    k_ghost_counts<<<1, 32, 0, c->stream>>>(nl->cell_start, plane, nci, n_local, sh.d_counts);
    CK_LAUNCH(c);
    NCK(n->GroupStart());
    NCK(n->Send(sh.d_counts + 0, 1, ncclUint32, sh.down, comm, c->stream));
    NCK(n->Send(sh.d_counts + 1, 1, ncclUint32, sh.up, comm, c->stream));
    NCK(n->Recv(sh.d_counts + 2, 1, ncclUint32, sh.up, comm, c->stream));
    NCK(n->Recv(sh.d_counts + 3, 1, ncclUint32, sh.down, comm, c->stream));
    NCK(n->GroupEnd());
    CK(cudaMemcpyAsync(sh.h_counts + 8, sh.d_counts, 20, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (sh.h_counts[12] != n_local) {
        parm_set_error("slab rank %d: %d atoms are more than one layer outside the slab after migration "
                       "(atoms must start in, or next to, their slab)", sh.rank, (int)n_local - (int)sh.h_counts[12]);
        return PARM_ERR_RUNTIME;
    }
    const uint32_t s_dn = sh.h_counts[8], s_up = sh.h_counts[9], g_up = sh.h_counts[10], g_dn = sh.h_counts[11];
    if ((uint64_t)n_local + g_up + g_dn > c->npad) { parm_set_error("slab rank %d: slot capacity %u too small for %u ghosts", sh.rank, c->npad, g_up + g_dn); return PARM_ERR_RUNTIME; }
    // ghosts land behind the owned atoms (over the leavers): first the upper neighbour's lowest layer, then the lower neighbour's highest
    PTRY(exchange_ranges(c, false, 0, s_dn, n_local - s_up, s_up, n_local, g_up, g_dn));
    PTRY(parm_nlist_append_ghosts(nl, n_local, g_up + g_dn));
    sh.n_local = n_local;
    sh.g_dn = g_dn;
    sh.g_up = g_up;
    sh.s_dn = s_dn;
    sh.s_up = s_up;
    sh.fast_rebuilds++;
    return parm_nlist_build_rows(nl);
}

int parm_shard_rebuild(parm_nlist *nl) {
    parm_ctx *c = nl->ctx;
    ShardState &sh = c->sh;
    sh.Ls = c->box.L[0] / sh.nranks;
    sh.lo = sh.rank * sh.Ls;
    PTRY(parm_nlist_prepare_grid(nl));
    PTRY(parm_prof_begin(c, PARM_PROF_REBUILD));
    CK(cudaMemsetAsync(nl->d_flags, 0, sizeof(NlistFlags), c->stream));
    if (sh.mig_fast) {
        bool fallback = false;
        PTRY(shard_rebuild_one_sort(nl, &fallback));
        if (!fallback) return 0;
    }
    sh.slow_rebuilds++;
    return shard_rebuild_two_sorts(nl);
}

static int shard_rebuild_two_sorts(parm_nlist *nl) {
    parm_ctx *c = nl->ctx;
    ShardState &sh = c->sh;
    const uint32_t plane = (uint32_t)nl->g.nc[1] * (uint32_t)nl->g.nc[2];
    const uint32_t nci = (uint32_t)nl->sd.nci; // layers 0..nci-1 interior, nci = halo above, nci+1 = halo below
    auto grid = [&](uint32_t n) { return std::max(1u, std::min<unsigned>((n + 255) / 256, (unsigned)c->num_sms * 8)); };

    // ---- phase 1: sort the owned atoms; the ones binned into a halo layer have left the slab:
    //      [interior | left upwards | left downwards], every range contiguous
    PTRY(parm_nlist_sort_permute(nl, nullptr, sh.n_local));
    CK(cudaMemcpyAsync(&sh.h_counts[8], nl->cell_start + (size_t)nci * plane, 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(&sh.h_counts[9], nl->cell_start + (size_t)(nci + 1) * plane, 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    const uint32_t keep = sh.h_counts[8], dn_begin = sh.h_counts[9];
    const uint32_t m_up = dn_begin - keep, m_dn = sh.n_local - dn_begin;
    uint32_t r_up = 0, r_dn = 0;
    PTRY(exchange_counts(c, m_dn, m_up, &r_up, &r_dn));
    if ((uint64_t)sh.n_local + r_up + r_dn > c->npad) { parm_set_error("slab rank %d: slot capacity %u too small for migration", sh.rank, c->npad); return PARM_ERR_RUNTIME; }
    PTRY(exchange_ranges(c, true, dn_begin, m_dn, keep, m_up, sh.n_local, r_up, r_dn));
    if (r_up + r_dn) {
        k_fill_u8<<<grid(r_up + r_dn), 256, 0, c->stream>>>(c->ghost + sh.n_local, r_up + r_dn, 0);
        CK_LAUNCH(c);
    }
    const uint32_t n_local = keep + r_up + r_dn;

    // ---- phase 2: sort the new owned set; its lowest / highest layers are the neighbours' ghosts
    k_iota2<<<grid(n_local), 256, 0, c->stream>>>(nl->cell_id_sorted, 0, keep, sh.n_local, r_up + r_dn);
    CK_LAUNCH(c);
    PTRY(parm_nlist_sort_permute(nl, nl->cell_id_sorted, n_local));
    const size_t idx[3] = {(size_t)plane, (size_t)(nci - 1) * plane, (size_t)nci * plane};
    for (int k = 0; k < 3; k++) CK(cudaMemcpyAsync(&sh.h_counts[8 + k], nl->cell_start + idx[k], 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (sh.h_counts[10] != n_local) {
        parm_set_error("slab rank %d: %u atoms are more than one layer outside the slab after migration "
                       "(atoms must start in, or next to, their slab)", sh.rank, n_local - sh.h_counts[10]);
        return PARM_ERR_RUNTIME;
    }
    const uint32_t s_dn = sh.h_counts[8], s_up = n_local - sh.h_counts[9];
    uint32_t g_up = 0, g_dn = 0;
    PTRY(exchange_counts(c, s_dn, s_up, &g_up, &g_dn));
    if ((uint64_t)n_local + g_up + g_dn > c->npad) { parm_set_error("slab rank %d: slot capacity %u too small for %u ghosts", sh.rank, c->npad, g_up + g_dn); return PARM_ERR_RUNTIME; }
    // ghosts land behind the owned atoms: first the upper neighbour's lowest layer, then the lower neighbour's highest
    PTRY(exchange_ranges(c, false, 0, s_dn, n_local - s_up, s_up, n_local, g_up, g_dn));

    // ---- phase 3: per-slot data of the ghosts (no third sort) and the rows of the owned atoms
    PTRY(parm_nlist_append_ghosts(nl, n_local, g_up + g_dn));
    sh.n_local = n_local;
    sh.g_dn = g_dn;
    sh.g_up = g_up;
    sh.s_dn = s_dn;
    sh.s_up = s_up;
    return parm_nlist_build_rows(nl);
}

int parm_shard_destroy(parm_ctx *c) {
    if (!c->sh.on) return 0;
    NcclApi *n = nccl_api();
    if (n && c->sh.comm) n->CommDestroy((ncclComm_t)c->sh.comm);
    if (c->sh.comm_stream) { cudaStreamDestroy(c->sh.comm_stream); cudaEventDestroy(c->sh.ev_k1); cudaEventDestroy(c->sh.ev_comm); cudaEventDestroy(c->sh.ev_rows); }
    if (c->sh.mig_list) cudaFree(c->sh.mig_list);
    if (c->sh.mig_cnt) cudaFree(c->sh.mig_cnt);
    for (int d = 0; d < 2; d++) {
        if (c->sh.mig_send[d]) cudaFree(c->sh.mig_send[d]);
        if (c->sh.mig_recv[d]) cudaFree(c->sh.mig_recv[d]);
    }
    if (c->sh.d_counts) cudaFree(c->sh.d_counts);
    if (c->sh.h_counts) cudaFreeHost(c->sh.h_counts);
    if (c->sh.d_gather) cudaFree(c->sh.d_gather);
    if (c->sh.h_gather) cudaFreeHost(c->sh.h_gather);
    return 0;
}
