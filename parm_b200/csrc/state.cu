// Context (OriginBox + AtomVec device state), host<->device transfers and the
// AtomGroup / Collection reductions.  Reference: box.hpp:97-158, 234-249, 441-479;
// box.cpp:239-260, 401-431; collection.cpp:21-29, 116-133.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>

#include "internal.cuh"

static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches{0};

void parm_set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
}
void parm_count_launch(parm_ctx *ctx, unsigned n) {
    g_launches += n;
    if (ctx) ctx->launches += n;
}

extern "C" const char *parm_b200_last_error(void) { return g_err.c_str(); }
extern "C" uint64_t parm_b200_launch_count(void) { return g_launches.load(); }
extern "C" const char *parm_b200_version(void) { return "parm_b200 0.1 (sm_100a)"; }

static inline unsigned grid_for(const parm_ctx *ctx, uint32_t n, unsigned block, unsigned per_sm = 8) {
    unsigned need = (n + block - 1) / block;
    unsigned cap = (unsigned)ctx->num_sms * per_sm;
    if (need < 1) need = 1;
    return need < cap ? need : cap;
}

int parm_ctx_ensure_red(parm_ctx *ctx, size_t doubles) {
    if (doubles <= ctx->d_red_doubles) return 0;
    if (ctx->d_red) cudaFree(ctx->d_red);
    if (ctx->h_red) cudaFreeHost(ctx->h_red);
    ctx->d_red = 0;
    ctx->h_red = 0;
    CK(cudaMalloc(&ctx->d_red, doubles * sizeof(double)));
    CK(cudaMallocHost(&ctx->h_red, doubles * sizeof(double)));
    ctx->d_red_doubles = doubles;
    return 0;
}

__global__ void k_init_order(uint32_t *order, uint32_t *slot_of, uint32_t n) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        order[i] = i;
        slot_of[i] = i;
    }
}

extern "C" int parm_ctx_create(int ndim, uint32_t n_atoms, int device, parm_ctx **out) {
    return parm_ctx_alloc(ndim, n_atoms, n_atoms, device, out);
}

// nid atom ids, cap_slots slots (== nid for a single-GPU context; local + ghost capacity when sharded)
int parm_ctx_alloc(int ndim, uint32_t n_atoms, uint32_t cap_slots, int device, parm_ctx **out) {
    if (!out) { parm_set_error("parm_ctx_create: out is NULL"); return PARM_ERR_INVALID; }
    *out = 0;
    if (ndim != 2 && ndim != 3) { parm_set_error("parm_ctx_create: ndim must be 2 or 3 (got %d)", ndim); return PARM_ERR_INVALID; }
    if (n_atoms > (1u << 30)) { parm_set_error("parm_ctx_create: too many atoms"); return PARM_ERR_INVALID; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        parm_set_error("parm_b200 needs a CUDA device (sm_100a); none usable: %s. There is no CPU fallback.",
                       e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return PARM_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { parm_set_error("parm_ctx_create: bad device %d (have %d)", device, ndev); return PARM_ERR_INVALID; }
    CK(cudaSetDevice(device));
    parm_ctx *c = new parm_ctx();
    c->D = ndim;
    c->n = n_atoms;
    c->nid = n_atoms;
    c->npad = ((cap_slots + 127u) / 128u) * 128u;
    if (c->npad == 0) c->npad = 128;
    c->nid_pad = ((n_atoms + 127u) / 128u) * 128u;
    if (c->nid_pad == 0) c->nid_pad = 128;
    if (c->nid_pad < c->npad && cap_slots == n_atoms) c->nid_pad = c->npad;
    c->device = device;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    c->num_sms = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    size_t np = c->npad;
    CK(cudaMalloc(&c->pos, np * sizeof(double4)));
    CK(cudaMalloc(&c->pos_alt, np * sizeof(double4)));
    double **vecs[6] = {&c->v, &c->a, &c->f, &c->v_alt, &c->a_alt, &c->f_alt};
    for (int k = 0; k < 6; k++) {
        CK(cudaMalloc(vecs[k], 3 * np * sizeof(double)));
        CK(cudaMemsetAsync(*vecs[k], 0, 3 * np * sizeof(double), c->stream));
    }
    CK(cudaMemsetAsync(c->pos, 0, np * sizeof(double4), c->stream));
    CK(cudaMemsetAsync(c->pos_alt, 0, np * sizeof(double4), c->stream));
    CK(cudaMalloc(&c->order, np * 4));
    CK(cudaMalloc(&c->order_alt, np * 4));
    CK(cudaMalloc(&c->slot_of, (size_t)std::max(c->nid_pad, c->npad) * 4));
    k_init_order<<<grid_for(c, c->npad, 256), 256, 0, c->stream>>>(c->order, c->slot_of, c->npad);
    CK_LAUNCH(c);
    PTRY(parm_ctx_ensure_red(c, 4096));
    for (int k = 0; k < 3; k++) { c->box.L[k] = 1.0; c->box.invL[k] = 1.0; c->box.halfL[k] = 0.5; }
    CK(cudaStreamSynchronize(c->stream));
    *out = c;
    return 0;
}

extern "C" int parm_ctx_destroy(parm_ctx *c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    parm_shard_destroy(c);
    parm_snapshot_free(c);
    void *ptrs[] = {c->pos, c->pos_alt, c->v, c->a, c->f, c->v_alt, c->a_alt, c->f_alt, c->order, c->order_alt,
                    c->slot_of, c->d_stage, c->d_red, c->ghost, c->ghost_alt};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (c->h_red) cudaFreeHost(c->h_red);
    for (auto &r : c->prof_pending) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : c->prof_pool) cudaEventDestroy(e);
    cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

extern "C" int parm_set_box(parm_ctx *c, const double *L) {
    if (!c || !L) { parm_set_error("parm_set_box: NULL argument"); return PARM_ERR_INVALID; }
    for (int k = 0; k < c->D; k++)
        if (!(L[k] > 0) || isinf(L[k])) { parm_set_error("parm_set_box: box lengths must be positive and finite"); return PARM_ERR_INVALID; }
    for (int k = 0; k < 3; k++) {
        double l = k < c->D ? L[k] : 1.0;
        c->box.L[k] = l;
        c->box.invL[k] = 1.0 / l;
        c->box.halfL[k] = l * 0.5;
    }
    c->box_set = true; // (the reference keeps its pair list across a resize: the grid is re-derived at every rebuild)
    for (parm_nlist *nl : c->nlists) parm_tile_invalidate(nl); // tile origins belong to the old box: gather kernel until the next rebuild
    return 0;
}

extern "C" int parm_get_box(parm_ctx *c, double *L) {
    for (int k = 0; k < c->D; k++) L[k] = c->box.L[k];
    return 0;
}

extern "C" int parm_sync(parm_ctx *c) {
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int parm_get_stream(parm_ctx *c, void **st) {
    *st = (void *)c->stream;
    return 0;
}

static int prof_event(parm_ctx *c, cudaEvent_t *e) {
    if (!c->prof_pool.empty()) {
        *e = c->prof_pool.back();
        c->prof_pool.pop_back();
        return 0;
    }
    CK(cudaEventCreate(e));
    return 0;
}
int parm_prof_begin(parm_ctx *c, int cls) {
    if (!c->prof_on) return 0;
    parm_ctx::ProfRec r;
    r.cls = cls;
    PTRY(prof_event(c, &r.a));
    PTRY(prof_event(c, &r.b));
    CK(cudaEventRecord(r.a, c->stream));
    c->prof_pending.push_back(r);
    return 0;
}
int parm_prof_end(parm_ctx *c) {
    if (!c->prof_on || c->prof_pending.empty()) return 0;
    CK(cudaEventRecord(c->prof_pending.back().b, c->stream));
    return 0;
}
extern "C" int parm_profile_enable(parm_ctx *c, int enable) {
    CK(cudaSetDevice(c->device));
    c->prof_on = enable != 0;
    return 0;
}
extern "C" int parm_profile_read(parm_ctx *c, double *ms, uint64_t *cnt) {
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    for (auto &r : c->prof_pending) {
        float t = 0;
        CK(cudaEventElapsedTime(&t, r.a, r.b));
        c->prof_ms[r.cls] += t;
        c->prof_cnt[r.cls]++;
        c->prof_pool.push_back(r.a);
        c->prof_pool.push_back(r.b);
    }
    c->prof_pending.clear();
    for (int k = 0; k < PARM_PROF_N; k++) {
        if (ms) ms[k] = c->prof_ms[k];
        if (cnt) cnt[k] = c->prof_cnt[k];
        c->prof_ms[k] = 0;
        c->prof_cnt[k] = 0;
    }
    return 0;
}

extern "C" int parm_host_register(void *ptr, size_t bytes) {
    if (!ptr || !bytes) return 0;
    CK(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
    return 0;
}
extern "C" int parm_host_unregister(void *ptr) {
    if (!ptr) return 0;
    CK(cudaHostUnregister(ptr));
    return 0;
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

// ---- transfers -------------------------------------------------------------
// stage layout: byte offsets of each field for atom 0 + common byte strides
struct StageDesc {
    long long off[5]; // x v a f m; < 0: field absent
    long long stride_vec, stride_m;
};

template <int D>
__global__ void k_scatter_from_stage(const char *__restrict__ stage, StageDesc sd, const uint32_t *__restrict__ slot_of,
                                     double4 *pos, double *v, double *a, double *f, uint32_t n, uint32_t npad) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t s = slot_of[i];
        if (sd.off[0] >= 0 || sd.off[4] >= 0) {
            double4 p = pos[s];
            if (sd.off[0] >= 0) {
                const double *src = (const double *)(stage + sd.off[0] + (long long)i * sd.stride_vec);
                p.x = src[0];
                p.y = src[1];
                p.z = D == 3 ? src[2] : 0.0;
            }
            if (sd.off[4] >= 0) p.w = *(const double *)(stage + sd.off[4] + (long long)i * sd.stride_m);
            pos[s] = p;
        }
        double *dst[3] = {v, a, f};
#pragma unroll
        for (int q = 0; q < 3; q++) {
            if (sd.off[1 + q] < 0) continue;
            const double *src = (const double *)(stage + sd.off[1 + q] + (long long)i * sd.stride_vec);
            dst[q][s] = src[0];
            dst[q][npad + s] = src[1];
            dst[q][2 * (size_t)npad + s] = D == 3 ? src[2] : 0.0;
        }
    }
}

template <int D>
__global__ void k_gather_to_stage(char *__restrict__ stage, StageDesc sd, const uint32_t *__restrict__ slot_of,
                                  const double4 *__restrict__ pos, const double *__restrict__ v,
                                  const double *__restrict__ a, const double *__restrict__ f, uint32_t n, uint32_t npad) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t s = slot_of[i];
        if (sd.off[0] >= 0 || sd.off[4] >= 0) {
            double4 p = pos[s];
            if (sd.off[0] >= 0) {
                double *dst = (double *)(stage + sd.off[0] + (long long)i * sd.stride_vec);
                dst[0] = p.x;
                dst[1] = p.y;
                if (D == 3) dst[2] = p.z;
            }
            if (sd.off[4] >= 0) *(double *)(stage + sd.off[4] + (long long)i * sd.stride_m) = p.w;
        }
        const double *src[3] = {v, a, f};
#pragma unroll
        for (int q = 0; q < 3; q++) {
            if (sd.off[1 + q] < 0) continue;
            double *dst = (double *)(stage + sd.off[1 + q] + (long long)i * sd.stride_vec);
            dst[0] = src[q][s];
            dst[1] = src[q][npad + s];
            if (D == 3) dst[2] = src[q][2 * (size_t)npad + s];
        }
    }
}

// Decide how the host data is laid out. Returns true when all requested fields live
// in one AoS block of n*stride bytes whose every byte belongs to a requested field
// (ParM's struct Atom with the full mask) so that one contiguous copy is valid.
static bool aos_block(const parm_ctx *c, unsigned mask, const double *const p[5], size_t sv, size_t sm, const char **base) {
    if (mask != PARM_ALL || sv != sm) return false;
    size_t vec = (size_t)c->D * 8;
    if (sv != 4 * vec + 8) return false;
    const char *b = (const char *)p[0];
    for (int q = 1; q < 5; q++)
        if ((const char *)p[q] != b + q * vec) return false;
    *base = b;
    return true;
}

static int transfer(parm_ctx *c, bool upload, unsigned mask, const double *const p[5], size_t sv, size_t sm) {
    CK(cudaSetDevice(c->device));
    if (c->sh.on) { parm_set_error("sharded context: use parm_shard_set_atoms / parm_shard_get_atoms"); return PARM_ERR_INVALID; }
    mask &= PARM_ALL;
    if (!mask || c->n == 0) return 0;
    const size_t vec = (size_t)c->D * 8;
    for (int q = 0; q < 5; q++)
        if ((mask >> q & 1) && !p[q]) { parm_set_error("parm_%s_atoms: field %d requested but pointer is NULL", upload ? "upload" : "download", q); return PARM_ERR_INVALID; }
    if (sv < vec || sm < 8) { parm_set_error("parm_%s_atoms: strides smaller than the fields", upload ? "upload" : "download"); return PARM_ERR_INVALID; }
    StageDesc sd;
    const char *base = 0;
    const size_t n = c->n;
    unsigned grid = grid_for(c, c->n, 256, 16);
    if (aos_block(c, mask, p, sv, sm, &base)) {
        PTRY(ensure_stage(c, n * sv));
        for (int q = 0; q < 5; q++) sd.off[q] = (long long)q * (long long)vec;
        sd.stride_vec = sd.stride_m = (long long)sv;
        if (upload) {
            CK(cudaMemcpyAsync(c->d_stage, base, n * sv, cudaMemcpyHostToDevice, c->stream));
            if (c->D == 3) k_scatter_from_stage<3><<<grid, 256, 0, c->stream>>>((const char *)c->d_stage, sd, c->slot_of, c->pos, c->v, c->a, c->f, c->n, c->npad);
            else k_scatter_from_stage<2><<<grid, 256, 0, c->stream>>>((const char *)c->d_stage, sd, c->slot_of, c->pos, c->v, c->a, c->f, c->n, c->npad);
            CK_LAUNCH(c);
        } else {
            if (c->D == 3) k_gather_to_stage<3><<<grid, 256, 0, c->stream>>>((char *)c->d_stage, sd, c->slot_of, c->pos, c->v, c->a, c->f, c->n, c->npad);
            else k_gather_to_stage<2><<<grid, 256, 0, c->stream>>>((char *)c->d_stage, sd, c->slot_of, c->pos, c->v, c->a, c->f, c->n, c->npad);
            CK_LAUNCH(c);
            CK(cudaMemcpyAsync((void *)base, c->d_stage, n * sv, cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
        }
        return 0;
    }
    // general path: each field becomes a dense column block in the stage
    PTRY(ensure_stage(c, n * (4 * vec + 8)));
    size_t off = 0;
    for (int q = 0; q < 5; q++) {
        if (!(mask >> q & 1)) { sd.off[q] = -1; continue; }
        sd.off[q] = (long long)off;
        off += n * (q < 4 ? vec : 8);
    }
    sd.stride_vec = (long long)vec;
    sd.stride_m = 8;
    if (upload) {
        for (int q = 0; q < 5; q++) {
            if (sd.off[q] < 0) continue;
            size_t w = q < 4 ? vec : 8, s = q < 4 ? sv : sm;
            CK(cudaMemcpy2DAsync((char *)c->d_stage + sd.off[q], w, p[q], s, w, n, cudaMemcpyHostToDevice, c->stream));
        }
        if (c->D == 3) k_scatter_from_stage<3><<<grid, 256, 0, c->stream>>>((const char *)c->d_stage, sd, c->slot_of, c->pos, c->v, c->a, c->f, c->n, c->npad);
        else k_scatter_from_stage<2><<<grid, 256, 0, c->stream>>>((const char *)c->d_stage, sd, c->slot_of, c->pos, c->v, c->a, c->f, c->n, c->npad);
        CK_LAUNCH(c);
        // the stage may be overwritten by the next call: host buffers may be pageable, so wait
        CK(cudaStreamSynchronize(c->stream));
    } else {
        if (c->D == 3) k_gather_to_stage<3><<<grid, 256, 0, c->stream>>>((char *)c->d_stage, sd, c->slot_of, c->pos, c->v, c->a, c->f, c->n, c->npad);
        else k_gather_to_stage<2><<<grid, 256, 0, c->stream>>>((char *)c->d_stage, sd, c->slot_of, c->pos, c->v, c->a, c->f, c->n, c->npad);
        CK_LAUNCH(c);
        for (int q = 0; q < 5; q++) {
            if (sd.off[q] < 0) continue;
            size_t w = q < 4 ? vec : 8, s = q < 4 ? sv : sm;
            CK(cudaMemcpy2DAsync((void *)p[q], s, (char *)c->d_stage + sd.off[q], w, w, n, cudaMemcpyDeviceToHost, c->stream));
        }
        CK(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

extern "C" int parm_upload_atoms(parm_ctx *c, unsigned mask, const double *x, const double *v, const double *a,
                                 const double *f, const double *m, size_t stride_vec, size_t stride_m) {
    if (!c) { parm_set_error("parm_upload_atoms: NULL context"); return PARM_ERR_INVALID; }
    const double *p[5] = {x, v, a, f, m};
    return transfer(c, true, mask, p, stride_vec, stride_m);
}
extern "C" int parm_download_atoms(parm_ctx *c, unsigned mask, double *x, double *v, double *a, double *f, double *m,
                                   size_t stride_vec, size_t stride_m) {
    if (!c) { parm_set_error("parm_download_atoms: NULL context"); return PARM_ERR_INVALID; }
    const double *p[5] = {x, v, a, f, m};
    return transfer(c, false, mask, p, stride_vec, stride_m);
}

// ---- OriginBox::diff on the device (box.hpp:103) -----------------------------
__global__ void k_box_diff(const double *r1, const double *r2, double *out, uint32_t npts, int D, BoxDev box) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < npts * D; i += gridDim.x * blockDim.x) {
        int k = i % D;
        out[i] = min_image_exact(__dsub_rn(r1[i], r2[i]), box.L[k], box.invL[k], box.halfL[k]);
    }
}
extern "C" int parm_box_diff(parm_ctx *c, uint32_t npts, const double *r1, const double *r2, double *out) {
    CK(cudaSetDevice(c->device));
    if (!npts) return 0;
    size_t nb = (size_t)npts * c->D * 8;
    PTRY(ensure_stage(c, 3 * nb));
    char *s = (char *)c->d_stage;
    CK(cudaMemcpyAsync(s, r1, nb, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(s + nb, r2, nb, cudaMemcpyHostToDevice, c->stream));
    k_box_diff<<<grid_for(c, npts * c->D, 256), 256, 0, c->stream>>>((double *)s, (double *)(s + nb), (double *)(s + 2 * nb), npts, c->D, c->box);
    CK_LAUNCH(c);
    CK(cudaMemcpyAsync(out, s + 2 * nb, nb, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

// ---- reductions ----------------------------------------------------------------
// Deterministic: fixed grid, per-block tree, then one block folds the partials.
#define RED_BLOCK 256
#define RED_MAXQ 4

__device__ __forceinline__ void block_sum(double *vals, int nq, double *smem /*[RED_MAXQ][8]*/) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int q = 0; q < nq; q++) {
        double x = vals[q];
#pragma unroll
        for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) smem[q * 8 + w] = x;
    }
    __syncthreads();
    if (w == 0) {
        for (int q = 0; q < nq; q++) {
            double x = lane < (RED_BLOCK / 32) ? smem[q * 8 + lane] : 0.0;
#pragma unroll
            for (int o = 4; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            vals[q] = x;
        }
    }
    __syncthreads();
}

template <int D>
__global__ void k_reduce(int what, const double4 *__restrict__ pos, const double *__restrict__ v,
                         const double *__restrict__ f, uint32_t n, uint32_t npad, double v0x, double v0y, double v0z,
                         double *partials) {
    __shared__ double smem[RED_MAXQ * 8];
    double acc[RED_MAXQ] = {0, 0, 0, 0};
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        double4 p = pos[s];
        double m = p.w;
        if (what == PARM_RED_MASS) {
            if (!frozen_le(m)) acc[0] += m;
        } else if (what == PARM_RED_MOMENTUM) {
            if (!frozen_le(m)) {
                acc[0] += v[s] * m;
                acc[1] += v[npad + s] * m;
                if (D == 3) acc[2] += v[2 * (size_t)npad + s] * m;
            }
        } else if (what == PARM_RED_KE) { // box.cpp:401-411: m/2 * curv.dot(curv), dot = e0+(e1+e2)
            if (!frozen_eq(m)) {
                double cx = v[s] - v0x, cy = v[npad + s] - v0y, cz = D == 3 ? v[2 * (size_t)npad + s] - v0z : 0.0;
                double d = D == 3 ? __dadd_rn(__dmul_rn(cx, cx), __dadd_rn(__dmul_rn(cy, cy), __dmul_rn(cz, cz)))
                                  : __dadd_rn(__dmul_rn(cx, cx), __dmul_rn(cy, cy));
                acc[0] += __dmul_rn(m / 2, d);
            }
        } else if (what == PARM_RED_COM) {
            if (!frozen_le(m)) {
                acc[0] += p.x * m;
                acc[1] += p.y * m;
                if (D == 3) acc[2] += p.z * m;
                acc[3] += m;
            }
        } else if (what == PARM_RED_NDOF) {
            if (!frozen_le(m)) acc[0] += (double)D;
        } else if (what == PARM_RED_COMFORCE) {
            acc[0] += f[s];
            acc[1] += f[npad + s];
            if (D == 3) acc[2] += f[2 * (size_t)npad + s];
        }
    }
    block_sum(acc, RED_MAXQ, smem);
    if (threadIdx.x == 0)
        for (int q = 0; q < RED_MAXQ; q++) partials[blockIdx.x * RED_MAXQ + q] = acc[q];
}

__global__ void k_reduce_final(const double *partials, int nblocks, double *out) {
    __shared__ double smem[RED_MAXQ * 8];
    double acc[RED_MAXQ] = {0, 0, 0, 0};
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x)
        for (int q = 0; q < RED_MAXQ; q++) acc[q] += partials[b * RED_MAXQ + q];
    block_sum(acc, RED_MAXQ, smem);
    if (threadIdx.x == 0)
        for (int q = 0; q < RED_MAXQ; q++) out[q] = acc[q];
}

extern "C" int parm_reduce(parm_ctx *c, int what, const double *v0, double *out) {
    if (!c || !out) { parm_set_error("parm_reduce: NULL argument"); return PARM_ERR_INVALID; }
    if (what < 0 || what > PARM_RED_COMFORCE) { parm_set_error("parm_reduce: unknown quantity %d", what); return PARM_ERR_INVALID; }
    CK(cudaSetDevice(c->device));
    const uint32_t no = parm_owned(c);
    int nb = (int)grid_for(c, no, RED_BLOCK, 4);
    PTRY(parm_ctx_ensure_red(c, (size_t)(nb + 1) * RED_MAXQ));
    double z[3] = {0, 0, 0};
    if (v0)
        for (int k = 0; k < c->D; k++) z[k] = v0[k];
    if (c->D == 3) k_reduce<3><<<nb, RED_BLOCK, 0, c->stream>>>(what, c->pos, c->v, c->f, no, c->npad, z[0], z[1], z[2], c->d_red + RED_MAXQ);
    else k_reduce<2><<<nb, RED_BLOCK, 0, c->stream>>>(what, c->pos, c->v, c->f, no, c->npad, z[0], z[1], z[2], c->d_red + RED_MAXQ);
    CK_LAUNCH(c);
    k_reduce_final<<<1, RED_BLOCK, 0, c->stream>>>(c->d_red + RED_MAXQ, nb, c->d_red);
    CK_LAUNCH(c);
    if (c->sh.on) PTRY(parm_shard_allreduce_sum(c, c->d_red, RED_MAXQ));
    CK(cudaMemcpyAsync(c->h_red, c->d_red, RED_MAXQ * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    switch (what) {
        case PARM_RED_MASS: case PARM_RED_KE: case PARM_RED_NDOF: out[0] = c->h_red[0]; break;
        case PARM_RED_MOMENTUM: case PARM_RED_COMFORCE:
            for (int k = 0; k < c->D; k++) out[k] = c->h_red[k];
            break;
        case PARM_RED_COM:
            for (int k = 0; k < c->D; k++) out[k] = c->h_red[k] / c->h_red[3];
            break;
    }
    return 0;
}

template <int D>
__global__ void k_scale_v(const double4 *__restrict__ pos, double *v, uint32_t n, uint32_t npad, double s) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (frozen_le(pos[i].w)) continue; // collection.cpp:24-26
        v[i] *= s;
        v[npad + i] *= s;
        if (D == 3) v[2 * (size_t)npad + i] *= s;
    }
}
extern "C" int parm_scale_velocities(parm_ctx *c, double s) {
    CK(cudaSetDevice(c->device));
    const uint32_t no = parm_owned(c);
    if (c->D == 3) k_scale_v<3><<<grid_for(c, no, 256), 256, 0, c->stream>>>(c->pos, c->v, no, c->npad, s);
    else k_scale_v<2><<<grid_for(c, no, 256), 256, 0, c->stream>>>(c->pos, c->v, no, c->npad, s);
    CK_LAUNCH(c);
    return 0;
}

__global__ void k_add_v(double *v, uint32_t n, uint32_t npad, double dx, double dy, double dz, int D) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        v[i] += dx; // box.cpp:413-417: every atom, frozen or not
        v[npad + i] += dy;
        if (D == 3) v[2 * (size_t)npad + i] += dz;
    }
}
extern "C" int parm_add_velocity(parm_ctx *c, const double *dv) {
    CK(cudaSetDevice(c->device));
    k_add_v<<<grid_for(c, parm_owned(c), 256), 256, 0, c->stream>>>(c->v, parm_owned(c), c->npad, dv[0], dv[1], c->D == 3 ? dv[2] : 0.0, c->D);
    CK_LAUNCH(c);
    return 0;
}

extern "C" int parm_reset_forces(parm_ctx *c) {
    CK(cudaSetDevice(c->device));
    CK(cudaMemsetAsync(c->f, 0, 3 * (size_t)c->npad * 8, c->stream));
    return 0;
}
