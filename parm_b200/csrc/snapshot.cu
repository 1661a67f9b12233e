// Trajectory output and the Grid cell structure: the two analysis-side helpers of SURVEY 8(f)4 that read the atoms
// without stopping the time-stepping.
//
// * Asynchronous frames (LJatoms.cpp:130-158 writes an XYZ frame every few hundred steps, pyparm/xyzfile.py:7-76 the
//   same from Python): parm_snapshot_begin gathers x (and v) by AtomVec index into a device staging buffer on the
//   context's stream -- a few microseconds -- and starts the device->host copy on a SECOND stream into page-locked
//   memory; timestep() calls issued afterwards overlap that copy. parm_snapshot_wait hands the frame to the caller.
// * Grid (trackers.hpp:227-309, trackers.cpp:192-219): the cell index of every atom is computed on the device from the
//   resident positions with the reference's get_loc rule (IEEE remainder image, `== widths -> 0`), so only 4 bytes per
//   atom travel; the facade builds the per-cell lists from them.
#include <string.h>
#include <algorithm>

#include "internal.cuh"

template <int D>
__global__ void k_snapshot_gather(const uint32_t *__restrict__ slot_of, const double4 *__restrict__ pos, const double *__restrict__ v,
                                  uint32_t n, uint32_t npad, unsigned mask, double *__restrict__ out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t s = slot_of[i];
        size_t o = 0;
        if (mask & PARM_X) {
            const double4 p = pos[s];
            double *dst = out + (size_t)i * D;
            dst[0] = p.x;
            dst[1] = p.y;
            if (D == 3) dst[2] = p.z;
            o = (size_t)n * D;
        }
        if (mask & PARM_V) {
            double *dst = out + o + (size_t)i * D;
            dst[0] = v[s];
            dst[1] = v[npad + s];
            if (D == 3) dst[2] = v[2 * (size_t)npad + s];
        }
    }
}

void parm_snapshot_free(parm_ctx *c) {
    if (!c->snap_init) return;
    cudaStreamSynchronize(c->snap_stream);
    cudaStreamDestroy(c->snap_stream);
    cudaEventDestroy(c->snap_ready);
    cudaEventDestroy(c->snap_done);
    if (c->d_snap) cudaFree(c->d_snap);
    if (c->h_snap) cudaFreeHost(c->h_snap);
    c->snap_init = false;
}

extern "C" int parm_snapshot_begin(parm_ctx *c, unsigned mask) {
    if (!c) { parm_set_error("parm_snapshot_begin: NULL context"); return PARM_ERR_INVALID; }
    CK(cudaSetDevice(c->device));
    if (c->sh.on) { parm_set_error("parm_snapshot_begin: slab-decomposed contexts use parm_shard_get_atoms"); return PARM_ERR_UNSUPPORTED; }
    mask &= PARM_X | PARM_V;
    if (!mask) { parm_set_error("parm_snapshot_begin: mask must hold PARM_X and/or PARM_V"); return PARM_ERR_INVALID; }
    if (c->snap_pending) { parm_set_error("parm_snapshot_begin: the previous frame has not been collected (parm_snapshot_wait)"); return PARM_ERR_INVALID; }
    if (!c->snap_init) {
        CK(cudaStreamCreateWithFlags(&c->snap_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&c->snap_ready, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->snap_done, cudaEventDisableTiming));
        c->snap_init = true;
    }
    const size_t need = 2 * (size_t)c->n * c->D;
    if (need > c->snap_doubles) {
        if (c->d_snap) cudaFree(c->d_snap);
        if (c->h_snap) cudaFreeHost(c->h_snap);
        c->d_snap = c->h_snap = nullptr;
        c->snap_doubles = 0;
        CK(cudaMalloc(&c->d_snap, need * 8));
        CK(cudaHostAlloc(&c->h_snap, need * 8, cudaHostAllocDefault));
        c->snap_doubles = need;
    }
    c->snap_mask = mask;
    if (c->n) {
        const unsigned grid = std::min<unsigned>((c->n + 255) / 256, (unsigned)c->num_sms * 8u);
        // the staging buffer may still be draining the previous frame's copy
        CK(cudaStreamWaitEvent(c->stream, c->snap_done, 0));
        if (c->D == 3) k_snapshot_gather<3><<<grid, 256, 0, c->stream>>>(c->slot_of, c->pos, c->v, c->n, c->npad, mask, c->d_snap);
        else k_snapshot_gather<2><<<grid, 256, 0, c->stream>>>(c->slot_of, c->pos, c->v, c->n, c->npad, mask, c->d_snap);
        CK_LAUNCH(c);
        CK(cudaEventRecord(c->snap_ready, c->stream));
        CK(cudaStreamWaitEvent(c->snap_stream, c->snap_ready, 0));
        const size_t nd = (size_t)c->n * c->D * ((mask & PARM_X ? 1 : 0) + (mask & PARM_V ? 1 : 0));
        CK(cudaMemcpyAsync(c->h_snap, c->d_snap, nd * 8, cudaMemcpyDeviceToHost, c->snap_stream));
        CK(cudaEventRecord(c->snap_done, c->snap_stream));
    }
    c->snap_pending = true;
    return 0;
}

extern "C" int parm_snapshot_wait(parm_ctx *c, double *x, double *v) {
    if (!c) { parm_set_error("parm_snapshot_wait: NULL context"); return PARM_ERR_INVALID; }
    CK(cudaSetDevice(c->device));
    if (!c->snap_pending) { parm_set_error("parm_snapshot_wait: no frame was started"); return PARM_ERR_INVALID; }
    if (c->n) CK(cudaEventSynchronize(c->snap_done));
    const size_t nd = (size_t)c->n * c->D;
    size_t o = 0;
    if (c->snap_mask & PARM_X) {
        if (x) memcpy(x, c->h_snap, nd * 8);
        o = nd;
    }
    if ((c->snap_mask & PARM_V) && v) memcpy(v, c->h_snap + o, nd * 8);
    c->snap_pending = false;
    return 0;
}

// ---- Grid::get_loc for every atom (trackers.cpp:192-219) ------------------------------------------------------------
template <int D>
__global__ void k_grid_loc(const uint32_t *__restrict__ slot_of, const double4 *__restrict__ pos, uint32_t n, BoxDev box,
                           uint32_t w0, uint32_t w1, uint32_t w2, uint32_t *__restrict__ loc) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double4 p = pos[slot_of[i]];
        const double x[3] = {p.x, p.y, p.z};
        const uint32_t w[3] = {w0, w1, w2};
        uint32_t k[3] = {0, 0, 0};
#pragma unroll
        for (int d = 0; d < D; d++) {
            // v = vec_mod(v - bsize/2, bsize) + bsize/2;  floor(v * widths / bsize);  == widths -> 0
            const double h = __ddiv_rn(box.L[d], 2.0);
            const double r = __dadd_rn(min_image_exact(__dadd_rn(x[d], -h), box.L[d], box.invL[d], box.halfL[d]), h);
            uint32_t q = (uint32_t)floor(__ddiv_rn(__dmul_rn(r, (double)w[d]), box.L[d]));
            if (q == w[d]) q = 0;
            k[d] = q;
        }
        loc[i] = D == 3 ? (k[2] * w1 + k[1]) * w0 + k[0] : k[1] * w0 + k[0];
    }
}

extern "C" int parm_grid_locs(parm_ctx *c, const uint32_t *widths, uint32_t *loc) {
    if (!c || !widths || !loc) { parm_set_error("parm_grid_locs: NULL argument"); return PARM_ERR_INVALID; }
    CK(cudaSetDevice(c->device));
    if (c->sh.on) { parm_set_error("parm_grid_locs: not available on slab-decomposed contexts"); return PARM_ERR_UNSUPPORTED; }
    if (!c->box_set) { parm_set_error("parm_grid_locs: the box has not been set"); return PARM_ERR_INVALID; }
    for (int d = 0; d < c->D; d++)
        if (!widths[d]) { parm_set_error("parm_grid_locs: widths must be positive"); return PARM_ERR_INVALID; }
    if (!c->n) return 0;
    uint32_t *d_loc = nullptr;
    CK(cudaMalloc(&d_loc, (size_t)c->n * 4));
    const unsigned grid = std::min<unsigned>((c->n + 255) / 256, (unsigned)c->num_sms * 8u);
    if (c->D == 3) k_grid_loc<3><<<grid, 256, 0, c->stream>>>(c->slot_of, c->pos, c->n, c->box, widths[0], widths[1], widths[2], d_loc);
    else k_grid_loc<2><<<grid, 256, 0, c->stream>>>(c->slot_of, c->pos, c->n, c->box, widths[0], widths[1], 1u, d_loc);
    parm_count_launch(c);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(loc, d_loc, (size_t)c->n * 4, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_loc);
    CK(e);
    return 0;
}
