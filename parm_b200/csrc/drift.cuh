// Skin-drift trigger of NeighborList::update_list(false) (trackers.cpp:23-53) as a parallel
// reduction. The reference keeps a running top-2 of |x_i - lastlocs_i| over a sequential
// loop and breaks as soon as bigdist + biggestdist >= skin; a prefix top-2 is monotone, so
// that is exactly "the two largest displacements of the whole set sum to >= skin".
// Displacements use raw unwrapped positions (trackers.cpp:27), association e0 + (e1 + e2).
#pragma once
#include "internal.cuh"

__device__ __forceinline__ double drift_dist(const double4 &p, double x0, double y0, double z0) {
    double dx = __dsub_rn(p.x, x0), dy = __dsub_rn(p.y, y0), dz = __dsub_rn(p.z, z0);
    return __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dadd_rn(__dmul_rn(dy, dy), __dmul_rn(dz, dz))));
}
// (b1 >= b2) running top-2, same comparisons as trackers.cpp:28-35 (NaN never enters)
__device__ __forceinline__ void top2_push(double &b1, double &b2, double d) {
    if (d > b1) {
        b2 = b1;
        b1 = d;
    } else if (d > b2) {
        b2 = d;
    }
}
__device__ __forceinline__ void top2_merge(double &b1, double &b2, double o1, double o2) {
    top2_push(b1, b2, o1);
    top2_push(b1, b2, o2);
}
__device__ __forceinline__ void top2_warp(double &b1, double &b2) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        double o1 = __shfl_xor_sync(0xffffffffu, b1, o), o2 = __shfl_xor_sync(0xffffffffu, b2, o);
        top2_merge(b1, b2, o1, o2);
    }
}

// Block top-2, publish per-block result, and let the last block to arrive fold all blocks
// and write the rebuild flag (to device memory and to the pinned host mirror).
// Must be called by every thread of every block of the grid (blockDim.x <= 1024).
// d_slot / h_slot (optional): per-step decision words read by the NEXT step's kernels (abort guard) and by
// the host, so the host can enqueue step s+1 before it has seen the decision of step s.
__device__ __forceinline__ void drift_finish(double b1, double b2, double skin, double *d_top2, unsigned int *counter,
                                             NlistFlags *dflags, NlistFlags *hflags, int *d_slot = nullptr,
                                             int *h_slot = nullptr, int step_no = -1) {
    __shared__ double s1[32], s2[32];
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    top2_warp(b1, b2);
    if (lane == 0) {
        s1[w] = b1;
        s2[w] = b2;
    }
    __syncthreads();
    if (w == 0) {
        b1 = lane < nw ? s1[lane] : 0.0;
        b2 = lane < nw ? s2[lane] : 0.0;
        top2_warp(b1, b2);
        if (lane == 0) {
            d_top2[2 * blockIdx.x] = b1;
            d_top2[2 * blockIdx.x + 1] = b2;
            __threadfence();
            unsigned int t = atomicInc(counter, gridDim.x - 1); // wraps back to 0 for the next launch
            s_last = (t == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    b1 = 0.0;
    b2 = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x)
        top2_merge(b1, b2, __ldcg(d_top2 + 2 * b), __ldcg(d_top2 + 2 * b + 1));
    top2_warp(b1, b2);
    if (lane == 0) {
        s1[w] = b1;
        s2[w] = b2;
    }
    __syncthreads();
    if (w == 0) {
        b1 = lane < nw ? s1[lane] : 0.0;
        b2 = lane < nw ? s2[lane] : 0.0;
        top2_warp(b1, b2);
        if (lane == 0) {
            int need = (__dadd_rn(b2, b1) >= skin) ? 1 : 0; // bigdist + biggestdist >= skin
            dflags->need_rebuild = need;
            dflags->top2[0] = b1;
            dflags->top2[1] = b2;
            hflags->top2[0] = b1;
            hflags->top2[1] = b2;
            hflags->need_rebuild = need;
            if (d_slot) *d_slot = need;
            if (h_slot) *h_slot = need;
            // batched stepping (parm_integ_timestep): the first step of the batch that asks for a rebuild leaves its
            // number for the host (the steps behind it abort, so there is no second writer)
            if (need && step_no >= 0) hflags->trigger = (uint32_t)step_no;
            __threadfence_system();
        }
    }
}
