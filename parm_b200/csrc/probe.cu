// Device peak probes for bench.py's roofline denominators ("fp64 CUDA-core throughput ... builder
// must measure it with an FMA-chain microbenchmark and record it"). Not part of the MD path: two tiny kernels
// timed with CUDA events on their own stream.
//   fp64: every thread runs 8 independent DFMA chains; flops = 2 per DFMA.
//   hbm : grid-stride copy of double4 vectors over buffers far larger than the 126 MB L2 (read + write bytes),
//         the same access pattern as the streaming integrator kernels K1/K3.
#include "internal.cuh"

__global__ void __launch_bounds__(256) k_probe_fp64(double *out, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 12345.678) out[blockIdx.x] = s; // never true: keeps the chains alive
}

__global__ void __launch_bounds__(256) k_probe_copy(const double4 *__restrict__ src, double4 *__restrict__ dst, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) { // four independent 32-byte loads in flight per thread
        const double4 a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
        dst[i] = a;
        dst[i + stride] = b;
        dst[i + 2 * stride] = c;
        dst[i + 3 * stride] = d;
    }
    for (; i < n; i += stride) dst[i] = src[i];
}

extern "C" int parm_b200_probe_peaks(int device, double *fp64_gflops, double *copy_gbs) {
    if (!fp64_gflops || !copy_gbs) {
        parm_set_error("parm_b200_probe_peaks: null output pointer");
        return PARM_ERR_INVALID;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
        parm_set_error("parm_b200_probe_peaks: no usable CUDA device %d (this library has no CPU fallback)", device);
        return PARM_ERR_CUDA;
    }
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    double *d_out = nullptr;
    const int blocks = prop.multiProcessorCount * 8, iters = 1024; // ~1 ms per launch
    CK(cudaMalloc(&d_out, sizeof(double) * blocks));
    float best = 1e30f, ms = 0;
    for (int rep = 0; rep < 5; rep++) { // first repetitions warm the clocks up
        CK(cudaEventRecord(e0, st));
        k_probe_fp64<<<blocks, 256, 0, st>>>(d_out, iters, 1.0 + rep);
        CK(cudaEventRecord(e1, st));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep >= 2 && ms < best) best = ms;
    }
    parm_count_launch(nullptr, 5);
    *fp64_gflops = 2.0 * 64.0 * iters * 256.0 * blocks / (best * 1e-3) / 1e9;
    CK(cudaFree(d_out));

    const size_t n = (size_t)1 << 25; // 2^25 double4 = 1 GiB per buffer
    double4 *a = nullptr, *b = nullptr;
    CK(cudaMalloc(&a, n * sizeof(double4)));
    CK(cudaMalloc(&b, n * sizeof(double4)));
    CK(cudaMemsetAsync(a, 0, n * sizeof(double4), st));
    best = 1e30f;
    for (int rep = 0; rep < 6; rep++) {
        CK(cudaEventRecord(e0, st));
        k_probe_copy<<<prop.multiProcessorCount * 16, 256, 0, st>>>(a, b, n);
        CK(cudaEventRecord(e1, st));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep >= 2 && ms < best) best = ms;
    }
    parm_count_launch(nullptr, 6);
    *copy_gbs = 2.0 * n * sizeof(double4) / (best * 1e-3) / 1e9;
    CK(cudaFree(a));
    CK(cudaFree(b));
    CK(cudaEventDestroy(e0));
    CK(cudaEventDestroy(e1));
    CK(cudaStreamDestroy(st));
    return PARM_OK;
}
