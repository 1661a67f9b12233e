// Microbenchmark: how long does one CTA wait for a 54 KB tile that arrives as N cp.async.bulk copies (B200)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_copies tma_copies.cu && ./tma_copies
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// every block: `reps` tiles of `total` bytes, each as `ncopies` copies from scattered places of a big buffer
__global__ void k(const char *src, size_t src_bytes, uint32_t total, uint32_t ncopies, uint32_t reps, unsigned long long *cyc) {
    extern __shared__ __align__(128) char smem[];
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) mbar_init(b, 1);
    __syncthreads();
    const uint32_t piece = (total / ncopies) & ~15u;
    unsigned long long t0 = clock64();
    for (uint32_t r = 0; r < reps; r++) {
        if (threadIdx.x < 32) {
            if (threadIdx.x == 0) mbar_arrive_expect_tx(b, piece * ncopies);
            __syncwarp();
            for (uint32_t i = threadIdx.x; i < ncopies; i += 32) {
                size_t off = ((size_t)(blockIdx.x * 7919u + r * 104729u + i * 15485863u) * 4096u) % (src_bytes - total);
                off &= ~(size_t)15;
                bulk_g2s((uint32_t)__cvta_generic_to_shared(smem + (size_t)i * piece), src + off, piece, b);
            }
        }
        mbar_wait(b, r & 1u);
        __syncthreads();
    }
    unsigned long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    const size_t src_bytes = 1ull << 30;
    char *src; cudaMalloc(&src, src_bytes); cudaMemset(src, 1, src_bytes);
    unsigned long long *cyc; cudaMalloc(&cyc, 4096 * 8);
    const uint32_t total = 54 * 1024, reps = 200;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 56 * 1024);
    for (int bps : {1, 2, 4}) {
        for (uint32_t nc : {1u, 2u, 4u, 9u, 18u, 36u, 72u}) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            k<<<148 * bps, 256, total, 0>>>(src, src_bytes, total, nc, 10, cyc);
            cudaEventRecord(e0);
            k<<<148 * bps, 256, total, 0>>>(src, src_bytes, total, nc, reps, cyc);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("blocks/SM %d copies/tile %2u piece %6u B: %.3f us per tile per block, %.1f GB/s total\n", bps, nc, (total / nc) & ~15u,
                   ms * 1e3 / reps, (double)((total / nc) & ~15u) * nc * reps * 148 * bps / (ms * 1e-3) / 1e9);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
