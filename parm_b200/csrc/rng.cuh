// Counter-based Gaussian noise for the Langevin integrators (CollectionSol, CollectionSolHT).
#pragma once
#include <stdint.h>

// ---- Philox4x32-10 counter RNG + Box-Muller (production noise of CollectionSol) ---------
__device__ __forceinline__ void philox4x32(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0;
        c[1] = lo1;
        c[2] = n2;
        c[3] = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
__device__ __forceinline__ void normal_pair(uint32_t id, uint64_t step, uint32_t stream, uint64_t seed, double &z0, double &z1) {
    uint32_t c[4] = {id, (uint32_t)step, (uint32_t)(step >> 32), stream};
    philox4x32(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    // 53-bit uniforms: u1 in (0,1], u2 in [0,1)
    double u1 = ((double)(((uint64_t)c[0] << 21) ^ (c[1] >> 11)) + 1.0) * (1.0 / 9007199254740992.0);
    double u2 = (double)(((uint64_t)c[2] << 21) ^ (c[3] >> 11)) * (1.0 / 9007199254740992.0);
    double r = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    z0 = r * cs;
    z1 = r * sn;
}

