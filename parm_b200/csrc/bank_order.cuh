// Bank-aware order of the 16-bit tile-local rows (EXPERIMENTAL, PARM_B200_TILE_BANKS=1, off by default; DESIGN.md
// section 9 lead 2, model: tools/bank_model.py).
//
// The pair kernel (force_tile.cuh, TEAM = 4, V = 8) reads, at step g of a row, entry 4 g + tl with lane tl of the
// atom's team; two teams share every LDS.128 phase (8 bank groups of 16 bytes) and four teams every LDS.64 phase
// (16 banks of 8 bytes). With rows in build order those reads hit effectively random banks: 10.1 wavefronts per warp
// step against 6 without conflicts. Here every row is re-ordered so that lane tl of team q (q = atom index in its
// chunk mod 4 = position of the team in its half-warp) finds, at step g, an entry of class
//        c(g, tl) = 4 * ((q + g) & 3) + tl          of   (tile index mod 16):
// the four teams of a half-warp then cover all 16 banks, the two teams of a quarter-warp all 8 bank groups.
// A class that has more entries than steps (G / 4 per class) moves its surplus to the sibling class c ^ 8 (same lane,
// same 16-byte bank group: only the z read can still collide), what is left goes to any free slot. Free slots keep a
// sentinel of THEIR class (16 sentinel slots S .. S + 15 behind the tile, S a multiple of 16), so pads never collide.
// The order inside a class is the build order, everything is integer and lane-local or exchanged through fixed
// positions: the result is deterministic.
//
// The two phases are plain functions of one lane so that tests/host/bank_order_test.cpp can run them on the CPU.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define BO_HD __host__ __device__ __forceinline__
#else
#define BO_HD inline
#endif

#define BO_RCAP 24 // surplus entries one lane may carry out of phase 1 (more: the row keeps its build order)

// position, inside a row in the pair kernel's lane-vector layout, of the entry lane tl reads at step g
BO_HD uint32_t bo_slot(uint32_t g, uint32_t tl) { return (g >> 3) * 32u + tl * 8u + (g & 7u); }
// natural (build order) entry k of a row stored in that layout
BO_HD uint32_t bo_nat(uint32_t k) { return (k & ~31u) + (k & 3u) * 8u + ((k & 31u) >> 2); }
BO_HD uint32_t bo_cnt(uint32_t packed, uint32_t h) { return (packed >> (8u * h)) & 0xffu; }

// Phase 1, lane tl: scans the whole row (in: lane-vector layout, `my` real entries, all < ntile), places the entries of
// its four classes into its own slots of `out` (pre-filled with class sentinels by the caller), then the surplus of a
// class into the free slots of the sibling class. Returns the packed per-subclass fill counts; rest[] / *nrest get the
// entries that found no slot in this lane. *overflow is set when more than BO_RCAP entries were left over.
BO_HD uint32_t bo_phase1(const uint16_t *in, uint32_t my, uint32_t G, uint32_t q, uint32_t tl, uint16_t *out, uint16_t *rest,
                         uint32_t *nrest, bool *overflow) {
    const uint32_t D = G >> 2; // steps per class
    uint32_t packed = 0, ns = 0;
    uint16_t spill[BO_RCAP];
    for (uint32_t k = 0; k < my; k++) {
        const uint32_t e = in[bo_nat(k)];
        if ((e & 3u) != tl) continue;
        const uint32_t h = (e >> 2) & 3u, kk = bo_cnt(packed, h);
        if (kk < D) {
            out[bo_slot(4u * kk + ((h - q) & 3u), tl)] = (uint16_t)e;
            packed += 1u << (8u * h);
        } else if (ns < BO_RCAP) {
            spill[ns++] = (uint16_t)e;
        } else {
            *overflow = true;
        }
    }
    uint32_t nr = 0;
    for (uint32_t i = 0; i < ns; i++) {
        const uint32_t e = spill[i], h2 = ((e >> 2) & 3u) ^ 2u, kk = bo_cnt(packed, h2);
        if (kk < D) {
            out[bo_slot(4u * kk + ((h2 - q) & 3u), tl)] = (uint16_t)e;
            packed += 1u << (8u * h2);
        } else {
            rest[nr++] = (uint16_t)e;
        }
    }
    *nrest = nr;
    return packed;
}

// Phase 2, lane tl: the entries left over by all four lanes (rest[lane * BO_RCAP + i], nrest[lane]) are dealt, in lane
// order, to the free slots of the lanes, in lane order; this lane fills its own share. packed[lane]: phase-1 counts.
BO_HD void bo_phase2(uint32_t G, uint32_t q, uint32_t tl, const uint32_t *packed, const uint16_t *rest, const uint32_t *nrest,
                     uint16_t *out) {
    const uint32_t D = G >> 2;
    uint32_t R = 0, hole0 = 0;
    for (uint32_t l = 0; l < 4; l++) {
        R += nrest[l];
        if (l < tl)
            for (uint32_t h = 0; h < 4; h++) hole0 += D - bo_cnt(packed[l], h);
    }
    if (hole0 >= R) return; // the lanes before this one absorb everything
    uint32_t idx = hole0;  // index, in the concatenated rest list, of the entry the next free slot of this lane takes
    for (uint32_t h = 0; h < 4 && idx < R; h++)
        for (uint32_t kk = bo_cnt(packed[tl], h); kk < D && idx < R; kk++, idx++) {
            uint32_t l = 0, i = idx;
            while (i >= nrest[l]) { i -= nrest[l]; l++; }
            out[bo_slot(4u * kk + ((h - q) & 3u), tl)] = rest[l * BO_RCAP + i];
        }
}

// class sentinel of the slot (g, tl) of team q; S: first of the 16 sentinel tile slots (a multiple of 16)
BO_HD uint16_t bo_sentinel(uint32_t S, uint32_t g, uint32_t q, uint32_t tl) { return (uint16_t)(S + 4u * ((q + g) & 3u) + tl); }
