// CPU check of parm_b200/csrc/bank_order.cuh (the per-lane phases of the bank-aware row order are plain functions):
// every entry of a row survives the re-ordering exactly once, free slots carry the sentinel of their class, and the
// replayed LDS.128 / LDS.64 wavefront counts of random 8-team warps drop the way tools/bank_model.py predicts.
//   g++ -O2 -std=c++17 -I parm_b200/csrc tests/host/bank_order_test.cpp -o /tmp/bank_order_test && /tmp/bank_order_test
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <random>
#include <set>
#include <vector>

#include "bank_order.cuh"

static std::mt19937 rng(12345);

struct Row {
    std::vector<uint16_t> in, out;
    uint32_t my, G;
};

static Row make_row(uint32_t ntile, uint32_t S, uint32_t my, uint32_t q, bool &overflow) {
    Row r;
    r.my = my;
    const uint32_t mypad = (my + 31u) & ~31u;
    r.G = mypad / 4;
    r.in.assign(mypad, (uint16_t)ntile);
    // neighbours: distinct tile indices, ascending inside 9 "columns" like the build order
    std::set<uint32_t> pick;
    std::uniform_int_distribution<uint32_t> U(0, ntile - 1);
    while (pick.size() < my) pick.insert(U(rng));
    std::vector<uint32_t> nat(pick.begin(), pick.end());
    for (uint32_t k = 0; k < my; k++) r.in[bo_nat(k)] = (uint16_t)nat[k];
    r.out.assign(mypad, 0);
    for (uint32_t g = 0; g < r.G; g++)
        for (uint32_t tl = 0; tl < 4; tl++) r.out[bo_slot(g, tl)] = bo_sentinel(S, g, q, tl);
    uint32_t packed[4], nrest[4];
    uint16_t rest[4 * BO_RCAP];
    overflow = false;
    for (uint32_t tl = 0; tl < 4; tl++) packed[tl] = bo_phase1(r.in.data(), my, r.G, q, tl, r.out.data(), rest + tl * BO_RCAP, &nrest[tl], &overflow);
    if (!overflow)
        for (uint32_t tl = 0; tl < 4; tl++) bo_phase2(r.G, q, tl, packed, rest, nrest, r.out.data());
    return r;
}

// wavefronts of one warp step: idx[team][lane]; 128-bit reads in quarter-warp phases (16-byte bank groups: idx mod 8),
// 64-bit reads in half-warp phases (8-byte banks: idx mod 16); a phase costs the largest number of distinct addresses
// that share a bank
static void wavefronts(const uint32_t idx[8][4], double &w128, double &w64) {
    for (int ph = 0; ph < 4; ph++) {
        std::map<uint32_t, std::set<uint32_t> > b;
        for (int t = 2 * ph; t < 2 * ph + 2; t++)
            for (int l = 0; l < 4; l++) b[idx[t][l] % 8].insert(idx[t][l]);
        size_t m = 0;
        for (auto &kv : b) m = std::max(m, kv.second.size());
        w128 += (double)m;
    }
    for (int ph = 0; ph < 2; ph++) {
        std::map<uint32_t, std::set<uint32_t> > b;
        for (int t = 4 * ph; t < 4 * ph + 4; t++)
            for (int l = 0; l < 4; l++) b[idx[t][l] % 16].insert(idx[t][l]);
        size_t m = 0;
        for (auto &kv : b) m = std::max(m, kv.second.size());
        w64 += (double)m;
    }
}

int main() {
    const uint32_t ntile = 2259, S = (ntile + 1 + 15u) & ~15u;
    int fails = 0, overflows = 0;
    double cur128 = 0, cur64 = 0, new128 = 0, new64 = 0;
    long steps = 0, misfit = 0, real = 0;
    std::normal_distribution<double> N(110.0, 9.0);
    for (int warp = 0; warp < 3000; warp++) {
        Row rows[8];
        for (uint32_t t = 0; t < 8; t++) {
            uint32_t my = (uint32_t)std::min(159.0, std::max(1.0, N(rng)));
            if (warp % 97 == 0) my = 1 + (warp + t) % 3;     // nearly empty rows
            if (warp % 89 == 0) my = 160;                   // full rows
            bool ovf;
            rows[t] = make_row(ntile, S, my, t & 3u, ovf);
            overflows += ovf;
            if (ovf) continue;
            // every entry exactly once, pads are class sentinels
            std::multiset<uint16_t> a, b;
            for (uint32_t k = 0; k < rows[t].my; k++) a.insert(rows[t].in[bo_nat(k)]);
            for (uint32_t g = 0; g < rows[t].G; g++)
                for (uint32_t tl = 0; tl < 4; tl++) {
                    const uint16_t e = rows[t].out[bo_slot(g, tl)];
                    if (e < ntile) {
                        b.insert(e);
                        real++;
                        misfit += (e % 16u) != 4u * (((t & 3u) + g) & 3u) + tl;
                    } else if (e != bo_sentinel(S, g, t & 3u, tl)) {
                        fails++;
                    }
                }
            if (a != b) fails++;
        }
        uint32_t gmax = 0;
        for (auto &r : rows) gmax = std::max(gmax, r.G);
        for (uint32_t g = 0; g < gmax; g++) {
            uint32_t ic[8][4], in_[8][4];
            for (uint32_t t = 0; t < 8; t++)
                for (uint32_t tl = 0; tl < 4; tl++) {
                    const bool act = g < rows[t].G;
                    ic[t][tl] = act ? rows[t].in[bo_slot(g, tl)] : 100000u + 16u * t + tl;   // idle team: no conflicts
                    in_[t][tl] = act ? rows[t].out[bo_slot(g, tl)] : 100000u + 16u * t + tl;
                }
            wavefronts(ic, cur128, cur64);
            wavefronts(in_, new128, new64);
            steps++;
        }
    }
    std::printf("rows checked: %d warps x 8, failures %d, phase-1 overflows %d\n", 3000, fails, overflows);
    std::printf("misfit entries %.2f %% of %ld\n", 100.0 * misfit / real, real);
    std::printf("build order : LDS.128 %.2f + LDS.64 %.2f = %.2f wavefronts per warp step\n", cur128 / steps, cur64 / steps, (cur128 + cur64) / steps);
    std::printf("bank order  : LDS.128 %.2f + LDS.64 %.2f = %.2f wavefronts per warp step\n", new128 / steps, new64 / steps, (new128 + new64) / steps);
    return fails ? 1 : 0;
}
