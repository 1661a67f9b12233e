// The pair-force kernel template of NListed<A,P> (interaction.hpp:2102-2291) and its launcher. The host
// API is in force.cu; the per-functor instantiations are spread over force_k*.cu so that they compile in
// parallel.
//
// One team of TEAM lanes per atom walks the atom's FULL neighbour row (no Newton's-third-law scatter, hence
// no atomics and a run-to-run deterministic sum); energy/virial/stress are reduced with warp shuffles, then
// per block, then by one folding block (deterministic), and halved because every pair is visited from both
// ends.
#pragma once
#include "pairs.cuh"

#define F_BLOCK 128
#define NPART 13 // E, virial, stress[9], contacts, overlaps

enum { MODE_F = 0, MODE_FALL = 1 };

struct ForceArgs {
    const double4 *pos;
    const uint32_t *nbr, *cnt;
    uint32_t kmax;
    const uint8_t *spec;     // SPEC 1: species id per slot
    const PairConst *table;  // SPEC 1: nspecies x nspecies
    int nspecies;
    PairConst P1;            // SPEC 0
    PairConst P4[4];         // SPEC 3: the 2 x 2 species table
    uint32_t mask;           // row entry & mask = neighbour slot (entries may carry a species id in their top bits)
    int packed;              // the species id in the top bits is this interaction's
    double *f;
    const double *vel;       // RepulsionDragPair only: v[3][npad]
    uint32_t n, npad;
    BoxDev box;
    int accumulate;          // add to f instead of overwriting
    int store;               // 0: observables only (energy()/pressure()/stress()), forces are not written
    double *partials;
    const double4 *par;      // SPEC 2: [lo | hi] halves, npad apart
    const double *eps_tab, *sig_tab;
    int ntypes, minmix;
    const int *abort_flag;
    uint32_t first;
};

// TEAM lanes share one atom: lane t of the team takes row entries t, t+TEAM, ... (the team reads TEAM
// consecutive indices = one or more full 32-byte sectors of the row), U entries per lane are in flight at
// once (index loads first, then the pos[j] gathers, then the arithmetic), and the team folds its partial
// force with xor-shuffles. TEAM*U divides 32 so rows (kmax % 32 == 0) are always readable up to the padded end.
// SPEC: 0 one species (constants in registers), 1 species table in shared memory, 2 per-atom parameters,
//       3 two species (this atom's table row in registers, selected per neighbour)
template <int KIND, int SPEC, int MODE, int TEAM, int U>
__global__ void __launch_bounds__(F_BLOCK) k_force(const ForceArgs A) {
    if (A.abort_flag && *A.abort_flag) return; // speculatively enqueued step whose predecessor asked for a rebuild
    extern __shared__ PairConst s_table[];
    const double4 *__restrict__ pos = A.pos;
    if (SPEC == 1) {
        for (int q = threadIdx.x; q < A.nspecies * A.nspecies; q += blockDim.x) s_table[q] = A.table[q];
        __syncthreads();
    }
    const uint32_t s = A.first + (blockIdx.x * blockDim.x + threadIdx.x) / TEAM; // slots [first, n)
    const uint32_t tl = threadIdx.x % TEAM;
    const bool want_obs = MODE != MODE_F;
    constexpr bool DRAG = KIND == PARM_PAIR_REPULSIONDRAG;
    constexpr bool HI = SPEC == 2 && parm_nparams(KIND) > 3;
    double fx = 0, fy = 0, fz = 0;
    double acc[NPART];
    if (want_obs)
#pragma unroll
        for (int q = 0; q < NPART; q++) acc[q] = 0.0;
    const bool valid = s < A.n;
    const uint32_t sc = valid ? s : 0;
    const uint32_t my = valid ? A.cnt[sc] : 0;
    const double4 pi = pos[sc];
    const uint32_t *row = A.nbr + (size_t)sc * A.kmax;
    const PairConst *prow = SPEC == 1 ? s_table + (int)A.spec[sc] * A.nspecies : nullptr;
    PairConst R0, R1; // SPEC 3: pair constants of this atom against species 0 and 1
    if (SPEC == 3) {
        const int si = (int)A.spec[sc];
        R0 = A.P4[2 * si];
        R1 = A.P4[2 * si + 1];
    }
    double qi[5] = {0, 0, 0, 0, 0};
    int ti = 0;
    if (SPEC == 2) {
        const double4 lo = A.par[sc];
        qi[0] = lo.x; qi[1] = lo.y; qi[2] = lo.z; ti = (int)lo.w;
        if (HI) {
            const double4 hi = A.par[A.npad + sc];
            qi[3] = hi.x; qi[4] = hi.y;
        }
    }
    double vix = 0, viy = 0, viz = 0;
    if (DRAG) {
        vix = A.vel[sc];
        viy = A.vel[A.npad + sc];
        viz = A.vel[2 * (size_t)A.npad + sc];
    }
    // The row words of a trip are requested one trip ahead -- those of the first trip together with the row length and
    // the atom's own position, before either has arrived (rows are allocated to kmax, a multiple of 32 = of TEAM * U, so
    // every word of a trip is readable; what lies beyond the row's length is discarded below). Short rows (contact
    // potentials: one trip) thus cost two dependent memory round trips instead of three.
    uint32_t ew[U];
#pragma unroll
    for (int u = 0; u < U; u++) ew[u] = __ldg(row + tl + u * TEAM);
    for (uint32_t k0 = tl; k0 < my; k0 += TEAM * U) {
        uint32_t j[U], sp[U];
        bool ok[U];
        double4 pj[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t k = k0 + u * TEAM;
            ok[u] = k < my;
            const uint32_t e = ok[u] ? ew[u] : sc;
            j[u] = e & A.mask;
            sp[u] = e >> PARM_NBR_SLOT_BITS;
        }
        if (k0 + TEAM * U < my) {
#pragma unroll
            for (int u = 0; u < U; u++) ew[u] = __ldg(row + k0 + TEAM * U + u * TEAM);
        }
#pragma unroll
        for (int u = 0; u < U; u++) pj[u] = ld_pos4(pos + j[u]);
#pragma unroll
        for (int u = 0; u < U; u++) {
            // OriginBox::diff(atom1->x, atom2->x), box.hpp:103
            double dx = min_image_fast(pi.x - pj[u].x, A.box.L[0], A.box.invL[0]);
            double dy = min_image_fast(pi.y - pj[u].y, A.box.L[1], A.box.invL[1]);
            double dz = min_image_fast(pi.z - pj[u].z, A.box.L[2], A.box.invL[2]);
            double dsq = dx * dx + (dy * dy + dz * dz);
            double vdotr = 0.0;
            if (DRAG) {
                const double wx = vix - __ldg(A.vel + j[u]);
                const double wy = viy - __ldg(A.vel + A.npad + j[u]);
                const double wz = viz - __ldg(A.vel + 2 * (size_t)A.npad + j[u]);
                vdotr = wx * dx + (wy * dy + wz * dz);
            }
            double scal, e;
            if (SPEC == 0) {
                pair_eval<KIND>(A.P1, dsq, vdotr, want_obs, scal, e);
            } else if (SPEC == 1) {
                const PairConst &P = prow[A.packed ? sp[u] : (uint32_t)__ldg(A.spec + j[u])];
                pair_eval<KIND>(P, dsq, vdotr, want_obs, scal, e);
            } else if (SPEC == 3) {
                const uint32_t sj = A.packed ? sp[u] : (uint32_t)__ldg(A.spec + j[u]);
                PairConst P;
                P.eps = sj ? R1.eps : R0.eps; P.sig = sj ? R1.sig : R0.sig; P.sig2 = sj ? R1.sig2 : R0.sig2;
                P.rc2 = sj ? R1.rc2 : R0.rc2; P.cutE = sj ? R1.cutE : R0.cutE;
                P.a = sj ? R1.a : R0.a; P.b = sj ? R1.b : R0.b; P.c = sj ? R1.c : R0.c;
                pair_eval<KIND>(P, dsq, vdotr, want_obs, scal, e);
            } else {
                double qj[5] = {0, 0, 0, 0, 0};
                const double4 lo = ld_pos4(A.par + j[u]);
                qj[0] = lo.x; qj[1] = lo.y; qj[2] = lo.z;
                if (HI) {
                    const double4 hi = ld_pos4(A.par + A.npad + j[u]);
                    qj[3] = hi.x; qj[4] = hi.y;
                }
                const PairConst P = mix_pair<KIND>(qi, ti, qj, (int)lo.w, A.eps_tab, A.sig_tab, A.ntypes, A.minmix != 0, want_obs);
                pair_eval<KIND>(P, dsq, vdotr, want_obs, scal, e);
            }
            if (!ok[u]) { // padding lane (j == self, dsq == 0): contributes nothing
                scal = 0.0;
                e = 0.0;
            }
            double gx = dx * scal, gy = dy * scal, gz = dz * scal;
            fx += gx;
            fy += gy;
            fz += gz;
            if (want_obs) {
                acc[0] += e;
                acc[1] += dx * gx + (dy * gy + dz * gz); // r.dot(f), :2241
                acc[2] += dx * gx; acc[3] += dx * gy; acc[4] += dx * gz; // stress += r * f^T, :2274
                acc[5] += dy * gx; acc[6] += dy * gy; acc[7] += dy * gz;
                acc[8] += dz * gx; acc[9] += dz * gy; acc[10] += dz * gz;
                acc[11] += (e != 0.0) ? 1.0 : 0.0; // contacts :2126-2137
                acc[12] += (e > 0.0) ? 1.0 : 0.0;  // overlaps :2140-2151
            }
        }
    }
    if (MODE == MODE_F || A.store) {
#pragma unroll
        for (int o = TEAM / 2; o; o >>= 1) {
            fx += __shfl_xor_sync(0xffffffffu, fx, o);
            fy += __shfl_xor_sync(0xffffffffu, fy, o);
            fz += __shfl_xor_sync(0xffffffffu, fz, o);
        }
        if (valid && tl == 0) {
            double *f = A.f;
            if (A.accumulate) {
                f[s] += fx;
                f[A.npad + s] += fy;
                f[2 * (size_t)A.npad + s] += fz;
            } else {
                f[s] = fx;
                f[A.npad + s] = fy;
                f[2 * (size_t)A.npad + s] = fz;
            }
        }
    }
    if (want_obs) {
        __shared__ double red[NPART][F_BLOCK / 32];
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
        for (int q = 0; q < NPART; q++) {
            double x = acc[q];
#pragma unroll
            for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == 0) red[q][w] = x;
        }
        __syncthreads();
        if (threadIdx.x < NPART) {
            double x = 0;
            for (int ww = 0; ww < F_BLOCK / 32; ww++) x += red[threadIdx.x][ww];
            A.partials[(size_t)blockIdx.x * NPART + threadIdx.x] = x;
        }
    }
}

template <int KIND, int SPEC, int TEAM>
static cudaError_t launch_mode(int mode, dim3 grid, size_t smem, cudaStream_t st, const ForceArgs &A) {
    if (mode == MODE_F) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(k_force<KIND, SPEC, MODE_F, TEAM, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_force<KIND, SPEC, MODE_F, TEAM, 2><<<grid, F_BLOCK, smem, st>>>(A);
    } else {
        if (smem > 48 * 1024) cudaFuncSetAttribute(k_force<KIND, SPEC, MODE_FALL, TEAM, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_force<KIND, SPEC, MODE_FALL, TEAM, 2><<<grid, F_BLOCK, smem, st>>>(A);
    }
    return cudaGetLastError();
}

// specmode 0/1/2 as SPEC above; team 4 or 8 (16: one species, small systems); natoms = number of slots in [first, n)
template <int KIND>
cudaError_t parm_launch_force_kind(int specmode, int team, int mode, uint32_t natoms, size_t smem, cudaStream_t st, const ForceArgs &A) {
    const dim3 grid((unsigned)(((size_t)natoms * team + F_BLOCK - 1) / F_BLOCK));
    if (team == 16 && specmode == 0) return launch_mode<KIND, 0, 16>(mode, grid, 0, st, A);
    if (team == 8) {
        if (specmode == 0) return launch_mode<KIND, 0, 8>(mode, grid, 0, st, A);
        if (specmode == 1) return launch_mode<KIND, 1, 8>(mode, grid, smem, st, A);
        if (specmode == 3) return launch_mode<KIND, 3, 8>(mode, grid, 0, st, A);
        return launch_mode<KIND, 2, 8>(mode, grid, 0, st, A);
    }
    if (specmode == 0) return launch_mode<KIND, 0, 4>(mode, grid, 0, st, A);
    if (specmode == 1) return launch_mode<KIND, 1, 4>(mode, grid, smem, st, A);
    if (specmode == 3) return launch_mode<KIND, 3, 4>(mode, grid, 0, st, A);
    return launch_mode<KIND, 2, 4>(mode, grid, 0, st, A);
}

#define PARM_INSTANTIATE_FORCE_KIND(K) \
    template cudaError_t parm_launch_force_kind<K>(int, int, int, uint32_t, size_t, cudaStream_t, const ForceArgs &);
