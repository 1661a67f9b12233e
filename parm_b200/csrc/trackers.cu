// Statistics trackers on the device, so that registering them with a Collection does not force a
// host synchronisation per step:
//   RsqTracker    constraints.hpp:318-368, constraints.cpp:404-565   per-atom <dx^2>, <dx^4>, <dr^4> at several lags
//   ISFTracker    constraints.hpp:370-414, constraints.cpp:567-710   per-atom self-intermediate scattering sums
//   EnergyTracker constraints.hpp:260-316, constraints.cpp:366-402   running sums of K, U, E and their squares
// The accumulators live in HBM, indexed by AtomVec index (the neighbour-list re-sort never touches them); an
// update is one streaming kernel per active lag (plus the centre-of-mass reduction when usecom is set), enqueued
// with the step's abort guard. Results are read on demand.
#include <math.h>
#include <string.h>

#include <algorithm>

#include "internal.cuh"

#define T_BLOCK 256
#define T_MAXBLOCKS 1024

struct parm_tracker {
    parm_ctx *ctx;
    int kind; // 0 Rsq, 1 ISF, 2 Energy
    bool usecom;
    uint64_t curt;
    std::vector<uint64_t> skips, counts;
    int nks;
    double *d_ks;
    double *d_com;                 // [4]: centre of mass (x, y, z) of the current update, 0 when !usecom
    double *d_part;                // reduction partials
    std::vector<double *> d_past;  // per lag: pastlocs [D][n]
    std::vector<double *> d_acc;   // per lag: Rsq: xyz2 [D][n], xyz4 [D][n], r4 [n]; ISF: [nks][n][D][2]
    // EnergyTracker
    std::vector<parm_inter *> inters;
    unsigned n_skip, n_skipped;
    uint64_t N;
    double U0;
    double *d_esums;               // Es, Us, Ks, Esq, Usq, Ksq
    double *d_etmp;                // curK, then one 13-double block per interaction
};

static inline unsigned tgrid(const parm_ctx *c, uint32_t n) {
    unsigned nb = (n + T_BLOCK - 1) / T_BLOCK;
    unsigned cap = std::min((unsigned)c->num_sms * 8, (unsigned)T_MAXBLOCKS);
    return nb < 1 ? 1 : std::min(nb, cap);
}
#define TLOOP for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x)
#define TGUARD if (abort_flag && *abort_flag) return

// what 0: sum x*m (3) and sum m over m > 0, finite (AtomGroup::com, box.cpp:228-247)
// what 1: sum v.v * m / 2 over every atom (EnergyTracker::update, constraints.cpp:374-378)
__global__ void __launch_bounds__(T_BLOCK)
k_t_reduce(int what, int D, const double4 *__restrict__ pos, const double *__restrict__ v, uint32_t n, uint32_t npad, double *partials,
           const int *__restrict__ abort_flag) {
    TGUARD;
    double q[4] = {0, 0, 0, 0};
    TLOOP {
        const double4 p = pos[s];
        if (what == 0) {
            if (p.w <= 0 || isinf(p.w)) continue;
            q[0] += p.x * p.w;
            q[1] += p.y * p.w;
            q[2] += p.z * p.w;
            q[3] += p.w;
        } else {
            const double vx = v[s], vy = v[npad + s], vz = D == 3 ? v[2 * (size_t)npad + s] : 0.0;
            const double vv = D == 3 ? __dadd_rn(__dmul_rn(vx, vx), __dadd_rn(__dmul_rn(vy, vy), __dmul_rn(vz, vz)))
                                     : __dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy));
            q[0] += __ddiv_rn(__dmul_rn(vv, p.w), 2.0);
        }
    }
    __shared__ double red[4][T_BLOCK / 32];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        double t = q[k];
#pragma unroll
        for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = t;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0;
        for (int w = 0; w < T_BLOCK / 32; w++) t += red[threadIdx.x][w];
        partials[4 * blockIdx.x + threadIdx.x] = t;
    }
}
// mode 0: out[0..2] = sum(x m) / sum(m)    mode 1: out[0] = sum
__global__ void k_t_fold(int mode, const double *partials, unsigned nblocks, double *out, const int *__restrict__ abort_flag) {
    TGUARD;
    __shared__ double red[4][T_BLOCK / 32];
    double tot[4];
    for (int k = 0; k < 4; k++) {
        double t = 0;
        for (unsigned b = threadIdx.x; b < nblocks; b += blockDim.x) t += partials[4 * b + k];
#pragma unroll
        for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < 4; k++) {
            tot[k] = 0;
            for (int w = 0; w < T_BLOCK / 32; w++) tot[k] += red[k][w];
        }
        if (mode == 0) {
            out[0] = tot[0] / tot[3];
            out[1] = tot[1] / tot[3];
            out[2] = tot[2] / tot[3];
        } else {
            out[0] = tot[0];
        }
    }
}

// pastlocs.row(i) = atoms[i].x - com   (ctors constraints.cpp:404-416, 567-582; reset :418-427, :584-593)
__global__ void __launch_bounds__(T_BLOCK)
k_t_init_past(int D, const double4 *__restrict__ pos, const uint32_t *__restrict__ order, uint32_t n, uint32_t nid,
              const double *__restrict__ com, double *past) {
    TLOOP {
        const double4 p = pos[s];
        const uint32_t id = order[s];
        past[id] = __dsub_rn(p.x, com[0]);
        past[nid + id] = __dsub_rn(p.y, com[1]);
        if (D == 3) past[2 * (size_t)nid + id] = __dsub_rn(p.z, com[2]);
    }
}

// RsqTracker1::update, constraints.cpp:429-455
template <int D>
__global__ void __launch_bounds__(T_BLOCK)
k_t_rsq(const double4 *__restrict__ pos, const uint32_t *__restrict__ order, uint32_t n, uint32_t nid, const double *__restrict__ com,
        double *past, double *acc, const int *__restrict__ abort_flag) {
    TGUARD;
    double *xyz2 = acc, *xyz4 = acc + (size_t)D * nid, *r4 = acc + 2 * (size_t)D * nid;
    TLOOP {
        const double4 p = pos[s];
        const uint32_t id = order[s];
        const double x[3] = {p.x, p.y, p.z};
        double dist4 = 0.0;
#pragma unroll
        for (int j = 0; j < D; j++) {
            const size_t q = (size_t)j * nid + id;
            const double r = __dsub_rn(x[j], com[j]);
            double d2 = __dsub_rn(r, past[q]);
            d2 = __dmul_rn(d2, d2);
            const double d4 = __dmul_rn(d2, d2);
            dist4 = __dadd_rn(dist4, d2);
            xyz2[q] = __dadd_rn(xyz2[q], d2);
            xyz4[q] = __dadd_rn(xyz4[q], d4);
            past[q] = r;
        }
        dist4 = __dmul_rn(dist4, dist4);
        r4[id] = __dadd_rn(r4[id], dist4);
    }
}

// ISFTracker1::update, constraints.cpp:595-619: ISFsums[ki][i][j] += exp(i ks[ki] dr[j])
template <int D>
__global__ void __launch_bounds__(T_BLOCK)
k_t_isf(const double4 *__restrict__ pos, const uint32_t *__restrict__ order, uint32_t n, uint32_t nid, const double *__restrict__ com,
        double *past, double *acc, const double *__restrict__ ks, int nks, const int *__restrict__ abort_flag) {
    TGUARD;
    TLOOP {
        const double4 p = pos[s];
        const uint32_t id = order[s];
        const double x[3] = {p.x, p.y, p.z};
        double dr[3];
#pragma unroll
        for (int j = 0; j < D; j++) {
            const size_t q = (size_t)j * nid + id;
            const double r = __dsub_rn(x[j], com[j]);
            dr[j] = __dsub_rn(r, past[q]);
            past[q] = r;
        }
        for (int ki = 0; ki < nks; ki++) {
            double *a = acc + (((size_t)ki * nid + id) * D) * 2;
#pragma unroll
            for (int j = 0; j < D; j++) {
                double sn, cs;
                sincos(__dmul_rn(ks[ki], dr[j]), &sn, &cs);
                a[2 * j] += cs;
                a[2 * j + 1] += sn;
            }
        }
    }
}

// EnergyTracker::update tail, constraints.cpp:385-393
__global__ void k_t_energy_accum(const double *tmp, int ninters, double U0, double *sums, const int *__restrict__ abort_flag) {
    TGUARD;
    if (threadIdx.x || blockIdx.x) return;
    const double curK = tmp[0];
    double curU = 0;
    for (int k = 0; k < ninters; k++) curU += tmp[4 + 13 * k];
    curU -= U0;
    sums[2] += curK;
    sums[1] += curU;
    sums[0] += curK + curU;
    sums[5] += curK * curK;
    sums[4] += curU * curU;
    sums[3] += (curK + curU) * (curK + curU);
}

// ---- host ------------------------------------------------------------------------------------------------------
int parm_inter_launch_obs(parm_inter *it, double *d_out, const int *abort_flag); // force.cu

static int com_to_device(parm_tracker *t, const int *abort_flag) {
    parm_ctx *c = t->ctx;
    const uint32_t n = parm_owned(c);
    if (!t->usecom || !n) return 0; // d_com stays zero
    const unsigned nb = tgrid(c, n);
    k_t_reduce<<<nb, T_BLOCK, 0, c->stream>>>(0, c->D, c->pos, c->v, n, c->npad, t->d_part, abort_flag);
    CK_LAUNCH(c);
    k_t_fold<<<1, T_BLOCK, 0, c->stream>>>(0, t->d_part, nb, t->d_com, abort_flag);
    CK_LAUNCH(c);
    return 0;
}

static size_t acc_doubles(const parm_tracker *t) {
    const parm_ctx *c = t->ctx;
    return t->kind == 0 ? (size_t)(2 * c->D + 1) * c->nid : (size_t)t->nks * c->nid * c->D * 2;
}

static int init_singles(parm_tracker *t) { // constructor / reset(): zero the sums, pastlocs = x - com
    parm_ctx *c = t->ctx;
    const uint32_t n = parm_owned(c);
    PTRY(com_to_device(t, nullptr));
    for (size_t k = 0; k < t->skips.size(); k++) {
        CK(cudaMemsetAsync(t->d_acc[k], 0, acc_doubles(t) * 8, c->stream));
        if (n) {
            k_t_init_past<<<tgrid(c, n), T_BLOCK, 0, c->stream>>>(c->D, c->pos, c->order, n, c->nid, t->d_com, t->d_past[k]);
            CK_LAUNCH(c);
        }
        t->counts[k] = 0;
    }
    return 0;
}

static int create_common(parm_ctx *c, int kind, const uint64_t *ns, int nns, int usecom, parm_tracker **out) {
    if (!c || !out || (nns > 0 && !ns)) { parm_set_error("tracker create: NULL argument"); return PARM_ERR_INVALID; }
    *out = 0;
    if (c->sh.on) { parm_set_error("statistics trackers run on single-GPU contexts only"); return PARM_ERR_UNSUPPORTED; }
    for (int k = 0; k < nns; k++)
        if (ns[k] == 0) { parm_set_error("tracker create: a lag of 0 steps (t %% skip) is undefined"); return PARM_ERR_INVALID; }
    CK(cudaSetDevice(c->device));
    parm_tracker *t = new parm_tracker();
    t->ctx = c;
    t->kind = kind;
    t->usecom = usecom != 0;
    t->skips.assign(ns, ns + nns);
    t->counts.assign(nns, 0);
    CK(cudaMalloc(&t->d_com, 4 * 8));
    CK(cudaMemset(t->d_com, 0, 4 * 8));
    CK(cudaMalloc(&t->d_part, 4 * T_MAXBLOCKS * 8));
    *out = t;
    return 0;
}

extern "C" int parm_rsq_create(parm_ctx *c, const uint64_t *ns, int nns, int usecom, parm_tracker **out) {
    PTRY(create_common(c, 0, ns, nns, usecom, out));
    parm_tracker *t = *out;
    for (int k = 0; k < nns; k++) {
        double *p = 0, *a = 0;
        CK(cudaMalloc(&p, std::max<size_t>((size_t)c->D * c->nid, 1) * 8));
        CK(cudaMalloc(&a, std::max<size_t>(acc_doubles(t), 1) * 8));
        t->d_past.push_back(p);
        t->d_acc.push_back(a);
    }
    return init_singles(t);
}

extern "C" int parm_isf_create(parm_ctx *c, const double *ks, int nks, const uint64_t *ns, int nns, int usecom, parm_tracker **out) {
    if (nks > 0 && !ks) { parm_set_error("parm_isf_create: NULL ks"); return PARM_ERR_INVALID; }
    PTRY(create_common(c, 1, ns, nns, usecom, out));
    parm_tracker *t = *out;
    t->nks = nks;
    CK(cudaMalloc(&t->d_ks, std::max(nks, 1) * 8));
    if (nks) CK(cudaMemcpy(t->d_ks, ks, (size_t)nks * 8, cudaMemcpyHostToDevice));
    for (int k = 0; k < nns; k++) {
        double *p = 0, *a = 0;
        CK(cudaMalloc(&p, std::max<size_t>((size_t)c->D * c->nid, 1) * 8));
        CK(cudaMalloc(&a, std::max<size_t>(acc_doubles(t), 1) * 8));
        t->d_past.push_back(p);
        t->d_acc.push_back(a);
    }
    return init_singles(t);
}

extern "C" int parm_energy_tracker_create(parm_ctx *c, parm_inter **inters, int ninters, unsigned n_skip, parm_tracker **out) {
    PTRY(create_common(c, 2, nullptr, 0, 0, out));
    parm_tracker *t = *out;
    for (int k = 0; k < ninters; k++) {
        if (!inters[k] || inters[k]->ctx != c) { parm_set_error("EnergyTracker: interaction belongs to another AtomVec"); return PARM_ERR_INVALID; }
        t->inters.push_back(inters[k]);
    }
    t->n_skip = std::max(n_skip, 1u);
    CK(cudaMalloc(&t->d_esums, 6 * 8));
    CK(cudaMemset(t->d_esums, 0, 6 * 8));
    CK(cudaMalloc(&t->d_etmp, (4 + 13 * (size_t)std::max(ninters, 1)) * 8));
    CK(cudaMemset(t->d_etmp, 0, (4 + 13 * (size_t)std::max(ninters, 1)) * 8));
    return 0;
}

extern "C" int parm_tracker_destroy(parm_tracker *t) {
    if (!t) return 0;
    parm_ctx *c = t->ctx;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (double *p : t->d_past) cudaFree(p);
    for (double *p : t->d_acc) cudaFree(p);
    if (t->d_ks) cudaFree(t->d_ks);
    if (t->d_com) cudaFree(t->d_com);
    if (t->d_part) cudaFree(t->d_part);
    if (t->d_esums) cudaFree(t->d_esums);
    if (t->d_etmp) cudaFree(t->d_etmp);
    delete t;
    return 0;
}

// StateTracker::update(box), enqueued on the context's stream (abort_flag: the speculative-step guard, may be NULL)
int parm_tracker_enqueue_update(parm_tracker *t, const int *abort_flag) {
    parm_ctx *c = t->ctx;
    const uint32_t n = parm_owned(c);
    if (t->kind == 2) { // EnergyTracker::update :366-393
        if (t->n_skipped + 1 < t->n_skip) {
            t->n_skipped += 1;
            return 0;
        }
        t->n_skipped = 0;
        const unsigned nb = tgrid(c, std::max(n, 1u));
        k_t_reduce<<<nb, T_BLOCK, 0, c->stream>>>(1, c->D, c->pos, c->v, n, c->npad, t->d_part, abort_flag);
        CK_LAUNCH(c);
        k_t_fold<<<1, T_BLOCK, 0, c->stream>>>(1, t->d_part, nb, t->d_etmp, abort_flag);
        CK_LAUNCH(c);
        for (size_t k = 0; k < t->inters.size(); k++) PTRY(parm_inter_launch_obs(t->inters[k], t->d_etmp + 4 + 13 * k, abort_flag));
        k_t_energy_accum<<<1, 32, 0, c->stream>>>(t->d_etmp, (int)t->inters.size(), t->U0, t->d_esums, abort_flag);
        CK_LAUNCH(c);
        t->N++;
        return 0;
    }
    t->curt++; // RsqTracker::update :492-499, ISFTracker::update :662-669
    bool any = false;
    for (size_t k = 0; k < t->skips.size(); k++) any = any || (t->curt % t->skips[k] == 0);
    if (!any || !n) {
        for (size_t k = 0; k < t->skips.size(); k++)
            if (t->curt % t->skips[k] == 0) t->counts[k]++;
        return 0;
    }
    PTRY(com_to_device(t, abort_flag));
    for (size_t k = 0; k < t->skips.size(); k++) {
        if (t->curt % t->skips[k] != 0) continue;
        const unsigned nb = tgrid(c, n);
        if (t->kind == 0) {
            if (c->D == 3) k_t_rsq<3><<<nb, T_BLOCK, 0, c->stream>>>(c->pos, c->order, n, c->nid, t->d_com, t->d_past[k], t->d_acc[k], abort_flag);
            else k_t_rsq<2><<<nb, T_BLOCK, 0, c->stream>>>(c->pos, c->order, n, c->nid, t->d_com, t->d_past[k], t->d_acc[k], abort_flag);
        } else {
            if (c->D == 3) k_t_isf<3><<<nb, T_BLOCK, 0, c->stream>>>(c->pos, c->order, n, c->nid, t->d_com, t->d_past[k], t->d_acc[k], t->d_ks, t->nks, abort_flag);
            else k_t_isf<2><<<nb, T_BLOCK, 0, c->stream>>>(c->pos, c->order, n, c->nid, t->d_com, t->d_past[k], t->d_acc[k], t->d_ks, t->nks, abort_flag);
        }
        CK_LAUNCH(c);
        t->counts[k]++;
    }
    return 0;
}
extern "C" int parm_tracker_update(parm_tracker *t) {
    if (!t) { parm_set_error("parm_tracker_update: NULL tracker"); return PARM_ERR_INVALID; }
    CK(cudaSetDevice(t->ctx->device));
    return parm_tracker_enqueue_update(t, nullptr);
}

extern "C" int parm_tracker_reset(parm_tracker *t) {
    if (!t) { parm_set_error("parm_tracker_reset: NULL tracker"); return PARM_ERR_INVALID; }
    parm_ctx *c = t->ctx;
    CK(cudaSetDevice(c->device));
    if (t->kind == 2) { // EnergyTracker::reset :284-293
        t->n_skipped = 0;
        t->N = 0;
        CK(cudaMemsetAsync(t->d_esums, 0, 6 * 8, c->stream));
        return 0;
    }
    t->curt = 0; // :501-508, :671-678
    return init_singles(t);
}

extern "C" int parm_tracker_counts(parm_tracker *t, uint64_t *counts, int cap) {
    if (!t || !counts) { parm_set_error("parm_tracker_counts: NULL argument"); return PARM_ERR_INVALID; }
    for (int k = 0; k < cap && k < (int)t->counts.size(); k++) counts[k] = t->counts[k];
    return 0;
}

// means over the updates of lag `single`: xyz2, xyz4 (n x NDIM row-major), r4 (n); any pointer may be NULL
extern "C" int parm_rsq_read(parm_tracker *t, int single, double *xyz2, double *xyz4, double *r4) {
    if (!t || t->kind != 0 || single < 0 || single >= (int)t->skips.size()) { parm_set_error("parm_rsq_read: bad tracker or lag index"); return PARM_ERR_INVALID; }
    parm_ctx *c = t->ctx;
    CK(cudaSetDevice(c->device));
    const uint32_t n = c->nid;
    const int D = c->D;
    std::vector<double> h((size_t)(2 * D + 1) * n);
    CK(cudaStreamSynchronize(c->stream));
    if (n) CK(cudaMemcpy(h.data(), t->d_acc[single], h.size() * 8, cudaMemcpyDeviceToHost));
    const double cnt = (double)t->counts[single];
    for (uint32_t i = 0; i < n; i++) {
        for (int j = 0; j < D; j++) {
            if (xyz2) xyz2[(size_t)i * D + j] = h[(size_t)j * n + i] / cnt;
            if (xyz4) xyz4[(size_t)i * D + j] = h[(size_t)(D + j) * n + i] / cnt;
        }
        if (r4) r4[i] = h[(size_t)2 * D * n + i] / cnt;
    }
    return 0;
}

// ISFxyz of lag `single`: out[nks][n][NDIM][2] (re, im), the sums divided by the count (constraints.cpp:637-650)
extern "C" int parm_isf_read(parm_tracker *t, int single, double *out) {
    if (!t || t->kind != 1 || !out || single < 0 || single >= (int)t->skips.size()) { parm_set_error("parm_isf_read: bad tracker or lag index"); return PARM_ERR_INVALID; }
    parm_ctx *c = t->ctx;
    CK(cudaSetDevice(c->device));
    const size_t tot = acc_doubles(t);
    CK(cudaStreamSynchronize(c->stream));
    if (tot) CK(cudaMemcpy(out, t->d_acc[single], tot * 8, cudaMemcpyDeviceToHost));
    const double cnt = (double)t->counts[single];
    for (size_t q = 0; q < tot; q++) out[q] /= cnt;
    return 0;
}

// out[8]: N, Es, Us, Ks, Esq, Usq, Ksq, U0 (the raw sums; E() = Es/N ... constraints.hpp:301-312)
extern "C" int parm_energy_tracker_read(parm_tracker *t, double *out) {
    if (!t || t->kind != 2 || !out) { parm_set_error("parm_energy_tracker_read: not an EnergyTracker"); return PARM_ERR_INVALID; }
    parm_ctx *c = t->ctx;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(out + 1, t->d_esums, 6 * 8, cudaMemcpyDeviceToHost));
    out[0] = (double)t->N;
    out[7] = t->U0;
    return 0;
}

// set_U0(flt) / set_U0(Box&), constraints.hpp:294-298, constraints.cpp:395-402
extern "C" int parm_energy_tracker_set_u0(parm_tracker *t, int from_box, double U0) {
    if (!t || t->kind != 2) { parm_set_error("parm_energy_tracker_set_u0: not an EnergyTracker"); return PARM_ERR_INVALID; }
    if (from_box) {
        double curU = 0;
        for (parm_inter *it : t->inters) {
            double e;
            PTRY(parm_inter_energy(it, &e));
            curU += e;
        }
        U0 = curU;
    }
    t->U0 = U0;
    return parm_tracker_reset(t);
}
