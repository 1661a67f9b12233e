// parm_b200 internal declarations shared by the .cu translation units.
// Device data layout (DESIGN.md "Data layout in HBM"):
//   pos   double4[npad]      (x, y, z, m) per SLOT; z == 0 in 2-D builds. Canonical
//                            storage of Atom::x and Atom::m (box.hpp:234-249).
//   v,a,f double[3][npad]    SoA per slot, component-major.
//   order uint32[npad]       order[slot] = AtomVec index; slot_of is the inverse.
// Slots are re-ordered by cell index at every neighbour-list rebuild so that the
// force kernel's gathers of pos[j] stay cache-local.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <set>
#include <string>
#include <utility>
#include <vector>

#include "../../include/parm_b200.h"

#define PARM_MAX_SPECIES 32   // distinct per-atom parameter tuples per interaction
#define PARM_NBR_SLOT_BITS 27  // neighbour-row entry = slot | species << 27 when packed (parm_nlist::packed_for)
#define PARM_NBR_SLOT_MASK ((1u << PARM_NBR_SLOT_BITS) - 1u)
#define PARM_NBR_MASK_OF(nl) (((nl)->packed || (nl)->tagged) ? PARM_NBR_SLOT_MASK : 0xffffffffu)
#define PARM_PACK_MIN_NEIGHBORS 24 // mean row length from which packing / the two-species register path pay

void parm_set_error(const char *fmt, ...);
void parm_count_launch(parm_ctx *ctx, unsigned n = 1);

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            parm_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, __LINE__, \
                           cudaGetErrorString(e_));                                           \
            return PARM_ERR_CUDA;                                                             \
        }                                                                                     \
    } while (0)
#define PTRY(expr)             \
    do {                       \
        int r_ = (expr);       \
        if (r_) return r_;     \
    } while (0)
#define CK_LAUNCH(ctx)        \
    do {                      \
        parm_count_launch(ctx); \
        CK(cudaGetLastError()); \
    } while (0)

struct BoxDev {
    double L[3], invL[3], halfL[3];
};

// Per species-pair constants: the state of one pair functor P after its constructor ran (csrc/pairs.cuh).
struct PairConst {
    double eps;  // epsilon_ij (EisMclachlan: c0)
    double sig;  // sigma_ij (LJAttractRepulseSigs: sig_r; EisMclachlan: cutoff)
    double sig2; // sig*sig
    double rc2;  // square of the distance beyond which P::forces and P::energy are exactly zero
    double cutE; // cut_energy
    double a, b, c; // functor-specific: exponent, repeps, gamma, C, l, f, c1, c2, eps_a, sig_a, sigcut ...
};

// Slab decomposition state (one process per GPU; csrc/shard.cu). The slab axis is x (axis 0, the
// slowest index of the cell id), rank r owns wrapped x in [lo, lo + Ls). Slot layout after a
// rebuild: [owned atoms, lowest layer first | ghosts from the upper neighbour | ghosts from the lower
// neighbour] -- every range that is ever communicated is contiguous.
struct ShardState {
    bool on;
    int rank, nranks, up, down;
    double lo, Ls;
    void *comm;                       // ncclComm_t
    uint32_t n_local, g_dn, g_up;     // local atoms, ghosts received from down / up
    uint32_t s_dn, s_up;              // boundary-layer atoms sent to down / up every step
    cudaStream_t comm_stream;         // per-step exchanges run here, overlapped with the interior forces
    cudaEvent_t ev_k1, ev_comm;
    cudaEvent_t ev_rows;              // behind an on-demand expansion of the 32-bit rows: both streams wait for it (nlist.cu)
    cudaStream_t main_stream;         // the context's own stream (parm_ctx::stream is swapped for the boundary launches of a step)
    uint32_t *d_counts, *h_counts;    // small exchange buffers (device, pinned host)
    double *d_gather, *h_gather;      // drift top-2 of every rank
    // one-sort rebuild (shard.cu: parm_shard_rebuild): the atoms that left the slab travel in fixed-capacity messages
    int mig_fast;                     // PARM_B200_SHARD_FAST (default 1)
    uint32_t mig_cap;                 // atoms per message (PARM_B200_SHARD_MIGCAP, default 4096); more leavers: two-sort path
    uint32_t *mig_list;               // [2][mig_cap] slots of the atoms leaving downwards / upwards
    uint32_t *mig_cnt;                // [4] leavers down, leavers up, overflow flag (max over ranks), spare
    char *mig_send[2], *mig_recv[2];  // [to down, to up] / [from up, from down]
    uint64_t fast_rebuilds, slow_rebuilds;
};

struct parm_ctx {
    int D;
    uint32_t n, npad;      // slots in use (local + ghost atoms), allocated slots
    uint32_t nid, nid_pad; // atom ids (AtomVec indices; global ids when sharded), allocated id entries
    ShardState sh;
    uint8_t *ghost, *ghost_alt; // per slot: 1 = ghost copy of a remote atom (sharded contexts only)
    int device;
    cudaStream_t stream;
    BoxDev box;
    bool box_set;
    double4 *pos, *pos_alt;
    double *v, *a, *f, *v_alt, *a_alt, *f_alt; // [3][npad]
    uint32_t *order, *order_alt, *slot_of;
    // transfer staging
    void *d_stage;
    size_t d_stage_bytes;
    // small results
    double *d_red;      // device scratch for reductions (partials + finals)
    size_t d_red_doubles;
    double *h_red;      // pinned host mirror
    uint64_t launches;
    std::vector<parm_nlist *> nlists;
    std::vector<parm_inter *> inters;
    int num_sms;
    // asynchronous trajectory frames (csrc/snapshot.cu)
    cudaStream_t snap_stream;
    cudaEvent_t snap_ready, snap_done;
    double *d_snap, *h_snap;   // device staging / pinned host copy of one frame (x and v by AtomVec index)
    size_t snap_doubles;
    unsigned snap_mask;
    bool snap_pending, snap_init;
    bool tile_prep_external;   // the caller refreshes prel itself (sharded step: owned slots after K1, ghosts after the exchange)
    // optional per-class CUDA-event timing
    bool prof_on;
    struct ProfRec { int cls; cudaEvent_t a, b; };
    std::vector<ProfRec> prof_pending;
    std::vector<cudaEvent_t> prof_pool;
    double prof_ms[PARM_PROF_N];
    uint64_t prof_cnt[PARM_PROF_N];
};
int parm_prof_begin(parm_ctx *c, int cls);
int parm_prof_end(parm_ctx *c);

struct GridDev {
    int nc[3];
    double scale[3]; // nc / L
};
// Cells have edge >= r_list/sub, the stencil is (2 sub + 1)^D cells.
struct StencilDev {
    int sub;     // cells per r_list
    int full[3]; // 1: nc >= 2 sub + 1 (offsets -sub..sub, unique images); 0: visit every cell of that axis once
    int open0;   // axis 0 is the slab axis of a sharded context: no wrap, halo layers at both ends
};
struct ShardDev {
    int on;
    double lo, Ls, L;
    int nci; // interior cell layers along the slab axis
};

struct NlistFlags { // device-written; a pinned host mirror receives need_rebuild / top2
    int need_rebuild;
    uint32_t maxcnt;              // longest row of the last build
    uint32_t nbmax;               // mask-mode build: most candidate blocks any warp group walked
    uint32_t trigger;             // (pinned host copy only) batched stepping: number of the step of the batch whose drift rule fired
    unsigned long long total;     // sum of row lengths of the last build
    unsigned long long xmax_bits; // bit pattern of max |coordinate| seen by the last binning pass
    double top2[2];
};

// ---- cell-tile staging for the pair kernel (csrc/tile.cu, csrc/force_tile.cuh) --------------------
// The owned slots are cut into CHUNKS of at most `ch` consecutive slots that never straddle an (x,y)
// cell column. The neighbours of a chunk's atoms all lie in <= TILE_MAXSEG contiguous slot runs (the
// 3 x 3 surrounding columns, z range of the chunk +- one cell): the chunk's TILE. The pair kernel
// stages the tile's positions in shared memory, relative to the chunk origin and minimum-imaged
// once per staged atom, and walks 16-bit tile-local neighbour rows (rows16) written once per rebuild.
#define TILE_MAXSEG 18
#define TILE_NT 256            // threads per block of the tile kernels
#define TILE_MAX_ATOMS 8150    // staged positions per tile: 24 B each in shared memory, and 8 * (index + sentinels) must fit 16 bits
#define TILE_IDX_SHIFT 3       // rows16 entries are 8 * tile index = the byte offset of z in the staged tile (16 * index for (x, y))
struct TileChunk {
    uint32_t s0, n;            // first slot, number of atoms (<= ch)
    uint32_t nseg, ntile;      // contiguous slot runs, staged atoms in total
    double o[3];               // origin the staged coordinates are relative to
    uint32_t flags, pad;       // bit 0: tile is wider than half the box on some axis: minimum image per pair as well
                               // bit 1: some run carries a periodic image shift (sh[][] below; bulk-copy staging only)
    uint32_t seg_start[TILE_MAXSEG];
    uint32_t seg_off[TILE_MAXSEG + 1]; // exclusive prefix of the run lengths
    int8_t sh[TILE_MAXSEG][3]; // image shift of every run in box lengths (-1, 0, +1), applied after the bulk copies land
    uint8_t pad2[2];
};
struct TileInfo { // device-written, copied to the host with the build flags
    uint32_t nchunks, max_tile, bad, wide;
};
struct TileState {
    int enabled;               // PARM_B200_TILE (default 1)
    int min_nbrs;              // mean row length from which the tile kernel is used (PARM_B200_TILE_MIN_NEIGHBORS, 32)
    int ch, team, v;           // chunk size; lanes per atom and row entries per lane and pass of the pair kernel
    bool planned;              // chunk table matches the current cell structure
    bool valid;                // rows16 match the current rows
    bool banked;               // rows16 were re-ordered by the experimental bank-aware pass (no half passes in the pair kernel)
    uint32_t nchunks, max_tile;
    uint32_t ncol;             // owned cell columns
    TileChunk *d_chunks;
    uint32_t *d_s0;            // [chunk_cap + 1] first slot of every chunk, then the end of the last one (the persistent pair
                               // kernel reads row lengths and row words one chunk ahead of the chunk table)
    uint32_t chunk_cap;
    uint16_t *rows16;          // [npad][kmax], permuted per (team, v) block, padded with the sentinel index
    size_t rows16_cap;
    uint32_t *d_col;           // [2][ncol + 1]: first slot / first chunk of every owned column
    uint32_t col_cap;
    uint32_t *h_col;           // pinned copy of d_col (same layout, stride col_cap)
    TileInfo *d_info, *h_info; // h_info pinned
    // bulk-copy staging (stage == 1): the pair kernel copies the runs of its tile from prel with cp.async.bulk (TMA)
    // instead of loading, re-imaging and storing them with its own threads
    int stage;                 // PARM_B200_TILE_STAGE (default 1); 0 = stage from pos inside the pair kernel
    bool stage_aligned;        // the current chunk table has even-aligned runs and shift codes (what stage == 1 needs)
    double2 *prel_xy;          // [npad] (x, y) - img * L: the image every atom had at the last rebuild, refreshed every step
    double *prel_z;            // [npad]
    float4 *img;               // [npad] image numbers (exact small integers), written with pw at every rebuild
};

// Mask-mode build (EXPERIMENTAL, PARM_B200_BUILD_MASKS=1, off by default): instead of expanded 32-bit rows the build
// kernel leaves one 32-bit pass mask per (atom, candidate block) and the first candidate slot (+ stencil column tag) of
// every block per warp group; tile.cu expands the masks straight into the 16-bit tile-local rows and the 32-bit rows
// are materialised only on demand (parm_nlist_ensure_rows32: pair download, gather kernel).
struct MaskOut { // kernel argument of the build
    uint32_t *masks;    // [mb_cap][npad]: masks[b * npad + slot]
    uint32_t *blk_base; // [ngroups][mb_cap]: first candidate slot | column tag << PARM_NBR_SLOT_BITS
    uint32_t *grp_nb;   // [ngroups]: candidate blocks of the group
    uint32_t mb_cap, npad;
    // direct mode: the build warp also keeps the masks of its 32 atoms in shared memory and expands them itself, lane =
    // atom, straight into the 16-bit tile-local rows (no separate localize pass); needs the tile plan of this rebuild
    uint32_t direct;                 // 1: write rows16 from the build kernel
    const TileChunk *chunks;
    const uint32_t *col_slot, *col_chunk; // first slot / first chunk of every owned column (tile.cu: k_tile_cols)
    uint32_t ch;                     // chunk size of the plan
    uint16_t *rows16;
    uint32_t *direct_fail;           // set when a warp group walked more candidate blocks than its shared memory holds
};
struct MaskState {
    int enabled;        // PARM_B200_BUILD_MASKS
    bool active;        // the current list was built in mask mode
    bool rows32_valid;  // nbr[] has been expanded from the masks since the last build
    bool direct;        // the build kernel wrote rows16 itself (no localize pass)
    uint32_t *d_fail, *h_fail; // direct mode: overflow flag of the in-kernel expansion (device / pinned)
    MaskOut out;
    size_t masks_cap, blk_cap, grp_cap;
    uint32_t ngroups, gpc;
    int zg;
};

struct parm_nlist {
    parm_ctx *ctx;
    TileState tile;
    MaskState mask;
    double skin;
    std::vector<double> h_diam; // by AtomVec index; < 0: not a member
    bool have_diam;
    double maxdiam, mindiam;
    bool uniform;        // every atom a member, one common diameter
    double *d_diam_id;   // by AtomVec index
    double *d_diam;      // by slot
    double *xlast;       // [3][npad] by slot: lastlocs (trackers.hpp:165)
    double4 *pw;         // wrapped copy (x,y,z in [0,L)) + conservative half threshold, rebuilt with the list
    int cell_sub;        // cells per r_list (1 or 2)
    uint32_t updatenum;
    bool ignorechanged;
    // Species packing: when one interaction with 2..32 tabulated species uses this list, the top 5 bits of every
    // row entry carry the neighbour's species id, so the force kernel needs no per-neighbour species gather.
    bool packed;                 // row entries carry a species id in their top bits (mask them)
    bool tagged;                 // row entries carry the stencil column (0..8) they were found in in their top bits:
                                 // written by the build kernel for tile.cu's localize pass, dropped by species packing
    parm_inter *packed_for;      // whose species they are (NULL once that interaction is destroyed)
    // NeighborList::ignore (trackers.hpp:190-193): excluded pairs, canonical (larger index, smaller index)
    std::set<std::pair<uint32_t, uint32_t> > ignored;
    bool ignore_dirty;           // the device CSR is older than `ignored`
    uint32_t *d_excl_start;      // by AtomVec index, nid + 1 entries
    uint32_t *d_excl;            // both directions: the excluded partners of each atom
    // cell grid
    int nc[3];
    GridDev g;
    StencilDev st;
    ShardDev sd;
    bool smallbox;
    double lmax, thr_min;
    uint32_t ncell;
    uint32_t *cell_id, *cell_id_sorted, *perm, *iota, *cell_start;
    uint32_t *cell_fill, *scan_sums, *sort_tmp; // counting-sort scratch
    uint32_t cell_start_cap;
    // list: row-major nbr[slot * kmax + k], cnt[slot]; kmax is a multiple of 32
    uint32_t kmax;
    uint32_t *nbr;
    size_t nbr_cap_entries;
    uint32_t *cnt;
    unsigned long long total_full; // sum of cnt
    uint32_t maxcnt;
    // drift / build flags
    double *d_top2;      // per-block top-2
    unsigned int *d_counter;
    NlistFlags *d_flags;
    NlistFlags *h_flags; // pinned
    int *d_slot, *h_slot; // [3 (+1)] per-step rebuild decisions (device / pinned), used round robin: see parm_integ_timestep
    uint64_t rebuilds;
};

struct parm_inter {
    parm_ctx *ctx;
    parm_nlist *nl;
    int kind;
    bool have_params;
    int nspecies;
    std::vector<uint8_t> h_spec_id; // by AtomVec index
    uint8_t *d_spec_id;             // by AtomVec index
    uint8_t *d_spec;                // by slot
    PairConst *d_table;             // nspecies x nspecies
    std::vector<PairConst> h_table;
    // more than PARM_MAX_SPECIES distinct tuples (continuous polydispersity): per-atom parameters are
    // gathered with the neighbour and the pair constructor runs per pair on the device
    bool generic;
    bool spec_stale;                // species ids changed since the list entries were packed: gather them instead
    bool minmix;                    // LoisOhernPairMinCLs / LoisLinPairMin constructors
    std::vector<double> h_par_id;   // 8 doubles per AtomVec index: p0 p1 p2 type | p3 p4 - - (geometric ones as sqrt)
    double4 *d_par_id, *d_par;      // by AtomVec index / by slot; two double4 per atom: [lo | hi] halves
    double *d_eps_table;            // ntypes x ntypes or NULL
    double *d_sig_table;            // ntypes x ntypes or NULL
    int ntypes;
    double *d_partials;
    size_t partial_doubles;
};

struct parm_integ {
    parm_ctx *ctx;
    int type; // PARM_INTEG_*
    // integ_extra.cu: gamma (Overdamped), Q (NoseHoover), corrector passes and derivative arrays (Gear4A-6A, by
    // AtomVec index: [3 arrays][3 components][nid_pad]), thermostat scalars on the device
    double gamma, Q, ndof_cached;
    int ncorrec;
    double *d_gear;
    struct IntegScalars *d_scal;
    double *d_xpart; // per-block partials of the thermostat reductions
    struct NlcgState *nlcg; // CollectionNLCG (nlcg.cu)
    struct SmallState *small; // scratch of the persistent small-system kernel (small.cu)
    double dt, damping, force_mag, desT;
    double c0, c1, c2, sigmar, sigmav, corr, x11, x21, x22;
    uint64_t seed;
    std::vector<parm_inter *> inters;
    std::vector<parm_nlist *> trackers;
    std::vector<parm_tracker *> stat_trackers; // RsqTracker / ISFTracker / EnergyTracker (trackers.cu)
    double *d_noise;
    size_t noise_len, noise_pos;
    uint32_t *d_mobile_rank; // by AtomVec index (noise injection addressing)
    uint32_t n_mobile;
    uint64_t steps, rebuilds;
    uint64_t noise_step0;    // g->steps when the noise was injected
    cudaEvent_t ev[2];
    bool ev_ok;
};

// slots [0, parm_owned(c)) hold the atoms this context integrates; ghost copies (sharded) follow them
static inline uint32_t parm_owned(const parm_ctx *c) { return c->sh.on ? c->sh.n_local : c->n; }

// thermostat state of CollectionNoseHoover / CollectionGaussianT, kept on the device so that steps queue up
struct IntegScalars {
    double xi, lns, Kt, ytov;
};

// CollectionNLCG members (collection.hpp:409-427)
struct NlcgState {
    double seceps;
    unsigned secmax;
    double kappa, alphamax, afrac, dxmax, stepmax, kmax, P0;
    double Knew, k, vl, fl, al;
    double alpha, beta, betaused, dxsum, alphavmax, maxdV;
    unsigned sec;
};
int parm_nlcg_timestep(parm_integ *g); // nlcg.cu
// small.cu: whole timestep(n) calls of small CollectionVerlet systems as one persistent kernel; *nsteps = steps left over
int parm_small_run(parm_integ *g, parm_nlist *nl, int *nsteps);
void parm_small_free(parm_integ *g);
void parm_nlcg_free(parm_integ *g);

int parm_tracker_enqueue_update(parm_tracker *t, const int *abort_flag); // trackers.cu

// ---- cross-TU host functions ----
int parm_integ_extra_enqueue(parm_integ *g, uint64_t step, const int *abort_flag, int slot); // integ_extra.cu
int parm_integ_extra_after_set_forces(parm_integ *g);  // CollectionGaussianT::set_forces -> set_xi
int parm_integ_launch_all_forces(parm_integ *g, const int *abort_flag);
int parm_ctx_alloc(int ndim, uint32_t nid, uint32_t cap_slots, int device, parm_ctx **out);
int parm_shard_halo_exchange(parm_ctx *c);                 // per step, after K1
int parm_shard_drift_decision(parm_nlist *nl, bool *rebuild); // after K3: global top-2 rule (synchronous)
int parm_shard_drift_enqueue(parm_nlist *nl, int *d_slot, int *h_slot); // same, decision left in the step slot
// per step: halo exchange + drift all-gather on the communication stream (after what is queued on the main
// stream so far), and the point where the main stream waits for them
int parm_shard_step_comm(parm_ctx *c, parm_nlist *nl, int *d_slot, int *h_slot);
int parm_shard_step_join(parm_ctx *c);
int parm_shard_rebuild(parm_nlist *nl);                    // migration + ghost selection + build
int parm_shard_allreduce_sum(parm_ctx *c, double *d_buf, int count);
int parm_shard_destroy(parm_ctx *c);
// rebuild steps shared by the single-GPU and the sharded path (csrc/nlist.cu)
int parm_nlist_prepare_grid(parm_nlist *nl);
int parm_nlist_sort_permute(parm_nlist *nl, const uint32_t *d_src, uint32_t nsrc); // d_src NULL: slots 0..nsrc-1
int parm_nlist_build_rows(parm_nlist *nl);
int parm_nlist_append_ghosts(parm_nlist *nl, uint32_t first, uint32_t count);
int parm_nlist_rebuild(parm_nlist *nl);
int parm_nlist_ensure_rows32(parm_nlist *nl);  // mask-mode lists: expand nbr[] from the pass masks if that has not happened yet
int parm_nlist_drift_check_async(parm_nlist *nl);   // standalone drift kernel (update_list(false) outside timestep)
int parm_inter_regather(parm_inter *inter);          // re-gather per-slot species after a re-sort
int parm_inter_launch_forces(parm_inter *inter, unsigned want, bool accumulate, double *d_out /*device, 13 doubles*/,
                             const int *abort_flag = nullptr, uint32_t first = 0, uint32_t count = 0xffffffffu);
int parm_ctx_ensure_red(parm_ctx *ctx, size_t doubles);
void parm_snapshot_free(parm_ctx *c); // csrc/snapshot.cu
// cell-tile path (csrc/tile.cu)
int parm_tile_plan_enqueue(parm_nlist *nl);   // chunk table from the cell structure (async, before the build's sync)
int parm_tile_plan_fetch(parm_nlist *nl);     // queue the device->host copy of the plan summary (before that sync)
int parm_tile_localize(parm_nlist *nl);       // after the rows are final (ignore applied, before species packing): write rows16
void parm_tile_invalidate(parm_nlist *nl);
void parm_tile_free(parm_nlist *nl);
int parm_tile_localize_masks(parm_nlist *nl); // mask-mode lists: rows16 straight from the pass masks
int parm_tile_rows16_reserve(parm_nlist *nl);  // rows16 allocated for the current npad x kmax (before a direct-mode build)
// prel of slots [first, first + count) from pos (bulk-copy staging); no-op when the list does not stage that way
int parm_tile_prep(parm_nlist *nl, uint32_t first, uint32_t count, cudaStream_t stream, const int *abort_flag);
bool parm_tile_all_fit(const parm_nlist *nl); // every interaction on the list can run on the tile kernel
bool parm_tile_usable(const parm_inter *it);  // this interaction can run on the tile kernel right now
bool parm_tile_chunk_range(const parm_nlist *nl, uint32_t first, uint32_t end, uint32_t *c0, uint32_t *c1);

// ---- device helpers ----
#ifdef __CUDACC__
// coordinate along the slab axis relative to the slab's lower face, continuous across both halos
__device__ __forceinline__ double shard_rel(double x, const ShardDev &sd) {
    double w = x - sd.L * floor(x / sd.L);
    double rel = w - sd.lo;
    if (rel < 0.0) rel += sd.L;
    if (rel >= sd.Ls + 0.5 * (sd.L - sd.Ls)) rel -= sd.L;
    return rel;
}
// IEEE remainder(dx, L) (box.hpp:69-72) without the libm loop: exact whenever rint picks
// the nearest integer; the fix-up restores exactness when dx*invL was mis-rounded next to a
// half-integer; exact ties follow remainder()'s round-half-even rule.
__device__ __forceinline__ double min_image_exact(double dx, double L, double invL, double halfL) {
    double q = rint(__dmul_rn(dx, invL));
    double r = __fma_rn(-q, L, dx);
    if (r > halfL) {
        r -= L;
        q += 1.0;
    } else if (r < -halfL) {
        r += L;
        q -= 1.0;
    }
    if (fabs(r) == halfL) {
        // tie: remainder() picks the even quotient
        double h = q * 0.5;
        if (h != rint(h)) r = -r;
    }
    return r;
}
// hot-loop form: same value as min_image_exact except on exact ties / mis-rounded
// half-integers (|r| ~ L/2, never inside a cutoff when L > 2 r_cut)
__device__ __forceinline__ double min_image_fast(double dx, double L, double invL) {
    double q = rint(dx * invL);
    return fma(-q, L, dx);
}
// 256-bit read-only gather of one (x,y,z,m) record: a single LDG.E.256 (one 32-byte sector)
__device__ __forceinline__ double4 ld_pos4(const double4 *p) {
    double4 r;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ bool frozen_le(double m) { return m <= 0.0 || isinf(m); }
__device__ __forceinline__ bool frozen_eq(double m) { return m == 0.0 || isinf(m); }
#endif
