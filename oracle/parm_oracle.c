/* TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or executed from
 * the product path (parm_b200/). Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * Plain-C restatement of ParM's per-timestep MD hot path (the "port" oracle).
 * Parity pin: tests/test_oracle_pin.py checks every function here bit-for-bit
 * (pairs, forces, energies, trajectories) against the UNMODIFIED reference
 * compiled into oracle/_ref/libparm_ref{2,3}d.so, and against the committed
 * fixtures in tests/golden/ that were generated from that reference build.
 *
 * Arithmetic conventions (must match oracle/shim/Eigen/Dense, which defines
 * the reference build's vector arithmetic since Eigen itself is absent):
 *   3-element sums associate e0 + (e1 + e2); 2-element sums e0 + e1;
 *   norm = sqrt(squaredNorm); no FMA contraction (-ffp-contract=off).
 *
 * Each function cites the reference file:line it restates (paths relative to
 * /root/reference/src).
 */
#include <math.h>
#include <stdint.h>
#ifndef M_PI
#define M_PI 3.14159265358979323846 /* glibc math.h value; hidden by -std=c11 */
#endif
#include <stdlib.h>
#include <string.h>

typedef struct {
    double eps, sig, c, cutE; /* c: exponent (kinds 1, 9) or cut_distance (2,3,4,5,7,8) */
    double e2, s2, x, y;      /* second epsilon / sigma and functor-specific members (see update_pairs) */
    uint32_t i, j;            /* atom1 = first() = i (later atom), atom2 = j */
} Pair;

typedef struct {
    int kind;
    double *params; /* n x NPER, layout per kind: include/parm_b200.h (parm_inter_set_params_ex) */
    uint32_t *type;
    double *eps_table, *sig_table; /* ntypes x ntypes; NULL = not indexed */
    int ntypes;
    int nl;
    uint32_t last_update;
    Pair *pairs;
    size_t npairs, cappairs;
} Inter;

typedef struct {
    double skin;
    uint8_t *member;
    uint32_t *ids; /* SubGroup order */
    uint32_t nids;
    double *diam;     /* per SubGroup slot */
    double *lastlocs; /* per SubGroup slot x D */
    uint32_t *first, *last;
    size_t npairs, cap;
    uint32_t updatenum;
    int ignorechanged;
    int injected;
    uint64_t *ignored; /* PairList ignorepairs: sorted keys (larger index << 32 | smaller index), trackers.hpp:41-128 */
    size_t nignored;
} NList;

typedef struct { /* RsqTracker / ISFTracker / EnergyTracker, constraints.hpp:260-414 */
    int kind; /* 0 Rsq, 1 ISF, 2 Energy */
    int usecom, nskips, nks;
    unsigned long long *skips, *counts, curt;
    double *ks;
    double **past; /* per lag: n x D */
    double **acc;  /* per lag: Rsq xyz2 (n x D), xyz4 (n x D), r4 (n); ISF nks x n x D x 2 */
    unsigned n_skip, n_skipped, N;
    double U0, Es, Us, Ks, Esq, Usq, Ksq;
} Stat;

typedef struct {
    Stat *stats;
    int nstats;
    int D;
    uint32_t n;
    double L[3];
    double *x, *v, *a, *f, *m;
    Inter *inters;
    int ninters;
    NList *nls;
    int nnls;
    int integrator; /* -1 none, else PARM_INTEG_* of include/parm_b200.h */
    double dt, damping, force_mag, desT;
    double gamma, Q, xi, lns; /* Overdamped; NoseHoover / GaussianT */
    unsigned ncorrec;         /* Gear4A-6A */
    double *bs, *cs, *ds;
    /* CollectionNLCG members, collection.hpp:409-427 */
    double seceps, kappa, alphamax, afrac, dxmax, stepmax, kmax, P0, Knew, k, vl, fl, al;
    double alpha, beta, betaused, dxsum, alphavmax, maxdV;
    unsigned secmax, sec;
    double c0, c1, c2, sigmar, sigmav, corr, x11, x21, x22;
    const double *noise;
    size_t noise_len, noise_pos;
} Sys;

/* Eigen-shim reduction order: e0 + (e1 + e2) / e0 + e1 */
static inline double sum3(int D, double e0, double e1, double e2) { return D == 3 ? e0 + (e1 + e2) : e0 + e1; }
static inline double dotD(int D, const double *a, const double *b) {
    return sum3(D, a[0] * b[0], a[1] * b[1], D == 3 ? a[2] * b[2] : 0.0);
}

/* OriginBox::diff, box.hpp:103 -> vec_mod, box.hpp:69-78 */
static inline void box_diff(const Sys *s, const double *r1, const double *r2, double *out) {
    for (int k = 0; k < s->D; k++) out[k] = remainder(r1[k] - r2[k], s->L[k]);
}

static int frozen_le(double m) { return m <= 0 || isinf(m); } /* collection.cpp:445,458 */
static int frozen_eq(double m) { return m == 0 || isinf(m); } /* box.cpp:406, collection.cpp:304 */

void *port_sys_create(int ndim, uint32_t n, const double *L, const double *x, const double *v, const double *m) {
    Sys *s = (Sys *)calloc(1, sizeof(Sys));
    s->D = ndim;
    s->n = n;
    for (int k = 0; k < ndim; k++) s->L[k] = L[k];
    size_t nd = (size_t)n * ndim;
    s->x = (double *)calloc(nd ? nd : 1, 8);
    s->v = (double *)calloc(nd ? nd : 1, 8);
    s->a = (double *)calloc(nd ? nd : 1, 8);
    s->f = (double *)calloc(nd ? nd : 1, 8);
    s->m = (double *)calloc(n ? n : 1, 8);
    memcpy(s->x, x, nd * 8);
    if (v) memcpy(s->v, v, nd * 8);
    memcpy(s->m, m, (size_t)n * 8);
    s->integrator = -1;
    return s;
}

void port_sys_destroy(void *h) {
    Sys *s = (Sys *)h;
    for (int k = 0; k < s->ninters; k++) {
        free(s->inters[k].params);
        free(s->inters[k].type);
        free(s->inters[k].eps_table);
        free(s->inters[k].sig_table);
        free(s->inters[k].pairs);
    }
    for (int k = 0; k < s->nnls; k++) {
        NList *l = &s->nls[k];
        free(l->member); free(l->ids); free(l->diam); free(l->lastlocs); free(l->first); free(l->last); free(l->ignored);
    }
    for (int k = 0; k < s->nstats; k++) {
        Stat *t = &s->stats[k];
        for (int q = 0; q < t->nskips; q++) { free(t->past[q]); free(t->acc[q]); }
        free(t->past); free(t->acc); free(t->skips); free(t->counts); free(t->ks);
    }
    free(s->stats);
    free(s->inters); free(s->nls);
    free(s->x); free(s->v); free(s->a); free(s->f); free(s->m);
    free(s->bs); free(s->cs); free(s->ds);
    free(s);
}

#define NPER 5
/* A::max_size(): interaction.hpp:864, :905, :953-959, :1017, :1091, :1168, :1340, :1422, :1464, :1517-1523,
 * :1610, :1690, :1779 */
static double max_size(const Inter *I, uint32_t i) {
    const double *p = I->params + (size_t)NPER * i;
    double sigma = p[1];
    if (I->sig_table) {
        const double *row = I->sig_table + (size_t)I->type[i] * I->ntypes;
        sigma = row[0];
        for (int k = 1; k < I->ntypes; k++)
            if (sigma < row[k]) sigma = row[k];
    }
    switch (I->kind) {
        case 0: case 1: case 9: return sigma;
        case 6: return p[1];
        case 8: return p[1] + p[4] * (p[2] - 1);
        case 10: case 12: return p[1] * (1 + p[2] + p[3]);
        case 11: case 13: return p[1] * (1 + p[3]);
        default: return sigma * p[2];
    }
}

static void collection_update_trackers(Sys *s);

int port_add_interaction_ex(void *h, int kind, double skin, const double *params, int nper, const uint32_t *type,
                            const double *eps_table, const double *sig_table, int ntypes, const uint8_t *member,
                            int injected, int share_nl) {
    Sys *s = (Sys *)h;
    if (kind < 0 || kind > 13 || nper < 1 || nper > NPER) return -1;
    int nl = share_nl;
    if (nl < 0) { /* NeighborList ctor, trackers.cpp:10-17 */
        s->nls = (NList *)realloc(s->nls, sizeof(NList) * (s->nnls + 1));
        NList *l = &s->nls[s->nnls];
        memset(l, 0, sizeof(NList));
        l->skin = skin;
        l->member = (uint8_t *)calloc(s->n ? s->n : 1, 1);
        l->ids = (uint32_t *)calloc(s->n ? s->n : 1, 4);
        l->diam = (double *)calloc(s->n ? s->n : 1, 8);
        l->lastlocs = (double *)calloc((size_t)(s->n ? s->n : 1) * s->D, 8);
        l->ignorechanged = 1;
        l->injected = injected;
        nl = s->nnls++;
    }
    s->inters = (Inter *)realloc(s->inters, sizeof(Inter) * (s->ninters + 1));
    Inter *I = &s->inters[s->ninters];
    memset(I, 0, sizeof(Inter));
    I->kind = kind;
    I->nl = nl;
    I->params = (double *)calloc((size_t)(s->n ? s->n : 1) * NPER, 8);
    for (uint32_t i = 0; i < s->n; i++) memcpy(I->params + (size_t)NPER * i, params + (size_t)nper * i, 8 * (size_t)nper);
    I->type = (uint32_t *)calloc(s->n ? s->n : 1, 4);
    if (type) memcpy(I->type, type, (size_t)s->n * 4);
    I->ntypes = ntypes > 0 ? ntypes : 1;
    if (eps_table) {
        I->eps_table = (double *)calloc((size_t)I->ntypes * I->ntypes, 8);
        memcpy(I->eps_table, eps_table, (size_t)I->ntypes * I->ntypes * 8);
    }
    if (sig_table) {
        I->sig_table = (double *)calloc((size_t)I->ntypes * I->ntypes, 8);
        memcpy(I->sig_table, sig_table, (size_t)I->ntypes * I->ntypes * 8);
    }
    NList *l = &s->nls[nl];
    for (uint32_t i = 0; i < s->n; i++) { /* NListed::add -> NeighborList::add, interaction.hpp:1906-1910, trackers.hpp:194-201 */
        if (member && !member[i]) continue;
        if (l->member[i]) return -3; /* SubGroup::add throws on duplicates, box.hpp:495-501 */
        l->member[i] = 1;
        l->ids[l->nids] = i;
        l->diam[l->nids] = max_size(I, i);
        memcpy(l->lastlocs + (size_t)l->nids * s->D, s->x + (size_t)i * s->D, 8 * s->D);
        l->nids++;
        l->ignorechanged = 1;
    }
    return s->ninters++;
}

int port_add_interaction(void *h, int kind, double skin, const double *params, const uint32_t *type,
                         const double *eps_table, int ntypes, const uint8_t *member, int injected, int share_nl) {
    return port_add_interaction_ex(h, kind, skin, params, 3, type, kind == 2 ? eps_table : NULL, NULL, ntypes, member,
                                   injected, share_nl);
}

static void push_pair(NList *l, uint32_t a, uint32_t b) {
    if (l->npairs == l->cap) {
        l->cap = l->cap ? l->cap * 2 : 1024;
        l->first = (uint32_t *)realloc(l->first, l->cap * 4);
        l->last = (uint32_t *)realloc(l->last, l->cap * 4);
    }
    l->first[l->npairs] = a;
    l->last[l->npairs] = b;
    l->npairs++;
}

static int cmp_u32(const void *a, const void *b) {
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return x < y ? -1 : x > y;
}

/* predicate of trackers.cpp:65-66 on SubGroup slots i, j */
static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}
/* PairList::has_pair, trackers.hpp:66-71 */
static int nl_ignored(const NList *l, uint32_t a, uint32_t b) {
    if (!l->nignored) return 0;
    uint64_t key = a > b ? ((uint64_t)a << 32 | b) : ((uint64_t)b << 32 | a);
    return bsearch(&key, l->ignored, l->nignored, 8, cmp_u64) != NULL;
}

static inline int nl_pred(const Sys *s, const NList *l, uint32_t i, uint32_t j) {
    if (nl_ignored(l, l->ids[i], l->ids[j])) return 0; /* trackers.cpp:64 */
    double d[3] = {0, 0, 0};
    double diam = (l->diam[i] + l->diam[j]) / 2;
    box_diff(s, s->x + (size_t)l->ids[i] * s->D, s->x + (size_t)l->ids[j] * s->D, d);
    return sqrt(sum3(s->D, d[0] * d[0], d[1] * d[1], d[2] * d[2])) < (diam + l->skin);
}

/* NeighborList::update_list, trackers.cpp:19-85 */
static int nl_update_list(Sys *s, NList *l, int force) {
    const int D = s->D;
    if (!force && !l->ignorechanged) { /* trackers.cpp:23-53 */
        double bigdist = 0, biggestdist = 0;
        for (uint32_t i = 0; i < l->nids; i++) {
            const double *x = s->x + (size_t)l->ids[i] * D, *x0 = l->lastlocs + (size_t)i * D;
            double d0 = x[0] - x0[0], d1 = x[1] - x0[1], d2 = D == 3 ? x[2] - x0[2] : 0;
            double curdist = sqrt(sum3(D, d0 * d0, d1 * d1, d2 * d2));
            if (curdist > biggestdist) {
                bigdist = biggestdist;
                biggestdist = curdist;
            } else if (curdist > bigdist) {
                bigdist = curdist;
            } else
                continue;
            if (bigdist + biggestdist >= l->skin) {
                force = 1;
                break;
            }
        }
        if (!force) return 0;
    }
    l->updatenum++;
    l->ignorechanged = 0;
    l->npairs = 0;
    const uint32_t N = l->nids;
    int usecells = l->injected;
    int nc[3] = {1, 1, 1};
    size_t ncell = 1;
    if (usecells) {
        double maxd = 0;
        for (uint32_t i = 0; i < N; i++)
            if (l->diam[i] > maxd) maxd = l->diam[i];
        double rc = (maxd + l->skin) * (1 + 1e-9) + 1e-9;
        for (int d = 0; d < D; d++) {
            nc[d] = (int)floor(s->L[d] / rc);
            if (nc[d] < 3) usecells = 0;
            if (nc[d] > 256) nc[d] = 256;
            ncell *= (size_t)(nc[d] < 1 ? 1 : nc[d]);
        }
    }
    if (!usecells) { /* trackers.cpp:59-69 verbatim order */
        for (uint32_t i = 0; i < N; i++) {
            memcpy(l->lastlocs + (size_t)i * D, s->x + (size_t)l->ids[i] * D, 8 * D);
            for (uint32_t j = 0; j < i; j++)
                if (nl_pred(s, l, i, j)) push_pair(l, l->ids[i], l->ids[j]);
        }
        return 1;
    }
    /* cell-list pair finder: same predicate, same (i asc, j asc) output order */
    int *cidx = (int *)malloc((size_t)N * D * sizeof(int));
    uint32_t *cellstart = (uint32_t *)calloc(ncell + 1, 4), *cellatoms = (uint32_t *)malloc((size_t)N * 4);
    size_t *cellof = (size_t *)malloc((size_t)N * sizeof(size_t));
    for (uint32_t i = 0; i < N; i++) {
        const double *x = s->x + (size_t)l->ids[i] * D;
        memcpy(l->lastlocs + (size_t)i * D, x, 8 * D);
        size_t c = 0;
        for (int d = 0; d < D; d++) {
            double w = x[d] - s->L[d] * floor(x[d] / s->L[d]);
            int k = (int)floor(w / s->L[d] * nc[d]);
            if (k < 0) k = 0;
            if (k >= nc[d]) k = nc[d] - 1;
            cidx[(size_t)i * D + d] = k;
            c = c * nc[d] + k;
        }
        cellof[i] = c;
        cellstart[c + 1]++;
    }
    for (size_t c = 0; c < ncell; c++) cellstart[c + 1] += cellstart[c];
    uint32_t *fill = (uint32_t *)malloc(ncell * 4);
    memcpy(fill, cellstart, ncell * 4);
    for (uint32_t i = 0; i < N; i++) cellatoms[fill[cellof[i]]++] = i;
    free(fill);
    int nst = D == 3 ? 27 : 9;
    uint32_t *js = NULL;
    size_t njs = 0, capjs = 0;
    for (uint32_t i = 0; i < N; i++) {
        njs = 0;
        for (int st = 0; st < nst; st++) {
            int off[3], t = st;
            for (int d = D - 1; d >= 0; d--) { off[d] = t % 3 - 1; t /= 3; }
            size_t c = 0;
            for (int d = 0; d < D; d++) c = c * nc[d] + (size_t)((cidx[(size_t)i * D + d] + off[d] + nc[d]) % nc[d]);
            for (uint32_t q = cellstart[c]; q < cellstart[c + 1]; q++) {
                uint32_t j = cellatoms[q];
                if (j >= i) break;
                if (nl_pred(s, l, i, j)) {
                    if (njs == capjs) { capjs = capjs ? capjs * 2 : 256; js = (uint32_t *)realloc(js, capjs * 4); }
                    js[njs++] = j;
                }
            }
        }
        qsort(js, njs, 4, cmp_u32);
        for (size_t q = 0; q < njs; q++) push_pair(l, l->ids[i], l->ids[js[q]]);
    }
    free(js); free(cidx); free(cellstart); free(cellatoms); free(cellof);
    return 1;
}

static double dmax(double a, double b) { return a < b ? b : a; } /* std::max */
static double dmin(double a, double b) { return a < b ? a : b; } /* (a1.C < a2.C ? a1.C : a2.C) */

/* LJAttract::energy(rsig), interaction.hpp:170-177 */
static double ljattract_energy_rsig(double rsig) {
    if (rsig < 1) return -1;
    double rsq = rsig * rsig;
    double rsix = rsq * rsq * rsq;
    double mid = (1 - 1 / rsix);
    return mid * mid - 1;
}

/* LJAttractRepulseSigsPair::attract_energy / repulse_energy :1195-1205 (eps_r = eps, eps_a = e2, sig_r = sig, sig_a = s2) */
static double sigs_attract_energy(const Pair *P, double r) {
    double r_over_sig = (r - P->sig + P->s2) / P->s2;
    double mid = (1 - pow(r_over_sig, -6));
    return P->e2 * (mid * mid) - P->e2;
}
static double sigs_repulse_energy(const Pair *P, double r) {
    double r_over_sig = r / P->sig;
    double mid = (1 - pow(r_over_sig, -6));
    return P->eps * (mid * mid) - P->e2;
}

/* NListed<A,P>::update_pairs, interaction.hpp:2102-2115; pair constructors:
 * LJRepulsePair :878-883, RepulsionPair :1531-1542, LJAttractRepulsePair :1255-1270,
 * LennardJonesCutPair :970-979 + LennardJonesCut ctor :247-252, LJAttractCutPair :1023-1040 + LJAttractCut ctor
 * :205-209, LJAttractFixedRepulsePair :1347-1364, EisMclachlanPair :1430-1439, LJishPair :1098-1114,
 * LJAttractRepulseSigsPair :1175-1193, RepulsionDragPair :1617-1623, LoisOhernPair :1696-1712 (+MinCLs :1747-1751),
 * LoisLinPair :1785-1801 (+Min :1832-1836) */
static void update_pairs(Sys *s, Inter *I) {
    NList *l = &s->nls[I->nl];
    if (I->last_update == l->updatenum) return;
    I->last_update = l->updatenum;
    if (l->npairs > I->cappairs) {
        I->cappairs = l->npairs;
        I->pairs = (Pair *)realloc(I->pairs, I->cappairs * sizeof(Pair));
    }
    I->npairs = l->npairs;
    const int nt = I->ntypes;
    for (size_t k = 0; k < l->npairs; k++) {
        uint32_t i = l->first[k], j = l->last[k];
        const double *p1 = I->params + (size_t)NPER * i, *p2 = I->params + (size_t)NPER * j;
        const size_t tij = (size_t)I->type[i] * nt + I->type[j]; /* a1.epsilons[a2.indx] */
        Pair P;
        memset(&P, 0, sizeof(P));
        P.i = i; P.j = j;
        const int kind = I->kind;
        if (kind == 0) {
            P.eps = sqrt(p1[0] * p2[0]);
            P.sig = (p1[1] + p2[1]) / 2;
        } else if (kind == 1) {
            if (I->eps_table) {
                P.eps = I->eps_table[tij];
                P.sig = I->sig_table[tij];
            } else {
                P.eps = sqrt(p1[0] * p2[0]);
                P.sig = (p1[1] + p2[1]) / 2.0;
            }
            P.c = (p1[2] + p2[2]) / 2.0;
        } else if (kind == 2) {
            P.eps = I->eps_table[tij];
            P.sig = (p1[1] + p2[1]) / 2.0;
            P.c = dmax(p1[2], p2[2]); /* max(a1.sigcut, a2.sigcut) */
            if (P.eps <= 0) {
                P.c = 1;
                P.cutE = 0;
                P.eps = fabs(P.eps);
            } else {
                double mid = (1 - pow(P.c, -6));
                P.cutE = P.eps * (mid * mid);
            }
        } else if (kind == 3 || kind == 4) {
            if (I->eps_table && I->sig_table) {
                P.eps = I->eps_table[tij];
                P.sig = I->sig_table[tij];
            } else if (I->eps_table) { /* LJAttractCutPair(IEpsSigCutAtom, IEpsSigCutAtom) :1035-1040 */
                P.eps = I->eps_table[tij];
                P.sig = (p1[1] + p2[1]) / 2;
            } else {
                P.eps = sqrt(p1[0] * p2[0]);
                P.sig = (p1[1] + p2[1]) / 2;
            }
            P.c = dmax(p1[2], p2[2]);
            if (kind == 3) {
                double rsix = pow(P.c, 6);
                double mid = (1 - 1 / rsix);
                P.cutE = P.eps * (mid * mid - 1);
            } else {
                P.cutE = ljattract_energy_rsig(P.c) * P.eps;
            }
        } else if (kind == 5) {
            double e12 = I->eps_table[tij];
            P.eps = fabs(e12);
            P.e2 = sqrt(p1[3] * p2[3]); /* repeps */
            P.sig = (p1[1] + p2[1]) / 2.0;
            P.c = dmax(p1[2], p2[2]);
            if (!(e12 > 0) || P.eps == 0) {
                P.c = 1;
                P.cutE = 0;
                P.eps = 0;
            } else {
                double mid = (1 - pow(P.c, -6.0));
                P.cutE = P.eps * (mid * mid);
            }
        } else if (kind == 6) { /* p = (sigmai, dist): c0 = eps, c1 = x, c2 = y, cutoff = sig */
            P.eps = -M_PI * (p1[0] * p2[1] - p2[0] * p1[1]) * (p1[1] * p1[1] - p2[1] * p2[1]);
            P.x = -2 * M_PI * (p1[0] * p2[1] * p2[1] + p2[0] * p1[1] * p1[1]);
            P.y = M_PI * (p1[0] * p2[1] + p2[0] * p1[1]);
            P.sig = p1[1] + p2[1];
        } else if (kind == 7) { /* n = x */
            P.eps = I->eps_table[tij];
            P.e2 = sqrt(p1[3] * p2[3]);
            P.sig = (p1[1] + p2[1]) / 2.0;
            P.x = (p1[4] + p2[4]) / 2;
            P.c = dmax(p1[2], p2[2]);
            if (P.eps <= 0) {
                P.c = 1;
                P.cutE = 0;
                P.eps = 0;
            } else {
                double mid = (1 - pow(P.c, -P.x));
                P.cutE = P.eps * (mid * mid);
            }
        } else if (kind == 8) {
            P.eps = sqrt(p1[0] * p2[0]);
            P.e2 = sqrt(p1[3] * p2[3]);
            P.sig = (p1[1] + p2[1]) / 2.0;
            P.s2 = (p1[4] + p2[4]) / 2.0;
            P.c = dmax(p1[2], p2[2]);
            if (P.c > 0) { /* (cut_energy stays uninitialised in the reference when cut_distance <= 0; 0 here) */
                double cdu = P.sig + P.s2 * (P.c - 1);
                if (P.c >= 1)
                    P.cutE = sigs_attract_energy(&P, cdu);
                else
                    P.cutE = sigs_repulse_energy(&P, cdu);
            }
        } else if (kind == 9) { /* gamma = x */
            P.eps = sqrt(p1[0] * p2[0]);
            P.sig = (p1[1] + p2[1]) / 2.0;
            P.c = (p1[2] + p2[2]) / 2.0;
            P.x = (p1[3] + p2[3]) / 2;
        } else if (kind == 10 || kind == 12) { /* C = x, l = y, sigcut = s2 */
            P.eps = sqrt(p1[0] * p2[0]);
            P.sig = (p1[1] + p2[1]) / 2.0;
            P.x = kind == 10 ? (p1[2] + p2[2]) / 2.0 : dmin(p1[2], p2[2]);
            P.y = kind == 10 ? (p1[3] + p2[3]) / 2.0 : dmin(p1[3], p2[3]);
            P.s2 = P.sig * (1 + P.x + P.y);
        } else { /* 11, 13: f = x, l = y, sigcut = s2 */
            P.eps = sqrt(p1[0] * p2[0]);
            P.sig = (p1[1] + p2[1]) / 2.0;
            P.x = kind == 11 ? (p1[2] + p2[2]) / 2.0 : dmin(p1[2], p2[2]);
            P.y = kind == 11 ? (p1[3] + p2[3]) / 2.0 : dmin(p1[3], p2[3]);
            P.s2 = P.sig + P.y;
        }
        I->pairs[k] = P;
    }
}

/* P::forces(box): LJRepulsive::forces interaction.hpp:135-151; RepulsionPair::forces :1544-1550;
 * LJAttractRepulsePair::forces :1289-1298; LennardJonesCut::forces :259-267; LJAttractCut::forces :222-232;
 * LJAttractFixedRepulsePair::forces :1388-1394; EisMclachlanPair::forces :1446-1452; LJishPair::forces :1130-1141;
 * LJAttractRepulseSigsPair::forces :1222-1243; RepulsionDragPair::forces :1631-1641; LoisOhernPair::forces
 * :1729-1743; LoisLinPair::forces :1816-1828.
 * Returns 0 when the force is exactly Vec::Zero(). rij = diff(atom1->x, atom2->x). */
static int pair_forces(const Sys *s, const Inter *I, const Pair *P, double *rij, double *f) {
    const int D = s->D;
    const int kind = I->kind;
    rij[2] = 0;
    f[0] = f[1] = f[2] = 0;
    if (kind == 2 && P->eps == 0) return 0;
    box_diff(s, s->x + (size_t)P->i * D, s->x + (size_t)P->j * D, rij);
    double dsq = sum3(D, rij[0] * rij[0], rij[1] * rij[1], rij[2] * rij[2]);
    double scal;
    if (kind == 0) {
        double rsq = dsq / (P->sig * P->sig);
        if (rsq > 1) return 0;
        double rsix = rsq * rsq * rsq;
        double fmagTimesR = 12 * P->eps / rsix * (1 / rsix - 1);
        scal = fmagTimesR / dsq;
    } else if (kind == 1) {
        if (dsq > P->sig * P->sig) return 0;
        double R = sqrt(dsq);
        scal = P->eps * pow(1.0 - (R / P->sig), P->c - 1) / P->sig / R;
    } else if (kind == 2) {
        double rsq = dsq / (P->sig * P->sig);
        if (rsq > (P->c * P->c)) return 0;
        double rsix = pow(rsq, -3);
        double fmagTimesR = 12 * P->eps * rsix * (rsix - 1);
        scal = fmagTimesR / dsq;
    } else if (kind == 3) {
        double rsq = dsq / (P->sig * P->sig);
        if (rsq > (P->c * P->c)) return 0;
        double rsix = rsq * rsq * rsq;
        double fmagTimesR = 12 * P->eps / rsix * (1 / rsix - 1);
        scal = fmagTimesR / dsq;
    } else if (kind == 4) {
        if (P->eps == 0) return 0;
        double rsq = dsq / (P->sig * P->sig);
        if (rsq < 1 || rsq > (P->c * P->c)) return 0;
        double rsix = rsq * rsq * rsq;
        double fmagTimesR = 12 * P->eps / rsix * (1 / rsix - 1);
        scal = fmagTimesR / dsq;
    } else if (kind == 5) {
        double rsq = dsq / (P->sig * P->sig);
        if (rsq > (P->c * P->c)) return 0;
        double rsix = pow(rsq, -3);
        double fmagTimesR = 12 * rsix * (rsix - 1);
        scal = rsq < 1 ? P->e2 * fmagTimesR / dsq : P->eps * fmagTimesR / dsq;
    } else if (kind == 6) {
        if (dsq > (P->sig * P->sig)) return 0;
        double R = sqrt(dsq);
        scal = (P->eps / dsq - P->y) / R;
    } else if (kind == 7) {
        double rsq = dsq / (P->sig * P->sig);
        if (rsq > P->c * P->c) return 0;
        double rmid = pow(rsq, -P->x / 2);
        double fmagTimesR = 2 * P->x * rmid * (rmid - 1);
        scal = rsq < 1 ? P->e2 * fmagTimesR / dsq : P->eps * fmagTimesR / dsq;
    } else if (kind == 8) {
        double cdu = P->sig + P->s2 * (P->c - 1);
        if (P->c > 0 && dsq > cdu * cdu) return 0;
        if (dsq > P->sig * P->sig) {
            double dist = sqrt(dsq);
            double rminus = dist - P->sig + P->s2;
            double r_over_sig = rminus / P->s2;
            double rsix = pow(r_over_sig, -6);
            double fmag_over_r = 12 * rsix * (rsix - 1) / (dist * rminus);
            scal = P->e2 * fmag_over_r;
        } else {
            double r_over_sigsq = dsq / (P->sig * P->sig);
            double rsix = pow(r_over_sigsq, -3);
            double fmagTimesR = 12 * rsix * (rsix - 1);
            scal = P->eps * fmagTimesR / dsq;
        }
    } else if (kind == 9) {
        if (dsq > P->sig * P->sig) return 0;
        double vij[3] = {0, 0, 0};
        for (int k = 0; k < D; k++) vij[k] = s->v[(size_t)P->i * D + k] - s->v[(size_t)P->j * D + k];
        double R = sqrt(dsq);
        double vr = sum3(D, vij[0] * rij[0], vij[1] * rij[1], vij[2] * rij[2]);
        double el = P->eps * pow(1.0 - (R / P->sig), P->c - 1) / P->sig / R;
        for (int k = 0; k < D; k++) { /* rij * el - (rij * vr / dsq) * gamma */
            double v_perp = rij[k] * vr / dsq;
            f[k] = rij[k] * el - v_perp * P->x;
        }
        return 1;
    } else if (kind == 10 || kind == 12) {
        if (dsq >= P->s2 * P->s2) return 0;
        double R = sqrt(dsq);
        double rsig = R / P->sig;
        if (rsig <= 1 + P->x) {
            double dR = rsig - 1;
            scal = -P->eps * dR / R;
        } else {
            double dR2 = rsig - (P->x + P->y + 1);
            scal = P->x * P->eps / P->y * dR2 / R;
        }
    } else {
        if (dsq >= P->s2 * P->s2) return 0;
        double R = sqrt(dsq);
        if (R <= P->sig) {
            double dR = 1.0 - (R / P->sig);
            scal = P->eps * dR / P->sig / R;
        } else {
            for (int k = 0; k < D; k++) f[k] = -rij[k] * (P->x / R);
            return 1;
        }
    }
    for (int k = 0; k < D; k++) f[k] = rij[k] * scal;
    return 1;
}

/* P::energy(box): LJRepulsive::energy :126-133; RepulsionPair::energy :1537-1543;
 * LJAttractRepulsePair::energy :1271-1288; LennardJonesCut::energy :253-258; LJAttractCut::energy :216-221;
 * LJAttractFixedRepulsePair::energy :1365-1387; EisMclachlanPair::energy :1440-1445; LJishPair::energy :1115-1129;
 * LJAttractRepulseSigsPair::energy :1207-1220; RepulsionDragPair::energy :1624-1630; LoisOhernPair::energy
 * :1714-1727; LoisLinPair::energy :1802-1814 */
static double pair_energy(const Sys *s, const Inter *I, const Pair *P) {
    const int D = s->D;
    const int kind = I->kind;
    double rij[3] = {0, 0, 0};
    box_diff(s, s->x + (size_t)P->i * D, s->x + (size_t)P->j * D, rij);
    double dsq = sum3(D, rij[0] * rij[0], rij[1] * rij[1], rij[2] * rij[2]);
    if (kind == 0) {
        double rsq = dsq / (P->sig * P->sig);
        if (rsq > 1) return 0;
        double rsix = rsq * rsq * rsq;
        double mid = (1 - 1 / rsix);
        return P->eps * (mid * mid);
    } else if (kind == 1 || kind == 9) {
        if (dsq > P->sig * P->sig) return 0.0;
        double R = sqrt(dsq);
        return P->eps * pow(1.0 - (R / P->sig), P->c) / P->c;
    } else if (kind == 2) {
        double rsq = dsq / (P->sig * P->sig);
        if (rsq > P->c * P->c) return 0;
        double mid = (1 - pow(rsq, -3));
        return P->eps * (mid * mid) - P->cutE;
    } else if (kind == 3) {
        double rsq = dsq / (P->sig * P->sig);
        if (rsq > (P->c * P->c)) return 0;
        double rsix = rsq * rsq * rsq;
        double mid = (1 - 1 / rsix);
        return P->eps * (mid * mid - 1) - P->cutE;
    } else if (kind == 4) {
        if (P->eps == 0) return 0;
        if (dsq > (P->c * P->c * P->sig * P->sig)) return 0;
        double rsq = dsq / (P->sig * P->sig); /* LJAttract::energy(diff, eps, sig) :161-168 */
        double lj;
        if (rsq < 1)
            lj = -P->eps;
        else {
            double rsix = rsq * rsq * rsq;
            double mid = (1 - 1 / rsix);
            lj = P->eps * (mid * mid - 1);
        }
        return lj - P->cutE;
    } else if (kind == 5) {
        double rsq = dsq / (P->sig * P->sig);
        if (rsq > P->c * P->c) return 0;
        double mid = (1 - pow(rsq, -3));
        if (rsq > 1) return P->eps * (mid * mid) - P->cutE;
        return P->e2 * (mid * mid) - P->cutE;
    } else if (kind == 6) {
        double R = sqrt(dsq);
        if (R > P->sig) return 0;
        return P->eps / R + P->x + P->y * R;
    } else if (kind == 7) {
        double rsq = dsq / (P->sig * P->sig);
        if (rsq > P->c * P->c) return 0;
        double mid = (1 - pow(rsq, -P->x / 2));
        if (rsq > 1) return P->eps * (mid * mid) - P->cutE;
        return P->e2 * (mid * mid) - P->cutE;
    } else if (kind == 8) {
        double dist = sqrt(dsq);
        double cdu = P->sig + P->s2 * (P->c - 1);
        if (P->c > 0 && dist > cdu) return 0;
        if (dist <= P->sig) return sigs_repulse_energy(P, dist) - P->cutE;
        return sigs_attract_energy(P, dist) - P->cutE;
    } else if (kind == 10 || kind == 12) {
        if (dsq >= P->s2 * P->s2) return 0.0;
        double R = sqrt(dsq) / P->sig;
        if (R <= 1 + P->x) {
            double dR = R - 1;
            return -P->eps * P->sig / 2 * (P->x * (P->x + P->y) - dR * dR);
        }
        double dR2 = P->x + P->y + 1 - R;
        return -P->x * P->eps * P->sig / 2 / P->y * dR2 * dR2;
    } else {
        if (dsq >= P->s2 * P->s2) return 0.0;
        double R = sqrt(dsq);
        if (R <= P->sig) {
            double dR = 1.0 - (R / P->sig);
            return P->eps / 2 * dR * dR - P->x * P->y;
        }
        return -P->x * (P->sig + P->y - R);
    }
}

/* NListed::set_forces :2166-2175, set_forces_get_pressure :2232-2245, pressure :2248-2261,
 * stress / set_forces_get_stress :2264-2291.  mode bit0: scatter forces; p_out: sum r.f; st: D x D */
static void inter_loop(Sys *s, Inter *I, int scatter, double *p_out, double *st) {
    const int D = s->D;
    update_pairs(s, I);
    double p = 0;
    if (st) memset(st, 0, sizeof(double) * D * D);
    for (size_t k = 0; k < I->npairs; k++) {
        const Pair *P = &I->pairs[k];
        double rij[3], f[3];
        pair_forces(s, I, P, rij, f);
        if (scatter)
            for (int d = 0; d < D; d++) {
                s->f[(size_t)P->i * D + d] += f[d];
                s->f[(size_t)P->j * D + d] -= f[d];
            }
        if (p_out || st) {
            double r[3] = {0, 0, 0};
            box_diff(s, s->x + (size_t)P->i * D, s->x + (size_t)P->j * D, r);
            if (p_out) p += sum3(D, r[0] * f[0], r[1] * f[1], r[2] * f[2]);
            if (st)
                for (int a = 0; a < D; a++)
                    for (int b = 0; b < D; b++) st[a * D + b] += r[a] * f[b];
        }
    }
    if (p_out) *p_out = p;
}

static double inter_energy(Sys *s, Inter *I) { /* NListed::energy :2154-2163 */
    update_pairs(s, I);
    double E = 0;
    for (size_t k = 0; k < I->npairs; k++) E += pair_energy(s, I, &I->pairs[k]);
    return E;
}

/* NListed::contacts :2126-2137 (E != 0), ::overlaps :2140-2151 (E > 0) */
static void inter_contacts(Sys *s, Inter *I, unsigned long long *c, unsigned long long *o) {
    update_pairs(s, I);
    *c = *o = 0;
    for (size_t k = 0; k < I->npairs; k++) {
        double E = pair_energy(s, I, &I->pairs[k]);
        if (E != 0.0) (*c)++;
        if (E > 0.0) (*o)++;
    }
}

/* ---- AtomGroup reductions, box.cpp:239-260, 401-431 ---- */
static double group_mass(const Sys *s) {
    double m = 0;
    for (uint32_t i = 0; i < s->n; i++)
        if (!frozen_le(s->m[i])) m += s->m[i];
    return m;
}
static void group_momentum(const Sys *s, double *tot) {
    tot[0] = tot[1] = tot[2] = 0;
    for (uint32_t i = 0; i < s->n; i++) {
        if (frozen_le(s->m[i])) continue;
        for (int d = 0; d < s->D; d++) tot[d] += s->v[(size_t)i * s->D + d] * s->m[i];
    }
}
static void group_com_velocity(const Sys *s, double *out) {
    group_momentum(s, out);
    double M = group_mass(s);
    for (int d = 0; d < s->D; d++) out[d] = out[d] / M;
}
static double group_ke(const Sys *s, const double *v0) {
    double totE = 0;
    for (uint32_t i = 0; i < s->n; i++) {
        if (frozen_eq(s->m[i])) continue;
        double c[3] = {0, 0, 0};
        for (int d = 0; d < s->D; d++) c[d] = s->v[(size_t)i * s->D + d] - v0[d];
        totE += s->m[i] / 2 * dotD(s->D, c, c);
    }
    return totE;
}

/* Collection::update_trackers, collection.cpp:45-50 -> NeighborList::update, trackers.hpp:173-176 */
static void stat_update(Sys *s, Stat *t);
static void collection_update_trackers(Sys *s) {
    for (int k = 0; k < s->nnls; k++) nl_update_list(s, &s->nls[k], 0);
    if (s->integrator >= 0) /* the statistics trackers are add_tracker()ed after the lists */
        for (int k = 0; k < s->nstats; k++) stat_update(s, &s->stats[k]);
}

/* Collection::set_forces, collection.cpp:159-179 */
static void collection_set_forces(Sys *s, int constraints_and_a) {
    const int D = s->D;
    memset(s->f, 0, (size_t)s->n * D * 8); /* reset_forces box.cpp:427-431 */
    for (int k = 0; k < s->ninters; k++) inter_loop(s, &s->inters[k], 1, NULL, NULL);
    if (!constraints_and_a) return;
    for (uint32_t i = 0; i < s->n; i++) {
        if (frozen_le(s->m[i])) {
            for (int d = 0; d < D; d++) s->a[(size_t)i * D + d] = 0;
            continue;
        }
        for (int d = 0; d < D; d++) s->a[(size_t)i * D + d] = s->f[(size_t)i * D + d] / s->m[i];
    }
}

/* CollectionSol::set_constants, collection.cpp:230-263; BivariateGauss::set vecrand.cpp:48-63 */
static int sol_set_constants(Sys *s) {
    if (s->force_mag <= 0.0) {
        s->c0 = 1; s->c1 = 1; s->c2 = .5; s->sigmar = 0; s->sigmav = 0; s->corr = 1;
    } else {
        double dampdt = s->force_mag * s->dt;
        s->c0 = exp(-dampdt);
        s->c1 = (-expm1(-dampdt)) / dampdt;
        s->c2 = (1 - s->c1) / dampdt;
        if (dampdt > 1e-4)
            s->sigmar = sqrt((1 / dampdt) * (2 - (-4 * expm1(-dampdt) + expm1(-2 * dampdt)) / dampdt));
        else
            s->sigmar = sqrt(2 * dampdt / 3 - dampdt * dampdt / 2 + 7 * dampdt * dampdt * dampdt / 30);
        s->sigmav = sqrt(-expm1(-2 * dampdt));
        double exdpdt = (-expm1(-dampdt));
        s->corr = exdpdt * exdpdt / dampdt / s->sigmar / s->sigmav;
    }
    if (!(s->sigmar >= 0) || !(s->sigmav >= 0) || !(s->corr >= 0) || !(s->corr <= 1)) return -1;
    s->x11 = s->sigmar;
    s->x21 = s->sigmav * s->corr;
    s->x22 = s->sigmav * sqrt(1 - s->corr * s->corr);
    return 0;
}

/* Collection ctor + initialize, collection.cpp:3-19; add_tracker/add_interaction collection.hpp:113-120 */
static void damped_set_constants(Sys *s) { /* CollectionDamped::set_constants, collection.cpp:342-354 */
    if (s->damping <= 0.0) {
        s->c0 = 1; s->c1 = 1; s->c2 = .5;
        return;
    }
    double dampdt = s->damping * s->dt;
    s->c0 = exp(-dampdt);
    s->c1 = (-expm1(-dampdt)) / dampdt;
    s->c2 = (1 - s->c1) / dampdt;
}

int port_make_collection_ex(void *h, int integrator, const double *p, int np) {
    Sys *s = (Sys *)h;
    if (integrator < 0 || integrator > 11 || np < 1) return -1;
    s->integrator = integrator;
    s->dt = p[0];
    s->xi = s->lns = 0;
    const size_t nd = (size_t)(s->n ? s->n : 1) * s->D;
    if (integrator == 1) {
        if (s->dt <= 0) return -2;
        s->damping = p[1];
        s->force_mag = p[1];
        s->desT = p[2];
        if (sol_set_constants(s)) return -2;
    } else if (integrator == 2) {
        if (s->dt <= 0) return -2;
        s->damping = p[1];
        damped_set_constants(s);
    } else if (integrator == 3) {
        s->damping = p[1];
        s->desT = p[2];
    } else if (integrator == 4) {
        s->gamma = np > 1 ? p[1] : 1.0;
    } else if (integrator == 5) {
        s->Q = p[1];
        s->desT = p[2];
    } else if (integrator == 11) { /* CollectionNLCG ctor, collection.cpp:494-523 */
        s->P0 = p[1];
        s->kappa = np > 2 ? p[2] : 10.0;
        s->kmax = np > 3 ? p[3] : 1000;
        s->secmax = (unsigned)(np > 4 ? p[4] : 40);
        s->seceps = np > 5 ? p[5] : 1e-20;
        s->alphamax = 2.0; s->afrac = 0; s->dxmax = 100; s->stepmax = 1e-3;
        s->Knew = 0; s->k = 0; s->vl = 0; s->fl = 0; s->al = 0; s->alpha = 0; s->beta = 0; s->betaused = 0;
        s->dxsum = 0; s->alphavmax = 0; s->maxdV = 0; s->sec = 0;
    } else if (integrator >= 8) {
        s->ncorrec = (unsigned)(np > 1 ? p[1] : 1);
        free(s->bs); free(s->cs); free(s->ds);
        s->bs = (double *)calloc(nd, 8); /* resetbs / resetbcs / resetbcds, collection.hpp:645, :679-683, :717-725 */
        s->cs = (double *)calloc(nd, 8);
        s->ds = (double *)calloc(nd, 8);
    }
    /* initialize() with empty interaction/tracker vectors: set_forces(true) zeroes f, a = f/m */
    memset(s->f, 0, (size_t)s->n * s->D * 8);
    for (uint32_t i = 0; i < s->n; i++)
        for (int d = 0; d < s->D; d++) s->a[(size_t)i * s->D + d] = frozen_le(s->m[i]) ? 0.0 : s->f[(size_t)i * s->D + d] / s->m[i];
    /* add_tracker(nl) x nnls: the k-th call updates trackers 0..k; add_interaction x ninters: all */
    for (int k = 0; k < s->nnls; k++)
        for (int q = 0; q <= k; q++) nl_update_list(s, &s->nls[q], 0);
    for (int k = 0; k < s->ninters; k++) collection_update_trackers(s);
    return 0;
}

void port_get_sol_constants(void *h, double *out) {
    Sys *s = (Sys *)h;
    out[0] = s->c0; out[1] = s->c1; out[2] = s->c2; out[3] = s->x11; out[4] = s->x21; out[5] = s->x22;
}

/* z: per step, per non-frozen atom in atom order: D normals (x1) then D normals (x2) */
void port_inject_noise(void *h, const double *z, size_t len) {
    Sys *s = (Sys *)h;
    s->noise = z;
    s->noise_len = len;
    s->noise_pos = 0;
}

int port_make_collection(void *h, int integrator, double dt, double damping, double T) {
    double p[3] = {dt, damping, T};
    if (integrator != 0 && integrator != 1) return -1;
    return port_make_collection_ex(h, integrator, p, 3);
}
void port_get_scalars(void *h, double *out) { out[0] = ((Sys *)h)->xi; out[1] = ((Sys *)h)->lns; }

/* CollectionVerlet::timestep, collection.cpp:442-469 */
static void verlet_timestep(Sys *s) {
    const int D = s->D;
    const double dt = s->dt;
    for (uint32_t i = 0; i < s->n; i++) {
        double *x = s->x + (size_t)i * D, *v = s->v + (size_t)i * D, *a = s->a + (size_t)i * D;
        if (frozen_le(s->m[i])) {
            for (int d = 0; d < D; d++) v[d] = 0;
            continue;
        }
        for (int d = 0; d < D; d++) {
            x[d] += v[d] * dt + a[d] * (dt * dt / 2);
            v[d] += a[d] * (dt / 2);
        }
    }
    collection_set_forces(s, 0);
    for (uint32_t i = 0; i < s->n; i++) {
        double *v = s->v + (size_t)i * D, *a = s->a + (size_t)i * D, *f = s->f + (size_t)i * D;
        if (frozen_le(s->m[i])) {
            for (int d = 0; d < D; d++) a[d] = 0;
            continue;
        }
        for (int d = 0; d < D; d++) {
            a[d] = f[d] / s->m[i];
            v[d] += a[d] * (dt / 2);
        }
    }
    collection_update_trackers(s);
}

/* CollectionSol::timestep, collection.cpp:265-322; gen_vecs vecrand.cpp:73-85 */
static void sol_timestep(Sys *s) {
    const int D = s->D;
    const double dt = s->dt;
    for (uint32_t i = 0; i < s->n; i++) {
        double *x = s->x + (size_t)i * D, *v = s->v + (size_t)i * D, *a = s->a + (size_t)i * D;
        if (frozen_le(s->m[i])) {
            for (int d = 0; d < D; d++) v[d] = 0;
            continue;
        }
        double v0 = sqrt(s->desT / s->m[i]);
        double r0 = dt * v0;
        double drG[3] = {0, 0, 0}, dvG[3] = {0, 0, 0};
        if (s->damping > 0) {
            double x1[3] = {0, 0, 0}, x2[3] = {0, 0, 0};
            if (s->noise && s->noise_pos + 2 * D <= s->noise_len) {
                for (int d = 0; d < D; d++) x1[d] = s->noise[s->noise_pos + d];
                for (int d = 0; d < D; d++) x2[d] = s->noise[s->noise_pos + D + d];
                s->noise_pos += 2 * D;
            }
            for (int d = 0; d < D; d++) {
                drG[d] = x1[d] * s->x11;
                dvG[d] = x1[d] * s->x21 + x2[d] * s->x22;
            }
        }
        for (int d = 0; d < D; d++) {
            double xn = x[d] + ((v[d] * (s->c1 * dt) + a[d] * (s->c2 * dt * dt)) + drG[d] * r0);
            double vn = (v[d] * s->c0 + a[d] * (dt * (s->c1 - s->c2))) + dvG[d] * v0;
            x[d] = xn;
            v[d] = vn;
        }
    }
    collection_set_forces(s, 0);
    for (uint32_t i = 0; i < s->n; i++) {
        double *a = s->a + (size_t)i * D, *f = s->f + (size_t)i * D;
        if (frozen_eq(s->m[i])) {
            for (int d = 0; d < D; d++) a[d] = 0;
            continue;
        }
        for (int d = 0; d < D; d++) a[d] = f[d] / s->m[i];
    }
    for (uint32_t i = 0; i < s->n; i++) {
        double *v = s->v + (size_t)i * D, *a = s->a + (size_t)i * D;
        if (frozen_eq(s->m[i])) {
            for (int d = 0; d < D; d++) v[d] = 0;
            continue;
        }
        for (int d = 0; d < D; d++) v[d] += a[d] * (dt * s->c2);
    }
    collection_update_trackers(s);
}

/* CollectionDamped::timestep, collection.cpp:356-381 */
static void damped_timestep(Sys *s) {
    const int D = s->D;
    const double dt = s->dt, c0 = s->c0, c1 = s->c1, c2 = s->c2;
    for (uint32_t i = 0; i < s->n; i++) {
        double *x = s->x + (size_t)i * D, *v = s->v + (size_t)i * D, *a = s->a + (size_t)i * D;
        if (frozen_le(s->m[i])) {
            for (int d = 0; d < D; d++) v[d] = 0;
            continue;
        }
        for (int d = 0; d < D; d++) {
            x[d] += v[d] * (c1 * dt) + a[d] * (c2 * dt * dt);
            v[d] = v[d] * c0 + a[d] * (dt * (c1 - c2));
        }
    }
    collection_set_forces(s, 1);
    for (uint32_t i = 0; i < s->n; i++) {
        if (frozen_le(s->m[i])) continue;
        for (int d = 0; d < D; d++) s->v[(size_t)i * D + d] += s->a[(size_t)i * D + d] * (dt * c2);
    }
    collection_update_trackers(s);
}

/* CollectionSolHT::timestep, collection.cpp:401-440; GaussVec(sqrt(2 T damping / dt)) :393, vecrand.hpp:227-240.
 * The Gaussian stream is injected (port_inject_noise): D standard normals per mobile atom per step, component order. */
static void solht_timestep(Sys *s) {
    const int D = s->D;
    const double dt = s->dt, damping = s->damping;
    double keepv = 1 - (damping * dt);
    double xpartfromv = dt - (dt * dt * damping / 2);
    double sigma = sqrt(2.0 * s->desT * damping / dt);
    for (uint32_t i = 0; i < s->n; i++) {
        double *x = s->x + (size_t)i * D, *v = s->v + (size_t)i * D, *a = s->a + (size_t)i * D;
        if (frozen_le(s->m[i])) {
            for (int d = 0; d < D; d++) v[d] = 0;
            continue;
        }
        for (int d = 0; d < D; d++) {
            x[d] += (v[d] * xpartfromv) + (a[d] * (.5 * dt * dt));
            v[d] = v[d] * keepv + a[d] * (dt / 2);
        }
    }
    collection_set_forces(s, 1);
    for (uint32_t i = 0; i < s->n; i++) {
        double *v = s->v + (size_t)i * D, *a = s->a + (size_t)i * D, *f = s->f + (size_t)i * D;
        if (frozen_le(s->m[i])) {
            for (int d = 0; d < D; d++) a[d] = 0;
            continue;
        }
        for (int d = 0; d < D; d++) {
            double z = 0;
            if (s->noise && s->noise_pos < s->noise_len) z = s->noise[s->noise_pos++];
            double g = z * sigma + 0.0; /* normal_distribution(0, sigma) */
            a[d] = (f[d] + g) / s->m[i];
            v[d] += a[d] * (dt / 2);
        }
    }
    collection_update_trackers(s);
}

/* CollectionOverdamped::timestep, collection.cpp:471-492 */
static void overdamped_timestep(Sys *s) {
    const int D = s->D;
    collection_set_forces(s, 0);
    for (uint32_t i = 0; i < s->n; i++) {
        double *v = s->v + (size_t)i * D, *a = s->a + (size_t)i * D, *f = s->f + (size_t)i * D;
        if (frozen_le(s->m[i])) {
            for (int d = 0; d < D; d++) v[d] = a[d] = 0;
            continue;
        }
        for (int d = 0; d < D; d++) {
            a[d] = f[d] / s->m[i];
            v[d] = a[d] * s->gamma;
        }
    }
    for (size_t q = 0; q < (size_t)s->n * D; q++) s->x[q] += s->v[q] * s->dt;
    collection_update_trackers(s);
}

/* solve_cubic, collection.cpp:2304-2353 */
static double solve_cubic(double a1, double a2, double a3, double closeto) {
    double Q = (a1 * a1 - 3 * a2) / 9;
    double Q3 = Q * Q * Q;
    double R = ((2 * a1 * a1 * a1) - (9 * a1 * a2) + 27 * a3) / 54;
    double R2 = R * R;
    if (Q3 >= R2) {
        double theta = acos(R / sqrt(Q3));
        double sqQ = -2 * sqrt(Q);
        double x1 = sqQ * cos(theta / 3) - (a1 / 3);
        double x2 = sqQ * cos((theta + (2 * M_PI)) / 3) - (a1 / 3);
        double x3 = sqQ * cos((theta + (4 * M_PI)) / 3) - (a1 / 3);
        double d1 = fabs(x1 - closeto), d2 = fabs(x2 - closeto), d3 = fabs(x3 - closeto);
        if (d1 < d2 && d1 < d3) return x1;
        if (d2 < d1 && d2 < d3) return x2;
        return x3;
    }
    double R2Q3 = cbrt(sqrt(R2 - Q3) + fabs(R));
    int sgn = (0 < R) - (R < 0);
    return -(sgn * (R2Q3 + (Q / R2Q3))) - (a1 / 3);
}

static double collection_ndof(const Sys *s) { /* collection.cpp:116-133 */
    int ndof = 0;
    for (uint32_t i = 0; i < s->n; i++)
        if (!frozen_le(s->m[i])) ndof += s->D;
    return ndof;
}

/* CollectionNoseHoover::timestep, collection.cpp:1170-1242 */
static void nosehoover_timestep(Sys *s) {
    const int D = s->D;
    const double dt = s->dt, Q = s->Q, T = s->desT;
    double z3[3] = {0, 0, 0};
    double ndof = collection_ndof(s);
    double Kt = 2 * group_ke(s, z3);
    for (uint32_t i = 0; i < s->n; i++) {
        double *x = s->x + (size_t)i * D, *v = s->v + (size_t)i * D, *a = s->a + (size_t)i * D;
        for (int d = 0; d < D; d++) {
            double Ftilde = (a[d] - (v[d] * s->xi));
            x[d] += v[d] * dt + Ftilde * (dt * dt / 2);
            v[d] += Ftilde * (dt / 2);
        }
    }
    s->lns += s->xi * dt + (Kt - ndof * T) * (dt * dt / 2 / Q);
    collection_set_forces(s, 0);
    for (uint32_t i = 0; i < s->n; i++)
        for (int d = 0; d < D; d++) {
            size_t q = (size_t)i * D + d;
            s->a[q] = s->f[q] / s->m[i];
            s->v[q] += s->a[q] * (dt / 2);
        }
    double Ky = 2 * group_ke(s, z3);
    double z0 = s->xi + (Kt - 2 * ndof * T) * (dt / 2 / Q);
    double z1 = Ky * 2 / dt / Q;
    s->xi = solve_cubic(4 / dt - z0, 4 / dt / dt - 4 * z0 / dt, -(z0 * 4 / dt / dt) - z1, s->xi);
    double ytov = 1 + s->xi * dt / 2;
    for (size_t q = 0; q < (size_t)s->n * D; q++) s->v[q] /= ytov;
    collection_update_trackers(s);
}

/* CollectionGaussianT::set_xi, collection.cpp:1251-1260 */
static double gaussiant_set_xi(Sys *s) {
    const int D = s->D;
    double num = 0, den = 0;
    for (uint32_t i = 0; i < s->n; i++) {
        const double *v = s->v + (size_t)i * D, *f = s->f + (size_t)i * D;
        num += dotD(D, f, v);
        den += dotD(D, v, v) * s->m[i];
    }
    s->xi = num / den;
    return s->xi;
}

/* CollectionGaussianT::timestep, collection.cpp:1268-1299 */
static void gaussiant_timestep(Sys *s) {
    const int D = s->D;
    const double dt = s->dt;
    for (uint32_t i = 0; i < s->n; i++) {
        double *x = s->x + (size_t)i * D, *v = s->v + (size_t)i * D, *a = s->a + (size_t)i * D;
        for (int d = 0; d < D; d++) {
            x[d] += v[d] * dt + a[d] * (dt * dt / 2);
            v[d] += (a[d] - (v[d] * s->xi)) * (dt / 2);
        }
    }
    collection_set_forces(s, 0);
    for (uint32_t i = 0; i < s->n; i++)
        for (int d = 0; d < D; d++) {
            size_t q = (size_t)i * D + d;
            s->a[q] = s->f[q] / s->m[i];
            s->v[q] += s->a[q] * (dt / 2);
        }
    double z = gaussiant_set_xi(s);
    s->xi = z / (1 - z * dt / 2);
    double ytov = 1 + s->xi * dt / 2;
    for (size_t q = 0; q < (size_t)s->n * D; q++) s->v[q] /= ytov;
    collection_update_trackers(s);
}

/* CollectionGear3A::timestep, collection.cpp:1301-1322 */
static void gear3a_timestep(Sys *s) {
    const int D = s->D;
    const double dt = s->dt;
    for (size_t q = 0; q < (size_t)s->n * D; q++) {
        s->x[q] += s->v[q] * dt + s->a[q] * (dt * dt / 2);
        s->v[q] += s->a[q] * dt;
    }
    collection_set_forces(s, 0);
    for (uint32_t i = 0; i < s->n; i++)
        for (int d = 0; d < D; d++) {
            size_t q = (size_t)i * D + d;
            double a = s->f[q] / s->m[i];
            double correction = a - s->a[q];
            s->v[q] += correction * (dt / 2);
            s->a[q] = a;
        }
    collection_update_trackers(s);
}

/* CollectionGear4A / 5A / 6A::timestep, collection.cpp:1324-1364, 1366-1407, 1409-1496 */
static void gear_timestep(Sys *s, int order) {
    const int D = s->D;
    const double dt = s->dt;
    const size_t nd = (size_t)s->n * D;
    double c0, c1, c3 = 0, c4 = 0, c5 = 0;
    if (order == 4) {
        for (size_t q = 0; q < nd; q++) {
            s->x[q] += s->v[q] * dt + s->a[q] * (dt * dt / 2) + s->bs[q] * (dt * dt * dt / 6);
            s->v[q] += s->a[q] * dt + s->bs[q] * (dt * dt / 2);
            s->a[q] += s->bs[q] * dt;
        }
        c0 = dt * dt / 12;
        c1 = 5 * dt / 12;
    } else {
        double dt2 = dt * dt / 2;
        double dt3 = dt * dt2 / 3;
        double dt4 = dt * dt3 / 4;
        double dt5 = dt * dt4 / 5;
        if (order == 5) {
            for (size_t q = 0; q < nd; q++) {
                s->x[q] += s->v[q] * dt + s->a[q] * dt2 + s->bs[q] * dt3 + s->cs[q] * dt4;
                s->v[q] += s->a[q] * dt + s->bs[q] * dt2 + s->cs[q] * dt3;
                s->a[q] += s->bs[q] * dt + s->cs[q] * dt2;
                s->bs[q] += s->cs[q] * dt;
            }
            c0 = 19 * dt * dt / 240; c1 = 3 * dt / 8;
            c3 = 3 / (2 * dt); c4 = 1 / (dt * dt);
        } else {
            for (size_t q = 0; q < nd; q++) {
                s->x[q] += s->v[q] * dt + s->a[q] * dt2 + s->bs[q] * dt3 + s->cs[q] * dt4 + s->ds[q] * dt5;
                s->v[q] += s->a[q] * dt + s->bs[q] * dt2 + s->cs[q] * dt3 + s->ds[q] * dt4;
                s->a[q] += s->bs[q] * dt + s->cs[q] * dt2 + s->ds[q] * dt3;
                s->bs[q] += s->cs[q] * dt + s->ds[q] * dt2;
                s->cs[q] += s->ds[q] * dt;
            }
            c0 = 3 * dt * dt / 40; c1 = 251 * dt / 720;
            c3 = 11 / (6 * dt); c4 = 2 / (dt * dt); c5 = 1 / (dt * dt * dt);
        }
    }
    for (unsigned m = 0; m < s->ncorrec; m++) {
        collection_set_forces(s, 0);
        for (uint32_t i = 0; i < s->n; i++)
            for (int d = 0; d < D; d++) {
                size_t q = (size_t)i * D + d;
                double a = s->f[q] / s->m[i];
                double correction = a - s->a[q];
                s->x[q] += correction * c0;
                s->v[q] += correction * c1;
                s->a[q] = a;
                if (order == 4) s->bs[q] += correction / dt;
                else s->bs[q] += correction * c3;
                if (order >= 5) s->cs[q] += correction * c4;
                if (order >= 6) s->ds[q] += correction * c5;
            }
    }
    collection_update_trackers(s);
}

/* ---- CollectionNLCG, collection.cpp:525-854 ---- */
static double box_volume(const Sys *s) { return s->D == 3 ? s->L[0] * s->L[1] * s->L[2] : s->L[0] * s->L[1]; }
static double nlcg_length_squared(const Sys *s) { return s->D == 3 ? pow(box_volume(s), 2.0 / 3.0) : box_volume(s); }
static double nlcg_dot(const Sys *s, const double *p, const double *q) { /* sum over mobile atoms of p_i . q_i */
    double r = 0;
    for (uint32_t i = 0; i < s->n; i++) {
        if (frozen_le(s->m[i])) continue;
        r += dotD(s->D, p + (size_t)i * s->D, q + (size_t)i * s->D);
    }
    return r;
}
static double nlcg_fdotf(const Sys *s) { return nlcg_dot(s, s->f, s->f) / nlcg_length_squared(s) + s->fl * s->fl; }
static double nlcg_fdota(const Sys *s) { return nlcg_dot(s, s->f, s->a) / nlcg_length_squared(s) + s->fl * s->al; }
static double nlcg_fdotv(const Sys *s) { return nlcg_dot(s, s->f, s->v) / nlcg_length_squared(s) + s->fl * s->vl; }
static double nlcg_vdotv(const Sys *s) { return nlcg_dot(s, s->v, s->v) / nlcg_length_squared(s) + s->vl * s->vl; }

static void nlcg_stepx(Sys *s, double dx) { /* :602-618 */
    double Lfac = exp(dx * s->vl / (s->kappa * s->D));
    for (size_t q = 0; q < (size_t)s->n * s->D; q++) {
        s->x[q] *= Lfac;
        s->x[q] += s->v[q] * dx;
    }
    for (int d = 0; d < s->D; d++) s->L[d] *= Lfac; /* OriginBox::resize(factor), box.cpp:3-6 */
}

/* Collection::set_forces_get_pressure(false), collection.cpp:181-208 */
static double collection_set_forces_get_pressure(Sys *s) {
    memset(s->f, 0, (size_t)s->n * s->D * 8);
    double p = 0;
    for (int k = 0; k < s->ninters; k++) {
        double pk;
        inter_loop(s, &s->inters[k], 1, &pk, NULL);
        p += pk;
    }
    for (uint32_t i = 0; i < s->n; i++)
        if (frozen_le(s->m[i]))
            for (int d = 0; d < s->D; d++) s->a[(size_t)i * s->D + d] = 0;
    return p;
}

static void nlcg_set_forces(Sys *s, int constraints_and_a, int setV) { /* :534-568 */
    double V = box_volume(s);
    if (setV) {
        double interacP = collection_set_forces_get_pressure(s);
        s->fl = ((interacP / s->D) - (s->P0 * V)) / s->kappa;
        if (constraints_and_a) {
            s->al = s->fl;
            s->vl = s->fl;
        }
    } else {
        collection_set_forces(s, 0);
    }
    if (constraints_and_a) {
        for (uint32_t i = 0; i < s->n; i++)
            for (int d = 0; d < s->D; d++) {
                size_t q = (size_t)i * s->D + d;
                double t = frozen_le(s->m[i]) ? 0.0 : s->f[q];
                s->a[q] = t;
                s->v[q] = t;
            }
        s->Knew = nlcg_fdota(s);
    }
}

static void nlcg_reset(Sys *s) { /* :525-532 */
    s->k = 0;
    nlcg_set_forces(s, 1, 1);
    memcpy(s->v, s->a, (size_t)s->n * s->D * 8);
    s->vl = s->al;
}

static void nlcg_descend(Sys *s) { /* :837-854 */
    nlcg_set_forces(s, 0, 1);
    for (uint32_t i = 0; i < s->n; i++) {
        if (frozen_le(s->m[i])) continue;
        for (int d = 0; d < s->D; d++) {
            size_t q = (size_t)i * s->D + d;
            s->v[q] = s->f[q];
            s->a[q] = s->f[q];
        }
    }
    s->al = s->fl;
    s->vl = s->fl;
    nlcg_stepx(s, s->dt);
    collection_update_trackers(s);
}

static void nlcg_timestep(Sys *s) { /* :682-835 */
    const double dt = s->dt;
    const int NDIM = s->D;
    nlcg_stepx(s, dt);
    nlcg_set_forces(s, 0, 1);
    double eta0 = -nlcg_fdotv(s);
    double eta;
    nlcg_stepx(s, -dt);
    collection_update_trackers(s);
    s->alpha = -dt;
    nlcg_set_forces(s, 0, 1);
    s->dxsum = 0;
    double vdv = nlcg_vdotv(s);
    for (s->sec = 0; s->sec < s->secmax; s->sec++) {
        eta = -nlcg_fdotv(s);
        double alphafac = -eta / fabs(eta0 - eta);
        if (fabs(eta0 - eta) <= 1e-12 * fabs(eta)) {
            alphafac = s->alphamax > 0 ? s->alphamax : 1.1;
            s->sec = s->sec > 0 ? s->sec * 2 - 1 : 1;
        }
        if (s->alphamax > 0 && alphafac > s->alphamax) alphafac = s->alphamax;
        if (s->alphamax > 0 && alphafac < -s->alphamax) alphafac = -s->alphamax;
        s->alpha = fabs(s->alpha) * alphafac;
        double newdxsum = fabs(s->dxsum + s->alpha);
        if (s->dxmax > 0 && newdxsum > s->dxmax) {
            s->k = 0;
            break;
        }
        double dVoverV = expm1(fabs(s->dxsum + s->alpha) * s->vl / (s->kappa * NDIM));
        if (s->maxdV > 0 && dVoverV > s->maxdV) {
            s->k = 0;
            break;
        }
        s->dxsum += s->alpha;
        nlcg_stepx(s, s->alpha);
        nlcg_set_forces(s, 0, 1);
        eta0 = eta;
        if (s->alpha * s->alpha * vdv < s->seceps * s->seceps) break;
        if ((s->sec > 1) && (s->afrac > 0) && (fabs(s->alpha) < fabs(s->dxsum)) && (fabs(s->alpha) / fabs(s->dxsum) < s->afrac)) break;
        if (s->stepmax > 0 && s->dxsum * s->dxsum * vdv > s->stepmax * s->stepmax) {
            s->k = 0;
            break;
        }
    }
    s->alphavmax = sqrt(s->alpha * s->alpha * vdv);
    double Kold = s->Knew;
    double Kmid = nlcg_fdota(s);
    for (uint32_t i = 0; i < s->n; i++)
        for (int d = 0; d < s->D; d++) {
            size_t q = (size_t)i * s->D + d;
            s->a[q] = frozen_le(s->m[i]) ? 0.0 : s->f[q];
        }
    s->al = s->fl;
    s->Knew = nlcg_fdota(s);
    s->beta = (s->Knew - Kmid) / Kold;
    s->betaused = s->beta;
    s->k++;
    if (s->k >= s->kmax || isinf(s->betaused) || isnan(s->betaused) || s->betaused <= 0) {
        s->k = 0;
        s->betaused = 0;
    } else if (s->betaused > 1) {
        s->betaused = 1;
    }
    for (uint32_t i = 0; i < s->n; i++)
        for (int d = 0; d < s->D; d++) {
            size_t q = (size_t)i * s->D + d;
            s->v[q] = frozen_le(s->m[i]) ? 0.0 : s->a[q] + s->v[q] * s->betaused;
        }
    s->vl = s->al + s->betaused * s->vl;
}

int port_nlcg_set(void *h, int which, double v) {
    Sys *s = (Sys *)h;
    if (s->integrator != 11) return -1;
    switch (which) {
        case 0: s->dt = v; nlcg_reset(s); break;
        case 1: s->P0 = v; nlcg_reset(s); break;
        case 2: s->kappa = v; nlcg_reset(s); break;
        case 3: s->alphamax = v; break;
        case 4: s->afrac = v; break;
        case 5: s->dxmax = v; break;
        case 6: s->stepmax = v; break;
        case 7: s->maxdV = v; break;
        case 8: s->kmax = v; break;
        case 9: s->secmax = (unsigned)v; break;
        case 10: s->seceps = v; break;
        default: return -1;
    }
    return 0;
}
int port_nlcg_get(void *h, double *o) {
    Sys *s = (Sys *)h;
    if (s->integrator != 11) return -1;
    o[0] = s->dt; o[1] = s->P0; o[2] = s->kappa; o[3] = s->Knew; o[4] = s->k; o[5] = s->vl; o[6] = s->fl; o[7] = s->al;
    o[8] = s->alpha; o[9] = s->beta; o[10] = s->betaused; o[11] = s->dxsum; o[12] = s->alphavmax; o[13] = s->sec;
    o[14] = s->kmax; o[15] = s->secmax;
    return 0;
}
int port_nlcg_set_forces(void *h, int caa, int setV) { Sys *s = (Sys *)h; if (s->integrator != 11) return -1; nlcg_set_forces(s, caa, setV); return 0; }
int port_nlcg_reset(void *h) { Sys *s = (Sys *)h; if (s->integrator != 11) return -1; nlcg_reset(s); return 0; }
int port_nlcg_descend(void *h) { Sys *s = (Sys *)h; if (s->integrator != 11) return -1; nlcg_descend(s); return 0; }
double port_potential_energy(void *h);
double port_nlcg_reduce(void *h, int what) {
    Sys *s = (Sys *)h;
    if (s->integrator != 11) return NAN;
    switch (what) {
        case 0: return nlcg_fdotf(s);
        case 1: return nlcg_fdota(s);
        case 2: return nlcg_fdotv(s);
        case 3: return nlcg_vdotv(s);
        case 4: { /* kinetic_energy :574-588 */
            double E = 0;
            double Lfac = exp(s->vl / (s->kappa * s->D));
            for (uint32_t i = 0; i < s->n; i++) {
                if (frozen_le(s->m[i])) continue;
                double w[3] = {0, 0, 0};
                for (int d = 0; d < s->D; d++) w[d] = s->v[(size_t)i * s->D + d] + (s->x[(size_t)i * s->D + d] * Lfac);
                E += dotD(s->D, w, w);
            }
            return E / 2.0;
        }
        case 5: { /* pressure :590-600 */
            double E = 0;
            for (int k = 0; k < s->ninters; k++) {
                double p;
                inter_loop(s, &s->inters[k], 0, &p, NULL);
                E += p;
            }
            return E / box_volume(s) / (double)s->D;
        }
        case 6: return port_potential_energy(h) + s->P0 * box_volume(s); /* :570-572 */
    }
    return NAN;
}
void port_get_box(void *h, double *L) { Sys *s = (Sys *)h; for (int d = 0; d < s->D; d++) L[d] = s->L[d]; }
void port_set_box(void *h, const double *L) { Sys *s = (Sys *)h; for (int d = 0; d < s->D; d++) s->L[d] = L[d]; }

/* ---- statistics trackers ---- */
static void group_com(const Sys *s, double *com) { /* AtomGroup::com, box.cpp:228-237 */
    double v[3] = {0, 0, 0};
    for (uint32_t i = 0; i < s->n; i++) {
        if (frozen_le(s->m[i])) continue;
        for (int d = 0; d < s->D; d++) v[d] += s->x[(size_t)i * s->D + d] * s->m[i];
    }
    double m = group_mass(s);
    for (int d = 0; d < s->D; d++) com[d] = v[d] / m;
}
static void stat_com(const Sys *s, const Stat *t, double *com) {
    com[0] = com[1] = com[2] = 0;
    if (t->usecom) group_com(s, com);
}
static size_t stat_acc_doubles(const Sys *s, const Stat *t) {
    return t->kind == 0 ? (size_t)(2 * s->D + 1) * s->n : (size_t)t->nks * s->n * s->D * 2;
}
static void stat_init_singles(Sys *s, Stat *t) { /* ctors constraints.cpp:404-416, 567-582; reset :418-427, :584-593 */
    double com[3];
    stat_com(s, t, com);
    for (int k = 0; k < t->nskips; k++) {
        memset(t->acc[k], 0, stat_acc_doubles(s, t) * 8);
        for (uint32_t i = 0; i < s->n; i++)
            for (int d = 0; d < s->D; d++) t->past[k][(size_t)i * s->D + d] = s->x[(size_t)i * s->D + d] - com[d];
        t->counts[k] = 0;
    }
}
static Stat *stat_new(Sys *s, int kind, const unsigned long long *ns, int nns, int usecom, const double *ks, int nks) {
    s->stats = (Stat *)realloc(s->stats, sizeof(Stat) * (s->nstats + 1));
    Stat *t = &s->stats[s->nstats++];
    memset(t, 0, sizeof(Stat));
    t->kind = kind;
    t->usecom = usecom;
    t->nskips = nns;
    t->nks = nks;
    t->skips = (unsigned long long *)calloc(nns ? nns : 1, 8);
    t->counts = (unsigned long long *)calloc(nns ? nns : 1, 8);
    if (nns) memcpy(t->skips, ns, 8 * (size_t)nns);
    t->ks = (double *)calloc(nks ? nks : 1, 8);
    if (nks) memcpy(t->ks, ks, 8 * (size_t)nks);
    t->past = (double **)calloc(nns ? nns : 1, sizeof(double *));
    t->acc = (double **)calloc(nns ? nns : 1, sizeof(double *));
    for (int k = 0; k < nns; k++) {
        t->past[k] = (double *)calloc((size_t)(s->n ? s->n : 1) * s->D, 8);
        t->acc[k] = (double *)calloc(stat_acc_doubles(s, t) ? stat_acc_doubles(s, t) : 1, 8);
    }
    return t;
}
static int stat_added(Sys *s) { /* Collection::add_tracker -> update_trackers(), collection.hpp:117-120 */
    if (s->integrator >= 0) collection_update_trackers(s);
    return s->nstats - 1;
}
int port_add_rsq_tracker(void *h, const unsigned long long *ns, int nns, int usecom) {
    Sys *s = (Sys *)h;
    stat_init_singles(s, stat_new(s, 0, ns, nns, usecom, NULL, 0));
    return stat_added(s);
}
int port_add_isf_tracker(void *h, const double *ks, int nks, const unsigned long long *ns, int nns, int usecom) {
    Sys *s = (Sys *)h;
    stat_init_singles(s, stat_new(s, 1, ns, nns, usecom, ks, nks));
    return stat_added(s);
}
int port_add_energy_tracker(void *h, unsigned n_skip) {
    Sys *s = (Sys *)h;
    Stat *t = stat_new(s, 2, NULL, 0, 0, NULL, 0);
    t->n_skip = n_skip > 1u ? n_skip : 1u;
    return stat_added(s);
}

static void stat_update(Sys *s, Stat *t) {
    const int D = s->D;
    if (t->kind == 2) { /* EnergyTracker::update, constraints.cpp:366-393 */
        if (t->n_skipped + 1 < t->n_skip) {
            t->n_skipped += 1;
            return;
        }
        t->n_skipped = 0;
        double curU = 0, curK = 0;
        for (uint32_t i = 0; i < s->n; i++) {
            const double *v = s->v + (size_t)i * D;
            curK += dotD(D, v, v) * s->m[i] / 2;
        }
        for (int k = 0; k < s->ninters; k++) curU += inter_energy(s, &s->inters[k]);
        curU -= t->U0;
        t->Ks += curK;
        t->Us += curU;
        t->Es += curK + curU;
        t->Ksq += curK * curK;
        t->Usq += curU * curU;
        t->Esq += (curK + curU) * (curK + curU);
        t->N++;
        return;
    }
    t->curt++; /* RsqTracker::update :492-499, ISFTracker::update :662-669 */
    double com[3];
    stat_com(s, t, com);
    for (int k = 0; k < t->nskips; k++) {
        if (t->curt % t->skips[k] != 0) continue;
        double *past = t->past[k], *acc = t->acc[k];
        for (uint32_t i = 0; i < s->n; i++) {
            double dr[3] = {0, 0, 0};
            for (int j = 0; j < D; j++) {
                double r = s->x[(size_t)i * D + j] - com[j];
                dr[j] = r - past[(size_t)i * D + j];
                past[(size_t)i * D + j] = r;
            }
            if (t->kind == 0) { /* RsqTracker1::update :429-455 */
                double *xyz2 = acc, *xyz4 = acc + (size_t)D * s->n, *r4 = acc + 2 * (size_t)D * s->n;
                double dist4 = 0;
                for (int j = 0; j < D; j++) {
                    double d2 = dr[j] * dr[j];
                    double d4 = d2 * d2;
                    dist4 += d2;
                    xyz2[(size_t)i * D + j] += d2;
                    xyz4[(size_t)i * D + j] += d4;
                }
                dist4 *= dist4;
                r4[i] += dist4;
            } else { /* ISFTracker1::update :595-619: += exp(i k dr_j) */
                for (int ki = 0; ki < t->nks; ki++)
                    for (int j = 0; j < D; j++) {
                        double *a = acc + (((size_t)ki * s->n + i) * D + j) * 2;
                        double arg = t->ks[ki] * dr[j];
                        a[0] += cos(arg); /* std::exp(complex(0, y)) = polar(1, y) = (cos y, sin y) */
                        a[1] += sin(arg);
                    }
            }
        }
        t->counts[k] += 1;
    }
}
void port_tracker_update(void *h, int t) { Sys *s = (Sys *)h; stat_update(s, &s->stats[t]); }
void port_tracker_reset(void *h, int k) {
    Sys *s = (Sys *)h;
    Stat *t = &s->stats[k];
    if (t->kind == 2) { /* EnergyTracker::reset, constraints.hpp:284-293 */
        t->n_skipped = 0; t->N = 0;
        t->Es = t->Us = t->Ks = t->Esq = t->Usq = t->Ksq = 0;
        return;
    }
    t->curt = 0;
    stat_init_singles(s, t);
}
void port_tracker_counts(void *h, int k, unsigned long long *out, int cap) {
    Stat *t = &((Sys *)h)->stats[k];
    for (int q = 0; q < cap && q < t->nskips; q++) out[q] = t->counts[q];
}
void port_rsq_read(void *h, int k, int single, double *xyz2, double *xyz4, double *r4) { /* :457-481 */
    Sys *s = (Sys *)h;
    Stat *t = &s->stats[k];
    const double *acc = t->acc[single];
    const double cnt = (double)t->counts[single];
    const size_t nd = (size_t)s->n * s->D;
    for (size_t q = 0; q < nd; q++) {
        if (xyz2) xyz2[q] = acc[q] / cnt;
        if (xyz4) xyz4[q] = acc[nd + q] / cnt;
    }
    for (uint32_t i = 0; i < s->n && r4; i++) r4[i] = acc[2 * nd + i] / cnt;
}
void port_isf_read(void *h, int k, int single, double *out) { /* ISFxyz :637-650: complex / complex(count, 0) */
    Sys *s = (Sys *)h;
    Stat *t = &s->stats[k];
    const double cnt = (double)t->counts[single];
    const size_t tot = stat_acc_doubles(s, t);
    for (size_t q = 0; q < tot; q++) out[q] = t->acc[single][q] / cnt;
}
void port_energy_tracker_read(void *h, int k, double *out) { /* accessors constraints.hpp:301-312 */
    Stat *t = &((Sys *)h)->stats[k];
    double N = (double)t->N;
    out[0] = N;
    out[1] = t->Es / N; out[2] = t->Us / N; out[3] = t->Ks / N;
    out[4] = t->Esq / t->N; out[5] = t->Usq / t->N; out[6] = t->Ksq / t->N;
    out[7] = t->U0;
}
void port_energy_tracker_set_U0(void *h, int k, int from_box, double U0) { /* :294-298, constraints.cpp:395-402 */
    Sys *s = (Sys *)h;
    Stat *t = &s->stats[k];
    if (from_box) {
        double curU = 0;
        for (int q = 0; q < s->ninters; q++) curU += inter_energy(s, &s->inters[q]);
        U0 = curU;
    }
    t->U0 = U0;
    port_tracker_reset(h, k);
}

void port_timestep(void *h, int nsteps) {
    Sys *s = (Sys *)h;
    for (int k = 0; k < nsteps; k++) {
        switch (s->integrator) {
            case 0: verlet_timestep(s); break;
            case 1: sol_timestep(s); break;
            case 2: damped_timestep(s); break;
            case 3: solht_timestep(s); break;
            case 4: overdamped_timestep(s); break;
            case 5: nosehoover_timestep(s); break;
            case 6: gaussiant_timestep(s); break;
            case 7: gear3a_timestep(s); break;
            case 8: gear_timestep(s, 4); break;
            case 9: gear_timestep(s, 5); break;
            case 10: gear_timestep(s, 6); break;
            case 11: nlcg_timestep(s); break;
        }
    }
}

int port_update_list(void *h, int nl, int force) { Sys *s = (Sys *)h; return nl_update_list(s, &s->nls[nl], force); }
/* NeighborList::ignore, trackers.hpp:190-193 (PairList::add_pair :76-83: a set, duplicates collapse) */
void port_ignore(void *h, int nl, const uint32_t *a, const uint32_t *b, uint64_t npairs) {
    NList *l = &((Sys *)h)->nls[nl];
    l->ignored = (uint64_t *)realloc(l->ignored, (l->nignored + npairs + 1) * 8);
    for (uint64_t k = 0; k < npairs; k++) {
        uint64_t key = a[k] > b[k] ? ((uint64_t)a[k] << 32 | b[k]) : ((uint64_t)b[k] << 32 | a[k]);
        l->ignored[l->nignored++] = key;
    }
    qsort(l->ignored, l->nignored, 8, cmp_u64);
    size_t w = 0;
    for (size_t k = 0; k < l->nignored; k++)
        if (w == 0 || l->ignored[w - 1] != l->ignored[k]) l->ignored[w++] = l->ignored[k];
    l->nignored = w;
    l->ignorechanged = 1;
}
uint32_t port_ignore_size(void *h, int nl) { return (uint32_t)((Sys *)h)->nls[nl].nignored; }
uint32_t port_which(void *h, int nl) { return ((Sys *)h)->nls[nl].updatenum; }
uint32_t port_numpairs(void *h, int nl) { return (uint32_t)((Sys *)h)->nls[nl].npairs; }
void port_get_pairs(void *h, int nl, uint32_t *first, uint32_t *last) {
    NList *l = &((Sys *)h)->nls[nl];
    memcpy(first, l->first, l->npairs * 4);
    memcpy(last, l->last, l->npairs * 4);
}
void port_set_atoms(void *h, const double *x, const double *v, const double *a, const double *f) {
    Sys *s = (Sys *)h;
    size_t nb = (size_t)s->n * s->D * 8;
    if (x) memcpy(s->x, x, nb);
    if (v) memcpy(s->v, v, nb);
    if (a) memcpy(s->a, a, nb);
    if (f) memcpy(s->f, f, nb);
}
void port_get_atoms(void *h, double *x, double *v, double *a, double *f) {
    Sys *s = (Sys *)h;
    size_t nb = (size_t)s->n * s->D * 8;
    if (x) memcpy(x, s->x, nb);
    if (v) memcpy(v, s->v, nb);
    if (a) memcpy(a, s->a, nb);
    if (f) memcpy(f, s->f, nb);
}
void port_box_diff(void *h, const double *r1, const double *r2, double *out) { box_diff((Sys *)h, r1, r2, out); }
double port_box_V(void *h) { /* box.hpp:122,127 */
    Sys *s = (Sys *)h;
    return s->D == 3 ? s->L[0] * s->L[1] * s->L[2] : s->L[0] * s->L[1];
}
void port_reset_forces(void *h) { Sys *s = (Sys *)h; memset(s->f, 0, (size_t)s->n * s->D * 8); }
void port_inter_set_forces(void *h, int k) { Sys *s = (Sys *)h; inter_loop(s, &s->inters[k], 1, NULL, NULL); }
double port_inter_set_forces_get_pressure(void *h, int k) { Sys *s = (Sys *)h; double p; inter_loop(s, &s->inters[k], 1, &p, NULL); return p; }
double port_inter_energy(void *h, int k) { Sys *s = (Sys *)h; return inter_energy(s, &s->inters[k]); }
double port_inter_pressure(void *h, int k) { Sys *s = (Sys *)h; double p; inter_loop(s, &s->inters[k], 0, &p, NULL); return p; }
int port_inter_contacts(void *h, int k, unsigned long long *c, unsigned long long *o) { Sys *s = (Sys *)h; inter_contacts(s, &s->inters[k], c, o); return 0; }
void port_inter_stress(void *h, int k, double *out) { Sys *s = (Sys *)h; inter_loop(s, &s->inters[k], 0, NULL, out); }

void port_set_forces(void *h, int constraints_and_a) {
    Sys *s = (Sys *)h;
    if (s->integrator == 11) { /* CollectionNLCG::set_forces(bool) -> set_forces(constraints_and_a, true), collection.hpp:449-451 */
        nlcg_set_forces(s, constraints_and_a, 1);
        return;
    }
    if (s->integrator == 6) { /* CollectionGaussianT::set_forces(bool) -> set_forces(true, true), collection.hpp:618 */
        collection_set_forces(s, 1);
        gaussiant_set_xi(s);
        return;
    }
    collection_set_forces(s, constraints_and_a);
}
double port_nlcg_reduce(void *h, int what);
double port_kinetic_energy(void *h) { /* virtual: CollectionNLCG overrides it, collection.cpp:574-588 */
    double z[3] = {0, 0, 0};
    if (((Sys *)h)->integrator == 11) return port_nlcg_reduce(h, 4);
    return group_ke((Sys *)h, z);
}
double port_potential_energy(void *h) { /* collection.cpp:98-108 */
    Sys *s = (Sys *)h;
    double E = 0;
    for (int k = 0; k < s->ninters; k++) E += inter_energy(s, &s->inters[k]);
    return E;
}
double port_energy(void *h) { return port_potential_energy(h) + port_kinetic_energy(h); } /* :110-114 */
double port_degrees_of_freedom(void *h) { /* :116-133 */
    Sys *s = (Sys *)h;
    int ndof = 0;
    for (uint32_t i = 0; i < s->n; i++)
        if (!frozen_le(s->m[i])) ndof += s->D;
    return ndof;
}
double port_temp(void *h, int minuscomv) { /* :135-142 */
    Sys *s = (Sys *)h;
    double v[3] = {0, 0, 0};
    if (minuscomv) group_com_velocity(s, v);
    int ndof = (int)port_degrees_of_freedom(h);
    if (minuscomv) ndof -= s->D;
    return group_ke(s, v) * 2 / ndof;
}
double port_virial(void *h) { /* :73-80 */
    Sys *s = (Sys *)h;
    double E = 0;
    for (int k = 0; k < s->ninters; k++) {
        double p;
        inter_loop(s, &s->inters[k], 0, &p, NULL);
        E += p;
    }
    return E;
}
double port_pressure(void *h) { /* :85-96 */
    if (((Sys *)h)->integrator == 11) return port_nlcg_reduce(h, 5); /* CollectionNLCG::pressure :590-600 */
    Sys *s = (Sys *)h;
    double V = port_box_V(h);
    double E = 2.0 * port_kinetic_energy(h);
    for (int k = 0; k < s->ninters; k++) {
        double p;
        inter_loop(s, &s->inters[k], 0, &p, NULL);
        E += p;
    }
    return E / V / (double)s->D;
}
void port_com_velocity(void *h, double *out) { double t[3]; group_com_velocity((Sys *)h, t); memcpy(out, t, 8 * ((Sys *)h)->D); }
void port_momentum(void *h, double *out) { double t[3]; group_momentum((Sys *)h, t); memcpy(out, t, 8 * ((Sys *)h)->D); }
double port_mass(void *h) { return group_mass((Sys *)h); }
double port_atoms_kinetic_energy(void *h, const double *v0) { double t[3] = {0, 0, 0}; memcpy(t, v0, 8 * ((Sys *)h)->D); return group_ke((Sys *)h, t); }
void port_add_velocity(void *h, const double *dv) { /* box.cpp:413-417 */
    Sys *s = (Sys *)h;
    for (uint32_t i = 0; i < s->n; i++)
        for (int d = 0; d < s->D; d++) s->v[(size_t)i * s->D + d] += dv[d];
}
void port_reset_com_velocity(void *h) { /* box.hpp:423 */
    Sys *s = (Sys *)h;
    double c[3];
    group_com_velocity(s, c);
    for (int d = 0; d < s->D; d++) c[d] = -c[d];
    port_add_velocity(h, c);
}
void port_scale_velocities(void *h, double scaleby) { /* collection.cpp:21-29 */
    Sys *s = (Sys *)h;
    for (uint32_t i = 0; i < s->n; i++) {
        if (frozen_le(s->m[i])) continue;
        for (int d = 0; d < s->D; d++) s->v[(size_t)i * s->D + d] *= scaleby;
    }
}
void port_scale_velocities_to_temp(void *h, double T, int minuscomv) { /* :31-35 */
    double t = port_temp(h, minuscomv);
    port_scale_velocities(h, sqrt(T / t));
}
void port_scale_velocities_to_energy(void *h, double E) { /* :37-43 */
    double E0 = port_energy(h);
    double k0 = port_kinetic_energy(h);
    double goalkinetic = k0 + (E - E0);
    port_scale_velocities(h, sqrt(goalkinetic / k0));
}
