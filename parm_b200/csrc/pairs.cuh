// Pair functors of the reference's NListed<A,P> family (interaction.hpp), as two pieces each:
//   mix_pair<KIND>  -- the pair CONSTRUCTOR (resolves eps_ij, sigma_ij, cut, cut_energy ... from the two
//                      per-atom A structs); __host__ __device__: evaluated once per species pair on the host,
//                      or per pair on the device when an interaction has more than PARM_MAX_SPECIES
//                      distinct parameter tuples (continuous polydispersity)
//   pair_eval<KIND> -- P::forces(box) / P::energy(box) on the device: f = rij * scal, energy e.
// Per-atom parameters p[0..4] follow the table in include/parm_b200.h (parm_inter_set_params_ex).
#pragma once
#include <math.h>

#include "internal.cuh"

// "Min" variants differ from their base functor only in the constructor: they share its kernels
#define PARM_KERNEL_KIND(kind) ((kind) == PARM_PAIR_LOISOHERNMIN ? PARM_PAIR_LOISOHERN : (kind) == PARM_PAIR_LOISLINMIN ? PARM_PAIR_LOISLIN : (kind))
#define PARM_NKERNEL_KINDS 12

// which per-atom parameters mix geometrically (sqrt(p1 p2)): on the device path these are stored as sqrt(p) so
// that the per-pair constructor is a multiply (identical to ~1 ulp)
__host__ __device__ constexpr unsigned parm_geo_mask(int kind, bool eps_indexed) {
    switch (kind) {
        case PARM_PAIR_LJATTRACTREPULSE: return 0u;
        case PARM_PAIR_LJATTRACTFIXEDREPULSE:
        case PARM_PAIR_LJISH: return 1u << 3;
        case PARM_PAIR_EISMCLACHLAN: return 0u;
        case PARM_PAIR_LJATTRACTREPULSESIGS: return (1u << 0) | (1u << 3);
        default: return eps_indexed ? 0u : 1u;
    }
}
__host__ __device__ constexpr int parm_nparams(int kind) {
    switch (kind) {
        case PARM_PAIR_LJREPULSE:
        case PARM_PAIR_EISMCLACHLAN: return 2;
        case PARM_PAIR_REPULSION:
        case PARM_PAIR_LJATTRACTREPULSE:
        case PARM_PAIR_LJCUT:
        case PARM_PAIR_LJATTRACTCUT: return 3;
        case PARM_PAIR_LJISH:
        case PARM_PAIR_LJATTRACTREPULSESIGS: return 5;
        default: return 4;
    }
}

// geometric mean of two per-atom parameters: the device copies hold sqrt(p)
__host__ __device__ inline double geo_mean(double x, double y) {
#ifdef __CUDA_ARCH__
    return x * y;
#else
    return sqrt(x * y);
#endif
}

// Tables: eps_tab / sig_tab are ntypes x ntypes (row = this atom's `epsilons` / `sigmas` vector, column = the
// other atom's indx) or NULL when that quantity is not indexed.
template <int KIND>
__host__ __device__ inline PairConst mix_pair(const double *p1, int t1, const double *p2, int t2, const double *eps_tab,
                                              const double *sig_tab, int nt, bool minmix, bool want_e) {
    PairConst P;
    P.eps = P.sig = P.sig2 = P.rc2 = P.cutE = P.a = P.b = P.c = 0.0;
    const double sig_mean = sig_tab ? sig_tab[t1 * nt + t2] : (p1[1] + p2[1]) / 2.0;
    if (KIND == PARM_PAIR_LJREPULSE) { // LJRepulsePair ctor :878-883
        P.eps = geo_mean(p1[0], p2[0]);
        P.sig = sig_mean;
        P.sig2 = P.sig * P.sig;
        P.rc2 = P.sig2;
    } else if (KIND == PARM_PAIR_REPULSION || KIND == PARM_PAIR_REPULSIONDRAG) { // :1531-1542, :1617-1623
        P.eps = eps_tab ? eps_tab[t1 * nt + t2] : geo_mean(p1[0], p2[0]);
        P.sig = sig_mean;
        P.sig2 = P.sig * P.sig;
        P.rc2 = P.sig2;
        P.a = (p1[2] + p2[2]) / 2.0;
        if (KIND == PARM_PAIR_REPULSIONDRAG) P.b = (p1[3] + p2[3]) / 2;
    } else if (KIND == PARM_PAIR_LJATTRACTREPULSE) { // :1255-1270
        double eps = eps_tab[t1 * nt + t2];
        double cut = fmax(p1[2], p2[2]);
        P.sig = sig_mean;
        if (eps <= 0) {
            cut = 1;
            eps = fabs(eps);
        } else if (want_e) {
            double mid = (1 - pow(cut, -6.0));
            P.cutE = eps * (mid * mid);
        }
        P.eps = eps;
        P.sig2 = P.sig * P.sig;
        P.rc2 = cut * cut * P.sig2;
    } else if (KIND == PARM_PAIR_LJCUT) { // LennardJonesCutPair ctors :970-979 + LennardJonesCut ctor :247-252
        P.eps = eps_tab ? eps_tab[t1 * nt + t2] : geo_mean(p1[0], p2[0]);
        P.sig = sig_mean;
        double cut = fmax(p1[2], p2[2]);
        if (want_e) {
            double rsix = pow(cut, 6.0);
            double mid = (1 - 1 / rsix);
            P.cutE = P.eps * (mid * mid - 1);
        }
        P.sig2 = P.sig * P.sig;
        P.rc2 = cut * cut * P.sig2;
    } else if (KIND == PARM_PAIR_LJATTRACTCUT) { // LJAttractCutPair ctors :1023-1040 + LJAttractCut ctor :205-209
        P.eps = eps_tab ? eps_tab[t1 * nt + t2] : geo_mean(p1[0], p2[0]);
        P.sig = sig_mean;
        double cut = fmax(p1[2], p2[2]);
        if (want_e) { // LJAttract::energy(rsig) :170-177, times epsilon
            double le = -1;
            if (!(cut < 1)) {
                double rsq = cut * cut;
                double rsix = rsq * rsq * rsq;
                double mid = (1 - 1 / rsix);
                le = mid * mid - 1;
            }
            P.cutE = le * P.eps;
        }
        P.sig2 = P.sig * P.sig;
        P.rc2 = cut * cut * P.sig2;
    } else if (KIND == PARM_PAIR_LJATTRACTFIXEDREPULSE) { // :1347-1364
        double e12 = eps_tab[t1 * nt + t2];
        P.eps = fabs(e12);
        P.a = geo_mean(p1[3], p2[3]); // repeps
        P.sig = sig_mean;
        double cut = fmax(p1[2], p2[2]);
        bool attract = e12 > 0;
        if (!attract || P.eps == 0) {
            cut = 1;
            P.eps = 0;
        } else if (want_e) {
            double mid = (1 - pow(cut, -6.0));
            P.cutE = P.eps * (mid * mid);
        }
        P.sig2 = P.sig * P.sig;
        P.rc2 = cut * cut * P.sig2;
    } else if (KIND == PARM_PAIR_EISMCLACHLAN) { // :1430-1439; p = (sigmai, dist)
        const double s1 = p1[0], d1 = p1[1], s2 = p2[0], d2 = p2[1];
        P.eps = -M_PI * (s1 * d2 - s2 * d1) * (d1 * d1 - d2 * d2); // c0
        P.a = -2 * M_PI * (s1 * d2 * d2 + s2 * d1 * d1);           // c1
        P.b = M_PI * (s1 * d2 + s2 * d1);                          // c2
        P.sig = d1 + d2;                                           // cutoff
        P.sig2 = P.sig * P.sig;
        P.rc2 = P.sig2;
    } else if (KIND == PARM_PAIR_LJISH) { // :1098-1114
        P.eps = eps_tab[t1 * nt + t2];
        P.a = geo_mean(p1[3], p2[3]); // repeps
        P.sig = sig_mean;
        P.b = (p1[4] + p2[4]) / 2; // n
        double cut = fmax(p1[2], p2[2]);
        if (P.eps <= 0) {
            cut = 1;
            P.eps = 0;
        } else if (want_e) {
            double mid = (1 - pow(cut, -P.b));
            P.cutE = P.eps * (mid * mid);
        }
        P.sig2 = P.sig * P.sig;
        P.rc2 = cut * cut * P.sig2;
    } else if (KIND == PARM_PAIR_LJATTRACTREPULSESIGS) { // :1175-1193; p = (eps_r, sig_r, sigcut, eps_a, sig_a)
        P.eps = geo_mean(p1[0], p2[0]);  // eps_r
        P.a = geo_mean(p1[3], p2[3]);    // eps_a
        P.sig = (p1[1] + p2[1]) / 2.0;   // sig_r
        P.b = (p1[4] + p2[4]) / 2.0;     // sig_a
        P.sig2 = P.sig * P.sig;
        const double cut = fmax(p1[2], p2[2]);
        if (cut <= 0) { // no cutoff; the reference leaves cut_energy uninitialised here, we define it as 0
            P.c = INFINITY;
            P.rc2 = INFINITY;
        } else {
            const double cdu = P.sig + P.b * (cut - 1); // cut_distance_units
            P.c = cdu;
            P.rc2 = cdu * cdu;
            if (want_e) {
                if (cut >= 1) { // attract_energy(cdu) :1195-1199
                    double ros = (cdu - P.sig + P.b) / P.b;
                    double mid = (1 - pow(ros, -6.0));
                    P.cutE = P.a * (mid * mid) - P.a;
                } else { // repulse_energy(cdu) :1201-1205
                    double ros = cdu / P.sig;
                    double mid = (1 - pow(ros, -6.0));
                    P.cutE = P.eps * (mid * mid) - P.a;
                }
            }
        }
    } else if (KIND == PARM_PAIR_LOISOHERN) { // :1696-1704 and LoisOhernPairMinCLs :1747-1751; p = (eps, sigma, C, l)
        P.eps = geo_mean(p1[0], p2[0]);
        P.sig = (p1[1] + p2[1]) / 2.0;
        P.a = minmix ? (p1[2] < p2[2] ? p1[2] : p2[2]) : (p1[2] + p2[2]) / 2.0; // C
        P.b = minmix ? (p1[3] < p2[3] ? p1[3] : p2[3]) : (p1[3] + p2[3]) / 2.0; // l
        P.c = P.sig * (1 + P.a + P.b);                                          // sigcut
        P.sig2 = P.sig * P.sig;
        P.rc2 = P.c * P.c;
    } else if (KIND == PARM_PAIR_LOISLIN) { // :1785-1793 and LoisLinPairMin :1832-1836; p = (eps, sigma, f, l)
        P.eps = geo_mean(p1[0], p2[0]);
        P.sig = (p1[1] + p2[1]) / 2.0;
        P.a = minmix ? (p1[2] < p2[2] ? p1[2] : p2[2]) : (p1[2] + p2[2]) / 2.0; // f
        P.b = minmix ? (p1[3] < p2[3] ? p1[3] : p2[3]) : (p1[3] + p2[3]) / 2.0; // l
        P.c = P.sig + P.b;                                                      // sigcut
        P.sig2 = P.sig * P.sig;
        P.rc2 = P.c * P.c;
    }
    return P;
}

#ifdef __CUDACC__
// pow(t, n-1) and pow(t, n) for the Hertzian family; n = 2 and 2.5 avoid the generic pow
__device__ __forceinline__ void pow_pair(double t, double n, double &pm1, double &p0) {
    if (n == 2.0) {
        pm1 = t;
    } else if (n == 2.5) {
        pm1 = t * sqrt(t);
    } else {
        pm1 = pow(t, n - 1.0);
    }
    p0 = pm1 * t;
}

// f = rij * scal, pair energy e (only when want_e). vdotr = (v1 - v2).rij, used by the drag functor only.
template <int KIND>
__device__ __forceinline__ void pair_eval(const PairConst &P, double dsq, double vdotr, bool want_e, double &scal, double &e) {
    scal = 0.0;
    e = 0.0;
    if (KIND == PARM_PAIR_REPULSION || KIND == PARM_PAIR_REPULSIONDRAG) {
        // RepulsionPair::forces / energy :1537-1550; RepulsionDragPair :1624-1641
        if (dsq > P.sig2) return;
        double R = sqrt(dsq);
        double t = 1.0 - R / P.sig;
        double pm1, p0;
        pow_pair(t, P.a, pm1, p0);
        scal = P.eps * pm1 / P.sig / R;
        if (KIND == PARM_PAIR_REPULSIONDRAG) scal -= P.b * vdotr / dsq; // - v_perp * gamma, v_perp = rij (vij.rij)/dsq
        if (want_e) e = P.eps * p0 / P.a;
    } else if (KIND == PARM_PAIR_LJREPULSE || KIND == PARM_PAIR_LJATTRACTREPULSE || KIND == PARM_PAIR_LJCUT) {
        // rsq = dsq/(sig*sig); if (rsq > cut*cut) -> 0; rsix = sigma^6/r^6
        // f = rij * (12 eps rsix (rsix - 1) / dsq)      (:135-151, :259-267, :1289-1298)
        if (dsq > P.rc2) return;
        double w = 1.0 / dsq;
        double s2 = P.sig2 * w;
        double ir6 = s2 * s2 * s2;
        scal = 12.0 * P.eps * ir6 * (ir6 - 1.0) * w;
        if (want_e) {
            double mid = 1.0 - ir6;
            if (KIND == PARM_PAIR_LJCUT)
                e = P.eps * (mid * mid - 1.0) - P.cutE; // :253-258
            else
                e = P.eps * (mid * mid) - P.cutE;       // :126-133 (cutE = 0), :1271-1288
        }
    } else if (KIND == PARM_PAIR_LJATTRACTCUT) {
        // LJAttractCut::forces :224-232, ::energy :216-221 + LJAttract::energy :161-168
        if (P.eps == 0 || dsq > P.rc2) return;
        if (dsq < P.sig2) { // inside the minimum: flat bottom
            if (want_e) e = -P.eps - P.cutE;
            return;
        }
        double w = 1.0 / dsq;
        double s2 = P.sig2 * w;
        double ir6 = s2 * s2 * s2;
        scal = 12.0 * P.eps * ir6 * (ir6 - 1.0) * w;
        if (want_e) {
            double mid = 1.0 - ir6;
            e = P.eps * (mid * mid - 1.0) - P.cutE;
        }
    } else if (KIND == PARM_PAIR_LJATTRACTFIXEDREPULSE) {
        // :1365-1404: eps beyond sigma, repeps inside
        if (dsq > P.rc2) return;
        double w = 1.0 / dsq;
        double s2 = P.sig2 * w;
        double ir6 = s2 * s2 * s2;
        double fm = 12.0 * ir6 * (ir6 - 1.0);
        scal = (dsq < P.sig2 ? P.a : P.eps) * fm * w;
        if (want_e) {
            double mid = 1.0 - ir6;
            e = (dsq > P.sig2 ? P.eps : P.a) * (mid * mid) - P.cutE;
        }
    } else if (KIND == PARM_PAIR_EISMCLACHLAN) {
        // :1440-1452: E = c0/R + c1 + c2 R inside r1 + r2
        if (dsq > P.rc2) return;
        double R = sqrt(dsq);
        scal = (P.eps / dsq - P.b) / R;
        if (want_e) e = P.eps / R + P.a + P.b * R;
    } else if (KIND == PARM_PAIR_LJISH) {
        // :1115-1141: exponent n instead of 6
        if (dsq > P.rc2) return;
        double rsq = dsq / P.sig2;
        double rmid = pow(rsq, -P.b / 2); // sigma^n / r^n
        double fm = 2 * P.b * rmid * (rmid - 1);
        scal = (rsq < 1 ? P.a : P.eps) * fm / dsq;
        if (want_e) {
            double mid = 1 - rmid;
            e = (rsq > 1 ? P.eps : P.a) * (mid * mid) - P.cutE;
        }
    } else if (KIND == PARM_PAIR_LJATTRACTREPULSESIGS) {
        // :1207-1243: LJ-repulsive core of width sig_r, attractive shell measured in units of sig_a
        if (dsq > P.rc2) return;
        if (dsq > P.sig2) {
            double dist = sqrt(dsq);
            double rminus = dist - P.sig + P.b;
            double ros = rminus / P.b;
            double r2 = 1.0 / (ros * ros);
            double rsix = r2 * r2 * r2; // pow(r_over_sig, -6)
            scal = P.a * (12 * rsix * (rsix - 1) / (dist * rminus));
            if (want_e) {
                double mid = 1 - rsix;
                e = (P.a * (mid * mid) - P.a) - P.cutE;
            }
        } else {
            double s2 = P.sig2 / dsq;
            double rsix = s2 * s2 * s2;
            scal = P.eps * (12 * rsix * (rsix - 1)) / dsq;
            if (want_e) {
                double mid = 1 - rsix;
                e = (P.eps * (mid * mid) - P.a) - P.cutE;
            }
        }
    } else if (KIND == PARM_PAIR_LOISOHERN) {
        // :1714-1743: harmonic to (1+C) sigma, then a linear-force rounding of width l sigma
        if (dsq >= P.rc2) return;
        double R = sqrt(dsq);
        double rsig = R / P.sig;
        if (rsig <= 1 + P.a) {
            double dR = rsig - 1;
            scal = -P.eps * dR / R;
            if (want_e) e = -P.eps * P.sig / 2 * (P.a * (P.a + P.b) - dR * dR);
        } else {
            double dR2 = rsig - (P.a + P.b + 1);
            scal = P.a * P.eps / P.b * dR2 / R;
            if (want_e) e = -P.a * P.eps * P.sig / 2 / P.b * dR2 * dR2;
        }
    } else if (KIND == PARM_PAIR_LOISLIN) {
        // :1802-1828: harmonic repulsion inside sigma, constant attractive force f out to sigma + l
        if (dsq >= P.rc2) return;
        double R = sqrt(dsq);
        if (R <= P.sig) {
            double dR = 1.0 - (R / P.sig);
            scal = P.eps * dR / P.sig / R;
            if (want_e) e = P.eps / 2 * dR * dR - P.a * P.b;
        } else {
            scal = -P.a / R;
            if (want_e) e = -P.a * (P.sig + P.b - R);
        }
    }
}
#endif
