/* rr_score.h -- the significance formula of the MaxCorrelation scan, written once for
 * host and device.
 *
 * Replaces, for one (group i, group j) pair whose four intersection counts are already
 * known, the arithmetic of
 *     PositiveSignificance      /root/reference/MaxCorrelation.c:421-434
 *     PositiveCumHypGeo_Log     /root/reference/MaxCorrelation.c:413-419
 *     F_beta(.,.,1.0)           /root/reference/MaxCorrelation.c:396-411
 * and of the GSL routine the reference calls at MaxCorrelation.c:415,
 *     gsl_cdf_hypergeometric_Q(schnitt-1, gr2, cov-gr2, gr1)
 * (GSL 2.x: cdf/hypergeometric.c -> randist/hypergeometric.c -> specfunc/gamma.c; GSL is a
 * link-time dependency of the reference, not vendored; algorithm restated from its
 * published sources).
 *
 * Exactness contract: every +,-,*,/ of the exact path is performed in IEEE double in the
 * reference's operation order with no fused multiply-add (RR_ADD/RR_MUL/RR_DIV map to
 * __dadd_rn/__dmul_rn/__ddiv_rn on the device), and ln(n!) comes from a host-built table
 * (rr_lnfact.c) so the only device/host difference is exp() and log10() (<= 2 ulp).
 *
 * The *bound* functions are this implementation's own (they have no reference
 * counterpart): rigorous upper bounds on the score that let the scan skip pairs which
 * cannot raise either running maximum.
 */
#ifndef RR_SCORE_H
#define RR_SCORE_H

#include <math.h>
#include <float.h>

#if defined(__CUDA_ARCH__)
#define RR_HD __host__ __device__ __forceinline__
#define RR_ADD(a, b) __dadd_rn((a), (b))
#define RR_SUB(a, b) __dsub_rn((a), (b))
#define RR_MUL(a, b) __dmul_rn((a), (b))
#define RR_DIV(a, b) __ddiv_rn((a), (b))
#define RR_LDG(p) __ldg(p)
#elif defined(__CUDACC__)
#define RR_HD __host__ __device__ __forceinline__
#define RR_ADD(a, b) ((a) + (b))
#define RR_SUB(a, b) ((a) - (b))
#define RR_MUL(a, b) ((a) * (b))
#define RR_DIV(a, b) ((a) / (b))
#define RR_LDG(p) (*(p))
#else
#define RR_HD static inline
#define RR_ADD(a, b) ((a) + (b))
#define RR_SUB(a, b) ((a) - (b))
#define RR_MUL(a, b) ((a) * (b))
#define RR_DIV(a, b) ((a) / (b))
#define RR_LDG(p) (*(p))
#endif

#define RR_LOG10E 0.43429448190325182765

/* gsl_sf_lnchoose(n, m) on the ln(n!) table: 0 for m==n or m==0, else with m folded to
 * min(m, n-m):  lnfact(n) - lnfact(m) - lnfact(n-m)   (this association order). */
RR_HD double rr_lnchoose(const double *lnf, unsigned int n, unsigned int m)
{
    if (m == n || m == 0) return 0.0;
    if (m * 2 > n) m = n - m;
    return RR_SUB(RR_SUB(RR_LDG(lnf + n), RR_LDG(lnf + m)), RR_LDG(lnf + (n - m)));
}

/* gsl_ran_hypergeometric_pdf(k, n1, n2, t) */
RR_HD double rr_hyper_lnpdf_unchecked(const double *lnf, unsigned int k, unsigned int n1, unsigned int n2,
                                      unsigned int t)
{
    double c1 = rr_lnchoose(lnf, n1, k);
    double c2 = rr_lnchoose(lnf, n2, t - k);
    double c3 = rr_lnchoose(lnf, n1 + n2, t);
    return RR_SUB(RR_ADD(c1, c2), c3);
}

RR_HD double rr_hyper_pdf(const double *lnf, unsigned int k, unsigned int n1, unsigned int n2, unsigned int t)
{
    if (t > n1 + n2) t = n1 + n2;
    if (k > n1 || k > t) return 0.0;
    if (t > n2 && k + n2 < t) return 0.0;
    return exp(rr_hyper_lnpdf_unchecked(lnf, k, n1, n2, t));
}

/* gsl_cdf_hypergeometric_Q(k, n1, n2, t), including GSL's mixed int/unsigned/double
 * sub-expressions in lower_tail / upper_tail. */
RR_HD double rr_hyper_Q(const double *lnf, unsigned int k, unsigned int n1, unsigned int n2, unsigned int t)
{
    double midpoint;
    if (k >= n1 || k >= t) return 0.0;
    midpoint = RR_DIV(RR_MUL((double)t, (double)n1), RR_ADD((double)n1, (double)n2));
    if ((double)k < midpoint) {
        /* Q = 1 - lower_tail(k, n1, n2, t) */
        int i = (int)k;
        double s = rr_hyper_pdf(lnf, (unsigned int)i, n1, n2, t);
        double P = s;
        if (s == 0.0) return 1.0; /* every later term is 0*factor = 0: P stays 0 */
        while (i > 0) {
            double f1 = RR_DIV((double)i, RR_ADD((double)(n1 - (unsigned int)i), 1.0));
            double f2 = RR_DIV((double)(n2 + (unsigned int)i - t), RR_ADD((double)(t - (unsigned int)i), 1.0));
            double factor = RR_MUL(f1, f2);
            double relerr;
            s = RR_MUL(s, factor);
            P = RR_ADD(P, s);
            relerr = RR_DIV(s, P);
            if (relerr < DBL_EPSILON) break;
            i--;
        }
        return RR_SUB(1.0, P);
    } else {
        /* Q = upper_tail(k, n1, n2, t) */
        unsigned int i = k + 1;
        double s = rr_hyper_pdf(lnf, i, n1, n2, t);
        double Q = s;
        if (s == 0.0) return 0.0; /* pdf underflow: the series stays 0 (GSL loops to i==t) */
        while (i < t) {
            double f1 = RR_DIV((double)(n1 - i), RR_ADD((double)i, 1.0));
            double f2 = RR_DIV((double)(t - i), RR_SUB(RR_ADD((double)(n2 + i), 1.0), (double)t));
            double factor = RR_MUL(f1, f2);
            double relerr;
            s = RR_MUL(s, factor);
            Q = RR_ADD(Q, s);
            relerr = RR_DIV(s, Q);
            if (relerr < DBL_EPSILON) break;
            i++;
        }
        return Q;
    }
}

/* 98.0 + F_beta(Gi, Gj, 1.0): all operands are exact integers in doubles, so
 * (1+1)*s / ((1+1*1)*s + |Gi\Gj| + |Gj\Gi|) needs no bitset pass. */
RR_HD double rr_saturated(unsigned int s, int sizei, int sizej)
{
    double sd = (double)s;
    double F = RR_MUL(2.0, sd);
    double den;
    if (F < 0.0001) return 98.0;
    den = RR_ADD(RR_ADD(RR_MUL(2.0, sd), (double)(sizei - (int)s)), (double)(sizej - (int)s));
    return RR_ADD(98.0, RR_DIV(F, den));
}

/* PositiveSignificance on counts: s=|Gi&Gj|, gr1=|Gi&Cj|, gr2=|Gj&Ci|, cov=|Ci&Cj|. */
RR_HD double rr_positive_significance(const double *lnf, unsigned int s, unsigned int gr1, unsigned int gr2,
                                      unsigned int cov, int sizei, int sizej)
{
    double Q, Z;
    if (gr1 == 0 || gr2 == 0) return 0.0;
    if (s < 1) return 0.0;
    Q = rr_hyper_Q(lnf, s - 1, gr2, cov - gr2, gr1);
    Z = RR_MUL(-1.0, log10(Q));
    if (isinf(Z) || Z > 99) Z = 99.0;
    if (isinf(Z) || Z > 98.0) Z = rr_saturated(s, sizei, sizej);
    return Z;
}

/* Group_PositiveSignificance (/root/reference/RepeatResolver.c:472-488) on counts: the same hypergeometric score,
 * without the s < 1 brake (its caller Cliquer only asks for s > mincov/4, 1215) and saturating at 97.90 + F1 (486). */
RR_HD double rr_group_significance(const double *lnf, unsigned int s, unsigned int gr1, unsigned int gr2,
                                   unsigned int cov, int sizei, int sizej)
{
    double Q, Z;
    if (gr1 == 0 || gr2 == 0) return 0.0;
    Q = rr_hyper_Q(lnf, s - 1, gr2, cov - gr2, gr1);
    Z = RR_MUL(-1.0, log10(Q));
    if (isinf(Z) || Z > 99) Z = 99.0;
    if (isinf(Z) || Z > 98.0) {
        const double sd = (double)s, F = RR_MUL(2.0, sd);
        if (F < 0.0001) return RR_ADD(97.90, 0.0);
        Z = RR_ADD(97.90, RR_DIV(F, RR_ADD(RR_ADD(RR_MUL(2.0, sd), (double)(sizei - (int)s)), (double)(sizej - (int)s))));
    }
    return Z;
}

/* ---- Relative_Group_Significance (/root/reference/RepeatResolver.c:506-523), the two-sided score of Relative_Vars ----
 * CumHypGeo_Log (490-504) needs gsl_cdf_hypergeometric_P beside Q.  The two tail series are written out again here
 * rather than shared with rr_hyper_Q above, so that the code the scan kernels compile is untouched. */
RR_HD double rr_hyper_lower_tail(const double *lnf, unsigned int k, unsigned int n1, unsigned int n2, unsigned int t)
{
    int i = (int)k;
    double s = rr_hyper_pdf(lnf, (unsigned int)i, n1, n2, t);
    double P = s;
    if (s == 0.0) return 0.0; /* every later term is 0*factor = 0 */
    while (i > 0) {
        double f1 = RR_DIV((double)i, RR_ADD((double)(n1 - (unsigned int)i), 1.0));
        double f2 = RR_DIV((double)(n2 + (unsigned int)i - t), RR_ADD((double)(t - (unsigned int)i), 1.0));
        double relerr;
        s = RR_MUL(s, RR_MUL(f1, f2));
        P = RR_ADD(P, s);
        relerr = RR_DIV(s, P);
        if (relerr < DBL_EPSILON) break;
        i--;
    }
    return P;
}

RR_HD double rr_hyper_upper_tail(const double *lnf, unsigned int k, unsigned int n1, unsigned int n2, unsigned int t)
{
    unsigned int i = k + 1;
    double s = rr_hyper_pdf(lnf, i, n1, n2, t);
    double Q = s;
    if (s == 0.0) return 0.0;
    while (i < t) {
        double f1 = RR_DIV((double)(n1 - i), RR_ADD((double)i, 1.0));
        double f2 = RR_DIV((double)(t - i), RR_SUB(RR_ADD((double)(n2 + i), 1.0), (double)t));
        double relerr;
        s = RR_MUL(s, RR_MUL(f1, f2));
        Q = RR_ADD(Q, s);
        relerr = RR_DIV(s, Q);
        if (relerr < DBL_EPSILON) break;
        i++;
    }
    return Q;
}

/* gsl_cdf_hypergeometric_P(k, n1, n2, t) */
RR_HD double rr_hyper_P(const double *lnf, unsigned int k, unsigned int n1, unsigned int n2, unsigned int t)
{
    double midpoint;
    if (k >= n1 || k >= t) return 1.0;
    midpoint = RR_DIV(RR_MUL((double)t, (double)n1), RR_ADD((double)n1, (double)n2));
    if ((double)k >= midpoint) return RR_SUB(1.0, rr_hyper_upper_tail(lnf, k, n1, n2, t));
    return rr_hyper_lower_tail(lnf, k, n1, n2, t);
}

/* the same Q as rr_hyper_Q, on the two series above (kept separate, see the note) */
RR_HD double rr_hyper_Q2(const double *lnf, unsigned int k, unsigned int n1, unsigned int n2, unsigned int t)
{
    double midpoint;
    if (k >= n1 || k >= t) return 0.0;
    midpoint = RR_DIV(RR_MUL((double)t, (double)n1), RR_ADD((double)n1, (double)n2));
    if ((double)k < midpoint) return RR_SUB(1.0, rr_hyper_lower_tail(lnf, k, n1, n2, t));
    return rr_hyper_upper_tail(lnf, k, n1, n2, t);
}

/* Relative_Group_Significance on counts: schnitt = |G1 & G2 & U|, gr1 = |G1 & U|, gr2 = |G2 & U|, cov = |U| */
RR_HD double rr_relative_significance(const double *lnf, unsigned int s, unsigned int gr1, unsigned int gr2, unsigned int cov)
{
    double P, Q, Z;
    if (gr1 == 0 || gr2 == 0) return 0.0;                                /* 517 */
    P = rr_hyper_P(lnf, s, gr2, cov - gr2, gr1);                         /* 492 */
    Q = rr_hyper_Q2(lnf, s - 1u, gr2, cov - gr2, gr1);                   /* 493: s - 1 wraps for 0, Q = 0 then */
    Z = RR_MUL(-1.0, log10((P < Q || s == 0) ? P : Q));                  /* 495-503 */
    if (isinf(Z) || Z > 99) Z = 99.0;
    return Z;
}

/* ---- pruning bounds (no reference counterpart) -------------------------------------
 * With X ~ Hypergeom(pop = cov, successes = gr2, draws = gr1) the score before the caps
 * is -log10 P[X >= s].  Two rigorous upper bounds:
 *  (1) a median of the hypergeometric lies within 1 of its mean, so if
 *      s <= mean - 2 then P[X >= s] >= 1/2 and the score is <= log10 2;
 *  (2) P[X >= s] >= pmf(x) for every x >= s inside the support, so the score is
 *      <= -log10 pmf(x); x = s when s is above the mean, else a point next to the mean
 *      (any admissible x is valid, so the mean may be computed approximately).
 * Bound (2) is compared with a 1e-6 margin, far above the 2.5e-11 noise of the lnfact
 * differences.  Saturation: a raw score above 98 is REPLACED by 98 + F in (98, 99], which can
 * exceed the raw score (raw 98.4 -> e.g. 98.95), so a bound above 98 proves nothing below 99:
 * rr_bound_effective() maps it to 99 and such pairs are never pruned.
 */
#define RR_SATURATION_START 98.0
#define RR_SCORE_MAX 99.0
#define RR_BOUND_MEDIAN 0.30103001

RR_HD int rr_below_median(unsigned int s, unsigned int gr1, unsigned int gr2, unsigned int cov)
{
    return (unsigned long long)(s + 2u) * cov <= (unsigned long long)gr1 * gr2;
}

RR_HD double rr_score_upper_bound(const double *lnf, unsigned int s, unsigned int gr1, unsigned int gr2,
                                  unsigned int cov)
{
    /* support of X: max(0, gr1+gr2-cov) <= x <= min(gr1, gr2) */
    unsigned int hi = gr1 < gr2 ? gr1 : gr2;
    unsigned int x = (unsigned int)(((float)gr1 * (float)gr2) / (float)cov) + 1u;
    double lp;
    if (x < s) x = s;
    if (x > hi) x = hi;
    if (x + cov < gr1 + gr2) x = gr1 + gr2 - cov;
    lp = rr_hyper_lnpdf_unchecked(lnf, x, gr2, cov - gr2, gr1);
    return -RR_LOG10E * lp + 1e-6;
}

/* what a bound on the raw score says about the final score */
RR_HD double rr_bound_effective(double raw_bound)
{
    return raw_bound > RR_SATURATION_START ? RR_SCORE_MAX + 1.0 : raw_bound;
}

#endif /* RR_SCORE_H */
