// rr_device.cuh -- device-side state and the fused significance epilogue shared by both
// count kernels (bitset AND+POPC and tcgen05 int8).
//
// Replaces the inner body of HilfsMaxCorrsRechner (/root/reference/MaxCorrelation.c:814-824):
// for one (row group i, column group j) pair whose counts are on chip, apply the size
// filters, evaluate PositiveSignificance (421-434, via rr_score.h) and fold the result into
// the running maximum of both groups.  The count matrix never reaches HBM.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "rr_score.h"

// running maximum of one group: value bits (non-negative doubles order like their u64
// patterns) + partner group id.  Updated with one 128-bit CAS so that (value, smallest
// partner) is exact and deterministic regardless of scheduling.
struct __align__(16) rr_best_t {
    unsigned long long z;  // __double_as_longlong(max), 0 = 0.0
    unsigned long long p;  // partner group id, ~0ull = none
};

struct rr_cand_p;

// everything a scan kernel needs, passed by value (kernel parameter)
struct rr_scan_params {
    int R, N;                    // rows kept, columns
    int W32;                     // u32 words per bitset (multiple of 4)
    int mincov;
    unsigned flags;
    const uint32_t *bits;        // [5N][W32] group bitsets in sorted-row order
    const int32_t *gsize;        // [5N] Groupsizearray
    const uint8_t *rowok;        // [5N] admissible as row group  (MaxCorrelation.c:802)
    const uint8_t *colok;        // [5N] admissible as column group (817)
    const int32_t *breakcol;     // [N]  first column the jj sweep of site ii does not reach (807-810)
    const int32_t *rowsites;     // [n_rowsites padded] sites with >=1 admissible row group, ascending; -1 pad
    int n_rowsites;
    const double *lnfact;        // [R+2] ln(n!)
    rr_best_t *best;             // [5N]
    unsigned long long *counters; // [8]: pair tests, exact evals, bound evals, work units, tier-2 evals
    // work decomposition (bitset kernel): row blocks x column blocks
    const int64_t *unit_prefix;  // [n_rowblocks+1] prefix sum of column-block counts
    const int32_t *unit_cb0;     // [n_rowblocks] first column block of each row block
    int n_rowblocks;
    int rb_lo, rb_hi;            // this part's row blocks
    const int32_t *word_hi;      // [n_classes][n_rowblocks] exclusive upper u32-word bound of contributing rows, per
    const int32_t *word_lo;      // [n_classes][n_colblocks] inclusive lower u32-word bound      length class (rr_plan.h)
    int n_colblocks;
    int n_classes;
    // deferred exact evaluation (tcgen05 kernel): what survives tier 2 raises the groups' maxima by a rigorous LOWER bound
    // of its score (a partner-less entry of best[], like an exchanged threshold) and is appended here; rr_k_deferred_exact
    // evaluates the list after the scan.  defer_mode 0: evaluate in place (RR_FLAG_NO_PRUNE, no list); 1: raise + append
    // (in place once the list is full); 2: raise only (seeding passes: their pairs come again in the full pass)
    rr_cand_p *deferred;
    unsigned long long deferred_cap;
    int defer_mode;
};

#ifndef RR_CPU_EMU
__device__ __forceinline__ rr_best_t rr_best_load(const rr_best_t *p)
{
    rr_best_t v;
    asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(v.z), "=l"(v.p) : "l"(p));
    return v;
}

__device__ __forceinline__ double rr_best_value(const rr_best_t *p)
{
    unsigned long long z;
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(z) : "l"(p));
    return __longlong_as_double((long long)z);
}

__device__ __forceinline__ bool rr_cas128(rr_best_t *addr, rr_best_t expected, rr_best_t desired, rr_best_t *old)
{
    unsigned long long o0, o1;
    asm volatile(
        "{\n\t.reg .b128 e, d, o;\n\t"
        "mov.b128 e, {%2, %3};\n\t"
        "mov.b128 d, {%4, %5};\n\t"
        "atom.global.relaxed.gpu.cas.b128 o, [%6], e, d;\n\t"
        "mov.b128 {%0, %1}, o;\n\t}"
        : "=l"(o0), "=l"(o1)
        : "l"(expected.z), "l"(expected.p), "l"(desired.z), "l"(desired.p), "l"(addr)
        : "memory");
    old->z = o0;
    old->p = o1;
    return o0 == expected.z && o1 == expected.p;
}

#else
// tests/emu compiles the kernel bodies with a host compiler: the three PTX helpers above as plain C++ behind one lock
// (same semantics: 16-byte load, 8-byte load, 16-byte compare-and-swap returning the old value)
rr_best_t rr_best_load(const rr_best_t *p);
double rr_best_value(const rr_best_t *p);
bool rr_cas128(rr_best_t *addr, rr_best_t expected, rr_best_t desired, rr_best_t *old);
#endif

// M[g] = max(M[g], Z) with strict > (MaxCorrelation.c:822-823); among equal values the
// smallest partner id is kept.  Z must be > 0 (NaN and -0.0 never win).
__device__ __forceinline__ void rr_best_update(rr_best_t *best, int g, double Z, int partner)
{
    if (!(Z > 0.0)) return;
    rr_best_t nw;
    nw.z = (unsigned long long)__double_as_longlong(Z);
    nw.p = (unsigned long long)(unsigned int)partner;
    rr_best_t cur = rr_best_load(best + g);
    while (nw.z > cur.z || (nw.z == cur.z && nw.p < cur.p)) {
        rr_best_t old;
        if (rr_cas128(best + g, cur, nw, &old)) break;
        cur = old;
    }
}

// a LOWER bound of a score some pair of group g attains: raises the group's maximum as a partner-less entry (it prunes
// like a maximum, loses every tie against a real pair and is overwritten by the pair's exact score, which is larger)
__device__ __forceinline__ void rr_best_raise(rr_best_t *best, int g, double lb)
{
    if (!(lb > 0.0)) return;
    rr_best_t nw;
    nw.z = (unsigned long long)__double_as_longlong(lb);
    nw.p = ~0ull;
    rr_best_t cur = rr_best_load(best + g);
    while (nw.z > cur.z) {
        rr_best_t old;
        if (rr_cas128(best + g, cur, nw, &old)) break;
        cur = old;
    }
}

// One pair test.  mi / mj are (possibly stale, hence conservative) running maxima of the
// two groups.  Returns the score, or a negative number if the pair was pruned (its score
// is provably below both maxima).
__device__ __forceinline__ double rr_pair_score(const rr_scan_params &P, unsigned s, unsigned gr1, unsigned gr2,
                                                unsigned cov, int sizei, int sizej, double mi, double mj,
                                                unsigned &n_exact, unsigned &n_bound)
{
    if (gr1 == 0 || gr2 == 0 || s < 1) return 0.0;  // MaxCorrelation.c:428-430
    if (!(P.flags & RR_FLAG_NO_PRUNE)) {
        // s at the lower end of the support: P[X >= s] = 1, and GSL returns exactly 1 (score 0): nothing to record
        if (s + cov == gr1 + gr2) return -1.0;
        double m = fmin(mi, mj);
        if (m > RR_BOUND_MEDIAN && rr_below_median(s, gr1, gr2, cov)) return -1.0;
        if (m > 0.0) {
            n_bound++;
            if (rr_bound_effective(rr_score_upper_bound(P.lnfact, s, gr1, gr2, cov)) < m) return -1.0;
        }
    }
    n_exact++;
    return rr_positive_significance(P.lnfact, s, gr1, gr2, cov, sizei, sizej);
}

// ---- tiered pruning with deferred, compacted evaluation -------------------------------------------
// Per pair test, in order of cost:
//   tier 0  trivial zero (MaxCorrelation.c:428-430)                          a few integer ops
//   tier 1  single-term pmf bound at max(s, ~mean+1) (rr_score.h), FP32       6 table look-ups
//   tier 2  windowed partial-sum bound in FP32 against FRESH maxima          ~100 FP32 ops
//   tier 3  the exact score (exp + hypergeometric series + log10)            10^3..10^4 FP64 ops
// Tiers 2 and 3 are needed by a few percent / per mille of the pairs.  Evaluating them in place
// would serialise the warp behind single lanes, so survivors are pushed into per-warp shared
// memory queues and evaluated 32 at a time, one candidate per lane.  In the tcgen05 kernel tier 3 does
// not run inside the scan at all (rr_scan_params::defer_mode): tier 2 leaves a lower bound of the score
// in the two maxima and the candidate in a list that a dense kernel evaluates after the scan.
struct __align__(8) rr_cand {
    uint32_t s, gr1, gr2, cov;   // the four counts of PositiveSignificance (423-426)
    int32_t gi, gj;              // row group, column group
};
constexpr int RR_QUEUE_CAP = 64;

// a queued candidate in 20 bytes (the per-warp queues are the largest shared-memory item of the tcgen05 kernel after the
// operand stages): four counts of 22 bits (the tcgen05 variants take R < 2^22) and two group ids of 26 bits (5N < 2^26)
struct rr_cand_p { uint32_t w[5]; };
__device__ __forceinline__ rr_cand_p rr_cand_pack(const rr_cand &c)
{
    rr_cand_p p;
    p.w[0] = c.s | (c.gr1 << 22);                                        // s[0:22) gr1[0:10)
    p.w[1] = (c.gr1 >> 10) | (c.gr2 << 12);                              // gr1[10:22) gr2[0:20)
    p.w[2] = (c.gr2 >> 20) | (c.cov << 2) | ((uint32_t)c.gi << 24);      // gr2[20:22) cov[0:22) gi[0:8)
    p.w[3] = ((uint32_t)c.gi >> 8) | ((uint32_t)c.gj << 18);             // gi[8:26) gj[0:14)
    p.w[4] = (uint32_t)c.gj >> 14;                                       // gj[14:26)
    return p;
}
__device__ __forceinline__ rr_cand rr_cand_unpack(const rr_cand_p &p)
{
    rr_cand c;
    c.s = p.w[0] & 0x3fffffu;
    c.gr1 = (p.w[0] >> 22) | ((p.w[1] & 0xfffu) << 10);
    c.gr2 = (p.w[1] >> 12) | ((p.w[2] & 0x3u) << 20);
    c.cov = (p.w[2] >> 2) & 0x3fffffu;
    c.gi = (int32_t)((p.w[2] >> 24) | ((p.w[3] & 0x3ffffu) << 8));
    c.gj = (int32_t)((p.w[3] >> 18) | (p.w[4] << 14));
    return c;
}

// ln(n!) look-up for the bounds: LT is a callable `double operator()(unsigned n)` supplied by the kernel
// (shared-memory table, HBM table, or both).
struct rr_lnf_global {
    const double *gmem;
    __device__ __forceinline__ double operator()(unsigned n) const { return __ldg(gmem + n); }
};
// ln C(n, m) for bounds only (association order irrelevant; ln 0! = 0 makes m = 0 / m = n come out as 0)
template <class LT>
__device__ __forceinline__ double rr_lnchoose_t(const LT &T, unsigned n, unsigned m)
{
    return T(n) - T(m) - T(n - m);
}

// tier 0 + 1, evaluated in FP32 on a float copy of the ln n! table: lnc3 = lnchoose(cov, gr1) is shared by the five
// column groups of a site; straight-line code (no branches) so the look-ups of a site overlap.
// (half the shared-memory bytes of a double table, no FP64 issue slots)
// Every rounding is covered by `margin` (log10 units): 9 table entries rounded to float plus 8 float additions
// of values <= ln(maxcov!) err by at most 17 * 2^-25 * ln(maxcov!) in ln units; the host passes
// 16 * 2^-24 * ln(maxcov!) * log10(e) (+2e-6), almost twice that.  thr is the
// threshold rounded DOWN to float and clamped by rr_thr_f32 (0 = "no maximum yet / never prune");
// meanfac = gr1 / cov (approximate).  LTF: callable float(unsigned n).
__device__ __forceinline__ float rr_thr_f32(double best, bool no_prune)
{
    // a raw score above 98 saturates to 98 + F, which may exceed it (rr_bound_effective): a bound only prunes when it
    // is below min(threshold, 98), so the clamp is folded into the threshold once per refresh
    return no_prune ? 0.0f : fminf(__double2float_rd(best), (float)RR_SATURATION_START);
}

template <class LTF>
__device__ __forceinline__ bool rr_tier1_f32(const LTF &T, unsigned s, unsigned gr1, unsigned gr2, unsigned cov,
                                             float thr, float lnc3, float meanfac, float margin)
{
    // s >= 1 implies gr1 >= 1 and gr2 >= 1 (s <= gr1, gr2): the three tests of 428-430 collapse to one
    const bool nz = s >= 1u;
    const unsigned hi = gr1 < gr2 ? gr1 : gr2;
    // pmf is evaluated at x = max(s, ~mean + 1) clamped to the support; x >= s >= max(0, gr1 + gr2 - cov) already
    unsigned x = (unsigned)(meanfac * (float)gr2) + 1u;
    x = x > s ? x : s;
    x = x < hi ? x : hi;
    const float lp = ((T(gr2) - T(x)) - T(gr2 - x)) + ((T(cov - gr2) - T(gr1 - x)) - T((cov + x) - (gr1 + gr2))) - lnc3;
    // U >= 0 (lp <= 0 up to the rounding the margin covers), so thr == 0 never prunes
    const float U = -(float)RR_LOG10E * lp + margin;
    return nz & !(U < thr);
}

// Tier 1 of the tcgen05 kernel: the same bound in FIXED POINT.  The table holds q(n) = round(ln(n!) * S), S a power of two
// with ln(maxcov!) * S < 2^30, addressed by 4 n (the accumulators hold 4 x count = the byte offset of an entry; every
// difference of scaled counts is again one, and x is kept a multiple of 4).  Sums and differences of entries are exact in
// integer arithmetic and take 3-input adds; what is left to cover is the rounding of the nine entries of ln pmf
// (|error| <= 4.5 units): RR_T1Q_SLACK.  A running maximum becomes the integer nq = SLACK - thrq, thrq = the threshold in
// table units rounded DOWN (rr_thr_q); the pair survives iff lnpmf_q <= max(nq_i, nq_j), i.e. unless
//   -log10 pmf  <=  log10(e) (thrq - 1.5) / S  <  min(threshold_i, threshold_j).
// x = (round(mean) + 1) << 2 comes out of one FFMA: adding 2^25 to meanfac * gr2 (= 4 x mean < 2^25) leaves round(mean) in
// the mantissa bits (ulp 4 there), so x = 4 * bits + (4 - 4 * 0x4C000000) in wrap-around arithmetic; any x >= s inside the
// support gives a valid bound.  lnc3q = q(cov) - q(gr1) - q(cov - gr1), or RR_T1Q_NEVER for rows without a pair test at the
// site (their ln pmf then comes out far above every nq).
constexpr int RR_T1Q_SLACK = 5;
constexpr int RR_T1Q_NEVER = -(1 << 30);
constexpr int RR_T1Q_INADMISSIBLE = -0x7fffffff;   // nq of a column group that is not admissible: nothing survives it

// S = 2^k, the largest with ln(maxcov!) * S < 2^30 (k <= 20)
RR_HD int rr_t1q_shift(double lnfact_maxcov)
{
    int k = 20;
    while (k > 0 && lnfact_maxcov * (double)(1 << k) >= 1073741824.0) k--;
    return k;
}

// nq of a running maximum; qscale = ln(10) * S rounded down to float.  A maximum of 0 ("none yet") and the exhaustive scan
// give nq = SLACK: ln pmf <= 0 always survives.  The clamp at 98 is rr_bound_effective folded in (a raw score above 98 is
// replaced by 98 + F, which can exceed it).  -2: the float product and qscale may each be up to one unit too high.
__device__ __forceinline__ int rr_thr_q(double best, bool no_prune, float qscale)
{
    if (no_prune) return RR_T1Q_SLACK;
    const float t = fminf(__double2float_rd(best), (float)RR_SATURATION_START);
    const int thrq = __float2int_rd(t * qscale) - 2;
    return RR_T1Q_SLACK - (thrq > 0 ? thrq : 0);
}

template <class LTQ>
__device__ __forceinline__ bool rr_tier1_q(const LTQ &T, unsigned s, unsigned gr1, unsigned gr2, unsigned cov, int nq, int lnc3q,
                                           float meanfac)
{
    const bool nz = s != 0u;
    const unsigned hi = gr1 < gr2 ? gr1 : gr2;
    unsigned x = __float_as_uint(__fmaf_rn(meanfac, (float)gr2, 33554432.0f)) * 4u + 0xD0000004u;
    x = x > s ? x : s;
    x = x < hi ? x : hi;
    const int lp = ((T(gr2) - T(x) - T(gr2 - x)) + (T(cov - gr2) - T(gr1 - x) - T((cov + x) - (gr1 + gr2)))) - lnc3q;
    return nz & (lp <= nq);
}

// tier 2: P[X >= s] >= sum_{x = x0}^{x0+m} pmf(x) for any x0 >= s; the window starts next to the mean
// (or at s above it) and the term ratios are accumulated in FP32, every factor rounded down.
template <class LT>
__device__ __forceinline__ bool rr_tier2(const LT &T, unsigned s, unsigned gr1, unsigned gr2, unsigned cov,
                                         double thr)
{
    if (!(thr > 0.0)) return true;
    const unsigned hi = gr1 < gr2 ? gr1 : gr2;
    const unsigned lo = gr1 + gr2 > cov ? gr1 + gr2 - cov : 0u;
    int xm = (int)__fdividef((float)gr1 * (float)gr2, (float)cov) - 3;
    unsigned x0 = xm > (int)s ? (unsigned)xm : s;
    if (x0 > hi) x0 = hi;
    if (x0 < lo) x0 = lo;
    const double lp = rr_lnchoose_t(T, gr2, x0) + rr_lnchoose_t(T, cov - gr2, gr1 - x0) - rr_lnchoose_t(T, cov, gr1);
    float S = 1.0f, term = 1.0f;
    unsigned x = x0;
#pragma unroll 1
    for (int m = 0; m < 24 && x < hi; m++, x++) {
        const float num = (float)(gr2 - x) * (float)(gr1 - x);
        const float den = (float)(x + 1u) * (float)((cov + x + 1u) - gr1 - gr2);
        term *= __fdividef(num, den) * 0.99999f;
        S += term;
        if (term < 1e-3f * S) break;
    }
    const double U2 = -RR_LOG10E * lp - (double)(__log2f(S) * 0.30103f * 0.99999f) + 2e-5;
    return !(rr_bound_effective(U2) < thr);
}

// Tier 2 with both ends: the pair survives unless an UPPER bound of its score is below thr (as rr_tier2), and *zlb receives
// a LOWER bound of its final score (0 if none is available), i.e. an upper bound of P[X >= s] = pmf(s) * (1 + r1 + r1 r2 +
// ...): the term ratios r_k = (gr1 - x)(gr2 - x) / ((x + 1)(cov + x + 1 - gr1 - gr2)) fall with x, so what lies behind the
// last term added is at most term * r / (1 - r) with the next ratio r < 1.  The upper sum takes every factor rounded up;
// the final value is lowered by 2e-5 relative + 2e-5 (FP32 accumulation over <= 24 factors of (1 + 2^-23)(1 + 1e-5), log2f
// to 2 ulp; the exact score's own error is orders below) and clamped at 98: a raw score above 98 becomes 98 + F > 98 (432).
template <class LT>
__device__ __forceinline__ bool rr_tier2_interval(const LT &T, unsigned s, unsigned gr1, unsigned gr2, unsigned cov, double thr,
                                                  double *zlb)
{
    *zlb = 0.0;
    const unsigned hi = gr1 < gr2 ? gr1 : gr2;
    const unsigned lo = gr1 + gr2 > cov ? gr1 + gr2 - cov : 0u;
    int xm = (int)__fdividef((float)gr1 * (float)gr2, (float)cov) - 3;
    unsigned x0 = xm > (int)s ? (unsigned)xm : s;
    if (x0 > hi) x0 = hi;
    if (x0 < lo) x0 = lo;
    const double lp = rr_lnchoose_t(T, gr2, x0) + rr_lnchoose_t(T, cov - gr2, gr1 - x0) - rr_lnchoose_t(T, cov, gr1);
    float S = 1.0f, term = 1.0f, Su = 1.0f, termu = 1.0f, r_next = 0.0f;
    unsigned x = x0;
#pragma unroll 1
    for (int m = 0; m < 24; m++, x++) {
        if (x >= hi) { r_next = 0.0f; break; }
        const float num = (float)(gr2 - x) * (float)(gr1 - x);
        const float den = (float)(x + 1u) * (float)((cov + x + 1u) - gr1 - gr2);
        const float r = __fdividef(num, den);
        r_next = r * 1.00002f;             // ratio of the term after `termu` to it, rounded up; ratios fall with x
        if (term < 1e-3f * S) break;
        term *= r * 0.99999f;
        termu *= r * 1.00002f;
        S += term;
        Su += termu;
        if (m == 23) {                     // ran out of terms: the ratio behind the last one
            const unsigned x1 = x + 1u;
            r_next = x1 >= hi ? 0.0f
                              : __fdividef((float)(gr2 - x1) * (float)(gr1 - x1), (float)(x1 + 1u) * (float)((cov + x1 + 1u) - gr1 - gr2)) * 1.00002f;
        }
    }
    const double U2 = -RR_LOG10E * lp - (double)(__log2f(S) * 0.30103f * 0.99999f) + 2e-5;
    const bool keep = !(thr > 0.0) || !(rr_bound_effective(U2) < thr);
    if (keep && x0 == s && r_next < 0.9f) {
        // P[X >= s] <= pmf(s) * (Su + termu * r / (1 - r))
        const float tail = termu * r_next / (1.0f - r_next) * 1.00002f;
        const float Sall = (Su + tail) * 1.00002f;
        double L = -RR_LOG10E * lp - (double)(__log2f(Sall) * 0.30103f * 1.00002f);
        L = L * (1.0 - 2e-5) - 2e-5;
        *zlb = L > RR_SATURATION_START ? RR_SATURATION_START : (L > 0.0 ? L : 0.0);
    }
    return keep;
}

__device__ __forceinline__ void rr_queue_push(rr_cand_p *q, int &count, bool need, const rr_cand &c, int lane)
{
    const unsigned mask = __ballot_sync(0xffffffffu, need);
    if (mask == 0u) return;
    if (need) q[count + __popc(mask & ((1u << lane) - 1u))] = rr_cand_pack(c);
    count += __popc(mask);
}

// tier 3 on queued candidates, 32 per round (all of them when flush is set); both groups' maxima
// are folded in with the 128-bit CAS
static __device__ __noinline__ void rr_drain_exact(const rr_scan_params &P, rr_cand_p *q, int &count, int lane,
                                                   unsigned &n_exact, bool flush)
{
    while (count >= 32 || (flush && count > 0)) {
        __syncwarp();
        const int take = count < 32 ? count : 32;
        if (lane < take) {
            const rr_cand c = rr_cand_unpack(q[count - take + lane]);
            n_exact++;
            const double Z = rr_positive_significance(P.lnfact, c.s, c.gr1, c.gr2, c.cov, __ldg(P.gsize + c.gi), __ldg(P.gsize + c.gj));
            if (Z > 0.0) {
                if (Z >= rr_best_value(P.best + c.gi)) rr_best_update(P.best, c.gi, Z, c.gj);
                if (Z >= rr_best_value(P.best + c.gj)) rr_best_update(P.best, c.gj, Z, c.gi);
            }
        }
        count -= take;
        __syncwarp();
    }
}

// tier 2 on queued tier-1 survivors; what survives moves on to the exact queue - or, with deferred evaluation
// (rr_scan_params::defer_mode), raises the two groups' maxima by a lower bound of its score and goes to the global list
// that rr_k_deferred_exact evaluates after the scan.  An epilogue warp then never sits in the FP64 series while the other
// warps of its CTA pair wait for it at the next tile.
template <class LT>
static __device__ __noinline__ void rr_drain_tier2(const rr_scan_params &P, const LT &T, rr_cand_p *q1, int &c1,
                                                   rr_cand_p *q2, int &c2, int lane, unsigned &n_tier2, unsigned &n_exact,
                                                   bool flush)
{
    while (c1 >= 32 || (flush && c1 > 0)) {
        __syncwarp();
        const int take = c1 < 32 ? c1 : 32;
        bool need = false;
        double zlb = 0.0;
        rr_cand c;
        c.s = c.gr1 = c.gr2 = c.cov = 0u; c.gi = c.gj = 0;
        if (lane < take) {
            c = rr_cand_unpack(q1[c1 - take + lane]);
            n_tier2++;
            need = true;
            if (!(P.flags & RR_FLAG_NO_PRUNE)) {
                const double thr = fmin(rr_best_value(P.best + c.gi), rr_best_value(P.best + c.gj));
                need = rr_tier2_interval(T, c.s, c.gr1, c.gr2, c.cov, thr, &zlb);
            }
        }
        c1 -= take;
        __syncwarp();
        if (P.defer_mode != 0) {
            if (need && zlb > 0.0) { rr_best_raise(P.best, c.gi, zlb); rr_best_raise(P.best, c.gj, zlb); }
            if (P.defer_mode == 2) need = false;            // seeding pass: thresholds only, the pair comes again
            else {
                const unsigned mask = __ballot_sync(0xffffffffu, need);
                if (mask != 0u) {
                    unsigned long long base = 0ull;
                    if (lane == 0) base = atomicAdd(P.counters + 5, (unsigned long long)__popc(mask));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    const unsigned long long idx = base + (unsigned long long)__popc(mask & ((1u << lane) - 1u));
                    if (need && idx < P.deferred_cap) { P.deferred[idx] = rr_cand_pack(c); need = false; }
                    // (entries past the capacity are evaluated in place, below)
                }
            }
        }
        rr_queue_push(q2, c2, need, c, lane);
        if (c2 >= 32) rr_drain_exact(P, q2, c2, lane, n_exact, false);
    }
    if (flush) rr_drain_exact(P, q2, c2, lane, n_exact, true);
}
