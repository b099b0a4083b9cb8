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
    unsigned long long *counters; // [4]: pair tests, exact evals, bound evals, work units
    // work decomposition (bitset kernel): row blocks x column blocks
    const int64_t *unit_prefix;  // [n_rowblocks+1] prefix sum of column-block counts
    const int32_t *unit_cb0;     // [n_rowblocks] first column block of each row block
    int n_rowblocks;
    int rb_lo, rb_hi;            // this part's row blocks
    const int32_t *word_hi;      // [n_rowblocks] exclusive upper u32-word bound of contributing rows
    const int32_t *word_lo;      // [n_colblocks] inclusive lower u32-word bound
};

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

// One pair test.  mi / mj are (possibly stale, hence conservative) running maxima of the
// two groups.  Returns the score, or a negative number if the pair was pruned (its score
// is provably below both maxima).
__device__ __forceinline__ double rr_pair_score(const rr_scan_params &P, unsigned s, unsigned gr1, unsigned gr2,
                                                unsigned cov, int sizei, int sizej, double mi, double mj,
                                                unsigned &n_exact, unsigned &n_bound)
{
    if (gr1 == 0 || gr2 == 0 || s < 1) return 0.0;  // MaxCorrelation.c:428-430
    if (!(P.flags & RR_FLAG_NO_PRUNE)) {
        double m = fmin(mi, mj);
        if (m > RR_BOUND_MEDIAN && rr_below_median(s, gr1, gr2, cov)) return -1.0;
        if (m > 0.0) {
            n_bound++;
            if (rr_score_upper_bound(P.lnfact, s, gr1, gr2, cov) < m) return -1.0;
        }
    }
    n_exact++;
    return rr_positive_significance(P.lnfact, s, gr1, gr2, cov, sizei, sizej);
}
