// rr_cliquer.cu -- Cliquer, the inner step of Group_Refinement (/root/reference/RepeatResolver.c:1179-1240), for a
// batch of query groups on the device bitsets of the packed MSA.
//
// The reference calls Cliquer once for every group whose MaxCorrs value is above the cutoff (1647-1649, 1723-1725);
// each call walks all 5 * (ende - anfang) candidate groups with one Schnitt (1213) and, where that intersection is
// above mincov/4 (1215), three more plus the hypergeometric score (Group_PositiveSignificance 472-488).  Here:
//
//   rr_k_cliquer_counts   one block = (batch of CLQ_QB queries, slab of CLQ_SLAB candidate sites).  The queries'
//                         group and coverage bitsets sit in shared memory; a warp takes one candidate site, streams
//                         its six bitsets (5 groups + coverage) 32 words at a time and accumulates, per query, the
//                         twelve AND+POPC counts of the site (5 x |Gk & Gq|, 5 x |Gk & Cq|, |Gq & Ck|, |Ck & Cq|).
//                         Integer work on HBM-resident bitsets: blocks of one slab are adjacent in launch order
//                         (blockIdx.x = batch), so a slab is read from HBM once and then served from L2 to every
//                         batch; two exact skips cut the words touched: 32-word chunks in which no query of the
//                         batch covers a read, and chunks in which the candidate site covers none (rows are in
//                         span order, so both are long runs).  Pairs that pass 1215 and whose rigorous score bound
//                         (rr_score.h) can exceed `greedy` are appended to a candidate list.
//   rr_k_cliquer_counts2  experiment, selectable with RR_CLIQUER_KERNEL=2, same results: the counts split in two steps
//                         (RR_CLQ_QB2 queries per block).  The stream over the site computes only the five |Gk & Gq|
//                         (the groups of a site partition its coverage, so their sum is |Gq & Ck|); the (query, group)
//                         pairs above mincov/4 then get |Gk & Cq| and |Ck & Cq| from a second pass over the site's
//                         words.  5 instead of 12 POPC per word and query - but measured SLOWER on a B200 (27.6 against
//                         18.5 ms for 1024 queries on a 4740 x 26594 MSA, profiles/r1_cliquer_ncu_summary.csv): the
//                         second pass fires for most sites near the query (their major group shares its reads), the
//                         kernel executes 1.48x the warp instructions and the POPC pipe drops from 62 % to 25 % busy.
//                         The one-step kernel stays the default.
//   rr_k_cliquer_score    one thread per listed candidate: the exact score in IEEE double, GSL's operation order
//                         (rr_group_significance); candidates above greedy (less a 1e-9 margin) go to the hit list.
//
// The host (rr_cliquer_batch in rr_abi.cu) sorts the hits per query, re-evaluates the few that decide the top
// maxclique-1 with the host libm (the same finalisation as the scan's RR_FLAG_HOST_FINALIZE) and applies
// TheBestUpdater's order (1156-1176).
#include <algorithm>
#include "rr_kernels.h"
#include "rr_score.h"

constexpr int CLQ_QB = RR_CLQ_QB;       // queries per block, one-step kernel
constexpr int CLQ_QB2 = RR_CLQ_QB2;     // queries per block, two-step kernel
constexpr int CLQ_SLAB = RR_CLQ_SLAB;   // candidate sites per block
constexpr int CLQ_WARPS = 8;
constexpr unsigned CLQ_FULL = 0xffffffffu;

static_assert(CLQ_QB * 5 <= 32 && CLQ_QB2 * 5 <= 32, "one lane per (query, group of the site) in the tail of the site loop");
static_assert(sizeof(rr_clq_rec) == 32, "record layout is shared with the host");

__device__ __forceinline__ void clq_append(rr_clq_rec *list, unsigned long long cap, unsigned long long *counter,
                                           const rr_clq_rec &r)
{
    const unsigned long long idx = atomicAdd(counter, 1ull);   // keeps counting past cap: the host sees the overflow
    if (idx < cap) list[idx] = r;
}

// the block's queries into shared memory: group and coverage bitsets, zero padded to whole 32-word chunks, and per chunk
// the set of queries that cover a read of it
template <int QB>
__device__ __forceinline__ void clq_stage_queries(const uint32_t *__restrict__ bits, const uint32_t *__restrict__ covbits,
                                                  int W32, const int32_t *__restrict__ queries, int nq, int slot0,
                                                  uint32_t *qg, uint32_t *qc, uint32_t *qmask)
{
    const int nchunks = (W32 + 31) >> 5, W32p = nchunks << 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int idx = threadIdx.x; idx < QB * W32p; idx += CLQ_WARPS * 32) {
        const int q = idx / W32p, w = idx - q * W32p;
        uint32_t y = 0u, cy = 0u;
        if (slot0 + q < nq && w < W32) {
            const int g = queries[slot0 + q];
            y = bits[(size_t)g * W32 + w];
            cy = covbits[(size_t)(g / 5) * W32 + w];
        }
        qg[idx] = y;
        qc[idx] = cy;
    }
    __syncthreads();
    for (int c = warp; c < nchunks; c += CLQ_WARPS) {
        unsigned m = 0u;
#pragma unroll
        for (int q = 0; q < QB; q++)
            if (__ballot_sync(CLQ_FULL, qc[q * W32p + c * 32 + lane] != 0u)) m |= 1u << q;
        if (lane == 0) qmask[c] = m;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(CLQ_WARPS * 32, 2)
rr_k_cliquer_counts(const uint32_t *__restrict__ bits, const uint32_t *__restrict__ covbits, int W32,
                    const int32_t *__restrict__ queries, int nq, int anfang, int ende, int min_s, double greedy,
                    const double *__restrict__ lnf, rr_clq_rec *__restrict__ cand, unsigned long long cap,
                    unsigned long long *__restrict__ counter)
{
    RR_DYN_SMEM(uint32_t, clq_smem);
    const int nchunks = (W32 + 31) >> 5, W32p = nchunks << 5;
    uint32_t *qg = clq_smem;                          // [CLQ_QB][W32p] query group bitsets, zero padded
    uint32_t *qc = qg + (size_t)CLQ_QB * W32p;        // [CLQ_QB][W32p] coverage of the queries' sites
    uint32_t *qmask = qc + (size_t)CLQ_QB * W32p;     // [nchunks] bit q: query q covers a read of this chunk
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot0 = blockIdx.x * CLQ_QB;
    clq_stage_queries<CLQ_QB>(bits, covbits, W32, queries, nq, slot0, qg, qc, qmask);

    // my (query, group of the candidate site) in the tail of the site loop
    const int my_q = lane / 5, my_k = lane - my_q * 5;
    const bool my_valid = lane < CLQ_QB * 5 && slot0 + my_q < nq;
    const int my_query = my_valid ? queries[slot0 + my_q] : -1;

    const int site_end = min(ende, anfang + ((int)blockIdx.y + 1) * CLQ_SLAB);
    for (int ii = anfang + (int)blockIdx.y * CLQ_SLAB + warp; ii < site_end; ii += CLQ_WARPS) {
        unsigned s[CLQ_QB][5], g1[CLQ_QB][5], g2[CLQ_QB], cv[CLQ_QB];
#pragma unroll
        for (int q = 0; q < CLQ_QB; q++) {
            g2[q] = cv[q] = 0u;
#pragma unroll
            for (int k = 0; k < 5; k++) s[q][k] = g1[q][k] = 0u;
        }
        const uint32_t *cb = covbits + (size_t)ii * W32;
        const uint32_t *gb = bits + (size_t)ii * 5 * W32;
        for (int c = 0; c < nchunks; c++) {
            const unsigned m = qmask[c];
            if (m == 0u) continue;                                   // no query of the batch covers a read here
            const int w = c * 32 + lane;
            const uint32_t cx = w < W32 ? __ldg(cb + w) : 0u;
            if (__ballot_sync(CLQ_FULL, cx != 0u) == 0u) continue;   // nor does the candidate site
            uint32_t x[5];
#pragma unroll
            for (int k = 0; k < 5; k++) x[k] = w < W32 ? __ldg(gb + (size_t)k * W32 + w) : 0u;
#pragma unroll
            for (int q = 0; q < CLQ_QB; q++) {
                if (!((m >> q) & 1u)) continue;                      // warp-uniform
                const uint32_t y = qg[q * W32p + w], cy = qc[q * W32p + w];
#pragma unroll
                for (int k = 0; k < 5; k++) {
                    s[q][k] += __popc(x[k] & y);
                    g1[q][k] += __popc(x[k] & cy);
                }
                g2[q] += __popc(y & cx);
                cv[q] += __popc(cx & cy);
            }
        }
        // warp totals (redux.sync), only as far as they are needed: every group of the site lies inside its coverage,
        // so |Gk & Gq| <= |Ck & Gq| and a site with |Ck & Gq| <= mincov/4 has no group that passes 1215
        int ms = 0, mg1 = 0, mg2 = 0, mcv = 0;
#pragma unroll
        for (int q = 0; q < CLQ_QB; q++) {
            const int t2 = (int)__reduce_add_sync(CLQ_FULL, g2[q]);
            if (t2 <= min_s) continue;                               // warp-uniform
            const int tc = (int)__reduce_add_sync(CLQ_FULL, cv[q]);
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const int ts = (int)__reduce_add_sync(CLQ_FULL, s[q][k]);
                if (ts <= min_s) continue;                           // warp-uniform
                const int t1 = (int)__reduce_add_sync(CLQ_FULL, g1[q][k]);
                if (lane == q * 5 + k) { ms = ts; mg1 = t1; mg2 = t2; mcv = tc; }
            }
        }
        const int group = ii * 5 + my_k;
        if (my_valid && ms > min_s && group != my_query) {           // 1210, 1215
            // Z <= bound; a bound above 98 proves nothing (the raw score is replaced by 97.90 + F1 there, 486)
            const double bound = rr_bound_effective(rr_score_upper_bound(lnf, (unsigned)ms, (unsigned)mg1, (unsigned)mg2, (unsigned)mcv));
            if (bound > greedy) {
                rr_clq_rec r;
                r.slot = slot0 + my_q; r.group = group; r.s = ms; r.gr1 = mg1; r.gr2 = mg2; r.cov = mcv; r.z = 0.0;
                clq_append(cand, cap, counter, r);
            }
        }
    }
}

// two-step counts: see the header comment
__global__ void __launch_bounds__(CLQ_WARPS * 32, 2)
rr_k_cliquer_counts2(const uint32_t *__restrict__ bits, const uint32_t *__restrict__ covbits, int W32,
                     const int32_t *__restrict__ queries, int nq, int anfang, int ende, int min_s, double greedy,
                     const double *__restrict__ lnf, rr_clq_rec *__restrict__ cand, unsigned long long cap,
                     unsigned long long *__restrict__ counter)
{
    RR_DYN_SMEM(uint32_t, clq_smem);
    const int nchunks = (W32 + 31) >> 5, W32p = nchunks << 5;
    uint32_t *qg = clq_smem;
    uint32_t *qc = qg + (size_t)CLQ_QB2 * W32p;
    uint32_t *qmask = qc + (size_t)CLQ_QB2 * W32p;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot0 = blockIdx.x * CLQ_QB2;
    clq_stage_queries<CLQ_QB2>(bits, covbits, W32, queries, nq, slot0, qg, qc, qmask);

    const int my_q = lane / 5, my_k = lane - my_q * 5;
    const bool my_valid = lane < CLQ_QB2 * 5 && slot0 + my_q < nq;
    const int my_query = my_valid ? queries[slot0 + my_q] : -1;

    const int site_end = min(ende, anfang + ((int)blockIdx.y + 1) * CLQ_SLAB);
    for (int ii = anfang + (int)blockIdx.y * CLQ_SLAB + warp; ii < site_end; ii += CLQ_WARPS) {
        unsigned s[CLQ_QB2][5];
#pragma unroll
        for (int q = 0; q < CLQ_QB2; q++)
#pragma unroll
            for (int k = 0; k < 5; k++) s[q][k] = 0u;
        const uint32_t *cb = covbits + (size_t)ii * W32;
        const uint32_t *gb = bits + (size_t)ii * 5 * W32;
        // step 1: |Gk & Gq| for the five groups of the site and every query of the block
        for (int c = 0; c < nchunks; c++) {
            const unsigned m = qmask[c];
            if (m == 0u) continue;
            const int w = c * 32 + lane;
            const uint32_t cx = w < W32 ? __ldg(cb + w) : 0u;
            if (__ballot_sync(CLQ_FULL, cx != 0u) == 0u) continue;
            uint32_t x[5];
#pragma unroll
            for (int k = 0; k < 5; k++) x[k] = w < W32 ? __ldg(gb + (size_t)k * W32 + w) : 0u;
#pragma unroll
            for (int q = 0; q < CLQ_QB2; q++) {
                if (!((m >> q) & 1u)) continue;                      // warp-uniform
                const uint32_t y = qg[q * W32p + w];
#pragma unroll
                for (int k = 0; k < 5; k++) s[q][k] += __popc(x[k] & y);
            }
        }
        int ms = 0, mg1 = 0, mg2 = 0, mcv = 0;
#pragma unroll
        for (int q = 0; q < CLQ_QB2; q++) {
            // the five groups of a site are disjoint and their union is its coverage (rr_k_pack_bits): |Gq & Ck| = sum
            const int t2 = (int)__reduce_add_sync(CLQ_FULL, s[q][0] + s[q][1] + s[q][2] + s[q][3] + s[q][4]);
            if (t2 <= min_s) continue;                               // warp-uniform: no group of the site passes 1215
            int ts[5];
            bool any = false;
#pragma unroll
            for (int k = 0; k < 5; k++) {
                ts[k] = (int)__reduce_add_sync(CLQ_FULL, s[q][k]);
                any = any || ts[k] > min_s;
            }
            if (!any) continue;                                      // warp-uniform
            // step 2: |Ck & Cq| and, for the groups that passed, |Gk & Cq|; the site's words are still in L1
            unsigned cacc = 0u, gacc[5] = {0u, 0u, 0u, 0u, 0u};
            for (int c = 0; c < nchunks; c++) {
                if (!((qmask[c] >> q) & 1u)) continue;
                const int w = c * 32 + lane;
                const uint32_t cx = w < W32 ? __ldg(cb + w) : 0u;
                if (__ballot_sync(CLQ_FULL, cx != 0u) == 0u) continue;
                const uint32_t cy = qc[q * W32p + w];
                cacc += __popc(cx & cy);
#pragma unroll
                for (int k = 0; k < 5; k++)
                    if (ts[k] > min_s) gacc[k] += __popc((w < W32 ? __ldg(gb + (size_t)k * W32 + w) : 0u) & cy);
            }
            const int tc = (int)__reduce_add_sync(CLQ_FULL, cacc);
#pragma unroll
            for (int k = 0; k < 5; k++) {
                if (ts[k] <= min_s) continue;                        // warp-uniform
                const int t1 = (int)__reduce_add_sync(CLQ_FULL, gacc[k]);
                if (lane == q * 5 + k) { ms = ts[k]; mg1 = t1; mg2 = t2; mcv = tc; }
            }
        }
        const int group = ii * 5 + my_k;
        if (my_valid && ms > min_s && group != my_query) {           // 1210, 1215
            const double bound = rr_bound_effective(rr_score_upper_bound(lnf, (unsigned)ms, (unsigned)mg1, (unsigned)mg2, (unsigned)mcv));
            if (bound > greedy) {
                rr_clq_rec r;
                r.slot = slot0 + my_q; r.group = group; r.s = ms; r.gr1 = mg1; r.gr2 = mg2; r.cov = mcv; r.z = 0.0;
                clq_append(cand, cap, counter, r);
            }
        }
    }
}

// EXPERIMENTAL (RR_CLIQUER_KERNEL=3), written from the profile of the one-step kernel (POPC pipe 62 % busy at 16 resident
// warps) when the round's GPU minutes were spent - never run on a GPU (its logic passes under the CPU emulation of tests/emu), opt-in GPU test only.  Same structure, three changes:
//   * the groups of a site partition its coverage (rr_k_pack_bits), so the fifth group's counts are differences:
//     |G4 & Gq| = |Ck & Gq| - sum of the other four, |G4 & Cq| = |Ck & Cq| - sum of the other four: 10 POPC and five
//     loads per word instead of 12 and six;
//   * a chunk is skipped unless some read is covered by the candidate site AND by one of the block's queries (word-wise
//     AND with the union of the queries' coverage, instead of the two chunk-level tests);
//   * three blocks per SM (launch bound 85 registers) instead of two;
//   * the coverage word of the next chunk is requested while the current chunk is counted: in the SASS view of the
//     one-step kernel 21 % of the warp-state samples sit on the consumers of the two dependent loads per chunk
//     (profiles/r1_cliquer_source_hotspots.txt), and with 4 warps per scheduler nothing hides them.
__global__ void __launch_bounds__(CLQ_WARPS * 32, 3)
rr_k_cliquer_counts3(const uint32_t *__restrict__ bits, const uint32_t *__restrict__ covbits, int W32,
                     const int32_t *__restrict__ queries, int nq, int anfang, int ende, int min_s, double greedy,
                     const double *__restrict__ lnf, rr_clq_rec *__restrict__ cand, unsigned long long cap,
                     unsigned long long *__restrict__ counter)
{
    RR_DYN_SMEM(uint32_t, clq_smem);
    const int nchunks = (W32 + 31) >> 5, W32p = nchunks << 5;
    uint32_t *qg = clq_smem;
    uint32_t *qc = qg + (size_t)CLQ_QB * W32p;
    uint32_t *qmask = qc + (size_t)CLQ_QB * W32p;
    uint32_t *qany = qmask + nchunks;                  // [W32p] union of the queries' coverage
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot0 = blockIdx.x * CLQ_QB;
    clq_stage_queries<CLQ_QB>(bits, covbits, W32, queries, nq, slot0, qg, qc, qmask);
    for (int w = threadIdx.x; w < W32p; w += CLQ_WARPS * 32) {
        uint32_t u = 0u;
#pragma unroll
        for (int q = 0; q < CLQ_QB; q++) u |= qc[q * W32p + w];
        qany[w] = u;
    }
    __syncthreads();

    const int my_q = lane / 5, my_k = lane - my_q * 5;
    const bool my_valid = lane < CLQ_QB * 5 && slot0 + my_q < nq;
    const int my_query = my_valid ? queries[slot0 + my_q] : -1;

    const int site_end = min(ende, anfang + ((int)blockIdx.y + 1) * CLQ_SLAB);
    for (int ii = anfang + (int)blockIdx.y * CLQ_SLAB + warp; ii < site_end; ii += CLQ_WARPS) {
        unsigned s[CLQ_QB][4], g1[CLQ_QB][4], g2[CLQ_QB], cv[CLQ_QB];
#pragma unroll
        for (int q = 0; q < CLQ_QB; q++) {
            g2[q] = cv[q] = 0u;
#pragma unroll
            for (int k = 0; k < 4; k++) s[q][k] = g1[q][k] = 0u;
        }
        const uint32_t *cb = covbits + (size_t)ii * W32;
        const uint32_t *gb = bits + (size_t)ii * 5 * W32;
        // chunks no query of the block touches are stepped over; the coverage word of the NEXT chunk is requested before
        // the current one is counted, so that its round trip to L2 overlaps the counting (in the one-step kernel the two
        // dependent loads per chunk - coverage word, then group words - are what the warps wait for)
        int c = 0;
        while (c < nchunks && qmask[c] == 0u) c++;
        uint32_t cx = (c < nchunks && c * 32 + lane < W32) ? __ldg(cb + c * 32 + lane) : 0u;
        while (c < nchunks) {
            int cn = c + 1;
            while (cn < nchunks && qmask[cn] == 0u) cn++;
            const uint32_t cx_next = (cn < nchunks && cn * 32 + lane < W32) ? __ldg(cb + cn * 32 + lane) : 0u;
            const unsigned m = qmask[c];
            const int w = c * 32 + lane;
            if (__ballot_sync(CLQ_FULL, (cx & qany[w]) != 0u) != 0u) {           // else: no read covered on both sides
                uint32_t x[4];
#pragma unroll
                for (int k = 0; k < 4; k++) x[k] = w < W32 ? __ldg(gb + (size_t)k * W32 + w) : 0u;
#pragma unroll
                for (int q = 0; q < CLQ_QB; q++) {
                    if (!((m >> q) & 1u)) continue;                              // warp-uniform
                    const uint32_t y = qg[q * W32p + w], cy = qc[q * W32p + w];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        s[q][k] += __popc(x[k] & y);
                        g1[q][k] += __popc(x[k] & cy);
                    }
                    g2[q] += __popc(y & cx);
                    cv[q] += __popc(cx & cy);
                }
            }
            c = cn;
            cx = cx_next;
        }
        int ms = 0, mg1 = 0, mg2 = 0, mcv = 0;
#pragma unroll
        for (int q = 0; q < CLQ_QB; q++) {
            const int t2 = (int)__reduce_add_sync(CLQ_FULL, g2[q]);
            if (t2 <= min_s) continue;                               // warp-uniform
            const int tc = (int)__reduce_add_sync(CLQ_FULL, cv[q]);
            int ts[5], t1[5];
            ts[4] = t2;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                ts[k] = (int)__reduce_add_sync(CLQ_FULL, s[q][k]);
                ts[4] -= ts[k];
            }
            const bool need5 = ts[4] > min_s;                        // the fifth |Gk & Cq| needs the other four
            t1[4] = tc;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                t1[k] = 0;
                if (ts[k] <= min_s && !need5) continue;              // warp-uniform
                t1[k] = (int)__reduce_add_sync(CLQ_FULL, g1[q][k]);
                t1[4] -= t1[k];
            }
#pragma unroll
            for (int k = 0; k < 5; k++)
                if (ts[k] > min_s && lane == q * 5 + k) { ms = ts[k]; mg1 = t1[k]; mg2 = t2; mcv = tc; }
        }
        const int group = ii * 5 + my_k;
        if (my_valid && ms > min_s && group != my_query) {           // 1210, 1215
            const double bound = rr_bound_effective(rr_score_upper_bound(lnf, (unsigned)ms, (unsigned)mg1, (unsigned)mg2, (unsigned)mcv));
            if (bound > greedy) {
                rr_clq_rec r;
                r.slot = slot0 + my_q; r.group = group; r.s = ms; r.gr1 = mg1; r.gr2 = mg2; r.cov = mcv; r.z = 0.0;
                clq_append(cand, cap, counter, r);
            }
        }
    }
}

__global__ void __launch_bounds__(128)
rr_k_cliquer_score(const rr_clq_rec *__restrict__ cand, unsigned long long cap, const unsigned long long *__restrict__ counter,
                   const int32_t *__restrict__ queries, const int32_t *__restrict__ gsize, const double *__restrict__ lnf,
                   double threshold, rr_clq_rec *__restrict__ hits, unsigned long long *__restrict__ hit_counter)
{
    const unsigned long long n = min(*counter, cap);
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < n;
         t += (unsigned long long)gridDim.x * blockDim.x) {
        rr_clq_rec r = cand[t];
        // Group1 = the candidate, Group2 = the query (1217): gr1 = |Gk & Cq|, gr2 = |Gq & Ck|
        r.z = rr_group_significance(lnf, (unsigned)r.s, (unsigned)r.gr1, (unsigned)r.gr2, (unsigned)r.cov, gsize[r.group],
                                    gsize[queries[r.slot]]);
        if (r.z > threshold) clq_append(hits, cap, hit_counter, r);
    }
}

static size_t clq_smem_bytes(int W32, int qb)
{
    const size_t nchunks = ((size_t)W32 + 31) / 32;
    return (2 * (size_t)qb * nchunks * 32 + nchunks) * sizeof(uint32_t);
}

static size_t clq_smem_bytes3(int W32) { return clq_smem_bytes(W32, CLQ_QB) + (((size_t)W32 + 31) / 32) * 32 * sizeof(uint32_t); }

size_t rr_cliquer_smem_bytes(int W32) { return std::max(clq_smem_bytes(W32, CLQ_QB2 > CLQ_QB ? CLQ_QB2 : CLQ_QB), clq_smem_bytes3(W32)); }

#ifndef RR_CPU_EMU   // tests/emu compiles the kernels above with a host compiler; the launch syntax below is nvcc only
cudaError_t rr_launch_cliquer(int kernel, const uint32_t *bits, const uint32_t *covbits, const int32_t *gsize, const double *lnf,
                              int W32, const int32_t *queries, int nq, int anfang, int ende, int min_s, double greedy,
                              double threshold, rr_clq_rec *cand, rr_clq_rec *hits, unsigned long long cap,
                              unsigned long long *counters /* [2]: candidates, hits */, int n_sm, cudaStream_t st)
{
    if (nq <= 0 || ende <= anfang) return cudaSuccess;
    const int qb = kernel == 2 ? CLQ_QB2 : CLQ_QB;
    const size_t smem = kernel == 3 ? clq_smem_bytes3(W32) : clq_smem_bytes(W32, qb);
    cudaError_t e = kernel == 2   ? cudaFuncSetAttribute(rr_k_cliquer_counts2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                    : kernel == 3 ? cudaFuncSetAttribute(rr_k_cliquer_counts3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                  : cudaFuncSetAttribute(rr_k_cliquer_counts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // blockIdx.x = batch of queries, blockIdx.y = slab of candidate sites: the batches of one slab are adjacent in launch order
    dim3 grid((unsigned)((nq + qb - 1) / qb), (unsigned)((ende - anfang + CLQ_SLAB - 1) / CLQ_SLAB));
    if (kernel == 3)
        rr_k_cliquer_counts3<<<grid, CLQ_WARPS * 32, smem, st>>>(bits, covbits, W32, queries, nq, anfang, ende, min_s, greedy, lnf,
                                                                 cand, cap, counters);
    else if (kernel == 2)
        rr_k_cliquer_counts2<<<grid, CLQ_WARPS * 32, smem, st>>>(bits, covbits, W32, queries, nq, anfang, ende, min_s, greedy, lnf,
                                                                 cand, cap, counters);
    else
        rr_k_cliquer_counts<<<grid, CLQ_WARPS * 32, smem, st>>>(bits, covbits, W32, queries, nq, anfang, ende, min_s, greedy, lnf,
                                                                cand, cap, counters);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    rr_k_cliquer_score<<<n_sm * 8, 128, 0, st>>>(cand, cap, counters, queries, gsize, lnf, threshold, hits, counters + 1);
    rr_count_launch(2);
    return cudaGetLastError();
}
#endif
