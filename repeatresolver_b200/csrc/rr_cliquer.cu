// rr_cliquer.cu -- Cliquer, the inner step of Group_Refinement (/root/reference/RepeatResolver.c:1179-1240), for a
// batch of query groups on the device bitsets of the packed MSA.
//
// The reference calls Cliquer once for every group whose MaxCorrs value is above the cutoff (1647-1649, 1723-1725);
// each call walks all 5 * (ende - anfang) candidate groups with one Schnitt (1213) and, where that intersection is
// above mincov/4 (1215), three more plus the hypergeometric score (Group_PositiveSignificance 472-488).  Here:
//
//   rr_k_cliquer_counts   one block = (batch of CLQ_QB queries, slab of CLQ_SLAB candidate sites).  The queries'
//                         group and coverage bitsets sit in shared memory; a warp takes one candidate site, streams
//                         its bitsets (coverage + 4 groups) 32 words at a time and accumulates, per query, ten AND+POPC
//                         counts (4 x |Gk & Gq|, 4 x |Gk & Cq|, |Gq & Ck|, |Ck & Cq|); the groups of a site partition
//                         its coverage (rr_k_pack_bits), so the fifth group's counts are differences.  Integer work on
//                         HBM-resident bitsets: blocks of one slab are adjacent in launch order (blockIdx.x = batch),
//                         so a slab is read from HBM once and then served from L2 to every batch.  A 32-word chunk is
//                         skipped unless some read is covered by the candidate site AND by one of the block's queries
//                         (rows are in span order, so these are long runs); the coverage word of the next chunk is
//                         requested while the current one is counted.  Pairs that pass 1215 and whose rigorous score
//                         bound (rr_score.h) can exceed `greedy` are appended to a candidate list.  Three blocks per SM
//                         (80 registers).  Measured on a B200 against the two earlier count kernels of round 1 (twelve
//                         counts per word: 18.5 ms; counts split in two passes: 27.6 ms): 14.4 ms for 1024 queries on a
//                         4740 x 26594 MSA, POPC pipe 67 %, issue slots 71 % (profiles/r2_cliquer_ncu_summary.csv).
//   rr_k_cliquer_score    one thread per listed candidate: the exact score in IEEE double, GSL's operation order
//                         (rr_group_significance); candidates above greedy (less a 1e-9 margin) go to the hit list.
//
// The host (rr_cliquer_batch in rr_abi.cu) sorts the hits per query, re-evaluates the few that decide the top
// maxclique-1 with the host libm (the same finalisation as the scan's RR_FLAG_HOST_FINALIZE) and applies
// TheBestUpdater's order (1156-1176).
#include <algorithm>
#include "rr_kernels.h"
#include "rr_score.h"

constexpr int CLQ_QB = RR_CLQ_QB;       // queries per block
constexpr int CLQ_SLAB = RR_CLQ_SLAB;   // candidate sites per block
constexpr int CLQ_WARPS = 8;
constexpr unsigned CLQ_FULL = 0xffffffffu;

static_assert(CLQ_QB * 5 <= 32, "one lane per (query, group of the site) in the tail of the site loop");
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

__global__ void __launch_bounds__(CLQ_WARPS * 32, 3)
rr_k_cliquer_counts(const uint32_t *__restrict__ bits, const uint32_t *__restrict__ covbits, int W32,
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

// ---------------------------------------------------------------------------------------------------------------
// CliqueGroup / CliqueCoverage (/root/reference/RepeatResolver.c:976-1008, 1064-1096), the step of Group_Refinement
// that follows Cliquer (1662-1664): the reads contained in MORE than c of a clique's member groups, and the reads
// covered at more than c of the members' sites.  The reference walks every read and tests it against every member
// (signumber x |clique| GrElement calls per clique); here a thread owns 32 reads = one word of the device bitsets and
// adds the members' words into bit-sliced counters (7 planes: a clique has at most 100 members, 986), then compares the
// counters with c plane by plane - 32 reads per instruction, every member word read once, coalesced over the block.
//   which = 0: member groups (bits), which = 1: the members' sites (covbits)
//   out[clique][W32] in RANK order (the order of the device bitsets); rr_k_rank_bits_to_rows turns it into the
//   reference's layout (read r = bit r % 64 of word r / 64 in MSA row order, GrAdd 211-217).
// HBM-bound integer work: |clique| x W32 words read per clique and kind.
constexpr int CLG_PLANES = 7;
__global__ void __launch_bounds__(128)
rr_k_clique_members(const uint32_t *__restrict__ bits, const uint32_t *__restrict__ covbits, int W32, int64_t n_cliques,
                    const int32_t *__restrict__ members, int stride, const int32_t *__restrict__ n_members,
                    const int32_t *__restrict__ cutoffs, int which, uint32_t *__restrict__ out)
{
    const int64_t q = blockIdx.y;
    if (q >= n_cliques) return;
    const int nm = min(n_members[q], stride);
    const int c = cutoffs[q];
    const int32_t *mem = members + q * stride;
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < W32; w += gridDim.x * blockDim.x) {
        uint32_t cnt[CLG_PLANES];
#pragma unroll
        for (int k = 0; k < CLG_PLANES; k++) cnt[k] = 0u;
        for (int m = 0; m < nm; m++) {
            const int g = mem[m];
            uint32_t x = which ? covbits[(size_t)(g / 5) * W32 + w] : bits[(size_t)g * W32 + w];
#pragma unroll
            for (int k = 0; k < CLG_PLANES; k++) {   // ripple-carry add of one bit per read
                const uint32_t carry = cnt[k] & x;
                cnt[k] ^= x;
                x = carry;
            }
        }
        // count > c, 32 reads at once: the first plane (from the top) where count and c differ decides
        uint32_t gt = 0u, eq = 0xffffffffu;
        if (c < 0) gt = 0xffffffffu;                       // every read, covered or not (986-1003 with c < 0)
        else if (c < (1 << CLG_PLANES) - 1) {
#pragma unroll
            for (int k = CLG_PLANES - 1; k >= 0; k--) {
                const uint32_t cb = (c >> k) & 1 ? 0xffffffffu : 0u;
                gt |= eq & cnt[k] & ~cb;
                eq &= ~(cnt[k] ^ cb);
            }
        }
        out[q * W32 + w] = gt;
    }
}

// Dropoff_Cutoff's member counts (RepeatResolver.c:1472-1486) for a batch of cliques: sizes[q * stride + k] = the reads
// contained in MORE than k of the first n_members[q] members of clique q, k < n_members[q] (at most 100 members, 1463).  One
// block row per clique, the same bit-sliced counters as above (every member word read once), then one comparison per k on 32
// reads at a time; a block sums its words in shared memory and adds one value per k to the result.  sizes must be zero on
// entry.  Bits beyond the last read are zero in every group, so they count nowhere.
__global__ void __launch_bounds__(128)
rr_k_clique_sizes(const uint32_t *__restrict__ bits, int W32, int64_t n_cliques, const int32_t *__restrict__ members, int stride,
                  const int32_t *__restrict__ n_members, uint32_t *__restrict__ sizes)
{
    __shared__ unsigned hist[128];
    const int64_t q = blockIdx.y;
    if (q >= n_cliques) return;                            // block-uniform
    const int nm = min(min(n_members[q], stride), 100);
    const int lane = threadIdx.x & 31;
    const int32_t *mem = members + q * stride;
    hist[threadIdx.x] = 0u;
    __syncthreads();
    // warp-uniform trip count: the reductions below are over whole warps
    for (int w0 = blockIdx.x * blockDim.x + (threadIdx.x - lane); w0 < W32; w0 += gridDim.x * blockDim.x) {
        const int w = w0 + lane;
        uint32_t cnt[CLG_PLANES];
#pragma unroll
        for (int k = 0; k < CLG_PLANES; k++) cnt[k] = 0u;
        if (w < W32)
            for (int m = 0; m < nm; m++) {
                uint32_t x = bits[(size_t)mem[m] * W32 + w];
#pragma unroll
                for (int k = 0; k < CLG_PLANES; k++) {
                    const uint32_t carry = cnt[k] & x;
                    cnt[k] ^= x;
                    x = carry;
                }
            }
        for (int c = 0; c < nm; c++) {                     // count > c; c <= 99 < 2^CLG_PLANES - 1
            uint32_t gt = 0u, eq = 0xffffffffu;
#pragma unroll
            for (int k = CLG_PLANES - 1; k >= 0; k--) {
                const uint32_t cb = (c >> k) & 1 ? 0xffffffffu : 0u;
                gt |= eq & cnt[k] & ~cb;
                eq &= ~(cnt[k] ^ cb);
            }
            const unsigned n = __reduce_add_sync(CLQ_FULL, (unsigned)__popc(gt));
            if (n == 0u) break;                            // warp-uniform; no read of these words is in more than c members
            if (lane == 0) atomicAdd(&hist[c], n);
        }
    }
    __syncthreads();
    if ((int)threadIdx.x < nm && hist[threadIdx.x]) atomicAdd(&sizes[q * stride + threadIdx.x], hist[threadIdx.x]);
}

// rank-order result words -> the reference's bitsets over MSA rows: one warp per 32 consecutive rows of one clique
__global__ void __launch_bounds__(256)
rr_k_rank_bits_to_rows(const uint32_t *__restrict__ in, int W32, int64_t n_cliques, const int32_t *__restrict__ rank_of_row,
                       int R, int words32 /* 2 * sc */, uint32_t *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t q = warp / words32;
    if (q >= n_cliques) return;                            // warp-uniform
    const int w = (int)(warp - q * words32);
    const int row = w * 32 + lane;
    int bit = 0;
    if (row < R) {
        const int r = rank_of_row[row];
        bit = (in[q * W32 + (r >> 5)] >> (r & 31)) & 1u;
    }
    const unsigned word = __ballot_sync(CLQ_FULL, bit);
    if (lane == 0) out[q * words32 + w] = word;
}

// query bitsets [2][CLQ_QB][W32p], the per-chunk query masks and the union of the queries' coverage [W32p]
size_t rr_cliquer_smem_bytes(int W32)
{
    const size_t nchunks = ((size_t)W32 + 31) / 32;
    return ((2 * (size_t)CLQ_QB + 1) * nchunks * 32 + nchunks) * sizeof(uint32_t);
}

#ifndef RR_CPU_EMU   // tests/emu compiles the kernels above with a host compiler; the launch syntax below is nvcc only
cudaError_t rr_launch_cliquer(const uint32_t *bits, const uint32_t *covbits, const int32_t *gsize, const double *lnf,
                              int W32, const int32_t *queries, int nq, int anfang, int ende, int min_s, double greedy,
                              double threshold, rr_clq_rec *cand, rr_clq_rec *hits, unsigned long long cap,
                              unsigned long long *counters /* [2]: candidates, hits */, int n_sm, cudaStream_t st)
{
    if (nq <= 0 || ende <= anfang) return cudaSuccess;
    const size_t smem = rr_cliquer_smem_bytes(W32);
    cudaError_t e = cudaFuncSetAttribute(rr_k_cliquer_counts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // blockIdx.x = batch of queries, blockIdx.y = slab of candidate sites: the batches of one slab are adjacent in launch order
    dim3 grid((unsigned)((nq + CLQ_QB - 1) / CLQ_QB), (unsigned)((ende - anfang + CLQ_SLAB - 1) / CLQ_SLAB));
    rr_k_cliquer_counts<<<grid, CLQ_WARPS * 32, smem, st>>>(bits, covbits, W32, queries, nq, anfang, ende, min_s, greedy, lnf,
                                                            cand, cap, counters);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    rr_k_cliquer_score<<<n_sm * 8, 128, 0, st>>>(cand, cap, counters, queries, gsize, lnf, threshold, hits, counters + 1);
    rr_count_launch(2);
    return cudaGetLastError();
}
// the member counts Dropoff_Cutoff needs: sizes [n_cliques][stride], zeroed here
cudaError_t rr_launch_clique_sizes(const uint32_t *bits, int W32, int64_t n_cliques, const int32_t *members, int stride,
                                   const int32_t *n_members, uint32_t *sizes, cudaStream_t st)
{
    if (n_cliques <= 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(sizes, 0, sizeof(uint32_t) * (size_t)n_cliques * stride, st);
    if (e != cudaSuccess) return e;
    for (int64_t q0 = 0; q0 < n_cliques; q0 += 65535) {   // grid.y limit
        const int64_t nq = std::min<int64_t>(65535, n_cliques - q0);
        dim3 grid((unsigned)std::max(1, std::min((W32 + 127) / 128, 64)), (unsigned)nq);
        rr_k_clique_sizes<<<grid, 128, 0, st>>>(bits, W32, nq, members + q0 * stride, stride, n_members + q0, sizes + q0 * stride);
        rr_count_launch(1);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    return cudaSuccess;
}
// CliqueGroup (which = 0) or CliqueCoverage (which = 1) of n_cliques cliques: tmp [n_cliques][W32] rank-order words,
// out [n_cliques][words32] words of 32 rows each (two per unsigned long of the reference)
cudaError_t rr_launch_clique_members(const uint32_t *bits, const uint32_t *covbits, int W32, int64_t n_cliques,
                                     const int32_t *members, int stride, const int32_t *n_members, const int32_t *cutoffs,
                                     int which, const int32_t *rank_of_row, int R, int words32, uint32_t *tmp, uint32_t *out,
                                     cudaStream_t st)
{
    if (n_cliques <= 0) return cudaSuccess;
    for (int64_t q0 = 0; q0 < n_cliques; q0 += 65535) {   // grid.y limit
        const int64_t nq = std::min<int64_t>(65535, n_cliques - q0);
        dim3 grid((unsigned)std::max(1, std::min((W32 + 127) / 128, 64)), (unsigned)nq);
        rr_k_clique_members<<<grid, 128, 0, st>>>(bits, covbits, W32, nq, members + q0 * stride, stride, n_members + q0,
                                                  cutoffs + q0, which, tmp + q0 * W32);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        const int64_t warps = nq * words32;
        rr_k_rank_bits_to_rows<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(tmp + q0 * W32, W32, nq, rank_of_row, R, words32,
                                                                                  out + q0 * words32);
        rr_count_launch(2);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    return cudaSuccess;
}
#endif
