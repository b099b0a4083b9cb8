/* TEST INFRASTRUCTURE ONLY.  Builds the kernel bodies of csrc/rr_cliquer.cu, rr_relvars.cu and rr_kmeans.cu with a host
 * compiler (see cuda_runtime.h in this directory) and exposes one C entry point per kernel that runs a whole grid, block
 * after block, every CUDA thread a pthread.  tests/test_kernel_emulation.py feeds them and compares with the oracle. */
#define RR_CPU_EMU 1
#include "cuda_runtime.h"
#include <vector>

thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;
thread_local emu_block *emu_blk;

extern "C" void rr_count_launch(int) {}
alignas(16) unsigned char emu_dynamic_smem[256 << 10];
#define RR_DYN_SMEM(type, name) type *name = (type *)emu_dynamic_smem

#include <mutex>
#include "../../repeatresolver_b200/csrc/rr_plan.cpp"           /* the host plan, as rr_scan builds it */
#include "../../repeatresolver_b200/csrc/rr_pack.cu"
#include "../../repeatresolver_b200/csrc/rr_scan_bitset.cu"
#include "../../repeatresolver_b200/csrc/rr_cliquer.cu"
#include "../../repeatresolver_b200/csrc/rr_relvars.cu"
#include "../../repeatresolver_b200/csrc/rr_kmeans.cu"

/* the three PTX helpers of rr_device.cuh (16-byte load, 8-byte load, 16-byte compare-and-swap) behind one lock */
static std::mutex emu_best_lock;
rr_best_t rr_best_load(const rr_best_t *p) { std::lock_guard<std::mutex> g(emu_best_lock); return *p; }
double rr_best_value(const rr_best_t *p) { std::lock_guard<std::mutex> g(emu_best_lock); return __longlong_as_double((long long)p->z); }
bool rr_cas128(rr_best_t *addr, rr_best_t expected, rr_best_t desired, rr_best_t *old)
{
    std::lock_guard<std::mutex> g(emu_best_lock);
    *old = *addr;
    if (old->z != expected.z || old->p != expected.p) return false;
    *addr = desired;
    return true;
}

template <typename F>
struct emu_job { F *body; emu_block *blk; dim3 grid, block; unsigned bx, by, tx; };

template <typename F>
static void *emu_thread(void *x)
{
    emu_job<F> *j = (emu_job<F> *)x;
    threadIdx.x = j->tx; threadIdx.y = threadIdx.z = 0;
    blockIdx.x = j->bx; blockIdx.y = j->by; blockIdx.z = 0;
    blockDim = j->block; gridDim = j->grid;
    emu_blk = j->blk;
    (*j->body)();
    return nullptr;
}

/* Run body() once per CUDA thread.  A thread that leaves the kernel early must not be waited for at a later barrier:
 * the kernels here only return early block-uniformly or after their last barrier. */
template <typename F>
static void emu_launch(dim3 grid, unsigned threads, F body)
{
    const unsigned nwarps = (threads + 31) / 32;
    for (unsigned by = 0; by < grid.y; by++)
        for (unsigned bx = 0; bx < grid.x; bx++) {
            emu_block blk;
            std::vector<emu_warp> warps(nwarps);
            pthread_barrier_init(&blk.bar, nullptr, threads);
            for (unsigned w = 0; w < nwarps; w++) pthread_barrier_init(&warps[w].bar, nullptr, std::min(32u, threads - 32 * w));
            blk.warps = warps.data();
            std::vector<pthread_t> th(threads);
            std::vector<emu_job<F>> jobs(threads);
            pthread_attr_t attr;
            pthread_attr_init(&attr);
            pthread_attr_setstacksize(&attr, 256 << 10);
            for (unsigned t = 0; t < threads; t++) {
                jobs[t] = emu_job<F>{&body, &blk, grid, dim3(threads), bx, by, t};
                pthread_create(&th[t], &attr, emu_thread<F>, &jobs[t]);
            }
            for (unsigned t = 0; t < threads; t++) pthread_join(th[t], nullptr);
            pthread_attr_destroy(&attr);
            pthread_barrier_destroy(&blk.bar);
            for (unsigned w = 0; w < nwarps; w++) pthread_barrier_destroy(&warps[w].bar);
        }
}

extern "C" {

/* The pruning tiers of the fused epilogue (rr_device.cuh) as plain functions, called the way rr_scan_umma.cu calls them:
 * tier 1 on a float copy of the ln n! table with the host-computed rounding margin, tier 2 on the double table.
 * quads: n x {s, gr1, gr2, cov}; best: n running maxima; keep1 / keep2: 1 = the pair survives the tier. */
void emu_tiers(const double *lnf, int max_cov, long long n, const uint32_t *quads, const double *best, unsigned char *keep1,
               unsigned char *keep2, unsigned char *keepq)
{
    std::vector<float> lnf32((size_t)max_cov + 2);
    for (int k = 0; k < max_cov + 2; k++) lnf32[k] = (float)lnf[k];
    const float margin = (float)(16.0 * 5.9604645e-8 * lnf[std::max(max_cov, 1)] * 0.4342944819 + 2e-6);   // rr_scan_umma.cu: U.t1_margin
    auto LT = [&](unsigned k) { return lnf32[k]; };
    const int shift = rr_t1q_shift(lnf[std::max(max_cov + 1, 1)]);
    std::vector<int> lnfq((size_t)max_cov + 2);
    for (int k = 0; k < max_cov + 2; k++) lnfq[k] = (int)nearbyint(lnf[k] * (double)(1 << shift));
    const float qscale = nextafterf((float)(2.302585092994046 * (double)(1 << shift)), 0.0f);
    rr_lnf_global T2{lnf};
    for (long long i = 0; i < n; i++) {
        const unsigned sc = quads[4 * i], gr1 = quads[4 * i + 1], gr2 = quads[4 * i + 2], cov = quads[4 * i + 3];
        const float lnc3 = (LT(cov) - LT(gr1)) - LT(cov - gr1);
        const float meanfac = (1.0f / (float)std::max(cov, 1u)) * (float)gr1;
        keep1[i] = rr_tier1_f32(LT, sc, gr1, gr2, cov, rr_thr_f32(best[i], false), lnc3, meanfac, margin) ? 1 : 0;
        // the form the tcgen05 kernel uses (rr_scan_umma.cu): counts scaled by 4 (= byte offsets into the table), the table in
        // fixed point with the scale um_fill_params chooses, thresholds as integer limits
        auto LTQ = [&](unsigned off) { return lnfq[off >> 2]; };
        const float meanfac_q = (1.0f / (float)std::max(4u * cov, 1u)) * (float)(4u * gr1);
        const int lnc3q = LTQ(4u * cov) - LTQ(4u * gr1) - LTQ(4u * (cov - gr1));
        keepq[i] = rr_tier1_q(LTQ, 4u * sc, 4u * gr1, 4u * gr2, 4u * cov, rr_thr_q(best[i], false, qscale), lnc3q, meanfac_q) ? 1 : 0;
        keep2[i] = rr_tier2(T2, sc, gr1, gr2, cov, best[i]) ? 1 : 0;
    }
}

/* rr_tier2_interval (rr_device.cuh): keep[i] = the pair survives against best[i]; zlb[i] = the lower bound of its score the
 * scan kernel raises the groups' maxima by (0 = none) */
void emu_tier2_interval(const double *lnf, long long n, const uint32_t *quads, const double *best, unsigned char *keep, double *zlb)
{
    rr_lnf_global T2{lnf};
    for (long long i = 0; i < n; i++)
        keep[i] = rr_tier2_interval(T2, quads[4 * i], quads[4 * i + 1], quads[4 * i + 2], quads[4 * i + 3], best[i], zlb + i) ? 1 : 0;
}

/* the packing kernels of rr_pack.cu with the grids their launchers use */
void emu_row_spans(const uint8_t *cells, int R, int N, int codes, int32_t *start, int32_t *end, int32_t *ncov)
{
    if (R > 0) emu_launch(dim3((unsigned)R), 256, [&] { rr_k_row_spans(cells, R, N, codes, start, end, ncov); });
}
/* cells = the rows [row_lo, row_hi) of the MSA only (a slice of a row-sliced pack; 0, R for the whole MSA) */
void emu_pack_bits(const uint8_t *cells, const int32_t *perm, int R, int N, int codes, uint32_t *bits, uint32_t *covbits, int W32,
                   int row_lo, int row_hi)
{
    if (N <= 0 || W32 <= 0) return;
    emu_launch(dim3((unsigned)((N + PK_COLS - 1) / PK_COLS), (unsigned)(W32 / 4)), PK_COLS,
               [&] { rr_k_pack_bits(cells, perm, R, N, codes, bits, covbits, W32, row_lo, row_hi); });
}
void emu_bitset_sizes(const uint32_t *sets, long long nsets, int W32, int32_t *sizes)
{
    if (nsets > 0) emu_launch(dim3((unsigned)((nsets + 7) / 8)), 256, [&] { rr_k_bitset_sizes(sets, nsets, W32, sizes); });
}
void emu_pair_counts(const uint32_t *bits, const uint32_t *covbits, int W32, long long n, const int32_t *gi, const int32_t *gj, int32_t *out)
{
    if (n > 0) emu_launch(dim3((unsigned)((n + 7) / 8)), 256, [&] { rr_k_pair_counts(bits, covbits, W32, n, gi, gj, out); });
}
/* CliqueGroup (which = 0) / CliqueCoverage (which = 1) as rr_launch_clique_members runs them: counts + threshold in rank
 * order, then the reference's row-order words */
void emu_clique_members(const uint32_t *bits, const uint32_t *covbits, int W32, long long n_cliques, const int32_t *members,
                        int stride, const int32_t *n_members, const int32_t *cutoffs, int which, const int32_t *rank_of_row, int R,
                        int words32, uint32_t *tmp, uint32_t *out)
{
    if (n_cliques <= 0) return;
    emu_launch(dim3(2, (unsigned)n_cliques), 128, [&] {
        rr_k_clique_members(bits, covbits, W32, n_cliques, members, stride, n_members, cutoffs, which, tmp);
    });
    emu_launch(dim3((unsigned)((n_cliques * words32 * 32 + 255) / 256)), 256, [&] {
        rr_k_rank_bits_to_rows(tmp, W32, n_cliques, rank_of_row, R, words32, out);
    });
}
/* Dropoff_Cutoff's member counts as rr_launch_clique_sizes runs them (sizes zeroed by the caller) */
void emu_clique_sizes(const uint32_t *bits, int W32, long long n_cliques, const int32_t *members, int stride,
                      const int32_t *n_members, uint32_t *sizes, int grid_x)
{
    if (n_cliques <= 0) return;
    emu_launch(dim3((unsigned)grid_x, (unsigned)n_cliques), 128, [&] {
        rr_k_clique_sizes(bits, W32, n_cliques, members, stride, n_members, sizes);
    });
}
void emu_general_break(const uint32_t *covbits, int W32, int N, int mincov, int32_t *breakcol)
{
    if (N > 0) emu_launch(dim3((unsigned)((N + 7) / 8)), 256, [&] { rr_k_general_break(covbits, W32, N, mincov, breakcol); });
}
void emu_bits_to_operand(const uint32_t *bits, long long nsets, int W32, int8_t *xb, long long Kp, int fp4)
{
    if (nsets > 0) emu_launch(dim3((unsigned)((nsets + 7) / 8)), 256, [&] { rr_k_bits_to_operand(bits, nsets, W32, xb, Kp, fp4); });
}

/* The AND+POPC variant of the scan, set up as rr_scan (rr_abi.cu) sets it up: plan from the spans in rank order, then
 * rr_k_scan_bitset over `blocks` persistent blocks.  bits/gsize/coverage in rank order as rr_pack leaves them; best:
 * [5N] {value bits, partner}; counters: [8]; returns the plan's pair count (what the device count must equal). */
long long emu_scan_bitset(int R, int N, int W32, int mincov, unsigned flags, const uint32_t *bits, const int32_t *gsize,
                          const int32_t *coverage, const int32_t *breakcol, const int32_t *start, const int32_t *end,
                          const int32_t *class_start, int n_classes, const double *lnfact, rr_best_t *best, unsigned long long *counters, int blocks,
                          int part_index, int part_count)
{
    rr_plan plan;
    rr_plan_build(plan, R, N, mincov, gsize, coverage, breakcol, start, end, class_start, n_classes, rr_bitset_ti(), rr_bitset_tj(), 32, 8, 0,
                  part_index, part_count);
    rr_scan_params P;
    memset(&P, 0, sizeof P);
    P.R = R; P.N = N; P.W32 = W32; P.mincov = mincov; P.flags = flags;
    P.bits = bits; P.gsize = gsize; P.rowok = plan.rowok.data(); P.colok = plan.colok.data();
    P.breakcol = breakcol; P.rowsites = plan.rowsites.data(); P.n_rowsites = plan.n_rowsites;
    P.lnfact = lnfact; P.best = best; P.counters = counters;
    P.unit_prefix = plan.unit_prefix.data(); P.unit_cb0 = plan.unit_cb0.data(); P.n_rowblocks = plan.n_rowblocks;
    P.rb_lo = plan.rb_lo; P.rb_hi = plan.rb_hi; P.word_hi = plan.k_hi.data(); P.word_lo = plan.k_lo.data();
    P.n_colblocks = plan.n_colblocks;
    P.n_classes = plan.n_classes;
    if (plan.rb_hi > plan.rb_lo && plan.unit_prefix[plan.rb_hi] > plan.unit_prefix[plan.rb_lo])
        emu_launch(dim3((unsigned)blocks), BS_TI * 32, [&] { rr_k_scan_bitset(P); });
    return (long long)plan.part_pairs;
}

int emu_clq_qb(void) { return CLQ_QB; }
int emu_clq_slab(void) { return CLQ_SLAB; }

/* rr_launch_cliquer's grid: the count kernel, then the score kernel */
int emu_cliquer(const uint32_t *bits, const uint32_t *covbits, const int32_t *gsize, const double *lnf, int W32,
                const int32_t *queries, int nq, int anfang, int ende, int min_s, double greedy, double threshold,
                rr_clq_rec *cand, rr_clq_rec *hits, unsigned long long cap, unsigned long long *counters)
{
    const size_t smem = rr_cliquer_smem_bytes(W32);
    if (smem > sizeof(emu_dynamic_smem) || nq <= 0 || ende <= anfang) return 1;
    dim3 grid((unsigned)((nq + CLQ_QB - 1) / CLQ_QB), (unsigned)((ende - anfang + CLQ_SLAB - 1) / CLQ_SLAB));
    emu_launch(grid, CLQ_WARPS * 32, [&] { rr_k_cliquer_counts(bits, covbits, W32, queries, nq, anfang, ende, min_s, greedy, lnf, cand, cap, counters); });
    emu_launch(dim3(3), 128, [&] { rr_k_cliquer_score(cand, cap, counters, queries, gsize, lnf, threshold, hits, counters + 1); });
    return 0;
}

void emu_masked_sizes(const uint32_t *bits, const uint32_t *umask, long long nsets, int W32, int32_t *sizes)
{
    if (nsets > 0) emu_launch(dim3((unsigned)((nsets + 7) / 8)), 256, [&] { rr_k_masked_sizes(bits, umask, nsets, W32, sizes); });
}

int emu_relvars_pairs(const uint32_t *bits, const uint32_t *umask, int W32, const int32_t *sel, int nsel, const int32_t *first_partner, const int32_t *gsize_u,
                      int cov_u, const double *lnf, double cutoff, unsigned char *mark, int4 *unsure, unsigned unsure_cap,
                      unsigned *unsure_count)
{
    if (nsel <= 0) return 0;
    const unsigned nt = (unsigned)((nsel + RV_TILE - 1) / RV_TILE);
    emu_launch(dim3(nt, nt), 256, [&] { rr_k_relvars_pairs(bits, umask, W32, sel, nsel, first_partner, gsize_u, cov_u, lnf, cutoff, mark, unsure, unsure_cap, unsure_count); });
    return 0;
}

/* the signatures of a part's reads as rr_launch_kmeans_signatures makes them */
void emu_kmeans_signatures(const uint8_t *rows, int cols, int codes, const int32_t *vars, int n_vars, int anzahl, int scv, uint64_t *sig)
{
    if (anzahl <= 0) return;
    emu_launch(dim3((unsigned)((2 * scv + 7) / 8), (unsigned)anzahl), 256, [&] {
        rr_k_km_signatures(rows, cols, codes, vars, n_vars, anzahl, scv, (uint32_t *)sig);
    });
}
/* one panel of pair scores as km_launch_pair_scores runs it */
static void emu_pair_scores(const uint64_t *sig, const uint64_t *X, const int32_t *xrows, int j0, int nj, int anzahl, int scv, int32_t *out,
                            long long stride_i, long long stride_j)
{
    if (anzahl <= 0 || nj <= 0) return;
    emu_launch(dim3((unsigned)((anzahl + KM_TILE - 1) / KM_TILE), (unsigned)((nj + KM_TILE - 1) / KM_TILE)), 256, [&] {
        rr_k_km_pair_scores(sig, X, xrows, j0, nj, anzahl, scv, out, stride_i, stride_j);
    });
}
/* the score table of the dissolution as rr_launch_kmeans_scores fills it */
void emu_kmeans_scores(const uint64_t *sig, const uint64_t *cen, const int32_t *J, int nJ, int anzahl, int scv, int32_t *S)
{
    emu_pair_scores(sig, cen, J, 0, nJ, anzahl, scv, S, nJ, 1);
}
/* rr_launch_kmeans_sweeps with a caller-chosen panel size, so that several panels per sweep are exercised on small inputs */
int emu_kmeans_sweeps(const uint64_t *sig, int anzahl, int scv, int panel_rows, int32_t *best_j, uint64_t *cen, int32_t *cluster)
{
    if (anzahl <= 0 || panel_rows < 1) return 1;
    std::vector<int32_t> panel((size_t)panel_rows * anzahl), state((size_t)6 * anzahl, 0);
    int32_t *best_s = state.data(), *best = state.data() + (size_t)5 * anzahl;
    for (int i = 0; i < 5 * anzahl; i++) best_j[i] = 0;
    for (int i = 0; i < anzahl; i++) cluster[i] = 0;
    const unsigned nb = (unsigned)((anzahl + 127) / 128);
    for (int j0 = 0; j0 < anzahl; j0 += panel_rows) {
        const int nj = std::min(panel_rows, anzahl - j0);
        emu_pair_scores(sig, sig, nullptr, j0, nj, anzahl, scv, panel.data(), 1, anzahl);
        emu_launch(dim3(nb), 128, [&] { rr_k_km_top5_seq(panel.data(), anzahl, j0, nj, best_s, best_j); });
    }
    const long long nw = (long long)anzahl * scv;
    emu_launch(dim3((unsigned)((nw + 255) / 256)), 256, [&] { rr_k_km_centroids(sig, best_j, anzahl, scv, cen); });
    for (int j0 = 0; j0 < anzahl; j0 += panel_rows) {
        const int nj = std::min(panel_rows, anzahl - j0);
        emu_pair_scores(sig, cen, nullptr, j0, nj, anzahl, scv, panel.data(), 1, anzahl);
        emu_launch(dim3(nb), 128, [&] { rr_k_km_assign_seq(panel.data(), anzahl, j0, nj, best, cluster); });
    }
    return 0;
}

}
