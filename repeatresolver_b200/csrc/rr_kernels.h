// rr_kernels.h -- launchers of the CUDA kernels (internal).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "rr_host.h"

struct rr_scan_params;
extern "C" void rr_count_launch(int n);

// the dynamic shared memory of a kernel (tests/emu substitutes a host buffer when it compiles kernel bodies on the CPU)
#ifndef RR_DYN_SMEM
#define RR_DYN_SMEM(type, name) extern __shared__ type name[]
#endif

// the reference's character classes (MaxCorrelation.c:304-329): aA->0 cC->1 gG->2 tT->3 '-' '_'->4, all else -> 5 (not covered)
#if defined(__CUDACC__) || defined(RR_CPU_EMU)
static __device__ __forceinline__ int rr_classify(unsigned int c, int codes)
{
    if (codes) return c < 5u ? (int)c : 5;
    unsigned int l = c | 0x20u;
    if (l == 'a') return 0;
    if (l == 'c') return 1;
    if (l == 'g') return 2;
    if (l == 't') return 3;
    if (c == '-' || c == '_') return 4;
    return 5;
}
#endif

// Device memory comes from the device's stream-ordered pool (cudaMallocAsync) with an unlimited release
// threshold: packing the next MSA reuses the previous one's blocks instead of paying cudaMalloc/cudaFree of
// multi-GB buffers (measured: 40-160 ms per pack/free cycle at config 2).  All work of a handle is on one
// stream; the ABI entry points select it with rr_alloc_stream() before allocating or freeing.
void rr_alloc_stream(cudaStream_t st);
cudaError_t rr_dev_malloc(void **p, size_t bytes);
void rr_dev_free(void *p);
struct rr_umma_plan;

cudaError_t rr_launch_row_spans(const uint8_t *cells, int R, int N, int codes, int32_t *start, int32_t *end,
                                int32_t *ncov, cudaStream_t st);
cudaError_t rr_launch_pack_bits(const uint8_t *cells /* rows [row_lo, row_hi) */, const int32_t *perm, int R, int N, int codes, uint32_t *bits,
                                uint32_t *covbits, int W32, int row_lo, int row_hi, cudaStream_t st);
cudaError_t rr_launch_bitset_sizes(const uint32_t *sets, int64_t nsets, int W32, int32_t *sizes, cudaStream_t st);
cudaError_t rr_launch_pair_counts(const uint32_t *bits, const uint32_t *covbits, int W32, int64_t n,
                                  const int32_t *gi, const int32_t *gj, int32_t *out, cudaStream_t st);
cudaError_t rr_launch_general_break(const uint32_t *covbits, int W32, int N, int mincov, int32_t *breakcol,
                                    cudaStream_t st);
cudaError_t rr_launch_bits_to_operand(const uint32_t *bits, int64_t nsets, int W32, int8_t *xb, int64_t Kp, int fp4, cudaStream_t st);
cudaError_t rr_launch_or_words(void *dst, const void *src, size_t bytes /* multiple of 16 */, cudaStream_t st);

// Cliquer (rr_cliquer.cu): one listed (query slot, candidate group) pair with its four counts and, once scored, Z
#define RR_CLQ_QB 4      // queries per block of rr_k_cliquer_counts
#define RR_CLQ_SLAB 256  // candidate sites per block
struct rr_clq_rec {
    int32_t slot, group, s, gr1, gr2, cov;
    double z;
};
size_t rr_cliquer_smem_bytes(int W32);
cudaError_t rr_launch_cliquer(const uint32_t *bits, const uint32_t *covbits, const int32_t *gsize, const double *lnf,
                              int W32, const int32_t *queries, int nq, int anfang, int ende, int min_s, double greedy,
                              double threshold, rr_clq_rec *cand, rr_clq_rec *hits, unsigned long long cap,
                              unsigned long long *counters, int n_sm, cudaStream_t st);
// CliqueGroup / CliqueCoverage of a batch of cliques (rr_cliquer.cu)
cudaError_t rr_launch_clique_members(const uint32_t *bits, const uint32_t *covbits, int W32, int64_t n_cliques,
                                     const int32_t *members, int stride, const int32_t *n_members, const int32_t *cutoffs,
                                     int which, const int32_t *rank_of_row, int R, int words32, uint32_t *tmp, uint32_t *out,
                                     cudaStream_t st);
// Dropoff_Cutoff's member counts of a batch of cliques (rr_cliquer.cu): sizes[q][k] = reads in more than k of the members
cudaError_t rr_launch_clique_sizes(const uint32_t *bits, int W32, int64_t n_cliques, const int32_t *members, int stride,
                                   const int32_t *n_members, uint32_t *sizes, cudaStream_t st);

// Relative_Vars (rr_relvars.cu): the all-pairs step on the packed rows of one part
cudaError_t rr_launch_masked_sizes(const uint32_t *bits, const uint32_t *umask, int64_t nsets, int W32, int32_t *sizes, cudaStream_t st);
cudaError_t rr_launch_relvars_pairs(const uint32_t *bits, const uint32_t *umask /* NULL: bits are the part's own rows */, int W32, const int32_t *sel, int nsel, const int32_t *first_partner,
                                    const int32_t *gsize_u, int cov_u, const double *lnf, double cutoff, unsigned char *mark,
                                    int4 *unsure, unsigned int unsure_cap, unsigned int *unsure_count, cudaStream_t st);

// Kmeans (rr_kmeans.cu): the two read x read sweeps and the centroids on the part's signatures
// panel: device room for panel_rows x anzahl int32 scores (any panel_rows >= 1: the sweeps go panel by panel); state: 6 x anzahl int32
cudaError_t rr_launch_kmeans_sweeps(const uint64_t *sig, int anzahl, int scv, int32_t *best_j, uint64_t *cen, int32_t *cluster,
                                    int32_t *panel, int panel_rows, int32_t *state, cudaStream_t st);
cudaError_t rr_launch_kmeans_scores(const uint64_t *sig, const uint64_t *cen, const int32_t *J, int nJ, int anzahl, int scv, int32_t *S,
                                    cudaStream_t st);
// the signatures of the part's reads from their rows on the device: rows [anzahl][cols], sig [anzahl][scv] (every word written)
cudaError_t rr_launch_kmeans_signatures(const uint8_t *rows, int cols, int codes, const int32_t *vars, int n_vars, int anzahl,
                                        int scv, uint64_t *sig, cudaStream_t st);

struct rr_best_t;
cudaError_t rr_launch_init_best(rr_best_t *best, int64_t n, cudaStream_t st);
cudaError_t rr_launch_raise_best(rr_best_t *best, const double *thr, int64_t n, cudaStream_t st);
cudaError_t rr_launch_best_values(const rr_best_t *best, double *values, int64_t n, cudaStream_t st);
cudaError_t rr_launch_scan_bitset(const rr_scan_params &P, int n_sm, cudaStream_t st);
int rr_bitset_ti(void);
int rr_bitset_tj(void);
