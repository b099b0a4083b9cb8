/* rr_debug.h -- test and measurement hooks of librr_maxcorr.so.  NOT part of the drop-in boundary
 * (include/rr_maxcorr.h): nothing here is needed to run the scan, and results obtained with a debug flag set are
 * not results.  Used by tests/ and tools/ only.
 */
#ifndef RR_DEBUG_H
#define RR_DEBUG_H

#include <stdint.h>
#include "rr_maxcorr.h"

#ifdef __cplusplus
extern "C" {
#endif

/* rr_scan_opts.flags, tcgen05 variants only: timing decomposition of the fused kernel (tools/probe_umma.py).
 * A scan with it leaves no maxima behind and rr_scan skips its pair-count check. */
#define RR_DEBUG_MMA_ONLY 0x400u       /* producer + MMA pipeline only: the epilogue releases every accumulator unread */

/* capacity (records) of the candidate / hit lists of rr_cliquer_batch; 0 restores the default.  Tests lower it to
 * force the list-overflow retry path. */
void rr_debug_set_cliquer_cap(unsigned long long cap);

/* capacity (candidates of 20 bytes) of the list in which the tcgen05 scan leaves the pairs whose exact score it evaluates
 * after the scan (csrc/rr_device.cuh: defer_mode); 0 restores the sizing from the plan.  Tests lower it so that the list
 * overflows and the kernel's in-place evaluation takes over in the middle of a scan. */
void rr_debug_set_deferred_cap(unsigned long long cap);

/* The raw accumulator of one (row tile, column tile) pair of the tcgen05 scan kernel, read back from TMEM by the
 * kernel's own epilogue after its own producer / MMA code (a separate instantiation of the same kernel template that
 * also stores what it reads).  Call after rr_scan(pk, opts) with a tcgen05 variant and the same opts.
 *   counts[128][240]  entry [r][c] = |G_row(r) & G_col(c)|  (Schnitt, MaxCorrelation.c:114-125)
 *   row_groups[128]   group id 5 * site + k of accumulator row r, -1 for padding rows (their counts are 0)
 *   col_groups[240]   group id of accumulator column c, -1 beyond the last site of the MSA
 * n_row_tiles / n_col_tiles (may be NULL) receive the plan's tile counts. */
int rr_debug_umma_counts(rr_packed *pk, const rr_scan_opts *opts, int row_tile, int col_tile, int32_t *counts,
                         int32_t *row_groups, int32_t *col_groups, int *n_row_tiles, int *n_col_tiles);

/* The tensor pipe's rate for the MMA kind, tile shape (M = 128, N = 240) and operand layout of RR_VARIANT_UMMA*: a
 * bare loop of back-to-back tcgen05.mma on every SM, nothing loaded, nothing read back.  *macs = multiply-accumulates
 * of one launch, *best_ms = fastest of reps launches (CUDA events); peak = 2 * macs / best_ms.  bench.py reports the
 * scan's executed MACs against this. */
int rr_debug_mma_peak(int device, int variant, int kblocks_per_sm, int reps, float *best_ms, double *macs);
/* the same loop for other issue forms: cta_group 1 (M = 128, n_cols = 240) or cta_group 2 (a CTA pair, M = 256, n_cols a
 * multiple of 16 in 32..256, each CTA holding half of the B tile) */
int rr_debug_mma_peak_shape(int device, int variant, int cta_group, int n_cols, int kblocks_per_sm, int reps, float *best_ms,
                            double *macs);

/* rr_kmeans_finish through the score table rr_kmeans fills on the device (reads x clusters that can ever hold two reads
 * during the dissolution, RepeatResolver.c:2727-2752), here filled on the host: must give what rr_kmeans_finish gives. */
int rr_debug_kmeans_finish_table(int anzahl, int scv, const uint64_t *sig, const uint64_t *cen, const int32_t *cluster_in,
                                 int mingroup, int32_t *cluster_out, int *n_clusters);

#ifdef __cplusplus
}
#endif
#endif /* RR_DEBUG_H */
