// rr_plan.h -- host-side scan plan (internal): admissibility flags, row-site compaction,
// tile ranges, exact contraction ranges and the pair-balanced multi-GPU partition.
// All of it is O(N + R) integer work derived from the filters of
// /root/reference/MaxCorrelation.c:796-817.
#pragma once
#include <stdint.h>
#include <algorithm>
#include <vector>
#include "rr_host.h"

struct rr_plan {
    std::vector<uint8_t> rowok, colok;  // [5N]  (802) / (817)
    std::vector<int32_t> rowsites;      // sites with >= 1 admissible row group, ascending, padded with -1
    int n_rowsites = 0, n_rowblocks = 0, n_colblocks = 0;
    std::vector<int64_t> unit_prefix;   // [n_rowblocks+1] prefix sum of column blocks per row block
    std::vector<int32_t> unit_cb0;      // [n_rowblocks] first column block
    // contraction ranges, one per LENGTH CLASS of rows (class c = ranks [class_start[c], class_start[c + 1]), each class
    // sorted by span start; boundaries are multiples of 256 rows): the k-units of class c that can contribute to (row block
    // rb, column block cb) are [k_lo[c * n_colblocks + cb], k_hi[c * n_rowblocks + rb]) - empty when lo >= hi
    int n_classes = 1;
    std::vector<int32_t> k_hi;          // [n_classes][n_rowblocks] exclusive upper bound of contributing k-units
    std::vector<int32_t> k_lo;          // [n_classes][n_colblocks] inclusive lower bound
    int kunits(int rb, int cb) const
    {
        int sum = 0;
        for (int c = 0; c < n_classes; c++) {
            const int a = k_hi[(size_t)c * std::max(n_rowblocks, 1) + rb] - k_lo[(size_t)c * std::max(n_colblocks, 1) + cb];
            sum += a > 0 ? a : 0;
        }
        return sum;
    }
    std::vector<int64_t> rb_pairs;      // [n_rowblocks] pair tests per row block
    int rb_lo = 0, rb_hi = 0;           // this part's row blocks
    int part_index = 0, part_count = 1;
    int max_cov = 0;                    // largest column coverage (bounds every ln(n!) argument)
    int64_t total_pairs = 0, part_pairs = 0;
    int64_t part_kunits = 0;            // sum over this part's units of contributing k-units
    int64_t executed_ops = 0;           // filled by the variant
};

// ti / tj: row / column sites per tile; kunit: rows (reads) per contraction unit
// (32 = one u32 word for the bitset kernel, the K block for the tcgen05 kernel); tile_cost: cost of a tile's
// epilogue in k-unit equivalents, overlap_pct: how much of the smaller of (epilogue, contraction) hides behind the
// larger one, 0 = none (multi-GPU balance).
// start/end: spans in rank order, or NULL when rows are not single spans (no skipping: one class over all rows);
// class_start[n_classes + 1]: rank boundaries of the length classes (rr_length_classes), each class sorted by span start.
constexpr int RR_MAX_CLASSES = 8;
// the rule rr_pack applies (rr_length_classes): the rows sorted by span length are cut at these cumulative fractions
// (measured on the bench workloads: one class 3.25e7 K blocks at config 2, {3/4, 1/4} 2.29e7, these four 2.05e7; more classes
// lose to the rounding of every class range to whole K blocks)
#define RR_LENGTH_CLASSES 4
#define RR_LENGTH_CLASS_FRACTIONS {0.4, 0.3, 0.2, 0.1}
void rr_plan_build(rr_plan &plan, int R, int N, int mincov, const int32_t *gsize, const int32_t *coverage,
                   const int32_t *breakcol, const int32_t *start, const int32_t *end, const int32_t *class_start, int n_classes,
                   int ti, int tj, int kunit, int tile_cost, int overlap_pct, int part_index, int part_count);

// ---- tcgen05 variant hooks (rr_scan_umma.cu) ---------------------------------------------
struct rr_umma_state;
struct rr_scan_params;
int rr_umma_available(void);
int rr_umma_row_sites(void);
int rr_umma_col_sites(void);
int rr_umma_kblock(int mode);
void rr_umma_free(rr_umma_state *s);
int rr_umma_scan(rr_umma_state *&s, int mode /* 0 int8, 1 e2m1 f8f6f4, 2 e2m1 mxf4 */, uint64_t plan_id, rr_scan_params &P, rr_plan &plan,
                 int n_sm, cudaStream_t st);
int rr_umma_dump_tile(rr_umma_state *s, int mode, uint64_t plan_id, rr_scan_params &P, rr_plan &plan, int rt, int ct,
                      int32_t *d_out /* device, [128][240] */, int n_sm, cudaStream_t st);
cudaError_t rr_umma_mma_peak(int mode, int n_sm, int kblocks_per_sm, double *macs, cudaStream_t st);
cudaError_t rr_umma_mma_peak_pair(int mode, int n_sm, int kblocks_per_sm, int n_cols, double *macs, cudaStream_t st);
