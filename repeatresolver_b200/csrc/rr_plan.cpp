// rr_plan.cpp -- see rr_plan.h
#include <algorithm>
#include <cuda_runtime.h>
#include "rr_plan.h"

void rr_plan_build(rr_plan &plan, int R, int N, int mincov, const int32_t *gsize, const int32_t *coverage,
                   const int32_t *breakcol, const int32_t *start, const int32_t *end, const int32_t *class_start, int n_classes,
                   int ti, int tj, int kunit, int tile_cost, int overlap_pct, int part_index, int part_count)
{
    const int q = mincov / 4;  // integer division, MaxCorrelation.c:802/817
    const size_t G = (size_t)5 * N;
    plan.rowok.assign(G, 0);
    plan.colok.assign(G, 0);
    std::vector<int32_t> colok_prefix((size_t)N + 1, 0);
    plan.rowsites.clear();
    std::vector<int32_t> nrow;
    plan.max_cov = 0;
    for (int ii = 0; ii < N; ii++) {
        const int32_t *gs = gsize + (size_t)5 * ii;
        const int baseno = gs[0] + gs[1] + gs[2] + gs[3];            // 798
        const bool basey = baseno > coverage[ii] / 2;                // 802
        plan.max_cov = std::max(plan.max_cov, (int)coverage[ii]);
        int nr = 0, nc = 0;
        for (int k = 0; k < 5; k++) {
            const bool sz = gs[k] > q && gs[k] < R;                  // 802 / 817 (maxgroup = signumber, 1008)
            plan.colok[(size_t)5 * ii + k] = sz;
            plan.rowok[(size_t)5 * ii + k] = sz && basey;
            nc += sz; nr += sz && basey;
        }
        colok_prefix[ii + 1] = colok_prefix[ii] + nc;
        if (nr > 0 && ii + 20 < N && breakcol[ii] > ii + 20) { plan.rowsites.push_back(ii); nrow.push_back(nr); }
    }
    plan.n_rowsites = (int)plan.rowsites.size();
    plan.n_rowblocks = (plan.n_rowsites + ti - 1) / ti;
    plan.n_colblocks = (N + tj - 1) / tj;
    plan.rowsites.resize((size_t)std::max(plan.n_rowblocks, 1) * ti, -1);

    // contraction range bounds from the row order.  Rows come in length classes, each sorted by span start: in a class
    // whose longest row spans L columns the rows that can reach a column block start at most L before it, so keeping the
    // long rows (an MSA of long reads always has rows spanning most of it) out of the classes of the short ones keeps the
    // ranges of the short ones tight.
    const int kunits_all = (R + kunit - 1) / kunit;
    const int nrb = std::max(plan.n_rowblocks, 1), ncb_all = std::max(plan.n_colblocks, 1);
    const bool spans = start && end;
    const int ncls = spans ? std::min(std::max(n_classes, 1), RR_MAX_CLASSES) : 1;
    plan.n_classes = ncls;
    plan.k_lo.assign((size_t)ncls * ncb_all, 0);
    plan.k_hi.assign((size_t)ncls * nrb, 0);
    std::vector<int> cls_r0(ncls, 0), cls_r1(ncls, R);
    if (spans)
        for (int c = 0; c < ncls; c++) {
            cls_r0[c] = std::min(std::max(class_start ? class_start[c] : 0, 0), R);
            cls_r1[c] = std::min(std::max(class_start ? class_start[c + 1] : R, cls_r0[c]), R);
        }
    if (spans) {
        std::vector<int32_t> minrank_end_ge((size_t)N + 1);
        for (int c = 0; c < ncls; c++) {
            const int r0 = cls_r0[c], r1 = cls_r1[c];
            std::fill(minrank_end_ge.begin(), minrank_end_ge.end(), r1);   // none: the empty range at the class's end
            int maxend = -1;
            for (int r = r0; r < r1; r++) {
                if (end[r] > maxend) {
                    for (int j = maxend + 1; j <= end[r] && j <= N; j++) minrank_end_ge[j] = r;
                    maxend = end[r];
                }
            }
            for (int cb = 0; cb < plan.n_colblocks; cb++) {
                const int first = minrank_end_ge[(size_t)cb * tj];
                plan.k_lo[(size_t)c * ncb_all + cb] = first >= r1 ? (r1 + kunit - 1) / kunit : first / kunit;
            }
        }
    }

    plan.unit_prefix.assign((size_t)plan.n_rowblocks + 1, 0);
    plan.unit_cb0.assign(std::max(plan.n_rowblocks, 1), 0);
    plan.rb_pairs.assign(std::max(plan.n_rowblocks, 1), 0);
    plan.total_pairs = 0;
    for (int rb = 0; rb < plan.n_rowblocks; rb++) {
        int ii_min = -1, ii_max = -1, maxbreak = 0;
        int64_t pairs = 0;
        for (int t = 0; t < ti; t++) {
            const int idx = rb * ti + t;
            const int ii = plan.rowsites[idx];
            if (ii < 0) continue;
            if (ii_min < 0) ii_min = ii;
            ii_max = ii;
            const int brk = std::min(breakcol[ii], N);
            maxbreak = std::max(maxbreak, brk);
            pairs += (int64_t)nrow[idx] * (colok_prefix[brk] - colok_prefix[ii + 20]);
        }
        const int cb0 = (ii_min + 20) / tj;
        const int cb1 = (maxbreak + tj - 1) / tj;
        plan.unit_cb0[rb] = cb0;
        plan.unit_prefix[rb + 1] = plan.unit_prefix[rb] + std::max(0, cb1 - cb0);
        plan.rb_pairs[rb] = pairs;
        plan.total_pairs += pairs;
        if (spans) {
            for (int c = 0; c < ncls; c++) {
                const int r0 = cls_r0[c], r1 = cls_r1[c];
                const int p = (int)(std::upper_bound(start + r0, start + r1, ii_max) - start);   // ranks [r0, p) start <= ii_max
                plan.k_hi[(size_t)c * nrb + rb] = p > r0 ? (p + kunit - 1) / kunit : r0 / kunit;
            }
        } else {
            plan.k_hi[rb] = kunits_all;   // no skipping: everything as one range from k-unit 0
        }
    }

    // cost-balanced contiguous partition of the row blocks: a tile costs tile_cost (the epilogue walks the whole
    // tile) and one per contributing k-unit (the contraction).  The bitset kernel does one after the other; in the
    // tcgen05 kernel the MMA pipeline runs beside the epilogue, so a tile costs the larger of the two plus the part
    // of the smaller one that does not hide (fitted on B200 part by part).
    std::vector<int64_t> rb_cost(std::max(plan.n_rowblocks, 1), 0);
    int64_t total_cost = 0;
    for (int rb = 0; rb < plan.n_rowblocks; rb++) {
        const int ncb = (int)(plan.unit_prefix[rb + 1] - plan.unit_prefix[rb]);
        int64_t c = 0;
        for (int k = 0; k < ncb; k++) {
            const int kb = plan.kunits(rb, plan.unit_cb0[rb] + k);
            c += 100 * (int64_t)std::max(tile_cost, kb) + (int64_t)(100 - overlap_pct) * std::min(tile_cost, kb);
        }
        rb_cost[rb] = c;
        total_cost += c;
    }
    std::vector<int> cut((size_t)part_count + 1, plan.n_rowblocks);
    cut[0] = 0;
    {
        int64_t acc = 0;
        int p = 1;
        for (int rb = 0; rb < plan.n_rowblocks && p < part_count; rb++) {
            acc += rb_cost[rb];
            while (p < part_count && acc * part_count >= total_cost * (int64_t)p) cut[p++] = rb + 1;
        }
    }
    plan.part_index = part_index;
    plan.part_count = part_count;
    plan.rb_lo = cut[part_index];
    plan.rb_hi = cut[part_index + 1];
    plan.part_pairs = 0;
    plan.part_kunits = 0;
    for (int rb = plan.rb_lo; rb < plan.rb_hi; rb++) {
        plan.part_pairs += plan.rb_pairs[rb];
        const int ncb = (int)(plan.unit_prefix[rb + 1] - plan.unit_prefix[rb]);
        for (int c = 0; c < ncb; c++) plan.part_kunits += plan.kunits(rb, plan.unit_cb0[rb] + c);
    }
    plan.executed_ops = plan.part_kunits * (int64_t)(25 * ti * tj);
}
