// rr_plan.cpp -- see rr_plan.h
#include <algorithm>
#include <cuda_runtime.h>
#include "rr_plan.h"

void rr_plan_build(rr_plan &plan, int R, int N, int mincov, const int32_t *gsize, const int32_t *coverage,
                   const int32_t *breakcol, const int32_t *start, const int32_t *end, int ti, int tj, int kunit,
                   int tile_cost, int overlap_pct, int part_index, int part_count)
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

    // contraction range bounds from the row order (rows sorted by span start)
    const int kunits_all = (R + kunit - 1) / kunit;
    std::vector<int32_t> minrank_end_ge;  // [N+1]
    if (start && end) {
        minrank_end_ge.assign((size_t)N + 1, R);
        int maxend = -1;
        for (int r = 0; r < R; r++) {
            if (end[r] > maxend) {
                for (int j = maxend + 1; j <= end[r] && j <= N; j++) minrank_end_ge[j] = r;
                maxend = end[r];
            }
        }
    }
    plan.k_lo.assign(std::max(plan.n_colblocks, 1), 0);
    for (int cb = 0; cb < plan.n_colblocks; cb++)
        plan.k_lo[cb] = (start && end) ? minrank_end_ge[(size_t)cb * tj] / kunit : 0;

    plan.unit_prefix.assign((size_t)plan.n_rowblocks + 1, 0);
    plan.unit_cb0.assign(std::max(plan.n_rowblocks, 1), 0);
    plan.k_hi.assign(std::max(plan.n_rowblocks, 1), 0);
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
        if (start && end) {
            const int p = (int)(std::upper_bound(start, start + R, ii_max) - start);
            plan.k_hi[rb] = (p + kunit - 1) / kunit;
        } else {
            plan.k_hi[rb] = kunits_all;
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
            const int kb = std::max(0, plan.k_hi[rb] - plan.k_lo[plan.unit_cb0[rb] + k]);
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
        for (int c = 0; c < ncb; c++)
            plan.part_kunits += std::max(0, plan.k_hi[rb] - plan.k_lo[plan.unit_cb0[rb] + c]);
    }
    plan.executed_ops = plan.part_kunits * (int64_t)(25 * ti * tj);
}
