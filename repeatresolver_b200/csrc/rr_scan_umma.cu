// rr_scan_umma.cu -- count-kernel variant A: the read-set intersections as a 0/1 int8 GEMM on
// the 5th-generation tensor cores (tcgen05.mma kind::i8, int32 accumulators in TMEM), operands
// staged by TMA into a 128B-swizzled shared-memory ring, with the significance epilogue fused
// behind the accumulator so that the count matrix never reaches HBM.
//
// Replaces the pair loop of HilfsMaxCorrsRechner (/root/reference/MaxCorrelation.c:796-830):
// counts = X^T X restricted to the band jj in [ii+20, break(ii)), X in {0,1}^(reads x groups).
//
//   A operand  xa[row tiles*128][Kp]  row sites only (>= 1 admissible row group), 6 sites per
//              32-row slab (5 groups each + 2 zero rows) so that a site never straddles a warp
//              of the epilogue; 24 sites per 128-row tile.
//   B operand  xb[5N][Kp]             every group, 48 sites = 240 columns per tile.
//   K          reads in span-start order; only the 128-read blocks [k_lo(col tile), k_hi(row tile))
//              can hold a read covering both tiles, the rest is skipped exactly.
//
// Warp roles (one persistent CTA per SM, 320 threads):
//   warp 0      TMA producer (one elected lane): A and B boxes of a K block -> smem stage
//   warp 1      TMEM allocator + MMA issuer (one elected lane): 4 x tcgen05.mma (K=32) per stage,
//               tcgen05.commit frees the stage / publishes the accumulator
//   warps 2-9   epilogue: tcgen05.ld the 128 x 240 int32 tile (two warps per 32-lane quarter, each
//               taking alternate 40-column chunks), 5x5 block sums (in-thread row sums, 5-lane
//               shuffle column sums), filters, pruning bounds, exact FP64 score, row max in
//               registers, column max through a 128-bit CAS.
// Two accumulators (2 x 256 TMEM columns) double-buffer MMA against the epilogue.
#include <cuda.h>
#include <vector>
#include <algorithm>
#include "rr_kernels.h"
#include "rr_device.cuh"
#include "rr_plan.h"

namespace {

constexpr int UM_ROW_SITES = 24;                 // row sites per tile
constexpr int UM_M = 128;                        // rows per A tile (4 slabs x 32)
constexpr int UM_COL_SITES = 48;                 // column sites per tile
constexpr int UM_N = UM_COL_SITES * 5;           // 240 columns per B tile
constexpr int UM_KB = 128;                       // reads per K block (= 128 B swizzle span)
constexpr int UM_STAGES = 4;
constexpr int UM_A_BYTES = UM_M * UM_KB;         // 16384
constexpr int UM_B_BYTES = UM_N * UM_KB;         // 30720
constexpr int UM_STAGE_BYTES = UM_A_BYTES + UM_B_BYTES;  // 47104 = 46 * 1024
constexpr int UM_EPI_WARPS = 8;
constexpr int UM_THREADS = (2 + UM_EPI_WARPS) * 32;
constexpr int UM_CHUNK_SITES = 8;                // column sites per epilogue chunk
constexpr int UM_CHUNK_COLS = UM_CHUNK_SITES * 5;  // 40 TMEM columns = x32 + x8
constexpr int UM_CHUNKS = UM_COL_SITES / UM_CHUNK_SITES;  // 6
constexpr int UM_TMEM_COLS = 512;
constexpr int UM_ACC_STRIDE = 256;               // TMEM columns between the two accumulators

struct um_meta {                                  // per accumulator buffer, column-side metadata
    double mj[UM_N];                              // running maxima of the tile's column groups
    int szj[UM_N];                                // Groupsizearray, or -1 when not admissible (817)
};

struct um_smem_tail {
    um_meta meta[2];
    unsigned long long full[UM_STAGES], empty[UM_STAGES], tfull[2], tempty[2];
    uint32_t tmem_base;
};

constexpr size_t UM_SMEM_BYTES = 1024 + (size_t)UM_STAGES * UM_STAGE_BYTES + sizeof(um_smem_tail);

struct um_unit { int32_t rt, ct0, ct1; };         // row tile, column tiles [ct0, ct1)

struct um_params {
    rr_scan_params P;
    const um_unit *units;
    int n_units;
    const int32_t *k_hi;      // [row tiles]   exclusive K-block bound
    const int32_t *k_lo;      // [column tiles] inclusive K-block bound
};

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, unsigned long long *bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(unsigned long long *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], int8 x int8 -> int32
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major operand tile in the canonical SWIZZLE_128B layout: rows of 128 B, 8-row atoms of
// 1024 B (SBO), LBO unused (1), descriptor version 1 (Blackwell), layout type 2.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                 // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset between 8-row atoms
    d |= (uint64_t)1 << 46;                 // version
    d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
    return d;
}
// kind::i8 instruction descriptor: D = S32, A = B = unsigned 8 bit, both K-major, M = 128, N = 240
__device__ __forceinline__ uint32_t make_idesc()
{
    return (2u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(UM_N >> 3) << 17) |
           ((uint32_t)(UM_M >> 4) << 24);
}

#define TMEM_LD_32(v, addr)                                                                                     \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                     \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                       \
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"       \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),        \
                   "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),      \
                   "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),      \
                   "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                                         \
                 : "r"(addr))
#define TMEM_LD_8(v, addr)                                                                                      \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"               \
                 : "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]),      \
                   "=r"(v[39])                                                                                   \
                 : "r"(addr))

// sum of x over the 5 lanes of a site (lanes 5t..5t+4), returned in every lane of the site
__device__ __forceinline__ int site_sum5(int x, int lane, int base_lane)
{
    int s1 = x + __shfl_down_sync(0xffffffffu, x, 1);
    int s2 = s1 + __shfl_down_sync(0xffffffffu, s1, 2);
    int s = s2 + __shfl_down_sync(0xffffffffu, x, 4);
    (void)lane;
    return __shfl_sync(0xffffffffu, s, base_lane);
}

__global__ void __launch_bounds__(UM_THREADS, 1)
rr_k_scan_umma(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const um_params U)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    um_smem_tail *T = reinterpret_cast<um_smem_tail *>(smem + (size_t)UM_STAGES * UM_STAGE_BYTES);
    const rr_scan_params &P = U.P;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < UM_STAGES; s++) { mbar_init(&T->full[s], 1); mbar_init(&T->empty[s], 1); }
        for (int a = 0; a < 2; a++) { mbar_init(&T->tfull[a], 1); mbar_init(&T->tempty[a], UM_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&T->tmem_base)), "n"(UM_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = T->tmem_base;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
            uint32_t it = 0;
            for (int u = blockIdx.x; u < U.n_units; u += gridDim.x) {
                const um_unit un = U.units[u];
                const int khi = U.k_hi[un.rt];
                for (int ct = un.ct0; ct < un.ct1; ct++) {
                    for (int kb = U.k_lo[ct]; kb < khi; kb++, it++) {
                        const int s = it % UM_STAGES;
                        const uint32_t ph = (it / UM_STAGES) & 1;
                        mbar_wait(&T->empty[s], ph ^ 1);
                        mbar_expect_tx(&T->full[s], UM_STAGE_BYTES);
                        uint8_t *sa = smem + (size_t)s * UM_STAGE_BYTES;
                        tma_load_2d(sa, &map_a, &T->full[s], kb * UM_KB, un.rt * UM_M);
                        tma_load_2d(sa + UM_A_BYTES, &map_b, &T->full[s], kb * UM_KB, ct * UM_N);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc = make_idesc();
            uint32_t it = 0, tile = 0;
            for (int u = blockIdx.x; u < U.n_units; u += gridDim.x) {
                const um_unit un = U.units[u];
                const int khi = U.k_hi[un.rt];
                for (int ct = un.ct0; ct < un.ct1; ct++) {
                    const int klo = U.k_lo[ct];
                    if (klo >= khi) continue;  // no read covers both tiles: the epilogue uses zeros
                    const int acc = tile & 1;
                    const uint32_t aph = (tile >> 1) & 1;
                    mbar_wait(&T->tempty[acc], aph ^ 1);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)(acc * UM_ACC_STRIDE);
                    for (int kb = klo; kb < khi; kb++, it++) {
                        const int s = it % UM_STAGES;
                        const uint32_t ph = (it / UM_STAGES) & 1;
                        mbar_wait(&T->full[s], ph);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(smem + (size_t)s * UM_STAGE_BYTES);
                        const uint64_t adesc = make_smem_desc(sa);
                        const uint64_t bdesc = make_smem_desc(sa + UM_A_BYTES);
#pragma unroll
                        for (int k = 0; k < UM_KB / 32; k++) {
                            // advance 32 bytes (one K=32 slice) inside the 128 B swizzle span: +2 in 16 B units
                            tc_mma_i8(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                                      (kb > klo || k > 0) ? 1u : 0u);
                        }
                        tc_commit(&T->empty[s]);  // frees the smem stage when these MMAs retire
                    }
                    tc_commit(&T->tfull[acc]);    // accumulator complete
                    tile++;
                }
            }
        }
    } else {
        // ================= epilogue =================
        const int ew = warp - 2;                 // 0..7
        const int quarter = warp & 3;            // TMEM lane quarter this warp may access
        const int half = ew >> 2;                // which of the two warps of the quarter
        const int et = threadIdx.x - 64;         // 0..255
        const int site_l = lane / 5;             // 0..5 (6 for the two pad lanes)
        const int a = lane - site_l * 5;         // group within the site
        const int base_lane = site_l * 5;
        const bool lane_row = lane < 30;
        unsigned n_pairs = 0, n_exact = 0, n_bound = 0, n_units = 0;
        uint32_t tile = 0, mtile = 0;

        for (int u = blockIdx.x; u < U.n_units; u += gridDim.x) {
            const um_unit un = U.units[u];
            const int khi = U.k_hi[un.rt];
            n_units += (et == 0);
            // ---- row-side state of this thread (one output row = one group of one row site) ----
            const int ii = lane_row ? P.rowsites[un.rt * UM_ROW_SITES + quarter * 6 + site_l] : -1;
            const int gi = ii >= 0 ? 5 * ii + a : -1;
            const bool row_ok = gi >= 0 && P.rowok[gi] != 0;
            const int szi = row_ok ? P.gsize[gi] : 0;
            const int brk = ii >= 0 ? min(P.breakcol[ii], P.N) : 0;
            // zi0: the group's maximum stored so far (other units / other warps); (lz, lp): the best
            // pair this thread has seen in this unit.  Pruning threshold = max(zi0, lz).
            const double zi0 = row_ok ? rr_best_value(P.best + gi) : 0.0;
            double lz = 0.0;
            int lp = 0x7fffffff;

            for (int ct = un.ct0; ct < un.ct1; ct++, mtile++) {
                const int klo = U.k_lo[ct];
                const bool has_counts = klo < khi;
                const int mb = mtile & 1;
                // column-side metadata of this tile (all 256 epilogue threads)
                um_meta &M = T->meta[mb];
                if (et < UM_N) {
                    const int j = ct * UM_N + et;
                    int sz = -1;
                    double m = 0.0;
                    if (j < 5 * P.N && P.colok[j]) { sz = P.gsize[j]; m = rr_best_value(P.best + j); }
                    M.szj[et] = sz;
                    M.mj[et] = m;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(UM_EPI_WARPS * 32) : "memory");

                int acc = 0;
                if (has_counts) {
                    acc = tile & 1;
                    const uint32_t aph = (tile >> 1) & 1;
                    mbar_wait(&T->tfull[acc], aph);
                    tc_fence_after();
                }
                const int jsite0 = ct * UM_COL_SITES;
                for (int ch = half; ch < UM_CHUNKS; ch += 2) {
                    const int cs0 = jsite0 + ch * UM_CHUNK_SITES;  // first column site of the chunk
                    if (cs0 >= P.N) break;
                    uint32_t v[UM_CHUNK_COLS];
                    if (has_counts) {
                        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) +
                                               (uint32_t)(acc * UM_ACC_STRIDE + ch * UM_CHUNK_COLS);
                        TMEM_LD_32(v, taddr);
                        TMEM_LD_8(v, taddr + 32);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    } else {
#pragma unroll
                        for (int q = 0; q < UM_CHUNK_COLS; q++) v[q] = 0u;
                    }
                    // warp-uniform skip: no lane of this warp pairs with this chunk
                    {
                        const bool any_lane = row_ok && cs0 + UM_CHUNK_SITES - 1 >= ii + 20 && cs0 < brk;
                        if (!__any_sync(0xffffffffu, any_lane)) continue;
                    }
#pragma unroll
                    for (int t = 0; t < UM_CHUNK_SITES; t++) {
                        const int jj = cs0 + t;
                        if (jj >= P.N) break;  // warp-uniform
                        const int c0 = (int)v[5 * t], c1 = (int)v[5 * t + 1], c2 = (int)v[5 * t + 2],
                                  c3 = (int)v[5 * t + 3], c4 = (int)v[5 * t + 4];
                        const int rowsum = c0 + c1 + c2 + c3 + c4;  // gr1 = |Gi & Cjj|
                        int colsum[5];                                // gr2 = |Gj & Cii| per column group
                        colsum[0] = site_sum5(c0, lane, base_lane);
                        colsum[1] = site_sum5(c1, lane, base_lane);
                        colsum[2] = site_sum5(c2, lane, base_lane);
                        colsum[3] = site_sum5(c3, lane, base_lane);
                        colsum[4] = site_sum5(c4, lane, base_lane);
                        const int cov = colsum[0] + colsum[1] + colsum[2] + colsum[3] + colsum[4];
                        const bool pair_site = row_ok && jj >= ii + 20 && jj < brk;
                        if (!pair_site) continue;
                        const int cc[5] = {c0, c1, c2, c3, c4};
#pragma unroll
                        for (int b = 0; b < 5; b++) {
                            const int col = (ch * UM_CHUNK_SITES + t) * 5 + b;
                            const int szj = M.szj[col];
                            if (szj < 0) continue;  // warp-uniform (817)
                            n_pairs++;
                            const double mj = M.mj[col];
                            const double Z = rr_pair_score(P, (unsigned)cc[b], (unsigned)rowsum, (unsigned)colsum[b],
                                                           (unsigned)cov, szi, szj, fmax(zi0, lz), mj, n_exact, n_bound);
                            if (Z > 0.0) {
                                const int gj = 5 * jj + b;
                                if (Z > lz || (Z == lz && gj < lp)) { lz = Z; lp = gj; }
                                if (Z >= mj) {
                                    rr_best_update(P.best, gj, Z, gi);
                                    if (Z > mj) M.mj[col] = Z;
                                }
                            }
                        }
                    }
                }
                if (has_counts) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&T->tempty[acc]);
                    tile++;
                }
            }
            if (lz > 0.0 && lz >= zi0) rr_best_update(P.best, gi, lz, lp);
        }

        unsigned long long v0 = n_pairs, v1 = n_exact, v2 = n_bound, v3 = n_units;
        for (int o = 16; o > 0; o >>= 1) {
            v0 += __shfl_xor_sync(0xffffffffu, v0, o);
            v1 += __shfl_xor_sync(0xffffffffu, v1, o);
            v2 += __shfl_xor_sync(0xffffffffu, v2, o);
            v3 += __shfl_xor_sync(0xffffffffu, v3, o);
        }
        if (lane == 0) {
            if (v0) atomicAdd(P.counters + 0, v0);
            if (v1) atomicAdd(P.counters + 1, v1);
            if (v2) atomicAdd(P.counters + 2, v2);
            if (v3) atomicAdd(P.counters + 3, v3);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(UM_TMEM_COLS) : "memory");
    }
}

// ---- A operand: gather the row sites' groups from xb into 32-row slabs -----------------------
__global__ void __launch_bounds__(256) rr_k_build_xa(const int8_t *__restrict__ xb, const int32_t *__restrict__ rowsites,
                                                      int64_t n_rows, int64_t Kp, int8_t *__restrict__ xa)
{
    const int64_t row = blockIdx.x;
    if (row >= n_rows) return;
    const int l = (int)(row & 31);
    const int64_t slab = row >> 5;
    int site = -1;
    if (l < 30) site = rowsites[slab * 6 + l / 5];
    const uint4 *src = site >= 0 ? reinterpret_cast<const uint4 *>(xb + ((size_t)5 * site + (l % 5)) * Kp) : nullptr;
    uint4 *dst = reinterpret_cast<uint4 *>(xa + (size_t)row * Kp);
    for (int64_t q = threadIdx.x; q < Kp / 16; q += blockDim.x) dst[q] = src ? src[q] : make_uint4(0, 0, 0, 0);
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(CUtensorMap *map, void *base, uint64_t rows, uint64_t Kp, uint32_t box_rows)
{
    static PFN_encodeTiled encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
            rr_set_error("cuTensorMapEncodeTiled is not available from this driver");
            return RR_E_CUDA;
        }
        encode = (PFN_encodeTiled)fn;
    }
    cuuint64_t dims[2] = {Kp, rows};
    cuuint64_t strides[1] = {Kp};
    cuuint32_t box[2] = {(cuuint32_t)UM_KB, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { rr_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return RR_E_CUDA; }
    return RR_OK;
}

}  // namespace

struct rr_umma_state {
    int8_t *xb = nullptr;
    int8_t *xa = nullptr;
    size_t xa_rows_cap = 0;
    int64_t Kp = 0;
    um_unit *d_units = nullptr;
    size_t units_cap = 0;
    int32_t *d_khi = nullptr, *d_klo = nullptr;
    size_t khi_cap = 0, klo_cap = 0;
    bool attr_set = false;
};

int rr_umma_available(void) { return 1; }
int rr_umma_row_sites(void) { return UM_ROW_SITES; }
int rr_umma_col_sites(void) { return UM_COL_SITES; }
int rr_umma_kblock(void) { return UM_KB; }

void rr_umma_free(rr_umma_state *s)
{
    if (!s) return;
    cudaFree(s->xb); cudaFree(s->xa); cudaFree(s->d_units); cudaFree(s->d_khi); cudaFree(s->d_klo);
    delete s;
}

#define UM_CUDA(call)                                                                                        \
    do {                                                                                                     \
        cudaError_t e__ = (call);                                                                            \
        if (e__ != cudaSuccess) {                                                                            \
            rr_set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e__), __FILE__, __LINE__, #call); \
            return RR_E_CUDA;                                                                                \
        }                                                                                                    \
    } while (0)

template <typename T>
static int grow(T **p, size_t *cap, size_t need)
{
    if (need <= *cap && *p) return RR_OK;
    cudaFree(*p);
    *p = nullptr;
    if (cudaMalloc((void **)p, std::max<size_t>(need, 1) * sizeof(T)) != cudaSuccess) {
        cudaGetLastError();
        rr_set_error("out of device memory (%zu bytes)", need * sizeof(T));
        return RR_E_NOMEM;
    }
    *cap = need;
    return RR_OK;
}

int rr_umma_scan(rr_umma_state *&S, rr_scan_params &P, rr_plan &plan, const uint8_t *d_cells, const int32_t *d_perm,
                 int codes, int n_sm, cudaStream_t st)
{
    int rc;
    if (!S) {
        S = new rr_umma_state();
        S->Kp = ((int64_t)P.R + UM_KB - 1) / UM_KB * UM_KB;
        if (S->Kp == 0) S->Kp = UM_KB;
        const size_t rows = (size_t)5 * P.N;
        if (cudaMalloc((void **)&S->xb, std::max<size_t>(rows * S->Kp, 16)) != cudaSuccess) {
            cudaGetLastError();
            rr_set_error("out of device memory for the int8 operand (%zu bytes)", rows * (size_t)S->Kp);
            return RR_E_NOMEM;
        }
        UM_CUDA(rr_launch_pack_int8(d_cells, d_perm, P.R, P.N, codes, S->xb, S->Kp, st));
    }
    // A operand for this plan's row sites
    const size_t xa_rows = (size_t)std::max(plan.n_rowblocks, 1) * UM_M;
    if (xa_rows > S->xa_rows_cap) {
        cudaFree(S->xa);
        S->xa = nullptr;
        if (cudaMalloc((void **)&S->xa, xa_rows * S->Kp) != cudaSuccess) {
            cudaGetLastError();
            rr_set_error("out of device memory for the A operand (%zu bytes)", xa_rows * (size_t)S->Kp);
            return RR_E_NOMEM;
        }
        S->xa_rows_cap = xa_rows;
    }
    rr_k_build_xa<<<(unsigned)xa_rows, 256, 0, st>>>(S->xb, P.rowsites, (int64_t)xa_rows, S->Kp, S->xa);
    rr_count_launch(1);
    UM_CUDA(cudaGetLastError());

    // work units: (row tile, up to UNIT_CT consecutive column tiles), this part's row tiles only
    constexpr int UNIT_CT = 16;
    std::vector<um_unit> units;
    int64_t kblocks = 0;
    for (int rb = plan.rb_lo; rb < plan.rb_hi; rb++) {
        const int cb0 = plan.unit_cb0[rb];
        const int ncb = (int)(plan.unit_prefix[rb + 1] - plan.unit_prefix[rb]);
        for (int c = 0; c < ncb; c += UNIT_CT) units.push_back({rb, cb0 + c, cb0 + std::min(ncb, c + UNIT_CT)});
        for (int c = 0; c < ncb; c++) kblocks += std::max(0, plan.k_hi[rb] - plan.k_lo[cb0 + c]);
    }
    plan.executed_ops = kblocks * (int64_t)(2LL * UM_M * UM_N * UM_KB);
    if (units.empty()) return RR_OK;
    if ((rc = grow(&S->d_units, &S->units_cap, units.size()))) return rc;
    if ((rc = grow(&S->d_khi, &S->khi_cap, plan.k_hi.size()))) return rc;
    if ((rc = grow(&S->d_klo, &S->klo_cap, plan.k_lo.size()))) return rc;
    UM_CUDA(cudaMemcpyAsync(S->d_units, units.data(), sizeof(um_unit) * units.size(), cudaMemcpyHostToDevice, st));
    UM_CUDA(cudaMemcpyAsync(S->d_khi, plan.k_hi.data(), sizeof(int32_t) * plan.k_hi.size(), cudaMemcpyHostToDevice, st));
    UM_CUDA(cudaMemcpyAsync(S->d_klo, plan.k_lo.data(), sizeof(int32_t) * plan.k_lo.size(), cudaMemcpyHostToDevice, st));
    UM_CUDA(cudaStreamSynchronize(st));  // units[] is a local

    CUtensorMap map_a, map_b;
    if ((rc = make_map(&map_a, S->xa, xa_rows, (uint64_t)S->Kp, UM_M))) return rc;
    if ((rc = make_map(&map_b, S->xb, (uint64_t)5 * P.N, (uint64_t)S->Kp, UM_N))) return rc;

    if (!S->attr_set) {
        UM_CUDA(cudaFuncSetAttribute(rr_k_scan_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UM_SMEM_BYTES));
        S->attr_set = true;
    }
    um_params U;
    U.P = P;
    U.units = S->d_units;
    U.n_units = (int)units.size();
    U.k_hi = S->d_khi;
    U.k_lo = S->d_klo;
    const int grid = std::min<int>(n_sm, (int)units.size());
    rr_k_scan_umma<<<grid, UM_THREADS, UM_SMEM_BYTES, st>>>(map_a, map_b, U);
    rr_count_launch(1);
    UM_CUDA(cudaGetLastError());
    return RR_OK;
}
