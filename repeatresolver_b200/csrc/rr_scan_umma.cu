// rr_scan_umma.cu -- count-kernel variant A: the read-set intersections as a 0/1 GEMM on the 5th-generation tensor
// cores (tcgen05.mma; operands as int8 with int32 accumulators, or as packed 4-bit e2m1 with fp32 accumulators, the
// latter also block-scaled kind::mxf4 at twice the rate), accumulators in TMEM, operands staged by TMA into a
// 128B-swizzled shared-memory ring, with the significance epilogue fused behind the accumulator so that the count
// matrix never reaches HBM.
//
// Replaces the pair loop of HilfsMaxCorrsRechner (/root/reference/MaxCorrelation.c:796-830):
// counts = X^T X restricted to the band jj in [ii+20, break(ii)), X in {0,1}^(reads x groups).
//
//   A operand  xa[row tiles*128][Kp]  row sites only (>= 1 admissible row group), 6 sites per
//              32-row slab (5 groups each + 2 zero rows) so that a site never straddles a warp
//              of the epilogue; 24 sites per 128-row tile.
//   B operand  xb[5N][Kp]             every group, 48 sites = 240 columns per tile.
//   K          reads in (length class, span start) order; only the K blocks [k_lo(col tile), k_hi(row tile)) of every class
//              (128 reads, 256 for mxf4) can hold a read covering both tiles, the rest is skipped exactly.
//
// Operand elements are 0 / 2 (not 0 / 1): a product is 4 = sizeof(float), so an accumulator holds 4 x count, which is
// the byte offset of ln(count!) in the table the epilogue's bounds look up - no shift-and-add per look-up.
//
// The kernel runs on CTA PAIRS (clusters of two = the two SMs of a TPC, tcgen05 cta_group::2): one MMA of M = 256 covers
// the row tiles of both CTAs (128 accumulator rows in each CTA's TMEM) against one B tile of which each CTA loads and holds
// HALF (120 rows) - a third less TMA traffic into and tensor-core traffic out of every SM's shared memory than two
// independent CTAs, and half the L2 -> SM traffic for B.  Rank 0 of the pair (the leader) issues the MMAs; both producers'
// boxes are counted on the leader's barrier; one multicast tcgen05.commit releases a stage / publishes an accumulator in
// both CTAs; the epilogue warps of both CTAs release an accumulator on the leader's barrier.
//
// Warp roles (one persistent CTA per SM, 640 threads):
//   warp 0      producer (one elected lane): per tile one bulk copy of the tile's 240 running maxima + admissibility
//               masks into a 2-deep shared-memory ring, then per K block this CTA's A box and its half of the B box by TMA
//               -> smem stage (4 stages of 31 KB)
//   warp 1      TMEM allocator (both CTAs) + MMA issuer (one elected lane of the leader): 4 x tcgen05.mma per stage,
//               tcgen05.commit frees the stage / publishes the accumulator
//   warp 2      per tile: turns the producer's copy of the 240 running maxima and the admissibility masks into the
//               tier-1 limits every epilogue warp reads (2-deep ring of its own), one tile ahead of the epilogue
//   warp 3      idle (keeps the epilogue warps aligned to the TMEM lane quarters); warps 0-3 give their registers
//               to the epilogue (setmaxnreg 32 / 112)
//   warps 4-19  epilogue, four warps per 32-lane TMEM quarter, each taking every fourth column site:
//               tcgen05.ld of the site's 5 counts, 5x5 block sums (in-thread row sum, 5-lane shuffle
//               column sums), filters, tiered pruning bounds; survivors are queued per warp and
//               evaluated 32 at a time (exact FP64 score), maxima folded in with a 128-bit CAS.
// Two accumulators (2 x 256 TMEM columns) double-buffer MMA against the epilogue.
#include <cuda.h>
#include <vector>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include "rr_kernels.h"
#include "rr_device.cuh"
#include "rr_plan.h"
#include "../../include/rr_debug.h"

namespace {

constexpr int UM_ROW_SITES = 24;                 // row sites per tile
constexpr int UM_M = 128;                        // rows per A tile (4 slabs x 32)
constexpr int UM_COL_SITES = 48;                 // column sites per tile
constexpr int UM_N = UM_COL_SITES * 5;           // 240 columns per B tile
constexpr int UM_KB = 128;                       // reads per K block (= 128 B swizzle span)
constexpr int UM_STAGES = 4;
constexpr int UM_A_BYTES = UM_M * UM_KB;         // 16384: this CTA's row tile
constexpr int UM_BH = UM_N / 2;                  // 120: the half of the B tile this CTA of the pair holds (cta_group::2)
constexpr int UM_B_BYTES = UM_BH * UM_KB;        // 15360
constexpr int UM_STAGE_BYTES = UM_A_BYTES + UM_B_BYTES;  // 31744 = 31 * 1024
static_assert(UM_STAGE_BYTES % 1024 == 0 && UM_BH % 8 == 0, "swizzled operand tiles are whole 1024-byte atoms");
constexpr int UM_FIRST_EPI_WARP = 4;
#ifndef RR_UM_EPI_WARPS
#define RR_UM_EPI_WARPS 16
#endif
constexpr int UM_EPI_WARPS = RR_UM_EPI_WARPS;    // 16 (four per scheduler) or 20
// registers: the CTA's pool is what it is launched with (96 per thread at 640 threads, 80 at 768); the four control warps
// drop to 32 and the epilogue warps can only take what that frees, in steps of 8 (16 warps: 96 + 4*64/16 = 112; 20 warps:
// 80 + 4*48/20 -> 88).  Asking for more blocks forever.
constexpr int UM_EPI_REGS = UM_EPI_WARPS == 16 ? 112 : 88;
static_assert((UM_FIRST_EPI_WARP + UM_EPI_WARPS) * (UM_EPI_WARPS == 16 ? 96 : 80) >= UM_FIRST_EPI_WARP * 32 + UM_EPI_WARPS * UM_EPI_REGS,
              "setmaxnreg.inc would wait for registers the CTA does not have");
constexpr int UM_THREADS = (UM_FIRST_EPI_WARP + UM_EPI_WARPS) * 32;  // 640
constexpr int UM_SUB = UM_EPI_WARPS / 4;         // epilogue warps per TMEM lane quarter
constexpr int UM_WSITES = (UM_COL_SITES + UM_SUB - 1) / UM_SUB;   // column sites per warp and tile (at most)
constexpr int UM_TMEM_COLS = 512;
constexpr int UM_ACC_STRIDE = 256;               // TMEM columns between the two accumulators
constexpr int UM_SF_COL = 240;                   // 16 spare TMEM columns behind accumulator 0: unit block scales (mxf4)
constexpr int UM_QSHIFT = 2;                     // accumulators hold count << UM_QSHIFT (operand elements are 2); rr_tier1_q assumes 2
constexpr int UM_CMASK_BYTES = 48;               // admissibility masks of a tile's column sites (one byte per site)

struct __align__(16) um_wsite {                   // one column site: read by the whole warp as two broadcasts (16 + 8 bytes)
    int nq[5];                                    // running maxima as fixed-point tier-1 limits (rr_thr_q); inadmissible: never
    int vmask;                                    // bit b set: group b of the site is admissible (817); 0 outside the MSA
    int pad[2];
};
struct __align__(16) um_tile_meta {               // per tile, written by the converter warp, read by every epilogue warp
    um_wsite site[UM_COL_SITES];
    int32_t has_counts;                           // some read covers both tiles (K blocks were issued)
    int32_t pad[3];
};
struct __align__(16) um_thr_buf {                 // per tile, written by the producer's bulk copies
    rr_best_t best[UM_N];                         // running maxima of the tile's column groups as they are in HBM
    uint8_t cmask[UM_CMASK_BYTES];                // [UM_COL_SITES] bit b: group b of the site is admissible (817)
    int32_t has_counts;                           // written by the producer: some read covers both tiles (K blocks were issued)
    int32_t pad[3];
};

struct um_smem_tail {
    um_tile_meta tmeta[2];
    rr_cand_p q1[UM_EPI_WARPS][RR_QUEUE_CAP];     // tier-1 survivors, one queue per epilogue warp
    rr_cand_p q2[UM_EPI_WARPS][RR_QUEUE_CAP];     // tier-2 survivors (exact evaluation pending)
    um_thr_buf thr[2];
    unsigned long long full[UM_STAGES], empty[UM_STAGES], tfull[2], tempty[2], bfull[2], bempty[2], mfull[2], mempty[2];
    uint32_t tmem_base;
};

constexpr size_t UM_TAIL_OFF = (size_t)UM_STAGES * UM_STAGE_BYTES;
constexpr size_t UM_LNF_OFF = (UM_TAIL_OFF + sizeof(um_smem_tail) + 15) & ~(size_t)15;
constexpr size_t UM_SMEM_MAX = 227 * 1024;
constexpr int UM_LNF_MAX = (int)((UM_SMEM_MAX - UM_LNF_OFF) / sizeof(int));  // ln(n!) table entries that fit
static_assert(UM_LNF_MAX >= 4096, "the float ln(n!) table should hold the depths of the bench workloads");

// a work unit of a CTA pair: row tile rt0 for the CTA of rank 0, rt1 for rank 1 (-1: none, that CTA only lends its tensor
// core and its half of the B tiles), column tiles [ct0, ct1) for both
struct um_unit { int32_t rt0, rt1, ct0, ct1; };

struct um_params {
    rr_scan_params P;
    const um_unit *units;
    int n_units;
    const int32_t *k_hi;      // [n_cls][n_rt] exclusive K-block bound per length class of rows (rr_plan.h) and row tile
    const int32_t *k_lo;      // [n_cls][n_ct] inclusive K-block bound per class and column tile
    int n_cls;
    int n_rt, n_ct;
    const uint8_t *colmask;   // [n_ct * 48 (+ padding)] per column site: bit b = group b admissible as a column group
    int lnf_smem;             // ln(n!) entries (fixed point, rr_tier1_q) staged in shared memory
    int t1q_shift;            // their scale S = 2^shift
    float t1q_scale;          // ln(10) * S rounded down: thresholds -> table units
    int preseed;              // pre-seed launch: subsample the rows of first-visit columns
    int32_t *dump;            // DUMP instantiation only: [UM_M][UM_N] counts of the (single) tile processed
    int dump_rank;            //   ... by the CTA of this rank in the pair
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
// try_wait suspends the thread until the phase completes or a time limit passes (the hint, in ns, asks for a long one), so
// a waiting warp issues an instruction per wake-up instead of polling
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(1000000u)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}
// the single-lane producer / MMA warps: back off between wake-ups so that they leave the issue slots to the epilogue
__device__ __forceinline__ void mbar_wait_sleep(unsigned long long *bar, uint32_t parity, unsigned ns)
{
    while (!mbar_try_wait(bar, parity)) __nanosleep(ns);
}
// plain bulk copy global -> shared (16-byte aligned on both sides, size a multiple of 16), completion on an mbarrier
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gsrc, uint32_t bytes, unsigned long long *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// ---- CTA pair (cluster of two, cta_group::2) ----
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same variable in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_u32(uint32_t cta_addr, uint32_t rank)
{
    uint32_t a;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(cta_addr), "r"(rank));
    return a;
}
// arrive on a barrier of the other CTA of the pair.  No cluster-scope release: nothing written through the generic proxy is
// handed over with these arrivals (TMEM reads are ordered by tcgen05.wait::ld + tcgen05.fence, operand tiles by the TMA's own
// completion), and the .release.cluster form costs a MEMBAR.ALL.GPU per arrival (measured: 10 % of all warp samples)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// a box of this CTA's operand half into its own shared memory; the bytes are counted on the LEADER CTA's mbarrier
// (cluster address), which the 2-CTA MMA waits for
__device__ __forceinline__ void tma_load_2d_pair(void *smem_dst, const CUtensorMap *map, uint32_t leader_bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1)
        : "memory");
}
// completion of every MMA issued so far -> one arrival on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit_pair(unsigned long long *bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// D[tmem of both CTAs, 2 x 128 rows] (+)= A[both CTAs' row tiles] * B[the two halves of the column tile]; issued by the leader
template <int MODE>
__device__ __forceinline__ void tc_mma_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                            uint32_t tsfa, uint32_t tsfb)
{
    if constexpr (MODE == 2)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(tsfa), "r"(tsfb)
            : "memory");
    else if constexpr (MODE == 1)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(unsigned long long *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem].  MODE 0: int8 x int8 -> int32 (kind::i8); 1: e2m1 x e2m1 -> fp32 at the 8-bit
// rate (kind::f8f6f4, operands unpacked to one element per byte in smem); 2: packed e2m1 with unit block scales
// (kind::mxf4.block_scale, UE8M0 scale 2^0 per 32 elements, K = 64 per instruction, twice the 8-bit rate)
template <int MODE>
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                       uint32_t tsfa, uint32_t tsfb)
{
    if constexpr (MODE == 2)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(tsfa), "r"(tsfb)
            : "memory");
    else if constexpr (MODE == 1)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    else
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
// instruction descriptor, both operands K-major, M = 128, N = 240:
//   kind::i8      D = S32 (2), A = B = unsigned 8 bit (0)
//   kind::f8f6f4  D = F32 (1), A = B = E2M1 (5)
//   kind::mxf4    block-scaled descriptor: A = B = E2M1 (1), scale format UE8M0 (bit 23), scale ids 0, K = 64
template <int MODE, int M = UM_M>
__device__ __forceinline__ uint32_t make_idesc()
{
    if constexpr (MODE == 2)
        return (1u << 7) | (1u << 10) | ((uint32_t)(UM_N >> 3) << 17) | (1u << 23) | ((uint32_t)(M >> 4) << 24);
    const uint32_t cfmt = MODE == 1 ? 1u : 2u, abfmt = MODE == 1 ? 5u : 0u;
    return (cfmt << 4) | (abfmt << 7) | (abfmt << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(UM_N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

#define TMEM_LD_8(v, addr)                                                                          \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"   \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) \
                 : "r"(addr))

// sum of x over the 5 lanes of a site (lanes 5t..5t+4), returned in every lane of the site
__device__ __forceinline__ int site_sum5(int x, int base_lane)
{
    int s1 = x + __shfl_down_sync(0xffffffffu, x, 1);
    int s2 = s1 + __shfl_down_sync(0xffffffffu, s1, 2);
    int s = s2 + __shfl_down_sync(0xffffffffu, x, 4);
    return __shfl_sync(0xffffffffu, s, base_lane);
}

// the same for two counts at once, packed 16 + 16 bits (every sum is <= 4 x the shared coverage <= 4 R < 65536)
__device__ __forceinline__ void site_sum5_x2(int x0, int x1, int base_lane, int &s0, int &s1)
{
    const int s = site_sum5(x0 | (x1 << 16), base_lane);
    s0 = s & 0xffff;
    s1 = (int)((unsigned)s >> 16);
}

extern __shared__ __align__(1024) uint8_t um_smem[];

// round(ln(n!) * S) for the bounds (rr_tier1_q), addressed by the BYTE OFFSET 4 n (what the accumulators hold).
// ALL_SMEM: every argument (<= largest column coverage) is inside the table staged in shared memory; otherwise the tail
// is computed from the double table in HBM/L2 with the same rounding.
template <bool ALL_SMEM>
struct um_lnf {
    uint32_t base;   // shared-window address of the table
    unsigned smem_bytes;
    const double *gmem;
    double scale;
    __device__ __forceinline__ int lds(unsigned off) const
    {
        int v;
        asm("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(base + off));
        return v;
    }
    __device__ __forceinline__ int operator()(unsigned off) const
    {
        if constexpr (ALL_SMEM) return lds(off);
        else return off < smem_bytes ? lds(off) : __double2int_rn(__ldg(gmem + (off >> UM_QSHIFT)) * scale);
    }
};

// a tier-1 limit from the high word of a running maximum (a non-negative double): dropping the low word can only lower the
// maximum (thresholds may be stale or low, never high)
__device__ __forceinline__ int um_nq_hi(uint32_t hi, bool no_prune, float qscale)
{
    return rr_thr_q(__hiloint2double((int)hi, 0), no_prune, qscale);
}
__device__ __forceinline__ uint32_t um_best_hi(const rr_best_t *p)
{
    uint32_t hi;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(hi) : "l"(reinterpret_cast<const char *>(p) + 4));
    return hi;
}

// exclusive K-block bound of length class seg for a unit: both CTAs of the pair run the same K blocks, the union of what
// their two row tiles need
__device__ __forceinline__ int um_khi(const um_params &U, int seg, const um_unit &un)
{
    int k = U.k_hi[seg * U.n_rt + un.rt0];
    if (un.rt1 >= 0) k = max(k, U.k_hi[seg * U.n_rt + un.rt1]);
    return k;
}

// PACK16: 4 R < 65536 is known at compile time (the fast path of the bench workloads); otherwise decided at run time
template <bool ALL_SMEM, int MODE, bool DUMP, bool PACK16>
__global__ void __launch_bounds__(UM_THREADS, 1)
rr_k_scan_umma(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const um_params U)
{
    uint8_t *smem = um_smem;
    um_smem_tail *T = reinterpret_cast<um_smem_tail *>(smem + UM_TAIL_OFF);
    int *lnf_s = reinterpret_cast<int *>(smem + UM_LNF_OFF);
    const rr_scan_params &P = U.P;
    // the warp index through a shuffle: the compiler then keeps it (and what derives from it) in uniform registers
    // instead of re-reading SR_TID inside the epilogue loop
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const bool mma_only = (P.flags & RR_DEBUG_MMA_ONLY) != 0;   // timing decomposition: accumulators released unread
    // the CTA pair: rank 0 (the leader) issues the MMAs for both; rank r owns one row tile of the unit and half r of every B tile
    const int cta_rank = (int)cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs_grid = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) __trap();  // the swizzled operand tiles need a 1024-byte aligned base
        // full: the leader's counts both producers (its own arrive + expect_tx, the peer's remote arrive) and the bytes of
        // both CTAs' boxes; empty / tfull: one multicast commit of the leader's MMA thread; tempty: the leader's counts the
        // epilogue warps of both CTAs
        for (int s = 0; s < UM_STAGES; s++) { mbar_init(&T->full[s], 2); mbar_init(&T->empty[s], 1); }
        for (int a = 0; a < 2; a++) {
            mbar_init(&T->tfull[a], 1); mbar_init(&T->tempty[a], 2 * UM_EPI_WARPS);
            mbar_init(&T->bfull[a], 1); mbar_init(&T->bempty[a], 1);
            mbar_init(&T->mfull[a], 1); mbar_init(&T->mempty[a], UM_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // the same warp of both CTAs
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&T->tmem_base)), "n"(UM_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    {
        const double scale = (double)(1 << U.t1q_shift);
        for (int n = threadIdx.x; n < U.lnf_smem; n += UM_THREADS) lnf_s[n] = __double2int_rn(U.P.lnfact[n] * scale);
    }
    tc_fence_before();
    cluster_sync_all();   // barriers of both CTAs initialised before either signals the other's
    tc_fence_after();
    const uint32_t tmem_base = T->tmem_base;
    if constexpr (MODE == 2) {
        // unit block scales: every byte of TMEM columns 240..255 = 0x7F (UE8M0 2^0), in all 128 lanes, so that the
        // MMA reads 1.0 wherever its scale-factor layout points inside that window
        if (warp >= UM_FIRST_EPI_WARP && warp < UM_FIRST_EPI_WARP + 4) {
            const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)UM_SF_COL;
            const uint32_t one = 0x7F7F7F7Fu;
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
                ::"r"(taddr), "r"(one) : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }

    // register reallocation: the control warpgroup (warps 0-3) gives registers to the 16 epilogue warps
    if (warp < UM_FIRST_EPI_WARP) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    if (warp == 0) {
        // ================= producer =================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
            uint32_t it = 0, tix = 0;
            const uint32_t leader_full0 = mapa_u32(smem_u32(&T->full[0]), 0);
            for (int u = pair; u < U.n_units; u += n_pairs_grid) {
                const um_unit un = U.units[u];
                for (int ct = un.ct0; ct < un.ct1; ct++, tix++) {
                    {   // the tile's running maxima and admissibility masks, one tile ahead of the epilogue at least
                        const int tb = tix & 1;
                        mbar_wait_sleep(&T->bempty[tb], ((tix >> 1) & 1) ^ 1, 64);
                        const int ng = min(UM_N, 5 * U.P.N - ct * UM_N);
                        // plain store, ordered before the waiters' reads by the release of the arrive below
                        int any = 0;
                        for (int seg = 0; seg < U.n_cls; seg++) any |= U.k_lo[seg * U.n_ct + ct] < um_khi(U, seg, un);
                        T->thr[tb].has_counts = any;
                        mbar_expect_tx(&T->bfull[tb], (uint32_t)(ng * sizeof(rr_best_t) + UM_CMASK_BYTES));
                        bulk_load(&T->thr[tb].best[0], U.P.best + (size_t)ct * UM_N, (uint32_t)(ng * sizeof(rr_best_t)), &T->bfull[tb]);
                        bulk_load(&T->thr[tb].cmask[0], U.colmask + (size_t)ct * UM_CMASK_BYTES, UM_CMASK_BYTES, &T->bfull[tb]);
                    }
                    for (int seg = 0; seg < U.n_cls; seg++)
                    for (int kb = U.k_lo[seg * U.n_ct + ct], ke = um_khi(U, seg, un); kb < ke; kb++, it++) {
                        const int s = it % UM_STAGES;
                        const uint32_t ph = (it / UM_STAGES) & 1;
                        mbar_wait_sleep(&T->empty[s], ph ^ 1, 64);
                        // the leader's barrier expects the boxes of both CTAs; the transaction count is in HBM-side bytes:
                        // packed e2m1 delivers half the smem footprint
                        const uint32_t lfull = leader_full0 + (uint32_t)(s * sizeof(unsigned long long));
                        if (cta_rank == 0) mbar_expect_tx(&T->full[s], 2 * (MODE == 1 ? UM_STAGE_BYTES / 2 : UM_STAGE_BYTES));
                        else mbar_arrive_cluster(lfull);
                        uint8_t *sa = smem + (size_t)s * UM_STAGE_BYTES;
                        constexpr int KREADS = MODE == 2 ? 2 * UM_KB : UM_KB;  // reads per 128-byte smem row
                        tma_load_2d_pair(sa, &map_a, lfull, kb * KREADS, (cta_rank == 0 || un.rt1 < 0 ? un.rt0 : un.rt1) * UM_M);
                        tma_load_2d_pair(sa + UM_A_BYTES, &map_b, lfull, kb * KREADS, ct * UM_N + cta_rank * UM_BH);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0 && cta_rank == 0) {
            const uint32_t idesc = make_idesc<MODE, 2 * UM_M>();
            uint32_t it = 0, tile = 0;
            for (int u = pair; u < U.n_units; u += n_pairs_grid) {
                const um_unit un = U.units[u];
                for (int ct = un.ct0; ct < un.ct1; ct++) {
                    int any = 0;
                    for (int seg = 0; seg < U.n_cls; seg++) any |= U.k_lo[seg * U.n_ct + ct] < um_khi(U, seg, un);
                    if (!any) continue;   // no read covers both tiles: the epilogue uses zeros
                    const int acc = tile & 1;
                    const uint32_t aph = (tile >> 1) & 1;
                    mbar_wait_sleep(&T->tempty[acc], aph ^ 1, 128);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)(acc * UM_ACC_STRIDE);
                    uint32_t accumulate = 0;   // the first MMA of a tile overwrites the accumulator
                    for (int seg = 0; seg < U.n_cls; seg++)
                    for (int kb = U.k_lo[seg * U.n_ct + ct], ke = um_khi(U, seg, un); kb < ke; kb++, it++) {
                        const int s = it % UM_STAGES;
                        const uint32_t ph = (it / UM_STAGES) & 1;
                        mbar_wait_sleep(&T->full[s], ph, 32);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(smem + (size_t)s * UM_STAGE_BYTES);
                        const uint64_t adesc = make_smem_desc(sa);
                        const uint64_t bdesc = make_smem_desc(sa + UM_A_BYTES);
#pragma unroll
                        for (int k = 0; k < UM_KB / 32; k++) {
                            // advance 32 bytes (one K=32 slice) inside the 128 B swizzle span: +2 in 16 B units
                            tc_mma_pair<MODE>(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, accumulate,
                                              tmem_base + UM_SF_COL, tmem_base + UM_SF_COL + 8);
                            accumulate = 1u;
                        }
                        tc_commit_pair(&T->empty[s]);  // frees the stage in both CTAs when these MMAs retire
                    }
                    tc_commit_pair(&T->tfull[acc]);    // accumulator complete, in both CTAs
                    tile++;
                }
            }
        }
    } else if (warp == 2) {
        // ================= tile metadata =================
        // the tile's running maxima as they were in HBM when the producer copied them (one to two tiles ago: thresholds may
        // be stale or low, never high) -> fixed-point tier-1 limits, once per tile for all sixteen epilogue warps
        const bool no_prune = (P.flags & RR_FLAG_NO_PRUNE) != 0;
        const float qscale = U.t1q_scale;
        uint32_t tix = 0;
        for (int u = pair; u < U.n_units; u += n_pairs_grid) {
            const um_unit un = U.units[u];
            for (int ct = un.ct0; ct < un.ct1; ct++, tix++) {
                const int tb = tix & 1;
                const uint32_t ph = (tix >> 1) & 1;
                const um_thr_buf &B = T->thr[tb];
                um_tile_meta &TM = T->tmeta[tb];
                mbar_wait_sleep(&T->mempty[tb], ph ^ 1, 64);   // the epilogue is done with the tile that used this slot
                mbar_wait_sleep(&T->bfull[tb], ph, 64);        // the producer's copy has landed
                __syncwarp();
                const int n_in = min(UM_COL_SITES, P.N - ct * UM_COL_SITES);   // column sites of the tile inside the MSA
#pragma unroll 1
                for (int e = lane; e < UM_N; e += 32) {
                    const int t = e / 5, b = e - 5 * t;
                    const bool ok = t < n_in && ((B.cmask[t] >> b) & 1) != 0;                            // (817)
                    const uint32_t hi = t < n_in ? (uint32_t)(B.best[e].z >> 32) : 0u;
                    TM.site[t].nq[b] = ok ? um_nq_hi(hi, no_prune, qscale) : RR_T1Q_INADMISSIBLE;
                }
                for (int t = lane; t < UM_COL_SITES; t += 32) TM.site[t].vmask = t < n_in ? (int)(B.cmask[t] & 31u) : 0;
                if (lane == 0) TM.has_counts = B.has_counts;
                __syncwarp();
                if (lane == 0) { mbar_arrive(&T->mfull[tb]); mbar_arrive(&T->bempty[tb]); }
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(UM_EPI_REGS));
        // ================= epilogue =================
        // all counts below are in units of 1/4 (count << UM_QSHIFT), as the accumulators deliver them
        const int ew = warp - UM_FIRST_EPI_WARP;  // 0..UM_EPI_WARPS-1
        const int quarter = warp & 3;            // TMEM lane quarter this warp may access
        const int sub = ew >> 2;                 // which of the UM_SUB warps of the quarter
        const int site_l = lane / 5;             // 0..5 (6 for the two pad lanes)
        const int a = lane - site_l * 5;         // group within the site
        const int base_lane = site_l * 5;
        const bool lane_row = lane < 30;
        unsigned n_pairs = 0, n_exact = 0, n_units = 0, n_tier2 = 0;
        uint32_t tile = 0, tix = 0;
        rr_cand_p *q1 = T->q1[ew], *q2 = T->q2[ew];
        int c1n = 0, c2n = 0;
        const bool no_prune = (P.flags & RR_FLAG_NO_PRUNE) != 0;
        const bool pack16 = PACK16 || P.R < (65536 >> UM_QSHIFT);
        const bool subsample = U.preseed != 0;   // set by the host for the pre-seed launch only
        um_lnf<ALL_SMEM> LT;     // fixed-point table, tier 1
        LT.base = (uint32_t)__shfl_sync(0xffffffffu, (int)smem_u32(lnf_s), 0); LT.smem_bytes = (unsigned)U.lnf_smem << UM_QSHIFT; LT.gmem = P.lnfact;
        LT.scale = (double)(1 << U.t1q_shift);
        const float qscale = U.t1q_scale;
        rr_lnf_global LG;        // double table in HBM/L2, tier 2 (rare, evaluated 32 at a time)
        LG.gmem = P.lnfact;

        const uint32_t leader_tempty0 = mapa_u32(smem_u32(&T->tempty[0]), 0);
        for (int u = pair; u < U.n_units; u += n_pairs_grid) {
            const um_unit un = U.units[u];
            n_units += (ew == 0 && lane == 0 && cta_rank == 0);
            // ---- row-side state of this thread (one output row = one group of one row site) ----
            const int rt = cta_rank == 0 ? un.rt0 : un.rt1;   // this CTA's row tile; a unit with one row tile leaves rank 1 without rows
            const int ii = lane_row && rt >= 0 ? P.rowsites[rt * UM_ROW_SITES + quarter * 6 + site_l] : -1;
            const int gi = ii >= 0 ? 5 * ii + a : -1;
            const bool row_ok = gi >= 0 && P.rowok[gi] != 0;
            const int brk = ii >= 0 ? min(P.breakcol[ii], P.N) : 0;
            const int row_lo = row_ok ? ii + 20 : 0x7fffffff;   // first column site this row group is tested against (796)
            uint32_t row_hi = row_ok ? um_best_hi(P.best + gi) : 0u;   // running max of row group i (high word), refreshed per tile

            for (int ct = un.ct0; ct < un.ct1; ct++, tix++) {
                const int jsite0 = ct * UM_COL_SITES;
                bool has_counts;
                // ---- the tile's column groups: tier-1 limits and admissibility masks as the converter warp left them
                int nq_i = um_nq_hi(row_hi, no_prune, qscale);
                const int tb = tix & 1;
                um_tile_meta &TM = T->tmeta[tb];
                mbar_wait(&T->mfull[tb], (tix >> 1) & 1);
                has_counts = TM.has_counts != 0;
                // the row group's maximum for the next tile of this unit (consumed at the top of the next iteration)
                if (row_ok && ct + 1 < un.ct1) row_hi = um_best_hi(P.best + gi);

                const int acc = tile & 1;
                if (has_counts) {
                    mbar_wait(&T->tfull[acc], (tile >> 1) & 1);
                    tc_fence_after();
                }
                const int t_end = mma_only ? 0 : min(UM_COL_SITES, P.N - jsite0);
                if (!has_counts) {
                    // no read covers both tiles: every count is 0, no pair can score, but the pair tests are counted
#pragma unroll 1
                    for (int t = sub; t < t_end; t += UM_SUB) {
                        const int jj = jsite0 + t;
                        if (row_ok && jj >= ii + 20 && jj < brk) n_pairs += __popc(TM.site[t].vmask);
                        if constexpr (DUMP)
                            if (cta_rank == U.dump_rank)
                                for (int b = 0; b < 5; b++) U.dump[(size_t)(quarter * 32 + lane) * UM_N + 5 * t + b] = 0;
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&T->mempty[tb]);
                    continue;
                }
                // 8 TMEM columns are fetched per site (5 used); the load of the next site is in flight
                // while the current one is processed
                uint32_t v[8];
                uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * UM_ACC_STRIDE) + 5 * sub;
                if (sub < t_end) TMEM_LD_8(v, taddr);
                // a counted loop over this warp's sites t = sub, sub + UM_SUB, ...: column jj, metadata *sp
                int jj = jsite0 + sub - UM_SUB;
                const um_wsite *sp = &TM.site[sub] - UM_SUB;
#pragma unroll 1
                for (int n_left = t_end > sub ? (t_end - sub + UM_SUB - 1) / UM_SUB : 0; n_left > 0; n_left--) {
                    int c[5];
                    jj += UM_SUB;
                    sp += UM_SUB;
                    // the site's limits and mask, requested before the wait for its counts (behind the vote below they sit in
                    // front of the first branch that needs them)
                    const int4 m0 = *reinterpret_cast<const int4 *>(&sp->nq[0]);
                    const int2 m1 = *reinterpret_cast<const int2 *>(&sp->nq[4]);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int b = 0; b < 5; b++) c[b] = MODE != 0 ? (int)__uint_as_float(v[b]) : (int)v[b];
                    taddr += 5 * UM_SUB;
                    if (n_left > 1) TMEM_LD_8(v, taddr);
                    if constexpr (DUMP) {
                        if (cta_rank == U.dump_rank) {
#pragma unroll
                            for (int b = 0; b < 5; b++) U.dump[(size_t)(quarter * 32 + lane) * UM_N + 5 * (jj - jsite0) + b] = c[b] >> UM_QSHIFT;
                        }
                    }
                    const bool pair_site = jj >= row_lo && jj < brk;   // (row_lo: never for rows that are no row groups)
                    if (!__any_sync(0xffffffffu, pair_site)) continue;
                    const int rowsum = c[0] + c[1] + c[2] + c[3] + c[4];  // gr1 = |Gi & Cjj|
                    int colsum[5];                                          // gr2 = |Gj & Cii| per column group
                    int cov;
                    bool need[5];
                    int mjw[5];
                    int vmask;
                    {
                        mjw[0] = m0.x; mjw[1] = m0.y; mjw[2] = m0.z; mjw[3] = m0.w; mjw[4] = m1.x;
                        vmask = m1.y;
                    }
                    if (pack16 && (vmask & (vmask - 1)) == 0) {
                        // ---- a site with ONE admissible column group (the deep insertion columns: only their gap group is
                        // large enough, 817): the shared coverage is the site sum of the row sums (the five groups of a site
                        // partition its coverage), so one packed shuffle round gives cov and the group's column sum
                        if (vmask == 0) continue;                           // warp-uniform: no admissible group, no pair test
                        const int b1 = __ffs(vmask) - 1;                    // warp-uniform
                        const int cb = b1 == 0 ? c[0] : b1 == 1 ? c[1] : b1 == 2 ? c[2] : b1 == 3 ? c[3] : c[4];
                        const int mj1 = b1 == 0 ? mjw[0] : b1 == 1 ? mjw[1] : b1 == 2 ? mjw[2] : b1 == 3 ? mjw[3] : mjw[4];
                        int cs1;
                        site_sum5_x2(rowsum, cb, base_lane, cov, cs1);
                        const int lnc3 = pair_site ? LT((unsigned)cov) - LT((unsigned)rowsum) - LT((unsigned)(cov - rowsum)) : RR_T1Q_NEVER;
                        float meanfac;
                        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(meanfac) : "f"((float)max(cov, 1)));
                        meanfac *= (float)rowsum;
                        const bool n1 = rr_tier1_q(LT, (unsigned)cb, (unsigned)rowsum, (unsigned)cs1, (unsigned)cov, max(nq_i, mj1), lnc3, meanfac);
                        n_pairs += pair_site;
                        if (__builtin_expect(!__any_sync(0xffffffffu, n1), 1)) continue;
#pragma unroll
                        for (int k = 0; k < 5; k++) { need[k] = n1 && b1 == k; colsum[k] = cs1; }
                    } else {
                        if (pack16) {   // 4 R < 65536 (warp-uniform): three shuffle rounds instead of five
                            site_sum5_x2(c[0], c[1], base_lane, colsum[0], colsum[1]);
                            site_sum5_x2(c[2], c[3], base_lane, colsum[2], colsum[3]);
                            colsum[4] = site_sum5(c[4], base_lane);
                        } else {
#pragma unroll
                            for (int b = 0; b < 5; b++) colsum[b] = site_sum5(c[b], base_lane);
                        }
                        cov = colsum[0] + colsum[1] + colsum[2] + colsum[3] + colsum[4];
                        // rows without a pair test at this site (filters 802 / 804-810) never survive tier 1 (RR_T1Q_NEVER)
                        const int lnc3 = pair_site ? LT((unsigned)cov) - LT((unsigned)rowsum) - LT((unsigned)(cov - rowsum)) : RR_T1Q_NEVER;
                        float meanfac;   // ~ gr1 / cov (only steers where the pmf bound is evaluated; cov = 0 has no pair test)
                        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(meanfac) : "f"((float)max(cov, 1)));
                        meanfac *= (float)rowsum;
                        // tier 0/1 for the admissible column groups of the site, then the queue pushes.  Sites whose
                        // five groups are all admissible (template columns, first insertion columns: ~2/3 of the pair
                        // tests) take a straight-line path so that the 30 table look-ups of the site overlap; the
                        // others skip their inadmissible groups with warp-uniform branches (817).
                        if (vmask == 31) {
#pragma unroll
                            for (int b = 0; b < 5; b++)
                                need[b] = rr_tier1_q(LT, (unsigned)c[b], (unsigned)rowsum, (unsigned)colsum[b], (unsigned)cov,
                                                     max(nq_i, mjw[b]), lnc3, meanfac);
                            n_pairs += pair_site ? 5 : 0;
                            // most sites yield no candidate at all: one vote on the predicates, before they become registers
                            if (__builtin_expect(!__any_sync(0xffffffffu, need[0] || need[1] || need[2] || need[3] || need[4]), 1)) continue;
                        } else {
#pragma unroll
                            for (int b = 0; b < 5; b++) {
                                need[b] = false;
                                if (vmask & (1 << b)) {  // warp-uniform
                                    need[b] = rr_tier1_q(LT, (unsigned)c[b], (unsigned)rowsum, (unsigned)colsum[b], (unsigned)cov,
                                                         max(nq_i, mjw[b]), lnc3, meanfac);
                                    n_pairs += pair_site;
                                }
                            }
                            if (__builtin_expect(!__any_sync(0xffffffffu, need[0] || need[1] || need[2] || need[3] || need[4]), 1)) continue;
                        }
                    }
#pragma unroll
                    for (int b = 0; b < 5; b++) {
                        // pre-seed pass only: a column seen for the first time (no maximum yet) would make every row
                        // of the tile a candidate at once; one row in eight is enough to seed it
                        if (subsample) need[b] &= mjw[b] < RR_T1Q_SLACK || ((lane + jj) & 7) == 0;   // (< SLACK: the group has a maximum)
                        // s at the lower end of the support (s = gr1 + gr2 - cov >= 1): P[X >= s] = 1 and GSL returns
                        // exactly that (its lower-tail sum starts from pdf(s-1) = 0), so the score is 0 and the pair can
                        // change nothing.  Without this test such pairs are candidates for every group whose maximum is
                        // still 0 or tiny - the gap groups of insertion columns, 60 % of the columns of config 2 and
                        // 88 % of its exact evaluations.  (Kept under RR_FLAG_NO_PRUNE: the exhaustive scan.)
                        need[b] &= no_prune || c[b] + cov != rowsum + colsum[b];
                        rr_cand cand;
                        cand.s = (uint32_t)c[b] >> UM_QSHIFT; cand.gr1 = (uint32_t)rowsum >> UM_QSHIFT;
                        cand.gr2 = (uint32_t)colsum[b] >> UM_QSHIFT; cand.cov = (uint32_t)cov >> UM_QSHIFT;
                        cand.gi = gi; cand.gj = 5 * jj + b;
                        rr_queue_push(q1, c1n, need[b], cand, lane);
                        if (c1n >= 32) {
                            rr_drain_tier2(P, LG, q1, c1n, q2, c2n, lane, n_tier2, n_exact, false);
                            // pick up what this and the other warps / CTAs have found meanwhile
                            if (row_ok) nq_i = rr_thr_q(rr_best_value(P.best + gi), no_prune, qscale);
                            // (the other warps of the same sites may do the same at the same time: every value written is a
                            // true maximum of its group, and a limit is read whole or not at all)
                            for (int e = lane; e < UM_WSITES * 5; e += 32) {
                                const int ts = sub + (e / 5) * UM_SUB;
                                if (ts < UM_COL_SITES && TM.site[ts].nq[e % 5] != RR_T1Q_INADMISSIBLE)   // (admissible implies inside the MSA)
                                    TM.site[ts].nq[e % 5] = rr_thr_q(rr_best_value(P.best + 5 * (jsite0 + ts) + (e % 5)), no_prune, qscale);
                            }
                            __syncwarp();
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_cluster(leader_tempty0 + (uint32_t)(acc * sizeof(unsigned long long)));   // the leader's MMA thread waits for both CTAs
                    mbar_arrive(&T->mempty[tb]);
                }
                tile++;
            }
            // (candidates stay queued across units - entries carry their group ids -, so tier 2 always runs 32 at a time)
        }
        rr_drain_tier2(P, LG, q1, c1n, q2, c2n, lane, n_tier2, n_exact, true);

        unsigned long long v0 = n_pairs, v1 = n_exact, v3 = n_units, v4 = n_tier2;
        for (int o = 16; o > 0; o >>= 1) {
            v0 += __shfl_xor_sync(0xffffffffu, v0, o);
            v1 += __shfl_xor_sync(0xffffffffu, v1, o);
            v3 += __shfl_xor_sync(0xffffffffu, v3, o);
            v4 += __shfl_xor_sync(0xffffffffu, v4, o);
        }
        if (lane == 0) {
            if (v0) atomicAdd(P.counters + 0, v0);
            if (v1) atomicAdd(P.counters + 1, v1);
            if (v3) atomicAdd(P.counters + 3, v3);
            if (v4) atomicAdd(P.counters + 4, v4);
        }
    }

    tc_fence_before();
    cluster_sync_all();   // neither CTA leaves while the other may still signal its barriers or read its operands
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(UM_TMEM_COLS) : "memory");
    }
}

// ---- deferred exact evaluation: the candidates the scan has listed (rr_drain_tier2), one per thread ----------------
// Every listed pair survived tiers 1 and 2 against the maxima of its time; by now the maxima hold lower bounds of every
// listed pair's score (and, as this kernel proceeds, exact scores), so tier 2 is asked again before the FP64 series.
__global__ void __launch_bounds__(128) rr_k_deferred_exact(const rr_scan_params P)
{
    const unsigned long long n = min(P.counters[5], P.deferred_cap);
    rr_lnf_global LG;
    LG.gmem = P.lnfact;
    unsigned n_exact = 0, n_tier2 = 0;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (unsigned long long)gridDim.x * blockDim.x) {
        const rr_cand c = rr_cand_unpack(P.deferred[t]);
        const double thr = fmin(rr_best_value(P.best + c.gi), rr_best_value(P.best + c.gj));
        n_tier2++;
        if (!rr_tier2(LG, c.s, c.gr1, c.gr2, c.cov, thr)) continue;
        n_exact++;
        const double Z = rr_positive_significance(P.lnfact, c.s, c.gr1, c.gr2, c.cov, __ldg(P.gsize + c.gi), __ldg(P.gsize + c.gj));
        if (Z > 0.0) {
            if (Z >= rr_best_value(P.best + c.gi)) rr_best_update(P.best, c.gi, Z, c.gj);
            if (Z >= rr_best_value(P.best + c.gj)) rr_best_update(P.best, c.gj, Z, c.gi);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        n_exact += __shfl_xor_sync(0xffffffffu, n_exact, o);
        n_tier2 += __shfl_xor_sync(0xffffffffu, n_tier2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (n_exact) atomicAdd(P.counters + 1, (unsigned long long)n_exact);
        if (n_tier2) atomicAdd(P.counters + 4, (unsigned long long)n_tier2);
    }
}

// ---- A operand: gather the row sites' groups from xb into 32-row slabs -----------------------
__global__ void __launch_bounds__(256) rr_k_build_xa(const int8_t *__restrict__ xb, const int32_t *__restrict__ rowsites,
                                                      int64_t n_rows, int64_t Kp /* bytes per row */, int8_t *__restrict__ xa)
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

static int make_map(CUtensorMap *map, void *base, uint64_t rows, uint64_t Kp, uint32_t box_rows, int mode)
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
    const bool fp4 = mode != 0;
    cuuint64_t strides[1] = {fp4 ? Kp / 2 : Kp};
    cuuint32_t box[2] = {(cuuint32_t)(mode == 2 ? 2 * UM_KB : UM_KB), box_rows};  // 128 bytes of smem per row
    cuuint32_t estr[2] = {1, 1};
    // fp4: packed 4-bit elements in HBM, expanded by the TMA unit to 16 elements per 16-byte chunk (8 data + 8
    // pad bytes) in shared memory - the layout kind::f8f6f4 reads; the box is still 128 bytes wide there
    // mxf4: the packed nibbles go to shared memory as they are (16U4_ALIGN8B), 256 reads per 128-byte row
    CUresult r = encode(map, mode == 2 ? CU_TENSOR_MAP_DATA_TYPE_16U4_ALIGN8B : mode == 1 ? CU_TENSOR_MAP_DATA_TYPE_16U4_ALIGN16B : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { rr_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return RR_E_CUDA; }
    return RR_OK;
}

}  // namespace

struct rr_umma_state {
    int8_t *xb[2] = {nullptr, nullptr};   // [0] int8, [1] packed e2m1 (elements 0 / 2)
    int8_t *xa[2] = {nullptr, nullptr};
    size_t xa_rows_cap[2] = {0, 0};
    int64_t Kp = 0;
    um_unit *d_units = nullptr;
    size_t units_cap = 0;
    uint8_t *d_colmask = nullptr;
    size_t colmask_cap = 0;
    int32_t *d_khi = nullptr, *d_klo = nullptr;
    size_t khi_cap = 0, klo_cap = 0;
    // what was built for the plan last seen (reused while plan_id and operand coding stay the same)
    uint64_t built_plan_id = 0;
    int built_md = -1;
    int n_units = 0, n_seed = 0, n_preseed = 0;
    int64_t executed_ops = 0;
    CUtensorMap map_a, map_b;
};

int rr_umma_available(void) { return 1; }
int rr_umma_row_sites(void) { return UM_ROW_SITES; }
int rr_umma_col_sites(void) { return UM_COL_SITES; }
int rr_umma_kblock(int mode) { return mode == 2 ? 2 * UM_KB : UM_KB; }

void rr_umma_free(rr_umma_state *s)
{
    if (!s) return;
    for (int m = 0; m < 2; m++) { rr_dev_free(s->xb[m]); rr_dev_free(s->xa[m]); }
    rr_dev_free(s->d_units); rr_dev_free(s->d_khi); rr_dev_free(s->d_klo); rr_dev_free(s->d_colmask);
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
    rr_dev_free(*p);
    *p = nullptr;
    if (rr_dev_malloc((void **)p, std::max<size_t>(need, 1) * sizeof(T)) != cudaSuccess) {
        cudaGetLastError();
        rr_set_error("out of device memory (%zu bytes)", need * sizeof(T));
        return RR_E_NOMEM;
    }
    *cap = need;
    return RR_OK;
}

// kernel parameters shared by every launch on the plan last built
static int um_fill_params(rr_umma_state *S, const rr_scan_params &P, const rr_plan &plan, um_params &U, size_t &smem_bytes, bool &all_smem)
{
    U.P = P;
    U.units = S->d_units;
    U.n_units = S->n_units;
    U.k_hi = S->d_khi;
    U.k_lo = S->d_klo;
    U.n_rt = std::max(plan.n_rowblocks, 1);
    U.n_ct = std::max(plan.n_colblocks, 1);
    U.n_cls = plan.n_classes;
    U.colmask = S->d_colmask;
    U.preseed = 0;
    U.dump = nullptr;
    U.dump_rank = 0;
    U.lnf_smem = std::min(std::min(plan.max_cov + 2, P.R + 2), UM_LNF_MAX);
    smem_bytes = UM_LNF_OFF + (size_t)U.lnf_smem * sizeof(int);
    // fixed-point scale of the tier-1 table (rr_tier1_q): the largest power of two that keeps ln(maxcov!) below 2^30
    U.t1q_shift = rr_t1q_shift(rr_lnfact((unsigned)std::max(std::min(plan.max_cov + 1, P.R + 1), 1)));
    U.t1q_scale = std::nextafterf((float)(2.302585092994046 * (double)(1 << U.t1q_shift)), 0.0f);
    all_smem = U.lnf_smem >= plan.max_cov + 1;
    return RR_OK;
}

// One launch over n_units work units: a persistent grid of CTA pairs (clusters of two), as many as the device runs at once
template <bool ALL_SMEM, int MODE, bool DUMP, bool PACK16 = false>
static cudaError_t um_launch_one(int n_units, size_t smem_bytes, cudaStream_t st, const CUtensorMap &map_a, const CUtensorMap &map_b,
                                 const um_params &prm)
{
    // per device, so set on every launch (a few microseconds)
    cudaError_t e = cudaFuncSetAttribute(rr_k_scan_umma<ALL_SMEM, MODE, DUMP, PACK16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UM_SMEM_MAX);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(UM_THREADS);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // how many pairs are resident at once (units are dealt out round robin: a pair that had to wait for a free SM pair would
    // run its share after everybody else's); asked once per device and instantiation
    static int resident_pairs[64] = {0};
    int dev = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    int &cached = resident_pairs[dev & 63];
    if (cached == 0) {
        int n_sm = 0, n = 0;
        if ((e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        cfg.gridDim = dim3((unsigned)(n_sm / 2 * 2));
        if ((e = cudaOccupancyMaxActiveClusters(&n, rr_k_scan_umma<ALL_SMEM, MODE, DUMP, PACK16>, &cfg)) != cudaSuccess) return e;
        cached = std::max(1, std::min(n, n_sm / 2));
        if (getenv("RR_TRACE")) fprintf(stderr, "[rr trace] scan kernel: %d CTA pairs resident on device %d (%d SMs)\n", n, dev, n_sm);
    }
    cfg.gridDim = dim3((unsigned)(2 * std::max(1, std::min(cached, n_units))));
    e = cudaLaunchKernelEx(&cfg, rr_k_scan_umma<ALL_SMEM, MODE, DUMP, PACK16>, map_a, map_b, prm);
    rr_count_launch(1);
    return e != cudaSuccess ? e : cudaGetLastError();
}

static cudaError_t um_launch(int mode, bool all_smem, bool dump, int n_units, size_t smem_bytes, cudaStream_t st, const CUtensorMap &map_a,
                             const CUtensorMap &map_b, const um_params &prm)
{
    if (dump) {   // test hook: any table size, one tile
        return mode == 2 ? um_launch_one<false, 2, true>(n_units, smem_bytes, st, map_a, map_b, prm)
             : mode == 1 ? um_launch_one<false, 1, true>(n_units, smem_bytes, st, map_a, map_b, prm)
                         : um_launch_one<false, 0, true>(n_units, smem_bytes, st, map_a, map_b, prm);
    }
    if (mode == 2 && all_smem && prm.P.R < (65536 >> UM_QSHIFT)) return um_launch_one<true, 2, false, true>(n_units, smem_bytes, st, map_a, map_b, prm);
    if (mode == 2) return all_smem ? um_launch_one<true, 2, false>(n_units, smem_bytes, st, map_a, map_b, prm) : um_launch_one<false, 2, false>(n_units, smem_bytes, st, map_a, map_b, prm);
    if (mode == 1) return all_smem ? um_launch_one<true, 1, false>(n_units, smem_bytes, st, map_a, map_b, prm) : um_launch_one<false, 1, false>(n_units, smem_bytes, st, map_a, map_b, prm);
    return all_smem ? um_launch_one<true, 0, false>(n_units, smem_bytes, st, map_a, map_b, prm) : um_launch_one<false, 0, false>(n_units, smem_bytes, st, map_a, map_b, prm);
}

static cudaError_t rr_launch_deferred_exact(const rr_scan_params &P, int n_sm, cudaStream_t st)
{
    rr_k_deferred_exact<<<n_sm * 16, 128, 0, st>>>(P);
    rr_count_launch(1);
    return cudaGetLastError();
}

int rr_umma_scan(rr_umma_state *&S, int mode, uint64_t plan_id, rr_scan_params &P, rr_plan &plan, int n_sm, cudaStream_t st)
{
    int rc;
    const int fp4 = mode != 0;            // modes 1 and 2 share the packed e2m1 operands
    const int md = fp4 ? 1 : 0;
    if (!S) {
        S = new rr_umma_state();
        S->Kp = ((int64_t)P.R + 2 * UM_KB - 1) / (2 * UM_KB) * (2 * UM_KB);  // whole mxf4 K blocks (256 reads)
        if (S->Kp == 0) S->Kp = 2 * UM_KB;
    }
    const int64_t row_bytes = fp4 ? S->Kp / 2 : S->Kp;
    if (!S->xb[md]) {
        const size_t rows = (size_t)5 * P.N;
        if (rr_dev_malloc((void **)&S->xb[md], std::max<size_t>(rows * row_bytes, 32)) != cudaSuccess) {
            cudaGetLastError();
            rr_set_error("out of device memory for the B operand (%zu bytes)", rows * (size_t)row_bytes);
            return RR_E_NOMEM;
        }
        UM_CUDA(rr_launch_bits_to_operand(P.bits, (int64_t)rows, P.W32, S->xb[md], S->Kp, fp4, st));
    }
    const bool seeding = !(P.flags & (RR_FLAG_NO_PRUNE | RR_FLAG_SKIP_SEED));
    if (S->built_plan_id != plan_id || S->built_md != mode) {
        // A operand for this plan's row sites
        const size_t xa_rows = (size_t)std::max(plan.n_rowblocks, 1) * UM_M;
        if (xa_rows > S->xa_rows_cap[md]) {
            rr_dev_free(S->xa[md]);
            S->xa[md] = nullptr;
            if (rr_dev_malloc((void **)&S->xa[md], xa_rows * row_bytes) != cudaSuccess) {
                cudaGetLastError();
                rr_set_error("out of device memory for the A operand (%zu bytes)", xa_rows * (size_t)row_bytes);
                return RR_E_NOMEM;
            }
            S->xa_rows_cap[md] = xa_rows;
        }
        rr_trace_mark("umma: operands allocated");
        rr_k_build_xa<<<(unsigned)xa_rows, 256, 0, st>>>(S->xb[md], P.rowsites, (int64_t)xa_rows, row_bytes, S->xa[md]);
        rr_count_launch(1);
        UM_CUDA(cudaGetLastError());

        // work units of a CTA pair: (two adjacent row tiles rb, rb + 1 of this part - the last pair of an odd part has
        // one -, aligned chunk of UNIT_CT column tiles over the union of the two tiles' column ranges; the epilogue's own
        // range test keeps a row out of the column tiles that are only the other tile's), ordered in 2-D blocks of GR (48)
        // row tiles x GC (3) chunks so that the units in flight at any time share a working set (GR A row-tiles +
        // GC*UNIT_CT B column-tiles, a few tens of MB) that stays resident in the 126 MB L2.
#ifndef RR_UM_GR
#define RR_UM_GR 48
#define RR_UM_GC 3
#define RR_UM_UNIT_CT 4
#endif
        constexpr int UNIT_CT = RR_UM_UNIT_CT, GR = RR_UM_GR, GC = RR_UM_GC;   // measured best of six block shapes at config 2
        static_assert(GR % 2 == 0, "row-tile pairs do not straddle a block");
        struct keyed { int64_t key; um_unit u; };
        struct colrange { int cb0, cb1; };      // column tiles [cb0, cb1) of a pair of row tiles (rb1 < 0: one tile)
        auto pair_range = [&](int rb0, int rb1) {
            colrange r = {0x7fffffff, -1};
            for (int t : {rb0, rb1}) {
                if (t < 0) continue;
                const int cb0 = plan.unit_cb0[t];
                const int ncb = (int)(plan.unit_prefix[t + 1] - plan.unit_prefix[t]);
                if (ncb <= 0) continue;
                r.cb0 = std::min(r.cb0, cb0);
                r.cb1 = std::max(r.cb1, cb0 + ncb);
            }
            return r;
        };
        auto pair_kblocks = [&](int rb0, int rb1, int cb) {   // K blocks both CTAs of the pair run for column tile cb
            int sum = 0;
            const int nrb = std::max(plan.n_rowblocks, 1), ncb = std::max(plan.n_colblocks, 1);
            for (int c = 0; c < plan.n_classes; c++) {
                int hi = plan.k_hi[(size_t)c * nrb + rb0];
                if (rb1 >= 0) hi = std::max(hi, plan.k_hi[(size_t)c * nrb + rb1]);
                sum += std::max(0, hi - plan.k_lo[(size_t)c * ncb + cb]);
            }
            return sum;
        };
        auto second = [&](int rb, int rb_end) { return rb + 1 < rb_end ? rb + 1 : -1; };   // the main pass pairs adjacent row tiles
        // generated directly in block order (row group, chunk group, row-tile pair, chunk): no sort needed
        std::vector<um_unit> units;
        int64_t kblocks = 0;   // K blocks issued, counted per CTA (a pair's K block is two)
        {
            int cc_min = 0x7fffffff, cc_max = -1;
            for (int rb = plan.rb_lo; rb < plan.rb_hi; rb += 2) {
                const colrange r = pair_range(rb, second(rb, plan.rb_hi));
                if (r.cb1 <= r.cb0) continue;
                cc_min = std::min(cc_min, r.cb0 / UNIT_CT);
                cc_max = std::max(cc_max, (r.cb1 - 1) / UNIT_CT);
                for (int c = r.cb0; c < r.cb1; c++) kblocks += 2 * (int64_t)pair_kblocks(rb, second(rb, plan.rb_hi), c);
            }
            for (int rg = plan.rb_lo; rg < plan.rb_hi; rg += GR)
                for (int cg = cc_max < 0 ? 1 : cc_min / GC; cc_max >= 0 && cg <= cc_max / GC; cg++)
                    for (int rb = rg; rb < std::min(rg + GR, plan.rb_hi); rb += 2) {
                        const colrange r = pair_range(rb, second(rb, plan.rb_hi));
                        if (r.cb1 <= r.cb0) continue;
                        const int c_lo = std::max(cg * GC, r.cb0 / UNIT_CT), c_hi = std::min(cg * GC + GC - 1, (r.cb1 - 1) / UNIT_CT);
                        for (int cc = c_lo; cc <= c_hi; cc++)
                            units.push_back({rb, second(rb, plan.rb_hi), std::max(r.cb0, cc * UNIT_CT), std::min(r.cb1, (cc + 1) * UNIT_CT)});
                    }
        }
        // Seeding pass: the same kernel over every SEED-th row tile of the WHOLE MSA first (a CTA pair takes two of them, SEED
        // row tiles apart: with two adjacent tiles every 2*SEED instead, a column's nearest seed row is twice as far away and
        // the full pass evaluates 2.2 x the candidates - measured).  It leaves
        // true (lower-bound) maxima in best[] for all column groups, so the full pass starts with thresholds close
        // to the final ones instead of 0 and the bounds prune from the first pair on.  Its pair statistics are
        // discarded.  With several parts (GPUs) the seed tiles are split by COLUMN chunk (chunk % parts == part):
        // a column's first visit, where every pair is a candidate, then happens on one GPU only, and the GPUs
        // exchange the seeded maxima (all-reduce MAX) before their full passes.
        // The seeding pass itself starts from zero thresholds; a pre-seed over every PRESEED-th seed pair
        // takes that warm-up on ~1/512 of the row tiles instead of 1/64.
#ifndef RR_UM_SEED
#define RR_UM_SEED 64        // measured with deferred evaluation: 32 -> 109.7 ms, 16 -> 112.7, 128 -> 110.3 against 109.2
#define RR_UM_PRESEED 8      // no pre-seed launch at all (0): 115.4 ms
#endif
        constexpr int SEED = RR_UM_SEED, PRESEED = RR_UM_PRESEED;   // PRESEED 0: no pre-seed launch
        std::vector<um_unit> seed_units, preseed_units;
        if (plan.n_rowblocks >= 2 * SEED) {
            std::vector<keyed> ks;
            for (int rb = SEED / 2; rb < plan.n_rowblocks; rb += 2 * SEED) {
                const int rb1 = rb + SEED < plan.n_rowblocks ? rb + SEED : -1;
                const colrange r = pair_range(rb, rb1);
                if (r.cb1 <= r.cb0) continue;
                for (int cc = r.cb0 / UNIT_CT; cc <= (r.cb1 - 1) / UNIT_CT; cc++) {
                    if (cc % plan.part_count != plan.part_index) continue;
                    // with several parts a GPU owns 1/parts of the chunks: one column tile per unit then, so that the few seed
                    // row tiles still fill its SMs
                    const int c_lo = std::max(r.cb0, cc * UNIT_CT), c_hi = std::min(r.cb1, (cc + 1) * UNIT_CT);
                    const int step = plan.part_count > 1 ? 1 : UNIT_CT;
                    for (int c = c_lo; c < c_hi; c += step) {
                        um_unit un = {rb, rb1, c, std::min(c_hi, c + step)};
                        ks.push_back({((int64_t)(cc / GC) << 32) + ((int64_t)rb << 12) + ((int64_t)(cc % (GC * 64)) << 3) + (c - c_lo), un});
                    }
                }
            }
            std::sort(ks.begin(), ks.end(), [](const keyed &x, const keyed &y) { return x.key < y.key; });
            for (const keyed &k : ks) {
                seed_units.push_back(k.u);
                if (PRESEED > 0 && (k.u.rt0 / (2 * SEED)) % (PRESEED > 0 ? PRESEED : 1) == 0) preseed_units.push_back(k.u);
            }
        }
        seed_units.insert(seed_units.end(), preseed_units.begin(), preseed_units.end());  // stored behind the seed list
        S->executed_ops = kblocks * (int64_t)(2LL * UM_M * UM_N * UM_KB) * (mode == 2 ? 2 : 1);
        S->n_units = (int)units.size();
        S->n_preseed = (int)preseed_units.size();
        S->n_seed = (int)seed_units.size() - S->n_preseed;
        if (!units.empty()) {
            if ((rc = grow(&S->d_units, &S->units_cap, units.size() + seed_units.size()))) return rc;
            if ((rc = grow(&S->d_khi, &S->khi_cap, plan.k_hi.size()))) return rc;
            if ((rc = grow(&S->d_klo, &S->klo_cap, plan.k_lo.size()))) return rc;
            // one byte per column site: which of its groups are admissible column groups (817); padded to whole tiles
            std::vector<uint8_t> colmask((size_t)std::max(plan.n_colblocks, 1) * UM_CMASK_BYTES + 16, 0);
            for (int j = 0; j < P.N; j++)
                for (int b = 0; b < 5; b++) colmask[j] |= (uint8_t)((plan.colok[(size_t)5 * j + b] ? 1 : 0) << b);
            if ((rc = grow(&S->d_colmask, &S->colmask_cap, colmask.size()))) return rc;
            UM_CUDA(cudaMemcpyAsync(S->d_colmask, colmask.data(), colmask.size(), cudaMemcpyHostToDevice, st));
            UM_CUDA(cudaMemcpyAsync(S->d_units, units.data(), sizeof(um_unit) * units.size(), cudaMemcpyHostToDevice, st));
            if (!seed_units.empty())
                UM_CUDA(cudaMemcpyAsync(S->d_units + units.size(), seed_units.data(), sizeof(um_unit) * seed_units.size(),
                                        cudaMemcpyHostToDevice, st));
            UM_CUDA(cudaMemcpyAsync(S->d_khi, plan.k_hi.data(), sizeof(int32_t) * plan.k_hi.size(), cudaMemcpyHostToDevice, st));
            UM_CUDA(cudaMemcpyAsync(S->d_klo, plan.k_lo.data(), sizeof(int32_t) * plan.k_lo.size(), cudaMemcpyHostToDevice, st));
            rr_trace_mark("umma: units built");
            UM_CUDA(cudaStreamSynchronize(st));  // units[] is a local
            rr_trace_mark("umma: uploads synced");
            if ((rc = make_map(&S->map_a, S->xa[md], xa_rows, (uint64_t)S->Kp, UM_M, mode))) return rc;
            if ((rc = make_map(&S->map_b, S->xb[md], (uint64_t)5 * P.N, (uint64_t)S->Kp, UM_BH, mode))) return rc;   // one CTA's half of a column tile
        }
        rr_trace_mark("umma: tensor maps");
        S->built_plan_id = plan_id;
        S->built_md = mode;
    }
    plan.executed_ops = S->executed_ops;
    if (S->n_units == 0) return RR_OK;

    um_params U;
    size_t smem_bytes;
    bool all_smem;
    if ((rc = um_fill_params(S, P, plan, U, smem_bytes, all_smem))) return rc;
    if (seeding && S->n_seed > 0) {
        um_params V = U;
        if (V.P.defer_mode) V.P.defer_mode = 2;   // thresholds only: the seed tiles' pairs come again in the full pass
        if (S->n_preseed > 0) {
            V.units = S->d_units + S->n_units + S->n_seed;
            V.n_units = S->n_preseed;
            V.preseed = 1;  // subsample first-visit columns
            UM_CUDA(um_launch(mode, all_smem, false, V.n_units, smem_bytes, st, S->map_a, S->map_b, V));
            V.preseed = 0;
        }
        V.units = S->d_units + S->n_units;
        V.n_units = S->n_seed;
        UM_CUDA(um_launch(mode, all_smem, false, V.n_units, smem_bytes, st, S->map_a, S->map_b, V));
        UM_CUDA(cudaMemsetAsync(P.counters, 0, sizeof(unsigned long long) * 8, st));
    }
    if (P.flags & RR_FLAG_SEED_ONLY) return RR_OK;
    UM_CUDA(um_launch(mode, all_smem, false, U.n_units, smem_bytes, st, S->map_a, S->map_b, U));
    // the exact scores of the candidates the full pass has listed (rr_device.cuh: defer_mode)
    if (P.defer_mode == 1 && P.deferred_cap > 0) UM_CUDA(rr_launch_deferred_exact(P, n_sm, st));
    return RR_OK;
}

// Test hook (include/rr_debug.h): the raw accumulator of ONE (row tile, column tile) pair as the tcgen05 kernel's
// epilogue reads it from TMEM, through the same producer / MMA code as the scan (DUMP instantiation of the kernel).
// Requires a preceding rr_umma_scan with the same mode and plan (operands, tensor maps and ranges are reused).
int rr_umma_dump_tile(rr_umma_state *S, int mode, uint64_t plan_id, rr_scan_params &P, rr_plan &plan, int rt, int ct,
                      int32_t *d_out /* device, [128][240] */, int n_sm, cudaStream_t st)
{
    int rc;
    if (!S || S->built_plan_id != plan_id || S->built_md != mode || S->n_units == 0) {
        rr_set_error("rr_debug_umma_counts: run rr_scan with this variant first");
        return RR_E_ARG;
    }
    if (rt < 0 || rt >= plan.n_rowblocks || ct < 0 || ct >= plan.n_colblocks) { rr_set_error("rr_debug_umma_counts: tile out of range"); return RR_E_ARG; }
    // odd row tiles are dumped by rank 1 of the pair (rt - 1, rt), even ones by rank 0 of (rt, rt + 1) - or of (rt) alone
    const int dump_rank = rt & 1;
    const int rt0 = rt - dump_rank;
    um_unit one = {rt0, rt0 + 1 < plan.n_rowblocks ? rt0 + 1 : -1, ct, ct + 1}, *d_one = nullptr;
    if (rr_dev_malloc((void **)&d_one, sizeof one) != cudaSuccess) { cudaGetLastError(); rr_set_error("out of device memory"); return RR_E_NOMEM; }
    um_params U;
    size_t smem_bytes;
    bool all_smem;
    rc = um_fill_params(S, P, plan, U, smem_bytes, all_smem);
    cudaError_t e = cudaSuccess;
    if (!rc) {
        U.units = d_one;
        U.n_units = 1;
        U.dump = d_out;
        U.dump_rank = dump_rank;
        U.P.flags |= RR_FLAG_NO_PRUNE;   // thresholds play no part
        e = cudaMemcpyAsync(d_one, &one, sizeof one, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(d_out, 0xff, sizeof(int32_t) * UM_M * UM_N, st);
        if (e == cudaSuccess) e = um_launch(mode, false, true, 1, smem_bytes, st, S->map_a, S->map_b, U);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    rr_dev_free(d_one);
    if (e != cudaSuccess) { rr_set_error("CUDA error %s in rr_debug_umma_counts", cudaGetErrorString(e)); return RR_E_CUDA; }
    return rc;
}

// ---------------------------------------------------------------------------------------------------------------
// Measurement hook (include/rr_debug.h): the tensor pipe's rate for the MMA kind, shape and operand layout the scan
// uses, with nothing else running - one CTA per SM issues back-to-back tcgen05.mma (M = 128, N = 240, both operands
// from the same shared-memory tile, two TMEM accumulators alternating) and nothing is loaded or read back.  This is
// the denominator of the hardware fraction bench.py reports (executed MACs / this rate).
// ---------------------------------------------------------------------------------------------------------------
namespace {
constexpr int PK_BATCH = 16;   // K blocks per commit
constexpr int PK_TILE_BYTES = (UM_M + UM_N) * UM_KB;   // one A tile (128 rows) and one whole B tile (240 rows) of 128 bytes per row
template <int MODE>
__global__ void __launch_bounds__(128, 1) rr_k_mma_peak(int batches)
{
    uint8_t *smem = um_smem;
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem + PK_TILE_BYTES);   // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + PK_TILE_BYTES + 16);
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < PK_TILE_BYTES / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = MODE == 0 ? 0x02020202u : 0x44444444u;
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1); mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes above, async-proxy reads by the MMA
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(UM_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if constexpr (MODE == 2) {
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)UM_SF_COL;
        const uint32_t one = 0x7F7F7F7Fu;
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
            ::"r"(taddr), "r"(one) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc<MODE>();
        const uint32_t sa = smem_u32(smem);
        const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sa + UM_A_BYTES);
        for (int b = 0; b < batches; b++) {
            const uint32_t tmem_d = tmem_base + (uint32_t)((b & 1) * UM_ACC_STRIDE);
            for (int kb = 0; kb < PK_BATCH; kb++)
#pragma unroll
                for (int k = 0; k < UM_KB / 32; k++)
                    tc_mma<MODE>(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u,
                                 tmem_base + UM_SF_COL, tmem_base + UM_SF_COL + 8);
            tc_commit(&bar[b & 1]);
            if (b >= 1) mbar_wait(&bar[(b - 1) & 1], ((b - 1) >> 1) & 1);   // at most two batches in flight
        }
        if (batches >= 1) mbar_wait(&bar[(batches - 1) & 1], ((batches - 1) >> 1) & 1);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(UM_TMEM_COLS) : "memory");
    }
}

// the same for a CTA pair: the leader issues tcgen05.mma.cta_group::2 (M = 256: 128 rows per CTA; N = n_cols, each CTA holding
// n_cols / 2 rows of B), nothing loaded, nothing read back
template <int MODE>
__global__ void __launch_bounds__(128, 1) rr_k_mma_peak_pair(int batches, int n_cols)
{
    uint8_t *smem = um_smem;
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem + PK_TILE_BYTES);   // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + PK_TILE_BYTES + 16);
    const int warp = threadIdx.x >> 5;
    const int rank = (int)cluster_ctarank();
    for (int i = threadIdx.x; i < PK_TILE_BYTES / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = MODE == 0 ? 0x02020202u : 0x44444444u;
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1); mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(UM_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if constexpr (MODE == 2) {
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)UM_SF_COL;
        const uint32_t one = 0x7F7F7F7Fu;
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
            ::"r"(taddr), "r"(one) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    if (threadIdx.x == 0 && rank == 0) {
        const uint32_t idesc = (make_idesc<MODE, 2 * UM_M>() & ~(0x3Fu << 17)) | ((uint32_t)(n_cols >> 3) << 17);
        const uint32_t sa = smem_u32(smem);
        const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sa + UM_A_BYTES);
        for (int b = 0; b < batches; b++) {
            const uint32_t tmem_d = tmem_base + (uint32_t)((b & 1) * UM_ACC_STRIDE);
            for (int kb = 0; kb < PK_BATCH; kb++)
#pragma unroll
                for (int k = 0; k < UM_KB / 32; k++)
                    tc_mma_pair<MODE>(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u,
                                      tmem_base + UM_SF_COL, tmem_base + UM_SF_COL + 8);
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         ::"r"(smem_u32(&bar[b & 1])), "h"((uint16_t)1) : "memory");
            if (b >= 1) mbar_wait(&bar[(b - 1) & 1], ((b - 1) >> 1) & 1);
        }
        if (batches >= 1) mbar_wait(&bar[(batches - 1) & 1], ((batches - 1) >> 1) & 1);
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(UM_TMEM_COLS) : "memory");
    }
}

template <int MODE>
cudaError_t mma_peak_pair_launch(int grid, int batches, int n_cols, cudaStream_t st)
{
    const int smem_bytes = PK_TILE_BYTES + 64;
    cudaError_t e = cudaFuncSetAttribute(rr_k_mma_peak_pair<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(grid / 2 * 2));
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, rr_k_mma_peak_pair<MODE>, batches, n_cols);
    rr_count_launch(1);
    return e != cudaSuccess ? e : cudaGetLastError();
}

template <int MODE>
cudaError_t mma_peak_launch(int grid, int batches, cudaStream_t st)
{
    const int smem_bytes = PK_TILE_BYTES + 64;
    cudaError_t e = cudaFuncSetAttribute(rr_k_mma_peak<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    rr_k_mma_peak<MODE><<<grid, 128, smem_bytes, st>>>(batches);
    rr_count_launch(1);
    return cudaGetLastError();
}
}  // namespace

// CTA pairs, N = n_cols (a multiple of 16, 32..256): returns the MACs one launch executes through *macs
cudaError_t rr_umma_mma_peak_pair(int mode, int n_sm, int kblocks_per_sm, int n_cols, double *macs, cudaStream_t st)
{
    const int batches = std::max(1, kblocks_per_sm / PK_BATCH);
    *macs = (double)(n_sm / 2 * 2) * batches * PK_BATCH * (double)UM_M * n_cols * UM_KB * (mode == 2 ? 2 : 1);
    return mode == 2 ? mma_peak_pair_launch<2>(n_sm, batches, n_cols, st) : mode == 1 ? mma_peak_pair_launch<1>(n_sm, batches, n_cols, st)
                                                                                         : mma_peak_pair_launch<0>(n_sm, batches, n_cols, st);
}

// mode as in rr_umma_scan; returns the MACs one launch executes (all CTAs) through *macs
cudaError_t rr_umma_mma_peak(int mode, int n_sm, int kblocks_per_sm, double *macs, cudaStream_t st)
{
    const int batches = std::max(1, kblocks_per_sm / PK_BATCH);
    *macs = (double)n_sm * batches * PK_BATCH * (double)UM_M * UM_N * UM_KB * (mode == 2 ? 2 : 1);
    return mode == 2 ? mma_peak_launch<2>(n_sm, batches, st) : mode == 1 ? mma_peak_launch<1>(n_sm, batches, st) : mma_peak_launch<0>(n_sm, batches, st);
}
