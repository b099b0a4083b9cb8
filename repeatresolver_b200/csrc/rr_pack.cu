// rr_pack.cu -- device-side packing of the MSA: the second half of Einlesen
// (/root/reference/MaxCorrelation.c:304-385) as HBM-bound byte kernels.
//
//   cells[R][N] (raw characters or 0..5 codes, row-major, as read from the file)
//     -> per-row covered span                       (feeds the row order and the first-break sweep)
//     -> bits[5N][W32]   group bitsets  (Groups,        348-351, 374) bit = rank of the row
//     -> covbits[N][W32] coverage sets  (LocalCoverage, 354-357, 380)  in span-start order
//     -> gsize[5N], coverage[N]         (Groupsizearray 385, Coverage 364-383)
//     -> xb[5N][Kp] / xa[slabs*32][Kp]  0/1 int8 operands of the tcgen05 variant (K-major)
//
// Character classes are the reference's (304-329): aA->0 cC->1 gG->2 tT->3 '-' '_'->4, all
// else -> 5 (not covered).
#include "rr_kernels.h"

// rr_classify: rr_kernels.h (shared with the signature kernel of rr_kmeans.cu)

// ---- per-row span: first / last covered column and number of covered cells ------------
__global__ void __launch_bounds__(256) rr_k_row_spans(const uint8_t *__restrict__ cells, int R, int N, int codes,
                                                       int32_t *__restrict__ start, int32_t *__restrict__ end,
                                                       int32_t *__restrict__ ncov)
{
    const int r = blockIdx.x;
    if (r >= R) return;
    const uint8_t *row = cells + (size_t)r * N;
    int first = 0x7fffffff, last = -1, cnt = 0;
    for (int c = threadIdx.x; c < N; c += blockDim.x) {
        if (rr_classify(row[c], codes) < 5) {
            first = min(first, c);
            last = max(last, c);
            cnt++;
        }
    }
    __shared__ int s_first[8], s_last[8], s_cnt[8];
    for (int o = 16; o > 0; o >>= 1) {
        first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
        last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { s_first[w] = first; s_last[w] = last; s_cnt[w] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; i++) { first = min(first, s_first[i]); last = max(last, s_last[i]); cnt += s_cnt[i]; }
        start[r] = first; end[r] = last; ncov[r] = cnt;
    }
}

// ---- cells -> bitsets --------------------------------------------------------------------
// One block: 128 ranks (4 u32 words of every bitset) x 128 columns.  The tile is staged in
// shared memory with coalesced row-segment reads, then one thread per column builds the
// 6 x 4 words and stores them as 16-byte vectors.
constexpr int PK_ROWS = 128;
constexpr int PK_COLS = 128;

// cells holds the rows [row_lo, row_hi) of the MSA only (a rank's slice of a row-sliced pack; the whole MSA for
// row_lo = 0, row_hi = R): ranks whose row lies outside the slice contribute no bit, so the buffers of all slices OR
// (= add) to the full bitsets.
__global__ void __launch_bounds__(PK_COLS) rr_k_pack_bits(const uint8_t *__restrict__ cells,
                                                           const int32_t *__restrict__ perm, int R, int N, int codes,
                                                           uint32_t *__restrict__ bits, uint32_t *__restrict__ covbits,
                                                           int W32, int row_lo, int row_hi)
{
    __shared__ uint8_t tile[PK_ROWS][PK_COLS + 4];
    const int c0 = blockIdx.x * PK_COLS;
    const int r0 = blockIdx.y * PK_ROWS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int rr = warp; rr < PK_ROWS; rr += PK_COLS / 32) {
        const int rank = r0 + rr;
        const int src = rank < R ? perm[rank] : -1;
        const uint8_t *row = src >= row_lo && src < row_hi ? cells + (size_t)(src - row_lo) * N : nullptr;
#pragma unroll
        for (int q = 0; q < PK_COLS / 32; q++) {
            const int c = c0 + q * 32 + lane;
            int code = 5;
            if (row && c < N) code = rr_classify(row[c], codes);
            tile[rr][q * 32 + lane] = (uint8_t)code;
        }
    }
    __syncthreads();
    const int col = c0 + threadIdx.x;
    if (col >= N) return;
    uint32_t w[6][4];
#pragma unroll
    for (int k = 0; k < 6; k++)
#pragma unroll
        for (int q = 0; q < 4; q++) w[k][q] = 0u;
#pragma unroll
    for (int q = 0; q < 4; q++) {
#pragma unroll 8
        for (int b = 0; b < 32; b++) {
            const int code = tile[q * 32 + b][threadIdx.x];
            const uint32_t bit = 1u << b;
#pragma unroll
            for (int k = 0; k < 5; k++) w[k][q] |= (code == k) ? bit : 0u;
            w[5][q] |= (code < 5) ? bit : 0u;
        }
    }
    const size_t wo = (size_t)blockIdx.y * 4;
#pragma unroll
    for (int k = 0; k < 5; k++)
        *reinterpret_cast<uint4 *>(bits + ((size_t)5 * col + k) * W32 + wo) = make_uint4(w[k][0], w[k][1], w[k][2], w[k][3]);
    *reinterpret_cast<uint4 *>(covbits + (size_t)col * W32 + wo) = make_uint4(w[5][0], w[5][1], w[5][2], w[5][3]);
}

// ---- bitset sizes: one warp per bitset ----------------------------------------------------
__global__ void __launch_bounds__(256) rr_k_bitset_sizes(const uint32_t *__restrict__ sets, int64_t nsets, int W32,
                                                          int32_t *__restrict__ sizes)
{
    const int64_t g = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (g >= nsets) return;
    const uint32_t *p = sets + (size_t)g * W32;
    int n = 0;
    for (int w = threadIdx.x & 31; w < W32; w += 32) n += __popc(p[w]);
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if ((threadIdx.x & 31) == 0) sizes[g] = n;
}

// ---- the four counts of PositiveSignificance (423-426) for explicit pairs ----------------
__global__ void __launch_bounds__(256) rr_k_pair_counts(const uint32_t *__restrict__ bits,
                                                         const uint32_t *__restrict__ covbits, int W32, int64_t n,
                                                         const int32_t *__restrict__ gi, const int32_t *__restrict__ gj,
                                                         int32_t *__restrict__ out)
{
    const int64_t t = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (t >= n) return;
    const int i = gi[t], j = gj[t];
    int s = 0, g1 = 0, g2 = 0, cv = 0;
    if (i >= 0 && j >= 0) {
        const uint32_t *a = bits + (size_t)i * W32, *b = bits + (size_t)j * W32;
        const uint32_t *ca = covbits + (size_t)(i / 5) * W32, *cb = covbits + (size_t)(j / 5) * W32;
        for (int w = threadIdx.x & 31; w < W32; w += 32) {
            const uint32_t x = a[w], y = b[w], cx = ca[w], cy = cb[w];
            s += __popc(x & y);
            g1 += __popc(x & cy);
            g2 += __popc(y & cx);
            cv += __popc(cx & cy);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        g1 += __shfl_xor_sync(0xffffffffu, g1, o);
        g2 += __shfl_xor_sync(0xffffffffu, g2, o);
        cv += __shfl_xor_sync(0xffffffffu, cv, o);
    }
    if ((threadIdx.x & 31) == 0) {
        out[4 * t + 0] = s; out[4 * t + 1] = g1; out[4 * t + 2] = g2; out[4 * t + 3] = cv;
    }
}

// ---- general first-break (MaxCorrelation.c:804-810) for MSAs whose rows are not single
// spans: one warp per site walks jj = ii+20, ii+21, ... and stops at the first column whose
// shared coverage with ii is below mincov.  Exact for any input; slow path.
__global__ void __launch_bounds__(256) rr_k_general_break(const uint32_t *__restrict__ covbits, int W32, int N,
                                                           int mincov, int32_t *__restrict__ breakcol)
{
    const int ii = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (ii >= N) return;
    const int lane = threadIdx.x & 31;
    const uint32_t *ci = covbits + (size_t)ii * W32;
    int jj = ii + 20;
    for (; jj < N; jj++) {
        const uint32_t *cj = covbits + (size_t)jj * W32;
        int n = 0;
        for (int w = lane; w < W32; w += 32) n += __popc(ci[w] & cj[w]);
        for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
        if (n < mincov) break;
    }
    if (lane == 0) breakcol[ii] = jj < ii + 20 ? ii + 20 : jj;
}

// ---- bitsets -> 0/2 operand of the tcgen05 variant, K-major: x[group][rank] ---------------------------------------
// (elements are 0 / 2 so that a product is 4 = sizeof(float): rr_scan_umma.cu).  One warp per group bitset; a lane
// expands one u32 word (32 ranks) per step into 32 int8 (two 16-byte stores) or 32 packed e2m1 nibbles (one 16-byte
// store; 2.0 = 0b0100, the lower rank in the low nibble).  HBM-bound: reads 5N W32 4 bytes, writes 5N Kp (or Kp / 2).
__device__ __forceinline__ uint32_t rr_spread8_nibbles(uint32_t b)   // bit i of b -> bit 4 i
{
    uint32_t t = b & 0xffu;
    t = (t | (t << 12)) & 0x000F000Fu;
    t = (t | (t << 6)) & 0x03030303u;
    t = (t | (t << 3)) & 0x11111111u;
    return t;
}
__device__ __forceinline__ uint32_t rr_spread4_bytes(uint32_t b)     // bit i of b -> bit 8 i
{
    uint32_t t = b & 0xfu;
    t = (t | (t << 14)) & 0x00030003u;
    t = (t | (t << 7)) & 0x01010101u;
    return t;
}

__global__ void __launch_bounds__(256) rr_k_bits_to_operand(const uint32_t *__restrict__ bits, int64_t nsets, int W32,
                                                             int8_t *__restrict__ xb, int64_t Kp, int fp4)
{
    const int64_t g = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (g >= nsets) return;
    const uint32_t *p = bits + (size_t)g * W32;
    const int words = (int)(Kp / 32);                 // Kp is a multiple of 256; words >= W32, the tail is zero
    if (fp4) {
        uint4 *out = reinterpret_cast<uint4 *>(xb + (size_t)g * (Kp / 2));
        for (int w = threadIdx.x & 31; w < words; w += 32) {
            const uint32_t x = w < W32 ? p[w] : 0u;
            out[w] = make_uint4(rr_spread8_nibbles(x) << 2, rr_spread8_nibbles(x >> 8) << 2, rr_spread8_nibbles(x >> 16) << 2,
                                rr_spread8_nibbles(x >> 24) << 2);
        }
    } else {
        uint4 *out = reinterpret_cast<uint4 *>(xb + (size_t)g * Kp);
        for (int w = threadIdx.x & 31; w < words; w += 32) {
            const uint32_t x = w < W32 ? p[w] : 0u;
            out[2 * w] = make_uint4(rr_spread4_bytes(x) << 1, rr_spread4_bytes(x >> 4) << 1, rr_spread4_bytes(x >> 8) << 1,
                                    rr_spread4_bytes(x >> 12) << 1);
            out[2 * w + 1] = make_uint4(rr_spread4_bytes(x >> 16) << 1, rr_spread4_bytes(x >> 20) << 1, rr_spread4_bytes(x >> 24) << 1,
                                        rr_spread4_bytes(x >> 28) << 1);
        }
    }
}

// word-wise OR of two bitset buffers (the merge step of a row-sliced pack inside one process)
__global__ void __launch_bounds__(256) rr_k_or_words(uint4 *__restrict__ dst, const uint4 *__restrict__ src, int64_t n16)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) {
        uint4 a = dst[i];
        const uint4 b = src[i];
        a.x |= b.x; a.y |= b.y; a.z |= b.z; a.w |= b.w;
        dst[i] = a;
    }
}

#ifndef RR_CPU_EMU   // tests/emu compiles the kernels above with a host compiler; the launch syntax below is nvcc only
// ---- launchers -------------------------------------------------------------------------
cudaError_t rr_launch_row_spans(const uint8_t *cells, int R, int N, int codes, int32_t *start, int32_t *end,
                                int32_t *ncov, cudaStream_t st)
{
    if (R > 0) rr_k_row_spans<<<R, 256, 0, st>>>(cells, R, N, codes, start, end, ncov);
    rr_count_launch(1);
    return cudaGetLastError();
}

cudaError_t rr_launch_pack_bits(const uint8_t *cells, const int32_t *perm, int R, int N, int codes, uint32_t *bits,
                                uint32_t *covbits, int W32, int row_lo, int row_hi, cudaStream_t st)
{
    if (N <= 0 || W32 <= 0) return cudaSuccess;
    dim3 grid((unsigned)((N + PK_COLS - 1) / PK_COLS), (unsigned)(W32 / 4));
    rr_k_pack_bits<<<grid, PK_COLS, 0, st>>>(cells, perm, R, N, codes, bits, covbits, W32, row_lo, row_hi);
    rr_count_launch(1);
    return cudaGetLastError();
}

cudaError_t rr_launch_bitset_sizes(const uint32_t *sets, int64_t nsets, int W32, int32_t *sizes, cudaStream_t st)
{
    if (nsets <= 0) return cudaSuccess;
    rr_k_bitset_sizes<<<(unsigned)((nsets + 7) / 8), 256, 0, st>>>(sets, nsets, W32, sizes);
    rr_count_launch(1);
    return cudaGetLastError();
}

cudaError_t rr_launch_pair_counts(const uint32_t *bits, const uint32_t *covbits, int W32, int64_t n,
                                  const int32_t *gi, const int32_t *gj, int32_t *out, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    rr_k_pair_counts<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(bits, covbits, W32, n, gi, gj, out);
    rr_count_launch(1);
    return cudaGetLastError();
}

cudaError_t rr_launch_general_break(const uint32_t *covbits, int W32, int N, int mincov, int32_t *breakcol,
                                    cudaStream_t st)
{
    if (N <= 0) return cudaSuccess;
    rr_k_general_break<<<(N + 7) / 8, 256, 0, st>>>(covbits, W32, N, mincov, breakcol);
    rr_count_launch(1);
    return cudaGetLastError();
}

cudaError_t rr_launch_bits_to_operand(const uint32_t *bits, int64_t nsets, int W32, int8_t *xb, int64_t Kp, int fp4, cudaStream_t st)
{
    if (nsets <= 0) return cudaSuccess;
    rr_k_bits_to_operand<<<(unsigned)((nsets + 7) / 8), 256, 0, st>>>(bits, nsets, W32, xb, Kp, fp4);
    rr_count_launch(1);
    return cudaGetLastError();
}

cudaError_t rr_launch_or_words(void *dst, const void *src, size_t bytes, cudaStream_t st)
{
    if (bytes == 0) return cudaSuccess;
    rr_k_or_words<<<1184, 256, 0, st>>>(reinterpret_cast<uint4 *>(dst), reinterpret_cast<const uint4 *>(src), (int64_t)(bytes / 16));
    rr_count_launch(1);
    return cudaGetLastError();
}
#endif
