// rr_relvars.cu -- the all-pairs step of Relative_Vars (/root/reference/RepeatResolver.c:2455-2476) on the device
// (SURVEY.md section 8f, row 3), the path behind rr_relative_vars and rr_relative_vars_packed; parity on a B200 in
// tests/test_zz_gpu_relvars.py, the same source under the CPU emulation of tests/emu.
//
// Input is either the packed copy of the part's rows (rr_pack of the reads with Unterteilung == u_no, umask == NULL): a
// group bitset is then already G & U, its size |G & U|, and |Gi & Gj & U| (Triple_Schnitt 150-161) a plain AND+POPC of two
// rows of `bits`; or the packed copy of the WHOLE MSA, as it sits on the device after the scan, with the part as a bitset
// `umask` over the same row order: the mask is ANDed into one operand while it is staged, rr_k_masked_sizes gives |G & U|.
// One block = a 64 x 64 tile of (earlier group a, later group b) of the selected list; the 2 x 64 bitsets go through
// shared memory 32 words at a time (coalesced 128-byte row segments), a thread accumulates a 4 x 4 block of counts in
// registers (POPC-issue bound, every staged word is used 64 times).  Tail, per pair with b at least 100 group ids after
// a (2461): Z <= -log10 pmf(s) for the two-sided score (both tails contain the point s), so pairs whose bound is at most
// `cutoff` are dropped; the rest get the exact score (rr_relative_significance, GSL's operation order in IEEE double).
// Z > cutoff marks both groups (2467-2471).  exp/log10 differ from the host libm by a few ulp, so a pair within 1e-9 of
// the cutoff is not decided here: it goes to a list the host re-evaluates with its own libm.
#include "rr_kernels.h"
#include "rr_score.h"

constexpr int RV_TILE = 64;

__global__ void __launch_bounds__(256)
rr_k_relvars_pairs(const uint32_t *__restrict__ bits, const uint32_t *__restrict__ umask, int W32, const int32_t *__restrict__ sel, int nsel,
                   const int32_t *__restrict__ first_partner, const int32_t *__restrict__ gsize_u, int cov_u,
                   const double *__restrict__ lnf, double cutoff, unsigned char *mark,
                   int4 *__restrict__ unsure, unsigned int unsure_cap, unsigned int *__restrict__ unsure_count)
{
    if (blockIdx.x < blockIdx.y) return;                       // x = tile of the later group: upper triangle only
    __shared__ uint32_t A[RV_TILE][33], B[RV_TILE][33];        // +1: the 16 rows a half-warp reads fall into 16 banks
    const int a0 = blockIdx.y * RV_TILE, b0 = blockIdx.x * RV_TILE;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // the tile has no admissible pair if even its last b comes before the first partner of its first a
    if (min(b0 + RV_TILE, nsel) - 1 < first_partner[a0]) return;
    unsigned s[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) s[i][j] = 0u;
    const int nchunks = (W32 + 31) >> 5;
    for (int c = 0; c < nchunks; c++) {
        const int w = c * 32 + lane;
        __syncthreads();
        for (int r = warp; r < RV_TILE; r += 8) {              // a warp stages one row segment of each operand
            const int a = a0 + r, b = b0 + r;
            const uint32_t u = w < W32 ? (umask ? __ldg(umask + w) : 0xffffffffu) : 0u;   // G & U: the mask goes into one operand
            A[r][lane] = (a < nsel && w < W32) ? __ldg(bits + (size_t)sel[a] * W32 + w) & u : 0u;
            B[r][lane] = (b < nsel && w < W32) ? __ldg(bits + (size_t)sel[b] * W32 + w) : 0u;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < 32; k++) {
            uint32_t x[4], y[4];
#pragma unroll
            for (int i = 0; i < 4; i++) { x[i] = A[i * 16 + ty][k]; y[i] = B[i * 16 + tx][k]; }
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) s[i][j] += __popc(x[i] & y[j]);
        }
    }
    const double margin = 1e-9 * fmax(1.0, fabs(cutoff));
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int a = a0 + i * 16 + ty, b = b0 + j * 16 + tx;
            if (a >= nsel || b >= nsel || b < first_partner[a]) continue;               // 2461: j >= i + 100
            // mark[] is racy by design: byte flags that are only ever set to 1, read here as a hint (a stale 0 costs a
            // redundant evaluation, never a wrong result)
            if (mark[a] && mark[b]) continue;                                           // the pair could only set marks that are set
            // Relative_Group_Significance(Groups[j], Groups[i], U): Group1 = the later group (2465)
            const unsigned gr1 = (unsigned)gsize_u[sel[b]], gr2 = (unsigned)gsize_u[sel[a]], sc = s[i][j];
            if (gr1 == 0u || gr2 == 0u) continue;                                       // 517: score 0, never above a cutoff >= 0
            const double bound = -RR_LOG10E * rr_hyper_lnpdf_unchecked(lnf, sc, gr2, (unsigned)cov_u - gr2, gr1) + 1e-6;
            if (bound <= cutoff - margin) continue;
            const double Z = rr_relative_significance(lnf, sc, gr1, gr2, (unsigned)cov_u);
            if (Z > cutoff + margin) { mark[a] = 1; mark[b] = 1; }
            else if (Z > cutoff - margin) {
                const unsigned idx = atomicAdd(unsure_count, 1u);
                if (idx < unsure_cap) unsure[idx] = make_int4(a, b, (int)sc, 0);
            }
        }
}

// |G & U| for every group (2449: Schnitt(U_Group, Groups[i])): one warp per bitset, as rr_k_bitset_sizes with a mask
__global__ void __launch_bounds__(256) rr_k_masked_sizes(const uint32_t *__restrict__ bits, const uint32_t *__restrict__ umask,
                                                          int64_t nsets, int W32, int32_t *__restrict__ sizes)
{
    const int64_t g = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (g >= nsets) return;
    const uint32_t *p = bits + (size_t)g * W32;
    unsigned n = 0;
    for (int w = threadIdx.x & 31; w < W32; w += 32) n += __popc(__ldg(p + w) & __ldg(umask + w));
    n = __reduce_add_sync(0xffffffffu, n);
    if ((threadIdx.x & 31) == 0) sizes[g] = (int32_t)n;
}

#ifndef RR_CPU_EMU   // tests/emu compiles the kernels above with a host compiler; the launch syntax below is nvcc only
cudaError_t rr_launch_masked_sizes(const uint32_t *bits, const uint32_t *umask, int64_t nsets, int W32, int32_t *sizes, cudaStream_t st)
{
    if (nsets <= 0) return cudaSuccess;
    rr_k_masked_sizes<<<(unsigned)((nsets + 7) / 8), 256, 0, st>>>(bits, umask, nsets, W32, sizes);
    rr_count_launch(1);
    return cudaGetLastError();
}

cudaError_t rr_launch_relvars_pairs(const uint32_t *bits, const uint32_t *umask, int W32, const int32_t *sel, int nsel, const int32_t *first_partner,
                                    const int32_t *gsize_u, int cov_u, const double *lnf, double cutoff, unsigned char *mark,
                                    int4 *unsure, unsigned int unsure_cap, unsigned int *unsure_count, cudaStream_t st)
{
    if (nsel <= 0) return cudaSuccess;
    const unsigned nt = (unsigned)((nsel + RV_TILE - 1) / RV_TILE);
    rr_k_relvars_pairs<<<dim3(nt, nt), 256, 0, st>>>(bits, umask, W32, sel, nsel, first_partner, gsize_u, cov_u, lnf, cutoff, mark, unsure,
                                                     unsure_cap, unsure_count);
    rr_count_launch(1);
    return cudaGetLastError();
}
#endif
