// rr_kmeans.cu -- the read x read sweeps of Kmeans (/root/reference/RepeatResolver.c:2604-2821) on the device
// ("the transposed Gram problem"); parity on a B200 in tests/test_zz_gpu_kmeans.py, the same
// source under the CPU emulation of tests/emu.  The integer rules are shared with the host through rr_kmeans.h and pinned
// there against the unmodified reference (tests/test_oracle_kmeans.py).
//
// sig[anzahl][scv]: the signatures of the part's reads over the selected groups (64-bit words, padding 0).
//   rr_k_km_signatures one warp per (read, 32 groups): the signatures from the part's rows (2626-2650)
//   rr_k_km_pair_scores GrMatch (163-175) of every (read i, row j of X) pair of a panel of rows j, X = the signatures (first
//                      sweep), the centroids (second sweep) or the centroids of the clusters the dissolution can ask for: a
//                      64 x 64 tile of pairs per block, the words of both sides staged through shared memory (each word read
//                      from L2 once per 64 pairs), 4 x 4 pairs per thread, XOR + POPC
//   rr_k_km_top5_seq   one thread per read i over a panel's scores in read order: the reference's five-slot rule (2662-2692),
//                      which depends on the order of the reads and so stays sequential per read; slots kept in HBM between panels
//   rr_k_km_centroids  one thread per (read, word): bitwise majority of the five kept reads' signatures (2697-2705)
//   rr_k_km_assign_seq one thread per read i over a panel's scores: the first best centroid of another read (2709-2725)
// Before (round 2, first version): one thread per read walked all reads itself, tiles of three signatures through shared
// memory and its own signature from L2 for every one of them - 6 SMs busy, tens of milliseconds per sweep.
#include <algorithm>
#include "rr_kernels.h"
#include "rr_kmeans.h"

constexpr int KM_TILE = 64;      // reads / rows of X per block tile
constexpr int KM_ZC = 32;        // signature words staged per step

// Signatures (2626-2650): bit j of read i = the read carries the symbol of group vars[j] at that group's site.  One warp per
// (read, 32 selected groups): the group ids are read coalesced (the same for every read: L2), the cells are byte gathers
// from the read's row (cols bytes, L1-resident), one ballot gives half a signature word.  Every 32-bit half of sig is
// written, bits beyond n_vars as 0.
__global__ void __launch_bounds__(256)
rr_k_km_signatures(const uint8_t *__restrict__ rows, int cols, int codes, const int32_t *__restrict__ vars, int n_vars, int anzahl,
                   int scv, uint32_t *__restrict__ sig32 /*[anzahl][2 * scv]*/)
{
    const int i = blockIdx.y;
    const int half = blockIdx.x * 8 + (threadIdx.x >> 5);                // 32-bit half word of the signature
    if (i >= anzahl || half >= 2 * scv) return;                          // warp-uniform
    const int j = half * 32 + (threadIdx.x & 31);
    int bit = 0;
    if (j < n_vars) {
        const int g = vars[j];
        bit = rr_classify(rows[(size_t)i * cols + g / 5], codes) == g % 5;
    }
    const unsigned word = __ballot_sync(0xffffffffu, bit);
    if ((threadIdx.x & 31) == 0) sig32[(size_t)i * 2 * scv + half] = word;
}

// scores of the pairs (read i, row j of X), j in [0, nj): out[i * stride_i + j * stride_j] = 64 * scv - Hamming distance.
// X row j is xrows[j] if xrows is given (the dissolution's candidate clusters), else j0 + j.
__global__ void __launch_bounds__(256)
rr_k_km_pair_scores(const uint64_t *__restrict__ sig, const uint64_t *__restrict__ X, const int32_t *__restrict__ xrows, int j0, int nj,
                    int anzahl, int scv, int32_t *__restrict__ out, int64_t stride_i, int64_t stride_j)
{
    __shared__ uint64_t sa[KM_TILE][KM_ZC + 1], sb[KM_TILE][KM_ZC + 1];   // +1: rows 16 apart fall into different banks
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int i0 = blockIdx.x * KM_TILE, jt0 = blockIdx.y * KM_TILE;
    int acc[4][4];
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int l = 0; l < 4; l++) acc[k][l] = 0;
    for (int z0 = 0; z0 < scv; z0 += KM_ZC) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < KM_TILE * KM_ZC; idx += 256) {  // consecutive threads, consecutive words of a row
            const int r = idx / KM_ZC, z = idx - r * KM_ZC;
            const bool zin = z0 + z < scv;
            const int i = i0 + r, j = jt0 + r;
            sa[r][z] = (zin && i < anzahl) ? sig[(size_t)i * scv + z0 + z] : 0;
            uint64_t b = 0;
            if (zin && j < nj) b = X[(size_t)(xrows ? xrows[j] : j0 + j) * scv + z0 + z];
            sb[r][z] = b;
        }
        __syncthreads();
#pragma unroll 4
        for (int z = 0; z < KM_ZC; z++) {
            uint64_t a[4], b[4];
#pragma unroll
            for (int k = 0; k < 4; k++) { a[k] = sa[tx + 16 * k][z]; b[k] = sb[ty + 16 * k][z]; }
#pragma unroll
            for (int k = 0; k < 4; k++)
#pragma unroll
                for (int l = 0; l < 4; l++) acc[k][l] += rr_km_popc64(a[k] ^ b[l]);
        }
    }
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int l = 0; l < 4; l++) {
            const int i = i0 + tx + 16 * k, j = jt0 + ty + 16 * l;
            if (i < anzahl && j < nj) out[(int64_t)i * stride_i + (int64_t)j * stride_j] = scv * 64 - acc[k][l];
        }
}

// the five-slot rule over the rows j0 .. j0 + nj of a panel P[nj][anzahl] (read i fastest); slots bs / bj [anzahl][5] live in
// HBM between panels (zero before the first)
__global__ void __launch_bounds__(128)
rr_k_km_top5_seq(const int32_t *__restrict__ P, int anzahl, int j0, int nj, int32_t *__restrict__ best_s, int32_t *__restrict__ best_j)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= anzahl) return;
    int bs[5], bj[5];
#pragma unroll
    for (int k = 0; k < 5; k++) { bs[k] = best_s[(size_t)i * 5 + k]; bj[k] = best_j[(size_t)i * 5 + k]; }
    for (int j = 0; j < nj; j++) rr_km_top5_step(bs, bj, P[(size_t)j * anzahl + i], j0 + j);
#pragma unroll
    for (int k = 0; k < 5; k++) { best_s[(size_t)i * 5 + k] = bs[k]; best_j[(size_t)i * 5 + k] = bj[k]; }
}

__global__ void __launch_bounds__(256)
rr_k_km_centroids(const uint64_t *__restrict__ sig, const int32_t *__restrict__ best_j, int anzahl, int scv,
                  uint64_t *__restrict__ cen)
{
    const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (t >= (int64_t)anzahl * scv) return;
    const int i = (int)(t / scv), z = (int)(t - (int64_t)i * scv);
    const int32_t *b = best_j + (size_t)i * 5;
    cen[t] = rr_km_majority5(sig[(size_t)b[0] * scv + z], sig[(size_t)b[1] * scv + z], sig[(size_t)b[2] * scv + z],
                             sig[(size_t)b[3] * scv + z], sig[(size_t)b[4] * scv + z]);
}

// the first best centroid of another read over the rows of a panel; best [anzahl] and cluster [anzahl] carry the state
// between panels (zero before the first)
__global__ void __launch_bounds__(128)
rr_k_km_assign_seq(const int32_t *__restrict__ P, int anzahl, int j0, int nj, int32_t *__restrict__ best, int32_t *__restrict__ cluster)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= anzahl) return;
    int b = best[i], bj = cluster[i];
    for (int j = 0; j < nj; j++) {
        const int score = P[(size_t)j * anzahl + i];
        if (score > b && i != j0 + j) { b = score; bj = j0 + j; }        // 2717: first best, not itself
    }
    best[i] = b;
    cluster[i] = bj;
}

#ifndef RR_CPU_EMU   // tests/emu compiles the kernels above with a host compiler; the launch syntax below is nvcc only
static cudaError_t km_launch_pair_scores(const uint64_t *sig, const uint64_t *X, const int32_t *xrows, int j0, int nj, int anzahl, int scv,
                                         int32_t *out, int64_t stride_i, int64_t stride_j, cudaStream_t st)
{
    if (anzahl <= 0 || nj <= 0) return cudaSuccess;
    for (int jb = 0; jb < nj; jb += 65535 * KM_TILE) {                   // grid.y limit
        const int n = std::min(nj - jb, 65535 * KM_TILE);
        dim3 grid((unsigned)((anzahl + KM_TILE - 1) / KM_TILE), (unsigned)((n + KM_TILE - 1) / KM_TILE));
        rr_k_km_pair_scores<<<grid, 256, 0, st>>>(sig, X, xrows ? xrows + jb : nullptr, j0 + jb, n, anzahl, scv, out + (int64_t)jb * stride_j,
                                                  stride_i, stride_j);
        rr_count_launch(1);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// the dissolution's table S[anzahl][nJ] (cluster fastest)
cudaError_t rr_launch_kmeans_scores(const uint64_t *sig, const uint64_t *cen, const int32_t *J, int nJ, int anzahl, int scv, int32_t *S,
                                    cudaStream_t st)
{
    return km_launch_pair_scores(sig, cen, J, 0, nJ, anzahl, scv, S, (int64_t)nJ, 1, st);
}

cudaError_t rr_launch_kmeans_signatures(const uint8_t *rows, int cols, int codes, const int32_t *vars, int n_vars, int anzahl,
                                        int scv, uint64_t *sig, cudaStream_t st)
{
    if (anzahl <= 0) return cudaSuccess;
    for (int i0 = 0; i0 < anzahl; i0 += 65535) {                        // grid.y limit
        const int n = std::min(65535, anzahl - i0);
        dim3 grid((unsigned)((2 * scv + 7) / 8), (unsigned)n);
        rr_k_km_signatures<<<grid, 256, 0, st>>>(rows + (size_t)i0 * cols, cols, codes, vars, n_vars, n, scv,
                                                 (uint32_t *)(sig + (size_t)i0 * scv));
        rr_count_launch(1);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// panel: room for panel_rows x anzahl scores; state: 6 x anzahl int32 (five slot scores + the assign sweep's best), zeroed here
cudaError_t rr_launch_kmeans_sweeps(const uint64_t *sig, int anzahl, int scv, int32_t *best_j, uint64_t *cen, int32_t *cluster,
                                    int32_t *panel, int panel_rows, int32_t *state, cudaStream_t st)
{
    if (anzahl <= 0) return cudaSuccess;
    if (panel_rows < 1) return cudaErrorInvalidValue;
    cudaError_t e;
    int32_t *best_s = state, *best = state + (size_t)5 * anzahl;
    if ((e = cudaMemsetAsync(state, 0, sizeof(int32_t) * 6 * (size_t)anzahl, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(best_j, 0, sizeof(int32_t) * 5 * (size_t)anzahl, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(cluster, 0, sizeof(int32_t) * (size_t)anzahl, st)) != cudaSuccess) return e;
    const unsigned nb = (unsigned)((anzahl + 127) / 128);
    for (int j0 = 0; j0 < anzahl; j0 += panel_rows) {                    // first sweep: every read against every read, in order
        const int nj = std::min(panel_rows, anzahl - j0);
        if ((e = km_launch_pair_scores(sig, sig, nullptr, j0, nj, anzahl, scv, panel, 1, (int64_t)anzahl, st)) != cudaSuccess) return e;
        rr_k_km_top5_seq<<<nb, 128, 0, st>>>(panel, anzahl, j0, nj, best_s, best_j);
        rr_count_launch(1);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    const int64_t nw = (int64_t)anzahl * scv;
    rr_k_km_centroids<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>(sig, best_j, anzahl, scv, cen);
    rr_count_launch(1);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    for (int j0 = 0; j0 < anzahl; j0 += panel_rows) {                    // second sweep: every read against every centroid
        const int nj = std::min(panel_rows, anzahl - j0);
        if ((e = km_launch_pair_scores(sig, cen, nullptr, j0, nj, anzahl, scv, panel, 1, (int64_t)anzahl, st)) != cudaSuccess) return e;
        rr_k_km_assign_seq<<<nb, 128, 0, st>>>(panel, anzahl, j0, nj, best, cluster);
        rr_count_launch(1);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    return cudaSuccess;
}
#endif
