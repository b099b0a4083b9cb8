// rr_kmeans.cu -- the read x read sweeps of Kmeans (/root/reference/RepeatResolver.c:2604-2821) on the device
// ("the transposed Gram problem"); parity on a B200 in tests/test_zz_gpu_kmeans.py, the same
// source under the CPU emulation of tests/emu.  The integer rules are shared with the host through rr_kmeans.h and pinned
// there against the unmodified reference (tests/test_oracle_kmeans.py).
//
// sig[anzahl][scv]: the signatures of the part's reads over the selected groups (64-bit words, padding 0).
//   rr_k_km_signatures one warp per (read, 32 groups): the signatures from the part's rows (2626-2650)
//   rr_k_km_top5       one thread per read i, all reads j in order (tiles of signatures through shared memory, every
//                      thread of the block reads the same word: a broadcast): GrMatch and the reference's five-slot rule
//                      (2662-2692), which depends on the order of the reads and so stays sequential per read
//   rr_k_km_centroids  one thread per (read, word): bitwise majority of the five kept reads' signatures (2697-2705)
//   rr_k_km_assign     one thread per read i: the first best centroid of another read (2709-2725)
//   rr_k_km_scores     one warp per (read, candidate cluster): the scores the dissolution of small clusters looks up (2735-2745)
// XOR+POPC work on a few KB of signatures per tile: POPC-issue bound, anzahl^2 * scv word operations per sweep.
#include <algorithm>
#include "rr_kernels.h"
#include "rr_kmeans.h"

constexpr int KM_THREADS = 128;

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

// stage signatures [j0, j0 + nj) into shared memory, coalesced
__device__ __forceinline__ void km_stage(uint64_t *tile, const uint64_t *__restrict__ src, int j0, int nj, int scv)
{
    for (int idx = threadIdx.x; idx < nj * scv; idx += KM_THREADS) tile[idx] = src[(size_t)j0 * scv + idx];
}

__global__ void __launch_bounds__(KM_THREADS)
rr_k_km_top5(const uint64_t *__restrict__ sig, int anzahl, int scv, int tile_reads, int32_t *__restrict__ best_j /*[anzahl][5]*/)
{
    RR_DYN_SMEM(uint64_t, km_smem);
    uint64_t *tile = km_smem;                                   // [tile_reads][scv]
    const int i = blockIdx.x * KM_THREADS + threadIdx.x;
    const uint64_t *mine = sig + (size_t)min(i, anzahl - 1) * scv;   // idle threads still take part in the staging
    int bs[5] = {0, 0, 0, 0, 0}, bj[5] = {0, 0, 0, 0, 0};
    for (int j0 = 0; j0 < anzahl; j0 += tile_reads) {
        const int nj = min(tile_reads, anzahl - j0);
        __syncthreads();
        km_stage(tile, sig, j0, nj, scv);
        __syncthreads();
        for (int j = 0; j < nj; j++) {
            int d = 0;
            for (int z = 0; z < scv; z++) d += __popcll(tile[j * scv + z] ^ __ldg(mine + z));
            rr_km_top5_step(bs, bj, scv * 64 - d, j0 + j);
        }
    }
    if (i < anzahl)
        for (int k = 0; k < 5; k++) best_j[(size_t)i * 5 + k] = bj[k];
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

__global__ void __launch_bounds__(KM_THREADS)
rr_k_km_assign(const uint64_t *__restrict__ sig, const uint64_t *__restrict__ cen, int anzahl, int scv, int tile_reads,
               int32_t *__restrict__ cluster)
{
    RR_DYN_SMEM(uint64_t, km_smem);
    uint64_t *tile = km_smem;
    const int i = blockIdx.x * KM_THREADS + threadIdx.x;
    const uint64_t *mine = sig + (size_t)min(i, anzahl - 1) * scv;
    int best = 0, best_j = 0;
    for (int j0 = 0; j0 < anzahl; j0 += tile_reads) {
        const int nj = min(tile_reads, anzahl - j0);
        __syncthreads();
        km_stage(tile, cen, j0, nj, scv);
        __syncthreads();
        for (int j = 0; j < nj; j++) {
            int d = 0;
            for (int z = 0; z < scv; z++) d += __popcll(tile[j * scv + z] ^ __ldg(mine + z));
            const int score = scv * 64 - d;
            if (score > best && i != j0 + j) { best = score; best_j = j0 + j; }     // 2717: first best, not itself
        }
    }
    if (i < anzahl) cluster[i] = best_j;
}

// The scores the dissolution of small clusters asks for (2735-2745): S[i * nJ + k] = GrMatch(Centroids[J[k]], VarSigs[i]) for
// the clusters J that can ever be admissible there; one warp per (read, cluster), lanes stride over the signature words.
__global__ void __launch_bounds__(256)
rr_k_km_scores(const uint64_t *__restrict__ sig, const uint64_t *__restrict__ cen, const int32_t *__restrict__ J, int nJ, int anzahl,
               int scv, int32_t *__restrict__ S)
{
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= (int64_t)anzahl * nJ) return;                               // warp-uniform
    const int i = (int)(w / nJ), k = (int)(w - (int64_t)i * nJ);
    const uint64_t *a = cen + (size_t)J[k] * scv, *b = sig + (size_t)i * scv;
    unsigned d = 0;
    for (int z = lane; z < scv; z += 32) d += (unsigned)rr_km_popc64(a[z] ^ b[z]);
    d = __reduce_add_sync(0xffffffffu, d);
    if (lane == 0) S[w] = scv * 64 - (int)d;
}

#ifndef RR_CPU_EMU   // tests/emu compiles the kernels above with a host compiler; the launch syntax below is nvcc only
cudaError_t rr_launch_kmeans_scores(const uint64_t *sig, const uint64_t *cen, const int32_t *J, int nJ, int anzahl, int scv, int32_t *S,
                                    cudaStream_t st)
{
    const int64_t warps = (int64_t)anzahl * nJ;
    if (warps <= 0) return cudaSuccess;
    rr_k_km_scores<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(sig, cen, J, nJ, anzahl, scv, S);
    rr_count_launch(1);
    return cudaGetLastError();
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

cudaError_t rr_launch_kmeans_sweeps(const uint64_t *sig, int anzahl, int scv, int32_t *best_j, uint64_t *cen, int32_t *cluster,
                                    cudaStream_t st)
{
    if (anzahl <= 0) return cudaSuccess;
    // signatures per shared-memory tile: up to 32 KB
    const int tile_reads = (int)std::max<size_t>(1, std::min<size_t>(256, ((size_t)32 << 10) / ((size_t)scv * sizeof(uint64_t))));
    const size_t smem = (size_t)tile_reads * scv * sizeof(uint64_t);
    if (smem > ((size_t)200 << 10)) return cudaErrorInvalidValue;   // more than 1.6 M selected groups: not this kernel
    cudaError_t e;
    if (smem > ((size_t)48 << 10)) {
        if ((e = cudaFuncSetAttribute(rr_k_km_top5, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(rr_k_km_assign, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    }
    const unsigned nb = (unsigned)((anzahl + KM_THREADS - 1) / KM_THREADS);
    rr_k_km_top5<<<nb, KM_THREADS, smem, st>>>(sig, anzahl, scv, tile_reads, best_j);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    const int64_t nw = (int64_t)anzahl * scv;
    rr_k_km_centroids<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>(sig, best_j, anzahl, scv, cen);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    rr_k_km_assign<<<nb, KM_THREADS, smem, st>>>(sig, cen, anzahl, scv, tile_reads, cluster);
    rr_count_launch(3);
    return cudaGetLastError();
}
#endif
