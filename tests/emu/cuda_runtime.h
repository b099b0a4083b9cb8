/* TEST INFRASTRUCTURE ONLY (tests/emu).  A stand-in for <cuda_runtime.h> that lets a HOST compiler build the kernel
 * bodies of csrc/*.cu (compiled with -DRR_CPU_EMU, this directory first on the include path): every CUDA thread of a
 * block runs as a pthread, __syncthreads() is a barrier over the block, the warp primitives exchange through a per-warp
 * buffer behind a 32-thread barrier.  It checks the kernels' logic (indexing, staging, reductions, tails) on the CPU; it
 * says nothing about performance and is no substitute for a run on the GPU.  Never linked into the product. */
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <pthread.h>

typedef int cudaError_t;
typedef void *cudaStream_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1 };
struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct int4 { int x, y, z, w; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 v = {x, y, z, w}; return v; }
static inline int4 make_int4(int x, int y, int z, int w) { int4 v = {x, y, z, w}; return v; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static            /* one block runs at a time */
#define __launch_bounds__(...)
#define __noinline__
#define __align__(n) alignas(n)

struct emu_warp { pthread_barrier_t bar; unsigned long long buf[32]; };
struct emu_block { pthread_barrier_t bar; emu_warp *warps; };
extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;
extern thread_local emu_block *emu_blk;

static inline void __syncthreads() { pthread_barrier_wait(&emu_blk->bar); }
static inline emu_warp *emu_my_warp() { return emu_blk->warps + (threadIdx.x >> 5); }

/* all 32 lanes must call these (the kernels use the full mask in warp-uniform control flow only) */
static inline unsigned __ballot_sync(unsigned, int pred)
{
    emu_warp *w = emu_my_warp();
    w->buf[threadIdx.x & 31] = pred ? 1u : 0u;
    pthread_barrier_wait(&w->bar);
    unsigned m = 0;
    for (int l = 0; l < 32; l++) m |= (unsigned)w->buf[l] << l;
    pthread_barrier_wait(&w->bar);
    return m;
}
static inline unsigned __reduce_add_sync(unsigned, unsigned v)
{
    emu_warp *w = emu_my_warp();
    w->buf[threadIdx.x & 31] = v;
    pthread_barrier_wait(&w->bar);
    unsigned s = 0;
    for (int l = 0; l < 32; l++) s += (unsigned)w->buf[l];
    pthread_barrier_wait(&w->bar);
    return s;
}
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int lane_mask)
{
    static_assert(sizeof(T) <= 8, "one 64-bit slot per lane");
    emu_warp *w = emu_my_warp();
    unsigned long long raw = 0;
    memcpy(&raw, &v, sizeof(T));
    w->buf[threadIdx.x & 31] = raw;
    pthread_barrier_wait(&w->bar);
    raw = w->buf[(threadIdx.x & 31) ^ (unsigned)lane_mask];
    pthread_barrier_wait(&w->bar);
    T out;
    memcpy(&out, &raw, sizeof(T));
    return out;
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src_lane)
{
    static_assert(sizeof(T) <= 8, "one 64-bit slot per lane");
    emu_warp *w = emu_my_warp();
    unsigned long long raw = 0;
    memcpy(&raw, &v, sizeof(T));
    w->buf[threadIdx.x & 31] = raw;
    pthread_barrier_wait(&w->bar);
    raw = w->buf[(unsigned)src_lane & 31u];
    pthread_barrier_wait(&w->bar);
    T out;
    memcpy(&out, &raw, sizeof(T));
    return out;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0u; }
static inline void __syncwarp() { pthread_barrier_wait(&emu_my_warp()->bar); }
static inline unsigned __activemask() { return 0xffffffffu; }     /* only meaningful where the whole warp is active */
static inline long long __double_as_longlong(double d) { long long x; memcpy(&x, &d, 8); return x; }
static inline double __longlong_as_double(long long x) { double d; memcpy(&d, &x, 8); return d; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline int __float2int_rd(float f) { return (int)floorf(f); }
static inline unsigned __float_as_uint(float f) { unsigned x; memcpy(&x, &f, 4); return x; }
static inline float emu_log2f(float x) { return log2f(x); }
#define __log2f emu_log2f             /* glibc declares a __log2f of its own */
static inline float __double2float_rd(double d) { float f = (float)d; return (double)f > d ? nextafterf(f, -INFINITY) : f; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
using std::max;
using std::min;
static inline int min(int a, unsigned b) { return a < (int)b ? a : (int)b; }
