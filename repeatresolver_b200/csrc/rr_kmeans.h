/* rr_kmeans.h -- the integer pieces of Kmeans (/root/reference/RepeatResolver.c:2604-2821), written once for host
 * and device.  A read's signature is a bit vector over the selected groups (bit j = the
 * read belongs to group Vars[j], 2634-2642), stored as scv = varzahl/64 + 1 words of 64 bits like the reference's
 * (2626); padding bits are 0 everywhere. */
#ifndef RR_KMEANS_H
#define RR_KMEANS_H
#include <stdint.h>

#if defined(__CUDACC__)
#define RR_KM_HD __host__ __device__ __forceinline__
#else
#define RR_KM_HD static inline
#endif

RR_KM_HD int rr_km_popc64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}

/* GrMatch (163-175): 64 * scv minus the Hamming distance */
RR_KM_HD int rr_km_match(const uint64_t *a, const uint64_t *b, int scv)
{
    int d = 0;
    for (int z = 0; z < scv; z++) d += rr_km_popc64(a[z] ^ b[z]);
    return scv * 64 - d;
}

/* One step of the five-slot rule of 2662-2692 for read j with similarity `score`: the slots are first put in ascending
 * order by the reference's exchange sort (k outer, l inner, exchange when slot l is smaller than slot k: not stable, so
 * which of several equal smallest entries ends up in slot 0 is part of the rule), then read j replaces slot 0 if its
 * score is strictly larger.  Slots start as (score 0, read 0). */
RR_KM_HD void rr_km_top5_step(int bs[5], int bj[5], int score, int j)
{
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 5; k++)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int l = k + 1; l < 5; l++)
            if (bs[l] < bs[k]) {
                int t = bs[l]; bs[l] = bs[k]; bs[k] = t;
                t = bj[l]; bj[l] = bj[k]; bj[k] = t;
            }
    if (score > bs[0]) { bs[0] = score; bj[0] = j; }
}

/* 2697-2705 for 64 groups at once: bit set where more than two of the five words have it */
RR_KM_HD uint64_t rr_km_majority5(uint64_t a, uint64_t b, uint64_t c, uint64_t d, uint64_t e)
{
    /* bit-sliced count of five one-bit inputs: (s2 s1 s0) = a + b + c + d + e */
    const uint64_t ab = a ^ b, ab_c = a & b;            /* a + b       = 2*ab_c + ab */
    const uint64_t cd = c ^ d, cd_c = c & d;            /* c + d       = 2*cd_c + cd */
    const uint64_t s0a = ab ^ cd, c0 = ab & cd;         /* low bits    = 2*c0 + s0a  */
    const uint64_t s0 = s0a ^ e, c1 = s0a & e;          /* + e                        */
    /* twos: ab_c + cd_c + c0 + c1 (c0 and c1 cannot both be set) */
    const uint64_t t = c0 | c1;
    const uint64_t x = ab_c ^ cd_c, y = ab_c & cd_c;    /* ab_c + cd_c = 2*y + x */
    const uint64_t s1 = x ^ t, c2 = x & t;              /* twos bit, carry into fours */
    const uint64_t s2 = y | c2;                         /* fours bit (y and c2 cannot both be set: the sum is at most 5) */
    (void)s0;
    return s2 | (s1 & s0) ;                             /* count >= 3: 4 or 5, or exactly 3 (= 2 + 1) */
}

#endif /* RR_KMEANS_H */
