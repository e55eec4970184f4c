// Helpers shared by the two exhaustive-scan kernels (popcount form: exhaustive.cu, tensor-core form:
// exhaustive_tc.cu).
#pragma once
#include <float.h>

#include "device_math.cuh"
#include "kernels.h"

namespace cpb {

constexpr int kExThreads = 256;
constexpr int kQT = 8;            // queries per CTA tile
constexpr int kCapMax = 2048;     // candidate slots per query in shared memory (power of two; 1024 for small k')
constexpr uint32_t kMaxKPrime = 1024;
constexpr unsigned long long kNoKey = 0xFFFFFFFFFFFFFFFFull;

__device__ __forceinline__ unsigned long long make_key(float est, uint32_t id) {
    return ((unsigned long long)__float_as_uint(est) << 32) | id;
}

// in-place ascending bitonic sort of n (power of two) keys in shared memory by the whole CTA
__device__ __forceinline__ void bitonic_sort(unsigned long long* a, uint32_t n) {
    for (uint32_t k = 2; k <= n; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
                const uint32_t p = i ^ j;
                if (p > i) {
                    const unsigned long long x = a[i], y = a[p];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { a[i] = y; a[p] = x; }
                }
            }
            __syncthreads();
        }
    }
}

// estimate of one (query, vertex) pair: the AVX2 lane of convert_to_distances_with_bounds (:138-173)
__device__ __forceinline__ float flat_estimate(float A, float Bc, float C, float aa, float ab, float floor_, float dqp,
                                               uint32_t sum, float pc, float nop, float ipqo) {
    if (dqp < 1e-12f) return __fmaf_rn(nop, nop, dqp);
    const float ip = __fmaf_rn(A, (float)sum, __fmaf_rn(Bc, pc, C));
    const float q = max_ps(ipqo, floor_);
    const float corr = __fsub_rn(ip, 0.0f);
    float e = q > 1e-10f ? __fdiv_rn(corr, q) : 0.0f;
    e = __fmaf_rn(aa, e, ab);
    const float d = __fmaf_rn(-__fmul_rn(2.0f, nop), e, __fmaf_rn(nop, nop, dqp));
    return max_ps(d, 0.0f);
}


// Division-free screen for the candidate test `est <= tau`: the same expression with ip/q replaced by
// ip * rq (rq = 1/q rounded, once per vertex) differs from the exact estimate by a few ulps of its
// largest term; an estimate that clears tau by 2^-16 of that scale cannot pass the exact test either, so
// almost every (query, vertex) pair is rejected here and the IEEE division only runs for near-candidates.
// Returns true when the exact estimate must be computed.
__device__ __forceinline__ bool flat_screen(float A, float Bc, float C, float aa, float ab, float dqp, uint32_t sum, float pc,
                                            float nop, float rq, float tau) {
    const float ip = __fmaf_rn(A, (float)sum, __fmaf_rn(Bc, pc, C));
    const float e = __fmaf_rn(aa, __fmul_rn(ip, rq), ab);
    const float base = __fmaf_rn(nop, nop, dqp);
    const float t2 = __fmul_rn(__fmul_rn(2.0f, nop), e);
    const float approx = __fsub_rn(base, t2);
    const float scale = __fadd_rn(base, fabsf(t2));
    return !(approx > __fmaf_rn(scale, 1.52587890625e-5f, tau));
}

}  // namespace cpb
