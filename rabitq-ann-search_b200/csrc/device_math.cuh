// Device functions shared by every kernel of the query path: FastScan integer sums, the float
// epilogues, exact distances.  Float code is written op by op (explicit __fmaf_rn / __fdiv_rn /
// __fsqrt_rn, compiled with -fmad=false) so that results are bit-equal to the reference's AVX2
// and GCC-contracted scalar sequences.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "device_index.h"

namespace cpb {

constexpr unsigned kFull = 0xFFFFFFFFu;

__device__ __forceinline__ float max_ps(float a, float b) { return a > b ? a : b; }  // _mm256_max_ps(a,b)
__device__ __forceinline__ float min_ps(float a, float b) { return a < b ? a : b; }  // _mm256_min_ps(a,b)

// Per-query estimator constants: RaBitQQuery<D> minus the LUT (core/codes.hpp:78-93).
struct QParams {
    float A, Bc, C;            // coeff_fastscan, coeff_popcount, coeff_constant
    float a, b, floor_, slack;  // affine_a, affine_b, ip_qo_floor, dot_slack
};

// ---------------------------------------------------------------------------------------------
// FastScan integer sums.  Replaces compute_inner_products (distance/fastscan_kernel.hpp:17-87):
// out[v] = sum_seg lut[seg][nibble_seg(v)] == sum_i bit_i(v) * u_i.  With the query's 4-bit
// values held as four bit-planes U_t this is sum_t 2^t popc(code & U_t): no table, no shuffle,
// exact in integers.  One lane = one neighbour slot; `planes` points at the block, `uq` at the
// query planes in shared memory ([t][chunk] uint4, broadcast reads).
// ---------------------------------------------------------------------------------------------
// sum_t 2^t * popc(w & u_t) over one 128-dim chunk (4 words x 4 query planes = 16 masked words).
// POPC issues at 4 lanes/clk/SMSP (the XU pipe), a quarter of the logic pipe's rate, and 16 of them per
// plane make the XU pipe the kernel's ceiling; so the 16 words are first compressed with carry-save
// adders (3 words of weight W -> 1 of weight W + 1 of weight 2W, two LOP3 each) down to 9 words:
// 9 POPC + 14 LOP3 instead of 16 POPC, the same issue slots, XU and logic pipes balanced.  Exact.
__device__ __forceinline__ void csa(uint32_t& s, uint32_t& c, uint32_t a, uint32_t b, uint32_t d) {
    s = a ^ b ^ d;
    c = (a & b) | (d & (a ^ b));
}

__device__ __forceinline__ uint32_t weighted_popc(const uint4& w, const uint4& u0, const uint4& u1,
                                                  const uint4& u2, const uint4& u3) {
    uint32_t s1, c1, s2, c2, s3, c3, s4, c4, s5, c5, s6, c6, s7, c7;
    csa(s1, c1, w.x & u0.x, w.y & u0.y, w.z & u0.z);                      // weight 1: s1, (w.w & u0.w); carry c1
    csa(s2, c2, w.x & u1.x, w.y & u1.y, w.z & u1.z);                      // weight 2
    csa(s3, c3, s2, w.w & u1.w, c1);                                       //   -> s3; carries c2, c3
    csa(s4, c4, w.x & u2.x, w.y & u2.y, w.z & u2.z);                      // weight 4
    csa(s5, c5, w.w & u2.w, c2, c3);                                       //   -> s4, s5; carries c4, c5
    csa(s6, c6, w.x & u3.x, w.y & u3.y, w.z & u3.z);                      // weight 8
    csa(s7, c7, w.w & u3.w, c4, c5);                                       //   -> s6, s7; carries c6, c7 (weight 16)
    return (__popc(s1) + __popc(w.w & u0.w)) + 2u * __popc(s3) + 4u * (__popc(s4) + __popc(s5)) +
           8u * (__popc(s6) + __popc(s7)) + 16u * (__popc(c6) + __popc(c7));
}

template <int B, bool SMEM = false>
__device__ __forceinline__ void plane_sums(const uint4* __restrict__ planes, uint32_t nch, uint32_t lane,
                                           const uint4* __restrict__ uq, uint32_t (&ps)[B]) {
#pragma unroll
    for (int b = 0; b < B; ++b) ps[b] = 0;
    for (uint32_t c = 0; c < nch; ++c) {
        const uint4 u0 = uq[0 * nch + c], u1 = uq[1 * nch + c], u2 = uq[2 * nch + c], u3 = uq[3 * nch + c];
#pragma unroll
        for (int b = 0; b < B; ++b) {
            const uint4* wp = planes + ((size_t)b * nch + c) * 32 + lane;
            const uint4 w = SMEM ? *wp : __ldg(wp);
            ps[b] += weighted_popc(w, u0, u1, u2, u3);
        }
    }
}

// One plane only (the N-bit path of the search kernel evaluates planes lazily).
template <bool SMEM>
__device__ __forceinline__ uint32_t plane_sum_one(const uint4* __restrict__ planes, uint32_t b, uint32_t nch, uint32_t lane,
                                                  const uint4* __restrict__ uq) {
    uint32_t s = 0;
    for (uint32_t c = 0; c < nch; ++c) {
        const uint4* wp = planes + ((size_t)b * nch + c) * 32 + lane;
        const uint4 w = SMEM ? *wp : __ldg(wp);
        s += weighted_popc(w, uq[0 * nch + c], uq[1 * nch + c], uq[2 * nch + c], uq[3 * nch + c]);
    }
    return s;
}

// compute_nbit_inner_products (:197-217): nbit = sum_b 2^(B-1-b) plane_b, msb = plane_0;
// compute_msb_only_inner_products (:349-368): msb2 = 2 plane_0 + plane_1.
template <int B>
__device__ __forceinline__ void combine_planes(const uint32_t (&ps)[B], uint32_t& nbit, uint32_t& msb,
                                               uint32_t& msb2) {
    nbit = 0;
#pragma unroll
    for (int b = 0; b < B; ++b) nbit += ps[b] << (B - 1 - b);
    msb = ps[0];
    msb2 = (B >= 2) ? 2u * ps[0] + ps[B >= 2 ? 1 : 0] : ps[0];
}

// ---------------------------------------------------------------------------------------------
// Float epilogues.
// ---------------------------------------------------------------------------------------------
// One lane of the AVX2 8-wide loops of convert_to_distances_with_bounds (:138-173) and
// convert_nbit_to_distances_with_bounds (:277-321).  (A_e,B_e,fs_e,pc_e) feed the estimate,
// (A_l,B_l,fs_l,pc_l) the plane-0 lower bound; `same` = 1-bit (both are one chain).
__device__ __forceinline__ void lane_avx(float A_e, float B_e, float fs_e, float pc_e, float A_l, float B_l,
                                         float fs_l, float pc_l, bool same, const QParams& p, float sqrt_dqp,
                                         float dqp, float nop, float ipqo, float ipcp, float& est, float& lower) {
    const float ip_approx = __fmaf_rn(A_e, fs_e, __fmaf_rn(B_e, pc_e, p.C));
    const float q = max_ps(ipqo, p.floor_);
    const float corr = __fsub_rn(ip_approx, ipcp);
    const bool good = q > 1e-10f;
    float e = good ? __fdiv_rn(corr, q) : 0.0f;
    e = __fmaf_rn(p.a, e, p.b);
    float d = __fmaf_rn(nop, nop, dqp);
    d = __fmaf_rn(-__fmul_rn(2.0f, nop), e, d);
    est = max_ps(d, 0.0f);

    float el = e;
    if (!same) {
        const float ip_msb = __fmaf_rn(A_l, fs_l, __fmaf_rn(B_l, pc_l, p.C));
        const float corr_m = __fsub_rn(ip_msb, ipcp);
        el = good ? __fdiv_rn(corr_m, q) : 0.0f;
        el = __fmaf_rn(p.a, el, p.b);
    }
    float cu = __fdiv_rn(__fadd_rn(el, p.slack), max_ps(sqrt_dqp, 1e-10f));
    cu = min_ps(max_ps(cu, -1.0f), 1.0f);
    float lo = __fmaf_rn(nop, nop, dqp);
    lo = __fmaf_rn(-__fmul_rn(__fmul_rn(2.0f, nop), sqrt_dqp), cu, lo);
    lo = max_ps(lo, 0.0f);
    lower = good ? lo : 0.0f;
}

// Scalar tails (:176-193, :324-345) and convert_msb_to_lower_bounds (:403-424) as GCC 13.3 -O3
// -mfma contracts them: t = A*fs; t = fma(pc,B,t); t += C; t -= ip_cp; t /= q;
// t = fma(t,a,b).
__device__ __forceinline__ float scalar_ip_est(float A, float Bc, float C, float fs, float pc, float ipcp, float q,
                                               float a, float b) {
    float t = __fmul_rn(A, fs);
    t = __fmaf_rn(pc, Bc, t);
    t = __fadd_rn(t, C);
    t = __fsub_rn(t, ipcp);
    t = __fdiv_rn(t, q);
    return __fmaf_rn(t, a, b);
}

// ...except the plane-0 chain inside the N-bit scalar tail (:339-342), which the same compiler
// contracts the other way round: t = B*pc; t = fma(A,fs,t); t += C  (pinned by oracle/probe_contraction.c)
__device__ __forceinline__ float scalar_ip_est_msbtail(float A, float Bc, float C, float fs, float pc, float ipcp,
                                                       float q, float a, float b) {
    float t = __fmul_rn(Bc, pc);
    t = __fmaf_rn(A, fs, t);
    t = __fadd_rn(t, C);
    t = __fsub_rn(t, ipcp);
    t = __fdiv_rn(t, q);
    return __fmaf_rn(t, a, b);
}

__device__ __forceinline__ float scalar_lower(float e, float slack, float sqrt_dqp, float nop, float dqp) {
    float cu = __fdiv_rn(__fadd_rn(e, slack), sqrt_dqp);
    if (cu < -1.0f) cu = -1.0f;
    if (cu > 1.0f) cu = 1.0f;
    const float lo = __fmaf_rn(-__fmul_rn(__fadd_rn(nop, nop), sqrt_dqp), cu, __fmaf_rn(nop, nop, dqp));
    return lo < 0.0f ? 0.0f : lo;
}

// Scalar-tail estimate + lower bound for one lane.
__device__ __forceinline__ void lane_scalar(float A_e, float B_e, float fs_e, float pc_e, float A_l, float B_l,
                                            float fs_l, float pc_l, bool same, const QParams& p, float sqrt_dqp,
                                            float dqp, float nop, float ipqo, float ipcp, float& est, float& lower) {
    const float q = ipqo > p.floor_ ? ipqo : p.floor_;  // std::max(ip_qo, floor)
    const bool good = q > 1e-10f;
    float e;
    if (good) e = scalar_ip_est(A_e, B_e, p.C, fs_e, pc_e, ipcp, q, p.a, p.b);
    else e = __fmaf_rn(0.0f, p.a, p.b);
    const float d = __fmaf_rn(-__fadd_rn(nop, nop), e, __fmaf_rn(nop, nop, dqp));
    est = d < 0.0f ? 0.0f : d;
    if (!good) { lower = 0.0f; return; }
    const float el = same ? e : scalar_ip_est_msbtail(A_l, B_l, p.C, fs_l, pc_l, ipcp, q, p.a, p.b);
    lower = scalar_lower(el, p.slack, sqrt_dqp, nop, dqp);
}

// convert_to_distances_with_bounds<D> (:89-194) for lane `lane` of a block with `count` neighbours.
// `sq` = __fsqrt_rn(dqp), computed once per block by the caller.
__device__ __forceinline__ void convert_1bit(const QParams& p, uint32_t sum, float nop, float ipqo, float ipcp,
                                             uint32_t pop, uint32_t lane, uint32_t count, float dqp, float sq,
                                             float& est, float& lower) {
    if (dqp < 1e-12f) { est = __fmaf_rn(nop, nop, dqp); lower = 0.0f; return; }
    if (lane < (count & ~7u))
        lane_avx(p.A, p.Bc, (float)sum, (float)pop, 0, 0, 0, 0, true, p, sq, dqp, nop, ipqo, ipcp, est, lower);
    else
        lane_scalar(p.A, p.Bc, (float)sum, (float)pop, 0, 0, 0, 0, true, p, sq, dqp, nop, ipqo, ipcp, est, lower);
}

// convert_msb_to_lower_bounds<D,B> (:371-425): K_PARTIAL = 3 with the plane-0 popcount.
template <int B>
__device__ __forceinline__ float convert_msb(const QParams& p, uint32_t msb2, float nop, float ipqo, float ipcp,
                                             uint32_t pop, float dqp, float sq) {
    if (dqp < 1e-12f) return 0.0f;
    constexpr float inv_kp = 1.0f / ((B < 2) ? 1.0f : 3.0f);   // constexpr float inv_K = 1.0f / K_PARTIAL (:378-380)
    const float A = __fmul_rn(p.A, inv_kp), Bc = __fmul_rn(p.Bc, inv_kp);
    const float q = ipqo > p.floor_ ? ipqo : p.floor_;
    if (!(q > 1e-10f)) return 0.0f;
    const float e = scalar_ip_est(A, Bc, p.C, (float)msb2, (float)pop, ipcp, q, p.a, p.b);
    return scalar_lower(e, p.slack, sq, nop, dqp);
}

// convert_nbit_to_distances_with_bounds<D,B> (:220-346).
template <int B>
__device__ __forceinline__ void convert_nbit(const QParams& p, uint32_t nbit, uint32_t msb, float nop, float ipqo,
                                             float ipcp, uint32_t pop, uint32_t wpop, uint32_t lane, uint32_t count,
                                             float dqp, float sq, float& est, float& lower) {
    if (dqp < 1e-12f) { est = __fmaf_rn(nop, nop, dqp); lower = 0.0f; return; }
    constexpr float inv_K = 1.0f / (float)((1u << B) - 1u);   // constexpr float inv_K = 1.0f / K (:232-233)
    const float A_n = __fmul_rn(p.A, inv_K), B_n = __fmul_rn(p.Bc, inv_K);
    if (lane < (count & ~7u))
        lane_avx(A_n, B_n, (float)nbit, (float)wpop, p.A, p.Bc, (float)msb, (float)pop, false, p, sq, dqp, nop, ipqo,
                 ipcp, est, lower);
    else
        lane_scalar(A_n, B_n, (float)nbit, (float)wpop, p.A, p.Bc, (float)msb, (float)pop, false, p, sq, dqp, nop,
                    ipqo, ipcp, est, lower);
}

// The two halves of convert_nbit_to_distances_with_bounds, separately callable (bit-identical to
// convert_nbit): the lower bound needs only plane 0 (`msb`), the estimate all planes (`nbit`).
template <int B>
__device__ __forceinline__ float nbit_lower(const QParams& p, uint32_t msb, float nop, float ipqo, float ipcp,
                                            uint32_t pop, uint32_t lane, uint32_t count, float dqp, float sq) {
    if (dqp < 1e-12f) return 0.0f;
    if (lane < (count & ~7u)) {   // AVX2 lanes (:277-321)
        const float q = max_ps(ipqo, p.floor_);
        if (!(q > 1e-10f)) return 0.0f;
        const float ip_msb = __fmaf_rn(p.A, (float)msb, __fmaf_rn(p.Bc, (float)pop, p.C));
        float el = __fdiv_rn(__fsub_rn(ip_msb, ipcp), q);
        el = __fmaf_rn(p.a, el, p.b);
        float cu = __fdiv_rn(__fadd_rn(el, p.slack), max_ps(sq, 1e-10f));
        cu = min_ps(max_ps(cu, -1.0f), 1.0f);
        const float lo = __fmaf_rn(-__fmul_rn(__fmul_rn(2.0f, nop), sq), cu, __fmaf_rn(nop, nop, dqp));
        return max_ps(lo, 0.0f);
    }
    const float q = ipqo > p.floor_ ? ipqo : p.floor_;   // scalar tail (:324-345)
    if (!(q > 1e-10f)) return 0.0f;
    const float el = scalar_ip_est_msbtail(p.A, p.Bc, p.C, (float)msb, (float)pop, ipcp, q, p.a, p.b);
    return scalar_lower(el, p.slack, sq, nop, dqp);
}

template <int B>
__device__ __forceinline__ float nbit_est(const QParams& p, uint32_t nbit, float nop, float ipqo, float ipcp,
                                          uint32_t wpop, uint32_t lane, uint32_t count, float dqp) {
    if (dqp < 1e-12f) return __fmaf_rn(nop, nop, dqp);
    constexpr float inv_K = 1.0f / (float)((1u << B) - 1u);
    const float A_n = __fmul_rn(p.A, inv_K), B_n = __fmul_rn(p.Bc, inv_K);
    if (lane < (count & ~7u)) {
        const float ip_approx = __fmaf_rn(A_n, (float)nbit, __fmaf_rn(B_n, (float)wpop, p.C));
        const float q = max_ps(ipqo, p.floor_);
        float e = q > 1e-10f ? __fdiv_rn(__fsub_rn(ip_approx, ipcp), q) : 0.0f;
        e = __fmaf_rn(p.a, e, p.b);
        const float d = __fmaf_rn(-__fmul_rn(2.0f, nop), e, __fmaf_rn(nop, nop, dqp));
        return max_ps(d, 0.0f);
    }
    const float q = ipqo > p.floor_ ? ipqo : p.floor_;
    float e;
    if (q > 1e-10f) e = scalar_ip_est(A_n, B_n, p.C, (float)nbit, (float)wpop, ipcp, q, p.a, p.b);
    else e = __fmaf_rn(0.0f, p.a, p.b);
    const float d = __fmaf_rn(-__fadd_rn(nop, nop), e, __fmaf_rn(nop, nop, dqp));
    return d < 0.0f ? 0.0f : d;
}

// ---------------------------------------------------------------------------------------------
// Exact distances.  dot_product_simd / l2_distance_simd (core/memory.hpp:65-96) keep eight FMA
// accumulators (element i -> accumulator i%8, increasing i) and reduce (lo+hi) -> hadd -> hadd.
// Eight lanes of a warp play the eight accumulators; vectors are stored accumulator-major
// (xT[l*T + t] = x[8t + l]) so each lane streams its own chain with 128-bit loads.  A warp runs
// four such groups (g = lane>>3) on four vectors at once.  qrow = this lane's row of the query
// in shared memory (row stride T+4 floats: conflict-free for 128-bit reads).
// ---------------------------------------------------------------------------------------------
// Bank swizzle of the accumulator-major layout.  Row l (T floats) is read by lane l in 16-byte chunks, all eight rows at
// the same chunk index at once: with a row stride of T floats those eight reads fall on the same shared-memory banks
// (4-way conflicts at T = 16: 48 of the 157 shared-memory wavefronts of an expansion).  So chunk j of row l is STORED at
// chunk position j ^ raw_swizzle(l, T / 4): the eight rows then cover eight different 16-byte bank groups, with no padding
// (the order in which a chain consumes its elements is unchanged).
__host__ __device__ __forceinline__ uint32_t raw_swizzle(uint32_t l, uint32_t nv) { return nv >= 8 ? l : (l * nv) >> 3; }
// position (in floats) of element t of row l inside the row
__host__ __device__ __forceinline__ uint32_t raw_chunk_pos(uint32_t l, uint32_t t, uint32_t T) {
    return (T & 3u) ? t : ((((t >> 2) ^ raw_swizzle(l, T >> 2)) << 2) | (t & 3u));
}

__device__ __forceinline__ float group_reduce8(float acc) {
    acc = __fadd_rn(acc, __shfl_xor_sync(kFull, acc, 4));  // s[l] = acc[l] + acc[l+4]
    acc = __fadd_rn(acc, __shfl_xor_sync(kFull, acc, 1));  // s0+s1 | s2+s3
    acc = __fadd_rn(acc, __shfl_xor_sync(kFull, acc, 2));  // (s0+s1)+(s2+s3)
    return acc;
}

// XSMEM: xrow points into shared memory (a staged vector) instead of HBM.  xrow = row (lane & 7) of a vector in the
// swizzled accumulator-major layout (raw_chunk_pos); qrow = the same row of the query, unswizzled, row stride T + 4.
template <bool L2, bool XSMEM = false>
__device__ __forceinline__ float group_chain(const float* __restrict__ xrow, const float* __restrict__ qrow,
                                             uint32_t T, bool active) {
    float acc = 0.0f;
    if (active) {
        if ((T & 3u) == 0) {
            const float4* xp = reinterpret_cast<const float4*>(xrow);
            const float4* qp = reinterpret_cast<const float4*>(qrow);
            const uint32_t nv = T >> 2;
            const uint32_t sw = raw_swizzle(threadIdx.x & 7u, nv);
#pragma unroll 4
            for (uint32_t j = 0; j < nv; ++j) {
                const float4 x = XSMEM ? xp[j ^ sw] : __ldg(xp + (j ^ sw));
                const float4 q = qp[j];
                if (L2) {
                    float d;
                    d = __fsub_rn(q.x, x.x); acc = __fmaf_rn(d, d, acc);
                    d = __fsub_rn(q.y, x.y); acc = __fmaf_rn(d, d, acc);
                    d = __fsub_rn(q.z, x.z); acc = __fmaf_rn(d, d, acc);
                    d = __fsub_rn(q.w, x.w); acc = __fmaf_rn(d, d, acc);
                } else {
                    acc = __fmaf_rn(q.x, x.x, acc);
                    acc = __fmaf_rn(q.y, x.y, acc);
                    acc = __fmaf_rn(q.z, x.z, acc);
                    acc = __fmaf_rn(q.w, x.w, acc);
                }
            }
        } else {
            for (uint32_t t = 0; t < T; ++t) {
                const float x = XSMEM ? xrow[t] : __ldg(xrow + t), q = qrow[t];
                if (L2) { const float d = __fsub_rn(q, x); acc = __fmaf_rn(d, d, acc); }
                else acc = __fmaf_rn(q, x, acc);
            }
        }
    }
    return group_reduce8(acc);
}

// exact_l2 lambda, search/rabitq_search.hpp:88-93: max(|q|^2 + norm_sq[id] - 2<q,x>, 0)
__device__ __forceinline__ float exact_from_dot(float qn, float norm, float dot) {
    const float r = __fsub_rn(__fadd_rn(qn, norm), __fmul_rn(2.0f, dot));
    return r < 0.0f ? 0.0f : r;
}

}  // namespace cpb
