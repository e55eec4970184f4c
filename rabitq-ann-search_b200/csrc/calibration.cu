// N4 -- the sample loop of Index::calibrate_estimator (api/hnsw_index.hpp:786-866, lambda process_query) on
// the device: for every sampled query q and its start vertex, the closer of the start vertex and its layer-0 neighbours
// (l2_distance_simd, strict <, stored order, stop at the first empty slot) becomes the parent p; then for each neighbour o
// of p the numbers the estimator is calibrated on:
//   nop            = max(nop[o], 1e-12)
//   ip_corrected   = ip_approx - ip_cp[o],   ip_approx = A' fs + Bc' pc + C   (A' = A / K, Bc' = Bc / K, pc = weighted popcount
//                    for N-bit codes; fs = compute_inner_products / compute_nbit_inner_products of p's block: the K2 sums)
//   ip_qo_denom    = max(|ip_qo[o]|, 1e-10)
//   true_ip        = <q - p, o - p> / nop,   accumulated dimension by dimension in f32
// What the host does with them afterwards (MAD floor, Huber IRLS affine fit, EVT tail) stays the reference's.
//
// Reuses K1 (the prepared query: bit-planes, coefficients, accumulator-major copy) and the K2 popcount sums.  One warp per
// sample: the eight-accumulator distance chains run four vectors at a time exactly as in the descent of K3; lane = slot of
// the parent's block for everything per neighbour.  Float sequences: the distances are the reference's AVX2 chains
// (device_math.cuh); ip_approx is fma(A', fs, Bc' * pc) + C and true_ip a separate multiply and add per dimension -- what
// GCC 13.3 -O3 -mfma makes of the two scalar expressions (found by enumeration against the compiled reference:
// tests/test_oracle_calibration.py).
#include <float.h>

#include "device_math.cuh"
#include "kernels.h"

namespace cpb {

constexpr int kCalWarps = 4;

template <int B>
__global__ void __launch_bounds__(kCalWarps * 32) calibration_samples_kernel(const DevIndex ix, const CalibrationArgs a) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t s = (uint64_t)blockIdx.x * kCalWarps + warp;
    const uint32_t D = ix.D, T = ix.T, Tp = T + 4, nch = ix.nch;
    const size_t per_warp = ((size_t)8 * Tp * 4 + (size_t)2 * D * 4 + (size_t)nch * 64 + 15) & ~(size_t)15;
    uint8_t* sm = smem_raw + (size_t)warp * per_warp;
    float* qs = reinterpret_cast<float*>(sm);                               // query, accumulator-major, padded rows
    float* qn = qs + (size_t)8 * Tp;                                        // query, natural order
    float* pv = qn + D;                                                     // parent vector, natural order
    uint4* uq = reinterpret_cast<uint4*>(pv + D);                           // query bit-planes
    if (s >= a.ns) return;

    for (uint32_t i = lane; i < D; i += 32) {
        qs[(i / T) * Tp + (i % T)] = a.qT[s * D + i];
        qn[i] = i < ix.dim ? a.queries[s * ix.dim + i] : 0.0f;
    }
    for (uint32_t i = lane; i < 4 * nch; i += 32) uq[i] = reinterpret_cast<const uint4*>(a.uplanes + s * 16 * nch)[i];
    __syncwarp();
    const float* cf = a.coeffs + s * kCoeffStride;
    const float A = cf[0], Bc = cf[1], C = cf[2];
    const uint32_t g = lane >> 3, l = lane & 7u;
    const float* qrow = qs + (size_t)l * Tp;

    // ---- parent: the start vertex or the closest of its neighbours (:790-803) --------------------------------------
    uint32_t parent = a.start_ids[s];
    float best = group_chain<true>(ix.rawT + (size_t)parent * D + (size_t)l * T, qrow, T, true);
    {
        const uint8_t* aux = ix.blocks + (size_t)parent * ix.block_stride + ix.aux_off;
        const uint32_t count = *reinterpret_cast<const uint32_t*>(aux + 640);
        const uint32_t myid = lane < count ? reinterpret_cast<const uint32_t*>(aux)[lane] : kInvalid;
        const unsigned holes = __ballot_sync(kFull, lane < count && myid == kInvalid);
        const uint32_t limit = holes ? (uint32_t)(__ffs(holes) - 1) : count;          // `if (nid == INVALID_NODE) break;`
        for (uint32_t j = 0; j < limit; j += 4) {
            const bool act = j + g < limit;
            const uint32_t nb = __shfl_sync(kFull, myid, act ? j + g : 0);
            const float d = group_chain<true>(ix.rawT + (size_t)nb * D + (size_t)l * T, qrow, T, act);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float dt = __shfl_sync(kFull, d, t * 8);
                const uint32_t nt = __shfl_sync(kFull, nb, t * 8);
                if (j + t < limit && dt < best) { best = dt; parent = nt; }
            }
        }
    }
    // dist_qp_sq is l2_distance_simd(q, parent) again (:808): the same chain on the same inputs
    if (lane == 0) { a.parent[s] = parent; a.nn_dist_sq[s] = best; a.dist_qp_sq[s] = best; }

    // ---- the parent's block (:810-823) -------------------------------------------------------------------------------
    const uint8_t* blk = ix.blocks + (size_t)parent * ix.block_stride;
    const uint8_t* aux = blk + ix.aux_off;
    const uint32_t count = *reinterpret_cast<const uint32_t*>(aux + 640);
    const uint32_t oid = lane < count ? reinterpret_cast<const uint32_t*>(aux)[lane] : kInvalid;
    const unsigned holes = __ballot_sync(kFull, lane < count && oid == kInvalid);
    const uint32_t limit = holes ? (uint32_t)(__ffs(holes) - 1) : count;
    uint32_t ps[B];
    plane_sums<B, false>(reinterpret_cast<const uint4*>(blk), nch, lane, uq, ps);
    uint32_t nbit, msb, msb2;
    combine_planes<B>(ps, nbit, msb, msb2);
    for (uint32_t i = lane; i < D; i += 32) {   // the parent's vector in natural order: element d = 8t + l sits at row l, raw_chunk_pos(l, t)
        const uint32_t ll = i & 7u, t = i >> 3;
        pv[i] = __ldg(ix.rawT + (size_t)parent * D + (size_t)ll * T + raw_chunk_pos(ll, t, T));
    }
    __syncwarp();

    float nop = 0.0f, ipc = 0.0f, den = 0.0f, tip = 0.0f;
    const bool have = lane < limit;
    if (have) {
        const float nop_raw = reinterpret_cast<const float*>(aux + 128)[lane];
        const float ipqo = reinterpret_cast<const float*>(aux + 256)[lane];
        const float ipcp = reinterpret_cast<const float*>(aux + 384)[lane];
        const uint32_t pops = reinterpret_cast<const uint32_t*>(aux + 512)[lane];
        nop = nop_raw > 1e-12f ? nop_raw : 1e-12f;                                     // std::max(nop, kSmall)
        constexpr float inv_K = 1.0f / (float)((1u << B) - 1u);
        const float Ae = B == 1 ? A : __fmul_rn(A, inv_K), Be = B == 1 ? Bc : __fmul_rn(Bc, inv_K);
        const float pc = B == 1 ? (float)(pops & 0xFFFFu) : (float)(pops >> 16);
        const float ip_approx = __fadd_rn(__fmaf_rn(Ae, (float)nbit, __fmul_rn(Be, pc)), C);
        ipc = __fsub_rn(ip_approx, ipcp);
        const float ab = fabsf(ipqo);
        den = ab > 1e-10f ? ab : 1e-10f;                                                 // std::max(|ip_qo|, kMedium)
        const float* ov = ix.rawT + (size_t)oid * D;
        float acc = 0.0f;
        for (uint32_t t = 0; t < T; ++t) {
#pragma unroll
            for (uint32_t ll = 0; ll < 8; ++ll) {
                const uint32_t d = 8 * t + ll;
                const float o = __ldg(ov + (size_t)ll * T + raw_chunk_pos(ll, t, T));
                acc = __fadd_rn(acc, __fmul_rn(__fsub_rn(qn[d], pv[d]), __fsub_rn(o, pv[d])));
            }
        }
        tip = __fdiv_rn(acc, nop);
    }
    const size_t o = s * 32 + lane;
    a.nop[o] = nop; a.ip_corrected[o] = ipc; a.ip_qo_denom[o] = den; a.true_ip[o] = tip; a.neighbor[o] = have ? oid : kInvalid;
}

#ifndef CPB_HOST_EMULATION
cudaError_t launch_calibration_samples(const DevIndex& ix, const CalibrationArgs& a, cudaStream_t stream) {
    if (a.ns == 0) return cudaSuccess;
    const size_t per_warp = ((size_t)8 * (ix.T + 4) * 4 + (size_t)2 * ix.D * 4 + (size_t)ix.nch * 64 + 15) & ~(size_t)15;
    const size_t smem = per_warp * kCalWarps;
    void (*kern)(const DevIndex, const CalibrationArgs) =
        ix.B == 1 ? calibration_samples_kernel<1> : ix.B == 2 ? calibration_samples_kernel<2> : calibration_samples_kernel<4>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)((a.ns + kCalWarps - 1) / kCalWarps), kCalWarps * 32, smem, stream>>>(ix, a);
    return cudaGetLastError();
}
#endif

}  // namespace cpb
