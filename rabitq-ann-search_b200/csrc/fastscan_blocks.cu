// K2 -- FastScan estimator over a list (or a contiguous range) of neighbour blocks.
//
// Replaces fastscan::compute_inner_products / compute_nbit_inner_products /
// compute_msb_only_inner_products (distance/fastscan_kernel.hpp:17-87,197-217,349-368) and the
// epilogues convert_to_distances_with_bounds / convert_msb_to_lower_bounds /
// convert_nbit_to_distances_with_bounds (:89-194,371-425,220-346).  This is the stand-alone form of
// the estimator the search kernel runs per expansion: the parity hook for integer sums and float
// estimates, and the kernel behind the "FastScan HBM GB/s" figure (a pure stream of blocks).
//
// One warp per block, lane = neighbour slot; persistent grid-stride over blocks; the query's
// bit-planes sit in shared memory per warp and are re-staged only when the block's query changes.
#include <float.h>

#include "device_math.cuh"
#include "kernels.h"

namespace cpb {

constexpr int kFsWarps = 8;

template <int B>
__global__ void __launch_bounds__(kFsWarps * 32) fastscan_blocks_kernel(const DevIndex ix, const FastScanArgs a) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t nch = ix.nch;
    uint4* uqs = reinterpret_cast<uint4*>(smem_raw) + (size_t)warp * 4 * nch;
    const Calib& cal = ix.calib;
    uint32_t staged_q = kInvalid;
    QParams qp;
    qp.a = cal.affine_a; qp.b = cal.affine_b; qp.floor_ = cal.ip_qo_floor;
    qp.A = qp.Bc = qp.C = 0.0f; qp.slack = 0.0f;

    const uint64_t stride = (uint64_t)gridDim.x * kFsWarps;
    for (uint64_t i = (uint64_t)blockIdx.x * kFsWarps + warp; i < a.nblocks; i += stride) {
        const uint32_t q = a.query_of_block ? __ldg(a.query_of_block + i) : 0u;
        if (q != staged_q) {
            __syncwarp();
            const uint4* us = reinterpret_cast<const uint4*>(a.uplanes + (size_t)q * 16 * nch);
            for (uint32_t j = lane; j < 4 * nch; j += 32) uqs[j] = us[j];
            const float* cf = a.coeffs + (size_t)q * kCoeffStride;
            qp.A = cf[0]; qp.Bc = cf[1]; qp.C = cf[2];
            staged_q = q;
            __syncwarp();
        }
        const uint64_t v = a.vertex_ids ? (uint64_t)__ldg(a.vertex_ids + i) : a.first_vertex + i;
        const uint8_t* blk = ix.blocks + v * ix.block_stride;
        const uint8_t* aux = blk + ix.aux_off;
        const uint32_t count = __ldg(reinterpret_cast<const uint32_t*>(aux + 640));
        const float nop = __ldg(reinterpret_cast<const float*>(aux + 128) + lane);
        const float ipqo = __ldg(reinterpret_cast<const float*>(aux + 256) + lane);
        const float ipcp = __ldg(reinterpret_cast<const float*>(aux + 384) + lane);
        const uint32_t pops = __ldg(reinterpret_cast<const uint32_t*>(aux + 512) + lane);
        const float dqp = __ldg(a.dqp + i);
        int li = a.slack_level ? __ldg(a.slack_level + i) : 0;
        if (cal.num_slack > 0) { li = li < cal.num_slack - 1 ? li : cal.num_slack - 1; qp.slack = cal.slack[li < 0 ? 0 : li]; }
        else qp.slack = 0.0f;

        uint32_t ps[B];
        plane_sums<B>(reinterpret_cast<const uint4*>(blk), nch, lane, uqs, ps);
        uint32_t nbit, msb, msb2;
        combine_planes<B>(ps, nbit, msb, msb2);
        float est, lower, msb_lower;
        const float sq = __fsqrt_rn(dqp);
        if (B == 1) {
            convert_1bit(qp, nbit, nop, ipqo, ipcp, pops & 0xFFFFu, lane, count, dqp, sq, est, lower);
            msb_lower = lower;
        } else {
            msb_lower = convert_msb<B>(qp, msb2, nop, ipqo, ipcp, pops & 0xFFFFu, dqp, sq);
            convert_nbit<B>(qp, nbit, msb, nop, ipqo, ipcp, pops & 0xFFFFu, pops >> 16, lane, count, dqp, sq, est, lower);
        }
        if (lane >= count) { est = FLT_MAX; lower = FLT_MAX; msb_lower = FLT_MAX; }
        const size_t o = (size_t)i * 32 + lane;
        if (a.nbit) a.nbit[o] = nbit;
        if (a.msb) a.msb[o] = msb;
        if (a.msb2) a.msb2[o] = msb2;
        if (a.est) a.est[o] = est;
        if (a.lower) a.lower[o] = lower;
        if (a.msb_lower) a.msb_lower[o] = msb_lower;
    }
}

cudaError_t launch_fastscan_blocks(const DevIndex& ix, const FastScanArgs& a, int num_sms, cudaStream_t stream) {
    if (a.nblocks == 0) return cudaSuccess;
    const size_t smem = (size_t)kFsWarps * 64 * ix.nch;
    uint64_t want = (a.nblocks + kFsWarps - 1) / kFsWarps;
    const uint64_t cap = (uint64_t)num_sms * 8;
    const int grid = (int)(want < cap ? want : cap);
    if (ix.B == 1) fastscan_blocks_kernel<1><<<grid, kFsWarps * 32, smem, stream>>>(ix, a);
    else if (ix.B == 2) fastscan_blocks_kernel<2><<<grid, kFsWarps * 32, smem, stream>>>(ix, a);
    else fastscan_blocks_kernel<4><<<grid, kFsWarps * 32, smem, stream>>>(ix, a);
    return cudaGetLastError();
}

}  // namespace cpb
