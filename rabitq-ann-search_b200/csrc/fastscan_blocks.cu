// K2 -- FastScan estimator over a list (or a contiguous range) of neighbour blocks.
//
// Replaces fastscan::compute_inner_products / compute_nbit_inner_products /
// compute_msb_only_inner_products (distance/fastscan_kernel.hpp:17-87,197-217,349-368) and the
// epilogues convert_to_distances_with_bounds / convert_msb_to_lower_bounds /
// convert_nbit_to_distances_with_bounds (:89-194,371-425,220-346).  This is the stand-alone form of
// the estimator the search kernel runs per expansion: the parity hook for integer sums and float
// estimates, and the kernel behind the "FastScan HBM GB/s" figure (a pure stream of blocks).
//
// One warp per block, lane = neighbour slot.  Each warp owns a ring of NS shared-memory stages; lane 0
// keeps NS bulk asynchronous copies (cp.async.bulk, the TMA engine) in flight, one block per stage,
// each completing on the stage's own mbarrier, while the warp runs the popcount / epilogue work of
// the oldest stage out of shared memory.  Persistent grid (SMs x resident CTAs), blocks strided over
// all warps.  The query's bit-planes sit in shared memory per warp and are re-staged only when the
// block's query changes.
#include <float.h>

#include "device_math.cuh"
#include "kernels.h"

namespace cpb {

constexpr int kFsMaxWarps = 8;

#ifndef CPB_HOST_EMULATION
__device__ __forceinline__ uint32_t fs_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fs_mbar_init(uint64_t* bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fs_smem_u32(bar)));
}
__device__ __forceinline__ void fs_mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// (Tried and not kept: warps PULLING blocks from a per-CTA shared-memory counter, one 32-warp CTA per SM, instead of a fixed
// stride per warp.  With the fixed stride ncu shows 23 of 32 warps active on average, every SM alike -- the schedulers'
// priorities let some warps finish early -- and the pull does keep all 32 busy to the end (warps active 36.6 % -> 49.7 %), but
// the issue rate does not follow (70.8 % -> 68.4 %) and the stream is 9 % slower on the same box: with the ALU pipe at 66 %,
// the XU pipe at 55 % and issue at 71 % the kernel sits on three nearly equal limits at once, and extra resident warps add
// contention, not throughput.  Staggering the warps' start by fractions of a block's time changes nothing either.)
// (Tried and not kept: two bulk copies per block that skip the 128 bytes of neighbour ids the estimator never reads -- 21
// lines of traffic instead of 22.  The second copy costs more than the line saves: 0.642 of the HBM peak against 0.668.)
__device__ __forceinline__ void fs_issue(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fs_smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(fs_smem_u32(dst)), "l"(src), "r"(bytes), "r"(fs_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fs_wait(uint64_t* bar, uint32_t phase) {
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(fs_smem_u32(bar)), "r"(phase) : "memory");
    } while (!done);
}
#else
// host emulation (tests/native/): the bulk copy is a memcpy by the issuing thread; the barrier word counts completed
// phases, and a parity wait returns once the phase of that parity is over -- the same protocol, minus the hardware
inline void fs_mbar_init(uint64_t* bar) { __atomic_store_n(bar, 0, __ATOMIC_RELEASE); }
inline void fs_mbar_init_fence() {}
inline void fs_issue(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    memcpy(dst, src, bytes);
    __atomic_fetch_add(bar, 1, __ATOMIC_RELEASE);
}
inline void fs_wait(uint64_t* bar, uint32_t phase) {
    while ((__atomic_load_n(bar, __ATOMIC_ACQUIRE) & 1u) == phase) std::this_thread::yield();
}
#endif

// LEAN = the streaming case (a contiguous range of blocks, one query, slack level 0, only est and lower written): no
// per-block look-ups of which query / which vertex / which outputs.
template <int B, bool LEAN>
__global__ void __launch_bounds__(kFsMaxWarps * 32) fastscan_blocks_kernel(const DevIndex ix, const FastScanArgs a,
                                                                            uint32_t ns, uint32_t stage_bytes) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const uint32_t nch = ix.nch;
    const size_t per_warp = (size_t)ns * stage_bytes + 64 + (size_t)nch * 64;
    uint8_t* sm = smem_raw + (size_t)warp * ((per_warp + 127) & ~(size_t)127);
    uint8_t* stages = sm;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(sm + (size_t)ns * stage_bytes);   // ns <= 8 barriers
    uint4* uqs = reinterpret_cast<uint4*>(sm + (size_t)ns * stage_bytes + 64);
    const uint32_t copy_bytes = (ix.aux_off + 644u + 15u) & ~15u;

    const Calib& cal = ix.calib;
    uint32_t staged_q = kInvalid;
    QParams qp;
    qp.a = cal.affine_a; qp.b = cal.affine_b; qp.floor_ = cal.ip_qo_floor;
    qp.A = qp.Bc = qp.C = 0.0f; qp.slack = 0.0f;

    if (lane == 0) {
        for (uint32_t s = 0; s < ns; ++s) fs_mbar_init(mbar + s);
        fs_mbar_init_fence();
    }
    __syncwarp();

    const uint64_t stride = (uint64_t)gridDim.x * nwarps;
    const uint64_t first = (uint64_t)blockIdx.x * nwarps + warp;
    auto block_ptr = [&](uint64_t i) -> const uint8_t* {
        const uint64_t v = (!LEAN && a.vertex_ids) ? (uint64_t)__ldg(a.vertex_ids + i) : a.first_vertex + i;
        return ix.blocks + v * ix.block_stride;
    };
    // prologue: fill the ring
    if (lane == 0)
        for (uint32_t s = 0; s < ns; ++s) {
            const uint64_t i = first + (uint64_t)s * stride;
            if (i < a.nblocks) fs_issue(stages + (size_t)s * stage_bytes, block_ptr(i), copy_bytes, mbar + s);
        }

    // slack level of blocks without an explicit one: looked up once, not per block (a dynamic index into the
    // parameter bank is a long-latency load)
    const float slack0 = cal.num_slack > 0 ? cal.slack[0] : 0.0f;
    uint32_t s = 0, phase = 0;   // ring position (no division by the runtime stage count in the loop)
    for (uint64_t i = first; i < a.nblocks; i += stride) {
        const uint32_t q = (!LEAN && a.query_of_block) ? __ldg(a.query_of_block + i) : 0u;
        if (q != staged_q) {
            __syncwarp();
            const uint4* us = reinterpret_cast<const uint4*>(a.uplanes + (size_t)q * 16 * nch);
            for (uint32_t t = lane; t < 4 * nch; t += 32) uqs[t] = us[t];
            const float* cf = a.coeffs + (size_t)q * kCoeffStride;
            qp.A = cf[0]; qp.Bc = cf[1]; qp.C = cf[2];
            staged_q = q;
            __syncwarp();
        }
        const float dqp = __ldg(a.dqp + i);
        qp.slack = slack0;
        if (!LEAN && a.slack_level && cal.num_slack > 0) {
            int li = __ldg(a.slack_level + i);
            li = li < cal.num_slack - 1 ? li : cal.num_slack - 1;
            qp.slack = cal.slack[li < 0 ? 0 : li];
        }

        fs_wait(mbar + s, phase);
        const uint8_t* blk = stages + (size_t)s * stage_bytes;
        const uint8_t* aux = blk + ix.aux_off;
        const uint32_t count = *reinterpret_cast<const uint32_t*>(aux + 640);
        const float nop = reinterpret_cast<const float*>(aux + 128)[lane];
        const float ipqo = reinterpret_cast<const float*>(aux + 256)[lane];
        const float ipcp = reinterpret_cast<const float*>(aux + 384)[lane];
        const uint32_t pops = reinterpret_cast<const uint32_t*>(aux + 512)[lane];
        uint32_t ps[B];
        plane_sums<B, true>(reinterpret_cast<const uint4*>(blk), nch, lane, uqs, ps);
        __syncwarp();   // every lane is done reading this stage: refill it
        if (lane == 0) {
            const uint64_t nxt = i + (uint64_t)ns * stride;
            if (nxt < a.nblocks) fs_issue(stages + (size_t)s * stage_bytes, block_ptr(nxt), copy_bytes, mbar + s);
        }
        uint32_t nbit, msb, msb2;
        combine_planes<B>(ps, nbit, msb, msb2);
        float est, lower, msb_lower;
        const float sq = __fsqrt_rn(dqp);
        if (B == 1) {
            convert_1bit(qp, nbit, nop, ipqo, ipcp, pops & 0xFFFFu, lane, count, dqp, sq, est, lower);
            msb_lower = lower;
        } else {
            msb_lower = (!LEAN && a.msb_lower) ? convert_msb<B>(qp, msb2, nop, ipqo, ipcp, pops & 0xFFFFu, dqp, sq) : 0.0f;   // only when asked for
            convert_nbit<B>(qp, nbit, msb, nop, ipqo, ipcp, pops & 0xFFFFu, pops >> 16, lane, count, dqp, sq, est, lower);
        }
        if (lane >= count) { est = FLT_MAX; lower = FLT_MAX; msb_lower = FLT_MAX; }
        const size_t o = (size_t)i * 32 + lane;
        if (LEAN) { a.est[o] = est; a.lower[o] = lower; }
        else {
            if (a.nbit) a.nbit[o] = nbit;
            if (a.msb) a.msb[o] = msb;
            if (a.msb2) a.msb2[o] = msb2;
            if (a.est) a.est[o] = est;
            if (a.lower) a.lower[o] = lower;
            if (a.msb_lower) a.msb_lower[o] = msb_lower;
        }
        if (++s == ns) { s = 0; phase ^= 1u; }
    }
}

#ifndef CPB_HOST_EMULATION
cudaError_t launch_fastscan_blocks(const DevIndex& ix, const FastScanArgs& a, int num_sms, cudaStream_t stream) {
    if (a.nblocks == 0) return cudaSuccess;
    const uint32_t copy_bytes = (ix.aux_off + 644u + 15u) & ~15u;
    const uint32_t stage_bytes = (copy_bytes + 127u) & ~127u;
    // double buffering per warp and as many warps as fit: the kernel is bound by the popcount (XU) pipe,
    // which wants many warps to stay busy, not by bytes in flight (32 warps x 2 stages per SM is ample)
    uint32_t ns = 2;
    const size_t per_warp = (((size_t)ns * stage_bytes + 64 + (size_t)ix.nch * 64) + 127) & ~(size_t)127;
    int warps = (int)((200u * 1024u) / per_warp);
    warps = warps < 1 ? 1 : (warps > kFsMaxWarps ? kFsMaxWarps : warps);
    int ctas_per_sm = (int)((220u * 1024u) / (per_warp * warps));
    ctas_per_sm = ctas_per_sm < 1 ? 1 : (ctas_per_sm > 8 ? 8 : ctas_per_sm);
    const size_t smem = per_warp * warps;
    const bool lean = !a.vertex_ids && !a.query_of_block && !a.slack_level && a.est && a.lower && !a.nbit && !a.msb && !a.msb2 && !a.msb_lower;
    void (*kern)(const DevIndex, const FastScanArgs, uint32_t, uint32_t) =
        lean ? (ix.B == 1 ? fastscan_blocks_kernel<1, true> : ix.B == 2 ? fastscan_blocks_kernel<2, true> : fastscan_blocks_kernel<4, true>)
             : (ix.B == 1 ? fastscan_blocks_kernel<1, false> : ix.B == 2 ? fastscan_blocks_kernel<2, false> : fastscan_blocks_kernel<4, false>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const uint64_t want = (a.nblocks + warps - 1) / warps;
    const uint64_t cap = (uint64_t)num_sms * ctas_per_sm;
    const int grid = (int)(want < cap ? want : cap);
    kern<<<grid, warps * 32, smem, stream>>>(ix, a, ns, stage_bytes);
    return cudaGetLastError();
}

#endif

}  // namespace cpb
