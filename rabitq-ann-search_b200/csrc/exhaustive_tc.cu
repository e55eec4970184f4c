// K5, tensor-core form -- the exhaustive scan's integer contraction on the 5th-generation tensor cores.
//
// sums[v][q] = sum_i bit_i(v) * u_i(q)  (compute_inner_products, distance/fastscan_kernel.hpp:17-87) over
// a tile of 128 vertices x 16 queries is a 128 x 16 x D u8 GEMM: A = the vertices' code bits expanded to
// bytes 0/1, B = the queries' 4-bit values as bytes, D = s32 accumulators in tensor memory.  Exact
// (products <= 15, sums <= 15 D).  Per 128-dim chunk: every thread expands its vertex's 128 code bits into
// the K-major, un-swizzled canonical shared-memory layout (8-row x 16-byte core matrices; LBO = 128 B
// between core matrices along K, SBO = 1 KB between 8-row groups), one thread issues four
// tcgen05.mma.cta_group::1.kind::i8 (M = 128, N = 16, K = 32) per tile and a tcgen05.commit onto an
// mbarrier, and the epilogue reads each vertex's 16 sums back with tcgen05.ld.32x32b.x16.  A CTA is 512
// threads = 4 such tiles in flight (64 TMEM columns), thread = vertex; estimates, the running threshold and
// the per-query candidate lists are exactly those of the popcount kernel (exhaustive.cu), which remains the
// path for the shapes this one does not take (k' > 512).
#include "exhaustive_common.cuh"

namespace cpb {

constexpr int kTcThreads = 512;   // 4 vertex tiles of 128
constexpr int kTcNQ = 16;         // queries per CTA tile = MMA N
constexpr int kTcCap = 1024;      // candidate slots per query (>= k' + kTcThreads)
constexpr uint32_t kTcCols = 64;  // TMEM columns: 4 tiles x 16

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor layout, version 1)
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}

__device__ __forceinline__ void tc_wait(uint64_t* bar, uint32_t phase) {
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(tc_smem_u32(bar)), "r"(phase) : "memory");
    } while (!done);
}

// 16 code bits -> 16 bytes (0/1), little-endian bit order
__device__ __forceinline__ uint4 expand16(uint32_t bits) {
    uint4 r;
    r.x = ((bits & 0xFu) * 0x00204081u) & 0x01010101u;
    r.y = (((bits >> 4) & 0xFu) * 0x00204081u) & 0x01010101u;
    r.z = (((bits >> 8) & 0xFu) * 0x00204081u) & 0x01010101u;
    r.w = (((bits >> 12) & 0xFu) * 0x00204081u) & 0x01010101u;
    return r;
}

__global__ void __launch_bounds__(kTcThreads, 1) exhaustive_scan_tc_kernel(const DevIndex ix, const ExhaustiveArgs a,
                                                                           uint32_t nslices, uint64_t slice_len,
                                                                           const uint8_t* __restrict__ ubytes,
                                                                           unsigned long long* __restrict__ partial) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t nch = ix.nch, W = nch * 4;
    uint8_t* As = smem_raw;                                                    // 4 tiles x 16 KB
    uint8_t* Bs = smem_raw + 65536;                                            // nch x 2 KB
    unsigned long long* cand = reinterpret_cast<unsigned long long*>(smem_raw + 65536 + (size_t)nch * 2048);   // [16][1024]
    float* par = reinterpret_cast<float*>(smem_raw + 65536 + (size_t)nch * 2048 + (size_t)kTcNQ * kTcCap * 8);  // [16][4]
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    __shared__ uint32_t cnt[kTcNQ];
    __shared__ float tau[kTcNQ];
    __shared__ uint32_t need_compact;

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t slice = blockIdx.x, q0 = blockIdx.y * kTcNQ;
    const uint32_t nqt = min((uint32_t)kTcNQ, a.nq - q0);
    const uint32_t kp = a.kprime;
    const Calib& cal = ix.calib;

    // ---- one-time setup: TMEM, mbarrier, the query operand B, per-query constants ------------------------
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&tmem_base_s)), "r"(kTcCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc_smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        need_compact = 0;
    }
    // B[c][n][k]: core matrices of 8 queries x 16 bytes; absent queries are zero rows
    for (uint32_t i = tid; i < (uint32_t)kTcNQ * nch * 8; i += blockDim.x) {
        const uint32_t kc = i & 7u, n = (i >> 3) % kTcNQ, c = i / (8u * kTcNQ);
        uint4 v = make_uint4(0, 0, 0, 0);
        if (n < nqt) v = *reinterpret_cast<const uint4*>(ubytes + ((size_t)(q0 + n) * nch + c) * 128 + kc * 16);
        *reinterpret_cast<uint4*>(Bs + (size_t)c * 2048 + (n >> 3) * 1024 + kc * 128 + (n & 7u) * 16) = v;
    }
    for (uint32_t i = tid; i < nqt; i += blockDim.x) {
        const float* cf = a.coeffs + (size_t)(q0 + i) * kCoeffStride;
        par[4 * i + 0] = cf[0]; par[4 * i + 1] = cf[1]; par[4 * i + 2] = cf[2]; par[4 * i + 3] = cf[4];
    }
    if (tid < kTcNQ) { cnt[tid] = 0; tau[tid] = FLT_MAX; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    // instruction descriptor: D = s32, A = B = unsigned 8-bit, both K-major, N = 16, M = 128
    const uint32_t idesc = (2u << 4) | ((uint32_t)(kTcNQ >> 3) << 17) | ((128u >> 4) << 24);
    uint32_t phase = 0;

    const uint64_t vb = a.id_begin + (uint64_t)slice * slice_len;
    const uint64_t ve = min(a.id_end, vb + slice_len);
    const uint64_t m = a.id_end - a.id_begin;
    const uint32_t tile = tid >> 7, row = tid & 127u;
    uint8_t* arow = As + (size_t)tile * 16384 + (row >> 3) * 1024 + (row & 7u) * 16;

    for (uint64_t base = vb; base < ve; base += kTcThreads) {
        const uint64_t v = base + tid;
        const bool live = v < ve;
        float nop = 0.0f, ipqo = 0.0f, pc = 0.0f, rq = 0.0f;
        if (live) {
            nop = __ldg(ix.flat_nop + v); ipqo = __ldg(ix.flat_ipqo + v); pc = (float)__ldg(ix.flat_pop + v);
            const float qq = max_ps(ipqo, cal.ip_qo_floor);
            rq = qq > 1e-10f ? __frcp_rn(qq) : 0.0f;
        }
        for (uint32_t c = 0; c < nch; ++c) {
            uint4 w = make_uint4(0, 0, 0, 0);
            if (live) w = __ldg(reinterpret_cast<const uint4*>(ix.flat_codes + v * W) + c);
            // this vertex's row of A: 8 pieces of 16 bytes, 128 B apart (one per core matrix along K)
            *reinterpret_cast<uint4*>(arow + 0 * 128) = expand16(w.x);
            *reinterpret_cast<uint4*>(arow + 1 * 128) = expand16(w.x >> 16);
            *reinterpret_cast<uint4*>(arow + 2 * 128) = expand16(w.y);
            *reinterpret_cast<uint4*>(arow + 3 * 128) = expand16(w.y >> 16);
            *reinterpret_cast<uint4*>(arow + 4 * 128) = expand16(w.z);
            *reinterpret_cast<uint4*>(arow + 5 * 128) = expand16(w.z >> 16);
            *reinterpret_cast<uint4*>(arow + 6 * 128) = expand16(w.w);
            *reinterpret_cast<uint4*>(arow + 7 * 128) = expand16(w.w >> 16);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor-core (async proxy) reads
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j)
#pragma unroll
                    for (uint32_t ks = 0; ks < 4; ++ks)
                        tc_mma_i8(tmem_base + j * kTcNQ, tc_desc(tc_smem_u32(As) + j * 16384 + ks * 256),
                                  tc_desc(tc_smem_u32(Bs) + c * 2048 + ks * 256), idesc, (c > 0 || ks > 0) ? 1u : 0u);
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(&mbar)) : "memory");
            }
            tc_wait(&mbar, phase);   // MMAs done: A may be overwritten, accumulators are readable
            phase ^= 1u;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t fs[kTcNQ];
        {
            const uint32_t taddr = tmem_base + (((warp & 3u) * 32u) << 16) + tile * kTcNQ;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                         : "=r"(fs[0]), "=r"(fs[1]), "=r"(fs[2]), "=r"(fs[3]), "=r"(fs[4]), "=r"(fs[5]), "=r"(fs[6]), "=r"(fs[7]),
                           "=r"(fs[8]), "=r"(fs[9]), "=r"(fs[10]), "=r"(fs[11]), "=r"(fs[12]), "=r"(fs[13]), "=r"(fs[14]), "=r"(fs[15])
                         : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
#pragma unroll
        for (int t = 0; t < kTcNQ; ++t) {
            if ((uint32_t)t < nqt && live) {
                const bool dense = a.sums || a.est;
                // (dist_qp_sq < 1e-12 takes another formula: no screen there; q <= 1e-10 makes the estimate
                //  nop^2 + dqp - 2 nop b, which the screen reproduces with rq = 0)
                if (dense || par[4 * t + 3] < 1e-12f ||
                    (kp && flat_screen(par[4 * t], par[4 * t + 1], par[4 * t + 2], cal.affine_a, cal.affine_b, par[4 * t + 3], fs[t],
                                       pc, nop, rq, tau[t]))) {
                    const float est = flat_estimate(par[4 * t], par[4 * t + 1], par[4 * t + 2], cal.affine_a, cal.affine_b,
                                                    cal.ip_qo_floor, par[4 * t + 3], fs[t], pc, nop, ipqo);
                    if (a.sums) a.sums[(size_t)(q0 + t) * m + (v - a.id_begin)] = fs[t];
                    if (a.est) a.est[(size_t)(q0 + t) * m + (v - a.id_begin)] = est;
                    if (kp && est <= tau[t]) {
                        const uint32_t pos = atomicAdd(&cnt[t], 1u);   // < capacity: lists are compacted before they can fill
                        cand[(size_t)t * kTcCap + pos] = make_key(est, (uint32_t)v);
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid < nqt && cnt[tid] + kTcThreads > kTcCap) need_compact = 1;
        __syncthreads();
        if (need_compact) {
            for (uint32_t t = 0; t < nqt; ++t) {
                const uint32_t c = cnt[t];
                if (c + kTcThreads > kTcCap) {
                    unsigned long long* lst = cand + (size_t)t * kTcCap;
                    for (uint32_t i = c + tid; i < kTcCap; i += blockDim.x) lst[i] = kNoKey;
                    __syncthreads();
                    bitonic_sort(lst, kTcCap);
                    if (tid == 0) {
                        cnt[t] = min(c, kp);
                        if (c >= kp) tau[t] = __uint_as_float((uint32_t)(lst[kp - 1] >> 32));
                    }
                    __syncthreads();
                }
            }
            if (tid == 0) need_compact = 0;
            __syncthreads();
        }
    }
    if (kp) {
        for (uint32_t t = 0; t < nqt; ++t) {
            const uint32_t c = cnt[t];
            unsigned long long* lst = cand + (size_t)t * kTcCap;
            for (uint32_t i = c + tid; i < kTcCap; i += blockDim.x) lst[i] = kNoKey;
            __syncthreads();
            bitonic_sort(lst, kTcCap);
            unsigned long long* out = partial + ((size_t)slice * a.nq + (q0 + t)) * kp;
            for (uint32_t i = tid; i < kp; i += blockDim.x) out[i] = lst[i];
            __syncthreads();
        }
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTcCols) : "memory");
}

bool exhaustive_tc_applicable(const DevIndex& ix, uint32_t kprime) {
    return kprime + kTcThreads <= (uint32_t)kTcCap && ix.nch <= 8;
}

cudaError_t launch_exhaustive_scan_tc(const DevIndex& ix, const ExhaustiveArgs& a, uint32_t nslices, uint64_t slice_len,
                                      const uint8_t* ubytes, unsigned long long* partial, cudaStream_t stream) {
    const size_t smem = 65536 + (size_t)ix.nch * 2048 + (size_t)kTcNQ * kTcCap * 8 + (size_t)kTcNQ * 16;
    cudaError_t e = cudaFuncSetAttribute(exhaustive_scan_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid(nslices, (a.nq + kTcNQ - 1) / kTcNQ);
    exhaustive_scan_tc_kernel<<<grid, kTcThreads, smem, stream>>>(ix, a, nslices, slice_len, ubytes, partial);
    return cudaGetLastError();
}

}  // namespace cpb
