// K5, tensor-core form -- the exhaustive scan's integer contraction on the 5th-generation tensor cores.
//
// sums[v][q] = sum_i bit_i(v) * u_i(q)  (compute_inner_products, distance/fastscan_kernel.hpp:17-87) over a
// tile of 128 vertices x 256 queries is a 128 x 256 x D u8 GEMM: A = the vertices' code bits expanded to
// bytes 0/1, B = the queries' 4-bit values as bytes, D = s32 accumulators in tensor memory.  Exact
// (products <= 15, sums <= 15 D).
//
// One persistent CTA per SM, 21 warps in three roles connected by mbarriers:
//   expanders (4 warps, thread = vertex row)  code bits -> bytes in the K-major, un-swizzled canonical layout
//       (8-row x 16-byte core matrices; LBO = 128 B along K, SBO = 1 KB between 8-row groups), a ring of 3
//       16 KB stages; they also leave each vertex's three screen numbers in shared memory for the epilogue;
//   issuer (1 thread)  four tcgen05.mma.cta_group::1.kind::i8 (M = 128, N = 256, K = 32) per 128-dim chunk,
//       tcgen05.commit releases the stage and, after the last chunk, publishes the accumulator; two
//       accumulators (2 x 256 TMEM columns) so the next tile's MMAs run under this tile's epilogue;
//   epilogue (16 warps: 4 lane quarters x 4 column groups, thread = vertex, 64 queries each)
//       tcgen05.ld.32x32b.x32, then per (vertex, query) pair a division-free candidate screen.
// The screen: est <= tau is, in exact arithmetic, fs >= thr(v,q) with
//       thr = [ s_v (w_v + dqp_q - tau_q) - pc_v Bc_q - C_q ] / A_q,   s_v = q_v / (2 nop_v a),  w_v = nop_v (nop_v - 2 b)
// a bilinear form of three per-vertex and four per-query numbers: three FMAs and a compare per pair, with
// the float error of both sides (bounded by 2^-19 of the sum of the magnitudes of all terms, see
// tc_query_params) and the rounding of the compare itself folded into the per-query constant.  Pairs that
// pass (a few per thousand) get the exact estimate -- the op-for-op AVX2 lane of
// convert_to_distances_with_bounds (:138-173), flat_estimate -- and, if est <= tau, join the query's
// candidate list.  Lists live in HBM/L2 per (CTA, query); when one is within a tile of full its owner warp
// selects the k' smallest keys in registers (bitwise search for the k'-th key) and lowers tau.  tau is only
// ever an upper bound of the final k'-th smallest estimate, so it is shared between CTAs through a global
// array (atomicMin): what ends in the lists differs from run to run, the k' smallest keys never do.
// A work item is (group of 256 queries, slice of vertices); each item leaves its <= k' best keys in
// partial[slice][query][k'], merged and re-ranked by exhaustive_select_rerank_kernel (exhaustive.cu).
#include "exhaustive_tc_common.cuh"

namespace cpb {

// {1/A, (dqp - tau)/A, -Bc/A, cut}: pair (v, q) passes the screen iff
//     (2^23 + fs) - alpha_v x - s_v y - pc_v z  >=  cut            (three FMAs at magnitude 2^23: <= 1.5 of rounding)
// cut = 2^23 + floor(-C/A - margin) - 2, margin = 2^-19 (sum of the magnitudes of every term of thr and of the
// estimate's own chain, in units of fs): the exact float estimate differs from the real-number one by at most
// 8 ulp of its largest intermediate, thr's own evaluation by as much again; 2^-19 = 32 ulp covers both.
// Queries the screen cannot serve (dqp < 1e-12 takes another formula; A = 0; no tau yet; magnitudes too large
// for the margin to stay under half a unit) get cut = -big: every pair passes.  Absent queries never pass.
__device__ __forceinline__ float4 tc_query_params(bool valid, float A, float Bc, float C, float dqp, float tau, const TcLimits& lim,
                                                  float dmax) {
    if (!valid) return make_float4(0.0f, 0.0f, 0.0f, kTcBig);
    const float4 all = make_float4(0.0f, 0.0f, 0.0f, -kTcBig);
    if (!(dqp >= 1e-12f) || !(A > 0.0f) || !(tau < kTcTauInf)) return all;
    const float ia = __fdiv_rn(1.0f, A);
    const float y = __fmul_rn(__fsub_rn(dqp, tau), ia), z = -__fmul_rn(Bc, ia), c0 = -__fmul_rn(C, ia);
    const float sabs = lim.alim * ia + lim.slim * (fabsf(dqp) + fabsf(tau)) * ia + dmax * fabsf(z) + fabsf(c0) + 15.0f * dmax;
    if (!(sabs < 262144.0f)) return all;
    const float cut = floorf(c0 - sabs * 1.9073486328125e-6f) - 2.0f + 8388608.0f;
    return make_float4(ia, y, z, cut);
}

// mean of s_v and |alpha_v| over the scanned range -> limits (16 x mean); also resets the shared thresholds
__global__ void exhaustive_tc_prepare_kernel(const DevIndex ix, uint64_t id_begin, uint64_t id_end, uint32_t nq, float* __restrict__ acc,
                                             uint32_t* __restrict__ taug, int reset_tau) {
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    if (reset_tau) for (uint64_t i = gid; i < nq; i += stride) taug[i] = __float_as_uint(FLT_MAX);
    const Calib& cal = ix.calib;
    float s = 0.0f, a = 0.0f, c = 0.0f;
    TcLimits none{FLT_MAX, FLT_MAX};
    for (uint64_t v = id_begin + gid; v < id_end; v += stride) {
        float sv, av; bool force;
        tc_vertex_params(cal, __ldg(ix.flat_nop + v), __ldg(ix.flat_ipqo + v), none, sv, av, force);
        if (!force && sv < 1e30f && fabsf(av) < 1e30f) { s += sv; a += fabsf(av); c += 1.0f; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(kFull, s, o); a += __shfl_xor_sync(kFull, a, o); c += __shfl_xor_sync(kFull, c, o);
    }
    if ((threadIdx.x & 31) == 0 && c > 0.0f) { atomicAdd(acc + 0, s); atomicAdd(acc + 1, a); atomicAdd(acc + 2, c); }
}

// Pairs that pass the screen are rare and scattered over lanes, so the lane that finds one only queues a 4-byte
// record {fs : 12 | column : 8 | row : 7 | tile within the checkpoint window : 2} in its warp's shared-memory queue
// (a handful of divergent instructions); the queue is drained by the whole warp, one pair per lane: the exact
// estimate (flat_estimate, the op-for-op AVX2 lane), the dense parity outputs, the append to the candidate list.
struct TcDrain {
    const float* nop; const float* ipqo; const uint16_t* pop;
    float aa, ab, floor_;
    uint32_t kp, q0;
    uint64_t id_begin, m;
    uint32_t* sums; float* est;
    unsigned long long* lists;
};

struct TcShared {
    float4 qpar[kTcNQ];    // screen constants
    float4 par[kTcNQ];     // A, Bc, C, |q-c|^2
    float tau[kTcNQ];
    uint32_t cnt[kTcNQ];
    uint64_t a_full[kTcStages], a_empty[kTcStages], acc_full[2], acc_empty[2];
    uint32_t qn[kTcEpiWarps];   // passer queue fill, per epilogue warp
    TcDrain drain;              // arguments of tc_drain for the current work item
    uint32_t tmem_base;
};

// `wbase` = id of row 0 of the first tile of the current checkpoint window
template <bool DENSE>
__device__ __noinline__ void tc_drain(TcShared& sh, uint32_t wq_s, uint32_t qn_s, uint32_t lane, uint32_t wbase) {
    // (queue and fill count by shared-window address, arguments from shared memory: the callers sit in the screen loop and
    //  must not carry generic pointers or a parameter block in registers)
    const TcDrain& d = sh.drain;
    __syncwarp();
    const uint32_t n = tc_lds(qn_s);
    for (uint32_t i = lane; i < n; i += 32) {
        const uint32_t e = tc_lds(wq_s + i * 4u);
        const uint32_t fs = e & 0xFFFu, col = (e >> 12) & 0xFFu, id = wbase + (e >> 20);   // (row | tile << 7) = offset in the window
        const float4 P = sh.par[col];
        const float est = flat_estimate(P.x, P.y, P.z, d.aa, d.ab, d.floor_, P.w, fs, (float)__ldg(d.pop + id), __ldg(d.nop + id), __ldg(d.ipqo + id));
        if (DENSE) {
            const size_t o = (size_t)(d.q0 + col) * d.m + (id - d.id_begin);
            if (d.sums) d.sums[o] = fs;
            if (d.est) d.est[o] = est;
        }
        if (d.kp && est <= sh.tau[col]) {
            const uint32_t pos = atomicAdd(&sh.cnt[col], 1u);   // < capacity: lists are trimmed G tiles ahead
            d.lists[(size_t)col * kTcListStride + pos] = make_key(est, id);
        }
    }
    __syncwarp();
    if (lane == 0) asm volatile("st.shared.u32 [%0], %1;" ::"r"(qn_s), "r"(0u) : "memory");
    __syncwarp();
}

template <bool DENSE>
__global__ void __launch_bounds__(kTcThreads, 1) exhaustive_scan_tc_kernel(const DevIndex ix, const ExhaustiveArgs a, uint32_t tiles_per_group,
                                                                           uint64_t units_per_cta, uint32_t ngroups,
                                                                           const float* __restrict__ vstat, uint32_t* __restrict__ taug,
                                                                           unsigned long long* __restrict__ lists,
                                                                           unsigned long long* __restrict__ partial) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t nch = ix.nch, W = nch * 4;
    uint8_t* As = smem_raw;                                           // kTcStages x 16 KB
    uint8_t* Bs = smem_raw + (size_t)kTcStages * 16384;               // nch x 32 KB
    TcShared& sh = *reinterpret_cast<TcShared*>(smem_raw + (size_t)kTcStages * 16384 + (size_t)nch * 32768);
    uint32_t* queues = reinterpret_cast<uint32_t*>(smem_raw + (size_t)kTcStages * 16384 + (size_t)nch * 32768 + ((sizeof(TcShared) + 15) & ~(size_t)15));   // [kTcEpiWarps][kTcQueue]
    float4* vring = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(queues) + (size_t)kTcEpiWarps * kTcQueue * 4);   // [kTcVRing][kTcM]

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t kp = a.kprime;
    const Calib& cal = ix.calib;
    const uint64_t m = a.id_end - a.id_begin;
    const float dmax = (float)(nch * 128u);
    unsigned long long* mylists = lists + (size_t)blockIdx.x * kTcNQ * kTcListStride;
    // tiles between list checks: a list is trimmed when it could overflow before the next check, i.e. at
    // cnt > cap - 128 G; G = 2 leaves room for ~150 new candidates between two trims of a k' = 100 list
    const uint32_t G = kp ? min(2u, (kTcCap - kp) / kTcM) : 1u;

    TcLimits lim;
    {
        const float c = vstat[2];
        lim.slim = c > 0.0f ? 16.0f * vstat[0] / c : 0.0f;
        lim.alim = c > 0.0f ? 16.0f * vstat[1] / c : 0.0f;
    }

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&sh.tmem_base)), "r"(kTcCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid < (uint32_t)kTcEpiWarps) sh.qn[tid] = 0;
    if (tid == 0) {
        for (int s = 0; s < kTcStages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(&sh.a_full[s])), "r"(kTcExpWarps));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc_smem_u32(&sh.a_empty[s])));
        }
        for (int b = 0; b < 2; ++b) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc_smem_u32(&sh.acc_full[b])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(&sh.acc_empty[b])), "r"(kTcEpiWarps));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = sh.tmem_base;
    // instruction descriptor: D = s32, A = B = unsigned 8-bit, both K-major, N = 256, M = 128
    const uint32_t idesc = (2u << 4) | ((uint32_t)(kTcNQ >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);

    uint32_t step = 0, tcount = 0;   // running counters of this thread's role (ring stage / accumulator phases)
    // Work = (query group, tile) pairs, group-major; this CTA owns the contiguous range [u0, u1) of them, i.e. at
    // most one run of tiles ("segment") per group: its candidate lists and thresholds live for the whole run.
    const uint64_t nunits = (uint64_t)ngroups * tiles_per_group;
    const uint64_t u0 = (uint64_t)blockIdx.x * units_per_cta;
    const uint64_t u1 = min(nunits, u0 + units_per_cta);

    for (uint64_t u = u0; u < u1;) {
        const uint32_t grp = (uint32_t)(u / tiles_per_group), tile0 = (uint32_t)(u % tiles_per_group);
        const uint32_t ntiles = (uint32_t)min((uint64_t)(tiles_per_group - tile0), u1 - u);
        u += ntiles;
        // ordinal of this segment among the segments of its group (CTAs cover the group in order)
        const uint32_t slice = blockIdx.x - (uint32_t)(((uint64_t)grp * tiles_per_group) / units_per_cta);
        const uint32_t q0 = grp * kTcNQ;
        const uint32_t nqt = min((uint32_t)kTcNQ, a.nq - q0);
        const uint64_t vb = a.id_begin + (uint64_t)tile0 * kTcM;
        const uint64_t ve = min(a.id_end, vb + (uint64_t)ntiles * kTcM);

        // ---- item prologue: the query operand B, per-query constants, empty lists -------------------------
        // B[c][n][k]: core matrices of 8 queries x 16 bytes; absent queries are zero rows
        for (uint32_t i = tid; i < (uint32_t)kTcNQ * nch * 8; i += blockDim.x) {
            const uint32_t kc = i & 7u, n = (i >> 3) % kTcNQ, c = i / (8u * kTcNQ);
            uint4 v = make_uint4(0, 0, 0, 0);
            if (n < nqt) v = *reinterpret_cast<const uint4*>(a.ubytes + ((size_t)(q0 + n) * nch + c) * 128 + kc * 16);
            *reinterpret_cast<uint4*>(Bs + (size_t)c * 32768 + (n >> 3) * 1024 + kc * 128 + (n & 7u) * 16) = v;
        }
        for (uint32_t i = tid; i < (uint32_t)kTcNQ; i += blockDim.x) {
            float4 p = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            float tau = FLT_MAX;
            if (i < nqt) {
                const float* cf = a.coeffs + (size_t)(q0 + i) * kCoeffStride;
                p = make_float4(cf[0], cf[1], cf[2], cf[4]);
                if (kp) tau = __uint_as_float(taug[q0 + i]);
            }
            sh.par[i] = p; sh.tau[i] = tau; sh.cnt[i] = 0;
            sh.qpar[i] = DENSE ? make_float4(0.0f, 0.0f, 0.0f, i < nqt ? -kTcBig : kTcBig)
                               : tc_query_params(i < nqt, p.x, p.y, p.z, p.w, tau, lim, dmax);
        }
        if (tid == 0)
            sh.drain = TcDrain{ix.flat_nop, ix.flat_ipqo, ix.flat_pop, cal.affine_a, cal.affine_b, cal.ip_qo_floor, kp, q0, a.id_begin, m, a.sums, a.est, mylists};
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();

        if (warp < (uint32_t)kTcExpWarps) {
            // ================= expanders: thread = vertex row ==============================================
            const uint32_t row = tid;
            uint8_t* rowoff = As + (row >> 3) * 1024 + (row & 7u) * 16;
            const uint32_t nsteps = ntiles * nch;
            uint4 nxt = make_uint4(0, 0, 0, 0);
            float nop_n = 0.0f, ipqo_n = 0.0f;
            uint32_t pop_n = 0;
            if (nsteps && vb + row < ve) {
                nxt = __ldg(reinterpret_cast<const uint4*>(ix.flat_codes + (vb + row) * W));
                nop_n = __ldg(ix.flat_nop + vb + row); ipqo_n = __ldg(ix.flat_ipqo + vb + row); pop_n = __ldg(ix.flat_pop + vb + row);
            }
            for (uint32_t i = 0; i < nsteps; ++i, ++step) {
                const uint4 w = nxt;
                const uint32_t c0 = i % nch;
                if (c0 == 0) {
                    // the screen numbers of this tile's vertices {-s_v, -alpha_v, -pc_v, what to OR onto the accumulator}:
                    // 2^23 + fs normally; 2^126 (1 + fs 2^-23) for rows the screen does not apply to -- it clears every
                    // cut of a present query and the three FMAs, whose vertex factors are then zero, leave it alone;
                    // NaN, which clears nothing, for rows past the end of the range
                    const uint64_t v0 = vb + (uint64_t)(i / nch) * kTcM + row;
                    float sv, av;
                    bool force;
                    tc_vertex_params(cal, nop_n, ipqo_n, lim, sv, av, force);
                    const uint32_t orv = !(v0 < ve) ? 0x7FC00000u : ((force || DENSE) ? 0x7E800000u : 0x4B000000u);
                    vring[(size_t)(tcount % kTcVRing) * kTcM + row] = make_float4(-sv, -av, force ? 0.0f : -(float)pop_n, __uint_as_float(orv));
                    ++tcount;
                    const uint64_t v1 = v0 + kTcM;
                    if (v1 < ve) { nop_n = __ldg(ix.flat_nop + v1); ipqo_n = __ldg(ix.flat_ipqo + v1); pop_n = __ldg(ix.flat_pop + v1); }
                }
                if (i + 1 < nsteps) {
                    const uint32_t t1 = (i + 1) / nch, c1 = (i + 1) % nch;
                    const uint64_t v1 = vb + (uint64_t)t1 * kTcM + row;
                    nxt = v1 < ve ? __ldg(reinterpret_cast<const uint4*>(ix.flat_codes + v1 * W) + c1) : make_uint4(0, 0, 0, 0);
                }
                const uint32_t s = step % kTcStages;
                tc_wait_relaxed<4000>(&sh.a_empty[s], ((step / kTcStages) & 1u) ^ 1u);
                uint8_t* arow = rowoff + (size_t)s * 16384;
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
                __syncwarp();
                if (lane == 0) tc_arrive(&sh.a_full[s]);
            }
        } else if (warp == (uint32_t)(kTcExpWarps + kTcEpiWarps)) {
            // ================= issuer: one thread ==========================================================
            if (lane == 0) {
                for (uint32_t t = 0; t < ntiles; ++t, ++tcount) {
                    const uint32_t buf = tcount & 1u;
                    tc_wait_relaxed<2000>(&sh.acc_empty[buf], ((tcount >> 1) & 1u) ^ 1u);
                    for (uint32_t c = 0; c < nch; ++c, ++step) {
                        const uint32_t s = step % kTcStages;
                        tc_wait_relaxed<1000>(&sh.a_full[s], (step / kTcStages) & 1u);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                        for (uint32_t ks = 0; ks < 4; ++ks)
                            tc_mma_i8(tmem_base + buf * kTcNQ, tc_desc(tc_smem_u32(As) + s * 16384 + ks * 256),
                                      tc_desc(tc_smem_u32(Bs) + c * 32768 + ks * 256), idesc, (c > 0 || ks > 0) ? 1u : 0u);
                        tc_commit(&sh.a_empty[s]);       // the stage may be overwritten once these MMAs have read it
                    }
                    tc_commit(&sh.acc_full[buf]);        // accumulator complete
                }
            }
        } else {
            // ================= epilogue: thread = vertex, 64 queries per warp ================================
            const uint32_t e = warp - kTcExpWarps, quarter = warp & 3u, cg = e >> 2;
            const uint32_t row = quarter * 32 + lane;
            const uint32_t colbase = cg * 64;
            const uint32_t own0 = colbase + quarter * 16;     // the 16 lists this warp maintains between tiles
            const uint32_t myq_s = tc_smem_u32(queues) + e * (uint32_t)(kTcQueue * 4), myqn_s = tc_smem_u32(&sh.qn[0]) + e * 4u;
            for (uint32_t t = 0; t < ntiles; ++t, ++tcount) {
                const uint32_t rowtag = (row | ((t % G) << 7)) << 20;                           // queue record: where this vertex is
                const uint32_t wbase = (uint32_t)(vb + (uint64_t)(t - t % G) * kTcM);           // ... relative to this id

                const uint32_t buf = tcount & 1u;
                tc_wait_relaxed<1000>(&sh.acc_full[buf], (tcount >> 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const float4 vp = vring[(size_t)(tcount % kTcVRing) * kTcM + row];   // written by the expanders before this tile's MMAs
                const float nsv = vp.x, nav = vp.y, npc = vp.z;
                const uint32_t orv = __float_as_uint(vp.w);
#pragma unroll 1
                for (uint32_t half = 0; half < 2; ++half) {
                    const uint32_t col0 = colbase + half * 32;
                    const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + buf * kTcNQ + col0;
                    uint32_t r[32];
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                        : "r"(taddr) : "memory");
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (half == 1) {   // both halves are in registers: the accumulator may be overwritten
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) tc_arrive(&sh.acc_empty[buf]);
                    }
                    // room for every pair of this half (32 columns x 32 lanes)?
                    if (tc_lds(myqn_s) > (uint32_t)(kTcQueue - 1024)) tc_drain<DENSE>(sh, myq_s, myqn_s, lane, wbase);
                    // the screen, 7 instructions per pair (LDS.128, LOP3, 3 FFMA, FSETP.OR), branch-free over 8 columns;
                    // only a lane with a hit among its 8 pairs looks at them one by one and queues the passers
#pragma unroll
                    for (int c8 = 0; c8 < 4; ++c8) {
                        float x8[8], w8[8];
                        bool hit = false;
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) {
                            const float4 Q = sh.qpar[col0 + c8 * 8 + jj];
                            float x = __uint_as_float(r[c8 * 8 + jj] | orv);   // 2^23 + fs, exact
                            x = __fmaf_rn(nsv, Q.y, x);
                            x = __fmaf_rn(npc, Q.z, x);
                            x = __fmaf_rn(nav, Q.x, x);
                            x8[jj] = x; w8[jj] = Q.w;
                            hit = hit || (x >= Q.w);
                        }
                        if (hit) {
#pragma unroll
                            for (int jj = 0; jj < 8; ++jj)
                                if (x8[jj] >= w8[jj]) tc_enqueue(myqn_s, myq_s, r[c8 * 8 + jj] | ((col0 + (uint32_t)(c8 * 8 + jj)) << 12) | rowtag);
                        }
                    }
                }

                const bool checkpoint = t % G == G - 1 || t + 1 == ntiles;
                if (checkpoint || DENSE) tc_drain<DENSE>(sh, myq_s, myqn_s, lane, wbase);
                if (kp && checkpoint) {
                    tc_group_sync(1 + cg);   // the four warps appending to these 64 lists are done with this tile
                    const uint32_t mycol = own0 + (lane & 15u);
                    const uint32_t need = __ballot_sync(kFull, lane < 16 && sh.cnt[mycol] + G * kTcM > (uint32_t)kTcCap);
                    uint32_t todo = need;
                    while (todo) {
                        const uint32_t jl = __ffs(todo) - 1;
                        todo &= todo - 1;
                        const uint32_t col = own0 + jl;
                        uint32_t tb;
                        const uint32_t nc = tc_select(mylists + (size_t)col * kTcListStride, sh.cnt[col], kp, lane, tb);
                        if (lane == 0) {
                            const uint32_t old = atomicMin(taug + q0 + col, tb);
                            sh.cnt[col] = nc;
                            sh.tau[col] = __uint_as_float(min(old, tb));
                        }
                    }
                    __syncwarp();
                    const bool refresh = (t / G) % 3u == 2u;
                    if (lane < 16 && mycol < nqt && (refresh || ((need >> lane) & 1u))) {
                        float tau = sh.tau[mycol];
                        if (refresh) { const float tg = __uint_as_float(taug[q0 + mycol]); if (tg < tau) tau = tg; }
                        const float4 P = sh.par[mycol];
                        sh.tau[mycol] = tau;
                        sh.qpar[mycol] = tc_query_params(true, P.x, P.y, P.z, P.w, tau, lim, dmax);
                    }
                    tc_group_sync(1 + cg);
                }
            }
        }

        // ---- item epilogue: every list down to its k' best, out to partial[slice][q][k'] ----------------------
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (kp) {
            for (uint32_t col = warp; col < nqt; col += blockDim.x >> 5) {
                unsigned long long* lst = mylists + (size_t)col * kTcListStride;
                uint32_t c = sh.cnt[col];
                if (c > kp) {
                    uint32_t tb;
                    c = tc_select(lst, c, kp, lane, tb);
                    if (lane == 0) atomicMin(taug + q0 + col, tb);
                }
                unsigned long long* out = partial + ((size_t)slice * a.nq + (q0 + col)) * kp;
                for (uint32_t i = lane; i < kp; i += 32) out[i] = i < c ? lst[i] : kNoKey;
            }
        }
        __syncthreads();
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTcCols) : "memory");
}

bool exhaustive_tc_applicable(const DevIndex& ix, uint32_t kprime) {
    return kprime <= kTcMaxKPrime && ix.nch <= 2 && ix.calib.affine_a > 0.0f;
}

size_t exhaustive_tc_workspace_bytes(uint32_t nq, uint32_t kprime, int num_sms) {
    return (size_t)64 * nq * (size_t)kprime * 8 + (size_t)num_sms * kTcNQ * kTcListStride * 8 + (size_t)nq * 4 + 1024;
}

cudaError_t launch_exhaustive_tc_prepare(const DevIndex& ix, uint64_t id_begin, uint64_t id_end, uint32_t nq, float* vstat, uint32_t* taug,
                                         bool reset_tau, int num_sms, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(vstat, 0, 16, stream);
    if (e != cudaSuccess) return e;
    exhaustive_tc_prepare_kernel<<<num_sms * 2, 256, 0, stream>>>(ix, id_begin, id_end, nq, vstat, taug, reset_tau ? 1 : 0);
    return cudaGetLastError();
}

cudaError_t launch_exhaustive_scan_tc_core(const DevIndex& ix, const ExhaustiveArgs& a, int num_sms, unsigned long long* partial,
                                           const TcWorkspace& w, uint32_t* nseg, cudaStream_t stream) {
    const TcSplit sp = tc_split(a.id_end - a.id_begin, a.nq, a.kprime, num_sms);
    *nseg = sp.nseg;
    cudaError_t e;
    if (a.kprime) {   // (segment, query) slots nobody writes must read as empty
        e = cudaMemsetAsync(partial, 0xFF, (size_t)sp.nseg * a.nq * (size_t)a.kprime * 8, stream);
        if (e != cudaSuccess) return e;
    }
    const size_t smem = (size_t)kTcStages * 16384 + (size_t)ix.nch * 32768 + sizeof(TcShared) + (size_t)kTcEpiWarps * kTcQueue * 4 + (size_t)kTcVRing * kTcM * 16 + 1024;
    const bool dense = a.sums || a.est;
    auto kern = dense ? exhaustive_scan_tc_kernel<true> : exhaustive_scan_tc_kernel<false>;
    // more than half an SM's shared memory: exactly one CTA (and its 512 TMEM columns) per SM
    const size_t smem_req = smem < (size_t)120 * 1024 ? (size_t)120 * 1024 : smem;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_req);
    if (e != cudaSuccess) return e;
    kern<<<sp.grid, kTcThreads, smem_req, stream>>>(ix, a, sp.tiles, sp.W, sp.ngroups, w.vstat, w.taug, w.lists, partial);
    return cudaGetLastError();
}

// Returns through *nseg the number of per-group segments (the "slices" exhaustive_select_rerank_kernel merges).
cudaError_t launch_exhaustive_scan_tc(const DevIndex& ix, const ExhaustiveArgs& a, int num_sms, unsigned long long* partial,
                                      uint32_t* nseg, cudaStream_t stream) {
    const TcWorkspace w = tc_workspace(partial, a.nq, a.kprime, num_sms);
    cudaError_t e = launch_exhaustive_tc_prepare(ix, a.id_begin, a.id_end, a.nq, w.vstat, w.taug, true, num_sms, stream);
    if (e != cudaSuccess) return e;
    if (a.tau_in && a.kprime) {
        e = launch_seed_tau(w.taug, a.tau_in, a.nq, stream);
        if (e != cudaSuccess) return e;
    }
    return launch_exhaustive_scan_tc_core(ix, a, num_sms, partial, w, nseg, stream);
}

}  // namespace cpb
