// K5 -- exhaustive batched scan over the per-vertex 1-bit RaBitQ codes, exact-L2 rerank, top-k.
//
// The reference has no brute-force mode; this composes its primitives exactly as the
// test oracle does: per query q: qc = q - centroid,
// encode_query_raw(qc) (K1 with center = 1), for every vertex v in [id_begin, id_end):
//   sum  = compute_inner_products(lut, code_v)                  (distance/fastscan_kernel.hpp:17-87)
//   est  = convert_to_distances_with_bounds, AVX2 lane, with nop/ip_qo of the vertex's own code
//          (RaBitQCode<D>, encoder/rabitq_encoder.hpp:225-262), ip_cp = 0, dist_qp_sq = |qc|^2,
//          slack level 0                                          (:138-173)
// keep the k' smallest (estimate, id), exact_l2 them (search/rabitq_search.hpp:88-93), return the k
// smallest (distance, id).
//
// Two kernels.  scan: grid = (vertex slices, query tiles); a CTA walks its slice 256 vertices at a
// time (thread = vertex, code in registers) against a tile of kQT queries whose bit-planes sit in
// shared memory, and keeps per query the candidates under a running threshold tau (the k'-th smallest
// key seen so far by this CTA) in shared memory, compacting with a bitonic sort when a list fills.
// Keys are (estimate bits << 32 | id): estimates are >= 0 so unsigned order is (estimate, id) order.
// select_rerank: one CTA per query merges the slices' lists, keeps k', re-ranks them with exact
// distances in the reference's 8-accumulator order and writes the k best.
//
// This file computes the integer sums with the popcount formulation shared with K2/K3 (exact).  The dense
// Q x N contraction is the one place of the query path that maps onto tensor cores: exhaustive_tc.cu is the
// tcgen05 kind::i8 form of the scan stage (used where it applies; this one serves D > 256, k' > 256 and a
// non-positive affine_a), both feed exhaustive_select_rerank_kernel below.
#include "exhaustive_common.cuh"

namespace cpb {

__global__ void __launch_bounds__(kExThreads) exhaustive_scan_kernel(const DevIndex ix, const ExhaustiveArgs a,
                                                                     uint32_t nslices, uint64_t slice_len, uint32_t kCap,
                                                                     unsigned long long* __restrict__ partial) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t nch = ix.nch, W = nch * 4;
    unsigned long long* cand = reinterpret_cast<unsigned long long*>(smem_raw);              // [kQT][kCap]
    uint4* uq = reinterpret_cast<uint4*>(smem_raw + (size_t)kQT * kCap * 8);                 // [kQT][4][nch]
    float* par = reinterpret_cast<float*>(smem_raw + (size_t)kQT * kCap * 8 + (size_t)kQT * nch * 64);   // [kQT][4]
    CPB_BLOCK_SHARED uint32_t cnt[kQT];
    CPB_BLOCK_SHARED float tau[kQT];
    CPB_BLOCK_SHARED uint32_t need_compact;

    const uint32_t slice = blockIdx.x, q0 = blockIdx.y * kQT;
    const uint32_t nqt = min((uint32_t)kQT, a.nq - q0);
    const uint32_t kp = a.kprime;
    const Calib& cal = ix.calib;

    for (uint32_t i = threadIdx.x; i < nqt * 4 * nch; i += blockDim.x)
        uq[i] = reinterpret_cast<const uint4*>(a.uplanes + (size_t)q0 * 16 * nch)[i];
    for (uint32_t i = threadIdx.x; i < nqt; i += blockDim.x) {
        const float* cf = a.coeffs + (size_t)(q0 + i) * kCoeffStride;
        par[4 * i + 0] = cf[0]; par[4 * i + 1] = cf[1]; par[4 * i + 2] = cf[2]; par[4 * i + 3] = cf[4];   // A, Bc, C, |qc|^2
    }
    if (threadIdx.x < kQT) {   // thresholds handed in by the caller (earlier pieces, other shards), else none yet
        cnt[threadIdx.x] = 0;
        tau[threadIdx.x] = (a.tau_in && threadIdx.x < nqt && a.tau_in[q0 + threadIdx.x] >= 0.0f) ? a.tau_in[q0 + threadIdx.x] : FLT_MAX;
    }
    if (threadIdx.x == 0) need_compact = 0;
    __syncthreads();

    const uint64_t vb = a.id_begin + (uint64_t)slice * slice_len;
    const uint64_t ve = min(a.id_end, vb + slice_len);
    const uint64_t m = a.id_end - a.id_begin;

    for (uint64_t base = vb; base < ve; base += kExThreads) {
        const uint64_t v = base + threadIdx.x;
        const bool live = v < ve;
        uint32_t fs[kQT];
#pragma unroll
        for (int t = 0; t < kQT; ++t) fs[t] = 0;
        float nop = 0.0f, ipqo = 0.0f, pc = 0.0f, rq = 0.0f;
        if (live) {
            nop = __ldg(ix.flat_nop + v); ipqo = __ldg(ix.flat_ipqo + v); pc = (float)__ldg(ix.flat_pop + v);
            { const float qq = max_ps(ipqo, cal.ip_qo_floor); rq = qq > 1e-10f ? __frcp_rn(qq) : 0.0f; }
            const uint4* code = reinterpret_cast<const uint4*>(ix.flat_codes + v * W);
            for (uint32_t c = 0; c < nch; ++c) {
                const uint4 w = __ldg(code + c);
#pragma unroll
                for (int t = 0; t < kQT; ++t)
                    if ((uint32_t)t < nqt)
                        fs[t] += weighted_popc(w, uq[(t * 4 + 0) * nch + c], uq[(t * 4 + 1) * nch + c], uq[(t * 4 + 2) * nch + c],
                                               uq[(t * 4 + 3) * nch + c]);
            }
        }
#pragma unroll
        for (int t = 0; t < kQT; ++t) {
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
                        cand[(size_t)t * kCap + pos] = make_key(est, (uint32_t)v);
                    }
                }
            }
        }
        __syncthreads();
        if (threadIdx.x < nqt && cnt[threadIdx.x] + kExThreads > kCap) need_compact = 1;
        __syncthreads();
        if (need_compact) {
            for (uint32_t t = 0; t < nqt; ++t) {
                const uint32_t c = cnt[t];
                if (c + kExThreads > kCap) {   // uniform per CTA
                    unsigned long long* lst = cand + (size_t)t * kCap;
                    for (uint32_t i = c + threadIdx.x; i < kCap; i += blockDim.x) lst[i] = kNoKey;
                    __syncthreads();
                    bitonic_sort(lst, kCap);
                    if (threadIdx.x == 0) {
                        cnt[t] = min(c, kp);
                        if (c >= kp) tau[t] = __uint_as_float((uint32_t)(lst[kp - 1] >> 32));
                    }
                    __syncthreads();
                }
            }
            if (threadIdx.x == 0) need_compact = 0;
            __syncthreads();
        }
    }
    // final: k' smallest of each list, ascending, to partial[slice][q][k']
    if (kp) {
        for (uint32_t t = 0; t < nqt; ++t) {
            const uint32_t c = cnt[t];
            unsigned long long* lst = cand + (size_t)t * kCap;
            for (uint32_t i = c + threadIdx.x; i < kCap; i += blockDim.x) lst[i] = kNoKey;
            __syncthreads();
            bitonic_sort(lst, kCap);
            unsigned long long* out = partial + ((size_t)slice * a.nq + (q0 + t)) * kp;
            for (uint32_t i = threadIdx.x; i < kp; i += blockDim.x) out[i] = lst[i];
            __syncthreads();
        }
    }
}

// one CTA per query: merge the slices' lists -> k' smallest keys -> exact distances -> k smallest (distance, id)
__global__ void __launch_bounds__(kExThreads) exhaustive_select_rerank_kernel(const DevIndex ix, const ExhaustiveArgs a,
                                                                              uint32_t nslices, uint32_t sort_n,
                                                                              const unsigned long long* __restrict__ partial) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);   // [sort_n]
    const uint32_t D = ix.D, T = ix.T, Tp = T + 4;
    float* qs = reinterpret_cast<float*>(smem_raw + (((size_t)sort_n * 8 + 15) & ~(size_t)15));   // query, accumulator-major, padded rows
    const uint32_t q = blockIdx.x, kp = a.kprime, k = a.k;
    const uint32_t total = nslices * kp;
    // the slices' lists are mostly padding once the thresholds are tight: only the keys that exist are gathered (in any
    // order) and sorted -- a power of two just above their number instead of slices x k'
    CPB_BLOCK_SHARED uint32_t nvalid;
    if (threadIdx.x == 0) nvalid = 0;
    __syncthreads();
    const uint32_t ntot = total + (a.prior_keys ? kp : 0u);
    for (uint32_t i0 = 0; i0 < ntot; i0 += blockDim.x) {
        const uint32_t i = i0 + threadIdx.x;
        unsigned long long key = kNoKey;
        if (i < total) { const uint32_t s = i / kp, j = i % kp; key = partial[((size_t)s * a.nq + q) * kp + j]; }
        else if (i < ntot) key = a.prior_keys[(size_t)q * kp + (i - total)];   // one more "slice": earlier pieces
        const unsigned have = __ballot_sync(kFull, key != kNoKey);
        uint32_t base = 0;
        if ((threadIdx.x & 31u) == 0 && have) base = atomicAdd(&nvalid, (uint32_t)__popc(have));
        base = __shfl_sync(kFull, base, 0);
        if (key != kNoKey) keys[base + __popc(have & ((1u << (threadIdx.x & 31u)) - 1u))] = key;
    }
    __syncthreads();
    uint32_t nsort = 2;
    while (nsort < nvalid || nsort < kp) nsort <<= 1;     // <= sort_n; the first k' slots are read below whatever exists
    for (uint32_t i = nvalid + threadIdx.x; i < nsort; i += blockDim.x) keys[i] = kNoKey;
    for (uint32_t i = threadIdx.x; i < D; i += blockDim.x) qs[(i / T) * Tp + (i % T)] = a.qT[(size_t)q * D + i];
    __syncthreads();
    bitonic_sort(keys, nsort);
    if (a.cand_keys) {   // candidate mode: the k' best keys themselves (ascending), the threshold they imply
        const bool final_piece = a.cand_dists != nullptr;
        for (uint32_t j = threadIdx.x; j < kp; j += blockDim.x) {
            const unsigned long long key = keys[j];
            a.cand_keys[(size_t)q * kp + j] = (key != kNoKey && final_piece) ? key + a.id_offset : key;   // ids stay below 2^32 (checked by the caller)
        }
        if (a.tau_out && threadIdx.x == 0)
            a.tau_out[q] = (kp && keys[kp - 1] != kNoKey) ? __uint_as_float((uint32_t)(keys[kp - 1] >> 32)) : FLT_MAX;
        if (!final_piece) return;
    }
    // exact distances of the k' best (estimate, id): 8 lanes per vector, the reference's accumulator order
    const float qn = a.coeffs[(size_t)q * kCoeffStride + 3];
    const uint32_t g = threadIdx.x >> 3, l = threadIdx.x & 7u, ngroups = blockDim.x >> 3;
    for (uint32_t j0 = 0; j0 < kp; j0 += ngroups) {
        const uint32_t j = j0 + g;
        const unsigned long long key = j < kp ? keys[j] : kNoKey;
        const bool act = key != kNoKey;
        const uint32_t id = act ? (uint32_t)key : 0u;
        const float dot = group_chain<false>(ix.rawT + (size_t)id * D + (size_t)l * T, qs + (size_t)l * Tp, T, act);
        const float ex = act ? exact_from_dot(qn, __ldg(ix.norm_sq + id), dot) : FLT_MAX;
        __syncthreads();
        if (j < kp && l == 0) {
            keys[j] = act ? make_key(ex, id) : kNoKey;
            if (a.cand_dists) a.cand_dists[(size_t)q * kp + j] = ex;
        }
        __syncthreads();
    }
    if (a.cand_keys) return;
    // re-sort the first k' by (distance, id); entries beyond k' are not results
    for (uint32_t i = kp + threadIdx.x; i < nsort; i += blockDim.x) keys[i] = kNoKey;
    __syncthreads();
    uint32_t n2 = 1;
    while (n2 < kp) n2 <<= 1;
    bitonic_sort(keys, n2);
    for (uint32_t j = threadIdx.x; j < k; j += blockDim.x) {
        const unsigned long long key = j < n2 ? keys[j] : kNoKey;
        const bool have = key != kNoKey;
        a.ids[(size_t)q * k + j] = have ? (int64_t)(uint32_t)key : (int64_t)-1;
        a.dists[(size_t)q * k + j] = have ? __uint_as_float((uint32_t)(key >> 32)) : FLT_MAX;
    }
}

// Intermediate pieces of a scan done in pieces need no order and no exact distances: only WHICH k' keys are the smallest
// so far and the estimate of the k'-th.  One warp per query: the keys that exist (the slices' lists are mostly padding once
// thresholds are tight) are gathered into the warp's buffer, kSelCap at a time, and reduced to the k' smallest by a bitwise
// search on register-resident keys (16 per lane); more keys than one buffer holds are folded in round by round.
constexpr uint32_t kSelCap = 512;      // keys per round = 16 per lane
constexpr int kSelWarps = 8;

__device__ __forceinline__ uint32_t warp_select_smallest(unsigned long long* __restrict__ lst, uint32_t c, uint32_t kp, uint32_t lane,
                                                         uint32_t& tau_bits) {
    constexpr int KPL = kSelCap / 32;
    uint32_t hi[KPL], lo[KPL];
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
        const uint32_t idx = (uint32_t)i * 32 + lane;
        const unsigned long long key = idx < c ? lst[idx] : kNoKey;
        hi[i] = (uint32_t)(key >> 32); lo[i] = (uint32_t)key;
    }
    uint32_t cur = 0;
    for (int bit = 30; bit >= 0; --bit) {   // estimates are non-negative floats: bit 31 is clear
        const uint32_t t = cur | (1u << bit);
        uint32_t nl = 0;
#pragma unroll
        for (int i = 0; i < KPL; ++i) nl += hi[i] < t ? 1u : 0u;
        nl = __reduce_add_sync(kFull, nl);
        if (nl < kp) cur = t;
    }
    uint32_t nlt = 0, neq = 0;
#pragma unroll
    for (int i = 0; i < KPL; ++i) { nlt += hi[i] < cur ? 1u : 0u; neq += hi[i] == cur ? 1u : 0u; }
    nlt = __reduce_add_sync(kFull, nlt); neq = __reduce_add_sync(kFull, neq);
    const uint32_t r = kp - nlt;   // 1 <= r <= neq of the keys with this estimate stay: those with the smallest ids
    uint32_t cutlo = 0xFFFFFFFFu;
    if (neq != r) {
        uint32_t cl = 0;
        for (int bit = 31; bit >= 0; --bit) {
            const uint32_t t = cl | (1u << bit);
            uint32_t nl = 0;
#pragma unroll
            for (int i = 0; i < KPL; ++i) nl += (hi[i] == cur && lo[i] < t) ? 1u : 0u;
            nl = __reduce_add_sync(kFull, nl);
            if (nl < r) cl = t;
        }
        cutlo = cl;
    }
    uint32_t mine = 0;
#pragma unroll
    for (int i = 0; i < KPL; ++i) mine += (hi[i] < cur || (hi[i] == cur && lo[i] <= cutlo)) ? 1u : 0u;
    uint32_t pos = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(kFull, pos, o); if (lane >= (uint32_t)o) pos += t; }
    pos -= mine;
    __syncwarp();
#pragma unroll
    for (int i = 0; i < KPL; ++i)
        if (hi[i] < cur || (hi[i] == cur && lo[i] <= cutlo)) lst[pos++] = ((unsigned long long)hi[i] << 32) | lo[i];
    __syncwarp();
    tau_bits = cur;
    return kp;
}

__global__ void __launch_bounds__(kSelWarps * 32) exhaustive_select_keys_kernel(const ExhaustiveArgs a, uint32_t nslices,
                                                                                const unsigned long long* __restrict__ partial) {
    CPB_BLOCK_SHARED unsigned long long bufs[kSelWarps][kSelCap];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t q = blockIdx.x * kSelWarps + warp;
    if (q >= a.nq) return;
    unsigned long long* buf = bufs[warp];
    const uint32_t kp = a.kprime, total = nslices * kp, ntot = total + (a.prior_keys ? kp : 0u);
    uint32_t c = 0, tau_bits = 0x7F7FFFFFu;   // FLT_MAX
    bool full = false;
    for (uint32_t i0 = 0; i0 < ntot; i0 += 32) {
        const uint32_t i = i0 + lane;
        unsigned long long key = kNoKey;
        if (i < total) { const uint32_t s = i / kp, j = i % kp; key = partial[((size_t)s * a.nq + q) * kp + j]; }
        else if (i < ntot) key = a.prior_keys[(size_t)q * kp + (i - total)];
        const unsigned have = __ballot_sync(kFull, key != kNoKey);
        if (c + (uint32_t)__popc(have) > kSelCap) {   // the buffer cannot take this row: reduce it to the k' smallest first
            __syncwarp();
            c = warp_select_smallest(buf, c, kp, lane, tau_bits);
            full = true;
        }
        if (key != kNoKey) buf[c + __popc(have & ((1u << lane) - 1u))] = key;
        c += (uint32_t)__popc(have);
        __syncwarp();
    }
    if (c > kp) { c = warp_select_smallest(buf, c, kp, lane, tau_bits); full = true; }
    else if (c == kp) {   // exactly k' keys: the threshold is their largest estimate
        uint32_t mx = 0;
        for (uint32_t i = lane; i < c; i += 32) { const uint32_t e = (uint32_t)(buf[i] >> 32); mx = e > mx ? e : mx; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const uint32_t e = __shfl_xor_sync(kFull, mx, o); mx = e > mx ? e : mx; }
        tau_bits = mx; full = true;
    }
    for (uint32_t j = lane; j < kp; j += 32) a.cand_keys[(size_t)q * kp + j] = j < c ? buf[j] : kNoKey;
    if (a.tau_out && lane == 0) a.tau_out[q] = full ? __uint_as_float(tau_bits) : FLT_MAX;
}

#ifndef CPB_HOST_EMULATION   // tests/native/ compiles the kernels above for the host (no PTX)
// Vertex slices per query tile.  Every (slice, query tile) CTA pays a fixed price in list compactions and a
// final sort, so slices should be long (>= 64 K vertices) -- but the grid still has to fill the GPU when there
// are few queries, and one CTA must be able to merge slices x k' keys in shared memory (16 K keys).
static uint32_t pick_slices(uint64_t m, uint32_t kprime, uint32_t nq, int num_sms) {
    const uint64_t qtiles = (nq + kQT - 1) / kQT;
    uint64_t s = ((uint64_t)8 * num_sms + qtiles - 1) / qtiles;          // enough CTAs for ~8 per SM
    const uint64_t by_len = (m + 4095) / 4096;                            // never shorter than 4 K vertices
    if (s > by_len) s = by_len;
    const uint64_t cap = kprime ? (16384 / (uint64_t)kprime > 1 ? 16384 / (uint64_t)kprime - 1 : 1) : 64;   // - 1: room for the list of earlier pieces
    if (s > cap) s = cap;
    if (s > 64) s = 64;
    return (uint32_t)(s < 1 ? 1 : s);
}

size_t exhaustive_workspace_bytes(const DevIndex&, uint32_t nq, uint64_t m, uint32_t kprime, int num_sms) {
    // upper bound over pick_slices(), plus the tensor-core scan's candidate lists and shared thresholds
    return exhaustive_tc_workspace_bytes(nq, kprime, num_sms) + 256;
}

cudaError_t launch_exhaustive(const DevIndex& ix, const ExhaustiveArgs& a, int num_sms, cudaStream_t stream) {
    if (a.nq == 0) return cudaSuccess;
    if (a.kprime > kMaxKPrime) return cudaErrorInvalidValue;
    const uint64_t m = a.id_end - a.id_begin;
    const bool tc = a.use_tensor_cores && a.ubytes && exhaustive_tc_applicable(ix, a.kprime);
    uint32_t nslices = tc ? 1u : pick_slices(m, a.kprime, a.nq, num_sms);   // (tensor-core form: decided by its launcher)
    const uint64_t slice_len = m ? (m + nslices - 1) / nslices : 1;
    unsigned long long* partial = static_cast<unsigned long long*>(a.workspace);
    const uint32_t cap = a.kprime <= 384 ? 1024u : (uint32_t)kCapMax;   // cap >= k' + 2 x 256 always
    const size_t smem = (size_t)kQT * cap * 8 + (size_t)kQT * ix.nch * 64 + (size_t)kQT * 16;
    cudaError_t e = cudaFuncSetAttribute(exhaustive_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // (a short range with no thresholds handed in would be scanned twice by the f16 form -- once by its own threshold
    //  prefix -- so it goes to the i8 form directly; results are identical)
    if (tc && a.use_tensor_cores >= 2 && exhaustive_tc16_applicable(ix, a.kprime) && (a.tau_in || m > 65536 || a.sums || a.est)) {
        e = launch_exhaustive_scan_tc16(ix, a, num_sms, partial, &nslices, stream);
        if (e != cudaSuccess) return e;
    } else if (tc) {
        e = launch_exhaustive_scan_tc(ix, a, num_sms, partial, &nslices, stream);
        if (e != cudaSuccess) return e;
    } else if (m > 0 || a.kprime) {
        dim3 grid(nslices, (a.nq + kQT - 1) / kQT);
        exhaustive_scan_kernel<<<grid, kExThreads, smem, stream>>>(ix, a, nslices, slice_len, cap, partial);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    if (a.cand_keys && !a.cand_dists && a.kprime && 2 * a.kprime <= kSelCap) {   // an intermediate piece: which keys, not in which order
        exhaustive_select_keys_kernel<<<(a.nq + kSelWarps - 1) / kSelWarps, kSelWarps * 32, 0, stream>>>(a, nslices, partial);
        return cudaGetLastError();
    }
    if (a.kprime && (a.k || a.cand_keys)) {
        uint32_t sort_n = 1;
        while (sort_n < (nslices + (a.prior_keys ? 1u : 0u)) * a.kprime) sort_n <<= 1;
        const size_t smem2 = (((size_t)sort_n * 8 + 15) & ~(size_t)15) + (size_t)8 * (ix.T + 4) * 4;
        e = cudaFuncSetAttribute(exhaustive_select_rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        if (e != cudaSuccess) return e;
        exhaustive_select_rerank_kernel<<<a.nq, kExThreads, smem2, stream>>>(ix, a, nslices, sort_n, partial);
        e = cudaGetLastError();
    }
    return e;
}

#endif

// ---- merge of per-shard candidate lists ---------------------------------------------------------------------------
// One CTA per query.  Pass 1: all keys into shared memory, sorted: the k'-th smallest is the cut.  Pass 2: the entries at
// or under the cut (at most k': keys are distinct, ids being global) are collected as (distance, id) keys, sorted, and the
// k best written -- what a single scan of the whole database returns (same candidates, same tie rules).
__global__ void __launch_bounds__(kExThreads) merge_candidates_kernel(const unsigned long long* __restrict__ gkeys,
                                                                      const float* __restrict__ gdists, uint32_t lists, uint32_t nq,
                                                                      uint32_t kp, uint32_t k, uint32_t sort_n, int64_t* __restrict__ ids_out,
                                                                      float* __restrict__ dists_out, float* __restrict__ tau_out) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);   // [sort_n]
    CPB_BLOCK_SHARED unsigned long long cut;
    CPB_BLOCK_SHARED uint32_t cnt;
    const uint32_t q = blockIdx.x, total = lists * kp;
    for (uint32_t i = threadIdx.x; i < sort_n; i += blockDim.x)
        keys[i] = i < total ? gkeys[((size_t)(i / kp) * nq + q) * kp + (i % kp)] : kNoKey;
    __syncthreads();
    bitonic_sort(keys, sort_n);
    if (threadIdx.x == 0) {
        cut = kp ? keys[kp - 1] : 0ull;   // kNoKey when fewer than k' candidates exist: everything is kept
        cnt = 0;
        if (tau_out) tau_out[q] = (kp && keys[kp - 1] != kNoKey) ? __uint_as_float((uint32_t)(keys[kp - 1] >> 32)) : FLT_MAX;
    }
    __syncthreads();
    if (!gdists || !k || !ids_out) return;
    const unsigned long long c = cut;
    uint32_t n2 = 1;
    while (n2 < kp) n2 <<= 1;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n2; i += blockDim.x) keys[i] = kNoKey;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
        const size_t o = ((size_t)(i / kp) * nq + q) * kp + (i % kp);
        const unsigned long long key = gkeys[o];
        if (key != kNoKey && key <= c) {
            const uint32_t slot = atomicAdd(&cnt, 1u);
            if (slot < n2) keys[slot] = make_key(gdists[o], (uint32_t)key);
        }
    }
    __syncthreads();
    bitonic_sort(keys, n2);
    for (uint32_t j = threadIdx.x; j < k; j += blockDim.x) {
        const unsigned long long key = j < n2 ? keys[j] : kNoKey;
        const bool have = key != kNoKey;
        ids_out[(size_t)q * k + j] = have ? (int64_t)(uint32_t)key : (int64_t)-1;
        dists_out[(size_t)q * k + j] = have ? __uint_as_float((uint32_t)(key >> 32)) : FLT_MAX;
    }
}

#ifndef CPB_HOST_EMULATION
cudaError_t launch_merge_candidates(const unsigned long long* keys, const float* dists, uint32_t lists, uint32_t nq, uint32_t kprime,
                                    uint32_t k, int64_t* ids_out, float* dists_out, float* tau_out, cudaStream_t stream) {
    if (nq == 0 || kprime == 0) return cudaSuccess;
    uint32_t sort_n = 1;
    while (sort_n < lists * kprime) sort_n <<= 1;
    if (sort_n < 2) sort_n = 2;
    const size_t smem = (size_t)sort_n * 8;
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(merge_candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    merge_candidates_kernel<<<nq, kExThreads, smem, stream>>>(keys, dists, lists, nq, kprime, k, sort_n, ids_out, dists_out, tau_out);
    return cudaGetLastError();
}
#endif

}  // namespace cpb
