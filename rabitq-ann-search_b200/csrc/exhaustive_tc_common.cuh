// Pieces shared by the two tensor-core forms of the exhaustive scan (exhaustive_tc.cu: kind::i8 contraction + float
// screen in the epilogue; exhaustive_tc16.cu: kind::f16 contraction with the screen's threshold folded into it).
#pragma once
#include "exhaustive_common.cuh"

namespace cpb {

constexpr int kTcNQ = 256;           // queries per work item = MMA N
constexpr int kTcM = 128;            // vertices per tile = MMA M
constexpr int kTcStages = 3;         // A-operand ring
constexpr int kTcVRing = 8;          // per-tile vertex screen parameters: ring deeper than expander lead + accumulators in flight
constexpr int kTcCap = 512;          // candidate slots per (CTA, query): what one register-resident selection handles
constexpr int kTcListStride = 1024;  // keys reserved per list in the workspace (the f16 form fills up to 1024 - k')
constexpr uint32_t kTcMaxKPrime = 256;   // k' + one tile of appends must fit the list
constexpr int kTcExpWarps = 4, kTcEpiWarps = 16;
constexpr int kTcThreads = (kTcExpWarps + kTcEpiWarps + 1) * 32;   // + the issuer warp
constexpr int kTcQueue = 1280;        // passer queue entries (4 B) per epilogue warp: 32 columns x 32 lanes + a quarter
constexpr uint32_t kTcCols = 512;    // TMEM columns: 2 accumulators x 256
constexpr float kTcBig = 3.0e38f;
constexpr float kTcTauInf = 1.0e37f;   // tau at or above this = no threshold yet

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

__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
// for the roles that run ahead and then wait long (expanders, issuer): do not spin in the epilogue's issue slots
// poll with a real sleep between polls (ns): for many warps waiting on the same event
template <int NS>
__device__ __forceinline__ void tc_wait_sleep(uint64_t* bar, uint32_t phase) {
    uint32_t done;
    for (;;) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(tc_smem_u32(bar)), "r"(phase) : "memory");
        if (done) break;
        __nanosleep(NS);
    }
}
template <int NS>
__device__ __forceinline__ void tc_wait_relaxed(uint64_t* bar, uint32_t phase) {
    uint32_t done;
    for (;;) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(tc_smem_u32(bar)), "r"(phase), "r"((uint32_t)NS) : "memory");   // suspend-time hint, ns
        if (done) break;
    }
}
__device__ __forceinline__ void tc_group_sync(uint32_t id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// 16 code bits -> 16 bytes (0/1), little-endian bit order
__device__ __forceinline__ uint4 expand16(uint32_t bits) {
    uint4 r;
    r.x = ((bits & 0xFu) * 0x00204081u) & 0x01010101u;
    r.y = (((bits >> 4) & 0xFu) * 0x00204081u) & 0x01010101u;
    r.z = (((bits >> 8) & 0xFu) * 0x00204081u) & 0x01010101u;
    r.w = (((bits >> 12) & 0xFu) * 0x00204081u) & 0x01010101u;
    return r;
}

// ---- the screen's per-vertex and per-query numbers ----------------------------------------------------
struct TcLimits { float slim, alim; };   // vertices with s_v or |alpha_v| above these skip the screen

// s_v, alpha_v = s_v w_v; `force` = the screen does not apply to this vertex (every pair gets the exact estimate)
__device__ __forceinline__ void tc_vertex_params(const Calib& cal, float nop, float ipqo, const TcLimits& lim, float& sv, float& av,
                                                 bool& force) {
    const float qv = max_ps(ipqo, cal.ip_qo_floor);
    const float den = __fmul_rn(__fmul_rn(2.0f, nop), cal.affine_a);
    sv = __fdiv_rn(qv, den);
    av = __fmul_rn(sv, __fmul_rn(nop, __fsub_rn(nop, __fmul_rn(2.0f, cal.affine_b))));
    force = !(qv > 1e-10f) || !(den > 0.0f) || !(sv <= lim.slim) || !(fabsf(av) <= lim.alim);
    if (force) { sv = 0.0f; av = 0.0f; }
}

// k' smallest of the c <= kTcCap keys of one list, in place, by one warp; returns the new count and the k'-th
// key's estimate bits.  Keys are distinct (ids are), empty slots compare as kNoKey.
__device__ __forceinline__ uint32_t tc_select(unsigned long long* __restrict__ lst, uint32_t c, uint32_t kp, uint32_t lane, uint32_t& tau_bits) {
    constexpr int KPL = kTcCap / 32;
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
    const uint32_t r = kp - nlt;   // 1 <= r <= neq of the keys with this estimate stay
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

// the same for lists of up to 2 kTcCap - k' keys: the k' best of the first kTcCap, the rest moved up behind them, again
__device__ __forceinline__ uint32_t tc_select_long(unsigned long long* __restrict__ lst, uint32_t c, uint32_t kp, uint32_t lane, uint32_t& tau_bits) {
    if (c <= (uint32_t)kTcCap) return tc_select(lst, c, kp, lane, tau_bits);
    tc_select(lst, kTcCap, kp, lane, tau_bits);
    const uint32_t rest = c - kTcCap;                     // <= kTcCap - k': source [kTcCap, c) and destination [k', k' + rest) do not overlap
    for (uint32_t i = lane; i < rest; i += 32) lst[kp + i] = lst[kTcCap + i];
    __syncwarp();
    return tc_select(lst, kp + rest, kp, lane, tau_bits);
}

__device__ __forceinline__ uint32_t tc_lds(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ void tc_enqueue(uint32_t qn_saddr, uint32_t q_saddr, uint32_t rec) {
    uint32_t pos;
    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(pos) : "r"(qn_saddr) : "memory");
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(q_saddr + pos * 4u), "r"(rec) : "memory");
}

// workspace carve shared by the launchers: partial | lists | taug | vstat
struct TcWorkspace { unsigned long long* lists; uint32_t* taug; float* vstat; };
inline TcWorkspace tc_workspace(unsigned long long* partial, uint32_t nq, uint32_t kprime, int num_sms) {
    uint8_t* ws = reinterpret_cast<uint8_t*>(partial) + (size_t)64 * nq * (size_t)kprime * 8;
    TcWorkspace w;
    w.lists = reinterpret_cast<unsigned long long*>(ws);
    ws += (size_t)num_sms * kTcNQ * kTcListStride * 8;
    w.taug = reinterpret_cast<uint32_t*>(ws);
    ws += (((size_t)nq * 4 + 15) & ~(size_t)15);
    w.vstat = reinterpret_cast<float*>(ws);
    return w;
}
// work split of a scan over m vertices: (group, tile) units, group-major, an equal contiguous share per CTA; a group may
// not be cut into more segments than one CTA can merge (slices x k' keys sorted in shared memory by the second kernel)
struct TcSplit { uint32_t ngroups, tiles, grid, nseg; uint64_t W; };
inline TcSplit tc_split(uint64_t m, uint32_t nq, uint32_t kprime, int num_sms) {
    TcSplit s;
    s.ngroups = (nq + kTcNQ - 1) / kTcNQ;
    s.tiles = (uint32_t)((m + kTcM - 1) / kTcM);
    const uint64_t units = (uint64_t)s.ngroups * s.tiles;
    uint64_t maxseg = kprime ? 16384 / (uint64_t)kprime - 1 : 64;   // - 1: the merge may take one more list (earlier pieces' candidates)
    if (maxseg > 64) maxseg = 64;
    if (maxseg < 1) maxseg = 1;
    uint64_t W = (units + num_sms - 1) / num_sms;
    if (maxseg > 1) { const uint64_t wmin = (s.tiles + maxseg - 2) / (maxseg - 1); if (W < wmin) W = wmin; }
    else W = s.tiles;
    if (W < 1) W = 1;
    s.W = W;
    s.grid = units ? (uint32_t)((units + W - 1) / W) : 1u;
    s.nseg = s.tiles ? (uint32_t)((s.tiles + W - 1) / W) + 1u : 1u;
    return s;
}

// (exhaustive_tc.cu) limits for the screen over [id_begin, id_end) into vstat; optionally resets the shared thresholds
cudaError_t launch_exhaustive_tc_prepare(const DevIndex& ix, uint64_t id_begin, uint64_t id_end, uint32_t nq, float* vstat, uint32_t* taug,
                                         bool reset_tau, int num_sms, cudaStream_t stream);
// (exhaustive_tc16.cu) thresholds handed in by the caller: taug = min(taug, tau_in)
cudaError_t launch_seed_tau(uint32_t* taug, const float* tau_in, uint32_t nq, cudaStream_t stream);
// (exhaustive_tc.cu) the kind::i8 scan alone, limits and thresholds as they stand
cudaError_t launch_exhaustive_scan_tc_core(const DevIndex& ix, const ExhaustiveArgs& a, int num_sms, unsigned long long* partial,
                                           const TcWorkspace& w, uint32_t* nseg, cudaStream_t stream);

}  // namespace cpb
