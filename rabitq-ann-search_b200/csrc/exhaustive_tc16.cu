// K5, tensor-core form with the candidate screen folded into the contraction (D <= 128, k' <= 256).
//
// exhaustive_tc.cu computes fs = sum_i bit_i(v) u_i(q) on the tensor cores (kind::i8) and then spends seven
// instructions per (vertex, query) pair on the screen  fs >= thr(v,q)  with
//     thr = alpha_v x_q + s_v y_q + pc_v z_q + c_q        (x = 1/A, y = (dqp - tau)/A, z = -Bc/A, c = -C/A).
// thr is bilinear, so it can ride in the same matrix product: with f16 operands and an f32 accumulator
// (kind::f16; bits and 4-bit query values are exact in f16, their sums in f32) sixteen extra K columns hold the
// vertex factors on the A side and the query factors on the B side, each split into a high and a low f16 part
// (hi = rn16(x), lo = rn16(x - hi): x to 2^-22), and the accumulator comes out as
//     acc(v,q) = fs - thr(v,q) + margin_q                    -- a pair passes the screen iff acc >= 0:
// ONE compare per pair in the epilogue.  margin_q covers the hi/lo truncation (2^-20 of the terms' magnitudes), the
// f32 accumulation inside the tensor core however it is ordered or rounded (<= 144 steps of one ulp of the largest
// partial sum), the float error of the exact estimate's own chain, and a torn read of a threshold that is being
// lowered while MMAs are in flight; see t16_query_columns.  Pairs that pass are queued and drained as in the i8
// form; the drain recomputes fs with the popcount formulation (exact), then the op-for-op AVX2 estimate.
//   extra columns j = 0..15 (k = 128 + j):     A side (vertex)            B side (query)
//     0,1  -alpha^_hi * x^_hi, x^_lo           2,3  -s^_hi * y^_hi, y^_lo  (one aligned 32-bit word: updated atomically)
//     4    -alpha^_lo * x^_hi                  5    -s^_lo * y^_hi
//     6,7  -pc * z_hi, z_lo                    8,9  1 * c'_hi, c'_lo       (c' = -c + margin)
//     10   force_v * BIG (rows the screen does not apply to pass every present query)      11..15  zero
//   (alpha^ = alpha 2^ea, x^ = x 2^-ea and s^ = s 2^es, y^ = y 2^-es: powers of two that centre the f16 ranges)
// Thresholds must be finite before this kernel starts (a query without one would pass every pair for a whole
// segment), so the launcher first runs the i8 form over a prefix of the range: its lists are thrown away, the
// thresholds it leaves in the shared array are the k'-th smallest estimates of that prefix -- valid upper bounds.
// Pipeline, roles, work split, candidate lists and the trim protocol are those of exhaustive_tc.cu.
#include <cuda_fp16.h>

#include "exhaustive_tc_common.cuh"

namespace cpb {

constexpr int k16K = 144;                         // halves per row: 128 code bits + 16 threshold columns
constexpr int k16KCores = k16K / 8;               // 16-byte core-matrix rows per operand row
constexpr uint32_t k16SBO = k16KCores * 128;      // bytes between 8-row groups
constexpr uint32_t k16StageA = 16 * k16SBO;       // 128 rows: 36 864 B
constexpr uint32_t k16BytesB = 32 * k16SBO;       // 256 rows: 73 728 B
constexpr int k16Stages = 2;
constexpr int k16Queue = 1280;                    // passer queue entries (2 B) per epilogue warp: 32 columns x 32 lanes + a quarter
constexpr float k16Big = 60000.0f;                // representable in f16, above every |thr| the screen admits

__device__ __forceinline__ uint64_t t16_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(k16SBO >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void t16_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ uint32_t h2bits(float hi, float lo) {
    return (uint32_t)__half_as_ushort(__float2half_rn(hi)) | ((uint32_t)__half_as_ushort(__float2half_rn(lo)) << 16);
}
__device__ __forceinline__ void split16(float x, float& hi, float& lo) {
    hi = __half2float(__float2half_rn(x));
    lo = __half2float(__float2half_rn(x - hi));
}

struct T16Scale { float sa, ss, isa, iss; };   // 2^ea, 2^es and their inverses

// the sixteen B-side halves of one query (two 16-byte core rows).  pass-all / never-pass / regular.
__device__ __forceinline__ void t16_query_columns(bool present, float A, float Bc, float C, float dqp, float tau, const T16Scale& sc,
                                                  float dmax, uint4& lo8, uint4& hi8) {
    lo8 = make_uint4(0, 0, 0, 0);
    hi8 = make_uint4(0, 0, 0, 0);
    if (!present) { hi8.x = h2bits(-k16Big, 0.0f); return; }                 // acc = fs - BIG < 0: never
    const uint32_t bigrow = (uint32_t)__half_as_ushort(__float2half_rn(k16Big));   // column 10 (low half of hi8.y)
    hi8.y = bigrow;
    hi8.x = h2bits(k16Big, 0.0f);                                            // pass every pair unless the screen applies
    if (!(dqp >= 1e-12f) || !(A > 0.0f) || !(tau < kTcTauInf)) return;
    const float ia = __fdiv_rn(1.0f, A);
    const float x = ia * sc.isa, y = (dqp - tau) * ia * sc.iss, z = -Bc * ia, c = -C * ia;
    // tau only falls during the scan (never below 0), so y moves from here towards ymax
    const float ymax = fmaxf(fabsf(y), fabsf(dqp * ia * sc.iss));
    // magnitudes, in units of fs, of everything that is summed (vertex factors are at most 1024 after scaling)
    const float sabs = 1024.0f * fabsf(x) + 1024.0f * ymax + dmax * fabsf(z) + fabsf(c) + 15.0f * dmax;
    if (!(sabs < 30000.0f) || !(fabsf(x) < 30000.0f) || !(ymax < 30000.0f)) return;
    // margin = 4e-5 sabs  (<= 144 accumulation steps of one ulp of the largest partial sum, doubled for truncation: 288 x 2^-23;
    //                      hi/lo truncation of the factors and the dropped lo x lo products: 2^-19; the exact estimate's
    //                      own float chain: 2^-19)
    //        + 0.5 (|y| + ymax)  (column 5 multiplies s_lo, |s_lo| <= 0.5, by a copy of y_hi that is written after the
    //                             (y_hi, y_lo) word and may lag behind it)
    //        + 0.75
    const float margin = 4.0e-5f * sabs + 0.5f * (fabsf(y) + ymax) + 0.75f;
    float xh, xl, yh, yl, zh, zl, ch, cl;
    split16(x, xh, xl); split16(y, yh, yl); split16(z, zh, zl); split16(-c + margin, ch, cl);
    lo8.x = h2bits(xh, xl);      // columns 0, 1
    lo8.y = h2bits(yh, yl);      // columns 2, 3
    lo8.z = h2bits(xh, yh);      // columns 4, 5
    lo8.w = h2bits(zh, zl);      // columns 6, 7
    hi8.x = h2bits(ch, cl);      // columns 8, 9
}

struct T16Drain {
    const uint32_t* codes; const float* nop; const float* ipqo; const uint16_t* pop; const uint32_t* uplanes;
    float aa, ab, floor_;
    uint32_t kp, q0;
    uint64_t id_begin, m;
    uint32_t* sums; float* est;
    unsigned long long* lists;
};

struct T16Shared {
    float4 par[kTcNQ];     // A, Bc, C, |q-c|^2
    float tau[kTcNQ];
    uint32_t cnt[kTcNQ];
    uint64_t a_full[k16Stages], a_empty[k16Stages], acc_full[2], acc_empty[2];
    uint32_t qn[kTcEpiWarps];
    T16Drain drain;
    uint32_t tmem_base;
};

// queue record (16 bits): column : 8 | lane : 5 | tile-in-window : 3   (the rows of a warp are those of its lane quarter)
template <bool DENSE>
__device__ __noinline__ void t16_drain(T16Shared& sh, uint32_t wq_s, uint32_t qn_s, uint32_t lane, uint32_t wbase) {
    // wq_s / qn_s: shared-window addresses of the warp's queue and its fill count (plain 32-bit values: the callers sit in
    // the compare loop and must not carry generic pointers); wbase = id of this warp's lane 0 in the first tile of the window
    const T16Drain& d = sh.drain;
    __syncwarp();
    const uint32_t n = tc_lds(qn_s);
    for (uint32_t i = lane; i < n; i += 32) {
        unsigned short e16;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(e16) : "r"(wq_s + i * 2u) : "memory");
        const uint32_t e = e16;
        const uint32_t col = e & 0xFFu, id = wbase + ((e >> 8) & 31u) + (e >> 13) * kTcM;
        // fs = sum_i bit_i u_i, popcount form (D <= 128: one chunk)
        const uint4 w = __ldg(reinterpret_cast<const uint4*>(d.codes) + id);
        const uint4* u = reinterpret_cast<const uint4*>(d.uplanes + (size_t)(d.q0 + col) * 16);
        const uint32_t fs = weighted_popc(w, __ldg(u + 0), __ldg(u + 1), __ldg(u + 2), __ldg(u + 3));
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
__global__ void __launch_bounds__(kTcThreads, 1) exhaustive_scan_tc16_kernel(const DevIndex ix, const ExhaustiveArgs a, uint32_t tiles_per_group,
                                                                             uint64_t units_per_cta, uint32_t ngroups,
                                                                             const float* __restrict__ vstat, uint32_t* __restrict__ taug,
                                                                             unsigned long long* __restrict__ lists,
                                                                             unsigned long long* __restrict__ partial) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* As = smem_raw;                                           // k16Stages x 36 KB
    uint8_t* Bs = smem_raw + (size_t)k16Stages * k16StageA;           // 72 KB
    T16Shared& sh = *reinterpret_cast<T16Shared*>(Bs + k16BytesB);
    unsigned short* queues = reinterpret_cast<unsigned short*>(Bs + k16BytesB + ((sizeof(T16Shared) + 15) & ~(size_t)15));   // [kTcEpiWarps][k16Queue]

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t kp = a.kprime;
    const Calib& cal = ix.calib;
    const uint64_t m = a.id_end - a.id_begin;
    const float dmax = 128.0f;
    unsigned long long* mylists = lists + (size_t)blockIdx.x * kTcNQ * kTcListStride;
    // lists of cap = 1024 - k' keys (two register-resident selections trim them); a checkpoint -- drain, trim, tau -- every G
    // tiles, G = the tiles a just-trimmed list of k' keys can take at 128 appends per tile (at most 7: 3 bits in the queue record)
    const uint32_t cap = 2u * kTcCap - kp;
    const uint32_t G = kp ? min(7u, (cap - kp) / kTcM) : 1u;

    TcLimits lim;
    T16Scale sc;
    {
        const float c = vstat[2];
        lim.slim = c > 0.0f ? 16.0f * vstat[0] / c : 0.0f;
        lim.alim = c > 0.0f ? 16.0f * vstat[1] / c : 0.0f;
        // powers of two that put the limits just under 1024
        sc.ss = lim.slim > 0.0f ? exp2f(9.0f - floorf(log2f(lim.slim))) : 1.0f;
        sc.sa = lim.alim > 0.0f ? exp2f(9.0f - floorf(log2f(lim.alim))) : 1.0f;
        sc.iss = 1.0f / sc.ss;
        sc.isa = 1.0f / sc.sa;
    }

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&sh.tmem_base)), "r"(kTcCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid < (uint32_t)kTcEpiWarps) sh.qn[tid] = 0;
    if (tid == 0) {
        for (int s = 0; s < k16Stages; ++s) {
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
    // instruction descriptor: D = f32, A = B = f16, both K-major, N = 256, M = 128
    const uint32_t idesc = (1u << 4) | ((uint32_t)(kTcNQ >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);

    uint32_t step = 0, tcount = 0;
    const uint64_t nunits = (uint64_t)ngroups * tiles_per_group;
    const uint64_t u0 = (uint64_t)blockIdx.x * units_per_cta;
    const uint64_t u1 = min(nunits, u0 + units_per_cta);

    for (uint64_t u = u0; u < u1;) {
        const uint32_t grp = (uint32_t)(u / tiles_per_group), tile0 = (uint32_t)(u % tiles_per_group);
        const uint32_t ntiles = (uint32_t)min((uint64_t)(tiles_per_group - tile0), u1 - u);
        u += ntiles;
        const uint32_t slice = blockIdx.x - (uint32_t)(((uint64_t)grp * tiles_per_group) / units_per_cta);
        const uint32_t q0 = grp * kTcNQ;
        const uint32_t nqt = min((uint32_t)kTcNQ, a.nq - q0);
        const uint64_t vb = a.id_begin + (uint64_t)tile0 * kTcM;
        const uint64_t ve = min(a.id_end, vb + (uint64_t)ntiles * kTcM);

        // ---- item prologue: the query operand B (values as f16 + threshold columns), constants, empty lists -----
        for (uint32_t i = tid; i < (uint32_t)kTcNQ * 16; i += blockDim.x) {       // 16 core rows of 8 dimensions per query
            const uint32_t kc = i & 15u, n = i >> 4;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (n < nqt) {
                const uint2 b8 = *reinterpret_cast<const uint2*>(a.ubytes + (size_t)(q0 + n) * 128 + kc * 8);   // 8 values 0..15
                // u8 -> f16: integers up to 2048 are (0x6400 | v) - 1024 in f16; here simply convert
                const uint32_t w[2] = {b8.x, b8.y};
                uint32_t o[4];
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const uint32_t lo = (w[p >> 1] >> ((p & 1) * 16)) & 0xFFu, hi = (w[p >> 1] >> ((p & 1) * 16 + 8)) & 0xFFu;
                    o[p] = (uint32_t)__half_as_ushort(__uint2half_rn(lo)) | ((uint32_t)__half_as_ushort(__uint2half_rn(hi)) << 16);
                }
                v = make_uint4(o[0], o[1], o[2], o[3]);
            }
            *reinterpret_cast<uint4*>(Bs + (size_t)(n >> 3) * k16SBO + kc * 128 + (n & 7u) * 16) = v;
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
            uint4 lo8, hi8;
            t16_query_columns(i < nqt, p.x, p.y, p.z, p.w, DENSE ? FLT_MAX : tau, sc, dmax, lo8, hi8);
            uint8_t* row = Bs + (size_t)(i >> 3) * k16SBO + (i & 7u) * 16;
            *reinterpret_cast<uint4*>(row + 16 * 128) = lo8;
            *reinterpret_cast<uint4*>(row + 17 * 128) = hi8;
        }
        if (tid == 0)
            sh.drain = T16Drain{ix.flat_codes, ix.flat_nop, ix.flat_ipqo, ix.flat_pop, a.uplanes, cal.affine_a, cal.affine_b, cal.ip_qo_floor,
                                kp, q0, a.id_begin, m, a.sums, a.est, mylists};
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();

        if (warp < (uint32_t)kTcExpWarps) {
            // ================= expanders: thread = vertex row ==============================================
            const uint32_t row = tid;
            uint8_t* rowoff = As + (row >> 3) * k16SBO + (row & 7u) * 16;
            uint4 nxt = make_uint4(0, 0, 0, 0);
            float nop_n = 0.0f, ipqo_n = 0.0f;
            uint32_t pop_n = 0;
            if (ntiles && vb + row < ve) {
                nxt = __ldg(reinterpret_cast<const uint4*>(ix.flat_codes + (vb + row) * 4));
                nop_n = __ldg(ix.flat_nop + vb + row); ipqo_n = __ldg(ix.flat_ipqo + vb + row); pop_n = __ldg(ix.flat_pop + vb + row);
            }
            for (uint32_t t = 0; t < ntiles; ++t, ++step) {
                const uint4 w = nxt;
                const uint64_t v0 = vb + (uint64_t)t * kTcM + row;
                float sv, av;
                bool force;
                tc_vertex_params(cal, nop_n, ipqo_n, lim, sv, av, force);
                if (DENSE || !(v0 < ve)) force = true;
                const float pcf = (float)pop_n;
                {
                    const uint64_t v1 = v0 + kTcM;
                    if (t + 1 < ntiles && v1 < ve) {
                        nxt = __ldg(reinterpret_cast<const uint4*>(ix.flat_codes + v1 * 4));
                        nop_n = __ldg(ix.flat_nop + v1); ipqo_n = __ldg(ix.flat_ipqo + v1); pop_n = __ldg(ix.flat_pop + v1);
                    } else nxt = make_uint4(0, 0, 0, 0);
                }
                // threshold columns of this vertex
                uint4 lo8 = make_uint4(0, 0, 0, 0), hi8 = make_uint4(0, 0, 0, 0);
                hi8.x = h2bits(1.0f, 1.0f);                                   // columns 8, 9: the per-query constant
                if (force) hi8.y = (uint32_t)__half_as_ushort(__float2half_rn(1.0f));   // column 10
                else {
                    float ah, al, sh_, sl;
                    split16(-av * sc.sa, ah, al); split16(-sv * sc.ss, sh_, sl);
                    lo8.x = h2bits(ah, ah);        // columns 0, 1
                    lo8.y = h2bits(sh_, sh_);      // columns 2, 3
                    lo8.z = h2bits(al, sl);        // columns 4, 5
                    lo8.w = h2bits(-pcf, -pcf);    // columns 6, 7
                }
                const uint32_t s = step % k16Stages;
                tc_wait_sleep<200>(&sh.a_empty[s], ((step / k16Stages) & 1u) ^ 1u);
                uint8_t* arow = rowoff + (size_t)s * k16StageA;
                // code bits -> f16 0.0 / 1.0, 8 dimensions (16 bytes) per core row
                const uint32_t wd[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int kc = 0; kc < 16; ++kc) {
                    const uint32_t b = (wd[kc >> 2] >> ((kc & 3) * 8)) & 0xFFu;
                    uint4 o;
                    o.x = ((b & 1u) | ((b & 2u) << 15)) * 0x3C00u;
                    o.y = (((b >> 2) & 1u) | (((b >> 2) & 2u) << 15)) * 0x3C00u;
                    o.z = (((b >> 4) & 1u) | (((b >> 4) & 2u) << 15)) * 0x3C00u;
                    o.w = (((b >> 6) & 1u) | (((b >> 6) & 2u) << 15)) * 0x3C00u;
                    *reinterpret_cast<uint4*>(arow + kc * 128) = o;
                }
                *reinterpret_cast<uint4*>(arow + 16 * 128) = lo8;
                *reinterpret_cast<uint4*>(arow + 17 * 128) = hi8;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) tc_arrive(&sh.a_full[s]);
            }
        } else if (warp == (uint32_t)(kTcExpWarps + kTcEpiWarps)) {
            // ================= issuer: one thread ==========================================================
            if (lane == 0) {
                for (uint32_t t = 0; t < ntiles; ++t, ++tcount, ++step) {
                    const uint32_t buf = tcount & 1u, s = step % k16Stages;
                    tc_wait_relaxed<2000>(&sh.acc_empty[buf], ((tcount >> 1) & 1u) ^ 1u);
                    tc_wait_relaxed<1000>(&sh.a_full[s], (step / k16Stages) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                    for (uint32_t j = 0; j < (uint32_t)(k16K / 16); ++j)
                        t16_mma(tmem_base + buf * kTcNQ, t16_desc(tc_smem_u32(As) + s * k16StageA + j * 256), t16_desc(tc_smem_u32(Bs) + j * 256),
                                idesc, j > 0 ? 1u : 0u);
                    tc_commit(&sh.a_empty[s]);
                    tc_commit(&sh.acc_full[buf]);
                }
            }
        } else {
            // ================= epilogue: thread = vertex, 64 queries per warp ================================
            const uint32_t e = warp - kTcExpWarps, quarter = warp & 3u, cg = e >> 2;
            const uint32_t row = quarter * 32 + lane;
            const uint32_t colbase = cg * 64;
            const uint32_t own0 = colbase + quarter * 16;
            const uint32_t myq_s = tc_smem_u32(queues) + e * (uint32_t)(k16Queue * 2), myqn_s = tc_smem_u32(&sh.qn[0]) + e * 4u;
            uint32_t tw = 0, nckpt = 0;   // tile within the checkpoint window, checkpoints so far
            for (uint32_t t = 0; t < ntiles; ++t, ++tcount) {
                const bool live = row < (uint32_t)min((uint64_t)kTcM, ve - (vb + (uint64_t)t * kTcM));
                const uint32_t rowtag = (lane | (tw << 5)) << 8;
                const uint32_t wbase = (uint32_t)(vb + (uint64_t)(t - tw) * kTcM) + quarter * 32u;
                const uint32_t buf = tcount & 1u;
                tc_wait_sleep<100>(&sh.acc_full[buf], (tcount >> 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
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
                    // one compare per pair: acc = fs - thr + margin >= 0 (the sign bit of a float is the sign of the int)
                    if (tc_lds(myqn_s) > (uint32_t)(k16Queue - 1024)) t16_drain<DENSE>(sh, myq_s, myqn_s, lane, wbase);   // room for this half
#pragma unroll
                    for (int c16 = 0; c16 < 2; ++c16) {
                        // the sign bit of an AND survives iff every acc in it is negative: four groups of four columns, then all
                        uint32_t g4[4];
#pragma unroll
                        for (int g = 0; g < 4; ++g)
                            g4[g] = r[c16 * 16 + g * 4] & r[c16 * 16 + g * 4 + 1] & r[c16 * 16 + g * 4 + 2] & r[c16 * 16 + g * 4 + 3];
                        const uint32_t any = g4[0] & g4[1] & g4[2] & g4[3];
                        if (live && !(any & 0x80000000u)) {
#pragma unroll
                            for (int g = 0; g < 4; ++g)
                                if (!(g4[g] & 0x80000000u)) {
#pragma unroll
                                    for (int jj = 0; jj < 4; ++jj)
                                        if (!(r[c16 * 16 + g * 4 + jj] & 0x80000000u)) {
                                            uint32_t pos;
                                            asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(pos) : "r"(myqn_s) : "memory");
                                            asm volatile("st.shared.u16 [%0], %1;" ::"r"(myq_s + pos * 2u),
                                                         "h"((unsigned short)((col0 + (uint32_t)(c16 * 16 + g * 4 + jj)) | rowtag)) : "memory");
                                        }
                                }
                        }
                    }
                }

                const bool checkpoint = tw == G - 1u || t + 1 == ntiles;
                if (checkpoint || DENSE) t16_drain<DENSE>(sh, myq_s, myqn_s, lane, wbase);
                tw = checkpoint ? 0u : tw + 1u;
                if (kp && checkpoint) {
                    tc_group_sync(1 + cg);
                    const uint32_t mycol = own0 + (lane & 15u);
                    const uint32_t need = __ballot_sync(kFull, lane < 16 && sh.cnt[mycol] + G * kTcM > cap);
                    uint32_t todo = need;
                    while (todo) {
                        const uint32_t jl = __ffs(todo) - 1;
                        todo &= todo - 1;
                        const uint32_t col = own0 + jl;
                        uint32_t tb;
                        const uint32_t nc = tc_select_long(mylists + (size_t)col * kTcListStride, sh.cnt[col], kp, lane, tb);
                        if (lane == 0) {
                            const uint32_t old = atomicMin(taug + q0 + col, tb);
                            sh.cnt[col] = nc;
                            sh.tau[col] = __uint_as_float(min(old, tb));
                        }
                    }
                    __syncwarp();
                    const bool refresh = (++nckpt & 1u) == 0u;
                    if (lane < 16 && mycol < nqt && (refresh || ((need >> lane) & 1u))) {
                        float tau = sh.tau[mycol];
                        if (refresh) { const float tg = __uint_as_float(taug[q0 + mycol]); if (tg < tau) tau = tg; }
                        const float4 P = sh.par[mycol];
                        sh.tau[mycol] = tau;
                        // a lower tau only moves y = (dqp - tau)/A: rewrite its words (the (hi, lo) pair is one aligned word;
                        // the copy of y_hi multiplied by s_lo may lag, which the margin allows for).  A query that had no
                        // regular columns (pass-all) keeps them: its constant column would have to change too.
                        uint8_t* rowp = Bs + (size_t)(mycol >> 3) * k16SBO + (mycol & 7u) * 16 + 16 * 128;
                        const uint32_t cur = *reinterpret_cast<const uint32_t*>(rowp + 4);
                        if (cur != 0u && P.x > 0.0f && tau < kTcTauInf) {
                            const float y = (P.w - tau) * __fdiv_rn(1.0f, P.x) * sc.iss;
                            float yh, yl;
                            split16(y, yh, yl);
                            if (fabsf(y) < 30000.0f) {
                                *reinterpret_cast<uint32_t*>(rowp + 4) = h2bits(yh, yl);
                                *reinterpret_cast<unsigned short*>(rowp + 10) = __half_as_ushort(__float2half_rn(yh));
                            }
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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
                    c = tc_select_long(lst, c, kp, lane, tb);
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

// thresholds from the prefix pass: the k'-th smallest key of everything its segments kept, one warp per query
__global__ void __launch_bounds__(128) t16_tau_from_partial_kernel(const unsigned long long* __restrict__ partial, uint32_t nseg, uint32_t nq,
                                                                   uint32_t kp, uint32_t* __restrict__ taug) {
    const uint32_t lane = threadIdx.x & 31, q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    const uint32_t total = nseg * kp;
    uint32_t cur = 0;
    for (int bit = 30; bit >= 0; --bit) {
        const uint32_t t = cur | (1u << bit);
        uint32_t nl = 0;
        for (uint32_t i = lane; i < total; i += 32) {
            const uint32_t s = i / kp, j = i % kp;
            nl += (uint32_t)(partial[((size_t)s * nq + q) * kp + j] >> 32) < t ? 1u : 0u;
        }
        nl = __reduce_add_sync(kFull, nl);
        if (nl < kp) cur = t;
    }
    // fewer than k' keys in all: the search ends on 0x7FFFFFFF, i.e. no threshold
    if (lane == 0 && cur < 0x7F000000u) atomicMin(taug + q, cur);
}

// thresholds handed in by the caller (other pieces of the scan, other shards): taug = min(taug, tau_in)
__global__ void t16_seed_tau_kernel(uint32_t* __restrict__ taug, const float* __restrict__ tau_in, uint32_t nq) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const float t = tau_in[q];
    if (t >= 0.0f && t < kTcTauInf) atomicMin(taug + q, __float_as_uint(t));
}

cudaError_t launch_seed_tau(uint32_t* taug, const float* tau_in, uint32_t nq, cudaStream_t stream) {
    t16_seed_tau_kernel<<<(nq + 255) / 256, 256, 0, stream>>>(taug, tau_in, nq);
    return cudaGetLastError();
}

bool exhaustive_tc16_applicable(const DevIndex& ix, uint32_t kprime) {
    return kprime <= kTcMaxKPrime && ix.nch == 1 && ix.calib.affine_a > 0.0f;
}

cudaError_t launch_exhaustive_scan_tc16(const DevIndex& ix, const ExhaustiveArgs& a, int num_sms, unsigned long long* partial,
                                        uint32_t* nseg, cudaStream_t stream) {
    const TcWorkspace w = tc_workspace(partial, a.nq, a.kprime, num_sms);
    const uint64_t m = a.id_end - a.id_begin;
    cudaError_t e;
    // thresholds first: the caller's, or the i8 form over a prefix of the range (its candidate lists are not kept)
    const bool seeded = a.tau_in != nullptr;
    const uint64_t prefix = seeded ? 0 : (m < 65536 ? m : 65536);
    if (a.kprime && prefix) {
        e = launch_exhaustive_tc_prepare(ix, a.id_begin, a.id_begin + prefix, a.nq, w.vstat, w.taug, true, num_sms, stream);
        if (e != cudaSuccess) return e;
        ExhaustiveArgs pa = a;
        pa.id_end = a.id_begin + prefix;
        pa.sums = nullptr; pa.est = nullptr;
        uint32_t ns = 0;
        e = launch_exhaustive_scan_tc_core(ix, pa, num_sms, partial, w, &ns, stream);
        if (e != cudaSuccess) return e;
        t16_tau_from_partial_kernel<<<(a.nq + 3) / 4, 128, 0, stream>>>(partial, ns, a.nq, a.kprime, w.taug);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    e = launch_exhaustive_tc_prepare(ix, a.id_begin, a.id_end, a.nq, w.vstat, w.taug, !(a.kprime && prefix), num_sms, stream);
    if (e != cudaSuccess) return e;
    if (seeded && a.kprime) {
        e = launch_seed_tau(w.taug, a.tau_in, a.nq, stream);
        if (e != cudaSuccess) return e;
    }
    const TcSplit sp = tc_split(m, a.nq, a.kprime, num_sms);
    *nseg = sp.nseg;
    if (a.kprime) {
        e = cudaMemsetAsync(partial, 0xFF, (size_t)sp.nseg * a.nq * (size_t)a.kprime * 8, stream);
        if (e != cudaSuccess) return e;
    }
    const size_t smem = (size_t)k16Stages * k16StageA + k16BytesB + sizeof(T16Shared) + (size_t)kTcEpiWarps * k16Queue * 2 + 1024;
    const bool dense = a.sums || a.est;
    auto kern = dense ? exhaustive_scan_tc16_kernel<true> : exhaustive_scan_tc16_kernel<false>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<sp.grid, kTcThreads, smem, stream>>>(ix, a, sp.tiles, sp.W, sp.ngroups, w.vstat, w.taug, w.lists, partial);
    return cudaGetLastError();
}

}  // namespace cpb
