// K1 -- batched query preparation: (optional centring) -> zero-pad -> 3 x (random sign diagonal,
// unnormalised Walsh-Hadamard) -> scale -> 4-bit scalar quantisation -> LUT / bit-planes /
// estimator coefficients.  One warp per query, everything in shared memory.
//
// Replaces, per query: Index::search's padding (api/hnsw_index.hpp:174-180),
// RandomHadamardRotation::apply_copy (encoder/rotation.hpp:34-50), fht
// (encoder/transform/fht.hpp:23-57), encode_query_raw_impl's scaling
// (encoder/rabitq_encoder.hpp:201-204) and build_lut (:98-136).
#include "device_math.cuh"
#include "kernels.h"

namespace cpb {

constexpr int kPrepWarps = 4;

__global__ void __launch_bounds__(kPrepWarps * 32)
query_prep_kernel(DevIndex ix, const float* __restrict__ queries, uint32_t nq, int center, float norm_factor,
                  float inv_sqrt_d, PrepOut out) {
    extern __shared__ float smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t q = blockIdx.x * kPrepWarps + warp;
    if (q >= nq) return;
    const uint32_t D = ix.D, dim = ix.dim, T = ix.T;
    float* buf = smem + (size_t)warp * (D + 32);
    uint8_t* u8 = reinterpret_cast<uint8_t*>(smem + (size_t)kPrepWarps * (D + 32)) + (size_t)warp * D;

    // raw query, zero-padded; the search kernels want it accumulator-major plus |q|^2
    const float* qsrc = queries + (size_t)q * dim;
    for (uint32_t i = lane; i < D; i += 32) buf[i] = i < dim ? qsrc[i] : 0.0f;
    __syncwarp();
    if (out.qT) {
        float* dst = out.qT + (size_t)q * D;
        for (uint32_t i = lane; i < D; i += 32) dst[(i & 7u) * T + (i >> 3)] = buf[i];
    }
    float qnorm = 0.0f, cnorm = 0.0f;
    {   // dot_product_simd(q, q): lanes 0..7 run the eight accumulator chains
        float acc = 0.0f;
        if (lane < 8) for (uint32_t t = 0; t < T; ++t) { const float v = buf[8 * t + lane]; acc = __fmaf_rn(v, v, acc); }
        qnorm = __shfl_sync(kFull, group_reduce8(acc), 0);
    }
    if (center) {
        __syncwarp();
        for (uint32_t i = lane; i < dim; i += 32) buf[i] = __fsub_rn(buf[i], ix.centroid[i]);
        __syncwarp();
        float acc = 0.0f;
        if (lane < 8) for (uint32_t t = 0; t < T; ++t) { const float v = buf[8 * t + lane]; acc = __fmaf_rn(v, v, acc); }
        cnorm = __shfl_sync(kFull, group_reduce8(acc), 0);
    }
    __syncwarp();

    // 3 x (diag, WHT).  Butterfly (a,b) at stride h -> (a+b, h<8 ? b-a : a-b)  
    for (int layer = 0; layer < 3; ++layer) {
        const float* sg = ix.signs + (size_t)layer * D;
        for (uint32_t i = lane; i < D; i += 32) buf[i] = __fmul_rn(buf[i], sg[i]);
        __syncwarp();
        for (uint32_t h = 1; h < D; h <<= 1) {
            for (uint32_t p = lane; p < D / 2; p += 32) {
                const uint32_t i = ((p / h) * 2 * h) + (p % h);
                const float a = buf[i], b = buf[i + h];
                buf[i] = __fadd_rn(a, b);
                buf[i + h] = (h < 8) ? __fsub_rn(b, a) : __fsub_rn(a, b);
            }
            __syncwarp();
        }
    }
    // min / max as the reference's sequential scan finds them: on equal values (+0 / -0) the
    // lowest index wins
    float vl = 0.0f, vmax = 0.0f;
    uint32_t il = kInvalid, ih = kInvalid;
    for (uint32_t i = lane; i < D; i += 32) {
        const float v = __fmul_rn(buf[i], norm_factor);
        buf[i] = v;
        if (il == kInvalid || v < vl) { vl = v; il = i; }
        if (ih == kInvalid || v > vmax) { vmax = v; ih = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ol = __shfl_xor_sync(kFull, vl, o), oh = __shfl_xor_sync(kFull, vmax, o);
        const uint32_t oil = __shfl_xor_sync(kFull, il, o), oih = __shfl_xor_sync(kFull, ih, o);
        if (oil != kInvalid && (il == kInvalid || ol < vl || (ol == vl && oil < il))) { vl = ol; il = oil; }
        if (oih != kInvalid && (ih == kInvalid || oh > vmax || (oh == vmax && oih < ih))) { vmax = oh; ih = oih; }
    }
    __syncwarp();
    if (out.rotated) for (uint32_t i = lane; i < D; i += 32) out.rotated[(size_t)q * D + i] = buf[i];

    float delta = __fdiv_rn(__fsub_rn(vmax, vl), 15.0f);
    if (delta < 1e-20f) delta = 1e-20f;
    const float inv_delta = __fdiv_rn(1.0f, delta);
    uint32_t usum = 0;
    const uint32_t nch = ix.nch;
    for (uint32_t i0 = 0; i0 < nch * 128; i0 += 32) {
        const uint32_t i = i0 + lane;
        int u = 0;
        if (i < D) {
            u = __float2int_rz(__fmaf_rn(__fsub_rn(buf[i], vl), inv_delta, 0.5f));
            u = u > 15 ? 15 : (u < 0 ? 0 : u);
            u8[i] = (uint8_t)u;
            usum += (uint32_t)u;
        }
        if (out.ubytes) out.ubytes[(size_t)q * nch * 128 + i] = (uint8_t)u;
        if (out.uplanes) {   // word i0/32 of plane t = bit t of u over these 32 dims
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const unsigned w = __ballot_sync(kFull, (u >> t) & 1);
                if (lane == 0) out.uplanes[((size_t)q * 4 + t) * nch * 4 + (i0 >> 5)] = w;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) usum += __shfl_xor_sync(kFull, usum, o);
    __syncwarp();
    if (out.lut) {
        uint8_t* lut = out.lut + (size_t)q * D * 4;
        for (uint32_t e = lane; e < D * 4; e += 32) {   // lut[j][p], e = 16 j + p
            const uint32_t j = e >> 4, pmask = e & 15;
            uint32_t s = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) if (pmask & (1u << b)) s += u8[4 * j + b];
            lut[e] = (uint8_t)s;
        }
    }
    if (lane == 0 && out.coeffs) {
        // sum_qu is a sequential f32 sum of integers <= 15 D < 2^24: exact, so order-free
        const float sum_qu = (float)usum, Df = (float)D;
        float* c = out.coeffs + (size_t)q * kCoeffStride;
        c[0] = __fmul_rn(__fmul_rn(2.0f, delta), inv_sqrt_d);
        c[1] = __fmul_rn(__fmul_rn(2.0f, vl), inv_sqrt_d);
        c[2] = __fmul_rn(-__fmaf_rn(vl, Df, __fmul_rn(delta, sum_qu)), inv_sqrt_d);
        c[3] = qnorm;
        c[4] = cnorm;
        c[5] = 0.0f; c[6] = 0.0f; c[7] = 0.0f;
    }
}

#ifndef CPB_HOST_EMULATION   // tests/native/ compiles the kernel above for the host (it has no PTX)
cudaError_t launch_query_prep(const DevIndex& ix, const float* d_queries, uint32_t nq, int center,
                              const PrepOut& out, cudaStream_t stream) {
    if (nq == 0) return cudaSuccess;
    const float Df = (float)ix.D;
    // constants of RaBitQEncoderBase's constructor (encoder/rabitq_encoder.hpp:37-39), host floats
    const float norm_factor = 1.0f / (Df * sqrtf(Df));
    const float inv_sqrt_d = 1.0f / sqrtf(Df);
    const size_t smem = (size_t)kPrepWarps * ((ix.D + 32) * sizeof(float) + ix.D);
    static bool attr_set = false;
    if (!attr_set && smem > 48 * 1024) {
        cudaFuncSetAttribute(query_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    const uint32_t grid = (nq + kPrepWarps - 1) / kPrepWarps;
    query_prep_kernel<<<grid, kPrepWarps * 32, smem, stream>>>(ix, d_queries, nq, center, norm_factor, inv_sqrt_d, out);
    return cudaGetLastError();
}

#endif

}  // namespace cpb
