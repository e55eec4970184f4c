// N3 -- build side: the codes of a vertex's neighbours relative to the vertex ("parent"), i.e. what
// prune_and_write stores in a neighbour block (graph/graph_refinement.hpp:46-67):
//   1-bit  RaBitQEncoder<D>::compute_neighbor_aux          (encoder/rabitq_encoder.hpp:138-181)
//   N-bit  NbitRaBitQEncoder<D,B>::compute_neighbor_aux_nbit (:287-323) with the coordinate-descent quantiser
//          caq_quantize (:371-467)
// both over rotate_raw_vector (:81-86) = the 3-layer sign/Hadamard rotation of K1.
//
// One warp per parent vertex, one LANE per neighbour: everything after the rotation is a strictly sequential
// float recurrence per (parent, neighbour) pair in the reference (running sums over the D coordinates; the
// coordinate descent updates two running sums coordinate by coordinate), so the parallelism is across pairs.
// Each lane owns one row of a [rows][D+1] shared-memory tile (the +1 makes "same coordinate, 32 rows" hit 32
// banks); all lanes walk the coordinates in step, so the sign diagonals, the rotated parent and the value table
// are broadcast reads.  Where the compiled reference fused a multiply-add and where it did not was found by
// search against it; every operation below is an explicit _rn intrinsic.
//
// The kernel uses no PTX and only warp-level primitives, so tests/native/ also compiles this file for the host
// (CPB_HOST_EMULATION: one thread per lane, barriers for the warp primitives) and checks it on the CPU.
#include "kernels.h"

namespace cpb {

namespace {

constexpr unsigned kFull = 0xFFFFFFFFu;

// S butterfly stages (h = H, 2H, ..) of the unnormalised Walsh-Hadamard transform of one row, by one thread, on 2^S
// elements H apart held in registers: one read and one write of the row per S stages.  Butterfly convention of the
// reference's FHT as in K1: (a, b) -> (a + b, h < 8 ? b - a : a - b).  pre != NULL: the row is multiplied
// by `scale` and then by the sign diagonal `pre` on the way in (the layer's diagonal; scale = 1/nop on the first layer).
template <int S>
__device__ __forceinline__ void fht_stages(float* x, uint32_t D, uint32_t cs, uint32_t H, const float* pre, bool scaled, float scale) {
    constexpr uint32_t N = 1u << S;
    for (uint32_t g = 0; g < D / N; ++g) {
        const uint32_t base = ((g & ~(H - 1)) << S) | (g & (H - 1));   // bits log2(H) .. log2(H)+S-1 are zero
        float v[N];
#pragma unroll
        for (uint32_t k = 0; k < N; ++k) {
            const uint32_t i = base + k * H;
            float t = x[i * cs];
            if (pre) {
                if (scaled) t = __fmul_rn(t, scale);
                t = __fmul_rn(t, pre[i]);
            }
            v[k] = t;
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const bool low = (H << s) < 8;
#pragma unroll
            for (uint32_t k = 0; k < N; ++k) {
                if (k & (1u << s)) continue;
                const float a = v[k], b = v[k | (1u << s)];
                v[k] = __fadd_rn(a, b);
                v[k | (1u << s)] = low ? __fsub_rn(b, a) : __fsub_rn(a, b);
            }
        }
#pragma unroll
        for (uint32_t k = 0; k < N; ++k) x[(base + k * H) * cs] = v[k];
    }
}

// one layer of the rotation on one row: (scale,) sign diagonal, transform
__device__ __forceinline__ void rotate_row(float* x, uint32_t D, uint32_t cs, const float* sg, bool scaled, float scale) {
    uint32_t H = 1;
    const float* pre = sg;
    while (H < D) {
        const uint32_t left = D / H;     // 2^(stages left)
        if (left >= 16) { fht_stages<4>(x, D, cs, H, pre, scaled, scale); H <<= 4; }
        else if (left == 8) { fht_stages<3>(x, D, cs, H, pre, scaled, scale); H <<= 3; }
        else if (left == 4) { fht_stages<2>(x, D, cs, H, pre, scaled, scale); H <<= 2; }
        else { fht_stages<1>(x, D, cs, H, pre, scaled, scale); H <<= 1; }
        pre = nullptr;
    }
}

template <int B>
__global__ void __launch_bounds__(256, 3)
neighbor_codes_kernel(NeighborCodesArgs a) {
    extern __shared__ float smem[];
    constexpr int K_INT = (1 << B) - 1;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t D = a.D, dim = a.dim, rows = a.rows;
    const uint32_t xs = D + 1;                       // row stride of the tile, floats
    const uint32_t us = D + 4;                       // row stride of the value tile, bytes (a multiple of 4, odd in words)
    float* ctab = smem;                              // the 2^B reconstruction values (2u - K) / K
    float* wbase = smem + 16 + (size_t)warp * a.warp_floats;
    float* praw = wbase;                             // the parent, zero-padded
    float* rp = wbase + D;                           // rotated and scaled parent
    float* x = wbase + 2 * D;                        // [rows][D+1]
    uint8_t* u8 = reinterpret_cast<uint8_t*>(x + (size_t)rows * xs);   // [rows][D+4]  (B > 1)
    // large D: the two tiles live in global memory instead (L2), as [coordinate][lane] -- one 128-byte line per
    // coordinate and warp access -- so that shared memory no longer caps the pairs in flight per SM
    const bool gtile = a.tile_x != nullptr;
    const uint32_t gw = blockIdx.x * (blockDim.x >> 5) + warp;
    if (gtile) { x = a.tile_x + (size_t)gw * D * 32; u8 = a.tile_u + (size_t)gw * D * 32; }
    const uint32_t cs = gtile ? 32u : 1u;                              // between coordinates of one row (both tiles)
    const uint32_t xrow = gtile ? 1u : xs, urow = gtile ? 1u : us;     // between rows

    if (threadIdx.x <= (unsigned)K_INT)
        ctab[threadIdx.x] = __fdiv_rn(__fsub_rn(__fmul_rn(2.0f, (float)threadIdx.x), (float)K_INT), (float)K_INT);
    __syncthreads();

    for (uint64_t p = gw; p < a.n_parents; p += a.total_warps) {
    const uint32_t pid = a.parent_ids ? a.parent_ids[p] : (uint32_t)p;
    const bool parent_ok = pid < a.n_vectors;

    // the parent: raw copy for the offsets, and rotate_raw_vector(parent) for ip_cp
    {
        const float* src = a.vectors + (size_t)(parent_ok ? pid : 0) * a.row_stride;
        for (uint32_t i = lane; i < D; i += 32) {
            const float v = (parent_ok && i < dim) ? src[i] : 0.0f;
            praw[i] = v;
            rp[i] = v;
        }
        __syncwarp();
        for (int layer = 0; layer < 3; ++layer) {
            const float* sg = a.signs + (size_t)layer * D;
            for (uint32_t i = lane; i < D; i += 32) rp[i] = __fmul_rn(rp[i], sg[i]);
            __syncwarp();
            for (uint32_t h = 1; h < D; h <<= 1) {
                for (uint32_t q = lane; q < D / 2; q += 32) {
                    const uint32_t i = ((q & ~(h - 1)) << 1) | (q & (h - 1));
                    const float va = rp[i], vb = rp[i + h];
                    rp[i] = __fadd_rn(va, vb);
                    rp[i + h] = (h < 8) ? __fsub_rn(vb, va) : __fsub_rn(va, vb);
                }
                __syncwarp();
            }
        }
        for (uint32_t i = lane; i < D; i += 32) rp[i] = __fmul_rn(rp[i], a.norm_factor);
        __syncwarp();
    }

    const uint32_t code_bytes = B * (D / 8);
    uint32_t count = 0;   // one past the last occupied slot
    for (uint32_t v0 = 0; v0 < kR; v0 += rows) {
        // offsets nb - parent of `rows` neighbours, written cooperatively (coalesced reads)
        const uint32_t my_slot = v0 + lane;
        const uint32_t my_nid = (lane < rows && parent_ok) ? a.nbr_ids[p * kR + my_slot] : kInvalid;
        const bool valid = my_nid < a.n_vectors;     // kInvalid marks an empty slot (fastscan_layout.hpp:51-92)
        for (uint32_t r = 0; r < rows; ++r) {
            const uint32_t nid = __shfl_sync(kFull, my_nid, r);
            if (nid >= a.n_vectors) continue;
            const float* src = a.vectors + (size_t)nid * a.row_stride;
            float* xr = x + (size_t)r * xrow;
            for (uint32_t i = lane; i < D; i += 32) xr[i * cs] = i < dim ? __fsub_rn(src[i], praw[i]) : 0.0f;
        }
        __syncwarp();

        float* xr = x + (size_t)lane * xrow;
        uint8_t* ur = u8 + (size_t)lane * urow;
        float nop = 0.0f, ip_qo = 0.0f, ip_cp = 0.0f;
        bool live = false;
        if (valid) {
            float nop_sq = 0.0f;                     // nop_sq += d * d: multiply, then add (not fused)
            for (uint32_t i = 0; i < dim; ++i) { const float d = xr[i * cs]; nop_sq = __fadd_rn(nop_sq, __fmul_rn(d, d)); }
            nop = __fsqrt_rn(nop_sq);
            live = !(nop < a.norm_eps);
        }
        if (live) {
            const float inv_nop = __fdiv_rn(1.0f, nop);   // the unit offset (x * inv_nop), then three (diagonal, transform) layers
            for (int layer = 0; layer < 3; ++layer) rotate_row(xr, D, cs, a.signs + (size_t)layer * D, layer == 0, inv_nop);
        }

        uint8_t* out_code = a.codes ? a.codes + ((size_t)p * kR + my_slot) * code_bytes : nullptr;
        // the reference's own block: plane b is packed[sp][slot] = the code byte of dimensions 8sp..8sp+7
        // (FastScanCodeBlock::store, distance/fastscan_layout.hpp:10-49), so the 32 lanes write 32 consecutive bytes
        uint8_t* blk = a.blocks ? a.blocks + (size_t)p * a.block_stride : nullptr;
        uint32_t pop = 0, wpop = 0;
        if (B == 1) {
            // sign bits, |rotated|_1 and the signed sum of the rotated parent, coordinate by coordinate
            float l1 = 0.0f, ip = 0.0f;
            if (lane < rows) {
                for (uint32_t j = 0; j < D / 8; ++j) {
                    uint32_t byte = 0;
                    if (live) {
#pragma unroll
                        for (uint32_t t = 0; t < 8; ++t) {
                            const uint32_t i = 8 * j + t;
                            const float r = __fmul_rn(xr[i * cs], a.norm_factor);
                            const bool bit = r >= 0.0f;
                            byte |= (bit ? 1u : 0u) << t;
                            l1 = __fadd_rn(l1, fabsf(r));
                            ip = bit ? __fadd_rn(ip, rp[i]) : __fsub_rn(ip, rp[i]);
                        }
                    }
                    pop += __popc(byte);
                    if (out_code) out_code[j] = (uint8_t)byte;
                    if (blk) blk[(size_t)j * kR + my_slot] = (uint8_t)byte;
                }
                ip_qo = __fmul_rn(l1, a.inv_sqrt_d);
                ip_cp = __fmul_rn(ip, a.inv_sqrt_d);
            }
        } else {
            if (live) {
                const float K = (float)K_INT;
                float mn, mx;
                {
                    const float v = __fmul_rn(xr[0], a.norm_factor);
                    xr[0] = v; mn = v; mx = v;
                }
                for (uint32_t i = 1; i < D; ++i) {
                    const float v = __fmul_rn(xr[i * cs], a.norm_factor);
                    xr[i * cs] = v;
                    if (v < mn) mn = v;
                    if (v > mx) mx = v;
                }
                float delta = __fdiv_rn(__fsub_rn(mx, mn), K);
                if (delta < a.coord_eps) delta = a.coord_eps;
                const float inv_delta = __fdiv_rn(1.0f, delta);
                // initial rounding; dot_co and norm_c_sq start as plain multiply-then-add sums
                float dot_co = 0.0f, norm_c_sq = 0.0f;
                for (uint32_t i = 0; i < D; ++i) {
                    const float xi = xr[i * cs];
                    int u = __float2int_rz(__fmaf_rn(__fsub_rn(xi, mn), inv_delta, 0.5f));
                    u = u < 0 ? 0 : (u > K_INT ? K_INT : u);
                    ur[i * cs] = (uint8_t)u;
                    const float c = ctab[u];
                    dot_co = __fadd_rn(__fmul_rn(c, xi), dot_co);
                    norm_c_sq = __fadd_rn(__fmul_rn(c, c), norm_c_sq);
                }
                // coordinate descent on cos^2(x, c) = dot_co^2 / norm_c_sq: at most 10 sweeps (rabitq_encoder.hpp:400-455)
                float prev_cos_sq = 0.0f;
                for (int iter = 0; iter < 10; ++iter) {
                    bool changed = false;
                    for (uint32_t i = 0; i < D; ++i) {
                        const int old_u = ur[i * cs];
                        const float xi = xr[i * cs];
                        const float old_c = ctab[old_u];
                        const float dot_without = __fmaf_rn(-old_c, xi, dot_co);
                        const float norm_without = __fmaf_rn(-old_c, old_c, norm_c_sq);
                        int best_u = old_u;
                        float best_dot = dot_co, best_norm = norm_c_sq;
                        if (B >= 4) {                // neighbours of the current value only
#pragma unroll
                            for (int s = -1; s <= 1; s += 2) {
                                const int ut = old_u + s;
                                if (ut < 0 || ut > K_INT) continue;
                                const float c = ctab[ut];
                                const float nd = __fmaf_rn(c, xi, dot_without), nn = __fmaf_rn(c, c, norm_without);
                                if (__fmul_rn(__fmul_rn(nd, nd), best_norm) > __fmul_rn(__fmul_rn(best_dot, best_dot), nn)) {
                                    best_u = ut; best_dot = nd; best_norm = nn;
                                }
                            }
                        } else {                     // every other value
#pragma unroll
                            for (int ut = 0; ut <= K_INT; ++ut) {
                                if (ut == old_u) continue;
                                const float c = ctab[ut];
                                const float nd = __fmaf_rn(c, xi, dot_without), nn = __fmaf_rn(c, c, norm_without);
                                if (__fmul_rn(__fmul_rn(nd, nd), best_norm) > __fmul_rn(__fmul_rn(best_dot, best_dot), nn)) {
                                    best_u = ut; best_dot = nd; best_norm = nn;
                                }
                            }
                        }
                        if (best_u != old_u) {       // the accepted trial's sums are the new running sums
                            dot_co = best_dot;
                            norm_c_sq = best_norm;
                            ur[i * cs] = (uint8_t)best_u;
                            changed = true;
                        }
                    }
                    if (!changed) break;
                    const float cos_sq = norm_c_sq > 0.0f ? __fdiv_rn(__fmul_rn(dot_co, dot_co), norm_c_sq) : 0.0f;
                    if (iter > 0 && __fsub_rn(cos_sq, prev_cos_sq) < 1e-4f) break;   // constants::kCaqEarlyExitTol
                    prev_cos_sq = cos_sq;
                }
            }
            // bit planes, MSB first (NbitCodeStorage::set_value, core/codes.hpp:107-116), and the two inner products
            if (lane < rows) {
                float sq = 0.0f, sc = 0.0f;
                for (uint32_t j = 0; j < D / 8; ++j) {
                    uint32_t bytes[B];
#pragma unroll
                    for (int b = 0; b < B; ++b) bytes[b] = 0;
                    if (live) {
#pragma unroll
                        for (uint32_t t = 0; t < 8; ++t) {
                            const uint32_t i = 8 * j + t;
                            const uint32_t u = ur[i * cs];
                            const float c = ctab[u];
                            sq = __fmaf_rn(c, xr[i * cs], sq);
                            sc = __fmaf_rn(c, rp[i], sc);
                            wpop += u;
#pragma unroll
                            for (int b = 0; b < B; ++b) bytes[b] |= ((u >> (B - 1 - b)) & 1u) << t;
                        }
                    }
                    pop += __popc(bytes[0]);
#pragma unroll
                    for (int b = 0; b < B; ++b) {
                        if (out_code) out_code[(size_t)b * (D / 8) + j] = (uint8_t)bytes[b];
                        if (blk) blk[(size_t)b * 4 * D + (size_t)j * kR + my_slot] = (uint8_t)bytes[b];
                    }
                }
                ip_qo = __fmul_rn(sq, a.inv_sqrt_d);
                ip_cp = __fmul_rn(sc, a.inv_sqrt_d);
            }
        }
        if (lane < rows) {
            if (!live) { ip_qo = 0.0f; ip_cp = 0.0f; }
            if (a.aux) {
                float* o = a.aux + ((size_t)p * kR + my_slot) * 3;
                o[0] = nop; o[1] = ip_qo; o[2] = ip_cp;
            }
            if (blk) {   // set_neighbor (fastscan_layout.hpp:73-87, 134-150)
                reinterpret_cast<float*>(blk + a.nop_off)[my_slot] = nop;
                reinterpret_cast<float*>(blk + a.nop_off + 128)[my_slot] = ip_qo;
                reinterpret_cast<float*>(blk + a.nop_off + 256)[my_slot] = ip_cp;
                reinterpret_cast<uint16_t*>(blk + a.nop_off + 384)[my_slot] = (uint16_t)pop;
                if (B > 1) reinterpret_cast<uint16_t*>(blk + a.nop_off + 448)[my_slot] = (uint16_t)wpop;
                reinterpret_cast<uint32_t*>(blk + a.ids_off)[my_slot] = valid ? my_nid : kInvalid;
            }
        }
        const unsigned vm = __ballot_sync(kFull, valid);
        if (vm) count = v0 + 32 - __clz(vm);
        __syncwarp();
    }
    if (lane == 0 && a.blocks) *reinterpret_cast<uint32_t*>(a.blocks + (size_t)p * a.block_stride + a.ids_off + 128) = count;
    }   // parents of this warp
}

}  // namespace

size_t neighbor_codes_warp_bytes(uint32_t D, uint32_t B, uint32_t rows) {
    return sizeof(float) * (2 * (size_t)D + (size_t)rows * (D + 1)) + (B > 1 ? (size_t)rows * (D + 4) : 0);
}

// Launch shape and the encoder's constants for (a.D, B).  global_tile: the per-warp tiles go to a scratch buffer of
// plan->scratch_bytes (the caller allocates it and sets a.tile_x / a.tile_u = tile_x + total_warps * D * 32 floats)
// and a bounded grid of warps loops over the parents; else they live in shared memory, one parent per warp.
cudaError_t neighbor_codes_plan(NeighborCodesArgs& a, uint32_t B, int num_sms, bool global_tile, NeighborCodesPlan* plan) {
    if (B != 1 && B != 2 && B != 4) return cudaErrorInvalidValue;
    constexpr size_t kBudget = 200 * 1024;
    uint32_t rows = kR, warps;
    size_t per_warp;
    if (global_tile) {
        per_warp = sizeof(float) * 2 * (size_t)a.D;
        warps = 8;
        const size_t cta = 64 + warps * per_warp;
        size_t per_sm = kBudget / cta;
        per_sm = per_sm < 1 ? 1 : (per_sm > 3 ? 3 : per_sm);   // 80 registers x 256 threads: three CTAs per SM
        const uint64_t want = (a.n_parents + warps - 1) / warps, cap = (uint64_t)num_sms * per_sm;
        plan->grid = (unsigned)(want < cap ? want : cap);
        plan->scratch_bytes = (size_t)plan->grid * warps * a.D * 32 * (sizeof(float) + 1);
    } else {
        while (rows > 1 && neighbor_codes_warp_bytes(a.D, B, rows) + 64 > kBudget) rows >>= 1;
        per_warp = (neighbor_codes_warp_bytes(a.D, B, rows) + 15) & ~(size_t)15;
        if (per_warp + 64 > kBudget) return cudaErrorInvalidValue;
        warps = (uint32_t)((kBudget - 64) / per_warp);
        warps = warps > 8 ? 8 : warps;
        if ((uint64_t)warps > a.n_parents) warps = (uint32_t)a.n_parents;
        plan->grid = (unsigned)((a.n_parents + warps - 1) / warps);
        plan->scratch_bytes = 0;
    }
    plan->warps = warps;
    plan->smem_bytes = 64 + (size_t)warps * per_warp;
    a.rows = rows;
    a.warp_floats = (uint32_t)(per_warp / sizeof(float));
    a.total_warps = plan->grid * warps;
    a.tile_x = nullptr;
    a.tile_u = nullptr;
    const float Df = (float)a.D;
    // constants of RaBitQEncoderBase's constructor (encoder/rabitq_encoder.hpp:37-39) and core/constants.hpp, host floats
    a.norm_factor = 1.0f / (Df * sqrtf(Df));
    a.inv_sqrt_d = 1.0f / sqrtf(Df);
    a.norm_eps = 1e-8f / Df;
    // FastScanNeighborBlock / NbitFastScanNeighborBlock (fastscan_layout.hpp:51-92, 114-155)
    a.nop_off = 4 * a.D * B;
    a.ids_off = a.nop_off + 384 + 64 * (B > 1 ? 2 : 1);
    a.coord_eps = 1e-10f / Df;
    return cudaSuccess;
}

#ifndef CPB_HOST_EMULATION
cudaError_t launch_neighbor_codes(const NeighborCodesArgs& a, uint32_t B, const NeighborCodesPlan& plan, cudaStream_t stream) {
    if (a.n_parents == 0) return cudaSuccess;
    auto go = [&](auto kernel) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem_bytes);
        if (e != cudaSuccess) return e;
        kernel<<<plan.grid, plan.warps * 32, plan.smem_bytes, stream>>>(a);
        return cudaGetLastError();
    };
    switch (B) {
        case 1: return go(neighbor_codes_kernel<1>);
        case 2: return go(neighbor_codes_kernel<2>);
        default: return go(neighbor_codes_kernel<4>);
    }
}
#endif

}  // namespace cpb
