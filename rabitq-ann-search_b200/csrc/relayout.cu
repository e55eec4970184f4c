// Index hand-off: turns the reference's in-memory / save-file records into this library's HBM
// layout (device_index.h).  The records are copied to the device as they are and re-laid out
// there, so the host never touches the 1-3 KB per vertex more than once.
//
// Source layout (reference): VertexSearchData<D,32,B> = { code ; neighbour block },
// neighbour block = planes[B] of packed[D/8][32] (byte sp of slot v = dims 8sp..8sp+7, LSB first;
// distance/fastscan_layout.hpp:10-49), then nop, ip_qo, ip_cp (f32[32]), popcounts,
// [weighted_popcounts] (u16[32]), neighbor_ids (u32[32]), count (:51-92,114-155).
#include "device_math.cuh"
#include "kernels.h"

namespace cpb {

// one CTA (128 threads) per vertex
__global__ void __launch_bounds__(128) relayout_blocks_kernel(const DevIndex ix, const uint8_t* __restrict__ rec,
                                                              uint64_t rec_size, uint32_t nb_off, uint64_t first,
                                                              uint32_t count, uint32_t* __restrict__ problems) {
    const uint32_t v = blockIdx.x;
    if (v >= count) return;
    const uint32_t D = ix.D, B = ix.B, nch = ix.nch;
    const uint8_t* src = rec + (size_t)v * rec_size;
    const uint8_t* nb = src + nb_off;
    uint8_t* dst = const_cast<uint8_t*>(ix.blocks) + (first + v) * (size_t)ix.block_stride;
    // code planes: dst[((b*nch + c)*32 + slot)*16 + t] = packed_b[16c + t][slot]
    const uint32_t plane_bytes = 4 * D, nsp = D / 8;
    const uint32_t total = B * nch * 32 * 16;
    for (uint32_t o = threadIdx.x; o < total; o += blockDim.x) {
        const uint32_t t = o & 15, slot = (o >> 4) & 31, bc = o >> 9, c = bc % nch, b = bc / nch;
        const uint32_t sp = 16 * c + t;
        dst[o] = sp < nsp ? nb[(size_t)b * plane_bytes + (size_t)sp * 32 + slot] : (uint8_t)0;
    }
    const uint32_t o0 = plane_bytes * B;
    const bool nbit = B > 1;
    const uint8_t* s_nop = nb + o0;
    const uint8_t* s_pop = nb + o0 + 384;
    const uint8_t* s_wpop = nb + o0 + 448;
    const uint8_t* s_ids = nb + o0 + (nbit ? 512 : 448);
    uint8_t* aux = dst + ix.aux_off;
    if (threadIdx.x < 32) {
        const uint32_t l = threadIdx.x;
        reinterpret_cast<uint32_t*>(aux)[l] = reinterpret_cast<const uint32_t*>(s_ids)[l];
        reinterpret_cast<uint32_t*>(aux + 128)[l] = reinterpret_cast<const uint32_t*>(s_nop)[l];
        reinterpret_cast<uint32_t*>(aux + 256)[l] = reinterpret_cast<const uint32_t*>(s_nop + 128)[l];
        reinterpret_cast<uint32_t*>(aux + 384)[l] = reinterpret_cast<const uint32_t*>(s_nop + 256)[l];
        const uint32_t pop = reinterpret_cast<const uint16_t*>(s_pop)[l];
        const uint32_t wpop = nbit ? reinterpret_cast<const uint16_t*>(s_wpop)[l] : 0u;
        reinterpret_cast<uint32_t*>(aux + 512)[l] = pop | (wpop << 16);
        const uint32_t cnt = reinterpret_cast<const uint32_t*>(s_ids + 128)[0];
        if (l == 0) {
            reinterpret_cast<uint32_t*>(aux + 640)[0] = cnt < 32 ? cnt : 32;
            reinterpret_cast<float*>(aux + 644)[0] = ix.norm_sq[first + v];   // the vertex's own |x|^2 rides along
        }
        // sanity of the graph: ids in range, and whether any id repeats inside the block
        const uint32_t id = reinterpret_cast<const uint32_t*>(s_ids)[l];
        const bool valid = l < cnt;
        const unsigned peers = __match_any_sync(kFull, valid ? id : (kInvalid - l));
        const bool dup = __any_sync(kFull, valid && (peers & (peers - 1)) != 0);
        const bool oob = __any_sync(kFull, valid && id >= ix.n);
        if (l == 0) { if (dup) atomicAdd(problems, 1u); if (oob) atomicAdd(problems + 1, 1u); }
    }
    // per-vertex 1-bit code (RaBitQCode<D>: signs @0, nop, ip_qo after the 64-B aligned sign words)
    if (B == 1 && ix.flat_codes && threadIdx.x >= 32 && threadIdx.x < 64) {
        const uint32_t l = threadIdx.x - 32;
        const uint32_t words = nch * 4;
        uint32_t* fc = const_cast<uint32_t*>(ix.flat_codes) + (first + v) * (size_t)words;
        uint32_t pc = 0;
        for (uint32_t wd = l; wd < words; wd += 32) {
            uint32_t x = 0;
            if (wd * 4 < D / 8) {
                for (uint32_t t = 0; t < 4 && wd * 4 + t < D / 8; ++t) x |= (uint32_t)src[wd * 4 + t] << (8 * t);
            }
            fc[wd] = x;
            pc += __popc(x);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) pc += __shfl_xor_sync(kFull, pc, o);
        if (l == 0) {
            const uint32_t storage = ((8 * ((D + 63) / 64)) + 63) / 64 * 64;
            const_cast<float*>(ix.flat_nop)[first + v] = *reinterpret_cast<const float*>(src + storage);
            const_cast<float*>(ix.flat_ipqo)[first + v] = *reinterpret_cast<const float*>(src + storage + 4);
            const_cast<uint16_t*>(ix.flat_pop)[first + v] = (uint16_t)pc;
        }
    }
}

__global__ void relayout_raw_kernel(const DevIndex ix, const float* __restrict__ raw, uint64_t first, uint32_t count) {
    const uint32_t D = ix.D, T = ix.T;
    const size_t total = (size_t)count * D;
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (size_t)gridDim.x * blockDim.x) {
        const size_t v = o / D;
        const uint32_t i = (uint32_t)(o % D);
        const uint32_t l = i & 7u, t = i >> 3;
        const_cast<float*>(ix.rawT)[(first + v) * D + (size_t)l * T + raw_chunk_pos(l, t, T)] = raw[o];
    }
}

#ifndef CPB_HOST_EMULATION   // tests/native/ compiles the kernels above for the host (no PTX)
cudaError_t launch_relayout_blocks(const DevIndex& ix, const uint8_t* d_records, uint64_t rec_size, uint32_t nb_off,
                                   uint64_t first, uint32_t count, uint32_t* d_problems, cudaStream_t stream) {
    if (count == 0) return cudaSuccess;
    relayout_blocks_kernel<<<count, 128, 0, stream>>>(ix, d_records, rec_size, nb_off, first, count, d_problems);
    return cudaGetLastError();
}

cudaError_t launch_relayout_raw(const DevIndex& ix, const float* d_raw, uint64_t first, uint32_t count,
                                cudaStream_t stream) {
    if (count == 0) return cudaSuccess;
    const size_t total = (size_t)count * ix.D;
    const int grid = (int)((total + 255) / 256 < 65535 ? (total + 255) / 256 : 65535);
    relayout_raw_kernel<<<grid, 256, 0, stream>>>(ix, d_raw, first, count);
    return cudaGetLastError();
}

#endif

}  // namespace cpb
