// Result post-processing OUTSIDE the parity path.
//
// The reference's result list holds a vertex once per time it was scored (search/rabitq_search.hpp:133,236,250;
// BoundedMaxHeap does not de-duplicate, :26-35 -- SURVEY F2) and returns internal, BFS-reordered ids
// (graph/rabitq_graph.hpp:208-278 -- SURVEY F1).  search_batch reproduces both bit for bit.  This kernel is the
// opt-in clean-up a user applies afterwards: from each ascending row of k_in (id, distance) pairs keep the
// first occurrence of every id, write the first k_out of them (padded with -1 / FLT_MAX like
// src/bindings.cpp:201-210) and, if a map is given, translate internal ids to the caller's original ids.
#include <float.h>

#include "kernels.h"

namespace cpb {

// one warp per row; a pair survives if no earlier pair of the row has its id
__global__ void __launch_bounds__(128) unique_topk_kernel(const int64_t* __restrict__ ids_in, const float* __restrict__ dists_in,
                                                          uint64_t nq, uint32_t kin, uint32_t kout,
                                                          const uint32_t* __restrict__ id_map, uint64_t map_size,
                                                          int64_t* __restrict__ ids_out, float* __restrict__ dists_out) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t row = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= nq) return;
    const int64_t* in = ids_in + row * kin;
    const float* din = dists_in + row * kin;
    int64_t* out = ids_out + row * kout;
    float* dout = dists_out + row * kout;
    uint32_t nout = 0;
    for (uint32_t base = 0; base < kin && nout < kout; base += 32) {
        const uint32_t j = base + lane;
        const int64_t id = j < kin ? in[j] : -1;
        bool keep = id >= 0;
        for (uint32_t e = 0; keep && e < j; ++e) keep = in[e] != id;   // k_in is small (tens): rows stay in L1
        const unsigned m = __ballot_sync(0xFFFFFFFFu, keep);
        const uint32_t pos = nout + __popc(m & ((1u << lane) - 1u));
        if (keep && pos < kout) {
            out[pos] = (id_map && (uint64_t)id < map_size) ? (int64_t)id_map[id] : id;
            dout[pos] = din[j];
        }
        nout += __popc(m);
    }
    if (nout > kout) nout = kout;
    for (uint32_t p = nout + lane; p < kout; p += 32) { out[p] = -1; dout[p] = FLT_MAX; }
}

cudaError_t launch_unique_topk(const int64_t* ids_in, const float* dists_in, uint64_t nq, uint32_t kin, uint32_t kout,
                               const uint32_t* id_map, uint64_t map_size, int64_t* ids_out, float* dists_out, cudaStream_t stream) {
    if (nq == 0 || kout == 0) return cudaSuccess;
    const uint32_t rows_per_cta = 4;
    unique_topk_kernel<<<(unsigned)((nq + rows_per_cta - 1) / rows_per_cta), rows_per_cta * 32, 0, stream>>>(
        ids_in, dists_in, nq, kin, kout, id_map, map_size, ids_out, dists_out);
    return cudaGetLastError();
}

}  // namespace cpb
