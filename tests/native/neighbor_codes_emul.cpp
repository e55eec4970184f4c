// The N3 kernel (rabitq-ann-search_b200/csrc/neighbor_codes.cu), compiled for the host over cuda_emul.h and exported
// with the argument list of cphnsw_b200_neighbor_codes -- host pointers instead of device pointers.  Built and
// called by tests/test_neighbor_codes_emulated.py; never part of the product.
#include "cuda_emul.h"

namespace cpb { namespace { alignas(16) float smem[64 * 1024]; } }   // 256 KB: more than any plan asks for

#define CPB_HOST_EMULATION 1
#include "../../rabitq-ann-search_b200/csrc/neighbor_codes.cu"

extern "C" int emul_neighbor_codes(uint32_t dim, uint32_t bits, const float* signs, const float* vectors, uint64_t row_stride,
                                   uint64_t n_vectors, const uint32_t* parent_ids, const uint32_t* nbr_ids, uint64_t n_parents,
                                   uint8_t* codes, float* aux, uint8_t* blocks, uint64_t block_stride, uint32_t max_warps,
                                   uint32_t* rows_used) {
    cpb::NeighborCodesArgs a{};
    uint32_t D = 16;
    while (D < dim) D <<= 1;
    a.D = D; a.dim = dim; a.signs = signs; a.vectors = vectors; a.row_stride = row_stride; a.n_vectors = n_vectors;
    a.parent_ids = parent_ids; a.nbr_ids = nbr_ids; a.n_parents = n_parents; a.codes = codes; a.aux = aux;
    a.blocks = blocks; a.block_stride = block_stride;
    cpb::NeighborCodesPlan plan{};
    const bool global_tile = (max_warps >> 16) != 0;       // high half of max_warps: tiles in "global" memory
    max_warps &= 0xFFFF;
    if (cpb::neighbor_codes_plan(a, bits, /*num_sms=*/2, global_tile, &plan) != cudaSuccess) return 1;
    if (plan.smem_bytes > sizeof(cpb::smem)) return 2;
    if (max_warps && plan.warps > max_warps && !global_tile) {
        plan.warps = max_warps;
        plan.grid = (unsigned)((n_parents + max_warps - 1) / max_warps);
        a.total_warps = plan.grid * plan.warps;
    }
    std::vector<uint8_t> scratch(plan.scratch_bytes + 16, 0xFF);
    if (plan.scratch_bytes) {
        a.tile_x = reinterpret_cast<float*>(scratch.data());
        a.tile_u = reinterpret_cast<uint8_t*>(a.tile_x + (size_t)a.total_warps * D * 32);
    }
    if (rows_used) *rows_used = a.rows;
    switch (bits) {
        case 1: cuda_emul::launch(cpb::neighbor_codes_kernel<1>, plan.grid, plan.warps * 32, cpb::smem, plan.smem_bytes, a); break;
        case 2: cuda_emul::launch(cpb::neighbor_codes_kernel<2>, plan.grid, plan.warps * 32, cpb::smem, plan.smem_bytes, a); break;
        default: cuda_emul::launch(cpb::neighbor_codes_kernel<4>, plan.grid, plan.warps * 32, cpb::smem, plan.smem_bytes, a); break;
    }
    return 0;
}
