// K1 (rabitq-ann-search_b200/csrc/query_prep.cu) compiled for the host over cuda_emul.h; host pointers in the
// argument list of cphnsw_b200_prepare_queries.  Built and called by tests/test_kernels_emulated.py; never part of
// the product.
#include "cuda_emul.h"

namespace cpb { alignas(16) float smem[16 * 1024]; }

#define CPB_HOST_EMULATION 1
#include "../../rabitq-ann-search_b200/csrc/query_prep.cu"

extern "C" int emul_prepare_queries(uint32_t dim, const float* signs, const float* centroid, const float* queries, uint32_t nq,
                                    int center, uint8_t* lut, float* coeffs, float* rotated, uint32_t* uplanes, float* qT) {
    cpb::DevIndex ix{};
    uint32_t D = 16;
    while (D < dim) D <<= 1;
    ix.D = D; ix.dim = dim; ix.B = 1; ix.nch = (D > 128 ? D : 128) / 128; ix.T = D / 8;
    ix.signs = signs; ix.centroid = centroid;
    cpb::PrepOut out{};
    out.lut = lut; out.coeffs = coeffs; out.rotated = rotated; out.uplanes = uplanes; out.qT = qT;
    const float Df = (float)D;
    const float norm_factor = 1.0f / (Df * sqrtf(Df)), inv_sqrt_d = 1.0f / sqrtf(Df);
    const size_t smem_bytes = (size_t)cpb::kPrepWarps * ((D + 32) * sizeof(float) + D);
    if (smem_bytes > sizeof(cpb::smem)) return 2;
    const unsigned grid = (nq + cpb::kPrepWarps - 1) / cpb::kPrepWarps;
    auto kernel = [&](int) { cpb::query_prep_kernel(ix, queries, nq, center, norm_factor, inv_sqrt_d, out); };
    cuda_emul::launch(kernel, grid, cpb::kPrepWarps * 32, cpb::smem, smem_bytes, 0);
    return 0;
}
