// N4 -- calibration.cu -- with the index hand-off kernels of relayout.cu, compiled for the host over cuda_emul.h: a
// finalized index in the reference's record layout in, re-laid out by the product's kernels, the sample loop run by the
// product's kernel.  The prepared queries come from the K1 harness.  Built and called by tests/test_kernels_emulated.py;
// never part of the product.
#include "cuda_emul.h"

namespace cpb { alignas(128) uint8_t smem_raw[224 * 1024]; }

#define CPB_HOST_EMULATION 1
#include "../../rabitq-ann-search_b200/csrc/relayout.cu"
#include "../../rabitq-ann-search_b200/csrc/calibration.cu"

extern "C" int emul_calibration_samples(uint32_t dim, uint32_t bits, const uint8_t* records, uint64_t rec_size, uint32_t nb_off, uint64_t n,
                                        const float* raw, const float* queries, const uint32_t* start_ids, uint64_t ns, const float* qT,
                                        const uint32_t* uplanes, const float* coeffs, uint32_t* parent, float* nn_dist_sq, float* dist_qp_sq,
                                        float* nop, float* ip_corrected, float* ip_qo_denom, float* true_ip, uint32_t* neighbor) {
    using namespace cpb;
    DevIndex ix{};
    uint32_t D = 16;
    while (D < dim) D <<= 1;
    ix.D = D; ix.B = bits; ix.dim = dim; ix.nch = (D > 128 ? D : 128) / 128; ix.T = D / 8; ix.n = n;
    ix.aux_off = ix.B * ix.nch * 512;
    ix.block_stride = (ix.aux_off + 644 + 127) / 128 * 128;
    std::vector<uint8_t> dev((size_t)n * ix.block_stride + 256, 0xEE);
    uint8_t* base = dev.data() + (128 - reinterpret_cast<uintptr_t>(dev.data()) % 128) % 128;
    std::vector<float> rawT((size_t)n * D + 64, -7.0f);
    float* rawT_al = rawT.data() + (16 - (reinterpret_cast<uintptr_t>(rawT.data()) / 4) % 16) % 16;
    std::vector<float> norm_sq(n, 0.0f);
    ix.blocks = base; ix.rawT = rawT_al; ix.norm_sq = norm_sq.data();
    uint32_t problems[2] = {0, 0};
    auto rl_blocks = [&](int) { relayout_blocks_kernel(ix, records, rec_size, nb_off, 0, (uint32_t)n, problems); };
    cuda_emul::launch(rl_blocks, (unsigned)n, 128, smem_raw, 0, 0);
    auto rl_raw = [&](int) { relayout_raw_kernel(ix, raw, 0, (uint32_t)n); };
    cuda_emul::launch(rl_raw, 2, 256, smem_raw, 0, 0);
    if (problems[1]) return 4;

    CalibrationArgs a{};
    a.queries = queries; a.start_ids = start_ids; a.ns = ns; a.qT = qT; a.uplanes = uplanes; a.coeffs = coeffs;
    a.parent = parent; a.nn_dist_sq = nn_dist_sq; a.dist_qp_sq = dist_qp_sq; a.nop = nop; a.ip_corrected = ip_corrected;
    a.ip_qo_denom = ip_qo_denom; a.true_ip = true_ip; a.neighbor = neighbor;
    const size_t per_warp = ((size_t)8 * (ix.T + 4) * 4 + (size_t)2 * ix.D * 4 + (size_t)ix.nch * 64 + 15) & ~(size_t)15;
    const size_t smem = per_warp * kCalWarps;
    if (smem > sizeof(smem_raw)) return 2;
    auto kern = [&](int) {
        if (bits == 1) calibration_samples_kernel<1>(ix, a);
        else if (bits == 2) calibration_samples_kernel<2>(ix, a);
        else calibration_samples_kernel<4>(ix, a);
    };
    cuda_emul::launch(kern, (unsigned)((ns + kCalWarps - 1) / kCalWarps), kCalWarps * 32, smem_raw, smem, 0);
    return 0;
}
