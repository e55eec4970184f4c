// The index hand-off (relayout.cu) and K2 (fastscan_blocks.cu) compiled for the host over cuda_emul.h: reference
// neighbour blocks in, re-laid out by the product's own kernel, then estimated by the product's own kernel (the
// bulk-copy/mbarrier helpers have a host form in the source).  Built and called by tests/test_kernels_emulated.py;
// never part of the product.
#include "cuda_emul.h"

namespace cpb { alignas(128) uint8_t smem_raw[224 * 1024]; }

#define CPB_HOST_EMULATION 1
#include "../../rabitq-ann-search_b200/csrc/relayout.cu"
#include "../../rabitq-ann-search_b200/csrc/fastscan_blocks.cu"

// blocks: n reference neighbour blocks, block_bytes apart; calib: affine_a, affine_b, ip_qo_floor, then num_slack levels.
// Outputs [nblocks][32]; any may be NULL.  lean != 0 asks for the streaming specialisation (needs the arguments it is for).
extern "C" int emul_fastscan_blocks(uint32_t dim, uint32_t bits, const uint8_t* blocks, uint64_t block_bytes, uint64_t n,
                                    const float* calib, int num_slack, const uint32_t* uplanes, const float* coeffs, uint32_t nq,
                                    const uint32_t* query_of_block, const uint32_t* vertex_ids, uint64_t nblocks,
                                    const float* dqp, const int32_t* slack_level, uint32_t* nbit, uint32_t* msb, uint32_t* msb2,
                                    float* est, float* lower, float* msb_lower, int lean, uint32_t* problems) {
    cpb::DevIndex ix{};
    uint32_t D = 16;
    while (D < dim) D <<= 1;
    ix.D = D; ix.B = bits; ix.dim = dim; ix.nch = (D > 128 ? D : 128) / 128; ix.T = D / 8; ix.n = n;
    ix.aux_off = ix.B * ix.nch * 512;
    ix.block_stride = (ix.aux_off + 644 + 127) / 128 * 128;
    std::vector<uint8_t> dev((size_t)n * ix.block_stride + 128, 0xEE);
    uint8_t* base = dev.data() + (128 - reinterpret_cast<uintptr_t>(dev.data()) % 128) % 128;
    std::vector<float> norm_sq(n, 1.0f);
    ix.blocks = base; ix.norm_sq = norm_sq.data();
    ix.calib.affine_a = calib[0]; ix.calib.affine_b = calib[1]; ix.calib.ip_qo_floor = calib[2];
    ix.calib.num_slack = num_slack;
    for (int i = 0; i < num_slack && i < 32; ++i) ix.calib.slack[i] = calib[3 + i];

    auto relayout = [&](int) { cpb::relayout_blocks_kernel(ix, blocks, block_bytes, 0, 0, (uint32_t)n, problems); };
    cuda_emul::launch(relayout, (unsigned)n, 128, cpb::smem_raw, 0, 0);

    cpb::FastScanArgs a{};
    a.uplanes = uplanes; a.coeffs = coeffs; a.nq = nq; a.query_of_block = query_of_block; a.vertex_ids = vertex_ids;
    a.first_vertex = 0; a.nblocks = nblocks; a.dqp = dqp; a.slack_level = slack_level;
    a.nbit = nbit; a.msb = msb; a.msb2 = msb2; a.est = est; a.lower = lower; a.msb_lower = msb_lower;
    // the launcher's shape (launch_fastscan_blocks), on a 2-SM device so that warps take several blocks each
    const uint32_t copy_bytes = (ix.aux_off + 644u + 15u) & ~15u, stage_bytes = (copy_bytes + 127u) & ~127u, ns = 2;
    const size_t per_warp = (((size_t)ns * stage_bytes + 64 + (size_t)ix.nch * 64) + 127) & ~(size_t)127;
    int warps = (int)((200u * 1024u) / per_warp);
    warps = warps < 1 ? 1 : (warps > cpb::kFsMaxWarps ? cpb::kFsMaxWarps : warps);
    const size_t smem = per_warp * warps;
    if (smem > sizeof(cpb::smem_raw)) return 2;
    const uint64_t want = (nblocks + warps - 1) / warps;
    const unsigned grid = (unsigned)(want < 2 ? want : 2);
    const bool can_lean = !vertex_ids && !query_of_block && !slack_level && est && lower && !nbit && !msb && !msb2 && !msb_lower;
    if (lean && !can_lean) return 3;
    auto k2 = [&](int) {
        if (lean) {
            if (bits == 1) cpb::fastscan_blocks_kernel<1, true>(ix, a, ns, stage_bytes);
            else if (bits == 2) cpb::fastscan_blocks_kernel<2, true>(ix, a, ns, stage_bytes);
            else cpb::fastscan_blocks_kernel<4, true>(ix, a, ns, stage_bytes);
        } else {
            if (bits == 1) cpb::fastscan_blocks_kernel<1, false>(ix, a, ns, stage_bytes);
            else if (bits == 2) cpb::fastscan_blocks_kernel<2, false>(ix, a, ns, stage_bytes);
            else cpb::fastscan_blocks_kernel<4, false>(ix, a, ns, stage_bytes);
        }
    };
    cuda_emul::launch(k2, grid, warps * 32, cpb::smem_raw, smem, 0);
    return 0;
}
