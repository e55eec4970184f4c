// K3 (+ fused K4) -- search.cu -- with the index hand-off kernels of relayout.cu, compiled for the host over cuda_emul.h:
// a finalized index in the reference's record layout in, re-laid out by the product's kernels, searched by the product's
// search kernel (the bulk-copy/mbarrier helpers have a host form in the source).  The prepared queries come from the K1
// harness.  Built and called by tests/test_kernels_emulated.py; never part of the product.
#include "cuda_emul.h"

#include <algorithm>

namespace cpb { alignas(128) uint8_t smem_raw[224 * 1024]; }

#define CPB_HOST_EMULATION 1
#include "../../rabitq-ann-search_b200/csrc/relayout.cu"
#include "../../rabitq-ann-search_b200/csrc/search.cu"

namespace {
template <typename T>
T rd(const uint8_t* p, size_t off) { T v; std::memcpy(&v, p + off, sizeof(T)); return v; }
}  // namespace

// records: n VertexSearchData records (rec_size apart, neighbour block at nb_off); raw [n][D]; calib: the 248-byte
// CalibrationSnapshot; levels: for level L = 1..n_levels the five arrays node, offs, nbr_node, nbr_slot, down at
// level_arrays[5 (L-1) ..] and its size at level_sizes[L-1] (the slot-addressed CSR cphnsw_b200_upload builds).
extern "C" int emul_search(uint32_t dim, uint32_t bits, const uint8_t* records, uint64_t rec_size, uint32_t nb_off, uint64_t n,
                           const float* raw, const float* norm_sq, const uint8_t* calib, int32_t max_level, uint32_t entry_point,
                           uint32_t entry_slot, uint32_t graph_entry_point, uint32_t n_levels, const uint32_t* const* level_arrays,
                           const uint32_t* level_sizes, const float* qT, const uint32_t* uplanes, const float* coeffs, uint32_t nq,
                           uint32_t k_user, int64_t* ids, float* dists, int warps, int ctas, uint32_t beam_capacity,
                           unsigned long long* stats_out /* 11 counters, may be NULL */, uint32_t* overflowed) {
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
    ix.blocks = base; ix.rawT = rawT_al; ix.norm_sq = norm_sq;
    ix.calib.affine_a = rd<float>(calib, 0); ix.calib.affine_b = rd<float>(calib, 4); ix.calib.ip_qo_floor = rd<float>(calib, 8);
    ix.calib.gamma_max = rd<float>(calib, 84); ix.calib.gamma_beta = rd<float>(calib, 88);
    ix.calib.gamma_warmup = rd<uint64_t>(calib, 96);
    for (int i = 0; i < 32; ++i) ix.calib.slack[i] = rd<float>(calib, 108 + 4 * i);
    ix.calib.num_slack = rd<int32_t>(calib, 236);
    ix.calib.gamma = rd<float>(calib, 240);
    ix.max_level = max_level; ix.entry_point = entry_point; ix.entry_slot = entry_slot; ix.graph_entry_point = graph_entry_point;
    ix.n_levels = n_levels;
    for (uint32_t L = 0; L < n_levels; ++L) {
        Level& lv = ix.levels[L];
        lv.node = level_arrays[5 * L]; lv.offs = level_arrays[5 * L + 1]; lv.nbr_node = level_arrays[5 * L + 2];
        lv.nbr_slot = level_arrays[5 * L + 3]; lv.down = level_arrays[5 * L + 4]; lv.size = level_sizes[L];
    }
    uint32_t problems[2] = {0, 0};
    auto rl_blocks = [&](int) { relayout_blocks_kernel(ix, records, rec_size, nb_off, 0, (uint32_t)n, problems); };
    cuda_emul::launch(rl_blocks, (unsigned)n, 128, smem_raw, 0, 0);
    auto rl_raw = [&](int) { relayout_raw_kernel(ix, raw, 0, (uint32_t)n); };
    cuda_emul::launch(rl_raw, 2, 256, smem_raw, 0, 0);
    ix.dup_neighbors = problems[0];
    if (problems[1]) return 4;

    // run_search's argument layout (capi.cu)
    const uint32_t k = std::max<uint32_t>(k_user, 1);
    while (warps > 1 && search_smem_per_warp(ix, k, stats_out != nullptr) * warps > sizeof(smem_raw)) --warps;
    const size_t smem = search_smem_per_warp(ix, k, stats_out != nullptr) * warps;
    if (smem > sizeof(smem_raw)) return 2;
    const int need = (int)((nq + warps - 1) / warps);
    if (ctas > need) ctas = need;
    SearchArgs a{};
    a.warp_smem = (uint32_t)search_smem_per_warp(ix, k, stats_out != nullptr);
    a.nq = nq; a.query_list = nullptr; a.k = k; a.kout = k_user; a.ids = ids; a.dists = dists;
    a.qT = qT; a.uplanes = uplanes; a.coeffs = coeffs; a.entry_out = nullptr;
    a.bitmap_words = (uint32_t)(((n + 31) / 32 + 31) & ~(uint64_t)31);
    const uint32_t cap = beam_capacity ? beam_capacity : (uint32_t)n + 1;
    size_t off = 0;
    a.heap_off = off; off += ((size_t)(cap + 2) * 16 + 127) & ~(size_t)127;
    a.nn_off = off; if (k > 128) off += ((size_t)k * 8 + 127) & ~(size_t)127;
    a.slot_stride = off; a.beam_capacity = cap;
    const size_t slots = (size_t)ctas * warps;
    std::vector<uint8_t> scratch(slots * a.slot_stride + 256, 0xEE);
    std::vector<uint32_t> bitmaps(slots * (size_t)a.bitmap_words, 0u), over(nq + 1, 0u);
    uint32_t counters[4] = {0, 0, 0, 0};
    Stats st{};
    a.scratch = scratch.data() + (128 - reinterpret_cast<uintptr_t>(scratch.data()) % 128) % 128;
    a.bitmaps = bitmaps.data(); a.overflow_list = over.data(); a.counters = counters; a.stats = &st;
    const bool stats = stats_out != nullptr;
    // the same selection as the launcher's pick_kernel
    auto kern = [&](int) {
#define CPB_EMUL_PICK(BITS, ST, DTV) search_kernel<BITS, ST, DTV>(ix, a)
#define CPB_EMUL_BITS(ST, DTV) do { if (bits == 1) CPB_EMUL_PICK(1, ST, DTV); else if (bits == 2) CPB_EMUL_PICK(2, ST, DTV); else CPB_EMUL_PICK(4, ST, DTV); } while (0)
        if (D == 128) {
            if (stats) CPB_EMUL_BITS(true, 128);
            else CPB_EMUL_BITS(false, 128);
        } else {
            if (stats) CPB_EMUL_BITS(true, 0);
            else CPB_EMUL_BITS(false, 0);
        }
#undef CPB_EMUL_BITS
#undef CPB_EMUL_PICK
    };
    cuda_emul::launch(kern, (unsigned)ctas, warps * 32, smem_raw, smem, 0);
    if (overflowed) *overflowed = counters[1];
    if (stats_out) std::memcpy(stats_out, &st, sizeof(Stats));
    for (uint32_t w : bitmaps) if (w) return 5;   // every slot must leave its bitmap clean for the next query
    return 0;
}

// K4 primitive (exact_l2_kernel): distances of listed ids, raw vectors re-laid out by relayout_raw_kernel
extern "C" int emul_exact_l2(uint32_t dim, uint64_t n, const float* raw, const float* norm_sq, const float* qT, const float* coeffs,
                             uint32_t nq, const uint32_t* ids, uint32_t m, float* out) {
    using namespace cpb;
    DevIndex ix{};
    uint32_t D = 16;
    while (D < dim) D <<= 1;
    ix.D = D; ix.B = 1; ix.dim = dim; ix.nch = (D > 128 ? D : 128) / 128; ix.T = D / 8; ix.n = n;
    std::vector<float> rawT((size_t)n * D + 64, -7.0f);
    ix.rawT = rawT.data() + (16 - (reinterpret_cast<uintptr_t>(rawT.data()) / 4) % 16) % 16;
    ix.norm_sq = norm_sq;
    auto rl_raw = [&](int) { relayout_raw_kernel(ix, raw, 0, (uint32_t)n); };
    cuda_emul::launch(rl_raw, 2, 256, smem_raw, 0, 0);
    const int warps = 4;
    const size_t smem = (size_t)warps * 8 * (ix.T + 4) * 4;
    auto kern = [&](int) { exact_l2_kernel(ix, qT, coeffs, nq, ids, m, out); };
    cuda_emul::launch(kern, (nq + warps - 1) / warps, warps * 32, smem_raw, smem, 0);
    return 0;
}

// slot_planes_by_warp against the lane-per-slot evaluation it replaces for a few slots: one warp, a block's code planes
// ([plane][chunk][slot] uint4) and the query's bit-planes ([t][chunk] uint4) as given.  Out, per lane: plane 1's sum and
// sum_{b >= 1} plane_b << (B-1-b) both ways (the warp's way only in the lanes named by `slots`).
extern "C" int emul_slot_planes(uint32_t bits, uint32_t nch, const uint8_t* planes, const uint32_t* uq, uint32_t slots,
                                uint32_t* p1_warp, uint32_t* rest_warp, uint32_t* p1_lane, uint32_t* rest_lane) {
    using namespace cpb;
    if (bits != 2 && bits != 4) return 1;
    auto kern = [&](int) {
        const uint32_t lane = threadIdx.x & 31;
        const uint4* u4 = reinterpret_cast<const uint4*>(uq);
        uint32_t p1 = 0, rest = 0, q1 = 0, qr = 0;
        if (bits == 4) slot_planes_by_warp<4>(planes, nch, u4, lane, slots, p1, rest);
        else slot_planes_by_warp<2>(planes, nch, u4, lane, slots, p1, rest);
        for (uint32_t b = 1; b < bits; ++b) {
            const uint32_t s = plane_sum_one<true>(reinterpret_cast<const uint4*>(planes), b, nch, lane, u4);
            if (b == 1) q1 = s;
            qr += s << (bits - 1 - b);
        }
        p1_warp[lane] = p1; rest_warp[lane] = rest; p1_lane[lane] = q1; rest_lane[lane] = qr;
    };
    cuda_emul::launch(kern, 1, 32, smem_raw, 0, 0);
    return 0;
}
