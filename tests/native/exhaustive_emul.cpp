// K5, popcount form: exhaustive_scan_kernel + exhaustive_select_rerank_kernel (exhaustive.cu; the second also ends the
// two tensor-core forms), and the scan in pieces: candidate mode + exhaustive_select_keys_kernel + merge_candidates_kernel, with the index hand-off of relayout.cu, compiled for the host over cuda_emul.h.  The prepared
// (centred) queries come from the K1 harness.  Built and called by tests/test_kernels_emulated.py; never part of the product.
#include "cuda_emul.h"

#include <float.h>

#include <algorithm>

namespace cpb { alignas(128) uint8_t smem_raw[224 * 1024]; }

#define CPB_HOST_EMULATION 1
#include "../../rabitq-ann-search_b200/csrc/relayout.cu"
#include "../../rabitq-ann-search_b200/csrc/exhaustive.cu"

extern "C" int emul_exhaustive(uint32_t dim, const uint8_t* records, uint64_t rec_size, uint32_t nb_off, uint64_t n, const float* raw,
                               const float* norm_sq, const float* calib3 /* affine_a, affine_b, ip_qo_floor */, const float* qT,
                               const uint32_t* uplanes, const float* coeffs, uint32_t nq, uint64_t id_begin, uint64_t id_end,
                               uint32_t k, uint32_t kprime, uint32_t nslices, uint32_t* sums, float* est, int64_t* ids, float* dists,
                               uint32_t npieces /* > 1: scan [id_begin, id_end) in that many pieces, as run_exhaustive / sharding.py do */) {
    using namespace cpb;
    DevIndex ix{};
    uint32_t D = 16;
    while (D < dim) D <<= 1;
    ix.D = D; ix.B = 1; ix.dim = dim; ix.nch = (D > 128 ? D : 128) / 128; ix.T = D / 8; ix.n = n;
    ix.aux_off = ix.B * ix.nch * 512;
    ix.block_stride = (ix.aux_off + 644 + 127) / 128 * 128;
    std::vector<uint8_t> dev((size_t)n * ix.block_stride + 256, 0xEE);
    std::vector<float> rawT((size_t)n * D + 64, -7.0f), fnop(n), fipqo(n);
    std::vector<uint32_t> fcodes((size_t)n * ix.nch * 4 + 16, 0xEEEEEEEEu);
    std::vector<uint16_t> fpop(n);
    ix.blocks = dev.data() + (128 - reinterpret_cast<uintptr_t>(dev.data()) % 128) % 128;
    ix.rawT = rawT.data() + (16 - (reinterpret_cast<uintptr_t>(rawT.data()) / 4) % 16) % 16;
    ix.flat_codes = fcodes.data() + (4 - (reinterpret_cast<uintptr_t>(fcodes.data()) / 4) % 4) % 4;
    ix.flat_nop = fnop.data(); ix.flat_ipqo = fipqo.data(); ix.flat_pop = fpop.data();
    ix.norm_sq = norm_sq;
    ix.calib.affine_a = calib3[0]; ix.calib.affine_b = calib3[1]; ix.calib.ip_qo_floor = calib3[2];
    uint32_t problems[2] = {0, 0};
    auto rl_blocks = [&](int) { relayout_blocks_kernel(ix, records, rec_size, nb_off, 0, (uint32_t)n, problems); };
    cuda_emul::launch(rl_blocks, (unsigned)n, 128, smem_raw, 0, 0);
    auto rl_raw = [&](int) { relayout_raw_kernel(ix, raw, 0, (uint32_t)n); };
    cuda_emul::launch(rl_raw, 2, 256, smem_raw, 0, 0);

    ExhaustiveArgs a{};
    a.uplanes = uplanes; a.coeffs = coeffs; a.qT = qT; a.ubytes = nullptr; a.nq = nq; a.id_begin = id_begin; a.id_end = id_end;
    a.k = k; a.kprime = kprime; a.sums = sums; a.est = est; a.ids = ids; a.dists = dists; a.use_tensor_cores = 0;
    if (kprime > kMaxKPrime) return 1;
    // launch_exhaustive's shapes, with the slice count given by the caller (several slices even on a small index)
    const uint64_t m = id_end - id_begin;
    if (nslices < 1) nslices = 1;
    const uint64_t slice_len = m ? (m + nslices - 1) / nslices : 1;
    std::vector<unsigned long long> partial((size_t)nslices * nq * (kprime ? kprime : 1) + 8, 0xEEEEEEEEEEEEEEEEull);
    const uint32_t cap = kprime <= 384 ? 1024u : (uint32_t)kCapMax;
    const size_t smem = (size_t)kQT * cap * 8 + (size_t)kQT * ix.nch * 64 + (size_t)kQT * 16;
    if (smem > sizeof(smem_raw)) return 2;
    if (npieces > 1 && kprime && k) {
        // the candidate interface: the k' best keys of (earlier pieces U this piece), thresholds handed on, exact distances on
        // the last piece, then the merge of the one list -- launch_exhaustive's kernel choices
        std::vector<unsigned long long> keys[2] = {std::vector<unsigned long long>((size_t)nq * kprime), std::vector<unsigned long long>((size_t)nq * kprime)};
        std::vector<float> tau(nq, FLT_MAX), cd((size_t)nq * kprime);
        const uint64_t plen = (m + npieces - 1) / npieces;
        for (uint32_t c = 0; c < npieces; ++c) {
            const bool last = c + 1 == npieces;
            ExhaustiveArgs p = a;
            p.id_begin = id_begin + std::min<uint64_t>(m, (uint64_t)c * plen); p.id_end = id_begin + std::min<uint64_t>(m, (uint64_t)(c + 1) * plen);
            p.k = 0; p.ids = nullptr; p.dists = nullptr; p.sums = nullptr; p.est = nullptr;
            p.prior_keys = c ? keys[(c - 1) & 1].data() : nullptr; p.tau_in = c ? tau.data() : nullptr;
            p.cand_keys = keys[c & 1].data(); p.tau_out = tau.data(); p.cand_dists = last ? cd.data() : nullptr; p.id_offset = 0;
            const uint64_t pm = p.id_end - p.id_begin, psl = pm ? (pm + nslices - 1) / nslices : 1;
            std::fill(partial.begin(), partial.end(), 0xEEEEEEEEEEEEEEEEull);
            auto scan = [&](int) { exhaustive_scan_kernel(ix, p, nslices, psl, cap, partial.data()); };
            cuda_emul::launch(scan, dim3(nslices, (nq + kQT - 1) / kQT), kExThreads, smem_raw, smem, 0);
            if (!last && 2 * kprime <= kSelCap) {
                auto sel = [&](int) { exhaustive_select_keys_kernel(p, nslices, partial.data()); };
                cuda_emul::launch(sel, (nq + kSelWarps - 1) / kSelWarps, kSelWarps * 32, smem_raw, 0, 0);
            } else {
                uint32_t sort_n = 1;
                while (sort_n < (nslices + (p.prior_keys ? 1u : 0u)) * kprime) sort_n <<= 1;
                const size_t smem2 = (((size_t)sort_n * 8 + 15) & ~(size_t)15) + (size_t)8 * (ix.T + 4) * 4;
                if (smem2 > sizeof(smem_raw)) return 3;
                auto sel = [&](int) { exhaustive_select_rerank_kernel(ix, p, nslices, sort_n, partial.data()); };
                cuda_emul::launch(sel, nq, kExThreads, smem_raw, smem2, 0);
            }
            if (last) {
                uint32_t sort_n = 2;
                while (sort_n < kprime) sort_n <<= 1;
                auto mrg = [&](int) { merge_candidates_kernel(keys[c & 1].data(), cd.data(), 1, nq, kprime, k, sort_n, ids, dists, nullptr); };
                cuda_emul::launch(mrg, nq, kExThreads, smem_raw, (size_t)sort_n * 8, 0);
            }
        }
        return 0;
    }
    if (m > 0 || kprime) {
        auto scan = [&](int) { exhaustive_scan_kernel(ix, a, nslices, slice_len, cap, partial.data()); };
        cuda_emul::launch(scan, dim3(nslices, (nq + kQT - 1) / kQT), kExThreads, smem_raw, smem, 0);
    }
    if (kprime && k) {
        uint32_t sort_n = 1;
        while (sort_n < nslices * kprime) sort_n <<= 1;
        const size_t smem2 = (((size_t)sort_n * 8 + 15) & ~(size_t)15) + (size_t)8 * (ix.T + 4) * 4;
        if (smem2 > sizeof(smem_raw)) return 3;
        auto sel = [&](int) { exhaustive_select_rerank_kernel(ix, a, nslices, sort_n, partial.data()); };
        cuda_emul::launch(sel, nq, kExThreads, smem_raw, smem2, 0);
    }
    return 0;
}
