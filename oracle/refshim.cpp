// TEST INFRASTRUCTURE ONLY -- never imported, linked or executed by the product path.
//
// C-ABI shim over the UNMODIFIED reference headers under $(REF)/include (they are included
// where they lie; nothing is copied).  The reference exports only CPIndex through pybind11
// (src/bindings.cpp:115-240); the kernel-level functions of the query hot path are C++
// templates with no binding, so the parity tests reach them through this file:
//
//   refshim_encode_queries   -> RaBitQEncoder<D>::encode_query_raw      (encoder/rabitq_encoder.hpp:73-79,197-209)
//   refshim_rotate           -> RaBitQEncoderBase::rotate_raw_vector    (encoder/rabitq_encoder.hpp:81-86)
//   refshim_fastscan         -> fastscan::compute_inner_products / compute_nbit_inner_products /
//                               compute_msb_only_inner_products          (distance/fastscan_kernel.hpp:17-87,197-217,349-368)
//   refshim_convert          -> fastscan::convert_to_distances_with_bounds / convert_msb_to_lower_bounds /
//                               convert_nbit_to_distances_with_bounds    (distance/fastscan_kernel.hpp:89-194,371-425,220-346)
//   refshim_dot / refshim_l2 -> dot_product_simd<D> / l2_distance_simd<D> (core/memory.hpp:65-96)
//   refshim_neighbor_aux     -> RaBitQEncoder<D>::compute_neighbor_aux / NbitRaBitQEncoder<D,B>::compute_neighbor_aux_nbit
//                               after rotate_raw_vector(parent)            (encoder/rabitq_encoder.hpp:81-86,138-181,287-323,
//                               371-467; called by prune_and_write, graph/graph_refinement.hpp:46-67) -- BUILD side
//                               (SURVEY section 8f, N3): oracle groundwork for the next round, no CUDA counterpart yet
//
// Built by oracle/Makefile into oracle/_ref/libcphnsw_refshim.so (git-ignored).
#include <cphnsw/core/codes.hpp>
#include <cphnsw/core/memory.hpp>
#include <cphnsw/core/types.hpp>
#include <cphnsw/core/util.hpp>
#include <cphnsw/distance/fastscan_kernel.hpp>
#include <cphnsw/distance/fastscan_layout.hpp>
#include <cphnsw/encoder/rabitq_encoder.hpp>

#include <cstdint>
#include <cstring>
#include <type_traits>
#include <vector>

using namespace cphnsw;

#define FOR_EACH_D(X) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048)

namespace {

template <size_t D>
int encode_queries_t(uint32_t dim, uint64_t nq, const float* q, uint8_t* lut, float* coeffs) {
    RaBitQEncoder<D> enc(dim, constants::kDefaultRotationSeed);
    std::vector<float> padded(D);
    for (uint64_t i = 0; i < nq; ++i) {
        // Index::search pads to D floats before encoding (api/hnsw_index.hpp:174-182)
        std::memcpy(padded.data(), q + i * dim, dim * sizeof(float));
        std::memset(padded.data() + dim, 0, (D - dim) * sizeof(float));
        RaBitQQuery<D> rq = enc.encode_query_raw(padded.data());
        std::memcpy(lut + i * (D / 4) * 16, rq.lut, (D / 4) * 16);
        coeffs[3 * i + 0] = rq.coeff_fastscan;
        coeffs[3 * i + 1] = rq.coeff_popcount;
        coeffs[3 * i + 2] = rq.coeff_constant;
    }
    return 0;
}

template <size_t D>
int rotate_t(uint32_t dim, uint64_t nq, const float* q, float* out) {
    RaBitQEncoder<D> enc(dim, constants::kDefaultRotationSeed);
    for (uint64_t i = 0; i < nq; ++i) enc.rotate_raw_vector(q + i * dim, out + i * D);
    return 0;
}

template <size_t D, size_t B>
int fastscan_t(const uint8_t* lut, const uint8_t* planes, uint32_t* nbit, uint32_t* msb, uint32_t* msb2) {
    alignas(64) uint8_t lut_local[D / 4][16];
    std::memcpy(lut_local, lut, sizeof(lut_local));
    if constexpr (B == 1) {
        FastScanCodeBlock<D, 32> blk;
        std::memcpy(blk.packed, planes, sizeof(blk.packed));
        fastscan::compute_inner_products<D>(lut_local, blk, nbit);
        std::memcpy(msb, nbit, 32 * sizeof(uint32_t));
        std::memcpy(msb2, nbit, 32 * sizeof(uint32_t));
    } else {
        NbitFastScanCodeBlock<D, B, 32> blk;
        static_assert(sizeof(blk) == B * 4 * D, "plane layout");
        std::memcpy(&blk, planes, sizeof(blk));
        fastscan::compute_nbit_inner_products<D, B>(lut_local, blk, nbit, msb);
        fastscan::compute_msb_only_inner_products<D, B>(lut_local, blk, msb2);
    }
    return 0;
}

struct QParams { float A, Bc, C, a, b, floor, slack; };

template <size_t D, size_t B>
int convert_t(const QParams& p, const uint32_t* nbit, const uint32_t* msb, const uint32_t* msb2,
              const float* nop, const float* ip_qo, const float* ip_cp,
              const uint16_t* pop, const uint16_t* wpop, uint32_t count, float dqp,
              float* est, float* lower, float* msb_lower) {
    RaBitQQuery<D> q;
    q.coeff_fastscan = p.A; q.coeff_popcount = p.Bc; q.coeff_constant = p.C;
    q.affine_a = p.a; q.affine_b = p.b; q.ip_qo_floor = p.floor; q.dot_slack = p.slack;
    if constexpr (B == 1) {
        fastscan::convert_to_distances_with_bounds<D>(q, nbit, nop, ip_qo, ip_cp, pop, count, est, lower, dqp);
        std::memcpy(msb_lower, lower, count * sizeof(float));
    } else {
        fastscan::convert_msb_to_lower_bounds<D, B>(q, msb2, nop, ip_qo, ip_cp, pop, count, msb_lower, dqp);
        fastscan::convert_nbit_to_distances_with_bounds<D, B>(q, nbit, msb, nop, ip_qo, ip_cp, pop, wpop,
                                                               count, est, lower, dqp);
    }
    return 0;
}

// build side: one parent vertex, n candidate neighbours (all vectors zero-padded to D floats, as RaBitQGraph stores them)
template <size_t D>
int neighbor_aux_1bit_t(uint32_t dim, uint64_t n, const float* parent, const float* nbrs, uint8_t* codes, float* aux) {
    RaBitQEncoder<D> enc(dim, constants::kDefaultRotationSeed);
    alignas(32) float rp[D];
    enc.rotate_raw_vector(parent, rp);
    for (uint64_t i = 0; i < n; ++i) {
        BinaryCodeStorage<D> c;
        VertexAuxData a = enc.compute_neighbor_aux(parent, nbrs + i * D, rp, c);
        std::memcpy(codes + i * (D / 8), c.signs, D / 8);
        aux[3 * i + 0] = a.nop; aux[3 * i + 1] = a.ip_qo; aux[3 * i + 2] = a.ip_cp;
    }
    return 0;
}

template <size_t D, size_t B>
int neighbor_aux_nbit_t(uint32_t dim, uint64_t n, const float* parent, const float* nbrs, uint8_t* codes, float* aux) {
    NbitRaBitQEncoder<D, B> enc(dim, constants::kDefaultRotationSeed);
    alignas(32) float rp[D];
    enc.rotate_raw_vector(parent, rp);
    for (uint64_t i = 0; i < n; ++i) {
        auto r = enc.compute_neighbor_aux_nbit(parent, nbrs + i * D, rp);
        for (size_t b = 0; b < B; ++b) std::memcpy(codes + (i * B + b) * (D / 8), r.code.planes[b], D / 8);   // planes, MSB first
        aux[3 * i + 0] = r.aux.nop; aux[3 * i + 1] = r.aux.ip_qo; aux[3 * i + 2] = r.aux.ip_cp;
    }
    return 0;
}

// The sample loop of Index::calibrate_estimator (api/hnsw_index.hpp:786-866, the lambda process_query) composed from the
// reference's own types and primitives -- l2_distance_simd, encode_query_raw, fastscan::compute_*_inner_products, the
// neighbour-block members -- with the scalar expressions written as the reference writes them and compiled with its flags.
// (The loop itself is a lambda over locals of a private member function and cannot be called; this is the same kind of
// composition as the exhaustive-scan oracle)  records: VertexSearchData<D,32,B>[n], raw: f32 [n][D].
template <size_t D, size_t B>
int calib_samples_t(uint32_t dim, uint64_t ns, const float* queries /* [ns][D], padded */, const uint32_t* start_ids, const float* raw,
                    const uint8_t* records, uint64_t rec_size, uint32_t nb_off, uint32_t* parent_out, float* nn_dist_sq, float* dist_qp_sq_out,
                    float* nop_out, float* ip_corrected_out, float* ip_qo_denom_out, float* true_ip_out, uint32_t* neighbor_out) {
    using NB = std::conditional_t<B == 1, FastScanNeighborBlock<D, 32>, NbitFastScanNeighborBlock<D, 32, B>>;
    using Enc = std::conditional_t<B == 1, RaBitQEncoder<D>, NbitRaBitQEncoder<D, B>>;
    Enc encoder_(dim, constants::kDefaultRotationSeed);
    auto get_vector = [&](uint32_t id) { return raw + (size_t)id * D; };
    auto get_neighbors = [&](uint32_t id) -> const NB& { return *reinterpret_cast<const NB*>(records + (size_t)id * rec_size + nb_off); };
    for (uint64_t s = 0; s < ns; ++s) {
        const float* query_vec = queries + s * D;
        uint32_t parent = start_ids[s];
        float best_dist = l2_distance_simd<D>(query_vec, get_vector(parent));
        const auto& nb = get_neighbors(parent);
        for (size_t i = 0; i < nb.size(); ++i) {
            uint32_t nid = nb.neighbor_ids[i];
            if (nid == INVALID_NODE) break;
            float d = l2_distance_simd<D>(query_vec, get_vector(nid));
            if (d < best_dist) {
                best_dist = d;
                parent = nid;
            }
        }
        nn_dist_sq[s] = best_dist;
        parent_out[s] = parent;
        const auto& pnb = get_neighbors(parent);
        auto encoded = encoder_.encode_query_raw(query_vec);
        float dist_qp_sq = l2_distance_simd<D>(query_vec, get_vector(parent));
        dist_qp_sq_out[s] = dist_qp_sq;
        for (size_t j = 0; j < 32; ++j) {
            nop_out[s * 32 + j] = 0.0f; ip_corrected_out[s * 32 + j] = 0.0f; ip_qo_denom_out[s * 32 + j] = 0.0f;
            true_ip_out[s * 32 + j] = 0.0f; neighbor_out[s * 32 + j] = INVALID_NODE;
        }
        size_t num_batches = (pnb.size() + constants::kFastScanBatch - 1) / constants::kFastScanBatch;
        for (size_t batch = 0; batch < num_batches; ++batch) {
            size_t batch_start = batch * constants::kFastScanBatch;
            size_t batch_count = std::min(constants::kFastScanBatch, pnb.size() - batch_start);
            alignas(64) uint32_t fastscan_sums[constants::kFastScanBatch];
            if constexpr (B == 1) {
                fastscan::compute_inner_products(encoded.lut, pnb.code_blocks[batch], fastscan_sums);
            } else {
                alignas(64) uint32_t msb_sums[constants::kFastScanBatch];
                fastscan::compute_nbit_inner_products<D, B>(encoded.lut, pnb.code_blocks[batch], fastscan_sums, msb_sums);
            }
            for (size_t j = 0; j < batch_count; ++j) {
                size_t ni = batch_start + j;
                uint32_t neighbor = pnb.neighbor_ids[ni];
                if (neighbor == INVALID_NODE) break;
                float ip_qo = pnb.ip_qo[ni];
                float nop = std::max(pnb.nop[ni], constants::eps::kSmall);
                float A = encoded.coeff_fastscan;
                float Bc = encoded.coeff_popcount;
                float C = encoded.coeff_constant;
                float ip_approx;
                if constexpr (B == 1) {
                    ip_approx = A * static_cast<float>(fastscan_sums[j]) + Bc * static_cast<float>(pnb.popcounts[ni]) + C;
                } else {
                    constexpr float K = static_cast<float>((1u << B) - 1);
                    constexpr float inv_K = 1.0f / K;
                    ip_approx = A * inv_K * static_cast<float>(fastscan_sums[j]) + Bc * inv_K * static_cast<float>(pnb.weighted_popcounts[ni]) + C;
                }
                float ip_corrected = ip_approx - pnb.ip_cp[ni];
                float ip_qo_denom = std::max(std::abs(ip_qo), constants::eps::kMedium);
                const float* p_vec = get_vector(parent);
                const float* o_vec = get_vector(neighbor);
                float true_ip = 0.0f;
                for (size_t d = 0; d < D; ++d) {
                    true_ip += (query_vec[d] - p_vec[d]) * (o_vec[d] - p_vec[d]);
                }
                true_ip /= nop;
                nop_out[s * 32 + ni] = nop; ip_corrected_out[s * 32 + ni] = ip_corrected; ip_qo_denom_out[s * 32 + ni] = ip_qo_denom;
                true_ip_out[s * 32 + ni] = true_ip; neighbor_out[s * 32 + ni] = neighbor;
            }
        }
    }
    return 0;
}

}  // namespace

extern "C" {

int refshim_calib_samples(uint32_t D, uint32_t B, uint32_t dim, uint64_t ns, const float* queries, const uint32_t* start_ids, const float* raw,
                          const uint8_t* records, uint64_t rec_size, uint32_t nb_off, uint32_t* parent, float* nn_dist_sq, float* dist_qp_sq,
                          float* nop, float* ip_corrected, float* ip_qo_denom, float* true_ip, uint32_t* neighbor) {
#define CALIB_CASE(DD) if (D == DD) { \
        if (B == 1) return calib_samples_t<DD, 1>(dim, ns, queries, start_ids, raw, records, rec_size, nb_off, parent, nn_dist_sq, dist_qp_sq, nop, ip_corrected, ip_qo_denom, true_ip, neighbor); \
        if (B == 2) return calib_samples_t<DD, 2>(dim, ns, queries, start_ids, raw, records, rec_size, nb_off, parent, nn_dist_sq, dist_qp_sq, nop, ip_corrected, ip_qo_denom, true_ip, neighbor); \
        if (B == 4) return calib_samples_t<DD, 4>(dim, ns, queries, start_ids, raw, records, rec_size, nb_off, parent, nn_dist_sq, dist_qp_sq, nop, ip_corrected, ip_qo_denom, true_ip, neighbor); }
    CALIB_CASE(32) CALIB_CASE(128) CALIB_CASE(1024)
#undef CALIB_CASE
    return -1;
}


int refshim_encode_queries(uint32_t dim, uint64_t nq, const float* q, uint8_t* lut, float* coeffs) {
    size_t pd = next_power_of_two(dim);
    switch (pd) {
#define X(DD) case DD: return encode_queries_t<DD>(dim, nq, q, lut, coeffs);
        FOR_EACH_D(X)
#undef X
    }
    return -1;
}

int refshim_rotate(uint32_t dim, uint64_t nq, const float* q, float* out) {
    size_t pd = next_power_of_two(dim);
    switch (pd) {
#define X(DD) case DD: return rotate_t<DD>(dim, nq, q, out);
        FOR_EACH_D(X)
#undef X
    }
    return -1;
}

int refshim_fastscan(uint32_t D, uint32_t B, const uint8_t* lut, const uint8_t* planes,
                     uint32_t* nbit, uint32_t* msb, uint32_t* msb2) {
#define X(DD) if (D == DD) { \
        if (B == 1) return fastscan_t<DD, 1>(lut, planes, nbit, msb, msb2); \
        if (B == 2) return fastscan_t<DD, 2>(lut, planes, nbit, msb, msb2); \
        if (B == 4) return fastscan_t<DD, 4>(lut, planes, nbit, msb, msb2); }
    FOR_EACH_D(X)
#undef X
    return -1;
}

// params = {coeff_fastscan, coeff_popcount, coeff_constant, affine_a, affine_b, ip_qo_floor, dot_slack}
int refshim_convert(uint32_t D, uint32_t B, const float* params,
                    const uint32_t* nbit, const uint32_t* msb, const uint32_t* msb2,
                    const float* nop, const float* ip_qo, const float* ip_cp,
                    const uint16_t* pop, const uint16_t* wpop, uint32_t count, float dqp,
                    float* est, float* lower, float* msb_lower) {
    QParams p{params[0], params[1], params[2], params[3], params[4], params[5], params[6]};
#define X(DD) if (D == DD) { \
        if (B == 1) return convert_t<DD, 1>(p, nbit, msb, msb2, nop, ip_qo, ip_cp, pop, wpop, count, dqp, est, lower, msb_lower); \
        if (B == 2) return convert_t<DD, 2>(p, nbit, msb, msb2, nop, ip_qo, ip_cp, pop, wpop, count, dqp, est, lower, msb_lower); \
        if (B == 4) return convert_t<DD, 4>(p, nbit, msb, msb2, nop, ip_qo, ip_cp, pop, wpop, count, dqp, est, lower, msb_lower); }
    FOR_EACH_D(X)
#undef X
    return -1;
}

float refshim_dot(uint32_t D, const float* a, const float* b) {
    switch (D) {
#define X(DD) case DD: return dot_product_simd<DD>(a, b);
        FOR_EACH_D(X)
#undef X
    }
    return -1.0f;
}

float refshim_l2(uint32_t D, const float* a, const float* b) {
    switch (D) {
#define X(DD) case DD: return l2_distance_simd<DD>(a, b);
        FOR_EACH_D(X)
#undef X
    }
    return -1.0f;
}

// codes: [n][B][D/8] bytes (bit i of a plane = bit i%8 of byte i/8; planes MSB first; B = 1: the sign bits);
// aux: [n][3] = nop, ip_qo, ip_cp.  D < 64 is not offered (the code storage is then a partial word).
int refshim_neighbor_aux(uint32_t D, uint32_t B, uint32_t dim, uint64_t n, const float* parent, const float* nbrs,
                         uint8_t* codes, float* aux) {
#define X(DD) if (D == DD && DD >= 64) { \
        if (B == 1) return neighbor_aux_1bit_t<DD>(dim, n, parent, nbrs, codes, aux); \
        if (B == 2) return neighbor_aux_nbit_t<DD, 2>(dim, n, parent, nbrs, codes, aux); \
        if (B == 4) return neighbor_aux_nbit_t<DD, 4>(dim, n, parent, nbrs, codes, aux); }
    FOR_EACH_D(X)
#undef X
    return -1;
}

}  // extern "C"
