// TEST / BENCH INFRASTRUCTURE ONLY -- never imported, linked or executed by the product path.
//
// Index builder over the UNMODIFIED reference headers (included where they lie under $(REF)).
// It exists because the stock reference cannot finalize an index larger than ~230k vectors:
// calibrate_estimator draws 15*sqrt(n) calibration queries x 32 neighbours = 480*sqrt(n) residuals
// (api/hnsw_index.hpp:732-733,873-890) but IndexProfile::derive demands an EVT tail of
// evt_min_tail = max(64, sqrt(n)) samples (core/adaptive_defaults.hpp:45-46), while fit_gpd_stable
// can offer at most sqrt(n_resid) = sqrt(480*sqrt(n)) (core/evt_crc.hpp:216-233) -- so for
// n > 230 400 finalize() always ends in "Calibration failed: EVT-CRC fit did not converge."
// (measured: 1M x 128, 4-bit, iid and clustered).  BASELINE.json's configs are 1M and larger.
//
// What this does: build() and finalize() exactly as shipped.  If -- and only if -- finalize()
// throws that calibration error, the graph is already complete (NNDescent, pruning, neighbour
// codes, BFS reorder, upper layers: everything before calibrate_estimator in
// api/hnsw_index.hpp:122-166), so the reference's own calibrate_estimator is run again with
// evt_min_tail lowered to half of what the residual sample can support; nothing else differs.
// The result is written by the reference's own save().  Both bench arms (the stock reference's
// search_batch and the CUDA path) then load that same file.
//
// usage: refbuild <dim> <bits> <n> <vectors.f32> <out.bin>
#define private public   // reach Index<>::profile_ / calibrate_estimator; the headers stay untouched
#include <cphnsw/api/hnsw_index.hpp>
#undef private

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace cphnsw;

template <size_t D, size_t B>
static int run(size_t dim, size_t n, const float* vecs, const char* out) {
    Index<D, 32, B> idx(dim);
    auto t0 = std::chrono::steady_clock::now();
    idx.build(vecs, n);
    bool relaxed = false;
    try {
        idx.finalize();
    } catch (const std::runtime_error& e) {
        if (std::strstr(e.what(), "EVT-CRC fit did not converge") == nullptr) throw;
        const size_t n_calib = std::min(idx.profile_.min_calib_samples, n);
        const double n_resid = 1.5 * (double)n_calib * 32.0;
        const size_t cap = (size_t)(std::sqrt(n_resid) / 2.0);
        std::fprintf(stderr, "refbuild: stock finalize() failed (%s); re-running calibrate_estimator with evt_min_tail %zu -> %zu\n",
                     e.what(), idx.profile_.evt_min_tail, cap);
        idx.profile_.evt_min_tail = std::max<size_t>(64, cap);
        idx.calibrate_estimator(n_calib);
        idx.needs_build_ = false;
        idx.finalized_ = true;
        relaxed = true;
    }
    idx.save(out);
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("{\"n\": %zu, \"dim\": %zu, \"bits\": %zu, \"build_s\": %.1f, \"calibration\": \"%s\"}\n", n, dim, B, s,
                relaxed ? "evt_min_tail_relaxed" : "stock");
    return 0;
}

int main(int argc, char** argv) {
    if (argc != 6) { std::fprintf(stderr, "usage: %s dim bits n vectors.f32 out.bin\n", argv[0]); return 2; }
    const size_t dim = std::strtoull(argv[1], nullptr, 10), bits = std::strtoull(argv[2], nullptr, 10),
                 n = std::strtoull(argv[3], nullptr, 10);
    std::vector<float> v(n * dim);
    FILE* f = std::fopen(argv[4], "rb");
    if (!f || std::fread(v.data(), sizeof(float), v.size(), f) != v.size()) { std::fprintf(stderr, "cannot read %s\n", argv[4]); return 2; }
    std::fclose(f);
    const size_t D = next_power_of_two(dim);
    try {
#define CASE(DD) if (D == DD) { \
        if (bits == 1) return run<DD, 1>(dim, n, v.data(), argv[5]); \
        if (bits == 2) return run<DD, 2>(dim, n, v.data(), argv[5]); \
        if (bits == 4) return run<DD, 4>(dim, n, v.data(), argv[5]); }
        CASE(128) CASE(1024)
#undef CASE
    } catch (const std::exception& e) {
        std::fprintf(stderr, "refbuild: %s\n", e.what());
        return 1;
    }
    std::fprintf(stderr, "refbuild: unsupported dim/bits (built for padded dims 128 and 1024)\n");
    return 2;
}
