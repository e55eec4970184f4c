// C ABI of the library (include/cphnsw_b200.h): index hand-off (save-file v2 reader, upload and
// re-layout), the search entry points and the kernel-level hooks.  Host-side glue only; the
// kernels are in query_prep.cu, fastscan_blocks.cu, search.cu, exhaustive.cu, relayout.cu.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "device_index.h"
#include "kernels.h"

using namespace cpb;

namespace {

thread_local std::string g_create_error;

int fail(cphnsw_b200_index* ix, int code, const std::string& msg) {
    if (ix) ix->err = msg; else g_create_error = msg;
    return code;
}

#define CUDA_TRY(ix, expr)                                                                         \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return fail(ix, CPHNSW_B200_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

template <typename T>
int dev_alloc(cphnsw_b200_index* ix, T** out, size_t count, bool zero = false) {
    void* p = nullptr;
    const size_t bytes = std::max<size_t>(count * sizeof(T), 16);
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return fail(ix, CPHNSW_B200_ENOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    if (zero) cudaMemset(p, 0, bytes);
    ix->allocs.push_back(p);
    ix->device_bytes += bytes;
    *out = static_cast<T*>(p);
    return 0;
}

template <typename T>
int dev_upload(cphnsw_b200_index* ix, const T** out, const T* host, size_t count) {
    T* p = nullptr;
    int rc = dev_alloc(ix, &p, count);
    if (rc) return rc;
    if (count) CUDA_TRY(ix, cudaMemcpy(p, host, count * sizeof(T), cudaMemcpyHostToDevice));
    *out = p;
    return 0;
}

void release_index(cphnsw_b200_index* ix) {
    for (void* p : ix->allocs) cudaFree(p);
    ix->allocs.clear();
    ix->device_bytes = 0;
    ix->loaded = false;
    ix->dev = DevIndex{};
    ix->frontier_budget = 0;
    // the bitmap arena is laid out for one n
    if (ix->bitmaps) { cudaFree(ix->bitmaps); ix->bitmaps = nullptr; ix->bitmaps_bytes = 0; }
}

uint32_t next_pow2(uint32_t v) { uint32_t p = 1; while (p < v) p <<= 1; return p; }

// sizeof(RaBitQCode<D>) / sizeof(NbitRaBitQCode<D,B>): sign words padded to 64 B, two floats, padded to 64 B
uint32_t code_bytes(uint32_t D, uint32_t B) {
    const uint32_t words = (D + 63) / 64;
    const uint32_t storage = (8 * words * B + 63) / 64 * 64;
    return (storage + 8 + 63) / 64 * 64;
}
uint32_t nb_bytes(uint32_t D, uint32_t B) {
    const uint32_t raw = 4 * D * B + 384 + 64 * (B > 1 ? 2 : 1) + 128 + 4;
    return (raw + 63) / 64 * 64;
}

// encoder/rotation.hpp:19-32: three sign diagonals drawn layer-major from one
// std::mt19937_64(seed) through std::uniform_int_distribution<int>(0,1) -- the same standard
// library calls, so the same values under the same libstdc++.
std::vector<float> rotation_signs(uint32_t D, uint64_t seed) {
    std::mt19937_64 rng(seed);
    std::uniform_int_distribution<int> coin(0, 1);
    std::vector<float> s((size_t)3 * D);
    for (uint32_t layer = 0; layer < 3; ++layer)
        for (uint32_t i = 0; i < D; ++i) s[(size_t)layer * D + i] = coin(rng) ? 1.0f : -1.0f;
    return s;
}

template <typename T>
T rd(const uint8_t* p, size_t off) { T v; std::memcpy(&v, p + off, sizeof(T)); return v; }

int ensure_buffer(cphnsw_b200_index* ix, void** buf, size_t* have, size_t want, bool zero) {
    if (*have >= want) return 0;
    if (*buf) { cudaFree(*buf); *buf = nullptr; *have = 0; }
    want = (want + (1 << 20)) & ~(size_t)((1 << 20) - 1);
    cudaError_t e = cudaMalloc(buf, want);
    if (e != cudaSuccess) return fail(ix, CPHNSW_B200_ENOMEM, std::string("cudaMalloc(scratch): ") + cudaGetErrorString(e));
    if (zero) { e = cudaMemset(*buf, 0, want); if (e != cudaSuccess) return fail(ix, CPHNSW_B200_ECUDA, cudaGetErrorString(e)); }
    *have = want;
    return 0;
}

struct QStateView { float* qT; uint32_t* uplanes; float* coeffs; uint8_t* ubytes; };

int ensure_qstate(cphnsw_b200_index* ix, uint64_t nq, QStateView* v) {
    const DevIndex& d = ix->dev;
    const size_t qT = (size_t)nq * d.D * 4, up = (size_t)nq * 16 * d.nch * 4, co = (size_t)nq * kCoeffStride * 4;
    const size_t a = (qT + 255) & ~(size_t)255, b = (up + 255) & ~(size_t)255, c = (co + 255) & ~(size_t)255;
    const size_t ub = (size_t)nq * d.nch * 128;
    int rc = ensure_buffer(ix, &ix->qstate, &ix->qstate_bytes, a + b + c + ub + 256, false);
    if (rc) return rc;
    uint8_t* p = static_cast<uint8_t*>(ix->qstate);
    v->qT = reinterpret_cast<float*>(p);
    v->uplanes = reinterpret_cast<uint32_t*>(p + a);
    v->coeffs = reinterpret_cast<float*>(p + a + b);
    v->ubytes = p + a + b + c;
    return 0;
}

int require_loaded(cphnsw_b200_index* ix) {
    if (!ix) return fail(nullptr, CPHNSW_B200_EINVAL, "null index handle");
    if (!ix->loaded) return fail(ix, CPHNSW_B200_ERUNTIME, "Index has no finalized data on the device (load or upload first).");
    cudaError_t e = cudaSetDevice(ix->device);
    if (e != cudaSuccess) return fail(ix, CPHNSW_B200_ECUDA, cudaGetErrorString(e));
    return 0;
}

}  // namespace

extern "C" {

int cphnsw_b200_create(int device, cphnsw_b200_index** out) {
    if (!out) return fail(nullptr, CPHNSW_B200_EINVAL, "out is null");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, CPHNSW_B200_ECUDA,
                    std::string("no CUDA device: this library has no CPU path (") + cudaGetErrorString(e) + ")");
    if (device < 0 || device >= ndev) return fail(nullptr, CPHNSW_B200_EINVAL, "device ordinal out of range");
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, CPHNSW_B200_ECUDA, cudaGetErrorString(e));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail(nullptr, CPHNSW_B200_ECUDA, cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, CPHNSW_B200_ECUDA, "this build carries sm_100a code only (B200); found sm_" +
                                                    std::to_string(prop.major) + std::to_string(prop.minor));
    auto* ix = new cphnsw_b200_index();
    ix->device = device;
    ix->num_sms = prop.multiProcessorCount;
    if (cudaMalloc(reinterpret_cast<void**>(&ix->d_stats), sizeof(Stats)) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&ix->d_counters), 16) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ix->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        [&] { for (auto& e : ix->ev) if (cudaEventCreate(&e) != cudaSuccess) return true; return false; }()) {
        delete ix;
        return fail(nullptr, CPHNSW_B200_ECUDA, "could not allocate control buffers");
    }
    *out = ix;
    return 0;
}

void cphnsw_b200_destroy(cphnsw_b200_index* ix) {
    if (!ix) return;
    cudaSetDevice(ix->device);
    release_index(ix);
    if (ix->scratch) cudaFree(ix->scratch);
    if (ix->bitmaps) cudaFree(ix->bitmaps);
    if (ix->qstate) cudaFree(ix->qstate);
    if (ix->stage) cudaFree(ix->stage);
    if (ix->d_stats) cudaFree(ix->d_stats);
    if (ix->d_counters) cudaFree(ix->d_counters);
    if (ix->enc_signs) cudaFree(ix->enc_signs);
    if (ix->enc_scratch) cudaFree(ix->enc_scratch);
    if (ix->own_stream) cudaStreamDestroy(ix->own_stream);
    for (auto& e : ix->ev) if (e) cudaEventDestroy(e);
    delete ix;
}

const char* cphnsw_b200_last_error(const cphnsw_b200_index* ix) { return ix ? ix->err.c_str() : g_create_error.c_str(); }

int cphnsw_b200_set_option(cphnsw_b200_index* ix, const char* name, int64_t value) {
    if (!ix || !name) return fail(ix, CPHNSW_B200_EINVAL, "null argument");
    const std::string n(name);
    if (n == "warps_per_cta") { if (value < 1 || value > 8) return fail(ix, CPHNSW_B200_EINVAL, "warps_per_cta must be 1..8"); ix->warps_per_cta = value; }
    else if (n == "ctas_per_sm") { if (value < 1 || value > 32) return fail(ix, CPHNSW_B200_EINVAL, "ctas_per_sm must be 1..32"); ix->ctas_per_sm = value; }
    else if (n == "collect_stats") ix->collect_stats = value ? 1 : 0;
    else if (n == "exhaustive_tensor_cores") ix->exhaustive_tensor_cores = value < 0 ? 0 : (value > 2 ? 2 : value);
    else if (n == "neighbor_codes_tile") ix->neighbor_codes_tile = value < 0 ? 0 : (value > 2 ? 2 : value);
    else if (n == "beam_capacity") { if (value < 64) return fail(ix, CPHNSW_B200_EINVAL, "beam_capacity must be >= 64"); ix->beam_capacity = value; }
    else return fail(ix, CPHNSW_B200_EINVAL, "unknown option " + n);
    return 0;
}

int cphnsw_b200_upload(cphnsw_b200_index* ix, const cphnsw_b200_host_index* h) {
    if (!ix || !h) return fail(ix, CPHNSW_B200_EINVAL, "null argument");
    CUDA_TRY(ix, cudaSetDevice(ix->device));
    // the factory's checks (src/bindings.cpp:77-113)
    if (h->bits != 1 && h->bits != 2 && h->bits != 4)
        return fail(ix, CPHNSW_B200_EINVAL, "Unsupported bits=" + std::to_string(h->bits) + ". Supported: 1, 2, 4.");
    if (h->dim == 0 || h->D != next_pow2(h->dim) || h->D < 16 || h->D > 2048)
        return fail(ix, CPHNSW_B200_EINVAL, "Unsupported dimension " + std::to_string(h->dim) +
                                                ". Supported padded dims: 16, 32, 64, 128, 256, 512, 1024, 2048.");
    if (h->n == 0) return fail(ix, CPHNSW_B200_ERUNTIME, "index is empty");
    if (h->n >= 0xFFFFFFFFull) return fail(ix, CPHNSW_B200_EINVAL, "too many vertices for 32-bit ids");
    if (h->nb_off != code_bytes(h->D, h->bits) || h->rec_size != (uint64_t)h->nb_off + nb_bytes(h->D, h->bits))
        return fail(ix, CPHNSW_B200_ERUNTIME, "Parameter mismatch: record size does not match D/R/BitWidth.");
    if (h->max_level > kMaxLevels || (h->max_level > 0 && h->n_layers > (uint32_t)kMaxLevels))
        return fail(ix, CPHNSW_B200_ERUNTIME, "too many HNSW levels");
    if (h->entry_point >= h->n || (h->max_level <= 0 && h->graph_entry_point >= h->n))
        return fail(ix, CPHNSW_B200_ERUNTIME, "Invalid entry point in index.");
    release_index(ix);

    DevIndex d{};
    d.D = h->D; d.B = h->bits; d.dim = h->dim; d.n = h->n;
    d.nch = std::max(h->D, 128u) / 128; d.T = h->D / 8;
    d.aux_off = d.B * d.nch * 512;
    d.block_stride = (d.aux_off + 644 + 127) / 128 * 128;
    d.max_level = h->max_level; d.entry_point = h->entry_point; d.graph_entry_point = h->graph_entry_point;

    // CalibrationSnapshot offsets: SURVEY App. B (api/hnsw_index.hpp:33-58)
    const uint8_t* cal = h->calibration;
    d.calib.affine_a = rd<float>(cal, 0); d.calib.affine_b = rd<float>(cal, 4); d.calib.ip_qo_floor = rd<float>(cal, 8);
    d.calib.gamma_max = rd<float>(cal, 84); d.calib.gamma_beta = rd<float>(cal, 88);
    d.calib.gamma_warmup = rd<uint64_t>(cal, 96);
    for (int i = 0; i < 32; ++i) d.calib.slack[i] = rd<float>(cal, 108 + 4 * i);
    d.calib.num_slack = rd<int32_t>(cal, 236);
    d.calib.gamma = rd<float>(cal, 240);
    if (d.calib.num_slack > 32) return fail(ix, CPHNSW_B200_ERUNTIME, "calibration snapshot is corrupt");

    int rc;
    uint8_t* blocks = nullptr; float* rawT = nullptr;
    if ((rc = dev_alloc(ix, &blocks, (size_t)d.n * d.block_stride, true))) { release_index(ix); return rc; }
    if ((rc = dev_alloc(ix, &rawT, (size_t)d.n * d.D))) { release_index(ix); return rc; }
    d.blocks = blocks; d.rawT = rawT;
    if ((rc = dev_upload(ix, &d.norm_sq, h->norm_sq, d.n))) { release_index(ix); return rc; }
    std::vector<float> cen(h->dim, 0.0f);
    if (h->centroid) std::memcpy(cen.data(), h->centroid, sizeof(float) * h->dim);
    if ((rc = dev_upload(ix, &d.centroid, cen.data(), h->dim))) { release_index(ix); return rc; }
    const std::vector<float> signs = rotation_signs(d.D, h->rotation_seed);
    if ((rc = dev_upload(ix, &d.signs, signs.data(), signs.size()))) { release_index(ix); return rc; }
    if (d.B == 1) {
        uint32_t* fc = nullptr; float* fn = nullptr; float* fq = nullptr; uint16_t* fp = nullptr;
        if ((rc = dev_alloc(ix, &fc, (size_t)d.n * d.nch * 4)) || (rc = dev_alloc(ix, &fn, d.n)) ||
            (rc = dev_alloc(ix, &fq, d.n)) || (rc = dev_alloc(ix, &fp, d.n))) { release_index(ix); return rc; }
        d.flat_codes = fc; d.flat_nop = fn; d.flat_ipqo = fq; d.flat_pop = fp;
    }

    // upper layers -> slot-addressed CSR
    d.n_levels = h->max_level > 0 ? h->n_layers : 0;
    d.entry_slot = kInvalid;
    std::vector<std::vector<uint32_t>> nbr_slot(d.n_levels), down(d.n_levels);
    for (uint32_t L = 0; L < d.n_levels; ++L) {
        const uint32_t sz = h->layer_sizes[L];
        const uint32_t* nodes = h->layer_nodes[L];
        const uint32_t* offs = h->layer_offs[L];
        const uint32_t total = sz ? offs[sz] : 0;
        auto slot_of = [&](const uint32_t* arr, uint32_t count, uint32_t node) -> uint32_t {
            const uint32_t* it = std::lower_bound(arr, arr + count, node);   // find_edge, hnsw_index.hpp:468-474
            return (it != arr + count && *it == node) ? (uint32_t)(it - arr) : kInvalid;
        };
        nbr_slot[L].resize(total);
        for (uint32_t j = 0; j < total; ++j) {
            if (h->layer_nbrs[L][j] >= d.n) { release_index(ix); return fail(ix, CPHNSW_B200_ERUNTIME, "upper-layer neighbour id out of range"); }
            nbr_slot[L][j] = slot_of(nodes, sz, h->layer_nbrs[L][j]);
        }
        down[L].resize(sz);
        for (uint32_t s = 0; s < sz; ++s) {
            if (nodes[s] >= d.n) { release_index(ix); return fail(ix, CPHNSW_B200_ERUNTIME, "upper-layer node id out of range"); }
            down[L][s] = L == 0 ? nodes[s] : slot_of(h->layer_nodes[L - 1], h->layer_sizes[L - 1], nodes[s]);
        }
        Level& lv = d.levels[L];
        lv.size = sz;
        const uint32_t zero = 0;
        if ((rc = dev_upload(ix, &lv.node, nodes, sz)) || (rc = dev_upload(ix, &lv.offs, sz ? offs : &zero, sz ? sz + 1 : 1)) ||
            (rc = dev_upload(ix, &lv.nbr_node, h->layer_nbrs[L], total)) ||
            (rc = dev_upload(ix, &lv.nbr_slot, nbr_slot[L].data(), total)) ||
            (rc = dev_upload(ix, &lv.down, down[L].data(), sz))) { release_index(ix); return rc; }
        if ((int32_t)L + 1 == h->max_level) d.entry_slot = slot_of(nodes, sz, h->entry_point);
    }

    // records and raw vectors: copy as they are, re-lay out on the device
    ix->dev = d;
    CUDA_TRY(ix, cudaMemset(ix->d_counters, 0, 16));
    const size_t chunk_bytes = (size_t)256 << 20;
    void* stage = nullptr;
    CUDA_TRY(ix, cudaMalloc(&stage, chunk_bytes));
    auto cleanup = [&](int code) { cudaFree(stage); release_index(ix); return code; };
    {
        const uint32_t per = (uint32_t)std::max<size_t>(1, chunk_bytes / h->rec_size);
        for (uint64_t first = 0; first < d.n; first += per) {
            const uint32_t cnt = (uint32_t)std::min<uint64_t>(per, d.n - first);
            cudaError_t e = cudaMemcpy(stage, h->search_data + first * h->rec_size, (size_t)cnt * h->rec_size, cudaMemcpyHostToDevice);
            if (e == cudaSuccess) e = launch_relayout_blocks(d, static_cast<const uint8_t*>(stage), h->rec_size, h->nb_off, first, cnt, ix->d_counters + 2, 0);
            if (e == cudaSuccess) e = cudaDeviceSynchronize();
            if (e != cudaSuccess) return cleanup(fail(ix, CPHNSW_B200_ECUDA, std::string("re-layout of neighbour blocks: ") + cudaGetErrorString(e)));
        }
        const uint32_t perv = (uint32_t)std::max<size_t>(1, chunk_bytes / ((size_t)d.D * 4));
        for (uint64_t first = 0; first < d.n; first += perv) {
            const uint32_t cnt = (uint32_t)std::min<uint64_t>(perv, d.n - first);
            cudaError_t e = cudaMemcpy(stage, h->raw + first * d.D, (size_t)cnt * d.D * 4, cudaMemcpyHostToDevice);
            if (e == cudaSuccess) e = launch_relayout_raw(d, static_cast<const float*>(stage), first, cnt, 0);
            if (e == cudaSuccess) e = cudaDeviceSynchronize();
            if (e != cudaSuccess) return cleanup(fail(ix, CPHNSW_B200_ECUDA, std::string("re-layout of raw vectors: ") + cudaGetErrorString(e)));
        }
    }
    cudaFree(stage);
    uint32_t problems[4] = {0, 0, 0, 0};
    CUDA_TRY(ix, cudaMemcpy(problems, ix->d_counters, 16, cudaMemcpyDeviceToHost));
    if (problems[3]) { release_index(ix); return fail(ix, CPHNSW_B200_ERUNTIME, "Index is corrupt: neighbour id out of range in " + std::to_string(problems[3]) + " blocks."); }
    ix->dev.dup_neighbors = problems[2];
    ix->loaded = true;
    return 0;
}

int cphnsw_b200_load(cphnsw_b200_index* ix, const char* path) {
    if (!ix || !path) return fail(ix, CPHNSW_B200_EINVAL, "null argument");
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(ix, CPHNSW_B200_ERUNTIME, std::string("Cannot open file for reading: ") + path);
    struct stat sb;
    if (fstat(fd, &sb) != 0 || sb.st_size < 68 + 248 + 72) { close(fd); return fail(ix, CPHNSW_B200_ERUNTIME, "Index file is truncated."); }
    const size_t fsize = (size_t)sb.st_size;
    void* map = mmap(nullptr, fsize, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (map == MAP_FAILED) return fail(ix, CPHNSW_B200_ERUNTIME, "mmap of the index file failed");
    const uint8_t* p = static_cast<const uint8_t*>(map);
    auto done = [&](int code) { munmap(map, fsize); return code; };

    // header (api/hnsw_index.hpp:217-245)
    if (rd<uint64_t>(p, 0) != 0x57534E48504300ull) return done(fail(ix, CPHNSW_B200_ERUNTIME, "Invalid magic bytes (not a CP-HNSW index file)."));
    if (rd<uint32_t>(p, 8) != 2) return done(fail(ix, CPHNSW_B200_ERUNTIME, "Unsupported index file version: " + std::to_string(rd<uint32_t>(p, 8))));
    cphnsw_b200_host_index h{};
    h.D = rd<uint32_t>(p, 12);
    const uint32_t R = rd<uint32_t>(p, 16);
    h.bits = rd<uint32_t>(p, 20);
    h.dim = rd<uint32_t>(p, 24);
    h.n = rd<uint64_t>(p, 28);
    h.max_level = rd<int32_t>(p, 36);
    h.entry_point = rd<uint32_t>(p, 40);
    h.graph_entry_point = h.entry_point;   // load() restores the graph entry from the header (:424-428)
    h.rotation_seed = rd<uint64_t>(p, 60);
    if (R != kR) return done(fail(ix, CPHNSW_B200_ERUNTIME, "Parameter mismatch: file R=" + std::to_string(R) + ", expected 32."));
    if ((h.bits != 1 && h.bits != 2 && h.bits != 4) || h.dim == 0 || h.D != next_pow2(h.dim) || h.D < 16 || h.D > 2048)
        return done(fail(ix, CPHNSW_B200_ERUNTIME, "Parameter mismatch: unsupported D/BitWidth/dim in index file."));
    size_t off = 68;
    h.calibration = p + off; off += 248 + 72;   // CalibrationSnapshot, IndexProfile
    h.nb_off = code_bytes(h.D, h.bits);
    h.rec_size = (uint64_t)h.nb_off + nb_bytes(h.D, h.bits);
    const size_t need = off + (size_t)4 * h.dim + (size_t)8 * h.n + (size_t)4 * h.n * h.D + (size_t)h.n * h.rec_size + 4;
    if (h.n == 0 || need > fsize) return done(fail(ix, CPHNSW_B200_ERUNTIME, "Index file is truncated."));
    h.centroid = reinterpret_cast<const float*>(p + off); off += (size_t)4 * h.dim;
    off += (size_t)4 * h.n;   // node_levels (build-time only)
    h.norm_sq = reinterpret_cast<const float*>(p + off); off += (size_t)4 * h.n;
    h.raw = reinterpret_cast<const float*>(p + off); off += (size_t)4 * h.n * h.D;
    h.search_data = p + off; off += (size_t)h.n * h.rec_size;
    const uint32_t n_layers = rd<uint32_t>(p, off); off += 4;
    if (n_layers > (uint32_t)kMaxLevels) return done(fail(ix, CPHNSW_B200_ERUNTIME, "Index file is corrupt (layer count)."));
    std::vector<std::vector<uint32_t>> nodes(n_layers), offs(n_layers), nbrs(n_layers);
    for (uint32_t L = 0; L < n_layers; ++L) {
        if (off + 4 > fsize) return done(fail(ix, CPHNSW_B200_ERUNTIME, "Index file is truncated."));
        const uint32_t ne = rd<uint32_t>(p, off); off += 4;
        nodes[L].reserve(ne); offs[L].reserve(ne + 1); offs[L].push_back(0);
        for (uint32_t e = 0; e < ne; ++e) {
            if (off + 8 > fsize) return done(fail(ix, CPHNSW_B200_ERUNTIME, "Index file is truncated."));
            const uint32_t node = rd<uint32_t>(p, off), cnt = rd<uint32_t>(p, off + 4); off += 8;
            if (off + (size_t)4 * cnt > fsize) return done(fail(ix, CPHNSW_B200_ERUNTIME, "Index file is truncated."));
            nodes[L].push_back(node);
            for (uint32_t j = 0; j < cnt; ++j) nbrs[L].push_back(rd<uint32_t>(p, off + 4 * j));
            off += (size_t)4 * cnt;
            offs[L].push_back((uint32_t)nbrs[L].size());
        }
        if (!std::is_sorted(nodes[L].begin(), nodes[L].end())) return done(fail(ix, CPHNSW_B200_ERUNTIME, "Index file is corrupt (layer edges not sorted)."));
    }
    std::vector<const uint32_t*> pn(n_layers), po(n_layers), pb(n_layers);
    std::vector<uint32_t> sizes(n_layers);
    const uint32_t dummy = 0;
    for (uint32_t L = 0; L < n_layers; ++L) {
        pn[L] = nodes[L].empty() ? &dummy : nodes[L].data();
        po[L] = offs[L].data();
        pb[L] = nbrs[L].empty() ? &dummy : nbrs[L].data();
        sizes[L] = (uint32_t)nodes[L].size();
    }
    h.n_layers = n_layers;
    h.layer_nodes = pn.data(); h.layer_offs = po.data(); h.layer_nbrs = pb.data(); h.layer_sizes = sizes.data();
    return done(cphnsw_b200_upload(ix, &h));
}

int cphnsw_b200_get_info(const cphnsw_b200_index* ix, cphnsw_b200_info* out) {
    if (!ix || !out) return CPHNSW_B200_EINVAL;
    if (!ix->loaded) return CPHNSW_B200_ERUNTIME;
    const DevIndex& d = ix->dev;
    std::memset(out, 0, sizeof(*out));
    out->D = d.D; out->bits = d.B; out->dim = d.dim; out->n = d.n; out->max_level = d.max_level;
    out->entry_point = d.entry_point; out->n_layers = d.n_levels; out->block_stride = d.block_stride;
    out->device_bytes = ix->device_bytes;
    out->affine_a = d.calib.affine_a; out->affine_b = d.calib.affine_b; out->ip_qo_floor = d.calib.ip_qo_floor;
    out->search_gamma = d.calib.gamma; out->gamma_max = d.calib.gamma_max; out->gamma_beta = d.calib.gamma_beta;
    out->gamma_warmup = d.calib.gamma_warmup; out->num_slack_levels = d.calib.num_slack;
    std::memcpy(out->slack_levels, d.calib.slack, sizeof(out->slack_levels));
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// search
// ---------------------------------------------------------------------------------------------------
static int run_search(cphnsw_b200_index* ix, const float* d_queries, uint64_t nq, uint64_t k_user, int64_t* d_ids,
                      float* d_dists, uint32_t* d_entry_out, cudaStream_t stream) {
    const DevIndex& d = ix->dev;
    if (nq == 0) return 0;
    if (nq > 0x7FFFFFFFull) return fail(ix, CPHNSW_B200_EINVAL, "too many queries in one batch");
    if (k_user > 0x7FFFFFFFull) return fail(ix, CPHNSW_B200_EINVAL, "k too large");
    const uint32_t k = (uint32_t)std::max<uint64_t>(k_user, 1);   // hnsw_index.hpp:187

    QStateView qs;
    int rc = ensure_qstate(ix, nq, &qs);
    if (rc) return rc;
    PrepOut po{};
    po.coeffs = qs.coeffs; po.uplanes = qs.uplanes; po.qT = qs.qT;
    CUDA_TRY(ix, cudaEventRecord(ix->ev[0], stream));
    CUDA_TRY(ix, launch_query_prep(d, d_queries, (uint32_t)nq, 0, po, stream));
    CUDA_TRY(ix, cudaEventRecord(ix->ev[1], stream));

    // launch geometry: persistent grid, one warp per in-flight query; as much of the frontier heap in
    // shared memory as still lets the requested number of warps reside
    const bool stats = ix->collect_stats != 0;
    int warps = (int)ix->warps_per_cta;
    const size_t smem_budget = 220 * 1024;
    while (warps > 1 && search_smem_per_warp(d, k) * warps > smem_budget) --warps;
    int per_sm = search_max_ctas_per_sm(d, k, warps, stats);
    if (per_sm <= 0) return fail(ix, CPHNSW_B200_ECUDA, "search kernel cannot be resident (shared memory / registers)");
    per_sm = std::min<int>(per_sm, (int)ix->ctas_per_sm);
    int ctas = ix->num_sms * per_sm;
    const int need = (int)((nq + warps - 1) / warps);
    if (ctas > need) ctas = need;

    auto layout = [&](uint32_t cap, SearchArgs& a) {
        const uint32_t words = (uint32_t)((d.n + 31) / 32);
        uint32_t chunk = 32, shift = 10;         // 32 chunks (one dirty bit each in a lane register),
        while (chunk * 32 < words) { chunk <<= 1; ++shift; }   // power-of-two sized so id -> chunk is a shift
        a.chunk_words = chunk;
        a.chunk_shift = shift;
        a.bitmap_words = chunk * 32;
        size_t off = 0;
        a.heap_off = off; off += ((size_t)(cap + 2) * 16 + 127) & ~(size_t)127;
        a.nn_off = off; if (k > 128) off += ((size_t)k * 8 + 127) & ~(size_t)127;
        a.slot_stride = off;
        a.beam_capacity = cap;
    };

    SearchArgs a{};
    a.nq = (uint32_t)nq; a.query_list = nullptr; a.k = k; a.kout = (uint32_t)k_user;
    a.ids = d_ids; a.dists = d_dists; a.qT = qs.qT; a.uplanes = qs.uplanes; a.coeffs = qs.coeffs;
    a.entry_out = d_entry_out;
    // frontier arena per slot: the option if set, else as much as an 8 GiB (or an eighth of free HBM) budget
    // gives every slot -- a frontier can hold at most n entries, and re-running an overflowed query
    // throws its first attempt away, so be generous where memory allows (1-bit indexes reach 40k+)
    const size_t slots = (size_t)ctas * warps;
    uint64_t cap64 = (uint64_t)ix->beam_capacity;
    if (cap64 == 0) {
        if (ix->frontier_budget == 0) {   // once per index: the answer must not drift from call to call
            size_t free_b = 0, total_b = 0;
            cudaMemGetInfo(&free_b, &total_b);
            ix->frontier_budget = std::min<size_t>((size_t)8 << 30, (free_b + ix->scratch_bytes) / 8);
        }
        cap64 = std::max<uint64_t>(4096, ix->frontier_budget / (slots * 16));
    }
    uint32_t cap = (uint32_t)std::min<uint64_t>(cap64, d.n + 1);
    layout(cap, a);
    const size_t list_bytes = ((size_t)nq * 4 + 255) & ~(size_t)255;
    rc = ensure_buffer(ix, &ix->scratch, &ix->scratch_bytes, list_bytes + slots * a.slot_stride, false);
    if (rc) return rc;
    rc = ensure_buffer(ix, &ix->bitmaps, &ix->bitmaps_bytes, slots * (size_t)a.bitmap_words * 4, true);
    if (rc) return rc;
    a.bitmaps = static_cast<uint32_t*>(ix->bitmaps);
    a.overflow_list = static_cast<uint32_t*>(ix->scratch);
    a.scratch = static_cast<uint8_t*>(ix->scratch) + list_bytes;
    a.counters = ix->d_counters;
    a.stats = ix->d_stats;
    CUDA_TRY(ix, cudaMemsetAsync(ix->d_counters, 0, 16, stream));
    CUDA_TRY(ix, cudaMemsetAsync(ix->d_stats, 0, sizeof(Stats), stream));
    CUDA_TRY(ix, cudaEventRecord(ix->ev[2], stream));
    CUDA_TRY(ix, launch_search(d, a, ctas, warps, stats, stream));
    CUDA_TRY(ix, cudaEventRecord(ix->ev[3], stream));

    // frontier overflow: re-run those queries with an arena that cannot overflow (each id enters
    // the frontier at most once, so n entries always suffice)
    uint32_t counters[4];
    CUDA_TRY(ix, cudaMemcpyAsync(counters, ix->d_counters, 16, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(ix, cudaStreamSynchronize(stream));
    ix->last_stats.overflow_retries = 0;
    cudaEventElapsedTime(&ix->prep_ms, ix->ev[0], ix->ev[1]);
    cudaEventElapsedTime(&ix->search_ms, ix->ev[2], ix->ev[3]);
    if (counters[1] > 0) {
        const uint32_t nover = counters[1];
        std::vector<uint32_t> list(nover);
        CUDA_TRY(ix, cudaMemcpy(list.data(), a.overflow_list, (size_t)nover * 4, cudaMemcpyDeviceToHost));
        SearchArgs b = a;
        layout((uint32_t)d.n + 1, b);
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        const size_t budget = std::max<size_t>(b.slot_stride, (free_b + ix->scratch_bytes) / 2);
        size_t rslots = std::min<size_t>(nover, budget / b.slot_stride);
        int rwarps = (int)std::min<size_t>(warps, rslots);
        int rctas = (int)std::max<size_t>(1, rslots / rwarps);
        rc = ensure_buffer(ix, &ix->scratch, &ix->scratch_bytes, list_bytes + (size_t)rctas * rwarps * b.slot_stride, false);
        if (rc) return rc;
        // the first pass left every bitmap clean, so the same bitmap arena serves the re-run
        rc = ensure_buffer(ix, &ix->bitmaps, &ix->bitmaps_bytes, (size_t)rctas * rwarps * b.bitmap_words * 4, true);
        if (rc) return rc;
        b.bitmaps = static_cast<uint32_t*>(ix->bitmaps);
        b.overflow_list = static_cast<uint32_t*>(ix->scratch);
        b.scratch = static_cast<uint8_t*>(ix->scratch) + list_bytes;
        uint32_t* d_list = nullptr;
        CUDA_TRY(ix, cudaMalloc(reinterpret_cast<void**>(&d_list), (size_t)nover * 4));
        cudaMemcpy(d_list, list.data(), (size_t)nover * 4, cudaMemcpyHostToDevice);
        b.query_list = d_list; b.nq = nover;
        cudaMemsetAsync(ix->d_counters, 0, 16, stream);
        cudaEventRecord(ix->ev[4], stream);
        cudaError_t e = launch_search(d, b, rctas, rwarps, stats, stream);
        cudaEventRecord(ix->ev[5], stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(counters, ix->d_counters, 16, cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        cudaFree(d_list);
        if (e != cudaSuccess) return fail(ix, CPHNSW_B200_ECUDA, std::string("overflow re-run: ") + cudaGetErrorString(e));
        if (counters[1] != 0) return fail(ix, CPHNSW_B200_ERUNTIME, "internal error: frontier overflow with a full-size arena");
        ix->last_stats.overflow_retries = nover;
        float rerun_ms = 0.0f;
        cudaEventElapsedTime(&rerun_ms, ix->ev[4], ix->ev[5]);
        ix->search_ms += rerun_ms;
    }
    return 0;
}

int cphnsw_b200_search_batch_device(cphnsw_b200_index* ix, const float* d_queries, uint64_t nq, uint64_t k,
                                    int64_t* d_ids, float* d_dists, void* stream) {
    int rc = require_loaded(ix);
    if (rc) return rc;
    if (nq && (!d_queries || (k && (!d_ids || !d_dists)))) return fail(ix, CPHNSW_B200_EINVAL, "null buffer");
    return run_search(ix, d_queries, nq, k, d_ids, d_dists, nullptr, static_cast<cudaStream_t>(stream));
}

int cphnsw_b200_search_batch(cphnsw_b200_index* ix, const float* queries, uint64_t nq, uint64_t k, int64_t* ids,
                             float* dists) {
    int rc = require_loaded(ix);
    if (rc) return rc;
    if (nq == 0) return 0;
    if (!queries || (k && (!ids || !dists))) return fail(ix, CPHNSW_B200_EINVAL, "null buffer");
    const DevIndex& d = ix->dev;
    const size_t qb = ((size_t)nq * d.dim * 4 + 255) & ~(size_t)255, ib = ((size_t)nq * k * 8 + 255) & ~(size_t)255,
                 db = (size_t)nq * k * 4;
    rc = ensure_buffer(ix, &ix->stage, &ix->stage_bytes, qb + ib + db + 256, false);
    if (rc) return rc;
    uint8_t* s = static_cast<uint8_t*>(ix->stage);
    float* d_q = reinterpret_cast<float*>(s);
    int64_t* d_i = reinterpret_cast<int64_t*>(s + qb);
    float* d_d = reinterpret_cast<float*>(s + qb + ib);
    cudaStream_t st = ix->own_stream;
    CUDA_TRY(ix, cudaMemcpyAsync(d_q, queries, (size_t)nq * d.dim * 4, cudaMemcpyHostToDevice, st));
    rc = run_search(ix, d_q, nq, k, d_i, d_d, nullptr, st);
    if (rc) return rc;
    if (k) {
        CUDA_TRY(ix, cudaMemcpyAsync(ids, d_i, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(ix, cudaMemcpyAsync(dists, d_d, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(ix, cudaStreamSynchronize(st));
    return 0;
}

int cphnsw_b200_last_timings(cphnsw_b200_index* ix, float* prep_ms, float* search_ms) {
    if (!ix) return CPHNSW_B200_EINVAL;
    if (prep_ms) *prep_ms = ix->prep_ms;
    if (search_ms) *search_ms = ix->search_ms;
    return 0;
}

int cphnsw_b200_last_stats(cphnsw_b200_index* ix, cphnsw_b200_stats* out) {
    int rc = require_loaded(ix);
    if (rc) return rc;
    if (!out) return fail(ix, CPHNSW_B200_EINVAL, "null argument");
    Stats s;
    CUDA_TRY(ix, cudaDeviceSynchronize());
    CUDA_TRY(ix, cudaMemcpy(&s, ix->d_stats, sizeof(s), cudaMemcpyDeviceToHost));
    const uint64_t retries = ix->last_stats.overflow_retries;
    out->pops = s.pops; out->expansions = s.expansions; out->exact_calls = s.exact_calls;
    out->beam_pushes = s.beam_pushes; out->max_beam = s.max_beam; out->nn_pushes = s.nn_pushes;
    out->lb_skips = s.lb_skips; out->gamma_terms = s.gamma_terms; out->msb_skipped = s.msb_skipped;
    out->estimated = s.estimated; out->descent_dists = s.descent_dists; out->overflow_retries = retries;
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// kernel-level hooks
// ---------------------------------------------------------------------------------------------------
int cphnsw_b200_prepare_queries(cphnsw_b200_index* ix, const float* d_queries, uint64_t nq, int center, uint8_t* d_lut,
                                float* d_coeffs, float* d_rotated, uint32_t* d_uplanes, void* stream) {
    int rc = require_loaded(ix);
    if (rc) return rc;
    if (nq == 0) return 0;
    if (!d_queries) return fail(ix, CPHNSW_B200_EINVAL, "null buffer");
    // the hook's coeffs are [nq][3]; the kernels' own stride is kCoeffStride, so go through qstate
    QStateView qs;
    rc = ensure_qstate(ix, nq, &qs);
    if (rc) return rc;
    PrepOut po{};
    po.lut = d_lut; po.coeffs = qs.coeffs; po.rotated = d_rotated; po.uplanes = d_uplanes; po.qT = nullptr;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUDA_TRY(ix, launch_query_prep(ix->dev, d_queries, (uint32_t)nq, center, po, st));
    if (d_coeffs)
        CUDA_TRY(ix, cudaMemcpy2DAsync(d_coeffs, 12, qs.coeffs, kCoeffStride * 4, 12, nq, cudaMemcpyDeviceToDevice, st));
    return 0;
}

int cphnsw_b200_fastscan_blocks(cphnsw_b200_index* ix, const uint32_t* d_uplanes, const float* d_coeffs, uint64_t nq,
                                const uint32_t* d_query_of_block, const uint32_t* d_vertex_ids, uint64_t first_vertex,
                                uint64_t nblocks, const float* d_dqp, const int32_t* d_slack_level, uint32_t* d_nbit,
                                uint32_t* d_msb, uint32_t* d_msb2, float* d_est, float* d_lower, float* d_msb_lower,
                                void* stream) {
    int rc = require_loaded(ix);
    if (rc) return rc;
    if (nblocks == 0) return 0;
    if (!d_uplanes || !d_coeffs || !d_dqp || nq == 0) return fail(ix, CPHNSW_B200_EINVAL, "null buffer");
    if (!d_vertex_ids && first_vertex + nblocks > ix->dev.n) return fail(ix, CPHNSW_B200_EINVAL, "vertex range out of bounds");
    // widen the hook's [nq][3] coefficients to the kernels' stride
    QStateView qs;
    rc = ensure_qstate(ix, nq, &qs);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUDA_TRY(ix, cudaMemsetAsync(qs.coeffs, 0, (size_t)nq * kCoeffStride * 4, st));
    CUDA_TRY(ix, cudaMemcpy2DAsync(qs.coeffs, kCoeffStride * 4, d_coeffs, 12, 12, nq, cudaMemcpyDeviceToDevice, st));
    FastScanArgs a{};
    a.uplanes = d_uplanes; a.coeffs = qs.coeffs; a.nq = (uint32_t)nq; a.query_of_block = d_query_of_block;
    a.vertex_ids = d_vertex_ids; a.first_vertex = first_vertex; a.nblocks = nblocks; a.dqp = d_dqp;
    a.slack_level = d_slack_level; a.nbit = d_nbit; a.msb = d_msb; a.msb2 = d_msb2; a.est = d_est; a.lower = d_lower;
    a.msb_lower = d_msb_lower;
    CUDA_TRY(ix, launch_fastscan_blocks(ix->dev, a, ix->num_sms, st));
    return 0;
}

int cphnsw_b200_exact_l2(cphnsw_b200_index* ix, const float* d_queries, uint64_t nq, const uint32_t* d_ids, uint64_t m,
                         float* d_out, void* stream) {
    int rc = require_loaded(ix);
    if (rc) return rc;
    if (nq == 0 || m == 0) return 0;
    if (!d_queries || !d_ids || !d_out) return fail(ix, CPHNSW_B200_EINVAL, "null buffer");
    QStateView qs;
    rc = ensure_qstate(ix, nq, &qs);
    if (rc) return rc;
    PrepOut po{};
    po.coeffs = qs.coeffs; po.qT = qs.qT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUDA_TRY(ix, launch_query_prep(ix->dev, d_queries, (uint32_t)nq, 0, po, st));
    CUDA_TRY(ix, launch_exact_l2(ix->dev, qs.qT, qs.coeffs, (uint32_t)nq, d_ids, (uint32_t)m, d_out, st));
    return 0;
}

int cphnsw_b200_greedy_descent(cphnsw_b200_index* ix, const float* d_queries, uint64_t nq, uint32_t* d_entry,
                               void* stream) {
    int rc = require_loaded(ix);
    if (rc) return rc;
    if (nq == 0) return 0;
    if (!d_queries || !d_entry) return fail(ix, CPHNSW_B200_EINVAL, "null buffer");
    return run_search(ix, d_queries, nq, 1, nullptr, nullptr, d_entry, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

extern "C" {

static int run_exhaustive(cphnsw_b200_index* ix, const float* d_queries, uint64_t nq, uint64_t k, uint64_t kprime,
                          uint64_t id_begin, uint64_t id_end, int64_t* d_ids, float* d_dists, uint32_t* d_sums,
                          float* d_est, cudaStream_t st) {
    const DevIndex& d = ix->dev;
    if (d.B != 1) return fail(ix, CPHNSW_B200_EINVAL, "the exhaustive scan is defined for bits=1 indexes (per-vertex 1-bit codes)");
    if (id_end > d.n || id_begin > id_end) return fail(ix, CPHNSW_B200_EINVAL, "id range out of bounds");
    if (nq == 0) return 0;
    if (nq > 0x7FFFFFFFull || k > 0x7FFFFFFFull) return fail(ix, CPHNSW_B200_EINVAL, "argument too large");
    if (kprime > 1024) return fail(ix, CPHNSW_B200_EINVAL, "kprime (rerank depth) must be <= 1024");
    QStateView qs;
    int rc = ensure_qstate(ix, nq, &qs);
    if (rc) return rc;
    PrepOut po{};
    po.coeffs = qs.coeffs; po.uplanes = qs.uplanes; po.qT = qs.qT; po.ubytes = qs.ubytes;
    CUDA_TRY(ix, cudaMemsetAsync(qs.ubytes, 0, (size_t)nq * d.nch * 128, st));   // padded dimensions carry 0
    CUDA_TRY(ix, launch_query_prep(d, d_queries, (uint32_t)nq, 1, po, st));
    ExhaustiveArgs a{};
    a.uplanes = qs.uplanes; a.coeffs = qs.coeffs; a.qT = qs.qT; a.ubytes = qs.ubytes; a.nq = (uint32_t)nq;
    a.use_tensor_cores = (int)ix->exhaustive_tensor_cores;
    a.id_begin = id_begin; a.id_end = id_end; a.k = (uint32_t)k; a.kprime = (uint32_t)kprime;
    a.sums = d_sums; a.est = d_est; a.ids = d_ids; a.dists = d_dists;
    const size_t wb = exhaustive_workspace_bytes(d, (uint32_t)nq, id_end - id_begin, (uint32_t)kprime, ix->num_sms);
    rc = ensure_buffer(ix, &ix->scratch, &ix->scratch_bytes, wb + 256, false);
    if (rc) return rc;
    a.workspace = ix->scratch; a.workspace_bytes = wb;
    CUDA_TRY(ix, launch_exhaustive(d, a, ix->num_sms, st));
    return 0;
}

int cphnsw_b200_exhaustive_search(cphnsw_b200_index* ix, const float* d_queries, uint64_t nq, uint64_t k, uint64_t kprime,
                                  uint64_t id_begin, uint64_t id_end, int64_t* d_ids, float* d_dists, void* stream) {
    int rc = require_loaded(ix);
    if (rc) return rc;
    if (nq && (!d_queries || (k && (!d_ids || !d_dists)))) return fail(ix, CPHNSW_B200_EINVAL, "null buffer");
    if (kprime < k) kprime = k;
    return run_exhaustive(ix, d_queries, nq, k, kprime, id_begin, id_end, d_ids, d_dists, nullptr, nullptr,
                          static_cast<cudaStream_t>(stream));
}

int cphnsw_b200_exhaustive_estimates(cphnsw_b200_index* ix, const float* d_queries, uint64_t nq, uint64_t id_begin,
                                     uint64_t id_end, uint32_t* d_sums, float* d_est, void* stream) {
    int rc = require_loaded(ix);
    if (rc) return rc;
    if (nq && !d_queries) return fail(ix, CPHNSW_B200_EINVAL, "null buffer");
    return run_exhaustive(ix, d_queries, nq, 0, 0, id_begin, id_end, nullptr, nullptr, d_sums, d_est,
                          static_cast<cudaStream_t>(stream));
}

int cphnsw_b200_unique_topk(cphnsw_b200_index* ix, const int64_t* d_ids_in, const float* d_dists_in, uint64_t nq,
                            uint64_t k_in, uint64_t k_out, const uint32_t* d_id_map, uint64_t map_size,
                            int64_t* d_ids_out, float* d_dists_out, void* stream) {
    if (!ix) return CPHNSW_B200_EINVAL;
    if (nq == 0 || k_out == 0) return 0;
    if (!d_ids_in || !d_dists_in || !d_ids_out || !d_dists_out) return fail(ix, CPHNSW_B200_EINVAL, "null buffer");
    if (k_in > 0x7FFFFFFFull || k_out > 0x7FFFFFFFull) return fail(ix, CPHNSW_B200_EINVAL, "argument too large");
    CUDA_TRY(ix, cudaSetDevice(ix->device));
    CUDA_TRY(ix, launch_unique_topk(d_ids_in, d_dists_in, nq, (uint32_t)k_in, (uint32_t)k_out, d_id_map, map_size, d_ids_out,
                                    d_dists_out, static_cast<cudaStream_t>(stream)));
    return 0;
}

int cphnsw_b200_neighbor_codes(cphnsw_b200_index* ix, uint32_t dim, uint32_t bits, uint64_t rotation_seed,
                               const float* d_vectors, uint64_t row_stride, uint64_t n_vectors,
                               const uint32_t* d_parent_ids, const uint32_t* d_nbr_ids, uint64_t n_parents,
                               uint8_t* d_codes, float* d_aux, uint8_t* d_blocks, uint64_t block_stride, void* stream) {
    if (!ix) return CPHNSW_B200_EINVAL;
    // the checks of the reference's factory (src/bindings.cpp:77-113)
    if (dim == 0 || dim > 2048) return fail(ix, CPHNSW_B200_EINVAL, "dim must be in [1, 2048]");
    if (bits != 1 && bits != 2 && bits != 4) return fail(ix, CPHNSW_B200_EINVAL, "bits must be 1, 2 or 4");
    if (n_parents == 0) return 0;
    if (!d_vectors || !d_nbr_ids || (!d_codes && !d_aux && !d_blocks)) return fail(ix, CPHNSW_B200_EINVAL, "null buffer");
    if (row_stride < dim) return fail(ix, CPHNSW_B200_EINVAL, "row_stride is smaller than dim");
    if (n_vectors >= kInvalid || n_parents > 0x7FFFFFFFull * 8) return fail(ix, CPHNSW_B200_EINVAL, "argument too large");
    CUDA_TRY(ix, cudaSetDevice(ix->device));
    const uint32_t D = next_pow2(dim) < 16 ? 16 : next_pow2(dim);
    if (d_blocks && (block_stride < nb_bytes(D, bits) || block_stride % 4 != 0 || reinterpret_cast<uintptr_t>(d_blocks) % 4 != 0))
        return fail(ix, CPHNSW_B200_EINVAL, "block_stride must be a multiple of 4 and at least the block size (" +
                                                std::to_string(nb_bytes(D, bits)) + " bytes), d_blocks 4-byte aligned");
    if (!ix->enc_signs || ix->enc_signs_D != D || ix->enc_signs_seed != rotation_seed) {
        if (ix->enc_signs) { cudaFree(ix->enc_signs); ix->enc_signs = nullptr; }
        const std::vector<float> signs = rotation_signs(D, rotation_seed);
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&ix->enc_signs), signs.size() * sizeof(float));
        if (e != cudaSuccess) { ix->enc_signs = nullptr; return fail(ix, CPHNSW_B200_ENOMEM, std::string("cudaMalloc(signs): ") + cudaGetErrorString(e)); }
        CUDA_TRY(ix, cudaMemcpy(ix->enc_signs, signs.data(), signs.size() * sizeof(float), cudaMemcpyHostToDevice));
        ix->enc_signs_D = D;
        ix->enc_signs_seed = rotation_seed;
    }
    NeighborCodesArgs a{};
    a.D = D; a.dim = dim; a.signs = ix->enc_signs;
    a.vectors = d_vectors; a.row_stride = row_stride; a.n_vectors = n_vectors;
    a.parent_ids = d_parent_ids; a.nbr_ids = d_nbr_ids; a.n_parents = n_parents;
    a.codes = d_codes; a.aux = d_aux; a.blocks = d_blocks; a.block_stride = block_stride;
    // per-warp tiles in a scratch buffer (L2) unless the option asks for shared memory: measured faster at every D
    // (profiles/n3_gpu_check_r01.log), shared memory caps the warps per SM
    const bool global_tile = ix->neighbor_codes_tile != 1;
    NeighborCodesPlan plan{};
    if (neighbor_codes_plan(a, bits, ix->num_sms, global_tile, &plan) != cudaSuccess)
        return fail(ix, CPHNSW_B200_EINVAL, "no launch shape for this dimension");
    if (plan.scratch_bytes) {
        int rc = ensure_buffer(ix, &ix->enc_scratch, &ix->enc_scratch_bytes, plan.scratch_bytes, false);
        if (rc) return rc;
        a.tile_x = static_cast<float*>(ix->enc_scratch);
        a.tile_u = reinterpret_cast<uint8_t*>(a.tile_x + (size_t)a.total_warps * D * 32);
    }
    CUDA_TRY(ix, launch_neighbor_codes(a, bits, plan, static_cast<cudaStream_t>(stream)));
    return 0;
}

}  // extern "C"
