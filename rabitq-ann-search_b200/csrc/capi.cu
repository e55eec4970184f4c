// C ABI of the library (include/cphnsw_b200.h): index hand-off (save-file v2 reader, upload and
// re-layout), the search entry points and the kernel-level hooks.  Host-side glue only; the
// kernels are in query_prep.cu, fastscan_blocks.cu, search.cu, exhaustive.cu, relayout.cu.
//
// Concurrency.  The index itself is read-only after upload; everything a call writes lives in a Lane
// (device_index.h).  A call takes the next lane round-robin, enqueues its work and leaves the lane "in flight";
// the lane's next user first waits for that work (one event) -- so calls never share scratch, two host threads
// can search one handle at the same time (as with the reference: api/hnsw_index.hpp:172), and nothing in the
// search path synchronises the host with the device except the entry points that return host data.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <new>
#include <random>
#include <string>
#include <vector>

#include "device_index.h"
#include "kernels.h"

using namespace cpb;

namespace {

thread_local std::string g_create_error;   // errors without a handle (create)
thread_local std::string g_last_error;     // what cphnsw_b200_last_error hands out (a copy: the handle's string may change)

int fail(cphnsw_b200_index* ix, int code, const std::string& msg) {
    if (ix) { std::lock_guard<std::mutex> g(ix->mu); ix->err = msg; }
    else g_create_error = msg;
    return code;
}
// same, for code that already holds ix->mu
int fail_locked(cphnsw_b200_index* ix, int code, const std::string& msg) { ix->err = msg; return code; }

#define CUDA_TRY(ix, expr)                                                                         \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return fail(ix, CPHNSW_B200_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

// No C++ exception crosses the C boundary (std::bad_alloc from a corrupt file's sizes, ...).
template <class F>
int guarded(cphnsw_b200_index* ix, F&& body) {
    try {
        return body();
    } catch (const std::bad_alloc&) {
        return fail(ix, CPHNSW_B200_ENOMEM, "out of host memory");
    } catch (const std::exception& e) {
        return fail(ix, CPHNSW_B200_ERUNTIME, std::string("internal error: ") + e.what());
    } catch (...) {
        return fail(ix, CPHNSW_B200_ERUNTIME, "internal error");
    }
}

// The caller's current device is restored on every exit path (a process whose torch device differs from the
// index's must not find it switched).
struct DeviceGuard {
    int prev = -1;
    cudaError_t err;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        err = prev == dev ? cudaSuccess : cudaSetDevice(dev);
        if (prev == dev) prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

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

// ---- lanes ------------------------------------------------------------------------------------------
void destroy_lane(Lane& L) {
    if (L.scratch) cudaFree(L.scratch);
    if (L.bitmaps) cudaFree(L.bitmaps);
    if (L.qstate) cudaFree(L.qstate);
    if (L.stage) cudaFree(L.stage);
    if (L.d_counters) cudaFree(L.d_counters);   // d_stats lives in the same allocation
    if (L.h_counters) cudaFreeHost(L.h_counters);
    if (L.stream) cudaStreamDestroy(L.stream);
    for (auto& e : L.ev) if (e) cudaEventDestroy(e);
    if (L.done) cudaEventDestroy(L.done);
    L = Lane{};
}

int create_lane(cphnsw_b200_index* ix, Lane& L) {
    if (L.created) return 0;
    bool bad = cudaMalloc(reinterpret_cast<void**>(&L.d_counters), 32 + sizeof(Stats)) != cudaSuccess ||   // one memset clears both
               cudaMallocHost(reinterpret_cast<void**>(&L.h_counters), 32) != cudaSuccess ||
               cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking) != cudaSuccess ||
               cudaEventCreateWithFlags(&L.done, cudaEventDisableTiming) != cudaSuccess;
    for (auto& e : L.ev) bad = bad || cudaEventCreate(&e) != cudaSuccess;
    if (bad) { cudaGetLastError(); destroy_lane(L); return fail(ix, CPHNSW_B200_ECUDA, "could not allocate the control buffers of a lane"); }
    std::memset(L.h_counters, 0, 32);
    L.d_stats = reinterpret_cast<Stats*>(L.d_counters + 8);
    cudaMemset(L.d_counters, 0, 32 + sizeof(Stats));
    L.created = true;
    return 0;
}

// Wait for the lane's in-flight work and fold its deferred outcome (overflow counters) into the lane.  The caller
// holds the lane (state 1).
int finish_lane(cphnsw_b200_index* ix, Lane& L) {
    if (!L.created) return 0;
    cudaError_t e = cudaEventSynchronize(L.done);
    if (e != cudaSuccess) {
        L.status = CPHNSW_B200_ECUDA; L.status_msg = std::string("search: ") + cudaGetErrorString(e);
        L.counters_pending = false;
        return fail(ix, L.status, L.status_msg);
    }
    if (L.counters_pending) {
        L.counters_pending = false;
        L.overflow_retries = L.h_counters[1];
        if (L.h_counters[3] != 0) {
            L.status = CPHNSW_B200_ERUNTIME; L.status_msg = "internal error: frontier overflow with a full-size arena";
            return fail(ix, L.status, L.status_msg);
        }
    }
    return 0;
}

// Take the next lane (round-robin).  If its previous call is still in flight, wait for it first: a deferred error of
// that call is reported here, once.
int acquire_lane(cphnsw_b200_index* ix, Lane** out) {
    int li;
    bool inflight;
    {
        std::unique_lock<std::mutex> lk(ix->mu);
        li = (int)(ix->next_lane++ % kLanes);
        ix->cv.wait(lk, [&] { return ix->lanes[li].state != 1; });
        inflight = ix->lanes[li].state == 2;
        ix->lanes[li].state = 1;
    }
    Lane& L = ix->lanes[li];
    int rc = create_lane(ix, L);
    if (rc == 0 && inflight) rc = finish_lane(ix, L);
    if (rc) {
        std::lock_guard<std::mutex> g(ix->mu);
        L.state = 0; L.status = 0;
        ix->cv.notify_all();
        return rc;
    }
    L.status = 0;
    L.timed = false;
    *out = &L;
    return 0;
}

// Hand the lane back.  inflight: work was enqueued and `done` recorded on its stream.
void release_lane(cphnsw_b200_index* ix, Lane* L, bool inflight, bool is_search = false) {
    std::lock_guard<std::mutex> g(ix->mu);
    L->state = inflight ? 2 : 0;
    if (is_search) ix->last_lane = (int)(L - ix->lanes);
    ix->cv.notify_all();
}

// Every lane idle (state 0), all deferred outcomes folded in.  Returns the first deferred error.
int drain_lanes(cphnsw_b200_index* ix) {
    int first = 0;
    for (int i = 0; i < kLanes; ++i) {
        Lane& L = ix->lanes[i];
        bool inflight;
        {
            std::unique_lock<std::mutex> lk(ix->mu);
            ix->cv.wait(lk, [&] { return L.state != 1; });
            inflight = L.state == 2;
            if (inflight) L.state = 1;
        }
        if (!inflight) continue;
        const int rc = finish_lane(ix, L);
        if (rc && !first) first = rc;
        std::lock_guard<std::mutex> g(ix->mu);
        L.state = 0;
        ix->cv.notify_all();
    }
    return first;
}

void release_index(cphnsw_b200_index* ix) {
    drain_lanes(ix);
    for (void* p : ix->allocs) cudaFree(p);
    ix->allocs.clear();
    ix->device_bytes = 0;
    ix->loaded = false;
    ix->dev = DevIndex{};
    ix->frontier_budget = 0;
    ix->occ_key[0] = 0;
    // the bitmap arenas are laid out for one n
    for (auto& L : ix->lanes)
        if (L.bitmaps) { cudaFree(L.bitmaps); L.bitmaps = nullptr; L.bitmaps_bytes = 0; }
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

int ensure_qstate(cphnsw_b200_index* ix, Lane& L, uint64_t nq, QStateView* v) {
    const DevIndex& d = ix->dev;
    const size_t qT = (size_t)nq * d.D * 4, up = (size_t)nq * 16 * d.nch * 4, co = (size_t)nq * kCoeffStride * 4;
    const size_t a = (qT + 255) & ~(size_t)255, b = (up + 255) & ~(size_t)255, c = (co + 255) & ~(size_t)255;
    const size_t ub = (size_t)nq * d.nch * 128;
    int rc = ensure_buffer(ix, &L.qstate, &L.qstate_bytes, a + b + c + ub + 256, false);
    if (rc) return rc;
    uint8_t* p = static_cast<uint8_t*>(L.qstate);
    v->qT = reinterpret_cast<float*>(p);
    v->uplanes = reinterpret_cast<uint32_t*>(p + a);
    v->coeffs = reinterpret_cast<float*>(p + a + b);
    v->ubytes = p + a + b + c;
    return 0;
}

int require_loaded(cphnsw_b200_index* ix) {
    if (!ix) return fail(nullptr, CPHNSW_B200_EINVAL, "null index handle");
    if (!ix->loaded) return fail(ix, CPHNSW_B200_ERUNTIME, "Index has no finalized data on the device (load or upload first).");
    return 0;
}

// A call that works in a lane: device guard, lane acquisition, `body(lane)` enqueues on `stream`, then the lane is
// left in flight behind an event on that stream.
template <class F>
int with_lane(cphnsw_b200_index* ix, cudaStream_t stream, bool is_search, F&& body) {
    DeviceGuard dg(ix->device);
    if (dg.err != cudaSuccess) return fail(ix, CPHNSW_B200_ECUDA, cudaGetErrorString(dg.err));
    Lane* L = nullptr;
    int rc = acquire_lane(ix, &L);
    if (rc) return rc;
    rc = guarded(ix, [&] { return body(*L); });
    const bool recorded = cudaEventRecord(L->done, stream) == cudaSuccess;
    if (!recorded) cudaGetLastError();
    release_lane(ix, L, recorded, is_search && rc == 0);
    return rc;
}

}  // namespace

extern "C" {

int cphnsw_b200_create(int device, cphnsw_b200_index** out) {
    if (!out) return fail(nullptr, CPHNSW_B200_EINVAL, "out is null");
    return guarded(nullptr, [&]() -> int {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            return fail(nullptr, CPHNSW_B200_ECUDA,
                        std::string("no CUDA device: this library has no CPU path (") + cudaGetErrorString(e) + ")");
        if (device < 0 || device >= ndev) return fail(nullptr, CPHNSW_B200_EINVAL, "device ordinal out of range");
        DeviceGuard dg(device);
        if (dg.err != cudaSuccess) return fail(nullptr, CPHNSW_B200_ECUDA, cudaGetErrorString(dg.err));
        cudaDeviceProp prop;
        e = cudaGetDeviceProperties(&prop, device);
        if (e != cudaSuccess) return fail(nullptr, CPHNSW_B200_ECUDA, cudaGetErrorString(e));
        if (prop.major != 10)
            return fail(nullptr, CPHNSW_B200_ECUDA, "this build carries sm_100a code only (B200); found sm_" +
                                                        std::to_string(prop.major) + std::to_string(prop.minor));
        auto* ix = new cphnsw_b200_index();
        ix->device = device;
        ix->num_sms = prop.multiProcessorCount;
        if (cudaMalloc(reinterpret_cast<void**>(&ix->d_problems), 16) != cudaSuccess) {
            delete ix;
            return fail(nullptr, CPHNSW_B200_ECUDA, "could not allocate control buffers");
        }
        *out = ix;
        return 0;
    });
}

void cphnsw_b200_destroy(cphnsw_b200_index* ix) {
    if (!ix) return;
    try {
        DeviceGuard dg(ix->device);
        release_index(ix);
        for (auto& L : ix->lanes) destroy_lane(L);
        if (ix->d_problems) cudaFree(ix->d_problems);
        if (ix->enc_signs) cudaFree(ix->enc_signs);
        if (ix->enc_scratch) cudaFree(ix->enc_scratch);
        if (ix->enc_done) cudaEventDestroy(ix->enc_done);
    } catch (...) {
    }
    delete ix;
}

const char* cphnsw_b200_last_error(const cphnsw_b200_index* ix) {
    if (!ix) return g_create_error.c_str();
    auto* m = const_cast<cphnsw_b200_index*>(ix);
    std::lock_guard<std::mutex> g(m->mu);
    g_last_error = ix->err;
    return g_last_error.c_str();
}

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

static int upload_impl(cphnsw_b200_index* ix, const cphnsw_b200_host_index* h) {
    DeviceGuard dg(ix->device);
    CUDA_TRY(ix, dg.err);
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
    CUDA_TRY(ix, cudaMemset(ix->d_problems, 0, 16));
    const size_t chunk_bytes = (size_t)256 << 20;
    void* stage = nullptr;
    CUDA_TRY(ix, cudaMalloc(&stage, chunk_bytes));
    auto cleanup = [&](int code) { cudaFree(stage); release_index(ix); return code; };
    {
        const uint32_t per = (uint32_t)std::max<size_t>(1, chunk_bytes / h->rec_size);
        for (uint64_t first = 0; first < d.n; first += per) {
            const uint32_t cnt = (uint32_t)std::min<uint64_t>(per, d.n - first);
            cudaError_t e = cudaMemcpy(stage, h->search_data + first * h->rec_size, (size_t)cnt * h->rec_size, cudaMemcpyHostToDevice);
            if (e == cudaSuccess) e = launch_relayout_blocks(d, static_cast<const uint8_t*>(stage), h->rec_size, h->nb_off, first, cnt, ix->d_problems + 2, 0);
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
    CUDA_TRY(ix, cudaMemcpy(problems, ix->d_problems, 16, cudaMemcpyDeviceToHost));
    if (problems[3]) { release_index(ix); return fail(ix, CPHNSW_B200_ERUNTIME, "Index is corrupt: neighbour id out of range in " + std::to_string(problems[3]) + " blocks."); }
    ix->dev.dup_neighbors = problems[2];
    ix->loaded = true;
    return 0;
}

int cphnsw_b200_upload(cphnsw_b200_index* ix, const cphnsw_b200_host_index* h) {
    if (!ix || !h) return fail(ix, CPHNSW_B200_EINVAL, "null argument");
    return guarded(ix, [&] { return upload_impl(ix, h); });
}

static int load_impl(cphnsw_b200_index* ix, const char* path) {
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(ix, CPHNSW_B200_ERUNTIME, std::string("Cannot open file for reading: ") + path);
    struct stat sb;
    if (fstat(fd, &sb) != 0 || sb.st_size < 68 + 248 + 72) { close(fd); return fail(ix, CPHNSW_B200_ERUNTIME, "Index file is truncated."); }
    const size_t fsize = (size_t)sb.st_size;
    void* map = mmap(nullptr, fsize, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (map == MAP_FAILED) return fail(ix, CPHNSW_B200_ERUNTIME, "mmap of the index file failed");
    const uint8_t* p = static_cast<const uint8_t*>(map);
    struct Unmap { void* p; size_t n; ~Unmap() { munmap(p, n); } } unmap{map, fsize};   // also when an exception unwinds
    auto done = [](int code) { return code; };

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
    // n comes from the file: bound it by what the file can hold before any size is computed from it (a vertex takes
    // at least 8 + 4 D + rec_size bytes), so none of the products below can wrap
    const uint64_t per_vertex = 8 + (uint64_t)4 * h.D + h.rec_size;
    if (h.n == 0 || h.n >= 0xFFFFFFFFull || h.n > fsize / per_vertex) return done(fail(ix, CPHNSW_B200_ERUNTIME, "Index file is truncated."));
    const size_t need = off + (size_t)4 * h.dim + (size_t)8 * h.n + (size_t)4 * h.n * h.D + (size_t)h.n * h.rec_size + 4;
    if (need > fsize) return done(fail(ix, CPHNSW_B200_ERUNTIME, "Index file is truncated."));
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
        if ((uint64_t)ne > (fsize - off) / 8) return done(fail(ix, CPHNSW_B200_ERUNTIME, "Index file is truncated."));   // 8 bytes per edge record at least
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
    return done(upload_impl(ix, &h));
}

int cphnsw_b200_load(cphnsw_b200_index* ix, const char* path) {
    if (!ix || !path) return fail(ix, CPHNSW_B200_EINVAL, "null argument");
    return guarded(ix, [&] { return load_impl(ix, path); });
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
// Enqueue K1 + K3 (+ the overflow re-run) for one batch on `stream`, in lane L.  Nothing here waits for the device.
static int run_search(cphnsw_b200_index* ix, Lane& L, const float* d_queries, uint64_t nq, uint64_t k_user, int64_t* d_ids,
                      float* d_dists, uint32_t* d_entry_out, cudaStream_t stream) {
    const DevIndex& d = ix->dev;
    if (nq == 0) return 0;
    if (nq > 0x7FFFFFFFull) return fail(ix, CPHNSW_B200_EINVAL, "too many queries in one batch");
    if (k_user > 0x7FFFFFFFull) return fail(ix, CPHNSW_B200_EINVAL, "k too large");
    const uint32_t k = (uint32_t)std::max<uint64_t>(k_user, 1);   // hnsw_index.hpp:187

    QStateView qs;
    int rc = ensure_qstate(ix, L, nq, &qs);
    if (rc) return rc;
    PrepOut po{};
    po.coeffs = qs.coeffs; po.uplanes = qs.uplanes; po.qT = qs.qT;

    // launch geometry: persistent grid, one warp per in-flight query
    const bool stats = ix->collect_stats != 0;
    int warps = (int)ix->warps_per_cta;
    const size_t smem_budget = 220 * 1024;
    const size_t spw = search_smem_per_warp(d, k, stats);
    while (warps > 1 && spw * warps > smem_budget) --warps;
    int per_sm;
    {   // the occupancy query costs tens of microseconds: once per (shared memory per warp, warps, stats)
        std::lock_guard<std::mutex> g(ix->mu);
        if (ix->occ_key[0] != (int)spw || ix->occ_key[1] != warps || ix->occ_key[2] != (int)stats) {
            ix->occ_key[3] = search_max_ctas_per_sm(d, k, warps, stats);
            ix->occ_key[0] = (int)spw; ix->occ_key[1] = warps; ix->occ_key[2] = (int)stats;
        }
        per_sm = ix->occ_key[3];
    }
    if (per_sm <= 0) return fail(ix, CPHNSW_B200_ECUDA, "search kernel cannot be resident (shared memory / registers)");
    per_sm = std::min<int>(per_sm, (int)ix->ctas_per_sm);
    int ctas = ix->num_sms * per_sm;
    const int need = (int)((nq + warps - 1) / warps);
    if (ctas > need) ctas = need;

    auto layout = [&](uint32_t cap, SearchArgs& a) {
        a.bitmap_words = (uint32_t)(((d.n + 31) / 32 + 31) & ~(uint64_t)31);   // one bit per vertex, whole 128-byte lines per slot
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
    // frontier arena per slot: the option if set, else as much as this lane's share of HBM (4 GiB, or a sixteenth of
    // what is free) gives every slot -- a frontier can hold at most n entries, and a query whose frontier outgrows
    // its arena is run again from scratch, so be generous where memory allows (1-bit indexes reach 40k+)
    const size_t slots = (size_t)ctas * warps;
    uint64_t cap64 = (uint64_t)ix->beam_capacity;
    if (cap64 == 0) {
        std::lock_guard<std::mutex> g(ix->mu);
        if (ix->frontier_budget == 0) {   // once per index: the answer must not drift from call to call
            size_t free_b = 0, total_b = 0;
            cudaMemGetInfo(&free_b, &total_b);
            size_t held = 0;
            for (const auto& l : ix->lanes) held += l.scratch_bytes;
            ix->frontier_budget = std::min<size_t>((size_t)4 << 30, (free_b + held) / (8 * kLanes));
        }
        cap64 = std::max<uint64_t>(4096, ix->frontier_budget / (slots * 16));
    }
    const uint32_t cap_full = (uint32_t)d.n + 1;   // each id enters the frontier at most once: this cannot overflow
    const uint32_t cap = (uint32_t)std::min<uint64_t>(cap64, cap_full);
    layout(cap, a);
    // re-run of overflowed queries, enqueued behind the first pass with no host round trip: the same kernel over the
    // device-side list of overflowed queries, with full-size arenas carved from the same scratch (the first pass is
    // complete by then), as many slots as fit
    const bool may_overflow = cap < cap_full && !d_entry_out;
    SearchArgs b = a;
    size_t rslots = 0;
    if (may_overflow) layout(cap_full, b);
    const size_t list_bytes = ((size_t)nq * 4 + 255) & ~(size_t)255;
    size_t arena_bytes = slots * a.slot_stride;
    if (may_overflow) {
        arena_bytes = std::max(arena_bytes, b.slot_stride);
        rslots = std::min<size_t>({arena_bytes / b.slot_stride, (size_t)256, (size_t)nq});
    }
    rc = ensure_buffer(ix, &L.scratch, &L.scratch_bytes, 2 * list_bytes + arena_bytes, false);
    if (rc) return rc;
    int rwarps = 0, rctas = 0;
    if (may_overflow) {
        rwarps = (int)std::min<size_t>(warps, rslots);
        rctas = (int)std::max<size_t>(1, rslots / rwarps);
        rslots = (size_t)rwarps * rctas;
    }
    rc = ensure_buffer(ix, &L.bitmaps, &L.bitmaps_bytes, std::max(slots, rslots) * (size_t)a.bitmap_words * 4, true);
    if (rc) return rc;
    a.bitmaps = static_cast<uint32_t*>(L.bitmaps);
    a.overflow_list = static_cast<uint32_t*>(L.scratch);
    a.scratch = static_cast<uint8_t*>(L.scratch) + 2 * list_bytes;
    a.counters = L.d_counters;
    a.stats = L.d_stats;
    CUDA_TRY(ix, cudaMemsetAsync(L.d_counters, 0, 32 + sizeof(Stats), stream));
    CUDA_TRY(ix, cudaEventRecord(L.ev[0], stream));
    CUDA_TRY(ix, launch_query_prep(d, d_queries, (uint32_t)nq, 0, po, stream));
    CUDA_TRY(ix, cudaEventRecord(L.ev[1], stream));
    CUDA_TRY(ix, cudaEventRecord(L.ev[2], stream));
    CUDA_TRY(ix, launch_search(d, a, ctas, warps, stats, stream));
    if (may_overflow) {
        b.bitmaps = a.bitmaps;   // the first pass leaves every bitmap clean
        b.scratch = a.scratch;
        b.query_list = a.overflow_list;                                   // what the first pass listed ...
        b.nq_ptr = L.d_counters + 1;                                      // ... and how many
        b.nq = (uint32_t)nq;
        b.overflow_list = a.overflow_list + list_bytes / 4;               // (cannot happen; checked when the lane is next used)
        b.counters = L.d_counters + 2;
        CUDA_TRY(ix, launch_search(d, b, rctas, rwarps, stats, stream));
    }
    CUDA_TRY(ix, cudaEventRecord(L.ev[3], stream));
    CUDA_TRY(ix, cudaMemcpyAsync(L.h_counters, L.d_counters, 32, cudaMemcpyDeviceToHost, stream));
    L.counters_pending = true;
    L.timed = true;
    L.launches = 2 + (may_overflow ? 1 : 0);
    return 0;
}

int cphnsw_b200_search_batch_device(cphnsw_b200_index* ix, const float* d_queries, uint64_t nq, uint64_t k,
                                    int64_t* d_ids, float* d_dists, void* stream) {
    int rc = require_loaded(ix);
    if (rc) return rc;
    if (nq && (!d_queries || (k && (!d_ids || !d_dists)))) return fail(ix, CPHNSW_B200_EINVAL, "null buffer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return with_lane(ix, st, true, [&](Lane& L) { return run_search(ix, L, d_queries, nq, k, d_ids, d_dists, nullptr, st); });
}

// Host buffers -> lane staging -> K1/K3 -> host buffers, all on the lane's own stream; returns with the lane held.
static int enqueue_host_search(cphnsw_b200_index* ix, Lane& L, const float* queries, uint64_t nq, uint64_t k, int64_t* ids,
                               float* dists) {
    const DevIndex& d = ix->dev;
    const size_t qb = ((size_t)nq * d.dim * 4 + 255) & ~(size_t)255, ib = ((size_t)nq * k * 8 + 255) & ~(size_t)255,
                 db = (size_t)nq * k * 4;
    int rc = ensure_buffer(ix, &L.stage, &L.stage_bytes, qb + ib + db + 256, false);
    if (rc) return rc;
    uint8_t* s = static_cast<uint8_t*>(L.stage);
    float* d_q = reinterpret_cast<float*>(s);
    int64_t* d_i = reinterpret_cast<int64_t*>(s + qb);
    float* d_d = reinterpret_cast<float*>(s + qb + ib);
    cudaStream_t st = L.stream;
    CUDA_TRY(ix, cudaMemcpyAsync(d_q, queries, (size_t)nq * d.dim * 4, cudaMemcpyHostToDevice, st));
    rc = run_search(ix, L, d_q, nq, k, d_i, d_d, nullptr, st);
    if (rc) return rc;
    if (k) {
        CUDA_TRY(ix, cudaMemcpyAsync(ids, d_i, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(ix, cudaMemcpyAsync(dists, d_d, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, st));
    }
    return 0;
}

int cphnsw_b200_search_batch_submit(cphnsw_b200_index* ix, const float* queries, uint64_t nq, uint64_t k, int64_t* ids,
                                    float* dists, uint64_t* ticket) {
    int rc = require_loaded(ix);
    if (rc) return rc;
    if (!ticket) return fail(ix, CPHNSW_B200_EINVAL, "null argument");
    if (nq && (!queries || (k && (!ids || !dists)))) return fail(ix, CPHNSW_B200_EINVAL, "null buffer");
    DeviceGuard dg(ix->device);
    CUDA_TRY(ix, dg.err);
    Lane* L = nullptr;
    rc = acquire_lane(ix, &L);
    if (rc) return rc;
    rc = nq ? guarded(ix, [&] { return enqueue_host_search(ix, *L, queries, nq, k, ids, dists); }) : 0;
    const bool recorded = cudaEventRecord(L->done, L->stream) == cudaSuccess;
    {
        std::lock_guard<std::mutex> g(ix->mu);
        L->ticket = ++ix->next_ticket;
        *ticket = L->ticket;
        L->state = recorded ? 2 : 0;
        if (rc == 0) ix->last_lane = (int)(L - ix->lanes);
        ix->cv.notify_all();
    }
    return rc;
}

int cphnsw_b200_search_batch_wait(cphnsw_b200_index* ix, uint64_t ticket) {
    if (!ix) return fail(nullptr, CPHNSW_B200_EINVAL, "null index handle");
    Lane* L = nullptr;
    {
        std::unique_lock<std::mutex> lk(ix->mu);
        if (ticket == 0 || ticket > ix->next_ticket) return fail_locked(ix, CPHNSW_B200_EINVAL, "unknown ticket");
        for (auto& l : ix->lanes) if (l.ticket == ticket) L = &l;
        if (!L) return 0;   // its lane has been reused since: that user waited for it (and reported its outcome)
        ix->cv.wait(lk, [&] { return L->state != 1 || L->ticket != ticket; });
        if (L->ticket != ticket || L->state == 0) return L->ticket == ticket ? L->status : 0;
        L->state = 1;
    }
    DeviceGuard dg(ix->device);
    const int rc = finish_lane(ix, *L);
    release_lane(ix, L, false);
    return rc;
}

int cphnsw_b200_search_batch(cphnsw_b200_index* ix, const float* queries, uint64_t nq, uint64_t k, int64_t* ids,
                             float* dists) {
    int rc = require_loaded(ix);
    if (rc) return rc;
    if (nq == 0) return 0;
    uint64_t ticket = 0;
    rc = cphnsw_b200_search_batch_submit(ix, queries, nq, k, ids, dists, &ticket);
    const int rw = ticket ? cphnsw_b200_search_batch_wait(ix, ticket) : 0;
    return rc ? rc : rw;
}

int cphnsw_b200_synchronize(cphnsw_b200_index* ix) {
    if (!ix) return fail(nullptr, CPHNSW_B200_EINVAL, "null index handle");
    DeviceGuard dg(ix->device);
    CUDA_TRY(ix, dg.err);
    return drain_lanes(ix);
}

// The lane of the most recent search, idle (its deferred outcome folded in), held by the caller.
static int hold_last_lane(cphnsw_b200_index* ix, Lane** out) {
    std::unique_lock<std::mutex> lk(ix->mu);
    if (ix->last_lane < 0) return fail_locked(ix, CPHNSW_B200_ERUNTIME, "no search has run on this handle yet");
    Lane& L = ix->lanes[ix->last_lane];
    ix->cv.wait(lk, [&] { return L.state != 1; });
    const bool inflight = L.state == 2;
    L.state = 1;
    lk.unlock();
    int rc = inflight ? finish_lane(ix, L) : 0;
    if (rc) { release_lane(ix, &L, false); return rc; }
    *out = &L;
    return 0;
}

int cphnsw_b200_last_timings(cphnsw_b200_index* ix, float* prep_ms, float* search_ms) {
    if (!ix) return CPHNSW_B200_EINVAL;
    DeviceGuard dg(ix->device);
    Lane* L = nullptr;
    int rc = hold_last_lane(ix, &L);
    if (rc) return rc;
    float p = 0.0f, s = 0.0f;
    if (L->timed) {
        cudaEventSynchronize(L->ev[3]);
        cudaEventElapsedTime(&p, L->ev[0], L->ev[1]);
        cudaEventElapsedTime(&s, L->ev[2], L->ev[3]);
    }
    release_lane(ix, L, false);
    if (prep_ms) *prep_ms = p;
    if (search_ms) *search_ms = s;
    return 0;
}

int cphnsw_b200_last_stats(cphnsw_b200_index* ix, cphnsw_b200_stats* out) {
    int rc = require_loaded(ix);
    if (rc) return rc;
    if (!out) return fail(ix, CPHNSW_B200_EINVAL, "null argument");
    DeviceGuard dg(ix->device);
    Lane* L = nullptr;
    rc = hold_last_lane(ix, &L);
    if (rc) return rc;
    Stats s;
    const cudaError_t e = cudaMemcpy(&s, L->d_stats, sizeof(s), cudaMemcpyDeviceToHost);
    const uint64_t retries = L->overflow_retries, launches = L->launches;
    release_lane(ix, L, false);
    if (e != cudaSuccess) return fail(ix, CPHNSW_B200_ECUDA, cudaGetErrorString(e));
    out->pops = s.pops; out->expansions = s.expansions; out->exact_calls = s.exact_calls;
    out->beam_pushes = s.beam_pushes; out->max_beam = s.max_beam; out->nn_pushes = s.nn_pushes;
    out->lb_skips = s.lb_skips; out->gamma_terms = s.gamma_terms; out->msb_skipped = s.msb_skipped;
    out->estimated = s.estimated; out->descent_dists = s.descent_dists; out->overflow_retries = retries; out->kernel_launches = launches;
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
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return with_lane(ix, st, false, [&](Lane& L) -> int {
        // the hook's coeffs are [nq][3]; the kernels' own stride is kCoeffStride, so go through qstate
        QStateView qs;
        int r = ensure_qstate(ix, L, nq, &qs);
        if (r) return r;
        PrepOut po{};
        po.lut = d_lut; po.coeffs = qs.coeffs; po.rotated = d_rotated; po.uplanes = d_uplanes; po.qT = nullptr;
        CUDA_TRY(ix, launch_query_prep(ix->dev, d_queries, (uint32_t)nq, center, po, st));
        if (d_coeffs)
            CUDA_TRY(ix, cudaMemcpy2DAsync(d_coeffs, 12, qs.coeffs, kCoeffStride * 4, 12, nq, cudaMemcpyDeviceToDevice, st));
        return 0;
    });
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
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return with_lane(ix, st, false, [&](Lane& L) -> int {
        // widen the hook's [nq][3] coefficients to the kernels' stride
        QStateView qs;
        int r = ensure_qstate(ix, L, nq, &qs);
        if (r) return r;
        CUDA_TRY(ix, cudaMemsetAsync(qs.coeffs, 0, (size_t)nq * kCoeffStride * 4, st));
        CUDA_TRY(ix, cudaMemcpy2DAsync(qs.coeffs, kCoeffStride * 4, d_coeffs, 12, 12, nq, cudaMemcpyDeviceToDevice, st));
        FastScanArgs a{};
        a.uplanes = d_uplanes; a.coeffs = qs.coeffs; a.nq = (uint32_t)nq; a.query_of_block = d_query_of_block;
        a.vertex_ids = d_vertex_ids; a.first_vertex = first_vertex; a.nblocks = nblocks; a.dqp = d_dqp;
        a.slack_level = d_slack_level; a.nbit = d_nbit; a.msb = d_msb; a.msb2 = d_msb2; a.est = d_est; a.lower = d_lower;
        a.msb_lower = d_msb_lower;
        CUDA_TRY(ix, launch_fastscan_blocks(ix->dev, a, ix->num_sms, st));
        return 0;
    });
}

int cphnsw_b200_exact_l2(cphnsw_b200_index* ix, const float* d_queries, uint64_t nq, const uint32_t* d_ids, uint64_t m,
                         float* d_out, void* stream) {
    int rc = require_loaded(ix);
    if (rc) return rc;
    if (nq == 0 || m == 0) return 0;
    if (!d_queries || !d_ids || !d_out) return fail(ix, CPHNSW_B200_EINVAL, "null buffer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return with_lane(ix, st, false, [&](Lane& L) -> int {
        QStateView qs;
        int r = ensure_qstate(ix, L, nq, &qs);
        if (r) return r;
        PrepOut po{};
        po.coeffs = qs.coeffs; po.qT = qs.qT;
        CUDA_TRY(ix, launch_query_prep(ix->dev, d_queries, (uint32_t)nq, 0, po, st));
        CUDA_TRY(ix, launch_exact_l2(ix->dev, qs.qT, qs.coeffs, (uint32_t)nq, d_ids, (uint32_t)m, d_out, st));
        return 0;
    });
}

int cphnsw_b200_calibration_samples(cphnsw_b200_index* ix, const float* d_queries, const uint32_t* d_start_ids, uint64_t ns,
                                    uint32_t* d_parent, float* d_nn_dist_sq, float* d_dist_qp_sq, float* d_nop, float* d_ip_corrected,
                                    float* d_ip_qo_denom, float* d_true_ip, uint32_t* d_neighbor, void* stream) {
    int rc = require_loaded(ix);
    if (rc) return rc;
    if (ns == 0) return 0;
    if (!d_queries || !d_start_ids || !d_parent || !d_nn_dist_sq || !d_dist_qp_sq || !d_nop || !d_ip_corrected || !d_ip_qo_denom ||
        !d_true_ip || !d_neighbor)
        return fail(ix, CPHNSW_B200_EINVAL, "null buffer");
    if (ns > 0x7FFFFFFFull) return fail(ix, CPHNSW_B200_EINVAL, "too many samples in one call");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return with_lane(ix, st, false, [&](Lane& L) -> int {
        QStateView qs;
        int r = ensure_qstate(ix, L, ns, &qs);
        if (r) return r;
        PrepOut po{};
        po.coeffs = qs.coeffs; po.uplanes = qs.uplanes; po.qT = qs.qT;
        CUDA_TRY(ix, launch_query_prep(ix->dev, d_queries, (uint32_t)ns, 0, po, st));
        CalibrationArgs a{};
        a.queries = d_queries; a.start_ids = d_start_ids; a.ns = ns; a.qT = qs.qT; a.uplanes = qs.uplanes; a.coeffs = qs.coeffs;
        a.parent = d_parent; a.nn_dist_sq = d_nn_dist_sq; a.dist_qp_sq = d_dist_qp_sq; a.nop = d_nop; a.ip_corrected = d_ip_corrected;
        a.ip_qo_denom = d_ip_qo_denom; a.true_ip = d_true_ip; a.neighbor = d_neighbor;
        CUDA_TRY(ix, launch_calibration_samples(ix->dev, a, st));
        return 0;
    });
}

int cphnsw_b200_greedy_descent(cphnsw_b200_index* ix, const float* d_queries, uint64_t nq, uint32_t* d_entry,
                               void* stream) {
    int rc = require_loaded(ix);
    if (rc) return rc;
    if (nq == 0) return 0;
    if (!d_queries || !d_entry) return fail(ix, CPHNSW_B200_EINVAL, "null buffer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return with_lane(ix, st, false, [&](Lane& L) { return run_search(ix, L, d_queries, nq, 1, nullptr, nullptr, d_entry, st); });
}

}  // extern "C"

extern "C" {

struct CandArgs {   // candidate mode of the scan (kernels.h: ExhaustiveArgs)
    const unsigned long long* prior_keys = nullptr;
    const float* tau_in = nullptr;
    unsigned long long* keys = nullptr;
    float* dists = nullptr;
    float* tau_out = nullptr;
    uint64_t id_offset = 0;
};

static int run_exhaustive(cphnsw_b200_index* ix, Lane& L, const float* d_queries, uint64_t nq, uint64_t k, uint64_t kprime,
                          uint64_t id_begin, uint64_t id_end, int64_t* d_ids, float* d_dists, uint32_t* d_sums,
                          float* d_est, cudaStream_t st, const CandArgs* cand = nullptr) {
    const DevIndex& d = ix->dev;
    QStateView qs;
    int rc = ensure_qstate(ix, L, nq, &qs);
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
    if (cand) {
        a.prior_keys = cand->prior_keys; a.tau_in = cand->tau_in; a.cand_keys = cand->keys; a.cand_dists = cand->dists;
        a.tau_out = cand->tau_out; a.id_offset = cand->id_offset;
    }
    const uint64_t m = id_end - id_begin;
    const size_t wb = (exhaustive_workspace_bytes(d, (uint32_t)nq, m, (uint32_t)kprime, ix->num_sms) + 255) & ~(size_t)255;
    // A long range is scanned in pieces (64K vertices, then each piece three times what came before): after a piece the
    // candidates are merged and the thresholds are the exact k'-th estimates so far, which keeps the per-CTA candidate lists
    // of the next piece short -- measured on 10M x 10k: 38.4 -> 32.4 ms at k' = 100, 185 -> 47 ms at k' = 256.  The result is
    // the single scan's (the candidate interface is the one the multi-GPU path uses: sharding.py).
    const bool pieces = !cand && !d_sums && !d_est && k > 0 && kprime > 0 && m > 4 * 65536;
    const size_t kb = ((size_t)nq * kprime * 8 + 255) & ~(size_t)255, tb = ((size_t)nq * 4 + 255) & ~(size_t)255,
                 db = ((size_t)nq * kprime * 4 + 255) & ~(size_t)255;
    rc = ensure_buffer(ix, &L.scratch, &L.scratch_bytes, wb + (pieces ? 2 * kb + tb + db : 0) + 256, false);
    if (rc) return rc;
    a.workspace = L.scratch; a.workspace_bytes = wb;
    if (!pieces) {
        CUDA_TRY(ix, launch_exhaustive(d, a, ix->num_sms, st));
        return 0;
    }
    uint8_t* extra = static_cast<uint8_t*>(L.scratch) + wb;
    unsigned long long* keys[2] = {reinterpret_cast<unsigned long long*>(extra), reinterpret_cast<unsigned long long*>(extra + kb)};
    float* tau = reinterpret_cast<float*>(extra + 2 * kb);
    float* cd = reinterpret_cast<float*>(extra + 2 * kb + tb);
    uint64_t b = 0, size = 65536;
    for (int c = 0;; ++c) {
        uint64_t e = std::min<uint64_t>(m, b + size);
        if (m - e < size / 2) e = m;     // no sliver at the end
        const bool last = e >= m;
        ExhaustiveArgs p = a;
        p.id_begin = id_begin + b; p.id_end = id_begin + e; p.k = 0; p.ids = nullptr; p.dists = nullptr;
        p.prior_keys = c ? keys[(c - 1) & 1] : nullptr;
        p.tau_in = c ? tau : nullptr;
        p.cand_keys = keys[c & 1]; p.tau_out = tau; p.cand_dists = last ? cd : nullptr; p.id_offset = 0;
        CUDA_TRY(ix, launch_exhaustive(d, p, ix->num_sms, st));
        if (last) {
            CUDA_TRY(ix, launch_merge_candidates(keys[c & 1], cd, 1, (uint32_t)nq, (uint32_t)kprime, (uint32_t)k, d_ids, d_dists, nullptr, st));
            return 0;
        }
        size = e * 3;
        b = e;
    }
}

static int check_exhaustive(cphnsw_b200_index* ix, uint64_t nq, uint64_t k, uint64_t kprime, uint64_t id_begin, uint64_t id_end) {
    const DevIndex& d = ix->dev;
    if (d.B != 1) return fail(ix, CPHNSW_B200_EINVAL, "the exhaustive scan is defined for bits=1 indexes (per-vertex 1-bit codes)");
    if (id_end > d.n || id_begin > id_end) return fail(ix, CPHNSW_B200_EINVAL, "id range out of bounds");
    if (nq > 0x7FFFFFFFull || k > 0x7FFFFFFFull) return fail(ix, CPHNSW_B200_EINVAL, "argument too large");
    if (kprime > 1024) return fail(ix, CPHNSW_B200_EINVAL, "kprime (rerank depth) must be <= 1024");
    return 0;
}

int cphnsw_b200_exhaustive_search(cphnsw_b200_index* ix, const float* d_queries, uint64_t nq, uint64_t k, uint64_t kprime,
                                  uint64_t id_begin, uint64_t id_end, int64_t* d_ids, float* d_dists, void* stream) {
    int rc = require_loaded(ix);
    if (rc) return rc;
    if (nq && (!d_queries || (k && (!d_ids || !d_dists)))) return fail(ix, CPHNSW_B200_EINVAL, "null buffer");
    if (kprime < k) kprime = k;
    if ((rc = check_exhaustive(ix, nq, k, kprime, id_begin, id_end))) return rc;
    if (nq == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return with_lane(ix, st, false, [&](Lane& L) {
        return run_exhaustive(ix, L, d_queries, nq, k, kprime, id_begin, id_end, d_ids, d_dists, nullptr, nullptr, st);
    });
}

int cphnsw_b200_exhaustive_candidates(cphnsw_b200_index* ix, const float* d_queries, uint64_t nq, uint64_t kprime, uint64_t id_begin,
                                      uint64_t id_end, uint64_t id_offset, const uint64_t* d_prior_keys, const float* d_tau_in,
                                      uint64_t* d_keys_out, float* d_dists_out, float* d_tau_out, void* stream) {
    int rc = require_loaded(ix);
    if (rc) return rc;
    if (nq && (!d_queries || !d_keys_out)) return fail(ix, CPHNSW_B200_EINVAL, "null buffer");
    if (kprime == 0) return fail(ix, CPHNSW_B200_EINVAL, "kprime must be at least 1");
    if ((rc = check_exhaustive(ix, nq, 0, kprime, id_begin, id_end))) return rc;
    if (id_offset + ix->dev.n > 0xFFFFFFFFull) return fail(ix, CPHNSW_B200_EINVAL, "global ids must stay below 2^32");
    if (nq == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CandArgs c;
    c.prior_keys = reinterpret_cast<const unsigned long long*>(d_prior_keys); c.tau_in = d_tau_in;
    c.keys = reinterpret_cast<unsigned long long*>(d_keys_out); c.dists = d_dists_out; c.tau_out = d_tau_out; c.id_offset = id_offset;
    return with_lane(ix, st, false, [&](Lane& L) {
        return run_exhaustive(ix, L, d_queries, nq, 0, kprime, id_begin, id_end, nullptr, nullptr, nullptr, nullptr, st, &c);
    });
}

int cphnsw_b200_merge_candidates(cphnsw_b200_index* ix, const uint64_t* d_keys, const float* d_dists, uint64_t lists, uint64_t nq,
                                 uint64_t kprime, uint64_t k, int64_t* d_ids_out, float* d_dists_out, float* d_tau_out, void* stream) {
    if (!ix) return CPHNSW_B200_EINVAL;
    if (nq == 0 || lists == 0) return 0;
    if (!d_keys || (k && d_ids_out && (!d_dists || !d_dists_out))) return fail(ix, CPHNSW_B200_EINVAL, "null buffer");
    if (kprime == 0 || kprime > 1024 || k > kprime || nq > 0x7FFFFFFFull || lists * kprime > 16384)
        return fail(ix, CPHNSW_B200_EINVAL, "merge_candidates: need 1 <= kprime <= 1024, k <= kprime, lists * kprime <= 16384");
    DeviceGuard dg(ix->device);
    CUDA_TRY(ix, dg.err);
    CUDA_TRY(ix, launch_merge_candidates(reinterpret_cast<const unsigned long long*>(d_keys), d_dists, (uint32_t)lists, (uint32_t)nq,
                                         (uint32_t)kprime, (uint32_t)k, d_ids_out, d_dists_out, d_tau_out, static_cast<cudaStream_t>(stream)));
    return 0;
}

int cphnsw_b200_exhaustive_estimates(cphnsw_b200_index* ix, const float* d_queries, uint64_t nq, uint64_t id_begin,
                                     uint64_t id_end, uint32_t* d_sums, float* d_est, void* stream) {
    int rc = require_loaded(ix);
    if (rc) return rc;
    if (nq && !d_queries) return fail(ix, CPHNSW_B200_EINVAL, "null buffer");
    if ((rc = check_exhaustive(ix, nq, 0, 0, id_begin, id_end))) return rc;
    if (nq == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return with_lane(ix, st, false, [&](Lane& L) {
        return run_exhaustive(ix, L, d_queries, nq, 0, 0, id_begin, id_end, nullptr, nullptr, d_sums, d_est, st);
    });
}

int cphnsw_b200_unique_topk(cphnsw_b200_index* ix, const int64_t* d_ids_in, const float* d_dists_in, uint64_t nq,
                            uint64_t k_in, uint64_t k_out, const uint32_t* d_id_map, uint64_t map_size,
                            int64_t* d_ids_out, float* d_dists_out, void* stream) {
    if (!ix) return CPHNSW_B200_EINVAL;
    if (nq == 0 || k_out == 0) return 0;
    if (!d_ids_in || !d_dists_in || !d_ids_out || !d_dists_out) return fail(ix, CPHNSW_B200_EINVAL, "null buffer");
    if (k_in > 0x7FFFFFFFull || k_out > 0x7FFFFFFFull) return fail(ix, CPHNSW_B200_EINVAL, "argument too large");
    DeviceGuard dg(ix->device);
    CUDA_TRY(ix, dg.err);
    CUDA_TRY(ix, launch_unique_topk(d_ids_in, d_dists_in, nq, (uint32_t)k_in, (uint32_t)k_out, d_id_map, map_size, d_ids_out,
                                    d_dists_out, static_cast<cudaStream_t>(stream)));
    return 0;
}

static int neighbor_codes_impl(cphnsw_b200_index* ix, uint32_t dim, uint32_t bits, uint64_t rotation_seed,
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
    DeviceGuard dg(ix->device);
    CUDA_TRY(ix, dg.err);
    std::lock_guard<std::mutex> enc(ix->enc_mu);   // the sign table and the tile scratch belong to one call at a time:
    if (ix->enc_done) cudaEventSynchronize(ix->enc_done);   // the previous call's kernel has left them
    else if (cudaEventCreateWithFlags(&ix->enc_done, cudaEventDisableTiming) != cudaSuccess) { ix->enc_done = nullptr; cudaGetLastError(); }
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
    if (ix->enc_done) cudaEventRecord(ix->enc_done, static_cast<cudaStream_t>(stream));
    return 0;
}

int cphnsw_b200_neighbor_codes(cphnsw_b200_index* ix, uint32_t dim, uint32_t bits, uint64_t rotation_seed,
                               const float* d_vectors, uint64_t row_stride, uint64_t n_vectors,
                               const uint32_t* d_parent_ids, const uint32_t* d_nbr_ids, uint64_t n_parents,
                               uint8_t* d_codes, float* d_aux, uint8_t* d_blocks, uint64_t block_stride, void* stream) {
    if (!ix) return CPHNSW_B200_EINVAL;
    return guarded(ix, [&] {
        return neighbor_codes_impl(ix, dim, bits, rotation_seed, d_vectors, row_stride, n_vectors, d_parent_ids, d_nbr_ids,
                                   n_parents, d_codes, d_aux, d_blocks, block_stride, stream);
    });
}

}  // extern "C"
