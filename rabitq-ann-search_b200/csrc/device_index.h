// Device-side view of a finalized CP-HNSW index as this library lays it out in HBM, plus the
// host-side owner that builds it.  See DESIGN.md "Data layout in HBM".
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <condition_variable>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/cphnsw_b200.h"

namespace cpb {

constexpr uint32_t kInvalid = 0xFFFFFFFFu;
constexpr int kMaxLevels = 24;
constexpr uint32_t kR = 32;  // fixed layer-0 degree (src/bindings.cpp:42)

// CalibrationSnapshot fields the query path consumes (api/hnsw_index.hpp:33-58)
struct Calib {
    float affine_a, affine_b, ip_qo_floor;
    float slack[32];
    int32_t num_slack;
    float gamma, gamma_max, gamma_beta;
    uint64_t gamma_warmup;
};

// One upper HNSW layer as slot-addressed CSR.  A "slot" is the position of a node in the layer's
// sorted node list; neighbour lists carry both the node id (for the distance) and the
// neighbour's own slot, so the greedy descent never searches (find_edge, hnsw_index.hpp:468-474,
// becomes a table walk).  down[s] = slot of node[s] one level below (level 1: the node id).
struct Level {
    const uint32_t* node;
    const uint32_t* offs;
    const uint32_t* nbr_node;
    const uint32_t* nbr_slot;  // kInvalid when the neighbour has no edge record at this level
    const uint32_t* down;      // kInvalid when absent below
    uint32_t size;
};

// Neighbour block of one vertex in HBM (block_stride bytes, 128-B aligned):
//   [plane b][chunk c][slot v] uint4   code bits of dims 128c..128c+127 of neighbour v, plane b
//                                      (bit j of word w <-> dim 128c+32w+j); planes MSB first
//   aux_off + 0    ids   u32[32]
//   aux_off + 128  nop   f32[32]
//   aux_off + 256  ip_qo f32[32]
//   aux_off + 384  ip_cp f32[32]
//   aux_off + 512  pops  u32[32]   lo16 = popcounts (plane-0 popcount), hi16 = weighted_popcounts
//   aux_off + 640  count u32
//   aux_off + 644  norm_sq f32     |x|^2 of the vertex itself (one bulk copy brings all an expansion reads)
struct DevIndex {
    uint32_t D, B, dim;
    uint32_t nch;  // 128-dim chunks per code plane = max(D,128)/128
    uint32_t T;    // D/8: length of one exact-L2 accumulator chain
    uint64_t n;
    const uint8_t* blocks;
    uint32_t block_stride, aux_off;
    uint32_t dup_neighbors;  // != 0: some block lists a neighbour id twice (never seen from the reference's builder)
    // raw vectors, accumulator-major with a bank swizzle: rawT[id][l*T + raw_chunk_pos(l, t, T)] = raw[id][8t + l]  (l = 0..7; device_math.cuh)
    const float* rawT;
    const float* norm_sq;
    const float* signs;     // [3][D]
    const float* centroid;  // [dim]
    int32_t max_level;
    uint32_t entry_point;        // node id
    uint32_t entry_slot;         // its slot in level max_level (kInvalid if it has no record)
    uint32_t graph_entry_point;  // used when max_level == 0
    uint32_t n_levels;
    Level levels[kMaxLevels];  // level L at [L-1]
    // per-vertex 1-bit codes for the exhaustive scan (B == 1 only)
    const uint32_t* flat_codes;  // [n][nch*4]
    const float* flat_nop;
    const float* flat_ipqo;
    const uint16_t* flat_pop;
    Calib calib;
};

// prepared query state in HBM (written by K1, read by K2/K3/K5)
struct QueryState {
    float* qT;          // [nq][D] padded raw query, accumulator-major
    uint32_t* uplanes;  // [nq][4][nch*4]
    float* coeffs;      // [nq][4] A, Bc, C, |q|^2
};

struct Stats {
    unsigned long long pops, expansions, exact_calls, beam_pushes, max_beam, nn_pushes;
    unsigned long long lb_skips, gamma_terms, msb_skipped, estimated, descent_dists;
};

}  // namespace cpb

namespace cpb {

// Everything one in-flight call needs beyond the (read-only) index: scratch arenas, the prepared-query state, staging
// buffers of the host-buffer entry points, counters and events.  A handle owns kLanes of them, handed out round-robin,
// so two calls can be in flight at once -- from two host threads (the reference's search is safe for concurrent
// callers: shared_lock + thread_local scratch, api/hnsw_index.hpp:172) or pipelined from one (submit / wait): the
// tail of one batch's persistent grid then overlaps the head of the next.
constexpr int kLanes = 2;
struct Lane {
    bool created = false;
    void* scratch = nullptr;      // frontier arenas, overflow lists, K5 workspace
    size_t scratch_bytes = 0;
    void* bitmaps = nullptr;      // zero between searches; slot i at i * bitmap_words
    size_t bitmaps_bytes = 0;
    void* qstate = nullptr;       // K1 outputs
    size_t qstate_bytes = 0;
    void* stage = nullptr;        // device copies of host queries / results
    size_t stage_bytes = 0;
    Stats* d_stats = nullptr;
    uint32_t* d_counters = nullptr;   // [0] work counter, [1] overflowed queries, [2] re-run work counter, [3] re-run overflows
    uint32_t* h_counters = nullptr;   // pinned copy, valid once `done` has completed
    cudaStream_t stream = nullptr;    // the host-buffer entry points run here
    cudaEvent_t ev[4] = {};           // K1 begin / end, K3 begin / end (re-run included)
    cudaEvent_t done = nullptr;       // recorded after the last operation of a call
    int state = 0;                    // 0 free, 1 held by a host thread, 2 work in flight (nobody holds it)
    uint64_t ticket = 0;              // of the call that used the lane last
    bool counters_pending = false;    // h_counters belongs to an unfinished search
    bool timed = false;               // ev[] belong to the last call
    int status = 0;                   // outcome of the last finished call
    std::string status_msg;
    uint64_t overflow_retries = 0;
    uint64_t launches = 0;            // kernels the last search enqueued
};

}  // namespace cpb

struct cphnsw_b200_index {
    int device = 0;
    bool loaded = false;
    std::string err;
    cpb::DevIndex dev{};
    uint64_t device_bytes = 0;
    std::vector<void*> allocs;  // everything dev points to
    int num_sms = 148;
    // options (do not change results)
    int64_t warps_per_cta = 2;       // small CTAs: a CTA leaves the SM as soon as its warps run out of queries, so the next batch moves in
    int64_t ctas_per_sm = 16;
    int64_t beam_capacity = 0;        // frontier entries per in-flight query (first attempt); 0 = sized from free HBM
    int64_t collect_stats = 0;        // per-batch counters (costs registers: off on the fast path)
    int64_t exhaustive_tensor_cores = 2;  // K5 scan: 2 = tcgen05 kind::f16 with the screen folded in, 1 = tcgen05 kind::i8 (each where
                                          // applicable, else the next), 0 = popcount form
    size_t frontier_budget = 0;  // bytes of HBM the frontier arenas of ONE lane may take (fixed when first needed)
    // lanes: per-call state.  mu guards the lane states, the ticket counter and err.
    std::mutex mu;
    std::condition_variable cv;
    cpb::Lane lanes[cpb::kLanes];
    uint64_t next_lane = 0, next_ticket = 0;
    int last_lane = -1;          // lane of the most recent search (last_stats / last_timings)
    int occ_key[4] = {0, 0, 0, 0};   // cached occupancy query of the search kernel: smem per warp, warps, stats -> CTAs per SM
    uint32_t* d_problems = nullptr;  // re-layout diagnostics (upload): [2] duplicate neighbour ids, [3] ids out of range
    // build side (neighbor_codes): one call at a time (enc_mu)
    std::mutex enc_mu;
    float* enc_signs = nullptr;      // rotation sign diagonals of the build-side encoder, [3][enc_signs_D]
    uint32_t enc_signs_D = 0;
    uint64_t enc_signs_seed = 0;
    void* enc_scratch = nullptr;     // global-tile mode of neighbor_codes
    size_t enc_scratch_bytes = 0;
    cudaEvent_t enc_done = nullptr;  // after the last neighbor_codes kernel
    int64_t neighbor_codes_tile = 0; // option: where the per-warp tiles live: 1 = shared memory, else global memory / L2 (results identical)
};
