// Launchers of the sm_100a kernels (one translation unit each).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "device_index.h"

// Block-scope shared variables: `static` when tests/native/ compiles a kernel for the host (blocks run one at a time)
#ifdef CPB_HOST_EMULATION
#define CPB_BLOCK_SHARED static
#else
#define CPB_BLOCK_SHARED __shared__
#endif

namespace cpb {

constexpr uint32_t kCoeffStride = 8;  // A, Bc, C, |q|^2, |q-c|^2, pad

// ---- K1 query preparation (query_prep.cu) ------------------------------------------------
struct PrepOut {
    uint8_t* lut;       // [nq][D/4][16]           (reference form; parity hook only)
    float* coeffs;      // [nq][kCoeffStride]
    float* rotated;     // [nq][D]
    uint32_t* uplanes;  // [nq][4][nch*4]
    float* qT;          // [nq][D] accumulator-major padded raw query
    uint8_t* ubytes;    // [nq][nch*128] the 4-bit query values, one byte per dimension (tensor-core operand of K5)
};
cudaError_t launch_query_prep(const DevIndex& ix, const float* d_queries, uint32_t nq, int center,
                              const PrepOut& out, cudaStream_t stream);

// ---- K2 FastScan over neighbour blocks (fastscan_blocks.cu) --------------------------------
struct FastScanArgs {
    const uint32_t* uplanes;  // [nq][4][nch*4]
    const float* coeffs;      // [nq][kCoeffStride]
    uint32_t nq;
    const uint32_t* query_of_block;  // may be NULL
    const uint32_t* vertex_ids;      // may be NULL
    uint64_t first_vertex, nblocks;
    const float* dqp;                // [nblocks]
    const int32_t* slack_level;      // may be NULL
    uint32_t *nbit, *msb, *msb2;     // [nblocks][32], may be NULL
    float *est, *lower, *msb_lower;  // [nblocks][32], may be NULL
};
cudaError_t launch_fastscan_blocks(const DevIndex& ix, const FastScanArgs& a, int num_sms, cudaStream_t stream);

// ---- K3 (+K4 fused) layer-0 Distance-Adaptive Beam Search (search.cu) --------------------------
struct SearchArgs {
    uint32_t nq;               // work items
    const uint32_t* nq_ptr;    // not NULL: the number of work items is read from device memory (re-run of overflowed queries)
    const uint32_t* query_list;  // NULL = identity; else work item i is query query_list[i]
    uint32_t k;                // max(user k, 1)
    uint32_t kout;             // user k (row stride of the outputs)
    int64_t* ids;              // [*][kout]
    float* dists;              // [*][kout]
    const float* qT;           // prepared queries
    const uint32_t* uplanes;
    const float* coeffs;
    uint8_t* scratch;          // per-slot arenas (frontier heap, large result lists): may hold garbage
    size_t slot_stride;        // bytes per slot
    size_t heap_off, nn_off;   // inside a slot
    uint32_t beam_capacity;    // frontier entries per slot
    uint32_t* bitmaps;         // per-slot "estimated" bitmaps: all-zero between queries
    uint32_t bitmap_words;     // words per slot (a multiple of 32), one bit per vertex
    uint32_t* counters;        // [0] work counter, [1] number of overflowed queries
    uint32_t* overflow_list;   // queries whose frontier overflowed (to be re-run)
    Stats* stats;              // may be NULL
    uint32_t* entry_out;       // descent-only mode: layer-0 entry point per query (else NULL)
    uint32_t warp_smem;        // bytes of shared memory per warp (filled in by launch_search)
};
size_t search_smem_per_warp(const DevIndex& ix, uint32_t k, bool stats);
int search_max_ctas_per_sm(const DevIndex& ix, uint32_t k, int warps_per_cta, bool stats);
cudaError_t launch_search(const DevIndex& ix, const SearchArgs& a, int ctas, int warps_per_cta, bool stats,
                          cudaStream_t stream);

// ---- K4 primitive: exact distances of listed ids (search.cu) ----------------------------------
cudaError_t launch_exact_l2(const DevIndex& ix, const float* qT, const float* coeffs, uint32_t nq,
                            const uint32_t* ids, uint32_t m, float* out, cudaStream_t stream);

// ---- index re-layout (relayout.cu) -------------------------------------------------------------
// d_problems[0] += blocks holding one neighbour id twice, d_problems[1] += blocks with an id >= n
cudaError_t launch_relayout_blocks(const DevIndex& ix, const uint8_t* d_records, uint64_t rec_size, uint32_t nb_off,
                                   uint64_t first, uint32_t count, uint32_t* d_problems, cudaStream_t stream);
cudaError_t launch_relayout_raw(const DevIndex& ix, const float* d_raw, uint64_t first, uint32_t count,
                                cudaStream_t stream);

// ---- K5 exhaustive scan (exhaustive.cu) ---------------------------------------------------------
struct ExhaustiveArgs {
    const uint32_t* uplanes; const float* coeffs; const float* qT; const uint8_t* ubytes;
    uint32_t nq; uint64_t id_begin, id_end;
    uint32_t k, kprime;
    uint32_t* sums; float* est;       // optional dense outputs [nq][id_end-id_begin]
    int64_t* ids; float* dists;       // [nq][k]
    void* workspace; size_t workspace_bytes;
    int use_tensor_cores;             // 2: tcgen05 kind::f16 scan with the screen folded in, 1: tcgen05 kind::i8 scan (each where
                                      // applicable, else the next), 0: popcount scan
    // ---- candidate mode (database sharded over GPUs, or a range scanned piece by piece): instead of the top-k, the k' best
    //      (estimate, id) keys of  prior_keys U [id_begin, id_end)  are written, ascending, to cand_keys ----
    const unsigned long long* prior_keys;  // [nq][kprime] keys of earlier pieces (ids of this index), kNoKey-padded; may be NULL
    const float* tau_in;                   // [nq] upper bounds of the final k'-th estimate known so far (from other pieces /
                                           // other shards): pairs above them are dropped.  NULL, or FLT_MAX entries = none
    unsigned long long* cand_keys;         // [nq][kprime] out; NULL = not candidate mode
    float* cand_dists;                     // [nq][kprime] out, exact distances of cand_keys (the last piece asks for them); may be NULL
    float* tau_out;                        // [nq] out: estimate of the k'-th key of cand_keys (FLT_MAX while there are fewer); may be NULL
    uint64_t id_offset;                    // added to the ids in cand_keys when cand_dists is written (shard-local -> global)
};

// k-way merge of candidate lists (one per shard): keys [lists][nq][kprime] ascending with their exact distances -> the k'
// smallest keys overall -> the k smallest (distance, id) of those.  Any of the outputs may be NULL; tau_out = estimate of the
// k'-th smallest key (FLT_MAX if the lists hold fewer), for which dists may be NULL.
cudaError_t launch_merge_candidates(const unsigned long long* keys, const float* dists, uint32_t lists, uint32_t nq, uint32_t kprime,
                                    uint32_t k, int64_t* ids_out, float* dists_out, float* tau_out, cudaStream_t stream);
size_t exhaustive_workspace_bytes(const DevIndex& ix, uint32_t nq, uint64_t m, uint32_t kprime, int num_sms);
cudaError_t launch_exhaustive(const DevIndex& ix, const ExhaustiveArgs& a, int num_sms, cudaStream_t stream);
// tensor-core form of the scan stage (exhaustive_tc.cu)
bool exhaustive_tc_applicable(const DevIndex& ix, uint32_t kprime);
size_t exhaustive_tc_workspace_bytes(uint32_t nq, uint32_t kprime, int num_sms);
// threshold folded into a kind::f16 contraction (exhaustive_tc16.cu)
bool exhaustive_tc16_applicable(const DevIndex& ix, uint32_t kprime);
cudaError_t launch_exhaustive_scan_tc16(const DevIndex& ix, const ExhaustiveArgs& a, int num_sms, unsigned long long* partial,
                                        uint32_t* nseg, cudaStream_t stream);
cudaError_t launch_exhaustive_scan_tc(const DevIndex& ix, const ExhaustiveArgs& a, int num_sms, unsigned long long* partial,
                                      uint32_t* nseg, cudaStream_t stream);

// ---- N4 calibration sampling (calibration.cu) -------------------------------------------------------------------
struct CalibrationArgs {
    const float* queries;        // [ns][dim] the sampled queries as given
    const uint32_t* start_ids;   // [ns] start vertex of each sample
    uint64_t ns;
    const float* qT;             // K1 outputs for the same queries (no centring)
    const uint32_t* uplanes;
    const float* coeffs;
    uint32_t* parent;            // [ns]
    float* nn_dist_sq;           // [ns]
    float* dist_qp_sq;           // [ns]
    float *nop, *ip_corrected, *ip_qo_denom, *true_ip;   // [ns][32]; 0 past the parent's last neighbour
    uint32_t* neighbor;          // [ns][32]; 0xFFFFFFFF past the parent's last neighbour
};
cudaError_t launch_calibration_samples(const DevIndex& ix, const CalibrationArgs& a, cudaStream_t stream);

// ---- result post-processing, outside the parity path (postprocess.cu) -----------------------------------
cudaError_t launch_unique_topk(const int64_t* ids_in, const float* dists_in, uint64_t nq, uint32_t kin, uint32_t kout,
                               const uint32_t* id_map, uint64_t map_size, int64_t* ids_out, float* dists_out, cudaStream_t stream);

// ---- N3 build side: neighbour codes relative to a parent vertex (neighbor_codes.cu) ---------------------
struct NeighborCodesArgs {
    uint32_t D, dim;
    const float* signs;          // [3][D] rotation sign diagonals
    const float* vectors;        // [n_vectors] rows of row_stride floats (the first dim are the vector)
    uint64_t row_stride, n_vectors;
    const uint32_t* parent_ids;  // [n_parents]; NULL = 0, 1, 2, ...
    const uint32_t* nbr_ids;     // [n_parents][32]; ids >= n_vectors (kInvalid) are empty slots
    uint64_t n_parents;
    uint8_t* codes;              // [n_parents][32][B][D/8]; may be NULL
    float* aux;                  // [n_parents][32][3] nop, ip_qo, ip_cp; may be NULL
    uint8_t* blocks;             // [n_parents] neighbour blocks in the reference's layout, block_stride apart; may be NULL
    uint64_t block_stride;
    // filled in by neighbor_codes_plan
    uint32_t rows, warp_floats, total_warps;
    uint32_t nop_off, ids_off;   // field offsets inside a block
    float norm_factor, inv_sqrt_d, norm_eps, coord_eps;
    // global-tile mode (large D): per-warp tiles [D][32] f32 and [D][32] u8; NULL = tiles in shared memory
    float* tile_x;
    uint8_t* tile_u;
};
struct NeighborCodesPlan { unsigned grid; uint32_t warps; size_t smem_bytes, scratch_bytes; };
cudaError_t neighbor_codes_plan(NeighborCodesArgs& a, uint32_t B, int num_sms, bool global_tile, NeighborCodesPlan* plan);
cudaError_t launch_neighbor_codes(const NeighborCodesArgs& a, uint32_t B, const NeighborCodesPlan& plan, cudaStream_t stream);

}  // namespace cpb
