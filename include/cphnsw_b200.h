/* cphnsw_b200 -- C ABI of the B200-native CP-HNSW query path.
 *
 * This is the drop-in boundary for the query-time hot path of the reference
 * (indrajeetadityaroy9/rabitq-ann-search).  The reference exposes that path only through the
 * pybind11 class CPIndex (src/bindings.cpp:115-240, over PyIndexBase :21-37); a maintainer
 * binds the functions below beside it (INTEGRATION.md shows the stub).  Plain pointers and
 * sizes only: no C++ types, no torch types, no exceptions cross this boundary.  Every function
 * returns 0 on success or a negative CPHNSW_B200_E* code; cphnsw_b200_last_error() gives the
 * message.  There is no CPU fallback: without a CUDA device every call fails.
 *
 * Pointers whose name starts with d_ are device addresses (e.g. torch.Tensor.data_ptr());
 * all others are host addresses.  `stream` is a cudaStream_t passed as void* (NULL = default).
 */
#ifndef CPHNSW_B200_H
#define CPHNSW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPHNSW_B200_OK 0
#define CPHNSW_B200_EINVAL (-1)    /* bad argument (std::invalid_argument -> ValueError)   */
#define CPHNSW_B200_ERUNTIME (-2)  /* bad file / state   (std::runtime_error  -> RuntimeError) */
#define CPHNSW_B200_ECUDA (-3)     /* CUDA runtime error                                    */
#define CPHNSW_B200_ENOMEM (-4)

typedef struct cphnsw_b200_index cphnsw_b200_index;

/* Host view of a finalized reference index, field for field what Index<D,32,BitWidth> holds
 * after finalize()/load() (api/hnsw_index.hpp:33-58,445-466; graph/rabitq_graph.hpp:286-290),
 * i.e. also what its save file v2 stores (api/hnsw_index.hpp:217-303).  All ids are the
 * reference's internal (BFS-reordered) ids. */
typedef struct {
    uint32_t D;          /* padded dimension (power of two, 16..2048) */
    uint32_t bits;       /* BitWidth: 1, 2 or 4 */
    uint32_t dim;        /* user dimension */
    uint64_t n;
    const uint8_t* search_data;  /* [n][rec_size] VertexSearchData<D,32,bits> records */
    uint64_t rec_size;           /* sizeof(VertexSearchData) */
    uint32_t nb_off;             /* offsetof(VertexSearchData, neighbors) = sizeof(code) */
    const float* raw;            /* [n][D] zero-padded raw vectors */
    const float* norm_sq;        /* [n] */
    const float* centroid;       /* [dim] (only the exhaustive-scan mode reads it) */
    const uint8_t* calibration;  /* CalibrationSnapshot, 248 bytes (api/hnsw_index.hpp:33-58) */
    int32_t max_level;
    uint32_t entry_point;        /* upper-layer entry point (== graph entry after load) */
    uint32_t graph_entry_point;  /* layer-0 hub, used when max_level == 0 */
    uint64_t rotation_seed;
    uint32_t n_layers;                   /* upper layers, level L at index L-1 */
    const uint32_t* const* layer_nodes;  /* [n_layers][layer_sizes[l]] sorted node ids */
    const uint32_t* const* layer_offs;   /* [n_layers][layer_sizes[l]+1] CSR offsets */
    const uint32_t* const* layer_nbrs;   /* [n_layers][...] neighbour node ids */
    const uint32_t* layer_sizes;
} cphnsw_b200_host_index;

typedef struct {
    uint32_t D, bits, dim;
    uint64_t n;
    int32_t max_level;
    uint32_t entry_point;
    uint32_t n_layers;
    uint32_t block_stride;      /* bytes per re-laid-out neighbour block in HBM */
    uint64_t device_bytes;      /* HBM held by the index */
    float affine_a, affine_b, ip_qo_floor;
    float search_gamma, gamma_max, gamma_beta;
    uint64_t gamma_warmup;
    int32_t num_slack_levels;
    float slack_levels[32];
} cphnsw_b200_info;

/* Per-batch counters of the last search (sums over queries; max_beam is a max). */
typedef struct {
    uint64_t pops, expansions, exact_calls, beam_pushes, max_beam, nn_pushes;
    uint64_t lb_skips, gamma_terms, msb_skipped, estimated, descent_dists;
    uint64_t overflow_retries;  /* queries re-run with a larger frontier arena */
    uint64_t kernel_launches;   /* kernels the call enqueued: K1, K3, and K3 again over the (usually empty) overflow list */
} cphnsw_b200_stats;

/* ---- lifetime: replaces PyIndexWrapper's unique_ptr<Index<...>> (src/bindings.cpp:39-75) ---- */
int cphnsw_b200_create(int device, cphnsw_b200_index** out);
void cphnsw_b200_destroy(cphnsw_b200_index* ix);
const char* cphnsw_b200_last_error(const cphnsw_b200_index* ix); /* ix may be NULL: create() errors */

/* ---- index hand-off ---------------------------------------------------------------------- */
/* Parse a save file v2 written by CPIndex.save (api/hnsw_index.hpp:217-303), validate it like
 * Index::load (:305-443: magic, version, R=32, bits, D = next_pow2(dim), truncation), re-lay it
 * out and upload it. */
int cphnsw_b200_load(cphnsw_b200_index* ix, const char* path);
/* Same from an in-memory view (what a binding beside src/bindings.cpp would pass). */
int cphnsw_b200_upload(cphnsw_b200_index* ix, const cphnsw_b200_host_index* host);
int cphnsw_b200_get_info(const cphnsw_b200_index* ix, cphnsw_b200_info* out);

/* ---- the hot path: replaces PyIndexBase::search_raw looped by search_batch
 *      (src/bindings.cpp:60-63,177-218 -> Index::search api/hnsw_index.hpp:168-211) ----------- */
/* Host buffers: queries [nq][dim] f32, ids [nq][k] i64, dists [nq][k] f32; rows padded with
 * -1 / FLT_MAX exactly like bindings.cpp:201-210.  k == 0 behaves as the reference does
 * (searches with k = 1, writes nothing). */
int cphnsw_b200_search_batch(cphnsw_b200_index* ix, const float* queries, uint64_t nq, uint64_t k,
                             int64_t* ids, float* dists);
/* Device buffers; asynchronous: everything (K1, K3, and the re-run of any query whose frontier outgrew its
 * arena) is enqueued on `stream` and the call returns without waiting for the device.  Results are ready in stream
 * order.  Calls issued on different streams (or from different host threads) may overlap on the device: each works
 * in its own lane of the handle (two lanes; a third call first waits for the oldest one), which is how the
 * reference's concurrent-reader guarantee (Index::search takes a shared_lock and uses thread_local scratch,
 * api/hnsw_index.hpp:172) is kept. */
int cphnsw_b200_search_batch_device(cphnsw_b200_index* ix, const float* d_queries, uint64_t nq,
                                    uint64_t k, int64_t* d_ids, float* d_dists, void* stream);
/* search_batch in two halves, for callers that stream batches (src/bindings.cpp:177-218 has no counterpart: its
 * OpenMP loop needs none).  submit copies the host queries in, enqueues the search and the copies of the results
 * out on an internal stream and returns a ticket; wait blocks until that batch's ids/dists are in the host buffers.
 * With two batches in flight the drain of one batch's persistent grid overlaps the start of the next.  For the
 * copies to be asynchronous the host buffers should be page-locked.  The buffers must stay valid until wait. */
int cphnsw_b200_search_batch_submit(cphnsw_b200_index* ix, const float* queries, uint64_t nq, uint64_t k,
                                    int64_t* ids, float* dists, uint64_t* ticket);
int cphnsw_b200_search_batch_wait(cphnsw_b200_index* ix, uint64_t ticket);
/* Wait for every call in flight on this handle (all lanes); returns the first deferred error, if any. */
int cphnsw_b200_synchronize(cphnsw_b200_index* ix);
/* Counters of the last search_batch* call (synchronises). */
int cphnsw_b200_last_stats(cphnsw_b200_index* ix, cphnsw_b200_stats* out);
/* Device time of the kernels of the last search_batch* call, from CUDA events on the launching
 * stream: K1 (query preparation) and K3 (descent + beam search + rerank; summed over the re-run
 * if a frontier overflowed). */
int cphnsw_b200_last_timings(cphnsw_b200_index* ix, float* prep_ms, float* search_ms);
/* Tuning knobs that do not change results: warps in flight and frontier arena size. */
int cphnsw_b200_set_option(cphnsw_b200_index* ix, const char* name, int64_t value);

/* ---- kernel-level hooks (parity tests and micro-benchmarks) --------------------------------- */
/* K1: replaces RaBitQEncoder::encode_query_raw (encoder/rabitq_encoder.hpp:73-79,197-209) after
 * the zero-padding of Index::search (api/hnsw_index.hpp:174-180).  Any output may be NULL.
 * d_lut [nq][D/4][16] u8, d_coeffs [nq][3] f32, d_rotated [nq][D] f32 (scaled rotated query),
 * d_uplanes [nq][4][max(D,128)/32] u32 (bit t of the 4-bit query values, the kernels' own form).
 * center != 0 subtracts the index centroid first (exhaustive-scan mode only). */
int cphnsw_b200_prepare_queries(cphnsw_b200_index* ix, const float* d_queries, uint64_t nq, int center,
                                uint8_t* d_lut, float* d_coeffs, float* d_rotated,
                                uint32_t* d_uplanes, void* stream);
/* K2: replaces fastscan::compute_inner_products / compute_nbit_inner_products /
 * compute_msb_only_inner_products and the three convert_* epilogues
 * (distance/fastscan_kernel.hpp:17-87,197-217,349-368 and :89-194,220-346,371-425) over the
 * neighbour blocks of `nblocks` vertices.  Block i uses query d_query_of_block[i] (or query 0
 * when NULL) of the nq queries prepared by K1 (d_uplanes, d_coeffs), dist_qp_sq d_dqp[i] and the
 * slack level d_slack_level[i] (or level 0 when NULL).  d_vertex_ids NULL = vertices
 * first_vertex .. first_vertex+nblocks-1.  Outputs are [nblocks][32]; any may be NULL. */
int cphnsw_b200_fastscan_blocks(cphnsw_b200_index* ix, const uint32_t* d_uplanes, const float* d_coeffs,
                                uint64_t nq, const uint32_t* d_query_of_block,
                                const uint32_t* d_vertex_ids, uint64_t first_vertex, uint64_t nblocks,
                                const float* d_dqp, const int32_t* d_slack_level,
                                uint32_t* d_nbit, uint32_t* d_msb, uint32_t* d_msb2,
                                float* d_est, float* d_lower, float* d_msb_lower, void* stream);
/* K4 primitive: replaces the exact_l2 lambda (search/rabitq_search.hpp:88-93, dot_product_simd
 * core/memory.hpp:81-96): out[q][j] = max(|q|^2 + norm_sq[id] - 2<q,x_id>, 0), ids [nq][m]. */
int cphnsw_b200_exact_l2(cphnsw_b200_index* ix, const float* d_queries, uint64_t nq,
                         const uint32_t* d_ids, uint64_t m, float* d_out, void* stream);
/* K3 prologue: replaces greedy_search_layer over the upper layers (api/hnsw_index.hpp:195-202,
 * 617-638): layer-0 entry point per query. */
int cphnsw_b200_greedy_descent(cphnsw_b200_index* ix, const float* d_queries, uint64_t nq,
                               uint32_t* d_entry, void* stream);

/* ---- exhaustive batched scan (bits == 1 indexes; composed from reference primitives, the
 *      reference has no such mode: see DESIGN.md) over internal ids [id_begin, id_end) -------- */
int cphnsw_b200_exhaustive_search(cphnsw_b200_index* ix, const float* d_queries, uint64_t nq,
                                  uint64_t k, uint64_t kprime, uint64_t id_begin, uint64_t id_end,
                                  int64_t* d_ids, float* d_dists, void* stream);
/* The scan in pieces -- a database sharded over GPUs, each shard scanned range by range -- with the single-scan result:
 * writes to d_keys_out [nq][kprime] the kprime smallest keys (estimate bits << 32 | id, 0xFF..FF padding; ascending on
 * the last piece, in no particular order before) of
 * d_prior_keys (the keys_out of earlier ranges of this index; may be NULL) and the vertices of [id_begin, id_end).
 * d_tau_in [nq] (may be NULL; FLT_MAX = none): upper bounds of the final kprime-th estimate learnt elsewhere (other
 * shards, through an all-reduce(min) of d_tau_out) -- pairs above them are dropped early, which is what keeps a shard's
 * work proportional to its size.  d_tau_out [nq] (may be NULL): the estimate of the kprime-th key written (FLT_MAX while
 * there are fewer).  The last piece passes d_dists_out [nq][kprime]: the exact distances (search/rabitq_search.hpp:88-93)
 * of the keys, whose ids are then also shifted by id_offset (shard-local -> global; must stay below 2^32). */
int cphnsw_b200_exhaustive_candidates(cphnsw_b200_index* ix, const float* d_queries, uint64_t nq, uint64_t kprime,
                                      uint64_t id_begin, uint64_t id_end, uint64_t id_offset,
                                      const uint64_t* d_prior_keys, const float* d_tau_in, uint64_t* d_keys_out,
                                      float* d_dists_out, float* d_tau_out, void* stream);
/* Merge of `lists` candidate lists (d_keys [lists][nq][kprime] with d_dists alongside: what every shard wrote, gathered):
 * the kprime smallest keys overall, then the k smallest (distance, id) of those -- exactly what
 * cphnsw_b200_exhaustive_search returns on the whole database.  d_tau_out [nq] (may be NULL): the estimate of the
 * kprime-th smallest key; with k == 0 or d_ids_out == NULL only that is computed (then d_dists may be NULL).
 * lists * kprime <= 16384. */
int cphnsw_b200_merge_candidates(cphnsw_b200_index* ix, const uint64_t* d_keys, const float* d_dists, uint64_t lists,
                                 uint64_t nq, uint64_t kprime, uint64_t k, int64_t* d_ids_out, float* d_dists_out,
                                 float* d_tau_out, void* stream);
/* Estimator only: integer sums [nq][id_end-id_begin] u32 and estimates f32 (either may be NULL). */
int cphnsw_b200_exhaustive_estimates(cphnsw_b200_index* ix, const float* d_queries, uint64_t nq,
                                     uint64_t id_begin, uint64_t id_end,
                                     uint32_t* d_sums, float* d_est, void* stream);

/* ---- opt-in post-processing, outside the parity path ---------------------------------------- */
/* The reference returns a vertex once per time it was scored (search/rabitq_search.hpp:133,236,250;
 * BoundedMaxHeap::push does not de-duplicate, :26-35) and internal BFS-reordered ids
 * (graph/rabitq_graph.hpp:208-278); search_batch reproduces both.  This call cleans a result up:
 * from each ascending row of k_in pairs (the output of search_batch*, device buffers) it keeps
 * the first occurrence of every id, writes the first k_out of them padded with -1 / FLT_MAX
 * (src/bindings.cpp:201-210) and, when d_id_map [map_size] is not NULL, replaces every internal
 * id by d_id_map[id] (the caller's original id).  In and out buffers must not overlap. */
int cphnsw_b200_unique_topk(cphnsw_b200_index* ix, const int64_t* d_ids_in, const float* d_dists_in,
                            uint64_t nq, uint64_t k_in, uint64_t k_out, const uint32_t* d_id_map,
                            uint64_t map_size, int64_t* d_ids_out, float* d_dists_out, void* stream);

/* ---- build side: neighbour codes relative to a parent vertex ------------------ */
/* What prune_and_write (graph/graph_refinement.hpp:46-67) stores for each selected neighbour of a
 * vertex: RaBitQEncoder<D>::compute_neighbor_aux (encoder/rabitq_encoder.hpp:138-181) for bits == 1,
 * NbitRaBitQEncoder<D,B>::compute_neighbor_aux_nbit (:287-323, quantiser caq_quantize :371-467) for
 * bits 2, 4 -- bit for bit.  The handle supplies the device and the error string; it need not hold
 * an index (the build side runs before one exists).  rotation_seed: the encoder's (42 in
 * api/hnsw_index.hpp).  d_vectors: n_vectors rows of row_stride floats, the first dim of each the
 * vector.  d_parent_ids [n_parents] (NULL = 0, 1, 2, ...) and d_nbr_ids [n_parents][32]: row ids;
 * an id >= n_vectors (0xFFFFFFFF) is an empty slot and yields zeros.  Outputs, D = next_pow2(dim)
 * (at least 16): d_codes u8 [n_parents][32][bits][D/8], planes MSB first, bit i%8 of byte i/8 =
 * dimension i (core/codes.hpp:107-116); d_aux f32 [n_parents][32][3] = nop, ip_qo, ip_cp; d_blocks:
 * one FastScanNeighborBlock / NbitFastScanNeighborBlock per parent in the reference's own layout
 * (distance/fastscan_layout.hpp:51-92, 114-155: packed planes, nop, ip_qo, ip_cp, popcounts,
 * [weighted_popcounts,] neighbor_ids, count = one past the last occupied slot), block_stride bytes
 * apart (>= the struct's size, a multiple of 4; bytes past `count` + 4 are left alone) -- i.e.
 * what nb.set_neighbor wrote, ready to be copied into search_data_[id].neighbors.  Any of the three
 * outputs may be NULL. */
int cphnsw_b200_neighbor_codes(cphnsw_b200_index* ix, uint32_t dim, uint32_t bits, uint64_t rotation_seed,
                               const float* d_vectors, uint64_t row_stride, uint64_t n_vectors,
                               const uint32_t* d_parent_ids, const uint32_t* d_nbr_ids, uint64_t n_parents,
                               uint8_t* d_codes, float* d_aux, uint8_t* d_blocks, uint64_t block_stride,
                               void* stream);

/* ---- calibration side: the sample loop of Index::calibrate_estimator -------------------------- */
/* What the lambda process_query (api/hnsw_index.hpp:786-866) records for each sampled query: d_queries [ns][dim] (a
 * database vector or the reference's synthetic perturbation of one -- the caller draws them, the reference's RNG being the
 * host library's) and d_start_ids [ns] (sample_ids[parent_cursor % n]).  Per sample: d_parent (the start vertex or the
 * closest of its neighbours, l2_distance_simd, strict <, stored order), d_nn_dist_sq = d_dist_qp_sq = that distance; per
 * neighbour slot of the parent's block, [ns][32]: d_nop = max(nop, 1e-12), d_ip_corrected = ip_approx - ip_cp,
 * d_ip_qo_denom = max(|ip_qo|, 1e-10), d_true_ip = <q - p, o - p> / nop, d_neighbor = the neighbour's id; slots from the
 * first empty one on hold 0 / 0xFFFFFFFF.  The index must be loaded (the graph is complete when calibration runs). */
int cphnsw_b200_calibration_samples(cphnsw_b200_index* ix, const float* d_queries, const uint32_t* d_start_ids, uint64_t ns,
                                    uint32_t* d_parent, float* d_nn_dist_sq, float* d_dist_qp_sq, float* d_nop,
                                    float* d_ip_corrected, float* d_ip_qo_denom, float* d_true_ip, uint32_t* d_neighbor,
                                    void* stream);

#ifdef __cplusplus
}
#endif
#endif
