/* TEST INFRASTRUCTURE ONLY -- the parity oracle.  Never imported, linked or executed by the
 * product path (rabitq-ann-search_b200/); only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may use it.
 *
 * Plain-C restatement of CP-HNSW's query-time hot path (reference:
 * indrajeetadityaroy9/rabitq-ann-search, file:line citations in cphnsw_oracle.c).
 *
 * Parity pinning: the reference ships NO tests or golden vectors, so the
 * oracle is pinned against the reference itself, compiled unmodified into oracle/_ref/ by
 * oracle/Makefile (tests/test_oracle_vs_ref.py, run wherever /root/reference-built _ref exists)
 * and against committed fixtures under tests/golden/ generated from that same build
 * (tests/golden/make_golden.py).
 */
#ifndef CPHNSW_ORACLE_H
#define CPHNSW_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- query preparation -------------------------------------------------------------- */
void cpo_rotation_signs(uint32_t D, uint64_t seed, float* signs /* [3][D] */);
void cpo_fht(float* x, uint32_t D);
/* rotated (optional, may be NULL) receives the rotated, norm_factor-scaled query (D floats) */
void cpo_set_encode_variant(int v);   /* contraction variant of coeff_constant, see cpo_encode_query (test use only; default 0) */
void cpo_encode_query(uint32_t dim, uint32_t D, const float* signs, const float* q,
                      uint8_t* lut /* [D/4][16] */, float coeffs[3], float* rotated);

/* BUILD side (oracle groundwork): RaBitQEncoder<D>::compute_neighbor_aux (encoder/rabitq_encoder.hpp:138-181) */
void cpo_neighbor_aux_1bit(uint32_t dim, uint32_t D, const float* signs, const float* parent, const float* nb,
                           int fused, uint8_t* code, float aux[3]);
/* NbitRaBitQEncoder<D,B>::compute_neighbor_aux_nbit + caq_quantize (:287-323, :371-467); flags: see cphnsw_oracle.c */
#define CPO_CAQ_FLAGS 0x1F9u   /* the contraction pattern of the compiled reference; see cphnsw_oracle.c */
void cpo_neighbor_aux_nbit(uint32_t dim, uint32_t D, uint32_t B, const float* signs, const float* parent, const float* nb,
                           uint32_t flags, uint8_t* planes, float aux[3]);

/* ---- FastScan ------------------------------------------------------------------------ */
void cpo_fastscan_plane(uint32_t D, const uint8_t* lut, const uint8_t* packed /* [D/8][32] */,
                        uint32_t out[32]);
void cpo_fastscan(uint32_t D, uint32_t B, const uint8_t* lut, const uint8_t* planes,
                  uint32_t nbit[32], uint32_t msb[32], uint32_t msb2[32]);

/* params = {coeff_fastscan, coeff_popcount, coeff_constant, affine_a, affine_b, ip_qo_floor, dot_slack} */
void cpo_convert_1bit(const float params[7], const uint32_t* sums, const float* nop,
                      const float* ip_qo, const float* ip_cp, const uint16_t* pop,
                      uint32_t count, float dqp, float* est, float* lower);
void cpo_convert_msb(uint32_t B, const float params[7], const uint32_t* msb2, const float* nop,
                     const float* ip_qo, const float* ip_cp, const uint16_t* pop,
                     uint32_t count, float dqp, float* lower);
void cpo_convert_nbit(uint32_t B, const float params[7], const uint32_t* nbit, const uint32_t* msb,
                      const float* nop, const float* ip_qo, const float* ip_cp,
                      const uint16_t* pop, const uint16_t* wpop,
                      uint32_t count, float dqp, float* est, float* lower);

/* ---- exact distances ----------------------------------------------------------------- */
float cpo_dot(uint32_t D, const float* a, const float* b);
float cpo_l2(uint32_t D, const float* a, const float* b);

/* ---- index view (filled from a reference save file by oracle/cphnsw_oracle.py) -------- */
typedef struct {
    uint32_t D, B, dim;
    uint64_t n;
    /* per-vertex AoS records exactly as the reference stores them */
    const uint8_t* search_data;
    uint64_t rec_size;     /* sizeof(VertexSearchData<D,32,B>) */
    uint32_t nb_off;       /* offset of the neighbour block inside a record */
    const float* raw;      /* [n][D] */
    const float* norm_sq;  /* [n] */
    /* calibration (api/hnsw_index.hpp:33-58) */
    float affine_a, affine_b, ip_qo_floor;
    float slack_levels[32];
    int32_t num_slack_levels;
    float search_gamma, gamma_max, gamma_beta;
    uint64_t gamma_warmup;
    /* upper layers, CSR per level (level L is at index L-1) */
    int32_t max_level;
    uint32_t entry_point;        /* header ep: upper entry point, and graph entry after load */
    uint32_t graph_entry_point;  /* graph_.entry_point() (== entry_point after load) */
    uint32_t n_layers;
    const uint32_t* const* layer_nodes; /* sorted node ids per layer */
    const uint32_t* const* layer_offs;  /* [n_edges+1] */
    const uint32_t* const* layer_nbrs;
    const uint32_t* layer_sizes;        /* n_edges per layer */
    const float* signs;                 /* [3][D] rotation signs */
} cpo_index;

typedef struct {
    uint64_t pops, expansions, exact_calls, beam_pushes, max_beam, nn_pushes;
    uint64_t lb_skips, gamma_terms, msb_skipped, estimated, descent_dists;
} cpo_stats;

/* One query (dim floats).  Writes up to k results (ascending distance), returns how many.
 * ep_override != 0xFFFFFFFF skips the upper-layer descent and starts layer 0 there. */
int cpo_search(const cpo_index* ix, const float* query, uint64_t k,
               uint32_t* ids, float* dists, cpo_stats* stats);
/* search_batch convention of src/bindings.cpp:177-218: int64 ids, rows padded with -1 / FLT_MAX */
int cpo_search_batch(const cpo_index* ix, const float* queries, uint64_t nq, uint64_t k,
                     int64_t* ids, float* dists, cpo_stats* stats_sum, int num_threads);
uint32_t cpo_greedy_descent(const cpo_index* ix, const float* query_padded, uint64_t* ndist);

/* Exhaustive-scan oracle composed from reference primitives (no reference mode). */
typedef struct {
    const uint8_t* codes;   /* [n][code_stride]: per-vertex 1-bit signs at offset 0 (u64 words) */
    uint64_t code_stride;   /* = rec_size (codes live at offset 0 of each SearchData record) */
    uint32_t nop_off, ipqo_off;
    const float* centroid;  /* [dim] */
} cpo_flat_view;
int cpo_exhaustive_search(const cpo_index* ix, const cpo_flat_view* fv, const float* query,
                          uint64_t k, uint64_t kprime, uint64_t id_begin, uint64_t id_end,
                          uint32_t* ids, float* dists, uint32_t* est_sums /* optional [id_end-id_begin] */,
                          float* est_out /* optional */);

/* Calibration sampling (N4): process_query of Index::calibrate_estimator (api/hnsw_index.hpp:786-866); outputs [32] per query.
 * flags: see the definition (contraction candidates of the two scalar expressions). */
void cpo_calibration_sample(const cpo_index* ix, const float* query_padded, uint32_t start_id, uint32_t flags, uint32_t* parent_out,
                            float* nn_dist_sq, float* dist_qp_sq, float* nop_out, float* ip_corrected, float* ip_qo_denom, float* true_ip,
                            uint32_t* neighbor);

#ifdef __cplusplus
}
#endif
#endif
