/* TEST INFRASTRUCTURE ONLY -- see cphnsw_oracle.h.
 *
 * Plain-C restatement of the CP-HNSW query hot path.  Each function cites the reference
 * file:line it follows (paths relative to the reference repository root).  Where the
 * reference's result depends on how GCC 13.3 -O3 -mfma contracted a scalar expression
 * (SURVEY.md F10), the fused operations found in the compiled reference are written here as
 * explicit fmaf() calls and this file is compiled with -ffp-contract=off.
 */
#include "cphnsw_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define INVALID_NODE 0xFFFFFFFFu

/* ======================================================================================
 * Rotation signs: encoder/rotation.hpp:19-32.  std::mt19937_64(seed) feeding
 * std::uniform_int_distribution<int>(0,1); with libstdc++ each coin is the top bit of one
 * 64-bit draw (Lemire multiply-shift with range 2 never rejects).  SURVEY App. A1.
 * ==================================================================================== */
typedef struct { uint64_t mt[312]; int idx; } mt64;

static void mt64_seed(mt64* s, uint64_t seed) {
    s->mt[0] = seed;
    for (int i = 1; i < 312; ++i)
        s->mt[i] = 6364136223846793005ULL * (s->mt[i - 1] ^ (s->mt[i - 1] >> 62)) + (uint64_t)i;
    s->idx = 312;
}

static uint64_t mt64_next(mt64* s) {
    if (s->idx >= 312) {
        for (int i = 0; i < 312; ++i) {
            uint64_t x = (s->mt[i] & 0xFFFFFFFF80000000ULL) | (s->mt[(i + 1) % 312] & 0x7FFFFFFFULL);
            uint64_t xa = x >> 1;
            if (x & 1ULL) xa ^= 0xB5026F5AA96619E9ULL;
            s->mt[i] = s->mt[(i + 156) % 312] ^ xa;
        }
        s->idx = 0;
    }
    uint64_t y = s->mt[s->idx++];
    y ^= (y >> 29) & 0x5555555555555555ULL;
    y ^= (y << 17) & 0x71D67FFFEDA60000ULL;
    y ^= (y << 37) & 0xFFF7EEE000000000ULL;
    y ^= (y >> 43);
    return y;
}

void cpo_rotation_signs(uint32_t D, uint64_t seed, float* signs) {
    mt64 rng;
    mt64_seed(&rng, seed);
    for (uint32_t layer = 0; layer < 3; ++layer)
        for (uint32_t i = 0; i < D; ++i)
            signs[layer * D + i] = (mt64_next(&rng) >> 63) ? 1.0f : -1.0f;
}

/* ======================================================================================
 * Unnormalised WHT with the reference's sign convention: encoder/transform/fht.hpp:23-57.
 * Strides 1,2,4 (inside one 8-lane group) produce (a+b, b-a); strides >= 8 produce
 * (a+b, a-b).  Only adds/subs, so any evaluation order of one butterfly level is bit-equal.
 * ==================================================================================== */
void cpo_fht(float* x, uint32_t D) {
    for (uint32_t h = 1; h < D; h <<= 1) {
        for (uint32_t i = 0; i < D; i += 2 * h) {
            for (uint32_t j = i; j < i + h; ++j) {
                float a = x[j], b = x[j + h];
                x[j] = a + b;
                x[j + h] = (h < 8) ? (b - a) : (a - b);
            }
        }
    }
}

/* ======================================================================================
 * Query encoding: api/hnsw_index.hpp:174-182 (pad), encoder/rotation.hpp:34-50 (3 x diag+WHT),
 * encoder/rabitq_encoder.hpp:201-204 (scale), :98-136 (build_lut), :37-39 (constants).
 * Fused ops as compiled: u = (int)fma(x - vl, inv_delta, 0.5f);
 *                        C = (-(fma(vl, D, delta*sum_qu))) * inv_sqrt_d.
 * ==================================================================================== */
static int cpo_encode_variant = 0;
void cpo_set_encode_variant(int v) { cpo_encode_variant = v; }

void cpo_encode_query(uint32_t dim, uint32_t D, const float* signs, const float* q,
                      uint8_t* lut, float coeffs[3], float* rotated) {
    float* buf = (float*)malloc(sizeof(float) * D);
    uint8_t* u8 = (uint8_t*)malloc(D);
    memcpy(buf, q, sizeof(float) * dim);
    for (uint32_t i = dim; i < D; ++i) buf[i] = 0.0f;

    for (uint32_t layer = 0; layer < 3; ++layer) {
        for (uint32_t i = 0; i < D; ++i) buf[i] = buf[i] * signs[layer * D + i];
        cpo_fht(buf, D);
    }
    float Df = (float)D;
    float norm_factor = 1.0f / (Df * sqrtf(Df));
    float inv_sqrt_d = 1.0f / sqrtf(Df);
    for (uint32_t i = 0; i < D; ++i) buf[i] = buf[i] * norm_factor;
    if (rotated) memcpy(rotated, buf, sizeof(float) * D);

    float vl = buf[0], vmax = buf[0];
    for (uint32_t i = 1; i < D; ++i) {
        if (buf[i] < vl) vl = buf[i];
        if (buf[i] > vmax) vmax = buf[i];
    }
    float delta = (vmax - vl) / 15.0f;
    if (delta < 1e-20f) delta = 1e-20f;
    float inv_delta = 1.0f / delta;

    float sum_qu = 0.0f;
    for (uint32_t i = 0; i < D; ++i) {
        float t = fmaf(buf[i] - vl, inv_delta, 0.5f);
        int u = (int)t;
        if (u > 15) u = 15;
        if (u < 0) u = 0;
        u8[i] = (uint8_t)u;
        sum_qu = sum_qu + (float)u;
    }
    for (uint32_t j = 0; j < D / 4; ++j) {
        for (uint32_t p = 0; p < 16; ++p) {
            uint8_t s = 0;
            for (uint32_t b = 0; b < 4; ++b)
                if (p & (1u << b)) s = (uint8_t)(s + u8[4 * j + b]);
            lut[j * 16 + p] = s;
        }
    }
    coeffs[0] = (2.0f * delta) * inv_sqrt_d;
    coeffs[1] = (2.0f * vl) * inv_sqrt_d;
    /* C = -(Df*vl + delta*sum_qu) * inv_sqrt_d: which product GCC fuses depends on the code around the inlined call.  The
     * reference module's search path (pinned by tests/golden/k1_golden.npz and every end-to-end test) has variant 0; the
     * calibration loop composed in oracle/refshim.cpp compiles to variant 1 (tests/test_oracle_calibration.py). */
    if (cpo_encode_variant == 1) coeffs[2] = (-fmaf(delta, sum_qu, Df * vl)) * inv_sqrt_d;
    else { float ds = delta * sum_qu; coeffs[2] = (-fmaf(vl, Df, ds)) * inv_sqrt_d; }
    free(buf);
    free(u8);
}

/* ======================================================================================
 * BUILD side (SURVEY section 8f, N3 -- oracle groundwork; no CUDA counterpart yet):
 * RaBitQEncoder<D>::compute_neighbor_aux, encoder/rabitq_encoder.hpp:138-181, with
 * rotate_raw_vector (:81-86) and compute_ip_cp (:88-96): the 1-bit code of a neighbour relative to
 * its parent vertex -- sign bits of the rotated unit offset, nop = |nb - parent|,
 * ip_qo = |rotated|_1 / sqrt(D), ip_cp = sum_i (+-1)_i rotated_parent_i / sqrt(D).
 * parent, nb: D floats (zero-padded); code: D/8 bytes, bit i = bit i%8 of byte i/8; aux = nop, ip_qo, ip_cp.
 * `fused` selects how `nop_sq += d*d` is taken to have been compiled (1: fmaf, 0: mul then add) -- pinned against
 * the compiled reference by tests/test_oracle_build_side.py: GCC 13.3 -O3 vectorises the squares and adds them one by
 * one, i.e. fused = 0 is what the reference computes (fused = 1 differs in nop for ~20 % of random inputs).
 * ==================================================================================== */
void cpo_neighbor_aux_1bit(uint32_t dim, uint32_t D, const float* signs, const float* parent, const float* nb,
                           int fused, uint8_t* code, float aux[3]) {
    float* diff = (float*)malloc(sizeof(float) * D);
    float* rp = (float*)malloc(sizeof(float) * D);
    float Df = (float)D;
    float norm_factor = 1.0f / (Df * sqrtf(Df));
    float inv_sqrt_d = 1.0f / sqrtf(Df);
    memset(code, 0, D / 8);
    /* rotate_raw_vector(parent) */
    memcpy(rp, parent, sizeof(float) * D);
    for (uint32_t layer = 0; layer < 3; ++layer) {
        for (uint32_t i = 0; i < D; ++i) rp[i] = rp[i] * signs[layer * D + i];
        cpo_fht(rp, D);
    }
    for (uint32_t i = 0; i < D; ++i) rp[i] = rp[i] * norm_factor;

    float nop_sq = 0.0f;
    for (uint32_t i = 0; i < dim; ++i) {
        diff[i] = nb[i] - parent[i];
        nop_sq = fused ? fmaf(diff[i], diff[i], nop_sq) : nop_sq + diff[i] * diff[i];
    }
    for (uint32_t i = dim; i < D; ++i) diff[i] = 0.0f;
    float nop = sqrtf(nop_sq);
    aux[0] = nop; aux[1] = 0.0f; aux[2] = 0.0f;
    if (nop < 1e-8f / Df) { free(diff); free(rp); return; }   /* constants::norm_epsilon(D) */
    float inv_nop = 1.0f / nop;
    for (uint32_t i = 0; i < D; ++i) diff[i] = diff[i] * inv_nop;
    for (uint32_t layer = 0; layer < 3; ++layer) {
        for (uint32_t i = 0; i < D; ++i) diff[i] = diff[i] * signs[layer * D + i];
        cpo_fht(diff, D);
    }
    float l1 = 0.0f, ip = 0.0f;
    for (uint32_t i = 0; i < D; ++i) {
        float r = diff[i] * norm_factor;
        if (r >= 0.0f) code[i >> 3] = (uint8_t)(code[i >> 3] | (1u << (i & 7)));
        l1 = l1 + fabsf(r);
    }
    for (uint32_t i = 0; i < D; ++i) {
        int bit = (code[i >> 3] >> (i & 7)) & 1;
        ip = bit ? ip + rp[i] : ip - rp[i];
    }
    aux[1] = l1 * inv_sqrt_d;
    aux[2] = ip * inv_sqrt_d;
    free(diff);
    free(rp);
}

/* ======================================================================================
 * BUILD side, N-bit: NbitRaBitQEncoder<D,B>::compute_neighbor_aux_nbit (encoder/rabitq_encoder.hpp:287-323) and the
 * coordinate-descent quantiser caq_quantize (:371-467).  All of it is scalar code whose `a*b + c` the compiler may or
 * may not fuse; `flags` says which (bit set = fused), one bit per expression, so that the combination the compiled
 * reference uses can be found by search (tests/test_oracle_build_side.py) and then pinned:
 *   bit 0  u = (int)((x - min) * inv_delta + 0.5f)          bit 1  dot_co   += c * x        (initial pass)
 *   bit 2  norm_c_sq += c * c                                bit 3  dot_without  = dot_co - old_c * x
 *   bit 4  norm_without = norm_c_sq - old_c * old_c          bit 5  new_dot  = dot_without + c * x   (trial and update)
 *   bit 6  new_norm = norm_without + c * c                   bit 7  ip_qo += c * x           (final pass)
 *   bit 8  ip_cp += c * rotated_parent                       bit 9  nop_sq += d * d
 * Pinned (CPO_CAQ_FLAGS = 0x1F9) against oracle/_ref on random (parent, neighbour) pairs, dims 64..960, B = 2, 4:
 *   decided by the search (any other value differs from the reference on some pair, B = 4): bits 1, 2, 9 unfused (the
 *   initial pass and the residual norm were vectorised: separate multiply, scalar reduction), bits 3, 5, 7, 8 fused;
 *   NOT decided by it (either value gave the reference's bits on all 1.6M pairs tried): bits 0, 4, 6 -- set to fused, which
 *   is what the same compiler did to the same-shaped expression next to each (bit 0: the query quantiser's
 *   (int)fma(x - vl, inv_delta, 0.5f), cpo_encode_query; bits 4, 6: the sibling of bits 3, 5 in the same statement pair).
 * codes: B planes of D/8 bytes, MSB first (NbitCodeStorage::set_value, core/codes.hpp:107-116).
 * ==================================================================================== */
static inline float caq_madd(int fused, float a, float b, float c) { return fused ? fmaf(a, b, c) : a * b + c; }
static inline float caq_msub(int fused, float c, float a, float b) { return fused ? fmaf(-a, b, c) : c - a * b; }

static float caq_quantize(uint32_t D, uint32_t B, const float* x, const float* rp, uint32_t flags, uint8_t* planes,
                          float inv_sqrt_d, float* out_ip_cp) {
    const int K_INT = (1 << B) - 1;
    const float K = (float)K_INT;
    const float Df = (float)D;
    float mn = x[0], mx = x[0];
    for (uint32_t i = 1; i < D; ++i) {
        if (x[i] < mn) mn = x[i];
        if (x[i] > mx) mx = x[i];
    }
    float delta = (mx - mn) / K;
    if (delta < 1e-10f / Df) delta = 1e-10f / Df;   /* constants::coordinate_epsilon(D) */
    const float inv_delta = 1.0f / delta;
    int* u = (int*)malloc(sizeof(int) * D);
    float dot_co = 0.0f, norm_c_sq = 0.0f;
    for (uint32_t i = 0; i < D; ++i) {
        const float t = caq_madd(flags & 1u, x[i] - mn, inv_delta, 0.5f);
        int v = (int)t;
        if (v < 0) v = 0;
        if (v > K_INT) v = K_INT;
        u[i] = v;
        const float c = (2.0f * (float)v - K) / K;
        dot_co = caq_madd(flags & 2u, c, x[i], dot_co);
        norm_c_sq = caq_madd(flags & 4u, c, c, norm_c_sq);
    }
    float prev_cos_sq = 0.0f;
    for (int iter = 0; iter < 10; ++iter) {
        int changed = 0;
        for (uint32_t i = 0; i < D; ++i) {
            const int old_u = u[i];
            const float old_c = (2.0f * (float)old_u - K) / K;
            const float dot_without = caq_msub(flags & 8u, dot_co, old_c, x[i]);
            const float norm_without = caq_msub(flags & 16u, norm_c_sq, old_c, old_c);
            int best_u = old_u;
            float best_dot = dot_co, best_norm = norm_c_sq;
            if (B >= 4) {
                for (int s = -1; s <= 1; s += 2) {
                    const int ut = old_u + s;
                    if (ut < 0 || ut > K_INT) continue;
                    const float c = (2.0f * (float)ut - K) / K;
                    const float nd = caq_madd(flags & 32u, c, x[i], dot_without);
                    const float nn = caq_madd(flags & 64u, c, c, norm_without);
                    if (nd * nd * best_norm > best_dot * best_dot * nn) { best_u = ut; best_dot = nd; best_norm = nn; }
                }
            } else {
                for (int ut = 0; ut <= K_INT; ++ut) {
                    if (ut == old_u) continue;
                    const float c = (2.0f * (float)ut - K) / K;
                    const float nd = caq_madd(flags & 32u, c, x[i], dot_without);
                    const float nn = caq_madd(flags & 64u, c, c, norm_without);
                    if (nd * nd * best_norm > best_dot * best_dot * nn) { best_u = ut; best_dot = nd; best_norm = nn; }
                }
            }
            if (best_u != old_u) {
                const float nc = (2.0f * (float)best_u - K) / K;
                dot_co = caq_madd(flags & 32u, nc, x[i], dot_without);
                norm_c_sq = caq_madd(flags & 64u, nc, nc, norm_without);
                u[i] = best_u;
                changed = 1;
            }
        }
        if (!changed) break;
        const float cos_sq = norm_c_sq > 0.0f ? (dot_co * dot_co / norm_c_sq) : 0.0f;
        if (iter > 0 && (cos_sq - prev_cos_sq) < 1e-4f) break;   /* constants::kCaqEarlyExitTol */
        prev_cos_sq = cos_sq;
    }
    float ip_qo = 0.0f, ip_cp = 0.0f;
    memset(planes, 0, (size_t)B * (D / 8));
    for (uint32_t i = 0; i < D; ++i) {
        for (uint32_t b = 0; b < B; ++b)
            if ((u[i] >> (B - 1 - b)) & 1) planes[b * (D / 8) + (i >> 3)] |= (uint8_t)(1u << (i & 7));
        const float c = (2.0f * (float)u[i] - K) / K;
        ip_qo = caq_madd(flags & 128u, c, x[i], ip_qo);
        ip_cp = caq_madd(flags & 256u, c, rp[i], ip_cp);
    }
    free(u);
    *out_ip_cp = ip_cp * inv_sqrt_d;
    return ip_qo * inv_sqrt_d;
}

void cpo_neighbor_aux_nbit(uint32_t dim, uint32_t D, uint32_t B, const float* signs, const float* parent, const float* nb,
                           uint32_t flags, uint8_t* planes, float aux[3]) {
    float* diff = (float*)malloc(sizeof(float) * D);
    float* rp = (float*)malloc(sizeof(float) * D);
    const float Df = (float)D;
    const float norm_factor = 1.0f / (Df * sqrtf(Df));
    const float inv_sqrt_d = 1.0f / sqrtf(Df);
    memset(planes, 0, (size_t)B * (D / 8));
    memcpy(rp, parent, sizeof(float) * D);
    for (uint32_t layer = 0; layer < 3; ++layer) {
        for (uint32_t i = 0; i < D; ++i) rp[i] = rp[i] * signs[layer * D + i];
        cpo_fht(rp, D);
    }
    for (uint32_t i = 0; i < D; ++i) rp[i] = rp[i] * norm_factor;
    float nop_sq = 0.0f;
    for (uint32_t i = 0; i < dim; ++i) {
        diff[i] = nb[i] - parent[i];
        nop_sq = caq_madd(flags & 512u, diff[i], diff[i], nop_sq);
    }
    for (uint32_t i = dim; i < D; ++i) diff[i] = 0.0f;
    const float nop = sqrtf(nop_sq);
    aux[0] = nop; aux[1] = 0.0f; aux[2] = 0.0f;
    if (nop < 1e-8f / Df) { free(diff); free(rp); return; }
    const float inv_nop = 1.0f / nop;
    for (uint32_t i = 0; i < D; ++i) diff[i] = diff[i] * inv_nop;
    for (uint32_t layer = 0; layer < 3; ++layer) {
        for (uint32_t i = 0; i < D; ++i) diff[i] = diff[i] * signs[layer * D + i];
        cpo_fht(diff, D);
    }
    for (uint32_t i = 0; i < D; ++i) diff[i] = diff[i] * norm_factor;
    float ip_cp = 0.0f;
    aux[1] = caq_quantize(D, B, diff, rp, flags, planes, inv_sqrt_d, &ip_cp);
    aux[2] = ip_cp;
    free(diff);
    free(rp);
}

/* ======================================================================================
 * FastScan integer sums: distance/fastscan_kernel.hpp:17-87 (one plane), :197-217 (N-bit),
 * :349-368 (top-two-planes "msb2").  Layout distance/fastscan_layout.hpp:10-49:
 * packed[sp][v] = (nibble(seg 2sp+1) << 4) | nibble(seg 2sp).  The AVX2 u8/u16 staging can
 * neither saturate nor wrap, so the result is the plain integer sum.
 * ==================================================================================== */
void cpo_fastscan_plane(uint32_t D, const uint8_t* lut, const uint8_t* packed, uint32_t out[32]) {
    uint32_t nsp = D / 8;
    for (uint32_t v = 0; v < 32; ++v) out[v] = 0;
    for (uint32_t sp = 0; sp < nsp; ++sp) {
        const uint8_t* lo = lut + (2 * sp) * 16;
        const uint8_t* hi = lut + (2 * sp + 1) * 16;
        for (uint32_t v = 0; v < 32; ++v) {
            uint8_t c = packed[sp * 32 + v];
            out[v] += (uint32_t)lo[c & 0x0F] + (uint32_t)hi[c >> 4];
        }
    }
}

void cpo_fastscan(uint32_t D, uint32_t B, const uint8_t* lut, const uint8_t* planes,
                  uint32_t nbit[32], uint32_t msb[32], uint32_t msb2[32]) {
    uint32_t plane_sums[32];
    size_t plane_bytes = (size_t)4 * D;
    for (uint32_t v = 0; v < 32; ++v) nbit[v] = 0;
    for (uint32_t b = 0; b < B; ++b) {
        cpo_fastscan_plane(D, lut, planes + b * plane_bytes, plane_sums);
        uint32_t w = 1u << (B - 1 - b);
        for (uint32_t v = 0; v < 32; ++v) {
            nbit[v] += w * plane_sums[v];
            if (b == 0) { msb[v] = plane_sums[v]; msb2[v] = plane_sums[v]; }
            if (b == 1) msb2[v] = 2 * msb2[v] + plane_sums[v];
        }
    }
}

/* ======================================================================================
 * Float epilogues.
 * ==================================================================================== */
static inline float max_ps(float a, float b) { return a > b ? a : b; } /* _mm256_max_ps(a,b) */
static inline float min_ps(float a, float b) { return a < b ? a : b; } /* _mm256_min_ps(a,b) */

/* One lane of the AVX2 8-wide loop of convert_to_distances_with_bounds
 * (distance/fastscan_kernel.hpp:138-173) == lower half of
 * convert_nbit_to_distances_with_bounds (:277-321) with its own (A,B,pc) for the estimate. */
static inline void lane_avx(float A_est, float B_est, float fs_est, float pc_est,
                            float A_lb, float B_lb, float fs_lb, float pc_lb, int same,
                            float C, float a, float b, float floor_, float slack,
                            float sqrt_dqp, float dqp, float nop, float ipqo, float ipcp,
                            float* est, float* lower) {
    float ip_approx = fmaf(A_est, fs_est, fmaf(B_est, pc_est, C));
    float q = max_ps(ipqo, floor_);
    float corr = ip_approx - ipcp;
    int good = q > 1e-10f;
    float e = good ? corr / q : 0.0f;
    e = fmaf(a, e, b);
    float d = fmaf(nop, nop, dqp);
    d = fmaf(-(2.0f * nop), e, d);
    *est = max_ps(d, 0.0f);

    float el = e;
    if (!same) {
        float ip_msb = fmaf(A_lb, fs_lb, fmaf(B_lb, pc_lb, C));
        float corr_m = ip_msb - ipcp;
        el = good ? corr_m / q : 0.0f;
        el = fmaf(a, el, b);
    }
    float cu = (el + slack) / max_ps(sqrt_dqp, 1e-10f);
    cu = min_ps(max_ps(cu, -1.0f), 1.0f);
    float lo = fmaf(nop, nop, dqp);
    lo = fmaf(-((2.0f * nop) * sqrt_dqp), cu, lo);
    lo = max_ps(lo, 0.0f);
    *lower = good ? lo : 0.0f;
}

/* Scalar tails (:176-193, :324-345) and convert_msb_to_lower_bounds (:403-424) as GCC 13.3
 * -O3 -mfma contracts them: t = A*fs; t = fma(pc, B, t); t += C; .... */
static inline float scalar_ip_est(float A, float Bc, float C, float fs, float pc, float ipcp,
                                  float q, float a, float b) {
    float t = A * fs;
    t = fmaf(pc, Bc, t);
    t = t + C;
    t = t - ipcp;
    t = t / q;
    return fmaf(t, a, b);
}

/* The plane-0 (lower-bound) chain inside the N-bit scalar tail (:339-342) is contracted the other way
 * round by the same compiler: t = B*pc; t = fma(A,fs,t); t += C  (pinned by oracle/probe_contraction.c). */
static inline float scalar_ip_est_msbtail(float A, float Bc, float C, float fs, float pc, float ipcp,
                                          float q, float a, float b) {
    float t = Bc * pc;
    t = fmaf(A, fs, t);
    t = t + C;
    t = t - ipcp;
    t = t / q;
    return fmaf(t, a, b);
}

static inline float scalar_lower(float e, float slack, float sqrt_dqp, float nop, float dqp) {
    float cu = (e + slack) / sqrt_dqp;
    if (cu < -1.0f) cu = -1.0f;
    if (cu > 1.0f) cu = 1.0f;
    float lo = fmaf(-((nop + nop) * sqrt_dqp), cu, fmaf(nop, nop, dqp));
    return lo < 0.0f ? 0.0f : lo;
}

void cpo_convert_1bit(const float p[7], const uint32_t* sums, const float* nop, const float* ip_qo,
                      const float* ip_cp, const uint16_t* pop, uint32_t count, float dqp,
                      float* est, float* lower) {
    float A = p[0], Bc = p[1], C = p[2], a = p[3], b = p[4], floor_ = p[5], slack = p[6];
    float sqrt_dqp = sqrtf(dqp);
    if (dqp < 1e-12f) {
        for (uint32_t i = 0; i < count; ++i) { est[i] = fmaf(nop[i], nop[i], dqp); lower[i] = 0.0f; }
        return;
    }
    uint32_t i = 0;
    for (; i + 8 <= count; i += 8)
        for (uint32_t l = i; l < i + 8; ++l)
            lane_avx(A, Bc, (float)sums[l], (float)pop[l], 0, 0, 0, 0, 1, C, a, b, floor_, slack,
                     sqrt_dqp, dqp, nop[l], ip_qo[l], ip_cp[l], &est[l], &lower[l]);
    for (; i < count; ++i) {
        float q = ip_qo[i] > floor_ ? ip_qo[i] : floor_;  /* std::max(ip_qo, floor) */
        float e;
        if (q > 1e-10f) e = scalar_ip_est(A, Bc, C, (float)sums[i], (float)pop[i], ip_cp[i], q, a, b);
        else e = fmaf(0.0f, a, b);
        float d = fmaf(-(nop[i] + nop[i]), e, fmaf(nop[i], nop[i], dqp));
        est[i] = d < 0.0f ? 0.0f : d;
        if (!(q > 1e-10f)) { lower[i] = 0.0f; continue; }
        lower[i] = scalar_lower(e, slack, sqrt_dqp, nop[i], dqp);
    }
}

void cpo_convert_msb(uint32_t B, const float p[7], const uint32_t* msb2, const float* nop,
                     const float* ip_qo, const float* ip_cp, const uint16_t* pop,
                     uint32_t count, float dqp, float* lower) {
    float kpartial = (B < 2) ? 1.0f : 3.0f;
    float inv_kp = 1.0f / kpartial;
    float A = p[0] * inv_kp, Bc = p[1] * inv_kp, C = p[2], a = p[3], b = p[4], floor_ = p[5], slack = p[6];
    float sqrt_dqp = sqrtf(dqp);
    if (dqp < 1e-12f) { for (uint32_t i = 0; i < count; ++i) lower[i] = 0.0f; return; }
    for (uint32_t i = 0; i < count; ++i) {
        float q = ip_qo[i] > floor_ ? ip_qo[i] : floor_;
        if (!(q > 1e-10f)) { lower[i] = 0.0f; continue; }
        float e = scalar_ip_est(A, Bc, C, (float)msb2[i], (float)pop[i], ip_cp[i], q, a, b);
        lower[i] = scalar_lower(e, slack, sqrt_dqp, nop[i], dqp);
    }
}

void cpo_convert_nbit(uint32_t B, const float p[7], const uint32_t* nbit, const uint32_t* msb,
                      const float* nop, const float* ip_qo, const float* ip_cp,
                      const uint16_t* pop, const uint16_t* wpop,
                      uint32_t count, float dqp, float* est, float* lower) {
    float K = (float)((1u << B) - 1);
    float inv_K = 1.0f / K;
    float A_n = p[0] * inv_K, B_n = p[1] * inv_K, A_m = p[0], B_m = p[1];
    float C = p[2], a = p[3], b = p[4], floor_ = p[5], slack = p[6];
    float sqrt_dqp = sqrtf(dqp);
    if (dqp < 1e-12f) {
        for (uint32_t i = 0; i < count; ++i) { est[i] = fmaf(nop[i], nop[i], dqp); lower[i] = 0.0f; }
        return;
    }
    uint32_t i = 0;
    for (; i + 8 <= count; i += 8)
        for (uint32_t l = i; l < i + 8; ++l)
            lane_avx(A_n, B_n, (float)nbit[l], (float)wpop[l], A_m, B_m, (float)msb[l], (float)pop[l], 0,
                     C, a, b, floor_, slack, sqrt_dqp, dqp, nop[l], ip_qo[l], ip_cp[l], &est[l], &lower[l]);
    for (; i < count; ++i) {
        float q = ip_qo[i] > floor_ ? ip_qo[i] : floor_;
        float e;
        if (q > 1e-10f) e = scalar_ip_est(A_n, B_n, C, (float)nbit[i], (float)wpop[i], ip_cp[i], q, a, b);
        else e = fmaf(0.0f, a, b);
        float d = fmaf(-(nop[i] + nop[i]), e, fmaf(nop[i], nop[i], dqp));
        est[i] = d < 0.0f ? 0.0f : d;
        if (!(q > 1e-10f)) { lower[i] = 0.0f; continue; }
        float em = scalar_ip_est_msbtail(A_m, B_m, C, (float)msb[i], (float)pop[i], ip_cp[i], q, a, b);
        lower[i] = scalar_lower(em, slack, sqrt_dqp, nop[i], dqp);
    }
}

/* ======================================================================================
 * Exact distances: core/memory.hpp:65-96.  Eight FMA accumulator lanes (element i goes to
 * lane i%8, in increasing i), then (lo+hi) -> hadd -> hadd: r = ((s0+s1)+(s2+s3)) with
 * s[l] = acc[l]+acc[l+4].  SURVEY App. A6.
 * ==================================================================================== */
float cpo_dot(uint32_t D, const float* a, const float* b) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t i = 0; i < D; i += 8)
        for (uint32_t l = 0; l < 8; ++l) acc[l] = fmaf(a[i + l], b[i + l], acc[l]);
    float s0 = acc[0] + acc[4], s1 = acc[1] + acc[5], s2 = acc[2] + acc[6], s3 = acc[3] + acc[7];
    return (s0 + s1) + (s2 + s3);
}

float cpo_l2(uint32_t D, const float* a, const float* b) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t i = 0; i < D; i += 8)
        for (uint32_t l = 0; l < 8; ++l) {
            float d = a[i + l] - b[i + l];
            acc[l] = fmaf(d, d, acc[l]);
        }
    float s0 = acc[0] + acc[4], s1 = acc[1] + acc[5], s2 = acc[2] + acc[6], s3 = acc[3] + acc[7];
    return (s0 + s1) + (s2 + s3);
}

/* ======================================================================================
 * Heaps.  The reference uses libstdc++'s std::push_heap / pop_heap / sort_heap
 * (search/rabitq_search.hpp:17-49, :79-80).  Which of two equal keys surfaces first is
 * decided by those algorithms' array movements, so they are restated exactly
 * (bits/stl_heap.h: __push_heap, __adjust_heap, __pop_heap).
 * ==================================================================================== */
typedef struct { float est, lower; uint32_t id; } beam_entry;   /* BeamEntry :53-58 */
typedef struct { uint32_t id; float dist; } nn_entry;            /* SearchResult core/types.hpp:12-23 */

/* min-heap on est: comp(a,b) = a.est > b.est */
static void beam_push_heap(beam_entry* h, size_t hole, size_t top, beam_entry v) {
    while (hole > top) {
        size_t parent = (hole - 1) / 2;
        if (!(h[parent].est > v.est)) break;
        h[hole] = h[parent];
        hole = parent;
    }
    h[hole] = v;
}

static void beam_adjust_heap(beam_entry* h, size_t hole, size_t len, beam_entry v) {
    size_t top = hole, child = hole;
    while ((ptrdiff_t)child < ((ptrdiff_t)len - 1) / 2) {
        child = 2 * (child + 1);
        if (h[child].est > h[child - 1].est) child--;
        h[hole] = h[child];
        hole = child;
    }
    if ((len & 1) == 0 && (ptrdiff_t)child == ((ptrdiff_t)len - 2) / 2) {
        child = 2 * (child + 1);
        h[hole] = h[child - 1];
        hole = child - 1;
    }
    beam_push_heap(h, hole, top, v);
}

typedef struct { beam_entry* d; size_t n, cap; } beam_heap;

static void beam_push(beam_heap* b, beam_entry v) {
    if (b->n == b->cap) {
        b->cap = b->cap ? b->cap * 2 : 1024;
        b->d = (beam_entry*)realloc(b->d, b->cap * sizeof(beam_entry));
    }
    b->d[b->n++] = v;
    beam_push_heap(b->d, b->n - 1, 0, v);
}

static beam_entry beam_pop(beam_heap* b) {
    beam_entry topv = b->d[0];
    if (b->n > 1) {
        beam_entry v = b->d[b->n - 1];
        b->d[b->n - 1] = b->d[0];
        beam_adjust_heap(b->d, 0, b->n - 1, v);
    }
    b->n--;
    return topv;
}

/* max-heap on distance: comp(a,b) = a.dist < b.dist */
static void nn_push_heap(nn_entry* h, size_t hole, size_t top, nn_entry v) {
    while (hole > top) {
        size_t parent = (hole - 1) / 2;
        if (!(h[parent].dist < v.dist)) break;
        h[hole] = h[parent];
        hole = parent;
    }
    h[hole] = v;
}

static void nn_adjust_heap(nn_entry* h, size_t hole, size_t len, nn_entry v) {
    size_t top = hole, child = hole;
    while ((ptrdiff_t)child < ((ptrdiff_t)len - 1) / 2) {
        child = 2 * (child + 1);
        if (h[child].dist < h[child - 1].dist) child--;
        h[hole] = h[child];
        hole = child;
    }
    if ((len & 1) == 0 && (ptrdiff_t)child == ((ptrdiff_t)len - 2) / 2) {
        child = 2 * (child + 1);
        h[hole] = h[child - 1];
        hole = child - 1;
    }
    nn_push_heap(h, hole, top, v);
}

/* std::pop_heap(first, first+len) */
static void nn_pop_heap(nn_entry* h, size_t len) {
    if (len > 1) {
        nn_entry v = h[len - 1];
        h[len - 1] = h[0];
        nn_adjust_heap(h, 0, len - 1, v);
    }
}

typedef struct { nn_entry* d; size_t n, cap; } nn_heap;

/* BoundedMaxHeap::push, search/rabitq_search.hpp:26-35 (no de-duplication: SURVEY F2) */
static void nn_push(nn_heap* h, nn_entry v) {
    if (h->n < h->cap) {
        h->d[h->n++] = v;
        nn_push_heap(h->d, h->n - 1, 0, v);
    } else if (v.dist < h->d[0].dist) {
        nn_pop_heap(h->d, h->n);
        h->d[h->n - 1] = v;
        nn_push_heap(h->d, h->n - 1, 0, v);
    }
}

static inline float nn_worst(const nn_heap* h) { return h->n == 0 ? FLT_MAX : h->d[0].dist; }

/* ======================================================================================
 * Upper-layer greedy descent: api/hnsw_index.hpp:617-638 with find_edge :468-474.
 * ==================================================================================== */
static int find_edge(const cpo_index* ix, int level, uint32_t node, uint32_t* begin, uint32_t* end) {
    uint32_t L = (uint32_t)(level - 1);
    const uint32_t* nodes = ix->layer_nodes[L];
    uint32_t lo = 0, hi = ix->layer_sizes[L];
    while (lo < hi) {
        uint32_t mid = lo + (hi - lo) / 2;
        if (nodes[mid] < node) lo = mid + 1; else hi = mid;
    }
    if (lo < ix->layer_sizes[L] && nodes[lo] == node) {
        *begin = ix->layer_offs[L][lo];
        *end = ix->layer_offs[L][lo + 1];
        return 1;
    }
    return 0;
}

static uint32_t greedy_layer(const cpo_index* ix, const float* q, uint32_t ep, int level, uint64_t* ndist) {
    uint32_t D = ix->D;
    float best = cpo_l2(D, q, ix->raw + (size_t)ep * D);
    if (ndist) (*ndist)++;
    uint32_t best_id = ep;
    int improved = 1;
    while (improved) {
        improved = 0;
        uint32_t b, e;
        if (!find_edge(ix, level, best_id, &b, &e)) break;
        for (uint32_t j = b; j < e; ++j) {
            uint32_t nb = ix->layer_nbrs[level - 1][j];
            float d = cpo_l2(D, q, ix->raw + (size_t)nb * D);
            if (ndist) (*ndist)++;
            if (d < best) { best = d; best_id = nb; improved = 1; }
        }
    }
    return best_id;
}

/* api/hnsw_index.hpp:195-202 */
uint32_t cpo_greedy_descent(const cpo_index* ix, const float* qpad, uint64_t* ndist) {
    uint32_t ep = ix->graph_entry_point;
    if (ix->max_level > 0) {
        ep = ix->entry_point;
        for (int level = ix->max_level; level >= 1; --level) ep = greedy_layer(ix, qpad, ep, level, ndist);
    }
    return ep;
}

/* ======================================================================================
 * Layer-0 Distance-Adaptive Beam Search: search/rabitq_search.hpp:60-277 driven by
 * Index::search api/hnsw_index.hpp:168-211.  Neighbour block field offsets: SURVEY App. B.
 * ==================================================================================== */
typedef struct {
    const uint8_t* planes; const float *nop, *ip_qo, *ip_cp;
    const uint16_t *pop, *wpop; const uint32_t* ids; uint32_t count;
} nb_view;

static nb_view get_nb(const cpo_index* ix, uint32_t id) {
    const uint8_t* nb = ix->search_data + (size_t)id * ix->rec_size + ix->nb_off;
    size_t o = (size_t)4 * ix->D * ix->B;
    nb_view v;
    v.planes = nb;
    v.nop = (const float*)(nb + o);
    v.ip_qo = (const float*)(nb + o + 128);
    v.ip_cp = (const float*)(nb + o + 256);
    v.pop = (const uint16_t*)(nb + o + 384);
    if (ix->B > 1) { v.wpop = (const uint16_t*)(nb + o + 448); o += 512; }
    else { v.wpop = NULL; o += 448; }
    v.ids = (const uint32_t*)(nb + o);
    v.count = *(const uint32_t*)(nb + o + 128);
    return v;
}

int cpo_search(const cpo_index* ix, const float* query, uint64_t k,
               uint32_t* out_ids, float* out_dists, cpo_stats* st) {
    const uint32_t D = ix->D, B = ix->B;
    cpo_stats local;
    if (!st) st = &local;
    memset(st, 0, sizeof(*st));

    float* qpad = (float*)calloc(D, sizeof(float));
    memcpy(qpad, query, sizeof(float) * ix->dim);
    uint8_t* lut = (uint8_t*)malloc((size_t)D * 4);
    float params[7];
    cpo_encode_query(ix->dim, D, ix->signs, qpad, lut, params, NULL);
    params[3] = ix->affine_a; params[4] = ix->affine_b; params[5] = ix->ip_qo_floor;
    params[6] = ix->slack_levels[0];
    if (k < 1) k = 1;
    const float gamma = ix->search_gamma;

    uint32_t ep = cpo_greedy_descent(ix, qpad, &st->descent_dists);

    /* TwoLevelVisitationTable (graph/visitation_table.hpp:49-108): two independent sets */
    uint8_t* estimated = (uint8_t*)calloc(ix->n, 1);
    uint8_t* visited = (uint8_t*)calloc(ix->n, 1);

    beam_heap beam = {NULL, 0, 0};
    nn_heap nn; nn.d = (nn_entry*)malloc(sizeof(nn_entry) * (k + 1)); nn.n = 0; nn.cap = k;

    float gamma_q = gamma;
    double ratio_sum = 0.0, ratio_sq_sum = 0.0;
    uint64_t ratio_count = 0;

    const float qn = cpo_dot(D, qpad, qpad);
#define EXACT_L2(ID) ({ float _d = cpo_dot(D, qpad, ix->raw + (size_t)(ID) * D); \
                        float _r = (qn + ix->norm_sq[(ID)]) - 2.0f * _d; st->exact_calls++; \
                        _r < 0.0f ? 0.0f : _r; })

    beam_entry e0 = {EXACT_L2(ep), 0.0f, ep};
    beam_push(&beam, e0); st->beam_pushes++;
    estimated[ep] = 1; st->estimated++;

    uint32_t sums[32], msb[32], msb2[32];
    float est[32], lower[32];
    int slack_batch_count = 0;

    while (beam.n > 0) {
        beam_entry cur; int found = 0;
        while (beam.n > 0) {
            cur = beam_pop(&beam); st->pops++;
            if (visited[cur.id]) continue;
            found = 1; break;
        }
        if (!found) break;

        if (nn.n >= k && cur.est >= gamma_q * nn_worst(&nn)) { st->gamma_terms++; break; }
        if (nn.n >= k && cur.lower > nn_worst(&nn)) { st->lb_skips++; continue; }

        visited[cur.id] = 1;
        float exact_dist = EXACT_L2(cur.id);
        nn_entry pe = {cur.id, exact_dist};
        nn_push(&nn, pe); st->nn_pushes++;
        st->expansions++;

        nb_view nb = get_nb(ix, cur.id);
        uint32_t nnb = nb.count;
        if (nnb == 0) continue;
        float dqp = exact_dist;

        if (ix->num_slack_levels > 0) {
            int li = slack_batch_count < ix->num_slack_levels - 1 ? slack_batch_count : ix->num_slack_levels - 1;
            params[6] = ix->slack_levels[li];
            ++slack_batch_count;
        }

        /* R = 32 = one batch (:147-207) */
        if (B == 1) {
            cpo_fastscan(D, 1, lut, nb.planes, sums, msb, msb2);
            cpo_convert_1bit(params, sums, nb.nop, nb.ip_qo, nb.ip_cp, nb.pop, nnb, dqp, est, lower);
        } else {
            cpo_fastscan(D, B, lut, nb.planes, sums, msb, msb2);
            cpo_convert_msb(B, params, msb2, nb.nop, nb.ip_qo, nb.ip_cp, nb.pop, nnb, dqp, lower);
            float threshold = nn_worst(&nn);
            int any = nn.n < k;
            if (!any) for (uint32_t j = 0; j < nnb; ++j) if (lower[j] < threshold) { any = 1; break; }
            if (any) {
                cpo_convert_nbit(B, params, sums, msb, nb.nop, nb.ip_qo, nb.ip_cp, nb.pop, nb.wpop,
                                 nnb, dqp, est, lower);
            } else {
                st->msb_skipped++;
                for (uint32_t j = 0; j < nnb; ++j) est[j] = FLT_MAX;
            }
        }

        int warmup = nn.n < k;
        for (uint32_t i = 0; i < nnb; ++i) {
            uint32_t nid = nb.ids[i];
            if (estimated[nid]) continue;
            estimated[nid] = 1; st->estimated++;

            float dabs = nn.n >= k ? gamma_q * nn_worst(&nn) : FLT_MAX;
            if (warmup) {
                float ex = EXACT_L2(nid);
                nn_entry ne = {nid, ex};
                nn_push(&nn, ne); st->nn_pushes++;
                if (ex < dabs) { beam_entry be = {ex, ex, nid}; beam_push(&beam, be); st->beam_pushes++; }
                continue;
            }
            float ed = est[i], lo = lower[i];
            if (nn.n >= k && lo >= nn_worst(&nn)) continue;
            if (ed < nn_worst(&nn)) {
                float ex = EXACT_L2(nid);
                nn_entry ne = {nid, ex};
                nn_push(&nn, ne); st->nn_pushes++;
                if (ex < dabs) { beam_entry be = {ex, lo, nid}; beam_push(&beam, be); st->beam_pushes++; }
                if (ex > 1e-12f) {
                    /* :255-267 as compiled: r = (double)(est/exact) (float divide);
                     * sums in double; see tests for the contraction check */
                    double r = (double)(ed / ex);
                    ratio_sum += r;
                    ratio_sq_sum = fma(r, r, ratio_sq_sum);
                    ++ratio_count;
                    if (ratio_count >= ix->gamma_warmup) {
                        double cnt = (double)ratio_count;
                        double r_mean = ratio_sum / cnt;
                        double r_var = fma(-r_mean, r_mean, ratio_sq_sum / cnt);
                        double r_std = sqrt(r_var > 0.0 ? r_var : 0.0);
                        float g = gamma * (float)fma((double)ix->gamma_beta, r_std, 1.0);
                        /* std::clamp(v, lo, hi) = v<lo ? lo : (hi<v ? hi : v) */
                        gamma_q = g < gamma ? gamma : (ix->gamma_max < g ? ix->gamma_max : g);
                    }
                }
            } else if (ed < dabs) {
                beam_entry be = {ed, lo, nid}; beam_push(&beam, be); st->beam_pushes++;
            }
        }
        if (beam.n > st->max_beam) st->max_beam = beam.n;
    }
#undef EXACT_L2

    /* extract_sorted: std::sort_heap */
    size_t m = nn.n;
    for (size_t len = m; len > 1; --len) nn_pop_heap(nn.d, len);
    size_t nout = m < k ? m : k;
    for (size_t i = 0; i < nout; ++i) { out_ids[i] = nn.d[i].id; out_dists[i] = nn.d[i].dist; }

    free(qpad); free(lut); free(estimated); free(visited); free(beam.d); free(nn.d);
    return (int)nout;
}

int cpo_search_batch(const cpo_index* ix, const float* queries, uint64_t nq, uint64_t k,
                     int64_t* ids, float* dists, cpo_stats* stats_sum, int num_threads) {
    uint64_t kk = k < 1 ? 1 : k;
    cpo_stats total;
    memset(&total, 0, sizeof(total));
#ifdef _OPENMP
    if (num_threads <= 0) num_threads = omp_get_max_threads();
#else
    (void)num_threads;
#endif
#pragma omp parallel num_threads(num_threads)
    {
        uint32_t* tid = (uint32_t*)malloc(sizeof(uint32_t) * kk);
        float* tdist = (float*)malloc(sizeof(float) * kk);
        cpo_stats acc;
        memset(&acc, 0, sizeof(acc));
#pragma omp for schedule(dynamic, 16)
        for (int64_t i = 0; i < (int64_t)nq; ++i) {
            cpo_stats s;
            int m = cpo_search(ix, queries + (size_t)i * ix->dim, k, tid, tdist, &s);
            uint64_t j = 0;
            for (; j < k && j < (uint64_t)m; ++j) { ids[i * k + j] = (int64_t)tid[j]; dists[i * k + j] = tdist[j]; }
            for (; j < k; ++j) { ids[i * k + j] = -1; dists[i * k + j] = FLT_MAX; }
            uint64_t* a = (uint64_t*)&acc; const uint64_t* b = (const uint64_t*)&s;
            for (size_t f = 0; f < sizeof(cpo_stats) / 8; ++f) {
                if (f == offsetof(cpo_stats, max_beam) / 8) { if (b[f] > a[f]) a[f] = b[f]; }
                else a[f] += b[f];
            }
        }
#pragma omp critical
        {
            uint64_t* a = (uint64_t*)&total; const uint64_t* b = (const uint64_t*)&acc;
            for (size_t f = 0; f < sizeof(cpo_stats) / 8; ++f) {
                if (f == offsetof(cpo_stats, max_beam) / 8) { if (b[f] > a[f]) a[f] = b[f]; }
                else a[f] += b[f];
            }
        }
        free(tid); free(tdist);
    }
    if (stats_sum) *stats_sum = total;
    return 0;
}

/* ======================================================================================
 * Exhaustive-scan oracle: the reference has no brute-force mode; this composes
 * its primitives: centre the query, encode, FastScan estimate of every per-vertex 1-bit code
 * (RaBitQCode<D>: signs @0, nop, ip_qo; encoder/rabitq_encoder.hpp:225-262) with ip_cp = 0 and
 * dist_qp_sq = ||q - c||^2, keep the k' smallest estimates, exact-L2 rerank, top-k.
 * Ties: (estimate, id) and (distance, id) ascending.
 * ==================================================================================== */
typedef struct { float key; uint32_t id; } kv;
static int kv_less(kv a, kv b) { return a.key < b.key || (a.key == b.key && a.id < b.id); }
static int kv_cmp(const void* a, const void* b) {
    kv x = *(const kv*)a, y = *(const kv*)b;
    return kv_less(x, y) ? -1 : (kv_less(y, x) ? 1 : 0);
}

int cpo_exhaustive_search(const cpo_index* ix, const cpo_flat_view* fv, const float* query,
                          uint64_t k, uint64_t kprime, uint64_t id_begin, uint64_t id_end,
                          uint32_t* out_ids, float* out_dists, uint32_t* est_sums, float* est_out) {
    const uint32_t D = ix->D;
    float* qc = (float*)calloc(D, sizeof(float));
    float* qpad = (float*)calloc(D, sizeof(float));
    memcpy(qpad, query, sizeof(float) * ix->dim);
    for (uint32_t i = 0; i < ix->dim; ++i) qc[i] = query[i] - fv->centroid[i];
    uint8_t* lut = (uint8_t*)malloc((size_t)D * 4);
    float params[7];
    cpo_encode_query(ix->dim, D, ix->signs, qc, lut, params, NULL);
    params[3] = ix->affine_a; params[4] = ix->affine_b; params[5] = ix->ip_qo_floor;
    params[6] = ix->slack_levels[0];
    float dqp = cpo_dot(D, qc, qc);

    uint64_t m = id_end - id_begin;
    kv* cand = (kv*)malloc(sizeof(kv) * (m ? m : 1));
    uint8_t* packed = (uint8_t*)malloc((size_t)4 * D);
    for (uint64_t g = id_begin; g < id_end; g += 32) {
        uint32_t cnt = (uint32_t)((id_end - g) < 32 ? (id_end - g) : 32);
        float nop[32], ipqo[32], ipcp[32], est[32], lower[32];
        uint16_t pop[32];
        uint32_t sums[32];
        memset(packed, 0, (size_t)4 * D);
        for (uint32_t j = 0; j < cnt; ++j) {
            const uint8_t* rec = fv->codes + (g + j) * fv->code_stride;
            uint32_t pc = 0;
            for (uint32_t sp = 0; sp < D / 8; ++sp) { packed[sp * 32 + j] = rec[sp]; pc += (uint32_t)__builtin_popcount(rec[sp]); }
            nop[j] = *(const float*)(rec + fv->nop_off);
            ipqo[j] = *(const float*)(rec + fv->ipqo_off);
            ipcp[j] = 0.0f;
            pop[j] = (uint16_t)pc;
        }
        for (uint32_t j = cnt; j < 32; ++j) { nop[j] = 0; ipqo[j] = 0; ipcp[j] = 0; pop[j] = 0; }
        cpo_fastscan_plane(D, lut, packed, sums);
        /* full 32-lane AVX2 path as for a complete block */
        cpo_convert_1bit(params, sums, nop, ipqo, ipcp, pop, 32, dqp, est, lower);
        for (uint32_t j = 0; j < cnt; ++j) {
            cand[g - id_begin + j].key = est[j];
            cand[g - id_begin + j].id = (uint32_t)(g + j);
            if (est_sums) est_sums[g - id_begin + j] = sums[j];
            if (est_out) est_out[g - id_begin + j] = est[j];
        }
    }
    qsort(cand, m, sizeof(kv), kv_cmp);
    uint64_t kp = kprime < m ? kprime : m;
    float qn = cpo_dot(D, qpad, qpad);
    for (uint64_t i = 0; i < kp; ++i) {
        uint32_t id = cand[i].id;
        float d = cpo_dot(D, qpad, ix->raw + (size_t)id * D);
        float r = (qn + ix->norm_sq[id]) - 2.0f * d;
        cand[i].key = r < 0.0f ? 0.0f : r;
    }
    qsort(cand, kp, sizeof(kv), kv_cmp);
    uint64_t nout = k < kp ? k : kp;
    for (uint64_t i = 0; i < nout; ++i) { out_ids[i] = cand[i].id; out_dists[i] = cand[i].key; }
    free(qc); free(qpad); free(lut); free(cand); free(packed);
    return (int)nout;
}

/* ======================================================================================
 * Calibration sampling: the lambda process_query of Index::calibrate_estimator,
 * api/hnsw_index.hpp:786-866 -- for one query and a start vertex: the closer of the start vertex and
 * its neighbours (strict <, stored order, stop at the first empty slot) becomes the parent; for every
 * neighbour of the parent: nop (floored at 1e-12), ip_corrected = ip_approx - ip_cp, the denominator
 * max(|ip_qo|, 1e-10) and the true inner product <q - p, o - p> / nop.
 * How GCC 13.3 -O3 -mfma contracts the two scalar expressions is not visible in the source; `flags`
 * carries the candidates and tests/test_oracle_calibration.py finds the one the compiled reference
 * uses (same method as probe_contraction.c):
 *   bit 0: ip_approx = fma(A', fs, Bc' * pc) + C   instead of   fma(Bc', pc, A' * fs) + C
 *   bit 1: true_ip accumulates with a separate multiply and add instead of one fma per dimension
 *   bit 2: ip_approx with no fused multiply-add at all: (A' * fs + Bc' * pc) + C
 * Outputs are [32] per query; slots past the block's end keep 0 / 0xFFFFFFFF.
 * ==================================================================================== */
void cpo_calibration_sample(const cpo_index* ix, const float* query_padded, uint32_t start_id, uint32_t flags, uint32_t* parent_out,
                            float* nn_dist_sq, float* dist_qp_sq, float* nop_out, float* ip_corrected, float* ip_qo_denom, float* true_ip,
                            uint32_t* neighbor) {
    const uint32_t D = ix->D, B = ix->B;
    uint32_t parent = start_id;
    float best = cpo_l2(D, query_padded, ix->raw + (size_t)parent * D);
    nb_view nb = get_nb(ix, parent);
    for (uint32_t i = 0; i < nb.count; ++i) {
        const uint32_t nid = nb.ids[i];
        if (nid == 0xFFFFFFFFu) break;
        const float d = cpo_l2(D, query_padded, ix->raw + (size_t)nid * D);
        if (d < best) { best = d; parent = nid; }
    }
    *nn_dist_sq = best;
    *parent_out = parent;
    *dist_qp_sq = cpo_l2(D, query_padded, ix->raw + (size_t)parent * D);
    uint8_t* lut = (uint8_t*)malloc((size_t)D * 4);
    float co[3];
    cpo_encode_query(D, D, ix->signs, query_padded, lut, co, NULL);
    nb_view pnb = get_nb(ix, parent);
    uint32_t nbit[32], msb[32], msb2[32];
    cpo_fastscan(D, B, lut, pnb.planes, nbit, msb, msb2);
    free(lut);
    for (uint32_t j = 0; j < 32; ++j) { nop_out[j] = 0.0f; ip_corrected[j] = 0.0f; ip_qo_denom[j] = 0.0f; true_ip[j] = 0.0f; neighbor[j] = 0xFFFFFFFFu; }
    const float inv_K = 1.0f / (float)((1u << B) - 1u);
    const float* p = ix->raw + (size_t)parent * D;
    for (uint32_t j = 0; j < pnb.count; ++j) {
        const uint32_t o_id = pnb.ids[j];
        if (o_id == 0xFFFFFFFFu) break;
        const float ipqo = pnb.ip_qo[j];
        const float nop = pnb.nop[j] > 1e-12f ? pnb.nop[j] : 1e-12f;   /* std::max(nop, kSmall) */
        const float fs = (float)nbit[j];
        const float pc = B == 1 ? (float)pnb.pop[j] : (float)pnb.wpop[j];
        const float Ae = B == 1 ? co[0] : co[0] * inv_K, Be = B == 1 ? co[1] : co[1] * inv_K;
        float t;
        if (flags & 4u) { const float m1 = Ae * fs, m2 = Be * pc; t = m1 + m2; }
        else if (flags & 1u) t = fmaf(Ae, fs, Be * pc);
        else t = fmaf(Be, pc, Ae * fs);
        const float ip_approx = t + co[2];
        ip_corrected[j] = ip_approx - pnb.ip_cp[j];
        const float a = fabsf(ipqo);
        ip_qo_denom[j] = a > 1e-10f ? a : 1e-10f;
        const float* o = ix->raw + (size_t)o_id * D;
        float acc = 0.0f;
        for (uint32_t d = 0; d < D; ++d) {
            const float x = query_padded[d] - p[d], y = o[d] - p[d];
            if (flags & 2u) { const float m = x * y; acc = acc + m; } else acc = fmaf(x, y, acc);
        }
        true_ip[j] = acc / nop;
        nop_out[j] = nop;
        neighbor[j] = o_id;
    }
}

