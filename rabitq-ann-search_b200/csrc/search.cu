// K3 (+K4 fused) -- upper-layer greedy descent and the layer-0 Distance-Adaptive Beam Search.
//
// Replaces Index::search (api/hnsw_index.hpp:168-211), greedy_search_layer (:617-638) and
// rabitq_search::search (search/rabitq_search.hpp:60-277) including BoundedMaxHeap (:17-49),
// the std::priority_queue frontier (:79-80) and TwoLevelVisitationTable
// (graph/visitation_table.hpp:49-108).
//
// Mapping: one WARP owns one query from descent to result (a CTA is just W such warps sharing
// an SM); the grid is persistent (ctas = SMs x resident CTAs) and warps pull queries from an
// atomic counter because work per query varies by 100x.  Inside a query the reference loop is
// strictly sequential (every decision reads the live k-th distance), so the parallel axes are:
// lane = neighbour slot for the 32-code FastScan block and its epilogue, 4 groups x 8 lanes =
// 4 exact distances at a time in the reference's 8-accumulator order, and thousands of queries
// in flight to cover HBM latency.
//
// Per-warp state: query (accumulator-major), query bit-planes, result list and the top of the
// frontier heap live in shared memory; the rest of the frontier heap and the "estimated" bitmap
// live in a per-slot HBM arena.  The frontier is the reference's binary heap, restated move for
// move (libstdc++ __push_heap/__adjust_heap) because equal estimates are common in large
// frontiers and their pop order decides the traversal.  The "visited" set of the reference is
// elided: an id enters the frontier only right after its first "estimated" mark, hence at most
// once, so is_visited() can never be true.
#include <float.h>

#include "device_math.cuh"
#include "kernels.h"

namespace cpb {

constexpr uint32_t kHeapCache = 64;   // frontier entries [0, kHeapCache) live in shared memory
constexpr uint32_t kNNSmem = 128;     // result lists up to this k live in shared memory

__host__ __device__ inline size_t smem_per_warp(uint32_t T, uint32_t nch, uint32_t k) {
    size_t s = (size_t)8 * (T + 4) * 4 + (size_t)nch * 64 + (size_t)kHeapCache * 16 + 32;
    if (k <= kNNSmem) s += (size_t)kNNSmem * 8;
    return (s + 15) & ~(size_t)15;
}

struct WarpCtx {
    // shared memory
    float* qrow;      // this lane's accumulator row of the query
    const uint4* uq;  // query bit-planes
    uint4* hs;        // frontier cache
    uint32_t* dirty;  // 256-bit summary of touched bitmap chunks
    // arena
    uint4* hg;        // frontier, physical index = logical + 1 (children share a 32-B sector)
    uint32_t* bitmap;
    float* nn_d;
    uint32_t* nn_i;
    uint32_t lane;
};

__device__ __forceinline__ uint4 hget(const WarpCtx& w, uint32_t i) { return i < kHeapCache ? w.hs[i] : w.hg[i + 1]; }
__device__ __forceinline__ void hset(const WarpCtx& w, uint32_t i, const uint4& e) {
    if (i < kHeapCache) w.hs[i] = e; else w.hg[i + 1] = e;
}
__device__ __forceinline__ float key(const uint4& e) { return __uint_as_float(e.x); }

// std::__push_heap with comp(a,b) = a.est > b.est (min-heap on the estimate)
__device__ __forceinline__ void heap_sift_up(const WarpCtx& w, uint32_t hole, const uint4& v) {
    const float vk = key(v);
    while (hole > 0) {
        const uint32_t parent = (hole - 1) >> 1;
        const uint4 pe = hget(w, parent);
        if (!(key(pe) > vk)) break;
        hset(w, hole, pe);
        hole = parent;
    }
    hset(w, hole, v);
}

// std::pop_heap + pop_back on a heap of n entries (lane 0 only)
__device__ __forceinline__ void heap_pop(const WarpCtx& w, uint32_t n) {
    if (n <= 1) return;
    const uint4 v = hget(w, n - 1);
    const int len = (int)n - 1;
    int hole = 0, child = 0;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        uint4 r = hget(w, child);
        const uint4 l = hget(w, child - 1);
        if (key(r) > key(l)) { child--; r = l; }
        hset(w, hole, r);
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        hset(w, hole, hget(w, child - 1));
        hole = child - 1;
    }
    heap_sift_up(w, (uint32_t)hole, v);
}

// BoundedMaxHeap::push (search/rabitq_search.hpp:26-35) on an ascending list: accept while not
// full, else replace the worst iff strictly closer.  No de-duplication.  Equal
// distances keep arrival order.  Warp-cooperative; all arguments warp-uniform.
__device__ __forceinline__ void nn_push(const WarpCtx& w, uint32_t& m, uint32_t k, uint32_t id, float dist) {
    if (m == k && !(dist < w.nn_d[k - 1])) return;
    const uint32_t newm = m < k ? m + 1 : k;
    for (int c = (int)((newm - 1) >> 5); c >= 0; --c) {
        const uint32_t i = (uint32_t)c * 32 + w.lane;
        float d0 = 0.0f, dm = 0.0f;
        uint32_t im = 0;
        const bool have0 = i < m, havem = i >= 1 && (i - 1) < m;
        if (have0) d0 = w.nn_d[i];
        if (havem) { dm = w.nn_d[i - 1]; im = w.nn_i[i - 1]; }
        const bool le0 = have0 && d0 <= dist;            // old[i] stays in place
        const bool lem = i == 0 || (havem && dm <= dist);  // old[i-1] stays in place
        __syncwarp();
        if (i < newm && !le0) {
            if (lem) { w.nn_d[i] = dist; w.nn_i[i] = id; }
            else { w.nn_d[i] = dm; w.nn_i[i] = im; }
        }
        __syncwarp();
        if (__all_sync(kFull, le0 || i >= newm)) break;  // nothing below this chunk moves
        if (__any_sync(kFull, i < newm && !le0 && lem)) break;  // insertion point passed
    }
    m = newm;
}

__device__ __forceinline__ float exact_group(const DevIndex& ix, const WarpCtx& w, uint32_t id, bool active,
                                             float qn) {
    const uint32_t l = w.lane & 7u;
    const float dot = group_chain<false>(ix.rawT + (size_t)id * ix.D + (size_t)l * ix.T, w.qrow, ix.T, active);
    const float norm = active ? __ldg(ix.norm_sq + id) : 0.0f;
    return exact_from_dot(qn, norm, dot);
}

// exact distances of the lanes named in `mask` (each lane's own `nid`), four per round; the
// result lands in that lane's return value.
__device__ __forceinline__ float exact_for_lanes(const DevIndex& ix, const WarpCtx& w, unsigned mask, uint32_t nid,
                                                 float qn, unsigned long long& calls) {
    float mine = 0.0f;
    const uint32_t g = w.lane >> 3;
    while (mask) {
        unsigned m = mask;
        int src[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { src[j] = m ? __ffs(m) - 1 : -1; m &= m - 1; }
        mask = m;
        const int mysrc = g == 0 ? src[0] : g == 1 ? src[1] : g == 2 ? src[2] : src[3];
        const uint32_t id = __shfl_sync(kFull, nid, mysrc < 0 ? 0 : mysrc);
        const float d = exact_group(ix, w, id, mysrc >= 0, qn);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float v = __shfl_sync(kFull, d, j * 8);
            if (src[j] >= 0) { ++calls; if ((int)w.lane == src[j]) mine = v; }
        }
    }
    return mine;
}

// greedy_search_layer over levels max_level..1 (api/hnsw_index.hpp:195-202, 617-638)
__device__ __forceinline__ uint32_t greedy_descent(const DevIndex& ix, const WarpCtx& w, unsigned long long& ndist) {
    if (ix.max_level <= 0) return ix.graph_entry_point;
    const uint32_t g = w.lane >> 3, l = w.lane & 7u;
    uint32_t node = ix.entry_point, slot = ix.entry_slot;
    for (int L = ix.max_level; L >= 1; --L) {
        const bool have_level = (uint32_t)L <= ix.n_levels;
        float best = group_chain<true>(ix.rawT + (size_t)node * ix.D + (size_t)l * ix.T, w.qrow, ix.T, true);
        ++ndist;
        uint32_t best_id = node, best_slot = slot;
        bool improved = have_level;
        while (improved) {
            improved = false;
            if (best_slot == kInvalid) break;
            const Level& lv = ix.levels[L - 1];
            const uint32_t b = __ldg(lv.offs + best_slot), e = __ldg(lv.offs + best_slot + 1);
            for (uint32_t j = b; j < e; j += 4) {
                const bool act = j + g < e;
                const uint32_t nb = act ? __ldg(lv.nbr_node + j + g) : 0;
                const uint32_t ns = act ? __ldg(lv.nbr_slot + j + g) : kInvalid;
                const float d = group_chain<true>(ix.rawT + (size_t)nb * ix.D + (size_t)l * ix.T, w.qrow, ix.T, act);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const float dt = __shfl_sync(kFull, d, t * 8);
                    const uint32_t nt = __shfl_sync(kFull, nb, t * 8), st = __shfl_sync(kFull, ns, t * 8);
                    if (j + t < e) {
                        ++ndist;
                        if (dt < best) { best = dt; best_id = nt; best_slot = st; improved = true; }
                    }
                }
            }
        }
        node = best_id;
        slot = (best_slot != kInvalid && have_level) ? __ldg(ix.levels[L - 1].down + best_slot) : kInvalid;
    }
    return node;
}

template <int B>
__global__ void __launch_bounds__(256, 2) search_kernel(const DevIndex ix, const SearchArgs a) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const uint32_t D = ix.D, T = ix.T, nch = ix.nch, Tp = T + 4;
    const uint32_t k = a.k;
    const bool nn_in_smem = k <= kNNSmem;

    // ---- carve shared memory -------------------------------------------------------------------
    const size_t per_warp = smem_per_warp(T, nch, k);
    uint8_t* sm = smem_raw + (size_t)warp * per_warp;
    WarpCtx w;
    w.lane = lane;
    float* qs = reinterpret_cast<float*>(sm);                     sm += (size_t)8 * Tp * 4;
    uint4* uqs = reinterpret_cast<uint4*>(sm);                    sm += (size_t)nch * 64;
    w.hs = reinterpret_cast<uint4*>(sm);                          sm += (size_t)kHeapCache * 16;
    w.dirty = reinterpret_cast<uint32_t*>(sm);                    sm += 32;
    w.qrow = qs + (size_t)(lane & 7u) * Tp;
    w.uq = uqs;
    const uint32_t slot = blockIdx.x * nwarps + warp;
    uint8_t* arena = a.scratch + (size_t)slot * a.slot_stride;
    w.hg = reinterpret_cast<uint4*>(arena + a.heap_off);
    w.bitmap = a.bitmaps + (size_t)slot * a.bitmap_words;
    if (nn_in_smem) {
        w.nn_d = reinterpret_cast<float*>(sm);
        w.nn_i = reinterpret_cast<uint32_t*>(sm + (size_t)kNNSmem * 4);
    } else {
        w.nn_d = reinterpret_cast<float*>(arena + a.nn_off);
        w.nn_i = reinterpret_cast<uint32_t*>(arena + a.nn_off + (size_t)k * 4);
    }
    if (lane < 8) w.dirty[lane] = 0;

    const Calib& cal = ix.calib;
    Stats st{};

    for (;;) {
        uint32_t wi = 0;
        if (lane == 0) wi = atomicAdd(a.counters, 1u);
        wi = __shfl_sync(kFull, wi, 0);
        if (wi >= a.nq) break;
        const uint32_t q = a.query_list ? a.query_list[wi] : wi;

        // ---- stage the prepared query --------------------------------------------------------
        __syncwarp();
        {
            const float* src = a.qT + (size_t)q * D;
            for (uint32_t i = lane; i < D; i += 32) qs[(i / T) * Tp + (i % T)] = src[i];
            const uint4* us = reinterpret_cast<const uint4*>(a.uplanes + (size_t)q * 16 * nch);
            for (uint32_t i = lane; i < 4 * nch; i += 32) uqs[i] = us[i];
        }
        const float* cf = a.coeffs + (size_t)q * kCoeffStride;
        QParams qp;
        qp.A = cf[0]; qp.Bc = cf[1]; qp.C = cf[2];
        qp.a = cal.affine_a; qp.b = cal.affine_b; qp.floor_ = cal.ip_qo_floor; qp.slack = cal.slack[0];
        const float qn = cf[3];
        __syncwarp();

        const uint32_t ep = greedy_descent(ix, w, st.descent_dists);
        if (a.entry_out) { if (lane == 0) a.entry_out[q] = ep; continue; }

        // ---- layer-0 search state (search/rabitq_search.hpp:77-97) ----------------------------
        uint32_t heap_n = 0, nn_m = 0;
        float gamma_q = cal.gamma;
        double ratio_sum = 0.0, ratio_sq_sum = 0.0;
        unsigned long long ratio_count = 0;
        int slack_batch_count = 0;
        bool overflow = false;
        uint32_t max_beam = 0;

        {
            const float d0 = exact_group(ix, w, ep, true, qn);
            ++st.exact_calls;
            if (lane == 0) {
                w.hs[0] = make_uint4(__float_as_uint(d0), __float_as_uint(0.0f), ep, 0u);
                atomicOr(&w.bitmap[ep >> 5], 1u << (ep & 31));
                const uint32_t ch = (ep >> 5) / a.chunk_words;
                w.dirty[ch >> 5] |= 1u << (ch & 31);
            }
            heap_n = 1; ++st.beam_pushes; ++st.estimated;
            __syncwarp();
        }

        while (heap_n > 0) {
            // ---- pop (:110-117).  is_visited() can never hit: see file header ---------------------
            const uint4 top = w.hs[0];
            const float cur_est = __uint_as_float(top.x), cur_lower = __uint_as_float(top.y);
            const uint32_t cur = top.z;
            if (lane == 0) heap_pop(w, heap_n);
            --heap_n; ++st.pops;
            __syncwarp();

            const bool full0 = nn_m >= k;
            float worst = full0 ? w.nn_d[k - 1] : FLT_MAX;
            if (full0 && cur_est >= __fmul_rn(gamma_q, worst)) { ++st.gamma_terms; break; }  // :120
            if (full0 && cur_lower > worst) { ++st.lb_skips; continue; }                      // :122

            const float exact_dist = exact_group(ix, w, cur, true, qn);   // :130-133
            ++st.exact_calls;
            nn_push(w, nn_m, k, cur, exact_dist);
            ++st.nn_pushes; ++st.expansions;

            const uint8_t* blk = ix.blocks + (size_t)cur * ix.block_stride;
            const uint8_t* aux = blk + ix.aux_off;
            const uint32_t count = __ldg(reinterpret_cast<const uint32_t*>(aux + 640));
            if (count == 0) continue;
            const float dqp = exact_dist;
            if (cal.num_slack > 0) {   // :141-145
                const int li = slack_batch_count < cal.num_slack - 1 ? slack_batch_count : cal.num_slack - 1;
                qp.slack = cal.slack[li];
                ++slack_batch_count;
            }

            // ---- FastScan over the 32-code block + epilogue (:150-207); lane = neighbour slot -----
            const uint32_t nid = __ldg(reinterpret_cast<const uint32_t*>(aux) + lane);
            const float nop = __ldg(reinterpret_cast<const float*>(aux + 128) + lane);
            const float ipqo = __ldg(reinterpret_cast<const float*>(aux + 256) + lane);
            const float ipcp = __ldg(reinterpret_cast<const float*>(aux + 384) + lane);
            const uint32_t pops = __ldg(reinterpret_cast<const uint32_t*>(aux + 512) + lane);
            uint32_t ps[B];
            plane_sums<B>(reinterpret_cast<const uint4*>(blk), nch, lane, w.uq, ps);
            uint32_t nbit, msb, msb2;
            combine_planes<B>(ps, nbit, msb, msb2);
            const bool valid = lane < count;
            float est, lower;
            if (B == 1) {
                convert_1bit(qp, nbit, nop, ipqo, ipcp, pops & 0xFFFFu, lane, count, dqp, est, lower);
            } else {
                lower = convert_msb<B>(qp, msb2, nop, ipqo, ipcp, pops & 0xFFFFu, dqp);
                const float threshold = nn_m ? w.nn_d[nn_m - 1] : FLT_MAX;   // nn.worst_distance()
                const bool any = nn_m < k || __any_sync(kFull, valid && lower < threshold);
                if (any) {
                    convert_nbit<B>(qp, nbit, msb, nop, ipqo, ipcp, pops & 0xFFFFu, pops >> 16, lane, count, dqp, est,
                                    lower);
                } else {
                    ++st.msb_skipped;
                    est = FLT_MAX;
                }
            }

            // ---- check_and_mark_estimated for all slots at once (:227) -----------------------------
            const unsigned peers = __match_any_sync(kFull, valid ? nid : (kInvalid - lane));
            bool isnew = false;
            if (valid && (uint32_t)(__ffs(peers) - 1) == lane) {
                const uint32_t wd = nid >> 5, bit = 1u << (nid & 31);
                const uint32_t old = atomicOr(&w.bitmap[wd], bit);
                isnew = !(old & bit);
                if (isnew) { const uint32_t ch = wd / a.chunk_words; atomicOr(&w.dirty[ch >> 5], 1u << (ch & 31)); }
            }
            unsigned rem = __ballot_sync(kFull, isnew);
            st.estimated += __popc(rem);

            // ---- the sequential neighbour loop (:218-273), batched between state changes -----------
            const bool warmup = nn_m < k;   // :210, fixed for the whole loop
            if (warmup) {
                const float myex = exact_for_lanes(ix, w, rem, nid, qn, st.exact_calls);
                while (rem) {
                    const int j = __ffs(rem) - 1;
                    rem &= rem - 1;
                    const float ex = __shfl_sync(kFull, myex, j);
                    const uint32_t id = __shfl_sync(kFull, nid, j);
                    const float dabs = nn_m >= k ? __fmul_rn(gamma_q, w.nn_d[k - 1]) : FLT_MAX;   // :230-232
                    nn_push(w, nn_m, k, id, ex);
                    ++st.nn_pushes;
                    if (ex < dabs) {
                        if (heap_n >= a.beam_capacity) { overflow = true; break; }
                        if (lane == 0) heap_sift_up(w, heap_n, make_uint4(__float_as_uint(ex), __float_as_uint(ex), id, 0u));
                        ++heap_n; ++st.beam_pushes;
                        __syncwarp();
                    }
                }
            } else {
                worst = w.nn_d[k - 1];
                // distances that may be needed: every new slot that passes both tests under the
                // current k-th distance (the k-th distance only shrinks, so this is a superset)
                const unsigned spec = __ballot_sync(kFull, isnew && !(lower >= worst) && est < worst);
                const float myex = exact_for_lanes(ix, w, spec, nid, qn, st.exact_calls);
                while (rem) {
                    const bool skip = lower >= worst;        // :246
                    const bool pex = !skip && est < worst;   // :248
                    const unsigned exm = __ballot_sync(kFull, pex) & rem;
                    const int first = exm ? __ffs(exm) - 1 : 32;
                    const unsigned batch = first < 32 ? (rem & ((1u << first) - 1u)) : rem;
                    const float dabs = __fmul_rn(gamma_q, worst);
                    unsigned pm = __ballot_sync(kFull, !skip && !pex && est < dabs) & batch;   // :269-271
                    while (pm) {
                        const int j = __ffs(pm) - 1;
                        pm &= pm - 1;
                        const uint4 e = make_uint4(__float_as_uint(__shfl_sync(kFull, est, j)),
                                                   __float_as_uint(__shfl_sync(kFull, lower, j)),
                                                   __shfl_sync(kFull, nid, j), 0u);
                        if (heap_n >= a.beam_capacity) { overflow = true; break; }
                        if (lane == 0) heap_sift_up(w, heap_n, e);
                        ++heap_n; ++st.beam_pushes;
                        __syncwarp();
                    }
                    if (overflow) break;
                    rem &= ~batch;
                    if (first < 32) {   // :248-267
                        rem &= ~(1u << first);
                        const float ex = __shfl_sync(kFull, myex, first), ed = __shfl_sync(kFull, est, first);
                        const float lo = __shfl_sync(kFull, lower, first);
                        const uint32_t id = __shfl_sync(kFull, nid, first);
                        nn_push(w, nn_m, k, id, ex);
                        ++st.nn_pushes;
                        if (ex < dabs) {
                            if (heap_n >= a.beam_capacity) { overflow = true; break; }
                            if (lane == 0) heap_sift_up(w, heap_n, make_uint4(__float_as_uint(ex), __float_as_uint(lo), id, 0u));
                            ++heap_n; ++st.beam_pushes;
                            __syncwarp();
                        }
                        if (ex > 1e-12f) {   // gamma_q adaptation (:255-267)
                            const double r = (double)__fdiv_rn(ed, ex);
                            ratio_sum = __dadd_rn(ratio_sum, r);
                            ratio_sq_sum = __fma_rn(r, r, ratio_sq_sum);
                            ++ratio_count;
                            if (ratio_count >= cal.gamma_warmup) {
                                const double cnt = (double)ratio_count;
                                const double r_mean = __ddiv_rn(ratio_sum, cnt);
                                const double r_var = __fma_rn(-r_mean, r_mean, __ddiv_rn(ratio_sq_sum, cnt));
                                const double r_std = __dsqrt_rn(r_var > 0.0 ? r_var : 0.0);
                                const float gq = __fmul_rn(cal.gamma, (float)__fma_rn((double)cal.gamma_beta, r_std, 1.0));
                                gamma_q = gq < cal.gamma ? cal.gamma : (cal.gamma_max < gq ? cal.gamma_max : gq);
                            }
                        }
                        worst = w.nn_d[k - 1];
                    }
                }
            }
            if (overflow) break;
            if (heap_n > max_beam) max_beam = heap_n;
        }

        // ---- results: extract_sorted (:37-40) then the padding of bindings.cpp:201-210 ------------
        __syncwarp();
        if (overflow) {
            if (lane == 0) { const uint32_t o = atomicAdd(a.counters + 1, 1u); a.overflow_list[o] = q; }
        } else if (a.kout > 0) {
            for (uint32_t j = lane; j < a.kout; j += 32) {
                const bool have = j < nn_m;
                a.ids[(size_t)q * a.kout + j] = have ? (int64_t)w.nn_i[j] : (int64_t)-1;
                a.dists[(size_t)q * a.kout + j] = have ? w.nn_d[j] : FLT_MAX;
            }
        }
        if ((unsigned long long)max_beam > st.max_beam) st.max_beam = max_beam;

        // ---- clear the touched chunks of the estimated bitmap -------------------------------------
        __syncwarp();
        for (uint32_t dw = 0; dw < 8; ++dw) {
            uint32_t bits = w.dirty[dw];
            while (bits) {
                const uint32_t ch = dw * 32 + (__ffs(bits) - 1);
                bits &= bits - 1;
                const uint32_t w0 = ch * a.chunk_words;
                uint4* p = reinterpret_cast<uint4*>(w.bitmap + w0);
                for (uint32_t i = lane; i < a.chunk_words / 4; i += 32) p[i] = make_uint4(0, 0, 0, 0);
            }
        }
        __syncwarp();
        if (lane < 8) w.dirty[lane] = 0;
    }

    if (a.stats && lane == 0) {
        atomicAdd(&a.stats->pops, st.pops);
        atomicAdd(&a.stats->expansions, st.expansions);
        atomicAdd(&a.stats->exact_calls, st.exact_calls);
        atomicAdd(&a.stats->beam_pushes, st.beam_pushes);
        atomicMax(&a.stats->max_beam, st.max_beam);
        atomicAdd(&a.stats->nn_pushes, st.nn_pushes);
        atomicAdd(&a.stats->lb_skips, st.lb_skips);
        atomicAdd(&a.stats->gamma_terms, st.gamma_terms);
        atomicAdd(&a.stats->msb_skipped, st.msb_skipped);
        atomicAdd(&a.stats->estimated, st.estimated);
        atomicAdd(&a.stats->descent_dists, st.descent_dists);
    }
}

size_t search_smem_per_warp(const DevIndex& ix, uint32_t k) { return smem_per_warp(ix.T, ix.nch, k); }

cudaError_t launch_search(const DevIndex& ix, const SearchArgs& a, int ctas, int warps_per_cta, cudaStream_t stream) {
    const size_t smem = search_smem_per_warp(ix, a.k) * warps_per_cta;
    void (*kern)(const DevIndex, const SearchArgs) =
        ix.B == 1 ? search_kernel<1> : ix.B == 2 ? search_kernel<2> : search_kernel<4>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<ctas, warps_per_cta * 32, smem, stream>>>(ix, a);
    return cudaGetLastError();
}

// ---- K4 primitive ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) exact_l2_kernel(const DevIndex ix, const float* __restrict__ qT,
                                                       const float* __restrict__ coeffs, uint32_t nq,
                                                       const uint32_t* __restrict__ ids, uint32_t m,
                                                       float* __restrict__ out) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const uint32_t q = blockIdx.x * nwarps + warp;
    if (q >= nq) return;
    const uint32_t D = ix.D, T = ix.T, Tp = T + 4;
    float* qs = reinterpret_cast<float*>(smem_raw) + (size_t)warp * 8 * Tp;
    for (uint32_t i = lane; i < D; i += 32) qs[(i / T) * Tp + (i % T)] = qT[(size_t)q * D + i];
    __syncwarp();
    const float qn = coeffs[(size_t)q * kCoeffStride + 3];
    const uint32_t g = lane >> 3, l = lane & 7u;
    for (uint32_t j = 0; j < m; j += 4) {
        const bool act = j + g < m;
        const uint32_t id = act ? ids[(size_t)q * m + j + g] : 0;
        const float dot = group_chain<false>(ix.rawT + (size_t)id * D + (size_t)l * T, qs + (size_t)l * Tp, T, act);
        if (act && l == 0) out[(size_t)q * m + j + g] = exact_from_dot(qn, __ldg(ix.norm_sq + id), dot);
    }
}

cudaError_t launch_exact_l2(const DevIndex& ix, const float* qT, const float* coeffs, uint32_t nq,
                            const uint32_t* ids, uint32_t m, float* out, cudaStream_t stream) {
    if (nq == 0 || m == 0) return cudaSuccess;
    const int warps = 4;
    const size_t smem = (size_t)warps * 8 * (ix.T + 4) * 4;
    exact_l2_kernel<<<(nq + warps - 1) / warps, warps * 32, smem, stream>>>(ix, qT, coeffs, nq, ids, m, out);
    return cudaGetLastError();
}

}  // namespace cpb
