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
// 4 exact distances at a time in the reference's 8-accumulator order, lane = tree node for the
// frontier heap, and thousands of queries in flight to cover HBM latency.
//
// Per-warp state: query (accumulator-major), query bit-planes, result list and the top levels of
// the frontier heap live in shared memory; the rest of the frontier and the "estimated" bitmap
// live in a per-slot HBM arena.
//
// Frontier.  The reference's frontier is std::priority_queue, i.e. libstdc++'s binary heap;
// equal estimates are common in large frontiers (a frontier of 10^4 floats in [200,300] has a few
// colliding pairs) and their pop order steers the traversal, so the heap is kept node for node
// identical to what __push_heap/__adjust_heap would build -- but executed by the whole warp:
//   push: the ancestors of the new leaf are an arithmetic sequence; lane l loads ancestor l, one
//         ballot finds how far the new entry rises, the displaced ancestors move down in parallel.
//   pop : __adjust_heap walks to a leaf always taking the child its comparator prefers (right,
//         unless right > left) and then sifts the former last entry v back up.  On a valid heap
//         the keys along that walk are non-decreasing, so v ends exactly below the last walk node
//         whose key is <= v's, and everything deeper returns to where it was.  Hence: descend only
//         while key(preferred child) <= key(v).  The walk is done five levels per step: the 31
//         sibling pairs under the hole are loaded one per lane, two ballots (which sibling, does it
//         still move) give every lane the whole 5-level path, and the moves happen in parallel.
// HBM reads of an expansion (the vertex's 32-code neighbour block and its raw vector) are staged into
// shared memory by two bulk asynchronous copies (cp.async.bulk, the TMA engine) completing on a per-warp
// mbarrier; they are issued as soon as the frontier top is known and fly while the heap is re-ordered.
// The "visited" set of the reference is elided: an id enters the frontier only right after its
// first "estimated" mark, hence at most once, so is_visited() can never be true.
#include <float.h>

#include "device_math.cuh"
#include "kernels.h"

namespace cpb {

constexpr uint32_t kNNSmem = 128;  // result lists up to this k live in shared memory
constexpr uint32_t kHC = 63;       // frontier entries [0, 63) = tree levels 0..5 live in shared memory

__host__ __device__ inline uint32_t nn_smem_entries(uint32_t k) { return k <= kNNSmem ? ((k + 31u) & ~31u) : 0u; }

// bytes of a block that an expansion reads (codes + aux + count), rounded for the bulk copy
__host__ __device__ inline uint32_t block_copy_bytes(uint32_t aux_off) { return (aux_off + 644u + 15u) & ~15u; }
// (16-byte granularity everywhere: what the bulk copies and the 128-bit reads need.  Every byte not spent here is L1 for the
// frontier arena -- the carve-out comes in steps, see launch_search.)
__host__ __device__ inline uint32_t stage_raw_off(uint32_t aux_off) { return block_copy_bytes(aux_off); }

__host__ __device__ inline size_t smem_per_warp(uint32_t D, uint32_t B, uint32_t k) {
    const uint32_t T = D / 8, nch = (D > 128 ? D : 128) / 128, aux_off = B * nch * 512;
    size_t s = (size_t)stage_raw_off(aux_off) + (size_t)D * 4 + 16 + 64;   // staged block + raw vector + mbarrier + WarpState
    s += (size_t)8 * (T + 4) * 4 + (size_t)nch * 64 + (size_t)(kHC + 1) * 16 + (size_t)nn_smem_entries(k) * 8;
    return (s + 15) & ~(size_t)15;
}

struct WarpCtx {
    // shared memory
    float* qrow;      // this lane's accumulator row of the query
    const uint4* uq;  // query bit-planes
    uint4* hs;        // frontier entries {key, lower bound bits, id, -} [0, kHC); physical index = logical + 1 here and in
                      // the arena, so a sibling pair is one aligned 32-byte unit (one sector in HBM)
    // arena
    uint4* hg;        // frontier entries beyond kHC
    uint32_t* bitmap;
    float* nn_d;
    uint32_t* nn_i;
    uint32_t lane;
    uint32_t D, T;    // padded dimension and D/8 (compile-time constants in the D = 128 instantiation)
    struct WarpState* ws;
    const uint4* walk;  // per-lane constants of the frontier walk (shared by the CTA), see init_walk_table
};

// Per-query state that is touched rarely or only by uniform reads: kept in shared memory, not in
// registers (the kernel is register-bound at 64/thread; a spilled in-flight load costs its latency).
struct WarpState {
    double ratio_sum, ratio_sq_sum;   // gamma_q adaptation (search/rabitq_search.hpp:255-267)
    float A, Bc, C, qn;               // estimator coefficients of the query, |q|^2
    uint4 last;                       // the frontier's last entry {key, payload x, payload y, valid}, when known without reading it
                                      // back: one 16-byte record, written by the push and read by the pop in one access each
    float gamma_q;
    uint32_t ratio_count;
    float slack;                      // dot_slack of the next non-empty expansion (:141-145)
    uint32_t pad_;
};
static_assert(sizeof(WarpState) == 64, "WarpState must fit its shared-memory slot");

// Entry i wherever it lives: one generic address (the shared window or the arena), so reads and writes of entries are
// single 16-byte accesses with no branch on the address space.
__device__ __forceinline__ uint4* eptr(const WarpCtx& w, uint32_t i) { return (i < kHC ? w.hs : w.hg) + (i + 1); }
__device__ __forceinline__ float kget(const WarpCtx& w, uint32_t i) { return __uint_as_float(eptr(w, i)->x); }
__device__ __forceinline__ void eget(const WarpCtx& w, uint32_t i, float& key, uint2& pay) {
    const uint4 e = *eptr(w, i);
    key = __uint_as_float(e.x); pay = make_uint2(e.y, e.z);
}
__device__ __forceinline__ void eset(const WarpCtx& w, uint32_t i, float key, uint2 pay) {
    *eptr(w, i) = make_uint4(__float_as_uint(key), pay.x, pay.y, 0u);
}

// Lane L < 31 owns sibling pair L of the 5-level subtree under the hole (pair L = the children of subtree
// node L).  walk[L] = {mask of the pairs above pair L on the way to the subtree root, which child (bit = 1:
// right) each of them must prefer for the walk to reach pair L, levels below the hole minus one, index of the pair in its level}.
__device__ __forceinline__ void init_walk_table(uint4* tab, uint32_t L) {
    uint32_t m = 0, need = 0, a = L;
    while (a > 0 && L < 31) {
        const uint32_t p = (a - 1) >> 1, r = (a - 1) & 1u;   // pair a hangs under child r of pair p
        m |= 1u << p;
        need |= r << p;
        a = p;
    }
    const uint32_t dlev = 32u - __clz(L + 1);      // pair L sits dlev levels below the hole, jpair-th in its level
    tab[L] = make_uint4(m, need, dlev - 1, L + 1 - (1u << (dlev - 1)));
}

// std::push_heap of (vk, vp) onto a heap of n entries, comp(a,b) = a.est > b.est.  Warp-cooperative.
// ws->lk/lp track the entry at the last index (what the next pop re-inserts) without reading it back.
__device__ __forceinline__ void heap_push(const WarpCtx& w, uint32_t n, float vk, uint2 vp) {
    const uint32_t m = n + 1;                      // 1-based index of the new leaf
    const uint32_t l = w.lane;
    const uint32_t depth = 31u - __clz(m);         // number of ancestors
    const uint32_t anc = l < depth ? (m >> (l + 1)) - 1 : 0u;   // ancestor l (l = 0: parent)
    float ka = 0.0f;
    if (l < depth) ka = kget(w, anc);
    const unsigned up = __ballot_sync(kFull, l < depth && ka > vk);
    const uint32_t cnt = __ffs(~up) - 1;           // the new entry passes ancestors 0 .. cnt-1
    if (cnt == 0) {                                 // the common case: it stays a leaf
        if (l == 0) {
            eset(w, n, vk, vp); w.ws->last = make_uint4(__float_as_uint(vk), vp.x, vp.y, 1u);
        }
    } else {
        uint2 pa = make_uint2(0, 0);
        if (l < cnt) { float t; eget(w, anc, t, pa); }
        __syncwarp();
        if (l < cnt) eset(w, (m >> l) - 1, ka, pa);    // ancestor l moves to where ancestor l-1 (or the leaf) was
        if (l == cnt) eset(w, (m >> cnt) - 1, vk, vp);
        if (l == 0) w.ws->last = make_uint4(__float_as_uint(ka), pa.x, pa.y, 1u);   // the parent now sits in the leaf
    }
    __syncwarp();
}

// One 5-level step of the pop walk under `hole`.  p = this lane's subtree node (parent of its sibling
// pair), kl/kr = its pair's keys (0 where absent).  Returns the number of levels moved (5 = go on) and
// updates hole.  SMEM: the whole step lives in shared memory (hole == 0).
template <bool SMEM>
__device__ __forceinline__ uint32_t pop_step(const WarpCtx& w, uint32_t& hole, uint32_t len, float vk) {
    const uint32_t L = w.lane;
    const uint4 wk = w.walk[L];
    const uint32_t p = SMEM ? L : ((hole + 1) << wk.z) - 1 + wk.w;
    const uint32_t left = 2 * p + 1;
    const bool hl = L < 31 && left < len, hr = L < 31 && left + 1 < len;
    // whole entries, the pair is one aligned 32-byte unit; below a hole outside shared memory every child is in the arena
    // loaded unconditionally (an absent child reads the pair at the end of the heap, inside the arena: len + 2 <= capacity + 1;
    // its keys are never used: hl / hr gate every decision)
    const uint4* pair = SMEM ? w.hs + ((L < 31 ? left : 0u) + 1) : w.hg + ((left < len ? left : len) + 1);
    const uint4 el = pair[0], er = pair[1];
    const float kl = __uint_as_float(el.x), kr = __uint_as_float(er.x);
    const bool right = hr && !(kr > kl);           // __adjust_heap: the right child unless right > left
    const bool ok = hl && (right ? kr : kl) <= vk;  // the preferred child still moves up
    const unsigned rmask = __ballot_sync(kFull, right);
    const unsigned omask = __ballot_sync(kFull, ok);
    const bool mv = ok && (omask & wk.x) == wk.x && (rmask & wk.x) == wk.y;
    const unsigned M = __ballot_sync(kFull, mv);   // the pairs on the walk: one per level, top down
    const uint32_t src = left + (right ? 1u : 0u);
    __syncwarp();
    if (mv) {
        const uint4 e = make_uint4(right ? er.x : el.x, right ? er.y : el.y, right ? er.z : el.z, 0u);   // (.w is never set)
        if (SMEM) w.hs[p + 1] = e; else *eptr(w, p) = e;
    }
    if (M) hole = __shfl_sync(kFull, src, 31 - __clz(M));
    return __popc(M);
}

// std::pop_heap + pop_back on a heap of n >= 1 entries.  Warp-cooperative.
__device__ __forceinline__ void heap_pop(const WarpCtx& w, uint32_t n) {
    if (n <= 1) { if (w.lane == 0) w.ws->last.w = 0u; return; }
    const uint32_t len = n - 1;                    // entries 0 .. len-1 remain, the hole starts at the root
    float vk;
    uint2 vp;
    const uint4 last = w.ws->last;
    if (last.w) { vk = __uint_as_float(last.x); vp = make_uint2(last.y, last.z); }   // known from the last push
    else eget(w, len, vk, vp);
    __syncwarp();
    if (w.lane == 0) w.ws->last.w = 0u;
    uint32_t hole = 0;
    uint32_t moved = pop_step<true>(w, hole, len, vk);
    while (moved == 5) {
        __syncwarp();
        moved = pop_step<false>(w, hole, len, vk);
    }
    __syncwarp();
    if (w.lane == 0) eset(w, hole, vk, vp);
    __syncwarp();
}

// ---- bulk asynchronous copies (TMA engine) completing on an mbarrier ----------------------------------
#ifndef CPB_HOST_EMULATION
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// The records an expansion reads (3 KB at a random place of a multi-GB index) are never read again by this query and
// hardly ever by another before L2 has turned over: they are fetched with an evict-first policy, which leaves L2 to the
// data that IS re-read -- the "estimated" bitmaps (32 atomics per expansion) and the frontier arenas.
// (The policy word is what `createpolicy.fractional.L2::evict_first.b64 p, 1.0` returns on sm_100a -- ptxas folds that
// instruction into seven uniform-datapath instructions per use; as a literal it is an operand.)
__device__ __forceinline__ uint64_t l2_evict_first_policy() { return 0x12F0000000000000ull; }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    } while (!done);
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#else
// host emulation (tests/native/): bar[0] counts completed phases, bar[1] the bytes the issuing thread still owes the
// current one (the kernel reserves 16 bytes per barrier); a bulk copy is a memcpy by the issuing thread
inline void mbar_init(uint64_t* bar, uint32_t) { bar[1] = 0; __atomic_store_n(&bar[0], 0, __ATOMIC_RELEASE); }
inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { bar[1] = bytes; }
inline uint64_t l2_evict_first_policy() { return 0; }
inline void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t) {
    memcpy(dst, src, bytes);
    bar[1] -= bytes;
    if (bar[1] == 0) __atomic_fetch_add(&bar[0], 1, __ATOMIC_RELEASE);
}
inline void mbar_wait(uint64_t* bar, uint32_t phase) {
    while ((__atomic_load_n(&bar[0], __ATOMIC_ACQUIRE) & 1u) == phase) std::this_thread::yield();
}
inline void prefetch_l2(const void*) {}
#endif

// BoundedMaxHeap (search/rabitq_search.hpp:17-49), node for node: the result set is libstdc++'s binary max-heap over
// (distance, id) entries compared by distance alone, filled with std::push_heap, and once full an entry strictly closer than
// front() replaces it through std::pop_heap / back() = v / std::push_heap.  No de-duplication.  WHICH of two entries
// of equal distance sits at the front -- and is the one evicted -- is a matter of the heap's layout, and with distinct ids of
// bit-equal distance (one query in ten thousand on the 250k benchmark index) it decides which id is in the answer; an
// ascending list that evicts its last entry, as rounds 1 and 2 had it, returns the other one.  So the heap is kept as the
// reference keeps it, by lane 0 alone: pushes that are accepted are rare once the set is full (the test against front() is
// warp-uniform and comes first), and the algorithms are sequential by nature.  d / id: shared memory up to k = 128, else
// the slot's arena.
__device__ __forceinline__ void nnh_push_heap(float* d, uint32_t* id, uint32_t hole, float vd, uint32_t vi) {   // std::__push_heap, top = 0
    while (hole > 0) {
        const uint32_t parent = (hole - 1) >> 1;
        if (!(d[parent] < vd)) break;
        d[hole] = d[parent]; id[hole] = id[parent];
        hole = parent;
    }
    d[hole] = vd; id[hole] = vi;
}
__device__ __forceinline__ void nnh_pop_heap(float* d, uint32_t* id, uint32_t len) {   // std::pop_heap(first, first + len)
    if (len <= 1) return;
    const uint32_t n = len - 1;                    // __adjust_heap(first, 0, n, value = the former last entry)
    const float vd = d[n]; const uint32_t vi = id[n];
    d[n] = d[0]; id[n] = id[0];
    uint32_t hole = 0, child = 0;
    while ((int)child < ((int)n - 1) / 2) {
        child = 2 * (child + 1);
        if (d[child] < d[child - 1]) --child;      // the larger child; the right one when equal
        d[hole] = d[child]; id[hole] = id[child];
        hole = child;
    }
    if ((n & 1u) == 0 && (int)child == ((int)n - 2) / 2) {
        child = 2 * (child + 1);
        d[hole] = d[child - 1]; id[hole] = id[child - 1];
        hole = child - 1;
    }
    nnh_push_heap(d, id, hole, vd, vi);
}
// BoundedMaxHeap::push.  m = entries held, front = front().distance (FLT_MAX while empty) -- both warp-uniform registers.
__device__ __forceinline__ void nn_push(const WarpCtx& w, uint32_t& m, uint32_t k, uint32_t id, float dist, float& front) {
    if (m == k && !(dist < front)) return;         // :31
    __syncwarp();
    if (w.lane == 0) {
        if (m < k) nnh_push_heap(w.nn_d, w.nn_i, m, dist, id);
        else {
            nnh_pop_heap(w.nn_d, w.nn_i, k);
            nnh_push_heap(w.nn_d, w.nn_i, k - 1, dist, id);
        }
    }
    if (m < k) ++m;
    __syncwarp();
    front = w.nn_d[0];
}

__device__ __forceinline__ float exact_group(const DevIndex& ix, const WarpCtx& w, uint32_t id, bool active,
                                             float qn) {
    const uint32_t l = w.lane & 7u;
    const float dot = group_chain<false>(ix.rawT + (size_t)id * w.D + (size_t)l * w.T, w.qrow, w.T, active);
    const float norm = active ? __ldg(ix.norm_sq + id) : 0.0f;
    return exact_from_dot(qn, norm, dot);
}

// exact distances of the lanes named in `mask` (each lane's own `nid`), four per round; the
// result lands in that lane's return value.
__device__ __forceinline__ float exact_for_lanes(const DevIndex& ix, const WarpCtx& w, unsigned mask, uint32_t nid,
                                                 float qn) {
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
            if ((int)w.lane == src[j]) mine = v;
        }
    }
    return mine;
}

// Planes 1 .. B-1 of a FEW slots of the staged block, evaluated by the whole warp.  Lane = slot (plane_sum_one) spends the
// same (B-1) x nch x 47 instructions whether one slot needs its other planes or all 32 do, and once the result list is full
// few do: on the 1M x 128 benchmark index 42 % of the expansions have no new slot under the plane-0 bound, 32 % one, 15 % two,
// 6 % three, 4.5 % more.  Here the (plane, chunk, word) items of a slot -- W = (B-1) x nch x 4 words, each worth
// sum_t 2^t popc(word & u_t) -- are dealt to the lanes of a group (16 lanes, two slots per pass, when W <= 16; else the warp,
// one slot per pass), summed across the group with shuffles and handed to the slot's own lane: the same integers, a
// quarter of the instructions.  Out, in the lanes named by `slots` (others: 0): p1 = plane 1's sum and
// rest = sum_{b >= 1} plane_b << (B-1-b), so that nbit = (plane_0 << (B-1)) + rest and msb2 = 2 plane_0 + p1.
template <int B>
__device__ __forceinline__ void slot_planes_by_warp(const uint8_t* stage, uint32_t nch, const uint4* uq, uint32_t lane,
                                                    unsigned slots, uint32_t& p1, uint32_t& rest) {
    p1 = 0; rest = 0;
    const uint32_t per_plane = nch * 4u, W = (uint32_t)(B - 1) * per_plane;   // nch is a power of two
    const uint32_t pshift = 31u - __clz(per_plane);
    const bool halves = W <= 16u;
    const uint32_t G = halves ? 16u : 32u, sub = lane & (G - 1u), grp = halves ? (lane >> 4) : 0u;
    while (slots) {
        const uint32_t j0 = __ffs(slots) - 1; slots &= slots - 1;
        uint32_t j1 = 32u;
        if (halves && slots) { j1 = __ffs(slots) - 1; slots &= slots - 1; }
        const uint32_t j = grp ? j1 : j0;
        uint32_t acc1 = 0, accr = 0;
        if (j < 32u) {
            for (uint32_t item = sub; item < W; item += G) {
                const uint32_t b = 1u + (item >> pshift), r = item & (per_plane - 1u), c = r >> 2, wi = r & 3u;
                const uint32_t word = reinterpret_cast<const uint32_t*>(stage)[(((size_t)b * nch + c) * 32u + j) * 4u + wi];
                const uint32_t* u = reinterpret_cast<const uint32_t*>(uq) + (size_t)c * 4u + wi;   // uq[t * nch + c], word wi
                const uint32_t v = __popc(word & u[0]) + 2u * __popc(word & u[(size_t)nch * 4u]) +
                                   4u * __popc(word & u[(size_t)nch * 8u]) + 8u * __popc(word & u[(size_t)nch * 12u]);
                accr += v << ((uint32_t)(B - 1) - b);
                if (b == 1u) acc1 += v;
            }
        }
        // sums over the group (xor shuffles up to G/2 stay inside an aligned group of G lanes)
        for (uint32_t o = G >> 1; o > 0; o >>= 1) {
            acc1 += __shfl_xor_sync(kFull, acc1, (int)o);
            accr += __shfl_xor_sync(kFull, accr, (int)o);
        }
        // to the slots' own lanes
        const uint32_t a1 = __shfl_sync(kFull, acc1, 16), ar = __shfl_sync(kFull, accr, 16);   // the second group's (halves)
        const uint32_t b1 = __shfl_sync(kFull, acc1, 0), br = __shfl_sync(kFull, accr, 0);     // the first group's
        if (lane == j0) { p1 = b1; rest = br; }
        if (lane == j1) { p1 = a1; rest = ar; }
    }
}

// greedy_search_layer over levels max_level..1 (api/hnsw_index.hpp:195-202, 617-638)
template <bool STATS>
__device__ __forceinline__ uint32_t greedy_descent(const DevIndex& ix, const WarpCtx& w, unsigned long long& ndist) {
    if (ix.max_level <= 0) return ix.graph_entry_point;
    const uint32_t g = w.lane >> 3, l = w.lane & 7u;
    uint32_t node = ix.entry_point, slot = ix.entry_slot;
    for (int L = ix.max_level; L >= 1; --L) {
        const bool have_level = (uint32_t)L <= ix.n_levels;
        float best = group_chain<true>(ix.rawT + (size_t)node * w.D + (size_t)l * w.T, w.qrow, w.T, true);
        if (STATS) ++ndist;
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
                const float d = group_chain<true>(ix.rawT + (size_t)nb * w.D + (size_t)l * w.T, w.qrow, w.T, act);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const float dt = __shfl_sync(kFull, d, t * 8);
                    const uint32_t nt = __shfl_sync(kFull, nb, t * 8), st = __shfl_sync(kFull, ns, t * 8);
                    if (j + t < e) {
                        if (STATS) ++ndist;
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

// DT = 128: the padded dimension is the compile-time constant 128 (SIFT/Deep shapes: one 128-dim chunk per
// code plane, 16-step distance chains, constant shared-memory offsets); DT = 0: any supported dimension.
// (Tried and not kept, again: an L2 prefetch of the block most likely to be expanded next -- the smaller child of the
// frontier's root -- one expansion ahead: 3 % slower with it on, and its mere presence behind a flag cost 9 %, the register
// allocation of this kernel being what it is.)
// (Tried and not kept: fetching the ancestors' keys of the next push right after the pop, so that the push at the end of the
// expansion does not wait for the arena -- 33 more instructions per expansion and more spills: 0.505 against 0.539.)
// (Tried and not kept: more registers for fewer warps -- 72 / 79 / 96 registers at 28 / 24 / 20 warps per SM lose 8 - 13 %,
// while 28 warps at 64 registers run as fast as 32.)
// (Tried and not kept: instantiations with the CTA shape and the per-warp shared-memory stride as compile-time constants.
// They remove the ~60 instructions per expansion that re-derive lane / warp / base from the thread id -- the kernel lives
// at the 64-register cap -- but run slower: 788 instead of 820 instructions per expansion, 66 % instead of 71 % issue
// utilisation, 3 % fewer QPS.  The kernel is bound by the latency of its dependent chain, not by issue slots.)
// (Tried and not kept, on one box against the same build without them, 250k x 128 x 4-bit, 541 k QPS: an L1 prefetch of the
// ancestors of the leaf the expansion's push will write, issued right after the pop with no register held -- the push's wait
// for those keys is the kernel's largest long-scoreboard stall, 7.4 % of warp time -- 530 k; plain popcounts instead of the
// carry-save compression, 12 issue slots and 33 ALU-pipe instructions fewer per expansion for 21 more POPC: 518 k; a larger L1
// (shared-memory carve-out 164 KB instead of 196 KB, which the trimmed per-warp footprint allows at 30 or 28 warps per SM):
// 540 k / 538 k.  The kernel sits on the issue, ALU and XU limits at once; none of them can be traded for another.)
// (Tried and not kept: the other planes evaluated BEFORE the probes' answers are back, i.e. whenever any slot, new or not, is
// under the plane-0 bound -- 98 % of the expansions instead of 58 %: -11 %.  The kernel does feel its ALU / XU instructions,
// which is what slot_planes_by_warp then saved.  A shared-memory cache of the key at the parent position of the leaf the last
// push wrote -- pushes and pops alternate, so push after push lands under the same parent, and a push that finds the key and
// stays a leaf needs neither the arena nor a ballot: -2 %.)
template <int B, bool STATS, int DT>
__global__ void __launch_bounds__(256, 4) search_kernel(const DevIndex ix, const SearchArgs a) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const uint32_t D = DT ? DT : ix.D, T = D / 8, nch = DT ? (DT > 128 ? DT : 128) / 128 : ix.nch, Tp = T + 4;
    const uint32_t aux_off = B * nch * 512;
    const uint32_t block_stride = DT ? ((aux_off + 644 + 127) & ~127u) : ix.block_stride;
    const uint32_t k = a.k;
    const uint32_t nq_work = a.nq_ptr ? *a.nq_ptr : a.nq;
    // up to how many slots slot_planes_by_warp is the cheaper way to their other planes (instruction estimates: lane per slot
    // 47 per plane and chunk; by warp 30 + 10 per item a lane handles, per pass of one or two slots)
    const uint32_t coop_W = (uint32_t)(B > 1 ? B - 1 : 0) * nch * 4u, coop_G = coop_W <= 16u ? 16u : 32u;
    const uint32_t coop_pass = 30u + 10u * ((coop_W + coop_G - 1u) / coop_G);
    const uint32_t coop_passes = (3u * (uint32_t)(B > 1 ? B - 1 : 0) * nch * 47u / 4u) / coop_pass;
    const uint32_t coop_max = min(8u, coop_passes * (coop_W <= 16u ? 2u : 1u));

    // ---- carve shared memory -------------------------------------------------------------------
    uint8_t* sm = smem_raw + warp * a.warp_smem;   // smem_per_warp(D, B, k), computed by the launcher
    WarpCtx w;
    w.lane = lane;
    w.D = D; w.T = T;
    CPB_BLOCK_SHARED uint4 walk_tab[32];
    if (warp == 0) init_walk_table(walk_tab, lane);
    w.walk = walk_tab;
    __syncthreads();
    const uint32_t blk_bytes = block_copy_bytes(aux_off), raw_off = stage_raw_off(aux_off);
    uint8_t* stage = sm;                                          sm += (size_t)raw_off + (size_t)D * 4;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(sm);             sm += 16;
    WarpState* ws = reinterpret_cast<WarpState*>(sm);             sm += 64;
    w.ws = ws;
    float* qs = reinterpret_cast<float*>(sm);                     sm += (size_t)8 * Tp * 4;
    uint4* uqs = reinterpret_cast<uint4*>(sm);                    sm += (size_t)nch * 64;
    w.hs = reinterpret_cast<uint4*>(sm);                          sm += (size_t)(kHC + 1) * 16;
    w.qrow = qs + (size_t)(lane & 7u) * Tp;
    w.uq = uqs;
    const uint32_t slot = blockIdx.x * nwarps + warp;
    uint8_t* arena = a.scratch + (size_t)slot * a.slot_stride;
    w.hg = reinterpret_cast<uint4*>(arena + a.heap_off);
    w.bitmap = a.bitmaps + (size_t)slot * a.bitmap_words;
    if (k <= kNNSmem) {
        w.nn_d = reinterpret_cast<float*>(sm);
        w.nn_i = reinterpret_cast<uint32_t*>(sm + (size_t)nn_smem_entries(k) * 4);
    } else {
        w.nn_d = reinterpret_cast<float*>(arena + a.nn_off);
        w.nn_i = reinterpret_cast<uint32_t*>(arena + a.nn_off + (size_t)k * 4);
    }
    const Calib& cal = ix.calib;
    Stats st{};
    if (lane == 0) mbar_init(mbar, 1);
    __syncwarp();
    uint32_t phase = 0;
    const uint8_t* aux = stage + aux_off;   // staged block: codes at 0, aux fields at aux_off

    for (;;) {
        uint32_t wi = 0;
        if (lane == 0) wi = atomicAdd(a.counters, 1u);
        wi = __shfl_sync(kFull, wi, 0);
        if (wi >= nq_work) break;
        const uint32_t q = a.query_list ? a.query_list[wi] : wi;

        // ---- stage the prepared query --------------------------------------------------------
        __syncwarp();
        {
            const float* src = a.qT + (size_t)q * D;
            for (uint32_t i = lane; i < D; i += 32) qs[(i / T) * Tp + (i % T)] = src[i];
            const uint4* us = reinterpret_cast<const uint4*>(a.uplanes + (size_t)q * 16 * nch);
            for (uint32_t i = lane; i < 4 * nch; i += 32) uqs[i] = us[i];
        }
        if (lane == 0) {
            const float* cf = a.coeffs + (size_t)q * kCoeffStride;
            ws->A = cf[0]; ws->Bc = cf[1]; ws->C = cf[2]; ws->qn = cf[3];
            ws->gamma_q = cal.gamma;
            ws->ratio_sum = 0.0; ws->ratio_sq_sum = 0.0; ws->ratio_count = 0u;
            ws->last = make_uint4(0u, 0u, 0u, 0u);
            ws->slack = cal.slack[0];
        }
        __syncwarp();
        const float qn = ws->qn;

        const uint32_t ep = greedy_descent<STATS>(ix, w, st.descent_dists);
        if (a.entry_out) { if (lane == 0) a.entry_out[q] = ep; continue; }

        // ---- layer-0 search state (search/rabitq_search.hpp:77-97) ----------------------------
        uint32_t heap_n = 0, nn_m = 0;
        float nn_front = FLT_MAX;   // front().distance of the result heap: the largest distance held (the k-th, once it is full)
        int slack_batch_count = 0;
        bool overflow = false;
        uint32_t max_beam = 0;

        {
            const float d0 = exact_group(ix, w, ep, true, qn);
            if (STATS) { ++st.exact_calls; ++st.beam_pushes; ++st.estimated; }
            if (lane == 0) {
                w.hs[1] = make_uint4(__float_as_uint(d0), __float_as_uint(0.0f), ep, 0u);
                atomicOr(&w.bitmap[ep >> 5], 1u << (ep & 31));
            }
            heap_n = 1;
            __syncwarp();
        }

        while (heap_n > 0) {
            // ---- pop (:110-117).  is_visited() can never hit: see file header ---------------------
            const uint4 top = w.hs[1];
            const float cur_est = __uint_as_float(top.x), cur_lower = __uint_as_float(top.y);
            const uint32_t cur = top.z;
            const bool full0 = nn_m >= k;
            float worst = full0 ? nn_front : FLT_MAX;
            const bool terminate = full0 && cur_est >= __fmul_rn(ws->gamma_q, worst);   // :120
            const bool lbskip = full0 && cur_lower > worst;                          // :122
            const bool expand = !terminate && !lbskip;
            // start the HBM reads of this expansion before the heap work: neighbour block and raw vector
            if (expand && lane == 0) {
                mbar_expect_tx(mbar, blk_bytes + D * 4);
                const uint64_t pol = l2_evict_first_policy();
                bulk_g2s(stage, ix.blocks + (size_t)cur * block_stride, blk_bytes, mbar, pol);
                bulk_g2s(stage + raw_off, ix.rawT + (size_t)cur * D, D * 4, mbar, pol);
            }
            heap_pop(w, heap_n);
            --heap_n;
            if (STATS) ++st.pops;
            if (terminate) { if (STATS) ++st.gamma_terms; break; }
            if (lbskip) { if (STATS) ++st.lb_skips; continue; }
            mbar_wait(mbar, phase);
            phase ^= 1u;

            // neighbour ids first: their "estimated" probes travel while the distances are computed
            const uint2 count_norm = *reinterpret_cast<const uint2*>(aux + 640);   // count, and the vertex's norm_sq riding in the block
            const uint32_t count = count_norm.x;
            const uint32_t nid = reinterpret_cast<const uint32_t*>(aux)[lane];
            const bool valid = lane < count;
            // ---- check_and_mark_estimated for all slots at once (:227); slots are distinct ids unless
            //      the index says otherwise, then only the first of equal ids may be new
            bool leader = valid;
            if (ix.dup_neighbors) {
                const unsigned peers = __match_any_sync(kFull, valid ? nid : (kInvalid - lane));
                leader = valid && (uint32_t)(__ffs(peers) - 1) == lane;
            }
            // (Tried and not kept: the bitmap belongs to this warp alone, so the probe could be a plain load here and a plain store
            // of the new bits where the result is first needed, with a __match_any_sync for the one expansion in sixty whose slots
            // share a word.  profiles/micro/random_records.cu has L2 executing 32 scattered atomics per record at about half the
            // rate of 32 loads plus stores -- but in the kernel the loads fetch two sectors per miss where the atomics fetch one
            // (DRAM reads per expansion 2 635 -> 2 994 bytes at 250k) and the time follows the DRAM bytes: 546 k -> 474 k QPS, with
            // cudaLimitMaxL2FetchGranularity at 32 bytes as without it.)
            uint32_t old = 0xFFFFFFFFu;
            if (leader) old = atomicOr(&w.bitmap[nid >> 5], 1u << (nid & 31));

            // plane-0 popcounts are independent of the distance chain below: issued together, the 16-step FMA
            // chain of the exact distance hides under them (a warp's expansion is one long dependent chain)
            uint32_t ps0 = 0;
            if (count > 0) ps0 = plane_sum_one<true>(reinterpret_cast<const uint4*>(stage), 0, nch, lane, w.uq);
            float exact_dist;   // :130-133, from the staged vector
            {
                const float dot = group_chain<false, true>(reinterpret_cast<const float*>(stage + raw_off) + (size_t)(lane & 7u) * T,
                                                           w.qrow, T, true);
                exact_dist = exact_from_dot(ws->qn, __uint_as_float(count_norm.y), dot);
            }
            nn_push(w, nn_m, k, cur, exact_dist, nn_front);
            if (STATS) { ++st.exact_calls; ++st.nn_pushes; ++st.expansions; }
            // (count == 0 -> `continue` in the reference: no lane is valid, nothing below acts)
            const float dqp = exact_dist;
            QParams qp;
            qp.A = ws->A; qp.Bc = ws->Bc; qp.C = ws->C;
            qp.a = cal.affine_a; qp.b = cal.affine_b; qp.floor_ = cal.ip_qo_floor;
            // :141-145: the slack level of this expansion = min(non-empty expansions so far, num_slack - 1); an empty block reads
            // no slack at all.  Kept in shared memory and advanced while it still changes, not looked up per expansion.
            qp.slack = ws->slack;
            if (count > 0 && slack_batch_count < cal.num_slack - 1) {
                ++slack_batch_count;
                __syncwarp();
                if (lane == 0) ws->slack = cal.slack[slack_batch_count];
                __syncwarp();
            }

            // ---- FastScan over the 32-code block + epilogue (:150-207); lane = neighbour slot -----
            // est / lower are only ever read for NEW slots of a non-warm-up expansion, and a slot whose lower
            // bound is not under the k-th distance is dropped before its estimate is looked at (:246).  The
            // N-bit lower bound needs plane 0 only, so: plane 0 and the bound now, the other planes (MSB-only
            // pre-filter :170-187, full estimate :189-200) only if some new slot survives the bound.
            const bool warmup = nn_m < k;   // :210, fixed for the whole neighbour loop
            float est = FLT_MAX, lower = 0.0f;
            const float sq = __fsqrt_rn(dqp);
            if (count > 0 && (!warmup || STATS)) {
                const float nop = reinterpret_cast<const float*>(aux + 128)[lane];
                const float ipqo = reinterpret_cast<const float*>(aux + 256)[lane];
                const float ipcp = reinterpret_cast<const float*>(aux + 384)[lane];
                const uint32_t pops = reinterpret_cast<const uint32_t*>(aux + 512)[lane];
                if (B == 1) convert_1bit(qp, ps0, nop, ipqo, ipcp, pops & 0xFFFFu, lane, count, dqp, sq, est, lower);
                else lower = nbit_lower<B>(qp, ps0, nop, ipqo, ipcp, pops & 0xFFFFu, lane, count, dqp, sq);
            }

            // planes 1 .. B-1, the MSB-pair bound (:170-187) and the full estimate (:189-200) in one go: the estimate is wanted unless
            // EVERY neighbour's MSB-pair bound reaches the k-th distance -- which this data never does -- so computing it before
            // that vote wastes nothing and lets its popcounts run under the bound's division chain
            float full_est = FLT_MAX, msb_lower = 0.0f;
            auto rest_of_estimate = [&]() {
                const float nop = reinterpret_cast<const float*>(aux + 128)[lane];
                const float ipqo = reinterpret_cast<const float*>(aux + 256)[lane];
                const float ipcp = reinterpret_cast<const float*>(aux + 384)[lane];
                const uint32_t pops = reinterpret_cast<const uint32_t*>(aux + 512)[lane];
                const uint32_t ps1 = plane_sum_one<true>(reinterpret_cast<const uint4*>(stage), 1, nch, lane, w.uq);
                uint32_t nbit = (ps0 << (B - 1)) + (ps1 << (B - 2));
#pragma unroll
                for (int b = 2; b < B; ++b)
                    nbit += plane_sum_one<true>(reinterpret_cast<const uint4*>(stage), b, nch, lane, w.uq) << (B - 1 - b);
                msb_lower = convert_msb<B>(qp, 2u * ps0 + ps1, nop, ipqo, ipcp, pops & 0xFFFFu, dqp, sq);
                full_est = nbit_est<B>(qp, nbit, nop, ipqo, ipcp, pops >> 16, lane, count, dqp);
            };
            const bool nbit_stage = B > 1 && count > 0 && (STATS || !warmup);
            const float w0 = nn_front;   // nn.worst_distance() (:179), the k-th distance when nn is full

            const bool isnew = leader && !(old & (1u << (nid & 31)));
            unsigned rem = __ballot_sync(kFull, isnew);
            if (STATS) st.estimated += __popc(rem);

            if (nbit_stage && (STATS || rem)) {
                const unsigned cand = __ballot_sync(kFull, isnew && !(lower >= w0));
                if (STATS || cand) {
                    bool done = false;
                    if (!STATS && (uint32_t)__popc(cand) <= coop_max) {
                        // few slots want their other planes: the warp evaluates just those (slot_planes_by_warp).  est is only
                        // ever read for new slots under the bound (a subset of cand: the k-th distance only shrinks), and the
                        // MSB-pair vote (:178-187) is decided as soon as ONE slot is under it -- if none of these is, the vote
                        // needs every slot's bound and the lane-per-slot evaluation below provides it.
                        uint32_t p1, rest;
                        slot_planes_by_warp<B>(stage, nch, w.uq, lane, cand, p1, rest);
                        const float nop = reinterpret_cast<const float*>(aux + 128)[lane];
                        const float ipqo = reinterpret_cast<const float*>(aux + 256)[lane];
                        const float ipcp = reinterpret_cast<const float*>(aux + 384)[lane];
                        const uint32_t pops = reinterpret_cast<const uint32_t*>(aux + 512)[lane];
                        const float ml = convert_msb<B>(qp, 2u * ps0 + p1, nop, ipqo, ipcp, pops & 0xFFFFu, dqp, sq);
                        const float fe = nbit_est<B>(qp, (ps0 << (B - 1)) + rest, nop, ipqo, ipcp, pops >> 16, lane, count, dqp);
                        if (__any_sync(kFull, ((cand >> lane) & 1u) && ml < w0)) { est = fe; done = true; }
                    }
                    if (!done) {
                        rest_of_estimate();
                        const bool any = nn_m < k || __any_sync(kFull, valid && msb_lower < w0);   // :178-187
                        if (any) est = full_est;
                        else {
                            if (STATS) ++st.msb_skipped;
                            est = FLT_MAX;
                            lower = msb_lower;
                        }
                    }
                }
                if (!warmup && !cand) rem = 0;   // no new slot passes :246 -- nothing in the neighbour loop can act
            }

            // ---- the sequential neighbour loop (:218-273), batched between state changes -----------
            if (warmup) {
                const float myex = exact_for_lanes(ix, w, rem, nid, qn);
                if (STATS) st.exact_calls += __popc(rem);
                while (rem) {
                    const int j = __ffs(rem) - 1;
                    rem &= rem - 1;
                    const float ex = __shfl_sync(kFull, myex, j);
                    const uint32_t id = __shfl_sync(kFull, nid, j);
                    const float dabs = nn_m >= k ? __fmul_rn(ws->gamma_q, nn_front) : FLT_MAX;   // :230-232
                    nn_push(w, nn_m, k, id, ex, nn_front);
                    if (STATS) ++st.nn_pushes;
                    if (ex < dabs) {
                        if (heap_n >= a.beam_capacity) { overflow = true; break; }
                        heap_push(w, heap_n, ex, make_uint2(__float_as_uint(ex), id));
                        ++heap_n;
                        if (STATS) ++st.beam_pushes;
                    }
                }
            } else if (rem) {
                worst = nn_front;
                // distances that may be needed: every new slot that passes both tests under the
                // current k-th distance (the k-th distance only shrinks, so this is a superset)
                const unsigned spec = __ballot_sync(kFull, isnew && !(lower >= worst) && est < worst);
                float myex = 0.0f;
                if (spec) { myex = exact_for_lanes(ix, w, spec, nid, qn); if (STATS) st.exact_calls += __popc(spec); }
                while (rem) {
                    const bool skip = lower >= worst;        // :246
                    const bool pex = !skip && est < worst;   // :248
                    const unsigned exm = __ballot_sync(kFull, pex) & rem;
                    const int first = exm ? __ffs(exm) - 1 : 32;
                    const unsigned batch = first < 32 ? (rem & ((1u << first) - 1u)) : rem;
                    const float dabs = __fmul_rn(ws->gamma_q, worst);
                    unsigned pm = __ballot_sync(kFull, !skip && !pex && est < dabs) & batch;   // :269-271
                    while (pm) {
                        const int j = __ffs(pm) - 1;
                        pm &= pm - 1;
                        if (heap_n >= a.beam_capacity) { overflow = true; break; }
                        heap_push(w, heap_n, __shfl_sync(kFull, est, j),
                                  make_uint2(__float_as_uint(__shfl_sync(kFull, lower, j)), __shfl_sync(kFull, nid, j)));
                        ++heap_n;
                        if (STATS) ++st.beam_pushes;
                    }
                    if (overflow) break;
                    rem &= ~batch;
                    if (first < 32) {   // :248-267
                        rem &= ~(1u << first);
                        const float ex = __shfl_sync(kFull, myex, first), ed = __shfl_sync(kFull, est, first);
                        const float lo = __shfl_sync(kFull, lower, first);
                        const uint32_t id = __shfl_sync(kFull, nid, first);
                        nn_push(w, nn_m, k, id, ex, nn_front);
                        if (STATS) ++st.nn_pushes;
                        if (ex < dabs) {
                            if (heap_n >= a.beam_capacity) { overflow = true; break; }
                            heap_push(w, heap_n, ex, make_uint2(__float_as_uint(lo), id));
                            ++heap_n;
                            if (STATS) ++st.beam_pushes;
                        }
                        if (ex > 1e-12f) {   // gamma_q adaptation (:255-267)
                            const double r = (double)__fdiv_rn(ed, ex);
                            const double rs = __dadd_rn(ws->ratio_sum, r), rq = __fma_rn(r, r, ws->ratio_sq_sum);
                            const uint32_t rc = ws->ratio_count + 1;
                            float gq = ws->gamma_q;
                            if ((unsigned long long)rc >= cal.gamma_warmup) {
                                const double cnt = (double)rc;
                                const double r_mean = __ddiv_rn(rs, cnt);
                                const double r_var = __fma_rn(-r_mean, r_mean, __ddiv_rn(rq, cnt));
                                const double r_std = __dsqrt_rn(r_var > 0.0 ? r_var : 0.0);
                                const float g = __fmul_rn(cal.gamma, (float)__fma_rn((double)cal.gamma_beta, r_std, 1.0));
                                gq = g < cal.gamma ? cal.gamma : (cal.gamma_max < g ? cal.gamma_max : g);
                            }
                            __syncwarp();
                            if (lane == 0) { ws->ratio_sum = rs; ws->ratio_sq_sum = rq; ws->ratio_count = rc; ws->gamma_q = gq; }
                            __syncwarp();
                        }
                        worst = nn_front;
                    }
                }
            }
            if (overflow) break;
            if (STATS && heap_n > max_beam) max_beam = heap_n;
        }

        // ---- results: extract_sorted (:37-40) then the padding of bindings.cpp:201-210 ------------
        __syncwarp();
        if (overflow) {
            if (lane == 0) { const uint32_t o = atomicAdd(a.counters + 1, 1u); a.overflow_list[o] = q; }
        } else if (a.kout > 0) {
            if (lane == 0)
                for (uint32_t len = nn_m; len > 1; --len) nnh_pop_heap(w.nn_d, w.nn_i, len);   // std::sort_heap: ascending
            __syncwarp();
            for (uint32_t j = lane; j < a.kout; j += 32) {
                const bool have = j < nn_m;
                a.ids[(size_t)q * a.kout + j] = have ? (int64_t)w.nn_i[j] : (int64_t)-1;
                a.dists[(size_t)q * a.kout + j] = have ? w.nn_d[j] : FLT_MAX;
            }
        }
        if (STATS && (unsigned long long)max_beam > st.max_beam) st.max_beam = max_beam;

        // ---- clear the estimated bitmap.  (Round 1 tracked which of 32 chunks had been touched and cleared only those: two
        //      instructions and a register per expansion, and after a few hundred marks of random ids every chunk is touched --
        //      a query marks tens of thousands.  The clear is 1 % of the bytes a query reads.) -----------------------------------
        __syncwarp();
        {
            uint4* p = reinterpret_cast<uint4*>(w.bitmap);
            for (uint32_t i = lane; i < a.bitmap_words / 4; i += 32) p[i] = make_uint4(0, 0, 0, 0);
        }
        __syncwarp();
    }

    if (STATS && a.stats && lane == 0) {
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

size_t search_smem_per_warp(const DevIndex& ix, uint32_t k, bool) { return smem_per_warp(ix.D, ix.B, k); }

#ifndef CPB_HOST_EMULATION   // tests/native/ compiles the kernels of this file for the host
typedef void (*SearchKernel)(const DevIndex, const SearchArgs);

template <int DT, bool ST>
static SearchKernel pick_kernel_b(uint32_t B) {
    return B == 1 ? search_kernel<1, ST, DT> : B == 2 ? search_kernel<2, ST, DT> : search_kernel<4, ST, DT>;
}
template <int DT>
static SearchKernel pick_kernel_d(uint32_t B, bool stats) {
    return stats ? pick_kernel_b<DT, true>(B) : pick_kernel_b<DT, false>(B);
}
static SearchKernel pick_kernel(const DevIndex& ix, bool stats, uint32_t) {
    return ix.D == 128 ? pick_kernel_d<128>(ix.B, stats) : pick_kernel_d<0>(ix.B, stats);
}

int search_max_ctas_per_sm(const DevIndex& ix, uint32_t k, int warps_per_cta, bool stats) {
    const size_t smem = search_smem_per_warp(ix, k, stats) * warps_per_cta;
    SearchKernel kern = pick_kernel(ix, stats, k);
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, warps_per_cta * 32, smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return nb;
}

cudaError_t launch_search(const DevIndex& ix, const SearchArgs& a, int ctas, int warps_per_cta, bool stats,
                          cudaStream_t stream) {
    const size_t smem = search_smem_per_warp(ix, a.k, stats) * warps_per_cta;
    SearchKernel kern = pick_kernel(ix, stats, a.k);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    SearchArgs b = a;
    b.warp_smem = (uint32_t)search_smem_per_warp(ix, a.k, stats);
    kern<<<ctas, warps_per_cta * 32, smem, stream>>>(ix, b);
    return cudaGetLastError();
}

#endif

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

#ifndef CPB_HOST_EMULATION
cudaError_t launch_exact_l2(const DevIndex& ix, const float* qT, const float* coeffs, uint32_t nq,
                            const uint32_t* ids, uint32_t m, float* out, cudaStream_t stream) {
    if (nq == 0 || m == 0) return cudaSuccess;
    const int warps = 4;
    const size_t smem = (size_t)warps * 8 * (ix.T + 4) * 4;
    exact_l2_kernel<<<(nq + warps - 1) / warps, warps * 32, smem, stream>>>(ix, qT, coeffs, nq, ids, m, out);
    return cudaGetLastError();
}

#endif

}  // namespace cpb
