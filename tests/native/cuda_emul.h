// Host emulation of the few CUDA device facilities a warp-synchronous kernel WITHOUT PTX needs, so that the
// kernel's own source can be compiled by g++ and checked on a machine with no GPU (test infrastructure only).
//
// One std::thread per CUDA thread of ONE block at a time.  threadIdx / blockIdx / blockDim are thread_local;
// __syncwarp / __shfl_sync / __ballot_sync meet at a per-warp std::barrier, __syncthreads at a per-block one;
// the *_rn intrinsics are the IEEE operations they name (compile with -ffp-contract=off so that g++ neither
// fuses nor splits anything).  Dynamic shared memory is the array the including file defines under the name
// the kernel declares (`extern __shared__ float smem[]` becomes a block-scope extern declaration).
#pragma once
#include <cuda_runtime.h>   // host side of the runtime headers: uint3/dim3, cudaError_t; __global__ etc. expand to nothing

#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include <algorithm>
using std::max;
using std::min;

#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif

namespace cuda_emul {
struct Warp {
    std::barrier<> bar{32};
    uint32_t slot[32];
};
struct Block {
    std::unique_ptr<std::barrier<>> bar;
    std::vector<std::unique_ptr<Warp>> warps;
};
inline thread_local Warp* t_warp = nullptr;
inline thread_local Block* t_block = nullptr;
}  // namespace cuda_emul

inline thread_local uint3 threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;

inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline float __fsqrt_rn(float a) { return std::sqrt(a); }
inline int __float2int_rz(float a) { return (int)a; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline float __uint_as_float(unsigned v) { float f; std::memcpy(&f, &v, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned v; std::memcpy(&v, &f, 4); return v; }
inline float __frcp_rn(float a) { return 1.0f / a; }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __dsqrt_rn(double a) { return std::sqrt(a); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }

inline void __syncwarp(unsigned = 0xFFFFFFFFu) { cuda_emul::t_warp->bar.arrive_and_wait(); }
inline void __syncthreads() { cuda_emul::t_block->bar->arrive_and_wait(); }

template <typename T>
inline T __shfl_sync(unsigned, T v, int src) {
    static_assert(sizeof(T) == 4, "32-bit shuffles only");
    cuda_emul::Warp* w = cuda_emul::t_warp;
    std::memcpy(&w->slot[threadIdx.x & 31], &v, 4);
    w->bar.arrive_and_wait();
    T r;
    std::memcpy(&r, &w->slot[src & 31], 4);
    w->bar.arrive_and_wait();
    return r;
}

template <typename T>
inline T __shfl_xor_sync(unsigned m, T v, int lane_mask) { return __shfl_sync(m, v, (int)(threadIdx.x & 31) ^ lane_mask); }
inline unsigned __reduce_add_sync(unsigned, unsigned v) {
    cuda_emul::Warp* w = cuda_emul::t_warp;
    w->slot[threadIdx.x & 31] = v;
    w->bar.arrive_and_wait();
    unsigned s = 0;
    for (int l = 0; l < 32; ++l) s += w->slot[l];
    w->bar.arrive_and_wait();
    return s;
}
// shfl.up: lanes below delta keep their own value
template <typename T>
inline T __shfl_up_sync(unsigned m, T v, unsigned delta) {
    const int lane = (int)(threadIdx.x & 31);
    return __shfl_sync(m, v, lane >= (int)delta ? lane - (int)delta : lane);
}
template <typename T>
inline T __ldg(const T* p) { return *p; }
template <typename T>
inline T __ldcg(const T* p) { return *p; }
template <typename T>
inline void __stcg(T* p, T v) { *p = v; }

inline unsigned __ballot_sync(unsigned, int pred) {
    cuda_emul::Warp* w = cuda_emul::t_warp;
    w->slot[threadIdx.x & 31] = pred ? 1u : 0u;
    w->bar.arrive_and_wait();
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) m |= w->slot[l] << l;
    w->bar.arrive_and_wait();
    return m;
}

inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }

inline unsigned __match_any_sync(unsigned, unsigned value) {
    cuda_emul::Warp* w = cuda_emul::t_warp;
    w->slot[threadIdx.x & 31] = value;
    w->bar.arrive_and_wait();
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) m |= (w->slot[l] == value ? 1u : 0u) << l;
    w->bar.arrive_and_wait();
    return m;
}

inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline uint32_t atomicOr(uint32_t* p, uint32_t v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) {
    unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xFFFFFFFFu; }

namespace cuda_emul {

// Runs kernel(args) for every block of the grid, one block after the other.  `smem`/`smem_bytes`: the shared-memory
// array, re-filled with garbage (0xFF: NaNs as floats) before each block so that a read of unwritten memory shows.
template <typename Kernel, typename Args>
void launch(Kernel kernel, dim3 grid, unsigned block, void* smem, size_t smem_bytes, const Args& args) {
    for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned b = 0; b < grid.x; ++b) {
        std::memset(smem, 0xFF, smem_bytes);
        Block blk;
        blk.bar = std::make_unique<std::barrier<>>(block);
        for (unsigned w = 0; w < (block + 31) / 32; ++w) blk.warps.push_back(std::make_unique<Warp>());
        std::vector<std::thread> threads;
        for (unsigned t = 0; t < block; ++t)
            threads.emplace_back([&, t] {
                threadIdx = uint3{t, 0, 0};
                blockIdx = uint3{b, by, 0};
                blockDim = dim3(block, 1, 1);
                gridDim = grid;
                t_block = &blk;
                t_warp = blk.warps[t / 32].get();
                kernel(args);
            });
        for (auto& th : threads) th.join();
    }
}

}  // namespace cuda_emul
