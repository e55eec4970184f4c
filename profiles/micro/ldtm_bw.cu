// Micro-benchmark: how fast can the warps of one SM read tensor memory?  (the K5 epilogue reads every accumulator once)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldtm_bw ldtm_bw.cu && ./ldtm_bw
// One CTA per SM, W warps (multiple of 4), each issuing `iters` tcgen05.ld.32x32b.x32 (32 lanes x 32 columns x 4 B = 4 KB)
// on its own lane quarter; reports bytes per clock per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void ldtm_kernel(int iters, unsigned long long* out, uint32_t* sink) {
    __shared__ uint32_t tmem_base_s;
    const uint32_t warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem_base_s + (((warp & 3u) * 32u) << 16) + ((warp >> 2) * 32u) % 480u;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        uint32_t r[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= r[j];
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc == 0x12345678u) sink[0] = acc;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(512u) : "memory");
}

int main() {
    unsigned long long* d_out; uint32_t* d_sink;
    cudaMalloc(&d_out, 148 * 8); cudaMalloc(&d_sink, 4);
    const int iters = 2000;
    for (int warps : {4, 8, 16, 32}) {
        ldtm_kernel<<<148, warps * 32>>>(iters, d_out, d_sink);
        ldtm_kernel<<<148, warps * 32>>>(iters, d_out, d_sink);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        unsigned long long h[148];
        cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
        double clk = 0; for (int i = 0; i < 148; ++i) clk += (double)h[i]; clk /= 148;
        const double bytes = (double)warps * iters * 4096.0;
        printf("warps %2d: %.0f clocks for %d x tcgen05.ld.32x32b.x32 per warp -> %.1f B/clk/SM (%.1f clk per 4 KB load)\n", warps, clk, iters, bytes / clk, clk / (iters * (warps / 4.0)) );
    }
    return 0;
}
