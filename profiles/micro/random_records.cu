// Micro-benchmark: what does HBM give a kernel that reads RANDOM records the way K3 does?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o random_records random_records.cu && ./random_records
// K3 (search.cu) reads, per expansion, one vertex record at a random place of a multi-GB index: a 2 704-byte neighbour block and a
// 512-byte raw vector (two bulk asynchronous copies), and it probes 32 random bits of a 128 KB per-warp bitmap with atomicOr.
// The roofline of bench.py divides by the STREAMING copy bandwidth (MEASURED_PEAKS.json); this program measures the ceiling of the
// access pattern itself: persistent warps, each keeping S records in flight (a ring of S shared-memory stages, one mbarrier
// each), optionally `think` nanoseconds of idle time per record (the compute of an expansion) and optionally the 32 atomics.
// Reports algorithmic GB/s (3 216 bytes per record) per configuration.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Args {
    const uint8_t* blocks; const uint8_t* raw; uint32_t* bitmaps;
    uint32_t nrec, block_stride, blk_bytes, raw_bytes, stages, iters, think_ns, atomics, evict_first, bitmap_words;
    unsigned long long* sink;
};

__global__ void __launch_bounds__(128) records_kernel(const Args a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const uint32_t stage_bytes = (a.blk_bytes + a.raw_bytes + 127u) & ~127u;
    uint8_t* st = smem + (size_t)warp * (a.stages * stage_bytes + 64);
    uint64_t* bar = reinterpret_cast<uint64_t*>(st + (size_t)a.stages * stage_bytes);
    const uint32_t slot = blockIdx.x * nw + warp;
    uint32_t* bitmap = a.bitmaps + (size_t)slot * a.bitmap_words;
    if (lane == 0) {
        for (uint32_t s = 0; s < a.stages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + s)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    uint64_t rng = 0x9E3779B97F4A7C15ull * (slot + 1);
    const uint64_t pol = a.evict_first ? 0x12F0000000000000ull : 0x1000000000000000ull;
    auto issue = [&](uint32_t s) {
        rng = rng * 6364136223846793005ull + 1442695040888963407ull;
        const uint32_t v = (uint32_t)((rng >> 33) % a.nrec);
        uint8_t* dst = st + (size_t)s * stage_bytes;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar + s)), "r"(a.blk_bytes + a.raw_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                     ::"r"(smem_u32(dst)), "l"(a.blocks + (size_t)v * a.block_stride), "r"(a.blk_bytes), "r"(smem_u32(bar + s)), "l"(pol) : "memory");
        if (a.raw_bytes)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                         ::"r"(smem_u32(dst + a.blk_bytes)), "l"(a.raw + (size_t)v * a.raw_bytes), "r"(a.raw_bytes), "r"(smem_u32(bar + s)), "l"(pol) : "memory");
    };
    if (lane == 0) for (uint32_t s = 0; s < a.stages; ++s) issue(s);
    uint32_t s = 0, phase = 0, acc = 0;
    for (uint32_t it = 0; it < a.iters; ++it) {
        uint32_t done;
        do {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(smem_u32(bar + s)), "r"(phase) : "memory");
        } while (!done);
        const uint32_t* w = reinterpret_cast<const uint32_t*>(st + (size_t)s * stage_bytes);
        const uint32_t id = w[512 + lane];   // where K3 finds the neighbour ids
        acc ^= id;
        if (a.atomics) {
            uint32_t h = (slot * 0x9E3779B1u) ^ (it * 0x85EBCA6Bu) ^ (lane * 0xC2B2AE35u) ^ id;
            h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
            const uint32_t b = h % (a.bitmap_words * 32u);
            if (a.atomics == 1) acc ^= atomicOr(&bitmap[b >> 5], 1u << (b & 31));
            else {   // 2: plain load + store (what a warp-private bitmap would allow, conflicts between lanes ignored here)
                const uint32_t old = __ldcg(&bitmap[b >> 5]);
                if (!(old & (1u << (b & 31)))) __stcg(&bitmap[b >> 5], old | (1u << (b & 31)));
                acc ^= old;
            }
        }
        if (a.think_ns) __nanosleep(a.think_ns);
        __syncwarp();
        if (lane == 0 && it + a.stages < a.iters) issue(s);
        if (++s == a.stages) { s = 0; phase ^= 1u; }
    }
    if (acc == 0x12345678u) a.sink[0] = acc;
}

int main(int argc, char** argv) {
    const uint32_t nrec = argc > 1 ? (uint32_t)atoi(argv[1]) : 1000000u;
    const uint32_t block_stride = 2816, blk = 2704, rawb = 512, bitmap_words = 32768;
    uint8_t *blocks, *raw; uint32_t* bitmaps; unsigned long long* sink;
    cudaMalloc(&blocks, (size_t)nrec * block_stride + 4096); cudaMalloc(&raw, (size_t)nrec * rawb + 4096);
    cudaMemset(blocks, 1, (size_t)nrec * block_stride + 4096); cudaMemset(raw, 1, (size_t)nrec * rawb + 4096);
    const int sms = 148;
    cudaMalloc(&bitmaps, (size_t)sms * 32 * bitmap_words * 4); cudaMemset(bitmaps, 0, (size_t)sms * 32 * bitmap_words * 4);   // 592 MB, as K3's
    cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    std::printf("# records: %u x (%u + %u) bytes, block stride %u; GB/s = records x 3216 B / time\n", nrec, blk, rawb, block_stride);
    std::printf("# warps/SM stages think_ns atomics evict_first  ->  ms  GB/s  records/us\n");
    struct Cfg { int warps_per_sm, stages, think, atomics, ef; uint32_t bw = 32768; };
    const Cfg cfgs[] = {
        {32, 1, 0, 0, 1}, {32, 2, 0, 0, 1}, {16, 4, 0, 0, 1}, {8, 8, 0, 0, 1}, {16, 2, 0, 0, 1}, {32, 2, 0, 0, 0},
        {32, 1, 0, 1, 1}, {32, 2, 0, 1, 1}, {16, 4, 0, 1, 1},
        {32, 1, 2000, 0, 1}, {32, 1, 4000, 0, 1}, {32, 1, 4000, 1, 1}, {32, 2, 4000, 1, 1}, {32, 2, 2000, 1, 1}, {28, 1, 4000, 1, 1},
        {32, 1, 3000, 1, 1}, {32, 1, 1000, 1, 1},
        // the same with smaller per-warp bitmaps (all of them together: 592 / 74 / 9 MB) and with plain load + store
        {32, 1, 0, 1, 1, 4096}, {32, 1, 0, 1, 1, 512}, {32, 1, 4000, 1, 1, 4096}, {32, 1, 4000, 1, 1, 512},
        {32, 1, 0, 2, 1, 32768}, {32, 1, 0, 2, 1, 4096}, {32, 1, 0, 2, 1, 512}, {32, 1, 4000, 2, 1, 32768}, {32, 1, 4000, 2, 1, 512},
    };
    const int only = argc > 2 ? atoi(argv[2]) : -1;   // one configuration by index (for ncu)
    int ci = -1;
    for (const Cfg& c : cfgs) {
        if (++ci != only && only >= 0) continue;
        Args a{};
        a.blocks = blocks; a.raw = raw; a.bitmaps = bitmaps; a.nrec = nrec; a.block_stride = block_stride; a.blk_bytes = blk; a.raw_bytes = rawb;
        a.stages = c.stages; a.think_ns = c.think; a.atomics = c.atomics; a.evict_first = c.ef; a.bitmap_words = c.bw; a.sink = sink;
        a.iters = c.think ? 1500 : 6000 / c.stages + 2000;
        const int ctas = sms * c.warps_per_sm / 4;
        const size_t smem = 4 * ((size_t)c.stages * 3328 + 64);
        cudaFuncSetAttribute(records_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            records_kernel<<<ctas, 128, smem>>>(a);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < best) best = ms;
        }
        cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) { std::printf("error: %s\n", cudaGetErrorString(err)); return 1; }
        const double recs = (double)ctas * 4 * a.iters;
        std::printf("%2d %d %4d %d %d bitmap %3.0f KB  ->  %.3f ms  %.0f GB/s  %.1f records/us\n", c.warps_per_sm, c.stages, c.think, c.atomics, c.ef, c.bw / 256.0, best,
                    recs * 3216.0 / best / 1e6, recs / best / 1e3);
    }
    return 0;
}
