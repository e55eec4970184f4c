// Stand-alone GPU check of cphnsw_b200_neighbor_codes against the C oracle (no Python): test infrastructure.
//   g++ -O1 -std=c++17 -I include -I oracle -I /usr/local/cuda/include tests/native/neighbor_codes_gpu_check.cpp \
//       rabitq-ann-search_b200/cphnsw_b200/libcphnsw_b200.so oracle/libcphnsw_oracle.so -L/usr/local/cuda/lib64 -lcudart -o <out>
// Prints one line per (dim, bits) case and "ALL OK" / "FAILED"; exit status 0 only when every bit matches.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "cphnsw_b200.h"
extern "C" {
#include "cphnsw_oracle.h"
}

static uint64_t g_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd() { g_state = g_state * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(g_state >> 33); }
static float rndf() {   // roughly normal: sum of four uniforms
    float s = 0.0f;
    for (int i = 0; i < 4; ++i) s += (float)(rnd() & 0xFFFFFF) / 16777216.0f;
    return (s - 2.0f) * 1.7320508f;
}

static int g_tile = 0;
static int run_case(cphnsw_b200_index* ix, uint32_t dim, uint32_t bits, uint32_t n_parents, uint32_t verify = 0xFFFFFFFFu) {
    if (verify > n_parents) verify = n_parents;
    uint32_t D = 16;
    while (D < dim) D <<= 1;
    const uint32_t n = n_parents * 8 + 40;
    std::vector<float> vec((size_t)n * dim);
    for (size_t i = 0; i < vec.size(); ++i) vec[i] = rndf();
    for (uint32_t r = n / 2; r < n; ++r)   // near-duplicates of the first half: short offsets
        for (uint32_t i = 0; i < dim; ++i) vec[(size_t)r * dim + i] = vec[(size_t)(r - n / 2) * dim + i] + 0.05f * rndf();
    std::vector<uint32_t> pids(n_parents), nbr((size_t)n_parents * 32);
    for (auto& p : pids) p = rnd() % n;
    for (auto& v : nbr) v = (rnd() % 7 == 0) ? 0xFFFFFFFFu : rnd() % n;
    nbr[3] = pids[0];   // nop == 0
    const size_t cb = (size_t)bits * (D / 8);
    std::vector<uint8_t> codes((size_t)n_parents * 32 * cb, 0xAA), want(codes.size(), 0);
    std::vector<float> aux((size_t)n_parents * 32 * 3, -1.0f), want_aux(aux.size(), 0.0f);

    const uint32_t nop_off = 4 * D * bits, pop_off = nop_off + 384, ids_off = pop_off + 64 * (bits > 1 ? 2 : 1);
    const uint32_t bsize = (ids_off + 128 + 4 + 63) / 64 * 64;
    std::vector<uint8_t> blocks((size_t)n_parents * bsize, 0xCC), want_blocks(blocks.size(), 0xCC);
    uint8_t* d_blocks;
    cudaMalloc(&d_blocks, blocks.size());
    cudaMemset(d_blocks, 0xCC, blocks.size());
    float *d_vec; uint32_t *d_pid, *d_nbr; uint8_t* d_codes; float* d_aux;
    cudaMalloc(&d_vec, vec.size() * 4); cudaMalloc(&d_pid, pids.size() * 4); cudaMalloc(&d_nbr, nbr.size() * 4);
    cudaMalloc(&d_codes, codes.size()); cudaMalloc(&d_aux, aux.size() * 4);
    cudaMemcpy(d_vec, vec.data(), vec.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_pid, pids.data(), pids.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_nbr, nbr.data(), nbr.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(d_codes, 0xAA, codes.size());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int rc = 0;
    for (int rep = 0; rep < 2 && rc == 0; ++rep) {   // the second launch is the timed one (the first may allocate scratch)
    cudaEventRecord(e0);
    rc = cphnsw_b200_neighbor_codes(ix, dim, bits, 42, d_vec, dim, n, d_pid, d_nbr, n_parents, d_codes, d_aux, d_blocks, bsize, nullptr);
    cudaEventRecord(e1);
    }
    cudaError_t ce = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    if (rc != 0 || ce != cudaSuccess) {
        std::printf("dim=%u bits=%u: call failed rc=%d (%s) cuda=%s\n", dim, bits, rc, cphnsw_b200_last_error(ix), cudaGetErrorString(ce));
        return 1;
    }
    cudaMemcpy(codes.data(), d_codes, codes.size(), cudaMemcpyDeviceToHost);
    cudaMemcpy(aux.data(), d_aux, aux.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(blocks.data(), d_blocks, blocks.size(), cudaMemcpyDeviceToHost);
    cudaFree(d_blocks);
    cudaFree(d_vec); cudaFree(d_pid); cudaFree(d_nbr); cudaFree(d_codes); cudaFree(d_aux);

    std::vector<float> signs((size_t)3 * D), par(D), nb(D);
    cpo_rotation_signs(D, 42, signs.data());
    for (uint32_t p = 0; p < verify; ++p) {
        std::fill(par.begin(), par.end(), 0.0f);
        std::memcpy(par.data(), &vec[(size_t)pids[p] * dim], dim * 4);
        for (uint32_t v = 0; v < 32; ++v) {
            const uint32_t id = nbr[(size_t)p * 32 + v];
            if (id >= n) continue;
            std::fill(nb.begin(), nb.end(), 0.0f);
            std::memcpy(nb.data(), &vec[(size_t)id * dim], dim * 4);
            uint8_t* c = &want[((size_t)p * 32 + v) * cb];
            float* a = &want_aux[((size_t)p * 32 + v) * 3];
            if (bits == 1) cpo_neighbor_aux_1bit(dim, D, signs.data(), par.data(), nb.data(), 0, c, a);
            else cpo_neighbor_aux_nbit(dim, D, bits, signs.data(), par.data(), nb.data(), CPO_CAQ_FLAGS, c, a);
        }
    }
    // the same codes in the reference's block layout (fastscan_layout.hpp:51-92, 114-155)
    for (uint32_t p = 0; p < verify; ++p) {
        uint8_t* b = &want_blocks[(size_t)p * bsize];
        uint32_t count = 0;
        for (uint32_t v = 0; v < 32; ++v) {
            const uint32_t id = nbr[(size_t)p * 32 + v];
            const uint8_t* c = &want[((size_t)p * 32 + v) * cb];
            uint32_t pop = 0, wpop = 0;
            for (uint32_t pl = 0; pl < bits; ++pl)
                for (uint32_t j = 0; j < D / 8; ++j) {
                    const uint8_t byte = c[(size_t)pl * (D / 8) + j];
                    b[(size_t)pl * 4 * D + (size_t)j * 32 + v] = byte;
                    const uint32_t pc = (uint32_t)__builtin_popcount(byte);
                    if (pl == 0) pop += pc;
                    wpop += pc << (bits - 1 - pl);
                }
            for (int k = 0; k < 3; ++k) std::memcpy(b + nop_off + 128 * k + 4 * v, &want_aux[((size_t)p * 32 + v) * 3 + k], 4);
            const uint16_t p16 = (uint16_t)pop, w16 = (uint16_t)wpop;
            std::memcpy(b + pop_off + 2 * v, &p16, 2);
            if (bits > 1) std::memcpy(b + pop_off + 64 + 2 * v, &w16, 2);
            const uint32_t idw = id < n ? id : 0xFFFFFFFFu;
            std::memcpy(b + ids_off + 4 * v, &idw, 4);
            if (id < n) count = v + 1;
        }
        std::memcpy(b + ids_off + 128, &count, 4);
    }
    size_t bad_blocks = 0;
    for (size_t i = 0; i < (size_t)verify * bsize; ++i) bad_blocks += blocks[i] != want_blocks[i];
    size_t bad_codes = 0, bad_aux = 0;
    for (size_t i = 0; i < (size_t)verify * 32 * cb; ++i) bad_codes += codes[i] != want[i];
    for (size_t i = 0; i < (size_t)verify * 96; ++i) bad_aux += std::memcmp(&aux[i], &want_aux[i], 4) != 0;
    std::printf("tile=%d dim=%u D=%u bits=%u parents=%u: %zu code bytes, %zu aux words, %zu block bytes differ; %.3f ms (%.1f ns/pair)\n", g_tile, dim, D,
                bits, n_parents, bad_codes, bad_aux, bad_blocks, ms, ms * 1e6 / (n_parents * 32.0));
    return (bad_codes || bad_aux || bad_blocks) ? 1 : 0;
}

int main() {
    setvbuf(stdout, nullptr, _IOLBF, 0);
    cphnsw_b200_index* ix = nullptr;
    if (cphnsw_b200_create(0, &ix) != 0) { std::printf("create failed: %s\n", cphnsw_b200_last_error(nullptr)); return 2; }
    int bad = 0;
    // tile = 1: per-warp tiles in shared memory; 2: in global memory (what large D defaults to); results identical
    const uint32_t dims[] = {128, 96, 20, 10, 300, 960, 1500};
    for (g_tile = 2; g_tile >= 1; --g_tile) {
        cphnsw_b200_set_option(ix, "neighbor_codes_tile", g_tile);
        for (uint32_t dim : dims)
            for (uint32_t bits : {1u, 2u, 4u}) bad += run_case(ix, dim, bits, dim > 256 ? 20 : 60);
    }
    // timing samples (the first parents of each are checked)
    for (g_tile = 1; g_tile <= 2; ++g_tile) {
        cphnsw_b200_set_option(ix, "neighbor_codes_tile", g_tile);
        for (uint32_t bits : {1u, 2u, 4u}) bad += run_case(ix, 128, bits, 30000, 200);
        for (uint32_t bits : {2u, 4u}) bad += run_case(ix, 200, bits, 10000, 50);
        bad += run_case(ix, 300, 2, 6000, 30);
        for (uint32_t bits : {1u, 2u, 4u}) bad += run_case(ix, 960, bits, 3000, 20);
    }
    cphnsw_b200_destroy(ix);
    std::printf(bad ? "FAILED\n" : "ALL OK\n");
    return bad ? 1 : 0;
}
