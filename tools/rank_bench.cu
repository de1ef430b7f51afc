// Microbenchmark: cost of the stable in-warp digit ranking of one onesweep tile (256 threads x 16 keys, 8-bit digit)
// for three ways of finding the lanes that hold the same digit:
//   0 = match.any (hardware MATCH, ADU pipe)   1 = 8 ballots   2 = shared-memory atomicOr match masks
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/rank_bench tools/rank_bench.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

constexpr int ITEMS = 16, THREADS = 256, RADIX = 256;

template <int MODE>
__global__ void __launch_bounds__(THREADS) rank_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t n, int shift) {
    __shared__ uint32_t warp_hist[8 * RADIX];
    __shared__ uint32_t mm[8 * RADIX];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 8 * RADIX; i += THREADS) { warp_hist[i] = 0; mm[i] = 0; }
    __syncthreads();
    const uint32_t tile0 = blockIdx.x * (THREADS * ITEMS);
    uint32_t key[ITEMS], rank[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const uint32_t idx = tile0 + warp * (32 * ITEMS) + j * 32 + lane;
        key[j] = idx < n ? in[idx] : 0;
    }
    const uint32_t lt = (1u << lane) - 1;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const uint32_t d = (key[j] >> shift) & 255u;
        uint32_t peers;
        if (MODE == 0) {
            peers = __match_any_sync(0xffffffffu, d);
        } else if (MODE == 1) {
            peers = 0xffffffffu;
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const bool p = (d >> b) & 1u;
                const uint32_t bal = __ballot_sync(0xffffffffu, p);
                peers &= p ? bal : ~bal;
            }
        } else if (MODE == 3) {
            // optimistic: one atomic with return; stable only if the hardware serialises same-address lanes in lane order
            rank[j] = atomicAdd(&warp_hist[warp * RADIX + d], 1u);
            continue;
        } else {
            atomicOr(&mm[warp * RADIX + d], 1u << lane);
            __syncwarp();
            peers = mm[warp * RADIX + d];
            __syncwarp();
        }
        const uint32_t lower = peers & lt;
        const uint32_t prev = warp_hist[warp * RADIX + d];
        __syncwarp();
        if (lower == 0) {
            warp_hist[warp * RADIX + d] = prev + __popc(peers);
            if (MODE == 2) mm[warp * RADIX + d] = 0;
        }
        __syncwarp();
        rank[j] = prev + __popc(lower);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const uint32_t idx = tile0 + warp * (32 * ITEMS) + j * 32 + lane;
        if (idx < n) out[idx] = rank[j] + warp_hist[warp * RADIX + ((key[j] >> shift) & 255u)];
    }
}

int main() {
    const uint32_t n = 9250000u * 8;
    uint32_t *in, *out;
    cudaMalloc(&in, n * 4ull);
    cudaMalloc(&out, n * 4ull);
    uint32_t* h = (uint32_t*)malloc(n * 4ull);
    uint64_t s = 88172645463325252ull;
    for (uint32_t i = 0; i < n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (uint32_t)(s >> 20); }
    cudaMemcpy(in, h, n * 4ull, cudaMemcpyHostToDevice);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int grid = (n + THREADS * ITEMS - 1) / (THREADS * ITEMS);
    uint32_t* ref = (uint32_t*)malloc(n * 4ull);
    for (int mode = 0; mode < 4; ++mode) {
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(a);
            if (mode == 0) rank_kernel<0><<<grid, THREADS>>>(in, out, n, 8);
            if (mode == 1) rank_kernel<1><<<grid, THREADS>>>(in, out, n, 8);
            if (mode == 2) rank_kernel<2><<<grid, THREADS>>>(in, out, n, 8);
            if (mode == 3) rank_kernel<3><<<grid, THREADS>>>(in, out, n, 8);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (ms < best) best = ms;
        }
        cudaMemcpy(h, out, n * 4ull, cudaMemcpyDeviceToHost);
        if (mode == 0) memcpy(ref, h, n * 4ull);
        size_t bad = 0;
        for (uint32_t i = 0; i < n; ++i) bad += h[i] != ref[i];
        printf("mode %d (%s): %.3f ms for %u keys = %.1f Gkeys/s, %.1f GB/s read+write, mismatches vs match.any: %zu  [%s]\n", mode,
               mode == 0 ? "match.any" : mode == 1 ? "8 ballots" : mode == 2 ? "atomicOr masks" : "atomicAdd return (lane order assumed)", best, n, n / best * 1e-6, n * 8.0 / best * 1e-6, bad,
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
