// Microbenchmark (not product code): throughput of scattered integer RED/ATOM on B200, to size the
// K[i][j] += c_i*c_j accumulate.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o atomic_bench atomic_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

// mode 0: every lane a random element of the whole footprint
// mode 1: every warp picks a random "row" (row_len elements), its lanes random columns inside it
template <typename T>
__global__ void red_kernel(T* buf, uint64_t n_elems, uint64_t row_len, int iters, int mode, uint64_t seed) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t warp = tid >> 5;
    for (int it = 0; it < iters; ++it) {
        uint64_t idx;
        if (mode == 0) {
            idx = mix(tid * 0x9e3779b97f4a7c15ULL + it + seed) % n_elems;
        } else {
            const uint64_t nrows = n_elems / row_len;
            const uint64_t row = mix(warp * 0x9e3779b97f4a7c15ULL + (it >> 2) + seed) % nrows;
            idx = row * row_len + mix(tid * 0xda942042e4dd58b5ULL + it + seed) % row_len;
        }
        atomicAdd(buf + idx, (T)1);
    }
}

__global__ void smem_atomic_kernel(unsigned* out, int iters, uint64_t seed) {
    extern __shared__ unsigned sm[];
    for (int i = threadIdx.x; i < 48 * 1024; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int it = 0; it < iters; ++it) atomicAdd(&sm[mix(tid * 0x9e3779b97f4a7c15ULL + it + seed) % (48 * 1024)], 1u);
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = sm[0];
}

// plain (non-atomic) shared read-modify-write on random addresses: upper bound for an exclusive-owner scheme
__global__ void smem_rmw_kernel(unsigned* out, int iters, uint64_t seed) {
    extern __shared__ unsigned sm[];
    for (int i = threadIdx.x; i < 48 * 1024; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int it = 0; it < iters; ++it) { unsigned a = mix(tid * 0x9e3779b97f4a7c15ULL + it + seed) % (48 * 1024); sm[a] = sm[a] + 1; }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = sm[0];
}

template <typename T>
double run(T* buf, uint64_t n_elems, uint64_t row_len, int mode, const char* label) {
    const int blocks = 148 * 16, threads = 256, iters = 256;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    red_kernel<T><<<blocks, threads>>>(buf, n_elems, row_len, 16, mode, 1);   // warm-up
    cudaEventRecord(a);
    red_kernel<T><<<blocks, threads>>>(buf, n_elems, row_len, iters, mode, 7);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double ups = (double)blocks * threads * iters / (ms * 1e-3);
    printf("%-58s %8.3f ms  %8.2f Gupd/s\n", label, ms, ups * 1e-9);
    return ups;
}

int main() {
    void* buf;
    const size_t bytes = 10ull << 30;
    if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(buf, 0, bytes);
    char label[128];
    const size_t foot[] = {16ull << 20, 32ull << 20, 48ull << 20, 64ull << 20, 96ull << 20, 128ull << 20, 256ull << 20, 1ull << 30, 5ull << 30, 10ull << 30};
    for (size_t f : foot) {
        snprintf(label, sizeof label, "RED.u64 random, footprint %6zu MB", f >> 20);
        run<unsigned long long>((unsigned long long*)buf, f / 8, 1, 0, label);
        snprintf(label, sizeof label, "RED.u32 random, footprint %6zu MB", f >> 20);
        run<unsigned>((unsigned*)buf, f / 4, 1, 0, label);
    }
    // warp-in-a-row pattern (row = 25k..50k elements like a row of the packed triangle)
    for (size_t f : {64ull << 20, 5ull << 30, 10ull << 30}) {
        snprintf(label, sizeof label, "RED.u64 warp-per-row (32768-elem rows), footprint %6zu MB", f >> 20);
        run<unsigned long long>((unsigned long long*)buf, f / 8, 32768, 1, label);
        snprintf(label, sizeof label, "RED.u32 warp-per-row (32768-elem rows), footprint %6zu MB", f >> 20);
        run<unsigned>((unsigned*)buf, f / 4, 32768, 1, label);
    }
    // shared-memory atomics, 192 KB per CTA, 1 CTA per SM
    {
        unsigned* out; cudaMalloc(&out, 4096 * 4);
        cudaFuncSetAttribute(smem_atomic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 1024);
        cudaFuncSetAttribute(smem_rmw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 1024);
        for (int threads : {256, 512, 1024}) {
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            const int iters = 4096;
            smem_atomic_kernel<<<148, threads, 192 * 1024>>>(out, 64, 1);
            cudaEventRecord(a);
            smem_atomic_kernel<<<148, threads, 192 * 1024>>>(out, iters, 3);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            printf("ATOMS.u32 random in 192 KB smem, 148 CTAs x %4d thr     %8.3f ms  %8.2f Gupd/s\n", threads, ms, 148.0 * threads * iters / (ms * 1e6));
            cudaEventRecord(a);
            smem_rmw_kernel<<<148, threads, 192 * 1024>>>(out, iters, 3);
            cudaEventRecord(b); cudaEventSynchronize(b);
            cudaEventElapsedTime(&ms, a, b);
            printf("LDS+STS  u32 random in 192 KB smem, 148 CTAs x %4d thr    %8.3f ms  %8.2f Gupd/s\n", threads, ms, 148.0 * threads * iters / (ms * 1e6));
        }
    }
    printf("last error: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
