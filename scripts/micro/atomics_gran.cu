// micro-benchmark: does cudaLimitMaxL2FetchGranularity change the cost of random 16-byte-slot updates on a table far
// larger than L2?  argv[1] = granularity in bytes (32 / 64 / 128), 0 = leave the default.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t h) { h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33; return h; }
template <int MODE>
__global__ void k(unsigned long long* tab, uint64_t mask, uint64_t n, uint64_t seed) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t s = mix(i + seed) & mask;
        if (MODE == 0) {  // volatile load of the key + RED on the count (what count_insert does for a known key)
            unsigned long long cur = *(volatile unsigned long long*)(tab + 2 * s);
            if (cur != 12345) atomicAdd(tab + 2 * s + 1, 1ULL);
        } else if (MODE == 1) {  // RED only
            atomicAdd(tab + 2 * s + 1, 1ULL);
        } else {  // load only
            unsigned long long cur = __ldcg(tab + 2 * s);
            if (cur == 12345) tab[0] = 1;
        }
    }
}
int main(int argc, char** argv) {
    const int gran = argc > 1 ? atoi(argv[1]) : 0;
    if (gran) {
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)gran);
        printf("set granularity %d: %s\n", gran, cudaGetErrorString(e));
    }
    size_t g = 0;
    cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
    printf("cudaLimitMaxL2FetchGranularity = %zu\n", g);
    const uint64_t n = 400000000ull;
    for (uint64_t mb : {4096ull, 16384ull}) {
        uint64_t slots = mb * 1024 * 1024 / 16;
        unsigned long long* tab;
        if (cudaMalloc(&tab, slots * 16) != cudaSuccess) return 1;
        cudaMemset(tab, 0, slots * 16);
        for (int mode = 0; mode < 3; ++mode) {
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            cudaEventRecord(a);
            if (mode == 0) k<0><<<148 * 8, 256>>>(tab, slots - 1, n, 7);
            if (mode == 1) k<1><<<148 * 8, 256>>>(tab, slots - 1, n, 7);
            if (mode == 2) k<2><<<148 * 8, 256>>>(tab, slots - 1, n, 7);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            printf("gran %3zu region %5llu MB mode %d : %.2f ms  %.1f G updates/s\n", g, (unsigned long long)mb, mode, ms, n / ms / 1e6);
        }
        cudaFree(tab);
    }
    return 0;
}
