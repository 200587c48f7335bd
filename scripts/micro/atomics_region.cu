// random slot updates over a 16 GB table, visited region by region (32 MB regions), like the two-phase counter
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t h) { h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33; return h; }
__global__ void k(unsigned long long* tab, uint64_t n, uint64_t per_region, uint32_t region_bits, uint64_t distinct_per_region) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = i / per_region;
        const uint64_t key = mix(i % distinct_per_region + r * 1000003ull);  // each distinct key repeats per_region / distinct times
        const uint64_t s = (r << region_bits) | (mix(key) & ((1ull << region_bits) - 1));
        unsigned long long cur = *(volatile unsigned long long*)(tab + 2 * s);
        if (cur != 12345) atomicAdd(tab + 2 * s + 1, 1ULL);
    }
}
int main() {
    const uint32_t region_bits = 21;  // 2M slots of 16 B = 32 MB
    const uint64_t regions = 512, slots = regions << region_bits;
    unsigned long long* tab;
    cudaMalloc(&tab, slots * 16);
    cudaMemset(tab, 0, slots * 16);
    const uint64_t per_region = 1875000, n = per_region * regions;
    for (uint64_t distinct : {1875000ull, 468750ull}) {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        k<<<148 * 8, 256>>>(tab, n, per_region, region_bits, distinct);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("region-ordered updates, %llu distinct per region: %.2f ms  %.1f G updates/s\n", (unsigned long long)distinct, ms, n / ms / 1e6);
    }
    return 0;
}
