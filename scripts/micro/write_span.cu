// micro-benchmark: write-only streams where every warp owns a contiguous span of the output and writes it 512 bytes per
// store instruction (the pattern of generate_kmers_*): bandwidth against the span per warp, persistent grid-stride warps.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) k(uint4* out, uint64_t total16, uint64_t span16) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t nspans = total16 / span16;
    for (uint64_t s = warp; s < nspans; s += nwarps) {
        uint4* o = out + s * span16;
        for (uint64_t i = lane; i < span16; i += 32) __stcs(o + i, make_uint4((uint32_t)i, (uint32_t)s, 3, 4));
    }
}
int main() {
    const uint64_t bytes = 8ull << 30, total16 = bytes / 16;
    uint4* out;
    cudaMalloc(&out, bytes);
    for (int ctas_per_sm : {4, 8}) {
        for (uint64_t span_kb : {1ull, 4ull, 8ull, 16ull, 64ull, 256ull}) {
            const uint64_t span16 = span_kb * 1024 / 16;
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            k<<<148 * ctas_per_sm, 256>>>(out, total16, span16);
            cudaEventRecord(a);
            k<<<148 * ctas_per_sm, 256>>>(out, total16, span16);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            printf("ctas/SM %d  span %4llu KB per warp : %.2f ms  %.0f GB/s\n", ctas_per_sm, (unsigned long long)span_kb, ms, bytes / ms / 1e6);
        }
    }
    return 0;
}
