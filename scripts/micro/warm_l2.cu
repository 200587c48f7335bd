// micro-benchmark: does bringing a table region into L2 ahead of time make the random slot updates (load + RED) on it faster?
// For regions of 32 / 64 / 128 MB of a 4 GB table: per region 0.75 updates per slot at random; the region is (a) cold, (b) warmed by
// prefetch.global.L2 of every line, (c) by streaming loads (ld.global.cg), (d) by loads with the L2::evict_last hint.  Times are
// the update kernels alone (the warm kernels are timed apart), averaged over the regions of the table.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t h) { h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33; return h; }
template <int OP>
__global__ void update(unsigned long long* tab, uint64_t mask, uint64_t n, uint64_t seed) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t s = mix(i + seed) & mask;
        if (OP == 0) {  // look at the key, then a fire-and-forget add on the count
            const unsigned long long cur = *(volatile unsigned long long*)(tab + 2 * s);
            if (cur != 12345) atomicAdd(tab + 2 * s + 1, 1ULL);
        } else if (OP == 1) {  // one returning add (a tagged count word)
            const unsigned long long old = atomicAdd(tab + 2 * s + 1, 1ULL);
            if (old == 0xFFFFFFFFFFFFull) tab[0] = 1;
        } else {  // fire-and-forget add only
            atomicAdd(tab + 2 * s + 1, 1ULL);
        }
    }
}
template <int HOW>
__global__ void warm(const unsigned long long* tab, uint64_t lines, unsigned long long* sink) {
    unsigned long long acc = 0;
    for (uint64_t l = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; l < lines * 4; l += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long* p = tab + l * 4;  // one 32-byte sector per thread
        if (HOW == 1) {
            if ((l & 3) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
        } else if (HOW == 2) {
            acc += __ldcg(p);
        } else {
            unsigned long long v, pol;
            asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
            asm volatile("ld.global.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
            acc += v;
        }
    }
    if (acc == 0x123456789ull) *sink = acc;
}
int main() {
    const uint64_t table_bytes = 4ull << 30;
    unsigned long long *tab, *sink;
    cudaMalloc(&tab, table_bytes);
    cudaMalloc(&sink, 8);
    cudaMemset(tab, 0, table_bytes);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (uint64_t mb : {32ull, 64ull, 128ull}) {
        const uint64_t rbytes = mb << 20, slots = rbytes / 16, nreg = table_bytes / rbytes, n = slots * 3 / 4;
        for (int op = 0; op < 3; ++op)
        for (int how = 0; how < 3; ++how) {
            float t_upd = 0, t_warm = 0;
            for (uint64_t r = 0; r < nreg; ++r) {
                unsigned long long* reg = tab + r * slots * 2;
                float ms;
                if (how) {
                    cudaEventRecord(a);
                    if (how == 1) warm<1><<<148 * 8, 256>>>(reg, rbytes / 128, sink);
                    if (how == 2) warm<2><<<148 * 8, 256>>>(reg, rbytes / 128, sink);
                    if (how == 3) warm<3><<<148 * 8, 256>>>(reg, rbytes / 128, sink);
                    cudaEventRecord(b);
                    cudaEventSynchronize(b);
                    cudaEventElapsedTime(&ms, a, b);
                    t_warm += ms;
                }
                cudaEventRecord(a);
                if (op == 0) update<0><<<148 * 8, 256>>>(reg, slots - 1, n, 7 + r);
                if (op == 1) update<1><<<148 * 8, 256>>>(reg, slots - 1, n, 7 + r);
                if (op == 2) update<2><<<148 * 8, 256>>>(reg, slots - 1, n, 7 + r);
                cudaEventRecord(b);
                cudaEventSynchronize(b);
                cudaEventElapsedTime(&ms, a, b);
                t_upd += ms;
            }
            const double total_updates = (double)n * nreg;
            printf("region %3llu MB op %d warm %d : updates %.2f ms (%.1f G/s), warm kernels %.2f ms (%.0f GB/s)\n", (unsigned long long)mb, op, how,
                   t_upd, total_updates / t_upd / 1e6, t_warm, how ? table_bytes / t_warm / 1e6 : 0.0);
        }
    }
    return 0;
}
