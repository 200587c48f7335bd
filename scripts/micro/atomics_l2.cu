// micro-benchmark: random 16-byte-slot updates (volatile load + RED add) over regions of different sizes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t h) { h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33; return h; }
template <int MODE>
__global__ void k(unsigned long long* tab, uint64_t mask, uint64_t n, uint64_t seed) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t s = mix(i + seed) & mask;
        if (MODE == 0) {  // volatile load of the key + RED on the count
            unsigned long long cur = *(volatile unsigned long long*)(tab + 2 * s);
            if (cur != 12345) atomicAdd(tab + 2 * s + 1, 1ULL);
        } else if (MODE == 1) {  // RED only
            atomicAdd(tab + 2 * s + 1, 1ULL);
        } else if (MODE == 2) {  // plain (non-volatile) load via ld.global.cg + RED
            unsigned long long cur = __ldcg(tab + 2 * s);
            if (cur != 12345) atomicAdd(tab + 2 * s + 1, 1ULL);
        } else if (MODE == 3) {  // load only
            unsigned long long cur = __ldcg(tab + 2 * s);
            if (cur == 12345) tab[0] = 1;
        } else if (MODE == 4) {  // load of one slot, RED on an UNRELATED slot of the region (is the cost of mode 0 tied to the same sector?)
            unsigned long long cur = __ldcg(tab + 2 * s);
            const uint64_t s2 = mix(i + seed + 0x9E3779B97F4A7C15ULL) & mask;
            if (cur != 12345) atomicAdd(tab + 2 * s2 + 1, 1ULL);
        } else if (MODE == 5) {  // RED first (fire and forget), then the load of the same slot, nothing depends on it until the end
            atomicAdd(tab + 2 * s + 1, 1ULL);
            unsigned long long cur = __ldcg(tab + 2 * s);
            if (cur == 12345) tab[0] = 1;
        } else if (MODE == 6) {  // one 128-bit load of the whole slot (key and count), then the RED
            const ulonglong2 e = __ldcg((const ulonglong2*)(tab + 2 * s));
            if (e.x != 12345) atomicAdd(tab + 2 * s + 1, 1ULL);
        } else if (MODE == 7) {  // atomicAdd WITH return on the count word only (one round trip, no load)
            unsigned long long old = atomicAdd(tab + 2 * s + 1, 1ULL);
            if (old == 0xFFFFFFFFFFFFull) tab[0] = 1;
        } else if (MODE == 8) {  // the key read by an ATOMIC (compare-and-swap that changes nothing), then the RED: atomics only
            unsigned long long cur = atomicCAS(tab + 2 * s, 12345ULL, 12345ULL);
            if (cur != 12345) atomicAdd(tab + 2 * s + 1, 1ULL);
        } else if (MODE == 9) {  // as 8 with atomicOr(key, 0)
            unsigned long long cur = atomicOr(tab + 2 * s, 0ULL);
            if (cur != 12345) atomicAdd(tab + 2 * s + 1, 1ULL);
        } else if (MODE == 11) {  // even CTAs only load, odd CTAs only RED (n / 2 of each: n / 2 "updates"): is the mix penalised in L2 or in the SM?
            if (blockIdx.x & 1) {
                atomicAdd(tab + 2 * s + 1, 1ULL);
            } else {
                unsigned long long cur = __ldcg(tab + 2 * s);
                if (cur == 12345) tab[0] = 1;
            }
        } else if (MODE == 12) {  // inside every CTA: warps 0-3 only load, warps 4-7 only RED
            if (threadIdx.x & 128) {
                atomicAdd(tab + 2 * s + 1, 1ULL);
            } else {
                unsigned long long cur = __ldcg(tab + 2 * s);
                if (cur == 12345) tab[0] = 1;
            }
        } else {  // MODE 10: one 128-bit compare-and-swap on the whole slot (key, count) with a guessed count of 0
            unsigned long long o0, o1;
            asm volatile("{\n\t.reg .b128 c, n, o;\n\tmov.b128 c, {%3, %4};\n\tmov.b128 n, {%5, %6};\n\tatom.cas.b128 o, [%2], c, n;\n\tmov.b128 {%0, %1}, o;\n\t}"
                         : "=l"(o0), "=l"(o1) : "l"(tab + 2 * s), "l"(0ULL), "l"(0ULL), "l"(0ULL), "l"(1ULL) : "memory");
            if (o0 == 12345) tab[0] = 1;
        }
    }
}
int main() {
    const uint64_t n = 400000000ull;
    for (uint64_t mb : {64ull}) {
        uint64_t slots = mb * 1024 * 1024 / 16;
        unsigned long long* tab;
        cudaMalloc(&tab, slots * 16);
        cudaMemset(tab, 0, slots * 16);
        for (int mode = 0; mode < 13; ++mode) {
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            cudaEventRecord(a);
            if (mode == 0) k<0><<<148 * 8, 256>>>(tab, slots - 1, n, 7);
            if (mode == 1) k<1><<<148 * 8, 256>>>(tab, slots - 1, n, 7);
            if (mode == 2) k<2><<<148 * 8, 256>>>(tab, slots - 1, n, 7);
            if (mode == 3) k<3><<<148 * 8, 256>>>(tab, slots - 1, n, 7);
            if (mode == 4) k<4><<<148 * 8, 256>>>(tab, slots - 1, n, 7);
            if (mode == 5) k<5><<<148 * 8, 256>>>(tab, slots - 1, n, 7);
            if (mode == 6) k<6><<<148 * 8, 256>>>(tab, slots - 1, n, 7);
            if (mode == 7) k<7><<<148 * 8, 256>>>(tab, slots - 1, n, 7);
            if (mode == 8) k<8><<<148 * 8, 256>>>(tab, slots - 1, n, 7);
            if (mode == 9) k<9><<<148 * 8, 256>>>(tab, slots - 1, n, 7);
            if (mode == 10) k<10><<<148 * 8, 256>>>(tab, slots - 1, n, 7);
            if (mode == 11) k<11><<<148 * 8, 256>>>(tab, slots - 1, n, 7);
            if (mode == 12) k<12><<<148 * 8, 256>>>(tab, slots - 1, n, 7);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            printf("region %5llu MB mode %d : %.2f ms  %.1f G updates/s\n", (unsigned long long)mb, mode, ms, (mode >= 11 ? n / 2 : n) / ms / 1e6);
        }
        cudaFree(tab);
    }
    return 0;
}
