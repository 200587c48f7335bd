// kmu_count_ops.cuh -- slot operations of the exact counting table (shared by kmu_count.cu and kmu_count_part.cu).
// Layout: see kmu_count.cu / DESIGN.md "counting table".
#pragma once
#include <cstdint>

#include "kmu_device.cuh"
#include "kmu_kernels.h"

namespace kmu {

__device__ __forceinline__ uint64_t fmix64(uint64_t h) {
    h ^= h >> 33;
    h *= 0xff51afd7ed558ccdULL;
    h ^= h >> 33;
    h *= 0xc4ceb9fe1a85ec53ULL;
    h ^= h >> 33;
    return h;
}

template <typename V>
struct CountOps;

template <>
struct CountOps<uint32_t> {
    // The home slot is claimed without looking first (one L2 round trip instead of two when it is free -- the common
    // case of a genome's mostly distinct k-mers); home() returns what the slot held (0: claimed, done) and rest()
    // finishes from there, looking before it claims.  (Four claims in flight per lane were slower: 10.4 -> 16.6 ms for 32 genomes.)
    static __device__ __forceinline__ unsigned long long home(const CountTable& t, uint32_t key, uint32_t add, uint64_t& i) {
        i = fmix64(key) & t.capmask;
        return atomicCAS((unsigned long long*)t.slots + i, 0ULL, ((unsigned long long)key << 32) | add);
    }
    static __device__ __forceinline__ bool rest(const CountTable& t, uint32_t key, uint32_t add, uint64_t i, unsigned long long e) {
        unsigned long long* tab = (unsigned long long*)t.slots;
        for (uint64_t probe = 0; probe <= t.capmask; ++probe) {
            if (probe) e = *(volatile unsigned long long*)(tab + i);
            if (e == 0) {
                unsigned long long old = atomicCAS(tab + i, 0ULL, ((unsigned long long)key << 32) | add);
                if (old == 0) return true;
                e = old;
            }
            if ((uint32_t)(e >> 32) == key) {
                unsigned long long old = atomicAdd(tab + i, (unsigned long long)add);
                if ((uint32_t)old + (uint64_t)add >= 0xFFFFFFF0ull) atomicAdd(tab + i, (unsigned long long)(-(long long)add));  // saturate
                return true;
            }
            i = (i + 1) & t.capmask;
        }
        return false;
    }
    static __device__ __forceinline__ bool insert(const CountTable& t, uint32_t key, uint32_t add) {
        uint64_t i;
        const unsigned long long e = home(t, key, add, i);
        return e == 0 ? true : rest(t, key, add, i, e);
    }
    static __device__ __forceinline__ uint32_t lookup(const CountTable& t, uint32_t key) {
        const unsigned long long* tab = (const unsigned long long*)t.slots;
        uint64_t i = fmix64(key) & t.capmask;
        for (uint64_t probe = 0; probe <= t.capmask; ++probe) {
            unsigned long long e = __ldg(tab + i);
            if (e == 0) return 0;
            if ((uint32_t)(e >> 32) == key) return (uint32_t)e;
            i = (i + 1) & t.capmask;
        }
        return 0;
    }
    static __device__ __forceinline__ bool occupied(const CountTable& t, uint64_t i, uint64_t& key, uint64_t& cnt) {
        unsigned long long e = ((const unsigned long long*)t.slots)[i];
        key = e >> 32;
        cnt = (uint32_t)e;
        return e != 0;
    }
};

template <>
struct CountOps<uint64_t> {
    static constexpr unsigned long long EMPTY = ~0ULL;
    static __device__ __forceinline__ bool insert(const CountTable& t, uint64_t key, uint32_t add) {
        if (key == EMPTY) {
            atomicAdd(t.special, (unsigned long long)add);
            return true;
        }
        unsigned long long* tab = (unsigned long long*)t.slots;
        uint64_t i = fmix64(key) & t.capmask;
        for (uint64_t probe = 0; probe <= t.capmask; ++probe) {
            // (looking first is the faster order here: claiming the home slot blind, as the u32 table does, made the
            // insertion of 960 M 31-mers 8 % slower -- most k-mers of a read set are repeats and the CAS is wasted)
            unsigned long long cur = *(volatile unsigned long long*)(tab + 2 * i);
            if (cur == EMPTY) {
                cur = atomicCAS(tab + 2 * i, EMPTY, (unsigned long long)key);
                if (cur == EMPTY) cur = key;
            }
            if (cur == key) {
                atomicAdd(tab + 2 * i + 1, (unsigned long long)add);
                return true;
            }
            i = (i + 1) & t.capmask;
        }
        return false;
    }
    static __device__ __forceinline__ uint32_t lookup(const CountTable& t, uint64_t key) {
        if (key == EMPTY) {
            unsigned long long c = *t.special;
            return c > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)c;
        }
        const unsigned long long* tab = (const unsigned long long*)t.slots;
        uint64_t i = fmix64(key) & t.capmask;
        for (uint64_t probe = 0; probe <= t.capmask; ++probe) {
            const ulonglong2 e = __ldg((const ulonglong2*)(tab + 2 * i));
            if (e.x == EMPTY) return 0;
            if (e.x == key) return e.y > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)e.y;
            i = (i + 1) & t.capmask;
        }
        return 0;
    }
    static __device__ __forceinline__ bool occupied(const CountTable& t, uint64_t i, uint64_t& key, uint64_t& cnt) {
        const ulonglong2 e = ((const ulonglong2*)t.slots)[i];
        key = e.x;
        cnt = e.y;
        return e.x != EMPTY;
    }
};

}  // namespace kmu
