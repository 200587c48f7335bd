// kmu_pmh3a_items.cu -- ProbMinHash3a over an explicit weighted set, on the whole GPU.
//
// Replaces ProbMinHash3a::hash_weigthed_hashmap as called with ONE multiplicity map for a whole file:
// ProbHash3aSketch::sketch_compressedkmer_seqs (src/sketching/setsketchert.rs:160-202, all contigs of a
// genome counted into one FnvHashMap<Val, u64>, then one signature) and the f64-weighted maps of
// BlockSeqSketcher (src/sketching/seqblocksketch.rs:121-138).
//
// The weighted set is either an explicit (key, weight) list or the slots of a counting table
// (kmu_count.cu) filled from the sequences.  Every CTA sketches a slice of the items into a partial
// signature in shared memory -- slot = {h bits, key}, 128-bit CAS -- and merges it into the global one
// with the same (h, key)-minimum rule.  Items are cut at a bound B (an item stops once
// winv * (i - 1) >= B); the result is exact iff every slot ends below B, which the host verifies,
// raising B and repeating otherwise (B starts at (m / D) ln(m / 1e-4) for D distinct items).
#include <cstdint>

#include "kmu_device.cuh"
#include "kmu_kernels.h"

namespace kmu {

// all points of one item below the bound
template <typename V>
__device__ __forceinline__ void pmh3a_item_points(V key, double winv, double bound, const Pmh3aItemsParams& P, Slot* slots) {
    Xoshiro256pp rng;
    rng.seed(nohash_seed(key));
    for (uint32_t i = 1;; ++i) {
        const double base = __dmul_rn(winv, (double)(i - 1));
        if (!(base < bound)) break;
        const double x = exp01_sample(P.e, rng);
        const double h = __dadd_rn(base, __dmul_rn(winv, x));
        const uint32_t s = rng.unif_range(0, P.m, P.slot_thresh);
        if (h < bound) slot_update_min(&slots[s], (uint64_t)__double_as_longlong(h), (uint64_t)key);
    }
}
template <typename V>
__device__ __forceinline__ void pmh3a_item(V key, double winv, double bound, const Pmh3aItemsParams& P, Slot* slots) {
    // most items of a large set die on their first point: half a seeding tells (first_point_alive, kmu_device.cuh)
    if (first_point_alive<V>(key, winv, bound, P.e.c1)) pmh3a_item_points<V>(key, winv, bound, P, slots);
}

struct ItemQ {
    unsigned long long key;
    double winv;
};
static_assert(sizeof(ItemQ) * 64 * 32 == PMH3A_ITEMS_QUEUE_BYTES, "one queue of 64 items per warp");

// SRC 0: explicit lists keys[n] (V), weights[n] (f64); SRC 1: u32-key counting table (8-byte slots);
// SRC 2: u64-key counting table (16-byte slots, empty key ~0)
template <typename V, int SRC>
__global__ void __launch_bounds__(1024, 1) pmh3a_items_kernel(const Pmh3aItemsParams P) {
    extern __shared__ __align__(16) uint8_t smem[];
    Slot* slots = P.slots_in_smem ? (Slot*)smem : P.global_slots;
    const bool partial = P.slots_in_smem != 0;
    if (partial) {
        for (uint32_t j = threadIdx.x; j < P.m; j += blockDim.x) {
            slots[j].hbits = F64_MAX_BITS;
            slots[j].key = 0;
        }
        __syncthreads();
    }
    const V header = (V)word_header(P.kmer_type, P.k);
    // Two phases, so that the lanes of a warp stay together: every lane reads its entry and tests the item's first point
    // (half a seeding); the few items that survive go to the warp's queue and are finished 32 at a time.  Without the
    // queue a warp runs the whole item loop for one or two lanes at a time (8 of 32 lanes active under ncu).
    const int lane = threadIdx.x & 31;
    ItemQ* wq = (ItemQ*)(smem + (P.slots_in_smem ? (size_t)P.m * sizeof(Slot) : 0)) + (threadIdx.x >> 5) * 64;
    uint32_t qn = 0;  // warp-uniform
    for (uint64_t i0 = blockIdx.x * (uint64_t)blockDim.x + (threadIdx.x & ~31u); i0 < P.n; i0 += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = i0 + lane;
        V key = 0;
        double w = 0.0;
        if (i < P.n) {
            if (SRC == 0) {
                key = ((const V*)P.keys)[i];
                w = P.weights[i];
            } else if (SRC == 1) {
                const unsigned long long e = ((const unsigned long long*)P.table)[i];
                if (e != 0) {
                    key = finalize_key<V>((V)(e >> 32), header, P.hash_kind);
                    w = (double)(uint32_t)e;
                }
            } else {
                const ulonglong2 e = ((const ulonglong2*)P.table)[i];
                if (e.x != ~0ULL) {
                    key = finalize_key<V>((V)e.x, header, P.hash_kind);
                    w = (double)e.y;
                }
            }
        }
        double winv = 0.0;
        bool alive = false;
        if (w > 0.0) {
            winv = 1.0 / w;
            alive = first_point_alive<V>(key, winv, P.bound, P.e.c1);
        }
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, alive);
        if (alive) wq[qn + __popc(bal & ((1u << lane) - 1u))] = ItemQ{(unsigned long long)key, winv};
        qn += __popc(bal);
        __syncwarp();
        if (qn >= 32) {
            qn -= 32;
            const ItemQ it = wq[qn + lane];
            __syncwarp();
            pmh3a_item_points<V>((V)it.key, it.winv, P.bound, P, slots);
        }
    }
    if ((uint32_t)lane < qn) {
        const ItemQ it = wq[lane];
        pmh3a_item_points<V>((V)it.key, it.winv, P.bound, P, slots);
    }
    if (SRC == 2 && blockIdx.x == 0 && threadIdx.x == 0) {  // the u64 key equal to the table's empty mark
        const unsigned long long c = *P.special;
        if (c) pmh3a_item<V>(finalize_key<V>((V)~0ULL, header, P.hash_kind), 1.0 / (double)c, P.bound, P, slots);
    }
    if (partial) {
        __syncthreads();
        for (uint32_t j = threadIdx.x; j < P.m; j += blockDim.x)
            if (slots[j].hbits != F64_MAX_BITS) slot_update_min(&P.global_slots[j], slots[j].hbits, slots[j].key);
    }
}

__global__ void pmh3a_items_init_kernel(Slot* slots, uint32_t m) {
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) {
        slots[j].hbits = F64_MAX_BITS;
        slots[j].key = 0;
    }
}

// signature out + the largest slot value (as bits) for the host's verification
template <typename V>
__global__ void pmh3a_items_finish_kernel(const Slot* slots, uint32_t m, V* sig, unsigned long long* max_hbits) {
    unsigned long long mx = 0;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) {
        sig[j] = (V)slots[j].key;
        mx = slots[j].hbits > mx ? slots[j].hbits : mx;
    }
    mx = warp_max_u64(mx);
    if ((threadIdx.x & 31) == 0) atomicMax(max_hbits, mx);
}

template <typename V, int SRC>
static cudaError_t launch_items_t(const Pmh3aItemsParams& P, int grid, size_t smem, cudaStream_t st) {
    auto kern = pmh3a_items_kernel<V, SRC>;
    smem += PMH3A_ITEMS_QUEUE_BYTES;  // the warps' queues follow the slots
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    kern<<<grid, 1024, smem, st>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_pmh3a_items(const Pmh3aItemsParams& P, bool key64, int src, int grid, size_t smem, cudaStream_t st) {
    if (src == 0) return key64 ? launch_items_t<uint64_t, 0>(P, grid, smem, st) : launch_items_t<uint32_t, 0>(P, grid, smem, st);
    if (src == 1) return launch_items_t<uint32_t, 1>(P, grid, smem, st);
    return launch_items_t<uint64_t, 2>(P, grid, smem, st);
}

cudaError_t launch_pmh3a_items_init(Slot* slots, uint32_t m, cudaStream_t st) {
    pmh3a_items_init_kernel<<<(m + 255) / 256, 256, 0, st>>>(slots, m);
    return cudaGetLastError();
}

cudaError_t launch_pmh3a_items_finish(const Slot* slots, uint32_t m, bool key64, void* sig, unsigned long long* max_hbits,
                                      cudaStream_t st) {
    if (key64) pmh3a_items_finish_kernel<uint64_t><<<(m + 255) / 256, 256, 0, st>>>(slots, m, (uint64_t*)sig, max_hbits);
    else pmh3a_items_finish_kernel<uint32_t><<<(m + 255) / 256, 256, 0, st>>>(slots, m, (uint32_t*)sig, max_hbits);
    return cudaGetLastError();
}

}  // namespace kmu
