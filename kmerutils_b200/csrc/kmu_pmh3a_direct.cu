// kmu_pmh3a_direct.cu -- ProbMinHash3a for long sequences over a small key space (u32 k-mers, k <= 8), one pass.
//
// Same result as pmh3a_sketch_kernel (kmu_pmh3a.cu); different organisation.  The first point of an item is
// h = x(key) / count with (x, slot) a function of the key only (per-key memo table).  Every k-mer OCCURRENCE raises
// the key's counter in the shared-memory histogram and offers x / (new count) to the key's slot: the last occurrence
// offers the true first point, the earlier offers are larger and harmless.  During the pass a slot is one 64-bit
// word, the top 48 bits of its smallest offer over the 16-bit index of the key that made it (64-bit CAS, entered
// only by offers below the current value).  After the pass each slot recomputes x / count of the key it names, which
// restores the low bits; two offers of different keys with the same top 48 bits are seen as a tie and the sequence
// is flagged.  A later point of an item is >= 1 / count, so only items with 1 / count < q1 (q1 = largest
// slot value) can still matter; one scan of the histogram -- the same sweep that wipes it -- lists them and they draw
// their later points from their own Xoshiro256++ stream with the usual 128-bit slot updates.  There is no second walk
// over the sequence.  Flagged sequences (tie, race, wrapped u8 counter, too many items) are redone by the general
// kernel.
#include <cstdint>
#include <cstdio>

#include "kmu_device.cuh"
#include "kmu_kernels.h"

namespace kmu {

constexpr uint32_t DIRECT_T = 8;          // positions per task
constexpr uint32_t DIRECT_ITEMS = 1024;   // items that draw later points
constexpr uint32_t F64_MAX_HI = (uint32_t)(F64_MAX_BITS >> 32);

struct DirectShared {
    uint64_t byte_off;
    unsigned long long qbits;  // largest slot value after the pass (bit pattern)
    uint32_t seq, nbases, valid, flag, cmax, nitems;
};

__device__ __forceinline__ bool direct_update(Slot* slots, uint32_t* hi, uint32_t s, double h, uint32_t key) {
    const uint64_t hbits = (uint64_t)__double_as_longlong(h);
    if ((uint32_t)(hbits >> 32) > *(volatile uint32_t*)(hi + s)) return false;
    if (slot_update_min(&slots[s], hbits, (uint64_t)key)) {
        atomicMin(hi + s, (uint32_t)(hbits >> 32));
        return true;
    }
    return false;
}

template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) pmh3a_direct_kernel(const Pmh3aParams P) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ double s_winv[64];
    const uint32_t m = P.m, k = P.k;
    uint8_t* hist = smem;
    Slot* slots = (Slot*)(smem + P.regionA_bytes);
    uint32_t* hi = (uint32_t*)(smem + P.regionA_bytes + (size_t)m * 16);
    uint32_t* items = (uint32_t*)(smem + P.regionA_bytes + P.slots_smem_bytes);  // pk | count << 16
    unsigned long long* best = (unsigned long long*)items;  // during the pass, per slot: top 48 bits of the lowest offer | key index
    DirectShared* ds = (DirectShared*)(items + DIRECT_ITEMS);
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < 64) s_winv[tid] = tid ? 1.0 / (double)tid : 0.0;
    for (uint32_t j = tid; j < P.regionA_bytes / 16; j += NT) ((uint4*)hist)[j] = make_uint4(0, 0, 0, 0);
    const bool canonical = hash_is_canonical(P.hash_kind);
    const uint4* memo = (const uint4*)P.memo_fast;

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            const unsigned long long w = atomicAdd(P.work_counter, 1ULL);
            ds->valid = w < P.count;
            if (w < P.count) {
                const uint32_t seq = P.order[P.first + w];
                ds->seq = seq;
                ds->nbases = (uint32_t)P.nbases[seq];
                ds->byte_off = P.byte_off[seq];
            }
            ds->flag = 0;
            ds->cmax = 0;
            ds->nitems = 0;
            ds->qbits = 0;
        }
        for (uint32_t j = tid; j < m; j += NT) best[j] = ~0ULL;
        __syncthreads();
        if (!ds->valid) break;
        const uint32_t seq = ds->seq, L = ds->nbases;
        const uint32_t* words = (const uint32_t*)(P.packed + ds->byte_off);
        const uint32_t nk = L >= k ? L - k + 1 : 0;
        const uint32_t ntasks = (nk + DIRECT_T - 1) / DIRECT_T;

        // ---- the pass: count, offer first points ----
        uint32_t mymax = 0;
        bool bad = false;
        for (uint32_t task = tid; task < ntasks; task += NT) {
            const uint32_t p0 = task * DIRECT_T;
            const uint32_t nv = min(DIRECT_T, nk - p0);
            TaskKmers<uint32_t> tk;
            tk.init(words, p0, k);
            uint32_t pk[DIRECT_T];
            uint4 e[DIRECT_T];
#pragma unroll
            for (uint32_t t = 0; t < DIRECT_T; ++t) {
                pk[t] = tk.get(t, canonical);
                if (t < nv) e[t] = __ldg(memo + pk[t]);  // eight independent L2 lookups in flight
            }
#pragma unroll
            for (uint32_t t = 0; t < DIRECT_T; ++t) {
                if (t < nv) {
                    const uint32_t sh = (pk[t] & 3u) * 8;
                    const uint32_t old = atomicAdd((uint32_t*)hist + (pk[t] >> 2), 1u << sh);
                    const uint32_t c = (old >> sh) & 0xFFu, cn = c + 1;
                    bad |= c == 0xFFu;
                    mymax = cn > mymax ? cn : mymax;
                    const double winv = cn < 64 ? s_winv[cn] : 1.0 / (double)cn;
                    const double h = __dmul_rn(winv, __hiloint2double((int)e[t].y, (int)e[t].x));
                    // top 48 bits of h | key index: smaller wins; equal top bits with another key are left to the general kernel
                    const unsigned long long mine = ((unsigned long long)__double_as_longlong(h) & ~0xFFFFULL) | pk[t];
                    unsigned long long* slot = best + e[t].z;
                    unsigned long long cur = *(volatile unsigned long long*)slot;
                    while (mine < cur) {
                        if (((mine ^ cur) >> 16) == 0) break;
                        const unsigned long long seen = atomicCAS(slot, cur, mine);
                        if (seen == cur) break;
                        cur = seen;
                    }
                    bad |= ((mine ^ cur) >> 16) == 0 && mine != cur;
                }
            }
        }
        mymax = __reduce_max_sync(0xFFFFFFFFu, mymax);
        if (lane == 0 && mymax) atomicMax(&ds->cmax, mymax);
        if (bad) ds->flag = 1;
        __syncthreads();

        // ---- slots: recompute the winner's first point, check it against the recorded high word; q1 ----
        {
            unsigned long long mx = 0;
            for (uint32_t j = tid; j < m; j += NT) {
                const unsigned long long b = best[j];
                unsigned long long hbits = F64_MAX_BITS;
                uint32_t key = 0;
                if (b != ~0ULL) {
                    const uint32_t pkey = (uint32_t)b & 0xFFFFu;
                    const uint4 ee = __ldg(memo + pkey);
                    const uint32_t cnt = hist[pkey];
                    const double winv = cnt < 64 ? s_winv[cnt] : 1.0 / (double)cnt;
                    const double h = __dmul_rn(winv, __hiloint2double((int)ee.y, (int)ee.x));
                    hbits = (unsigned long long)__double_as_longlong(h);
                    if (cnt == 0 || ee.z != j || ((hbits ^ b) >> 16) != 0) ds->flag = 3;
                    key = ee.w;
                }
                hi[j] = (uint32_t)(hbits >> 32);
                slots[j].hbits = hbits;
                slots[j].key = key;
                mx = hbits > mx ? hbits : mx;
            }
            if (tid < ((m + 31) & ~31u)) {
                mx = warp_max_u64(mx);
                if (lane == 0) atomicMax(&ds->qbits, mx);
            }
        }
        __syncthreads();
        const double q1 = __longlong_as_double((long long)ds->qbits);
        const uint32_t cmax = ds->cmax;
        // smallest count whose items may place a later point: 1 / c < q1
        uint32_t cneed = 1;
        while (cneed < 64 && !(s_winv[cneed] < q1)) ++cneed;
        const double winv_cmax = cmax < 64 ? s_winv[cmax] : 1.0 / (double)cmax;
        const bool later = ds->flag == 0 && nk && !(winv_cmax >= q1);
        if (later && (cneed >= 64 || cneed < 2) && tid == 0) ds->flag = 2;  // every item (or huge counts): general kernel
        const bool scan = later && cneed >= 2 && cneed < 64;

        // ---- one sweep over the histogram: list the items with count >= cneed, wipe ----
        {
            const uint32_t need4 = cneed * 0x01010101u;
            for (uint32_t j = tid; j < P.regionA_bytes / 16; j += NT) {
                const uint4 v = ((uint4*)hist)[j];
                ((uint4*)hist)[j] = make_uint4(0, 0, 0, 0);
                if (scan) {
                    const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (uint32_t w = 0; w < 4; ++w) {
                        uint32_t hit = __vcmpgeu4(wv[w], need4);
                        while (hit) {
                            const uint32_t b = (__ffs(hit) - 1) >> 3;
                            hit &= ~(0xFFu << (b * 8));
                            const uint32_t pos = atomicAdd(&ds->nitems, 1u);
                            if (pos < DIRECT_ITEMS) items[pos] = (j * 16 + w * 4 + b) | (((wv[w] >> (b * 8)) & 0xFFu) << 16);
                        }
                    }
                }
            }
        }
        __syncthreads();
        if (scan) {
            const uint32_t nitems = ds->nitems;
            if (nitems > DIRECT_ITEMS) {
                if (tid == 0) ds->flag = 4;
            } else {
                for (uint32_t i = tid; i < nitems; i += NT) {
                    const uint32_t it = items[i];
                    const uint32_t cnt = it >> 16;
                    const uint32_t key = __ldg(memo + (it & 0xFFFFu)).w;
                    const double winv = cnt < 64 ? s_winv[cnt] : 1.0 / (double)cnt;
                    Xoshiro256pp rng;
                    rng.seed(nohash_seed(key));
                    (void)exp01_sample(P.e, rng);  // the first point was offered from the memo: keep the stream aligned
                    (void)rng.unif_range(0, m, P.slot_thresh);
                    for (uint32_t ip = 2;; ++ip) {
                        const double base = __dmul_rn(winv, (double)(ip - 1));
                        if (!(base < q1)) break;
                        const double x = exp01_sample(P.e, rng);
                        const double h = __dadd_rn(base, __dmul_rn(winv, x));
                        const uint32_t s = rng.unif_range(0, m, P.slot_thresh);
                        if (h < q1) direct_update(slots, hi, s, h, key);
                    }
                }
            }
            __syncthreads();
        }

        // ---- signature out, or hand the sequence to the general kernel ----
        if (ds->flag == 0) {
            uint32_t* out = (uint32_t*)P.sig + (size_t)seq * m;
            for (uint32_t j = tid; j < m; j += NT) out[j] = (uint32_t)slots[j].key;
        } else if (tid == 0) {
#ifdef KMU_DIRECT_DEBUG
            if (atomicAdd(P.overflow_count, 0ULL) < 40) printf("flag %u nk %u cmax %u q1 %g cneed %u nitems %u\n", ds->flag, nk, cmax, q1, cneed, ds->nitems);
#endif
            P.overflow_list[atomicAdd(P.overflow_count, 1ULL)] = seq;
        }
    }
}

size_t pmh3a_direct_smem_bytes(uint32_t k, uint32_t m) {
    size_t hist = (size_t)1 << (2 * k);
    if (hist < 16) hist = 16;
    const size_t slots = (((size_t)m * 20) + 15) & ~(size_t)15;
    if ((size_t)m * 8 > DIRECT_ITEMS * 4) return ~(size_t)0;  // the pass keeps its slots in the item list's space
    return hist + slots + DIRECT_ITEMS * 4 + sizeof(DirectShared) + 16;
}

template <int NT, int MINB>
static cudaError_t launch_direct_t(const Pmh3aParams& P, int grid, size_t smem, cudaStream_t stream) {
    auto kern = pmh3a_direct_kernel<NT, MINB>;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    kern<<<grid, NT, smem, stream>>>(P);
    return cudaGetLastError();
}

// variant: 0 = 256 threads x 3 CTAs / SM, 1 = 256 x 2, 2 = 512 x 1, 3 = 512 x 2
int pmh3a_direct_ctas_per_sm(int variant) { return variant == 0 ? 3 : (variant == 2 ? 1 : 2); }
cudaError_t launch_pmh3a_direct(const Pmh3aParams& P, int grid, int variant, cudaStream_t stream) {
    const size_t smem = pmh3a_direct_smem_bytes(P.k, P.m);
    switch (variant) {
        case 0: return launch_direct_t<256, 3>(P, grid, smem, stream);
        case 1: return launch_direct_t<256, 2>(P, grid, smem, stream);
        case 2: return launch_direct_t<512, 1>(P, grid, smem, stream);
        default: return launch_direct_t<512, 2>(P, grid, smem, stream);
    }
}

}  // namespace kmu
