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

struct DirectWork {
    uint64_t byte_off;
    uint32_t seq, nbases, valid, pad;
};
struct DirectState {
    unsigned long long qbits;  // largest slot value after the pass (bit pattern)
    uint32_t flag, cmax, nitems, pad;
};
struct DirectShared {
    DirectWork work[2];   // [parity of the sequence's turn]: fetched one turn ahead
    DirectState st[2];
    double winv[256];     // 1 / count
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

__device__ __forceinline__ void direct_fetch(const Pmh3aParams& P, DirectWork* w, unsigned long long i) {
    w->valid = i < P.count;
    if (i < P.count) {
        const uint32_t seq = P.order[P.first + i];
        w->seq = seq;
        w->nbases = (uint32_t)P.nbases[seq];
        w->byte_off = P.byte_off[seq];
    }
}

// an offer that may lower its slot (rare): 64-bit CAS; equal top 48 bits with another key -> tie
__device__ __forceinline__ bool direct_offer_slow(unsigned long long* slot, unsigned long long mine) {
    unsigned long long cur = *(volatile unsigned long long*)slot;
    while (mine < cur) {
        if (((mine ^ cur) >> 16) == 0) break;
        const unsigned long long seen = atomicCAS(slot, cur, mine);
        if (seen == cur) break;
        cur = seen;
    }
    return ((mine ^ cur) >> 16) == 0 && mine != cur;
}

template <int NT, int MINB, bool SPLIT>
__global__ void __launch_bounds__(NT, MINB) pmh3a_direct_kernel(const Pmh3aParams P) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t m = P.m, k = P.k;
    uint8_t* hist = smem;
    Slot* slots = (Slot*)(smem + P.regionA_bytes);
    uint32_t* hi = (uint32_t*)(smem + P.regionA_bytes + (size_t)m * 16);
    uint32_t* items = (uint32_t*)(smem + P.regionA_bytes + P.slots_smem_bytes);  // pk | count << 16
    unsigned long long* best = (unsigned long long*)items;  // during the pass, per slot: top 48 bits of the lowest offer | key index
    DirectShared* ds = (DirectShared*)(items + DIRECT_ITEMS);
    const double* s_winv = ds->winv;
    const int tid = threadIdx.x, lane = tid & 31;
    for (uint32_t j = tid; j < 256; j += NT) ds->winv[j] = j ? 1.0 / (double)j : 0.0;
    for (uint32_t j = tid; j < P.regionA_bytes / 16; j += NT) ((uint4*)hist)[j] = make_uint4(0, 0, 0, 0);
    for (uint32_t j = tid; j < m; j += NT) best[j] = ~0ULL;
    if (tid == 0) {
        direct_fetch(P, &ds->work[0], atomicAdd(P.work_counter, 1ULL));
        ds->st[0] = DirectState{0, 0, 0, 0, 0};
    }
    const bool canonical = hash_is_canonical(P.hash_kind);
    const uint4* memo = (const uint4*)P.memo_fast;
    __syncthreads();

    for (uint32_t turn = 0;; ++turn) {
        const DirectWork* wk = &ds->work[turn & 1];
        DirectState* st = &ds->st[turn & 1];
        if (!wk->valid) break;
        const uint32_t seq = wk->seq, L = wk->nbases;
        const uint32_t* words = (const uint32_t*)(P.packed + wk->byte_off);
        const uint32_t nk = L >= k ? L - k + 1 : 0;
        // next turn's sequence, off the critical path: the ticket is drawn now, its loads wait until after the pass
        unsigned long long ticket = 0;
        if (tid == NT - 1) ticket = atomicAdd(P.work_counter, 1ULL);

        // ---- the pass: count, offer first points.  Tasks of tlen <= 8 consecutive positions, thread-interleaved;
        //      tlen is chosen so that the last round of tasks is nearly full ----
        uint32_t mymax = 0;
        bool bad = false;
        const uint32_t rounds = (nk + NT * DIRECT_T - 1) / (NT * DIRECT_T);
        const uint32_t tlen = rounds ? (nk + NT * rounds - 1) / (NT * rounds) : 1;
        for (uint32_t p0 = tid * tlen; p0 < nk; p0 += NT * tlen) {
            const uint32_t nv = min(tlen, nk - p0);
            TaskKmers<uint32_t> tk;
            tk.init(words, p0, k);
            uint32_t pk[DIRECT_T], old[DIRECT_T], xlo[DIRECT_T], xhi[DIRECT_T], sl[DIRECT_T];
#pragma unroll
            for (uint32_t t = 0; t < DIRECT_T; ++t) {
                pk[t] = tk.get(t, canonical);
                if (t < nv) {  // eight independent L2 lookups in flight
                    const uint4 e = __ldg(memo + pk[t]);
                    xlo[t] = e.x;
                    xhi[t] = e.y;
                    sl[t] = e.z;
                }
            }
            if (SPLIT) {
#pragma unroll
                for (uint32_t t = 0; t < DIRECT_T; ++t)
                    if (t < nv) old[t] = atomicAdd((uint32_t*)hist + (pk[t] >> 2), 1u << ((pk[t] & 3u) * 8));
            }
#pragma unroll
            for (uint32_t t = 0; t < DIRECT_T; ++t) {
                if (t < nv) {
                    if (!SPLIT) old[t] = atomicAdd((uint32_t*)hist + (pk[t] >> 2), 1u << ((pk[t] & 3u) * 8));
                    const uint32_t cn = ((old[t] >> ((pk[t] & 3u) * 8)) & 0xFFu) + 1;  // 256: the u8 counter wrapped
                    mymax = cn > mymax ? cn : mymax;
                    const double h = __dmul_rn(s_winv[cn & 0xFFu], __hiloint2double((int)xhi[t], (int)xlo[t]));
                    unsigned long long* slot = best + sl[t];
                    if ((uint32_t)__double2hiint(h) <= ((volatile uint32_t*)slot)[1])
                        bad |= direct_offer_slow(slot, ((unsigned long long)__double_as_longlong(h) & ~0xFFFFULL) | pk[t]);
                }
            }
        }
        mymax = __reduce_max_sync(0xFFFFFFFFu, mymax);
        if (lane == 0 && mymax) atomicMax(&st->cmax, mymax);
        if (bad || mymax > 255) st->flag = 1;
        __syncthreads();
        if (tid == NT - 1) {
            direct_fetch(P, &ds->work[(turn + 1) & 1], ticket);
            ds->st[(turn + 1) & 1] = DirectState{0, 0, 0, 0, 0};
        }

        // ---- slots: recompute the winner's first point (restores the low bits), q1 ----
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
                    const double h = __dmul_rn(s_winv[cnt], __hiloint2double((int)ee.y, (int)ee.x));
                    hbits = (unsigned long long)__double_as_longlong(h);
                    if (cnt == 0 || ee.z != j || ((hbits ^ b) >> 16) != 0) st->flag = 3;
                    key = ee.w;
                }
                hi[j] = (uint32_t)(hbits >> 32);
                slots[j].hbits = hbits;
                slots[j].key = key;
                mx = hbits > mx ? hbits : mx;
            }
            if (tid < ((m + 31) & ~31u)) {
                mx = warp_max_u64(mx);
                if (lane == 0) atomicMax(&st->qbits, mx);
            }
        }
        __syncthreads();
        const double q1 = __longlong_as_double((long long)st->qbits);
        const uint32_t cmax = st->cmax;
        // smallest count whose items may place a later point: 1 / c < q1
        uint32_t cneed = q1 > 1.0 ? 1u : (q1 < 1.0 / 256.0 ? 256u : min(256u, (uint32_t)(1.0 / q1)));
        while (cneed > 1 && s_winv[cneed - 1] < q1) --cneed;
        while (cneed < 256 && !(s_winv[cneed] < q1)) ++cneed;
        const bool later = st->flag == 0 && nk && cmax >= cneed;
        const bool scan = later && cneed >= 2 && cneed < 256;
        if (later && !scan && tid == 0) st->flag = 2;  // every item needs later points: general kernel

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
                            const uint32_t pos = atomicAdd(&st->nitems, 1u);
                            if (pos < DIRECT_ITEMS) items[pos] = (j * 16 + w * 4 + b) | (((wv[w] >> (b * 8)) & 0xFFu) << 16);
                        }
                    }
                }
            }
        }
        __syncthreads();
        if (scan) {
            const uint32_t nitems = st->nitems;
            if (nitems > DIRECT_ITEMS) {
                if (tid == 0) st->flag = 4;
            } else {
                for (uint32_t i = tid; i < nitems; i += NT) {
                    const uint32_t it = items[i];
                    const uint32_t cnt = it >> 16;
                    const uint32_t key = __ldg(memo + (it & 0xFFFFu)).w;
                    const double winv = s_winv[cnt];
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

        // ---- signature out, or hand the sequence to the general kernel; the pass slots of the next turn ----
        if (st->flag == 0) {
            uint32_t* out = (uint32_t*)P.sig + (size_t)seq * m;
            for (uint32_t j = tid; j < m; j += NT) out[j] = (uint32_t)slots[j].key;
        } else if (tid == 0) {
            P.overflow_list[atomicAdd(P.overflow_count, 1ULL)] = seq;
        }
        for (uint32_t j = tid; j < m; j += NT) best[j] = ~0ULL;
        __syncthreads();
    }
}

size_t pmh3a_direct_smem_bytes(uint32_t k, uint32_t m) {
    size_t hist = (size_t)1 << (2 * k);
    if (hist < 16) hist = 16;
    const size_t slots = (((size_t)m * 20) + 15) & ~(size_t)15;
    if ((size_t)m * 8 > DIRECT_ITEMS * 4) return ~(size_t)0;  // the pass keeps its slots in the item list's space
    return hist + slots + DIRECT_ITEMS * 4 + sizeof(DirectShared) + 16;
}

template <int NT, int MINB, bool SPLIT>
static cudaError_t launch_direct_t(const Pmh3aParams& P, int grid, size_t smem, cudaStream_t stream) {
    auto kern = pmh3a_direct_kernel<NT, MINB, SPLIT>;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    kern<<<grid, NT, smem, stream>>>(P);
    return cudaGetLastError();
}

// variant: threads x CTAs / SM; 0 = 256 x 3, 1 = 256 x 2, 2 = 512 x 1, 3 = 512 x 2 (64 registers), 4 = 384 x 2, 5 = 3 without the
// split atomics loop
int pmh3a_direct_ctas_per_sm(int variant) { return variant == 0 ? 3 : (variant == 2 ? 1 : 2); }
cudaError_t launch_pmh3a_direct(const Pmh3aParams& P, int grid, int variant, cudaStream_t stream) {
    const size_t smem = pmh3a_direct_smem_bytes(P.k, P.m);
    switch (variant) {
        case 0: return launch_direct_t<256, 3, true>(P, grid, smem, stream);
        case 1: return launch_direct_t<256, 2, true>(P, grid, smem, stream);
        case 2: return launch_direct_t<512, 1, true>(P, grid, smem, stream);
        case 3: return launch_direct_t<512, 2, true>(P, grid, smem, stream);
        case 4: return launch_direct_t<384, 2, true>(P, grid, smem, stream);
        case 5: return launch_direct_t<512, 2, false>(P, grid, smem, stream);
        default: return launch_direct_t<384, 2, false>(P, grid, smem, stream);
    }
}

}  // namespace kmu
