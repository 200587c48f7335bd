// kmu_pmh3a_direct.cu -- ProbMinHash3a over a small key space (u32 k-mers, k <= 8), one pass over the sequence.
//
// Same result as pmh3a_sketch_kernel (kmu_pmh3a.cu); different organisation.  Point i of an item is
// h_i = (i - 1 + x_i(key)) / count with (x_i, slot_i) a function of the key only; the per-key memo table holds the
// first two.  Every k-mer OCCURRENCE raises the key's counter in the shared-memory histogram and offers its first NP
// points (NP = 1 for long sequences, 2 for short ones) computed with the NEW count to their slots: the last occurrence
// offers the true points, the earlier offers are larger and harmless.  During the pass a slot is one 64-bit word: the
// top 47 bits of its smallest offer, the point index, the 16-bit index of the key that made it (64-bit CAS, entered
// only by offers below the current value).  After the pass each slot recomputes the point it names from the final
// count, which restores the low bits; two offers with the same top 47 bits are seen as a tie and the sequence is
// flagged.  Point NP + 1 of an item is >= NP / count, so only items with NP / count < q1 (q1 = largest slot value) can
// still matter: they are listed -- by one scan of the histogram, the sweep that wipes it (long sequences), or from
// the keys whose count reached 2 during the pass (short sequences) -- and draw their later points from their own
// Xoshiro256++ stream with the usual 128-bit slot updates.  There is no second walk over the sequence.  Flagged
// sequences (tie, wrapped counter, too many items, every item needs later points) are redone by the general
// kernel.
//
// Why the result is the reference's: the sketch is, slot by slot, the minimum of (h, key) over ALL points of ALL
// items; the sequential algorithm only skips points that cannot be that minimum.  Here every point below q1 is
// offered (the memoised ones during the pass, the later ones from the item list), and a point that is not offered is
// >= NP / count >= q1 >= the value its slot ends with.
//
// One CTA per sequence, several sequences per SM (4-bit counters: 32 KB of histogram, four CTAs of 256 threads;
// 8-bit counters for the few sequences long enough to push a count past 15: two CTAs of 512 threads).  The packed
// bytes arrive by TMA bulk copies in segments of 4 STAGE positions, the next segment -- or the first one of the next
// sequence -- in flight while the current one is processed; work descriptors are fetched two turns ahead.  The
// kernel asks for the smallest shared-memory carve-out that holds its CTAs: L1 is where loads in flight wait.
#include <algorithm>
#include <cstdint>

#include "kmu_device.cuh"
#include "kmu_kernels.h"

namespace kmu {

constexpr uint32_t DIRECT_ITEMS = 512;   // items that draw later points (at least; 2 m when that is more: the pass slots live there)
constexpr uint32_t DIRECT_LIST2 = 1024;  // keys whose count reached 2 (short sequences)

struct DirectWork {
    uint64_t byte_off;
    uint32_t seq, nbases, valid, pad;
};
struct DirectState {
    unsigned long long qbits;  // largest slot value after the pass (bit pattern)
    uint32_t flag, cmax, nitems, n2;
};
struct DirectShared {
    DirectWork work[4];  // [turn & 3]: fetched two turns ahead (the first segment of the next sequence is staged during this one)
    DirectState st[2];
    long long t_mark;    // profiling runs: clock at the last phase mark
    long long pad;       // sizeof is a multiple of 16: the reciprocal table and the staging buffers follow
};
static_assert(sizeof(DirectShared) % 16 == 0, "staging buffers must be 16-byte aligned");

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

// pass slot word: h bits 63..17 | point index (bit 16) | key index
__device__ __forceinline__ unsigned long long direct_word(double h, uint32_t pt, uint32_t pk) {
    return ((unsigned long long)__double_as_longlong(h) & ~0x1FFFFULL) | (pt << 16) | pk;
}

// an offer that may lower its slot (rare): 64-bit minimum; equal top 47 bits with the value it met -> tie (the
// sequence is flagged and redone by the general kernel, so what the slot holds after a tie does not matter)
__device__ __forceinline__ bool direct_offer_slow(unsigned long long* slot, unsigned long long mine) {
    const unsigned long long old = atomicMin(slot, mine);
    return ((mine ^ old) >> 17) == 0 && mine != old;
}

// point pt (0, 1) of an item from its memo entry {x lo, x hi, slot, key} and 1 / count
__device__ __forceinline__ double direct_point(const uint4& e, uint32_t pt, double winv) {
    const double wx = __dmul_rn(winv, __hiloint2double((int)e.y, (int)e.x));
    return pt ? __dadd_rn(winv, wx) : wx;  // base of point 2 = winv * 1
}

// NT threads per sequence, T positions per task, NP memoised points offered per occurrence, LIST: the items that
// may need later points come from the keys seen twice (else from a scan of the histogram)
// HB: bits per histogram counter (8, or 4: half the shared memory, twice the sequences in flight per SM; a
// count above 15 flags the sequence)
// STAGE: bytes per staging buffer (two per CTA): the packed bytes travel global -> shared by TMA bulk copies, one
// segment of 4 STAGE positions ahead of the pass (0: the pass reads global memory)
template <int NT, int MINB, int T, int NP, bool LIST, int HB, int STAGE>
__global__ void __launch_bounds__(NT, MINB) pmh3a_direct_kernel(const Pmh3aParams P) {
    constexpr uint32_t CMAX = (1u << HB) - 1;
    constexpr uint32_t PS = STAGE * 4;            // positions per segment
    constexpr uint32_t STAGE_BUF = STAGE + 32;    // segment + halo + the word pair read past the last window
    auto hist_count = [](const uint8_t* h, uint32_t pkey) -> uint32_t {
        return HB == 8 ? (uint32_t)h[pkey] : ((uint32_t)h[pkey >> 1] >> ((pkey & 1u) * 4)) & 15u;
    };
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t m = P.m, k = P.k;
    uint8_t* hist = smem;
    Slot* slots = (Slot*)(smem + P.regionA_bytes);
    uint32_t* hi = (uint32_t*)(smem + P.regionA_bytes + (size_t)m * 16);
    uint32_t* items = (uint32_t*)(smem + P.regionA_bytes + P.slots_smem_bytes);  // pk | count << 16
    unsigned long long* best = (unsigned long long*)items;  // during the pass, per slot: the lowest offer (direct_word)
    const uint32_t items_cap = (max(DIRECT_ITEMS, 2 * m) + 3) & ~3u;  // what follows stays 16-byte aligned
    uint16_t* list2 = (uint16_t*)(items + items_cap);         // LIST forms only
    DirectShared* ds = (DirectShared*)(list2 + (LIST ? DIRECT_LIST2 : 0));
    double* winv_tab = (double*)(ds + 1);                      // [CMAX + 1]: 1 / count
    uint8_t* stage = (uint8_t*)(winv_tab + CMAX + 1);          // [2][STAGE_BUF]
    uint64_t* bars = (uint64_t*)(stage + 2 * STAGE_BUF);       // [2] "segment landed"
    uint32_t buf = 0, parity = 0;                              // staging buffer of the next segment; bit b: phase of bars[b]
    auto issue = [&](const uint8_t* row, uint32_t nb, uint32_t seg, uint32_t b) {  // one thread
        const uint32_t row_padded = (((nb + 3) / 4) + 15) & ~15u, off = seg * STAGE;
        const uint32_t bytes = min((uint32_t)STAGE + 16, row_padded - off);
        mbar_arrive_expect_tx(&bars[b], bytes);
        tma_load_bytes(stage + b * STAGE_BUF, row + off, bytes, &bars[b]);
    };
    const double* s_winv = winv_tab;
    const int tid = threadIdx.x, lane = tid & 31;
    for (uint32_t j = tid; j <= CMAX; j += NT) winv_tab[j] = j ? 1.0 / (double)j : 0.0;
    for (uint32_t j = tid; j < P.regionA_bytes / 16; j += NT) ((uint4*)hist)[j] = make_uint4(0, 0, 0, 0);
    for (uint32_t j = tid; j < m; j += NT) best[j] = ~0ULL;
    if (tid == 0) {
        direct_fetch(P, &ds->work[0], atomicAdd(P.work_counter, 1ULL));
        direct_fetch(P, &ds->work[1], atomicAdd(P.work_counter, 1ULL));
        ds->st[0] = DirectState{0, 0, 0, 0, 0};
        if (STAGE) {
            mbar_init(&bars[0], 1);
            mbar_init(&bars[1], 1);
            mbar_fence_init();
            if (ds->work[0].valid && ds->work[0].nbases >= k) issue(P.packed + ds->work[0].byte_off, ds->work[0].nbases, 0, 0);
        }
    }
    const bool canonical = hash_is_canonical(P.hash_kind);
    const uint4* memo = (const uint4*)P.memo_fast;  // [2 pk] first point, [2 pk + 1] second point
    __syncthreads();

    // optional phase timing (profiling runs only): thread 0 accumulates clock deltas.
    // 0 pass, 1 wait for the slowest thread of the pass, 2 slots + q1, 3 sweep, 4 later points, 5 output
    // (the running mark lives in shared memory: no register is held across the pass for it)
    if (P.phase_clocks && tid == 0) ds->t_mark = clock64();
    auto mark = [&](int phase) {
        if (P.phase_clocks && tid == 0) {
            const long long now = clock64();
            atomicAdd(P.phase_clocks + phase, (unsigned long long)(now - ds->t_mark));
            ds->t_mark = now;
        }
    };
    for (uint32_t turn = 0;; ++turn) {
        const DirectWork* wk = &ds->work[turn & 3];
        DirectState* st = &ds->st[turn & 1];
        if (!wk->valid) break;
        const uint32_t seq = wk->seq, L = wk->nbases;
        const uint32_t* words = (const uint32_t*)(P.packed + wk->byte_off);
        const uint32_t nk = L >= k ? L - k + 1 : 0;
        // the sequence after next, off the critical path: the ticket is drawn now, its loads wait until after the pass
        unsigned long long ticket = 0;
        if (tid == NT - 1) ticket = atomicAdd(P.work_counter, 1ULL);

        // ---- the pass: count, offer the memoised points.  Tasks of tlen <= T consecutive positions,
        //      thread-interleaved; tlen is chosen so that the last round of tasks is nearly full ----
        uint32_t mymax = 0;
        bool bad = false;
        const uint32_t nseg = STAGE ? (nk + PS - 1) / PS : 1;
        for (uint32_t seg = 0; seg < nseg; ++seg) {
        const uint32_t* seg_words = words;
        uint32_t seg_nk = nk;
        if (STAGE) {
            seg_nk = min(PS, nk - seg * PS);
            if (tid == NT - 1) {  // the next segment, or the first one of the next sequence, lands while this one is processed
                const DirectWork* nw = &ds->work[(turn + 1) & 3];
                if (seg + 1 < nseg) issue((const uint8_t*)words, L, seg + 1, buf ^ 1);
                else if (nw->valid && nw->nbases >= k) issue(P.packed + nw->byte_off, nw->nbases, 0, buf ^ 1);
            }
            mbar_wait(&bars[buf], (parity >> buf) & 1u);
            parity ^= 1u << buf;
            seg_words = (const uint32_t*)(stage + buf * STAGE_BUF);
        }
        const uint32_t rounds = (seg_nk + NT * T - 1) / (NT * T);
        const uint32_t tlen = rounds ? (seg_nk + NT * rounds - 1) / (NT * rounds) : 1;
        for (uint32_t p0 = tid * tlen; p0 < seg_nk; p0 += NT * tlen) {
            const uint32_t nv = min(tlen, seg_nk - p0);
            TaskKmers<uint32_t> tk;
            tk.init(seg_words, p0, k);
            tk.template narrow<T>();  // k <= 8, T <= 8: 2 (T - 1) + 2 k <= 30 bits
            uint32_t pk[T];
            uint4 e1[T], e2[T];
#pragma unroll
            for (uint32_t t = 0; t < T; ++t) {
                pk[t] = tk.template get_narrow<T>(t, canonical);
                if (t < nv) {  // independent L2 lookups in flight (both points sit in one 32-byte sector)
                    e1[t] = __ldcg(memo + 2 * pk[t]);  // L2 only: L1 keeps the sequence bytes
                    if (NP == 2) e2[t] = __ldcg(memo + 2 * pk[t] + 1);
                }
            }
#pragma unroll
            for (uint32_t t = 0; t < T; ++t) {
                if (t < nv) {
                    const uint32_t sh = HB == 8 ? (pk[t] & 3u) * 8 : (pk[t] & 7u) * 4;
                    const uint32_t old = atomicAdd((uint32_t*)hist + (pk[t] >> (HB == 8 ? 2 : 3)), 1u << sh);
                    const uint32_t cn = ((old >> sh) & CMAX) + 1;  // CMAX + 1: the counter wrapped
                    mymax = cn > mymax ? cn : mymax;
                    if (LIST && cn == 2) {
                        const uint32_t pos = atomicAdd(&st->n2, 1u);
                        if (pos < DIRECT_LIST2) list2[pos] = (uint16_t)pk[t];
                    }
                    const double winv = s_winv[cn & CMAX];
                    {
                        const double h = direct_point(e1[t], 0, winv);
                        unsigned long long* slot = best + e1[t].z;
                        if ((uint32_t)__double2hiint(h) <= ((volatile uint32_t*)slot)[1]) bad |= direct_offer_slow(slot, direct_word(h, 0, pk[t]));
                    }
                    if (NP == 2) {
                        const double h = direct_point(e2[t], 1, winv);
                        unsigned long long* slot = best + e2[t].z;
                        if ((uint32_t)__double2hiint(h) <= ((volatile uint32_t*)slot)[1]) bad |= direct_offer_slow(slot, direct_word(h, 1, pk[t]));
                    }
                }
            }
        }
        if (STAGE) {
            if (seg + 1 < nseg) __syncthreads();  // everybody is done with this buffer before the segment after next lands in it
            buf ^= 1;
        }
        }
        mymax = __reduce_max_sync(0xFFFFFFFFu, mymax);
        if (lane == 0 && mymax) atomicMax(&st->cmax, mymax);
        if (bad || mymax > CMAX) st->flag = 1;
        mark(0);
        __syncthreads();
        mark(1);
        if (tid == NT - 1) {
            direct_fetch(P, &ds->work[(turn + 2) & 3], ticket);
            ds->st[(turn + 1) & 1] = DirectState{0, 0, 0, 0, 0};
            if (STAGE && nseg == 0) {  // nothing was staged behind this (empty) sequence: the next one's first segment
                const DirectWork* nw = &ds->work[(turn + 1) & 3];
                if (nw->valid && nw->nbases >= k) issue(P.packed + nw->byte_off, nw->nbases, 0, buf);
            }
        }

        // ---- slots: recompute the winning point (restores the low bits), q1 ----
        {
            unsigned long long mx = 0;
            for (uint32_t j = tid; j < m; j += NT) {
                const unsigned long long b = best[j];
                unsigned long long hbits = F64_MAX_BITS;
                uint32_t key = 0;
                if (b != ~0ULL) {
                    const uint32_t pkey = (uint32_t)b & 0xFFFFu, pt = ((uint32_t)b >> 16) & 1u;
                    const uint4 ee = __ldg(memo + 2 * pkey + pt);
                    const uint32_t cnt = hist_count(hist, pkey);
                    hbits = (unsigned long long)__double_as_longlong(direct_point(ee, pt, s_winv[cnt]));
                    if (cnt == 0 || ee.z != j || ((hbits ^ b) >> 17) != 0) st->flag = 3;
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
        mark(2);
        const double q1 = __longlong_as_double((long long)st->qbits);
        const uint32_t cmax = st->cmax;
        // smallest count whose items may place point NP + 1: NP / c < q1
        const double np = (double)NP;
        constexpr uint32_t CNONE = CMAX + 1;  // no count qualifies
        uint32_t cneed = q1 > np ? 1u : (q1 < np / (double)CNONE ? CNONE : min(CNONE, (uint32_t)(np / q1)));
        while (cneed > 1 && __dmul_rn(s_winv[cneed - 1], np) < q1) --cneed;
        while (cneed < CNONE && !(__dmul_rn(s_winv[cneed], np) < q1)) ++cneed;
        const bool later = st->flag == 0 && nk && cmax >= cneed;
        const bool scan = later && cneed >= 2 && cneed < CNONE;
        if (later && !scan && tid == 0) st->flag = 2;  // every item needs later points: general kernel

        // ---- the items with count >= cneed.  Scan: one sweep over the histogram lists them and wipes it ----
        if (!LIST) {
            if (scan) {
                const uint32_t need4 = cneed * 0x01010101u;
#pragma unroll 2
                for (uint32_t j = tid; j < P.regionA_bytes / 16; j += NT) {
                    const uint4 v = ((uint4*)hist)[j];
                    ((uint4*)hist)[j] = make_uint4(0, 0, 0, 0);
                    const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
                    if (HB == 8) {
#pragma unroll
                        for (uint32_t w = 0; w < 4; ++w) {
                            uint32_t hit = __vcmpgeu4(wv[w], need4);
                            while (hit) {
                                const uint32_t b = (__ffs(hit) - 1) >> 3;
                                hit &= ~(0xFFu << (b * 8));
                                const uint32_t pos = atomicAdd(&st->nitems, 1u);
                                if (pos < items_cap) items[pos] = (j * 16 + w * 4 + b) | (((wv[w] >> (b * 8)) & 0xFFu) << 16);
                            }
                        }
                    } else {
                        // even / odd nibbles as bytes; one test for the whole 16 bytes (32 counters), the listing loop only
                        // where a counter qualifies -- almost never
                        uint32_t he[4], ho[4], any = 0;
#pragma unroll
                        for (uint32_t w = 0; w < 4; ++w) {
                            he[w] = __vcmpgeu4(wv[w] & 0x0F0F0F0Fu, need4);
                            ho[w] = __vcmpgeu4((wv[w] >> 4) & 0x0F0F0F0Fu, need4);
                            any |= he[w] | ho[w];
                        }
                        if (any) {
#pragma unroll
                            for (uint32_t w = 0; w < 4; ++w) {
#pragma unroll
                                for (uint32_t half = 0; half < 2; ++half) {
                                    const uint32_t nib = (wv[w] >> (half * 4)) & 0x0F0F0F0Fu;
                                    uint32_t hit = half ? ho[w] : he[w];
                                    while (hit) {
                                        const uint32_t b = (__ffs(hit) - 1) >> 3;
                                        hit &= ~(0xFFu << (b * 8));
                                        const uint32_t pos = atomicAdd(&st->nitems, 1u);
                                        if (pos < items_cap) items[pos] = (j * 32 + w * 8 + b * 2 + half) | (((nib >> (b * 8)) & 0xFFu) << 16);
                                    }
                                }
                            }
                        }
                    }
                }
            } else {
#pragma unroll 4
                for (uint32_t j = tid; j < P.regionA_bytes / 16; j += NT) ((uint4*)hist)[j] = make_uint4(0, 0, 0, 0);
            }
            __syncthreads();
        }
        mark(3);
        if (scan) {
            const uint32_t nitems = LIST ? st->n2 : st->nitems;
            if (nitems > (LIST ? DIRECT_LIST2 : items_cap)) {
                if (tid == 0) st->flag = 4;
            } else {
                for (uint32_t i = tid; i < nitems; i += NT) {
                    uint32_t pkey, cnt;
                    if (LIST) {
                        pkey = list2[i];
                        cnt = hist_count(hist, pkey);
                        if (cnt < cneed) continue;
                    } else {
                        pkey = items[i] & 0xFFFFu;
                        cnt = items[i] >> 16;
                    }
                    const double winv = s_winv[cnt];
                    const uint4 ee = __ldg(memo + 2 * pkey + 1);  // the item's second point
                    const uint32_t key = ee.w;
                    if (NP == 1) {
                        const double h2 = direct_point(ee, 1, winv);
                        if (h2 < q1) direct_update(slots, hi, ee.z, h2, key);
                        if (!(__dmul_rn(winv, 2.0) < q1)) continue;
                    }
                    Xoshiro256pp rng;  // third point on: the key's own stream, aligned past the two memoised points
                    rng.seed(nohash_seed(key));
                    (void)exp01_sample(P.e, rng);
                    (void)rng.unif_range(0, m, P.slot_thresh);
                    (void)exp01_sample(P.e, rng);
                    (void)rng.unif_range(0, m, P.slot_thresh);
                    for (uint32_t ip = 3;; ++ip) {
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
            mark(4);
        }
        if (LIST) {  // the later points read the counts: wipe afterwards
#pragma unroll 4
            for (uint32_t j = tid; j < P.regionA_bytes / 16; j += NT) ((uint4*)hist)[j] = make_uint4(0, 0, 0, 0);
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
        mark(5);
    }
}

size_t pmh3a_direct_hist_bytes(uint32_t k, int variant);
size_t pmh3a_direct_stage_bytes(int variant);
bool pmh3a_direct_lists(int variant);
static size_t hist_full(uint32_t k) { return std::max<size_t>(16, (size_t)1 << (2 * k)); }
size_t pmh3a_direct_smem_bytes(uint32_t k, uint32_t m, int variant) {
    size_t hist = pmh3a_direct_hist_bytes(k, variant);
    const size_t slots = (((size_t)m * 20) + 15) & ~(size_t)15;
    const size_t items = (size_t)((std::max<uint32_t>(DIRECT_ITEMS, 2 * m) + 3) & ~3u) * 4;
    const size_t list2 = pmh3a_direct_lists(variant) ? DIRECT_LIST2 * 2 : 0;
    const size_t winv = (pmh3a_direct_hist_bytes(k, variant) < hist_full(k) ? 16 : 256) * sizeof(double);
    const size_t stage = pmh3a_direct_stage_bytes(variant) ? 2 * (pmh3a_direct_stage_bytes(variant) + 32) + 16 : 0;
    return hist + slots + items + list2 + sizeof(DirectShared) + winv + stage + 16;
}

template <int NT, int MINB, int T, int NP, bool LIST, int HB = 8, int STAGE = 0>
static cudaError_t launch_direct_t(const Pmh3aParams& P, int grid, size_t smem, cudaStream_t stream) {
    auto kern = pmh3a_direct_kernel<NT, MINB, T, NP, LIST, HB, STAGE>;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        // the smallest shared-memory carve-out that holds MINB CTAs: what is left is L1, and L1 is where loads in
        // flight wait -- the kernel's memory-level parallelism
        const int pct = (int)std::min<size_t>(100, ((smem + 1024) * MINB * 100 + 233471) / 233472);
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    kern<<<grid, NT, smem, stream>>>(P);
    return cudaGetLastError();
}

// The forms of the kernel (kmu_capi.cu picks one per length class):
//   0  very long sequences: 8-bit counters, 512 threads, two sequences per SM, one point per occurrence
//   1  long sequences: 4-bit counters, 256 threads, four sequences per SM, one point per occurrence, TMA staging
//   2  short sequences: as 1 without staging, two points per occurrence, items from the keys seen twice
//   3  as 1 without staging (measurements)
struct DirectForm {
    int threads, ctas_per_sm, hist_bits, stage_bytes;
    bool lists;
};
static DirectForm direct_form(int variant) {
    switch (variant) {
        case 0: return {512, 2, 8, 0, false};
        case 2: return {256, 4, 4, 0, true};
        case 3: return {256, 4, 4, 0, false};
        default: return {256, 4, 4, 4096, false};
    }
}
int pmh3a_direct_ctas_per_sm(int variant) { return direct_form(variant).ctas_per_sm; }
int pmh3a_direct_threads(int variant) { return direct_form(variant).threads; }
bool pmh3a_direct_lists(int variant) { return direct_form(variant).lists; }
size_t pmh3a_direct_stage_bytes(int variant) { return (size_t)direct_form(variant).stage_bytes; }
size_t pmh3a_direct_hist_bytes(uint32_t k, int variant) {
    const size_t hist = ((size_t)1 << (2 * k)) * direct_form(variant).hist_bits / 8;
    return hist < 16 ? 16 : hist;
}
cudaError_t launch_pmh3a_direct(const Pmh3aParams& P, int grid, int variant, cudaStream_t stream) {
    const size_t smem = pmh3a_direct_smem_bytes(P.k, P.m, variant);
    switch (variant) {
        case 0: return launch_direct_t<512, 2, 8, 1, false>(P, grid, smem, stream);
        case 2: return launch_direct_t<256, 4, 4, 2, true, 4>(P, grid, smem, stream);
        case 3: return launch_direct_t<256, 4, 4, 1, false, 4>(P, grid, smem, stream);
        default: return launch_direct_t<256, 4, 4, 1, false, 4, 4096>(P, grid, smem, stream);
    }
}

}  // namespace kmu
