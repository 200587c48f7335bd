// kmu_pmh3a.cu -- per-sequence ProbMinHash3a sketching on sm_100a.
//
// Replaces the CPU loop of SeqSketcher::sketch_probminhash3a
// (src/sketching/seqsketchjaccard.rs:211-260) and ProbHash3aSketch::sketch_compressedkmer
// (src/sketching/setsketchert.rs:121-157): per sequence
//     KmerSeqIterator -> fhash -> FnvHashMap<Val,u64> multiplicities
//     -> ProbMinHash3a::hash_weigthed_hashmap -> signature of m slots.
//
// GPU formulation (see DESIGN.md "ProbMinHash3a kernel"):
//   * a TEAM (1..32 warps of one CTA) owns one sequence at a time; teams pull
//     sequences from a global work counter in descending-length order.
//   * pass 1 walks the packed 2-bit bases (rolling forward / reverse-complement
//     windows in registers) and counts pre-keys in shared memory: a direct
//     u16 histogram when 4^k bins fit, else an open-addressing table (shared
//     memory, or an L2-resident global scratch for long sequences).
//   * pass 2 walks the sequence again; the occurrence that atomically *claims*
//     a pre-key owns that distinct item.  Claimed (key, multiplicity) pairs are
//     compacted into per-warp queues with ballots so that the expensive part --
//     seeding Xoshiro256++ and drawing the item's exponential points -- always
//     runs on full warps.
//   * the signature slots are 16-byte {h, key} records updated with one 128-bit
//     CAS; the result is argmin over all points per slot, which is what the
//     sequential algorithm computes (its qmax tests only prune points that cannot
//     win), so every item can run depth-first against a lazily refreshed qmax.
#include <cstdint>
#include <cstdio>

#include "kmu_device.cuh"
#include "kmu_kernels.h"

namespace kmu {

// ---- multiplicity stores -----------------------------------------------------------
// u32 pre-keys: one u64 per entry = key << 32 | claimed << 31 | count ; 0 == empty
// u64 pre-keys: 16-byte entry {key, claimed << 63 | count} ; count == 0 == empty
// insert() tells whether this call created the entry (first occurrence of the key).
template <typename V>
struct TableOps;

template <>
struct TableOps<uint32_t> {
    using Entry = unsigned long long;
    static __device__ __forceinline__ uint32_t slot_of(uint32_t key, uint32_t log2cap) {
        return (key * 0x9E3779B1u) >> (32 - log2cap);
    }
    static __device__ __forceinline__ bool insert(Entry* tab, uint32_t capmask, uint32_t log2cap, uint32_t key) {
        uint32_t i = slot_of(key, log2cap);
        for (;;) {
            Entry e = *(volatile Entry*)(tab + i);
            if (e == 0) {
                Entry old = atomicCAS(tab + i, 0ULL, ((Entry)key << 32) | 1ULL);
                if (old == 0) return true;
                e = old;
            }
            if ((uint32_t)(e >> 32) == key) {
                atomicAdd(tab + i, 1ULL);
                return false;
            }
            i = (i + 1) & capmask;
        }
    }
    // multiplicity of a key that is in the table (plain loads; after the pass-1 barrier)
    static __device__ __forceinline__ uint32_t lookup(const Entry* tab, uint32_t capmask, uint32_t log2cap, uint32_t key) {
        uint32_t i = slot_of(key, log2cap);
        for (;;) {
            Entry e = *(volatile const Entry*)(tab + i);
            if (e == 0) return 0;
            if ((uint32_t)(e >> 32) == key) return (uint32_t)e & 0x7FFFFFFFu;
            i = (i + 1) & capmask;
        }
    }
    // returns the multiplicity if this call claimed the key, else 0
    static __device__ __forceinline__ uint32_t claim(Entry* tab, uint32_t capmask, uint32_t log2cap, uint32_t key) {
        uint32_t i = slot_of(key, log2cap);
        for (;;) {
            Entry e = *(volatile Entry*)(tab + i);
            if (e == 0) return 0;  // cannot happen for a key inserted in pass 1
            if ((uint32_t)(e >> 32) == key) {
                if (e & 0x80000000ULL) return 0;
                Entry old = atomicOr(tab + i, 0x80000000ULL);
                return (old & 0x80000000ULL) ? 0u : (uint32_t)(old & 0x7FFFFFFFULL);
            }
            i = (i + 1) & capmask;
        }
    }
};

template <>
struct TableOps<uint64_t> {
    struct __align__(16) Entry {
        unsigned long long key;
        unsigned long long cnt;
    };
    static __device__ __forceinline__ uint32_t slot_of(uint64_t key, uint32_t log2cap) {
        return (uint32_t)((key * 0x9E3779B97F4A7C15ULL) >> (64 - log2cap));
    }
    static __device__ __forceinline__ bool insert(Entry* tab, uint32_t capmask, uint32_t log2cap, uint64_t key) {
        uint32_t i = slot_of(key, log2cap);
        for (;;) {
            unsigned long long c = *(volatile unsigned long long*)&tab[i].cnt;
            if (c == 0) {
                uint64_t oh, ok;
                cas128((Slot*)(tab + i), 0, 0, key, 1, oh, ok);
                if (oh == 0 && ok == 0) return true;
            }
            unsigned long long kk = *(volatile unsigned long long*)&tab[i].key;
            if (kk == key) {
                atomicAdd(&tab[i].cnt, 1ULL);
                return false;
            }
            i = (i + 1) & capmask;
        }
    }
    static __device__ __forceinline__ uint32_t lookup(const Entry* tab, uint32_t capmask, uint32_t log2cap, uint64_t key) {
        uint32_t i = slot_of(key, log2cap);
        for (;;) {
            unsigned long long c = *(volatile const unsigned long long*)&tab[i].cnt;
            if (c == 0) return 0;
            if (*(volatile const unsigned long long*)&tab[i].key == key) return (uint32_t)(c & 0x7FFFFFFFULL);
            i = (i + 1) & capmask;
        }
    }
    static __device__ __forceinline__ uint32_t claim(Entry* tab, uint32_t capmask, uint32_t log2cap, uint64_t key) {
        uint32_t i = slot_of(key, log2cap);
        for (;;) {
            unsigned long long c = *(volatile unsigned long long*)&tab[i].cnt;
            if (c == 0) return 0;
            unsigned long long kk = *(volatile unsigned long long*)&tab[i].key;
            if (kk == key) {
                if (c >> 63) return 0;
                unsigned long long old = atomicOr(&tab[i].cnt, 1ULL << 63);
                return (old >> 63) ? 0u : (uint32_t)(old & 0x7FFFFFFFULL);
            }
            i = (i + 1) & capmask;
        }
    }
};

template <typename V>
struct QItem {
    V key;  // fhash(kmer): the value that goes into the signature
    uint32_t cnt;
};

constexpr int QCAP = 64;  // per-warp queue ring (power of two, >= 2 * 32)

// per-team bookkeeping in shared memory (double buffered work descriptors + TMA barriers)
struct __align__(16) TeamShared {
    uint64_t mbar[2];
    uint64_t byte_off[2];
    uint32_t seq[2];
    uint32_t nbases[2];
    uint32_t valid[2];
    uint32_t staged[2];
    uint32_t qmax_hi;
    uint32_t flag;
    uint32_t next_task[2];  // dynamic task counters: [0] pass 1 and fill phase, [1] pass 2
};
static_assert(sizeof(TeamShared) == PMH3A_TEAM_SHARED_BYTES, "TeamShared size");

// Sketch state of one team: the 16-byte slots and a mirror of the high words of their h values.
struct SlotArray {
    Slot* slots;
    uint32_t* hi;  // hi[j] >= high word of slots[j].hbits at any time (stale values are larger)
    uint32_t* s_qmax_hi;
};

// Cheap upper bound of the slot maxima (MaxValueTracker's root): only the mirrored high words
// are scanned (conflict-free), the bound is (max_hi, 0xFFFFFFFF).  Any stale or loose bound is
// safe: qmax tests only prune points that cannot win a slot.
__device__ __forceinline__ void refresh_qmax(const SlotArray& S, uint32_t m, int lane) {
    uint32_t mx = 0;
    for (uint32_t j = lane; j < m; j += 32) {
        uint32_t hi = *(volatile const uint32_t*)(S.hi + j);
        mx = hi > mx ? hi : mx;
    }
    mx = __reduce_max_sync(0xFFFFFFFFu, mx);
    if (lane == 0 && mx < *(volatile uint32_t*)S.s_qmax_hi) atomicMin(S.s_qmax_hi, mx);
    __syncwarp();
}
__device__ __forceinline__ double load_qmax(const SlotArray& S) {
    uint32_t hi = *(volatile const uint32_t*)S.s_qmax_hi;
    return __hiloint2double((int)hi, (int)0xFFFFFFFFu);
}
__device__ __forceinline__ void sketch_update(const SlotArray& S, uint32_t s, double h, uint64_t key) {
    const uint64_t hbits = (uint64_t)__double_as_longlong(h);
    if (slot_update_min(&S.slots[s], hbits, key)) *(volatile uint32_t*)(S.hi + s) = (uint32_t)(hbits >> 32);
}

constexpr uint32_t QFLAG_SKIP_FIRST = 0x80000000u;  // QItem.cnt: the item's first point is already in the sketch
constexpr uint32_t QFLAG_SKIP_TWO = 0x40000000u;    // QItem.cnt: so are its first two points
constexpr uint32_t QFLAG_MASK = QFLAG_SKIP_FIRST | QFLAG_SKIP_TWO;

// --------------------------------------------------------------------------------
// Process up to 32 queued distinct items with one warp: every lane owns one item and
// emits its points i = 1, 2, ... while winv * (i - 1) < qmax
// (ProbMinHash3a::hash_weigthed_hashmap).
// --------------------------------------------------------------------------------
template <typename V>
__device__ __forceinline__ void process_items(const QItem<V>* queue, uint32_t head, uint32_t n, int lane,
                                              const Pmh3aParams& P, const SlotArray& S, uint32_t refresh_period) {
    bool act = (uint32_t)lane < n;
    V key = 0;
    double winv = 0.0;
    Xoshiro256pp rng;
    rng.s0 = rng.s1 = rng.s2 = rng.s3 = 0;
    uint32_t i = 1;
    if (act) {
        QItem<V> it = queue[(head + lane) & (QCAP - 1)];
        winv = 1.0 / (double)(it.cnt & ~QFLAG_MASK);
        key = it.key;
        rng.seed(nohash_seed(key));
        if (it.cnt & QFLAG_MASK) {  // first point(s) applied by the caller: keep the stream aligned
            (void)exp01_sample(P.e, rng);
            (void)rng.unif_range(0, P.m, P.slot_thresh);
            i = 2;
            if (it.cnt & QFLAG_SKIP_TWO) {
                (void)exp01_sample(P.e, rng);
                (void)rng.unif_range(0, P.m, P.slot_thresh);
                i = 3;
            }
        }
    }
    uint32_t iter = 0;
    double qmax = load_qmax(S);
    while (__any_sync(0xFFFFFFFFu, act)) {
        if (iter && iter % refresh_period == 0) {  // only when some lane needs more than one point
            refresh_qmax(S, P.m, lane);
            qmax = load_qmax(S);
        }
        ++iter;
        if (act) {
            double base = __dmul_rn(winv, (double)(i - 1));
            if (!(base < qmax)) {
                act = false;
            } else {
                double x = exp01_sample(P.e, rng);
                double h = __dadd_rn(base, __dmul_rn(winv, x));
                if (i == 1 && !(h < qmax)) {
                    act = false;  // first point already above every slot: the item is dead
                } else {
                    uint32_t s = rng.unif_range(0, P.m, P.slot_thresh);
                    sketch_update(S, s, h, (uint64_t)key);
                    ++i;
                    if (!(__dmul_rn(winv, (double)(i - 1)) < qmax)) act = false;
                }
            }
        }
    }
}

// MODE 0: direct histogram of 4^k u8 counters (four per u32) ; MODE 1: open-addressing table
// MEMO  : small key spaces (u32 k-mers, k <= 10): the first point of every pre-key -- exp01 sample,
//         slot and hashed key, all functions of the key only -- comes from a table built once per
//         (k, type, hash, m); otherwise the first point is half-recomputed by first_point_alive().
template <typename V, int MODE, bool MEMO, bool AA>
__global__ void __launch_bounds__(1024, 1) pmh3a_sketch_kernel(const Pmh3aParams P) {
    using TK = typename KmerSource<V, AA>::Task;
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ double s_winv[64];
    using TO = TableOps<V>;
    using Entry = typename TO::Entry;

    Team team;
    team.size = P.team_warps * 32;
    team.id = threadIdx.x / team.size;
    team.tid = threadIdx.x - team.id * team.size;
    team.warp = team.tid >> 5;
    team.lane = threadIdx.x & 31;
    const int nteams = blockDim.x / team.size;

    // ---- carve the team's shared memory ------------------------------------------
    // [region A: histogram / table][slots 16 m][slot hi mirror 4 m][queues][stage 0][stage 1][masks][TeamShared]
    uint8_t* tbase = smem + (size_t)team.id * P.team_smem_bytes;
    uint8_t* regionA = tbase;
    SlotArray S;
    S.slots = (Slot*)(tbase + P.regionA_bytes);
    S.hi = (uint32_t*)(tbase + P.regionA_bytes + (size_t)P.m * 16);
    QItem<V>* queues = (QItem<V>*)(tbase + P.regionA_bytes + P.slots_smem_bytes);
    uint8_t* stage0 = (uint8_t*)queues + (size_t)P.team_warps * QCAP * sizeof(QItem<V>);
    uint8_t* masks = stage0 + 2 * (size_t)P.stage_bytes;
    TeamShared* ts = (TeamShared*)(masks + P.stage_bytes / 2);
    S.s_qmax_hi = &ts->qmax_hi;
    if (P.slots_smem_bytes == 0) {
        uint8_t* g = (uint8_t*)P.slot_scratch + ((size_t)blockIdx.x * nteams + team.id) * ((size_t)P.m * 20);
        S.slots = (Slot*)g;
        S.hi = (uint32_t*)(g + (size_t)P.m * 16);
    }
    QItem<V>* myq = queues + (size_t)team.warp * QCAP;

    if (threadIdx.x < 64) s_winv[threadIdx.x] = threadIdx.x ? 1.0 / (double)threadIdx.x : 0.0;
    // region A starts clean; both passes leave it clean again
    for (uint32_t j = team.tid; j < P.regionA_bytes / 4; j += team.size) ((uint32_t*)regionA)[j] = 0;
    if (team.tid == 0) {
        mbar_init(&ts->mbar[0], 1);
        mbar_init(&ts->mbar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const V header = (V)word_header(P.kmer_type, P.k);
    const bool canonical = hash_is_canonical(P.hash_kind);
    const uint32_t k = P.k;
    const uint32_t refresh_period = P.m <= 2048 ? 1u : P.m / 2048;
    Entry* gtab = nullptr;
    if (MODE == 1 && P.table_scratch)
        gtab = (Entry*)P.table_scratch + ((size_t)blockIdx.x * nteams + team.id) * P.table_scratch_entries;

    // fetch a work item into descriptor `d` and start the TMA copy of its packed bytes
    auto fetch = [&](int d) {
        if (team.tid == 0) {
            const unsigned long long w = atomicAdd(P.work_counter, 1ULL);
            uint32_t valid = w < P.count, staged = 0;
            if (valid) {
                const uint32_t seq = P.order[P.first + w];
                const uint64_t off = P.byte_off[seq];
                const uint32_t L = (uint32_t)P.nbases[seq];
                // bytes the walkers may touch: ceil(L/4) plus one look-ahead word, rounded to 16
                const uint32_t bytes = (((L + 3) >> 2) + 4 + 15) & ~15u;
                ts->seq[d] = seq;
                ts->byte_off[d] = off;
                ts->nbases[d] = L;
                if (bytes <= P.stage_bytes) {
                    staged = 1;
                    mbar_arrive_expect_tx(&ts->mbar[d], bytes);
                    tma_load_bytes(stage0 + (size_t)d * P.stage_bytes, P.packed + off, bytes, &ts->mbar[d]);
                }
            }
            ts->valid[d] = valid;
            ts->staged[d] = staged;
        }
    };
    // next block of 32 tasks from counter `which` for this warp
    auto next_tasks = [&](int which) -> uint32_t {
        uint32_t base = 0;
        if (team.lane == 0) base = atomicAdd(&ts->next_task[which], 32u);
        return __shfl_sync(0xFFFFFFFFu, base, 0);
    };
    // optional phase timing (profiling runs only): team thread 0 accumulates clock deltas
    long long t_mark = 0;
    auto mark = [&](int phase) {
        if (P.phase_clocks && team.tid == 0) {
            long long now = clock64();
            atomicAdd(P.phase_clocks + phase, (unsigned long long)(now - t_mark));
            t_mark = now;
        }
    };

    uint32_t phase0 = 0, phase1 = 0;
    int cur = 0;
    fetch(0);
    team.sync();
    if (P.phase_clocks && team.tid == 0) t_mark = clock64();
    for (;;) {
        if (!ts->valid[cur]) break;
        fetch(cur ^ 1);  // prefetch the next sequence while this one is processed
        const uint32_t seq = ts->seq[cur];
        const uint32_t L = ts->nbases[cur];
        const uint32_t* words = (const uint32_t*)(P.packed + ts->byte_off[cur]);
        // sequences that fit the staging buffer also fit the first-occurrence masks
        const bool use_masks = ts->staged[cur] != 0;
        if (use_masks) {
            words = (const uint32_t*)(stage0 + (size_t)cur * P.stage_bytes);
            if (cur == 0) {
                mbar_wait(&ts->mbar[0], phase0);
                phase0 ^= 1;
            } else {
                mbar_wait(&ts->mbar[1], phase1);
                phase1 ^= 1;
            }
        }
        const uint32_t nk = L >= k ? L - k + 1 : 0;

        for (uint32_t j = team.tid; j < P.m; j += team.size) {
            S.slots[j].hbits = F64_MAX_BITS;
            S.slots[j].key = 0;
            S.hi[j] = (uint32_t)(F64_MAX_BITS >> 32);
        }
        // Speculative start value of qmax.  The slot minima behave like Exp(nk / m) variables, so
        // all of them end below B = (m / nk) ln(m / 1e-4) except with probability ~1e-4; pruning
        // against min(B, actual) from the first item on is exactly as safe as pruning against the
        // actual qmax PROVIDED the final maximum is below B -- which is verified before the
        // signature is accepted (a failed sequence is redone without speculation).
        uint32_t spec_hi = (uint32_t)(F64_MAX_BITS >> 32);
        if (P.speculate && nk) {
            const double B = P.spec_factor / (double)nk;
            const uint32_t bhi = (uint32_t)__double2hiint(B);
            if (bhi < spec_hi) spec_hi = bhi;
        }
        if (team.tid == 0) {
            ts->qmax_hi = spec_hi;
            ts->flag = 0;
            ts->next_task[0] = 0;
            ts->next_task[1] = 0;
        }
        // positions per task: 16 (one word), or 8 when there are too few words to occupy the team
        const uint32_t log2T = nk < (uint32_t)team.size * 16 ? 3 : 4;
        const uint32_t T = 1u << log2T;
        const uint32_t ntasks = (nk + T - 1) >> log2T;

        // table geometry for this sequence
        Entry* tab = (Entry*)regionA;
        uint32_t log2cap = 0, capmask = 0;
        if (MODE == 1) {
            uint64_t want = nk < 32 ? 64 : (uint64_t)nk * 2;
            log2cap = 64 - __clzll((long long)(want - 1));
            uint64_t cap = 1ULL << log2cap;
            if (cap * sizeof(Entry) > P.regionA_bytes) tab = gtab;  // long sequence: L2 / HBM scratch
            capmask = (uint32_t)(cap - 1);
        }
        team.sync();
        mark(0);  // fetch + wait for the staged bytes + slot init

        // ---------------- pass 1 : multiplicities + first-occurrence masks -------------------
        const bool may_wrap = nk > 0xFFu;
        for (;;) {
            const uint32_t base = next_tasks(0);
            if (base >= ntasks) break;
            const uint32_t task = base + team.lane;
            if (task < ntasks) {
                TK tk;
                uint32_t p = task << log2T;
                tk.init(words, p, k);
                const uint32_t pend = min(p + T, nk);
                uint32_t first = 0;
#pragma unroll 1
                for (uint32_t t = 0; p < pend; ++t, ++p) {
                    V pk = tk.get(t, canonical);
                    if (MODE == 0) {
                        const uint32_t sh = ((uint32_t)pk & 3u) * 8;
                        const uint32_t old = atomicAdd((uint32_t*)regionA + ((uint32_t)pk >> 2), 1u << sh);
                        const uint32_t c = (old >> sh) & 0xFFu;
                        first |= (uint32_t)(c == 0) << t;
                        if (may_wrap && c == 0xFFu) ts->flag = 1;  // u8 counter wrapped
                    } else {
                        first |= (uint32_t)TO::insert(tab, capmask, log2cap, pk) << t;
                    }
                }
                if (use_masks) {
                    if (log2T == 4) ((uint16_t*)masks)[task] = (uint16_t)first;
                    else masks[task] = (uint8_t)first;
                }
            }
        }
        team.sync();
        mark(1);  // pass 1
        const bool overflow = MODE == 0 && ts->flag != 0;

        // ---------------- pass 2 : owners sketch their distinct items -------------------------
        // Every owner reads its item's multiplicity, clears the counter and runs the first-point
        // filter; items that are still alive are compacted into the warp queue and processed on
        // full warps.  qmax starts from the speculative bound set up above instead of +inf.
        uint32_t qhead = 0, qtail = 0;
        for (;;) {
            const uint32_t base = next_tasks(1);
            if (base >= ntasks) break;
            const uint32_t task = base + team.lane;
            const bool tact = task < ntasks;
            TK tk;
            uint32_t p = task << log2T;
            uint32_t own = 0xFFFFu;
            if (tact) {
                tk.init(words, p, k);
                if (use_masks) own = log2T == 4 ? (uint32_t)((const uint16_t*)masks)[task] : (uint32_t)masks[task];
            }
            refresh_qmax(S, P.m, team.lane);
            double qmax = load_qmax(S);
            bool filter_off = qmax >= 0.75;
#pragma unroll 1
            for (uint32_t t = 0; t < T; ++t, ++p) {
                uint32_t cnt = 0;
                V pk = 0;
                const bool in_range = tact && p < nk;
                const bool mine = in_range && ((own >> t) & 1u);
                if (TK::SEQUENTIAL ? in_range : mine) pk = tk.get(t, canonical);  // rolling walkers must see every position
                if (mine) {
                    if (MODE == 0) {
                        if (overflow) {
                            ((uint32_t*)regionA)[(uint32_t)pk >> 2] = 0;
                        } else if (use_masks) {
                            cnt = regionA[(uint32_t)pk];
                            regionA[(uint32_t)pk] = 0;
                        } else {
                            const uint32_t sh = ((uint32_t)pk & 3u) * 8;
                            const uint32_t old = atomicAnd((uint32_t*)regionA + ((uint32_t)pk >> 2), ~(0xFFu << sh));
                            cnt = (old >> sh) & 0xFFu;
                        }
                    } else {
                        cnt = use_masks ? TO::lookup(tab, capmask, log2cap, pk) : TO::claim(tab, capmask, log2cap, pk);
                    }
                }
                V key = 0;
                if (cnt) {
                    if (MEMO) {
                        // first and second point straight from the per-key tables; only an item that may
                        // need a third point (2 winv < qmax) goes to the queue
                        const uint4 e = __ldg((const uint4*)P.memo_fast + 2 * (uint32_t)pk);
                        const double winv = cnt < 64 ? s_winv[cnt] : 1.0 / (double)cnt;
                        const bool second = winv < qmax;
                        uint4 e2 = make_uint4(0, 0, 0, 0);
                        if (second) e2 = __ldg((const uint4*)P.memo_fast + 2 * (uint32_t)pk + 1);
                        const double h = __dmul_rn(winv, __hiloint2double((int)e.y, (int)e.x));
                        key = (V)e.w;
                        if (h < qmax) sketch_update(S, e.z, h, (uint64_t)e.w);
                        if (second && cnt < QFLAG_SKIP_TWO) {
                            const double h2 = __dadd_rn(winv, __dmul_rn(winv, __hiloint2double((int)e2.y, (int)e2.x)));
                            if (h2 < qmax) sketch_update(S, e2.z, h2, (uint64_t)e.w);
                            cnt = __dmul_rn(winv, 2.0) < qmax ? (cnt | QFLAG_SKIP_TWO) : 0u;
                        } else if (!second) {
                            cnt = 0;
                        }
                    } else {
                        key = finalize_key<V>(pk, header, P.hash_kind);
                        // while qmax is large nearly every item survives its first point: skip the filter
                        if (!filter_off) {
                            const double winv = cnt < 64 ? s_winv[cnt] : 1.0 / (double)cnt;
                            if (!first_point_alive<V>(key, winv, qmax, P.e.c1)) cnt = 0;
                        }
                    }
                }
                const uint32_t bal = __ballot_sync(0xFFFFFFFFu, cnt != 0);
                if (bal) {
                    if (cnt) {
                        uint32_t pos = qtail + __popc(bal & ((1u << team.lane) - 1));
                        QItem<V> it;
                        it.key = key;
                        it.cnt = cnt;
                        myq[pos & (QCAP - 1)] = it;
                    }
                    qtail += __popc(bal);
                    __syncwarp();
                    if (qtail - qhead >= 32) {
                        process_items<V>(myq, qhead, 32, team.lane, P, S, refresh_period);
                        qhead += 32;
                        __syncwarp();
                        qmax = load_qmax(S);
                        filter_off = qmax >= 0.75;
                    }
                }
            }
        }
        if (qtail != qhead) process_items<V>(myq, qhead, qtail - qhead, team.lane, P, S, refresh_period);
        team.sync();
        mark(3);  // pass 2

        // ---------------- signature out, leave region A clean --------------------------
        if (!overflow) {
            V* out = (V*)P.sig + (size_t)seq * P.m;
            bool bad = false;
            for (uint32_t j = team.tid; j < P.m; j += team.size) {
                out[j] = (V)S.slots[j].key;
                bad |= (uint32_t)(S.slots[j].hbits >> 32) > spec_hi;  // a slot ended above the speculative bound
            }
            if (bad) ts->flag = 2;
        }
        if (MODE == 1) {
            const uint32_t nvec = (capmask + 1) / (16 / sizeof(Entry));  // 16-byte vectors
            for (uint32_t j = team.tid; j < nvec; j += team.size) ((uint4*)tab)[j] = make_uint4(0, 0, 0, 0);
        }
        cur ^= 1;
        team.sync();  // descriptor `cur` was written before the passes; staging buffer cur^1 is free again
        if (team.tid == 0 && ts->flag != 0) {  // u8 counter wrapped or speculation failed: redo this sequence
            unsigned long long pos = atomicAdd(P.overflow_count, 1ULL);
            P.overflow_list[pos] = seq;
        }
        mark(4);  // signature out + table clear
    }
}

// first two points of every pre-key, one 32-byte sector per key: fast[2 pk] = {x1 bits lo, x1 bits hi, slot1, hashed key},
// fast[2 pk + 1] = {x2 bits lo, x2 bits hi, slot2, hashed key}
__global__ void pmh3a_memo_kernel(uint4* fast, uint32_t nkeys, Pmh3aParams P) {
    const uint32_t header = word_header(P.kmer_type, P.k);
    for (uint32_t pk = blockIdx.x * blockDim.x + threadIdx.x; pk < nkeys; pk += gridDim.x * blockDim.x) {
        const uint32_t key = finalize_key<uint32_t>(pk, header, P.hash_kind);
        Xoshiro256pp rng;
        rng.seed(nohash_seed(key));
        const double x = exp01_sample(P.e, rng);
        const uint32_t s = rng.unif_range(0, P.m, P.slot_thresh);
        fast[2 * pk] = make_uint4((uint32_t)__double2loint(x), (uint32_t)__double2hiint(x), s, key);
        const double x2 = exp01_sample(P.e, rng);
        const uint32_t s2 = rng.unif_range(0, P.m, P.slot_thresh);
        fast[2 * pk + 1] = make_uint4((uint32_t)__double2loint(x2), (uint32_t)__double2hiint(x2), s2, key);
    }
}

cudaError_t launch_pmh3a_memo(const Pmh3aParams& P, void* fast, uint32_t nkeys, cudaStream_t stream) {
    int block = 256;
    int grid = (int)((nkeys + block - 1) / block);
    if (grid > 148 * 8) grid = 148 * 8;
    pmh3a_memo_kernel<<<grid, block, 0, stream>>>((uint4*)fast, nkeys, P);
    return cudaGetLastError();
}

// --------------------------------------------------------------------------------
// host-side launcher
// --------------------------------------------------------------------------------
template <typename V, int MODE, bool MEMO, bool AA = false>
static cudaError_t launch_one(const Pmh3aParams& P, int grid, int block, size_t smem, cudaStream_t stream) {
    auto kern = pmh3a_sketch_kernel<V, MODE, MEMO, AA>;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    kern<<<grid, block, smem, stream>>>(P);
    return cudaGetLastError();
}

size_t pmh3a_qitem_bytes(bool key64) { return key64 ? sizeof(QItem<uint64_t>) : sizeof(QItem<uint32_t>); }
size_t pmh3a_entry_bytes(bool key64) {
    return key64 ? sizeof(TableOps<uint64_t>::Entry) : sizeof(TableOps<uint32_t>::Entry);
}

cudaError_t launch_pmh3a(const Pmh3aParams& P, bool key64, int mode, int grid, int block, size_t smem,
                         cudaStream_t stream) {
    if (P.kmer_type == KMU_KMERAA32) return launch_one<uint32_t, 1, false, true>(P, grid, block, smem, stream);
    if (P.kmer_type == KMU_KMERAA64) return launch_one<uint64_t, 1, false, true>(P, grid, block, smem, stream);
    if (key64) {
        return mode == 0 ? launch_one<uint64_t, 0, false>(P, grid, block, smem, stream)
                         : launch_one<uint64_t, 1, false>(P, grid, block, smem, stream);
    }
    if (P.memo_fast)
        return mode == 0 ? launch_one<uint32_t, 0, true>(P, grid, block, smem, stream)
                         : launch_one<uint32_t, 1, true>(P, grid, block, smem, stream);
    return mode == 0 ? launch_one<uint32_t, 0, false>(P, grid, block, smem, stream)
                     : launch_one<uint32_t, 1, false>(P, grid, block, smem, stream);
}

}  // namespace kmu
