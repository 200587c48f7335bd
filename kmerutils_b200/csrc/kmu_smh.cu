// kmu_smh.cu -- SuperMinHash sketching on sm_100a.
//
// Replaces SeqSketcher::sketch_superminhash (src/sketching/seqsketchjaccard.rs:328-380, FnvHasher on the
// key), SuperHashSketch::sketch_compressedkmer / sketch_compressedkmer_seqs
// (src/sketching/setsketchert.rs:255-335, NoHashHasher) and sketch_seqrange_superminhash
// (src/sketching/seqminhash.rs:19-62): every k-mer is streamed through probminhash's
// SuperMinHash::sketch (Ertl 2017).
//
// GPU formulation.  Item x draws (r_j, k_j), j = 0, 1, ... from its private Xoshiro256++ stream, runs
// an incremental Fisher-Yates shuffle (slot_j = p[j] after swapping p[j], p[k_j]) and offers the value
// r_j + j to slot_j; the signature is the per-slot MINIMUM over all items and all j.  The sequential
// algorithm only cuts an item's loop at j > a, where a = floor(max slot value): those points cannot
// win.  Hence
//   * fast path: every item evaluates j = 0 .. a_spec for a speculative a_spec derived from the k-mer
//     count; the result is exact iff all slots end below a_spec + 1, which is verified -- sequences
//     that fail are redone by the exact path.  For a_spec = 0 (the usual case: many more k-mers than
//     slots) an item is one seed, two draws and one atomicMin in shared memory.
//   * exact path (short sequences relative to m): one warp per sequence, 32 items per round with a
//     full lazily-reset permutation per lane in global scratch, a = floor(max slot) between rounds.
#include <algorithm>
#include <cstdint>

#include "kmu_device.cuh"
#include "kmu_kernels.h"

namespace kmu {

template <typename S>
struct FloatOps;
template <>
struct FloatOps<double> {
    using B = unsigned long long;
    static __device__ __forceinline__ double draw(Xoshiro256pp& rng) { return rng.unif01(); }
    static __device__ __forceinline__ B bits(double v) { return (B)__double_as_longlong(v); }
    static __device__ __forceinline__ double value(B b) { return __longlong_as_double((long long)b); }
    static __device__ __forceinline__ double add(double r, uint32_t j) { return __dadd_rn(r, (double)j); }
    static __device__ __forceinline__ B large() { return bits(4294967295.0); }  // F::from(u32::MAX)
};
template <>
struct FloatOps<float> {
    using B = unsigned int;
    static __device__ __forceinline__ float draw(Xoshiro256pp& rng) { return rng.unif01_f32(); }
    static __device__ __forceinline__ B bits(float v) { return __float_as_uint(v); }
    static __device__ __forceinline__ float value(B b) { return __uint_as_float(b); }
    static __device__ __forceinline__ float add(float r, uint32_t j) { return __fadd_rn(r, (float)j); }
    static __device__ __forceinline__ B large() { return bits(4294967296.0f); }  // u32::MAX rounds up in f32
};

template <typename V>
__device__ __forceinline__ uint64_t item_seed(V key, int hasher) {
    return hasher == 0 ? nohash_seed(key) : fnv1a_seed<V>(key);
}

// Uniform::<usize>::new(low, low + range).sample (rand 0.9, 32-bit path)
__device__ __forceinline__ uint32_t unif_from(Xoshiro256pp& rng, uint32_t low, uint32_t range) {
    const uint32_t thresh = (0u - range) % range;
    return rng.unif_range(low, range, thresh);
}

constexpr uint32_t SMH_FAST_MAX = 15;  // largest a_spec of the fast path (sparse permutation of <= 16 entries)

// points j = 0 .. a_spec of one item against the slots h[0..m)
template <typename S>
__device__ __forceinline__ void smh_item_points(Xoshiro256pp& rng, uint32_t m, uint32_t a_spec,
                                                typename FloatOps<S>::B* h) {
    using F = FloatOps<S>;
    using B = typename F::B;
    if (a_spec == 0) {
        const S r = F::draw(rng);
        const uint32_t slot = unif_from(rng, 0, m);
        const B vb = F::bits(r);
        if (vb < *(volatile B*)(h + slot)) atomicMin(h + slot, vb);
        return;
    }
    uint32_t idx[SMH_FAST_MAX + 1], val[SMH_FAST_MAX + 1];
    uint32_t n = 0;
    for (uint32_t j = 0; j <= a_spec; ++j) {
        const S r = F::draw(rng);
        const uint32_t kk = unif_from(rng, j, m - j);
        uint32_t vj = j, vk = kk;
        int pos_k = -1;
        for (uint32_t e = 0; e < n; ++e) {
            if (idx[e] == j) vj = val[e];
            if (idx[e] == kk) {
                vk = val[e];
                pos_k = (int)e;
            }
        }
        uint32_t slot = vj;
        if (kk != j) {
            slot = vk;  // p[j] after the swap is the old p[k]
            if (pos_k >= 0) {
                val[pos_k] = vj;
            } else {
                idx[n] = kk;
                val[n] = vj;
                ++n;
            }
        }
        const B vb = F::bits(F::add(r, j));
        if (vb < *(volatile B*)(h + slot)) atomicMin(h + slot, vb);
    }
}

struct SmhTeamShared {
    uint32_t seq, nbases, valid, flag;
    uint64_t byte_off;
};

template <typename V, typename S, bool AA>
__global__ void __launch_bounds__(1024, 1) smh_fast_kernel(const SmhParams P) {
    using TK = typename KmerSource<V, AA>::Task;
    using F = FloatOps<S>;
    using B = typename F::B;
    extern __shared__ __align__(16) uint8_t smem[];
    Team team;
    team.size = P.team_warps * 32;
    team.id = threadIdx.x / team.size;
    team.tid = threadIdx.x - team.id * team.size;
    team.warp = team.tid >> 5;
    team.lane = threadIdx.x & 31;
    uint8_t* tbase = smem + (size_t)team.id * P.team_smem_bytes;
    B* h = (B*)tbase;
    SmhTeamShared* ts = (SmhTeamShared*)(tbase + (((size_t)P.m * sizeof(B) + 15) & ~(size_t)15));
    const V header = (V)word_header(P.kmer_type, P.k);
    const bool canonical = hash_is_canonical(P.hash_kind);
    const uint32_t k = P.k, m = P.m;

    for (;;) {
        team.sync();  // the previous sequence's shared state is no longer read
        if (team.tid == 0) {
            const unsigned long long w = atomicAdd(P.work_counter, 1ULL);
            ts->valid = w < P.count;
            if (w < P.count) {
                const uint32_t seq = P.order[P.first + w];
                ts->seq = seq;
                ts->nbases = (uint32_t)P.nbases[seq];
                ts->byte_off = P.byte_off[seq];
            }
            ts->flag = 0;
        }
        team.sync();
        if (!ts->valid) break;
        const uint32_t seq = ts->seq, L = ts->nbases;
        const uint32_t* words = (const uint32_t*)(P.packed + ts->byte_off);
        const uint32_t nk = L >= k ? L - k + 1 : 0;
        // speculative loop bound: with a_spec + 1 points per item every slot is hit except with
        // probability ~1e-4 when (a_spec + 1) nk >= m ln(1e4 m)
        uint32_t a_spec = 0;
        if (nk) {
            const double need = (double)m / (double)nk * P.ln_term;
            const uint32_t a1 = need >= (double)m ? m : (uint32_t)ceil(need);
            a_spec = (a1 < 1 ? 1 : a1) - 1;
        }
        // value cut of the one-point path: with D >= nk / 4 distinct items every slot ends below 4 m ln(1e4 m) / nk except
        // with probability ~1e-4 (1 when the sequence is too short for it to help)
        S cut = (S)1;
        bool bound_cut = false;
        if (nk && P.value_cut) {
            const double c = 4.0 * (double)m / (double)nk * P.ln_term;
            if (c < 0.9) cut = (S)c;
        }
        if (a_spec > SMH_FAST_MAX) {  // short sequence: exact path
            if (team.tid == 0) P.slow_list[atomicAdd(P.slow_count, 1ULL)] = seq;
            continue;
        }
        for (uint32_t j = team.tid; j < m; j += team.size) h[j] = F::large();
        team.sync();
        if (!AA && sizeof(V) == 4 && P.memo && a_spec == 0) {
            // one point per item, straight from the per-key table: eight L2 lookups in flight per thread
            const uint4* memo = (const uint4*)P.memo;
            const uint32_t ntasks8 = (nk + 7) >> 3;
            for (uint32_t task = team.tid; task < ntasks8; task += team.size) {
                const uint32_t p0 = task << 3, nv = min(8u, nk - p0);
                TaskKmers<uint32_t> tk8;
                tk8.init(words, p0, k);
                uint4 e[8];
#pragma unroll
                for (uint32_t t = 0; t < 8; ++t)
                    if (t < nv) e[t] = __ldcg(memo + tk8.get(t, canonical));
#pragma unroll
                for (uint32_t t = 0; t < 8; ++t) {
                    if (t < nv) {
                        const B vb = sizeof(B) == 8 ? (B)(((unsigned long long)e[t].y << 32) | e[t].x) : (B)e[t].x;
                        if (vb < *(volatile B*)(h + e[t].z)) atomicMin(h + e[t].z, vb);
                    }
                }
            }
        } else if (a_spec == 0 && cut < (S)1) {
            // Long sequence, one point per item, value = the FIRST output of the item's generator -- which needs only s0
            // and s3 (SplitMix64 outputs 1 and 4 of the seed).  Every slot ends below `cut` (verified below, like the
            // a_spec bound), so an item whose value is not below it is dropped after half a seeding; the others are
            // queued per warp and finished 32 at a time with the full generator.
            const uint32_t ntasks = (nk + 15) >> 4;
            V* wq = (V*)(smem + (size_t)P.team_smem_bytes * (blockDim.x / team.size)) + (threadIdx.x >> 5) * 64;
            uint32_t qn = 0;  // warp-uniform
            for (uint32_t task0 = 0; task0 < ntasks; task0 += team.size) {  // the same trip count for every lane of a warp
                const uint32_t task = task0 + team.tid;
                TK tk;
                uint32_t p = task << 4, pend = p;
                if (task < ntasks) {
                    tk.init(words, p, k);
                    pend = min(p + 16, nk);
                }
#pragma unroll 1
                for (uint32_t t = 0; t < 16; ++t, ++p) {
                    bool alive = false;
                    V key = 0;
                    if (p < pend) {
                        key = finalize_key<V>(tk.get(t, canonical), header, P.hash_kind);
                        const uint64_t sd = item_seed<V>(key, P.hasher);
                        uint64_t x0 = sd, x3 = sd + 3ULL * 0x9E3779B97F4A7C15ULL;
                        const uint64_t s0 = Xoshiro256pp::splitmix(x0), s3 = Xoshiro256pp::splitmix(x3);
                        const uint64_t r0 = rotl64(s0 + s3, 23) + s0;
                        S r;
                        if (sizeof(S) == 8) r = (S)(__longlong_as_double((long long)((r0 >> 12) | 0x3FF0000000000000ULL)) - 1.0);
                        else r = (S)(__uint_as_float(((uint32_t)(r0 >> 32) >> 9) | 0x3F800000u) - 1.0f);
                        alive = r < cut;
                    }
                    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, alive);
                    if (alive) wq[qn + __popc(bal & ((1u << team.lane) - 1u))] = key;
                    qn += __popc(bal);
                    __syncwarp();
                    if (qn >= 32) {
                        qn -= 32;
                        const V kq = wq[qn + team.lane];
                        __syncwarp();
                        Xoshiro256pp rng;
                        rng.seed(item_seed<V>(kq, P.hasher));
                        smh_item_points<S>(rng, m, 0, h);
                    }
                }
            }
            if ((uint32_t)team.lane < qn) {
                Xoshiro256pp rng;
                rng.seed(item_seed<V>(wq[team.lane], P.hasher));
                smh_item_points<S>(rng, m, 0, h);
            }
            __syncwarp();
            bound_cut = true;
        } else {
        const uint32_t ntasks = (nk + 15) >> 4;
        for (uint32_t task = team.tid; task < ntasks; task += team.size) {
            TK tk;
            uint32_t p = task << 4;
            tk.init(words, p, k);
            const uint32_t pend = min(p + 16, nk);
#pragma unroll 1
            for (uint32_t t = 0; p < pend; ++t, ++p) {
                const V key = finalize_key<V>(tk.get(t, canonical), header, P.hash_kind);
                Xoshiro256pp rng;
                rng.seed(item_seed<V>(key, P.hasher));
                smh_item_points<S>(rng, m, a_spec, h);
            }
        }
        }
        team.sync();
        S* out = (S*)P.sig + (size_t)seq * m;
        const S bound = bound_cut ? cut : (S)(a_spec + 1);
        bool bad = false;
        for (uint32_t j = team.tid; j < m; j += team.size) {
            const S v = F::value(h[j]);
            out[j] = v;
            bad |= !(v < bound);
        }
        if (bad && nk) ts->flag = 1;
        team.sync();
        if (team.tid == 0 && ts->flag) P.slow_list[atomicAdd(P.slow_count, 1ULL)] = seq;  // speculation failed
    }
}

// exact path: one warp per sequence
template <typename V, typename S, bool AA>
__global__ void __launch_bounds__(32) smh_exact_kernel(const SmhParams P) {
    using Walker = typename KmerSource<V, AA>::Walker;
    using F = FloatOps<S>;
    using B = typename F::B;
    const int lane = threadIdx.x;
    const uint32_t k = P.k, m = P.m;
    const V header = (V)word_header(P.kmer_type, P.k);
    const bool canonical = hash_is_canonical(P.hash_kind);
    // scratch of this warp: h[m] (B), then per lane p[m], q[m] (u32)
    uint8_t* base = P.scratch + (size_t)blockIdx.x * P.scratch_per_warp;
    B* h = (B*)base;
    uint32_t* pl = (uint32_t*)(base + (((size_t)m * sizeof(B) + 15) & ~(size_t)15)) + (size_t)lane * 2 * m;
    uint32_t* ql = pl + m;
    uint32_t stamp = 0;  // q is zeroed before the launch; stamps only grow
    for (;;) {
        unsigned long long w = 0;
        if (lane == 0) w = atomicAdd(P.work_counter, 1ULL);
        w = __shfl_sync(0xFFFFFFFFu, w, 0);
        if (w >= P.count) break;
        const uint32_t seq = P.order[P.first + w];
        const uint32_t L = (uint32_t)P.nbases[seq];
        const uint32_t* words = (const uint32_t*)(P.packed + P.byte_off[seq]);
        const uint32_t nk = L >= k ? L - k + 1 : 0;
        for (uint32_t j = lane; j < m; j += 32) h[j] = F::large();
        __syncwarp();
        uint32_t a_cur = m - 1;
        for (uint32_t p0 = 0; p0 < nk; p0 += 32) {
            const uint32_t pos = p0 + lane;
            if (pos < nk) {
                Walker wk;
                wk.start(words, pos, k);
                wk.roll();
                const V key = finalize_key<V>(wk.prekey(canonical), header, P.hash_kind);
                Xoshiro256pp rng;
                rng.seed(item_seed<V>(key, P.hasher));
                ++stamp;
                for (uint32_t j = 0; j <= a_cur; ++j) {
                    const S r = F::draw(rng);
                    const uint32_t kk = unif_from(rng, j, m - j);
                    if (ql[j] != stamp) {
                        ql[j] = stamp;
                        pl[j] = j;
                    }
                    if (ql[kk] != stamp) {
                        ql[kk] = stamp;
                        pl[kk] = kk;
                    }
                    const uint32_t t = pl[j];
                    pl[j] = pl[kk];
                    pl[kk] = t;
                    const uint32_t slot = pl[j];
                    const B vb = F::bits(F::add(r, j));
                    if (vb < *(volatile B*)(h + slot)) atomicMin(h + slot, vb);
                }
            } else {
                ++stamp;
            }
            __syncwarp();
            // a = floor(max slot value), capped at m - 1 (the sequential a_upper)
            B mx = 0;
            for (uint32_t j = lane; j < m; j += 32) {
                const B v = *(volatile B*)(h + j);
                mx = v > mx ? v : mx;
            }
            for (int o = 16; o; o >>= 1) {
                const B other = __shfl_xor_sync(0xFFFFFFFFu, mx, o);
                mx = other > mx ? other : mx;
            }
            const double top = (double)F::value(mx);
            a_cur = top >= (double)(m - 1) ? m - 1 : (uint32_t)top;
            __syncwarp();
        }
        S* out = (S*)P.sig + (size_t)seq * m;
        for (uint32_t j = lane; j < m; j += 32) out[j] = F::value(h[j]);
        __syncwarp();
    }
}

// ---- one sketch for the whole batch (SuperHashSketch::sketch_compressedkmer_seqs, setsketchert.rs:299-335) --------
// every CTA keeps partial slots in shared memory over its share of the 64-byte chunks of the packed buffer and merges
// them into the global slots with atomicMin; the host verifies the speculative bound (all slots < a_spec + 1)
template <typename V, typename S, bool AA>
__global__ void __launch_bounds__(512, 2) smh_whole_kernel(const SmhParams P, SeqView b, uint64_t total_bytes, uint32_t a_spec,
                                                            typename FloatOps<S>::B* gslots) {
    using F = FloatOps<S>;
    using B = typename F::B;
    extern __shared__ __align__(16) uint8_t smem[];
    B* h = (B*)smem;
    const uint32_t m = P.m;
    for (uint32_t j = threadIdx.x; j < m; j += blockDim.x) h[j] = F::large();
    __syncthreads();
    const V header = (V)word_header(P.kmer_type, P.k);
    const bool canonical = hash_is_canonical(P.hash_kind);
    const uint64_t nchunks = (total_bytes + CHUNK_BYTES - 1) / CHUNK_BYTES;
    for (uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; c < nchunks; c += (uint64_t)gridDim.x * blockDim.x)
        for_each_kmer_in_chunk<V, AA>(b, total_bytes, c, P.k, canonical, [&](V pk) {
            const V key = finalize_key<V>(pk, header, P.hash_kind);
            Xoshiro256pp rng;
            rng.seed(item_seed<V>(key, P.hasher));
            smh_item_points<S>(rng, m, a_spec, h);
        });
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < m; j += blockDim.x)
        if (h[j] != F::large()) atomicMin(gslots + j, h[j]);
}

// DNA, long inputs (one point per item): warp-cooperative form of the whole-batch kernel with the value cut of
// smh_fast_kernel -- the lanes take 128 consecutive positions of the warp's slice at a time (four per lane), half a seeding gives an item's
// value, an item not below `cut` is dropped there, the others are queued per warp and finished 32 at a time.  The host
// verifies that every merged slot ends below the cut.
constexpr uint32_t SMH_WHOLE_WARPS = 16;
template <typename V, typename S>
__global__ void __launch_bounds__(32 * SMH_WHOLE_WARPS, 2) smh_whole_warp_kernel(const SmhParams P, SeqView b, uint64_t total_bytes, S cut,
                                                                                   typename FloatOps<S>::B* gslots) {
    using F = FloatOps<S>;
    using B = typename F::B;
    extern __shared__ __align__(16) uint8_t smem[];
    B* h = (B*)smem;
    const uint32_t m = P.m;
    V* wq = (V*)(smem + (((size_t)m * sizeof(B) + 15) & ~(size_t)15)) + (threadIdx.x >> 5) * 64;
    for (uint32_t j = threadIdx.x; j < m; j += blockDim.x) h[j] = F::large();
    __syncthreads();
    const V header = (V)word_header(P.kmer_type, P.k);
    const bool canonical = hash_is_canonical(P.hash_kind);
    const int lane = threadIdx.x & 31;
    const uint64_t ngroups = (total_bytes + GROUP_BYTES - 1) / GROUP_BYTES;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    uint32_t qn = 0;  // warp-uniform
    for (uint64_t g = warp; g < ngroups; g += nwarps)
        warp_for_each_kmer<V>(b, total_bytes, g, P.k, canonical, lane, [&](V pk, bool active) {
            const V key = finalize_key<V>(pk, header, P.hash_kind);
            bool alive = false;
            if (active) {
                const uint64_t sd = item_seed<V>(key, P.hasher);
                uint64_t x0 = sd, x3 = sd + 3ULL * 0x9E3779B97F4A7C15ULL;
                const uint64_t s0 = Xoshiro256pp::splitmix(x0), s3 = Xoshiro256pp::splitmix(x3);
                const uint64_t r0 = rotl64(s0 + s3, 23) + s0;
                S r;
                if (sizeof(S) == 8) r = (S)(__longlong_as_double((long long)((r0 >> 12) | 0x3FF0000000000000ULL)) - 1.0);
                else r = (S)(__uint_as_float(((uint32_t)(r0 >> 32) >> 9) | 0x3F800000u) - 1.0f);
                alive = r < cut;
            }
            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, alive);
            if (alive) wq[qn + __popc(bal & ((1u << lane) - 1u))] = key;
            qn += __popc(bal);
            __syncwarp();
            if (qn >= 32) {
                qn -= 32;
                const V kq = wq[qn + lane];
                __syncwarp();
                Xoshiro256pp rng;
                rng.seed(item_seed<V>(kq, P.hasher));
                smh_item_points<S>(rng, m, 0, h);
            }
        });
    if ((uint32_t)lane < qn) {
        Xoshiro256pp rng;
        rng.seed(item_seed<V>(wq[lane], P.hasher));
        smh_item_points<S>(rng, m, 0, h);
    }
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < m; j += blockDim.x)
        if (h[j] != F::large()) atomicMin(gslots + j, h[j]);
}

template <typename V, typename S>
static cudaError_t launch_whole_cut_t(const SmhParams& P, const SeqView& b, uint64_t total_bytes, double cut, void* gslots, int sm_count,
                                      cudaStream_t st) {
    auto kern = smh_whole_warp_kernel<V, S>;
    const size_t smem = (((size_t)P.m * sizeof(S) + 15) & ~(size_t)15) + SMH_WHOLE_WARPS * 64 * sizeof(V);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    const uint64_t ngroups = (total_bytes + GROUP_BYTES - 1) / GROUP_BYTES;
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((ngroups + SMH_WHOLE_WARPS - 1) / SMH_WHOLE_WARPS, (uint64_t)sm_count * 2));
    kern<<<grid, 32 * SMH_WHOLE_WARPS, smem, st>>>(P, b, total_bytes, (S)cut, (typename FloatOps<S>::B*)gslots);
    return cudaGetLastError();
}
// DNA batches only; `cut` as the kernel will compare it (already rounded to S by the caller for the verification)
cudaError_t launch_smh_whole_cut(const SmhParams& P, bool key64, bool f64, const SeqView& b, uint64_t total_bytes, double cut,
                                 void* gslots, int sm_count, cudaStream_t st) {
    if (key64) return f64 ? launch_whole_cut_t<uint64_t, double>(P, b, total_bytes, cut, gslots, sm_count, st)
                          : launch_whole_cut_t<uint64_t, float>(P, b, total_bytes, cut, gslots, sm_count, st);
    return f64 ? launch_whole_cut_t<uint32_t, double>(P, b, total_bytes, cut, gslots, sm_count, st)
               : launch_whole_cut_t<uint32_t, float>(P, b, total_bytes, cut, gslots, sm_count, st);
}

template <typename B>
__global__ void smh_fill_kernel(B* slots, uint32_t m, B v) {
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) slots[j] = v;
}

// column minimum of nseq rows of m slots (element-wise merge of per-sequence sketches)
template <typename B>
__global__ void smh_colmin_kernel(const B* rows, uint64_t nseq, uint32_t m, B* out) {
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) {
        B mn = ~(B)0;
        for (uint64_t r = blockIdx.y; r < nseq; r += gridDim.y) {
            const B v = rows[r * m + j];
            mn = v < mn ? v : mn;
        }
        atomicMin(out + j, mn);
    }
}

template <typename V, typename S, bool AA>
static cudaError_t launch_whole_t(const SmhParams& P, const SeqView& b, uint64_t total_bytes, uint32_t a_spec, void* gslots,
                                  int grid, size_t smem, cudaStream_t st) {
    auto kern = smh_whole_kernel<V, S, AA>;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    kern<<<grid, 512, smem, st>>>(P, b, total_bytes, a_spec, (typename FloatOps<S>::B*)gslots);
    return cudaGetLastError();
}
template <typename V, bool AA>
static cudaError_t launch_whole_s(const SmhParams& P, bool f64, const SeqView& b, uint64_t total_bytes, uint32_t a_spec,
                                  void* gslots, int grid, size_t smem, cudaStream_t st) {
    return f64 ? launch_whole_t<V, double, AA>(P, b, total_bytes, a_spec, gslots, grid, smem, st)
               : launch_whole_t<V, float, AA>(P, b, total_bytes, a_spec, gslots, grid, smem, st);
}
cudaError_t launch_smh_whole(const SmhParams& P, bool key64, bool f64, const SeqView& b, uint64_t total_bytes, uint32_t a_spec,
                             void* gslots, int grid, size_t smem, cudaStream_t st) {
    if (P.kmer_type == KMU_KMERAA32) return launch_whole_s<uint32_t, true>(P, f64, b, total_bytes, a_spec, gslots, grid, smem, st);
    if (P.kmer_type == KMU_KMERAA64) return launch_whole_s<uint64_t, true>(P, f64, b, total_bytes, a_spec, gslots, grid, smem, st);
    return key64 ? launch_whole_s<uint64_t, false>(P, f64, b, total_bytes, a_spec, gslots, grid, smem, st)
                 : launch_whole_s<uint32_t, false>(P, f64, b, total_bytes, a_spec, gslots, grid, smem, st);
}
// slots <- F::from(u32::MAX) as bit patterns
cudaError_t launch_smh_fill_large(void* slots, uint32_t m, bool f64, cudaStream_t st) {
    if (f64) smh_fill_kernel<unsigned long long><<<(m + 255) / 256, 256, 0, st>>>((unsigned long long*)slots, m, 0x41EFFFFFFFE00000ULL);
    else smh_fill_kernel<unsigned int><<<(m + 255) / 256, 256, 0, st>>>((unsigned int*)slots, m, 0x4F800000u);
    return cudaGetLastError();
}
cudaError_t launch_smh_colmin(const void* rows, uint64_t nseq, uint32_t m, bool f64, void* out, cudaStream_t st) {
    dim3 grid((m + 255) / 256, (unsigned)std::min<uint64_t>(nseq ? nseq : 1, 256));
    if (f64) smh_colmin_kernel<unsigned long long><<<grid, 256, 0, st>>>((const unsigned long long*)rows, nseq, m, (unsigned long long*)out);
    else smh_colmin_kernel<unsigned int><<<grid, 256, 0, st>>>((const unsigned int*)rows, nseq, m, (unsigned int*)out);
    return cudaGetLastError();
}

template <typename V, typename S, bool AA>
static cudaError_t launch_fast(const SmhParams& P, int grid, int block, size_t smem, cudaStream_t st) {
    auto kern = smh_fast_kernel<V, S, AA>;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    kern<<<grid, block, smem, st>>>(P);
    return cudaGetLastError();
}

template <typename V, bool AA>
static cudaError_t launch_fast_s(const SmhParams& P, bool f64, int grid, int block, size_t smem, cudaStream_t st) {
    return f64 ? launch_fast<V, double, AA>(P, grid, block, smem, st) : launch_fast<V, float, AA>(P, grid, block, smem, st);
}

// point 0 of every pre-key of a small key space: what smh_item_points does for a_spec == 0, once per key
template <typename S>
__global__ void smh_memo_kernel(uint4* memo, uint32_t nkeys, SmhParams P) {
    using F = FloatOps<S>;
    const uint32_t header = word_header(P.kmer_type, P.k);
    for (uint32_t pk = blockIdx.x * blockDim.x + threadIdx.x; pk < nkeys; pk += gridDim.x * blockDim.x) {
        const uint32_t key = finalize_key<uint32_t>(pk, header, P.hash_kind);
        Xoshiro256pp rng;
        rng.seed(item_seed<uint32_t>(key, P.hasher));
        const S r = F::draw(rng);
        const uint32_t slot = unif_from(rng, 0, P.m);
        const unsigned long long vb = (unsigned long long)F::bits(r);
        memo[pk] = make_uint4((uint32_t)vb, (uint32_t)(vb >> 32), slot, 0u);
    }
}

cudaError_t launch_smh_memo(const SmhParams& P, bool f64, void* memo, uint32_t nkeys, cudaStream_t st) {
    const int grid = (int)std::min<uint32_t>((nkeys + 255) / 256, 148 * 8);
    if (f64) smh_memo_kernel<double><<<grid, 256, 0, st>>>((uint4*)memo, nkeys, P);
    else smh_memo_kernel<float><<<grid, 256, 0, st>>>((uint4*)memo, nkeys, P);
    return cudaGetLastError();
}

cudaError_t launch_smh_fast(const SmhParams& P, bool key64, bool f64, int grid, int block, size_t smem, cudaStream_t st) {
    if (P.kmer_type == KMU_KMERAA32) return launch_fast_s<uint32_t, true>(P, f64, grid, block, smem, st);
    if (P.kmer_type == KMU_KMERAA64) return launch_fast_s<uint64_t, true>(P, f64, grid, block, smem, st);
    return key64 ? launch_fast_s<uint64_t, false>(P, f64, grid, block, smem, st)
                 : launch_fast_s<uint32_t, false>(P, f64, grid, block, smem, st);
}

template <typename V, bool AA>
static void launch_exact_s(const SmhParams& P, bool f64, int grid, cudaStream_t st) {
    if (f64) smh_exact_kernel<V, double, AA><<<grid, 32, 0, st>>>(P);
    else smh_exact_kernel<V, float, AA><<<grid, 32, 0, st>>>(P);
}

cudaError_t launch_smh_exact(const SmhParams& P, bool key64, bool f64, int grid, cudaStream_t st) {
    if (P.kmer_type == KMU_KMERAA32) launch_exact_s<uint32_t, true>(P, f64, grid, st);
    else if (P.kmer_type == KMU_KMERAA64) launch_exact_s<uint64_t, true>(P, f64, grid, st);
    else if (key64) launch_exact_s<uint64_t, false>(P, f64, grid, st);
    else launch_exact_s<uint32_t, false>(P, f64, grid, st);
    return cudaGetLastError();
}

}  // namespace kmu
