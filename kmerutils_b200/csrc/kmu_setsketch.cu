// kmu_setsketch.cu -- SetSketch (HyperLogLog-like) registers on sm_100a.
//
// Replaces HyperLogLogSketch::sketch_compressedkmer (one sketch per sequence,
// src/sketching/setsketchert.rs:758-802), ::sketch_compressedkmer_seqs(_block) (one sketch for a
// whole file, :677-724, 811-895; its block split + SetSketcher::merge is an element-wise max) and
// the amino-acid mirror (src/aautils/setsketchert.rs:790-1011): every k-mer is streamed through
// probminhash's SetSketcher::sketch (Ertl 2021, SetSketch1).
//
// GPU formulation.  Item x draws an increasing sequence x_0 < x_1 < ... (exponential spacings,
// rand_distr's ziggurat Exp1), turns x_j into the level k_j = clamp(floor(1 - log_b x_j), 0, q + 1)
// and offers it to register slot_j of an incremental Fisher-Yates shuffle; a register keeps the
// MAXIMUM.  The sequential algorithm stops an item as soon as k_j <= lower_k, a lower bound of the
// smallest register: such points cannot raise any register.  Hence
//   * speculative path: every item is cut at a level K_spec derived from the number of k-mers; the
//     result is exact iff the smallest register ends >= K_spec, which is verified.  A failed sketch
//     is redone with K_spec' = the smallest register it reached (a true lower bound), which cannot
//     fail.  Almost every item stops at j = 0 on a comparison of x_0 with a precomputed cut, before
//     any logarithm.
//   * exact path (few k-mers relative to m): one warp per sketch, 32 items per round with a full
//     lazily-reset permutation per lane in global scratch, lower_k = min register between rounds.
// log / exp are the deterministic routines of kmu_detmath.cuh (bit-identical to the CPU oracle).
#include <algorithm>
#include <cstdint>

#include "kmu_detmath.cuh"
#include "kmu_device.cuh"
#include "kmu_kernels.h"
#include "zig_exp_tables.h"

namespace kmu {

__constant__ double c_zig_x[257] = ZIG_EXP_TABLE_X;
__constant__ double c_zig_f[257] = ZIG_EXP_TABLE_F;

struct ZigTables {
    const double* x;  // 257 entries each, in shared memory
    const double* f;
};

__device__ __forceinline__ void load_zig_tables(double* sx, double* sf) {
    for (int i = threadIdx.x; i < 257; i += blockDim.x) {
        sx[i] = c_zig_x[i];
        sf[i] = c_zig_f[i];
    }
}

// rand 0.9 StandardUniform for f64: 53 bits, multiply method
__device__ __forceinline__ double std_uniform_f64(Xoshiro256pp& rng) {
    return __dmul_rn((double)(rng.next_u64() >> 11), 1.0 / 9007199254740992.0);
}

// rand_distr 0.5 Exp1 (ziggurat)
__device__ __forceinline__ double exp1_sample(Xoshiro256pp& rng, const ZigTables& z) {
    for (;;) {
        const uint64_t bits = rng.next_u64();
        const uint32_t i = (uint32_t)bits & 0xffu;
        const double u = __dsub_rn(__longlong_as_double((long long)((bits >> 12) | 0x3FF0000000000000ULL)),
                                   1.0 - 2.220446049250313e-16 / 2.0);
        const double x = __dmul_rn(u, z.x[i]);
        if (x < z.x[i + 1]) return x;
        if (i == 0) return __dsub_rn(ZIG_EXP_R, det_log(std_uniform_f64(rng)));
        const double f1 = z.f[i + 1];
        if (__dadd_rn(f1, __dmul_rn(__dsub_rn(z.f[i], f1), std_uniform_f64(rng))) < det_exp(-x)) return x;
    }
}

constexpr uint32_t SSK_QUEUE = 64;  // keys per warp queue of the team kernel (SSK_QUEUE_BYTES for 32 warps, kmu_kernels.h)
constexpr uint32_t SSK_SPARSE_CAP = 24;  // points an item may place in the speculative kernels

// Points of one item against the registers.  Returns false if the item needed more than
// SSK_SPARSE_CAP points (the caller redoes the sketch on the exact path).
__device__ __forceinline__ bool ssk_item_points(Xoshiro256pp& rng, const ZigTables& zt, const SskConsts& C, uint32_t kspec,
                                                double xcut, uint32_t* regs) {
    uint32_t idx[SSK_SPARSE_CAP], val[SSK_SPARSE_CAP];
    uint32_t n = 0;
    double x = 0.0;
    const uint32_t m = C.m;
    for (uint32_t j = 0; j < m; ++j) {
        const double e = exp1_sample(rng, zt);
        // almost every item stops at j = 0: its spacing is a constant, no division
        x = __dadd_rn(x, __dmul_rn(j == 0 ? C.inva_m0 : __ddiv_rn(C.inva, (double)(m - j)), e));
        if (x > xcut) break;  // certainly log_b(x) > -K_spec: the level is <= K_spec
        const double lb = __ddiv_rn(det_log(x), C.lnb);
        if (lb > -(double)kspec) break;
        const int z = __double2int_rd(__dsub_rn(1.0, lb));  // floor, saturating
        const int kk = max(0, min(C.iq1, z));
        if ((uint32_t)kk <= kspec) break;
        if (j >= SSK_SPARSE_CAP) return false;
        // FYshuffle::next: idx = lastidx + (usize)(U * (m - lastidx)), swap, value at idx is the slot
        const double xsi = rng.unif01();
        const uint32_t pi = j + (uint32_t)__double2uint_rz(__dmul_rn(xsi, (double)(m - j)));
        uint32_t vj = j, vk = pi;
        int pos_k = -1;
        for (uint32_t t = 0; t < n; ++t) {
            if (idx[t] == j) vj = val[t];
            if (idx[t] == pi) {
                vk = val[t];
                pos_k = (int)t;
            }
        }
        uint32_t slot = vj;
        if (pi != j) {
            slot = vk;
            if (pos_k >= 0) {
                val[pos_k] = vj;
            } else {
                idx[n] = pi;
                val[n] = vj;
                ++n;
            }
        }
        if ((uint32_t)kk > *(volatile uint32_t*)(regs + slot)) atomicMax(regs + slot, (uint32_t)kk);
    }
    return true;
}

// speculative level for a sequence of nk k-mers: with D distinct items every register ends >= K_spec except with
// probability ~m e^-spec_ln:  K_spec = 1 + floor(log_b(D a / spec_ln)),  D = spec_dfrac * (expected distinct keys among nk
// draws from the key space).  An item then places about m spec_ln / D points below the cut; a sketch that ends with a
// register below K_spec (a few percent of them, more for repetitive sequences) is redone from the level
// it reached.  Any K_spec gives the exact result; this one minimises the work.
__device__ __forceinline__ uint32_t ssk_kspec(uint64_t nk, const SskConsts& C) {
    double d = (double)nk;
    if (C.spec_keyspace > 0.0) d = C.spec_keyspace * (1.0 - exp(-d / C.spec_keyspace));
    const double ratio = d * C.spec_dfrac * C.a / C.spec_ln;
    if (!(ratio > 1.0)) return 0;
    const double kf = 1.0 + floor(log(ratio) / C.lnb);
    const double cap = (double)(C.iq1 - 1);
    return (uint32_t)(kf < cap ? kf : cap);
}
// the cautious level (D >= nk / 4, failure odds ~1e-4 per sketch): what a failed speculation is redone with when it
// left a register below it -- the redo pass must not fail in its turn, the path after it is one warp per sequence
__device__ __forceinline__ uint32_t ssk_kspec_cautious(uint64_t nk, const SskConsts& C) {
    const double ratio = (double)nk * 0.25 * C.a / C.ln_term;
    if (!(ratio > 1.0)) return 0;
    const double kf = 1.0 + floor(log(ratio) / C.lnb);
    const double cap = (double)(C.iq1 - 1);
    return (uint32_t)(kf < cap ? kf : cap);
}
__device__ __forceinline__ double ssk_xcut(uint32_t kspec, const SskConsts& C) {
    return exp(-(double)kspec * C.lnb) * (1.0 + 1e-6);
}

__device__ __forceinline__ void ssk_store(void* out, int sig_bytes, size_t i, uint32_t v) {
    if (sig_bytes == 2) ((uint16_t*)out)[i] = (uint16_t)v;
    else if (sig_bytes == 4) ((uint32_t*)out)[i] = v;
    else ((uint64_t*)out)[i] = v;
}

struct SskTeamShared {
    uint32_t seq, nbases, valid, flag, minreg;
    uint32_t pad;
    uint64_t byte_off;
};

// ---- one sketch per sequence, one team per sequence ---------------------------------------------
template <typename V, bool AA>
__global__ void __launch_bounds__(1024, 1) ssk_team_kernel(const SskParams P) {
    using TK = typename KmerSource<V, AA>::Task;
    extern __shared__ __align__(16) uint8_t smem[];
    double* zx = (double*)smem;
    double* zf = zx + 264;
    load_zig_tables(zx, zf);
    const ZigTables zt{zx, zf};
    Team team;
    team.size = P.team_warps * 32;
    team.id = threadIdx.x / team.size;
    team.tid = threadIdx.x - team.id * team.size;
    team.warp = team.tid >> 5;
    team.lane = threadIdx.x & 31;
    V* wq = (V*)(smem + 2 * 264 * sizeof(double)) + (threadIdx.x >> 5) * SSK_QUEUE;  // this warp's queue of keys
    uint8_t* tbase = smem + 2 * 264 * sizeof(double) + SSK_QUEUE_BYTES + (size_t)team.id * P.team_smem_bytes;
    uint32_t* regs = (uint32_t*)tbase;
    SskTeamShared* ts = (SskTeamShared*)(tbase + (((size_t)P.C.m * 4 + 15) & ~(size_t)15));
    const V header = (V)word_header(P.kmer_type, P.k);
    const bool canonical = hash_is_canonical(P.hash_kind);
    const uint32_t k = P.k, m = P.C.m;
    __syncthreads();

    for (;;) {
        team.sync();
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
            ts->minreg = 0xFFFFFFFFu;
        }
        team.sync();
        if (!ts->valid) break;
        const uint32_t seq = ts->seq, L = ts->nbases;
        const uint32_t* words = (const uint32_t*)(P.packed + ts->byte_off);
        const uint32_t nk = L >= k ? L - k + 1 : 0;
        const uint32_t kspec = P.kspec_in ? P.kspec_in[seq] : ssk_kspec(nk, P.C);
        if (nk && (kspec == 0 || nk <= P.exact_nk_max)) {  // few k-mers: the exact path is cheaper and always right
            if (team.tid == 0) P.exact_list[atomicAdd(P.exact_count, 1ULL)] = seq;
            continue;
        }
        const double xcut = ssk_xcut(kspec, P.C);
        for (uint32_t j = team.tid; j < m; j += team.size) regs[j] = 0;
        team.sync();
        const uint32_t ntasks = (nk + 15) >> 4;
        bool ok = true;
        // Two phases, so that the lanes of a warp stay together: (1) every lane seeds its item and draws the first
        // point -- the same work for all; the items whose first point falls below the cut (a minority) go to the
        // warp's queue; (2) whenever 32 are queued, every lane takes one and places all its points.  Without the queue
        // a warp waits at every item for its unluckiest lane (9.8 of 32 lanes active, profiles/r1m_ncu_ssk_team.txt).
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
                bool surv = false;
                V key = 0;
                if (p < pend) {
                    key = finalize_key<V>(tk.get(t, canonical), header, P.hash_kind);
                    // the first draw from half a seeding (s0 and s3 of Xoshiro256++, as in ssk_whole_warp_kernel); an item
                    // whose ziggurat draw is not accepted at once is queued like a survivor
                    const uint64_t sd = nohash_seed(key);
                    uint64_t x0 = sd, x3 = sd + 3ULL * 0x9E3779B97F4A7C15ULL;
                    const uint64_t s0 = Xoshiro256pp::splitmix(x0), s3 = Xoshiro256pp::splitmix(x3);
                    const uint64_t bits = rotl64(s0 + s3, 23) + s0;
                    const uint32_t zi = (uint32_t)bits & 0xffu;
                    const double e0 = __dmul_rn(__dsub_rn(__longlong_as_double((long long)((bits >> 12) | 0x3FF0000000000000ULL)),
                                                          1.0 - 2.220446049250313e-16 / 2.0),
                                                zx[zi]);
                    surv = !(e0 < zx[zi + 1] && __dmul_rn(P.C.inva_m0, e0) > xcut);
                }
                const uint32_t bal = __ballot_sync(0xFFFFFFFFu, surv);
                if (surv) wq[qn + __popc(bal & ((1u << team.lane) - 1u))] = key;
                qn += __popc(bal);
                __syncwarp();
                if (qn >= 32) {
                    qn -= 32;
                    const V kq = wq[qn + team.lane];
                    __syncwarp();
                    Xoshiro256pp rng;
                    rng.seed(nohash_seed(kq));
                    ok &= ssk_item_points(rng, zt, P.C, kspec, xcut, regs);
                }
            }
        }
        if ((uint32_t)team.lane < qn) {
            Xoshiro256pp rng;
            rng.seed(nohash_seed(wq[team.lane]));
            ok &= ssk_item_points(rng, zt, P.C, kspec, xcut, regs);
        }
        __syncwarp();
        if (!ok) ts->flag = 2;
        team.sync();
        uint32_t mn = 0xFFFFFFFFu;
        for (uint32_t j = team.tid; j < m; j += team.size) {
            const uint32_t v = regs[j];
            ssk_store(P.sig, P.sig_bytes, (size_t)seq * m + j, v);
            mn = v < mn ? v : mn;
        }
        mn = __reduce_min_sync(0xFFFFFFFFu, mn);
        if (team.lane == 0 && nk) atomicMin(&ts->minreg, mn);
        team.sync();
        if (team.tid == 0 && nk) {
            if (ts->flag == 2) {  // an item overflowed the sparse permutation: exact path
                P.exact_list[atomicAdd(P.exact_count, 1ULL)] = seq;
            } else if (ts->minreg < kspec) {
                // speculation failed: redo with the level reached (a true lower bound of every final register, cannot
                // fail) or, when that is lower than the cautious level, with the cautious one
                P.kmin_out[seq] = max(ts->minreg, ssk_kspec_cautious(nk, P.C));
                P.slow_list[atomicAdd(P.slow_count, 1ULL)] = seq;
            }
        }
    }
}

// ---- one sketch for the whole batch: every CTA keeps partial registers in shared memory ----------
template <typename V, bool AA>
__global__ void __launch_bounds__(512, 2) ssk_whole_kernel(const SskParams P, SeqView b, uint64_t total_bytes, uint32_t kspec,
                                                            double xcut) {
    extern __shared__ __align__(16) uint8_t smem[];
    double* zx = (double*)smem;
    double* zf = zx + 264;
    uint32_t* regs = (uint32_t*)(zf + 264);
    load_zig_tables(zx, zf);
    const ZigTables zt{zx, zf};
    const uint32_t m = P.C.m;
    for (uint32_t j = threadIdx.x; j < m; j += blockDim.x) regs[j] = 0;
    __syncthreads();
    const V header = (V)word_header(P.kmer_type, P.k);
    const bool canonical = hash_is_canonical(P.hash_kind);
    const uint64_t nchunks = (total_bytes + CHUNK_BYTES - 1) / CHUNK_BYTES;
    bool ok = true;
    for (uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; c < nchunks; c += (uint64_t)gridDim.x * blockDim.x)
        for_each_kmer_in_chunk<V, AA>(b, total_bytes, c, P.k, canonical, [&](V pk) {
            const V key = finalize_key<V>(pk, header, P.hash_kind);
            Xoshiro256pp rng;
            rng.seed(nohash_seed(key));
            ok &= ssk_item_points(rng, zt, P.C, kspec, xcut, regs);
        });
    if (!ok) *P.slow_count = 1ULL;  // overflow flag of the whole-batch path
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < m; j += blockDim.x) {
        const uint32_t v = regs[j];
        if (v) atomicMax(P.whole_regs + j, v);
    }
}

// DNA form of the whole-batch kernel, warp-cooperative: the lanes of a warp take 128 consecutive positions of the warp's
// slice at a time (four per lane, warp_for_each_kmer) and stay together, which allows two phases -- (1) half a seeding per item: the first output of
// Xoshiro256++ needs only s0 and s3 (SplitMix64 outputs 1 and 4 of the seed), and almost every item of a long input stops
// on that draw; (2) the items whose ziggurat draw is not accepted at once, or whose first point falls below the cut
// (about one in 90), are queued per warp and finished 32 at a time with the full generator.  (In the per-thread-chunk
// form above the lanes that take the rare path never reconverge with the others before the end of their chunk.)
constexpr uint32_t SSK_WHOLE_WARPS = 16;
template <typename V>
__global__ void __launch_bounds__(32 * SSK_WHOLE_WARPS, 2) ssk_whole_warp_kernel(const SskParams P, SeqView b, uint64_t total_bytes,
                                                                                   uint32_t kspec, double xcut) {
    extern __shared__ __align__(16) uint8_t smem[];
    double* zx = (double*)smem;
    double* zf = zx + 264;
    V* wq = (V*)(zf + 264) + (threadIdx.x >> 5) * SSK_QUEUE;
    uint32_t* regs = (uint32_t*)((V*)(zf + 264) + SSK_WHOLE_WARPS * SSK_QUEUE);
    load_zig_tables(zx, zf);
    const ZigTables zt{zx, zf};
    const uint32_t m = P.C.m;
    for (uint32_t j = threadIdx.x; j < m; j += blockDim.x) regs[j] = 0;
    __syncthreads();
    const V header = (V)word_header(P.kmer_type, P.k);
    const bool canonical = hash_is_canonical(P.hash_kind);
    const int lane = threadIdx.x & 31;
    const uint64_t ngroups = (total_bytes + GROUP_BYTES - 1) / GROUP_BYTES;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    bool ok = true;
    uint32_t qn = 0;  // warp-uniform
    for (uint64_t g = warp; g < ngroups; g += nwarps)
        warp_for_each_kmer<V>(b, total_bytes, g, P.k, canonical, lane, [&](V pk, bool active) {
            const V key = finalize_key<V>(pk, header, P.hash_kind);
            bool full = false;
            if (active) {
                const uint64_t sd = nohash_seed(key);
                uint64_t x0 = sd, x3 = sd + 3ULL * 0x9E3779B97F4A7C15ULL;
                const uint64_t s0 = Xoshiro256pp::splitmix(x0), s3 = Xoshiro256pp::splitmix(x3);
                const uint64_t bits = rotl64(s0 + s3, 23) + s0;
                const uint32_t zi = (uint32_t)bits & 0xffu;
                const double e0 = __dmul_rn(__dsub_rn(__longlong_as_double((long long)((bits >> 12) | 0x3FF0000000000000ULL)),
                                                      1.0 - 2.220446049250313e-16 / 2.0),
                                            zx[zi]);
                full = !(e0 < zx[zi + 1] && __dmul_rn(P.C.inva_m0, e0) > xcut);
            }
            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, full);
            if (full) wq[qn + __popc(bal & ((1u << lane) - 1u))] = key;
            qn += __popc(bal);
            __syncwarp();
            if (qn >= 32) {
                qn -= 32;
                const V kq = wq[qn + lane];
                __syncwarp();
                Xoshiro256pp rng;
                rng.seed(nohash_seed(kq));
                ok &= ssk_item_points(rng, zt, P.C, kspec, xcut, regs);
            }
        });
    if ((uint32_t)lane < qn) {
        Xoshiro256pp rng;
        rng.seed(nohash_seed(wq[lane]));
        ok &= ssk_item_points(rng, zt, P.C, kspec, xcut, regs);
    }
    if (!ok) *P.slow_count = 1ULL;  // overflow flag of the whole-batch path
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < m; j += blockDim.x) {
        const uint32_t v = regs[j];
        if (v) atomicMax(P.whole_regs + j, v);
    }
}

// registers (u32) -> signature type
__global__ void ssk_store_kernel(const uint32_t* regs, uint32_t m, void* out, int sig_bytes) {
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x)
        ssk_store(out, sig_bytes, j, regs[j]);
}

// ---- exact path: one warp per sketch --------------------------------------------------------------
// group mode (P.group != 0): one warp sketches ALL sequences order[first .. first+count) into one
// signature (row 0); otherwise warps pull single sequences.
template <typename V, bool AA>
__global__ void __launch_bounds__(32) ssk_exact_kernel(const SskParams P) {
    using Walker = typename KmerSource<V, AA>::Walker;
    __shared__ double zx[264], zf[264];
    load_zig_tables(zx, zf);
    __syncwarp();
    const ZigTables zt{zx, zf};
    const int lane = threadIdx.x;
    const uint32_t k = P.k, m = P.C.m;
    const V header = (V)word_header(P.kmer_type, P.k);
    const bool canonical = hash_is_canonical(P.hash_kind);
    uint8_t* base = P.scratch + (size_t)blockIdx.x * P.scratch_per_warp;
    uint32_t* regs = (uint32_t*)base;
    uint32_t* pl = regs + ((m + 3) & ~3u) + (size_t)lane * 2 * m;
    uint32_t* ql = pl + m;
    uint32_t stamp = 0;
    for (;;) {
        unsigned long long w = 0;
        if (lane == 0) w = atomicAdd(P.work_counter, 1ULL);
        w = __shfl_sync(0xFFFFFFFFu, w, 0);
        const uint64_t nwork = P.group ? 1 : P.count;
        if (w >= nwork) break;
        for (uint32_t j = lane; j < m; j += 32) regs[j] = 0;
        __syncwarp();
        uint32_t lower = 0;
        const uint64_t s_begin = P.group ? 0 : w, s_end = P.group ? P.count : w + 1;
        for (uint64_t si = s_begin; si < s_end; ++si) {
            const uint32_t seq = P.order[P.first + si];
            const uint32_t L = (uint32_t)P.nbases[seq];
            const uint32_t* words = (const uint32_t*)(P.packed + P.byte_off[seq]);
            const uint32_t nk = L >= k ? L - k + 1 : 0;
            for (uint32_t p0 = 0; p0 < nk; p0 += 32) {
                const uint32_t pos = p0 + lane;
                ++stamp;
                if (pos < nk) {
                    Walker wk;
                    wk.start(words, pos, k);
                    wk.roll();
                    const V key = finalize_key<V>(wk.prekey(canonical), header, P.hash_kind);
                    Xoshiro256pp rng;
                    rng.seed(nohash_seed(key));
                    double x = 0.0;
                    for (uint32_t j = 0; j < m; ++j) {
                        const double e = exp1_sample(rng, zt);
                        x = __dadd_rn(x, __dmul_rn(__ddiv_rn(P.C.inva, (double)(m - j)), e));
                        const double lb = __ddiv_rn(det_log(x), P.C.lnb);
                        if (lb > -(double)lower) break;
                        const int z = __double2int_rd(__dsub_rn(1.0, lb));
                        const int kk = max(0, min(P.C.iq1, z));
                        if ((uint32_t)kk <= lower) break;
                        const double xsi = rng.unif01();
                        const uint32_t pi = j + (uint32_t)__double2uint_rz(__dmul_rn(xsi, (double)(m - j)));
                        if (ql[j] != stamp) {
                            ql[j] = stamp;
                            pl[j] = j;
                        }
                        if (ql[pi] != stamp) {
                            ql[pi] = stamp;
                            pl[pi] = pi;
                        }
                        const uint32_t slot = pl[pi];
                        pl[pi] = pl[j];
                        pl[j] = slot;
                        if ((uint32_t)kk > *(volatile uint32_t*)(regs + slot)) atomicMax(regs + slot, (uint32_t)kk);
                    }
                }
                __syncwarp();
                uint32_t mn = 0xFFFFFFFFu;
                for (uint32_t j = lane; j < m; j += 32) {
                    const uint32_t v = *(volatile uint32_t*)(regs + j);
                    mn = v < mn ? v : mn;
                }
                lower = __reduce_min_sync(0xFFFFFFFFu, mn);
                __syncwarp();
            }
        }
        const size_t row = P.group ? 0 : (size_t)P.order[P.first + w];
        for (uint32_t j = lane; j < m; j += 32) ssk_store(P.sig, P.sig_bytes, row * m + j, regs[j]);
        __syncwarp();
    }
}

// ---- launchers ---------------------------------------------------------------------------------
template <typename V, bool AA>
static cudaError_t launch_team_t(const SskParams& P, int grid, int block, size_t smem, cudaStream_t st) {
    auto kern = ssk_team_kernel<V, AA>;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    kern<<<grid, block, smem, st>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_ssk_team(const SskParams& P, int grid, int block, size_t smem, cudaStream_t st) {
    switch (P.kmer_type) {
        case KMU_KMERAA32: return launch_team_t<uint32_t, true>(P, grid, block, smem, st);
        case KMU_KMERAA64: return launch_team_t<uint64_t, true>(P, grid, block, smem, st);
        case KMU_KMER64: return launch_team_t<uint64_t, false>(P, grid, block, smem, st);
        default: return launch_team_t<uint32_t, false>(P, grid, block, smem, st);
    }
}

template <typename V, bool AA>
static cudaError_t launch_whole_t(const SskParams& P, const SeqView& b, uint64_t total_bytes, uint32_t kspec, double xcut,
                                  int grid, size_t smem, cudaStream_t st) {
    if constexpr (!AA) {  // DNA: the warp-cooperative form (queues of SSK_QUEUE keys per warp in front of the registers)
        auto kern = ssk_whole_warp_kernel<V>;
        smem += SSK_WHOLE_WARPS * SSK_QUEUE * sizeof(V);
        static size_t configured = 0;
        if (smem > configured) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            configured = smem;
        }
        const uint64_t ngroups = (total_bytes + GROUP_BYTES - 1) / GROUP_BYTES;
        const uint64_t want = (ngroups + SSK_WHOLE_WARPS - 1) / SSK_WHOLE_WARPS;
        // `grid` is sized for 512 chunks of 64 bytes per CTA: at most two CTAs per SM
        const int g2 = (int)std::max<uint64_t>(1, std::min<uint64_t>(want, (uint64_t)std::max(grid, 1) >= want ? want : (uint64_t)grid));
        kern<<<g2, 32 * SSK_WHOLE_WARPS, smem, st>>>(P, b, total_bytes, kspec, xcut);
        return cudaGetLastError();
    }
    auto kern = ssk_whole_kernel<V, AA>;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    kern<<<grid, 512, smem, st>>>(P, b, total_bytes, kspec, xcut);
    return cudaGetLastError();
}

cudaError_t launch_ssk_whole(const SskParams& P, const SeqView& b, uint64_t total_bytes, uint32_t kspec, double xcut,
                             int grid, size_t smem, cudaStream_t st) {
    switch (P.kmer_type) {
        case KMU_KMERAA32: return launch_whole_t<uint32_t, true>(P, b, total_bytes, kspec, xcut, grid, smem, st);
        case KMU_KMERAA64: return launch_whole_t<uint64_t, true>(P, b, total_bytes, kspec, xcut, grid, smem, st);
        case KMU_KMER64: return launch_whole_t<uint64_t, false>(P, b, total_bytes, kspec, xcut, grid, smem, st);
        default: return launch_whole_t<uint32_t, false>(P, b, total_bytes, kspec, xcut, grid, smem, st);
    }
}

cudaError_t launch_ssk_store(const uint32_t* regs, uint32_t m, void* out, int sig_bytes, cudaStream_t st) {
    ssk_store_kernel<<<(m + 255) / 256, 256, 0, st>>>(regs, m, out, sig_bytes);
    return cudaGetLastError();
}

cudaError_t launch_ssk_exact(const SskParams& P, int grid, cudaStream_t st) {
    switch (P.kmer_type) {
        case KMU_KMERAA32: ssk_exact_kernel<uint32_t, true><<<grid, 32, 0, st>>>(P); break;
        case KMU_KMERAA64: ssk_exact_kernel<uint64_t, true><<<grid, 32, 0, st>>>(P); break;
        case KMU_KMER64: ssk_exact_kernel<uint64_t, false><<<grid, 32, 0, st>>>(P); break;
        default: ssk_exact_kernel<uint32_t, false><<<grid, 32, 0, st>>>(P); break;
    }
    return cudaGetLastError();
}

}  // namespace kmu
