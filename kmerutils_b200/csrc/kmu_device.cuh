// kmu_device.cuh -- device-side building blocks shared by every kernel of the
// B200 k-mer engine: 2-bit window arithmetic, the reference's hash closures, the
// per-item random streams of the sketchers and the 128-bit slot update.
//
// Citations are path:line in the reference tree (jean-pierreBoth/kmerutils v0.0.14);
// "App. A.x" refers to SURVEY.md appendix A (arithmetic of the un-vendored
// probminhash / rand / rand_xoshiro crates).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/kmerutils_b200.h"
#include "kmu_detmath.cuh"
#include "kmu_kernels.h"

namespace kmu {

constexpr uint64_t F64_MAX_BITS = 0x7FEFFFFFFFFFFFFFULL;  // f64::MAX, MaxValueTracker's initial value

// ----------------------------------------------------------------------------
//  2-bit words
// ----------------------------------------------------------------------------
// packed bytes hold the first base in their most significant bits (alphabet.rs:162-168);
// a byte-swapped 32-bit load therefore gives 16 bases, first base in bits 31..30.
__device__ __forceinline__ uint32_t be32(uint32_t le_word) { return __byte_perm(le_word, 0, 0x0123); }

// reverse complement of a full 32-bit / 64-bit word of bases (kmer16b32bit.rs:43-54,
// kmer64bit.rs:83-96 before the final shift): complement, reverse the bits, swap the
// two bits of every pair back.
__device__ __forceinline__ uint32_t revcomp_word32(uint32_t w) {
    uint32_t r = __brev(~w);
    return ((r & 0x55555555u) << 1) | ((r & 0xAAAAAAAAu) >> 1);
}
__device__ __forceinline__ uint64_t revcomp_word64(uint64_t w) {
    uint64_t r = __brevll(~w);
    return ((r & 0x5555555555555555ULL) << 1) | ((r & 0xAAAAAAAAAAAAAAAAULL) >> 1);
}
// reverse complement of a right-aligned k-base value (k >= 1)
__device__ __forceinline__ uint32_t revcomp_val(uint32_t v, uint32_t k) { return revcomp_word32(v) >> (32 - 2 * k); }
__device__ __forceinline__ uint64_t revcomp_val(uint64_t v, uint32_t k) { return revcomp_word64(v) >> (64 - 2 * k); }

template <typename V>
__device__ __forceinline__ V value_mask(uint32_t nbits) {
    return nbits >= 8 * sizeof(V) ? ~V(0) : ((V(1) << nbits) - 1);
}

// probminhash::invhash (App. A.6)
__device__ __forceinline__ uint32_t int32_hash(uint32_t key) {
    key += ~(key << 15);
    key ^= (key >> 10);
    key += (key << 3);
    key ^= (key >> 6);
    key += ~(key << 11);
    key ^= (key >> 16);
    return key;
}
__device__ __forceinline__ uint64_t int64_hash(uint64_t key) {
    key = (~key) + (key << 21);
    key = key ^ (key >> 24);
    key = (key + (key << 3)) + (key << 8);
    key = key ^ (key >> 14);
    key = (key + (key << 2)) + (key << 4);
    key = key ^ (key >> 28);
    key = key + (key << 31);
    return key;
}
__device__ __forceinline__ uint32_t inv_hash(uint32_t k) { return int32_hash(k); }
__device__ __forceinline__ uint64_t inv_hash(uint64_t k) { return int64_hash(k); }

// ----------------------------------------------------------------------------
//  hash closures.  The walker hands out a *pre-key*: the forward
//  value or, for the canonical kinds, min(value, revcomp value) -- Ord of all
//  three k-mer types reduces to the value when k is equal (kmer32bit.rs:47-55,
//  kmer64bit.rs:45-53).  finalize_key() turns the pre-key into fhash(kmer).
// ----------------------------------------------------------------------------
__host__ __device__ __forceinline__ bool hash_is_canonical(int hash_kind) {
    return hash_kind == KMU_HASH_CANON_INVHASH || hash_kind == KMU_HASH_CANON_RAW;
}
// header of the k-mer word: Kmer32bit keeps k in its top 4 bits (kmer32bit.rs:68-76)
__host__ __device__ __forceinline__ uint32_t word_header(int kmer_type, uint32_t k) {
    return kmer_type == KMU_KMER32 ? (k << 28) : 0u;
}
template <typename V>
__device__ __forceinline__ V finalize_key(V prekey, V header, int hash_kind) {
    switch (hash_kind) {
        case KMU_HASH_MASKED_VALUE: return prekey;  // value & (2^(bits*k) - 1): the header is masked off
        case KMU_HASH_CANON_INVHASH:
        case KMU_HASH_INVHASH: return inv_hash(V(prekey | header));
        default: return V(prekey | header);  // IDENTITY_RAW, CANON_RAW: the raw word `.0`
    }
}

// ----------------------------------------------------------------------------
//  Per-item random stream: Xoshiro256++ seeded by SplitMix64 from the
//  NoHashHasher / FNV value of the key (App. A.1, src/nohasher.rs:22-48)
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint64_t nohash_seed(uint32_t key) { return (uint64_t)__byte_perm(key, 0, 0x0123); }
__device__ __forceinline__ uint64_t nohash_seed(uint64_t key) {
    uint32_t lo = (uint32_t)key, hi = (uint32_t)(key >> 32);
    return ((uint64_t)__byte_perm(lo, 0, 0x0123) << 32) | (uint64_t)__byte_perm(hi, 0, 0x0123);
}
template <typename V>
__device__ __forceinline__ uint64_t fnv1a_seed(V key) {
    uint64_t h = 0xcbf29ce484222325ULL;
#pragma unroll
    for (int i = 0; i < (int)sizeof(V); ++i) {
        h ^= (uint64_t)((key >> (8 * i)) & 0xFF);
        h *= 0x100000001b3ULL;
    }
    return h;
}

__device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

struct Xoshiro256pp {
    uint64_t s0, s1, s2, s3;
    __device__ __forceinline__ void seed(uint64_t seed) {
        uint64_t x = seed;
        s0 = splitmix(x);
        s1 = splitmix(x);
        s2 = splitmix(x);
        s3 = splitmix(x);
    }
    static __device__ __forceinline__ uint64_t splitmix(uint64_t& x) {
        x += 0x9E3779B97F4A7C15ULL;
        uint64_t z = x;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
    __device__ __forceinline__ uint64_t next_u64() {
        uint64_t r = rotl64(s0 + s3, 23) + s0;
        uint64_t t = s1 << 17;
        s2 ^= s0;
        s3 ^= s1;
        s1 ^= s2;
        s0 ^= s3;
        s2 ^= t;
        s3 = rotl64(s3, 45);
        return r;
    }
    __device__ __forceinline__ uint32_t next_u32() { return (uint32_t)(next_u64() >> 32); }
    // rand 0.9 Uniform::<f64>::new(0., 1.): 52 mantissa bits, [1,2) - 1 (App. A.2)
    __device__ __forceinline__ double unif01() {
        return __longlong_as_double((long long)((next_u64() >> 12) | 0x3FF0000000000000ULL)) - 1.0;
    }
    __device__ __forceinline__ float unif01_f32() { return __uint_as_float((next_u32() >> 9) | 0x3F800000u) - 1.0f; }
    // rand 0.9 UniformUsize (range fits u32): widening multiply on next_u32, reject lo < thresh
    __device__ __forceinline__ uint32_t unif_range(uint32_t low, uint32_t range, uint32_t thresh) {
        for (;;) {
            uint32_t v = next_u32();
            uint32_t lo = v * range;
            if (lo >= thresh) return low + __umulhi(v, range);
        }
    }
};

// ExpRestricted01 (App. A.3). The constants are computed on the host with libm
// (the reference gets them from the same libm through Rust's f64::exp_m1 / ln / exp).
// every product / sum is an explicit round-to-nearest intrinsic: Rust never
// contracts a*b+c into an FMA, so neither may nvcc.
__device__ __forceinline__ double exp01_sample(const Exp01Params& p, Xoshiro256pp& rng) {
    double x = __dmul_rn(p.c1, rng.unif01());
    if (x < 1.0) return x;
    for (;;) {
        x = rng.unif01();
        if (x < p.c2) return x;
        double y = __dmul_rn(0.5, rng.unif01());
        if (y > __dsub_rn(1.0, x)) {
            x = __dsub_rn(1.0, x);
            y = __dsub_rn(1.0, y);
        }
        if (x <= __dmul_rn(p.c3, __dsub_rn(1.0, y))) return x;
        if (__dmul_rn(p.c1, y) <= __dsub_rn(1.0, x)) return x;
        // deterministic expm1 shared with the CPU oracle (kmu_detmath.cuh == oracle/det_math.hpp): CUDA's expm1 and glibc's
        // may differ by one ulp
        if (__dmul_rn(__dmul_rn(y, p.c1), p.lambda) <= det_expm1(__dmul_rn(p.lambda, __dsub_rn(1.0, x)))) return x;
    }
}

// ----------------------------------------------------------------------------
//  Sketch slots: 16-byte records {hbits, key}.  A point (h, key) replaces the
//  record when (h, key) is lexicographically smaller; h > 0 so the IEEE bit
//  pattern orders like the value.  One 128-bit compare-and-swap (ATOMS.CAS.128 /
//  ATOMG.CAS.128 on sm_100a) makes the update atomic for both fields.
// ----------------------------------------------------------------------------
__device__ __forceinline__ void cas128(Slot* addr, uint64_t cmp_h, uint64_t cmp_k, uint64_t new_h, uint64_t new_k,
                                       uint64_t& old_h, uint64_t& old_k) {
    asm volatile(
        "{\n\t.reg .b128 c, n, o;\n\t"
        "mov.b128 c, {%3, %4};\n\t"
        "mov.b128 n, {%5, %6};\n\t"
        "atom.cas.b128 o, [%2], c, n;\n\t"
        "mov.b128 {%0, %1}, o;\n\t}"
        : "=l"(old_h), "=l"(old_k)
        : "l"(addr), "l"(cmp_h), "l"(cmp_k), "l"(new_h), "l"(new_k)
        : "memory");
}

// returns true when the record was replaced
__device__ __forceinline__ bool slot_update_min(Slot* slot, uint64_t hbits, uint64_t key) {
    uint64_t cur_h = *(volatile uint64_t*)&slot->hbits;
    if (hbits > cur_h) return false;
    uint64_t cur_k = *(volatile uint64_t*)&slot->key;
    for (;;) {
        if (hbits > cur_h || (hbits == cur_h && key >= cur_k)) return false;
        uint64_t old_h, old_k;
        cas128(slot, cur_h, cur_k, hbits, key, old_h, old_k);
        if (old_h == cur_h && old_k == cur_k) return true;
        cur_h = old_h;
        cur_k = old_k;
    }
}

// warp-wide maximum of a u64 through two redux.sync passes (redux is 32-bit)
__device__ __forceinline__ uint64_t warp_max_u64(uint64_t v) {
    uint32_t hi = (uint32_t)(v >> 32);
    uint32_t mhi = __reduce_max_sync(0xFFFFFFFFu, hi);
    uint32_t lo = hi == mhi ? (uint32_t)v : 0u;
    uint32_t mlo = __reduce_max_sync(0xFFFFFFFFu, lo);
    return ((uint64_t)mhi << 32) | mlo;
}
__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v) { return ~warp_max_u64(~v); }

// Is an item still alive after its first point, i.e. winv * x1 < qmax ?  x1 = c1 * U where U comes
// from the first output of the key's Xoshiro256++, which needs only the state words s0 and s3
// (SplitMix64 outputs 1 and 4 of the seed): half a seeding, no memory.  Items on the rare
// rejection branch of ExpRestricted01 (c1 * U >= 1) are left to the full path.
template <typename V>
__device__ __forceinline__ bool first_point_alive(V key, double winv, double qmax, double c1) {
    const uint64_t seed = nohash_seed(key);
    uint64_t z0 = seed + 0x9E3779B97F4A7C15ULL;
    uint64_t z3 = seed + 4ULL * 0x9E3779B97F4A7C15ULL;
    z0 = (z0 ^ (z0 >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z3 = (z3 ^ (z3 >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z0 = (z0 ^ (z0 >> 27)) * 0x94D049BB133111EBULL;
    z3 = (z3 ^ (z3 >> 27)) * 0x94D049BB133111EBULL;
    const uint64_t s0 = z0 ^ (z0 >> 31), s3 = z3 ^ (z3 >> 31);
    const uint64_t r = rotl64(s0 + s3, 23) + s0;
    const double u = __longlong_as_double((long long)((r >> 12) | 0x3FF0000000000000ULL)) - 1.0;
    const double x = __dmul_rn(c1, u);
    return !(x < 1.0) || __dmul_rn(winv, x) < qmax;
}

// ----------------------------------------------------------------------------
//  TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier: stages the packed bytes of
//  the next sequence into shared memory while the current one is processed.
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_bytes(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    // dst, src 16-byte aligned; bytes a multiple of 16
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "KMU_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra KMU_DONE;\n\t"
        "bra KMU_WAIT;\n\t"
        "KMU_DONE:\n\t}" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}

// --------------------------------------------------------------------------------
// team = group of warps cooperating on one sequence
// --------------------------------------------------------------------------------
struct Team {
    int id;      // team index inside the CTA
    int tid;     // thread index inside the team
    int size;    // threads in the team
    int warp;    // warp index inside the team
    int lane;
    __device__ __forceinline__ void sync() const {
        if (size == 32) {
            __syncwarp();
        } else {
            // named barrier id+1 (0 is left to __syncthreads)
            asm volatile("bar.sync %0, %1;" ::"r"(id + 1), "r"(size) : "memory");
        }
    }
};

// counter based SplitMix64 used by the synthetic generator
__host__ __device__ __forceinline__ uint64_t synth_z(uint64_t seed, uint64_t i) {
    uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

// ----------------------------------------------------------------------------
//  K-mer walker.  A thread positions the walker on any base of a sequence and
//  then pulls one k-mer per roll(): bases come from a left-aligned 64-bit shift
//  register refilled from big-endian 32-bit words, and the forward and
//  reverse-complement windows are rolled in registers (KmerSeqIterator::next,
//  kmergenerator.rs:75-106, yields the same windows one `push` at a time).
//  The reader looks ahead by at most one 32-bit word past the word holding the
//  last base it is asked for; batches keep that much slack after every sequence.
// ----------------------------------------------------------------------------
template <typename V>
struct KmerWalker {
    V fwd, rc, mask;
    uint64_t sr;  // upcoming bases, left aligned
    int nb;       // valid bits in sr; > 32 after every refill
    const uint32_t* wp;
    uint32_t rc_shift;  // 2k - 2

    __device__ __forceinline__ void refill() {
        if (nb <= 32) {
            sr |= (uint64_t)be32(*wp) << (32 - nb);
            ++wp;
            nb += 32;
        }
    }
    // next n bases (1 <= n <= 16) as a right-aligned value
    __device__ __forceinline__ uint32_t take(uint32_t n) {
        uint32_t v = (uint32_t)(sr >> (64 - 2 * n));
        sr <<= 2 * n;
        nb -= 2 * n;
        refill();
        return v;
    }
    // after start(words, p0, k) the first roll() yields the k-mer starting at base p0
    __device__ __forceinline__ void start(const uint32_t* words, uint64_t p0, uint32_t k) {
        mask = value_mask<V>(2 * k);
        rc_shift = 2 * k - 2;
        wp = words + (p0 >> 4);
        uint32_t o = (uint32_t)(p0 & 15);
        sr = (uint64_t)be32(*wp) << (32 + 2 * o);
        ++wp;
        nb = 32 - 2 * o;
        refill();
        uint32_t kk = k - 1;
        V pre = 0;
        if (kk > 16) {
            pre = (V)take(16);
            kk -= 16;
            pre = (V)(pre << (2 * kk)) | (V)take(kk);
        } else if (kk > 0) {
            pre = (V)take(kk);
        }
        fwd = pre;
        rc = k > 1 ? (V)(revcomp_val(pre, k - 1) << 2) : V(0);
    }
    __device__ __forceinline__ void roll() {
        uint32_t b = take(1);
        fwd = (V)(((fwd << 2) | (V)b) & mask);
        rc = (V)((rc >> 2) | ((V)(3u - b) << rc_shift));
    }
    __device__ __forceinline__ V prekey(bool canonical) const { return canonical ? (fwd < rc ? fwd : rc) : fwd; }
};

// ----------------------------------------------------------------------------
//  TaskKmers: the k-mers of one task = T consecutive start positions (T in
//  {4, 8, 16}, the first one a multiple of T).  For u32 k-mers (k <= 16) two
//  big-endian words hold every window of the task, so each k-mer is two funnel
//  shifts and a mask on the word pair and on its reverse complement; u64 k-mers
//  roll through KmerWalker.  get(t) must be called for t = 0, 1, 2, ... in order
//  and only for positions that exist.
// ----------------------------------------------------------------------------
template <typename V>
struct TaskKmers;

template <>
struct TaskKmers<uint32_t> {
    static constexpr bool SEQUENTIAL = false;  // get(t) may be called for any subset of t, in increasing order
    uint64_t W, RC;
    uint32_t mask, sh0, j0;
    __device__ __forceinline__ void init(const uint32_t* words, uint64_t p0, uint32_t k) {
        const uint32_t* w = words + (p0 >> 4);
        init_loaded(w[0], w[1], p0, k);
    }
    // the two words words[p0 >> 4], words[(p0 >> 4) + 1] were loaded by the caller (software pipelining)
    __device__ __forceinline__ void init_loaded(uint32_t w0_le, uint32_t w1_le, uint64_t p0, uint32_t k) {
        W = ((uint64_t)be32(w0_le) << 32) | be32(w1_le);
        RC = revcomp_word64(W);
        mask = value_mask<uint32_t>(2 * k);
        sh0 = 64 - 2 * k;
        j0 = (uint32_t)(p0 & 15);
    }
    __device__ __forceinline__ uint32_t get(uint32_t t, bool canonical) {
        uint32_t j = j0 + t;
        uint32_t f = (uint32_t)(W >> (sh0 - 2 * j)) & mask;
        if (!canonical) return f;
        uint32_t r = (uint32_t)(RC >> (2 * j)) & mask;
        return f < r ? f : r;
    }
    // Tasks of NT positions with 2 (NT - 1) + 2 k <= 32: all windows of the task sit in one 32-bit word per strand, so
    // that get_narrow(t) with a compile-time t is a constant shift and a mask.  Call narrow<NT>() once after init.
    uint32_t w32, rc32;
    template <int NT>
    __device__ __forceinline__ void narrow() {
        w32 = (uint32_t)(W >> (sh0 - 2 * (j0 + NT - 1)));
        rc32 = (uint32_t)(RC >> (2 * j0));
    }
    template <int NT>
    __device__ __forceinline__ uint32_t get_narrow(uint32_t t, bool canonical) const {
        const uint32_t f = (w32 >> (2 * (NT - 1 - t))) & mask;
        if (!canonical) return f;
        const uint32_t r = (rc32 >> (2 * t)) & mask;
        return f < r ? f : r;
    }
};

template <>
struct TaskKmers<uint64_t> {
    static constexpr bool SEQUENTIAL = true;  // get(t) must be called for every t = 0, 1, 2, ...
    KmerWalker<uint64_t> wk;
    __device__ __forceinline__ void init(const uint32_t* words, uint64_t p0, uint32_t k) { wk.start(words, p0, k); }
    __device__ __forceinline__ uint64_t get(uint32_t, bool canonical) {
        wk.roll();
        return wk.prekey(canonical);
    }
};

// ----------------------------------------------------------------------------
//  Amino-acid k-mers (src/aautils/kmeraa.rs): sequences are one 5-bit code per byte
//  (Alphabet::encode, kmeraa.rs:85-109, applied once at ingest); a k-mer is
//  sum code_i << 5 (k - 1 - i) in a u32 (KmerAA32bit, k <= 6) or u64 (KmerAA64bit,
//  k <= 12); push = ((v << 5) & mask) | code (kmeraa.rs:171-182, 301-312).  There is
//  no reverse complement (kmeraa.rs:185-187 panics), hence no canonical form.
// ----------------------------------------------------------------------------
template <typename V>
struct KmerWalkerAA {
    V fwd, mask;
    const uint8_t* bp;
    __device__ __forceinline__ void start(const void* codes, uint64_t p0, uint32_t k) {
        mask = value_mask<V>(5 * k);
        bp = (const uint8_t*)codes + p0;
        fwd = 0;
        for (uint32_t i = 0; i + 1 < k; ++i) fwd = (V)((fwd << 5) | (V)(*bp++));
    }
    __device__ __forceinline__ void roll() { fwd = (V)(((fwd << 5) | (V)(*bp++)) & mask); }
    __device__ __forceinline__ V prekey(bool) const { return fwd; }
};

template <typename V>
struct TaskKmersAA {
    static constexpr bool SEQUENTIAL = true;
    KmerWalkerAA<V> wk;
    __device__ __forceinline__ void init(const uint32_t* codes, uint64_t p0, uint32_t k) { wk.start(codes, p0, k); }
    __device__ __forceinline__ V get(uint32_t, bool) {
        wk.roll();
        return wk.fwd;
    }
};

template <typename V, bool AA>
struct KmerSource {
    using Task = TaskKmers<V>;
    using Walker = KmerWalker<V>;
};
template <typename V>
struct KmerSource<V, true> {
    using Task = TaskKmersAA<V>;
    using Walker = KmerWalkerAA<V>;
};

// ----------------------------------------------------------------------------
//  Chunked traversal of a whole batch: the byte buffer is cut into 64-byte chunks (256 bases or 64
//  residues); a thread owns the k-mers that START in its chunk, finds the sequence(s) overlapping
//  it by binary search in the batch's byte offsets and rolls the windows in registers.
// ----------------------------------------------------------------------------
constexpr uint32_t CHUNK_BYTES = 64;

// last sequence s with byte_off[s] <= byte (byte_off ascending, byte_off[0] == 0)
__device__ __forceinline__ uint64_t seq_of_byte(const uint64_t* __restrict__ byte_off, uint64_t nseq, uint64_t byte) {
    uint64_t lo = 0, hi = nseq;
    while (hi - lo > 1) {
        uint64_t mid = (lo + hi) >> 1;
        if (__ldg(byte_off + mid) <= byte) lo = mid; else hi = mid;
    }
    return lo;
}

// the same search by a whole (converged) warp: 32 probes per round, four rounds for a million sequences instead of twenty
// dependent loads
__device__ __forceinline__ uint64_t seq_of_byte_warp(const uint64_t* __restrict__ byte_off, uint64_t nseq, uint64_t byte) {
    const uint32_t lane = threadIdx.x & 31;
    uint64_t lo = 0, hi = nseq;  // byte_off[lo] <= byte, the answer is in [lo, hi)
    while (hi - lo > 1) {
        const uint64_t step = (hi - lo + 30) >> 5;
        const uint64_t idx = lo + (uint64_t)(lane + 1) * step;
        const bool le = idx < hi && __ldg(byte_off + idx) <= byte;
        const uint32_t c = __popc(__ballot_sync(0xFFFFFFFFu, le));  // the offsets ascend: the first c probes
        lo += c * step;
        hi = min(hi, lo + step);
    }
    return lo;
}

// Calls f(pre-key) for every k-mer that starts inside chunk `c` (pre-key: canonical or forward value)
template <typename V, bool AA, typename F>
__device__ __forceinline__ void for_each_kmer_in_chunk(const SeqView& b, uint64_t total_bytes, uint64_t c, uint32_t k,
                                                       bool canonical, F&& f) {
    constexpr uint64_t PER_BYTE = AA ? 1 : 4;  // sequence positions per byte
    const uint64_t byte0 = c * CHUNK_BYTES;
    const uint64_t byte1 = min(byte0 + (uint64_t)CHUNK_BYTES, total_bytes);
    uint64_t s = seq_of_byte(b.byte_off, b.nseq, byte0);
    while (s < b.nseq) {
        const uint64_t sb = __ldg(b.byte_off + s);
        if (sb >= byte1) break;
        const uint64_t L = __ldg(b.nbases + s);
        const uint64_t nk = L >= k ? L - k + 1 : 0;
        const uint64_t p_lo = byte0 > sb ? (byte0 - sb) * PER_BYTE : 0;
        const uint64_t p_hi = min(nk, (byte1 - sb) * PER_BYTE);
        if (p_lo < p_hi) {
            typename KmerSource<V, AA>::Walker wk;
            wk.start((const uint32_t*)(b.packed + sb), p_lo, k);
            for (uint64_t p = p_lo; p < p_hi; ++p) {
                wk.roll();
                f(wk.prekey(canonical));
            }
        }
        ++s;
    }
}

// ----------------------------------------------------------------------------
//  Direct k-mer read: the forward value of the k-mer starting at base p, straight from the packed
//  words (two or three big-endian words re-aligned with funnel shifts) -- no rolling state, so the
//  32 lanes of a warp can take 32 consecutive positions.
// ----------------------------------------------------------------------------
template <typename V>
__device__ __forceinline__ V kmer_at(const uint32_t* __restrict__ words, uint64_t p, uint32_t k);
template <>
__device__ __forceinline__ uint32_t kmer_at<uint32_t>(const uint32_t* __restrict__ words, uint64_t p, uint32_t k) {
    const uint32_t* w = words + (p >> 4);
    const uint32_t sh = (uint32_t)(p & 15) * 2;
    const uint32_t x = __funnelshift_l(be32(__ldg(w + 1)), be32(__ldg(w)), sh);  // 16 bases from p (k <= 16)
    return x >> (32 - 2 * k);
}
template <>
__device__ __forceinline__ uint64_t kmer_at<uint64_t>(const uint32_t* __restrict__ words, uint64_t p, uint32_t k) {
    const uint32_t* w = words + (p >> 4);
    const uint32_t sh = (uint32_t)(p & 15) * 2;
    const uint32_t a = be32(__ldg(w)), b = be32(__ldg(w + 1)), c = be32(__ldg(w + 2));
    const uint64_t x = ((uint64_t)__funnelshift_l(b, a, sh) << 32) | __funnelshift_l(c, b, sh);  // 32 bases from p
    return x >> (64 - 2 * k);
}

// Warp-cooperative traversal of a whole batch: the byte buffer is cut into groups of 2 KB; a warp
// owns the k-mers that START in its group and hands 128 consecutive positions at a time to its lanes,
// four per lane (warp_for_each_kmer below).  f(pre-key, active) is called four times per turn with all
// 32 lanes converged (inactive lanes: active == false).
constexpr uint32_t GROUP_BYTES = 2048;

// the 4 k-mers that start at positions q .. q + 3 of the word stream w (canonical: the smaller strand): one window of packed
// words and ONE reverse complement of the window serve all four, on both strands (as generate_kmers_run_kernel, kmu_extract.cu)
__device__ __forceinline__ uint32_t revcomp16_word(uint32_t w) {
    const uint32_t r = __brev(w);
    uint32_t d;  // ~(((r << 1) & 0xAAAAAAAA) | ((r >> 1) & 0x55555555)): lut 0x27 of (a, b, c) = ~(c ? b : a)
    asm("lop3.b32 %0, %1, %2, 0x55555555, 0x27;" : "=r"(d) : "r"(r << 1), "r"(r >> 1));
    return d;
}
template <bool GLOBAL = true>
__device__ __forceinline__ uint32_t packed_word(const uint32_t* a) {
    return GLOBAL ? __ldg(a) : *a;  // !GLOBAL: a staged copy in shared memory
}
template <bool GLOBAL = true>
__device__ __forceinline__ void kmers4_at(const uint32_t* __restrict__ w, uint32_t q, uint32_t k, bool canonical, uint32_t keys[4]) {
    const uint32_t* a = w + (q >> 4);
    const uint32_t sh = (q & 15) * 2;
    const uint32_t wa = be32(packed_word<GLOBAL>(a)), wb = be32(packed_word<GLOBAL>(a + 1)), wc = be32(packed_word<GLOBAL>(a + 2));
    const uint32_t xh = __funnelshift_l(wb, wa, sh), xl = __funnelshift_l(wc, wb, sh);
    const uint64_t x = ((uint64_t)xh << 32) | xl;
    const uint64_t rc = canonical ? (((uint64_t)revcomp16_word(xl) << 32) | revcomp16_word(xh)) : 0;
    const uint32_t mask = value_mask<uint32_t>(2 * k);
#pragma unroll
    for (uint32_t t = 0; t < 4; ++t) {
        uint32_t key = (uint32_t)(x >> (64 - 2 * k - 2 * t)) & mask;
        if (canonical) key = min(key, (uint32_t)(rc >> (2 * t)) & mask);
        keys[t] = key;
    }
}
template <bool GLOBAL = true>
__device__ __forceinline__ void kmers4_at(const uint32_t* __restrict__ w, uint32_t q, uint32_t k, bool canonical, uint64_t keys[4]) {
    const uint32_t* a = w + (q >> 4);
    const uint32_t sh = (q & 15) * 2;
    const uint32_t w0 = be32(packed_word<GLOBAL>(a)), w1 = be32(packed_word<GLOBAL>(a + 1)), w2 = be32(packed_word<GLOBAL>(a + 2)),
                   w3 = be32(packed_word<GLOBAL>(a + 3));
    const uint32_t a0 = __funnelshift_l(w1, w0, sh), a1 = __funnelshift_l(w2, w1, sh), a2 = __funnelshift_l(w3, w2, sh);
    const uint32_t r0 = canonical ? revcomp16_word(a0) : 0, r1 = canonical ? revcomp16_word(a1) : 0, r2 = canonical ? revcomp16_word(a2) : 0;
    const uint64_t mask = value_mask<uint64_t>(2 * k);
#pragma unroll
    for (uint32_t t = 0; t < 4; ++t) {
        const uint64_t x = ((uint64_t)__funnelshift_l(a1, a0, 2 * t) << 32) | __funnelshift_l(a2, a1, 2 * t);
        uint64_t key = x >> (64 - 2 * k);
        if (canonical) {
            const uint64_t y = (((uint64_t)__funnelshift_r(r1, r2, 2 * t) << 32) | __funnelshift_r(r0, r1, 2 * t)) & mask;
            key = key < y ? key : y;
        }
        keys[t] = key;
    }
}

// group_bytes: the slice of the packed buffer one warp takes (a multiple of 16; GROUP_BYTES unless the input is so
// small that slices of that size would leave most of the GPU idle).  The warp hands 128 consecutive positions of one
// sequence at a time to its lanes, four consecutive positions per lane (kmers4_at); f is called four times per turn.
template <typename V, typename F>
__device__ __forceinline__ void warp_for_each_kmer(const SeqView& b, uint64_t total_bytes, uint64_t group, uint32_t k,
                                                   bool canonical, int lane, F&& f, uint32_t group_bytes = GROUP_BYTES) {
    const uint64_t byte0 = group * group_bytes;
    const uint64_t byte1 = min(byte0 + (uint64_t)group_bytes, total_bytes);
    uint64_t s = seq_of_byte_warp(b.byte_off, b.nseq, byte0);
    while (s < b.nseq) {
        const uint64_t sb = __ldg(b.byte_off + s);
        if (sb >= byte1) break;
        const uint64_t L = __ldg(b.nbases + s);
        const uint64_t nk = L >= k ? L - k + 1 : 0;
        const uint64_t p_lo = byte0 > sb ? (byte0 - sb) * 4 : 0;
        const uint64_t p_hi = min(nk, (byte1 - sb) * 4);
        ++s;
        if (p_lo >= p_hi) continue;
        const uint32_t n = (uint32_t)(p_hi - p_lo);
        const uint32_t* wbase = (const uint32_t*)(b.packed + sb) + (p_lo >> 4);
        const uint32_t q0 = (uint32_t)p_lo & 15;
        for (uint32_t r0 = 0; r0 < n; r0 += 128) {
            const uint32_t r = r0 + 4 * (uint32_t)lane;
            V keys[4] = {0, 0, 0, 0};
            if (r < n) kmers4_at(wbase, q0 + r, k, canonical, keys);
#pragma unroll
            for (uint32_t t = 0; t < 4; ++t) f(keys[t], r + t < n);
        }
    }
}

}  // namespace kmu
