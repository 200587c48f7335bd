// ============================================================================
//  oracle/kmer_oracle.cpp  --  TEST INFRASTRUCTURE ONLY (see kmer_oracle.hpp).
//  CPU restatement of the kmerutils hot path; every function cites the
//  reference file:line it follows.  Built by oracle/Makefile into
//  oracle/libkmer_oracle.so with -ffp-contract=off (Rust never fuses a*b+c).
// ============================================================================
#include "kmer_oracle.hpp"

#include "det_math.hpp"
#include "zig_exp_tables.h"

#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

// ---------------------------------------------------------------- alphabet ----
// Alphabet2b::encode, src/base/alphabet.rs:119-127 (case-insensitive; A0 C1 G2 T3)
// to_ascii_uppercase only touches a-z
inline int encode2b_strict(uint8_t c) {
    if (c >= 'a' && c <= 'z') c = (uint8_t)(c - 32);
    switch (c) {
        case 'A': return 0;
        case 'C': return 1;
        case 'G': return 2;
        case 'T': return 3;
        default: return -1;
    }
}
const char DECODE2B[4] = {'A', 'C', 'G', 'T'};

// IterSequence::next reduced to its meaning: base `pos` of the packed sequence,
// first base of a byte in its two most significant bits (sequence.rs:605-648,
// alphabet.rs:162-168)
inline uint8_t base_at(const uint8_t* packed, uint64_t pos) {
    return (packed[pos >> 2] >> (6 - 2 * (pos & 3))) & 3;
}

// ---------------------------------------------------------------- k-mer words ---
inline uint64_t value_mask(int nbits) { return nbits >= 64 ? ~0ULL : ((1ULL << nbits) - 1); }

inline uint32_t brev32(uint32_t x) {
    x = ((x & 0x55555555u) << 1) | ((x >> 1) & 0x55555555u);
    x = ((x & 0x33333333u) << 2) | ((x >> 2) & 0x33333333u);
    x = ((x & 0x0F0F0F0Fu) << 4) | ((x >> 4) & 0x0F0F0F0Fu);
    x = ((x & 0x00FF00FFu) << 8) | ((x >> 8) & 0x00FF00FFu);
    return (x << 16) | (x >> 16);
}
inline uint64_t brev64(uint64_t x) {
    return ((uint64_t)brev32((uint32_t)x) << 32) | brev32((uint32_t)(x >> 32));
}

uint64_t kmer_build(uint64_t value, int k, int type) {
    switch (type) {
        case ORC_KMER32: return (uint32_t)(((uint32_t)k << 28) | (uint32_t)value);  // kmer32bit.rs:212-216
        case ORC_KMER16B32: return (uint32_t)value;                                 // kmer16b32bit.rs:130-135
        default: return value;                                                      // kmer64bit.rs / kmeraa.rs
    }
}

uint64_t kmer_push(uint64_t word, int k, int type, uint8_t base) {
    switch (type) {
        case ORC_KMER32: {  // kmer32bit.rs:98-113
            uint32_t w = (uint32_t)word;
            uint32_t hdr = w & 0xF0000000u;
            uint32_t nb = (w >> 28) & 0xF;
            uint32_t vmask = (1u << (2 * nb)) - 1;
            uint32_t nk = ((w << 2) & vmask) | (base & 3u);
            (void)k;
            return nk | hdr;
        }
        case ORC_KMER16B32:  // kmer16b32bit.rs:57-61
            return (uint32_t)(((uint32_t)word << 2) | (base & 3u));
        case ORC_KMER64: {  // kmer64bit.rs:68-80 ; (1<<64) wraps to mask 0 in release at k=32:
            // we restate the *intended* semantics for k = 32 (full 64-bit window) and say so in DESIGN.md
            uint64_t vmask = value_mask(2 * k);
            return ((word << 2) & vmask) | (uint64_t)(base & 3u);
        }
        default: return 0;
    }
}

uint64_t kmer_revcomp(uint64_t word, int k, int type) {
    switch (type) {
        case ORC_KMER32: {  // kmer32bit.rs:119-137
            uint32_t w = (uint32_t)word;
            uint32_t hdr = w & 0xF0000000u;
            uint32_t nb = (w >> 28) & 0xF;
            uint32_t r = ~w;
            r = brev32(r);
            r = ((r & 0x55555555u) << 1) | ((r & 0xAAAAAAAAu) >> 1);
            uint32_t sh = 32 - 2 * nb;
            r = sh >= 32 ? 0 : (r >> sh);
            r = (r & 0x0FFFFFFFu) | hdr;
            (void)k;
            return r;
        }
        case ORC_KMER16B32: {  // kmer16b32bit.rs:43-54
            uint32_t r = ~(uint32_t)word;
            r = brev32(r);
            r = ((r & 0x55555555u) << 1) | ((r & 0xAAAAAAAAu) >> 1);
            return r;
        }
        case ORC_KMER64: {  // kmer64bit.rs:83-96
            uint64_t r = ~word;
            r = brev64(r);
            r = ((r & 0x5555555555555555ULL) << 1) | ((r & 0xAAAAAAAAAAAAAAAAULL) >> 1);
            int sh = 64 - 2 * k;
            r = sh >= 64 ? 0 : (r >> sh);
            return r;
        }
        default: return word;  // AA k-mers have no reverse complement (kmeraa.rs:185-187 panics)
    }
}

int kmer_cmp(uint64_t a, uint64_t b, int type) {
    if (type == ORC_KMER32) {  // kmer32bit.rs:47-55 : header first, then value
        uint32_t ha = (uint32_t)a & 0xF0000000u, hb = (uint32_t)b & 0xF0000000u;
        if (ha != hb) return ha < hb ? -1 : 1;
        uint32_t va = (uint32_t)a & 0x0FFFFFFFu, vb = (uint32_t)b & 0x0FFFFFFFu;
        return va < vb ? -1 : (va > vb ? 1 : 0);
    }
    // derived Ord on u32 (kmer16b32bit.rs:20) ; (k, value) with equal k (kmer64bit.rs:45-53)
    return a < b ? -1 : (a > b ? 1 : 0);
}

uint64_t kmer_compressed_value(uint64_t word, int type) {
    if (type == ORC_KMER32) return (uint32_t)word & 0x0FFFFFFFu;  // kmer32bit.rs:173-178
    return word;
}

bool kmer_type_accepts(int k, int type) {
    switch (type) {
        case ORC_KMER32: return k >= 1 && k <= 14;   // kmergenerator.rs:311 / kmer32bit.rs:160
        case ORC_KMER16B32: return k == 16;          // kmergenerator.rs:218-220
        case ORC_KMER64: return k >= 1 && k <= 32;   // kmergenerator.rs:415
        case ORC_KMERAA32: return k >= 1 && k <= 6;  // aautils/kmeraa.rs:212-214, 727-732
        case ORC_KMERAA64: return k >= 1 && k <= 12; // aautils/kmeraa.rs:822-824
        default: return false;
    }
}

inline bool is_aa_type(int type) { return type == ORC_KMERAA32 || type == ORC_KMERAA64; }

// aautils Alphabet::encode (kmeraa.rs:85-109): 5-bit codes, 14 is skipped; -1 = not in the alphabet (the reference panics)
inline int encode_aa(uint8_t c) {
    switch (c) {
        case 'A': return 1; case 'C': return 2; case 'D': return 3; case 'E': return 4; case 'F': return 5;
        case 'G': return 6; case 'H': return 7; case 'I': return 8; case 'K': return 9; case 'L': return 10;
        case 'M': return 11; case 'N': return 12; case 'P': return 13; case 'Q': return 15; case 'R': return 16;
        case 'S': return 17; case 'T': return 18; case 'V': return 19; case 'W': return 20; case 'Y': return 21;
        default: return -1;
    }
}
const char AA_LETTERS[21] = "ACDEFGHIKLMNPQRSTVWY";

// KmerAA32bit::push / KmerAA64bit::push (kmeraa.rs:171-182, 301-312): the ASCII residue is encoded here
inline uint64_t kmer_push_aa(uint64_t aa, int k, uint8_t residue) {
    const uint64_t vmask = (1ULL << (5 * k)) - 1;
    return ((aa << 5) & vmask) | ((uint64_t)encode_aa(residue) & 0x1F);
}

// All k-mer words of [begin, end) of one sequence, in order.  DNA: `seq` is 2-bit packed
// (KmerSeqIterator::next, kmergenerator.rs:75-106); amino acids: `seq` is one ASCII residue per byte
// (aautils KmerSeqIterator::next, kmeraa.rs:568-627).  f(word) gets the k-mer word (`.0` / `.aa`).
template <typename F>
void for_each_kmer(const uint8_t* seq, uint64_t nbases, uint64_t begin, uint64_t end, int k, int type, F&& f) {
    if (!kmer_type_accepts(k, type)) return;
    if (end > nbases) end = nbases;
    if (end < begin + (uint64_t)k) return;
    if (is_aa_type(type)) {
        uint64_t v = 0;
        for (int i = 0; i < k; ++i) v = (v << 5) | ((uint64_t)encode_aa(seq[begin + i]) & 0x1F);  // kmeraa.rs:598-621
        f(v);
        for (uint64_t p = begin + k; p < end; ++p) {
            v = kmer_push_aa(v, k, seq[p]);
            f(v);
        }
        return;
    }
    uint64_t val = 0;
    for (int i = 0; i < k - 1; ++i) val = (val << 2) | base_at(seq, begin + i);
    uint64_t word = kmer_build(val, k, type);
    for (uint64_t p = begin + k - 1; p < end; ++p) {
        word = kmer_push(word, k, type, base_at(seq, p));
        f(word);
    }
}

// invhash (probminhash::invhash, Thomas Wang / Heng Li) -- SURVEY App. A.6
inline uint32_t int32_hash(uint32_t key) {
    key += ~(key << 15);
    key ^= (key >> 10);
    key += (key << 3);
    key ^= (key >> 6);
    key += ~(key << 11);
    key ^= (key >> 16);
    return key;
}
inline uint64_t int64_hash(uint64_t key) {
    key = (~key) + (key << 21);
    key = key ^ (key >> 24);
    key = (key + (key << 3)) + (key << 8);
    key = key ^ (key >> 14);
    key = (key + (key << 2)) + (key << 4);
    key = key ^ (key >> 28);
    key = key + (key << 31);
    return key;
}

inline bool is_u32_type(int type) { return type == ORC_KMER32 || type == ORC_KMER16B32 || type == ORC_KMERAA32; }

uint64_t apply_hash(uint64_t word, int k, int type, int kind) {
    switch (kind) {
        case ORC_HASH_IDENTITY_RAW: return word;
        case ORC_HASH_MASKED_VALUE: {
            int bits = (type == ORC_KMERAA32 || type == ORC_KMERAA64) ? 5 * k : 2 * k;
            return kmer_compressed_value(word, type) & value_mask(bits);
        }
        case ORC_HASH_CANON_INVHASH:
        case ORC_HASH_CANON_RAW: {
            // kmer.reverse_complement().min(*kmer)  (datasketcher.rs:223)
            uint64_t rc = kmer_revcomp(word, k, type);
            uint64_t canon = kmer_cmp(word, rc, type) < 0 ? word : rc;
            if (kind == ORC_HASH_CANON_RAW) return canon;
            return is_u32_type(type) ? (uint64_t)int32_hash((uint32_t)canon) : int64_hash(canon);
        }
        case ORC_HASH_INVHASH: return is_u32_type(type) ? (uint64_t)int32_hash((uint32_t)word) : int64_hash(word);
        default: return word;
    }
}

// ---------------------------------------------------------------- ntHash -------
// seeds nthash.rs:17-20 ; BASE_MAPPING_2B nthash.rs:28-30
const uint64_t NT_SEED[4] = {0x3c8bfbb395c60474ULL, 0x3193c18562a02b4cULL, 0x20323ed082572324ULL,
                             0x295549f54be24456ULL};
inline uint64_t rotl64(uint64_t x, unsigned r) { r &= 63; return r ? (x << r) | (x >> (64 - r)) : x; }
inline uint64_t rotr64(uint64_t x, unsigned r) { r &= 63; return r ? (x >> r) | (x << (64 - r)) : x; }
inline uint8_t kmer_base(uint64_t word, int k, int type, int i) {
    // i-th base from the left of the k-mer (value right aligned)
    (void)type;
    return (uint8_t)((word >> (2 * (k - 1 - i))) & 3);
}

// ---------------------------------------------------------------- RNG ----------
// rand_xoshiro::SplitMix64 + Xoshiro256PlusPlus::seed_from_u64
inline uint64_t splitmix64_next(uint64_t& x) {
    x += 0x9E3779B97F4A7C15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
struct Xoshiro256pp {
    uint64_t s[4];
    explicit Xoshiro256pp(uint64_t seed) {
        uint64_t x = seed;
        for (int i = 0; i < 4; ++i) s[i] = splitmix64_next(x);
    }
    inline uint64_t next_u64() {
        uint64_t r = rotl64(s[0] + s[3], 23) + s[0];
        uint64_t t = s[1] << 17;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl64(s[3], 45);
        return r;
    }
    inline uint32_t next_u32() { return (uint32_t)(next_u64() >> 32); }
    // rand 0.9 Uniform::<f64>::new(0.,1.).sample : 52 random mantissa bits (App. A.2)
    inline double unif01() {
        uint64_t bits = (next_u64() >> 12) | 0x3FF0000000000000ULL;
        double v;
        std::memcpy(&v, &bits, 8);
        return v - 1.0;
    }
    inline float unif01_f32() {
        uint32_t bits = (next_u32() >> 9) | 0x3F800000u;
        float v;
        std::memcpy(&v, &bits, 4);
        return v - 1.0f;
    }
    // rand 0.9 UniformUsize::sample for a range that fits in u32: Lemire widening
    // multiply on next_u32 with rejection threshold (2^32 - range) % range (App. A.2)
    inline uint32_t unif_range_u32(uint32_t low, uint32_t range) {
        uint32_t thresh = (uint32_t)(0u - range) % range;
        for (;;) {
            uint64_t prod = (uint64_t)next_u32() * (uint64_t)range;
            uint32_t lo = (uint32_t)prod;
            if (lo >= thresh) return low + (uint32_t)(prod >> 32);
        }
    }
};

// NoHashHasher (src/nohasher.rs:22-48): the key's native-endian bytes read big-endian
inline uint64_t bswap32(uint32_t v) { return __builtin_bswap32(v); }
uint64_t nohash_seed(uint64_t key, int key_bytes) {
    return key_bytes == 4 ? (uint64_t)__builtin_bswap32((uint32_t)key) : __builtin_bswap64(key);
}
// fnv::FnvHasher over the native-endian key bytes (seqsketchjaccard.rs:346-349)
uint64_t fnv1a_seed(uint64_t key, int key_bytes) {
    uint64_t h = 0xcbf29ce484222325ULL;
    for (int i = 0; i < key_bytes; ++i) {
        h ^= (key >> (8 * i)) & 0xFF;
        h *= 0x100000001b3ULL;
    }
    return h;
}

// ---------------------------------------------------------------- ProbMinHash3a --
// probminhash::probminhasher::ExpRestricted01
struct ExpRestricted01 {
    double lambda, c1, c2, c3;
    explicit ExpRestricted01(double l) : lambda(l) {
        c1 = std::expm1(lambda) / lambda;
        c2 = std::log(2.0 / (1.0 + std::exp(-lambda))) / lambda;
        c3 = (1.0 - std::exp(-lambda)) / lambda;
    }
    inline double sample(Xoshiro256pp& rng) const {
        double x = c1 * rng.unif01();
        if (x < 1.0) return x;
        for (;;) {
            x = rng.unif01();
            if (x < c2) return x;
            double y = 0.5 * rng.unif01();
            if (y > 1.0 - x) {
                x = 1.0 - x;
                y = 1.0 - y;
            }
            if (x <= c3 * (1.0 - y)) return x;
            if (c1 * y <= 1.0 - x) return x;
            if (y * c1 * lambda <= det_expm1(lambda * (1.0 - x))) return x;  // the same expm1 as the kernels (det_math.hpp)
        }
    }
};

// MaxValueTracker: slot values + running maximum (tournament tree; root == qmax)
struct MaxTracker {
    uint32_t m;
    std::vector<double> v;  // 2m-1 nodes, leaves first
    explicit MaxTracker(uint32_t m_) : m(m_), v(2 * (size_t)m_ - 1, std::numeric_limits<double>::max()) {}
    inline double max_value() const { return v.back(); }
    inline double at(uint32_t k) const { return v[k]; }
    void update(uint32_t k, double value) {
        size_t last = v.size() - 1;
        size_t cur = k;
        double curv = value;
        if (!(curv < v[cur])) return;
        for (;;) {
            v[cur] = curv;
            size_t p = m + cur / 2;
            if (p > last) break;
            size_t sib = cur ^ 1;
            if (v[sib] >= v[p] && v[cur] >= v[p]) break;
            if (curv < v[sib]) curv = v[sib];
            cur = p;
            if (curv >= v[cur]) break;
        }
    }
};

struct PendingItem {
    uint64_t key;
    double winv;
    Xoshiro256pp rng;
};

// ProbMinHash3a::hash_weigthed_hashmap. Items are visited in
// the order given (ascending key: the reference's hashbrown order is not
// reproducible and only matters for exact f64 ties).
void pmh3a(const uint64_t* keys, const double* weights, uint64_t n, uint32_t m, int key_bytes, uint64_t* sig) {
    for (uint32_t j = 0; j < m; ++j) sig[j] = 0;  // Val::default()
    if (m < 2) return;                            // reference asserts m >= 2
    const double lambda = std::log((double)m / (double)(m - 1));
    ExpRestricted01 exp01(lambda);
    MaxTracker q(m);
    std::vector<PendingItem> todo;
    double qmax = q.max_value();
    for (uint64_t it = 0; it < n; ++it) {
        const uint64_t key = keys[it];
        const double winv = 1.0 / weights[it];
        Xoshiro256pp rng(nohash_seed(key, key_bytes));
        const double h = winv * exp01.sample(rng);
        qmax = q.max_value();
        if (h < qmax) {
            uint32_t k = rng.unif_range_u32(0, m);
            if (h < q.at(k)) {
                sig[k] = key;
                q.update(k, h);
                qmax = q.max_value();
            }
            if (winv < qmax) todo.push_back(PendingItem{key, winv, rng});
        }
    }
    uint64_t i = 2;
    while (!todo.empty()) {
        size_t insert_pos = 0;
        for (size_t j = 0; j < todo.size(); ++j) {
            PendingItem& p = todo[j];
            double h = p.winv * (double)(i - 1);
            if (h < q.max_value()) {
                h = h + p.winv * exp01.sample(p.rng);
                uint32_t k = p.rng.unif_range_u32(0, m);
                if (h < q.at(k)) {
                    sig[k] = p.key;
                    q.update(k, h);
                    qmax = q.max_value();
                }
                if (p.winv * (double)i < qmax) {
                    todo[insert_pos] = p;
                    ++insert_pos;
                }
            }
        }
        todo.resize(insert_pos, PendingItem{0, 0.0, Xoshiro256pp(0)});
        ++i;
    }
}

// FnvHashMap<Val, u64> stand-in (fnv + hashbrown in the reference,
// seqsketchjaccard.rs:226-227): open addressing, FNV-1a over the key bytes,
// count == 0 marks an empty slot.  Iteration order = slot order (the reference's
// hashbrown order is not reproducible and only matters for exact f64 ties).
struct FlatCountMap {
    std::vector<uint64_t> keys;
    std::vector<uint64_t> counts;
    uint64_t mask = 0;
    uint64_t used = 0;
    void reset(uint64_t expected) {
        uint64_t cap = 16;
        while (cap < 2 * expected) cap <<= 1;
        if (keys.size() != cap) {
            keys.assign(cap, 0);
            counts.assign(cap, 0);
        } else {
            std::fill(counts.begin(), counts.end(), 0);
        }
        mask = cap - 1;
        used = 0;
    }
    void grow() {
        std::vector<uint64_t> ok, oc;
        ok.swap(keys);
        oc.swap(counts);
        uint64_t cap = ok.size() * 2;
        keys.assign(cap, 0);
        counts.assign(cap, 0);
        mask = cap - 1;
        for (size_t i = 0; i < ok.size(); ++i)
            if (oc[i]) add(ok[i], oc[i]);
    }
    static inline uint64_t fnv(uint64_t key) {
        uint64_t h = 0xcbf29ce484222325ULL;
        for (int i = 0; i < 8; ++i) {
            h ^= (key >> (8 * i)) & 0xFF;
            h *= 0x100000001b3ULL;
        }
        return h ^ (h >> 32);
    }
    inline void add(uint64_t key, uint64_t c) {
        uint64_t i = fnv(key) & mask;
        for (;;) {
            if (counts[i] == 0) {
                keys[i] = key;
                counts[i] = c;
                if (++used * 2 > keys.size()) grow();
                return;
            }
            if (keys[i] == key) {
                counts[i] += c;
                return;
            }
            i = (i + 1) & mask;
        }
    }
};

// multiplicity map of one sequence range: fhash(kmer) -> count
// (seqsketchjaccard.rs:226-234)
void count_kmers(const uint8_t* packed, uint64_t nbases, uint64_t begin, uint64_t end, int k, int type,
                 int hash_kind, FlatCountMap& wb) {
    for_each_kmer(packed, nbases, begin, end, k, type,
                  [&](uint64_t word) { wb.add(apply_hash(word, k, type, hash_kind), 1); });
}

void sketch_from_map(const FlatCountMap& wb, uint32_t m, int key_bytes, uint64_t* sig) {
    std::vector<uint64_t> keys;
    std::vector<double> w;
    keys.reserve(wb.used);
    w.reserve(wb.used);
    for (size_t i = 0; i < wb.keys.size(); ++i)
        if (wb.counts[i]) {
            keys.push_back(wb.keys[i]);
            w.push_back((double)wb.counts[i]);
        }
    pmh3a(keys.data(), w.data(), keys.size(), m, key_bytes, sig);
}

// ---------------------------------------------------------------- SuperMinHash ----
// probminhash::superminhasher::SuperMinHash<F, T, H> (Ertl 2017), restated from the
// published algorithm.  hsketch starts at F::from(u32::MAX) ("large"), q = -1, b[m-1] = m, a = m-1.
template <typename S>
struct SuperMinHashOrc {
    uint32_t m;
    std::vector<S> h;
    std::vector<int64_t> q;
    std::vector<uint32_t> p;
    std::vector<int64_t> b;
    int64_t item_rank = 0;
    uint32_t a_upper;
    explicit SuperMinHashOrc(uint32_t m_) : m(m_), h(m_, (S)4294967295.0), q(m_, -1), p(m_, 0), b(m_, 0), a_upper(m_ - 1) {
        b[m - 1] = m;
    }
    static inline S unif01(Xoshiro256pp& rng);
    void sketch(uint64_t seed) {
        Xoshiro256pp rng(seed);
        const int64_t irank = item_rank;
        uint32_t j = 0;
        while (j <= a_upper) {
            const S r = unif01(rng);
            const uint32_t k = rng.unif_range_u32(j, m - j);  // Uniform::<usize>::new(j, m)
            if (q[j] != irank) {
                q[j] = irank;
                p[j] = j;
            }
            if (q[k] != irank) {
                q[k] = irank;
                p[k] = k;
            }
            std::swap(p[j], p[k]);
            const S rpj = r + (S)j;
            if (rpj < h[p[j]]) {
                const double cur = (double)h[p[j]];
                const uint32_t j2 = cur >= (double)(m - 1) ? m - 1 : (uint32_t)cur;
                h[p[j]] = rpj;
                if (j < j2) {
                    b[j2] -= 1;
                    b[j] += 1;
                    while (b[a_upper] == 0) --a_upper;
                }
            }
            ++j;
        }
        ++item_rank;
    }
};
template <>
inline double SuperMinHashOrc<double>::unif01(Xoshiro256pp& rng) { return rng.unif01(); }
template <>
inline float SuperMinHashOrc<float>::unif01(Xoshiro256pp& rng) { return rng.unif01_f32(); }

template <typename S>
void superminhash_seqs(const uint8_t* packed, const uint64_t* byte_off, const uint64_t* nbases, uint64_t nseq, int k,
                       int type, int hash_kind, uint32_t m, int hasher, S* out) {
    SuperMinHashOrc<S> smh(m);
    const int key_bytes = is_u32_type(type) ? 4 : 8;
    for (uint64_t s = 0; s < nseq; ++s)
        for_each_kmer(packed + byte_off[s], nbases[s], 0, nbases[s], k, type, [&](uint64_t word) {
            const uint64_t key = apply_hash(word, k, type, hash_kind);
            smh.sketch(hasher == 0 ? nohash_seed(key, key_bytes) : fnv1a_seed(key, key_bytes));
        });
    for (uint32_t j = 0; j < m; ++j) out[j] = smh.h[j];
}

// ---------------------------------------------------------------- SetSketch -------
// rand_distr 0.5 Exp1 (ziggurat; tables regenerated by scripts/gen_ziggurat_tables.py) on top of
// rand 0.9's StandardUniform f64 (53 bits, multiply method).
const double ZIG_X[257] = ZIG_EXP_TABLE_X;
const double ZIG_F[257] = ZIG_EXP_TABLE_F;
inline double std_uniform_f64(Xoshiro256pp& rng) { return (double)(rng.next_u64() >> 11) * (1.0 / 9007199254740992.0); }
inline double exp1_sample(Xoshiro256pp& rng) {
    for (;;) {
        const uint64_t bits = rng.next_u64();
        const unsigned i = (unsigned)(bits & 0xff);
        const double u = detmath_from_bits((bits >> 12) | 0x3FF0000000000000ULL) - (1.0 - 2.220446049250313e-16 / 2.0);
        const double x = u * ZIG_X[i];
        if (x < ZIG_X[i + 1]) return x;
        if (i == 0) return ZIG_EXP_R - det_log(std_uniform_f64(rng));
        if (ZIG_F[i + 1] + (ZIG_F[i] - ZIG_F[i + 1]) * std_uniform_f64(rng) < det_exp(-x)) return x;
    }
}

// probminhash::setsketcher::SetSketcher (Ertl 2021, SetSketch1) with its FYshuffle
// (incremental Fisher-Yates driven by Uniform<f64>).  The permutation is reset lazily (stamps) --
// same values as the crate's O(m) reset.
struct SetSketchOrc {
    double b, a, lnb;
    uint64_t m, q;
    std::vector<uint64_t> kvec;
    double lower_k = 0.0;
    uint64_t nbmin = 0;
    std::vector<uint32_t> v, stamp;
    uint32_t cur_stamp = 0;
    SetSketchOrc(double b_, uint64_t m_, double a_, uint64_t q_)
        : b(b_), a(a_), lnb(det_log(b_)), m(m_), q(q_), kvec(m_, 0), v(m_, 0), stamp(m_, 0) {}
    inline uint32_t& perm(uint64_t i) {
        if (stamp[i] != cur_stamp) {
            stamp[i] = cur_stamp;
            v[i] = (uint32_t)i;
        }
        return v[i];
    }
    void sketch(uint64_t seed) {
        Xoshiro256pp rng(seed);
        ++cur_stamp;  // permut_generator.reset()
        uint64_t lastidx = 0;
        const int32_t iq1 = (int32_t)q + 1;
        const double inva = 1.0 / a;
        double x_pred = 0.0;
        for (uint64_t j = 0; j < m; ++j) {
            const double x_j = x_pred + (inva / (double)(m - j)) * exp1_sample(rng);
            x_pred = x_j;
            const double lb = det_log(x_j) / lnb;
            if (lb > -lower_k) break;
            const double fl = std::floor(1.0 - lb);
            const int32_t z = fl >= 2147483647.0 ? 2147483647 : (fl <= -2147483648.0 ? (int32_t)(-2147483647 - 1) : (int32_t)fl);
            const int32_t k = std::max(0, std::min(iq1, z));
            if ((double)k <= lower_k) break;
            // FYshuffle::next
            const double xsi = rng.unif01();
            const uint64_t idx = lastidx + (uint64_t)(xsi * (double)(m - lastidx));
            const uint32_t val = perm(idx);
            perm(idx) = perm(lastidx);
            perm(lastidx) = val;
            ++lastidx;
            if ((double)k > (double)kvec[val]) {
                kvec[val] = (uint64_t)k;
                ++nbmin;
                if (nbmin % m == 0) {
                    const double flow = (double)*std::min_element(kvec.begin(), kvec.end());
                    if (flow > lower_k) lower_k = flow;
                }
            }
        }
    }
};

}  // namespace

// ================================================================= C API =========
extern "C" {

int64_t orc_pack_2bit(const uint8_t* ascii, uint64_t n, uint8_t* out) {
    // Sequence::new(raw, 2)  sequence.rs:25-106 ; tail padded with 'A' (00) :66-71
    uint64_t nbytes = (n + 3) / 4;
    for (uint64_t b = 0; b < nbytes; ++b) {
        uint8_t packed = 0;
        for (int i = 0; i < 4; ++i) {
            uint64_t p = 4 * b + i;
            int code = 0;
            if (p < n) {
                code = encode2b_strict(ascii[p]);
                if (code < 0) return -1;
            }
            packed |= (uint8_t)(code << (6 - 2 * i));
        }
        out[b] = packed;
    }
    return (int64_t)nbytes;
}

uint64_t orc_encode_and_add_2bit(const uint8_t* ascii, uint64_t n, uint8_t* out) {
    // Sequence::encode_and_add  sequence.rs:388-451 : invalid characters are skipped
    uint64_t kept = 0;
    for (uint64_t i = 0; i < n; ++i) {
        int code = encode2b_strict(ascii[i]);
        if (code < 0) continue;
        if ((kept & 3) == 0) out[kept >> 2] = 0;
        out[kept >> 2] |= (uint8_t)(code << (6 - 2 * (kept & 3)));
        ++kept;
    }
    return kept;
}

uint64_t orc_count_non_acgt(const uint8_t* ascii, uint64_t n) {  // alphabet.rs:28-31
    uint64_t c = 0;
    for (uint64_t i = 0; i < n; ++i) c += encode2b_strict(ascii[i]) < 0;
    return c;
}

uint8_t orc_get_base(const uint8_t* packed, uint64_t pos) { return base_at(packed, pos); }

void orc_unpack_2bit(const uint8_t* packed, uint64_t nbases, uint8_t* ascii_out) {
    for (uint64_t i = 0; i < nbases; ++i) ascii_out[i] = (uint8_t)DECODE2B[base_at(packed, i)];
}

void orc_seq_revcomp_2bit(const uint8_t* packed, uint64_t nbases, uint8_t* out) {
    // meaning of get_reverse_complement_2bitseq (sequence.rs:252-295): base i of the
    // result is the complement of base n-1-i; same description (tail zero padded... the
    // reference leaves complemented padding bits in the tail; we reproduce that below)
    uint64_t len = (nbases + 3) / 4;
    unsigned inlast = (unsigned)(nbases & 3);
    unsigned shift_amount = inlast ? 8 - 2 * inlast : 0;
    unsigned shift_mask = inlast ? (1u << shift_amount) - 1 : 0;
    for (uint64_t i = 0; i < len; ++i) {
        uint8_t byte = packed[len - 1 - i];
        uint8_t rev = byte;
        if (inlast) {
            rev = (uint8_t)(rev >> shift_amount);
            if (i < len - 1) rev |= (uint8_t)((packed[len - 1 - i - 1] & shift_mask) << (8 - shift_amount));
        }
        rev = (uint8_t)(((rev & 0x33) << 2) | ((rev & 0xCC) >> 2));
        rev = (uint8_t)(((rev & 0x0F) << 4) | ((rev & 0xF0) >> 4));
        rev = (uint8_t)~rev;
        out[i] = rev;
    }
}

uint64_t orc_kmer_build(uint64_t value, int k, int type) { return kmer_build(value, k, type); }
uint64_t orc_kmer_push(uint64_t word, int k, int type, uint8_t base) { return kmer_push(word, k, type, base); }
uint64_t orc_kmer_revcomp(uint64_t word, int k, int type) { return kmer_revcomp(word, k, type); }
int orc_kmer_cmp(uint64_t a, uint64_t b, int k, int type) {
    (void)k;
    return kmer_cmp(a, b, type);
}
uint64_t orc_kmer_compressed_value(uint64_t word, int k, int type) {
    (void)k;
    return kmer_compressed_value(word, type);
}

uint64_t orc_generate_kmers(const uint8_t* packed, uint64_t nbases, uint64_t begin, uint64_t end, int k, int type,
                            uint64_t* out) {
    // KmerSeqIterator::next  kmergenerator.rs:75-106 ; range semantics sequence.rs:562-585
    if (!kmer_type_accepts(k, type)) return ~0ULL;
    if (end > nbases || end <= begin) return 0;  // set_range returns Err -> callers unwrap/panic
    if (end - begin < (uint64_t)k) return 0;
    if (is_aa_type(type)) {
        for (uint64_t i = begin; i < end; ++i)
            if (encode_aa(packed[i]) < 0) return ~0ULL - 1;  // Alphabet::encode panics (kmeraa.rs:106)
        uint64_t na = 0;
        for_each_kmer(packed, nbases, begin, end, k, type, [&](uint64_t w) { out[na++] = w; });
        return na;
    }
    uint64_t val = 0;
    for (int i = 0; i < k; ++i) val = (val << 2) | base_at(packed, begin + i);  // first k-mer via KmerBuilder::build
    uint64_t word = kmer_build(val, k, type);
    uint64_t n = 0;
    out[n++] = word;
    for (uint64_t p = begin + k; p < end; ++p) {
        word = kmer_push(word, k, type, base_at(packed, p));
        out[n++] = word;
    }
    return n;
}

uint64_t orc_apply_hash(uint64_t word, int k, int type, int hash_kind) { return apply_hash(word, k, type, hash_kind); }
uint32_t orc_int32_hash(uint32_t key) { return int32_hash(key); }
uint64_t orc_int64_hash(uint64_t key) { return int64_hash(key); }

uint64_t orc_nthash_init(uint64_t word, int k, int type) {  // kmer.rs:48-61
    uint64_t h = 0;
    for (int i = 0; i < k; ++i) h ^= rotl64(NT_SEED[kmer_base(word, k, type, i)], (unsigned)(k - i - 1));
    return h;
}

int orc_nthash_canonical_init(uint64_t word, int k, int type, uint64_t* fhash, uint64_t* rhash, uint64_t* canon) {
    // kmer.rs:74-94
    uint64_t f = 0, r = 0;
    for (int i = 0; i < k; ++i) {
        uint8_t b = kmer_base(word, k, type, i);
        f ^= rotl64(NT_SEED[b], (unsigned)(k - i - 1));
        r ^= rotl64(NT_SEED[3 - b], (unsigned)i);
    }
    *fhash = f;
    *rhash = r;
    if (f <= r) {
        *canon = f;
        return 0;
    }
    *canon = r;
    return 1;
}

void orc_nthash_mult(uint64_t h0, int k, uint64_t* hashed, int n) {  // nthash.rs:63-72 (wrapping)
    if (n <= 0) return;
    hashed[0] = h0;
    for (int i = 1; i < n; ++i) {
        uint64_t t = h0 * ((uint64_t)i ^ ((uint64_t)k * 0x90b45d39fb6da1faULL));
        t ^= t >> 27;
        hashed[i] = t;
    }
}

uint64_t orc_nthash_cycle(uint64_t word, int k, int type, uint64_t hashval, uint8_t new_base) {
    // kmer.rs:63-71 : old_base = leftmost base of self; push result is discarded
    uint8_t old_base = kmer_base(word, k, type, 0);
    return rotl64(hashval, 1) ^ rotl64(NT_SEED[old_base], (unsigned)k) ^ NT_SEED[new_base & 3];
}

int orc_nthash_canonical_cycle(uint64_t word, int k, int type, uint8_t new_base, uint64_t* fhash, uint64_t* rhash,
                               uint64_t* canon) {
    // kmer.rs:96-117 : fhash/rhash are zeroed first (bug-compatible)
    uint8_t old_base = kmer_base(word, k, type, 0);
    uint64_t f = 0, r = 0;
    f = rotl64(f, 1) ^ rotl64(NT_SEED[old_base], (unsigned)k) ^ NT_SEED[new_base & 3];
    r = rotr64(r, 1) ^ rotl64(NT_SEED[3 - old_base], (unsigned)k) ^ rotl64(NT_SEED[3 - (new_base & 3)], (unsigned)(k - 1));
    *fhash = f;
    *rhash = r;
    if (f <= r) {
        *canon = f;
        return 0;
    }
    *canon = r;
    return 1;
}

uint64_t orc_nohash_seed(uint64_t key, int key_bytes) { return nohash_seed(key, key_bytes); }
uint64_t orc_fnv1a_seed(uint64_t key, int key_bytes) { return fnv1a_seed(key, key_bytes); }

void orc_xoshiro_seed(uint64_t seed, uint64_t s[4]) {
    Xoshiro256pp r(seed);
    std::memcpy(s, r.s, 32);
}
uint64_t orc_xoshiro_next(uint64_t s[4]) {
    Xoshiro256pp r(0);
    std::memcpy(r.s, s, 32);
    uint64_t v = r.next_u64();
    std::memcpy(s, r.s, 32);
    return v;
}

void orc_pmh3a_weighted(const uint64_t* keys, const double* weights, uint64_t n, uint32_t m, int key_bytes,
                        uint64_t* sig) {
    pmh3a(keys, weights, n, m, key_bytes, sig);
}

void orc_sketch_pmh3a_seq(const uint8_t* packed, uint64_t nbases, int k, int type, int hash_kind, uint32_t m,
                          uint64_t* sig) {
    FlatCountMap wb;
    wb.reset(std::min<uint64_t>(nbases, 1ULL << 26));  // get_nbkmer_guess, kmergenerator.rs:207-211
    count_kmers(packed, nbases, 0, nbases, k, type, hash_kind, wb);
    sketch_from_map(wb, m, is_u32_type(type) ? 4 : 8, sig);
}

void orc_sketch_pmh3a_batch(const uint8_t* packed, const uint64_t* byte_off, const uint64_t* nbases, uint64_t nseq,
                            int k, int type, int hash_kind, uint32_t m, void* sig_out, int sig_bytes, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    std::atomic<uint64_t> next(0);
    auto worker = [&]() {
        std::vector<uint64_t> sig(m);
        for (;;) {
            uint64_t i = next.fetch_add(1);
            if (i >= nseq) break;
            orc_sketch_pmh3a_seq(packed + byte_off[i], nbases[i], k, type, hash_kind, m, sig.data());
            if (sig_bytes == 4) {
                uint32_t* o = (uint32_t*)sig_out + i * (uint64_t)m;
                for (uint32_t j = 0; j < m; ++j) o[j] = (uint32_t)sig[j];
            } else {
                std::memcpy((uint64_t*)sig_out + i * (uint64_t)m, sig.data(), 8 * (size_t)m);
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
}

void orc_sketch_pmh3a_seqs(const uint8_t* packed, const uint64_t* byte_off, const uint64_t* nbases, uint64_t nseq, int k,
                           int type, int hash_kind, uint32_t m, uint64_t* sig) {
    FlatCountMap wb;
    wb.reset(1024);
    for (uint64_t i = 0; i < nseq; ++i) count_kmers(packed + byte_off[i], nbases[i], 0, nbases[i], k, type, hash_kind, wb);
    sketch_from_map(wb, m, is_u32_type(type) ? 4 : 8, sig);
}

uint64_t orc_blocksketch_seq(const uint8_t* packed, uint64_t nbases, int k, uint32_t m, uint64_t block_size,
                             uint32_t* sig_out, uint64_t max_blocks) {
    // BlockSeqSketcher::blocksketch_sequence (seqblocksketch.rs:97-149): the number of
    // blocks comes from BASES, the blocks themselves are runs of block_size K-MERS of
    // one continuous iterator; trailing blocks may be empty (signature all zero).
    // fhash = canonical + int32_hash on Kmer32bit (seqblocksketch.rs:462-466)
    if (nbases == 0 || block_size == 0) return 0;
    uint64_t nb_blocks = nbases % block_size == 0 ? nbases / block_size : 1 + nbases / block_size;
    uint64_t nkmers = nbases >= (uint64_t)k ? nbases - k + 1 : 0;
    std::vector<uint64_t> sig(m);
    for (uint64_t b = 0; b < nb_blocks && b < max_blocks; ++b) {
        FlatCountMap wa;
        wa.reset(block_size);
        uint64_t first = b * block_size;  // first k-mer index of the block
        if (first < nkmers) {
            uint64_t cnt = std::min(block_size, nkmers - first);
            count_kmers(packed, nbases, first, first + cnt + k - 1, k, ORC_KMER32, ORC_HASH_CANON_INVHASH, wa);
        }
        sketch_from_map(wa, m, 4, sig.data());
        for (uint32_t j = 0; j < m; ++j) sig_out[b * (uint64_t)m + j] = (uint32_t)sig[j];
    }
    return nb_blocks;
}

double orc_jaccard_equal_fraction(const void* a, const void* b, uint32_t m, int sig_bytes) {
    // compute_probminhash_jaccard (seqsketchjaccard.rs:86-108)
    uint32_t eq = 0;
    for (uint32_t i = 0; i < m; ++i)
        eq += std::memcmp((const uint8_t*)a + (size_t)i * sig_bytes, (const uint8_t*)b + (size_t)i * sig_bytes, sig_bytes) == 0;
    return (double)eq / (double)m;
}

// counter-based SplitMix64: output number i (0-based) of stream `seed`
static inline uint64_t synth_z(uint64_t seed, uint64_t i) {
    uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

void orc_synth_packed(uint64_t seed, uint64_t first_base, uint64_t nbases, uint8_t* packed_out) {
    uint64_t nbytes = (nbases + 3) / 4;
    for (uint64_t b = 0; b < nbytes; ++b) {
        uint8_t v = 0;
        for (int i = 0; i < 4; ++i) {
            uint64_t p = 4 * b + i;
            unsigned code = p < nbases ? (unsigned)(synth_z(seed, first_base + p) >> 62) : 0u;
            v |= (uint8_t)(code << (6 - 2 * i));
        }
        packed_out[b] = v;
    }
}

void orc_synth_ascii(uint64_t seed, uint64_t first_base, uint64_t nbases, uint8_t* ascii_out) {
    for (uint64_t p = 0; p < nbases; ++p) ascii_out[p] = (uint8_t)DECODE2B[synth_z(seed, first_base + p) >> 62];
}

// ---------------------------------------------------------------- counting (A15) ----
// Exact restatement of KmerCounter's meaning (kmercount.rs:241-288) with zero filter false
// positives: the key is kmer.get_compressed_value() of the canonical k-mer
// (kmer.reverse_complement().min(kmer), kmercount.rs:313,827,938); get_count = 0 if never
// inserted, 1 if inserted once, else min(multiplicity, 2^nb_bits - 1) (counting Bloom saturation,
// test kmercount.rs:1615 expects 255 for 8 bits); nb_distinct = #keys, nb_unique = #keys seen once.
uint64_t orc_count_kmers(const uint8_t* packed, const uint64_t* byte_off, const uint64_t* nbases, uint64_t nseq, int k,
                         int type, int canonical, uint64_t* keys_out, uint64_t* counts_out, uint64_t cap) {
    if (!kmer_type_accepts(k, type)) return ~0ULL;
    std::vector<uint64_t> all;
    for (uint64_t s = 0; s < nseq; ++s) {
        const uint8_t* p = packed + byte_off[s];
        const uint64_t L = nbases[s];
        if (L < (uint64_t)k) continue;
        uint64_t val = 0;
        for (int i = 0; i < k - 1; ++i) val = (val << 2) | base_at(p, i);
        uint64_t word = kmer_build(val, k, type);
        for (uint64_t q = k - 1; q < L; ++q) {
            word = kmer_push(word, k, type, base_at(p, q));
            uint64_t w = canonical ? apply_hash(word, k, type, ORC_HASH_CANON_RAW) : word;
            all.push_back(kmer_compressed_value(w, type));
        }
    }
    std::sort(all.begin(), all.end());
    uint64_t n = 0;
    for (size_t i = 0; i < all.size();) {
        size_t j = i;
        while (j < all.size() && all[j] == all[i]) ++j;
        if (n < cap) {
            keys_out[n] = all[i];
            counts_out[n] = j - i;
        }
        ++n;
        i = j;
    }
    return n;
}

// DispatchableT::dispatch (kmercount.rs:382-420): owner of a compressed k-mer value among nb_receiver
uint64_t orc_dispatch(uint64_t compressed_value, int type, uint64_t nb_receiver) {
    if (is_u32_type(type)) return (uint64_t)(int32_hash((uint32_t)compressed_value) % (uint32_t)nb_receiver);
    return int64_hash(compressed_value) % nb_receiver;
}

// SuperMinHash of one group of sequences (nseq = 1: SeqSketcher::sketch_superminhash /
// SuperHashSketch::sketch_compressedkmer per sequence, seqsketchjaccard.rs:328-380, setsketchert.rs:255-296;
// nseq > 1: SuperHashSketch::sketch_compressedkmer_seqs, setsketchert.rs:299-335).
// hasher: 0 = NoHashHasher (setsketchert.rs:267-269), 1 = fnv::FnvHasher (seqsketchjaccard.rs:346-349)
// sig_bytes: 4 = f32, 8 = f64
void orc_sketch_superminhash(const uint8_t* packed, const uint64_t* byte_off, const uint64_t* nbases, uint64_t nseq, int k,
                             int type, int hash_kind, uint32_t m, int hasher, int sig_bytes, void* out) {
    if (sig_bytes == 4) superminhash_seqs<float>(packed, byte_off, nbases, nseq, k, type, hash_kind, m, hasher, (float*)out);
    else superminhash_seqs<double>(packed, byte_off, nbases, nseq, k, type, hash_kind, m, hasher, (double*)out);
}

void orc_sketch_superminhash_batch(const uint8_t* packed, const uint64_t* byte_off, const uint64_t* nbases, uint64_t nseq,
                                   int k, int type, int hash_kind, uint32_t m, int hasher, int sig_bytes, void* out,
                                   int nthreads) {
    if (nthreads < 1) nthreads = 1;
    std::atomic<uint64_t> next(0);
    auto worker = [&]() {
        for (;;) {
            uint64_t i = next.fetch_add(1);
            if (i >= nseq) break;
            orc_sketch_superminhash(packed, byte_off + i, nbases + i, 1, k, type, hash_kind, m, hasher, sig_bytes,
                                    (uint8_t*)out + i * (uint64_t)m * sig_bytes);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
}

// synthetic protein: residue i of stream `seed` = "ACDEFGHIKLMNPQRSTVWY"[z_i % 20]
void orc_synth_aa(uint64_t seed, uint64_t first_res, uint64_t nres, uint8_t* ascii_out) {
    for (uint64_t p = 0; p < nres; ++p) ascii_out[p] = (uint8_t)AA_LETTERS[synth_z(seed, first_res + p) % 20];
}

// SequenceAA::new_filtered (kmeraa.rs:447-456): keeps the residues of the alphabet; returns how many
uint64_t orc_aa_filter(const uint8_t* ascii, uint64_t n, uint8_t* out) {
    uint64_t kept = 0;
    for (uint64_t i = 0; i < n; ++i)
        if (encode_aa(ascii[i]) >= 0) out[kept++] = ascii[i];
    return kept;
}

// SetSketch of one group of sequences (nseq = 1: HyperLogLogSketch::sketch_compressedkmer per sequence,
// setsketchert.rs:758-802; nseq > 1: sketch_compressedkmer_seqs(_block), :677-724, 811-895 -- the block split
// + merge of the reference is an element-wise max and gives the same registers).  Keys are hashed by
// NoHashHasher (:702-704).  sig_bytes 2 / 4 / 8 = u16 / u32 / u64 registers.
void orc_sketch_setsketch(const uint8_t* packed, const uint64_t* byte_off, const uint64_t* nbases, uint64_t nseq, int k,
                          int type, int hash_kind, double b, uint64_t m, double a, uint64_t q, int sig_bytes, void* out) {
    SetSketchOrc ss(b, m, a, q);
    const int key_bytes = is_u32_type(type) ? 4 : 8;
    for (uint64_t s = 0; s < nseq; ++s)
        for_each_kmer(packed + byte_off[s], nbases[s], 0, nbases[s], k, type, [&](uint64_t word) {
            ss.sketch(nohash_seed(apply_hash(word, k, type, hash_kind), key_bytes));
        });
    for (uint64_t j = 0; j < m; ++j) {
        if (sig_bytes == 2) ((uint16_t*)out)[j] = (uint16_t)ss.kvec[j];
        else if (sig_bytes == 4) ((uint32_t*)out)[j] = (uint32_t)ss.kvec[j];
        else ((uint64_t*)out)[j] = ss.kvec[j];
    }
}

void orc_sketch_setsketch_batch(const uint8_t* packed, const uint64_t* byte_off, const uint64_t* nbases, uint64_t nseq,
                                int k, int type, int hash_kind, double b, uint64_t m, double a, uint64_t q, int sig_bytes,
                                void* out, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    std::atomic<uint64_t> next(0);
    auto worker = [&]() {
        for (;;) {
            uint64_t i = next.fetch_add(1);
            if (i >= nseq) break;
            orc_sketch_setsketch(packed, byte_off + i, nbases + i, 1, k, type, hash_kind, b, m, a, q, sig_bytes,
                                 (uint8_t*)out + i * m * (uint64_t)sig_bytes);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
}

double orc_det_log(double x) { return det_log(x); }
double orc_det_exp(double x) { return det_exp(x); }
double orc_det_expm1(double x) { return det_expm1(x); }

// How often does the deterministic ln / exp / expm1 of det_math.hpp lead to a DIFFERENT RESULT than the platform libm
// the Rust reference calls (f64::ln / exp / exp_m1 -> glibc on Linux)?  Draws the arguments exactly as the sketchers
// draw them and evaluates both at every call:
//   SetSketch (setsketcher.rs: lb = ln(x_j) / ln(b), k = floor(1 - lb)): `points` points of `nkeys` keys with NoHash seeds
//   synth(seed, i); the ziggurat's rare paths (ln(U) in the tail, exp(-x) in the wedge test) are counted too;
//   ExpRestricted01 (ProbMinHash3a, lambda = ln(m / (m - 1))): its last rejection test, y c1 lambda <= expm1(lambda (1 - x)).
// out[0] ln evaluations            out[1] bit patterns differ   out[2] KEYS whose register values k_1 .. k_points differ between an
//                                                                all-deterministic and an all-libm evaluation of the key
// out[3] ziggurat ln/exp evaluations out[4] bit patterns differ out[5] wedge accept decisions that differ
// out[6] expm1 evaluations         out[7] bit patterns differ   out[8] accept decision differs
void orc_libm_divergence(uint64_t nkeys, uint64_t seed, double b, uint64_t m, double a, uint64_t q, uint32_t points, uint32_t m_pmh,
                         int nthreads, uint64_t* out) {
    if (nthreads <= 0) nthreads = (int)std::max(1u, std::thread::hardware_concurrency());
    std::vector<std::array<uint64_t, 9>> acc((size_t)nthreads);
    const double lnb_det = det_log(b), lnb_std = std::log(b), inva = 1.0 / a;
    const int32_t iq1 = (int32_t)q + 1;
    const ExpRestricted01 e01(std::log((double)m_pmh / (double)(m_pmh - 1)));
    auto regval = [&](double lb) {
        const double fl = std::floor(1.0 - lb);
        const int32_t z = fl >= 2147483647.0 ? 2147483647 : (fl <= -2147483648.0 ? (int32_t)(-2147483647 - 1) : (int32_t)fl);
        return std::max(0, std::min(iq1, z));
    };
    auto work = [&](int t) {
        std::array<uint64_t, 9>& c = acc[(size_t)t];
        c.fill(0);
        for (uint64_t i = (uint64_t)t; i < nkeys; i += (uint64_t)nthreads) {
            const uint64_t key = synth_z(seed, i);
            {  // SetSketch points: the same key through an all-deterministic and an all-libm pipeline
                int32_t kd[64], ks[64];
                const uint32_t np_ = std::min<uint32_t>(std::min<uint64_t>(points, m), 64);
                for (int pass = 0; pass < 2; ++pass) {
                    const bool det = pass == 0;
                    Xoshiro256pp rng(key);
                    double x_pred = 0.0;
                    for (uint32_t j = 0; j < np_; ++j) {
                        double e;
                        for (;;) {  // Exp1 (ziggurat)
                            const uint64_t bits = rng.next_u64();
                            const unsigned zi = (unsigned)(bits & 0xff);
                            const double u = detmath_from_bits((bits >> 12) | 0x3FF0000000000000ULL) - (1.0 - 2.220446049250313e-16 / 2.0);
                            const double x = u * ZIG_X[zi];
                            if (x < ZIG_X[zi + 1]) {
                                e = x;
                                break;
                            }
                            if (zi == 0) {
                                const double uu = std_uniform_f64(rng);
                                if (det) {
                                    ++c[3];
                                    c[4] += detmath_bits(det_log(uu)) != detmath_bits(std::log(uu));
                                }
                                e = ZIG_EXP_R - (det ? det_log(uu) : std::log(uu));
                                break;
                            }
                            const double lhs = ZIG_F[zi + 1] + (ZIG_F[zi] - ZIG_F[zi + 1]) * std_uniform_f64(rng);
                            const double ex = det ? det_exp(-x) : std::exp(-x);
                            if (det) {
                                ++c[3];
                                c[4] += detmath_bits(ex) != detmath_bits(std::exp(-x));
                                c[5] += (lhs < ex) != (lhs < std::exp(-x));
                            }
                            if (lhs < ex) {
                                e = x;
                                break;
                            }
                        }
                        const double x_j = x_pred + (inva / (double)(m - j)) * e;
                        x_pred = x_j;
                        if (det) {
                            ++c[0];
                            c[1] += detmath_bits(det_log(x_j)) != detmath_bits(std::log(x_j));
                        }
                        (det ? kd : ks)[j] = regval(det ? det_log(x_j) / lnb_det : std::log(x_j) / lnb_std);
                        (void)rng.unif01();  // the FYshuffle draw between two points
                    }
                }
                bool same = true;
                for (uint32_t j = 0; j < np_; ++j) same &= kd[j] == ks[j];
                c[2] += !same;
            }
            {  // ExpRestricted01, rejection branch entered with probability 1 - 1/c1
                Xoshiro256pp rng(key ^ 0x5DEECE66DULL);
                for (uint32_t rep = 0; rep < points; ++rep) {
                    double x = e01.c1 * rng.unif01();
                    if (x < 1.0) continue;
                    for (;;) {
                        x = rng.unif01();
                        if (x < e01.c2) break;
                        double y = 0.5 * rng.unif01();
                        if (y > 1.0 - x) {
                            x = 1.0 - x;
                            y = 1.0 - y;
                        }
                        if (x <= e01.c3 * (1.0 - y)) break;
                        if (e01.c1 * y <= 1.0 - x) break;
                        const double arg = e01.lambda * (1.0 - x), lhs = y * e01.c1 * e01.lambda;
                        const double m1 = det_expm1(arg), m2 = std::expm1(arg);
                        ++c[6];
                        c[7] += detmath_bits(m1) != detmath_bits(m2);
                        c[8] += (lhs <= m1) != (lhs <= m2);
                        if (lhs <= m1) break;
                    }
                }
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(work, t);
    for (auto& t : th) t.join();
    for (int j = 0; j < 9; ++j) {
        out[j] = 0;
        for (int t = 0; t < nthreads; ++t) out[j] += acc[(size_t)t][(size_t)j];
    }
}
double orc_exp1_from_seed(uint64_t seed, int skip) {
    Xoshiro256pp rng(seed);
    double v = 0;
    for (int i = 0; i <= skip; ++i) v = exp1_sample(rng);
    return v;
}

// read `r` of the synthetic short-read set, as ASCII: see kmu_seqbatch_sample_reads
void orc_sample_read(const uint8_t* genome_packed, uint64_t glen, uint64_t seed, uint64_t r, uint32_t read_len,
                     uint32_t err_ppm, uint8_t* ascii_out) {
    const uint64_t eseed = seed ^ 0x5bd1e995a5a5a5a5ULL;
    const uint64_t start = synth_z(seed, 2 * r) % (glen - read_len + 1);
    const bool rev = synth_z(seed, 2 * r + 1) & 1;
    for (uint32_t j = 0; j < read_len; ++j) {
        unsigned code = rev ? 3u - base_at(genome_packed, start + read_len - 1 - j) : base_at(genome_packed, start + j);
        const uint64_t e = synth_z(eseed, r * read_len + j);
        if ((uint32_t)(e % 1000000ULL) < err_ppm) code = (code + 1u + (unsigned)((e >> 32) % 3u)) & 3u;
        ascii_out[j] = (uint8_t)DECODE2B[code];
    }
}

int orc_hardware_threads(void) {
    unsigned n = std::thread::hardware_concurrency();
    return n ? (int)n : 1;
}

}  // extern "C"
