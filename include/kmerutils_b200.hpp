// ============================================================================
//  kmerutils_b200.hpp -- C++17 host layer above the C ABI (kmerutils_b200.h).
//
//  The reference (jean-pierreBoth/kmerutils, Rust) is compiled code and its toolchain is absent from
//  this image, so the host side of the boundary is written in C++ and mirrors the reference's public
//  surface for the hot path: same type and method names, same argument meaning, same error
//  behaviour (where the reference panics, these throw kmerutils::Panic).  Header only; everything
//  that computes goes through the extern "C" entry points of libkmerutils_b200.so, i.e. to the
//  CUDA kernels -- there is no CPU implementation behind any sketch / generation / counting call.
//  (The value types Kmer32bit / Kmer16b32bit / Kmer64bit carry the few single-word bit operations
//  of the KmerT trait: they are the *types* of the API, not a data path.)
//
//  Reference item                                                        here
//  --------------------------------------------------------------------  --------------------------------
//  base::sequence::Sequence          src/base/sequence.rs:14-106          kmerutils::base::Sequence
//  base::kmertraits::{KmerT,CompressedKmerT,KmerBuilder} kmertraits.rs    the three k-mer structs below
//  base::kmergenerator::KmerGenerator src/base/kmergenerator.rs:148-186   kmerutils::base::KmerGenerator<T>
//  base::kmergenerator::KmerSeqIterator src/base/kmergenerator.rs:30-107  kmerutils::base::KmerSeqIterator<T> (streams from the GPU)
//  base::nthash::NtHash              src/base/nthash.rs:76-120, kmer.rs:45-145  methods of Kmer32bit / Kmer16b32bit
//  base::kmercount::DispatchableT    src/base/kmercount.rs:382-420        dispatch() of the three k-mer types
//  base::kmercount::{KmerCountT,KmerCounter,KmerCounterPool,
//        count_kmer_threaded_one_to_many} src/base/kmercount.rs:48-98,881 kmerutils::base::KmerCounter<T> ...
//  sketching::seqsketchjaccard::SeqSketcher  seqsketchjaccard.rs:117-414  kmerutils::sketching::SeqSketcher
//  sketching::setsketchert::{SeqSketcherT,ProbHash3aSketch,SuperHashSketch,
//        HyperLogLogSketch}          src/sketching/setsketchert.rs:54-896 same names
//  jaccard_index_probminhash3a       seqsketchjaccard.rs:423-495          same name
//  dump_signatures_block_u32 / SigSketchFileReader  :577-712              same names
//  aautils::kmeraa::{Alphabet,KmerAA32bit,KmerAA64bit,SequenceAA,KmerGenerator},
//  aautils::setsketchert::{SeqSketcher,SeqSketcherAAT,ProbHash3aSketch,SuperHashSketch,HyperLogLogSketch}   kmerutils::aautils::*
//  sketching::seqminhash::sketch_seqrange_superminhash  src/sketching/seqminhash.rs:19-62   same name
//
//  The Rust closure `fhash: Fn(&Kmer) -> Kmer::Val` cannot cross into CUDA: the five closures the
//  reference actually passes are the constants of kmerutils::KmerHash.
// ============================================================================
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <mutex>
#include <optional>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <utility>
#include <vector>

#include "kmerutils_b200.h"

namespace kmerutils {

/// what a Rust `panic!` / `unwrap()` of the reference becomes on this side
struct Panic : std::runtime_error {
    int32_t code;
    Panic(int32_t c, const std::string& what) : std::runtime_error(what), code(c) {}
};

inline void check(int32_t rc, const char* where) {
    if (rc != KMU_OK) throw Panic(rc, std::string(where) + ": " + kmu_last_error());
}

/// One GPU context per process and device (the C ABI serialises calls on a context).
class Context {
  public:
    explicit Context(int device = 0) { check(kmu_ctx_create(device, &ctx_), "kmu_ctx_create"); }
    ~Context() { kmu_ctx_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    kmu_ctx* get() const { return ctx_; }
    uint64_t launch_count() const { return kmu_launch_count(ctx_); }
    static Context& global(int device = 0) {
        static std::mutex mu;
        static std::unique_ptr<Context> one;
        std::lock_guard<std::mutex> g(mu);
        if (!one) one.reset(new Context(device));
        return *one;
    }

  private:
    kmu_ctx* ctx_ = nullptr;
};

/// the `fhash` closures of the reference, by the place they are written
struct KmerHash {
    int32_t kind;
    /// |k| k.0                                            seqsketchjaccard.rs:775
    static constexpr KmerHash identity() { return {KMU_HASH_IDENTITY_RAW}; }
    /// |k| k.get_compressed_value() & mask                setsketchert.rs:1098-1104
    static constexpr KmerHash masked_value() { return {KMU_HASH_MASKED_VALUE}; }
    /// |k| intNN_hash(k.reverse_complement().min(*k).0)   datasketcher.rs:222-226
    static constexpr KmerHash canonical_invhash() { return {KMU_HASH_CANON_INVHASH}; }
    /// |k| k.reverse_complement().min(*k).0               kmercount.rs:313
    static constexpr KmerHash canonical() { return {KMU_HASH_CANON_RAW}; }
    /// |k| intNN_hash(k.0)                                minhash.rs:226
    static constexpr KmerHash invhash() { return {KMU_HASH_INVHASH}; }
};

namespace base {

// ---------------------------------------------------------------- k-mer value types
inline uint32_t swap_pairs32(uint32_t v) { return ((v & 0x55555555u) << 1) | ((v & 0xAAAAAAAAu) >> 1); }
inline uint64_t swap_pairs64(uint64_t v) { return ((v & 0x5555555555555555ull) << 1) | ((v & 0xAAAAAAAAAAAAAAAAull) >> 1); }
inline uint32_t bitrev32(uint32_t v) {
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
    return __builtin_bswap32(v);
}
inline uint64_t bitrev64(uint64_t v) { return ((uint64_t)bitrev32((uint32_t)v) << 32) | bitrev32((uint32_t)(v >> 32)); }

/// probminhash::invhash::int32_hash / int64_hash (Thomas Wang's invertible mixes), used by DispatchableT::dispatch
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

/// ntHash on 2-bit k-mers: seeds and multi-hash expansion (src/base/nthash.rs:10-30, 63-72)
namespace nthash {
constexpr uint64_t MULTISEED = 0x90b45d39fb6da1faull;
constexpr unsigned MULTISHIFT = 27;
// BASE_MAPPING_2B: A C G T, then the complements T G C A (offset OFFSET_COMP_2B = 4)
constexpr uint64_t BASE_MAPPING_2B[8] = {0x3c8bfbb395c60474ull, 0x3193c18562a02b4cull, 0x20323ed082572324ull, 0x295549f54be24456ull,
                                         0x295549f54be24456ull, 0x20323ed082572324ull, 0x3193c18562a02b4cull, 0x3c8bfbb395c60474ull};
inline uint64_t rotl(uint64_t x, unsigned r) { return (r &= 63) ? (x << r) | (x >> (64 - r)) : x; }
inline uint64_t rotr(uint64_t x, unsigned r) { return (r &= 63) ? (x >> r) | (x << (64 - r)) : x; }
/// from_one_hash_val_to_mult_hash (nthash.rs:63-72; release builds wrap on overflow)
inline void from_one_hash_val_to_mult_hash(uint64_t ksize, std::vector<uint64_t>& hashed) {
    for (size_t i = 1; i < hashed.size(); ++i) {
        uint64_t t = hashed[0] * ((uint64_t)i ^ (ksize * MULTISEED));
        t ^= t >> MULTISHIFT;
        hashed[i] = t;
    }
}
}  // namespace nthash

/// The NtHash trait (src/base/nthash.rs:76-120) as the macro implement_nthash_for! writes it for the two u32 k-mer types
/// (src/base/kmer.rs:45-145), bug for bug: the *_cycle methods call `self.push(new_base)` and drop the result, so `self`
/// never advances (:69, :109), and nthash_canonical_cycle zeroes fhash / rhash before rolling (:97-98).  Only the *_init
/// methods are meaningful; the batch form on the GPU (kmu_nthash_canonical) is defined against them.
template <typename Derived>
struct NtHashMethods {
    uint64_t nthash_init() const {
        uint64_t f = 0, r = 0;
        init(f, r);
        return f;
    }
    uint64_t nthash_cycle(uint64_t hashval, uint8_t new_base) {
        const uint32_t ksize = self().get_nb_base();
        return nthash::rotl(hashval, 1) ^ nthash::rotl(nthash::BASE_MAPPING_2B[old_base()], ksize) ^ nthash::BASE_MAPPING_2B[new_base & 3];
    }
    std::pair<uint64_t, uint8_t> nthash_canonical_init(uint64_t& fhash, uint64_t& rhash) const {
        init(fhash, rhash);
        return fhash <= rhash ? std::make_pair(fhash, (uint8_t)0) : std::make_pair(rhash, (uint8_t)1);
    }
    std::pair<uint64_t, uint8_t> nthash_canonical_cycle(uint8_t new_base, uint64_t& fhash, uint64_t& rhash) {
        fhash = 0;
        rhash = 0;
        const uint32_t ksize = self().get_nb_base();
        const uint8_t ob = old_base();
        fhash = nthash::rotl(fhash, 1) ^ nthash::rotl(nthash::BASE_MAPPING_2B[ob], ksize) ^ nthash::BASE_MAPPING_2B[new_base & 3];
        rhash = nthash::rotr(rhash, 1) ^ nthash::rotl(nthash::BASE_MAPPING_2B[4 + ob], ksize) ^
                nthash::rotl(nthash::BASE_MAPPING_2B[4 + (new_base & 3)], ksize - 1);
        return fhash <= rhash ? std::make_pair(fhash, (uint8_t)0) : std::make_pair(rhash, (uint8_t)1);
    }
    uint8_t nthash_mult_canonical_init(uint64_t& fhash, uint64_t& rhash, std::vector<uint64_t>& hashed) const {
        const auto res = nthash_canonical_init(fhash, rhash);
        hashed[0] = res.first;
        nthash::from_one_hash_val_to_mult_hash(self().get_nb_base(), hashed);
        return res.second;
    }
    uint8_t nthash_mult_canonical_cycle(uint8_t new_base, uint64_t& fhash, uint64_t& rhash, std::vector<uint64_t>& hashed) {
        const auto res = nthash_canonical_cycle(new_base, fhash, rhash);
        hashed[0] = res.first;
        nthash::from_one_hash_val_to_mult_hash(self().get_nb_base(), hashed);
        return res.second;
    }

  private:
    const Derived& self() const { return *static_cast<const Derived*>(this); }
    uint8_t old_base() const {  // leftmost base of the k-mer
        const uint32_t k = self().get_nb_base();
        return (uint8_t)((self().v >> (2 * (k - 1))) & 3u);
    }
    void init(uint64_t& fhash, uint64_t& rhash) const {
        fhash = 0;
        rhash = 0;
        const uint32_t k = self().get_nb_base();
        for (uint32_t i = 0; i < k; ++i) {
            const uint32_t base = (self().v >> (2 * (k - 1 - i))) & 3u;
            fhash ^= nthash::rotl(nthash::BASE_MAPPING_2B[base], k - i - 1);
            rhash ^= nthash::rotl(nthash::BASE_MAPPING_2B[4 + base], i);
        }
    }
};

/// Kmer32bit (src/base/kmer32bit.rs:22): up to 14 bases, the number of bases in the top 4 bits
struct Kmer32bit : NtHashMethods<Kmer32bit> {
    using Val = uint32_t;
    static constexpr int32_t kmu_type = KMU_KMER32;
    uint32_t v = 0;  // the reference's `.0`
    Kmer32bit() = default;
    explicit Kmer32bit(uint8_t nb_bases) {  // Kmer32bit::new
        if (nb_bases >= 15) throw Panic(KMU_EINVAL, "Kmer32bit cannot store more than 14 bases");
        v = (uint32_t)nb_bases << 28;
    }
    static Kmer32bit build(Val val, uint8_t kmer_size) {  // KmerBuilder::build
        Kmer32bit k(kmer_size);
        k.v |= val & 0x0FFFFFFFu;
        return k;
    }
    static Kmer32bit from_word(Val word, uint8_t) {
        Kmer32bit k;
        k.v = word;
        return k;
    }
    static size_t get_nb_base_max() { return 14; }
    uint8_t get_nb_base() const { return (uint8_t)(v >> 28); }
    /// the value field without the length header (kmer32bit.rs:173-178)
    Val get_compressed_value() const { return v & 0x0FFFFFFFu; }
    size_t get_bitsize() const { return 32; }
    /// DispatchableT::dispatch (kmercount.rs:399-408)
    size_t dispatch(size_t nb_receiver) const { return int32_hash(get_compressed_value()) % (uint32_t)nb_receiver; }
    Kmer32bit push(uint8_t b) const {
        const uint32_t mask = (1u << (2 * get_nb_base())) - 1;
        Kmer32bit r;
        r.v = (((v << 2) & mask) | (b & 3u)) | (v & 0xF0000000u);
        return r;
    }
    Kmer32bit reverse_complement() const {
        const uint32_t nb = v >> 28;
        uint32_t r = swap_pairs32(bitrev32(~v));
        r = nb ? r >> (32 - 2 * nb) : 0;
        Kmer32bit out;
        out.v = (r & 0x0FFFFFFFu) | (v & 0xF0000000u);
        return out;
    }
    std::vector<uint8_t> get_uncompressed_kmer() const {
        const int nb = get_nb_base();
        std::vector<uint8_t> s(nb);
        for (int i = 0; i < nb; ++i) s[i] = "ACGT"[(v >> (2 * (nb - 1 - i))) & 3];
        return s;
    }
    // Ord: number of bases first, then the value field (kmer32bit.rs:47-55)
    friend bool operator==(Kmer32bit a, Kmer32bit b) { return a.v == b.v; }
    friend bool operator<(Kmer32bit a, Kmer32bit b) {
        if ((a.v & 0xF0000000u) != (b.v & 0xF0000000u)) return (a.v & 0xF0000000u) < (b.v & 0xF0000000u);
        return (a.v & 0x0FFFFFFFu) < (b.v & 0x0FFFFFFFu);
    }
};

/// Kmer16b32bit (src/base/kmer16b32bit.rs:21): exactly 16 bases in a u32
struct Kmer16b32bit : NtHashMethods<Kmer16b32bit> {
    using Val = uint32_t;
    static constexpr int32_t kmu_type = KMU_KMER16B32;
    uint32_t v = 0;
    static Kmer16b32bit build(Val val, uint8_t kmer_size) {
        if (kmer_size != 16) throw Panic(KMU_EINVAL, "Kmer16b32bit has 16 bases!!");
        return from_word(val, 16);
    }
    static Kmer16b32bit from_word(Val word, uint8_t) {
        Kmer16b32bit k;
        k.v = word;
        return k;
    }
    static size_t get_nb_base_max() { return 16; }
    uint8_t get_nb_base() const { return 16; }
    Val get_compressed_value() const { return v; }
    size_t get_bitsize() const { return 32; }
    /// DispatchableT::dispatch (kmercount.rs:388-397)
    size_t dispatch(size_t nb_receiver) const { return int32_hash(v) % (uint32_t)nb_receiver; }
    Kmer16b32bit push(uint8_t b) const { return from_word((v << 2) | (b & 3u), 16); }
    Kmer16b32bit reverse_complement() const { return from_word(swap_pairs32(bitrev32(~v)), 16); }
    std::vector<uint8_t> get_uncompressed_kmer() const {
        std::vector<uint8_t> s(16);
        for (int i = 0; i < 16; ++i) s[i] = "ACGT"[(v >> (2 * (15 - i))) & 3];
        return s;
    }
    friend bool operator==(Kmer16b32bit a, Kmer16b32bit b) { return a.v == b.v; }
    friend bool operator<(Kmer16b32bit a, Kmer16b32bit b) { return a.v < b.v; }
};

/// Kmer64bit (src/base/kmer64bit.rs:24): up to 32 bases in a u64, the number of bases kept beside it
struct Kmer64bit {
    using Val = uint64_t;
    static constexpr int32_t kmu_type = KMU_KMER64;
    uint64_t v = 0;      // `.0`
    uint8_t nb_base = 0; // `.1`
    Kmer64bit() = default;
    explicit Kmer64bit(uint8_t nb) : v(0), nb_base(nb) {}
    static Kmer64bit build(Val val, uint8_t kmer_size) { return from_word(val, kmer_size); }
    static Kmer64bit from_word(Val word, uint8_t kmer_size) {
        Kmer64bit k(kmer_size);
        k.v = word;
        return k;
    }
    static size_t get_nb_base_max() { return 32; }
    uint8_t get_nb_base() const { return nb_base; }
    Val get_compressed_value() const { return v; }
    size_t get_bitsize() const { return 64; }
    /// DispatchableT::dispatch (kmercount.rs:410-419)
    size_t dispatch(size_t nb_receiver) const { return (size_t)(int64_hash(v) % (uint64_t)nb_receiver); }
    Kmer64bit push(uint8_t b) const {
        const uint64_t mask = nb_base >= 32 ? ~0ull : ((1ull << (2 * nb_base)) - 1);
        return from_word(((v << 2) & mask) | (b & 3u), nb_base);
    }
    Kmer64bit reverse_complement() const {
        const uint64_t r = swap_pairs64(bitrev64(~v));
        return from_word(nb_base ? r >> (64 - 2 * nb_base) : 0, nb_base);
    }
    std::vector<uint8_t> get_uncompressed_kmer() const {
        std::vector<uint8_t> s(nb_base);
        for (int i = 0; i < nb_base; ++i) s[i] = "ACGT"[(v >> (2 * (nb_base - 1 - i))) & 3];
        return s;
    }
    friend bool operator==(Kmer64bit a, Kmer64bit b) { return a.v == b.v && a.nb_base == b.nb_base; }
    friend bool operator<(Kmer64bit a, Kmer64bit b) { return a.nb_base != b.nb_base ? a.nb_base < b.nb_base : a.v < b.v; }
};

// ---------------------------------------------------------------- Sequence
/// Sequence (src/base/sequence.rs:14-20): bases packed 4 per byte, first base in the two most significant bits.
/// `Sequence::new(raw, 2)` (:25-106) packs on the GPU (kmu_seqbatch_from_ascii) and panics on a non-ACGT character
/// (alphabet.rs:125); use Sequence::new_batch to pack many reads with one upload.
class Sequence {
  public:
    Sequence() = default;
    Sequence(const uint8_t* raw, size_t n, uint8_t nb_bits) { *this = std::move(new_batch({std::string((const char*)raw, n)}, nb_bits)[0]); }
    Sequence(const std::string& raw, uint8_t nb_bits) : Sequence((const uint8_t*)raw.data(), raw.size(), nb_bits) {}
    /// a sequence that is already packed (e.g. read back from a batch)
    static Sequence from_packed(std::vector<uint8_t> packed, size_t nb_base) {
        Sequence s;
        s.seq_ = std::move(packed);
        s.nb_base_ = nb_base;
        return s;
    }
    static std::vector<Sequence> new_batch(const std::vector<std::string>& raws, uint8_t nb_bits, bool drop_invalid = false) {
        if (nb_bits != 2) throw Panic(KMU_EINVAL, "only the 2-bit alphabet is on the GPU path");
        std::vector<uint64_t> off(raws.size() + 1, 0);
        for (size_t i = 0; i < raws.size(); ++i) off[i + 1] = off[i] + raws[i].size();
        std::vector<uint8_t> ascii(off.back() + 1);
        for (size_t i = 0; i < raws.size(); ++i) std::memcpy(ascii.data() + off[i], raws[i].data(), raws[i].size());
        kmu_ctx* ctx = Context::global().get();
        kmu_seqbatch* b = nullptr;
        check(kmu_seqbatch_from_ascii(ctx, ascii.data(), off.data(), raws.size(), drop_invalid ? 1 : 0, nullptr, &b),
              "Sequence::new");
        std::vector<uint8_t> packed(kmu_seqbatch_packed_bytes(b));
        std::vector<uint64_t> boff(raws.size()), nb(raws.size());
        const int32_t rc = kmu_seqbatch_download(ctx, b, packed.data(), boff.data(), nb.data());
        kmu_seqbatch_destroy(b);
        check(rc, "Sequence::new");
        std::vector<Sequence> out(raws.size());
        for (size_t i = 0; i < raws.size(); ++i) {
            out[i].nb_base_ = nb[i];
            out[i].seq_.assign(packed.begin() + boff[i], packed.begin() + boff[i] + (nb[i] + 3) / 4);
        }
        return out;
    }
    uint8_t nb_bits_by_base() const { return 2; }
    size_t size() const { return nb_base_; }
    size_t compressed_length() const { return seq_.size(); }
    const std::vector<uint8_t>& packed() const { return seq_; }
    /// 2-bit code of base `pos` (sequence.rs:120-140)
    uint8_t get_base(size_t pos) const {
        if (pos >= nb_base_) throw Panic(KMU_EINVAL, "Sequence::get_base: position beyond the end");
        return (seq_[pos >> 2] >> (6 - 2 * (pos & 3))) & 3;
    }
    /// get_reverse_complement (sequence.rs:298-316)
    Sequence get_reverse_complement() const {
        Sequence r;
        r.nb_base_ = nb_base_;
        r.seq_.assign(seq_.size(), 0);
        for (size_t i = 0; i < nb_base_; ++i) {
            const uint8_t c = 3 - get_base(nb_base_ - 1 - i);
            r.seq_[i >> 2] |= (uint8_t)(c << (6 - 2 * (i & 3)));
        }
        return r;
    }
    std::vector<uint8_t> decompress() const {
        std::vector<uint8_t> s(nb_base_);
        for (size_t i = 0; i < nb_base_; ++i) s[i] = "ACGT"[get_base(i)];
        return s;
    }

  private:
    std::vector<uint8_t> seq_;
    size_t nb_base_ = 0;
};

/// RAII device batch made of `&[&Sequence]`
class DeviceBatch {
  public:
    explicit DeviceBatch(const std::vector<const Sequence*>& vseq) {
        std::vector<const uint8_t*> ptrs(vseq.size());
        std::vector<uint64_t> nb(vseq.size());
        for (size_t i = 0; i < vseq.size(); ++i) {
            ptrs[i] = vseq[i]->packed().data();
            nb[i] = vseq[i]->size();
        }
        check(kmu_seqbatch_from_ptrs(Context::global().get(), ptrs.data(), nb.data(), vseq.size(), &b_), "kmu_seqbatch_from_ptrs");
    }
    /// a sub-range of one sequence (KmerSeqIterator::set_range, kmergenerator.rs:56-65)
    DeviceBatch(const Sequence& s, size_t begin, size_t end) {
        DeviceBatch whole(std::vector<const Sequence*>{&s});
        const uint64_t idx = 0, b = begin, e = end;
        check(kmu_seqbatch_slices(Context::global().get(), whole.get(), &idx, &b, &e, 1, &b_), "kmu_seqbatch_slices");
    }
    ~DeviceBatch() { kmu_seqbatch_destroy(b_); }
    DeviceBatch(const DeviceBatch&) = delete;
    DeviceBatch& operator=(const DeviceBatch&) = delete;
    kmu_seqbatch* get() const { return b_; }

  private:
    kmu_seqbatch* b_ = nullptr;
};

inline std::vector<const Sequence*> as_refs(const std::vector<Sequence>& v) {
    std::vector<const Sequence*> r(v.size());
    for (size_t i = 0; i < v.size(); ++i) r[i] = &v[i];
    return r;
}

// ---------------------------------------------------------------- KmerSeqIterator
/// hash functor so that the k-mer value types key std::unordered_map (the reference derives Hash on them)
struct KmerStdHash {
    template <typename K>
    size_t operator()(const K& k) const {
        return (size_t)int64_hash((uint64_t)k.v ^ ((uint64_t)k.get_nb_base() << 58));
    }
};
/// FnvHashMap<T, u32> of generate_kmer_distribution
template <typename T>
using KmerDistribution = std::unordered_map<T, uint32_t, KmerStdHash>;

/// hashmap_count_to_vec_count (kmergenerator.rs:189-203)
template <typename T>
std::vector<std::pair<T, uint32_t>> hashmap_count_to_vec_count(const KmerDistribution<T>& kmer_distribution) {
    return std::vector<std::pair<T, uint32_t>>(kmer_distribution.begin(), kmer_distribution.end());
}

/// KmerSeqIterator<T> (src/base/kmergenerator.rs:30-107): `new(ksize, &sequence)`, `set_range(begin, end)`, `next()`.
/// The sequence is uploaded once; next() hands out k-mers from a window of FETCH k-mers generated on the GPU
/// (kmu_seqbatch_slices + kmu_generate_kmers), refilled when it runs dry -- a streaming consumer never holds more.
template <typename T>
class KmerSeqIterator {
  public:
    static constexpr size_t FETCH = 1u << 20;
    KmerSeqIterator(uint8_t ksize, const Sequence& sequence)
        : nb_base_(ksize), batch_(std::vector<const Sequence*>{&sequence}), seq_size_(sequence.size()), begin_(0), end_(sequence.size()) {
        if ((size_t)ksize > T::get_nb_base_max())  // kmergenerator.rs:48-53
            throw Panic(KMU_EINVAL, "KmerSeqIterator cannot support so many bases for given kmer type");
        if (seq_size_ == 0) throw Panic(KMU_EINVAL, "attempt to subtract with overflow: IterSequence::new on an empty sequence");  // sequence.rs:531
    }
    /// Result<(), ()> of IterSequence::set_range (sequence.rs:562-585): false = Err(())
    bool set_range(size_t begin, size_t end) {
        if (end <= begin || end > seq_size_) return false;
        begin_ = begin;
        end_ = end;
        pos_ = begin;
        window_.clear();
        wpos_ = 0;
        return true;
    }
    std::optional<T> next() {
        if (wpos_ >= window_.size() && !refill()) return std::nullopt;
        return T::from_word(window_[wpos_++], nb_base_);
    }

  private:
    bool refill() {
        // k-mers starting at pos_ .. : the window covers bases [pos_, pos_ + FETCH + k - 1) clamped to the range
        if (pos_ + nb_base_ > end_) return false;
        const uint64_t idx = 0, b = pos_, e = std::min<uint64_t>(end_, pos_ + FETCH + nb_base_ - 1);
        kmu_ctx* ctx = Context::global().get();
        kmu_seqbatch* part = nullptr;
        check(kmu_seqbatch_slices(ctx, batch_.get(), &idx, &b, &e, 1, &part), "KmerSeqIterator::next");
        window_.assign(kmu_kmer_count(part, nb_base_), 0);
        const int32_t rc = kmu_generate_kmers(ctx, part, nb_base_, T::kmu_type, KMU_HASH_IDENTITY_RAW, window_.data(), nullptr, 0);
        kmu_seqbatch_destroy(part);
        check(rc, "KmerSeqIterator::next");
        wpos_ = 0;
        pos_ += window_.size();
        return !window_.empty();
    }
    uint8_t nb_base_;
    DeviceBatch batch_;
    size_t seq_size_, begin_, end_, pos_ = 0, wpos_ = 0;
    std::vector<typename T::Val> window_;
};

// ---------------------------------------------------------------- KmerGenerator
/// KmerGenerator<T> (src/base/kmergenerator.rs:148-186).  `KmerGenerator::new(ksize)` panics on a size the type
/// cannot hold (:48-53, 218, 311, 415) -- here the first generate call throws.
template <typename T>
class KmerGenerator {
  public:
    explicit KmerGenerator(uint8_t ksize) : kmer_size_(ksize) {}
    size_t get_kmer_size() const { return kmer_size_; }
    std::vector<T> generate_kmer(const Sequence& seq) const {
        DeviceBatch b(std::vector<const Sequence*>{&seq});
        return run(b);
    }
    std::vector<T> generate_kmer_in_range(const Sequence& seq, size_t begin, size_t end) const {
        if (begin >= end || end > seq.size()) throw Panic(KMU_EINVAL, "KmerSeqIterator::set_range failed");
        DeviceBatch b(seq, begin, end);
        return run(b);
    }
    /// generate_weighted_kmer = generate_kmer_distribution (kmergenerator.rs:177-186, per type :262-303, 343-408, 459-526):
    /// the distinct (forward) k-mers of the sequence with their multiplicities -- counted in an exact table on the GPU
    /// (kmu_count_insert_seqs, canonical off) and read back; iteration order of a hash map is unspecified in the reference too.
    KmerDistribution<T> generate_weighted_kmer(const Sequence& seq) const { return generate_kmer_distribution(seq); }
    KmerDistribution<T> generate_kmer_distribution(const Sequence& seq) const {
        DeviceBatch b(std::vector<const Sequence*>{&seq});
        const uint64_t nk = kmu_kmer_count(b.get(), kmer_size_);
        KmerDistribution<T> map;
        kmu_ctx* ctx = Context::global().get();
        kmu_counter* c = nullptr;
        check(kmu_count_create(ctx, kmer_size_, T::kmu_type, 32, std::max<uint64_t>(nk, 16), &c), "generate_kmer_distribution");
        std::vector<typename T::Val> keys(nk + 1);
        std::vector<uint32_t> counts(nk + 1);
        uint64_t n = 0;
        int32_t rc = kmu_count_insert_seqs(ctx, c, b.get(), 0);
        if (rc == KMU_OK) rc = kmu_count_export(ctx, c, 1, keys.data(), counts.data(), nk + 1, &n);
        kmu_count_destroy(c);
        check(rc, "generate_kmer_distribution");
        map.reserve(n);
        for (uint64_t i = 0; i < n; ++i) map.emplace(T::build(keys[i], kmer_size_), counts[i]);
        return map;
    }
    /// all sequences with one upload and one launch (what a caller looping over generate_kmer wants on a GPU)
    std::vector<std::vector<T>> generate_kmer_batch(const std::vector<const Sequence*>& vseq) const {
        DeviceBatch b(vseq);
        std::vector<uint64_t> off(vseq.size() + 1);
        std::vector<typename T::Val> words(kmu_kmer_count(b.get(), kmer_size_));
        check(kmu_generate_kmers(Context::global().get(), b.get(), kmer_size_, T::kmu_type, KMU_HASH_IDENTITY_RAW, words.data(),
                                 off.data(), 0),
              "KmerGenerator::generate_kmer");
        std::vector<std::vector<T>> out(vseq.size());
        for (size_t i = 0; i < vseq.size(); ++i) {
            out[i].reserve(off[i + 1] - off[i]);
            for (uint64_t j = off[i]; j < off[i + 1]; ++j) out[i].push_back(T::from_word(words[j], kmer_size_));
        }
        return out;
    }

  private:
    std::vector<T> run(const DeviceBatch& b) const {
        std::vector<typename T::Val> words(kmu_kmer_count(b.get(), kmer_size_));
        check(kmu_generate_kmers(Context::global().get(), b.get(), kmer_size_, T::kmu_type, KMU_HASH_IDENTITY_RAW, words.data(),
                                 nullptr, 0),
              "KmerGenerator::generate_kmer");
        std::vector<T> out;
        out.reserve(words.size());
        for (auto w : words) out.push_back(T::from_word(w, kmer_size_));
        return out;
    }
    uint8_t kmer_size_;
};

// ---------------------------------------------------------------- counting
/// KmerCountT + KmerCounter (src/base/kmercount.rs:48-98): exact table in HBM instead of the cuckoo + counting Bloom pair
/// (semantics of the reference with zero filter false positives; counts saturate at 2^nb_bits - 1).
/// insert_kmer buffers on the host and flushes in one upload before any query.
template <typename Kmer>
class KmerCounter {
  public:
    /// KmerCounter::new(fpr, capacity, nb_bits) (:88-98); fpr has no meaning for an exact table
    KmerCounter(float /*fpr*/, size_t capacity, size_t nb_bits, uint8_t kmer_size) : k_(kmer_size), nb_bits_((uint8_t)nb_bits) {
        check(kmu_count_create(Context::global().get(), kmer_size, Kmer::kmu_type, (uint32_t)nb_bits, capacity, &c_), "KmerCounter::new");
    }
    ~KmerCounter() { kmu_count_destroy(c_); }
    KmerCounter(const KmerCounter&) = delete;
    KmerCounter& operator=(const KmerCounter&) = delete;
    uint8_t get_count_nb_bits() const { return nb_bits_; }
    void insert_kmer(Kmer kmer) {
        pending_.push_back(kmer.get_compressed_value());
        if (pending_.size() >= (1u << 20)) flush();
    }
    /// every k-mer of every sequence, canonical as count_kmer does (:313)
    void insert_sequences(const std::vector<const Sequence*>& vseq, bool canonical = true) {
        flush();
        DeviceBatch b(vseq);
        check(kmu_count_insert_seqs(Context::global().get(), c_, b.get(), canonical ? 1 : 0), "count_kmer");
    }
    uint32_t get_count(Kmer kmer) {
        flush();
        const typename Kmer::Val key = kmer.get_compressed_value();
        uint32_t cnt = 0;
        check(kmu_count_query(Context::global().get(), c_, &key, 1, &cnt, 0), "KmerCounter::get_count");
        return cnt;
    }
    std::vector<uint32_t> get_counts(const std::vector<Kmer>& kmers) {
        flush();
        std::vector<typename Kmer::Val> keys(kmers.size());
        for (size_t i = 0; i < kmers.size(); ++i) keys[i] = kmers[i].get_compressed_value();
        std::vector<uint32_t> cnt(kmers.size());
        check(kmu_count_query(Context::global().get(), c_, keys.data(), keys.size(), cnt.data(), 0), "KmerCounter::get_count");
        return cnt;
    }
    /// multiplicity if the k-mer was seen at least twice, else 0 (:100-107)
    uint32_t get_above2_count(Kmer kmer) {
        const uint32_t c = get_count(kmer);
        return c >= 2 ? c : 0;
    }
    uint64_t get_nb_distinct() { return stat(0); }
    uint64_t get_nb_unique() { return stat(1); }
    kmu_counter* handle() {
        flush();
        return c_;
    }

  private:
    void flush() {
        if (pending_.empty()) return;
        check(kmu_count_insert_kmers(Context::global().get(), c_, pending_.data(), pending_.size(), 0), "KmerCounter::insert_kmer");
        pending_.clear();
    }
    uint64_t stat(int which) {
        flush();
        uint64_t d = 0, u = 0;
        check(kmu_count_stats(Context::global().get(), c_, &d, &u, nullptr, nullptr), "KmerCounter stats");
        return which ? u : d;
    }
    kmu_counter* c_ = nullptr;
    uint8_t k_, nb_bits_;
    std::vector<typename Kmer::Val> pending_;
};

/// KmerCounterPool (kmercount.rs:424-565): one counter per "thread", a k-mer lives in counter `kmer.dispatch(n)`
/// (DispatchableT, :382-420).  Here every counter is a table in HBM; the pool keeps the reference's layout so that code
/// written against `pool.counters[i]` / `get_above2_count` / `get_count_nb_bits` keeps working.
template <typename Kmer>
class KmerCounterPool {
  public:
    std::vector<std::unique_ptr<KmerCounter<Kmer>>> counters;
    explicit KmerCounterPool(std::vector<std::unique_ptr<KmerCounter<Kmer>>> c) : counters(std::move(c)) {}
    uint32_t get_above2_count(Kmer kmer) { return counters[kmer.dispatch(counters.size())]->get_above2_count(kmer); }
    uint8_t get_count_nb_bits() const { return counters.empty() ? 0 : counters[0]->get_count_nb_bits(); }
    // KmerCountT for the pool (:533-563)
    void insert_kmer(Kmer kmer) { counters[kmer.dispatch(counters.size())]->insert_kmer(kmer); }
    uint32_t get_count(Kmer kmer) { return counters[kmer.dispatch(counters.size())]->get_count(kmer); }
    /// many k-mers with one query per counter (what a caller looping over get_count wants on a GPU)
    std::vector<uint32_t> get_counts(const std::vector<Kmer>& kmers) {
        std::vector<std::vector<Kmer>> per(counters.size());
        std::vector<std::vector<size_t>> idx(counters.size());
        for (size_t i = 0; i < kmers.size(); ++i) {
            const size_t loc = kmers[i].dispatch(counters.size());
            per[loc].push_back(kmers[i]);
            idx[loc].push_back(i);
        }
        std::vector<uint32_t> out(kmers.size(), 0);
        for (size_t loc = 0; loc < counters.size(); ++loc) {
            if (per[loc].empty()) continue;
            const auto c = counters[loc]->get_counts(per[loc]);
            for (size_t j = 0; j < c.size(); ++j) out[idx[loc][j]] = c[j];
        }
        return out;
    }
    uint64_t get_nb_distinct() {
        uint64_t n = 0;
        for (auto& c : counters) n += c->get_nb_distinct();
        return n;
    }
    uint64_t get_nb_unique() {
        uint64_t n = 0;
        for (auto& c : counters) n += c->get_nb_unique();
        return n;
    }
};

/// the common body of the two threaded drivers: the canonical k-mers of all sequences are bucketed by
/// DispatchableT::dispatch on the GPU (kmu_count_partition) and bucket i is inserted into counter i
template <typename Kmer>
std::unique_ptr<KmerCounterPool<Kmer>> count_kmer_into_pool(const std::vector<Sequence>& seqvec, size_t nb_threads, size_t nb_bits,
                                                            size_t kmer_size) {
    if (nb_threads < 1 || nb_threads > 64) throw Panic(KMU_EINVAL, "nb_threads must be in 1..64");
    kmu_ctx* ctx = Context::global().get();
    DeviceBatch b(as_refs(seqvec));
    const uint64_t nk = kmu_kmer_count(b.get(), (uint32_t)kmer_size);
    std::vector<std::unique_ptr<KmerCounter<Kmer>>> counters;
    for (size_t i = 0; i < nb_threads; ++i)  // the reference sizes every filter for 1e9 / 3e9 keys (:888-892); a table is sized by its input
        counters.emplace_back(new KmerCounter<Kmer>(0.03f, (size_t)(nk / nb_threads * 1.3) + 1024, nb_bits, (uint8_t)kmer_size));
    if (nk) {
        void* dev = nullptr;
        uint8_t handle[64];
        check(kmu_ipc_alloc(ctx, nk * sizeof(typename Kmer::Val), &dev, handle), "count_kmer: device buffer");
        std::vector<uint64_t> part(nb_threads, 0);
        int32_t rc = kmu_count_partition(ctx, b.get(), (uint32_t)kmer_size, Kmer::kmu_type, 1, (uint32_t)nb_threads, dev, part.data(), 1);
        uint64_t off = 0;
        for (size_t i = 0; i < nb_threads && rc == KMU_OK; ++i) {
            rc = kmu_count_insert_kmers(ctx, counters[i]->handle(), (const uint8_t*)dev + off * sizeof(typename Kmer::Val), part[i], 1);
            off += part[i];
        }
        kmu_ipc_free(ctx, dev);
        check(rc, "count_kmer");
    }
    return std::make_unique<KmerCounterPool<Kmer>>(std::move(counters));
}

/// count_kmer_threaded_one_to_many (kmercount.rs:881-974): one producer, nb_threads counters, `count_size` BITS per count
/// (:893), canonical k-mers (:938), counter = dispatch (:941)
template <typename Kmer>
std::unique_ptr<KmerCounterPool<Kmer>> count_kmer_threaded_one_to_many(const std::vector<Sequence>& seqvec, size_t nb_threads,
                                                                       size_t count_size, size_t kmer_size) {
    return count_kmer_into_pool<Kmer>(seqvec, nb_threads, count_size, kmer_size);
}

/// count_kmer_thread_independant (kmercount.rs:797-867): every thread walks all k-mers and keeps those it owns; 8-bit counts
/// (:802).  Same pool as the one-to-many driver.
template <typename Kmer>
std::unique_ptr<KmerCounterPool<Kmer>> count_kmer_thread_independant(const std::vector<Sequence>& seqvec, size_t nb_threads,
                                                                     size_t kmer_size) {
    return count_kmer_into_pool<Kmer>(seqvec, nb_threads, 8, kmer_size);
}

}  // namespace base

namespace sketching {

using base::DeviceBatch;
using base::Sequence;

/// SeqSketcherParams (src/sketching/sketcharg.rs analogue used by setsketchert.rs:93)
struct SeqSketcherParams {
    size_t kmer_size;
    size_t sketch_size;
    size_t get_kmer_size() const { return kmer_size; }
    size_t get_sketch_size() const { return sketch_size; }
};

template <typename Val>
inline std::vector<std::vector<Val>> rows_of(const std::vector<Val>& flat, size_t nrows, size_t m) {
    std::vector<std::vector<Val>> out(nrows);
    for (size_t i = 0; i < nrows; ++i) out[i].assign(flat.begin() + i * m, flat.begin() + (i + 1) * m);
    return out;
}

/// the reference unwraps KmerSeqIterator::set_range(0, size) on every sequence: an empty sequence panics
/// (seqsketchjaccard.rs:230)
inline void reject_empty(const std::vector<const Sequence*>& vseq) {
    for (const Sequence* s : vseq)
        if (s->size() == 0) throw Panic(KMU_EINVAL, "called `Result::unwrap()` on an `Err` value: set_range on an empty sequence");
}

/// SeqSketcher (src/sketching/seqsketchjaccard.rs:117-414)
class SeqSketcher {
  public:
    SeqSketcher(size_t kmer_size, size_t sketch_size) : kmer_size_(kmer_size), sketch_size_(sketch_size) {}
    size_t get_kmer_size() const { return kmer_size_; }
    size_t get_sketch_size() const { return sketch_size_; }

    /// sketch_probminhash3a (:211-260): one signature of sketch_size hashed k-mers per sequence, input order
    template <typename Kmer>
    std::vector<std::vector<typename Kmer::Val>> sketch_probminhash3a(const std::vector<const Sequence*>& vseq, KmerHash fhash) const {
        reject_empty(vseq);
        DeviceBatch b(vseq);
        std::vector<typename Kmer::Val> flat(vseq.size() * sketch_size_);
        check(kmu_sketch_pmh3a(Context::global().get(), b.get(), (uint32_t)kmer_size_, Kmer::kmu_type, fhash.kind, (uint32_t)sketch_size_,
                               flat.data(), 0),
              "sketch_probminhash3a");
        return rows_of(flat, vseq.size(), sketch_size_);
    }
    /// sketch_superminhash (:328-380): S = float or double, FNV-hashed keys (:346-349)
    template <typename Kmer, typename S>
    std::vector<std::vector<S>> sketch_superminhash(const std::vector<const Sequence*>& vseq, KmerHash fhash) const {
        static_assert(std::is_same<S, float>::value || std::is_same<S, double>::value, "S is f32 or f64");
        reject_empty(vseq);
        DeviceBatch b(vseq);
        std::vector<S> flat(vseq.size() * sketch_size_);
        check(kmu_sketch_superminhash(Context::global().get(), b.get(), (uint32_t)kmer_size_, Kmer::kmu_type, fhash.kind,
                                      (uint32_t)sketch_size_, KMU_HASHER_FNV, (int32_t)sizeof(S), flat.data(), 0),
              "sketch_superminhash");
        return rows_of(flat, vseq.size(), sketch_size_);
    }
    /// create_signature_dump (:390-414): header of the signature file
    kmu_sigdump* create_signature_dump(const std::string& dumpfname) const {
        kmu_sigdump* d = nullptr;
        check(kmu_sigdump_create(dumpfname.c_str(), (uint32_t)sketch_size_, (uint32_t)kmer_size_, &d), "create_signature_dump");
        return d;
    }

  private:
    size_t kmer_size_, sketch_size_;
};

/// dump_signatures_block_u32 (seqsketchjaccard.rs:577-585)
inline void dump_signatures_block_u32(const std::vector<std::vector<uint32_t>>& signatures, kmu_sigdump* out) {
    for (const auto& s : signatures) check(kmu_sigdump_write(out, s.data(), 1), "dump_signatures_block_u32");
}

/// SigSketchFileReader (seqsketchjaccard.rs:588-712)
class SigSketchFileReader {
  public:
    explicit SigSketchFileReader(const std::string& fname) : fname_(fname) {
        check(kmu_sigdump_read(fname.c_str(), &sig_size_, &sketch_size_, &kmer_size_, &nsig_, nullptr, 0, 0), "SigSketchFileReader::new");
    }
    uint8_t get_kmer_size() const { return (uint8_t)kmer_size_; }
    size_t get_signature_length() const { return sketch_size_; }
    size_t get_signature_size() const { return sig_size_; }
    std::optional<std::vector<uint32_t>> next() {
        if (pos_ >= nsig_) return std::nullopt;
        std::vector<uint32_t> sig(sketch_size_);
        check(kmu_sigdump_read(fname_.c_str(), nullptr, nullptr, nullptr, nullptr, sig.data(), pos_++, 1), "SigSketchFileReader::next");
        return sig;
    }

  private:
    std::string fname_;
    uint32_t sig_size_ = 0, sketch_size_ = 0, kmer_size_ = 0;
    uint64_t nsig_ = 0, pos_ = 0;
};

/// fraction of equal slots: compute_probminhash_jaccard / probminhash_get_jaccard_objects (seqsketchjaccard.rs:86-108)
template <typename D>
double compute_probminhash_jaccard(const std::vector<D>& siga, const std::vector<D>& sigb) {
    if (siga.size() != sigb.size()) throw Panic(KMU_EINVAL, "signatures of different sizes");
    double j = 0;
    check(kmu_signature_jaccard(Context::global().get(), siga.data(), 1, sigb.data(), 1, (uint32_t)siga.size(), (int32_t)sizeof(D), &j, 0),
          "compute_probminhash_jaccard");
    return j;
}

/// jaccard_index_probminhash3a (seqsketchjaccard.rs:423-495): J(seqa, b) for every b of vseqb
template <typename Kmer>
std::vector<double> jaccard_index_probminhash3a(const Sequence& seqa, const std::vector<Sequence>& vseqb, size_t sketch_size,
                                                size_t kmer_size, KmerHash fhash) {
    std::vector<const Sequence*> all{&seqa};
    for (const Sequence& b : vseqb) all.push_back(&b);
    const auto sigs = SeqSketcher(kmer_size, sketch_size).template sketch_probminhash3a<Kmer>(all, fhash);
    std::vector<typename Kmer::Val> flat;
    for (size_t i = 1; i < sigs.size(); ++i) flat.insert(flat.end(), sigs[i].begin(), sigs[i].end());
    std::vector<double> j(vseqb.size());
    if (!vseqb.empty())
        check(kmu_signature_jaccard(Context::global().get(), sigs[0].data(), 1, flat.data(), vseqb.size(), (uint32_t)sketch_size,
                                    (int32_t)sizeof(typename Kmer::Val), j.data(), 0),
              "jaccard_index_probminhash3a");
    return j;
}

/// BlockSketched / BlockSketchedSeq / BlockSeqSketcher / DistBlockSketched (src/sketching/seqblocksketch.rs:36-167, 419-440):
/// every sequence is cut into runs of block_size consecutive k-mers (Kmer32bit), one ProbMinHash3a signature per block.
/// The number of blocks comes from the BASES (ceil(size / block_size)), so trailing blocks may hold no k-mer at all.
struct BlockSketched {
    uint32_t numseq = 0, numblock = 0;
    std::vector<uint32_t> sketch;
    const std::vector<uint32_t>& get_skech_slice() const { return sketch; }
};
struct BlockSketchedSeq {
    size_t numseq = 0;
    std::vector<std::vector<BlockSketched>> sketch;  // one vector of length 1 per block, as the reference keeps them for hnsw_rs
};
class BlockSeqSketcher {
  public:
    BlockSeqSketcher(size_t block_size, size_t kmer_size, size_t sketch_size) : block_size_(block_size), kmer_size_(kmer_size), sketch_size_(sketch_size) {}
    /// a pack of (numseq, sequence): all blocks of all sequences in one upload and one sketch call
    std::vector<BlockSketchedSeq> blocksketch_sequences(const std::vector<std::pair<uint32_t, const Sequence*>>& pack_seq, KmerHash fhash) const {
        std::vector<const Sequence*> vseq;
        std::vector<uint64_t> idx, begin, end;
        std::vector<size_t> nblocks;
        for (size_t s = 0; s < pack_seq.size(); ++s) {
            const Sequence* seq = pack_seq[s].second;
            if (seq->size() == 0) throw Panic(KMU_EINVAL, "assertion failed: seq.size() > 0");  // seqblocksketch.rs:108
            vseq.push_back(seq);
            const size_t nb = (seq->size() + block_size_ - 1) / block_size_;
            nblocks.push_back(nb);
            for (size_t b = 0; b < nb; ++b) {
                idx.push_back(s);
                begin.push_back(b * block_size_);
                end.push_back(b * block_size_ + block_size_ + kmer_size_ - 1);  // block_size k-mers: clamped to the sequence
            }
        }
        DeviceBatch whole(vseq);
        kmu_seqbatch* blocks = nullptr;
        check(kmu_seqbatch_slices(Context::global().get(), whole.get(), idx.data(), begin.data(), end.data(), idx.size(), &blocks),
              "BlockSeqSketcher");
        std::vector<uint32_t> flat(idx.size() * sketch_size_);
        const int32_t rc = kmu_sketch_pmh3a(Context::global().get(), blocks, (uint32_t)kmer_size_, KMU_KMER32, fhash.kind,
                                            (uint32_t)sketch_size_, flat.data(), 0);
        kmu_seqbatch_destroy(blocks);
        check(rc, "BlockSeqSketcher");
        std::vector<BlockSketchedSeq> out(pack_seq.size());
        size_t row = 0;
        for (size_t s = 0; s < pack_seq.size(); ++s) {
            out[s].numseq = pack_seq[s].first;
            for (size_t b = 0; b < nblocks[s]; ++b, ++row) {
                BlockSketched bs;
                bs.numseq = pack_seq[s].first;
                bs.numblock = (uint32_t)b;
                bs.sketch.assign(flat.begin() + row * sketch_size_, flat.begin() + (row + 1) * sketch_size_);
                out[s].sketch.push_back({std::move(bs)});
            }
        }
        return out;
    }
    BlockSketchedSeq blocksketch_sequence(size_t numseq, const Sequence& seq, KmerHash fhash) const {
        return blocksketch_sequences({{(uint32_t)numseq, &seq}}, fhash)[0];
    }

  private:
    size_t block_size_, kmer_size_, sketch_size_;
};
/// 1 inside a sequence (reads are to be paired across sequences), else the fraction of differing slots (:419-433)
struct DistBlockSketched {
    float eval(const std::vector<BlockSketched>& va, const std::vector<BlockSketched>& vb) const {
        if (va.size() != 1 || vb.size() != 1) throw Panic(KMU_EINVAL, "assertion failed: va.len() == 1 && vb.len() == 1");
        if (va[0].numseq == vb[0].numseq) return 1.f;
        if (va[0].sketch.size() != vb[0].sketch.size()) throw Panic(KMU_EINVAL, "assertion failed: va.len() == vb.len()");
        return (float)(1.0 - compute_probminhash_jaccard(va[0].sketch, vb[0].sketch));
    }
};

/// SeqSketcherT (src/sketching/setsketchert.rs:54-79): sketch_compressedkmer = one signature per sequence,
/// sketch_compressedkmer_seqs = ONE signature for the whole vector (a genome in several contigs)
template <typename Kmer, typename SigT>
struct SeqSketcherT {
    using Sig = SigT;
    virtual ~SeqSketcherT() = default;
    virtual size_t get_kmer_size() const = 0;
    virtual size_t get_sketch_size() const = 0;
    virtual std::vector<std::vector<Sig>> sketch_compressedkmer(const std::vector<const Sequence*>& vseq, KmerHash fhash) const = 0;
    virtual std::vector<std::vector<Sig>> sketch_compressedkmer_seqs(const std::vector<const Sequence*>& vseq, KmerHash fhash) const = 0;
};

/// ProbHash3aSketch (setsketchert.rs:85-203)
template <typename Kmer>
class ProbHash3aSketch : public SeqSketcherT<Kmer, typename Kmer::Val> {
  public:
    using Sig = typename Kmer::Val;
    explicit ProbHash3aSketch(const SeqSketcherParams& p) : p_(p) {}
    size_t get_kmer_size() const override { return p_.kmer_size; }
    size_t get_sketch_size() const override { return p_.sketch_size; }
    std::vector<std::vector<Sig>> sketch_compressedkmer(const std::vector<const Sequence*>& vseq, KmerHash fhash) const override {
        return SeqSketcher(p_.kmer_size, p_.sketch_size).template sketch_probminhash3a<Kmer>(vseq, fhash);
    }
    std::vector<std::vector<Sig>> sketch_compressedkmer_seqs(const std::vector<const Sequence*>& vseq, KmerHash fhash) const override {
        DeviceBatch b(vseq);
        std::vector<Sig> sig(p_.sketch_size);
        check(kmu_sketch_pmh3a_whole(Context::global().get(), b.get(), (uint32_t)p_.kmer_size, Kmer::kmu_type, fhash.kind,
                                     (uint32_t)p_.sketch_size, sig.data(), 0),
              "ProbHash3aSketch::sketch_compressedkmer_seqs");
        return {sig};
    }

  private:
    SeqSketcherParams p_;
};

/// SuperHashSketch (setsketchert.rs:211-335): NoHashHasher keys (:267-269)
template <typename Kmer, typename S>
class SuperHashSketch : public SeqSketcherT<Kmer, S> {
  public:
    using Sig = S;
    explicit SuperHashSketch(const SeqSketcherParams& p) : p_(p) {}
    size_t get_kmer_size() const override { return p_.kmer_size; }
    size_t get_sketch_size() const override { return p_.sketch_size; }
    std::vector<std::vector<S>> sketch_compressedkmer(const std::vector<const Sequence*>& vseq, KmerHash fhash) const override {
        reject_empty(vseq);
        DeviceBatch b(vseq);
        std::vector<S> flat(vseq.size() * p_.sketch_size);
        check(kmu_sketch_superminhash(Context::global().get(), b.get(), (uint32_t)p_.kmer_size, Kmer::kmu_type, fhash.kind,
                                      (uint32_t)p_.sketch_size, KMU_HASHER_NOHASH, (int32_t)sizeof(S), flat.data(), 0),
              "SuperHashSketch::sketch_compressedkmer");
        return rows_of(flat, vseq.size(), p_.sketch_size);
    }
    std::vector<std::vector<S>> sketch_compressedkmer_seqs(const std::vector<const Sequence*>& vseq, KmerHash fhash) const override {
        DeviceBatch b(vseq);
        std::vector<S> sig(p_.sketch_size);
        check(kmu_sketch_superminhash_whole(Context::global().get(), b.get(), (uint32_t)p_.kmer_size, Kmer::kmu_type, fhash.kind,
                                            (uint32_t)p_.sketch_size, KMU_HASHER_NOHASH, (int32_t)sizeof(S), sig.data(), 0),
              "SuperHashSketch::sketch_compressedkmer_seqs");
        return {sig};
    }

  private:
    SeqSketcherParams p_;
};

/// sketch_seqrange_superminhash (src/sketching/seqminhash.rs:19-62): SuperMinHash (f64, NoHashHasher) of the k-mers inside
/// `range` of one sequence, canonical + int32_hash hard-coded (:37-38, 48-49); k = 16 -> Kmer16b32bit, 9..=15 -> Kmer32bit,
/// anything else panics (:55-60).  (k = 15 panics one step later in the reference, inside KmerSeqIterator::<Kmer32bit>::new.)
inline std::vector<double> sketch_seqrange_superminhash(const Sequence& seq, size_t range_start, size_t range_end, size_t kmer_size,
                                                        size_t sketch_size) {
    if (kmer_size != 16 && (kmer_size < 9 || kmer_size > 15)) throw Panic(KMU_EINVAL, "sketch_sequence_superminhash , unimplemented kmer_size");
    if (range_end <= range_start || range_end > seq.size())  // set_range(..).unwrap() (:33, :44)
        throw Panic(KMU_EINVAL, "called `Result::unwrap()` on an `Err` value: set_range");
    DeviceBatch b(seq, range_start, range_end);
    std::vector<double> sig(sketch_size);
    check(kmu_sketch_superminhash(Context::global().get(), b.get(), (uint32_t)kmer_size, kmer_size == 16 ? KMU_KMER16B32 : KMU_KMER32,
                                  KMU_HASH_CANON_INVHASH, (uint32_t)sketch_size, KMU_HASHER_NOHASH, 8, sig.data(), 0),
          "sketch_seqrange_superminhash");
    return sig;
}

/// SetSketchParams of probminhash (default b 1.001, m 4096, a 20, q 2^16 - 2)
struct SetSketchParams {
    double b = 1.001;
    uint64_t m = 4096;
    double a = 20.0;
    uint64_t q = 65534;
};

/// HyperLogLogSketch (setsketchert.rs:648-896): S = uint16_t / uint32_t / uint64_t registers
template <typename Kmer, typename S>
class HyperLogLogSketch : public SeqSketcherT<Kmer, S> {
  public:
    using Sig = S;
    HyperLogLogSketch(const SeqSketcherParams& p, const SetSketchParams& hll) : p_(p), hll_(hll) { hll_.m = p.sketch_size; }
    size_t get_kmer_size() const override { return p_.kmer_size; }
    size_t get_sketch_size() const override { return p_.sketch_size; }
    std::vector<std::vector<S>> sketch_compressedkmer(const std::vector<const Sequence*>& vseq, KmerHash fhash) const override {
        return run(vseq, fhash, 0);
    }
    std::vector<std::vector<S>> sketch_compressedkmer_seqs(const std::vector<const Sequence*>& vseq, KmerHash fhash) const override {
        return run(vseq, fhash, 1);
    }

  private:
    std::vector<std::vector<S>> run(const std::vector<const Sequence*>& vseq, KmerHash fhash, int whole) const {
        DeviceBatch b(vseq);
        const size_t nrows = whole ? 1 : vseq.size();
        std::vector<S> flat(nrows * hll_.m);
        const kmu_setsketch_params prm{hll_.b, hll_.m, hll_.a, hll_.q};
        check(kmu_sketch_setsketch(Context::global().get(), b.get(), (uint32_t)p_.kmer_size, Kmer::kmu_type, fhash.kind, &prm,
                                   (int32_t)sizeof(S), whole, flat.data(), 0),
              "HyperLogLogSketch");
        return rows_of(flat, nrows, hll_.m);
    }
    SeqSketcherParams p_;
    SetSketchParams hll_;
};

}  // namespace sketching
// ================================================================== amino acids (src/aautils)
namespace aautils {

/// Alphabet (src/aautils/kmeraa.rs:29-130): 20 residues on 5 bits, codes 1..21 in the order of "ACDEFGHIKLMNPQRSTVWY"
/// (code 14 is skipped: Q = 15)
struct Alphabet {
    static constexpr const char* bases = "ACDEFGHIKLMNPQRSTVWY";
    uint8_t len() const { return 20; }
    uint8_t get_nb_bits() const { return 5; }
    bool is_valid_base(uint8_t c) const { return c && std::strchr(bases, (char)c) != nullptr; }
    uint8_t encode(uint8_t c) const {
        const char* p = c ? std::strchr(bases, (char)c) : nullptr;
        if (!p) throw Panic(KMU_EINVAL, "encode: not a code in alpahabet for amino acid");
        const uint8_t i = (uint8_t)(p - bases) + 1;
        return i >= 14 ? i + 1 : i;
    }
    uint8_t decode(uint8_t code) const {
        const uint8_t i = code > 14 ? code - 1 : code;
        if (code == 0 || code == 14 || i > 20) throw Panic(KMU_EINVAL, "decode: not a code in alpahabet for amino acid");
        return (uint8_t)bases[i - 1];
    }
};

/// KmerAA32bit / KmerAA64bit (kmeraa.rs:147-400): k residues of 5 bits, first residue in the most significant bits
template <typename W, int32_t TYPE>
struct KmerAAbits {
    using Val = W;
    static constexpr int32_t kmu_type = TYPE;
    W aa = 0;
    uint8_t nb_base = 0;
    KmerAAbits() = default;
    explicit KmerAAbits(uint8_t nb) : aa(0), nb_base(nb) {  // ::new panics from sizeof(W) * 8 / 5 residues on (kmeraa.rs:153-161)
        if (nb >= sizeof(W) * 8 / 5) throw Panic(KMU_EINVAL, "KmerAA: nb_base too large for the word");
    }
    static KmerAAbits build(Val val, uint8_t kmer_size) { return from_word(val, kmer_size); }
    static KmerAAbits from_word(Val word, uint8_t kmer_size) {
        KmerAAbits k;
        k.aa = word;
        k.nb_base = kmer_size;
        return k;
    }
    static size_t get_nb_base_max() { return sizeof(W) * 8 / 5; }
    uint8_t get_nb_base() const { return nb_base; }
    Val get_compressed_value() const { return aa; }
    size_t get_bitsize() const { return sizeof(W) * 8; }
    /// push takes the residue as its ASCII letter and encodes it (kmeraa.rs:171-182)
    KmerAAbits push(uint8_t c) const {
        const W mask = ((W)1 << (5 * nb_base)) - 1;
        return from_word((W)(((aa << 5) & mask) | (W)(Alphabet().encode(c) & 31)), nb_base);
    }
    KmerAAbits reverse_complement() const { throw Panic(KMU_EINVAL, "KmerAA reverse_complement not yet implemented"); }
    std::vector<uint8_t> get_uncompressed_kmer() const {
        std::vector<uint8_t> s(nb_base);
        for (int i = 0; i < nb_base; ++i) s[i] = Alphabet().decode((uint8_t)((aa >> (5 * (nb_base - 1 - i))) & 31));
        return s;
    }
    friend bool operator==(KmerAAbits a, KmerAAbits b) { return a.aa == b.aa && a.nb_base == b.nb_base; }
    friend bool operator<(KmerAAbits a, KmerAAbits b) { return a.nb_base != b.nb_base ? a.nb_base < b.nb_base : a.aa < b.aa; }
};
using KmerAA32bit = KmerAAbits<uint32_t, KMU_KMERAA32>;
using KmerAA64bit = KmerAAbits<uint64_t, KMU_KMERAA64>;

/// SequenceAA (kmeraa.rs:404-456): the residues as ASCII letters
class SequenceAA {
  public:
    SequenceAA() = default;
    explicit SequenceAA(const std::string& str) : seq_(str.begin(), str.end()) {}  // SequenceAA::new / from_str
    static SequenceAA new_filtered(const std::string& buf, const Alphabet& alphabet) {  // :447-456
        SequenceAA s;
        for (char c : buf)
            if (alphabet.is_valid_base((uint8_t)c)) s.seq_.push_back((uint8_t)c);
        return s;
    }
    size_t len() const { return seq_.size(); }
    size_t size() const { return seq_.size(); }
    bool is_empty() const { return seq_.empty(); }
    uint8_t get_base(size_t pos) const {
        if (pos >= seq_.size()) throw Panic(KMU_EINVAL, "SequenceAA::get_base: position beyond the end");
        return seq_[pos];
    }
    std::string to_string() const { return std::string(seq_.begin(), seq_.end()); }
    const std::vector<uint8_t>& residues() const { return seq_; }

  private:
    std::vector<uint8_t> seq_;
};

/// RAII device batch of proteins (5-bit codes in HBM; an invalid residue panics as Alphabet::encode does)
class DeviceBatchAA {
  public:
    explicit DeviceBatchAA(const std::vector<const SequenceAA*>& vseq, size_t begin = 0, size_t end = ~(size_t)0) {
        std::vector<uint64_t> off(vseq.size() + 1, 0);
        std::vector<uint8_t> ascii;
        for (size_t i = 0; i < vseq.size(); ++i) {
            const auto& r = vseq[i]->residues();
            const size_t b = std::min(begin, r.size()), e = std::min(end, r.size());
            ascii.insert(ascii.end(), r.begin() + b, r.begin() + std::max(b, e));
            off[i + 1] = ascii.size();
        }
        ascii.push_back(0);
        check(kmu_seqbatch_from_aa(Context::global().get(), ascii.data(), off.data(), vseq.size(), 0, nullptr, &b_), "SequenceAA");
    }
    ~DeviceBatchAA() { kmu_seqbatch_destroy(b_); }
    DeviceBatchAA(const DeviceBatchAA&) = delete;
    DeviceBatchAA& operator=(const DeviceBatchAA&) = delete;
    kmu_seqbatch* get() const { return b_; }

  private:
    kmu_seqbatch* b_ = nullptr;
};

/// KmerGenerator<KmerAA..> (kmeraa.rs:646-684)
template <typename T>
class KmerGenerator {
  public:
    explicit KmerGenerator(uint8_t ksize) : kmer_size_(ksize) {}
    size_t get_kmer_size() const { return kmer_size_; }
    std::vector<T> generate_kmer(const SequenceAA& seq) const { return run(DeviceBatchAA({&seq})); }
    /// KmerSeqIterator::set_range(first, last) (kmeraa.rs:540-557): k-mers inside residues [begin, end)
    std::vector<T> generate_kmer_in_range(const SequenceAA& seq, size_t begin, size_t end) const {
        if (begin >= end || end > seq.size()) throw Panic(KMU_EINVAL, "KmerSeqIterator::set_range failed");
        return run(DeviceBatchAA({&seq}, begin, end));
    }

  private:
    std::vector<T> run(const DeviceBatchAA& b) const {
        std::vector<typename T::Val> words(kmu_kmer_count(b.get(), kmer_size_));
        check(kmu_generate_kmers(Context::global().get(), b.get(), kmer_size_, T::kmu_type, KMU_HASH_IDENTITY_RAW, words.data(), nullptr, 0),
              "KmerGenerator::generate_kmer");
        std::vector<T> out;
        out.reserve(words.size());
        for (auto w : words) out.push_back(T::from_word(w, kmer_size_));
        return out;
    }
    uint8_t kmer_size_;
};

/// SeqSketcher of the amino-acid module (src/aautils/setsketchert.rs:1020-1200)
class SeqSketcher {
  public:
    SeqSketcher(size_t kmer_size, size_t sketch_size) : kmer_size_(kmer_size), sketch_size_(sketch_size) {}
    size_t get_kmer_size() const { return kmer_size_; }
    size_t get_sketch_size() const { return sketch_size_; }
    template <typename Kmer>
    std::vector<std::vector<typename Kmer::Val>> sketch_probminhash3a(const std::vector<const SequenceAA*>& vseq, KmerHash fhash) const {
        DeviceBatchAA b(vseq);
        std::vector<typename Kmer::Val> flat(vseq.size() * sketch_size_);
        check(kmu_sketch_pmh3a(Context::global().get(), b.get(), (uint32_t)kmer_size_, Kmer::kmu_type, fhash.kind, (uint32_t)sketch_size_,
                               flat.data(), 0),
              "sketch_probminhash3a");
        return sketching::rows_of(flat, vseq.size(), sketch_size_);
    }
    template <typename Kmer, typename S = double>
    std::vector<std::vector<S>> sketch_superminhash(const std::vector<const SequenceAA*>& vseq, KmerHash fhash) const {
        DeviceBatchAA b(vseq);
        std::vector<S> flat(vseq.size() * sketch_size_);
        check(kmu_sketch_superminhash(Context::global().get(), b.get(), (uint32_t)kmer_size_, Kmer::kmu_type, fhash.kind,
                                      (uint32_t)sketch_size_, KMU_HASHER_FNV, (int32_t)sizeof(S), flat.data(), 0),
              "sketch_superminhash");
        return sketching::rows_of(flat, vseq.size(), sketch_size_);
    }

  private:
    size_t kmer_size_, sketch_size_;
};

/// SeqSketcherAAT (src/aautils/setsketchert.rs:42-72) and two of its implementations
template <typename Kmer, typename SigT>
struct SeqSketcherAAT {
    using Sig = SigT;
    virtual ~SeqSketcherAAT() = default;
    virtual size_t get_kmer_size() const = 0;
    virtual size_t get_sketch_size() const = 0;
    virtual std::vector<std::vector<Sig>> sketch_compressedkmeraa(const std::vector<const SequenceAA*>& vseq, KmerHash fhash) const = 0;
};
template <typename Kmer>
class ProbHash3aSketch : public SeqSketcherAAT<Kmer, typename Kmer::Val> {
  public:
    explicit ProbHash3aSketch(const sketching::SeqSketcherParams& p) : p_(p) {}
    size_t get_kmer_size() const override { return p_.kmer_size; }
    size_t get_sketch_size() const override { return p_.sketch_size; }
    std::vector<std::vector<typename Kmer::Val>> sketch_compressedkmeraa(const std::vector<const SequenceAA*>& vseq,
                                                                         KmerHash fhash) const override {
        return SeqSketcher(p_.kmer_size, p_.sketch_size).template sketch_probminhash3a<Kmer>(vseq, fhash);
    }

  private:
    sketching::SeqSketcherParams p_;
};
/// SuperHashSketch of the amino-acid module (src/aautils/setsketchert.rs:203-329): NoHashHasher keys (:250-252, 302-304)
template <typename Kmer, typename S>
class SuperHashSketch : public SeqSketcherAAT<Kmer, S> {
  public:
    static_assert(std::is_same<S, float>::value || std::is_same<S, double>::value, "S is f32 or f64");
    explicit SuperHashSketch(const sketching::SeqSketcherParams& p) : p_(p) {}
    size_t get_kmer_size() const override { return p_.kmer_size; }
    size_t get_sketch_size() const override { return p_.sketch_size; }
    std::vector<std::vector<S>> sketch_compressedkmeraa(const std::vector<const SequenceAA*>& vseq, KmerHash fhash) const override {
        for (const SequenceAA* q : vseq)  // set_range(0, seqb.size()).unwrap() (:255): an empty sequence panics
            if (q->size() == 0) throw Panic(KMU_EINVAL, "called `Result::unwrap()` on an `Err` value: set_range on an empty sequence");
        DeviceBatchAA b(vseq);
        std::vector<S> flat(vseq.size() * p_.sketch_size);
        check(kmu_sketch_superminhash(Context::global().get(), b.get(), (uint32_t)p_.kmer_size, Kmer::kmu_type, fhash.kind,
                                      (uint32_t)p_.sketch_size, KMU_HASHER_NOHASH, (int32_t)sizeof(S), flat.data(), 0),
              "SuperHashSketch::sketch_compressedkmeraa");
        return sketching::rows_of(flat, vseq.size(), p_.sketch_size);
    }
    /// ONE signature for the whole collection (sketch_compressedkmeraa_seqs, :291-328)
    std::vector<std::vector<S>> sketch_compressedkmeraa_seqs(const std::vector<const SequenceAA*>& vseq, KmerHash fhash) const {
        DeviceBatchAA b(vseq);
        std::vector<S> sig(p_.sketch_size);
        check(kmu_sketch_superminhash_whole(Context::global().get(), b.get(), (uint32_t)p_.kmer_size, Kmer::kmu_type, fhash.kind,
                                            (uint32_t)p_.sketch_size, KMU_HASHER_NOHASH, (int32_t)sizeof(S), sig.data(), 0),
              "SuperHashSketch::sketch_compressedkmeraa_seqs");
        return {sig};
    }

  private:
    sketching::SeqSketcherParams p_;
};
template <typename Kmer, typename S>
class HyperLogLogSketch : public SeqSketcherAAT<Kmer, S> {
  public:
    HyperLogLogSketch(const sketching::SeqSketcherParams& p, const sketching::SetSketchParams& hll) : p_(p), hll_(hll) { hll_.m = p.sketch_size; }
    size_t get_kmer_size() const override { return p_.kmer_size; }
    size_t get_sketch_size() const override { return p_.sketch_size; }
    std::vector<std::vector<S>> sketch_compressedkmeraa(const std::vector<const SequenceAA*>& vseq, KmerHash fhash) const override {
        return run(vseq, fhash, 0);
    }
    /// ONE sketch for the whole collection (sketch_compressedkmeraa_seqs)
    std::vector<std::vector<S>> sketch_compressedkmeraa_seqs(const std::vector<const SequenceAA*>& vseq, KmerHash fhash) const {
        return run(vseq, fhash, 1);
    }

  private:
    std::vector<std::vector<S>> run(const std::vector<const SequenceAA*>& vseq, KmerHash fhash, int whole) const {
        DeviceBatchAA b(vseq);
        const size_t nrows = whole ? 1 : vseq.size();
        std::vector<S> flat(nrows * hll_.m);
        const kmu_setsketch_params prm{hll_.b, hll_.m, hll_.a, hll_.q};
        check(kmu_sketch_setsketch(Context::global().get(), b.get(), (uint32_t)p_.kmer_size, Kmer::kmu_type, fhash.kind, &prm,
                                   (int32_t)sizeof(S), whole, flat.data(), 0),
              "HyperLogLogSketch");
        return sketching::rows_of(flat, nrows, hll_.m);
    }
    sketching::SeqSketcherParams p_;
    sketching::SetSketchParams hll_;
};

}  // namespace aautils
}  // namespace kmerutils
