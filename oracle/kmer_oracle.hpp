// ============================================================================
//  oracle/kmer_oracle.hpp  --  TEST INFRASTRUCTURE ONLY.
//
//  CPU restatement of the kmerutils hot path (reference: jean-pierreBoth/kmerutils
//  v0.0.14, Rust).  It exists so that tests/, __graft_entry__.smoke() and
//  bench.py's cpu_baseline / --impl reference legs can CHECK (and time beside)
//  the CUDA path.  Nothing under kmerutils_b200/ may include, link or call it.
//
//  Parity status
//  -------------
//  * First-party arithmetic (2-bit packing, k-mer words, reverse complement,
//    generation, ntHash, NoHashHasher, AA k-mers, counting semantics) follows
//    the reference sources line by line and is PINNED against every
//    known-answer value the reference's own unit tests hold (tests/test_oracle_kat.py).
//  * Sketch arithmetic (ProbMinHash3a / SuperMinHash / SetSketch, int32_hash /
//    int64_hash, Xoshiro256++ / SplitMix64, rand 0.9 Uniform samplers) lives in
//    the un-vendored crates probminhash ^0.1, rand 0.9, rand_distr 0.5,
//    rand_xoshiro 0.7 (reference Cargo.toml:74-89; Cargo.lock is git-ignored).
//    It is restated here from the published algorithms; the reference holds no
//    golden signature, so for signatures this oracle is "PARITY UNPINNED"
//    (statistical reference tests are re-run instead, see tests/).
//
//  All citations are path:line under /root/reference/.
// ============================================================================
#pragma once
#include <cstdint>
#include <cstddef>

#ifdef __cplusplus
extern "C" {
#endif

// k-mer word types (the reference's three DNA k-mer structs + two AA structs)
enum {
    ORC_KMER32 = 0,    // Kmer32bit   src/base/kmer32bit.rs:22   (k <= 14, k in top 4 bits)
    ORC_KMER16B32 = 1, // Kmer16b32bit src/base/kmer16b32bit.rs:21 (k == 16)
    ORC_KMER64 = 2,    // Kmer64bit   src/base/kmer64bit.rs:24   (k <= 32, value only; k separate)
    ORC_KMERAA32 = 3,  // KmerAA32bit src/aautils/kmeraa.rs:146  (k <= 6, 5 bits/residue)
    ORC_KMERAA64 = 4   // KmerAA64bit src/aautils/kmeraa.rs:280  (k <= 12)
};

// the hash closures `fhash` the reference actually uses
enum {
    ORC_HASH_IDENTITY_RAW = 0,  // kmer.0                              seqsketchjaccard.rs:775
    ORC_HASH_MASKED_VALUE = 1,  // get_compressed_value() & mask       setsketchert.rs:1098-1104
    ORC_HASH_CANON_INVHASH = 2, // intNN_hash(min(kmer, revcomp).0)    datasketcher.rs:222-226
    ORC_HASH_CANON_RAW = 3,     // min(kmer, revcomp).0                kmercount.rs:313 (counting key)
    ORC_HASH_INVHASH = 4        // intNN_hash(kmer.0)                  minhash.rs:226
};

// ---- A1/A2 : alphabet + Sequence ------------------------------------------------
// Sequence::new(raw, 2): returns number of bytes written (ceil(n/4)) or -1 on a
// non-ACGT character (the reference panics, alphabet.rs:125).
int64_t orc_pack_2bit(const uint8_t* ascii, uint64_t n, uint8_t* out);
// Sequence::encode_and_add: silently drops non-ACGT; returns bases kept.
uint64_t orc_encode_and_add_2bit(const uint8_t* ascii, uint64_t n, uint8_t* out);
uint64_t orc_count_non_acgt(const uint8_t* ascii, uint64_t n);
uint8_t orc_get_base(const uint8_t* packed, uint64_t pos);
void orc_unpack_2bit(const uint8_t* packed, uint64_t nbases, uint8_t* ascii_out);
// Sequence::get_reverse_complement (2-bit), sequence.rs:252-295
void orc_seq_revcomp_2bit(const uint8_t* packed, uint64_t nbases, uint8_t* out);

// ---- A3-A7 : k-mer words ---------------------------------------------------------
uint64_t orc_kmer_build(uint64_t value, int k, int type);
uint64_t orc_kmer_push(uint64_t word, int k, int type, uint8_t base);
uint64_t orc_kmer_revcomp(uint64_t word, int k, int type);
// -1 / 0 / +1 following the type's Ord impl
int orc_kmer_cmp(uint64_t a, uint64_t b, int k, int type);
uint64_t orc_kmer_compressed_value(uint64_t word, int k, int type);
// KmerSeqIterator over [begin, end): writes words (as u64) and returns the count.
// returns UINT64_MAX if (k, type) is rejected by the reference (it panics).
uint64_t orc_generate_kmers(const uint8_t* packed, uint64_t nbases, uint64_t begin, uint64_t end,
                            int k, int type, uint64_t* out);
uint64_t orc_apply_hash(uint64_t word, int k, int type, int hash_kind);
uint32_t orc_int32_hash(uint32_t key);
uint64_t orc_int64_hash(uint64_t key);

// ---- A8 : ntHash (2-bit *_init functions; kmer.rs:48-94, nthash.rs:63-72) ----------
uint64_t orc_nthash_init(uint64_t word, int k, int type);
// returns strand (0 if fhash <= rhash)
int orc_nthash_canonical_init(uint64_t word, int k, int type, uint64_t* fhash, uint64_t* rhash, uint64_t* canon);
void orc_nthash_mult(uint64_t h0, int k, uint64_t* hashed, int n);
// bug-compatible single-step shims (kmer.rs:63-71, 96-117)
uint64_t orc_nthash_cycle(uint64_t word, int k, int type, uint64_t hashval, uint8_t new_base);
int orc_nthash_canonical_cycle(uint64_t word, int k, int type, uint8_t new_base, uint64_t* fhash,
                               uint64_t* rhash, uint64_t* canon);

// ---- A17 : NoHashHasher / FNV seeds -------------------------------------------------
uint64_t orc_nohash_seed(uint64_t key, int key_bytes);
uint64_t orc_fnv1a_seed(uint64_t key, int key_bytes);

// ---- RNG building blocks (exposed so the tests can pin them) ------------------------
void orc_xoshiro_seed(uint64_t seed, uint64_t s[4]);
uint64_t orc_xoshiro_next(uint64_t s[4]);

// ---- A11 : ProbMinHash3a ---------------------------------------------------------
// keys/counts: the multiplicity map in ascending key order. sig: m entries (u64).
// key_bytes 4 or 8 (selects the NoHashHasher seed width).
void orc_pmh3a_weighted(const uint64_t* keys, const double* weights, uint64_t n, uint32_t m,
                        int key_bytes, uint64_t* sig);
// one sequence: KmerSeqIterator -> fhash -> map -> ProbMinHash3a (seqsketchjaccard.rs:224-243)
void orc_sketch_pmh3a_seq(const uint8_t* packed, uint64_t nbases, int k, int type, int hash_kind,
                          uint32_t m, uint64_t* sig);
// batch, one task per sequence over nthreads (rayon analogue, seqsketchjaccard.rs:245-248).
// sig_out is nseq*m elements of sig_bytes (4 or 8) each.
void orc_sketch_pmh3a_batch(const uint8_t* packed, const uint64_t* byte_off, const uint64_t* nbases,
                            uint64_t nseq, int k, int type, int hash_kind, uint32_t m,
                            void* sig_out, int sig_bytes, int nthreads);
// whole-file variant (setsketchert.rs:160-202): one multiplicity map over all sequences
void orc_sketch_pmh3a_seqs(const uint8_t* packed, const uint64_t* byte_off, const uint64_t* nbases,
                           uint64_t nseq, int k, int type, int hash_kind, uint32_t m, uint64_t* sig);
// block sketch (seqblocksketch.rs:97-149): returns number of blocks; sig_out nblocks*m u32
uint64_t orc_blocksketch_seq(const uint8_t* packed, uint64_t nbases, int k, uint32_t m,
                             uint64_t block_size, uint32_t* sig_out, uint64_t max_blocks);
double orc_jaccard_equal_fraction(const void* a, const void* b, uint32_t m, int sig_bytes);

// ---- A12 : SuperMinHash (probminhash::superminhasher, Ertl 2017; PARITY UNPINNED) ---------
// one sketch over a group of nseq sequences; hasher 0 = NoHashHasher, 1 = FnvHasher; sig_bytes 4 = f32, 8 = f64
void orc_sketch_superminhash(const uint8_t* packed, const uint64_t* byte_off, const uint64_t* nbases,
                             uint64_t nseq, int k, int type, int hash_kind, uint32_t m, int hasher,
                             int sig_bytes, void* out);
// one sketch per sequence (rayon analogue), out: nseq * m values
void orc_sketch_superminhash_batch(const uint8_t* packed, const uint64_t* byte_off, const uint64_t* nbases,
                                   uint64_t nseq, int k, int type, int hash_kind, uint32_t m, int hasher,
                                   int sig_bytes, void* out, int nthreads);

// ---- A13 : SetSketch / HyperLogLogSketch (probminhash::setsketcher, Ertl 2021; PARITY UNPINNED:
//      rand_distr's ziggurat tables are regenerated, ln/exp are the deterministic det_math.hpp ones) ----
void orc_sketch_setsketch(const uint8_t* packed, const uint64_t* byte_off, const uint64_t* nbases,
                          uint64_t nseq, int k, int type, int hash_kind, double b, uint64_t m, double a,
                          uint64_t q, int sig_bytes, void* out);
void orc_sketch_setsketch_batch(const uint8_t* packed, const uint64_t* byte_off, const uint64_t* nbases,
                                uint64_t nseq, int k, int type, int hash_kind, double b, uint64_t m,
                                double a, uint64_t q, int sig_bytes, void* out, int nthreads);
double orc_det_log(double x);
double orc_det_exp(double x);
double orc_det_expm1(double x);
// statistics of the deviations of det_log / det_exp / det_expm1 from the platform libm on arguments drawn as the sketchers
// draw them (see kmer_oracle.cpp); out: 9 counters
void orc_libm_divergence(uint64_t nkeys, uint64_t seed, double b, uint64_t m, double a, uint64_t q, uint32_t points, uint32_t m_pmh,
                         int nthreads, uint64_t* out);
double orc_exp1_from_seed(uint64_t seed, int skip);

// ---- A15 : counting (exact multiset semantics of KmerCounter, kmercount.rs:241-288) -----
// distinct canonical compressed k-mer values in ascending order with their multiplicities;
// returns the number of distinct keys (only the first `cap` are written); UINT64_MAX on a bad (k, type)
uint64_t orc_count_kmers(const uint8_t* packed, const uint64_t* byte_off, const uint64_t* nbases,
                         uint64_t nseq, int k, int type, int canonical, uint64_t* keys_out,
                         uint64_t* counts_out, uint64_t cap);
// DispatchableT::dispatch (kmercount.rs:382-420)
uint64_t orc_dispatch(uint64_t compressed_value, int type, uint64_t nb_receiver);

// ---- synthetic data ---------------------------------------------------
// base i of stream `seed` = top 2 bits of SplitMix64 output number i (counter based)
void orc_synth_packed(uint64_t seed, uint64_t first_base, uint64_t nbases, uint8_t* packed_out);
void orc_synth_ascii(uint64_t seed, uint64_t first_base, uint64_t nbases, uint8_t* ascii_out);

// ---- A14 : amino acids.  For the KMERAA types every `packed` argument above is the SequenceAA
// payload: one ASCII residue per byte (kmeraa.rs:404-406), lengths in residues.
void orc_synth_aa(uint64_t seed, uint64_t first_res, uint64_t nres, uint8_t* ascii_out);
uint64_t orc_aa_filter(const uint8_t* ascii, uint64_t n, uint8_t* out);

void orc_sample_read(const uint8_t* genome_packed, uint64_t glen, uint64_t seed, uint64_t r, uint32_t read_len,
                     uint32_t err_ppm, uint8_t* ascii_out);
int orc_hardware_threads(void);

#ifdef __cplusplus
}
#endif
