/* ============================================================================
 *  kmerutils_b200.h -- C ABI of the B200-native k-mer engine.
 *
 *  This is the drop-in boundary for the data-parallel hot path of
 *  jean-pierreBoth/kmerutils (Rust, v0.0.14).  The reference has no FFI of its
 *  own; the entry points below are what a `extern "C"` shim
 *  inside the Rust crate binds, each one replacing the
 *  reference function cited next to it (paths relative to the reference tree).
 *
 *  Conventions
 *   - every function returns an int32 status (KMU_OK == 0); the message of the
 *     last failure on the calling thread is kmu_last_error().
 *   - plain pointers and sizes only; the caller owns every input/output buffer.
 *   - the library never falls back to a CPU implementation: without a CUDA
 *     device every compute entry point returns KMU_ECUDA.
 *   - results come back in input order (the reference asserts this,
 *     src/sketching/setsketchert.rs:154-155).
 *   - a context is bound to one GPU; calls on one context are serialised by an
 *     internal mutex, different contexts run concurrently.
 * ==========================================================================*/
#ifndef KMERUTILS_B200_H
#define KMERUTILS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes --------------------------------------------------------- */
#define KMU_OK 0
#define KMU_EINVAL 1 /* the reference would panic on this input (bad k for the type, m < 2, ...) */
#define KMU_ECUDA 2  /* CUDA runtime failure or no device */
#define KMU_ENOMEM 3
#define KMU_EOVERFLOW 4

/* ---- k-mer word types ---------------------------------- */
#define KMU_KMER32 0    /* Kmer32bit    src/base/kmer32bit.rs:22    k <= 14, k in the top 4 bits, u32 */
#define KMU_KMER16B32 1 /* Kmer16b32bit src/base/kmer16b32bit.rs:21 k == 16, u32 */
#define KMU_KMER64 2    /* Kmer64bit    src/base/kmer64bit.rs:24    k <= 32, u64 value (k kept apart) */
#define KMU_KMERAA32 3  /* KmerAA32bit  src/aautils/kmeraa.rs:146   k <= 6, 5 bits / residue, u32 */
#define KMU_KMERAA64 4  /* KmerAA64bit  src/aautils/kmeraa.rs:280   k <= 12, u64 */

/* ---- the hash closures `fhash` the reference passes to its sketchers.
 * A Rust closure cannot cross into CUDA, so the boundary takes an enumerated kind. */
#define KMU_HASH_IDENTITY_RAW 0  /* |k| k.0                                 seqsketchjaccard.rs:775     */
#define KMU_HASH_MASKED_VALUE 1  /* |k| k.get_compressed_value() & mask     setsketchert.rs:1098-1104   */
#define KMU_HASH_CANON_INVHASH 2 /* |k| intNN_hash(k.reverse_complement().min(*k).0)  datasketcher.rs:222-226 */
#define KMU_HASH_CANON_RAW 3     /* |k| k.reverse_complement().min(*k).0    kmercount.rs:313,827,938    */
#define KMU_HASH_INVHASH 4       /* |k| intNN_hash(k.0)                      minhash.rs:226              */

typedef struct kmu_ctx kmu_ctx;
typedef struct kmu_seqbatch kmu_seqbatch;
typedef struct kmu_counter kmu_counter;

/* ---- context ------------------------------------------------------------------ */
int32_t kmu_ctx_create(int32_t device, kmu_ctx** ctx);
void kmu_ctx_destroy(kmu_ctx* ctx);
const char* kmu_last_error(void);
/* library version / build info string (static storage) */
const char* kmu_version(void);
/* number of kernels of this library launched on the context since creation */
uint64_t kmu_launch_count(const kmu_ctx* ctx);
/* the CUDA stream (cudaStream_t) all work of this context is enqueued on */
void* kmu_ctx_stream(kmu_ctx* ctx);
/* block the calling thread until the context's stream is idle */
int32_t kmu_ctx_sync(kmu_ctx* ctx);

/* ---- sequence batches (replaces Vec<Sequence>, src/base/sequence.rs:14-20) --------
 * A batch is a set of 2-bit packed sequences resident in HBM: one byte buffer in
 * which every sequence starts on a 16-byte boundary with the reference's byte
 * layout (4 bases / byte, first base in the two most significant bits,
 * src/base/alphabet.rs:162-168), plus per-sequence byte offsets and base counts. */

/* from `nseq` separate host buffers -- the `&[&Sequence]` the Rust entry points get
 * (setsketchert.rs:70-79).  seq_ptrs[i] holds ceil(nbases[i]/4) bytes. */
int32_t kmu_seqbatch_from_ptrs(kmu_ctx* ctx, const uint8_t* const* seq_ptrs, const uint64_t* nbases, uint64_t nseq,
                               kmu_seqbatch** batch);
/* from one concatenated host buffer; sequence i starts at packed + byte_off[i].
 * If every byte_off[i] is a multiple of 16 the buffer is copied as is (pin it for
 * full PCIe speed); otherwise it is re-laid out on the host first. */
int32_t kmu_seqbatch_from_packed(kmu_ctx* ctx, const uint8_t* packed, uint64_t packed_bytes, const uint64_t* byte_off,
                                 const uint64_t* nbases, uint64_t nseq, kmu_seqbatch** batch);
/* ASCII (FASTA/FASTQ payload) -> 2-bit on the GPU: Sequence::new(raw, 2)
 * (sequence.rs:25-106).  ascii is one concatenated host buffer, sequence i is
 * ascii[ascii_off[i] .. ascii_off[i+1]).  invalid_counts (nseq entries, may be NULL)
 * receives count_non_acgt (alphabet.rs:28-31); when drop_invalid == 0 a sequence
 * with any invalid character makes the call fail with KMU_EINVAL (the reference
 * panics, alphabet.rs:125); when drop_invalid != 0 invalid characters are skipped
 * as by Sequence::encode_and_add (sequence.rs:388-451). */
int32_t kmu_seqbatch_from_ascii(kmu_ctx* ctx, const uint8_t* ascii, const uint64_t* ascii_off, uint64_t nseq,
                                int32_t drop_invalid, uint64_t* invalid_counts, kmu_seqbatch** batch);
/* synthetic uniform ACGT reads generated on the device (bench / tests):
 * base j of sequence i = top 2 bits of SplitMix64 output (first_base[i] + j) of stream `seed`. */
int32_t kmu_seqbatch_synth(kmu_ctx* ctx, uint64_t seed, const uint64_t* nbases, uint64_t nseq, kmu_seqbatch** batch);
/* amino-acid sequences: SequenceAA (src/aautils/kmeraa.rs:404-456), one ASCII residue per byte in,
 * 5-bit codes (Alphabet::encode, :85-109) in HBM.  drop_invalid != 0 is SequenceAA::new_filtered
 * (:447-456); otherwise any residue outside "ACDEFGHIKLMNPQRSTVWY" fails with KMU_EINVAL (the
 * reference panics in Alphabet::encode, :106).  Use with KMU_KMERAA32 (k <= 6) / KMU_KMERAA64 (k <= 12). */
int32_t kmu_seqbatch_from_aa(kmu_ctx* ctx, const uint8_t* ascii, const uint64_t* ascii_off, uint64_t nseq,
                             int32_t drop_invalid, uint64_t* invalid_counts, kmu_seqbatch** batch);
/* synthetic proteins: residue j of sequence i = "ACDEFGHIKLMNPQRSTVWY"[z % 20] */
int32_t kmu_seqbatch_synth_aa(kmu_ctx* ctx, uint64_t seed, const uint64_t* nres, uint64_t nseq, kmu_seqbatch** batch);
/* 0 = DNA (2 bits / base), 1 = amino acids (one 5-bit code per byte) */
int32_t kmu_seqbatch_alphabet(const kmu_seqbatch* batch);
/* synthetic short reads (bench / tests): reads first_read .. first_read + nreads of length
 * read_len drawn uniformly from the one sequence of `genome`, random strand, substitution errors at
 * err_ppm per million bases.  The read index seeds every draw, so shards of one read set can be
 * generated independently on several GPUs. */
int32_t kmu_seqbatch_sample_reads(kmu_ctx* ctx, const kmu_seqbatch* genome, uint64_t seed, uint64_t first_read,
                                  uint64_t nreads, uint32_t read_len, uint32_t err_ppm, kmu_seqbatch** batch);
/* sub-ranges [begin[i], end[i]) (in bases / residues, clamped to the sequence) of sequences seq_idx[i] of `src`
 * as a new batch: KmerSeqIterator::set_range (kmergenerator.rs:56-65), the blocks of BlockSeqSketcher
 * (seqblocksketch.rs:97-149: block b of a sequence = bases [b*bs, b*bs + bs + k - 1)), the ranges of
 * sketch_seqrange_superminhash (seqminhash.rs:19-62). */
int32_t kmu_seqbatch_slices(kmu_ctx* ctx, const kmu_seqbatch* src, const uint64_t* seq_idx, const uint64_t* begin,
                            const uint64_t* end, uint64_t nslices, kmu_seqbatch** batch);
/* a view of nseq consecutive sequences of `src` (e.g. the contigs of one genome of a multi-genome batch) that shares the
 * packed bases of `src`; `src` must outlive the view.  Destroy it with kmu_seqbatch_destroy. */
int32_t kmu_seqbatch_view(const kmu_seqbatch* src, uint64_t first_seq, uint64_t nseq, kmu_seqbatch** view);
void kmu_seqbatch_destroy(kmu_seqbatch* batch);
uint64_t kmu_seqbatch_nseq(const kmu_seqbatch* batch);
uint64_t kmu_seqbatch_total_bases(const kmu_seqbatch* batch);
uint64_t kmu_seqbatch_packed_bytes(const kmu_seqbatch* batch);
/* copy the device layout back to the host (packed_out: kmu_seqbatch_packed_bytes();
 * byte_off_out / nbases_out: nseq entries each; any may be NULL) */
int32_t kmu_seqbatch_download(kmu_ctx* ctx, const kmu_seqbatch* batch, uint8_t* packed_out, uint64_t* byte_off_out,
                              uint64_t* nbases_out);

/* ---- k-mer generation (KmerGenerator::generate_kmer, src/base/kmergenerator.rs:162-167;
 *      KmerSeqIterator::next :75-106) fused with the hash closure -------------------------
 * For every sequence all windows of k bases, forward strand, in order; each k-mer
 * word is mapped through `hash_kind` (KMU_HASH_IDENTITY_RAW gives the reference's
 * Vec<Kmer>).  out holds sum_i max(0, nbases[i]-k+1) elements of 4 bytes (u32 types)
 * or 8 bytes (u64 types); out_off (nseq+1 entries, may be NULL) receives the
 * element offset of each sequence's first k-mer.  out is a host pointer unless
 * out_on_device != 0. */
uint64_t kmu_kmer_count(const kmu_seqbatch* batch, uint32_t k);
int32_t kmu_generate_kmers(kmu_ctx* ctx, const kmu_seqbatch* batch, uint32_t k, int32_t kmer_type, int32_t hash_kind,
                           void* out, uint64_t* out_off, int32_t out_on_device);

/* ---- ntHash (NtHash::nthash_canonical_init / nthash_mult_canonical_init for the 2-bit
 *      k-mer types, src/base/kmer.rs:74-94,120-129; nthash.rs:63-72) ------------------------
 * out_hash: n_kmers * n_multi u64 (k-mer major); hash 0 is min(fhash, rhash), hashes
 * 1.. follow from_one_hash_val_to_mult_hash.  out_strand (may be NULL): 0 if
 * fhash <= rhash else 1.  Defined for k <= 32 (the reference implements the trait
 * for the two u32 types only; k > 16 is the same formula on Kmer64bit). */
int32_t kmu_nthash_canonical(kmu_ctx* ctx, const kmu_seqbatch* batch, uint32_t k, uint32_t n_multi, uint64_t* out_hash,
                             uint8_t* out_strand, int32_t out_on_device);

/* ---- ProbMinHash3a per-sequence sketch
 *      SeqSketcher::sketch_probminhash3a   src/sketching/seqsketchjaccard.rs:211-260
 *      ProbHash3aSketch::sketch_compressedkmer   src/sketching/setsketchert.rs:121-157
 * One signature of m slots per sequence: slot j holds the hashed k-mer whose
 * weighted exponential point is minimal in slot j (0 for an untouched slot).
 * sig: nseq * m elements of 4 bytes (u32 k-mer types) or 8 bytes (u64 types),
 * row i = sequence i.  m >= 2. */
int32_t kmu_sketch_pmh3a(kmu_ctx* ctx, const kmu_seqbatch* batch, uint32_t k, int32_t kmer_type, int32_t hash_kind,
                         uint32_t m, void* sig, int32_t sig_on_device);
/* one-shot host form: sequences in host memory in, signatures in host memory out.  The input is
 * cut into chunks that are uploaded, sketched and downloaded on three streams (double buffered);
 * pin the buffers (cudaHostRegister / cudaMallocHost) for full PCIe speed. */
int32_t kmu_sketch_pmh3a_host(kmu_ctx* ctx, const uint8_t* packed, uint64_t packed_bytes, const uint64_t* byte_off,
                              const uint64_t* nbases, uint64_t nseq, uint32_t k, int32_t kmer_type, int32_t hash_kind,
                              uint32_t m, void* sig);

/* the same with the sequences as nseq SEPARATE host allocations -- what `sketch_compressedkmer(&self, vseq: &[&Sequence], ..)`
 * (setsketchert.rs:70-79) and `SeqSketcher::sketch_probminhash3a(&self, vseq: &[&Sequence], ..)` (seqsketchjaccard.rs:211-220)
 * receive: seq_ptrs[i] = Sequence::seq of sequence i, ceil(nbases[i] / 4) bytes.  The sequences are gathered chunk by
 * chunk into pinned staging memory by several host threads while the previous chunk is being sketched. */
int32_t kmu_sketch_pmh3a_host_ptrs(kmu_ctx* ctx, const uint8_t* const* seq_ptrs, const uint64_t* nbases, uint64_t nseq,
                                   uint32_t k, int32_t kmer_type, int32_t hash_kind, uint32_t m, void* sig);

/* whole-file form: ONE signature for the batch (all contigs of a genome counted into one multiplicity
 * map), ProbHash3aSketch::sketch_compressedkmer_seqs  src/sketching/setsketchert.rs:160-202.  sig: m values. */
int32_t kmu_sketch_pmh3a_whole(kmu_ctx* ctx, const kmu_seqbatch* batch, uint32_t k, int32_t kmer_type, int32_t hash_kind,
                               uint32_t m, void* sig, int32_t sig_on_device);
/* the same for many genomes in one call: group g = the next group_sizes[g] sequences of the batch (the groups cover the
 * batch), one signature per group, no host round trip between groups.  This is the loop gsearch runs over
 * sketch_compressedkmer_seqs, one genome after the other.  sig: ngroups rows of m values. */
int32_t kmu_sketch_pmh3a_groups(kmu_ctx* ctx, const kmu_seqbatch* batch, const uint64_t* group_sizes, uint64_t ngroups,
                                uint32_t k, int32_t kmer_type, int32_t hash_kind, uint32_t m, void* sig, int32_t sig_on_device);
/* ProbMinHash3a::hash_weigthed_hashmap on an explicit weighted set (host arrays): n distinct keys of key_bytes
 * (4 or 8) with positive f64 weights -- the f64-weighted maps of BlockSeqSketcher (seqblocksketch.rs:121-138)
 * or any multiplicity map built elsewhere.  sig: m keys. */
int32_t kmu_pmh3a_weighted(kmu_ctx* ctx, const void* keys, const double* weights, uint64_t n, int32_t key_bytes,
                           uint32_t m, void* sig);

/* Peer-to-peer form of the exchange: ONE kernel extracts the canonical k-mers, buckets them by owner and stores
 * every bucket straight into its owner's receive buffer -- a buffer of another GPU of the box mapped with CUDA IPC,
 * so the stores cross NVLink from inside the kernel (no staging copy, no separate collective).
 *   1. kmu_count_partition_counts : per-owner counts of this rank's k-mers;
 *   2. the ranks share the counts (a few integers) and derive, for every destination, the offset of each sender's
 *      bucket; receive buffers are allocated with kmu_ipc_alloc and opened on the senders with kmu_ipc_open;
 *   3. kmu_count_partition_scatter : the same walk writes bucket p at dests[p] + dest_offsets[p] (elements);
 *   4. after a barrier each rank feeds its receive buffer to kmu_count_insert_kmers. */
int32_t kmu_count_partition_counts(kmu_ctx* ctx, const kmu_seqbatch* batch, uint32_t k, int32_t kmer_type,
                                   int32_t canonical, uint32_t nparts, uint64_t* part_counts);
int32_t kmu_count_partition_scatter(kmu_ctx* ctx, const kmu_seqbatch* batch, uint32_t k, int32_t kmer_type,
                                    int32_t canonical, uint32_t nparts, void* const* dests, const uint64_t* dest_offsets);
/* Fused exchange, one walk (replaces the one-producer / N-consumer hand-off of count_kmer_threaded_one_to_many,
 * kmercount.rs:881-974, owner = DispatchableT::dispatch :382-420): ONE kernel extracts the canonical k-mers of the batch,
 * buckets them by owner and appends every bucket to its slab inside the owner's receive buffer -- dests[o], local memory
 * for o == self, a peer GPU's buffer opened with kmu_ipc_open otherwise: the stores cross NVLink from inside the kernel.
 * Every rank creates its counter with the same arguments (same capacity); a receive buffer holds nregions * nowners slabs of
 * slab_cap keys, slab (r, s) = what sender s found for region r.  Among several owners nregions is 1 (one slab per sender:
 * long runs for the NVLink stores; the receiver cuts what it got by region of its table inside kmu_count_insert_slabs);
 * with one owner the buckets are the regions of the local table.
 *   kmu_count_exchange_geometry : nregions for this table and this number of owners;
 *   kmu_count_exchange_scatter  : the kernel; sent_counts[o * nregions + r] = keys this rank appended for (o, r);
 *                                 *overflowed != 0 when a slab was too small (nothing may be inserted from this round);
 *   kmu_count_insert_slabs      : after the ranks have shared their sent_counts (and a barrier), each rank inserts its
 *                                 buffer region after region -- the updates of a region hit L2;
 *                                 counts[s * nregions + r] = sender s's sent_counts[self * nregions + r]. */
int32_t kmu_count_exchange_geometry(const kmu_counter* counter, uint32_t nowners, uint32_t* nregions);
int32_t kmu_count_exchange_scatter(kmu_ctx* ctx, const kmu_seqbatch* batch, const kmu_counter* counter, int32_t canonical,
                                   uint32_t nowners, uint32_t self, uint64_t slab_cap, void* const* dests,
                                   uint64_t* sent_counts, int32_t* overflowed);
int32_t kmu_count_insert_slabs(kmu_ctx* ctx, kmu_counter* counter, const void* slabs, uint64_t slab_cap, uint32_t nsend,
                               const uint64_t* counts);
int32_t kmu_ipc_alloc(kmu_ctx* ctx, uint64_t bytes, void** dev_ptr, uint8_t handle[64]);
int32_t kmu_ipc_free(kmu_ctx* ctx, void* dev_ptr);
int32_t kmu_ipc_open(kmu_ctx* ctx, const uint8_t handle[64], void** peer_ptr);
int32_t kmu_ipc_close(kmu_ctx* ctx, void* peer_ptr);

/* ---- SuperMinHash per-sequence sketch
 *      SeqSketcher::sketch_superminhash          src/sketching/seqsketchjaccard.rs:328-380  (key_hasher FNV, :346-349)
 *      SuperHashSketch::sketch_compressedkmer    src/sketching/setsketchert.rs:255-296      (key_hasher NOHASH, :267-269)
 * Every k-mer (duplicates included) is streamed through SuperMinHash::sketch; the signature is
 * get_hsketch(): m floating point values per sequence, f32 (sig_bytes 4) or f64 (sig_bytes 8),
 * mergeable by element-wise minimum.  A sequence without any k-mer keeps the initial value
 * F::from(u32::MAX) in every slot. */
#define KMU_HASHER_NOHASH 0 /* NoHashHasher  src/nohasher.rs:22-48 */
#define KMU_HASHER_FNV 1    /* fnv::FnvHasher (FNV-1a 64 over the native-endian key bytes) */
int32_t kmu_sketch_superminhash(kmu_ctx* ctx, const kmu_seqbatch* batch, uint32_t k, int32_t kmer_type,
                                int32_t hash_kind, uint32_t m, int32_t key_hasher, int32_t sig_bytes, void* sig,
                                int32_t sig_on_device);

/* ---- SetSketch / HyperLogLog registers
 *      HyperLogLogSketch::sketch_compressedkmer        src/sketching/setsketchert.rs:758-802  (whole == 0: one sketch per sequence)
 *      HyperLogLogSketch::sketch_compressedkmer_seqs   src/sketching/setsketchert.rs:811-895  (whole != 0: ONE sketch for the batch;
 *                                                      the reference's block split + SetSketcher::merge is an element-wise max)
 *      amino-acid mirror                               src/aautils/setsketchert.rs:790-1011
 * Every k-mer is streamed through probminhash's SetSketcher::sketch with NoHashHasher keys; the
 * signature is get_signature(): m registers of u16 / u32 / u64 (sig_bytes 2 / 4 / 8), mergeable by
 * element-wise maximum.  params == NULL means SetSketchParams::default() = {b 1.001, m 4096, a 20, q 2^16 - 2}. */
typedef struct kmu_setsketch_params {
    double b;   /* base of the geometric levels, > 1 */
    uint64_t m; /* number of registers */
    double a;   /* rate of the exponential */
    uint64_t q; /* registers saturate at q + 1 */
} kmu_setsketch_params;
int32_t kmu_sketch_setsketch(kmu_ctx* ctx, const kmu_seqbatch* batch, uint32_t k, int32_t kmer_type, int32_t hash_kind,
                             const kmu_setsketch_params* params, int32_t sig_bytes, int32_t whole, void* sig,
                             int32_t sig_on_device);

/* ---- Jaccard estimates between signatures: out[i * nb + j] = #{s : a_i[s] == b_j[s]} / m
 *      compute_probminhash_jaccard / probminhash_get_jaccard_objects (src/sketching/seqsketchjaccard.rs:86-108), the
 *      comparison step of jaccard_index_probminhash3a (:423-495), SuperMinHash::get_jaccard_index_estimate, and the
 *      DistHamming of datasketcher (src/bin/datasketcher.rs:179-185) as 1 - out.  slot_bytes 2 / 4 / 8; slots are
 *      compared as bit patterns. */
int32_t kmu_signature_jaccard(kmu_ctx* ctx, const void* sig_a, uint64_t na, const void* sig_b, uint64_t nb, uint32_t m,
                              int32_t slot_bytes, double* out, int32_t on_device);

/* whole-file form: ONE SuperMinHash signature for the batch (SuperHashSketch::sketch_compressedkmer_seqs,
 * src/sketching/setsketchert.rs:299-335); equals the element-wise minimum of the per-sequence signatures */
int32_t kmu_sketch_superminhash_whole(kmu_ctx* ctx, const kmu_seqbatch* batch, uint32_t k, int32_t kmer_type,
                                      int32_t hash_kind, uint32_t m, int32_t key_hasher, int32_t sig_bytes, void* sig,
                                      int32_t sig_on_device);

/* ---- k-mer counting (replaces KmerCounter: cuckoo filter + counting Bloom filter,
 *      src/base/kmercount.rs:70-83) -------------------------------------------------------
 * One exact open-addressing table in HBM keyed by kmer.get_compressed_value().  Semantics are
 * those of the reference with zero filter false positives: get_count = 0 (never inserted), 1
 * (inserted once) or min(multiplicity, 2^count_bits - 1) (the counting Bloom filter saturates,
 * test kmercount.rs:1615); nb_distinct = number of different k-mers; nb_unique = k-mers seen
 * exactly once (KmerCountT, kmercount.rs:48-59).  `capacity` is the number of distinct k-mers
 * the table must hold (KmerCounter::new(fpr, capacity, nb_bits), :88-98); inserting more fails
 * with KMU_EOVERFLOW instead of degrading. */
int32_t kmu_count_create(kmu_ctx* ctx, uint32_t k, int32_t kmer_type, uint32_t count_bits, uint64_t capacity,
                         kmu_counter** counter);
void kmu_count_destroy(kmu_counter* counter);
uint64_t kmu_count_capacity(const kmu_counter* counter); /* slots allocated */
/* count_kmer / count_kmer_threaded_one_to_many (kmercount.rs:293-362, 881-974): every k-mer of every
 * sequence, canonical != 0 -> kmer.reverse_complement().min(kmer) first (:313,827,938) */
int32_t kmu_count_insert_seqs(kmu_ctx* ctx, kmu_counter* counter, const kmu_seqbatch* batch, int32_t canonical);
/* KmerCountT::insert_kmer for an array of compressed k-mer values (u32 or u64 by k-mer type) */
int32_t kmu_count_insert_kmers(kmu_ctx* ctx, kmu_counter* counter, const void* kmers, uint64_t n, int32_t on_device);
/* KmerCountT::get_count for an array of compressed k-mer values */
int32_t kmu_count_query(kmu_ctx* ctx, const kmu_counter* counter, const void* kmers, uint64_t n, uint32_t* counts,
                        int32_t on_device);
/* get_nb_distinct / get_nb_unique, total multiplicity, and hist256[c] = number of k-mers with
 * min(multiplicity, 255) == c (any output may be NULL) */
int32_t kmu_count_stats(kmu_ctx* ctx, const kmu_counter* counter, uint64_t* nb_distinct, uint64_t* nb_unique,
                        uint64_t* nb_inserted, uint64_t* hist256);
/* the k-mers with multiplicity >= min_count and their counts, unordered (what
 * threaded_dump_kmer_counter writes with min_count = 2, kmercount.rs:653-791); host buffers of `cap`
 * entries; *n_out = number found (KMU_EOVERFLOW if > cap) */
int32_t kmu_count_export(kmu_ctx* ctx, const kmu_counter* counter, uint32_t min_count, void* kmers, uint32_t* counts,
                         uint64_t cap, uint64_t* n_out);
/* the multiple k-mer dump of threaded_dump_kmer_counter / dump_in_file_multiple_kmer (kmercount.rs:139-145, 584-791):
 * `u32 0xcea2bbff | u8 kmer_size | u8 nb_bytes_by_count | u64 nb_kmer`, then `kmer.dump()` + count per k-mer seen at
 * least twice (each once, unordered).  count_bytes 1 or 2. */
int32_t kmu_count_dump_multiple(kmu_ctx* ctx, const kmu_counter* counter, const char* path, int32_t count_bytes,
                                uint64_t* nb_dumped);

/* KmerCountReload::load_multiple_kmers_from_file (kmercount.rs:1209-1351): header fields, the number of records in the file
 * (*n_read; the reference also reads to the end of file rather than trusting nb_declared) and, when kmers / counts are
 * given (cap entries each), the records: kmers[i] = the dumped word (`.0`; for Kmer32bit it carries k in its top four bits). */
int32_t kmu_count_reload_multiple(const char* path, uint32_t* kmer_size, uint32_t* count_bytes, uint64_t* nb_declared,
                                  uint64_t* kmers, uint32_t* counts, uint64_t cap, uint64_t* n_read);

/* partial registers of a counting table, for the multi-GPU whole-file sketch (every rank counts the keys it owns,
 * sketches them; the ranks merge with an allreduce-min on h and a key select): slots = m records {u64 h bits, u64 key},
 * h = largest f64 where no point fell below `bound`.  The table's keys are pre-keys (as inserted); hash_kind maps them. */
int32_t kmu_pmh3a_counter_slots(kmu_ctx* ctx, const kmu_counter* counter, int32_t hash_kind, uint32_t m, double bound,
                                void* slots, int32_t slots_on_device);
/* DispatchableT::dispatch (kmercount.rs:382-420) for a whole batch: all (canonical) compressed k-mer
 * values bucketed by owner = intNN_hash(value) % nparts.  kmers_out holds kmu_kmer_count() values,
 * bucket p first-to-last at offset sum(part_counts[0..p)).  This is the send side of the multi-GPU
 * exchange: bucket p goes to rank p, which feeds what it receives to kmu_count_insert_kmers. */
int32_t kmu_count_partition(kmu_ctx* ctx, const kmu_seqbatch* batch, uint32_t k, int32_t kmer_type, int32_t canonical,
                            uint32_t nparts, void* kmers_out, uint64_t* part_counts, int32_t out_on_device);

/* ---- host-side feeders and writers (no device work) --------------------------------------------
 * FASTA / FASTQ reader (plain or gzip-compressed, as needletail reads them): packs of ACCEPTED reads as one ASCII buffer + offsets, ready for
 * kmu_seqbatch_from_ascii.  A record holding any non-ACGT character is dropped and counted, as
 * parse_with_needletail (src/io.rs:12-72) and readblockseq (src/bin/datasketcher.rs:358-388) do. */
typedef struct kmu_fastx kmu_fastx;
int32_t kmu_fastx_open(const char* path, kmu_fastx** reader);
void kmu_fastx_close(kmu_fastx* reader);
int32_t kmu_fastx_next_pack(kmu_fastx* reader, uint64_t max_seqs, uint8_t* ascii, uint64_t ascii_cap, uint64_t* ascii_off,
                            uint64_t* nseq_out);
void kmu_fastx_stats(const kmu_fastx* reader, uint64_t* nb_read, uint64_t* nb_bad_read, uint64_t* nb_bases,
                     uint64_t* nb_bad_bases);
/* Multi-threaded form of the feeder (kmu_ingest.cu): a reader thread cuts the file into blocks at record boundaries, `nthreads`
 * parser threads (0 = all host cores) check and copy the accepted reads of a block into a pinned pack buffer, and the packs come
 * out in file order while the parsers work ahead -- the loop of src/bin/datasketcher.rs:236-300 with the sketch of pack i
 * overlapping the parsing of packs i + 1 ...  FASTA (sequences on any number of lines) and 4-line FASTQ, plain or gzip-compressed
 * (the inflation of a gzip stream stays serial); a FASTQ with sequences on several lines is refused (use kmu_fastx_open).
 *   kmu_ingest_next    : *ascii / *ascii_off / *nseq = the next pack (for kmu_seqbatch_from_ascii), valid until
 *                        kmu_ingest_release(token); *nseq == 0 at end of file. */
typedef struct kmu_ingest kmu_ingest;
int32_t kmu_ingest_open(const char* path, uint32_t nthreads, uint64_t block_bytes, kmu_ingest** reader);
int32_t kmu_ingest_next(kmu_ingest* reader, const uint8_t** ascii, const uint64_t** ascii_off, uint64_t* nseq, void** token);
int32_t kmu_ingest_release(kmu_ingest* reader, void* token);
void kmu_ingest_stats(const kmu_ingest* reader, uint64_t* nb_read, uint64_t* nb_bad_read, uint64_t* nb_bases,
                      uint64_t* nb_bad_bases);
void kmu_ingest_close(kmu_ingest* reader);
/* signature dump: `u32 0xceabeadd | u32 sig_size = 4 | u32 sketch_size | u32 kmer_size`, then sketch_size u32 per
 * sequence in input order (SeqSketcher::create_signature_dump, dump_signatures_block_u32,
 * src/sketching/seqsketchjaccard.rs:390-414, 577-585) */
typedef struct kmu_sigdump kmu_sigdump;
int32_t kmu_sigdump_create(const char* path, uint32_t sketch_size, uint32_t kmer_size, kmu_sigdump** dump);
int32_t kmu_sigdump_write(kmu_sigdump* dump, const uint32_t* sig, uint64_t nseq);
/* block signature dump: `u32 0xceabbadd | u8 sig_size = 4 | u32 sketch_size | u32 kmer_size | u32 block_size`, then per
 * sequence `u32 numseq | u32 nbblocks` and per block `u32 numseq | u32 numblock | sketch`
 * (BlockSeqSketcher::create_signature_dump / dump_blocks, src/sketching/seqblocksketch.rs:59-65, 172-226) */
int32_t kmu_blockdump_create(const char* path, uint32_t sketch_size, uint32_t kmer_size, uint32_t block_size,
                             kmu_sigdump** dump);
int32_t kmu_blockdump_write(kmu_sigdump* dump, const uint32_t* sig, const uint32_t* numseq, const uint32_t* numblock,
                            uint64_t nblocks);
int32_t kmu_sigdump_close(kmu_sigdump* dump);
/* SigSketchFileReader (seqsketchjaccard.rs:588-712): header, number of signatures, and (sig != NULL) signatures
 * first .. first + count */
int32_t kmu_sigdump_read(const char* path, uint32_t* sig_size, uint32_t* sketch_size, uint32_t* kmer_size, uint64_t* nsig,
                         uint32_t* sig, uint64_t first, uint64_t count);

/* ---- timing of the last compute call on the context (CUDA events on its stream) ----- */
typedef struct kmu_times {
    float kernel_ms;    /* device time of the compute kernels of the last call */
    float h2d_ms;       /* host->device copies of the last call (0 if none) */
    float d2h_ms;       /* device->host copies of the last call (0 if none) */
    uint64_t h2d_bytes; /* bytes copied host->device by the last call */
    uint64_t d2h_bytes; /* bytes copied device->host by the last call */
    uint64_t launches;  /* kernels launched by the last call */
    float host_ms;      /* wall time of the whole call on the host (one-shot host entry points) */
} kmu_times;
int32_t kmu_last_times(const kmu_ctx* ctx, kmu_times* out);

/* ---- optional per-launch profile of the last sketch call (CUDA events around every launch
 *      of the sketch kernel; switch on with kmu_ctx_set_profiling before the call) ---------- */
typedef struct kmu_launch_rec {
    int32_t mode;           /* 0 = u8 histogram, 1 = open-addressing table, 2 = one-pass kernel for long sequences */
    int32_t table_global;   /* table in global scratch instead of shared memory */
    uint32_t team_warps;    /* warps cooperating on one sequence */
    uint32_t teams_per_cta;
    uint32_t grid, block, smem_bytes;
    uint64_t nseq;          /* sequences handled by the launch */
    uint64_t nbases;        /* bases handled by the launch */
    uint64_t nk_max;        /* largest k-mer count the launch was sized for */
    float ms;               /* device time of the launch */
    uint32_t counter_idx;
    /* SM clocks summed over teams: [0] fetch/wait/init [1] pass 1 [2] fill [3] pass 2 [4] output */
    uint64_t phase_clocks[8];
} kmu_launch_rec;
int32_t kmu_ctx_set_profiling(kmu_ctx* ctx, int32_t on);
/* returns the number of launches of the last sketch call; fills at most cap records */
uint32_t kmu_last_launch_profile(kmu_ctx* ctx, kmu_launch_rec* out, uint32_t cap);

#ifdef __cplusplus
}
#endif
#endif /* KMERUTILS_B200_H */
