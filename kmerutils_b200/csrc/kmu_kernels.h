// kmu_kernels.h -- host-visible launch parameters and launchers of the CUDA kernels.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace kmu {

// ExpRestricted01 constants, computed on the host with libm
struct Exp01Params {
    double lambda, c1, c2, c3;
};

// one sketch slot: {bit pattern of the winning point h, key that produced it}
struct alignas(16) Slot {
    uint64_t hbits;
    uint64_t key;
};

// device view of a sequence batch (see include/kmerutils_b200.h "sequence batches")
struct SeqView {
    const uint8_t* packed;     // every sequence starts 16-byte aligned; >= 64 bytes of slack at the end
    const uint64_t* byte_off;  // nseq
    const uint64_t* nbases;    // nseq
    uint64_t nseq;
};

struct Pmh3aParams {
    const uint8_t* packed;
    const uint64_t* byte_off;
    const uint64_t* nbases;
    const uint32_t* order;  // sequence ids in processing order (descending length)
    uint64_t first, count;  // this launch handles order[first .. first+count)
    unsigned long long* work_counter;
    uint32_t k;
    int kmer_type;
    int hash_kind;
    uint32_t m;
    uint32_t slot_thresh;  // 2^32 mod m : rejection threshold of UniformUsize (App. A.2)
    Exp01Params e;
    void* sig;  // nseq * m values of V
    // team / shared memory geometry
    uint32_t team_warps;
    uint32_t team_smem_bytes;   // bytes of shared memory per team
    uint32_t regionA_bytes;     // histogram / table bytes per team (multiple of 16)
    uint32_t slots_smem_bytes;  // 16 * m rounded up, or 0 when the slots live in slot_scratch
    Slot* slot_scratch;
    uint8_t* table_scratch;  // global table scratch (zeroed), table_scratch_entries entries per team
    uint64_t table_scratch_entries;
    unsigned long long* overflow_count;
    uint32_t* overflow_list;
    // first two points of every pre-key, 2 x {x bits lo, x bits hi, slot, hashed key} (u32 key types with a
    // small key space), or nullptr
    const void* memo_fast;
    // speculative qmax start: B = spec_factor / nk with spec_factor = m ln(m / 1e-4); 0 = off
    uint32_t speculate;
    double spec_factor;
    // TMA staging: bytes per staging buffer (two per team); 0 disables staging
    uint32_t stage_bytes;
    // profiling only: 8 counters per launch, SM clocks spent per phase by thread 0 of every team
    unsigned long long* phase_clocks;
};

constexpr size_t PMH3A_TEAM_SHARED_BYTES = 80;

cudaError_t launch_pmh3a_memo(const Pmh3aParams& P, void* fast, uint32_t nkeys, cudaStream_t stream);
size_t pmh3a_qitem_bytes(bool key64);
size_t pmh3a_entry_bytes(bool key64);
cudaError_t launch_pmh3a(const Pmh3aParams& P, bool key64, int mode, int grid, int block, size_t smem,
                         cudaStream_t stream);

// one-pass kernel for long sequences over a small key space (kmu_pmh3a_direct.cu); P.regionA_bytes = 4^k,
// P.slots_smem_bytes = 20 m rounded to 16, P.memo_fast set, P.order/first/count = the sequences to sketch
size_t pmh3a_direct_smem_bytes(uint32_t k, uint32_t m, int variant = 0);
size_t pmh3a_direct_hist_bytes(uint32_t k, int variant);
int pmh3a_direct_ctas_per_sm(int variant);
int pmh3a_direct_threads(int variant);
cudaError_t launch_pmh3a_direct(const Pmh3aParams& P, int grid, int variant, cudaStream_t stream);

// ---- batch utilities (kmu_batch.cu) ------------------------------------------------------
// synthetic packed bases: base j of sequence i = SplitMix64 stream `seed` output first_base[i] + j
cudaError_t launch_synth_packed(uint8_t* packed, const uint64_t* byte_off, const uint64_t* nbases,
                                const uint64_t* first_base, uint64_t nseq, uint64_t total_words, uint64_t seed,
                                cudaStream_t stream);
// ASCII -> 2 bit.  mode 0: strict (invalid char counted, packed as A); mode 1: drop invalid
cudaError_t launch_count_invalid(const uint8_t* ascii, const uint64_t* ascii_off, uint64_t nseq, uint64_t* invalid,
                                 cudaStream_t stream);
cudaError_t launch_pack_ascii(const uint8_t* ascii, const uint64_t* ascii_off, const uint64_t* byte_off,
                              const uint64_t* nbases, uint64_t nseq, int drop_invalid, uint8_t* packed,
                              cudaStream_t stream);
// amino-acid sequences: one 5-bit code per byte (kmeraa.rs:85-109)
cudaError_t launch_aa_count_invalid(const uint8_t* ascii, const uint64_t* ascii_off, uint64_t nseq, uint64_t* invalid,
                                    cudaStream_t stream);
cudaError_t launch_aa_encode(const uint8_t* ascii, const uint64_t* ascii_off, const uint64_t* byte_off, uint64_t nseq,
                             uint8_t* codes, cudaStream_t stream);
cudaError_t launch_synth_aa(uint8_t* codes, const uint64_t* byte_off, const uint64_t* nres, const uint64_t* first_res,
                            uint64_t nseq, uint64_t total_bytes, uint64_t seed, cudaStream_t stream);
// reads sampled from a packed genome with substitution errors; every read takes words_per_read u32
cudaError_t launch_sample_reads(const uint8_t* genome, uint64_t glen, uint64_t seed, uint64_t first_read, uint64_t nreads,
                                uint32_t read_len, uint32_t err_ppm, uint32_t words_per_read, uint8_t* out,
                                cudaStream_t stream);
cudaError_t launch_slice_copy(const uint8_t* src, const uint64_t* src_byte_off, const uint64_t* seq_idx, const uint64_t* begin,
                              const uint64_t* dst_byte_off, const uint64_t* dst_len, uint64_t nslices, uint64_t total_words,
                              int alphabet, uint8_t* dst, cudaStream_t stream);
// length classes: bucket = 8 per octave of the k-mer count, bucket 0 = longest
constexpr int LEN_BUCKETS = 512;
cudaError_t launch_len_hist(const uint64_t* nbases, uint64_t nseq, uint32_t k, unsigned long long* hist,
                            cudaStream_t stream);
cudaError_t launch_len_scatter(const uint64_t* nbases, uint64_t nseq, uint32_t k, unsigned long long* cursor,
                               uint32_t* order, cudaStream_t stream);
inline int len_bucket_host(uint64_t nk) {
    if (nk == 0) return LEN_BUCKETS - 1;
    int e = 63 - __builtin_clzll(nk);
    int frac = e >= 3 ? (int)((nk >> (e - 3)) & 7) : (int)((nk << (3 - e)) & 7);
    return LEN_BUCKETS - 1 - (e * 8 + frac);
}
// smallest k-mer count that falls in `bucket`
inline uint64_t len_bucket_min_nk(int bucket) {
    int key = LEN_BUCKETS - 1 - bucket;
    int e = key / 8, frac = key % 8;
    if (e >= 3) return (uint64_t)(8 + frac) << (e - 3);
    return ((uint64_t)(8 + frac)) >> (3 - e);
}

// ---- k-mer generation / ntHash (kmu_extract.cu) -----------------------------------------
cudaError_t launch_kmer_offsets(const uint64_t* nbases, uint64_t nseq, uint32_t k, uint64_t* out_off,
                                cudaStream_t stream);
cudaError_t launch_generate_kmers(const SeqView& b, uint64_t total_bytes, uint32_t k, int kmer_type, int hash_kind,
                                  const uint64_t* out_off, void* out, cudaStream_t stream);
cudaError_t launch_nthash(const SeqView& b, uint64_t total_bytes, uint32_t k, uint32_t n_multi, const uint64_t* out_off,
                          uint64_t* out_hash, uint8_t* out_strand, cudaStream_t stream);

inline bool hash_kind_is_canonical_host(int hash_kind) { return hash_kind == 2 || hash_kind == 3; }  // KMU_HASH_CANON_*

// ---- ProbMinHash3a over a weighted set (kmu_pmh3a_items.cu) -----------------------------------
struct Pmh3aItemsParams {
    const void* keys;        // SRC 0: n keys (u32 / u64)
    const double* weights;   // SRC 0: n weights
    const void* table;       // SRC 1 / 2: counting table slots, n = number of slots
    const unsigned long long* special;  // SRC 2: multiplicity of the key ~0
    uint64_t n;
    uint32_t k;
    int kmer_type, hash_kind;  // SRC 1 / 2: table keys are pre-keys, mapped through the hash closure
    uint32_t m, slot_thresh;
    Exp01Params e;
    double bound;
    int slots_in_smem;
    Slot* global_slots;
};
// shared memory of the item kernel beside the slots: one queue of 64 (key, 1 / weight) items per warp; the slots live in
// shared memory when 16 m + this fits (callers decide with P.slots_in_smem and pass the slot bytes only)
constexpr size_t PMH3A_ITEMS_QUEUE_BYTES = 32 * 64 * 16;
cudaError_t launch_pmh3a_items(const Pmh3aItemsParams& P, bool key64, int src, int grid, size_t smem, cudaStream_t st);
cudaError_t launch_pmh3a_items_init(Slot* slots, uint32_t m, cudaStream_t st);
cudaError_t launch_pmh3a_items_finish(const Slot* slots, uint32_t m, bool key64, void* sig, unsigned long long* max_hbits,
                                      cudaStream_t st);

// ---- SuperMinHash (kmu_smh.cu) ----------------------------------------------------------------
constexpr size_t SMH_QUEUE_BYTES = 32 * 64 * 8;
struct SmhParams {
    const uint8_t* packed;
    const uint64_t* byte_off;
    const uint64_t* nbases;
    const uint32_t* order;
    uint64_t first, count;
    unsigned long long* work_counter;
    uint32_t k;
    int kmer_type, hash_kind;
    uint32_t m;
    int hasher;      // 0 NoHashHasher, 1 FnvHasher
    void* sig;       // nseq * m values of S
    uint32_t team_warps, team_smem_bytes;
    double ln_term;  // ln(1e4 m)
    uint32_t value_cut;  // 1: the CTA's shared memory ends with one queue of 64 keys per warp (SMH_QUEUE_BYTES): long sequences use the value cut
    unsigned long long* slow_count;
    uint32_t* slow_list;
    uint8_t* scratch;  // exact path
    uint64_t scratch_per_warp;
    // small key spaces (u32 DNA k-mers, k <= 8): point 0 of every pre-key, {value bits lo, value bits hi, slot, 0}; or nullptr
    const void* memo;
};
cudaError_t launch_smh_memo(const SmhParams& P, bool f64, void* memo, uint32_t nkeys, cudaStream_t st);
cudaError_t launch_smh_fast(const SmhParams& P, bool key64, bool f64, int grid, int block, size_t smem, cudaStream_t st);
cudaError_t launch_smh_whole_cut(const SmhParams& P, bool key64, bool f64, const SeqView& b, uint64_t total_bytes, double cut,
                                 void* gslots, int sm_count, cudaStream_t st);
cudaError_t launch_smh_whole(const SmhParams& P, bool key64, bool f64, const SeqView& b, uint64_t total_bytes, uint32_t a_spec,
                             void* gslots, int grid, size_t smem, cudaStream_t st);
cudaError_t launch_smh_fill_large(void* slots, uint32_t m, bool f64, cudaStream_t st);
cudaError_t launch_smh_colmin(const void* rows, uint64_t nseq, uint32_t m, bool f64, void* out, cudaStream_t st);
cudaError_t launch_smh_exact(const SmhParams& P, bool key64, bool f64, int grid, cudaStream_t st);

// ---- SetSketch (kmu_setsketch.cu) ---------------------------------------------------------------
constexpr size_t SSK_QUEUE_BYTES = 32 * 64 * 8;  // team kernel: one queue of 64 keys per warp, after the ziggurat tables
struct SskConsts {
    double a, inva, lnb;  // SetSketchParams.a, 1 / a, ln(b) (deterministic log)
    double ln_term;       // ln(1e4 m)
    double inva_m0;       // inva / m: the spacing of an item's first point (what the loop computes for j = 0)
    // per-sequence speculation (a failed sketch is redone from the level it reached, so this only trades the cost of the
    // points below the cut against the cost of the redo): D = spec_dfrac * distinct(nk), failure odds ~ m e^-spec_ln
    double spec_ln, spec_dfrac, spec_keyspace;  // spec_keyspace: number of possible keys, 0 = far more than any nk
    uint32_t m;
    int iq1;              // q + 1
};
struct SskParams {
    const uint8_t* packed;
    const uint64_t* byte_off;
    const uint64_t* nbases;
    const uint32_t* order;
    uint64_t first, count;
    unsigned long long* work_counter;
    uint32_t k;
    int kmer_type, hash_kind;
    SskConsts C;
    void* sig;      // registers as u16 / u32 / u64 (sig_bytes 2 / 4 / 8)
    int sig_bytes;
    uint32_t team_warps, team_smem_bytes;
    const uint32_t* kspec_in;  // per-sequence level for a redo launch, or nullptr
    uint32_t* kmin_out;        // smallest register reached by a failed speculation
    unsigned long long* slow_count;
    uint32_t* slow_list;
    unsigned long long* exact_count;
    uint32_t* exact_list;
    uint32_t exact_nk_max;
    uint32_t* whole_regs;  // whole-batch mode: m global registers
    uint8_t* scratch;      // exact path
    uint64_t scratch_per_warp;
    int group;             // exact path: all listed sequences go into ONE sketch
};
cudaError_t launch_ssk_team(const SskParams& P, int grid, int block, size_t smem, cudaStream_t st);
cudaError_t launch_ssk_whole(const SskParams& P, const SeqView& b, uint64_t total_bytes, uint32_t kspec, double xcut,
                             int grid, size_t smem, cudaStream_t st);
cudaError_t launch_ssk_store(const uint32_t* regs, uint32_t m, void* out, int sig_bytes, cudaStream_t st);
cudaError_t launch_ssk_exact(const SskParams& P, int grid, cudaStream_t st);

// ---- counting table (kmu_count.cu) ----------------------------------------------------------
struct CountTable {
    void* slots;                  // u32 keys: u64 slot (key << 32 | count); u64 keys: 16-byte slot {key, count}
    uint64_t capmask;             // capacity - 1 (capacity is a power of two)
    unsigned long long* special;  // multiplicity of the one u64 key equal to the empty sentinel
    unsigned long long* overflow; // set to 1 when an insertion found no free slot
};
cudaError_t launch_count_init(const CountTable& t, bool key64, int sm_count, cudaStream_t st);
cudaError_t launch_pmh3a_prefilter(const SeqView& b, uint64_t total_bytes, uint32_t k, bool key64, bool canonical, int kmer_type,
                                   int hash_kind, double bound, double c1, const CountTable& t, uint32_t* seen, uint64_t bitmask,
                                   int sm_count, cudaStream_t st);
cudaError_t launch_count_insert_seqs(const SeqView& b, uint64_t total_bytes, uint32_t k, bool key64, bool canonical,
                                     const CountTable& t, int sm_count, cudaStream_t st, uint64_t byte_begin = 0);
cudaError_t launch_count_insert_keys(const void* keys, uint64_t n, bool key64, const CountTable& t, int sm_count,
                                     cudaStream_t st);
cudaError_t launch_count_query(const void* keys, uint64_t n, bool key64, const CountTable& t, uint32_t max_count,
                               uint32_t* out, int sm_count, cudaStream_t st);
// stats: 3 + 256 counters {distinct (unused here), unique (unused here), total multiplicity, hist[min(count,255)]}
cudaError_t launch_count_stats(const CountTable& t, bool key64, unsigned long long* stats, int sm_count, cudaStream_t st);
cudaError_t launch_count_export(const CountTable& t, bool key64, uint32_t min_count, void* keys, uint32_t* counts,
                                unsigned long long* cursor, uint64_t cap, int sm_count, cudaStream_t st);
int count_partition_grid(uint64_t total_bytes, int sm_count);
// block_counts: nparts * grid entries (scratch); part_totals: 2 * nparts entries; out: all k-mers, part major
cudaError_t launch_count_partition(const SeqView& b, uint64_t total_bytes, uint32_t k, bool key64, bool canonical,
                                   uint32_t nparts, int grid, unsigned long long* block_counts,
                                   unsigned long long* part_totals, void* out, cudaStream_t st);
cudaError_t launch_count_partition_counts(const SeqView& b, uint64_t total_bytes, uint32_t k, bool key64, bool canonical,
                                          uint32_t nparts, int grid, unsigned long long* block_counts,
                                          unsigned long long* part_totals, cudaStream_t st);
cudaError_t launch_count_partition_scatter(const SeqView& b, uint64_t total_bytes, uint32_t k, bool key64, bool canonical,
                                           uint32_t nparts, int grid, unsigned long long* block_counts,
                                           const unsigned long long* part_base, void* const* dests, cudaStream_t st);
cudaError_t launch_count_partition_by_region(const SeqView& b, uint64_t total_bytes, uint32_t k, bool key64, bool canonical,
                                             const CountTable& t, uint32_t nparts, int grid, unsigned long long* block_counts,
                                             unsigned long long* part_totals, void* out, cudaStream_t st);


// ---- two-phase insertion / fused exchange (kmu_count_part.cu) ---------------------------------
// bucket of a key = owner * nregions + region; owner = intNN_hash(key) % nowners (DispatchableT::dispatch,
// kmercount.rs:382-420), region = (fmix64(key) & capmask) >> shift = the slice of the owner's table the key hashes to.
// Bucket (o, r) is appended to the slab at dests[o] + (r * nsend + self) * slab_cap (elements).
struct PartGeom {
    uint32_t nowners;   // >= 1
    uint32_t nregions;  // power of two; nowners * nregions <= 4096
    uint64_t capmask;   // capacity - 1 of an owner's table (all owners use the same capacity)
    uint32_t shift;     // log2(slots per region)
    uint32_t nsend, self;
    uint64_t slab_cap;  // keys per slab (< 2^32)
};
// key-array form of the partition: nseg segments `stride` keys apart, segment s holding min(counts[s * count_stride], nkeys)
// keys (counts == nullptr: one array of nkeys keys); skip_flag != 0 -> the kernel does nothing
struct KeySegs {
    uint64_t stride = 0;
    uint32_t nseg = 1;
    const unsigned long long* counts = nullptr;
    uint64_t count_stride = 0;
    const unsigned long long* skip_flag = nullptr;
};
size_t count_part_smem_bytes(bool key64, uint32_t nbuckets);
cudaError_t launch_count_part_seqs(const SeqView& b, uint64_t byte_begin, uint64_t byte_end, uint64_t total_bytes, uint32_t k,
                                   bool key64, bool canonical, const PartGeom& g, void* const* dests,
                                   unsigned long long* cursors, unsigned long long* flag, int sm_count, cudaStream_t st);
cudaError_t launch_count_part_keys(const void* keys, uint64_t nkeys, const KeySegs& segs, bool key64, const PartGeom& g,
                                   void* const* dests, unsigned long long* cursors, unsigned long long* flag, int sm_count,
                                   cudaStream_t st);
cudaError_t launch_count_insert_slabs(const void* slabs, uint64_t slab_cap, uint32_t nregions, uint32_t nsend,
                                      const unsigned long long* counts, const CountTable& t, bool key64, uint32_t shift,
                                      const unsigned long long* skip_flag, bool prefetch, unsigned int* done, int sm_count,
                                      cudaStream_t st, uint32_t region0 = 0, uint64_t count_stride = 0);

}  // namespace kmu
