// kmu_host.h -- host-side internals shared by the C-ABI translation units: error reporting,
// growable device / pinned buffers, the context and the sequence batch.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/kmerutils_b200.h"
#include "kmu_kernels.h"

// sets the calling thread's last-error message (kmu_last_error) and returns `code`
int32_t kmu_fail(int32_t code, const char* fmt, ...);
#define fail kmu_fail

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) return fail(KMU_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)

constexpr size_t SMEM_BUDGET = 227 * 1024 - 1024;  // opt-in shared memory per CTA on sm_100 minus static use
constexpr size_t SEQ_ALIGN = 16;
constexpr size_t TAIL_SLACK = 64;

// growable device buffer
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMallocHost(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

// double-buffered host <-> device pipeline of the one-shot host entry points
struct HostPipe {
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    cudaEvent_t in_begin[2]{}, in_done[2]{}, compute_done[2]{}, out_begin[2]{}, out_done[2]{};
    DevBuf packed[2], meta[2], sig[2], order[2], cursor[2];
    PinnedBuf stage[2], meta_host[2];
};

struct kmu_ctx {
    HostPipe pipe;
    int device = 0;
    cudaStream_t stream = nullptr;
    std::mutex mu;
    uint64_t launches = 0;
    kmu_times last{};
    cudaEvent_t ev[6]{};  // k0 k1 h0 h1 d0 d1
    int sm_count = 148;
    // scratch
    DevBuf order, counters, table_scratch, slot_scratch, overflow, sig_dev, misc;
    DevBuf ascii_dev, ascii_off_dev, ascii_bad_dev;  // staging of kmu_seqbatch_from_ascii / _from_aa (grow-only: no cudaMalloc per pack)
    // kmu_sketch_pmh3a_groups: genomes are sketched on several streams at once, each with its own table and slots
    static constexpr int GROUP_STREAMS = 4;
    cudaStream_t group_stream[GROUP_STREAMS]{};
    cudaEvent_t group_ev[GROUP_STREAMS + 1]{};
    DevBuf group_table[GROUP_STREAMS], group_slots[GROUP_STREAMS], group_seen[GROUP_STREAMS];
    DevBuf part_fine;  // level-2 slabs of the two-phase counting insertion (kmu_capi_count.cu)
    DevBuf whole_table, items_slots;  // whole-file ProbMinHash3a: counting table + global slots
    int p2p_grid = 0;  // kmu_count_partition_counts -> _scatter hand-over
    uint32_t p2p_nparts = 0;
    const kmu_seqbatch* p2p_batch = nullptr;
    DevBuf counters_alt, overflow_alt;  // second set of sketch counters / redo list (host pipeline: two chunks in flight)
    bool table_scratch_clean = false;
    PinnedBuf pinned, pinned_small;
    cudaEvent_t phase_ev[2]{};  // end of the main sketch launches of a chunk (host pipeline)
    DevBuf smh_memo;  // SuperMinHash point 0 per pre-key (small key spaces), valid for the parameters below
    uint32_t smh_memo_k = 0, smh_memo_m = 0;
    int smh_memo_type = -1, smh_memo_hash = -1, smh_memo_hasher = -1, smh_memo_bytes = 0;
    cudaStream_t aux_stream = nullptr;  // the few-CTA launch of the very long sequences runs beside the main launches
    cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
    cudaStream_t redo_stream = nullptr;  // host pipeline: the redo launch of a chunk runs beside the next chunk
    bool redo_on_side = false, redo_on_main = false;  // ... where the last phase-2 call put its redo launch, if any
    // first-point table of the ProbMinHash3a kernels (small key spaces), see kmu_pmh3a.cu
    DevBuf memo;
    uint32_t memo_k = 0, memo_m = 0;
    int memo_type = -1, memo_hash = -1;
    // optional per-launch profile of the last sketch call
    bool profiling = false;
    std::vector<cudaEvent_t> lev;
    std::vector<kmu_launch_rec> lrec;
};

// processing order of a batch for the sketch kernels (longest sequences first), per k
struct OrderCache {
    uint32_t k = 0;
    std::vector<unsigned long long> hist, cursor;
    uint64_t nk_longest = 0;
    DevBuf order, cursor_dev;
};
struct OctaveClass {
    uint64_t first, count;  // range of the order array
    uint64_t nk_max;        // largest k-mer count in the class (the last class also holds sequences without k-mers)
};

// output offsets of the materialising kernels (exclusive prefix of the k-mer counts, nseq + 1 entries), per k: computed on
// the host and uploaded once per (batch, k) -- 746 333 reads cost ~1 ms of host loop + copy + synchronisation per call otherwise
struct KmerOffCache {
    uint32_t k = 0;
    uint64_t total = 0;
    std::vector<uint64_t> host;
    DevBuf dev;
};

struct kmu_seqbatch {
    mutable OrderCache order_cache;
    mutable KmerOffCache koff_cache;
    int device = 0;
    uint8_t* packed = nullptr;
    uint64_t* byte_off = nullptr;
    uint64_t* nbases = nullptr;
    uint64_t nseq = 0;
    uint64_t packed_bytes = 0;  // without the tail slack
    uint64_t total_bases = 0;
    bool owns = true;           // false: the buffers belong to a context arena or to a parent batch
    bool owns_byte_off = false; // a view: byte_off is the view's own (rebased) array
    int alphabet = 0;           // 0: DNA, 2 bits per base packed; 1: amino acids, one 5-bit code per byte
    std::vector<uint64_t> h_nbases;
    std::vector<uint64_t> h_byte_off;
    // sum over the sequences of max(0, L - k + 1), cached for the last k asked (the batch is immutable)
    mutable uint32_t kmer_total_k = 0;
    mutable uint64_t kmer_total = 0;
    mutable uint64_t min_nbases = ~0ull;  // shortest sequence (computed on first use)
    mutable uint64_t max_nbases = ~0ull;  // longest sequence (computed on first use)
    uint64_t longest() const {
        if (max_nbases == ~0ull) {
            uint64_t m = 0;
            for (uint64_t L : h_nbases) m = L > m ? L : m;
            max_nbases = m;
        }
        return max_nbases;
    }
    uint64_t kmer_count(uint32_t k) const {
        if (k == 0) return 0;
        if (min_nbases == ~0ull) {
            uint64_t m = ~0ull - 1;
            for (uint64_t L : h_nbases) m = L < m ? L : m;
            min_nbases = m;
        }
        if (nseq && min_nbases >= k) return total_bases - nseq * (uint64_t)(k - 1);  // every sequence holds a k-mer
        if (kmer_total_k != k) {
            uint64_t n = 0;
            for (uint64_t L : h_nbases) n += L >= k ? L - k + 1 : 0;
            kmer_total = n;
            kmer_total_k = k;
        }
        return kmer_total;
    }
};

struct ScopedDevice {
    int prev = -1;
    explicit ScopedDevice(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
    }
    ~ScopedDevice() {
        int cur = -1;
        cudaGetDevice(&cur);
        if (prev >= 0 && cur != prev) cudaSetDevice(prev);
    }
};

inline uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }
inline bool kmer_type_is_aa(int type) { return type == KMU_KMERAA32 || type == KMU_KMERAA64; }
inline bool kmer_type_is_u64(int type) { return type == KMU_KMER64 || type == KMU_KMERAA64; }
// k-mer type, k, hash closure and batch alphabet must agree; returns KMU_OK or sets the error
int32_t kmu_check_kmer_args(const kmu_seqbatch* b, uint32_t k, int kmer_type, int hash_kind);

// does the k-mer type accept this k? (the reference panics otherwise)
inline bool kmer_type_accepts(uint32_t k, int type) {
    switch (type) {
        case KMU_KMER32: return k >= 1 && k <= 14;   // src/base/kmergenerator.rs:311, kmer32bit.rs:68-76
        case KMU_KMER16B32: return k == 16;          // src/base/kmergenerator.rs:218-220
        case KMU_KMER64: return k >= 1 && k <= 32;   // src/base/kmergenerator.rs:415
        case KMU_KMERAA32: return k >= 1 && k <= 6;  // src/aautils/kmeraa.rs:212-214,727-732
        case KMU_KMERAA64: return k >= 1 && k <= 12; // src/aautils/kmeraa.rs:822-824
        default: return false;
    }
}

int32_t kmu_ensure_order(kmu_ctx* ctx, const kmu_seqbatch* b, uint32_t k, uint64_t* launches);
std::vector<OctaveClass> kmu_octave_classes(const kmu_seqbatch* b);
