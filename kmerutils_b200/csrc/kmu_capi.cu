// kmu_capi.cu -- the extern "C" boundary (include/kmerutils_b200.h): contexts, sequence
// batches resident in HBM, launch policy of the kernels.  No CPU fallback anywhere: every
// compute entry point needs a CUDA device and fails with KMU_ECUDA otherwise.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <mutex>
#include <string>
#include <thread>
#include <functional>
#include <condition_variable>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/kmerutils_b200.h"
#include "kmu_host.h"
#include "kmu_kernels.h"

namespace {
thread_local std::string g_err;
}  // namespace

#undef fail
int32_t kmu_fail(int32_t code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define fail kmu_fail

int32_t kmu_check_kmer_args(const kmu_seqbatch* b, uint32_t k, int kmer_type, int hash_kind) {
    if (kmer_type < KMU_KMER32 || kmer_type > KMU_KMERAA64) return fail(KMU_EINVAL, "unknown kmer type %d", kmer_type);
    if (!kmer_type_accepts(k, kmer_type))
        return fail(KMU_EINVAL, "KmerSeqIterator cannot support kmer size %u for kmer type %d", k, kmer_type);
    if (hash_kind < 0 || hash_kind > KMU_HASH_INVHASH) return fail(KMU_EINVAL, "unknown hash kind %d", hash_kind);
    const bool aa = kmer_type_is_aa(kmer_type);
    if (b && aa != (b->alphabet == 1))
        return fail(KMU_EINVAL, "kmer type %d does not match the alphabet of the batch (%s)", kmer_type,
                    b->alphabet ? "amino acids" : "DNA");
    if (aa && (hash_kind == KMU_HASH_CANON_INVHASH || hash_kind == KMU_HASH_CANON_RAW))
        return fail(KMU_EINVAL, "amino-acid k-mers have no reverse complement (kmeraa.rs:185-187 panics)");
    return KMU_OK;
}

namespace {

// byte layout of a batch: every sequence on a 16-byte boundary
uint64_t layout_offsets(const uint64_t* nbases, uint64_t nseq, std::vector<uint64_t>& off, int alphabet = 0) {
    off.resize(nseq);
    uint64_t cur = 0;
    for (uint64_t i = 0; i < nseq; ++i) {
        off[i] = cur;
        cur += align_up(alphabet ? nbases[i] : (nbases[i] + 3) / 4, SEQ_ALIGN);
    }
    return cur;
}

int32_t batch_alloc(kmu_ctx* ctx, const uint64_t* nbases, uint64_t nseq, kmu_seqbatch** out, int alphabet = 0) {
    auto* b = new kmu_seqbatch();
    b->device = ctx->device;
    b->nseq = nseq;
    b->alphabet = alphabet;
    b->h_nbases.assign(nbases, nbases + nseq);
    b->packed_bytes = layout_offsets(nbases, nseq, b->h_byte_off, alphabet);
    uint64_t shortest = ~0ull - 1, longest = 0;  // kept with the batch: kmer_count(k) is then O(1) whenever every sequence holds a k-mer
    for (uint64_t i = 0; i < nseq; ++i) {
        b->total_bases += nbases[i];
        shortest = nbases[i] < shortest ? nbases[i] : shortest;
        longest = nbases[i] > longest ? nbases[i] : longest;
    }
    b->min_nbases = shortest;
    b->max_nbases = longest;
    cudaError_t e = cudaMalloc((void**)&b->packed, b->packed_bytes + TAIL_SLACK);
    if (e == cudaSuccess) e = cudaMalloc((void**)&b->byte_off, sizeof(uint64_t) * (nseq + 1));
    if (e == cudaSuccess) e = cudaMalloc((void**)&b->nbases, sizeof(uint64_t) * (nseq + 1));
    if (e != cudaSuccess) {
        kmu_seqbatch_destroy(b);
        return fail(KMU_ENOMEM, "cudaMalloc of a %llu byte batch failed: %s", (unsigned long long)b->packed_bytes,
                    cudaGetErrorString(e));
    }
    *out = b;
    return KMU_OK;
}

int32_t batch_upload_meta(kmu_ctx* ctx, kmu_seqbatch* b) {
    if (b->nseq == 0) return KMU_OK;
    CUDA_TRY(cudaMemcpyAsync(b->byte_off, b->h_byte_off.data(), sizeof(uint64_t) * b->nseq, cudaMemcpyHostToDevice,
                             ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(b->nbases, b->h_nbases.data(), sizeof(uint64_t) * b->nseq, cudaMemcpyHostToDevice,
                             ctx->stream));
    return KMU_OK;
}

// A small persistent pool for the host-side gathers (one batch of separately allocated sequences is ~750 000 memcpy calls
// in a dozen chunks: starting 16 threads per chunk cost more than the copies of a small chunk).
class CopyPool {
  public:
    static CopyPool& get() {
        static CopyPool pool;
        return pool;
    }
    unsigned size() const { return (unsigned)workers.size() + 1; }
    // runs job(t) for t in [0, njobs) on the pool's threads and the caller; returns when all are done
    void run(unsigned njobs, const std::function<void(unsigned)>& job) {
        if (njobs <= 1 || workers.empty()) {
            for (unsigned t = 0; t < njobs; ++t) job(t);
            return;
        }
        std::lock_guard<std::mutex> one_at_a_time(run_mu);
        {
            std::lock_guard<std::mutex> lk(mu);
            cur_job = &job;
            cur_njobs = njobs;
            next = 0;
            pending = njobs;
            ++generation;
        }
        cv.notify_all();
        work();
        std::unique_lock<std::mutex> lk(mu);
        done_cv.wait(lk, [&] { return pending == 0; });
        cur_job = nullptr;
    }

  private:
    CopyPool() {
        const unsigned nt = std::min<unsigned>(16, std::max(1u, std::thread::hardware_concurrency()));
        for (unsigned t = 1; t < nt; ++t) workers.emplace_back([this] { loop(); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv.notify_all();
        for (auto& w : workers) w.join();
    }
    void work() {
        for (;;) {
            unsigned t;
            const std::function<void(unsigned)>* job;
            {
                std::lock_guard<std::mutex> lk(mu);
                if (!cur_job || next >= cur_njobs) return;
                t = next++;
                job = cur_job;
            }
            (*job)(t);
            std::lock_guard<std::mutex> lk(mu);
            if (--pending == 0) done_cv.notify_all();
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || generation != seen; });
                if (stop) return;
                seen = generation;
            }
            work();
        }
    }
    std::vector<std::thread> workers;
    std::mutex mu, run_mu;
    std::condition_variable cv, done_cv;
    const std::function<void(unsigned)>* cur_job = nullptr;
    unsigned cur_njobs = 0, next = 0, pending = 0;
    uint64_t generation = 0;
    bool stop = false;
};

// dst_off: byte offset of every sequence in dst (ascending).  The sequences are dealt out in pieces of equal BYTES.
void parallel_copy(uint8_t* dst, const std::vector<uint64_t>& dst_off, const uint8_t* const* ptrs, const uint8_t* base,
                   const uint64_t* src_off, const uint64_t* nbases, uint64_t nseq) {
    auto work = [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; ++i) {
            const uint8_t* src = ptrs ? ptrs[i] : base + src_off[i];
            uint64_t nb = (nbases[i] + 3) / 4;
            std::memcpy(dst + dst_off[i], src, nb);
            uint64_t padded = align_up(nb, SEQ_ALIGN);
            if (padded > nb) std::memset(dst + dst_off[i] + nb, 0, padded - nb);
        }
    };
    CopyPool& pool = CopyPool::get();
    if (nseq < 4096 || pool.size() == 1) {
        work(0, nseq);
        return;
    }
    // pieces of ~equal bytes, four per thread (the threads take the next piece when they are done with one)
    const unsigned npieces = pool.size() * 4;
    const uint64_t total = dst_off[nseq - 1] + align_up((nbases[nseq - 1] + 3) / 4, SEQ_ALIGN) - dst_off[0];
    std::vector<uint64_t> cut(npieces + 1, nseq);
    cut[0] = 0;
    for (unsigned p = 1; p < npieces; ++p) {
        const uint64_t target = dst_off[0] + total / npieces * p;
        cut[p] = (uint64_t)(std::lower_bound(dst_off.begin(), dst_off.begin() + nseq, target) - dst_off.begin());
    }
    pool.run(npieces, [&](unsigned p) { work(cut[p], cut[p + 1]); });
}

}  // namespace

// =====================================================================================
extern "C" {

const char* kmu_last_error(void) { return g_err.c_str(); }
const char* kmu_version(void) { return "kmerutils_b200 0.1.0 (sm_100a)"; }

int32_t kmu_ctx_create(int32_t device, kmu_ctx** out) {
    if (!out) return fail(KMU_EINVAL, "ctx output pointer is null");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(KMU_ECUDA, "no CUDA device (%s): this library has no CPU path", cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(KMU_EINVAL, "device %d out of range (0..%d)", device, n - 1);
    ScopedDevice sd(device);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(KMU_ECUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    auto* c = new kmu_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    for (int i = 0; i < 6 && e == cudaSuccess; ++i) e = cudaEventCreate(&c->ev[i]);
    if (e != cudaSuccess) {
        kmu_ctx_destroy(c);
        return fail(KMU_ECUDA, "context creation failed: %s", cudaGetErrorString(e));
    }
    *out = c;
    return KMU_OK;
}

void kmu_ctx_destroy(kmu_ctx* c) {
    if (!c) return;
    ScopedDevice sd(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (auto* b : {&c->order, &c->counters, &c->table_scratch, &c->slot_scratch, &c->overflow, &c->sig_dev, &c->misc, &c->memo,
                    &c->whole_table, &c->items_slots, &c->part_fine, &c->group_table[0], &c->group_table[1], &c->group_table[2], &c->group_table[3],
                    &c->group_slots[0], &c->group_slots[1], &c->group_slots[2], &c->group_slots[3], &c->group_seen[0], &c->group_seen[1],
                    &c->group_seen[2], &c->group_seen[3], &c->ascii_dev, &c->ascii_off_dev, &c->ascii_bad_dev, &c->counters_alt, &c->overflow_alt, &c->smh_memo})
        b->release();
    c->pinned.release();
    c->pinned_small.release();
    for (auto& ev : c->phase_ev)
        if (ev) cudaEventDestroy(ev);
    for (int i = 0; i < 2; ++i) {
        c->pipe.packed[i].release();
        c->pipe.meta[i].release();
        c->pipe.sig[i].release();
        c->pipe.stage[i].release();
        c->pipe.meta_host[i].release();
        c->pipe.order[i].release();
        c->pipe.cursor[i].release();
        for (cudaEvent_t ev : {c->pipe.in_begin[i], c->pipe.in_done[i], c->pipe.compute_done[i], c->pipe.out_begin[i],
                               c->pipe.out_done[i]})
            if (ev) cudaEventDestroy(ev);
    }
    if (c->pipe.copy_in) cudaStreamDestroy(c->pipe.copy_in);
    if (c->pipe.copy_out) cudaStreamDestroy(c->pipe.copy_out);
    for (auto& ev : c->ev)
        if (ev) cudaEventDestroy(ev);
    for (auto& ev : c->lev) cudaEventDestroy(ev);
    if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
    if (c->redo_stream) cudaStreamDestroy(c->redo_stream);
    if (c->fork_ev) cudaEventDestroy(c->fork_ev);
    if (c->join_ev) cudaEventDestroy(c->join_ev);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

uint64_t kmu_launch_count(const kmu_ctx* ctx) { return ctx ? ctx->launches : 0; }
void* kmu_ctx_stream(kmu_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int32_t kmu_ctx_sync(kmu_ctx* ctx) {
    if (!ctx) return fail(KMU_EINVAL, "null context");
    ScopedDevice sd(ctx->device);
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return KMU_OK;
}

int32_t kmu_ctx_set_profiling(kmu_ctx* ctx, int32_t on) {
    if (!ctx) return fail(KMU_EINVAL, "null context");
    ctx->profiling = on != 0;
    return KMU_OK;
}

uint32_t kmu_last_launch_profile(kmu_ctx* ctx, kmu_launch_rec* out, uint32_t cap) {
    if (!ctx) return 0;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScopedDevice sd(ctx->device);
    uint32_t n = (uint32_t)ctx->lrec.size();
    std::vector<unsigned long long> ph(8 * 128, 0);
    if (ctx->counters.p)
        cudaMemcpy(ph.data(), (unsigned long long*)ctx->counters.p + 2 * kmu::LEN_BUCKETS + 256, sizeof(unsigned long long) * 8 * 128,
                   cudaMemcpyDeviceToHost);
    for (uint32_t i = 0; i < n; ++i) {
        if (2 * i + 1 < ctx->lev.size()) cudaEventElapsedTime(&ctx->lrec[i].ms, ctx->lev[2 * i], ctx->lev[2 * i + 1]);
        for (int j = 0; j < 8; ++j) ctx->lrec[i].phase_clocks[j] = ph[8 * (ctx->lrec[i].counter_idx & 127) + j];
    }
    if (out)
        for (uint32_t i = 0; i < n && i < cap; ++i) out[i] = ctx->lrec[i];
    return n;
}

int32_t kmu_last_times(const kmu_ctx* ctx, kmu_times* out) {
    if (!ctx || !out) return fail(KMU_EINVAL, "null argument");
    *out = ctx->last;
    return KMU_OK;
}

// ---- batches -------------------------------------------------------------------------
void kmu_seqbatch_destroy(kmu_seqbatch* b) {
    if (!b) return;
    ScopedDevice sd(b->device);
    b->order_cache.order.release();
    b->order_cache.cursor_dev.release();
    b->koff_cache.dev.release();
    if (!b->owns && b->owns_byte_off && b->byte_off) cudaFree(b->byte_off);
    if (b->owns) {
        if (b->packed) cudaFree(b->packed);
        if (b->byte_off) cudaFree(b->byte_off);
        if (b->nbases) cudaFree(b->nbases);
    }
    delete b;
}
uint64_t kmu_seqbatch_nseq(const kmu_seqbatch* b) { return b ? b->nseq : 0; }
uint64_t kmu_seqbatch_total_bases(const kmu_seqbatch* b) { return b ? b->total_bases : 0; }
uint64_t kmu_seqbatch_packed_bytes(const kmu_seqbatch* b) { return b ? b->packed_bytes : 0; }

static int32_t batch_from_host(kmu_ctx* ctx, const uint8_t* const* ptrs, const uint8_t* base, uint64_t base_bytes,
                               const uint64_t* src_off, const uint64_t* nbases, uint64_t nseq, kmu_seqbatch** out) {
    if (!ctx || !out || (nseq && !nbases)) return fail(KMU_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScopedDevice sd(ctx->device);
    kmu_seqbatch* b = nullptr;
    int32_t rc = batch_alloc(ctx, nbases, nseq, &b);
    if (rc) return rc;
    ctx->last = kmu_times{};
    cudaEventRecord(ctx->ev[2], ctx->stream);
    // already in the batch layout? then one straight copy from the caller's buffer
    bool same_layout = base != nullptr;
    if (same_layout) {
        for (uint64_t i = 0; i < nseq; ++i)
            if (src_off[i] != b->h_byte_off[i]) {
                same_layout = false;
                break;
            }
        if (same_layout && base_bytes < b->packed_bytes) same_layout = false;
    }
    cudaError_t e = cudaSuccess;
    if (b->packed_bytes) {
        if (same_layout) {
            e = cudaMemcpyAsync(b->packed, base, b->packed_bytes, cudaMemcpyHostToDevice, ctx->stream);
        } else {
            e = ctx->pinned.reserve(b->packed_bytes);
            if (e == cudaSuccess) {
                parallel_copy((uint8_t*)ctx->pinned.p, b->h_byte_off, ptrs, base, src_off, nbases, nseq);
                e = cudaMemcpyAsync(b->packed, ctx->pinned.p, b->packed_bytes, cudaMemcpyHostToDevice, ctx->stream);
            }
        }
    }
    if (e == cudaSuccess) e = cudaMemsetAsync(b->packed + b->packed_bytes, 0, TAIL_SLACK, ctx->stream);
    if (e == cudaSuccess) rc = batch_upload_meta(ctx, b);
    cudaEventRecord(ctx->ev[3], ctx->stream);
    if (e == cudaSuccess && rc == KMU_OK) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess || rc) {
        kmu_seqbatch_destroy(b);
        return rc ? rc : fail(KMU_ECUDA, "batch upload failed: %s", cudaGetErrorString(e));
    }
    cudaEventElapsedTime(&ctx->last.h2d_ms, ctx->ev[2], ctx->ev[3]);
    ctx->last.h2d_bytes = b->packed_bytes + 2 * sizeof(uint64_t) * nseq;
    *out = b;
    return KMU_OK;
}

int32_t kmu_seqbatch_from_ptrs(kmu_ctx* ctx, const uint8_t* const* seq_ptrs, const uint64_t* nbases, uint64_t nseq,
                               kmu_seqbatch** batch) {
    if (nseq && !seq_ptrs) return fail(KMU_EINVAL, "null sequence pointer array");
    return batch_from_host(ctx, seq_ptrs, nullptr, 0, nullptr, nbases, nseq, batch);
}

int32_t kmu_seqbatch_from_packed(kmu_ctx* ctx, const uint8_t* packed, uint64_t packed_bytes, const uint64_t* byte_off,
                                 const uint64_t* nbases, uint64_t nseq, kmu_seqbatch** batch) {
    if (nseq && (!packed || !byte_off)) return fail(KMU_EINVAL, "null packed buffer / offsets");
    for (uint64_t i = 0; i < nseq; ++i)
        if (byte_off[i] + (nbases[i] + 3) / 4 > packed_bytes)
            return fail(KMU_EINVAL, "sequence %llu overruns the packed buffer", (unsigned long long)i);
    return batch_from_host(ctx, nullptr, packed, packed_bytes, byte_off, nbases, nseq, batch);
}

int32_t kmu_seqbatch_synth(kmu_ctx* ctx, uint64_t seed, const uint64_t* nbases, uint64_t nseq, kmu_seqbatch** out) {
    if (!ctx || !out || (nseq && !nbases)) return fail(KMU_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScopedDevice sd(ctx->device);
    kmu_seqbatch* b = nullptr;
    int32_t rc = batch_alloc(ctx, nbases, nseq, &b);
    if (rc) return rc;
    rc = batch_upload_meta(ctx, b);
    // first_base = exclusive prefix sum of nbases (one stream for the whole batch)
    std::vector<uint64_t> first(nseq);
    uint64_t acc = 0;
    for (uint64_t i = 0; i < nseq; ++i) {
        first[i] = acc;
        acc += nbases[i];
    }
    cudaError_t e = ctx->misc.reserve(sizeof(uint64_t) * (nseq + 1));
    if (e == cudaSuccess && nseq)
        e = cudaMemcpyAsync(ctx->misc.p, first.data(), sizeof(uint64_t) * nseq, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess)
        e = kmu::launch_synth_packed(b->packed, b->byte_off, b->nbases, (const uint64_t*)ctx->misc.p, nseq,
                                     (b->packed_bytes + TAIL_SLACK) / 4, seed, ctx->stream);
    ctx->launches += 1;
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess || rc) {
        kmu_seqbatch_destroy(b);
        return rc ? rc : fail(KMU_ECUDA, "synthetic batch generation failed: %s", cudaGetErrorString(e));
    }
    *out = b;
    return KMU_OK;
}

int32_t kmu_seqbatch_from_ascii(kmu_ctx* ctx, const uint8_t* ascii, const uint64_t* ascii_off, uint64_t nseq,
                                int32_t drop_invalid, uint64_t* invalid_counts, kmu_seqbatch** out) {
    if (!ctx || !out || (nseq && (!ascii || !ascii_off))) return fail(KMU_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScopedDevice sd(ctx->device);
    ctx->last = kmu_times{};
    const uint64_t total_ascii = nseq ? ascii_off[nseq] : 0;
    // staging buffers owned by the context (grow-only): a feeder calls this once per pack, and cudaMalloc / cudaFree per
    // call cost more than the upload itself
    DevBuf& d_ascii = ctx->ascii_dev;
    DevBuf& d_off = ctx->ascii_off_dev;
    DevBuf& d_bad = ctx->ascii_bad_dev;
    auto cleanup = [&]() {};
    cudaError_t e = d_ascii.reserve(total_ascii + 16);
    if (e == cudaSuccess) e = d_off.reserve(sizeof(uint64_t) * (nseq + 1));
    if (e == cudaSuccess) e = d_bad.reserve(sizeof(uint64_t) * (nseq + 1));
    if (e != cudaSuccess) {
        cleanup();
        return fail(KMU_ENOMEM, "device allocation failed: %s", cudaGetErrorString(e));
    }
    cudaEventRecord(ctx->ev[2], ctx->stream);
    if (total_ascii) e = cudaMemcpyAsync(d_ascii.p, ascii, total_ascii, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(d_off.p, ascii_off, sizeof(uint64_t) * (nseq + 1), cudaMemcpyHostToDevice, ctx->stream);
    cudaEventRecord(ctx->ev[3], ctx->stream);
    if (e == cudaSuccess)
        e = kmu::launch_count_invalid((const uint8_t*)d_ascii.p, (const uint64_t*)d_off.p, nseq, (uint64_t*)d_bad.p,
                                      ctx->stream);
    std::vector<uint64_t> bad(nseq), nb(nseq);
    if (e == cudaSuccess && nseq)
        e = cudaMemcpyAsync(bad.data(), d_bad.p, sizeof(uint64_t) * nseq, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        cleanup();
        return fail(KMU_ECUDA, "ASCII upload / validation failed: %s", cudaGetErrorString(e));
    }
    ctx->launches += 1;
    uint64_t nbad_total = 0;
    for (uint64_t i = 0; i < nseq; ++i) {
        uint64_t len = ascii_off[i + 1] - ascii_off[i];
        nbad_total += bad[i];
        nb[i] = drop_invalid ? len - bad[i] : len;
    }
    if (invalid_counts) std::copy(bad.begin(), bad.end(), invalid_counts);
    if (!drop_invalid && nbad_total) {
        cleanup();
        // Alphabet2b::encode panics on the first such character (alphabet.rs:125)
        return fail(KMU_EINVAL, "pattern not a code in alphabet_2b: %llu non-ACGT characters",
                    (unsigned long long)nbad_total);
    }
    kmu_seqbatch* b = nullptr;
    int32_t rc = batch_alloc(ctx, nb.data(), nseq, &b);
    if (rc) {
        cleanup();
        return rc;
    }
    rc = batch_upload_meta(ctx, b);
    e = cudaMemsetAsync(b->packed, 0, b->packed_bytes + TAIL_SLACK, ctx->stream);
    if (e == cudaSuccess)
        e = kmu::launch_pack_ascii((const uint8_t*)d_ascii.p, (const uint64_t*)d_off.p, b->byte_off, b->nbases, nseq,
                                   drop_invalid, b->packed, ctx->stream);
    ctx->launches += 1;
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cleanup();
    if (e != cudaSuccess || rc) {
        kmu_seqbatch_destroy(b);
        return rc ? rc : fail(KMU_ECUDA, "ASCII packing failed: %s", cudaGetErrorString(e));
    }
    cudaEventElapsedTime(&ctx->last.h2d_ms, ctx->ev[2], ctx->ev[3]);
    ctx->last.h2d_bytes = total_ascii + sizeof(uint64_t) * (nseq + 1);
    ctx->last.launches = 2;
    *out = b;
    return KMU_OK;
}

// SequenceAA (src/aautils/kmeraa.rs:404-456): residues are validated and encoded to their 5-bit
// codes once, on the GPU.  drop_invalid != 0 is SequenceAA::new_filtered (:447-456); otherwise a
// residue outside the 20-letter upper-case alphabet fails the call (the reference panics in
// Alphabet::encode as soon as a k-mer reaches it, :106).
int32_t kmu_seqbatch_from_aa(kmu_ctx* ctx, const uint8_t* ascii, const uint64_t* ascii_off, uint64_t nseq,
                             int32_t drop_invalid, uint64_t* invalid_counts, kmu_seqbatch** out) {
    if (!ctx || !out || (nseq && (!ascii || !ascii_off))) return fail(KMU_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScopedDevice sd(ctx->device);
    ctx->last = kmu_times{};
    const uint64_t total_ascii = nseq ? ascii_off[nseq] : 0;
    DevBuf d_ascii, d_off, d_bad;
    auto cleanup = [&]() {
        d_ascii.release();
        d_off.release();
        d_bad.release();
    };
    cudaError_t e = d_ascii.reserve(total_ascii + 16);
    if (e == cudaSuccess) e = d_off.reserve(sizeof(uint64_t) * (nseq + 1));
    if (e == cudaSuccess) e = d_bad.reserve(sizeof(uint64_t) * (nseq + 1));
    if (e != cudaSuccess) {
        cleanup();
        return fail(KMU_ENOMEM, "device allocation failed: %s", cudaGetErrorString(e));
    }
    if (total_ascii) e = cudaMemcpyAsync(d_ascii.p, ascii, total_ascii, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(d_off.p, ascii_off, sizeof(uint64_t) * (nseq + 1), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess)
        e = kmu::launch_aa_count_invalid((const uint8_t*)d_ascii.p, (const uint64_t*)d_off.p, nseq, (uint64_t*)d_bad.p,
                                         ctx->stream);
    std::vector<uint64_t> bad(nseq), nb(nseq);
    if (e == cudaSuccess && nseq)
        e = cudaMemcpyAsync(bad.data(), d_bad.p, sizeof(uint64_t) * nseq, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        cleanup();
        return fail(KMU_ECUDA, "amino-acid upload / validation failed: %s", cudaGetErrorString(e));
    }
    ctx->launches += 1;
    uint64_t nbad_total = 0;
    for (uint64_t i = 0; i < nseq; ++i) {
        nbad_total += bad[i];
        nb[i] = ascii_off[i + 1] - ascii_off[i] - bad[i];
    }
    if (invalid_counts) std::copy(bad.begin(), bad.end(), invalid_counts);
    if (!drop_invalid && nbad_total) {
        cleanup();
        return fail(KMU_EINVAL, "encode: not a code in alphabet for amino acid: %llu invalid residues",
                    (unsigned long long)nbad_total);
    }
    kmu_seqbatch* b = nullptr;
    int32_t rc = batch_alloc(ctx, nb.data(), nseq, &b, 1);
    if (rc) {
        cleanup();
        return rc;
    }
    rc = batch_upload_meta(ctx, b);
    e = cudaMemsetAsync(b->packed, 0, b->packed_bytes + TAIL_SLACK, ctx->stream);
    if (e == cudaSuccess)
        e = kmu::launch_aa_encode((const uint8_t*)d_ascii.p, (const uint64_t*)d_off.p, b->byte_off, nseq, b->packed,
                                  ctx->stream);
    ctx->launches += 1;
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cleanup();
    if (e != cudaSuccess || rc) {
        kmu_seqbatch_destroy(b);
        return rc ? rc : fail(KMU_ECUDA, "amino-acid encoding failed: %s", cudaGetErrorString(e));
    }
    ctx->last.h2d_bytes = total_ascii + sizeof(uint64_t) * (nseq + 1);
    ctx->last.launches = 2;
    *out = b;
    return KMU_OK;
}

int32_t kmu_seqbatch_synth_aa(kmu_ctx* ctx, uint64_t seed, const uint64_t* nres, uint64_t nseq, kmu_seqbatch** out) {
    if (!ctx || !out || (nseq && !nres)) return fail(KMU_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScopedDevice sd(ctx->device);
    kmu_seqbatch* b = nullptr;
    int32_t rc = batch_alloc(ctx, nres, nseq, &b, 1);
    if (rc) return rc;
    rc = batch_upload_meta(ctx, b);
    std::vector<uint64_t> first(nseq);
    uint64_t acc = 0;
    for (uint64_t i = 0; i < nseq; ++i) {
        first[i] = acc;
        acc += nres[i];
    }
    cudaError_t e = ctx->misc.reserve(sizeof(uint64_t) * (nseq + 1));
    if (e == cudaSuccess && nseq)
        e = cudaMemcpyAsync(ctx->misc.p, first.data(), sizeof(uint64_t) * nseq, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess)
        e = kmu::launch_synth_aa(b->packed, b->byte_off, b->nbases, (const uint64_t*)ctx->misc.p, nseq,
                                 b->packed_bytes + TAIL_SLACK, seed, ctx->stream);
    ctx->launches += 1;
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess || rc) {
        kmu_seqbatch_destroy(b);
        return rc ? rc : fail(KMU_ECUDA, "synthetic protein generation failed: %s", cudaGetErrorString(e));
    }
    *out = b;
    return KMU_OK;
}

int32_t kmu_seqbatch_sample_reads(kmu_ctx* ctx, const kmu_seqbatch* genome, uint64_t seed, uint64_t first_read,
                                  uint64_t nreads, uint32_t read_len, uint32_t err_ppm, kmu_seqbatch** out) {
    if (!ctx || !genome || !out) return fail(KMU_EINVAL, "null argument");
    if (genome->alphabet != 0 || genome->nseq != 1) return fail(KMU_EINVAL, "the genome must be a batch of one DNA sequence");
    const uint64_t glen = genome->h_nbases[0];
    if (read_len < 1 || read_len > glen) return fail(KMU_EINVAL, "read length %u does not fit the genome", read_len);
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScopedDevice sd(ctx->device);
    std::vector<uint64_t> nb(nreads, read_len);
    kmu_seqbatch* b = nullptr;
    int32_t rc = batch_alloc(ctx, nb.data(), nreads, &b);
    if (rc) return rc;
    rc = batch_upload_meta(ctx, b);
    const uint32_t words_per_read = (uint32_t)(align_up((read_len + 3) / 4, SEQ_ALIGN) / 4);
    cudaError_t e = kmu::launch_sample_reads(genome->packed, glen, seed, first_read, nreads, read_len, err_ppm,
                                             words_per_read, b->packed, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(b->packed + b->packed_bytes, 0, TAIL_SLACK, ctx->stream);
    ctx->launches += 1;
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess || rc) {
        kmu_seqbatch_destroy(b);
        return rc ? rc : fail(KMU_ECUDA, "read sampling failed: %s", cudaGetErrorString(e));
    }
    *out = b;
    return KMU_OK;
}

int32_t kmu_seqbatch_slices(kmu_ctx* ctx, const kmu_seqbatch* src, const uint64_t* seq_idx, const uint64_t* begin,
                            const uint64_t* end, uint64_t nslices, kmu_seqbatch** out) {
    if (!ctx || !src || !out || (nslices && (!seq_idx || !begin || !end))) return fail(KMU_EINVAL, "null argument");
    std::vector<uint64_t> len(nslices), beg(nslices);
    for (uint64_t i = 0; i < nslices; ++i) {
        if (seq_idx[i] >= src->nseq) return fail(KMU_EINVAL, "slice %llu names sequence %llu of %llu", (unsigned long long)i,
                                                 (unsigned long long)seq_idx[i], (unsigned long long)src->nseq);
        const uint64_t L = src->h_nbases[seq_idx[i]];
        const uint64_t e = std::min(end[i], L), b0 = std::min(begin[i], e);
        beg[i] = b0;
        len[i] = e - b0;
    }
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScopedDevice sd(ctx->device);
    kmu_seqbatch* b = nullptr;
    int32_t rc = batch_alloc(ctx, len.data(), nslices, &b, src->alphabet);
    if (rc) return rc;
    rc = batch_upload_meta(ctx, b);
    cudaError_t e = ctx->misc.reserve(2 * sizeof(uint64_t) * (nslices + 1));
    uint64_t* d_idx = (uint64_t*)ctx->misc.p;
    uint64_t* d_beg = d_idx + (nslices + 1);
    if (e == cudaSuccess && nslices) {
        e = cudaMemcpyAsync(d_idx, seq_idx, sizeof(uint64_t) * nslices, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_beg, beg.data(), sizeof(uint64_t) * nslices, cudaMemcpyHostToDevice, ctx->stream);
    }
    if (e == cudaSuccess)
        e = kmu::launch_slice_copy(src->packed, src->byte_off, d_idx, d_beg, b->byte_off, b->nbases, nslices, b->packed_bytes / 4,
                                   src->alphabet, b->packed, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(b->packed + b->packed_bytes, 0, TAIL_SLACK, ctx->stream);
    ctx->launches += 1;
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);  // beg / len are host temporaries
    if (e != cudaSuccess || rc) {
        kmu_seqbatch_destroy(b);
        return rc ? rc : fail(KMU_ECUDA, "slicing failed: %s", cudaGetErrorString(e));
    }
    *out = b;
    return KMU_OK;
}

// a view of `nseq` consecutive sequences of a batch (one genome = its contigs): shares the device buffers of `src`,
// which must outlive the view
int32_t kmu_seqbatch_view(const kmu_seqbatch* src, uint64_t first_seq, uint64_t nseq, kmu_seqbatch** out) {
    if (!src || !out) return fail(KMU_EINVAL, "null argument");
    if (first_seq > src->nseq || nseq > src->nseq - first_seq)
        return fail(KMU_EINVAL, "view %llu..%llu exceeds the %llu sequences of the batch", (unsigned long long)first_seq,
                    (unsigned long long)(first_seq + nseq), (unsigned long long)src->nseq);
    auto* v = new kmu_seqbatch();
    v->device = src->device;
    v->owns = false;
    v->alphabet = src->alphabet;
    v->nseq = nseq;
    // the view starts at its first sequence: byte offsets are rebased into an array of its own
    const uint64_t base = nseq ? src->h_byte_off[first_seq] : 0;
    v->packed = src->packed + base;
    v->nbases = src->nbases + first_seq;
    v->h_nbases.assign(src->h_nbases.begin() + first_seq, src->h_nbases.begin() + first_seq + nseq);
    v->h_byte_off.resize(nseq);
    for (uint64_t i = 0; i < nseq; ++i) v->h_byte_off[i] = src->h_byte_off[first_seq + i] - base;
    for (uint64_t L : v->h_nbases) v->total_bases += L;
    if (nseq) {
        const uint64_t L = v->h_nbases.back();
        v->packed_bytes = v->h_byte_off.back() + align_up(src->alphabet ? L : (L + 3) / 4, SEQ_ALIGN);
    }
    ScopedDevice sd(src->device);
    cudaError_t e = cudaMalloc((void**)&v->byte_off, sizeof(uint64_t) * (nseq + 1));
    if (e == cudaSuccess && nseq) e = cudaMemcpy(v->byte_off, v->h_byte_off.data(), sizeof(uint64_t) * nseq, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        if (v->byte_off) cudaFree(v->byte_off);
        delete v;
        return fail(KMU_ECUDA, "view creation failed: %s", cudaGetErrorString(e));
    }
    v->owns_byte_off = true;
    *out = v;
    return KMU_OK;
}

int32_t kmu_seqbatch_alphabet(const kmu_seqbatch* b) { return b ? b->alphabet : -1; }

int32_t kmu_seqbatch_download(kmu_ctx* ctx, const kmu_seqbatch* b, uint8_t* packed_out, uint64_t* byte_off_out,
                              uint64_t* nbases_out) {
    if (!ctx || !b) return fail(KMU_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScopedDevice sd(ctx->device);
    if (packed_out && b->packed_bytes)
        CUDA_TRY(cudaMemcpyAsync(packed_out, b->packed, b->packed_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (byte_off_out) std::copy(b->h_byte_off.begin(), b->h_byte_off.end(), byte_off_out);
    if (nbases_out) std::copy(b->h_nbases.begin(), b->h_nbases.end(), nbases_out);
    return KMU_OK;
}

// ---- k-mer generation / ntHash ----------------------------------------------------------
uint64_t kmu_kmer_count(const kmu_seqbatch* b, uint32_t k) {
    if (!b) return 0;
    return b->kmer_count(k);
}

// -> *d_off: device array of nseq + 1 output offsets for this k (kept with the batch: KmerOffCache)
static int32_t upload_kmer_offsets(kmu_ctx* ctx, const kmu_seqbatch* b, uint32_t k, uint64_t* out_off_host,
                                   uint64_t* total_out, const uint64_t** d_off) {
    KmerOffCache& c = b->koff_cache;
    if (c.k != k || c.host.size() != b->nseq + 1 || !c.dev.p) {
        c.k = 0;
        c.host.resize(b->nseq + 1);
        uint64_t acc = 0;
        for (uint64_t i = 0; i < b->nseq; ++i) {
            c.host[i] = acc;
            const uint64_t L = b->h_nbases[i];
            acc += L >= k ? L - k + 1 : 0;
        }
        c.host[b->nseq] = acc;
        c.total = acc;
        CUDA_TRY(c.dev.reserve(sizeof(uint64_t) * (b->nseq + 1)));
        CUDA_TRY(cudaMemcpyAsync(c.dev.p, c.host.data(), sizeof(uint64_t) * (b->nseq + 1), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // pageable source
        c.k = k;
    }
    *total_out = c.total;
    *d_off = (const uint64_t*)c.dev.p;
    if (out_off_host) std::copy(c.host.begin(), c.host.end(), out_off_host);
    return KMU_OK;
}

int32_t kmu_generate_kmers(kmu_ctx* ctx, const kmu_seqbatch* b, uint32_t k, int32_t kmer_type, int32_t hash_kind,
                           void* out, uint64_t* out_off, int32_t out_on_device) {
    if (!ctx || !b) return fail(KMU_EINVAL, "null argument");
    if (int32_t a = kmu_check_kmer_args(b, k, kmer_type, hash_kind)) return a;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScopedDevice sd(ctx->device);
    ctx->last = kmu_times{};
    uint64_t total = 0;
    const uint64_t* d_koff = nullptr;
    int32_t rc = upload_kmer_offsets(ctx, b, k, out_off, &total, &d_koff);
    if (rc) return rc;
    if (total == 0) return KMU_OK;
    if (!out) return fail(KMU_EINVAL, "null output buffer");
    const size_t esz = kmer_type_is_u64(kmer_type) ? 8 : 4;
    void* dout = out;
    if (!out_on_device) {
        CUDA_TRY(ctx->sig_dev.reserve(total * esz));
        dout = ctx->sig_dev.p;
    }
    kmu::SeqView v{b->packed, b->byte_off, b->nbases, b->nseq};
    cudaEventRecord(ctx->ev[0], ctx->stream);
    CUDA_TRY(kmu::launch_generate_kmers(v, b->packed_bytes, k, kmer_type, hash_kind, d_koff, dout, ctx->stream));
    cudaEventRecord(ctx->ev[1], ctx->stream);
    ctx->launches += 1;
    ctx->last.launches = 1;
    if (!out_on_device) {
        cudaEventRecord(ctx->ev[4], ctx->stream);
        CUDA_TRY(cudaMemcpyAsync(out, dout, total * esz, cudaMemcpyDeviceToHost, ctx->stream));
        cudaEventRecord(ctx->ev[5], ctx->stream);
        ctx->last.d2h_bytes = total * esz;
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->last.kernel_ms, ctx->ev[0], ctx->ev[1]);
    if (!out_on_device) cudaEventElapsedTime(&ctx->last.d2h_ms, ctx->ev[4], ctx->ev[5]);
    return KMU_OK;
}

int32_t kmu_nthash_canonical(kmu_ctx* ctx, const kmu_seqbatch* b, uint32_t k, uint32_t n_multi, uint64_t* out_hash,
                             uint8_t* out_strand, int32_t out_on_device) {
    if (!ctx || !b) return fail(KMU_EINVAL, "null argument");
    if (k < 1 || k > 32) return fail(KMU_EINVAL, "ntHash is defined here for 1 <= k <= 32, got %u", k);
    if (b->alphabet != 0) return fail(KMU_EINVAL, "ntHash is defined on DNA sequences");
    if (n_multi < 1) return fail(KMU_EINVAL, "n_multi must be >= 1");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScopedDevice sd(ctx->device);
    ctx->last = kmu_times{};
    uint64_t total = 0;
    const uint64_t* d_koff = nullptr;
    int32_t rc = upload_kmer_offsets(ctx, b, k, nullptr, &total, &d_koff);
    if (rc) return rc;
    if (total == 0) return KMU_OK;
    if (!out_hash) return fail(KMU_EINVAL, "null output buffer");
    uint64_t* dh = out_hash;
    uint8_t* ds = out_strand;
    const size_t hbytes = total * n_multi * sizeof(uint64_t);
    if (!out_on_device) {
        CUDA_TRY(ctx->sig_dev.reserve(hbytes + total + 16));
        dh = (uint64_t*)ctx->sig_dev.p;
        ds = out_strand ? (uint8_t*)ctx->sig_dev.p + hbytes : nullptr;
    }
    kmu::SeqView v{b->packed, b->byte_off, b->nbases, b->nseq};
    cudaEventRecord(ctx->ev[0], ctx->stream);
    CUDA_TRY(kmu::launch_nthash(v, b->packed_bytes, k, n_multi, d_koff, dh, ds, ctx->stream));
    cudaEventRecord(ctx->ev[1], ctx->stream);
    ctx->launches += 1;
    ctx->last.launches = 1;
    if (!out_on_device) {
        cudaEventRecord(ctx->ev[4], ctx->stream);
        CUDA_TRY(cudaMemcpyAsync(out_hash, dh, hbytes, cudaMemcpyDeviceToHost, ctx->stream));
        if (out_strand) CUDA_TRY(cudaMemcpyAsync(out_strand, ds, total, cudaMemcpyDeviceToHost, ctx->stream));
        cudaEventRecord(ctx->ev[5], ctx->stream);
        ctx->last.d2h_bytes = hbytes + (out_strand ? total : 0);
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->last.kernel_ms, ctx->ev[0], ctx->ev[1]);
    if (!out_on_device) cudaEventElapsedTime(&ctx->last.d2h_ms, ctx->ev[4], ctx->ev[5]);
    return KMU_OK;
}

// ---- ProbMinHash3a --------------------------------------------------------------------
namespace {

struct LaunchClass {
    uint64_t first, count;  // range of `order`
    uint64_t nk_max;
    int mode;               // 0 histogram, 1 table
    bool table_global;
};

// geometry of one launch for sequences of at most nk_max k-mers
struct Geometry {
    uint32_t team_warps, teams_per_cta, regionA_bytes, slots_smem_bytes, team_smem_bytes, stage_bytes;
    int block;
    size_t smem;
    uint64_t table_entries_global;  // per team, 0 if the table lives in shared memory
};

uint64_t pow2_at_least(uint64_t x) {
    uint64_t p = 1;
    while (p < x) p <<= 1;
    return p;
}

Geometry make_geometry(uint64_t nk_max, int mode, uint32_t k, uint32_t m, bool key64, bool force_global_table,
                       bool no_stage = false) {
    Geometry g{};
    const size_t entry = kmu::pmh3a_entry_bytes(key64);
    const size_t qitem = kmu::pmh3a_qitem_bytes(key64);
    // about 512 k-mers per warp and pass (16 per lane)
    uint32_t tw = 1;
    while (tw < 32 && (uint64_t)tw * 512 < nk_max) tw <<= 1;
    uint64_t regionA;
    g.table_entries_global = 0;
    if (mode == 0) {
        regionA = (1ull << (2 * k));  // u8 counters
        if (regionA < 16) regionA = 16;
    } else {
        uint64_t entries = pow2_at_least(std::max<uint64_t>(64, 2 * nk_max));
        regionA = entries * entry;
        if (force_global_table || regionA > 128 * 1024) {
            g.table_entries_global = entries;
            regionA = 16;
        }
    }
    const uint64_t slots = align_up((uint64_t)m * 20, 16);  // 16-byte records + u32 mirror of the high words
    // TMA staging buffers: two per team, each large enough for the longest sequence of the
    // launch up to 16 KB (64 k bases); longer sequences are read from global memory
    // (amino-acid sequences are one byte per residue and are read from global memory)
    const uint64_t stage = no_stage ? 0 : std::min<uint64_t>(16 * 1024, align_up(nk_max / 4 + k + 32, 32));
    g.stage_bytes = (uint32_t)stage;
    if (tw == 32) {
        // two 16-warp teams keep more sequences in flight than one 32-warp team, if both fit
        const uint64_t fixed16 = regionA + 16ull * 64 * qitem + 2 * stage + stage / 2 + kmu::PMH3A_TEAM_SHARED_BYTES;
        const uint64_t team16 = align_up(fixed16 + slots <= SMEM_BUDGET ? fixed16 + slots : fixed16, 16);
        if (2 * team16 <= SMEM_BUDGET) tw = 16;
    }
    for (;;) {
        const uint64_t fixed = regionA + (uint64_t)tw * 64 * qitem + 2 * stage + stage / 2 + kmu::PMH3A_TEAM_SHARED_BYTES;
        const bool slots_in_smem = fixed + slots <= SMEM_BUDGET;
        const uint64_t team_bytes = align_up(slots_in_smem ? fixed + slots : fixed, 16);
        const uint32_t teams_fit = std::max<uint32_t>(1, (uint32_t)(SMEM_BUDGET / team_bytes));
        uint32_t max_teams = 32 / tw;
        if (tw > 1 && max_teams > 15) max_teams = 15;  // named barriers 1..15
        // shared memory limits the number of teams: widen the teams so that the SM still gets 32 warps
        if (tw < 32 && teams_fit * tw < 24) {
            tw <<= 1;
            continue;
        }
        const uint32_t teams = std::min(max_teams, teams_fit);
        g.team_warps = tw;
        g.teams_per_cta = teams;
        g.regionA_bytes = (uint32_t)regionA;
        g.slots_smem_bytes = slots_in_smem ? (uint32_t)slots : 0;
        g.team_smem_bytes = (uint32_t)team_bytes;
        g.block = (int)(tw * 32 * teams);
        g.smem = (size_t)team_bytes * teams;
        return g;
    }
}

}  // namespace

}  // extern "C"

// processing order of a batch for the per-sequence sketch kernels: longest sequences first
// (8 buckets per octave of the k-mer count); built once per (batch, k) and kept with the batch
int32_t kmu_ensure_order(kmu_ctx* ctx, const kmu_seqbatch* b, uint32_t k, uint64_t* launches) {
    OrderCache& oc = b->order_cache;
    if (oc.k == k && !oc.hist.empty()) return KMU_OK;
    cudaStream_t st = ctx->stream;
    oc.hist.assign(kmu::LEN_BUCKETS, 0);
    oc.cursor.assign(kmu::LEN_BUCKETS, 0);
    oc.nk_longest = 0;
    for (uint64_t L : b->h_nbases) {
        const uint64_t nk = L >= k ? L - k + 1 : 0;
        ++oc.hist[kmu::len_bucket_host(nk)];
        oc.nk_longest = std::max(oc.nk_longest, nk);
    }
    unsigned long long acc = 0;
    for (int i = 0; i < kmu::LEN_BUCKETS; ++i) {
        oc.cursor[i] = acc;
        acc += oc.hist[i];
    }
    CUDA_TRY(oc.order.reserve(sizeof(uint32_t) * (b->nseq + 1)));
    CUDA_TRY(oc.cursor_dev.reserve(sizeof(unsigned long long) * kmu::LEN_BUCKETS));
    CUDA_TRY(cudaMemcpyAsync(oc.cursor_dev.p, oc.cursor.data(), sizeof(unsigned long long) * kmu::LEN_BUCKETS,
                             cudaMemcpyHostToDevice, st));
    CUDA_TRY(kmu::launch_len_scatter(b->nbases, b->nseq, k, (unsigned long long*)oc.cursor_dev.p, (uint32_t*)oc.order.p, st));
    if (launches) ++*launches;
    oc.k = k;
    return KMU_OK;
}

// octave classes of the cached order: sequences whose k-mer count lies in [2^oct, 2^(oct+1)), longest first
std::vector<OctaveClass> kmu_octave_classes(const kmu_seqbatch* b) {
    const OrderCache& oc = b->order_cache;
    std::vector<OctaveClass> out;
    for (int oct = 63; oct >= 0; --oct) {
        int b_hi = kmu::LEN_BUCKETS - 1 - (oct * 8 + 7), b_lo = kmu::LEN_BUCKETS - 1 - oct * 8;
        uint64_t cnt = 0;
        for (int bb = b_hi; bb <= b_lo; ++bb) cnt += oc.hist[bb];
        if (!cnt) continue;
        OctaveClass c;
        c.first = oc.cursor[b_hi];
        c.count = cnt;
        c.nk_max = oct == 63 ? ~0ULL : ((2ULL << oct) - 1);
        out.push_back(c);
    }
    if (!out.empty()) out.front().nk_max = std::min(out.front().nk_max, oc.nk_longest);
    return out;
}

extern "C" {

// phase 0: everything; phase 1: the main launches only (nothing waits on the device); phase 2: wait for
// them and redo the sequences they flagged.  The host pipeline runs 1 and 2 apart so that it can
// prepare the next chunk while the device works.
static int32_t sketch_pmh3a_locked(kmu_ctx* ctx, const kmu_seqbatch* b, uint32_t k, int32_t kmer_type,
                                   int32_t hash_kind, uint32_t m, void* d_sig, int phase = 0, int scratch_set = 0,
                                   bool redo_aside = false) {
    DevBuf& counters = scratch_set ? ctx->counters_alt : ctx->counters;
    DevBuf& overflow = scratch_set ? ctx->overflow_alt : ctx->overflow;
    const bool key64 = kmer_type_is_u64(kmer_type);
    const bool aa = kmer_type_is_aa(kmer_type);
    const size_t vsz = key64 ? 8 : 4;
    const uint64_t nseq = b->nseq;
    cudaStream_t st = ctx->stream;
    uint64_t launches = 0;
    if (phase != 2) ctx->lrec.clear();

    // ---- order sequences longest first (8 buckets per octave of the k-mer count) -------
    // histogram and cursors come from the host copy of the lengths; the processing order is
    // built on the device once per (batch, k) and kept with the batch
    CUDA_TRY(counters.reserve(sizeof(unsigned long long) * (2 * kmu::LEN_BUCKETS + 256 + 8 * 128)));
    unsigned long long* d_hist = (unsigned long long*)counters.p;
    unsigned long long* d_cursor = d_hist + kmu::LEN_BUCKETS;
    unsigned long long* d_work = d_cursor + kmu::LEN_BUCKETS;  // 128 work counters
    unsigned long long* d_ovf_count = d_work + 128;
    unsigned long long* d_phase = d_work + 256;  // 8 per launch, profiling only
    if (phase != 2)
        CUDA_TRY(cudaMemsetAsync(counters.p, 0, sizeof(unsigned long long) * (2 * kmu::LEN_BUCKETS + 256 + 8 * 128), st));
    {
        int32_t orc = kmu_ensure_order(ctx, b, k, &launches);
        if (orc) return orc;
    }
    OrderCache& oc = b->order_cache;
    (void)d_hist;
    const std::vector<unsigned long long>& hist = oc.hist;
    const std::vector<unsigned long long>& cursor = oc.cursor;
    const uint64_t nk_longest = oc.nk_longest;
    const uint32_t* d_order = (const uint32_t*)oc.order.p;

    // ---- launch classes: one per octave of the k-mer count ------------------------------
    const bool hist_ok = k <= 8 && !aa;
    const size_t entry = kmu::pmh3a_entry_bytes(key64);
    // sequences over a small key space go to the one-pass kernel (kmu_pmh3a_direct.cu).  Three forms, by length
    // class (key = 8 * octave of the k-mer count + its next three bits), longest first in the processing order:
    //   keys >= key_vlong          8-bit counters, one point per occurrence (counts above 15 are likely)
    //   [key_long, key_vlong)      4-bit counters, one point per occurrence
    //   [key_short, key_long)      4-bit counters, two points per occurrence
    int form_key[4] = {kmu::LEN_BUCKETS, 16 * 8 + 4, 11 * 8, 9 * 8 + 4};  // 98304, 2048, 768 k-mers
    int form_variant[3] = {0, 1, 2};
    if (const char* env = std::getenv("KMU_DIRECT_KEYS")) std::sscanf(env, "%d,%d,%d", &form_key[1], &form_key[2], &form_key[3]);
    if (const char* env = std::getenv("KMU_DIRECT_VARIANTS")) std::sscanf(env, "%d,%d,%d", &form_variant[0], &form_variant[1], &form_variant[2]);
    form_key[2] = std::min(form_key[2], form_key[1]);
    form_key[3] = std::min(form_key[3], form_key[2]);
    const bool direct_ok = !key64 && !aa && k <= 8 && kmu::pmh3a_direct_smem_bytes(k, m) <= SMEM_BUDGET &&
                           !std::getenv("KMU_NO_DIRECT");
    if (!direct_ok) form_key[1] = form_key[2] = form_key[3] = kmu::LEN_BUCKETS;
    const int key_short = form_key[3];
    uint64_t form_count[3] = {0, 0, 0}, form_nk_min[3];
    for (int f = 0; f < 3; ++f) {
        for (int key = form_key[f] - 1; key >= form_key[f + 1]; --key) form_count[f] += hist[kmu::LEN_BUCKETS - 1 - key];
        form_nk_min[f] = kmu::len_bucket_min_nk(kmu::LEN_BUCKETS - 1 - std::min(form_key[f + 1], kmu::LEN_BUCKETS - 1));
    }
    std::vector<LaunchClass> classes;
    for (int oct = std::min(63, (key_short - 1) / 8); oct >= 0 && key_short > 0; --oct) {
        // buckets of this octave: keys oct*8 .. oct*8+7 (below key_short)  -> bucket index LEN_BUCKETS-1-key
        const int top_key = std::min(oct * 8 + 7, key_short - 1);
        int b_hi = kmu::LEN_BUCKETS - 1 - top_key, b_lo = kmu::LEN_BUCKETS - 1 - oct * 8;
        uint64_t cnt = 0;
        for (int bb = b_hi; bb <= b_lo; ++bb) cnt += hist[bb];
        if (!cnt) continue;
        uint64_t nk_max = oct == 63 ? ~0ULL : ((2ULL << oct) - 1);
        uint64_t table_bytes = pow2_at_least(std::max<uint64_t>(64, 2 * nk_max)) * entry;
        LaunchClass c{};
        c.first = cursor[b_hi];
        c.count = cnt;
        c.nk_max = nk_max;
        const uint64_t hist_bytes = (1ull << (2 * k));  // u8 counters
        if (hist_ok && (table_bytes > 128 * 1024 || hist_bytes <= 4 * table_bytes)) {
            c.mode = 0;
            c.table_global = false;
        } else {
            c.mode = 1;
            c.table_global = table_bytes > 128 * 1024;
        }
        // merge with the previous (longer) class when the team geometry is the same and the counting
        // structure does not depend on the length (histogram, or table in global memory): one launch
        if (!classes.empty()) {
            LaunchClass& p = classes.back();
            // a handful of sequences is not worth a launch of its own (0.1 ms whatever its size): they join the
            // class above, whose histogram / table is large enough for anything shorter
            // (so does a class with less than ~0.1 ms of work in it: the short reads of one chunk of the host pipeline)
            if ((c.count <= 2 * (uint64_t)ctx->sm_count || c.count * c.nk_max <= 2000000) && p.first + p.count == c.first) {
                p.count += c.count;
                continue;
            }
            if (p.mode == c.mode && p.table_global == c.table_global && (c.mode == 0 || c.table_global)) {
                Geometry gp = make_geometry(p.nk_max, p.mode, k, m, key64, p.table_global, aa);
                Geometry gc = make_geometry(c.nk_max, c.mode, k, m, key64, c.table_global, aa);
                if (gp.team_warps == gc.team_warps && gp.teams_per_cta == gc.teams_per_cta) {
                    p.count += c.count;
                    continue;
                }
            }
        }
        classes.push_back(c);
    }
    // exact largest k-mer count (the first class is sized by it, not by its octave bound)
    if (!classes.empty()) classes.front().nk_max = std::min(classes.front().nk_max, nk_longest);

    // ---- parameters common to all launches ------------------------------------------------
    kmu::Pmh3aParams P{};
    P.packed = b->packed;
    P.byte_off = b->byte_off;
    P.nbases = b->nbases;
    P.order = d_order;
    P.k = k;
    P.kmer_type = kmer_type;
    P.hash_kind = hash_kind;
    P.m = m;
    P.slot_thresh = (uint32_t)((0x100000000ULL) % m);
    {
        // ProbMinHash3a::new : lambda = ln(m / (m-1)); ExpRestricted01::new
        double lambda = std::log((double)m / (double)(m - 1));
        P.e.lambda = lambda;
        P.e.c1 = std::expm1(lambda) / lambda;
        P.e.c2 = std::log(2.0 / (1.0 + std::exp(-lambda))) / lambda;
        P.e.c3 = (1.0 - std::exp(-lambda)) / lambda;
    }
    P.sig = d_sig;
    CUDA_TRY(overflow.reserve(sizeof(uint32_t) * (nseq + 1)));
    P.overflow_count = d_ovf_count;
    P.overflow_list = (uint32_t*)overflow.p;
    // first and second point of every possible pre-key when the key space is small (u32 key types, k <= 10):
    // built once per (k, type, hash, m) and kept in the context (2 x 16 B per key, L2 resident)
    P.memo_fast = nullptr;
    if (!key64 && !aa && 2 * k <= 20) {
        const uint32_t nkeys = 1u << (2 * k);
        if (!(ctx->memo.p && ctx->memo_k == k && ctx->memo_m == m && ctx->memo_type == kmer_type &&
              ctx->memo_hash == hash_kind)) {
            CUDA_TRY(ctx->memo.reserve((size_t)nkeys * 32));
            CUDA_TRY(kmu::launch_pmh3a_memo(P, ctx->memo.p, nkeys, st));
            ++launches;
            ctx->memo_k = k;
            ctx->memo_m = m;
            ctx->memo_type = kmer_type;
            ctx->memo_hash = hash_kind;
        }
        P.memo_fast = ctx->memo.p;
    }

    auto run_class = [&](const LaunchClass& c, const uint32_t* order, int counter_idx, bool speculate, cudaStream_t st) -> int32_t {
        Geometry g = make_geometry(c.nk_max, c.mode, k, m, key64, c.table_global, aa);
        uint64_t teams_needed = c.count;
        uint64_t ctas_needed = (teams_needed + g.teams_per_cta - 1) / g.teams_per_cta;
        // CTAs per SM that fit (threads and shared memory)
        uint32_t per_sm = (uint32_t)std::min<uint64_t>(std::min<uint64_t>(2048 / g.block, 65536 / (64 * (uint64_t)g.block)),
                                                       SMEM_BUDGET / std::max<size_t>(g.smem, 1));
        if (per_sm == 0) per_sm = 1;
        if (per_sm > 8) per_sm = 8;
        int grid = (int)std::min<uint64_t>(ctas_needed, (uint64_t)ctx->sm_count * per_sm);
        kmu::Pmh3aParams Q = P;
        Q.order = order;
        Q.first = c.first;
        Q.count = c.count;
        Q.work_counter = d_work + counter_idx;
        Q.speculate = speculate ? 1u : 0u;
        Q.spec_factor = (double)m * std::log((double)m / 1e-4);
        Q.phase_clocks = ctx->profiling ? d_phase + 8 * counter_idx : nullptr;
        Q.team_warps = g.team_warps;
        Q.team_smem_bytes = g.team_smem_bytes;
        Q.regionA_bytes = g.regionA_bytes;
        Q.slots_smem_bytes = g.slots_smem_bytes;
        Q.stage_bytes = g.stage_bytes;
        Q.slot_scratch = nullptr;
        Q.table_scratch = nullptr;
        Q.table_scratch_entries = 0;
        const uint64_t nteams_total = (uint64_t)grid * g.teams_per_cta;
        if (g.slots_smem_bytes == 0) {
            CUDA_TRY(ctx->slot_scratch.reserve(nteams_total * m * 20));
            Q.slot_scratch = (kmu::Slot*)ctx->slot_scratch.p;
        }
        if (g.table_entries_global) {
            // bound the scratch: shrink the grid rather than ask for more than 32 GiB
            uint64_t per_team = g.table_entries_global * entry;
            const uint64_t budget = 32ULL << 30;
            if (per_team * nteams_total > budget) {
                uint64_t fit_teams = std::max<uint64_t>(1, budget / per_team);
                grid = (int)std::max<uint64_t>(1, fit_teams / g.teams_per_cta);
            }
            uint64_t need = per_team * (uint64_t)grid * g.teams_per_cta;
            if (need > ctx->table_scratch.cap) ctx->table_scratch_clean = false;
            cudaError_t e = ctx->table_scratch.reserve(need);
            if (e != cudaSuccess)
                return fail(KMU_ENOMEM, "table scratch of %llu bytes: %s", (unsigned long long)need, cudaGetErrorString(e));
            if (!ctx->table_scratch_clean) {
                CUDA_TRY(cudaMemsetAsync(ctx->table_scratch.p, 0, ctx->table_scratch.cap, st));
                ctx->table_scratch_clean = true;  // kernels leave their tables clean
            }
            Q.table_scratch = (uint8_t*)ctx->table_scratch.p;
            Q.table_scratch_entries = g.table_entries_global;
        }
        size_t li = ctx->lrec.size();
        if (ctx->profiling) {
            while (ctx->lev.size() < 2 * (li + 1)) {
                cudaEvent_t ev;
                CUDA_TRY(cudaEventCreate(&ev));
                ctx->lev.push_back(ev);
            }
            cudaEventRecord(ctx->lev[2 * li], st);
        }
        CUDA_TRY(kmu::launch_pmh3a(Q, key64, c.mode, grid, g.block, g.smem, st));
        if (ctx->profiling) {
            cudaEventRecord(ctx->lev[2 * li + 1], st);
            kmu_launch_rec r{};
            r.mode = c.mode;
            r.table_global = c.table_global;
            r.team_warps = g.team_warps;
            r.teams_per_cta = g.teams_per_cta;
            r.grid = (uint32_t)grid;
            r.block = (uint32_t)g.block;
            r.smem_bytes = (uint32_t)g.smem;
            r.nseq = c.count;
            r.nk_max = c.nk_max;
            r.counter_idx = (uint32_t)counter_idx;
            ctx->lrec.push_back(r);
        }
        ++launches;
        return KMU_OK;
    };

    if (classes.size() > 120) return fail(KMU_EINVAL, "too many launch classes");  // work counters 124..126: one-pass launches, 127: redo launch
    int ci = 0;
    uint64_t form_first = 0;
    bool join_aux = false;
    const bool use_side = !ctx->profiling && phase != 2 && form_count[1] > 0;
    if (use_side && !ctx->aux_stream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&ctx->join_ev, cudaEventDisableTiming));
    }
    for (int form = 0; form < 3 && phase != 2; form_first += form_count[form], ++form) {
        const uint64_t count = form_count[form];
        if (!count) continue;
        kmu::Pmh3aParams Q = P;
        Q.order = d_order;
        Q.first = form_first;
        Q.count = count;
        Q.work_counter = d_work + 126 - form;
        Q.phase_clocks = ctx->profiling ? d_phase + 8 * (126 - form) : nullptr;
        const int variant = form_variant[form];
        Q.regionA_bytes = (uint32_t)kmu::pmh3a_direct_hist_bytes(k, variant);
        Q.slots_smem_bytes = (uint32_t)align_up((uint64_t)m * 20, 16);
        const int grid = (int)std::min<uint64_t>(count, (uint64_t)ctx->sm_count * kmu::pmh3a_direct_ctas_per_sm(variant));
        size_t li = ctx->lrec.size();
        if (ctx->profiling) {
            while (ctx->lev.size() < 2 * (li + 1)) {
                cudaEvent_t ev;
                CUDA_TRY(cudaEventCreate(&ev));
                ctx->lev.push_back(ev);
            }
            cudaEventRecord(ctx->lev[2 * li], st);
        }
        // the very long sequences are few (a handful of CTAs, each busy for a long time): their launch runs on a
        // side stream next to the other forms instead of in front of them (profiling runs time it alone).  Putting
        // the short form and the team-kernel classes there as well was measured: slower (the kernels compete).
        const bool aside = use_side && form == 0 && count < (uint64_t)grid + 1 && count < 2 * (uint64_t)ctx->sm_count;
        if (aside) {
            if (!join_aux) {
                CUDA_TRY(cudaEventRecord(ctx->fork_ev, st));
                CUDA_TRY(cudaStreamWaitEvent(ctx->aux_stream, ctx->fork_ev, 0));
                join_aux = true;
            }
            CUDA_TRY(kmu::launch_pmh3a_direct(Q, grid, variant, ctx->aux_stream));
        } else {
            CUDA_TRY(kmu::launch_pmh3a_direct(Q, grid, variant, st));
        }
        if (ctx->profiling) {
            cudaEventRecord(ctx->lev[2 * li + 1], st);
            kmu_launch_rec r{};
            r.mode = 2;  // one-pass kernel
            r.team_warps = (uint32_t)kmu::pmh3a_direct_threads(variant) / 32;
            r.teams_per_cta = 1;
            r.grid = (uint32_t)grid;
            r.block = (uint32_t)kmu::pmh3a_direct_threads(variant);
            r.smem_bytes = (uint32_t)kmu::pmh3a_direct_smem_bytes(k, m, variant);
            r.nseq = count;
            r.nk_max = form == 0 ? nk_longest : std::min<uint64_t>(nk_longest, form_nk_min[form - 1] - 1);
            r.counter_idx = 126 - form;
            for (uint64_t L : b->h_nbases) {
                const uint64_t nk = L >= k ? L - k + 1 : 0;
                if (nk >= form_nk_min[form] && (form == 0 || nk < form_nk_min[form - 1])) r.nbases += L;
            }
            ctx->lrec.push_back(r);
        }
        ++launches;
    }
    const size_t lrec_base = ctx->lrec.size();
    if (phase != 2)
        for (const LaunchClass& c : classes) {
            int32_t rc = run_class(c, d_order, ci++, true, st);
            if (rc) return rc;
        }
    if (join_aux) {
        CUDA_TRY(cudaEventRecord(ctx->join_ev, ctx->aux_stream));
        CUDA_TRY(cudaStreamWaitEvent(st, ctx->join_ev, 0));
    }
    if (ctx->profiling && phase != 2) {
        // bases per launch: class i covers the sequences whose length bucket start lies in its range
        for (uint64_t L : b->h_nbases) {
            uint64_t nk = L >= k ? L - k + 1 : 0;
            uint64_t pos = cursor[kmu::len_bucket_host(nk)];
            for (size_t i = 0; i < classes.size() && lrec_base + i < ctx->lrec.size(); ++i)
                if (pos >= classes[i].first && pos < classes[i].first + classes[i].count) {
                    ctx->lrec[lrec_base + i].nbases += L;
                    break;
                }
        }
    }
    // ---- sequences whose u8 histogram counters wrapped or whose speculative qmax bound failed:
    //      redo them with u32 table counters and without speculation ---------------------------
    if (phase == 1) {
        // the redo count travels to pinned host memory right behind the launches: phase 2 only waits for this
        // chunk's event, not for whatever was enqueued on the stream afterwards
        CUDA_TRY(ctx->pinned_small.reserve(64));
        unsigned long long* h_novf = (unsigned long long*)ctx->pinned_small.p + (scratch_set ? 1 : 0);
        CUDA_TRY(cudaMemcpyAsync(h_novf, d_ovf_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        if (!ctx->phase_ev[scratch_set]) CUDA_TRY(cudaEventCreateWithFlags(&ctx->phase_ev[scratch_set], cudaEventDisableTiming));
        CUDA_TRY(cudaEventRecord(ctx->phase_ev[scratch_set], st));
    }
    if (phase != 1) {
        unsigned long long novf = 0;
        if (phase == 2) {
            CUDA_TRY(cudaEventSynchronize(ctx->phase_ev[scratch_set]));
            novf = ((unsigned long long*)ctx->pinned_small.p)[scratch_set ? 1 : 0];
        } else {
            CUDA_TRY(cudaMemcpyAsync(&novf, d_ovf_count, sizeof(novf), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
        }
        if (novf) {
            LaunchClass c{};
            c.first = 0;
            c.count = novf;
            c.nk_max = nk_longest;
            c.mode = 1;
            c.table_global = true;
            // host pipeline: the redo launch of this chunk runs on a side stream while the next chunk's main launches
            // (already queued on the main stream) proceed, unless those share the global scratch (slots / tables) with it
            bool aside = redo_aside && phase == 2;
            for (const LaunchClass& mc : classes) {
                Geometry mg = make_geometry(mc.nk_max, mc.mode, k, m, key64, mc.table_global, aa);
                if (mg.slots_smem_bytes == 0 || mg.table_entries_global) aside = false;
            }
            cudaStream_t rs = st;
            if (aside) {
                if (!ctx->redo_stream) CUDA_TRY(cudaStreamCreateWithFlags(&ctx->redo_stream, cudaStreamNonBlocking));
                rs = ctx->redo_stream;  // the host has waited for this chunk's main launches: no event needed
                ctx->redo_on_side = true;
            } else {
                ctx->redo_on_main = true;
            }
            int32_t rc = run_class(c, (const uint32_t*)overflow.p, 127, false, rs);
            if (rc) return rc;
        }
    }
    ctx->launches += launches;
    ctx->last.launches += launches;
    (void)vsz;
    return KMU_OK;
}

int32_t kmu_sketch_pmh3a(kmu_ctx* ctx, const kmu_seqbatch* b, uint32_t k, int32_t kmer_type, int32_t hash_kind,
                         uint32_t m, void* sig, int32_t sig_on_device) {
    if (!ctx || !b) return fail(KMU_EINVAL, "null argument");
    if (int32_t a = kmu_check_kmer_args(b, k, kmer_type, hash_kind)) return a;
    if (m < 2) return fail(KMU_EINVAL, "ProbMinHash3a needs at least 2 hash values (m = %u)", m);
    if (b->nseq == 0) return KMU_OK;
    if (!sig) return fail(KMU_EINVAL, "null signature buffer");
    if (b->nseq >= 0xFFFFFFFFull) return fail(KMU_EINVAL, "more than 2^32-1 sequences in one batch");
    if (b->longest() >= (1ull << 30)) return fail(KMU_EINVAL, "a single sequence is limited to 2^30 bases here (kmu_sketch_pmh3a_whole has no limit)");
    // Genome-sized sequences over a large key space: the team kernel keeps one sequence on one SM (a 5 Mb genome is
    // ~70 ms of one SM), the whole-file procedure (counting table in HBM + item kernel, the whole GPU on one sequence,
    // ~0.4 ms per 5 Mb) gives the same signature.  Taken when every sequence is long enough for its fixed ~0.1 ms.
    if (!kmer_type_is_aa(kmer_type) && k > 8 && !std::getenv("KMU_PMH3A_TEAM_ONLY")) {  // (tests compare the two paths)
        (void)b->kmer_count(k);  // fills min_nbases
        const uint64_t min_nk = b->min_nbases >= k && b->min_nbases < ~0ull - 1 ? b->min_nbases - k + 1 : 0;
        if (min_nk >= 2000000 || (b->nseq * 2 <= (uint64_t)ctx->sm_count && min_nk >= 250000)) {
            std::vector<uint64_t> ones(b->nseq, 1);
            return kmu_sketch_pmh3a_groups(ctx, b, ones.data(), b->nseq, k, kmer_type, hash_kind, m, sig, sig_on_device);
        }
    }
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScopedDevice sd(ctx->device);
    ctx->last = kmu_times{};
    const size_t vsz = kmer_type_is_u64(kmer_type) ? 8 : 4;
    const size_t sig_bytes = (size_t)b->nseq * m * vsz;
    void* d_sig = sig;
    if (!sig_on_device) {
        CUDA_TRY(ctx->sig_dev.reserve(sig_bytes));
        d_sig = ctx->sig_dev.p;
    }
    cudaEventRecord(ctx->ev[0], ctx->stream);
    int32_t rc = sketch_pmh3a_locked(ctx, b, k, kmer_type, hash_kind, m, d_sig);
    cudaEventRecord(ctx->ev[1], ctx->stream);
    if (rc) {
        cudaStreamSynchronize(ctx->stream);
        return rc;
    }
    if (!sig_on_device) {
        cudaEventRecord(ctx->ev[4], ctx->stream);
        CUDA_TRY(cudaMemcpyAsync(sig, d_sig, sig_bytes, cudaMemcpyDeviceToHost, ctx->stream));
        cudaEventRecord(ctx->ev[5], ctx->stream);
        ctx->last.d2h_bytes = sig_bytes;
    }
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        ctx->table_scratch_clean = false;
        return fail(KMU_ECUDA, "ProbMinHash3a sketch kernels failed: %s", cudaGetErrorString(e));
    }
    cudaEventElapsedTime(&ctx->last.kernel_ms, ctx->ev[0], ctx->ev[1]);
    if (!sig_on_device) cudaEventElapsedTime(&ctx->last.d2h_ms, ctx->ev[4], ctx->ev[5]);
    return KMU_OK;
}

// One-shot host form.  The sequences are cut into chunks of about 96 MB of packed bases; chunk
// c+1 travels host->device on the copy-in stream while chunk c is sketched on the compute stream
// and the signatures of chunk c-1 travel device->host on the copy-out stream (two device slots,
// grow-only, owned by the context: no cudaMalloc / cudaFree per call).  A host buffer already in
// the batch layout (every sequence on a 16-byte boundary, back to back) is copied straight from
// the caller's memory -- pin it for full PCIe speed; any other layout is re-laid out through a
// pinned staging buffer first.
// ptrs != nullptr: the sequences are nseq separate host allocations (the `&[&Sequence]` of the Rust entry points); they are
// gathered chunk by chunk into the pinned staging buffer by up to 16 host threads while the previous chunk is sketched.
static int32_t sketch_pmh3a_host_impl(kmu_ctx* ctx, const uint8_t* packed, uint64_t packed_bytes, const uint64_t* byte_off,
                                      const uint8_t* const* ptrs, const uint64_t* nbases, uint64_t nseq, uint32_t k,
                                      int32_t kmer_type, int32_t hash_kind, uint32_t m, void* sig) {
    if (!ctx || (nseq && (!nbases || (!ptrs && (!packed || !byte_off))))) return fail(KMU_EINVAL, "null argument");
    if (kmer_type_is_aa(kmer_type)) return fail(KMU_EINVAL, "the one-shot host form takes 2-bit DNA sequences");
    if (int32_t a = kmu_check_kmer_args(nullptr, k, kmer_type, hash_kind)) return a;
    if (m < 2) return fail(KMU_EINVAL, "ProbMinHash3a needs at least 2 hash values (m = %u)", m);
    if (nseq == 0) return KMU_OK;
    if (!sig) return fail(KMU_EINVAL, "null signature buffer");
    if (nseq >= 0xFFFFFFFFull) return fail(KMU_EINVAL, "more than 2^32-1 sequences in one batch");
    // one pass: bounds, and whether the caller's buffer already is in the batch layout (then its offsets are used as they are)
    bool same_layout = ptrs == nullptr;
    uint64_t total_bytes = 0;
    for (uint64_t i = 0; i < nseq; ++i) {
        if (!ptrs && byte_off[i] + (nbases[i] + 3) / 4 > packed_bytes)
            return fail(KMU_EINVAL, "sequence %llu overruns the packed buffer", (unsigned long long)i);
        if (ptrs && nbases[i] && !ptrs[i]) return fail(KMU_EINVAL, "sequence %llu: null pointer", (unsigned long long)i);
        if (nbases[i] >= (1ull << 30)) return fail(KMU_EINVAL, "a single sequence is limited to 2^30 bases here (kmu_sketch_pmh3a_whole has no limit)");
        if (!ptrs) same_layout &= byte_off[i] == total_bytes;
        total_bytes += align_up((nbases[i] + 3) / 4, SEQ_ALIGN);
    }
    if (!ptrs) same_layout &= packed_bytes >= total_bytes;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScopedDevice sd(ctx->device);
    const auto t_begin = std::chrono::steady_clock::now();
    const bool trace = std::getenv("KMU_TRACE") != nullptr;
    auto stamp = [&](const char* what, size_t c) {
        if (trace)
            std::fprintf(stderr, "[kmu host pipe] %8.3f ms  %s %zu\n",
                         std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_begin).count(), what, c);
    };
    ctx->last = kmu_times{};
    HostPipe& hp = ctx->pipe;
    if (!hp.copy_in) {
        CUDA_TRY(cudaStreamCreateWithFlags(&hp.copy_in, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&hp.copy_out, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CUDA_TRY(cudaEventCreate(&hp.in_begin[i]));
            CUDA_TRY(cudaEventCreate(&hp.in_done[i]));
            CUDA_TRY(cudaEventCreate(&hp.compute_done[i]));
            CUDA_TRY(cudaEventCreate(&hp.out_begin[i]));
            CUDA_TRY(cudaEventCreate(&hp.out_done[i]));
        }
    }
    const size_t vsz = kmer_type == KMU_KMER64 ? 8 : 4;
    // batch layout of the whole input and the chunk boundaries
    std::vector<uint64_t> lay_own;
    if (!same_layout) layout_offsets(nbases, nseq, lay_own);
    const uint64_t* lay = same_layout ? byte_off : lay_own.data();
    // chunk sizes grow geometrically from a small first chunk: the upload of chunk c + 1 must hide behind the sketching
    // of chunk c, and PCIe moves packed bases only ~1.3 x as fast as the kernels consume them (50 GB/s against 39 GB/s
    // on C2), so a chunk may be at most ~1.3 x its predecessor: 25 % growth from 1/48 of the input (measured: 32.4 ms per
    // call; 50 % from 1/128 stalls on every upload of the growth phase, 34.0 ms; scripts/e2e_sweep.sh); large chunks
    // later (few launch tails); a small last chunk (short drain of the final download)
    uint64_t big = std::max<uint64_t>(64ull << 20, total_bytes / 4);
    uint64_t small = std::max<uint64_t>(8ull << 20, total_bytes / 48);
    uint64_t first = std::max<uint64_t>(4ull << 20, total_bytes / 48);
    bool grow = true;
    if (const char* env = std::getenv("KMU_HOST_CHUNK_BYTES")) {  // tests: force many small chunks
        const uint64_t v = std::strtoull(env, nullptr, 10);
        if (v) {
            big = small = first = v;
            grow = false;
        }
    }
    uint32_t grow_pct = 25;  // experiments: KMU_HOST_GROW_PCT, KMU_HOST_FIRST_DIV
    if (const char* env = std::getenv("KMU_HOST_GROW_PCT")) grow_pct = (uint32_t)std::strtoul(env, nullptr, 10);
    if (const char* env = std::getenv("KMU_HOST_FIRST_DIV")) {
        const uint64_t d = std::strtoull(env, nullptr, 10);
        if (d) first = std::max<uint64_t>(4ull << 20, total_bytes / d);
    }
    uint64_t next_target = first;
    std::vector<uint64_t> cut{0};
    for (uint64_t i = 0, start = 0; i < nseq; ++i) {
        const uint64_t end = i + 1 < nseq ? lay[i + 1] : total_bytes;
        const uint64_t left = total_bytes - lay[start];
        uint64_t target = next_target;
        if (cut.size() > 1 && left > small && left - small < target) target = left - small;  // leave a small tail chunk
        if (end - lay[start] >= target || i + 1 == nseq) {
            cut.push_back(i + 1);
            start = i + 1;
            if (grow) next_target = std::min<uint64_t>(big, next_target + next_target * grow_pct / 100);
        }
    }
    const size_t nchunks = cut.size() - 1;
    uint64_t max_bytes = 0, max_seqs = 0;
    for (size_t c = 0; c < nchunks; ++c) {
        const uint64_t e = cut[c + 1] < nseq ? lay[cut[c + 1]] : total_bytes;
        max_bytes = std::max(max_bytes, e - lay[cut[c]]);
        max_seqs = std::max(max_seqs, cut[c + 1] - cut[c]);
    }
    for (int sl = 0; sl < 2 && sl < (int)nchunks; ++sl) {
        cudaError_t e = hp.packed[sl].reserve(max_bytes + TAIL_SLACK);
        if (e == cudaSuccess) e = hp.meta[sl].reserve(2 * sizeof(uint64_t) * (max_seqs + 1));
        if (e == cudaSuccess) e = hp.sig[sl].reserve(max_seqs * m * vsz);
        if (e == cudaSuccess) e = hp.meta_host[sl].reserve(2 * sizeof(uint64_t) * (max_seqs + 1));
        // the processing order of a chunk lives in pipeline-owned buffers: no cudaMalloc (a device-wide
        // synchronisation) between chunks
        if (e == cudaSuccess) e = hp.order[sl].reserve(sizeof(uint32_t) * (max_seqs + 1));
        if (e == cudaSuccess) e = hp.cursor[sl].reserve(sizeof(unsigned long long) * kmu::LEN_BUCKETS);
        if (e == cudaSuccess && !same_layout) e = hp.stage[sl].reserve(max_bytes);
        if (e != cudaSuccess) return fail(KMU_ENOMEM, "host pipeline buffers: %s", cudaGetErrorString(e));
    }
    std::vector<kmu_seqbatch> views(nchunks);
    std::vector<char> timed_in(2, 0), timed_out(2, 0);
    auto harvest = [&](int sl) {  // add the finished copies of a slot to the totals
        float ms = 0;
        if (timed_in[sl] && cudaEventElapsedTime(&ms, hp.in_begin[sl], hp.in_done[sl]) == cudaSuccess) ctx->last.h2d_ms += ms;
        if (timed_out[sl] && cudaEventElapsedTime(&ms, hp.out_begin[sl], hp.out_done[sl]) == cudaSuccess) ctx->last.d2h_ms += ms;
        timed_in[sl] = timed_out[sl] = 0;
    };
    auto upload = [&](size_t c) -> int32_t {
        const int sl = (int)(c & 1);
        const uint64_t s0 = cut[c], s1 = cut[c + 1], n = s1 - s0;
        const uint64_t b0 = lay[s0], b1 = s1 < nseq ? lay[s1] : total_bytes;
        kmu_seqbatch& v = views[c];
        v.device = ctx->device;
        v.owns = false;
        v.packed = (uint8_t*)hp.packed[sl].p;
        v.byte_off = (uint64_t*)hp.meta[sl].p;
        v.nbases = (uint64_t*)hp.meta[sl].p + (max_seqs + 1);
        v.nseq = n;
        v.packed_bytes = b1 - b0;
        v.order_cache.order = hp.order[sl];  // borrowed: handed back when the chunk is done
        v.order_cache.cursor_dev = hp.cursor[sl];
        v.h_nbases.assign(nbases + s0, nbases + s1);
        v.h_byte_off.resize(n);
        for (uint64_t i = 0; i < n; ++i) {
            v.h_byte_off[i] = lay[s0 + i] - b0;
            v.total_bases += nbases[s0 + i];
        }
        // the slot's previous occupant (chunk c-2) must have been sketched and its copies harvested
        if (c >= 2) {
            // chunk c - 2 (same slot) has been sketched; its redo launch may still be running: the copy waits for it
            CUDA_TRY(cudaStreamWaitEvent(hp.copy_in, hp.compute_done[sl], 0));
            harvest(sl);
        }
        uint64_t* mh = (uint64_t*)hp.meta_host[sl].p;
        std::memcpy(mh, v.h_byte_off.data(), sizeof(uint64_t) * n);
        std::memcpy(mh + (max_seqs + 1), v.h_nbases.data(), sizeof(uint64_t) * n);
        const uint8_t* src = same_layout ? packed + b0 : nullptr;
        if (!same_layout) {
            if (c >= 2) {
                // the staging buffer of this slot fed the upload of chunk c - 2: that copy must be over before it is refilled
                CUDA_TRY(cudaEventSynchronize(hp.in_done[sl]));
            }
            if (ptrs) parallel_copy((uint8_t*)hp.stage[sl].p, v.h_byte_off, ptrs + s0, nullptr, nullptr, nbases + s0, n);
            else parallel_copy((uint8_t*)hp.stage[sl].p, v.h_byte_off, nullptr, packed, byte_off + s0, nbases + s0, n);
            src = (const uint8_t*)hp.stage[sl].p;
        }
        CUDA_TRY(cudaEventRecord(hp.in_begin[sl], hp.copy_in));
        CUDA_TRY(cudaMemcpyAsync(v.packed, src, v.packed_bytes, cudaMemcpyHostToDevice, hp.copy_in));
        CUDA_TRY(cudaMemsetAsync(v.packed + v.packed_bytes, 0, TAIL_SLACK, hp.copy_in));
        CUDA_TRY(cudaMemcpyAsync(hp.meta[sl].p, mh, 2 * sizeof(uint64_t) * (max_seqs + 1), cudaMemcpyHostToDevice, hp.copy_in));
        CUDA_TRY(cudaEventRecord(hp.in_done[sl], hp.copy_in));
        timed_in[sl] = 1;
        ctx->last.h2d_bytes += v.packed_bytes + 2 * sizeof(uint64_t) * n;
        return KMU_OK;
    };
    stamp("layout done, chunks", nchunks);
    // Per chunk c: upload chunk c + 1, build its processing order and queue its sketch launches [phase 1] behind
    // chunk c's on the main stream -- the GPU goes from one chunk to the next without waiting for the host; then wait
    // for chunk c's launches, redo the sequences it flagged [phase 2] on a side stream beside chunk c + 1, and start
    // the download.
    cudaEventRecord(ctx->ev[0], ctx->stream);
    // main launches of chunk c: chunk-local order + phase 1; compute_done[slot] marks their end (phase 2 moves it
    // behind the redo launch when there is one)
    auto enqueue_main = [&](size_t c) -> int32_t {
        const int sl = (int)(c & 1);
        CUDA_TRY(cudaStreamWaitEvent(ctx->stream, hp.in_done[sl], 0));
        if (c >= 2) {
            CUDA_TRY(cudaStreamWaitEvent(ctx->stream, hp.out_done[sl], 0));      // signature slot free again
            CUDA_TRY(cudaStreamWaitEvent(ctx->stream, hp.compute_done[sl], 0));  // and the redo launch that used this scratch set
        }
        int32_t r = kmu_ensure_order(ctx, &views[c], k, nullptr);
        // two sets of counters / redo lists: the redo launch of chunk c runs while chunk c + 1 is sketched
        if (!r) r = sketch_pmh3a_locked(ctx, &views[c], k, kmer_type, hash_kind, m, hp.sig[sl].p, 1, sl);
        if (!r) CUDA_TRY(cudaEventRecord(hp.compute_done[sl], ctx->stream));
        stamp("sketch enqueued", c);
        return r;
    };
    int32_t rc = upload(0);
    stamp("upload enqueued", 0);
    if (!rc) rc = enqueue_main(0);
    for (size_t c = 0; c < nchunks && rc == KMU_OK; ++c) {
        const int sl = (int)(c & 1);
        if (c + 1 < nchunks) {
            rc = upload(c + 1);
            if (!rc) rc = enqueue_main(c + 1);
            if (rc) break;
        }
        ctx->redo_on_side = ctx->redo_on_main = false;
        rc = sketch_pmh3a_locked(ctx, &views[c], k, kmer_type, hash_kind, m, hp.sig[sl].p, 2, sl, true);
        stamp("sketch finished", c);
        if (rc) break;
        hp.order[sl] = views[c].order_cache.order;
        hp.cursor[sl] = views[c].order_cache.cursor_dev;
        views[c].order_cache.order = DevBuf{};
        views[c].order_cache.cursor_dev = DevBuf{};
        // a CUDA failure here must not return before the three streams are idle: copies into the caller's `sig` and the
        // pinned staging buffers may still be in flight (every exit goes through the synchronisation below)
#define CUDA_BRK(expr)                                                                           \
    {                                                                                            \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            rc = fail(KMU_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(_e));               \
            break;                                                                               \
        }                                                                                        \
    }
        if (ctx->redo_on_side) CUDA_BRK(cudaEventRecord(hp.compute_done[sl], ctx->redo_stream))
        else if (ctx->redo_on_main) CUDA_BRK(cudaEventRecord(hp.compute_done[sl], ctx->stream))  // behind the next chunk's launches
        CUDA_BRK(cudaStreamWaitEvent(hp.copy_out, hp.compute_done[sl], 0))
        const size_t out_bytes = (size_t)views[c].nseq * m * vsz;
        CUDA_BRK(cudaEventRecord(hp.out_begin[sl], hp.copy_out))
        CUDA_BRK(cudaMemcpyAsync((uint8_t*)sig + (size_t)cut[c] * m * vsz, hp.sig[sl].p, out_bytes, cudaMemcpyDeviceToHost,
                                 hp.copy_out))
        CUDA_BRK(cudaEventRecord(hp.out_done[sl], hp.copy_out))
#undef CUDA_BRK
        timed_out[sl] = 1;
        ctx->last.d2h_bytes += out_bytes;
    }
    if (ctx->redo_stream) {  // the last redo launch belongs to the timed region
        if (!ctx->join_ev) cudaEventCreateWithFlags(&ctx->join_ev, cudaEventDisableTiming);
        cudaEventRecord(ctx->join_ev, ctx->redo_stream);
        cudaStreamWaitEvent(ctx->stream, ctx->join_ev, 0);
    }
    cudaEventRecord(ctx->ev[1], ctx->stream);
    cudaError_t e1 = cudaStreamSynchronize(hp.copy_in);
    cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    cudaError_t e3 = cudaStreamSynchronize(hp.copy_out);
    if (rc) return rc;
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
        ctx->table_scratch_clean = false;
        return fail(KMU_ECUDA, "host sketch pipeline failed: %s",
                    cudaGetErrorString(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3)));
    }
    harvest(0);
    harvest(1);
    {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]) == cudaSuccess) ctx->last.kernel_ms = ms;
    }
    stamp("all streams idle", 0);
    ctx->last.host_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    return KMU_OK;
}

int32_t kmu_sketch_pmh3a_host(kmu_ctx* ctx, const uint8_t* packed, uint64_t packed_bytes, const uint64_t* byte_off,
                              const uint64_t* nbases, uint64_t nseq, uint32_t k, int32_t kmer_type, int32_t hash_kind,
                              uint32_t m, void* sig) {
    return sketch_pmh3a_host_impl(ctx, packed, packed_bytes, byte_off, nullptr, nbases, nseq, k, kmer_type, hash_kind, m, sig);
}

int32_t kmu_sketch_pmh3a_host_ptrs(kmu_ctx* ctx, const uint8_t* const* seq_ptrs, const uint64_t* nbases, uint64_t nseq, uint32_t k,
                                   int32_t kmer_type, int32_t hash_kind, uint32_t m, void* sig) {
    if (nseq && !seq_ptrs) return fail(KMU_EINVAL, "null argument");
    return sketch_pmh3a_host_impl(ctx, nullptr, 0, nullptr, seq_ptrs, nbases, nseq, k, kmer_type, hash_kind, m, sig);
}

}  // extern "C"
