// kmu_capi_count.cu -- C ABI of the counting table (include/kmerutils_b200.h "k-mer counting").
// Replaces KmerCounter / KmerCounterPool and the count_kmer* drivers of src/base/kmercount.rs.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <vector>

#include "kmu_host.h"

struct kmu_counter {
    int device = 0;
    uint32_t k = 0;
    int kmer_type = 0;
    uint32_t count_bits = 8;
    bool key64 = false;
    uint64_t capacity = 0;  // slots (power of two)
    DevBuf slots, aux;      // aux: [0] special, [1] overflow, [2] export cursor, [8 .. 8+3+256) stats
    uint64_t inserted = 0;  // k-mers inserted so far (multiplicity total)
    kmu::CountTable view() const {
        kmu::CountTable t;
        t.slots = slots.p;
        t.capmask = capacity - 1;
        t.special = (unsigned long long*)aux.p;
        t.overflow = (unsigned long long*)aux.p + 1;
        return t;
    }
    uint32_t max_count() const { return count_bits >= 32 ? 0xFFFFFFFFu : ((1u << count_bits) - 1u); }
};

namespace {

constexpr size_t AUX_WORDS = 8 + 3 + 256;

int32_t check_overflow(kmu_ctx* ctx, kmu_counter* c) {
    unsigned long long ovf = 0;
    CUDA_TRY(cudaMemcpyAsync(&ovf, (unsigned long long*)c->aux.p + 1, sizeof(ovf), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (ovf)
        return fail(KMU_EOVERFLOW, "counting table of %llu slots is full: create the counter with a larger capacity",
                    (unsigned long long)c->capacity);
    return KMU_OK;
}


// ---- two-phase insertion (kmu_count_part.cu) -----------------------------------------------------------------
// The table is cut into regions of REGION_BYTES (64 MB: load + RED updates of a region that sits in L2 run at 64 G/s
// against 15.5 G/s on the whole table, profiles/r1d_micro_atomics.txt); more than max_buckets regions -> larger regions.
constexpr uint32_t MAX_BUCKETS = 4096;

double now_ms() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

// KMU_COUNT_REGION_KB / KMU_COUNT_TWO_PHASE_MIN_KEYS: test knobs (small tables through the regioned path)
uint64_t region_bytes_setting() {
    if (const char* e = std::getenv("KMU_COUNT_REGION_KB")) {
        const uint64_t kb = (uint64_t)std::atoll(e);
        if (kb >= 1) return kb << 10;
    }
    return 128ull << 20;
}

// The table is cut into `nfine_total` regions of region_bytes_setting() (128 MB, the size of the L2: measured on 3.2 G 31-mers
// into a 64 GB table, partition + insertion: 256 regions of 256 MB 31.1 + 92.3 ms, 512 of 128 MB 32.9 + 75.4, 1024 of 64 MB
// 39.3 + 73.0, 2048 of 32 MB 55.0 + 77.9 -- fewer buckets make longer runs per tile in the partition kernel, and the
// insertion, bound by the L2's dependent load + RED rate rather than by misses, loses little until the region is twice the L2).  The partition kernel is efficient up to a few hundred buckets per pass (a tile of 4096 k-mers sorted in shared
// memory: with thousands of buckets every bucket gets one key per tile, i.e. one global atomic and one 8-byte store per
// key -- measured 14.7 G keys/s at 4096 buckets against 77 G at 512), so large tables are partitioned in TWO levels:
// level 1 (from the reads; across GPUs: owner x coarse region) makes ncoarse coarse regions per owner, level 2 cuts
// each coarse slab into nfine fine regions just before they are inserted.
struct RegionGeom {
    uint32_t ncoarse = 1;       // level-1 regions of an owner's table (== regions when nfine == 1)
    uint32_t nfine = 1;         // level-2 regions per coarse region
    uint32_t coarse_shift = 0;  // log2(slots per coarse region)
    uint32_t fine_shift = 0;    // log2(slots per fine region)
    uint32_t regions() const { return ncoarse * nfine; }
};

uint32_t level1_bucket_target() {
    if (const char* e = std::getenv("KMU_COUNT_LEVEL1_BUCKETS")) {
        const long v = std::atol(e);
        if (v >= 1) return (uint32_t)v;
    }
    return 1024;
}

RegionGeom region_geometry(uint64_t capacity, bool key64, uint32_t nowners) {
    const uint64_t slot_bytes = key64 ? 16 : 8;
    uint64_t region_slots = 1;
    while (region_slots * 2 * slot_bytes <= region_bytes_setting()) region_slots <<= 1;
    if (nowners < 1) nowners = 1;
    uint64_t total = capacity > region_slots ? capacity / region_slots : 1;  // a power of two
    while (total > 65536) {  // never more than 64 K regions: larger regions instead
        total >>= 1;
        region_slots <<= 1;
    }
    RegionGeom g;
    uint64_t ncoarse = total, nfine = 1;
    while (ncoarse > 1 && ncoarse * nowners > level1_bucket_target()) {
        ncoarse >>= 1;
        nfine <<= 1;
    }
    while (nfine > 1024) {  // level 2 too wide: give level 1 more buckets (MAX_BUCKETS bounds it)
        nfine >>= 1;
        ncoarse <<= 1;
    }
    g.ncoarse = (uint32_t)ncoarse;
    g.nfine = (uint32_t)nfine;
    while ((1ull << g.fine_shift) < region_slots) ++g.fine_shift;
    g.coarse_shift = g.fine_shift;
    while ((1ull << g.coarse_shift) < region_slots * nfine) ++g.coarse_shift;
    return g;
}

// slab capacity for `n` keys spread over `nbuckets` by a hash: the expected share + 64 sqrt(share) + slack.  The keys
// are k-mer OCCURRENCES: a key seen c times puts c entries into one bucket, so the spread of a bucket is sqrt(c) times
// that of distinct keys -- 64 sqrt(share) is 8 sigma at a mean multiplicity of 64 (the C3 shape has 10-40).
uint64_t slab_capacity(uint64_t n, uint64_t nbuckets) {
    const double mean = (double)n / (double)nbuckets;
    return (uint64_t)(mean + 64.0 * std::sqrt(mean) + 1024.0);
}

bool two_phase_wanted(const kmu_counter* c, uint64_t nkeys) {
    if (std::getenv("KMU_COUNT_DIRECT")) return false;
    const uint64_t table_bytes = c->capacity * (c->key64 ? 16 : 8);
    uint64_t min_keys = 1ull << 20;
    if (const char* e = std::getenv("KMU_COUNT_TWO_PHASE_MIN_KEYS")) min_keys = (uint64_t)std::atoll(e);
    return table_bytes >= 4 * region_bytes_setting() && nkeys >= min_keys;
}

// bytes the partition slabs of one chunk may take: 32 GB, at most half of the free memory (the buffer already held counts as free)
uint64_t slab_budget_bytes(uint64_t held) {
    uint64_t budget = 32ull << 30;
    if (const char* e = std::getenv("KMU_COUNT_SLAB_MB")) return (uint64_t)std::atoll(e) << 20;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    return std::min<uint64_t>(budget, ((uint64_t)free_b + held) / 2);
}

// scratch in ctx->counters (u64 words): [0, MAX_BUCKETS) level-1 cursors | 64 chunk flags | 8 (one device pointer) | 64 pointers
// (exchange destinations) | 65536 level-2 cursors | 4096 level-2 flags (one per coarse region) | 65536 u32 `done` counters
constexpr size_t OFF_FLAGS = MAX_BUCKETS, OFF_PTR = OFF_FLAGS + 64, OFF_DESTS = OFF_PTR + 8, OFF_CUR2 = OFF_DESTS + 64,
                 OFF_FLAGS2 = OFF_CUR2 + 65536, OFF_DONE = OFF_FLAGS2 + 4096, PART_SCRATCH_WORDS = OFF_DONE + 65536 / 2 + 8;

// Regioned insertion of level-1 slabs: slab (c, s) = keys of sender s for coarse region c at slabs + (c * nsend + s) * slab_cap,
// counts_dev[s * ncoarse + c] keys each (device array).  nfine == 1: one launch over all regions.  Otherwise, per coarse
// region: partition its slabs into nfine fine slabs (ctx->part_fine), insert those.  skip_flag (device, may be null): the
// whole call does nothing when set.  A fine slab that overflows (flags2[c]) makes the host redo that coarse region directly.
int32_t insert_level1_slabs(kmu_ctx* ctx, kmu_counter* c, const void* slabs, uint64_t slab_cap, uint32_t nsend,
                            const unsigned long long* counts_dev, const RegionGeom& rg, const unsigned long long* skip_flag,
                            uint64_t* launches) {
    const size_t esz = c->key64 ? 8 : 4;
    // the L2 prefetch of the next region is OFF: measured -2 to -7 ms per 3.2 G keys without it -- ncu: the prefetched lines (31 GB
    // more DRAM reads) do not serve the atomics, whose L2 misses stay the same (KMU_COUNT_PREFETCH=1 turns it back on)
    const bool prefetch = std::getenv("KMU_COUNT_PREFETCH") != nullptr;
    unsigned long long* scratch = (unsigned long long*)ctx->counters.p;
    unsigned int* done = std::getenv("KMU_COUNT_FREE_RUNNING") ? nullptr : (unsigned int*)(scratch + OFF_DONE);
    if (done) CUDA_TRY(cudaMemsetAsync(done, 0, sizeof(unsigned int) * rg.regions(), ctx->stream));
    if (rg.nfine == 1) {
        CUDA_TRY(kmu::launch_count_insert_slabs(slabs, slab_cap, rg.ncoarse, nsend, counts_dev, c->view(), c->key64, rg.fine_shift,
                                                skip_flag, prefetch, done, ctx->sm_count, ctx->stream));
        *launches += 1;
        return KMU_OK;
    }
    unsigned long long* cur2 = scratch + OFF_CUR2;
    unsigned long long* flags2 = scratch + OFF_FLAGS2;
    void** d_fine = (void**)(scratch + OFF_PTR);
    const uint64_t fine_cap = slab_capacity((uint64_t)nsend * slab_cap, rg.nfine);
    if (fine_cap >= 0xFFFFFFFFull) return fail(KMU_EINVAL, "partition slab too large");
    CUDA_TRY(ctx->part_fine.reserve(fine_cap * rg.nfine * esz));
    void* fine_ptr = ctx->part_fine.p;
    CUDA_TRY(cudaMemcpyAsync(d_fine, &fine_ptr, sizeof(void*), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(cur2, 0, sizeof(unsigned long long) * (65536 + 4096), ctx->stream));
    kmu::PartGeom g{};
    g.nowners = 1;
    g.nregions = rg.nfine;
    g.capmask = ((1ull << rg.coarse_shift) - 1);  // the fine region inside the coarse one
    g.shift = rg.fine_shift;
    g.nsend = 1;
    g.self = 0;
    g.slab_cap = fine_cap;
    for (uint32_t cr = 0; cr < rg.ncoarse; ++cr) {
        kmu::KeySegs segs;
        segs.stride = slab_cap;
        segs.nseg = nsend;
        segs.counts = counts_dev + cr;
        segs.count_stride = rg.ncoarse;
        segs.skip_flag = skip_flag;
        const uint8_t* base = (const uint8_t*)slabs + (uint64_t)cr * nsend * slab_cap * esz;
        CUDA_TRY(kmu::launch_count_part_keys(base, slab_cap, segs, c->key64, g, d_fine, cur2 + (uint64_t)cr * rg.nfine, flags2 + cr,
                                             ctx->sm_count, ctx->stream));
        CUDA_TRY(kmu::launch_count_insert_slabs(fine_ptr, fine_cap, rg.nfine, 1, cur2 + (uint64_t)cr * rg.nfine, c->view(), c->key64,
                                                rg.fine_shift, flags2 + cr, prefetch, done ? done + (uint64_t)cr * rg.nfine : nullptr,
                                                ctx->sm_count, ctx->stream, cr * rg.nfine, rg.nfine));
        *launches += 2;
    }
    // level-2 overflows (a key repeated millions of times inside one coarse region): those coarse regions go in directly
    std::vector<unsigned long long> f2(rg.ncoarse), skip(1, 0);
    CUDA_TRY(cudaMemcpyAsync(f2.data(), flags2, sizeof(unsigned long long) * rg.ncoarse, cudaMemcpyDeviceToHost, ctx->stream));
    if (skip_flag) CUDA_TRY(cudaMemcpyAsync(skip.data(), skip_flag, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (skip[0]) return KMU_OK;
    std::vector<unsigned long long> hc;
    for (uint32_t cr = 0; cr < rg.ncoarse; ++cr) {
        if (!f2[cr]) continue;
        if (hc.empty()) {
            hc.resize((size_t)nsend * rg.ncoarse);
            CUDA_TRY(cudaMemcpy(hc.data(), counts_dev, sizeof(unsigned long long) * hc.size(), cudaMemcpyDeviceToHost));
        }
        for (uint32_t sd = 0; sd < nsend; ++sd) {
            const uint64_t n = std::min<uint64_t>(hc[(size_t)sd * rg.ncoarse + cr], slab_cap);
            const uint8_t* seg = (const uint8_t*)slabs + ((uint64_t)cr * nsend + sd) * slab_cap * esz;
            CUDA_TRY(kmu::launch_count_insert_keys(seg, n, c->key64, c->view(), ctx->sm_count, ctx->stream));
            *launches += 1;
        }
    }
    return KMU_OK;
}

// insert the k-mers of the batch (src == nullptr) or the device key array `src` through partition + regioned insertion.
// *launches is increased by the kernels launched.  Chunks whose partition overflowed a slab are redone directly.
int32_t insert_two_phase(kmu_ctx* ctx, kmu_counter* c, const kmu_seqbatch* b, bool canonical, const void* src, uint64_t nsrc,
                         uint64_t* launches) {
    const double t_in = now_ms();
    const size_t esz = c->key64 ? 8 : 4;
    const RegionGeom rg = region_geometry(c->capacity, c->key64, 1);
    // chunks: a bound of the k-mers of a chunk = 4 per packed byte (sequences) or the keys themselves
    const uint64_t unit_total = b ? b->packed_bytes : nsrc;          // bytes or keys
    const uint64_t keys_per_unit = b ? 4 : 1;
    // the exact number of k-mers (known on the host from the lengths): 150-base reads hold 120 31-mers, not 150 -- sizing
    // the slabs by it lets a 4 Gbase step go through in ONE chunk, and every chunk costs a full sweep of the table
    const uint64_t keys_total = b ? b->kmer_count(c->k) : nsrc;
    // one chunk if its slabs fit the buffer already held (no query of the free memory on the hot path)
    uint64_t chunk_units = unit_total;
    if (std::getenv("KMU_COUNT_SLAB_MB") || slab_capacity(keys_total, rg.ncoarse) * rg.ncoarse * esz > ctx->sig_dev.cap) {
        const uint64_t budget = slab_budget_bytes(ctx->sig_dev.cap);
        if (slab_capacity(keys_total, rg.ncoarse) * rg.ncoarse * esz > budget) {
            chunk_units = std::max<uint64_t>(1, budget / (esz * keys_per_unit) * 9 / 10);
            if (b) chunk_units = std::max<uint64_t>(2048, chunk_units / 2048 * 2048);  // GROUP_BYTES of kmu_device.cuh
            chunk_units = std::min(chunk_units, unit_total);
        }
    }
    const uint64_t nchunks = (unit_total + chunk_units - 1) / chunk_units;
    if (nchunks > 64) return fail(KMU_ENOMEM, "not enough free device memory for the partition slabs (%llu chunks)", (unsigned long long)nchunks);
    const uint64_t bound = std::min(chunk_units * keys_per_unit, keys_total);
    const uint64_t slab_cap = slab_capacity(bound, rg.ncoarse);
    if (slab_cap >= 0xFFFFFFFFull) return fail(KMU_EINVAL, "partition slab too large");
    CUDA_TRY(ctx->sig_dev.reserve(slab_cap * rg.ncoarse * esz));
    CUDA_TRY(ctx->counters.reserve(sizeof(unsigned long long) * PART_SCRATCH_WORDS));
    unsigned long long* cursors = (unsigned long long*)ctx->counters.p;
    unsigned long long* flags = cursors + OFF_FLAGS;
    void** d_dest = (void**)(cursors + OFF_DESTS);
    void* slab_ptr = ctx->sig_dev.p;
    CUDA_TRY(cudaMemcpyAsync(d_dest, &slab_ptr, sizeof(void*), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(flags, 0, sizeof(unsigned long long) * 64, ctx->stream));
    kmu::PartGeom g{};
    g.nowners = 1;
    g.nregions = rg.ncoarse;
    g.capmask = c->capacity - 1;
    g.shift = rg.coarse_shift;
    g.nsend = 1;
    g.self = 0;
    g.slab_cap = slab_cap;
    kmu::SeqView v{};
    if (b) v = kmu::SeqView{b->packed, b->byte_off, b->nbases, b->nseq};
    const bool timing = std::getenv("KMU_COUNT_TIMING") != nullptr;
    cudaEvent_t tev[3] = {nullptr, nullptr, nullptr};
    if (timing)
        for (auto& e : tev) cudaEventCreate(&e);
    for (uint64_t ch = 0; ch < nchunks; ++ch) {
        const uint64_t u0 = ch * chunk_units, u1 = std::min(unit_total, u0 + chunk_units);
        CUDA_TRY(cudaMemsetAsync(cursors, 0, sizeof(unsigned long long) * rg.ncoarse, ctx->stream));
        if (timing) cudaEventRecord(tev[0], ctx->stream);
        if (b)
            CUDA_TRY(kmu::launch_count_part_seqs(v, u0, u1, b->packed_bytes, c->k, c->key64, canonical, g, d_dest, cursors, flags + ch,
                                                 ctx->sm_count, ctx->stream));
        else
            CUDA_TRY(kmu::launch_count_part_keys((const uint8_t*)src + u0 * esz, u1 - u0, kmu::KeySegs{}, c->key64, g, d_dest, cursors,
                                                 flags + ch, ctx->sm_count, ctx->stream));
        *launches += 1;
        if (timing) cudaEventRecord(tev[1], ctx->stream);
        int32_t rc = insert_level1_slabs(ctx, c, slab_ptr, slab_cap, 1, cursors, rg, flags + ch, launches);
        if (rc) return rc;
        if (timing) {
            cudaEventRecord(tev[2], ctx->stream);
            cudaEventSynchronize(tev[2]);
            float a = 0, bms = 0;
            cudaEventElapsedTime(&a, tev[0], tev[1]);
            cudaEventElapsedTime(&bms, tev[1], tev[2]);
            std::fprintf(stderr, "[kmu count] chunk %llu/%llu: %u x %u regions of %llu KB, slab_cap %llu, level-1 partition %.2f ms, "
                                 "level-2 partition + insertion %.2f ms (host %.2f ms so far)\n",
                         (unsigned long long)ch, (unsigned long long)nchunks, rg.ncoarse, rg.nfine,
                         (unsigned long long)(((c->key64 ? 16ull : 8ull) << rg.fine_shift) >> 10), (unsigned long long)slab_cap, a, bms,
                         now_ms() - t_in);
        }
    }
    if (timing)
        for (auto& e : tev) cudaEventDestroy(e);
    unsigned long long hflags[64];
    CUDA_TRY(cudaMemcpyAsync(hflags, flags, sizeof(hflags), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (uint64_t ch = 0; ch < nchunks; ++ch) {
        if (!hflags[ch]) continue;
        const uint64_t u0 = ch * chunk_units, u1 = std::min(unit_total, u0 + chunk_units);
        if (b)
            CUDA_TRY(kmu::launch_count_insert_seqs(v, u1, c->k, c->key64, canonical, c->view(), ctx->sm_count, ctx->stream, u0));
        else
            CUDA_TRY(kmu::launch_count_insert_keys((const uint8_t*)src + u0 * esz, u1 - u0, c->key64, c->view(), ctx->sm_count, ctx->stream));
        *launches += 1;
    }
    return KMU_OK;
}

// The keys of `nseg` segments `stride` keys apart (counts_host[s] = counts_dev[s] keys in segment s: what the senders of
// an exchange stored into this rank's receive buffer): ONE partition by region of the table over all segments, then the
// regioned insertion -- one sweep of the table whatever the number of senders.  Segments too large for the slab budget
// go through insert_two_phase one by one.
int32_t insert_segments(kmu_ctx* ctx, kmu_counter* c, const void* keys, uint64_t stride, uint32_t nseg, const uint64_t* counts_host,
                        const unsigned long long* counts_dev, uint64_t total, uint64_t* launches) {
    const size_t esz = c->key64 ? 8 : 4;
    if (total == 0) return KMU_OK;
    const RegionGeom rg = region_geometry(c->capacity, c->key64, 1);
    const uint64_t slab_cap = slab_capacity(total, rg.ncoarse);
    const uint64_t need = slab_cap * rg.ncoarse * esz;
    if (!two_phase_wanted(c, total) || slab_cap >= 0xFFFFFFFFull || (need > ctx->sig_dev.cap && need > slab_budget_bytes(ctx->sig_dev.cap))) {
        for (uint32_t sg = 0; sg < nseg; ++sg) {
            if (!counts_host[sg]) continue;
            const uint8_t* seg = (const uint8_t*)keys + (uint64_t)sg * stride * esz;
            if (two_phase_wanted(c, counts_host[sg])) {
                int32_t rc = insert_two_phase(ctx, c, nullptr, false, seg, counts_host[sg], launches);
                if (rc) return rc;
            } else {
                CUDA_TRY(kmu::launch_count_insert_keys(seg, counts_host[sg], c->key64, c->view(), ctx->sm_count, ctx->stream));
                *launches += 1;
            }
        }
        return KMU_OK;
    }
    CUDA_TRY(ctx->sig_dev.reserve(need));
    CUDA_TRY(ctx->counters.reserve(sizeof(unsigned long long) * PART_SCRATCH_WORDS));
    unsigned long long* cursors = (unsigned long long*)ctx->counters.p;
    unsigned long long* flags = cursors + OFF_FLAGS;
    void** d_dest = (void**)(cursors + OFF_DESTS);
    void* slab_ptr = ctx->sig_dev.p;
    CUDA_TRY(cudaMemcpyAsync(d_dest, &slab_ptr, sizeof(void*), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(cursors, 0, sizeof(unsigned long long) * (MAX_BUCKETS + 64), ctx->stream));
    kmu::PartGeom g{};
    g.nowners = 1;
    g.nregions = rg.ncoarse;
    g.capmask = c->capacity - 1;
    g.shift = rg.coarse_shift;
    g.nsend = 1;
    g.self = 0;
    g.slab_cap = slab_cap;
    kmu::KeySegs segs;
    segs.stride = stride;
    segs.nseg = nseg;
    segs.counts = counts_dev;
    segs.count_stride = 1;
    CUDA_TRY(kmu::launch_count_part_keys(keys, stride, segs, c->key64, g, d_dest, cursors, flags, ctx->sm_count, ctx->stream));
    *launches += 1;
    int32_t rc = insert_level1_slabs(ctx, c, slab_ptr, slab_cap, 1, cursors, rg, flags, launches);
    if (rc) return rc;
    unsigned long long lost = 0;
    CUDA_TRY(cudaMemcpyAsync(&lost, flags, sizeof(lost), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (lost) {  // a region slab overflowed (one key repeated millions of times): the insertion was skipped, go in directly
        for (uint32_t sg = 0; sg < nseg; ++sg) {
            if (!counts_host[sg]) continue;
            CUDA_TRY(kmu::launch_count_insert_keys((const uint8_t*)keys + (uint64_t)sg * stride * esz, counts_host[sg], c->key64, c->view(),
                                                   ctx->sm_count, ctx->stream));
            *launches += 1;
        }
    }
    return KMU_OK;
}

}  // namespace

extern "C" {

int32_t kmu_count_create(kmu_ctx* ctx, uint32_t k, int32_t kmer_type, uint32_t count_bits, uint64_t capacity,
                         kmu_counter** out) {
    if (!ctx || !out) return fail(KMU_EINVAL, "null argument");
    *out = nullptr;
    if (kmer_type_is_aa(kmer_type)) return fail(KMU_EINVAL, "the counter takes DNA k-mer types (KmerCounter, kmercount.rs:70)");
    if (!kmer_type_accepts(k, kmer_type))
        return fail(KMU_EINVAL, "kmer size %u is not supported by kmer type %d", k, kmer_type);
    if (count_bits < 1 || count_bits > 32) return fail(KMU_EINVAL, "count_bits must be in 1..32, got %u", count_bits);
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScopedDevice sd(ctx->device);
    auto* c = new kmu_counter();
    c->device = ctx->device;
    c->k = k;
    c->kmer_type = kmer_type;
    c->count_bits = count_bits;
    c->key64 = kmer_type == KMU_KMER64;
    // load factor <= 0.5 at the stated capacity; never more slots than there are possible keys (x2)
    uint64_t want = std::max<uint64_t>(1024, capacity * 2);
    if (2 * k < 40) want = std::min<uint64_t>(want, std::max<uint64_t>(1024, 2ull << (2 * k)));
    uint64_t cap = 1024;
    while (cap < want) cap <<= 1;
    c->capacity = cap;
    const size_t slot_bytes = c->key64 ? 16 : 8;
    cudaError_t e = c->slots.reserve(cap * slot_bytes);
    if (e == cudaSuccess) e = c->aux.reserve(sizeof(unsigned long long) * AUX_WORDS);
    if (e != cudaSuccess) {
        c->slots.release();
        c->aux.release();
        delete c;
        return fail(KMU_ENOMEM, "counting table of %llu slots (%llu bytes): %s", (unsigned long long)cap,
                    (unsigned long long)(cap * slot_bytes), cudaGetErrorString(e));
    }
    e = cudaMemsetAsync(c->aux.p, 0, sizeof(unsigned long long) * AUX_WORDS, ctx->stream);
    if (e == cudaSuccess) e = kmu::launch_count_init(c->view(), c->key64, ctx->sm_count, ctx->stream);
    ctx->launches += 1;
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        c->slots.release();
        c->aux.release();
        delete c;
        return fail(KMU_ECUDA, "counting table initialisation failed: %s", cudaGetErrorString(e));
    }
    *out = c;
    return KMU_OK;
}

void kmu_count_destroy(kmu_counter* c) {
    if (!c) return;
    ScopedDevice sd(c->device);
    c->slots.release();
    c->aux.release();
    delete c;
}

uint64_t kmu_count_capacity(const kmu_counter* c) { return c ? c->capacity : 0; }

int32_t kmu_count_insert_seqs(kmu_ctx* ctx, kmu_counter* c, const kmu_seqbatch* b, int32_t canonical) {
    const double t_enter = now_ms();
    if (!ctx || !c || !b) return fail(KMU_EINVAL, "null argument");
    if (b->alphabet != 0) return fail(KMU_EINVAL, "the counter takes DNA sequences");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScopedDevice sd(ctx->device);
    ctx->last = kmu_times{};
    if (b->nseq == 0) return KMU_OK;
    kmu::SeqView v{b->packed, b->byte_off, b->nbases, b->nseq};
    const uint64_t total = b->kmer_count(c->k);
    // Tables far larger than L2 take the batch in two phases (partition by table region, then region after region with
    // the updates hitting L2, kmu_count_part.cu); small tables and small batches are inserted directly.
    cudaEventRecord(ctx->ev[0], ctx->stream);
    if (two_phase_wanted(c, total)) {
        uint64_t nl = 0;
        int32_t rc2 = insert_two_phase(ctx, c, b, canonical != 0, nullptr, 0, &nl);
        if (rc2) return rc2;
        ctx->launches += nl;
        ctx->last.launches = nl;
    } else {
        CUDA_TRY(kmu::launch_count_insert_seqs(v, b->packed_bytes, c->k, c->key64, canonical != 0, c->view(), ctx->sm_count,
                                               ctx->stream));
        ctx->launches += 1;
        ctx->last.launches = 1;
    }
    cudaEventRecord(ctx->ev[1], ctx->stream);
    const double t_q = now_ms();
    int32_t rc = check_overflow(ctx, c);
    cudaEventElapsedTime(&ctx->last.kernel_ms, ctx->ev[0], ctx->ev[1]);
    if (std::getenv("KMU_COUNT_TIMING"))
        std::fprintf(stderr, "[kmu count] insert_seqs: host %.2f ms until queued, %.2f ms in all; events %.2f ms\n", t_q - t_enter,
                     now_ms() - t_enter, ctx->last.kernel_ms);
    if (rc) return rc;
    c->inserted += total;
    return KMU_OK;
}

int32_t kmu_count_insert_kmers(kmu_ctx* ctx, kmu_counter* c, const void* kmers, uint64_t n, int32_t on_device) {
    if (!ctx || !c || (n && !kmers)) return fail(KMU_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScopedDevice sd(ctx->device);
    ctx->last = kmu_times{};
    if (n == 0) return KMU_OK;
    const size_t esz = c->key64 ? 8 : 4;
    const void* d = kmers;
    if (!on_device) {
        CUDA_TRY(ctx->misc.reserve(n * esz));
        cudaEventRecord(ctx->ev[2], ctx->stream);
        CUDA_TRY(cudaMemcpyAsync(ctx->misc.p, kmers, n * esz, cudaMemcpyHostToDevice, ctx->stream));
        cudaEventRecord(ctx->ev[3], ctx->stream);
        ctx->last.h2d_bytes = n * esz;
        d = ctx->misc.p;
    }
    cudaEventRecord(ctx->ev[0], ctx->stream);
    if (two_phase_wanted(c, n)) {
        uint64_t nl = 0;
        int32_t rc2 = insert_two_phase(ctx, c, nullptr, false, d, n, &nl);
        if (rc2) return rc2;
        ctx->launches += nl;
        ctx->last.launches = nl;
    } else {
        CUDA_TRY(kmu::launch_count_insert_keys(d, n, c->key64, c->view(), ctx->sm_count, ctx->stream));
        ctx->launches += 1;
        ctx->last.launches = 1;
    }
    cudaEventRecord(ctx->ev[1], ctx->stream);
    int32_t rc = check_overflow(ctx, c);
    cudaEventElapsedTime(&ctx->last.kernel_ms, ctx->ev[0], ctx->ev[1]);
    if (!on_device) cudaEventElapsedTime(&ctx->last.h2d_ms, ctx->ev[2], ctx->ev[3]);
    if (rc) return rc;
    c->inserted += n;
    return KMU_OK;
}

int32_t kmu_count_query(kmu_ctx* ctx, const kmu_counter* c, const void* kmers, uint64_t n, uint32_t* counts,
                        int32_t on_device) {
    if (!ctx || !c || (n && (!kmers || !counts))) return fail(KMU_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScopedDevice sd(ctx->device);
    ctx->last = kmu_times{};
    if (n == 0) return KMU_OK;
    const size_t esz = c->key64 ? 8 : 4;
    const void* dk = kmers;
    uint32_t* dc = counts;
    if (!on_device) {
        CUDA_TRY(ctx->misc.reserve(n * (esz + 4) + 16));
        CUDA_TRY(cudaMemcpyAsync(ctx->misc.p, kmers, n * esz, cudaMemcpyHostToDevice, ctx->stream));
        dk = ctx->misc.p;
        dc = (uint32_t*)((uint8_t*)ctx->misc.p + align_up(n * esz, 16));
    }
    cudaEventRecord(ctx->ev[0], ctx->stream);
    CUDA_TRY(kmu::launch_count_query(dk, n, c->key64, c->view(), c->max_count(), dc, ctx->sm_count, ctx->stream));
    cudaEventRecord(ctx->ev[1], ctx->stream);
    ctx->launches += 1;
    ctx->last.launches = 1;
    if (!on_device) CUDA_TRY(cudaMemcpyAsync(counts, dc, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->last.kernel_ms, ctx->ev[0], ctx->ev[1]);
    return KMU_OK;
}

int32_t kmu_count_stats(kmu_ctx* ctx, const kmu_counter* c, uint64_t* nb_distinct, uint64_t* nb_unique,
                        uint64_t* nb_inserted, uint64_t* hist256) {
    if (!ctx || !c) return fail(KMU_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScopedDevice sd(ctx->device);
    unsigned long long* stats = (unsigned long long*)c->aux.p + 8;
    CUDA_TRY(cudaMemsetAsync(stats, 0, sizeof(unsigned long long) * (3 + 256), ctx->stream));
    CUDA_TRY(kmu::launch_count_stats(c->view(), c->key64, stats, ctx->sm_count, ctx->stream));
    ctx->launches += 1;
    std::vector<unsigned long long> h(3 + 256);
    CUDA_TRY(cudaMemcpyAsync(h.data(), stats, sizeof(unsigned long long) * (3 + 256), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    uint64_t distinct = 0;
    for (int i = 1; i < 256; ++i) distinct += h[3 + i];
    if (nb_distinct) *nb_distinct = distinct;
    if (nb_unique) *nb_unique = h[3 + 1];
    if (nb_inserted) *nb_inserted = h[2];
    if (hist256)
        for (int i = 0; i < 256; ++i) hist256[i] = h[3 + i];
    return KMU_OK;
}

int32_t kmu_count_export(kmu_ctx* ctx, const kmu_counter* c, uint32_t min_count, void* kmers, uint32_t* counts,
                         uint64_t cap, uint64_t* n_out) {
    if (!ctx || !c || !n_out || (cap && (!kmers || !counts))) return fail(KMU_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScopedDevice sd(ctx->device);
    const size_t esz = c->key64 ? 8 : 4;
    unsigned long long* cursor = (unsigned long long*)c->aux.p + 2;
    CUDA_TRY(cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), ctx->stream));
    CUDA_TRY(ctx->misc.reserve(cap * (esz + 4) + 32));
    void* dk = ctx->misc.p;
    uint32_t* dc = (uint32_t*)((uint8_t*)ctx->misc.p + align_up(cap * esz, 16));
    CUDA_TRY(kmu::launch_count_export(c->view(), c->key64, min_count, dk, dc, cursor, cap, ctx->sm_count, ctx->stream));
    ctx->launches += 1;
    unsigned long long n = 0;
    CUDA_TRY(cudaMemcpyAsync(&n, cursor, sizeof(n), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    // the u64 key equal to the empty sentinel lives outside the table
    unsigned long long special = 0;
    if (c->key64) CUDA_TRY(cudaMemcpy(&special, c->aux.p, sizeof(special), cudaMemcpyDeviceToHost));
    const uint64_t take = std::min<uint64_t>(n, cap);
    if (take) {
        CUDA_TRY(cudaMemcpyAsync(kmers, dk, take * esz, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(counts, dc, take * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    uint64_t total = n;
    if (special >= min_count && special > 0) {
        if (total < cap) {
            ((uint64_t*)kmers)[total] = ~0ULL;
            counts[total] = special > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)special;
        }
        ++total;
    }
    *n_out = total;
    if (total > cap) return fail(KMU_EOVERFLOW, "%llu k-mers to export, buffer holds %llu", (unsigned long long)total,
                                 (unsigned long long)cap);
    return KMU_OK;
}

int32_t kmu_count_partition(kmu_ctx* ctx, const kmu_seqbatch* b, uint32_t k, int32_t kmer_type, int32_t canonical,
                            uint32_t nparts, void* kmers_out, uint64_t* part_counts, int32_t out_on_device) {
    if (!ctx || !b || !part_counts) return fail(KMU_EINVAL, "null argument");
    if (kmer_type_is_aa(kmer_type) || b->alphabet != 0) return fail(KMU_EINVAL, "partitioning takes DNA k-mers");
    if (!kmer_type_accepts(k, kmer_type))
        return fail(KMU_EINVAL, "kmer size %u is not supported by kmer type %d", k, kmer_type);
    if (nparts < 1 || nparts > 64) return fail(KMU_EINVAL, "nparts must be in 1..64, got %u", nparts);
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScopedDevice sd(ctx->device);
    ctx->last = kmu_times{};
    for (uint32_t p = 0; p < nparts; ++p) part_counts[p] = 0;
    const uint64_t total = b->kmer_count(k);
    if (total == 0) return KMU_OK;
    if (!kmers_out) return fail(KMU_EINVAL, "null output buffer");
    const bool key64 = kmer_type == KMU_KMER64;
    const size_t esz = key64 ? 8 : 4;
    const int grid = kmu::count_partition_grid(b->packed_bytes, ctx->sm_count);
    CUDA_TRY(ctx->counters.reserve(sizeof(unsigned long long) * ((size_t)nparts * grid + 2 * nparts + 64)));
    unsigned long long* block_counts = (unsigned long long*)ctx->counters.p;
    unsigned long long* part_totals = block_counts + (size_t)nparts * grid;
    void* dout = kmers_out;
    if (!out_on_device) {
        CUDA_TRY(ctx->sig_dev.reserve(total * esz));
        dout = ctx->sig_dev.p;
    }
    kmu::SeqView v{b->packed, b->byte_off, b->nbases, b->nseq};
    cudaEventRecord(ctx->ev[0], ctx->stream);
    CUDA_TRY(kmu::launch_count_partition(v, b->packed_bytes, k, key64, canonical != 0, nparts, grid, block_counts,
                                         part_totals, dout, ctx->stream));
    cudaEventRecord(ctx->ev[1], ctx->stream);
    ctx->launches += 5;
    ctx->last.launches = 5;
    std::vector<unsigned long long> pt(nparts);
    CUDA_TRY(cudaMemcpyAsync(pt.data(), part_totals, sizeof(unsigned long long) * nparts, cudaMemcpyDeviceToHost, ctx->stream));
    if (!out_on_device) {
        CUDA_TRY(cudaMemcpyAsync(kmers_out, dout, total * esz, cudaMemcpyDeviceToHost, ctx->stream));
        ctx->last.d2h_bytes = total * esz;
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->last.kernel_ms, ctx->ev[0], ctx->ev[1]);
    for (uint32_t p = 0; p < nparts; ++p) part_counts[p] = pt[p];
    return KMU_OK;
}

// ---- dump of the multiple k-mers (threaded_dump_kmer_counter / dump_in_file_multiple_kmer,
//      src/base/kmercount.rs:139-145, 584-791): `u32 0xcea2bbff | u8 kmer_size | u8 nb_bytes_by_count | u64 nb_kmer`,
//      then per k-mer `kmer.dump()` (4 bytes for the u32 types, kmer32bit.rs / kmer16b32bit.rs:65-68; `u8 k + u64 value`
//      for Kmer64bit, kmer64bit.rs:98-104) followed by the count on nb_bytes_by_count bytes.  Every k-mer seen at
//      least twice appears ONCE with its saturated count; nb_kmer is exact (the reference writes an estimate and, in
//      the threaded writer, u16 counts whatever the header says, SURVEY App. B.8 -- count_bytes = 2 matches it).
int32_t kmu_count_dump_multiple(kmu_ctx* ctx, const kmu_counter* c, const char* path, int32_t count_bytes, uint64_t* nb_dumped) {
    if (!ctx || !c || !path) return fail(KMU_EINVAL, "null argument");
    if (count_bytes != 1 && count_bytes != 2) return fail(KMU_EINVAL, "count_bytes must be 1 or 2");
    uint64_t distinct = 0, unique = 0;
    int32_t rc = kmu_count_stats(ctx, c, &distinct, &unique, nullptr, nullptr);
    if (rc) return rc;
    const uint64_t n = distinct - unique;
    const size_t esz = c->key64 ? 8 : 4;
    std::vector<uint8_t> keys((n + 1) * esz);
    std::vector<uint32_t> counts(n + 1);
    uint64_t got = 0;
    rc = kmu_count_export(ctx, c, 2, keys.data(), counts.data(), n, &got);
    if (rc) return rc;
    FILE* f = std::fopen(path, "wb");
    if (!f) return fail(KMU_EINVAL, "cannot open %s", path);
    const uint32_t magic = 0xcea2bbffu;
    const uint8_t ksz = (uint8_t)c->k, cb = (uint8_t)count_bytes;
    std::fwrite(&magic, 4, 1, f);
    std::fwrite(&ksz, 1, 1, f);
    std::fwrite(&cb, 1, 1, f);
    std::fwrite(&got, 8, 1, f);
    const uint32_t sat = c->max_count() < (count_bytes == 1 ? 0xFFu : 0xFFFFu) ? c->max_count() : (count_bytes == 1 ? 0xFFu : 0xFFFFu);
    std::vector<uint8_t> rec;
    rec.reserve(got * (esz + 3));
    for (uint64_t i = 0; i < got; ++i) {
        if (c->key64) {
            rec.push_back(ksz);
            const uint8_t* p = keys.data() + i * 8;
            rec.insert(rec.end(), p, p + 8);
        } else {
            uint32_t w;
            std::memcpy(&w, keys.data() + i * 4, 4);
            if (c->kmer_type == KMU_KMER32) w |= c->k << 28;  // Kmer32bit.0 carries k in its top four bits (kmer32bit.rs:68-76)
            const uint8_t* p = (const uint8_t*)&w;
            rec.insert(rec.end(), p, p + 4);
        }
        const uint32_t cnt = counts[i] < sat ? counts[i] : sat;
        rec.push_back((uint8_t)cnt);
        if (count_bytes == 2) rec.push_back((uint8_t)(cnt >> 8));
    }
    std::fwrite(rec.data(), 1, rec.size(), f);
    std::fclose(f);
    if (nb_dumped) *nb_dumped = got;
    return KMU_OK;
}

// ---- peer-to-peer exchange: extraction + bucketing + NVLink stores in one kernel -------------------
// Step 1: per-owner counts of this rank's k-mers (kept with the context for step 2).
int32_t kmu_count_partition_counts(kmu_ctx* ctx, const kmu_seqbatch* b, uint32_t k, int32_t kmer_type, int32_t canonical,
                                   uint32_t nparts, uint64_t* part_counts) {
    if (!ctx || !b || !part_counts) return fail(KMU_EINVAL, "null argument");
    if (kmer_type_is_aa(kmer_type) || b->alphabet != 0) return fail(KMU_EINVAL, "partitioning takes DNA k-mers");
    if (!kmer_type_accepts(k, kmer_type))
        return fail(KMU_EINVAL, "kmer size %u is not supported by kmer type %d", k, kmer_type);
    if (nparts < 1 || nparts > 64) return fail(KMU_EINVAL, "nparts must be in 1..64, got %u", nparts);
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScopedDevice sd(ctx->device);
    ctx->last = kmu_times{};
    const bool key64 = kmer_type == KMU_KMER64;
    const int grid = kmu::count_partition_grid(b->packed_bytes, ctx->sm_count);
    CUDA_TRY(ctx->counters.reserve(sizeof(unsigned long long) * ((size_t)nparts * grid + 2 * nparts + 2 * 64 + 64)));
    unsigned long long* block_counts = (unsigned long long*)ctx->counters.p;
    unsigned long long* part_totals = block_counts + (size_t)nparts * grid;
    kmu::SeqView v{b->packed, b->byte_off, b->nbases, b->nseq};
    std::vector<unsigned long long> pt(nparts, 0);
    if (b->nseq && b->packed_bytes) {
        cudaEventRecord(ctx->ev[0], ctx->stream);
        CUDA_TRY(kmu::launch_count_partition_counts(v, b->packed_bytes, k, key64, canonical != 0, nparts, grid, block_counts,
                                                    part_totals, ctx->stream));
        cudaEventRecord(ctx->ev[1], ctx->stream);
        CUDA_TRY(cudaMemcpyAsync(pt.data(), part_totals, sizeof(unsigned long long) * nparts, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        cudaEventElapsedTime(&ctx->last.kernel_ms, ctx->ev[0], ctx->ev[1]);
        ctx->launches += 2;
        ctx->last.launches = 2;
    }
    for (uint32_t p = 0; p < nparts; ++p) part_counts[p] = pt[p];
    ctx->p2p_grid = grid;
    ctx->p2p_nparts = nparts;
    ctx->p2p_batch = b;
    return KMU_OK;
}

// Step 2: the same walk writes bucket p at dests[p] + dest_offsets[p] (elements).  dests[p] is a device pointer of
// this process: local memory or a peer GPU's buffer opened with kmu_ipc_open -- the stores then cross NVLink from
// inside the kernel, no separate collective.  Must follow kmu_count_partition_counts on the same batch / k / nparts.
int32_t kmu_count_partition_scatter(kmu_ctx* ctx, const kmu_seqbatch* b, uint32_t k, int32_t kmer_type, int32_t canonical,
                                    uint32_t nparts, void* const* dests, const uint64_t* dest_offsets) {
    if (!ctx || !b || !dests || !dest_offsets) return fail(KMU_EINVAL, "null argument");
    if (ctx->p2p_batch != b || ctx->p2p_nparts != nparts || ctx->p2p_grid <= 0)
        return fail(KMU_EINVAL, "kmu_count_partition_scatter must follow kmu_count_partition_counts on the same batch");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScopedDevice sd(ctx->device);
    ctx->last = kmu_times{};
    const bool key64 = kmer_type == KMU_KMER64;
    const int grid = ctx->p2p_grid;
    unsigned long long* block_counts = (unsigned long long*)ctx->counters.p;
    unsigned long long* part_base = block_counts + (size_t)nparts * grid + 2 * nparts;  // 64 entries
    void** d_dests = (void**)(part_base + 64);                                          // 64 entries
    std::vector<unsigned long long> base(dest_offsets, dest_offsets + nparts);
    CUDA_TRY(cudaMemcpyAsync(part_base, base.data(), sizeof(unsigned long long) * nparts, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(d_dests, dests, sizeof(void*) * nparts, cudaMemcpyHostToDevice, ctx->stream));
    kmu::SeqView v{b->packed, b->byte_off, b->nbases, b->nseq};
    if (b->nseq && b->packed_bytes) {
        cudaEventRecord(ctx->ev[0], ctx->stream);
        CUDA_TRY(kmu::launch_count_partition_scatter(v, b->packed_bytes, k, key64, canonical != 0, nparts, grid, block_counts,
                                                     part_base, d_dests, ctx->stream));
        cudaEventRecord(ctx->ev[1], ctx->stream);
        ctx->launches += 2;
        ctx->last.launches = 2;
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (b->nseq && b->packed_bytes) cudaEventElapsedTime(&ctx->last.kernel_ms, ctx->ev[0], ctx->ev[1]);
    ctx->p2p_batch = nullptr;
    return KMU_OK;
}

// ---- fused exchange: extraction + (owner, region) bucketing + NVLink stores in ONE kernel and ONE walk -----------------
// Every rank holds a table of the same capacity (kmu_count_create with the same arguments).  The receive buffer of a
// rank is nregions * nowners slabs of slab_cap keys: slab (r, s) holds what sender s found for region r of the table.
// Among several owners the exchange buckets by OWNER ONLY (one slab per sender in every receive buffer: runs of thousands of
// keys per tile, which is what NVLink stores want) and the receiver partitions what it got by region of its table itself
// (kmu_count_insert_slabs): bucketing by (owner, region) in one pass needs owners x 1024 buckets, and past ~1000 buckets
// a tile of 4096 k-mers leaves two keys per run (measured at 2 GPUs: 93 ms of scatter + 116 ms of two-level insertion
// against 46 + 78 on one GPU).  One owner: the local table's regions, i.e. the partition of the single-GPU path.
// The one-pass form (buckets by owner AND region, plain regioned insertion at the receiver) stays available behind
// KMU_COUNT_EXCHANGE_BUCKETS (the largest owners x regions product that still takes it; default 1: always by owner).  Measured
// again with regions of 128 MB at 2 GPUs (2 x 512 buckets): scatter 94 ms + insertion 72 ms against 52 + 109 for the
// owner-only form -- the runs of two or three keys that a 1024-bucket tile leaves are what NVLink stores are worst at.
uint32_t exchange_regions(const kmu_counter* c, uint32_t nowners) {
    const uint32_t ncoarse = region_geometry(c->capacity, c->key64, 1).ncoarse;
    if (nowners <= 1) return ncoarse;
    uint64_t limit = 1;
    if (const char* e = std::getenv("KMU_COUNT_EXCHANGE_BUCKETS")) limit = (uint64_t)std::max(1ll, std::atoll(e));
    return (uint64_t)nowners * ncoarse <= limit ? ncoarse : 1u;
}
int32_t kmu_count_exchange_geometry(const kmu_counter* c, uint32_t nowners, uint32_t* nregions) {
    if (!c || !nregions || nowners < 1 || nowners > 64) return fail(KMU_EINVAL, "bad argument");
    *nregions = exchange_regions(c, nowners);
    return KMU_OK;
}

int32_t kmu_count_exchange_scatter(kmu_ctx* ctx, const kmu_seqbatch* b, const kmu_counter* c, int32_t canonical, uint32_t nowners,
                                   uint32_t self, uint64_t slab_cap, void* const* dests, uint64_t* sent_counts,
                                   int32_t* overflowed) {
    if (!ctx || !b || !c || !dests || !sent_counts || !overflowed) return fail(KMU_EINVAL, "null argument");
    if (b->alphabet != 0) return fail(KMU_EINVAL, "the counter takes DNA sequences");
    if (nowners < 1 || nowners > 64 || self >= nowners) return fail(KMU_EINVAL, "nowners must be in 1..64 and self below it");
    if (slab_cap == 0 || slab_cap >= 0xFFFFFFFFull) return fail(KMU_EINVAL, "slab_cap out of range");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScopedDevice sd(ctx->device);
    ctx->last = kmu_times{};
    const RegionGeom rg = region_geometry(c->capacity, c->key64, 1);
    const uint32_t nreg_x = exchange_regions(c, nowners);
    const uint32_t nb = nowners * nreg_x;
    if (nb > MAX_BUCKETS) return fail(KMU_EINVAL, "too many (owner, region) buckets: %u", nb);
    CUDA_TRY(ctx->counters.reserve(sizeof(unsigned long long) * PART_SCRATCH_WORDS));
    unsigned long long* cursors = (unsigned long long*)ctx->counters.p;
    unsigned long long* flags = cursors + OFF_FLAGS;
    void** d_dests = (void**)(cursors + OFF_DESTS);
    CUDA_TRY(cudaMemcpyAsync(d_dests, dests, sizeof(void*) * nowners, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(cursors, 0, sizeof(unsigned long long) * (MAX_BUCKETS + 64), ctx->stream));
    kmu::PartGeom g{};
    g.nowners = nowners;
    g.nregions = nreg_x;
    g.capmask = c->capacity - 1;
    g.shift = rg.coarse_shift;
    g.nsend = nowners;
    g.self = self;
    g.slab_cap = slab_cap;
    kmu::SeqView v{b->packed, b->byte_off, b->nbases, b->nseq};
    cudaEventRecord(ctx->ev[0], ctx->stream);
    if (b->nseq && b->packed_bytes) {
        CUDA_TRY(kmu::launch_count_part_seqs(v, 0, b->packed_bytes, b->packed_bytes, c->k, c->key64, canonical != 0, g, d_dests, cursors,
                                             flags, ctx->sm_count, ctx->stream));
        ctx->launches += 1;
        ctx->last.launches = 1;
    }
    cudaEventRecord(ctx->ev[1], ctx->stream);
    std::vector<unsigned long long> h(nb + 1);
    CUDA_TRY(cudaMemcpyAsync(h.data(), cursors, sizeof(unsigned long long) * nb, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(h.data() + nb, flags, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->last.kernel_ms, ctx->ev[0], ctx->ev[1]);
    for (uint32_t i = 0; i < nb; ++i) sent_counts[i] = h[i];
    *overflowed = h[nb] ? 1 : 0;
    return KMU_OK;
}

int32_t kmu_count_insert_slabs(kmu_ctx* ctx, kmu_counter* c, const void* slabs, uint64_t slab_cap, uint32_t nsend,
                               const uint64_t* counts) {
    if (!ctx || !c || !slabs || !counts) return fail(KMU_EINVAL, "null argument");
    if (nsend < 1 || nsend > 64) return fail(KMU_EINVAL, "nsend must be in 1..64");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScopedDevice sd(ctx->device);
    ctx->last = kmu_times{};
    const RegionGeom rg = region_geometry(c->capacity, c->key64, 1);
    const size_t n = (size_t)nsend * exchange_regions(c, nsend);
    uint64_t total = 0;
    for (size_t i = 0; i < n; ++i) {
        if (counts[i] > slab_cap) return fail(KMU_EOVERFLOW, "a slab holds %llu keys, capacity %llu", (unsigned long long)counts[i], (unsigned long long)slab_cap);
        total += counts[i];
    }
    CUDA_TRY(ctx->counters.reserve(sizeof(unsigned long long) * PART_SCRATCH_WORDS));
    CUDA_TRY(ctx->misc.reserve(sizeof(unsigned long long) * n));
    CUDA_TRY(cudaMemcpyAsync(ctx->misc.p, counts, sizeof(unsigned long long) * n, cudaMemcpyHostToDevice, ctx->stream));
    cudaEventRecord(ctx->ev[0], ctx->stream);
    uint64_t nl = 0;
    int32_t rc0 = KMU_OK;
    if (nsend == 1 || exchange_regions(c, nsend) > 1)
        rc0 = insert_level1_slabs(ctx, c, slabs, slab_cap, nsend, (const unsigned long long*)ctx->misc.p, rg, nullptr, &nl);
    else rc0 = insert_segments(ctx, c, slabs, slab_cap, nsend, counts, (const unsigned long long*)ctx->misc.p, total, &nl);
    if (rc0) return rc0;
    cudaEventRecord(ctx->ev[1], ctx->stream);
    ctx->launches += nl;
    ctx->last.launches = nl;
    int32_t rc = check_overflow(ctx, c);
    cudaEventElapsedTime(&ctx->last.kernel_ms, ctx->ev[0], ctx->ev[1]);
    if (rc) return rc;
    c->inserted += total;
    return KMU_OK;
}

// receive buffers shared between the processes of one box (CUDA IPC): export a buffer of this GPU ...
int32_t kmu_ipc_alloc(kmu_ctx* ctx, uint64_t bytes, void** dev_ptr, uint8_t handle[64]) {
    if (!ctx || !dev_ptr || !handle) return fail(KMU_EINVAL, "null argument");
    ScopedDevice sd(ctx->device);
    void* p = nullptr;
    CUDA_TRY(cudaMalloc(&p, bytes ? bytes : 256));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return fail(KMU_ECUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    std::memcpy(handle, &h, 64);
    *dev_ptr = p;
    return KMU_OK;
}
int32_t kmu_ipc_free(kmu_ctx* ctx, void* dev_ptr) {
    if (!ctx) return fail(KMU_EINVAL, "null argument");
    ScopedDevice sd(ctx->device);
    if (dev_ptr) CUDA_TRY(cudaFree(dev_ptr));
    return KMU_OK;
}
// ... and map a peer's buffer into this process (the returned pointer is valid in this context's kernels)
int32_t kmu_ipc_open(kmu_ctx* ctx, const uint8_t handle[64], void** peer_ptr) {
    if (!ctx || !handle || !peer_ptr) return fail(KMU_EINVAL, "null argument");
    ScopedDevice sd(ctx->device);
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    CUDA_TRY(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return KMU_OK;
}
int32_t kmu_ipc_close(kmu_ctx* ctx, void* peer_ptr) {
    if (!ctx) return fail(KMU_EINVAL, "null argument");
    ScopedDevice sd(ctx->device);
    if (peer_ptr) CUDA_TRY(cudaIpcCloseMemHandle(peer_ptr));
    return KMU_OK;
}

// ---- ProbMinHash3a over a weighted set --------------------------------------------------------
static void fill_exp01(kmu::Exp01Params& e, uint32_t m) {
    // ProbMinHash3a::new : lambda = ln(m / (m-1)); ExpRestricted01::new
    const double lambda = std::log((double)m / (double)(m - 1));
    e.lambda = lambda;
    e.c1 = std::expm1(lambda) / lambda;
    e.c2 = std::log(2.0 / (1.0 + std::exp(-lambda))) / lambda;
    e.c3 = (1.0 - std::exp(-lambda)) / lambda;
}

// runs the item kernel with growing bounds until every slot ended below the bound
// wmax: a lower bound of the largest weight in the set (1 for multiplicity tables).  The heaviest item alone fills every
// slot below m (ln m + 40) / wmax with probability 1 - e^-40: the bound never has to grow beyond that.
static int32_t run_pmh3a_items(kmu_ctx* ctx, kmu::Pmh3aItemsParams P, bool key64, int src, uint64_t distinct, void* d_sig,
                               uint64_t* launches, double wmax = 1.0) {
    cudaStream_t st = ctx->stream;
    const uint32_t m = P.m;
    CUDA_TRY(ctx->items_slots.reserve(sizeof(kmu::Slot) * m + 64));
    P.global_slots = (kmu::Slot*)ctx->items_slots.p;
    unsigned long long* d_max = (unsigned long long*)((uint8_t*)ctx->items_slots.p + sizeof(kmu::Slot) * m);
    const size_t smem = sizeof(kmu::Slot) * (size_t)m;
    P.slots_in_smem = smem + kmu::PMH3A_ITEMS_QUEUE_BYTES <= SMEM_BUDGET ? 1 : 0;
    P.slot_thresh = (uint32_t)(0x100000000ULL % m);
    fill_exp01(P.e, m);
    const uint64_t work = (P.n + 511) / 512;
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(work, (uint64_t)ctx->sm_count));
    const double bound_max = (double)m * (std::log((double)m) + 40.0) / wmax;
    double bound = distinct ? std::min(bound_max, (double)m / (double)distinct * std::log((double)m / 1e-4)) : 1.0;
    for (int attempt = 0; attempt < 200; ++attempt) {
        P.bound = bound;
        CUDA_TRY(kmu::launch_pmh3a_items_init(P.global_slots, m, st));
        CUDA_TRY(cudaMemsetAsync(d_max, 0, sizeof(unsigned long long), st));
        if (P.n) CUDA_TRY(kmu::launch_pmh3a_items(P, key64, src, grid, P.slots_in_smem ? smem : 0, st));
        CUDA_TRY(kmu::launch_pmh3a_items_finish(P.global_slots, m, key64, d_sig, d_max, st));
        *launches += 3;
        unsigned long long mx = 0;
        CUDA_TRY(cudaMemcpyAsync(&mx, d_max, sizeof(mx), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        double top;
        std::memcpy(&top, &mx, 8);
        if (distinct == 0 || top < bound) return KMU_OK;  // every slot filled below the bound: nothing pruned could win
        if (bound >= bound_max) break;
        bound = std::min(bound_max, bound * 4.0);
    }
    return fail(KMU_ECUDA, "ProbMinHash3a item sketch did not converge");
}

int32_t kmu_pmh3a_weighted(kmu_ctx* ctx, const void* keys, const double* weights, uint64_t n, int32_t key_bytes, uint32_t m,
                           void* sig) {
    if (!ctx || !sig || (n && (!keys || !weights))) return fail(KMU_EINVAL, "null argument");
    if (key_bytes != 4 && key_bytes != 8) return fail(KMU_EINVAL, "key_bytes must be 4 or 8");
    if (m < 2) return fail(KMU_EINVAL, "ProbMinHash3a needs at least 2 hash values (m = %u)", m);
    for (uint64_t i = 0; i < n; ++i)
        if (!(weights[i] > 0.0)) return fail(KMU_EINVAL, "weight %llu is not positive", (unsigned long long)i);
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScopedDevice sd(ctx->device);
    ctx->last = kmu_times{};
    cudaStream_t st = ctx->stream;
    const size_t kb = n * (size_t)key_bytes, wb = n * sizeof(double), sb = (size_t)m * key_bytes;
    CUDA_TRY(ctx->misc.reserve(align_up(kb, 16) + align_up(wb, 16) + sb + 64));
    uint8_t* d_keys = (uint8_t*)ctx->misc.p;
    double* d_w = (double*)(d_keys + align_up(kb, 16));
    void* d_sig = (uint8_t*)d_w + align_up(wb, 16);
    if (n) {
        CUDA_TRY(cudaMemcpyAsync(d_keys, keys, kb, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_w, weights, wb, cudaMemcpyHostToDevice, st));
    }
    kmu::Pmh3aItemsParams P{};
    P.keys = d_keys;
    P.weights = d_w;
    P.n = n;
    P.m = m;
    uint64_t launches = 0;
    cudaEventRecord(ctx->ev[0], st);
    double wmax = 0.0;
    for (uint64_t i = 0; i < n; ++i) wmax = std::max(wmax, weights[i]);
    int32_t rc = run_pmh3a_items(ctx, P, key_bytes == 8, 0, n, d_sig, &launches, n ? wmax : 1.0);
    cudaEventRecord(ctx->ev[1], st);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(sig, d_sig, sb, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&ctx->last.kernel_ms, ctx->ev[0], ctx->ev[1]);
    ctx->launches += launches;
    ctx->last.launches = launches;
    return KMU_OK;
}

int32_t kmu_sketch_pmh3a_whole(kmu_ctx* ctx, const kmu_seqbatch* b, uint32_t k, int32_t kmer_type, int32_t hash_kind,
                               uint32_t m, void* sig, int32_t sig_on_device) {
    if (!ctx || !b || !sig) return fail(KMU_EINVAL, "null argument");
    if (int32_t a = kmu_check_kmer_args(b, k, kmer_type, hash_kind)) return a;
    if (kmer_type_is_aa(kmer_type)) return fail(KMU_EINVAL, "whole-file ProbMinHash3a takes DNA sequences");
    if (m < 2) return fail(KMU_EINVAL, "ProbMinHash3a needs at least 2 hash values (m = %u)", m);
    const uint64_t total = b->kmer_count(k);
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScopedDevice sd(ctx->device);
    ctx->last = kmu_times{};
    cudaStream_t st = ctx->stream;
    const bool key64 = kmer_type == KMU_KMER64;
    const size_t vsz = key64 ? 8 : 4;
    void* d_sig = sig;
    if (!sig_on_device) {
        CUDA_TRY(ctx->sig_dev.reserve((size_t)m * vsz));
        d_sig = ctx->sig_dev.p;
    }
    // multiplicity table over all sequences (the FnvHashMap of setsketchert.rs:171-189); the hash closures are
    // injective on the pre-key, so counting pre-keys counts hashed keys
    uint64_t want = std::max<uint64_t>(1024, total * 2);
    if (2 * k < 40) want = std::min<uint64_t>(want, std::max<uint64_t>(1024, 2ull << (2 * k)));
    uint64_t cap = 1024;
    while (cap < want) cap <<= 1;
    const size_t slot_bytes = key64 ? 16 : 8;
    cudaError_t me = ctx->whole_table.reserve(cap * slot_bytes + sizeof(unsigned long long) * AUX_WORDS);
    if (me != cudaSuccess)
        return fail(KMU_ENOMEM, "multiplicity table of %llu slots: %s", (unsigned long long)cap, cudaGetErrorString(me));
    kmu::CountTable t;
    t.slots = ctx->whole_table.p;
    t.capmask = cap - 1;
    unsigned long long* aux = (unsigned long long*)((uint8_t*)ctx->whole_table.p + cap * slot_bytes);
    t.special = aux;
    t.overflow = aux + 1;
    unsigned long long* stats = aux + 8;
    uint64_t launches = 0;
    cudaEventRecord(ctx->ev[0], st);
    CUDA_TRY(cudaMemsetAsync(aux, 0, sizeof(unsigned long long) * AUX_WORDS, st));
    CUDA_TRY(kmu::launch_count_init(t, key64, ctx->sm_count, st));
    kmu::SeqView v{b->packed, b->byte_off, b->nbases, b->nseq};
    CUDA_TRY(kmu::launch_count_insert_seqs(v, b->packed_bytes, k, key64, kmu::hash_kind_is_canonical_host(hash_kind), t,
                                           ctx->sm_count, st));
    CUDA_TRY(kmu::launch_count_stats(t, key64, stats, ctx->sm_count, st));
    launches += 3;
    std::vector<unsigned long long> h(3 + 256);
    CUDA_TRY(cudaMemcpyAsync(h.data(), stats, sizeof(unsigned long long) * (3 + 256), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    uint64_t distinct = 0;
    for (int i = 1; i < 256; ++i) distinct += h[3 + i];
    kmu::Pmh3aItemsParams P{};
    P.table = t.slots;
    P.special = t.special;
    P.n = cap;
    P.k = k;
    P.kmer_type = kmer_type;
    P.hash_kind = hash_kind;
    P.m = m;
    int32_t rc = run_pmh3a_items(ctx, P, key64, key64 ? 2 : 1, distinct, d_sig, &launches);
    cudaEventRecord(ctx->ev[1], st);
    if (rc) return rc;
    if (!sig_on_device) {
        CUDA_TRY(cudaMemcpyAsync(sig, d_sig, (size_t)m * vsz, cudaMemcpyDeviceToHost, st));
        ctx->last.d2h_bytes = (size_t)m * vsz;
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&ctx->last.kernel_ms, ctx->ev[0], ctx->ev[1]);
    ctx->launches += launches;
    ctx->last.launches = launches;
    return KMU_OK;
}

// Partial ProbMinHash3a registers of a counting table (the multi-GPU whole-file sketch: every rank owns the complete
// counts of the keys that hash to it, sketches them, and the ranks merge their registers with an allreduce-min on h
// and a key select).  slots: m records {h bits (u64; F64 max when the slot saw no point below `bound`),
// key (u64)}; items are cut at `bound` as in kmu_sketch_pmh3a_whole -- the caller checks the MERGED maximum against it.
int32_t kmu_pmh3a_counter_slots(kmu_ctx* ctx, const kmu_counter* c, int32_t hash_kind, uint32_t m, double bound, void* slots,
                                int32_t slots_on_device) {
    if (!ctx || !c || !slots) return fail(KMU_EINVAL, "null argument");
    if (m < 2) return fail(KMU_EINVAL, "ProbMinHash3a needs at least 2 hash values (m = %u)", m);
    if (hash_kind < KMU_HASH_IDENTITY_RAW || hash_kind > KMU_HASH_INVHASH) return fail(KMU_EINVAL, "unknown hash kind %d", hash_kind);
    if (!(bound > 0.0)) return fail(KMU_EINVAL, "the item bound must be positive");
    // an item emits (bound x count) points: one key alone fills all m slots below m (ln m + 40) with probability
    // 1 - e^-40, so nothing larger is ever needed -- and a much larger bound would keep the GPU busy for hours
    if (bound > (double)m * (std::log((double)m) + 40.0))
        return fail(KMU_EINVAL, "item bound %g is beyond m (ln m + 40) = %g", bound, (double)m * (std::log((double)m) + 40.0));
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScopedDevice sd(ctx->device);
    cudaStream_t st = ctx->stream;
    CUDA_TRY(ctx->items_slots.reserve(sizeof(kmu::Slot) * m + 64));
    kmu::Pmh3aItemsParams P{};
    const kmu::CountTable t = c->view();
    P.table = t.slots;
    P.special = t.special;
    P.n = c->capacity;
    P.k = c->k;
    P.kmer_type = c->kmer_type;
    P.hash_kind = hash_kind;
    P.m = m;
    P.global_slots = (kmu::Slot*)ctx->items_slots.p;
    const size_t smem = sizeof(kmu::Slot) * (size_t)m;
    P.slots_in_smem = smem + kmu::PMH3A_ITEMS_QUEUE_BYTES <= SMEM_BUDGET ? 1 : 0;
    P.slot_thresh = (uint32_t)(0x100000000ULL % m);
    fill_exp01(P.e, m);
    P.bound = bound;
    const uint64_t work = (P.n + 511) / 512;
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(work, (uint64_t)ctx->sm_count));
    CUDA_TRY(kmu::launch_pmh3a_items_init(P.global_slots, m, st));
    CUDA_TRY(kmu::launch_pmh3a_items(P, c->key64, c->key64 ? 2 : 1, grid, P.slots_in_smem ? smem : 0, st));
    CUDA_TRY(cudaMemcpyAsync(slots, P.global_slots, sizeof(kmu::Slot) * m, slots_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    ctx->launches += 2;
    ctx->last.launches = 2;
    return KMU_OK;
}

// One whole-file signature per group of consecutive sequences (a genome = its contigs): what gsearch does with
// ProbHash3aSketch::sketch_compressedkmer_seqs (src/sketching/setsketchert.rs:160-202) genome after genome.  The groups
// run back to back on the stream with no host round trip in between: the multiplicity table of a group is reused
// for the next one (it stays L2 resident: 2 x k-mers x 8 bytes), the item bound comes from the k-mer count instead of
// the distinct count (1.5 m ln(m / 1e-4) / k-mers: exact whenever at least two thirds of the k-mers are distinct),
// and all verifications are read back once at the end; a group that failed its check is redone by
// kmu_sketch_pmh3a_whole.  sig: ngroups rows of m values.
int32_t kmu_sketch_pmh3a_groups(kmu_ctx* ctx, const kmu_seqbatch* b, const uint64_t* group_sizes, uint64_t ngroups, uint32_t k,
                                int32_t kmer_type, int32_t hash_kind, uint32_t m, void* sig, int32_t sig_on_device) {
    if (!ctx || !b || (ngroups && (!group_sizes || !sig))) return fail(KMU_EINVAL, "null argument");
    if (int32_t a = kmu_check_kmer_args(b, k, kmer_type, hash_kind)) return a;
    if (kmer_type_is_aa(kmer_type)) return fail(KMU_EINVAL, "whole-file ProbMinHash3a takes DNA sequences");
    if (m < 2) return fail(KMU_EINVAL, "ProbMinHash3a needs at least 2 hash values (m = %u)", m);
    uint64_t covered = 0;
    for (uint64_t g = 0; g < ngroups; ++g) covered += group_sizes[g];
    if (covered != b->nseq) return fail(KMU_EINVAL, "the groups cover %llu sequences, the batch has %llu", (unsigned long long)covered,
                                        (unsigned long long)b->nseq);
    if (ngroups == 0) return KMU_OK;
    const bool key64 = kmer_type == KMU_KMER64;
    const size_t vsz = key64 ? 8 : 4, slot_bytes = key64 ? 16 : 8;
    // per group: first sequence, k-mers, bytes; byte offsets rebased to the group's first sequence
    std::vector<uint64_t> first(ngroups + 1, 0), nk(ngroups, 0), rebased(b->nseq);
    uint64_t cap_max = 1024;
    for (uint64_t g = 0; g < ngroups; ++g) {
        first[g + 1] = first[g] + group_sizes[g];
        const uint64_t base = group_sizes[g] ? b->h_byte_off[first[g]] : 0;
        for (uint64_t i = first[g]; i < first[g + 1]; ++i) {
            rebased[i] = b->h_byte_off[i] - base;
            nk[g] += b->h_nbases[i] >= k ? b->h_nbases[i] - k + 1 : 0;
        }
        uint64_t want = std::max<uint64_t>(1024, nk[g] + nk[g] / 2 + nk[g] / 16);  // load <= 0.64: a 5 Mb genome's table (64 MB) stays in L2
        if (2 * k < 40) want = std::min<uint64_t>(want, std::max<uint64_t>(1024, 2ull << (2 * k)));
        uint64_t cap = 1024;
        while (cap < want) cap <<= 1;
        cap_max = std::max(cap_max, cap);
    }
    std::vector<unsigned long long> tops(ngroups, 0), ovfs(ngroups, 0);
    std::vector<double> bounds(ngroups, 0.0);
    // prefiltered insertion (pmh3a_prefilter_kernel, kmu_count.cu) for groups whose bound is selective: the table holds
    // the ~(bound + 15 %) of the k-mers that can matter instead of all of them; bits: 8 per k-mer in each of the two bitmaps
    const bool prefilter_on = std::getenv("KMU_GROUP_NO_PREFILTER") == nullptr;
    auto prefiltered = [&](uint64_t g, uint64_t* cap_f, uint64_t* bits) {
        if (!prefilter_on || nk[g] < (1ull << 20)) return false;
        const double bound = 1.5 * (double)m / (double)nk[g] * std::log((double)m / 1e-4);
        if (!(bound < 0.25)) return false;
        uint64_t want = (uint64_t)(1.5 * (double)nk[g] * (bound + 0.15)) + 1024, cap = 1024, nb = 1ull << 16;
        while (cap < want) cap <<= 1;
        while (nb < 8 * nk[g]) nb <<= 1;
        *cap_f = cap;
        *bits = nb;
        return true;
    };
    uint64_t seen_bytes_max = 0;
    for (uint64_t g = 0; g < ngroups; ++g) {
        uint64_t cf = 0, nb = 0;
        if (prefiltered(g, &cf, &nb)) seen_bytes_max = std::max<uint64_t>(seen_bytes_max, 2 * nb / 8);
    }
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        ScopedDevice sd(ctx->device);
        ctx->last = kmu_times{};
        cudaStream_t st = ctx->stream;
        // the genomes are independent: NS of them are in flight at once, each on its own stream with its own table and
        // slots (the launches of one genome are short and leave SMs idle; the tables of the genomes in flight must fit L2)
        // measured on 148 genomes of 5 Mb (ms per call): unfiltered insertion (64 MB tables) 1 stream 50.5, 2: 37.6, 3: 38.6, 4: 43.9;
        // prefiltered insertion (16 MB tables + 16 MB filters) 1: 33.5, 2: 23.1, 3: 22.2-22.5, 4: 23.3
        int ns_want = 3;
        if (const char* e = std::getenv("KMU_GROUP_STREAMS")) ns_want = std::max(1, std::min(kmu_ctx::GROUP_STREAMS, std::atoi(e)));  // measurements
        const int NS = (int)std::min<uint64_t>((uint64_t)ns_want, ngroups);
        cudaError_t me = cudaSuccess;
        for (int q = 0; q < NS && me == cudaSuccess; ++q) {
            if (!ctx->group_stream[q]) me = cudaStreamCreateWithFlags(&ctx->group_stream[q], cudaStreamNonBlocking);
            if (me == cudaSuccess && !ctx->group_ev[q]) me = cudaEventCreateWithFlags(&ctx->group_ev[q], cudaEventDisableTiming);
            if (me == cudaSuccess) me = ctx->group_table[q].reserve(cap_max * slot_bytes + sizeof(unsigned long long) * AUX_WORDS);
            if (me == cudaSuccess) me = ctx->group_slots[q].reserve(sizeof(kmu::Slot) * m + 64);
            if (me == cudaSuccess && seen_bytes_max) me = ctx->group_seen[q].reserve(seen_bytes_max);
        }
        if (me == cudaSuccess && !ctx->group_ev[kmu_ctx::GROUP_STREAMS])
            me = cudaEventCreateWithFlags(&ctx->group_ev[kmu_ctx::GROUP_STREAMS], cudaEventDisableTiming);
        if (me == cudaSuccess) me = ctx->misc.reserve(sizeof(uint64_t) * (b->nseq + 1) + 2 * sizeof(unsigned long long) * ngroups + 64);
        if (me == cudaSuccess && !sig_on_device) me = ctx->sig_dev.reserve((size_t)ngroups * m * vsz);
        if (me != cudaSuccess) return fail(KMU_ENOMEM, "group sketch buffers: %s", cudaGetErrorString(me));
        uint64_t* d_rebased = (uint64_t*)ctx->misc.p;
        unsigned long long* d_tops = (unsigned long long*)(d_rebased + b->nseq + 1);
        unsigned long long* d_ovfs = d_tops + ngroups;  // a group whose filtered table overflowed (many repeated k-mers) is redone
        uint8_t* d_sig = sig_on_device ? (uint8_t*)sig : (uint8_t*)ctx->sig_dev.p;
        CUDA_TRY(cudaMemcpyAsync(d_rebased, rebased.data(), sizeof(uint64_t) * b->nseq, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemsetAsync(d_tops, 0, 2 * sizeof(unsigned long long) * ngroups, st));
        kmu::Pmh3aItemsParams P{};
        P.k = k;
        P.kmer_type = kmer_type;
        P.hash_kind = hash_kind;
        P.m = m;
        const size_t smem = sizeof(kmu::Slot) * (size_t)m;
        P.slots_in_smem = smem + kmu::PMH3A_ITEMS_QUEUE_BYTES <= SMEM_BUDGET ? 1 : 0;
        P.slot_thresh = (uint32_t)(0x100000000ULL % m);
        fill_exp01(P.e, m);
        uint64_t launches = 0;
        cudaEventRecord(ctx->ev[0], st);
        CUDA_TRY(cudaEventRecord(ctx->group_ev[kmu_ctx::GROUP_STREAMS], st));
        for (int q = 0; q < NS; ++q) CUDA_TRY(cudaStreamWaitEvent(ctx->group_stream[q], ctx->group_ev[kmu_ctx::GROUP_STREAMS], 0));
        for (uint64_t g = 0; g < ngroups; ++g) {
            const int q = (int)(g % (uint64_t)NS);
            cudaStream_t gs = ctx->group_stream[q];
            if (nk[g] == 0) {  // no k-mer: the all-zero signature of an empty sketch
                CUDA_TRY(cudaMemsetAsync(d_sig + g * (size_t)m * vsz, 0, (size_t)m * vsz, gs));
                continue;
            }
            uint64_t want = std::max<uint64_t>(1024, nk[g] + nk[g] / 2 + nk[g] / 16);  // load <= 0.64
            if (2 * k < 40) want = std::min<uint64_t>(want, std::max<uint64_t>(1024, 2ull << (2 * k)));
            uint64_t cap = 1024;
            while (cap < want) cap <<= 1;
            uint64_t cap_f = 0, seen_bits = 0;
            const bool filtered = prefiltered(g, &cap_f, &seen_bits) && cap_f < cap;
            if (filtered) cap = cap_f;
            kmu::CountTable t;
            t.slots = ctx->group_table[q].p;
            t.capmask = cap - 1;
            unsigned long long* aux = (unsigned long long*)((uint8_t*)ctx->group_table[q].p + cap_max * slot_bytes);
            t.special = aux;
            t.overflow = aux + 1;
            CUDA_TRY(cudaMemsetAsync(aux, 0, sizeof(unsigned long long) * 8, gs));
            CUDA_TRY(kmu::launch_count_init(t, key64, ctx->sm_count, gs));
            const uint64_t s0 = first[g], ns = group_sizes[g];
            const uint64_t Llast = b->h_nbases[s0 + ns - 1];
            const uint64_t bytes = rebased[s0 + ns - 1] + align_up((Llast + 3) / 4, SEQ_ALIGN);
            kmu::SeqView v{b->packed + b->h_byte_off[s0], d_rebased + s0, b->nbases + s0, ns};
            bounds[g] = 1.5 * (double)m / (double)nk[g] * std::log((double)m / 1e-4);
            if (filtered) {
                CUDA_TRY(cudaMemsetAsync(ctx->group_seen[q].p, 0, 2 * seen_bits / 8, gs));
                CUDA_TRY(kmu::launch_pmh3a_prefilter(v, bytes, k, key64, kmu::hash_kind_is_canonical_host(hash_kind), kmer_type, hash_kind,
                                                     bounds[g], P.e.c1, t, (uint32_t*)ctx->group_seen[q].p, seen_bits - 1, ctx->sm_count, gs));
                CUDA_TRY(cudaMemcpyAsync(d_ovfs + g, t.overflow, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, gs));
                launches += 1;
            } else {
                CUDA_TRY(kmu::launch_count_insert_seqs(v, bytes, k, key64, kmu::hash_kind_is_canonical_host(hash_kind), t, ctx->sm_count, gs));
            }
            P.global_slots = (kmu::Slot*)ctx->group_slots[q].p;
            P.table = t.slots;
            P.special = t.special;
            P.n = cap;
            P.bound = bounds[g];
            const uint64_t work = (P.n + 511) / 512;
            const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(work, (uint64_t)ctx->sm_count));
            CUDA_TRY(kmu::launch_pmh3a_items_init(P.global_slots, m, gs));
            CUDA_TRY(kmu::launch_pmh3a_items(P, key64, key64 ? 2 : 1, grid, P.slots_in_smem ? smem : 0, gs));
            CUDA_TRY(kmu::launch_pmh3a_items_finish(P.global_slots, m, key64, d_sig + g * (size_t)m * vsz, d_tops + g, gs));
            launches += 5;
        }
        for (int q = 0; q < NS; ++q) {
            CUDA_TRY(cudaEventRecord(ctx->group_ev[q], ctx->group_stream[q]));
            CUDA_TRY(cudaStreamWaitEvent(st, ctx->group_ev[q], 0));
        }
        cudaEventRecord(ctx->ev[1], st);
        CUDA_TRY(cudaMemcpyAsync(tops.data(), d_tops, sizeof(unsigned long long) * ngroups, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(ovfs.data(), d_ovfs, sizeof(unsigned long long) * ngroups, cudaMemcpyDeviceToHost, st));
        if (!sig_on_device) {
            CUDA_TRY(cudaMemcpyAsync(sig, d_sig, (size_t)ngroups * m * vsz, cudaMemcpyDeviceToHost, st));
            ctx->last.d2h_bytes = (size_t)ngroups * m * vsz;
        }
        CUDA_TRY(cudaStreamSynchronize(st));
        cudaEventElapsedTime(&ctx->last.kernel_ms, ctx->ev[0], ctx->ev[1]);
        ctx->launches += launches;
        ctx->last.launches = launches;
    }
    // groups whose largest slot value did not stay below their bound: the full procedure (distinct count, growing bound)
    for (uint64_t g = 0; g < ngroups; ++g) {
        if (nk[g] == 0) continue;
        double top;
        std::memcpy(&top, &tops[g], 8);
        if (top < bounds[g] && !ovfs[g]) continue;
        kmu_seqbatch* view = nullptr;
        int32_t rc = kmu_seqbatch_view(b, first[g], group_sizes[g], &view);
        if (rc) return rc;
        rc = kmu_sketch_pmh3a_whole(ctx, view, k, kmer_type, hash_kind, m, (uint8_t*)sig + g * (size_t)m * vsz, sig_on_device);
        kmu_seqbatch_destroy(view);
        if (rc) return rc;
    }
    return KMU_OK;
}

}  // extern "C"
