// kmu_capi_sketch.cu -- C ABI of the SuperMinHash and SetSketch sketchers
// (include/kmerutils_b200.h "SuperMinHash", "SetSketch").
#include <algorithm>
#include <cmath>
#include <vector>

#include "kmu_host.h"

namespace {

struct TeamGeometry {
    uint32_t team_warps, teams_per_cta, team_smem_bytes;
    int block;
    size_t smem;
};

// one team (1..32 warps) per sequence: about 1024 k-mers per warp, as many teams per CTA as fit
TeamGeometry team_geometry(uint64_t nk_max, size_t team_bytes, size_t cta_fixed_bytes = 0) {
    TeamGeometry g{};
    uint32_t tw = 1;
    while (tw < 32 && (uint64_t)tw * 1024 < nk_max) tw <<= 1;
    team_bytes = align_up(team_bytes, 16);
    for (;;) {
        uint32_t fit = (uint32_t)std::max<size_t>(1, (SMEM_BUDGET - cta_fixed_bytes) / team_bytes);
        uint32_t max_teams = 32 / tw;
        if (tw > 1 && max_teams > 15) max_teams = 15;  // named barriers 1..15
        if (tw < 32 && fit * tw < 16) {  // shared memory leaves too few warps: widen the teams
            tw <<= 1;
            continue;
        }
        g.team_warps = tw;
        g.teams_per_cta = std::min(max_teams, fit);
        g.team_smem_bytes = (uint32_t)team_bytes;
        g.block = (int)(tw * 32 * g.teams_per_cta);
        g.smem = cta_fixed_bytes + team_bytes * g.teams_per_cta;
        return g;
    }
}

}  // namespace

// per-sequence SuperMinHash into device memory (the context's mutex is held by the caller)
static int32_t smh_per_sequence_device(kmu_ctx* ctx, const kmu_seqbatch* b, uint32_t k, int32_t kmer_type, int32_t hash_kind,
                                       uint32_t m, int32_t key_hasher, int32_t sig_bytes, void* d_sig, uint64_t* launches) {
    cudaStream_t st = ctx->stream;
    const bool key64 = kmer_type_is_u64(kmer_type), f64 = sig_bytes == 8;
    const size_t team_bytes = align_up((size_t)m * sig_bytes, 16) + 32;
    int32_t rc = kmu_ensure_order(ctx, b, k, launches);
    if (rc) return rc;
    std::vector<OctaveClass> classes = kmu_octave_classes(b);
    if (classes.size() > 120) return fail(KMU_EINVAL, "too many launch classes");
    CUDA_TRY(ctx->counters.reserve(sizeof(unsigned long long) * 256));
    CUDA_TRY(cudaMemsetAsync(ctx->counters.p, 0, sizeof(unsigned long long) * 256, st));
    unsigned long long* d_work = (unsigned long long*)ctx->counters.p;  // [0..127] work counters, [128] slow count
    CUDA_TRY(ctx->overflow.reserve(sizeof(uint32_t) * (b->nseq + 1)));
    kmu::SmhParams P{};
    P.packed = b->packed;
    P.byte_off = b->byte_off;
    P.nbases = b->nbases;
    P.order = (const uint32_t*)b->order_cache.order.p;
    P.k = k;
    P.kmer_type = kmer_type;
    P.hash_kind = hash_kind;
    P.m = m;
    P.hasher = key_hasher;
    P.sig = d_sig;
    P.ln_term = std::log(1e4 * (double)m);
    P.slow_count = d_work + 128;
    P.slow_list = (uint32_t*)ctx->overflow.p;
    // small key spaces: point 0 of every possible pre-key, built once per parameter set and kept in the context
    P.memo = nullptr;
    if (!key64 && !kmer_type_is_aa(kmer_type) && k <= 8) {
        const uint32_t nkeys = 1u << (2 * k);
        if (!(ctx->smh_memo.p && ctx->smh_memo_k == k && ctx->smh_memo_m == m && ctx->smh_memo_type == kmer_type &&
              ctx->smh_memo_hash == hash_kind && ctx->smh_memo_hasher == key_hasher && ctx->smh_memo_bytes == sig_bytes)) {
            CUDA_TRY(ctx->smh_memo.reserve((size_t)nkeys * 16));
            CUDA_TRY(kmu::launch_smh_memo(P, f64, ctx->smh_memo.p, nkeys, st));
            ++*launches;
            ctx->smh_memo_k = k;
            ctx->smh_memo_m = m;
            ctx->smh_memo_type = kmer_type;
            ctx->smh_memo_hash = hash_kind;
            ctx->smh_memo_hasher = key_hasher;
            ctx->smh_memo_bytes = sig_bytes;
        }
        P.memo = ctx->smh_memo.p;
    }
    // the warps' key queues of the value-cut path follow the teams in shared memory, when there is room
    const size_t smh_queue = team_bytes + kmu::SMH_QUEUE_BYTES <= SMEM_BUDGET ? kmu::SMH_QUEUE_BYTES : 0;
    P.value_cut = smh_queue ? 1u : 0u;
    // merge neighbouring octave classes that get the same team geometry: one launch each
    struct Launch { uint64_t first, count; TeamGeometry g; };
    std::vector<Launch> ls;
    for (const OctaveClass& c : classes) {
        TeamGeometry g = team_geometry(c.nk_max, team_bytes, smh_queue);
        if (!ls.empty() && ls.back().g.team_warps == g.team_warps && ls.back().g.teams_per_cta == g.teams_per_cta &&
            ls.back().first + ls.back().count == c.first) {
            ls.back().count += c.count;
        } else {
            ls.push_back({c.first, c.count, g});
        }
    }
    int ci = 0;
    for (const Launch& l : ls) {
        kmu::SmhParams Q = P;
        Q.first = l.first;
        Q.count = l.count;
        Q.work_counter = d_work + ci++;
        Q.team_warps = l.g.team_warps;
        Q.team_smem_bytes = l.g.team_smem_bytes;
        const uint64_t ctas_needed = (l.count + l.g.teams_per_cta - 1) / l.g.teams_per_cta;
        uint32_t per_sm = (uint32_t)std::min<uint64_t>(2048 / l.g.block, SMEM_BUDGET / std::max<size_t>(l.g.smem, 1));
        per_sm = std::max(1u, std::min(per_sm, 8u));
        const int grid = (int)std::min<uint64_t>(ctas_needed, (uint64_t)ctx->sm_count * per_sm);
        CUDA_TRY(kmu::launch_smh_fast(Q, key64, f64, grid, l.g.block, l.g.smem, st));
        ++*launches;
    }
    // short sequences (relative to m) and failed speculations: exact path
    unsigned long long nslow = 0;
    CUDA_TRY(cudaMemcpyAsync(&nslow, P.slow_count, sizeof(nslow), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (nslow) {
        kmu::SmhParams Q = P;
        Q.order = P.slow_list;
        Q.first = 0;
        Q.count = nslow;
        Q.work_counter = d_work + ci++;
        const uint64_t per_warp = align_up((uint64_t)m * sig_bytes, 16) + 32ull * 2 * m * sizeof(uint32_t);
        uint64_t warps = std::min<uint64_t>(nslow, (uint64_t)ctx->sm_count * 8);
        const uint64_t budget = 8ull << 30;
        if (warps * per_warp > budget) warps = std::max<uint64_t>(1, budget / per_warp);
        CUDA_TRY(ctx->table_scratch.reserve(warps * per_warp));
        ctx->table_scratch_clean = false;  // shared with the ProbMinHash3a tables, which expect zeros
        CUDA_TRY(cudaMemsetAsync(ctx->table_scratch.p, 0, warps * per_warp, st));
        Q.scratch = (uint8_t*)ctx->table_scratch.p;
        Q.scratch_per_warp = per_warp;
        CUDA_TRY(kmu::launch_smh_exact(Q, key64, f64, (int)warps, st));
        ++*launches;
    }
    return KMU_OK;
}

// deterministic natural logarithm, host copy of kmu_detmath.cuh (same operations, same results)
static double host_det_log(double x);

extern "C" {

int32_t kmu_sketch_superminhash(kmu_ctx* ctx, const kmu_seqbatch* b, uint32_t k, int32_t kmer_type, int32_t hash_kind,
                                uint32_t m, int32_t key_hasher, int32_t sig_bytes, void* sig, int32_t sig_on_device) {
    if (!ctx || !b) return fail(KMU_EINVAL, "null argument");
    if (int32_t a = kmu_check_kmer_args(b, k, kmer_type, hash_kind)) return a;
    if (key_hasher != KMU_HASHER_NOHASH && key_hasher != KMU_HASHER_FNV) return fail(KMU_EINVAL, "unknown key hasher %d", key_hasher);
    if (sig_bytes != 4 && sig_bytes != 8) return fail(KMU_EINVAL, "sig_bytes must be 4 (f32) or 8 (f64)");
    if (m < 1) return fail(KMU_EINVAL, "SuperMinHash needs a sketch size >= 1");
    const size_t team_bytes = align_up((size_t)m * sig_bytes, 16) + 32;
    if (team_bytes > SMEM_BUDGET)
        return fail(KMU_EINVAL, "sketch size %u does not fit the shared memory of one SM (max %zu slots of %d bytes)", m,
                    (SMEM_BUDGET - 32) / sig_bytes, sig_bytes);
    if (b->nseq == 0) return KMU_OK;
    if (!sig) return fail(KMU_EINVAL, "null signature buffer");
    if (b->nseq >= 0xFFFFFFFFull) return fail(KMU_EINVAL, "more than 2^32-1 sequences in one batch");
    for (uint64_t L : b->h_nbases)
        if (L >= 0xFFFFFF00ull) return fail(KMU_EINVAL, "a single sequence is limited to 2^32 - 256 bases");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScopedDevice sd(ctx->device);
    ctx->last = kmu_times{};
    cudaStream_t st = ctx->stream;
    const bool key64 = kmer_type_is_u64(kmer_type), f64 = sig_bytes == 8;
    const size_t out_bytes = (size_t)b->nseq * m * sig_bytes;
    void* d_sig = sig;
    if (!sig_on_device) {
        CUDA_TRY(ctx->sig_dev.reserve(out_bytes));
        d_sig = ctx->sig_dev.p;
    }
    uint64_t launches = 0;
    cudaEventRecord(ctx->ev[0], st);
    {
        int32_t rc = smh_per_sequence_device(ctx, b, k, kmer_type, hash_kind, m, key_hasher, sig_bytes, d_sig, &launches);
        if (rc) return rc;
    }
    cudaEventRecord(ctx->ev[1], st);
    if (!sig_on_device) {
        cudaEventRecord(ctx->ev[4], st);
        CUDA_TRY(cudaMemcpyAsync(sig, d_sig, out_bytes, cudaMemcpyDeviceToHost, st));
        cudaEventRecord(ctx->ev[5], st);
        ctx->last.d2h_bytes = out_bytes;
    }
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(KMU_ECUDA, "SuperMinHash sketch kernels failed: %s", cudaGetErrorString(e));
    cudaEventElapsedTime(&ctx->last.kernel_ms, ctx->ev[0], ctx->ev[1]);
    if (!sig_on_device) cudaEventElapsedTime(&ctx->last.d2h_ms, ctx->ev[4], ctx->ev[5]);
    ctx->launches += launches;
    ctx->last.launches = launches;
    return KMU_OK;
}

int32_t kmu_sketch_superminhash_whole(kmu_ctx* ctx, const kmu_seqbatch* b, uint32_t k, int32_t kmer_type, int32_t hash_kind,
                                      uint32_t m, int32_t key_hasher, int32_t sig_bytes, void* sig, int32_t sig_on_device) {
    if (!ctx || !b || !sig) return fail(KMU_EINVAL, "null argument");
    if (int32_t a = kmu_check_kmer_args(b, k, kmer_type, hash_kind)) return a;
    if (key_hasher != KMU_HASHER_NOHASH && key_hasher != KMU_HASHER_FNV) return fail(KMU_EINVAL, "unknown key hasher %d", key_hasher);
    if (sig_bytes != 4 && sig_bytes != 8) return fail(KMU_EINVAL, "sig_bytes must be 4 (f32) or 8 (f64)");
    if (m < 1) return fail(KMU_EINVAL, "SuperMinHash needs a sketch size >= 1");
    if (align_up((size_t)m * sig_bytes, 16) + 32 > SMEM_BUDGET)
        return fail(KMU_EINVAL, "sketch size %u does not fit the shared memory of one SM", m);
    if (b->nseq >= 0xFFFFFFFFull) return fail(KMU_EINVAL, "more than 2^32-1 sequences in one batch");
    uint64_t total = 0;
    for (uint64_t L : b->h_nbases) {
        if (L >= 0xFFFFFF00ull) return fail(KMU_EINVAL, "a single sequence is limited to 2^32 - 256 bases");
        total += L >= k ? L - k + 1 : 0;
    }
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScopedDevice sd(ctx->device);
    ctx->last = kmu_times{};
    cudaStream_t st = ctx->stream;
    const bool key64 = kmer_type_is_u64(kmer_type), f64 = sig_bytes == 8;
    const size_t row_bytes = (size_t)m * sig_bytes;
    CUDA_TRY(ctx->items_slots.reserve(row_bytes + 64));
    void* d_row = ctx->items_slots.p;  // the merged slots (bit patterns of S)
    uint64_t launches = 0;
    cudaEventRecord(ctx->ev[0], st);
    CUDA_TRY(kmu::launch_smh_fill_large(d_row, m, f64, st));
    ++launches;
    bool done = total == 0;
    // speculative loop bound from the total k-mer count (same rule as the per-sequence kernel)
    uint32_t a_spec = 0;
    if (total) {
        const double need = (double)m / (double)total * std::log(1e4 * (double)m);
        const uint32_t a1 = need >= (double)m ? m : (uint32_t)std::ceil(need);
        a_spec = std::max<uint32_t>(a1, 1) - 1;
    }
    // long DNA inputs: the value cut (an item whose value is not below 4 m ln(1e4 m) / n is dropped after half a seeding);
    // exact iff every merged slot ends below the cut, which is checked here -- otherwise the general kernel below
    if (!done && a_spec == 0 && !kmer_type_is_aa(kmer_type) && !std::getenv("KMU_SMH_NO_CUT") &&
        align_up(row_bytes, 16) + 16 * 64 * (key64 ? 8 : 4) <= SMEM_BUDGET) {
        double cut = 4.0 * (double)m / (double)total * std::log(1e4 * (double)m);
        if (!f64) cut = (double)(float)cut;
        if (cut < 0.9) {
            kmu::SmhParams P{};
            P.k = k;
            P.kmer_type = kmer_type;
            P.hash_kind = hash_kind;
            P.m = m;
            P.hasher = key_hasher;
            kmu::SeqView v{b->packed, b->byte_off, b->nbases, b->nseq};
            CUDA_TRY(kmu::launch_smh_whole_cut(P, key64, f64, v, b->packed_bytes, cut, d_row, ctx->sm_count, st));
            ++launches;
            std::vector<uint8_t> h(row_bytes);
            CUDA_TRY(cudaMemcpyAsync(h.data(), d_row, row_bytes, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            done = true;
            for (uint32_t j = 0; j < m && done; ++j)
                done = (f64 ? ((const double*)h.data())[j] : (double)((const float*)h.data())[j]) < cut;
            if (!done) {
                CUDA_TRY(kmu::launch_smh_fill_large(d_row, m, f64, st));
                ++launches;
            }
        }
    }
    if (!done && a_spec <= 15) {
        kmu::SmhParams P{};
        P.k = k;
        P.kmer_type = kmer_type;
        P.hash_kind = hash_kind;
        P.m = m;
        P.hasher = key_hasher;
        kmu::SeqView v{b->packed, b->byte_off, b->nbases, b->nseq};
        const uint64_t nchunks = (b->packed_bytes + 63) / 64;
        const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((nchunks + 511) / 512, (uint64_t)ctx->sm_count * 2));
        CUDA_TRY(kmu::launch_smh_whole(P, key64, f64, v, b->packed_bytes, a_spec, d_row, grid, align_up(row_bytes, 16), st));
        ++launches;
        std::vector<uint8_t> h(row_bytes);
        CUDA_TRY(cudaMemcpyAsync(h.data(), d_row, row_bytes, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        done = true;
        const double bound = (double)(a_spec + 1);
        for (uint32_t j = 0; j < m && done; ++j)
            done = (f64 ? ((const double*)h.data())[j] : (double)((const float*)h.data())[j]) < bound;
        if (!done) {
            CUDA_TRY(kmu::launch_smh_fill_large(d_row, m, f64, st));
            ++launches;
        }
    }
    if (!done) {
        // few k-mers relative to m, or the bound failed: per-sequence sketches (exact path inside), merged by minimum
        CUDA_TRY(ctx->sig_dev.reserve((size_t)b->nseq * row_bytes));
        int32_t rc = smh_per_sequence_device(ctx, b, k, kmer_type, hash_kind, m, key_hasher, sig_bytes, ctx->sig_dev.p, &launches);
        if (rc) return rc;
        CUDA_TRY(kmu::launch_smh_colmin(ctx->sig_dev.p, b->nseq, m, f64, d_row, st));
        ++launches;
    }
    cudaEventRecord(ctx->ev[1], st);
    CUDA_TRY(cudaMemcpyAsync(sig, d_row, row_bytes, sig_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    if (!sig_on_device) ctx->last.d2h_bytes = row_bytes;
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(KMU_ECUDA, "SuperMinHash whole-file sketch failed: %s", cudaGetErrorString(e));
    cudaEventElapsedTime(&ctx->last.kernel_ms, ctx->ev[0], ctx->ev[1]);
    ctx->launches += launches;
    ctx->last.launches = launches;
    return KMU_OK;
}

int32_t kmu_sketch_setsketch(kmu_ctx* ctx, const kmu_seqbatch* b, uint32_t k, int32_t kmer_type, int32_t hash_kind,
                             const kmu_setsketch_params* prm, int32_t sig_bytes, int32_t whole, void* sig,
                             int32_t sig_on_device) {
    if (!ctx || !b) return fail(KMU_EINVAL, "null argument");
    if (int32_t a = kmu_check_kmer_args(b, k, kmer_type, hash_kind)) return a;
    kmu_setsketch_params def{1.001, 4096, 20.0, 65534};  // SetSketchParams::default()
    const kmu_setsketch_params P0 = prm ? *prm : def;
    if (sig_bytes != 2 && sig_bytes != 4 && sig_bytes != 8) return fail(KMU_EINVAL, "sig_bytes must be 2, 4 or 8 (u16 / u32 / u64 registers)");
    if (!(P0.b > 1.0) || !(P0.a > 0.0) || P0.m < 1) return fail(KMU_EINVAL, "SetSketch needs b > 1, a > 0, m >= 1");
    const uint64_t reg_max = sig_bytes == 2 ? 0xFFFFull : 0x7FFFFFFEull;
    if (P0.q + 1 > reg_max) return fail(KMU_EINVAL, "q + 1 = %llu does not fit the register type", (unsigned long long)(P0.q + 1));
    const uint32_t m = (uint32_t)P0.m;
    const size_t table_bytes = 2 * 264 * sizeof(double);
    const size_t team_bytes = align_up((size_t)m * 4, 16) + 32;
    if (P0.m > 0xFFFFFFull || team_bytes + table_bytes + kmu::SSK_QUEUE_BYTES > SMEM_BUDGET)
        return fail(KMU_EINVAL, "m = %llu registers do not fit the shared memory of one SM", (unsigned long long)P0.m);
    if (b->nseq == 0 && !whole) return KMU_OK;
    if (!sig) return fail(KMU_EINVAL, "null signature buffer");
    if (b->nseq >= 0xFFFFFFFFull) return fail(KMU_EINVAL, "more than 2^32-1 sequences in one batch");
    uint64_t total_kmers = 0;
    for (uint64_t L : b->h_nbases) {
        if (L >= 0xFFFFFF00ull) return fail(KMU_EINVAL, "a single sequence is limited to 2^32 - 256 bases");
        total_kmers += L >= k ? L - k + 1 : 0;
    }
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScopedDevice sd(ctx->device);
    ctx->last = kmu_times{};
    cudaStream_t st = ctx->stream;
    const size_t nrows = whole ? 1 : b->nseq;
    const size_t out_bytes = nrows * m * (size_t)sig_bytes;
    void* d_sig = sig;
    if (!sig_on_device) {
        CUDA_TRY(ctx->sig_dev.reserve(out_bytes));
        d_sig = ctx->sig_dev.p;
    }
    uint64_t launches = 0;
    cudaEventRecord(ctx->ev[0], st);
    int32_t rc = kmu_ensure_order(ctx, b, k, &launches);
    if (rc) return rc;
    // counters: [0..127] work counters, [128] slow count, [129] exact count
    CUDA_TRY(ctx->counters.reserve(sizeof(unsigned long long) * 256));
    CUDA_TRY(cudaMemsetAsync(ctx->counters.p, 0, sizeof(unsigned long long) * 256, st));
    unsigned long long* d_work = (unsigned long long*)ctx->counters.p;
    // lists: slow | exact | kmin, nseq + 1 entries each; then m whole-batch registers
    CUDA_TRY(ctx->overflow.reserve(sizeof(uint32_t) * (3 * (b->nseq + 1) + m)));
    uint32_t* d_slow = (uint32_t*)ctx->overflow.p;
    uint32_t* d_exact = d_slow + (b->nseq + 1);
    uint32_t* d_kmin = d_exact + (b->nseq + 1);
    uint32_t* d_whole = d_kmin + (b->nseq + 1);
    kmu::SskParams P{};
    P.packed = b->packed;
    P.byte_off = b->byte_off;
    P.nbases = b->nbases;
    P.order = (const uint32_t*)b->order_cache.order.p;
    P.k = k;
    P.kmer_type = kmer_type;
    P.hash_kind = hash_kind;
    P.C.a = P0.a;
    P.C.inva = 1.0 / P0.a;
    P.C.inva_m0 = P.C.inva / (double)m;  // IEEE division, as __ddiv_rn on the device
    P.C.lnb = host_det_log(P0.b);
    P.C.ln_term = std::log(1e4 * (double)m);
    P.C.m = m;
    P.C.iq1 = (int)P0.q + 1;
    P.sig = d_sig;
    P.sig_bytes = sig_bytes;
    P.kmin_out = d_kmin;
    P.slow_count = d_work + 128;
    P.slow_list = d_slow;
    P.exact_count = d_work + 129;
    P.exact_list = d_exact;
    // per-sequence sketches speculate less cautiously than the whole-file one (a redo costs one sequence, not the file):
    // one sketch in ~20 is redone, every item places 4-5 times fewer points
    P.C.spec_ln = std::log(8.0 * (double)m);
    P.C.spec_dfrac = 0.85;
    {
        const double bits = (b->alphabet == 0 ? 2.0 : std::log2(20.0)) * (double)k;  // keys the sequence can hold
        P.C.spec_keyspace = bits < 40.0 ? std::exp2(bits) * (kmu::hash_kind_is_canonical_host(hash_kind) ? 0.5 : 1.0) : 0.0;
    }
    // below this many k-mers an item places too many points for the sparse permutation of the speculative kernels
    // (on average m spec_ln / (spec_dfrac nk) of them, capped at SSK_SPARSE_CAP = 24): the exact path takes the sequence
    P.exact_nk_max = (uint64_t)std::ceil((double)m * P.C.spec_ln / (P.C.spec_dfrac * 7.5));
    P.whole_regs = d_whole;
    int ci = 0;
    auto run_exact = [&](const uint32_t* list, uint64_t count, bool group) -> int32_t {
        kmu::SskParams Q = P;
        Q.order = list;
        Q.first = 0;
        Q.count = count;
        Q.group = group ? 1 : 0;
        Q.work_counter = d_work + ci++;
        const uint64_t per_warp = align_up((uint64_t)m * 4, 16) + 32ull * 2 * m * sizeof(uint32_t);
        // one-warp CTAs waiting on chains of f64 operations (0.24 IPC with 8 per SM): as many as an SM holds
        uint64_t warps = group ? 1 : std::min<uint64_t>(count, (uint64_t)ctx->sm_count * 24);
        const uint64_t budget = 8ull << 30;
        if (warps * per_warp > budget) warps = std::max<uint64_t>(1, budget / per_warp);
        CUDA_TRY(ctx->table_scratch.reserve(warps * per_warp));
        ctx->table_scratch_clean = false;
        CUDA_TRY(cudaMemsetAsync(ctx->table_scratch.p, 0, warps * per_warp, st));
        Q.scratch = (uint8_t*)ctx->table_scratch.p;
        Q.scratch_per_warp = per_warp;
        CUDA_TRY(kmu::launch_ssk_exact(Q, (int)warps, st));
        ++launches;
        return KMU_OK;
    };

    if (whole) {
        // the speculative level from the total number of k-mers (same formula as the device side)
        const double ratio = (double)total_kmers * 0.25 * P.C.a / P.C.ln_term;
        uint32_t kspec = 0;
        if (ratio > 1.0) {
            const double kf = 1.0 + std::floor(std::log(ratio) / P.C.lnb);
            kspec = (uint32_t)std::min(kf, (double)(P.C.iq1 - 1));
        }
        bool done = false;
        // (the whole-file cut keeps the cautious bound: an item places ~4 m ln(1e4 m) / nk points below it)
        if ((double)total_kmers > 4.0 * (double)m * P.C.ln_term / 7.5 && kspec > 0) {
            kmu::SeqView v{b->packed, b->byte_off, b->nbases, b->nseq};
            const size_t smem = table_bytes + (size_t)m * 4;
            std::vector<uint32_t> regs(m);
            for (int pass = 0; pass < 2 && !done; ++pass) {
                CUDA_TRY(cudaMemsetAsync(d_whole, 0, sizeof(uint32_t) * m, st));
                CUDA_TRY(cudaMemsetAsync(P.slow_count, 0, sizeof(unsigned long long), st));
                const double xcut = std::exp(-(double)kspec * P.C.lnb) * (1.0 + 1e-6);
                const uint64_t nchunks = (b->packed_bytes + 63) / 64;
                const int grid = (int)std::min<uint64_t>((nchunks + 511) / 512, (uint64_t)ctx->sm_count * 2);
                CUDA_TRY(kmu::launch_ssk_whole(P, v, b->packed_bytes, kspec, xcut, std::max(grid, 1), smem, st));
                ++launches;
                unsigned long long ovf = 0;
                CUDA_TRY(cudaMemcpyAsync(regs.data(), d_whole, sizeof(uint32_t) * m, cudaMemcpyDeviceToHost, st));
                CUDA_TRY(cudaMemcpyAsync(&ovf, P.slow_count, sizeof(ovf), cudaMemcpyDeviceToHost, st));
                CUDA_TRY(cudaStreamSynchronize(st));
                const uint32_t mn = *std::min_element(regs.begin(), regs.end());
                if (!ovf && mn >= kspec) done = true;
                else if (ovf || mn == 0) break;  // exact path
                else kspec = mn;                 // a true lower bound of every final register: cannot fail again
            }
            if (done) {
                CUDA_TRY(kmu::launch_ssk_store(d_whole, m, d_sig, sig_bytes, st));
                ++launches;
            }
        }
        if (!done) {
            if (b->nseq == 0) {
                CUDA_TRY(cudaMemsetAsync(d_sig, 0, out_bytes, st));
            } else {
                rc = run_exact((const uint32_t*)b->order_cache.order.p, b->nseq, true);
                if (rc) return rc;
            }
        }
    } else {
        std::vector<OctaveClass> classes = kmu_octave_classes(b);
        if (classes.size() > 100) return fail(KMU_EINVAL, "too many launch classes");
        struct Launch { uint64_t first, count; TeamGeometry g; };
        std::vector<Launch> ls;
        for (const OctaveClass& c : classes) {
            TeamGeometry g = team_geometry(c.nk_max, team_bytes, table_bytes + kmu::SSK_QUEUE_BYTES);
            if (!ls.empty() && ls.back().g.team_warps == g.team_warps && ls.back().g.teams_per_cta == g.teams_per_cta &&
                ls.back().first + ls.back().count == c.first) {
                ls.back().count += c.count;
            } else {
                ls.push_back({c.first, c.count, g});
            }
        }
        auto run_team = [&](const kmu::SskParams& base, const uint32_t* order, uint64_t first, uint64_t count,
                            const TeamGeometry& g) -> int32_t {
            kmu::SskParams Q = base;
            Q.order = order;
            Q.first = first;
            Q.count = count;
            Q.work_counter = d_work + ci++;
            Q.team_warps = g.team_warps;
            Q.team_smem_bytes = g.team_smem_bytes;
            const uint64_t ctas_needed = (count + g.teams_per_cta - 1) / g.teams_per_cta;
            const int grid = (int)std::min<uint64_t>(ctas_needed, (uint64_t)ctx->sm_count);
            CUDA_TRY(kmu::launch_ssk_team(Q, std::max(grid, 1), g.block, g.smem, st));
            ++launches;
            return KMU_OK;
        };
        for (const Launch& l : ls) {
            rc = run_team(P, P.order, l.first, l.count, l.g);
            if (rc) return rc;
        }
        unsigned long long cnt[2] = {0, 0};
        CUDA_TRY(cudaMemcpyAsync(cnt, P.slow_count, sizeof(cnt), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (cnt[0]) {  // failed speculations: once more with the level they reached; anything left goes to the exact path
            kmu::SskParams Q = P;
            Q.kspec_in = d_kmin;
            Q.slow_count = P.exact_count;
            Q.slow_list = P.exact_list;
            rc = run_team(Q, d_slow, 0, cnt[0], team_geometry(b->order_cache.nk_longest, team_bytes, table_bytes + kmu::SSK_QUEUE_BYTES));
            if (rc) return rc;
            CUDA_TRY(cudaMemcpyAsync(cnt, P.slow_count, sizeof(cnt), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
        }
        if (cnt[1]) {
            rc = run_exact(d_exact, cnt[1], false);
            if (rc) return rc;
        }
    }
    cudaEventRecord(ctx->ev[1], st);
    if (!sig_on_device) {
        cudaEventRecord(ctx->ev[4], st);
        CUDA_TRY(cudaMemcpyAsync(sig, d_sig, out_bytes, cudaMemcpyDeviceToHost, st));
        cudaEventRecord(ctx->ev[5], st);
        ctx->last.d2h_bytes = out_bytes;
    }
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(KMU_ECUDA, "SetSketch kernels failed: %s", cudaGetErrorString(e));
    cudaEventElapsedTime(&ctx->last.kernel_ms, ctx->ev[0], ctx->ev[1]);
    if (!sig_on_device) cudaEventElapsedTime(&ctx->last.d2h_ms, ctx->ev[4], ctx->ev[5]);
    ctx->launches += launches;
    ctx->last.launches = launches;
    return KMU_OK;
}

}  // extern "C"

// ---- host copy of det_log (kmu_detmath.cuh): ln(b) must be the same double on the host, on the
// device and in the oracle.  This translation unit is compiled with --fmad=false.
#include <cstring>
static inline uint64_t hd_bits(double v) { uint64_t b; std::memcpy(&b, &v, 8); return b; }
static inline double hd_from(uint64_t b) { double v; std::memcpy(&v, &b, 8); return v; }
static double host_det_log(double x) {
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                 Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                 Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    uint64_t bits = hd_bits(x);  // callers pass 1 < b < 2^1000: normal, positive
    int32_t hx = (int32_t)(bits >> 32);
    const uint32_t lx = (uint32_t)bits;
    int32_t k = (hx >> 20) - 1023;
    hx &= 0x000fffff;
    const int32_t i0 = (hx + 0x95f64) & 0x100000;
    x = hd_from(((uint64_t)(uint32_t)(hx | (i0 ^ 0x3ff00000)) << 32) | lx);
    k += (i0 >> 20);
    const volatile double f = x - 1.0;
    const double dk = (double)k;
    if ((0x000fffff & (2 + hx)) < 3) {
        if (f == 0.0) return k == 0 ? 0.0 : dk * ln2_hi + dk * ln2_lo;
        const volatile double ff = f * f;
        const volatile double inner = 0.33333333333333333 * f;
        const volatile double R = ff * (0.5 - inner);
        if (k == 0) return f - R;
        const volatile double t = dk * ln2_lo;
        const volatile double u = R - t;
        const volatile double hi = dk * ln2_hi;
        return hi - (u - f);
    }
    const volatile double s = f / (2.0 + f);
    const volatile double z = s * s;
    int32_t i = hx - 0x6147a;
    const volatile double w = z * z;
    const int32_t j = 0x6b851 - hx;
    volatile double a1 = w * Lg6; a1 = Lg4 + a1; a1 = w * a1; a1 = Lg2 + a1;
    const volatile double t1 = w * a1;
    volatile double a2 = w * Lg7; a2 = Lg5 + a2; a2 = w * a2; a2 = Lg3 + a2; a2 = w * a2; a2 = Lg1 + a2;
    const volatile double t2 = z * a2;
    i |= j;
    const volatile double R = t2 + t1;
    if (i > 0) {
        volatile double hfsq = 0.5 * f; hfsq = hfsq * f;
        volatile double v1 = hfsq + R; v1 = s * v1;
        if (k == 0) { const volatile double d = hfsq - v1; return f - d; }
        const volatile double t = dk * ln2_lo;
        volatile double v2 = v1 + t; v2 = hfsq - v2; v2 = v2 - f;
        const volatile double hi = dk * ln2_hi;
        return hi - v2;
    }
    volatile double v1 = f - R; v1 = s * v1;
    if (k == 0) return f - v1;
    const volatile double t = dk * ln2_lo;
    volatile double v2 = v1 - t; v2 = v2 - f;
    const volatile double hi = dk * ln2_hi;
    return hi - v2;
}
