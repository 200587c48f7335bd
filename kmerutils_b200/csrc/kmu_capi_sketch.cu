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
TeamGeometry team_geometry(uint64_t nk_max, size_t team_bytes) {
    TeamGeometry g{};
    uint32_t tw = 1;
    while (tw < 32 && (uint64_t)tw * 1024 < nk_max) tw <<= 1;
    team_bytes = align_up(team_bytes, 16);
    for (;;) {
        uint32_t fit = (uint32_t)std::max<size_t>(1, SMEM_BUDGET / team_bytes);
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
        g.smem = team_bytes * g.teams_per_cta;
        return g;
    }
}

}  // namespace

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
    int32_t rc = kmu_ensure_order(ctx, b, k, &launches);
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
    // merge neighbouring octave classes that get the same team geometry: one launch each
    struct Launch { uint64_t first, count; TeamGeometry g; };
    std::vector<Launch> ls;
    for (const OctaveClass& c : classes) {
        TeamGeometry g = team_geometry(c.nk_max, team_bytes);
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
        ++launches;
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
        ++launches;
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

}  // extern "C"
