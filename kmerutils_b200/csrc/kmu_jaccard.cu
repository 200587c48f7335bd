// kmu_jaccard.cu -- Jaccard estimates between signatures: the fraction of equal slots
// (probminhash::compute_probminhash_jaccard, its local restatement probminhash_get_jaccard_objects,
// src/sketching/seqsketchjaccard.rs:86-108, and the comparison step of jaccard_index_probminhash3a, :423-495;
// SuperMinHash::get_jaccard_index_estimate is the same count on f32 / f64 slots).
// One CTA keeps up to 8 rows of A in shared memory (fewer when the signatures are long: gsearch's 12 000 slots of u32
// are 48 KB a row; rows too long for even one to fit are read through L1 / L2 instead); its warps stream rows of B once
// each and compare them against all of them.
#include <cstdint>

#include "kmu_host.h"

namespace kmu {

constexpr int JA_ROWS = 8;

template <typename T>
__global__ void __launch_bounds__(256) jaccard_kernel(const T* __restrict__ a, uint64_t na, const T* __restrict__ b,
                                                       uint64_t nb, uint32_t m, double* __restrict__ out, int rows_cap,
                                                       int a_in_smem) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint64_t a0 = (uint64_t)blockIdx.x * rows_cap;
    const int rows = (int)min((uint64_t)rows_cap, na - a0);
    const T* sa = a + a0 * m;  // rows * m
    if (a_in_smem) {
        T* s = (T*)smem;
        for (uint32_t i = threadIdx.x; i < (uint32_t)rows * m; i += blockDim.x) s[i] = a[a0 * m + i];
        sa = s;
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (uint64_t j = (uint64_t)blockIdx.y * nwarps + warp; j < nb; j += (uint64_t)gridDim.y * nwarps) {
        uint32_t eq[JA_ROWS];
#pragma unroll
        for (int r = 0; r < JA_ROWS; ++r) eq[r] = 0;
        const T* brow = b + j * m;
        for (uint32_t s = lane; s < m; s += 32) {
            const T v = brow[s];
#pragma unroll
            for (int r = 0; r < JA_ROWS; ++r)
                if (r < rows) eq[r] += sa[(uint32_t)r * m + s] == v;
        }
#pragma unroll
        for (int r = 0; r < JA_ROWS; ++r) {
            const uint32_t t = __reduce_add_sync(0xFFFFFFFFu, eq[r]);
            if (lane == 0 && r < rows) out[(a0 + r) * nb + j] = (double)t / (double)m;
        }
    }
}

template <typename T>
static cudaError_t launch_jaccard_t(const void* a, uint64_t na, const void* b, uint64_t nb, uint32_t m, double* out,
                                    int sm_count, cudaStream_t st) {
    auto kern = jaccard_kernel<T>;
    int rows_cap = (int)std::min<size_t>(JA_ROWS, SMEM_BUDGET / ((size_t)m * sizeof(T)));
    const int a_in_smem = rows_cap >= 1;
    if (!a_in_smem) rows_cap = JA_ROWS;
    const size_t smem = a_in_smem ? (size_t)rows_cap * m * sizeof(T) : 0;
    static size_t configured = 0;
    if (smem > configured && smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    const uint64_t gx = (na + rows_cap - 1) / rows_cap;
    uint64_t gy = (nb + 7) / 8;
    const uint64_t want_y = std::max<uint64_t>(1, (uint64_t)sm_count * 4 / std::max<uint64_t>(gx, 1));
    if (gy > want_y) gy = want_y;
    dim3 grid((unsigned)std::min<uint64_t>(gx, 0x7FFFFFFF), (unsigned)std::min<uint64_t>(gy, 65535));
    kern<<<grid, 256, smem, st>>>((const T*)a, na, (const T*)b, nb, m, out, rows_cap, a_in_smem);
    return cudaGetLastError();
}

}  // namespace kmu

extern "C" int32_t kmu_signature_jaccard(kmu_ctx* ctx, const void* sig_a, uint64_t na, const void* sig_b, uint64_t nb, uint32_t m,
                                         int32_t slot_bytes, double* out, int32_t on_device) {
    if (!ctx || (na && nb && (!sig_a || !sig_b || !out))) return fail(KMU_EINVAL, "null argument");
    if (slot_bytes != 2 && slot_bytes != 4 && slot_bytes != 8) return fail(KMU_EINVAL, "slot_bytes must be 2, 4 or 8");
    if (m < 1) return fail(KMU_EINVAL, "empty signatures");
    if (na > 0x7FFFFFFFull) return fail(KMU_EINVAL, "too many signatures");
    if (na == 0 || nb == 0) return KMU_OK;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScopedDevice sd(ctx->device);
    ctx->last = kmu_times{};
    cudaStream_t st = ctx->stream;
    const size_t ab = (size_t)na * m * slot_bytes, bb = (size_t)nb * m * slot_bytes, ob = (size_t)na * nb * sizeof(double);
    const void *da = sig_a, *db = sig_b;
    double* dout = out;
    if (!on_device) {
        CUDA_TRY(ctx->misc.reserve(align_up(ab, 256) + align_up(bb, 256) + ob));
        uint8_t* p = (uint8_t*)ctx->misc.p;
        CUDA_TRY(cudaMemcpyAsync(p, sig_a, ab, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(p + align_up(ab, 256), sig_b, bb, cudaMemcpyHostToDevice, st));
        da = p;
        db = p + align_up(ab, 256);
        dout = (double*)(p + align_up(ab, 256) + align_up(bb, 256));
        ctx->last.h2d_bytes = ab + bb;
    }
    cudaEventRecord(ctx->ev[0], st);
    cudaError_t e;
    // equality of slots is bit equality for the integer signatures; the float signatures of SuperMinHash never hold NaN
    if (slot_bytes == 2) e = kmu::launch_jaccard_t<uint16_t>(da, na, db, nb, m, dout, ctx->sm_count, st);
    else if (slot_bytes == 4) e = kmu::launch_jaccard_t<uint32_t>(da, na, db, nb, m, dout, ctx->sm_count, st);
    else e = kmu::launch_jaccard_t<uint64_t>(da, na, db, nb, m, dout, ctx->sm_count, st);
    CUDA_TRY(e);
    cudaEventRecord(ctx->ev[1], st);
    ctx->launches += 1;
    ctx->last.launches = 1;
    if (!on_device) {
        CUDA_TRY(cudaMemcpyAsync(out, dout, ob, cudaMemcpyDeviceToHost, st));
        ctx->last.d2h_bytes = ob;
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&ctx->last.kernel_ms, ctx->ev[0], ctx->ev[1]);
    return KMU_OK;
}
