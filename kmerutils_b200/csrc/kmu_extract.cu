// kmu_extract.cu -- materialising kernels: all k-mers of every sequence mapped through a hash
// closure (KmerGenerator::generate_kmer, src/base/kmergenerator.rs:162-167 +
// KmerSeqIterator::next :75-106), and canonical ntHash (src/base/kmer.rs:74-94,
// src/base/nthash.rs:63-72).  Both are streaming kernels bounded by their HBM writes.
#include <cstdint>
#include <cstdlib>

#include "kmu_device.cuh"
#include "kmu_kernels.h"

namespace kmu {

// first sequence s with out_off[s+1] > e  (out_off has nseq+1 entries, non-decreasing)
__device__ __forceinline__ uint64_t seq_of_element(const uint64_t* __restrict__ out_off, uint64_t nseq, uint64_t e) {
    uint64_t lo = 0, hi = nseq;  // invariant: out_off[lo] <= e < out_off[hi]
    while (hi - lo > 1) {
        uint64_t mid = (lo + hi) >> 1;
        if (out_off[mid] <= e) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void kmer_offsets_kernel(const uint64_t* __restrict__ nbases, uint64_t nseq, uint32_t k, uint64_t* nk) {
    for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < nseq; s += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t L = nbases[s];
        nk[s] = L >= k ? L - k + 1 : 0;
    }
}

cudaError_t launch_kmer_offsets(const uint64_t* nbases, uint64_t nseq, uint32_t k, uint64_t* out, cudaStream_t stream) {
    if (nseq == 0) return cudaSuccess;
    int block = 256;
    uint64_t want = (nseq + block - 1) / block;
    int grid = (int)(want < 148ull * 4 ? want : 148ull * 4);
    kmer_offsets_kernel<<<grid, block, 0, stream>>>(nbases, nseq, k, out);
    return cudaGetLastError();
}

// every thread produces T consecutive output elements
template <typename V, int T, bool AA>
__global__ void __launch_bounds__(256) generate_kmers_kernel(SeqView b, uint32_t k, int kmer_type, int hash_kind,
                                                              const uint64_t* __restrict__ out_off, V* __restrict__ out) {
    const uint64_t total = out_off[b.nseq];
    const V header = (V)word_header(kmer_type, k);
    const bool canonical = hash_is_canonical(hash_kind);
    for (uint64_t e0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * T; e0 < total;
         e0 += (uint64_t)gridDim.x * blockDim.x * T) {
        uint64_t s = seq_of_element(out_off, b.nseq, e0);
        uint64_t s_begin = out_off[s], s_end = out_off[s + 1];
        typename KmerSource<V, AA>::Walker wk;
        wk.start((const uint32_t*)(b.packed + b.byte_off[s]), e0 - s_begin, k);
        V vals[T];
#pragma unroll
        for (int t = 0; t < T; ++t) {
            uint64_t e = e0 + t;
            vals[t] = 0;
            if (e < total) {
                if (e >= s_end) {  // crossed into the next non-empty sequence
                    do {
                        ++s;
                        s_end = out_off[s + 1];
                    } while (e >= s_end);
                    wk.start((const uint32_t*)(b.packed + b.byte_off[s]), 0, k);
                }
                wk.roll();
                vals[t] = finalize_key<V>(wk.prekey(canonical), header, hash_kind);
            }
        }
        if (e0 + T <= total) {
            constexpr int VEC = 16 / sizeof(V);
#pragma unroll
            for (int t = 0; t < T; t += VEC) {
                uint4 v;
                if (sizeof(V) == 4) {
                    v = make_uint4((uint32_t)vals[t], (uint32_t)vals[t + 1], (uint32_t)vals[t + 2], (uint32_t)vals[t + 3]);
                } else {
                    uint64_t a = (uint64_t)vals[t], c = (uint64_t)vals[t + 1 < T ? t + 1 : t];
                    v = make_uint4((uint32_t)a, (uint32_t)(a >> 32), (uint32_t)c, (uint32_t)(c >> 32));
                }
                *(uint4*)(out + e0 + t) = v;
            }
        } else {
            for (int t = 0; t < T && e0 + t < total; ++t) out[e0 + t] = vals[t];
        }
    }
}

// DNA batches: warp-cooperative version.  A warp owns the k-mers starting in one 2 KB group of the packed
// buffer; lane l takes positions p0 + l, p0 + l + 32, ... read straight from the packed words (kmer_at), so every
// store instruction of the warp writes one contiguous run of 32 values.
template <typename V, int U>
__global__ void __launch_bounds__(256) generate_kmers_warp_kernel(SeqView b, uint64_t total_bytes, uint32_t k, int kmer_type,
                                                                   int hash_kind, const uint64_t* __restrict__ out_off,
                                                                   V* __restrict__ out) {
    const V header = (V)word_header(kmer_type, k);
    const bool canonical = hash_is_canonical(hash_kind);
    const uint64_t ngroups = (total_bytes + GROUP_BYTES - 1) / GROUP_BYTES;
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t g = warp; g < ngroups; g += nwarps) {
        const uint64_t byte0 = g * GROUP_BYTES;
        const uint64_t byte1 = min(byte0 + (uint64_t)GROUP_BYTES, total_bytes);
        uint64_t s = seq_of_byte(b.byte_off, b.nseq, byte0);
        while (s < b.nseq) {
            const uint64_t sb = __ldg(b.byte_off + s);
            if (sb >= byte1) break;
            const uint64_t L = __ldg(b.nbases + s);
            const uint64_t nk = L >= k ? L - k + 1 : 0;
            const uint64_t p_lo = byte0 > sb ? (byte0 - sb) * 4 : 0;
            const uint64_t p_hi = min(nk, (byte1 - sb) * 4);
            const uint32_t* words = (const uint32_t*)(b.packed + sb);
            V* o = out + __ldg(out_off + s);
            for (uint64_t p0 = p_lo; p0 < p_hi; p0 += 32 * U) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint64_t p = p0 + lane + 32 * u;
                    if (p < p_hi) {
                        V key = kmer_at<V>(words, p, k);
                        if (canonical) {
                            const V rc = revcomp_val(key, k);
                            key = key < rc ? key : rc;
                        }
                        o[p] = finalize_key<V>(key, header, hash_kind);
                    }
                }
            }
            ++s;
        }
    }
}

// Vector form of the warp kernel: a lane takes Q = 16 / sizeof(V) consecutive k-mers (one 64-bit window of 32 bases
// read from three packed words serves all of them, forward and reverse strand) and writes them with ONE 16-byte store,
// so every store instruction of the warp writes 512 contiguous bytes.  Quads are aligned on the OUTPUT element index
// (out must be 16-byte aligned); the ragged quads at the ends of a sequence / group fall back to scalar stores.
template <typename V, int U>
__global__ void __launch_bounds__(256) generate_kmers_vec_kernel(SeqView b, uint64_t total_bytes, uint32_t k, int kmer_type,
                                                                  int hash_kind, const uint64_t* __restrict__ out_off,
                                                                  V* __restrict__ out) {
    constexpr int Q = 16 / (int)sizeof(V);
    const V header = (V)word_header(kmer_type, k);
    const bool canonical = hash_is_canonical(hash_kind);
    const V mask = value_mask<V>(2 * k);
    const uint64_t ngroups = (total_bytes + GROUP_BYTES - 1) / GROUP_BYTES;
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t g = warp; g < ngroups; g += nwarps) {
        const uint64_t byte0 = g * GROUP_BYTES;
        const uint64_t byte1 = min(byte0 + (uint64_t)GROUP_BYTES, total_bytes);
        uint64_t s = seq_of_byte(b.byte_off, b.nseq, byte0);
        while (s < b.nseq) {
            const uint64_t sb = __ldg(b.byte_off + s);
            if (sb >= byte1) break;
            const uint64_t L = __ldg(b.nbases + s);
            const uint64_t nk = L >= k ? L - k + 1 : 0;
            const uint64_t p_lo = byte0 > sb ? (byte0 - sb) * 4 : 0;
            const uint64_t p_hi = min(nk, (byte1 - sb) * 4);
            const uint32_t* words = (const uint32_t*)(b.packed + sb);
            const uint64_t ob = __ldg(out_off + s);
            const uint64_t e_lo = ob + p_lo, e_hi = ob + p_hi;
            for (uint64_t e0 = e_lo & ~(uint64_t)(Q - 1); e0 < e_hi; e0 += 32 * Q * U) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint64_t e = e0 + (uint64_t)(u * 32 + lane) * Q;
                    if (e < e_hi) {
                        const uint32_t t_lo = e < e_lo ? (uint32_t)(e_lo - e) : 0u;
                        const uint32_t t_hi = (uint32_t)min((uint64_t)Q, e_hi - e);
                        const uint64_t p = e + t_lo - ob;  // first position this lane produces
                        const uint32_t* w = words + (p >> 4);
                        const uint32_t sh = (uint32_t)(p & 15) * 2;
                        const uint32_t wa = be32(__ldg(w)), wb = be32(__ldg(w + 1)), wc = be32(__ldg(w + 2));
                        V vals[Q];
                        if (sizeof(V) == 4) {
                            // 32 bases from p; k-mer t = bases t .. t + k - 1 of the window, its reverse complement sits
                            // 2 t bits above the bottom of the window's reverse complement
                            const uint64_t x = ((uint64_t)__funnelshift_l(wb, wa, sh) << 32) | __funnelshift_l(wc, wb, sh);
                            const uint64_t rc = revcomp_word64(x);
#pragma unroll
                            for (int t = 0; t < Q; ++t) {
                                V key = (V)(x >> (64 - 2 * k - 2 * t)) & mask;
                                if (canonical) {
                                    const V r = (V)(rc >> (2 * t)) & mask;
                                    key = key < r ? key : r;
                                }
                                vals[t] = finalize_key<V>(key, header, hash_kind);
                            }
                        } else {
#pragma unroll
                            for (int t = 0; t < Q; ++t) {
                                const uint32_t st = sh + 2 * t;  // <= 32: the clamping funnel shift
                                const uint64_t x = ((uint64_t)__funnelshift_lc(wb, wa, st) << 32) | __funnelshift_lc(wc, wb, st);
                                V key = (V)(x >> (64 - 2 * k));
                                if (canonical) {
                                    const V r = (V)revcomp_val((uint64_t)key, k);
                                    key = key < r ? key : r;
                                }
                                vals[t] = finalize_key<V>(key, header, hash_kind);
                            }
                        }
                        if (t_lo == 0 && t_hi == Q) {
                            uint4 v;
                            if (sizeof(V) == 4) {
                                v = make_uint4((uint32_t)vals[0], (uint32_t)vals[1], (uint32_t)vals[Q > 2 ? 2 : 0], (uint32_t)vals[Q > 2 ? 3 : 0]);
                            } else {
                                const uint64_t a0 = (uint64_t)vals[0], a1 = (uint64_t)vals[1];
                                v = make_uint4((uint32_t)a0, (uint32_t)(a0 >> 32), (uint32_t)a1, (uint32_t)(a1 >> 32));
                            }
                            __stcs((uint4*)(out + e), v);  // streaming: the output is not read again by this kernel
                        } else {
                            for (uint32_t t = 0; t + t_lo < t_hi; ++t) out[e + t_lo + t] = vals[t];
                        }
                    }
                }
            }
            ++s;
        }
    }
}

cudaError_t launch_generate_kmers(const SeqView& b, uint64_t total_bytes, uint32_t k, int kmer_type, int hash_kind,
                                  const uint64_t* out_off, void* out, cudaStream_t stream) {
    if (b.nseq == 0) return cudaSuccess;
    const int block = 256;
    const int grid = 148 * 8;
    if (kmer_type == KMU_KMERAA64)
        generate_kmers_kernel<uint64_t, 8, true><<<grid, block, 0, stream>>>(b, k, kmer_type, hash_kind, out_off, (uint64_t*)out);
    else if (kmer_type == KMU_KMERAA32)
        generate_kmers_kernel<uint32_t, 8, true><<<grid, block, 0, stream>>>(b, k, kmer_type, hash_kind, out_off, (uint32_t*)out);
    else if (kmer_type == KMU_KMER64 && ((uintptr_t)out & 15) == 0)
        generate_kmers_vec_kernel<uint64_t, 2><<<grid, block, 0, stream>>>(b, total_bytes, k, kmer_type, hash_kind, out_off, (uint64_t*)out);
    else if (kmer_type != KMU_KMER64 && ((uintptr_t)out & 15) == 0)
        generate_kmers_vec_kernel<uint32_t, 2><<<grid, block, 0, stream>>>(b, total_bytes, k, kmer_type, hash_kind, out_off, (uint32_t*)out);
    else if (kmer_type == KMU_KMER64)
        generate_kmers_warp_kernel<uint64_t, 4><<<grid, block, 0, stream>>>(b, total_bytes, k, kmer_type, hash_kind, out_off, (uint64_t*)out);
    else
        generate_kmers_warp_kernel<uint32_t, 4><<<grid, block, 0, stream>>>(b, total_bytes, k, kmer_type, hash_kind, out_off, (uint32_t*)out);
    return cudaGetLastError();
}

// ---- ntHash -----------------------------------------------------------------------------------
// seeds: src/base/nthash.rs:17-20
__constant__ uint64_t NT_SEED[4] = {0x3c8bfbb395c60474ULL, 0x3193c18562a02b4cULL, 0x20323ed082572324ULL,
                                    0x295549f54be24456ULL};
__device__ __forceinline__ uint64_t rotl_var(uint64_t x, uint32_t r) {
    r &= 63;
    return r ? (x << r) | (x >> (64 - r)) : x;
}
__device__ __forceinline__ uint64_t nt_seed(uint32_t b) {
    // select without a memory access: 4 immediates
    uint64_t lo = (b & 1) ? 0x3193c18562a02b4cULL : 0x3c8bfbb395c60474ULL;
    uint64_t hi = (b & 1) ? 0x295549f54be24456ULL : 0x20323ed082572324ULL;
    return (b & 2) ? hi : lo;
}

// warp-cooperative version: a warp owns the k-mers starting in one 2 KB group and takes them in tiles of
// 32 x NT_RUN positions; every lane computes a run of NT_RUN consecutive positions (one O(k) initialisation,
// then the O(1) ntHash recurrence) into the warp's shared-memory tile, and the tile goes out with every store
// instruction writing 32 consecutive values
constexpr uint32_t NT_RUN = 32;
constexpr uint32_t NT_PITCH = NT_RUN + 1;                                          // u64 elements per run: lanes 2 banks apart
constexpr uint32_t NT_TILE_BYTES = 32 * NT_PITCH * 8 + ((32 * NT_PITCH + 15) & ~15u);  // hashes + strand bytes
__global__ void __launch_bounds__(128) nthash_warp_kernel(SeqView b, uint64_t total_bytes, uint32_t k, uint32_t n_multi,
                                                           const uint64_t* __restrict__ out_off, uint64_t* __restrict__ out_hash,
                                                           uint8_t* __restrict__ out_strand) {
    extern __shared__ __align__(16) uint8_t nt_smem[];
    // per-launch tables: the seeds, their complements, and the recurrence's two XOR terms for every (outgoing base,
    // incoming base) pair:  f' = rotl(f, 1) ^ FD[out][in],  r' = rotl(r, 63) ^ RD[out][in]
    __shared__ uint64_t SEED[4], SEEDC[4], FD[16], RD[16], F2[16], R2[16];
    if (threadIdx.x < 16) {
        const uint32_t ob = threadIdx.x >> 2, nb = threadIdx.x & 3;
        // initialisation two bases at a time: F2[(x << 2) | y] for consecutive bases x, y (forward strand, x first),
        // R2[(b << 2) | a] for the reverse strand taken from the last base backwards (a first)
        F2[threadIdx.x] = rotl_var(nt_seed(ob), 1) ^ nt_seed(nb);
        R2[threadIdx.x] = rotl_var(nt_seed(3u - nb), 1) ^ nt_seed(3u - ob);
        FD[threadIdx.x] = rotl_var(nt_seed(ob), k) ^ nt_seed(nb);
        RD[threadIdx.x] = rotl_var(nt_seed(3u - ob), 63) ^ rotl_var(nt_seed(3u - nb), k - 1);
        if (threadIdx.x < 4) {
            SEED[threadIdx.x] = nt_seed(threadIdx.x);
            SEEDC[threadIdx.x] = nt_seed(3u - threadIdx.x);
        }
    }
    __syncthreads();
    const uint64_t mult = (uint64_t)k * 0x90b45d39fb6da1faULL;  // nthash.rs:13,68 (wrapping)
    const uint64_t ngroups = (total_bytes + GROUP_BYTES - 1) / GROUP_BYTES;
    const int lane = threadIdx.x & 31;
    uint64_t* tile_h = (uint64_t*)(nt_smem + (threadIdx.x >> 5) * NT_TILE_BYTES);
    uint8_t* tile_s = (uint8_t*)(tile_h + 32 * NT_PITCH);
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t g = warp; g < ngroups; g += nwarps) {
        const uint64_t byte0 = g * GROUP_BYTES;
        const uint64_t byte1 = min(byte0 + (uint64_t)GROUP_BYTES, total_bytes);
        uint64_t s = seq_of_byte(b.byte_off, b.nseq, byte0);
        while (s < b.nseq) {
            const uint64_t sb = __ldg(b.byte_off + s);
            if (sb >= byte1) break;
            const uint64_t L = __ldg(b.nbases + s);
            const uint64_t nk = L >= k ? L - k + 1 : 0;
            const uint64_t p_lo = byte0 > sb ? (byte0 - sb) * 4 : 0;
            const uint64_t p_hi = min(nk, (byte1 - sb) * 4);
            const uint32_t* words = (const uint32_t*)(b.packed + sb);
            const uint64_t obase = __ldg(out_off + s);
            for (uint64_t t0 = p_lo; t0 < p_hi; t0 += 32 * NT_RUN) {  // one tile
                const uint64_t p0 = t0 + (uint64_t)lane * NT_RUN;
                if (p0 < p_hi) {
                    const uint32_t n = (uint32_t)min((uint64_t)NT_RUN, p_hi - p0);
                    // the run needs its first k-mer and, per step, the outgoing base (position p0 + j) and the incoming one
                    // (position p0 + j + k): two 64-bit windows of 32 bases, read once -- no rolling k-mer value
                    const uint32_t* wo = words + (p0 >> 4);
                    const uint32_t sho = (uint32_t)(p0 & 15) * 2;
                    const uint32_t o0 = be32(__ldg(wo)), o1 = be32(__ldg(wo + 1)), o2 = be32(__ldg(wo + 2));
                    const uint64_t OUT = ((uint64_t)__funnelshift_l(o1, o0, sho) << 32) | __funnelshift_l(o2, o1, sho);
                    const uint64_t q0 = p0 + k;
                    const uint32_t* wi = words + (q0 >> 4);
                    const uint32_t shi = (uint32_t)(q0 & 15) * 2;
                    const uint32_t i0 = be32(__ldg(wi)), i1 = be32(__ldg(wi + 1)), i2 = be32(__ldg(wi + 2));
                    const uint64_t IN = ((uint64_t)__funnelshift_l(i1, i0, shi) << 32) | __funnelshift_l(i2, i1, shi);
                    const uint64_t kv = OUT >> (64 - 2 * k);  // the k-mer at p0
                    // nthash_canonical_init (kmer.rs:74-94), Horner form: f = XOR_i rotl(seed[b_i], k-1-i),
                    // r = XOR_i rotl(seed[3 - b_i], i)
                    uint64_t f = 0, r = 0;
                    uint32_t i = 0;
                    if (k & 1) {  // odd k: the first base of each direction alone, then pairs
                        f = SEED[(uint32_t)(kv >> (2 * (k - 1))) & 3u];
                        r = SEEDC[(uint32_t)kv & 3u];
                        i = 1;
                    }
                    for (; i < k; i += 2) {
                        f = ((f << 2) | (f >> 62)) ^ F2[(uint32_t)(kv >> (2 * (k - 2 - i))) & 15u];
                        r = ((r << 2) | (r >> 62)) ^ R2[(uint32_t)(kv >> (2 * i)) & 15u];
                    }
                    uint64_t* th = tile_h + lane * NT_PITCH;
                    uint8_t* ts = tile_s + lane * NT_PITCH;
                    static_assert(NT_RUN == 32, "one 64-bit window per run");
#pragma unroll
                    for (uint32_t j = 0; j < NT_RUN; ++j) {
                        if (j < n) {
                            th[j] = f <= r ? f : r;
                            ts[j] = f <= r ? 0 : 1;
                            // ntHash recurrence: identical values to re-initialising on the new window
                            const uint32_t t = ((uint32_t)(OUT >> (62 - 2 * j)) & 3u) * 4 + ((uint32_t)(IN >> (62 - 2 * j)) & 3u);
                            f = ((f << 1) | (f >> 63)) ^ FD[t];
                            r = ((r >> 1) | (r << 63)) ^ RD[t];
                        }
                    }
                }
                __syncwarp();
                // tile out: run j of the tile = positions t0 + 32 j .. + 31, one per lane
                const uint32_t tile_n = (uint32_t)min((uint64_t)32 * NT_RUN, p_hi - t0);
                for (uint32_t j = 0; j * NT_RUN < tile_n; ++j) {
                    const uint32_t q = j * NT_RUN + lane;
                    if (q < tile_n) {
                        const uint64_t h0 = tile_h[j * NT_PITCH + lane];
                        const uint64_t p = t0 + q;
                        uint64_t* o = out_hash + (obase + p) * n_multi;
                        o[0] = h0;
                        for (uint32_t i = 1; i < n_multi; ++i) {
                            uint64_t tmp = h0 * ((uint64_t)i ^ mult);
                            tmp ^= tmp >> 27;
                            o[i] = tmp;
                        }
                        if (out_strand) out_strand[obase + p] = tile_s[j * NT_PITCH + lane];
                    }
                }
                __syncwarp();
            }
            ++s;
        }
    }
}

cudaError_t launch_nthash(const SeqView& b, uint64_t total_bytes, uint32_t k, uint32_t n_multi, const uint64_t* out_off,
                          uint64_t* out_hash, uint8_t* out_strand, cudaStream_t stream) {
    if (b.nseq == 0) return cudaSuccess;
    const int wpb = 4, cps = 5;  // warps per CTA (one 9.3 KB tile each), CTAs per SM
    const size_t smem = wpb * (size_t)NT_TILE_BYTES;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(nthash_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    nthash_warp_kernel<<<148 * cps, 32 * wpb, smem, stream>>>(b, total_bytes, k, n_multi, out_off, out_hash, out_strand);
    return cudaGetLastError();
}

}  // namespace kmu
