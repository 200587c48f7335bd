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
        uint64_t s = seq_of_byte_warp(b.byte_off, b.nseq, byte0);
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

// Run form of the warp kernel: a lane takes R = 32 / sizeof(V) CONSECUTIVE k-mers (8 u32 / 4 u64) and writes them with ONE
// 256-bit store (st.global.v8.b32, SASS STG.E.ENL2.256), so every store instruction of the warp writes 1 KB of contiguous
// output.  The lane reads the window of packed words its run needs once; ONE reverse complement of the window serves all
// k-mers of the run on both strands, there is no rolling state.  Everything inside a (group, sequence) segment is 32-bit
// arithmetic relative to the segment; runs are aligned on the OUTPUT index (out must be 32-byte aligned), the < 2 R ragged
// elements at the two ends of a segment go through the scalar form.  The issue budget is what bounds this kernel (the
// alu pipe takes shifts / logic / compares at one warp instruction per two cycles per SM sub-partition, the fma pipe the
// multiply-adds at the same rate), hence:
//   * the two `key += ~(key << n)` steps of int32_hash are one multiply-add each (key * (1 - 2^n) - 1, fma pipe) and the
//     header OR is folded into the first one's constant;
//   * PACK16 (2 k <= 16 bits): k-mers t and t + 4 of a run sit exactly one byte apart in the window, one shift + one byte
//     permute packs both into the two halves of a register, and the strand minimum is one 16-bit SIMD instruction.
__device__ __forceinline__ void st256(void* p, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4, uint32_t a5,
                                      uint32_t a6, uint32_t a7) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(a4), "r"(a5),
                 "r"(a6), "r"(a7)
                 : "memory");
}
// reverse complement of 16 bases: one BREV, the complement folded into the pair swap (one three-input logic op)
__device__ __forceinline__ uint32_t revcomp16(uint32_t w) {
    const uint32_t r = __brev(w);
    uint32_t d;  // ~(((r << 1) & 0xAAAAAAAA) | ((r >> 1) & 0x55555555)): lut 0x27 of (a, b, c) = ~(c ? b : a)
    asm("lop3.b32 %0, %1, %2, 0x55555555, 0x27;" : "=r"(d) : "r"(r << 1), "r"(r >> 1));
    return d;
}
// x >> n (0 <= n < 32, n a compile-time constant after unrolling) as a high multiply: the fma pipe's shift
__device__ __forceinline__ uint32_t shr_fma(uint32_t x, uint32_t n) { return n == 0 ? x : __umulhi(x, 1u << (32 - n)); }
// int32_hash(v | header) with c0 = header * 0xFFFF8001 - 1 (v and header share no bit)
__device__ __forceinline__ uint32_t int32_hash_folded(uint32_t v, uint32_t c0) {
    uint32_t key = v * 0xFFFF8001u + c0;  // key + ~(key << 15) = key * (1 - 2^15) - 1
    key ^= key >> 10;
    key *= 9u;
    key ^= key >> 6;
    key = key * 0xFFFFF801u - 1u;  // key + ~(key << 11)
    key ^= key >> 16;  // (the hash's shifts stay on the alu pipe: as high multiplies, all three 787 -> 739 Gbases/s, one 822 -> 811)
    return key;
}
template <typename V, bool HASH>
__device__ __forceinline__ V finish_key(V v, V header, uint32_t c0) {
    if (sizeof(V) == 4) return HASH ? (V)int32_hash_folded((uint32_t)v, c0) : (V)(v | header);
    return HASH ? (V)int64_hash((uint64_t)(v | header)) : (V)(v | header);
}

template <typename V, bool CANON, bool HASH, int PACK16>  // PACK16: 0 no, 1 for 2 k < 16, 2 for 2 k == 16 (no mask)
__global__ void __launch_bounds__(256) generate_kmers_run_kernel(SeqView b, uint64_t total_bytes, uint32_t k, V header,
                                                                  const uint64_t* __restrict__ out_off, V* __restrict__ out) {
    constexpr uint32_t R = 32 / (uint32_t)sizeof(V);
    const V mask = value_mask<V>(2 * k);
    const uint32_t c0 = (uint32_t)header * 0xFFFF8001u - 1u;
    const uint64_t ngroups = (total_bytes + GROUP_BYTES - 1) / GROUP_BYTES;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t g = warp; g < ngroups; g += nwarps) {
        const uint64_t byte0 = g * GROUP_BYTES;
        const uint64_t byte1 = min(byte0 + (uint64_t)GROUP_BYTES, total_bytes);
        uint64_t s = seq_of_byte_warp(b.byte_off, b.nseq, byte0);
        while (s < b.nseq) {
            const uint64_t sb = __ldg(b.byte_off + s);
            if (sb >= byte1) break;
            const uint64_t L = __ldg(b.nbases + s);
            const uint64_t nk = L >= k ? L - k + 1 : 0;
            const uint64_t p_lo = byte0 > sb ? (byte0 - sb) * 4 : 0;
            const uint64_t p_hi = min(nk, (byte1 - sb) * 4);
            ++s;
            if (p_lo >= p_hi) continue;
            // the segment: n k-mers from position p_lo of the sequence, output o[0 .. n)
            const uint32_t n = (uint32_t)(p_hi - p_lo);
            const uint64_t e0 = __ldg(out_off + s - 1) + p_lo;
            V* o = out + e0;
            const uint32_t* wbase = (const uint32_t*)(b.packed + sb) + (p_lo >> 4);
            const uint32_t q0 = (uint32_t)p_lo & 15;
            const uint32_t head = min(n, (R - (uint32_t)(e0 & (R - 1))) & (R - 1));
            const uint32_t nruns = (n - head) / R;
            const uint32_t nscalar = n - nruns * R;  // < 2 R <= 16: the ragged ends, one lane each
            if (lane < nscalar) {
                const uint32_t r = lane < head ? lane : lane + nruns * R;
                V key = kmer_at<V>(wbase, q0 + r, k);
                if (CANON) {
                    const V rc = revcomp_val(key, k);
                    key = key < rc ? key : rc;
                }
                o[r] = finish_key<V, HASH>(key, header, c0);
            }
#pragma unroll 2
            for (uint32_t j = lane; j < nruns; j += 32) {
                const uint32_t r = head + j * R;
                const uint32_t q = q0 + r;
                const uint32_t* w = wbase + (q >> 4);
                const uint32_t sh = (q & 15) * 2;
                if (sizeof(V) == 4 && PACK16) {
                    // 2 k <= 16: the run's 8 k-mers lie in the 16 bases from position p_lo + r (two words).  k-mer t + 4 sits
                    // in bits 0 .. 2k of z = xh >> (24 - 2k - 2t), k-mer t in bits 8 .. 8 + 2k: one byte permute packs them
                    // as (t | t + 4); on strand - (reverse complement of the 16 bases) k-mer t is at bit 2t, t + 4 at bit
                    // 2t + 8.  The right shifts by constants are high multiplies (fma pipe): the alu pipe is the bound.
                    const uint32_t xh = __funnelshift_l(be32(__ldg(w + 1)), be32(__ldg(w)), sh);
                    const uint32_t rl = CANON ? revcomp16(xh) : 0;
                    const uint32_t mask2 = (uint32_t)mask * 0x10001u;
                    uint32_t vals[8];
#pragma unroll
                    for (uint32_t t = 0; t < 4; ++t) {
                        const uint32_t z = PACK16 == 2 ? shr_fma(xh, 8 - 2 * t) : xh >> (24 - 2 * k - 2 * t);
                        uint32_t pk = __byte_perm(z, 0, 0x2110);
                        if (PACK16 == 1) pk &= mask2;
                        if (CANON) {
                            uint32_t pr = __byte_perm(shr_fma(rl, 2 * t), 0, 0x1021);
                            if (PACK16 == 1) pr &= mask2;
                            pk = __vminu2(pk, pr);
                        }
                        const uint32_t hi = shr_fma(pk, 16), lo = pk - (hi << 16);
                        vals[t] = (uint32_t)finish_key<V, HASH>((V)hi, header, c0);
                        vals[t + 4] = (uint32_t)finish_key<V, HASH>((V)lo, header, c0);
                    }
                    st256(o + r, vals[0], vals[1], vals[2], vals[3], vals[4], vals[5], vals[6], vals[7]);
                } else if (sizeof(V) == 4) {
                    // 32 bases from position p_lo + r: the 8 k-mers of the run start at bases 0 .. 7 of the window
                    const uint32_t wa = be32(__ldg(w)), wb = be32(__ldg(w + 1)), wc = be32(__ldg(w + 2));
                    const uint32_t xh = __funnelshift_l(wb, wa, sh), xl = __funnelshift_l(wc, wb, sh);
                    const uint64_t x = ((uint64_t)xh << 32) | xl;
                    const uint64_t rc = CANON ? (((uint64_t)revcomp16(xl) << 32) | revcomp16(xh)) : 0;
                    uint32_t vals[8];
                    {
#pragma unroll
                        for (uint32_t t = 0; t < 8; ++t) {
                            uint32_t key = (uint32_t)(x >> (64 - 2 * k - 2 * t)) & (uint32_t)mask;
                            if (CANON) key = min(key, (uint32_t)(rc >> (2 * t)) & (uint32_t)mask);
                            vals[t] = (uint32_t)finish_key<V, HASH>((V)key, header, c0);
                        }
                    }
                    st256(o + r, vals[0], vals[1], vals[2], vals[3], vals[4], vals[5], vals[6], vals[7]);
                } else {
                    // 48 bases from position p_lo + r in (a0, a1, a2); the window's reverse complement is (r2, r1, r0) read
                    // from the top, so strand - of k-mer t is the window's reverse complement shifted right by 2 t
                    const uint32_t w0 = be32(__ldg(w)), w1 = be32(__ldg(w + 1)), w2 = be32(__ldg(w + 2)), w3 = be32(__ldg(w + 3));
                    const uint32_t a0 = __funnelshift_l(w1, w0, sh), a1 = __funnelshift_l(w2, w1, sh), a2 = __funnelshift_l(w3, w2, sh);
                    const uint32_t r0 = CANON ? revcomp16(a0) : 0, r1 = CANON ? revcomp16(a1) : 0, r2 = CANON ? revcomp16(a2) : 0;
                    uint32_t vals[8];
#pragma unroll
                    for (uint32_t t = 0; t < 4; ++t) {
                        const uint64_t x = ((uint64_t)__funnelshift_l(a1, a0, 2 * t) << 32) | __funnelshift_l(a2, a1, 2 * t);
                        uint64_t key = x >> (64 - 2 * k);
                        if (CANON) {
                            const uint64_t y = (((uint64_t)__funnelshift_r(r1, r2, 2 * t) << 32) | __funnelshift_r(r0, r1, 2 * t)) & (uint64_t)mask;
                            key = key < y ? key : y;
                        }
                        key = (uint64_t)finish_key<V, HASH>((V)key, header, c0);
                        vals[2 * t] = (uint32_t)key;
                        vals[2 * t + 1] = (uint32_t)(key >> 32);
                    }
                    st256(o + r, vals[0], vals[1], vals[2], vals[3], vals[4], vals[5], vals[6], vals[7]);
                }
            }
        }
    }
}

// grid of a grid-stride kernel: exactly the CTAs that are resident at once (a grid of 8 CTAs per SM with 40 registers per
// thread runs as one full wave of 6 per SM plus a second wave at a third of the occupancy: every CTA has the same share
// of the groups, so the second wave costs as much time as the first)
template <typename K>
static int resident_grid(K kernel, int block) {
    int per_sm = 0, dev = 0, sms = 148;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return per_sm * sms;
}

template <typename V>
static void launch_run_kernel(const SeqView& b, uint64_t total_bytes, uint32_t k, int kmer_type, int hash_kind,
                              const uint64_t* out_off, V* out, cudaStream_t stream) {
    const int block = 256;
    const bool canon = hash_is_canonical(hash_kind);
    const bool hashed = hash_kind == KMU_HASH_CANON_INVHASH || hash_kind == KMU_HASH_INVHASH;
    const V header = hash_kind == KMU_HASH_MASKED_VALUE ? (V)0 : (V)word_header(kmer_type, k);
#define KMU_RUN(C, H, P)                                                                                              \
    generate_kmers_run_kernel<V, C, H, P><<<resident_grid(generate_kmers_run_kernel<V, C, H, P>, block), block, 0, stream>>>( \
        b, total_bytes, k, header, out_off, out)
    if constexpr (sizeof(V) == 4) {
        if (k == 8) {
            if (canon) { if (hashed) KMU_RUN(true, true, 2); else KMU_RUN(true, false, 2); }
            else { if (hashed) KMU_RUN(false, true, 2); else KMU_RUN(false, false, 2); }
            return;
        }
        if (k < 8) {
            if (canon) { if (hashed) KMU_RUN(true, true, 1); else KMU_RUN(true, false, 1); }
            else { if (hashed) KMU_RUN(false, true, 1); else KMU_RUN(false, false, 1); }
            return;
        }
    }
    {
        if (canon) { if (hashed) KMU_RUN(true, true, 0); else KMU_RUN(true, false, 0); }
        else { if (hashed) KMU_RUN(false, true, 0); else KMU_RUN(false, false, 0); }
    }
#undef KMU_RUN
}

cudaError_t launch_generate_kmers(const SeqView& b, uint64_t total_bytes, uint32_t k, int kmer_type, int hash_kind,
                                  const uint64_t* out_off, void* out, cudaStream_t stream) {
    if (b.nseq == 0) return cudaSuccess;
    const int block = 256;
    const int grid = 148 * 6;  // 40 registers per thread: 6 CTAs of 256 threads per SM
    if (kmer_type == KMU_KMERAA64)
        generate_kmers_kernel<uint64_t, 8, true><<<grid, block, 0, stream>>>(b, k, kmer_type, hash_kind, out_off, (uint64_t*)out);
    else if (kmer_type == KMU_KMERAA32)
        generate_kmers_kernel<uint32_t, 8, true><<<grid, block, 0, stream>>>(b, k, kmer_type, hash_kind, out_off, (uint32_t*)out);
    else if (kmer_type == KMU_KMER64 && ((uintptr_t)out & 31) == 0)
        launch_run_kernel<uint64_t>(b, total_bytes, k, kmer_type, hash_kind, out_off, (uint64_t*)out, stream);
    else if (kmer_type != KMU_KMER64 && ((uintptr_t)out & 31) == 0)
        launch_run_kernel<uint32_t>(b, total_bytes, k, kmer_type, hash_kind, out_off, (uint32_t*)out, stream);
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
        uint64_t s = seq_of_byte_warp(b.byte_off, b.nseq, byte0);
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

// Run form of ntHash (one hash per k-mer, 32-byte aligned outputs): the k-mers of a (group, sequence) segment are cut
// into 32 contiguous stretches, one per lane, each a multiple of 16 positions (a run) that starts on a multiple of 16 of
// the OUTPUT index.  A lane initialises once (O(k)) and then only rolls; every 4 steps it writes its 4 hashes with one 256-bit
// store (a full 32-byte sector: no shared-memory transposition, no partial sectors), every 16 steps the 16 strand bytes
// with one 128-bit store.  The elements before the first aligned index of a segment and behind its last whole run (< 32
// together) are computed one per lane by initialisation alone.  The search for the group's first sequence is the warp's
// (seq_of_byte_warp); the table index of every step comes from two nibble words built once per run.
struct NtTables {
    uint64_t SEED[4], SEEDC[4], FD[16], RD[16], F2[16], R2[16];
};
__device__ __forceinline__ void nt_tables_fill(NtTables& T, uint32_t k) {
    if (threadIdx.x < 16) {
        const uint32_t ob = threadIdx.x >> 2, nb = threadIdx.x & 3;
        T.F2[threadIdx.x] = rotl_var(nt_seed(ob), 1) ^ nt_seed(nb);
        T.R2[threadIdx.x] = rotl_var(nt_seed(3u - nb), 1) ^ nt_seed(3u - ob);
        T.FD[threadIdx.x] = rotl_var(nt_seed(ob), k) ^ nt_seed(nb);
        T.RD[threadIdx.x] = rotl_var(nt_seed(3u - ob), 63) ^ rotl_var(nt_seed(3u - nb), k - 1);
        if (threadIdx.x < 4) {
            T.SEED[threadIdx.x] = nt_seed(threadIdx.x);
            T.SEEDC[threadIdx.x] = nt_seed(3u - threadIdx.x);
        }
    }
}
// 32 bases from position q of the word stream w
__device__ __forceinline__ uint64_t window32(const uint32_t* __restrict__ w, uint32_t q) {
    const uint32_t* a = w + (q >> 4);
    const uint32_t sh = (q & 15) * 2;
    const uint32_t x0 = be32(__ldg(a)), x1 = be32(__ldg(a + 1)), x2 = be32(__ldg(a + 2));
    return ((uint64_t)__funnelshift_l(x1, x0, sh) << 32) | __funnelshift_l(x2, x1, sh);
}
// nthash_canonical_init (kmer.rs:74-94), Horner form, two bases per step
__device__ __forceinline__ void nt_init(const NtTables& T, uint64_t kv, uint32_t k, uint64_t& f, uint64_t& r) {
    f = 0;
    r = 0;
    uint32_t i = 0;
    if (k & 1) {
        f = T.SEED[(uint32_t)(kv >> (2 * (k - 1))) & 3u];
        r = T.SEEDC[(uint32_t)kv & 3u];
        i = 1;
    }
    for (; i < k; i += 2) {
        f = ((f << 2) | (f >> 62)) ^ T.F2[(uint32_t)(kv >> (2 * (k - 2 - i))) & 15u];
        r = ((r << 2) | (r >> 62)) ^ T.R2[(uint32_t)(kv >> (2 * i)) & 15u];
    }
}
__device__ __forceinline__ void st128(void* p, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3) {
    asm volatile("st.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(a0), "r"(a1), "r"(a2), "r"(a3) : "memory");
}

// 16 bases from position q of the word stream w
__device__ __forceinline__ uint32_t window16(const uint32_t* __restrict__ w, uint32_t q) {
    const uint32_t* a = w + (q >> 4);
    return __funnelshift_l(be32(__ldg(a + 1)), be32(__ldg(a)), (q & 15) * 2);
}
// c ? b : a, bit by bit (one three-input logic op)
__device__ __forceinline__ uint32_t bitsel(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xD8;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

constexpr uint32_t NT_STEPS = 16;  // positions per run: the granularity of a lane's stretch
template <bool STRAND>
__global__ void __launch_bounds__(128) nthash_run_kernel(SeqView b, uint64_t total_bytes, uint32_t k,
                                                          const uint64_t* __restrict__ out_off, uint64_t* __restrict__ out_hash,
                                                          uint8_t* __restrict__ out_strand) {
    __shared__ NtTables T;
    nt_tables_fill(T, k);
    __syncthreads();
    const uint64_t ngroups = (total_bytes + GROUP_BYTES - 1) / GROUP_BYTES;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t g = warp; g < ngroups; g += nwarps) {
        const uint64_t byte0 = g * GROUP_BYTES;
        const uint64_t byte1 = min(byte0 + (uint64_t)GROUP_BYTES, total_bytes);
        uint64_t s = seq_of_byte_warp(b.byte_off, b.nseq, byte0);
        while (s < b.nseq) {
            const uint64_t sb = __ldg(b.byte_off + s);
            if (sb >= byte1) break;
            const uint64_t L = __ldg(b.nbases + s);
            const uint64_t nk = L >= k ? L - k + 1 : 0;
            const uint64_t p_lo = byte0 > sb ? (byte0 - sb) * 4 : 0;
            const uint64_t p_hi = min(nk, (byte1 - sb) * 4);
            ++s;
            if (p_lo >= p_hi) continue;
            const uint32_t n = (uint32_t)(p_hi - p_lo);
            const uint64_t e0 = __ldg(out_off + s - 1) + p_lo;
            uint64_t* oh = out_hash + e0;
            uint8_t* os = out_strand + e0;
            const uint32_t* wbase = (const uint32_t*)(b.packed + sb) + (p_lo >> 4);
            const uint32_t q0 = (uint32_t)p_lo & 15;
            // [0, head) and [head + 16 nruns, n): fewer than 32 positions, one per lane by initialisation alone;
            // the nruns runs of 16 in between: lane l takes runs [l nruns / 32, (l + 1) nruns / 32)
            const uint32_t head = min(n, (NT_STEPS - (uint32_t)(e0 & (NT_STEPS - 1))) & (NT_STEPS - 1));
            const uint32_t nruns = (n - head) / NT_STEPS;
            const uint32_t nscalar = n - nruns * NT_STEPS;
            const uint32_t run0 = (lane * nruns) >> 5, run1 = ((lane + 1) * nruns) >> 5;
            // the first window of the lane's stretch is requested before the scalar positions are worked on
            uint64_t w_run = 0;
            if (run0 < run1) w_run = window32(wbase, q0 + head + run0 * NT_STEPS);
            if (lane < nscalar) {
                const uint32_t q = lane < head ? lane : lane + nruns * NT_STEPS;
                uint64_t f, r;
                nt_init(T, window32(wbase, q0 + q) >> (64 - 2 * k), k, f, r);
                const bool rev = r < f;
                oh[q] = rev ? r : f;
                if (STRAND) os[q] = rev ? 1 : 0;
            }
            if (run0 < run1) {
                uint32_t q = head + run0 * NT_STEPS;
                const uint32_t q_end = head + run1 * NT_STEPS;
                uint64_t f, r;
                nt_init(T, w_run >> (64 - 2 * k), k, f, r);
                // outgoing bases (positions q + j) and incoming ones (q + j + k); the windows of the next run are loaded
                // before this run's steps (the read past the stretch stays inside the batch: 64 bytes of slack)
                uint32_t OUTn = window16(wbase, q0 + q), INn = window16(wbase, q0 + q + k);
                for (; q < q_end; q += NT_STEPS) {
                    const uint32_t OUT = OUTn, IN = INn;
                    OUTn = window16(wbase, q0 + q + NT_STEPS);
                    INn = window16(wbase, q0 + q + NT_STEPS + k);
                    uint32_t hv[8], sv[4] = {0, 0, 0, 0};
                    // table index (outgoing base, incoming base) of every step as a nibble: the odd steps in TO, the even
                    // ones in TE, so that a step costs one shift and one mask instead of two of each
                    const uint32_t TO = bitsel(IN, OUT << 2, 0xCCCCCCCCu), TE = bitsel(IN >> 2, OUT, 0xCCCCCCCCu);
#pragma unroll
                    for (uint32_t j = 0; j < NT_STEPS; ++j) {
                        const uint32_t rev = r < f ? 0xFFFFFFFFu : 0u;
                        hv[2 * (j & 3)] = bitsel((uint32_t)f, (uint32_t)r, rev);
                        hv[2 * (j & 3) + 1] = bitsel((uint32_t)(f >> 32), (uint32_t)(r >> 32), rev);
                        if (STRAND) sv[j >> 2] |= rev & (1u << (8 * (j & 3)));
                        if ((j & 3) == 3) st256(oh + q + j - 3, hv[0], hv[1], hv[2], hv[3], hv[4], hv[5], hv[6], hv[7]);
                        const uint32_t W = (j & 1) ? TO : TE, nib = (j & 1) ? (15 - j) / 2 : (14 - j) / 2;
                        const uint32_t off = (nib ? W >> (4 * nib - 3) : W << 3) & 0x78u;  // 8 t, a byte offset
                        f = ((f << 1) | (f >> 63)) ^ *(const uint64_t*)((const uint8_t*)T.FD + off);
                        r = ((r >> 1) | (r << 63)) ^ *(const uint64_t*)((const uint8_t*)T.RD + off);
                    }
                    if (STRAND) st128(os + q, sv[0], sv[1], sv[2], sv[3]);
                }
            }
            __syncwarp();
        }
    }
}

cudaError_t launch_nthash(const SeqView& b, uint64_t total_bytes, uint32_t k, uint32_t n_multi, const uint64_t* out_off,
                          uint64_t* out_hash, uint8_t* out_strand, cudaStream_t stream) {
    if (b.nseq == 0) return cudaSuccess;
    if (n_multi == 1 && ((uintptr_t)out_hash & 31) == 0 && ((uintptr_t)out_strand & 15) == 0) {
        if (out_strand)
            nthash_run_kernel<true><<<resident_grid(nthash_run_kernel<true>, 128), 128, 0, stream>>>(b, total_bytes, k, out_off, out_hash, out_strand);
        else
            nthash_run_kernel<false><<<resident_grid(nthash_run_kernel<false>, 128), 128, 0, stream>>>(b, total_bytes, k, out_off, out_hash, out_strand);
        return cudaGetLastError();
    }
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
