// kmu_batch.cu -- kernels that build sequence batches in HBM: ASCII -> 2-bit packing
// (Alphabet2b / Sequence::new / Sequence::encode_and_add), the synthetic read
// generator of the benchmark, and the length-class ordering used to schedule
// sequences longest-first.
#include <cstdint>

#include "kmu_device.cuh"
#include "kmu_kernels.h"

namespace kmu {

// Alphabet2b::encode (src/base/alphabet.rs:119-127): case-insensitive A0 C1 G2 T3, else invalid (4)
__device__ __forceinline__ uint32_t encode2b(uint8_t c) {
    if (c >= 'a' && c <= 'z') c -= 32;
    return c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 4u;
}

// ---- synthetic reads -----------------------------------------------------------------------
// one thread per 32-bit word (16 bases); word index -> sequence by binary search on byte_off
__global__ void synth_packed_kernel(uint8_t* packed, const uint64_t* __restrict__ byte_off,
                                    const uint64_t* __restrict__ nbases, const uint64_t* __restrict__ first_base,
                                    uint64_t nseq, uint64_t total_words, uint64_t seed) {
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < total_words;
         w += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t byte = w * 4;
        // last sequence whose byte_off <= byte
        uint64_t lo = 0, hi = nseq;
        while (hi - lo > 1) {
            uint64_t mid = (lo + hi) >> 1;
            if (byte_off[mid] <= byte) lo = mid; else hi = mid;
        }
        uint64_t p0 = (byte - byte_off[lo]) * 4;  // first base of this word inside the sequence
        uint64_t L = nbases[lo];
        uint64_t fb = first_base[lo];
        uint32_t word = 0;  // big-endian base order
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            uint64_t p = p0 + i;
            uint32_t code = p < L ? (uint32_t)(synth_z(seed, fb + p) >> 62) : 0u;
            word |= code << (30 - 2 * i);
        }
        ((uint32_t*)packed)[w] = __byte_perm(word, 0, 0x0123);
    }
}

cudaError_t launch_synth_packed(uint8_t* packed, const uint64_t* byte_off, const uint64_t* nbases,
                                const uint64_t* first_base, uint64_t nseq, uint64_t total_words, uint64_t seed,
                                cudaStream_t stream) {
    if (total_words == 0 || nseq == 0) return cudaSuccess;
    int block = 256;
    uint64_t want = (total_words + block - 1) / block;
    int grid = (int)(want < 148ull * 16 ? want : 148ull * 16);
    synth_packed_kernel<<<grid, block, 0, stream>>>(packed, byte_off, nbases, first_base, nseq, total_words, seed);
    return cudaGetLastError();
}

// ---- count_non_acgt (alphabet.rs:28-31): one warp per sequence --------------------------------
__global__ void count_invalid_kernel(const uint8_t* __restrict__ ascii, const uint64_t* __restrict__ ascii_off,
                                     uint64_t nseq, uint64_t* invalid) {
    uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    int lane = threadIdx.x & 31;
    for (uint64_t s = warp; s < nseq; s += nwarps) {
        uint64_t b = ascii_off[s], e = ascii_off[s + 1];
        uint32_t bad = 0;
        for (uint64_t i = b + lane; i < e; i += 32) bad += encode2b(ascii[i]) > 3;
        bad = __reduce_add_sync(0xFFFFFFFFu, bad);
        if (lane == 0) invalid[s] = bad;
    }
}

cudaError_t launch_count_invalid(const uint8_t* ascii, const uint64_t* ascii_off, uint64_t nseq, uint64_t* invalid,
                                 cudaStream_t stream) {
    if (nseq == 0) return cudaSuccess;
    int block = 256;
    uint64_t want = (nseq * 32 + block - 1) / block;
    int grid = (int)(want < 148ull * 8 ? want : 148ull * 8);
    count_invalid_kernel<<<grid, block, 0, stream>>>(ascii, ascii_off, nseq, invalid);
    return cudaGetLastError();
}

// ---- ASCII -> packed: one warp per sequence; 32 characters per step; kept bases are compacted
// with a ballot so that dropping invalid characters (Sequence::encode_and_add, sequence.rs:388-451)
// and strict packing (Sequence::new, sequence.rs:25-106) share one kernel.  The destination bytes
// are zero-filled first (tail padding = 'A', sequence.rs:66-71) and OR-ed into.
__global__ void pack_ascii_kernel(const uint8_t* __restrict__ ascii, const uint64_t* __restrict__ ascii_off,
                                  const uint64_t* __restrict__ byte_off, const uint64_t* __restrict__ nbases,
                                  uint64_t nseq, int drop_invalid, uint8_t* packed) {
    uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    int lane = threadIdx.x & 31;
    for (uint64_t s = warp; s < nseq; s += nwarps) {
        uint64_t b = ascii_off[s], e = ascii_off[s + 1];
        uint8_t* dst = packed + byte_off[s];
        uint64_t kept = 0;  // warp-uniform
        (void)nbases;
        for (uint64_t i0 = b; i0 < e; i0 += 32) {
            uint64_t i = i0 + lane;
            uint32_t code = i < e ? encode2b(ascii[i]) : 5u;
            bool keep = drop_invalid ? code <= 3 : code <= 4;
            if (code == 4) code = 0;  // strict mode only reaches here after validation failed: pack as 'A'
            uint32_t bal = __ballot_sync(0xFFFFFFFFu, keep);
            if (keep) {
                uint64_t pos = kept + __popc(bal & ((1u << lane) - 1));
                // 16 bases per aligned u32 (little-endian word of four big-endian-ordered bytes)
                uint32_t byte_in_word = (uint32_t)((pos >> 2) & 3);
                uint32_t shift = byte_in_word * 8 + (6 - 2 * (uint32_t)(pos & 3));
                atomicOr((uint32_t*)dst + (pos >> 4), code << shift);
            }
            kept += __popc(bal);
        }
    }
}

cudaError_t launch_pack_ascii(const uint8_t* ascii, const uint64_t* ascii_off, const uint64_t* byte_off,
                              const uint64_t* nbases, uint64_t nseq, int drop_invalid, uint8_t* packed,
                              cudaStream_t stream) {
    if (nseq == 0) return cudaSuccess;
    int block = 256;
    uint64_t want = (nseq * 32 + block - 1) / block;
    int grid = (int)(want < 148ull * 8 ? want : 148ull * 8);
    pack_ascii_kernel<<<grid, block, 0, stream>>>(ascii, ascii_off, byte_off, nbases, nseq, drop_invalid, packed);
    return cudaGetLastError();
}

// ---- amino-acid sequences (src/aautils/kmeraa.rs) ---------------------------------------------
// Alphabet::encode (kmeraa.rs:85-109): 20 upper-case residues -> 5-bit codes (14 is skipped); 0 = invalid
__device__ __forceinline__ uint32_t encode_aa(uint8_t c) {
    switch (c) {
        case 'A': return 1; case 'C': return 2; case 'D': return 3; case 'E': return 4; case 'F': return 5;
        case 'G': return 6; case 'H': return 7; case 'I': return 8; case 'K': return 9; case 'L': return 10;
        case 'M': return 11; case 'N': return 12; case 'P': return 13; case 'Q': return 15; case 'R': return 16;
        case 'S': return 17; case 'T': return 18; case 'V': return 19; case 'W': return 20; case 'Y': return 21;
        default: return 0;
    }
}

__global__ void aa_count_invalid_kernel(const uint8_t* __restrict__ ascii, const uint64_t* __restrict__ ascii_off,
                                        uint64_t nseq, uint64_t* invalid) {
    uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    int lane = threadIdx.x & 31;
    for (uint64_t s = warp; s < nseq; s += nwarps) {
        uint64_t b = ascii_off[s], e = ascii_off[s + 1];
        uint32_t bad = 0;
        for (uint64_t i = b + lane; i < e; i += 32) bad += encode_aa(ascii[i]) == 0;
        bad = __reduce_add_sync(0xFFFFFFFFu, bad);
        if (lane == 0) invalid[s] = bad;
    }
}

// one warp per sequence; valid residues are compacted (SequenceAA::new_filtered, kmeraa.rs:447-456) and
// stored as one code per byte
__global__ void aa_encode_kernel(const uint8_t* __restrict__ ascii, const uint64_t* __restrict__ ascii_off,
                                 const uint64_t* __restrict__ byte_off, uint64_t nseq, uint8_t* codes) {
    uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    int lane = threadIdx.x & 31;
    for (uint64_t s = warp; s < nseq; s += nwarps) {
        uint64_t b = ascii_off[s], e = ascii_off[s + 1];
        uint8_t* dst = codes + byte_off[s];
        uint64_t kept = 0;
        for (uint64_t i0 = b; i0 < e; i0 += 32) {
            uint64_t i = i0 + lane;
            uint32_t code = i < e ? encode_aa(ascii[i]) : 0u;
            uint32_t bal = __ballot_sync(0xFFFFFFFFu, code != 0);
            if (code) dst[kept + __popc(bal & ((1u << lane) - 1))] = (uint8_t)code;
            kept += __popc(bal);
        }
    }
}

cudaError_t launch_aa_count_invalid(const uint8_t* ascii, const uint64_t* ascii_off, uint64_t nseq, uint64_t* invalid,
                                    cudaStream_t stream) {
    if (nseq == 0) return cudaSuccess;
    uint64_t want = (nseq * 32 + 255) / 256;
    int grid = (int)(want < 148ull * 8 ? want : 148ull * 8);
    aa_count_invalid_kernel<<<grid, 256, 0, stream>>>(ascii, ascii_off, nseq, invalid);
    return cudaGetLastError();
}

cudaError_t launch_aa_encode(const uint8_t* ascii, const uint64_t* ascii_off, const uint64_t* byte_off, uint64_t nseq,
                             uint8_t* codes, cudaStream_t stream) {
    if (nseq == 0) return cudaSuccess;
    uint64_t want = (nseq * 32 + 255) / 256;
    int grid = (int)(want < 148ull * 8 ? want : 148ull * 8);
    aa_encode_kernel<<<grid, 256, 0, stream>>>(ascii, ascii_off, byte_off, nseq, codes);
    return cudaGetLastError();
}

// synthetic proteins: residue j of sequence i = "ACDEFGHIKLMNPQRSTVWY"[z % 20],
// z = SplitMix64 output first_res[i] + j of stream `seed`
__global__ void synth_aa_kernel(uint8_t* codes, const uint64_t* __restrict__ byte_off, const uint64_t* __restrict__ nres,
                                const uint64_t* __restrict__ first_res, uint64_t nseq, uint64_t total_bytes,
                                uint64_t seed) {
    for (uint64_t byte = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; byte < total_bytes;
         byte += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t lo = 0, hi = nseq;
        while (hi - lo > 1) {
            uint64_t mid = (lo + hi) >> 1;
            if (byte_off[mid] <= byte) lo = mid; else hi = mid;
        }
        const uint64_t p = byte - byte_off[lo];
        uint32_t code = 0;
        if (p < nres[lo]) {
            const uint32_t r = (uint32_t)(synth_z(seed, first_res[lo] + p) % 20u);
            code = r < 13 ? r + 1 : r + 2;  // codes 1..13, 15..21
        }
        codes[byte] = (uint8_t)code;
    }
}

cudaError_t launch_synth_aa(uint8_t* codes, const uint64_t* byte_off, const uint64_t* nres, const uint64_t* first_res,
                            uint64_t nseq, uint64_t total_bytes, uint64_t seed, cudaStream_t stream) {
    if (total_bytes == 0 || nseq == 0) return cudaSuccess;
    uint64_t want = (total_bytes + 255) / 256;
    int grid = (int)(want < 148ull * 16 ? want : 148ull * 16);
    synth_aa_kernel<<<grid, 256, 0, stream>>>(codes, byte_off, nres, first_res, nseq, total_bytes, seed);
    return cudaGetLastError();
}

// ---- reads sampled from a genome ----------------------------------------
// read r: start = z(2r) % (G - len + 1), strand = z(2r + 1) & 1 (1: reverse complement),
// base j substituted when e = z'(r * len + j) has e % 1e6 < err_ppm, by (base + 1 + (e >> 32) % 3) & 3;
// z = SplitMix64 stream `seed`, z' = stream `seed ^ 0x5bd1e995a5a5a5a5`.
__device__ __forceinline__ uint32_t genome_base(const uint8_t* __restrict__ g, uint64_t pos) {
    return (g[pos >> 2] >> (6 - 2 * (pos & 3))) & 3u;
}

__global__ void sample_reads_kernel(const uint8_t* __restrict__ genome, uint64_t glen, uint64_t seed, uint64_t first_read,
                                    uint64_t nreads, uint32_t read_len, uint32_t err_ppm, uint32_t words_per_read,
                                    uint8_t* out) {
    const uint64_t total_words = nreads * words_per_read;
    const uint64_t eseed = seed ^ 0x5bd1e995a5a5a5a5ULL;
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < total_words;
         w += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = w / words_per_read;
        const uint32_t wi = (uint32_t)(w - r * words_per_read);
        const uint64_t gr = first_read + r;
        const uint64_t start = synth_z(seed, 2 * gr) % (glen - read_len + 1);
        const bool rev = synth_z(seed, 2 * gr + 1) & 1;
        uint32_t word = 0;
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
            const uint32_t j = wi * 16 + i;
            uint32_t code = 0;
            if (j < read_len) {
                code = rev ? 3u - genome_base(genome, start + read_len - 1 - j) : genome_base(genome, start + j);
                const uint64_t e = synth_z(eseed, gr * read_len + j);
                if ((uint32_t)(e % 1000000ULL) < err_ppm) code = (code + 1u + (uint32_t)((e >> 32) % 3u)) & 3u;
            }
            word |= code << (30 - 2 * i);
        }
        ((uint32_t*)out)[w] = __byte_perm(word, 0, 0x0123);
    }
}

cudaError_t launch_sample_reads(const uint8_t* genome, uint64_t glen, uint64_t seed, uint64_t first_read, uint64_t nreads,
                                uint32_t read_len, uint32_t err_ppm, uint32_t words_per_read, uint8_t* out,
                                cudaStream_t stream) {
    if (nreads == 0) return cudaSuccess;
    const uint64_t total_words = nreads * words_per_read;
    uint64_t want = (total_words + 255) / 256;
    int grid = (int)(want < 148ull * 16 ? want : 148ull * 16);
    sample_reads_kernel<<<grid, 256, 0, stream>>>(genome, glen, seed, first_read, nreads, read_len, err_ppm,
                                                  words_per_read, out);
    return cudaGetLastError();
}

// ---- slices: sub-ranges of sequences as a new batch -------------------------------------------
// (blocks of BlockSeqSketcher, src/sketching/seqblocksketch.rs:97-149; ranges of
//  sketch_seqrange_superminhash, src/sketching/seqminhash.rs:19-62; KmerSeqIterator::set_range)
// one thread per destination 32-bit word (DNA: 16 bases re-aligned with a funnel shift; amino acids: 4 codes)
__global__ void slice_copy_kernel(const uint8_t* __restrict__ src, const uint64_t* __restrict__ src_byte_off,
                                  const uint64_t* __restrict__ seq_idx, const uint64_t* __restrict__ begin,
                                  const uint64_t* __restrict__ dst_byte_off, const uint64_t* __restrict__ dst_len,
                                  uint64_t nslices, uint64_t total_words, int alphabet, uint8_t* dst) {
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < total_words;
         w += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t byte = w * 4;
        uint64_t lo = 0, hi = nslices;
        while (hi - lo > 1) {
            uint64_t mid = (lo + hi) >> 1;
            if (dst_byte_off[mid] <= byte) lo = mid; else hi = mid;
        }
        const uint64_t len = dst_len[lo];
        const uint8_t* sp = src + src_byte_off[seq_idx[lo]];
        uint32_t out = 0;
        if (alphabet) {
            const uint64_t p0 = byte - dst_byte_off[lo];
            for (int i = 0; i < 4; ++i)
                if (p0 + i < len) out |= (uint32_t)sp[begin[lo] + p0 + i] << (8 * i);
            ((uint32_t*)dst)[w] = out;
        } else {
            const uint64_t p0 = (byte - dst_byte_off[lo]) * 4;  // first base of this word inside the slice
            if (p0 < len) {
                const uint64_t q = begin[lo] + p0;
                const uint32_t* sw = (const uint32_t*)sp + (q >> 4);
                const uint32_t sh = (uint32_t)(q & 15) * 2;
                uint32_t v = be32(sw[0]);
                if (sh) v = (v << sh) | (be32(sw[1]) >> (32 - sh));
                const uint64_t left = len - p0;
                if (left < 16) v &= ~0u << (32 - 2 * (uint32_t)left);  // pad with 'A' like Sequence::new
                out = v;
            }
            ((uint32_t*)dst)[w] = __byte_perm(out, 0, 0x0123);
        }
    }
}

cudaError_t launch_slice_copy(const uint8_t* src, const uint64_t* src_byte_off, const uint64_t* seq_idx, const uint64_t* begin,
                              const uint64_t* dst_byte_off, const uint64_t* dst_len, uint64_t nslices, uint64_t total_words,
                              int alphabet, uint8_t* dst, cudaStream_t stream) {
    if (total_words == 0 || nslices == 0) return cudaSuccess;
    uint64_t want = (total_words + 255) / 256;
    int grid = (int)(want < 148ull * 16 ? want : 148ull * 16);
    slice_copy_kernel<<<grid, 256, 0, stream>>>(src, src_byte_off, seq_idx, begin, dst_byte_off, dst_len, nslices, total_words,
                                                alphabet, dst);
    return cudaGetLastError();
}

// ---- length classes ---------------------------------------------------------------------------
__device__ __forceinline__ int len_bucket(uint64_t nk) {
    if (nk == 0) return LEN_BUCKETS - 1;
    int e = 63 - __clzll((long long)nk);
    int frac = e >= 3 ? (int)((nk >> (e - 3)) & 7) : (int)((nk << (3 - e)) & 7);
    return LEN_BUCKETS - 1 - (e * 8 + frac);
}

__global__ void len_hist_kernel(const uint64_t* __restrict__ nbases, uint64_t nseq, uint32_t k,
                                unsigned long long* hist) {
    __shared__ unsigned int sh[LEN_BUCKETS];
    for (int i = threadIdx.x; i < LEN_BUCKETS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < nseq; s += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t L = nbases[s];
        atomicAdd(&sh[len_bucket(L >= k ? L - k + 1 : 0)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < LEN_BUCKETS; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}

__global__ void len_scatter_kernel(const uint64_t* __restrict__ nbases, uint64_t nseq, uint32_t k,
                                   unsigned long long* cursor, uint32_t* order) {
    for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < nseq; s += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t L = nbases[s];
        unsigned long long pos = atomicAdd(&cursor[len_bucket(L >= k ? L - k + 1 : 0)], 1ULL);
        order[pos] = (uint32_t)s;
    }
}

cudaError_t launch_len_hist(const uint64_t* nbases, uint64_t nseq, uint32_t k, unsigned long long* hist,
                            cudaStream_t stream) {
    if (nseq == 0) return cudaSuccess;
    int block = 256;
    uint64_t want = (nseq + block - 1) / block;
    int grid = (int)(want < 148ull * 4 ? want : 148ull * 4);
    len_hist_kernel<<<grid, block, 0, stream>>>(nbases, nseq, k, hist);
    return cudaGetLastError();
}

cudaError_t launch_len_scatter(const uint64_t* nbases, uint64_t nseq, uint32_t k, unsigned long long* cursor,
                               uint32_t* order, cudaStream_t stream) {
    if (nseq == 0) return cudaSuccess;
    int block = 256;
    uint64_t want = (nseq + block - 1) / block;
    int grid = (int)(want < 148ull * 4 ? want : 148ull * 4);
    len_scatter_kernel<<<grid, block, 0, stream>>>(nbases, nseq, k, cursor, order);
    return cudaGetLastError();
}

}  // namespace kmu
