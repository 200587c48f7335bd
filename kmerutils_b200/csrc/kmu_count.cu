// kmu_count.cu -- exact k-mer multiplicity counting on sm_100a.
//
// Replaces KmerCounter (cuckoo filter for "seen once" + counting Bloom filter for "seen at least
// twice", src/base/kmercount.rs:70-83,241-288) and its drivers count_kmer /
// count_kmer_threaded_one_to_many (:293-362, :881-974) by ONE open-addressing table in HBM:
// the key is kmer.get_compressed_value() of the canonical k-mer (kmer.reverse_complement().min(kmer),
// :313,827,938), claimed with atomicCAS and counted with atomicAdd.  get_count / nb_distinct /
// nb_unique are what the reference returns when its filters have no false positive.
//
// Layout (see DESIGN.md "counting table"):
//   u32 k-mer types (Kmer32bit, Kmer16b32bit): 8-byte slot  key << 32 | count ; 0 == empty
//   u64 k-mer type  (Kmer64bit)              : 16-byte slot {key, count}     ; key == ~0 == empty
//     (the one key equal to the sentinel, 32 T's inserted non-canonically, is counted apart)
// Work decomposition: the packed batch is cut into 64-byte chunks (256 bases); a thread owns the
// k-mers that START in its chunk, finds the sequence(s) overlapping it by binary search in the
// batch's byte offsets and rolls forward / reverse-complement windows in registers.
#include <cstdint>

#include "kmu_count_ops.cuh"
#include "kmu_device.cuh"
#include "kmu_kernels.h"

namespace kmu {

__global__ void count_init_kernel(CountTable t, int key64) {
    const uint64_t n = t.capmask + 1;
    if (key64) {
        ulonglong2* s = (ulonglong2*)t.slots;
        for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
            s[i] = make_ulonglong2(~0ULL, 0ULL);
    } else {
        unsigned long long* s = (unsigned long long*)t.slots;
        for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
            s[i] = 0ULL;
    }
}

template <typename V>
__global__ void __launch_bounds__(256) count_insert_seqs_kernel(SeqView b, uint64_t total_bytes, uint32_t k, int canonical,
                                                                 CountTable t, uint32_t group_bytes, uint64_t first_group) {
    const uint64_t ngroups = (total_bytes + group_bytes - 1) / group_bytes;
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    bool ok = true;
    for (uint64_t g = first_group + warp; g < ngroups; g += nwarps)
        warp_for_each_kmer<V>(
            b, total_bytes, g, k, canonical != 0, lane,
            [&](V key, bool active) {
                if (active) ok &= CountOps<V>::insert(t, key, 1u);
            },
            group_bytes);
    if (!ok) *t.overflow = 1ULL;
}

// ---- whole-file ProbMinHash3a: only the k-mers that can place a point below the bound are counted ------------------
// The item kernel cuts every item at the bound B (kmu_pmh3a_items.cu) and the host verifies that every slot ended below
// B: an item of weight w matters only if its first point x1 / w is below B, and x1 is a function of the key alone (half a
// seeding, first_point_alive).  For a 5 Mb genome and m = 12 000, B = 0.067: 93 % of the k-mers are items of weight 1
// with x1 >= B -- they cannot matter, yet counting them exactly is what the per-genome table (64 MB, beyond what stays
// in L2) and its 5 M claims are for.  Two walks over the sequences instead:
//   PASS 0  a k-mer with x1 < B goes into the table (exact count, as before); any other sets its "seen" bit in the filter, or,
//           when that bit was already set (a second occurrence -- or a collision), its "seen again" bit;
//   PASS 1  a k-mer with x1 >= B whose "seen again" bit is set goes into the table: every key that occurs twice or more is
//           counted exactly (all its occurrences come here), colliding single keys too (harmless).
// What is left out are exactly keys of weight 1 with x1 >= B.  The table holds ~(B + repeats + collisions) of the k-mers.
template <typename V, int PASS>
__global__ void __launch_bounds__(256) pmh3a_prefilter_kernel(SeqView b, uint64_t total_bytes, uint32_t k, int canonical, int kmer_type,
                                                               int hash_kind, double bound, double c1, CountTable t,
                                                               uint32_t* __restrict__ seen, uint64_t bitmask, uint32_t group_bytes) {
    const uint64_t ngroups = (total_bytes + group_bytes - 1) / group_bytes;
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const V header = (V)word_header(kmer_type, k);
    bool ok = true;
    for (uint64_t g = warp; g < ngroups; g += nwarps)
        warp_for_each_kmer<V>(
            b, total_bytes, g, k, canonical != 0, lane,
            [&](V key, bool active) {
                if (!active) return;
                // two bits per position of the filter, side by side: "seen" and "seen again"
                const uint64_t h = (fmix64((uint64_t)key) >> 24) & bitmask;  // (the table index uses the low bits of the same hash)
                uint32_t* word = seen + (h >> 4);
                const uint32_t once = 1u << ((h & 15) * 2), again = once << 1;
                if (PASS == 0) {
                    if (first_point_alive<V>(finalize_key<V>(key, header, hash_kind), 1.0, bound, c1)) {
                        ok &= CountOps<V>::insert(t, key, 1u);
                    } else if (atomicOr(word, once) & once) {
                        atomicOr(word, again);
                    }
                } else if ((__ldcg(word) & again) && !first_point_alive<V>(finalize_key<V>(key, header, hash_kind), 1.0, bound, c1)) {
                    ok &= CountOps<V>::insert(t, key, 1u);  // (a key below the bound went in with all its occurrences in pass 0)
                }
            },
            group_bytes);
    if (!ok) *t.overflow = 1ULL;
}

template <typename V>
__global__ void __launch_bounds__(256) count_insert_keys_kernel(const V* __restrict__ keys, uint64_t n, CountTable t) {
    bool ok = true;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        ok &= CountOps<V>::insert(t, keys[i], 1u);
    if (!ok) *t.overflow = 1ULL;
}

template <typename V>
__global__ void __launch_bounds__(256) count_query_kernel(const V* __restrict__ keys, uint64_t n, CountTable t,
                                                           uint32_t max_count, uint32_t* __restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t c = CountOps<V>::lookup(t, keys[i]);
        out[i] = c < max_count ? c : max_count;
    }
}

// stats[0] = distinct, stats[1] = unique, stats[2] = total multiplicity, stats[3 + c] = #keys with min(count, 255) == c
template <typename V>
__global__ void __launch_bounds__(256) count_stats_kernel(CountTable t, unsigned long long* stats) {
    __shared__ unsigned int hist[256];
    __shared__ unsigned long long tot;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    if (threadIdx.x == 0) tot = 0;
    __syncthreads();
    const uint64_t n = t.capmask + 1;
    unsigned long long mytot = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t key, cnt;
        if (CountOps<V>::occupied(t, i, key, cnt)) {
            atomicAdd(&hist[cnt < 255 ? (uint32_t)cnt : 255u], 1u);
            mytot += cnt;
        }
    }
    if (sizeof(V) == 8 && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long c = *t.special;
        if (c) {
            atomicAdd(&hist[c < 255 ? (uint32_t)c : 255u], 1u);
            mytot += c;
        }
    }
    atomicAdd(&tot, mytot);
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        if (hist[i]) atomicAdd(stats + 3 + i, (unsigned long long)hist[i]);
    if (threadIdx.x == 0 && tot) atomicAdd(stats + 2, tot);
}

// compacts the occupied slots into (keys, counts) -- the iteration of a counter (dump of the multiple k-mers)
template <typename V>
__global__ void __launch_bounds__(256) count_export_kernel(CountTable t, uint32_t min_count, V* __restrict__ keys,
                                                            uint32_t* __restrict__ counts, unsigned long long* cursor,
                                                            uint64_t cap) {
    const uint64_t n = t.capmask + 1;
    for (uint64_t i0 = blockIdx.x * (uint64_t)blockDim.x; i0 < n; i0 += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = i0 + threadIdx.x;
        uint64_t key = 0, cnt = 0;
        bool take = i < n && CountOps<V>::occupied(t, i, key, cnt) && cnt >= min_count;
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, take);
        uint64_t base = 0;
        const int lane = threadIdx.x & 31;
        if (lane == 0 && bal) base = atomicAdd(cursor, (unsigned long long)__popc(bal));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (take) {
            const uint64_t pos = base + __popc(bal & ((1u << lane) - 1));
            if (pos < cap) {
                keys[pos] = (V)key;
                counts[pos] = cnt > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)cnt;
            }
        }
    }
}

// ---- partition by owner: DispatchableT::dispatch (kmercount.rs:382-420) -------------------------
// owner = intNN_hash(compressed canonical value) % nparts.  Two walks: (A) every block counts its
// k-mers per owner into block_counts[part][block]; an exclusive scan of that matrix (part major)
// gives every block its private output ranges; (B) the block recomputes the keys and writes them.
constexpr int MAX_PARTS = 4096;

// PMODE 0: owner = DispatchableT::dispatch = intNN_hash(value) % nparts (multi-GPU exchange)
// PMODE 1: bucket = region of the counting table the key hashes to: (fmix64(key) & capmask) >> shift
//          (two-phase insertion: k-mers are grouped by table region so that the inserts of a region hit L2)
template <typename V, int PMODE>
__device__ __forceinline__ uint32_t part_of(V key, uint32_t nparts, uint64_t capmask, uint32_t shift) {
    if (PMODE == 0) {
        const V h = inv_hash(key);
        return (nparts & (nparts - 1)) == 0 ? (uint32_t)h & (nparts - 1) : (uint32_t)(h % (V)nparts);  // a mask for 2 / 4 / 8 parts
    }
    return (uint32_t)((fmix64((uint64_t)key) & capmask) >> shift);
}

// WRITE: bucket p is written at dests[p] + (offset kept in block_counts) when `dests` is given -- dests[p] may be
// peer-GPU memory mapped over NVLink: extraction, bucketing and the exchange are then ONE kernel -- else at out + offset.
template <typename V, bool WRITE, int PMODE>
__global__ void __launch_bounds__(256) count_partition_kernel(SeqView b, uint64_t total_bytes, uint32_t k, int canonical,
                                                               uint32_t nparts, uint64_t capmask, uint32_t shift,
                                                               unsigned long long* block_counts, V* __restrict__ out,
                                                               V* const* __restrict__ dests) {
    extern __shared__ unsigned long long cur[];  // nparts cursors
    for (uint32_t p = threadIdx.x; p < nparts; p += blockDim.x)
        cur[p] = WRITE ? block_counts[(size_t)p * gridDim.x + blockIdx.x] : 0ULL;
    __syncthreads();
    const uint64_t ngroups = (total_bytes + GROUP_BYTES - 1) / GROUP_BYTES;
    const int lane = threadIdx.x & 31;
    const uint32_t wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    // a block owns the groups blockIdx.x * wpb + wib + j * gridDim.x * wpb in both walks: same k-mers, same block
    for (uint64_t g = (uint64_t)blockIdx.x * wpb + wib; g < ngroups; g += (uint64_t)gridDim.x * wpb)
        warp_for_each_kmer<V>(b, total_bytes, g, k, canonical != 0, lane, [&](V key, bool active) {
            // lanes going to the same part reserve their output positions with one shared-memory atomic
            const uint32_t p = active ? part_of<V, PMODE>(key, nparts, capmask, shift) : 0xFFFFFFFFu;
            const uint32_t peers = __match_any_sync(0xFFFFFFFFu, p);
            const int leader = __ffs(peers) - 1;
            unsigned long long base = 0;
            if (active && lane == leader) base = atomicAdd(&cur[p], (unsigned long long)__popc(peers));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (WRITE && active) {
                V* o = dests ? dests[p] : out;
                o[base + __popc(peers & ((1u << lane) - 1))] = key;
            }
        });
    if (!WRITE) {
        __syncthreads();
        for (uint32_t p = threadIdx.x; p < nparts; p += blockDim.x) block_counts[(size_t)p * gridDim.x + blockIdx.x] = cur[p];
    }
}

// block_counts (nparts * nblocks entries, part major) -> exclusive offsets, in three steps
__global__ void count_partition_rowsum_kernel(const unsigned long long* block_counts, uint32_t nparts, uint32_t nblocks,
                                              unsigned long long* part_totals) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nparts) return;
    unsigned long long s = 0;
    for (uint32_t j = 0; j < nblocks; ++j) s += block_counts[(size_t)p * nblocks + j];
    part_totals[p] = s;
}
__global__ void count_partition_base_kernel(const unsigned long long* part_totals, uint32_t nparts, unsigned long long* part_base) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long acc = 0;
        for (uint32_t p = 0; p < nparts; ++p) {
            part_base[p] = acc;
            acc += part_totals[p];
        }
    }
}
__global__ void count_partition_offsets_kernel(unsigned long long* block_counts, uint32_t nparts, uint32_t nblocks,
                                               const unsigned long long* part_base) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nparts) return;
    unsigned long long base = part_base[p];
    for (uint32_t j = 0; j < nblocks; ++j) {
        const unsigned long long v = block_counts[(size_t)p * nblocks + j];
        block_counts[(size_t)p * nblocks + j] = base;
        base += v;
    }
}

// ---- launchers ---------------------------------------------------------------------------------
static int grid_for(uint64_t work_items, int block, int sm_count, int per_sm) {
    uint64_t want = (work_items + block - 1) / block;
    uint64_t cap = (uint64_t)sm_count * per_sm;
    return (int)(want < cap ? (want ? want : 1) : cap);
}

cudaError_t launch_count_init(const CountTable& t, bool key64, int sm_count, cudaStream_t st) {
    count_init_kernel<<<grid_for(t.capmask + 1, 256, sm_count, 16), 256, 0, st>>>(t, key64 ? 1 : 0);
    return cudaGetLastError();
}

// byte_begin (a multiple of GROUP_BYTES) .. total_bytes: the k-mers that START in that range of the packed buffer
cudaError_t launch_count_insert_seqs(const SeqView& b, uint64_t total_bytes, uint32_t k, bool key64, bool canonical,
                                     const CountTable& t, int sm_count, cudaStream_t st, uint64_t byte_begin) {
    if (b.nseq == 0 || total_bytes <= byte_begin) return cudaSuccess;
    // a warp takes one slice of the packed buffer; small inputs (one genome) get smaller slices so that every SM has
    // its eight CTAs of work: the kernel lives on memory-level parallelism
    uint32_t group_bytes = GROUP_BYTES;
    const uint64_t want_warps = (uint64_t)sm_count * 8 * 8;
    while (group_bytes > 64 && (total_bytes - byte_begin + group_bytes - 1) / group_bytes < want_warps) group_bytes /= 2;
    const uint64_t first_group = byte_begin / group_bytes;
    const uint64_t ngroups = (total_bytes + group_bytes - 1) / group_bytes - first_group;
    const int grid = grid_for(ngroups * 32, 256, sm_count, 8);
    if (key64) count_insert_seqs_kernel<uint64_t><<<grid, 256, 0, st>>>(b, total_bytes, k, canonical, t, group_bytes, first_group);
    else count_insert_seqs_kernel<uint32_t><<<grid, 256, 0, st>>>(b, total_bytes, k, canonical, t, group_bytes, first_group);
    return cudaGetLastError();
}

// both walks of the prefiltered insertion (pmh3a_prefilter_kernel); seen: 2 * (bitmask + 1) / 8 zeroed bytes
cudaError_t launch_pmh3a_prefilter(const SeqView& b, uint64_t total_bytes, uint32_t k, bool key64, bool canonical, int kmer_type,
                                   int hash_kind, double bound, double c1, const CountTable& t, uint32_t* seen, uint64_t bitmask,
                                   int sm_count, cudaStream_t st) {
    if (b.nseq == 0 || total_bytes == 0) return cudaSuccess;
    uint32_t group_bytes = GROUP_BYTES;
    const uint64_t want_warps = (uint64_t)sm_count * 8 * 8;
    while (group_bytes > 64 && (total_bytes + group_bytes - 1) / group_bytes < want_warps) group_bytes /= 2;
    const uint64_t ngroups = (total_bytes + group_bytes - 1) / group_bytes;
    const int grid = grid_for(ngroups * 32, 256, sm_count, 8);
    if (key64) {
        pmh3a_prefilter_kernel<uint64_t, 0><<<grid, 256, 0, st>>>(b, total_bytes, k, canonical, kmer_type, hash_kind, bound, c1, t, seen, bitmask, group_bytes);
        pmh3a_prefilter_kernel<uint64_t, 1><<<grid, 256, 0, st>>>(b, total_bytes, k, canonical, kmer_type, hash_kind, bound, c1, t, seen, bitmask, group_bytes);
    } else {
        pmh3a_prefilter_kernel<uint32_t, 0><<<grid, 256, 0, st>>>(b, total_bytes, k, canonical, kmer_type, hash_kind, bound, c1, t, seen, bitmask, group_bytes);
        pmh3a_prefilter_kernel<uint32_t, 1><<<grid, 256, 0, st>>>(b, total_bytes, k, canonical, kmer_type, hash_kind, bound, c1, t, seen, bitmask, group_bytes);
    }
    return cudaGetLastError();
}

cudaError_t launch_count_insert_keys(const void* keys, uint64_t n, bool key64, const CountTable& t, int sm_count,
                                     cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const int grid = grid_for(n, 256, sm_count, 8);
    if (key64) count_insert_keys_kernel<uint64_t><<<grid, 256, 0, st>>>((const uint64_t*)keys, n, t);
    else count_insert_keys_kernel<uint32_t><<<grid, 256, 0, st>>>((const uint32_t*)keys, n, t);
    return cudaGetLastError();
}

cudaError_t launch_count_query(const void* keys, uint64_t n, bool key64, const CountTable& t, uint32_t max_count,
                               uint32_t* out, int sm_count, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const int grid = grid_for(n, 256, sm_count, 8);
    if (key64) count_query_kernel<uint64_t><<<grid, 256, 0, st>>>((const uint64_t*)keys, n, t, max_count, out);
    else count_query_kernel<uint32_t><<<grid, 256, 0, st>>>((const uint32_t*)keys, n, t, max_count, out);
    return cudaGetLastError();
}

cudaError_t launch_count_stats(const CountTable& t, bool key64, unsigned long long* stats, int sm_count, cudaStream_t st) {
    const int grid = grid_for(t.capmask + 1, 256, sm_count, 8);
    if (key64) count_stats_kernel<uint64_t><<<grid, 256, 0, st>>>(t, stats);
    else count_stats_kernel<uint32_t><<<grid, 256, 0, st>>>(t, stats);
    return cudaGetLastError();
}

cudaError_t launch_count_export(const CountTable& t, bool key64, uint32_t min_count, void* keys, uint32_t* counts,
                                unsigned long long* cursor, uint64_t cap, int sm_count, cudaStream_t st) {
    const int grid = grid_for(t.capmask + 1, 256, sm_count, 8);
    if (key64) count_export_kernel<uint64_t><<<grid, 256, 0, st>>>(t, min_count, (uint64_t*)keys, counts, cursor, cap);
    else count_export_kernel<uint32_t><<<grid, 256, 0, st>>>(t, min_count, (uint32_t*)keys, counts, cursor, cap);
    return cudaGetLastError();
}

int count_partition_grid(uint64_t total_bytes, int sm_count) {
    const uint64_t ngroups = (total_bytes + GROUP_BYTES - 1) / GROUP_BYTES;
    return grid_for(ngroups * 32, 256, sm_count, 4);
}

template <typename V, int PMODE>
static cudaError_t launch_partition_t(const SeqView& b, uint64_t total_bytes, uint32_t k, bool canonical, uint32_t nparts,
                                      uint64_t capmask, uint32_t shift, int grid, unsigned long long* block_counts,
                                      unsigned long long* part_totals, void* out, cudaStream_t st) {
    const size_t smem = sizeof(unsigned long long) * nparts;
    unsigned long long* part_base = part_totals + nparts;
    count_partition_kernel<V, false, PMODE><<<grid, 256, smem, st>>>(b, total_bytes, k, canonical, nparts, capmask, shift,
                                                                      block_counts, nullptr, nullptr);
    count_partition_rowsum_kernel<<<(nparts + 127) / 128, 128, 0, st>>>(block_counts, nparts, (uint32_t)grid, part_totals);
    count_partition_base_kernel<<<1, 32, 0, st>>>(part_totals, nparts, part_base);
    count_partition_offsets_kernel<<<(nparts + 127) / 128, 128, 0, st>>>(block_counts, nparts, (uint32_t)grid, part_base);
    count_partition_kernel<V, true, PMODE><<<grid, 256, smem, st>>>(b, total_bytes, k, canonical, nparts, capmask, shift,
                                                                     block_counts, (V*)out, nullptr);
    return cudaGetLastError();
}

// the two halves apart (peer-to-peer exchange): (1) count per owner; the host turns the job-wide counts into the
// offset of this rank's bucket inside every destination; (2) write every bucket at dests[p] + part_base[p]
cudaError_t launch_count_partition_counts(const SeqView& b, uint64_t total_bytes, uint32_t k, bool key64, bool canonical,
                                          uint32_t nparts, int grid, unsigned long long* block_counts,
                                          unsigned long long* part_totals, cudaStream_t st) {
    const size_t smem = sizeof(unsigned long long) * nparts;
    if (key64) count_partition_kernel<uint64_t, false, 0><<<grid, 256, smem, st>>>(b, total_bytes, k, canonical, nparts, 0, 0, block_counts, nullptr, nullptr);
    else count_partition_kernel<uint32_t, false, 0><<<grid, 256, smem, st>>>(b, total_bytes, k, canonical, nparts, 0, 0, block_counts, nullptr, nullptr);
    count_partition_rowsum_kernel<<<(nparts + 127) / 128, 128, 0, st>>>(block_counts, nparts, (uint32_t)grid, part_totals);
    return cudaGetLastError();
}
cudaError_t launch_count_partition_scatter(const SeqView& b, uint64_t total_bytes, uint32_t k, bool key64, bool canonical,
                                           uint32_t nparts, int grid, unsigned long long* block_counts,
                                           const unsigned long long* part_base, void* const* dests, cudaStream_t st) {
    const size_t smem = sizeof(unsigned long long) * nparts;
    count_partition_offsets_kernel<<<(nparts + 127) / 128, 128, 0, st>>>(block_counts, nparts, (uint32_t)grid, part_base);
    if (key64) count_partition_kernel<uint64_t, true, 0><<<grid, 256, smem, st>>>(b, total_bytes, k, canonical, nparts, 0, 0, block_counts, nullptr, (uint64_t* const*)dests);
    else count_partition_kernel<uint32_t, true, 0><<<grid, 256, smem, st>>>(b, total_bytes, k, canonical, nparts, 0, 0, block_counts, nullptr, (uint32_t* const*)dests);
    return cudaGetLastError();
}

// block_counts: nparts * grid entries (scratch); part_totals: 2 * nparts entries (totals, then scratch); out: all k-mers, part major
cudaError_t launch_count_partition(const SeqView& b, uint64_t total_bytes, uint32_t k, bool key64, bool canonical,
                                   uint32_t nparts, int grid, unsigned long long* block_counts,
                                   unsigned long long* part_totals, void* out, cudaStream_t st) {
    if (key64) return launch_partition_t<uint64_t, 0>(b, total_bytes, k, canonical, nparts, 0, 0, grid, block_counts, part_totals, out, st);
    return launch_partition_t<uint32_t, 0>(b, total_bytes, k, canonical, nparts, 0, 0, grid, block_counts, part_totals, out, st);
}

// two-phase insertion: group the k-mers of the batch by table region (nparts = regions, a power of two)
cudaError_t launch_count_partition_by_region(const SeqView& b, uint64_t total_bytes, uint32_t k, bool key64, bool canonical,
                                             const CountTable& t, uint32_t nparts, int grid, unsigned long long* block_counts,
                                             unsigned long long* part_totals, void* out, cudaStream_t st) {
    uint32_t shift = 0;
    while (((t.capmask + 1) >> shift) > nparts) ++shift;
    if (key64) return launch_partition_t<uint64_t, 1>(b, total_bytes, k, canonical, nparts, t.capmask, shift, grid, block_counts, part_totals, out, st);
    return launch_partition_t<uint32_t, 1>(b, total_bytes, k, canonical, nparts, t.capmask, shift, grid, block_counts, part_totals, out, st);
}

}  // namespace kmu
