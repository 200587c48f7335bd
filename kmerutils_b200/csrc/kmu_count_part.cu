// kmu_count_part.cu -- two-phase insertion into the exact counting table, and the fused multi-GPU exchange.
//
// Why: a k-mer inserted straight into a table far larger than L2 costs one random DRAM read-modify-write of a
// 32-byte sector per k-mer -- measured 163 B of DRAM traffic per 31-mer against 32 B algorithmic, 15.5 G updates/s
// whatever the kernel does (profiles/r1d_micro_atomics.txt).  So the batch is first PARTITIONED by the region of the
// table each k-mer hashes to (phase 1, streaming), then inserted region after region by the whole grid (phase 2): a
// region's sectors are fetched once per pass, take all their updates in L2 and are written back once (36.8 B of DRAM
// traffic per 31-mer for both phases, profiles/r2c_count_traffic.csv; the updates themselves run at the L2's rate for a
// look followed by a RED, 48-54 G/s, profiles/r2c_micro_warm_l2.txt).
//
//   phase 1  count_part_kernel      one walk over the packed reads (or over a key array): canonical k-mer ->
//                                   bucket = owner * nregions + region.  A CTA takes tiles of 1 KB of packed bytes; the
//                                   tile's bytes and the offsets / lengths of its sequences are copied into shared memory
//                                   a tile ahead (cp.async); the sequence segments inside the tile are cut into chunks of
//                                   128 positions that the 16 warps take in turn, four consecutive k-mers per lane from
//                                   one window (kmers4_at).  The tile's <= 4096 k-mers are sorted by bucket in shared
//                                   memory (histogram, scan, rank) and every bucket's run is appended to that bucket's
//                                   slab with ONE global atomic per (tile, bucket) and stores from consecutive threads.
//                                   Slabs have a fixed capacity (expected share + 8 sigma): an overflow (pathological
//                                   input: one k-mer repeated millions of times) raises a flag and the caller falls back
//                                   to direct insertion.
//   phase 2  count_insert_slabs_kernel   for r in regions: all CTAs insert the keys of region r (look, atomicCAS claim,
//                                   RED add), kept within two regions of each other by `done` counters.
//
// Multi-GPU (reference: DispatchableT::dispatch, kmercount.rs:382-420, and the one-producer / N-consumer hand-off of
// count_kmer_threaded_one_to_many, :881-974): owner = intNN_hash(key) % nowners.  dests[o] is the receive buffer of
// rank o -- for o != self a peer GPU's memory mapped with CUDA IPC -- so extraction, canonicalisation, bucketing
// by (owner, region) and the all-to-all over NVLink are ONE kernel and one walk: no staging buffer, no separate
// collective on the data path.  Cursors are sender-local; the receive buffer of every rank is cut into
// [region][sender] slabs, so the receiver's phase 2 needs only the senders' final cursor values (a few KB).
#include <cstdint>

#include "kmu_count_ops.cuh"
#include "kmu_device.cuh"
#include "kmu_kernels.h"

namespace kmu {

constexpr int PART_THREADS = 512;
constexpr uint32_t PART_TILE_BYTES = 1024;               // packed bytes per tile
constexpr uint32_t PART_TILE_KEYS = PART_TILE_BYTES * 4;  // k-mers per tile (at most one per base)
constexpr uint32_t PART_STAGE_BYTES = PART_TILE_BYTES + 32;  // a window of the tile's last position ends < 16 bytes behind it
constexpr uint32_t PART_STAGE_SEQS = 64;
static_assert(PART_STAGE_BYTES / 16 + 2 * PART_STAGE_SEQS <= PART_THREADS, "one staging copy per thread");

__device__ __forceinline__ void cp_async(void* smem_dst, const void* src, int bytes16) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    if (bytes16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
}
// the copies of one tile (bytes [T0, T0 + PART_STAGE_BYTES) as far as the batch's tail slack goes; offsets and lengths of the
// sequences first .. first + 63): one copy per thread, committed as one group
__device__ __forceinline__ void part_stage(const SeqView& b, uint64_t T0, uint64_t total_bytes, uint64_t first, unsigned char* st_bytes,
                                           unsigned long long* st_off, unsigned long long* st_len, int tid) {
    constexpr int NP = PART_STAGE_BYTES / 16;
    if (tid < NP) {
        if (T0 + 16ull * tid + 16 <= total_bytes + 64) cp_async(st_bytes + 16 * tid, b.packed + T0 + 16ull * tid, 1);
    } else if (tid < NP + (int)PART_STAGE_SEQS) {
        const uint64_t q = first + (tid - NP);
        if (q < b.nseq) cp_async(st_off + (tid - NP), b.byte_off + q, 0);
    } else if (tid < NP + 2 * (int)PART_STAGE_SEQS) {
        const uint64_t q = first + (tid - NP - PART_STAGE_SEQS);
        if (q < b.nseq) cp_async(st_len + (tid - NP - PART_STAGE_SEQS), b.nbases + q, 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

template <typename V>
__device__ __forceinline__ uint32_t part_bucket(V key, const PartGeom& g) {
    uint32_t b = g.nregions > 1 ? (uint32_t)((fmix64((uint64_t)key) & g.capmask) >> g.shift) : 0u;
    if (g.nowners > 1) {
        // DispatchableT::dispatch: intNN_hash(v) % N -- a mask for 2 / 4 / 8 GPUs (a 64-bit remainder by a run-time divisor is
        // ~100 instructions, four times per lane and chunk)
        const V h = inv_hash(key);
        const uint32_t o = (g.nowners & (g.nowners - 1)) == 0 ? (uint32_t)h & (g.nowners - 1) : (uint32_t)(h % (V)g.nowners);
        b += o * g.nregions;
    }
    return b;
}

// last sequence s >= s_hint with byte_off[s] <= byte (byte_off ascending): the warp looks at 32 entries at a time
__device__ __forceinline__ uint64_t seq_forward(const uint64_t* __restrict__ byte_off, uint64_t nseq, uint64_t s, uint64_t byte,
                                                int lane) {
    for (;;) {
        const uint64_t idx = s + 1 + lane;
        const bool ok = idx < nseq && __ldg(byte_off + idx) <= byte;
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, ok);
        s += __popc(bal);
        if (bal != 0xFFFFFFFFu) return s;
    }
}

template <typename V, bool FROM_SEQ>
__global__ void __launch_bounds__(PART_THREADS, 2)
    count_part_kernel(SeqView b, uint64_t byte_begin, uint64_t byte_end, uint64_t total_bytes, uint32_t k, int canonical,
                      const V* __restrict__ keys, uint64_t nkeys, KeySegs segs, PartGeom g, V* const* __restrict__ dests,
                      unsigned long long* __restrict__ cursors, unsigned long long* __restrict__ flag) {
    if (segs.skip_flag && *segs.skip_flag) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t NB = g.nowners * g.nregions;
    V* tkeys = (V*)smem_raw;
    V* sorted = tkeys + PART_TILE_KEYS;
    uint16_t* tb = (uint16_t*)(sorted + PART_TILE_KEYS);
    uint16_t* sb = tb + PART_TILE_KEYS;
    uint32_t* hist = (uint32_t*)(sb + PART_TILE_KEYS);  // counts, then the fill cursor of the bucket inside `sorted`
    uint32_t* jend = hist + NB;       // one past the last sorted index of the bucket that still fits its slab
    V** dptr = (V**)(jend + NB);  // slab address of sorted index 0 of the bucket (hist and jend take 8 NB bytes: aligned)
    __shared__ uint32_t tile_n, warp_sums[PART_THREADS / 32];
    // read form: the tile's packed bytes and the offsets / lengths of the sequences around it, copied a tile ahead
    // (cp.async, behind the sort and copy-out of the tile before) -- phase 1 then starts without a global load
    __shared__ __align__(16) unsigned char st_bytes[PART_STAGE_BYTES];
    __shared__ unsigned long long st_off[PART_STAGE_SEQS], st_len[PART_STAGE_SEQS], st_first, st_next;
    __shared__ unsigned long long seg_pref[65], seg_n[64];  // key-array form: tiles before segment s, keys of segment s
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;

    if (!FROM_SEQ) {
        // the keys are one array of nkeys, or segs.nseg segments `segs.stride` keys apart holding
        // min(segs.counts[s * segs.count_stride], nkeys) keys each (the slabs a coarse region received from its senders)
        if (tid == 0) {
            unsigned long long acc = 0;
            const uint32_t ns = segs.counts ? segs.nseg : 1u;
            for (uint32_t q = 0; q < ns; ++q) {
                unsigned long long nq = nkeys;
                if (segs.counts) {
                    nq = segs.counts[(uint64_t)q * segs.count_stride];
                    if (nq > nkeys) nq = nkeys;
                }
                seg_pref[q] = acc;
                seg_n[q] = nq;
                acc += (nq + PART_TILE_KEYS - 1) / PART_TILE_KEYS;
            }
            seg_pref[ns] = acc;
            for (uint32_t q = ns + 1; q < 65; ++q) seg_pref[q] = ~0ULL;
        }
        __syncthreads();
    }
    const uint32_t nsegs = (!FROM_SEQ && segs.counts) ? segs.nseg : 1u;
    const uint64_t ntiles = FROM_SEQ ? (byte_end - byte_begin + PART_TILE_BYTES - 1) / PART_TILE_BYTES : seg_pref[nsegs];
    // a CTA owns a contiguous range of tiles: the sequence index only moves forward
    const uint64_t per = (ntiles + gridDim.x - 1) / gridDim.x;
    const uint64_t t0 = (uint64_t)blockIdx.x * per, t1 = min(ntiles, t0 + per);
    for (uint32_t p = tid; p < NB; p += PART_THREADS) hist[p] = 0;
    if (tid == 0) tile_n = 0;
    uint64_t s = 0;
    if (FROM_SEQ && t0 < t1) {
        s = seq_of_byte_warp(b.byte_off, b.nseq, byte_begin + t0 * PART_TILE_BYTES);
        part_stage(b, byte_begin + t0 * PART_TILE_BYTES, total_bytes, s, st_bytes, st_off, st_len, tid);
        if (tid == 0) st_first = s;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    bool lost = false;
    const uint32_t lg_regions = (uint32_t)__ffs((int)g.nregions) - 1;

    for (uint64_t t = t0; t < t1; ++t) {
        // ---- phase 1: the tile's keys and bucket ids into shared memory, histogram of the buckets
        if (FROM_SEQ) {
            // the tile's bytes cut the sequences into segments; the segments are cut into chunks of 128 positions and the
            // warps of the CTA take the chunks in turn, four consecutive k-mers per lane from one packed window
            const uint64_t T0 = byte_begin + t * PART_TILE_BYTES;
            const uint64_t T1 = min(min(T0 + (uint64_t)PART_TILE_BYTES, byte_end), total_bytes);
            if (T0 < T1) {
                // s: the last sequence that starts at or before T0 -- among the staged offsets (they start at the last
                // sequence of the tile before), else by the forward search in global memory
                const uint64_t first = st_first;
                {
                    const uint64_t q = first + lane;
                    const bool le = q < b.nseq && st_off[lane] <= T0;
                    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, le);
                    s = first + (bal ? __popc(bal) - 1 : 0);
                    if (bal == 0xFFFFFFFFu) s = seq_forward(b.byte_off, b.nseq, s, T0, lane);
                }
                uint32_t cbase = 0;  // chunks of the sequences before this batch of 32
                uint64_t last_inside = s;
                for (uint64_t qb = s; qb < b.nseq; qb += 32) {
                    const uint64_t q = qb + lane;
                    const bool staged = q - first < PART_STAGE_SEQS;  // q >= first
                    uint64_t sbyte = ~0ULL, p_lo = 0;
                    uint32_t nseg = 0;
                    if (q < b.nseq) sbyte = staged ? st_off[q - first] : __ldg(b.byte_off + q);
                    const bool inside = sbyte < T1;  // the offsets ascend: a prefix of the lanes
                    if (inside) {
                        const uint64_t L = staged ? st_len[q - first] : __ldg(b.nbases + q);
                        const uint64_t nk = L >= k ? L - k + 1 : 0;
                        p_lo = T0 > sbyte ? (T0 - sbyte) * 4 : 0;
                        const uint64_t p_hi = min(nk, (T1 - sbyte) * 4);
                        nseg = p_hi > p_lo ? (uint32_t)(p_hi - p_lo) : 0u;
                    }
                    const uint32_t nin = __popc(__ballot_sync(0xFFFFFFFFu, inside));
                    if (nin) last_inside = qb + nin - 1;
                    const uint32_t nch = (nseg + 127) >> 7;
                    uint32_t incl = nch;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                        if (lane >= d) incl += o;
                    }
                    const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
                    const uint32_t excl = incl - nch;
                    for (uint32_t cl = ((uint32_t)wib - cbase) & (PART_THREADS / 32 - 1); cl < total; cl += PART_THREADS / 32) {
                        const int j = __popc(__ballot_sync(0xFFFFFFFFu, incl <= cl));  // the lane that holds the chunk's segment
                        const uint64_t sb_j = __shfl_sync(0xFFFFFFFFu, sbyte, j), pc = __shfl_sync(0xFFFFFFFFu, p_lo, j) +
                                                                                       (uint64_t)(cl - __shfl_sync(0xFFFFFFFFu, excl, j)) * 128;
                        const uint32_t left = __shfl_sync(0xFFFFFFFFu, nseg, j) - (cl - __shfl_sync(0xFFFFFFFFu, excl, j)) * 128;
                        const uint32_t cn = min(128u, left);
                        V keys[4] = {0, 0, 0, 0};
                        // the chunk's first word lies at or behind T0 (a sequence starts on a multiple of 16 bytes, the tile on
                        // a multiple of 1024), its last window ends inside the staged bytes
                        if (4u * lane < cn)
                            kmers4_at<false>((const uint32_t*)(st_bytes + (sb_j + (pc >> 4) * 4 - T0)), ((uint32_t)pc & 15u) + 4u * lane, k,
                                             canonical != 0, keys);
                        uint32_t base = 0;
                        if (lane == 0) base = atomicAdd(&tile_n, cn);
                        base = __shfl_sync(0xFFFFFFFFu, base, 0);
                        // position 4 lane + t of the chunk goes to slot (positions with a smaller t) + lane: consecutive
                        // lanes, consecutive slots (the order inside the tile is irrelevant, it is sorted next)
#pragma unroll
                        for (uint32_t tt = 0; tt < 4; ++tt) {
                            const uint32_t cnt = cn > tt ? (cn - tt + 3) >> 2 : 0u;
                            if ((uint32_t)lane < cnt) {
                                const uint32_t bk = part_bucket<V>(keys[tt], g);
                                tkeys[base + lane] = keys[tt];
                                tb[base + lane] = (uint16_t)bk;
                                atomicAdd(&hist[bk], 1u);
                            }
                            base += cnt;
                        }
                    }
                    cbase += total;
                    if (!__all_sync(0xFFFFFFFFu, inside)) break;
                }
                if (tid == 0) st_next = last_inside;  // the next tile's staged metadata starts here (read behind the barrier)
            } else if (tid == 0) {
                st_next = s;
            }
        } else {
            uint32_t sg = 0;
            while (sg + 1 < nsegs && seg_pref[sg + 1] <= t) ++sg;
            const uint64_t i0 = (t - seg_pref[sg]) * PART_TILE_KEYS;
            const uint32_t n = (uint32_t)min((uint64_t)PART_TILE_KEYS, (uint64_t)seg_n[sg] - i0);
            const V* src = keys + (uint64_t)sg * segs.stride + i0;
            for (uint32_t i = tid; i < n; i += PART_THREADS) {
                const V key = src[i];
                const uint32_t bk = part_bucket<V>(key, g);
                tkeys[i] = key;
                tb[i] = (uint16_t)bk;
                atomicAdd(&hist[bk], 1u);
            }
            if (tid == 0) tile_n = n;
        }
        __syncthreads();
        const uint32_t n = tile_n;
        if (FROM_SEQ && t + 1 < t1) {
            // this tile's staged bytes are dead from here on: the next tile's take their place while this one is sorted
            part_stage(b, byte_begin + (t + 1) * PART_TILE_BYTES, total_bytes, st_next, st_bytes, st_off, st_len, tid);
        }
        // ---- phase 2: exclusive scan of the histogram; one global atomic per non-empty bucket reserves its run.
        // The slab positions come back from L2 while the tile is sorted: the first two buckets of a thread (all of them up to
        // 1024 buckets) keep theirs in registers until then.
        auto place = [&](uint32_t bk, uint32_t cnt, uint32_t o0, unsigned long long at) {
            if (at + cnt > g.slab_cap) lost = true;
            // sorted[o0 + i] goes to slab[at + i]: one pointer and one bound per bucket, not per key
            const uint32_t o = bk >> lg_regions, r = bk & (g.nregions - 1);  // nregions is a power of two
            dptr[bk] = dests[o] + ((uint64_t)r * g.nsend + g.self) * g.slab_cap + at - o0;
            const unsigned long long room = at < g.slab_cap ? g.slab_cap - at : 0ull;
            jend[bk] = o0 + (uint32_t)(room < cnt ? room : cnt);
        };
        unsigned long long at2[2] = {0, 0};
        uint32_t off2[2] = {0, 0}, c2[2] = {0, 0};
        {
            const uint32_t per_t = (NB + PART_THREADS - 1) / PART_THREADS;  // <= 8
            const uint32_t first = tid * per_t;
            uint32_t c[8], sum = 0;
#pragma unroll
            for (uint32_t j = 0; j < 8; ++j) {
                c[j] = (j < per_t && first + j < NB) ? hist[first + j] : 0u;
                sum += c[j];
            }
            uint32_t incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += o;
            }
            if (lane == 31) warp_sums[wib] = incl;
            __syncthreads();
            uint32_t wbase = 0;
            for (int w = 0; w < wib; ++w) wbase += warp_sums[w];
            uint32_t off = wbase + incl - sum;
#pragma unroll
            for (uint32_t j = 0; j < 8; ++j) {
                if (j < per_t && first + j < NB) {
                    const uint32_t bk = first + j;
                    hist[bk] = off;
                    if (j < 2) off2[j] = off;
                    if (c[j]) {
                        const unsigned long long at = atomicAdd(&cursors[bk], (unsigned long long)c[j]);
                        if (j < 2)
                            at2[j] = at;
                        else
                            place(bk, c[j], off, at);
                    }
                    off += c[j];
                }
            }
            c2[0] = c[0];
            c2[1] = c[1];
        }
        __syncthreads();
        // ---- phase 3: the tile sorted by bucket (four keys in flight per thread, then the rest one by one)
        {
            uint32_t i = tid;
            for (; i + 3 * PART_THREADS < n; i += 4 * PART_THREADS) {
                uint32_t bk[4], r[4];
#pragma unroll
                for (uint32_t u = 0; u < 4; ++u) bk[u] = tb[i + u * PART_THREADS];
#pragma unroll
                for (uint32_t u = 0; u < 4; ++u) r[u] = atomicAdd(&hist[bk[u]], 1u);
#pragma unroll
                for (uint32_t u = 0; u < 4; ++u) {
                    sorted[r[u]] = tkeys[i + u * PART_THREADS];
                    sb[r[u]] = (uint16_t)bk[u];
                }
            }
            for (; i < n; i += PART_THREADS) {
                const uint32_t bk = tb[i];
                const uint32_t r = atomicAdd(&hist[bk], 1u);
                sorted[r] = tkeys[i];
                sb[r] = (uint16_t)bk;
            }
        }
        {
            const uint32_t per_t = (NB + PART_THREADS - 1) / PART_THREADS, first = tid * per_t;
#pragma unroll
            for (uint32_t j = 0; j < 2; ++j)
                if (j < per_t && first + j < NB && c2[j]) place(first + j, c2[j], off2[j], at2[j]);
        }
        __syncthreads();
        // ---- phase 4: every bucket's run goes to its slab (consecutive threads, consecutive addresses)
        {
            uint32_t j = tid;
            for (; j + 3 * PART_THREADS < n; j += 4 * PART_THREADS) {
                uint32_t bk[4], je[4];
                V* dp[4];
#pragma unroll
                for (uint32_t u = 0; u < 4; ++u) bk[u] = sb[j + u * PART_THREADS];
#pragma unroll
                for (uint32_t u = 0; u < 4; ++u) {
                    je[u] = jend[bk[u]];
                    dp[u] = dptr[bk[u]];
                }
#pragma unroll
                for (uint32_t u = 0; u < 4; ++u) {
                    const uint32_t ju = j + u * PART_THREADS;
                    if (ju < je[u]) dp[u][ju] = sorted[ju];
                }
            }
            for (; j < n; j += PART_THREADS) {
                const uint32_t bk = sb[j];
                if (j < jend[bk]) dptr[bk][j] = sorted[j];
            }
        }
        __syncthreads();
        for (uint32_t p = tid; p < NB; p += PART_THREADS) hist[p] = 0;
        if (tid == 0) tile_n = 0;
        if (FROM_SEQ) {
            if (tid == 0) st_first = st_next;
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
    }
    if (lost) *flag = 1ULL;
}

// phase 2: counts[s * count_stride + r] keys of sender s for region r sit at slabs + (r * nsend + s) * slab_cap; the
// regions are regions region0 .. region0 + nregions of the table (region0 only matters for the prefetch addresses).
// All CTAs take the regions in the same order and are kept within two regions of each other: a CTA may start region r
// only when every CTA has finished region r - 2 (done[] counters; cooperative launch, so that all CTAs are resident).
// Without the coupling the CTAs drift apart by tens of regions and the working set leaves L2 (measured: no faster than
// random insertion).
template <typename V>
__global__ void __launch_bounds__(256) count_insert_slabs_kernel(const V* __restrict__ slabs, uint64_t slab_cap, uint32_t nregions,
                                                                 uint32_t nsend, const unsigned long long* __restrict__ counts,
                                                                 CountTable t, uint32_t shift,
                                                                 const unsigned long long* __restrict__ skip_flag, int prefetch,
                                                                 unsigned int* __restrict__ done, uint32_t region0,
                                                                 uint64_t count_stride) {
    if (skip_flag && *skip_flag) return;  // phase 1 overflowed a slab: the caller inserts this chunk directly
    const uint64_t gtid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x, gsize = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t slot_bytes = sizeof(V) == 8 ? 16 : 8;
    const uint64_t region_bytes = nregions > 1 ? (slot_bytes << shift) : (t.capmask + 1) * slot_bytes;
    const uint64_t lines = region_bytes / 128;
    bool ok = true;
    for (uint32_t r = 0; r < nregions; ++r) {
        if (done && r >= 2) {
            if (threadIdx.x == 0) {
                while (*(volatile unsigned int*)(done + r - 2) < gridDim.x) __nanosleep(200);
                __threadfence();
            }
            __syncthreads();
        }
        if (prefetch) {
            // stream the next region of the table into L2 (region 0: the region itself as well): coalesced lines
            // instead of the cold random sector fetches of the updates
            const uint32_t first = r == 0 ? 0 : r + 1, last = r + 1;
            for (uint32_t pr = first; pr <= last && pr < nregions; ++pr) {
                const char* base = (const char*)t.slots + (uint64_t)(region0 + pr) * region_bytes;
                for (uint64_t l = gtid; l < lines; l += gsize)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(base + l * 128));
            }
        }
        for (uint32_t s = 0; s < nsend; ++s) {
            const unsigned long long cnt0 = counts[(uint64_t)s * count_stride + r];
            const uint64_t cnt = cnt0 < slab_cap ? cnt0 : slab_cap;
            const V* seg = slabs + ((uint64_t)r * nsend + s) * slab_cap;
            for (uint64_t i = gtid; i < cnt; i += gsize) ok &= CountOps<V>::insert(t, seg[i], 1u);
        }
        if (done) {
            __syncthreads();
            if (threadIdx.x == 0) {
                __threadfence();
                atomicAdd(done + r, 1u);
            }
        }
    }
    if (!ok) *t.overflow = 1ULL;
}

size_t count_part_smem_bytes(bool key64, uint32_t nbuckets) {
    const size_t esz = key64 ? 8 : 4;
    return (size_t)PART_TILE_KEYS * (2 * esz + 4) + (size_t)nbuckets * 16 + 8;
}

int count_part_grid(int sm_count) { return sm_count * 2; }

cudaError_t launch_count_part_seqs(const SeqView& b, uint64_t byte_begin, uint64_t byte_end, uint64_t total_bytes, uint32_t k,
                                   bool key64, bool canonical, const PartGeom& g, void* const* dests,
                                   unsigned long long* cursors, unsigned long long* flag, int sm_count, cudaStream_t st) {
    if (byte_end <= byte_begin) return cudaSuccess;
    const size_t smem = count_part_smem_bytes(key64, g.nowners * g.nregions);
    const int grid = count_part_grid(sm_count);
    if (key64) {
        static bool attr = false;
        if (!attr) {
            cudaFuncSetAttribute(count_part_kernel<uint64_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 136 * 1024);
            attr = true;
        }
        count_part_kernel<uint64_t, true><<<grid, PART_THREADS, smem, st>>>(b, byte_begin, byte_end, total_bytes, k, canonical, nullptr, 0,
                                                                             KeySegs{}, g, (uint64_t* const*)dests, cursors, flag);
    } else {
        static bool attr = false;
        if (!attr) {
            cudaFuncSetAttribute(count_part_kernel<uint32_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 136 * 1024);
            attr = true;
        }
        count_part_kernel<uint32_t, true><<<grid, PART_THREADS, smem, st>>>(b, byte_begin, byte_end, total_bytes, k, canonical, nullptr, 0,
                                                                             KeySegs{}, g, (uint32_t* const*)dests, cursors, flag);
    }
    return cudaGetLastError();
}

cudaError_t launch_count_part_keys(const void* keys, uint64_t nkeys, const KeySegs& segs, bool key64, const PartGeom& g,
                                   void* const* dests, unsigned long long* cursors, unsigned long long* flag, int sm_count,
                                   cudaStream_t st) {
    if (nkeys == 0) return cudaSuccess;
    const size_t smem = count_part_smem_bytes(key64, g.nowners * g.nregions);
    const int grid = count_part_grid(sm_count);
    SeqView none{nullptr, nullptr, nullptr, 0};
    if (key64) {
        static bool attr = false;
        if (!attr) {
            cudaFuncSetAttribute(count_part_kernel<uint64_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 136 * 1024);
            attr = true;
        }
        count_part_kernel<uint64_t, false><<<grid, PART_THREADS, smem, st>>>(none, 0, 0, 0, 0, 0, (const uint64_t*)keys, nkeys, segs, g,
                                                                              (uint64_t* const*)dests, cursors, flag);
    } else {
        static bool attr = false;
        if (!attr) {
            cudaFuncSetAttribute(count_part_kernel<uint32_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 136 * 1024);
            attr = true;
        }
        count_part_kernel<uint32_t, false><<<grid, PART_THREADS, smem, st>>>(none, 0, 0, 0, 0, 0, (const uint32_t*)keys, nkeys, segs, g,
                                                                              (uint32_t* const*)dests, cursors, flag);
    }
    return cudaGetLastError();
}

// done: nregions zeroed counters (the coupling of the CTAs), or nullptr for free-running CTAs
cudaError_t launch_count_insert_slabs(const void* slabs, uint64_t slab_cap, uint32_t nregions, uint32_t nsend,
                                      const unsigned long long* counts, const CountTable& t, bool key64, uint32_t shift,
                                      const unsigned long long* skip_flag, bool prefetch, unsigned int* done, int sm_count,
                                      cudaStream_t st, uint32_t region0, uint64_t count_stride) {
    if (count_stride == 0) count_stride = nregions;
    const void* fn = key64 ? (const void*)count_insert_slabs_kernel<uint64_t> : (const void*)count_insert_slabs_kernel<uint32_t>;
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 256, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    const int grid = sm_count * per_sm;
    int pf = prefetch ? 1 : 0;
    void* args[] = {(void*)&slabs, (void*)&slab_cap, (void*)&nregions, (void*)&nsend, (void*)&counts, (void*)&t,
                    (void*)&shift,  (void*)&skip_flag, (void*)&pf,      (void*)&done,     (void*)&region0,  (void*)&count_stride};
    if (done) return cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(256), args, 0, st);
    return cudaLaunchKernel(fn, dim3(grid), dim3(256), args, 0, st);
}

}  // namespace kmu
