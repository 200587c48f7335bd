"""Engine: one GPU context of the C ABI plus numpy-friendly wrappers.

Everything here is plumbing around include/kmerutils_b200.h; the arithmetic runs in the CUDA
kernels under kmerutils_b200/csrc/.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import (HASH_CANON_INVHASH, KMER16B32, KMER32, KMER64, KMERAA32, KMERAA64, KmuTimes, check, u64p)

_U64_TYPES = (KMER64, KMERAA64)


def val_dtype(kmer_type):
    """numpy dtype of Kmer::Val for a k-mer type (u32 or u64)."""
    return np.uint64 if kmer_type in _U64_TYPES else np.uint32


def _as_u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


class SeqBatch:
    """A set of 2-bit packed sequences resident in HBM (kmu_seqbatch)."""

    def __init__(self, engine, handle):
        self.engine = engine
        self._h = handle

    @property
    def handle(self):
        if self._h is None:
            raise ValueError("batch already destroyed")
        return self._h

    def __len__(self):
        return int(self.engine.lib.kmu_seqbatch_nseq(self.handle))

    @property
    def total_bases(self):
        return int(self.engine.lib.kmu_seqbatch_total_bases(self.handle))

    @property
    def packed_bytes(self):
        return int(self.engine.lib.kmu_seqbatch_packed_bytes(self.handle))

    def kmer_count(self, k):
        return int(self.engine.lib.kmu_kmer_count(self.handle, k))

    def download(self):
        """-> (packed bytes in the batch layout, byte_off[nseq], nbases[nseq])"""
        n = len(self)
        packed = np.zeros(self.packed_bytes, dtype=np.uint8)
        off = np.zeros(n, dtype=np.uint64)
        nb = np.zeros(n, dtype=np.uint64)
        check(self.engine.lib.kmu_seqbatch_download(self.engine.ctx, self.handle, _p(packed), _p(off, u64p),
                                                    _p(nb, u64p)))
        return packed, off, nb

    def download_meta(self):
        """-> (None, byte_off[nseq], nbases[nseq]) without copying the packed bytes"""
        n = len(self)
        off = np.zeros(n, dtype=np.uint64)
        nb = np.zeros(n, dtype=np.uint64)
        check(self.engine.lib.kmu_seqbatch_download(self.engine.ctx, self.handle, None, _p(off, u64p), _p(nb, u64p)))
        return None, off, nb

    def destroy(self):
        if self._h is not None:
            self.engine.lib.kmu_seqbatch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class Engine:
    """One CUDA context (one B200).  Raises KmuError when no GPU is present."""

    def __init__(self, device=0):
        self.lib = _lib.load_library()
        ctx = C.c_void_p()
        check(self.lib.kmu_ctx_create(device, C.byref(ctx)))
        self.ctx = ctx
        self.device = device

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.kmu_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- bookkeeping --------------------------------------------------------------------
    def launch_count(self):
        return int(self.lib.kmu_launch_count(self.ctx))

    def stream(self):
        return self.lib.kmu_ctx_stream(self.ctx)

    def sync(self):
        check(self.lib.kmu_ctx_sync(self.ctx))

    def last_times(self):
        t = KmuTimes()
        check(self.lib.kmu_last_times(self.ctx, C.byref(t)))
        return {f: getattr(t, f) for f, _ in KmuTimes._fields_}

    def set_profiling(self, on=True):
        check(self.lib.kmu_ctx_set_profiling(self.ctx, int(bool(on))))

    def last_launch_profile(self):
        """Per-launch records of the last sketch call (needs set_profiling(True) before it)."""
        recs = (_lib.KmuLaunchRec * 160)()
        n = self.lib.kmu_last_launch_profile(self.ctx, recs, 160)
        out = []
        for i in range(min(n, 160)):
            d = {f: getattr(recs[i], f) for f, _ in _lib.KmuLaunchRec._fields_}
            d["phase_clocks"] = list(d["phase_clocks"])
            out.append(d)
        return out

    # ---- batches ------------------------------------------------------------------------
    def batch_from_sequences(self, packed_list, nbases):
        """packed_list: one uint8 array per sequence (the reference's Vec<Sequence>)."""
        arrs = [np.ascontiguousarray(a, dtype=np.uint8) for a in packed_list]
        nb = _as_u64(nbases)
        n = len(arrs)
        ptrs = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in arrs])
        h = C.c_void_p()
        check(self.lib.kmu_seqbatch_from_ptrs(self.ctx, ptrs, _p(nb, u64p), n, C.byref(h)))
        return SeqBatch(self, h)

    def batch_from_packed(self, packed, byte_off, nbases):
        """One concatenated buffer; sequence i starts at packed[byte_off[i]]."""
        off = _as_u64(byte_off)
        nb = _as_u64(nbases)
        if isinstance(packed, np.ndarray):
            packed = np.ascontiguousarray(packed, dtype=np.uint8)
            ptr, nbytes = packed.ctypes.data, packed.nbytes
        else:  # (address, nbytes) of e.g. a pinned torch tensor
            ptr, nbytes = packed
        h = C.c_void_p()
        check(self.lib.kmu_seqbatch_from_packed(self.ctx, C.c_void_p(ptr), nbytes, _p(off, u64p), _p(nb, u64p),
                                                len(nb), C.byref(h)))
        return SeqBatch(self, h)

    def batch_from_ascii(self, seqs, drop_invalid=False):
        """seqs: list of bytes.  -> (SeqBatch, invalid_counts).  Sequence::new / encode_and_add on the GPU."""
        lens = np.array([len(s) for s in seqs], dtype=np.uint64)
        off = np.zeros(len(seqs) + 1, dtype=np.uint64)
        np.cumsum(lens, out=off[1:])
        buf = np.frombuffer(b"".join(bytes(s) for s in seqs) + b"\0", dtype=np.uint8)
        bad = np.zeros(max(len(seqs), 1), dtype=np.uint64)
        h = C.c_void_p()
        check(self.lib.kmu_seqbatch_from_ascii(self.ctx, _p(buf), _p(off, u64p), len(seqs), int(bool(drop_invalid)),
                                               _p(bad, u64p), C.byref(h)))
        return SeqBatch(self, h), bad[: len(seqs)]

    def batch_from_aa(self, seqs, drop_invalid=False):
        """seqs: list of bytes (proteins).  -> (SeqBatch of 5-bit codes, invalid_counts).  SequenceAA::new /
        new_filtered (src/aautils/kmeraa.rs:404-456) on the GPU."""
        lens = np.array([len(s) for s in seqs], dtype=np.uint64)
        off = np.zeros(len(seqs) + 1, dtype=np.uint64)
        np.cumsum(lens, out=off[1:])
        buf = np.frombuffer(b"".join(bytes(s) for s in seqs) + b"\0", dtype=np.uint8)
        bad = np.zeros(max(len(seqs), 1), dtype=np.uint64)
        h = C.c_void_p()
        check(self.lib.kmu_seqbatch_from_aa(self.ctx, _p(buf), _p(off, u64p), len(seqs), int(bool(drop_invalid)),
                                            _p(bad, u64p), C.byref(h)))
        return SeqBatch(self, h), bad[: len(seqs)]

    def batch_synth_aa(self, seed, nres):
        nb = _as_u64(nres)
        h = C.c_void_p()
        check(self.lib.kmu_seqbatch_synth_aa(self.ctx, seed, _p(nb, u64p), len(nb), C.byref(h)))
        return SeqBatch(self, h)

    def batch_synth(self, seed, nbases):
        nb = _as_u64(nbases)
        h = C.c_void_p()
        check(self.lib.kmu_seqbatch_synth(self.ctx, seed, _p(nb, u64p), len(nb), C.byref(h)))
        return SeqBatch(self, h)

    def batch_sample_reads(self, genome, seed, first_read, nreads, read_len=150, err_ppm=5000):
        """Short reads drawn from a one-sequence genome batch: random start and strand, substitutions."""
        h = C.c_void_p()
        check(self.lib.kmu_seqbatch_sample_reads(self.ctx, genome.handle, seed, first_read, nreads, read_len, err_ppm,
                                                 C.byref(h)))
        return SeqBatch(self, h)

    def batch_view(self, src, first_seq, nseq):
        """View of nseq consecutive sequences of `src` (shares its packed bases; keep `src` alive)."""
        h = C.c_void_p()
        check(self.lib.kmu_seqbatch_view(src.handle, int(first_seq), int(nseq), C.byref(h)))
        v = SeqBatch(self, h)
        v._parent = src
        return v

    def sketch_groups(self, batch, group_sizes, algo, k, kmer_type, hash_kind=HASH_CANON_INVHASH, **kw):
        """One whole-file signature per group of consecutive sequences (a genome = its contigs), the way gsearch drives
        the `sketch_compressedkmer_seqs` entry points.  algo: "pmh3a" | "superminhash" | "setsketch"; kw as for the
        whole-file calls.  -> (ngroups, m) array"""
        if algo == "pmh3a":
            return self.sketch_pmh3a_groups(batch, group_sizes, k, kmer_type, hash_kind, **kw)
        fn = {"pmh3a": self.sketch_pmh3a_whole, "superminhash": self.sketch_superminhash_whole,
              "setsketch": lambda b, k_, t, h, **kk: self.sketch_setsketch(b, k_, t, h, whole=True, **kk)}[algo]
        rows, first = [], 0
        for n in group_sizes:
            view = self.batch_view(batch, first, int(n))
            rows.append(fn(view, k, kmer_type, hash_kind, **kw))
            view.destroy()
            first += int(n)
        return np.stack(rows) if rows else np.zeros((0, 0))

    def batch_slices(self, src, seq_idx, begin, end):
        """New batch made of the ranges [begin[i], end[i]) of sequences seq_idx[i] of `src` (copied on the device)."""
        idx, b, e = _as_u64(seq_idx), _as_u64(begin), _as_u64(end)
        h = C.c_void_p()
        check(self.lib.kmu_seqbatch_slices(self.ctx, src.handle, _p(idx, u64p), _p(b, u64p), _p(e, u64p), len(idx),
                                           C.byref(h)))
        return SeqBatch(self, h)

    def blocksketch(self, batch, k, m, block_size, hash_kind=HASH_CANON_INVHASH):
        """BlockSeqSketcher::blocksketch_sequences (seqblocksketch.rs:97-167): every sequence is cut into runs of
        block_size consecutive k-mers (Kmer32bit), one ProbMinHash3a signature per block.  The number of blocks
        comes from the BASES (ceil(L / block_size)), so trailing blocks may be empty (all-zero signature).
        -> (sig[nblocks_total, m], numseq[nblocks_total], numblock[nblocks_total])"""
        _, _, nb = batch.download_meta()
        nblocks = (nb + np.uint64(block_size) - np.uint64(1)) // np.uint64(block_size)
        numseq = np.repeat(np.arange(len(nb), dtype=np.uint64), nblocks.astype(np.int64))
        first = np.concatenate([[0], np.cumsum(nblocks)[:-1]]).astype(np.uint64) if len(nb) else np.zeros(0, np.uint64)
        numblock = np.arange(int(nblocks.sum()), dtype=np.uint64) - np.repeat(first, nblocks.astype(np.int64))
        begin = numblock * np.uint64(block_size)
        end = begin + np.uint64(block_size + k - 1)
        blocks = self.batch_slices(batch, numseq, begin, end)
        sig = self.sketch_pmh3a(blocks, k, KMER32, hash_kind, m)
        blocks.destroy()
        return sig, numseq.astype(np.uint32), numblock.astype(np.uint32)

    # ---- k-mers ---------------------------------------------------------------------------
    def generate_kmers(self, batch, k, kmer_type, hash_kind=_lib.HASH_IDENTITY_RAW):
        """-> (values, out_off): all k-mers of all sequences mapped through the hash closure."""
        n = len(batch)
        total = batch.kmer_count(k)
        out = np.zeros(total, dtype=val_dtype(kmer_type))
        off = np.zeros(n + 1, dtype=np.uint64)
        check(self.lib.kmu_generate_kmers(self.ctx, batch.handle, k, kmer_type, hash_kind, _p(out), _p(off, u64p), 0))
        return out, off

    def nthash_canonical(self, batch, k, n_multi=1, want_strand=True):
        total = batch.kmer_count(k)
        h = np.zeros((total, n_multi), dtype=np.uint64)
        strand = np.zeros(total, dtype=np.uint8) if want_strand else None
        check(self.lib.kmu_nthash_canonical(self.ctx, batch.handle, k, n_multi, _p(h),
                                            _p(strand) if want_strand else None, 0))
        return h, strand

    # ---- sketches -------------------------------------------------------------------------
    def sketch_pmh3a(self, batch, k, kmer_type, hash_kind=HASH_CANON_INVHASH, m=200, out=None, out_device_ptr=None):
        """ProbMinHash3a signature per sequence -> (nseq, m) array of Kmer::Val.

        out_device_ptr: raw device address to leave the signatures in HBM (returns None)."""
        n = len(batch)
        if out_device_ptr is not None:
            check(self.lib.kmu_sketch_pmh3a(self.ctx, batch.handle, k, kmer_type, hash_kind, m,
                                            C.c_void_p(out_device_ptr), 1))
            return None
        if out is None:
            out = np.zeros((n, m), dtype=val_dtype(kmer_type))
        check(self.lib.kmu_sketch_pmh3a(self.ctx, batch.handle, k, kmer_type, hash_kind, m, _p(out), 0))
        return out

    def sketch_superminhash(self, batch, k, kmer_type, hash_kind=HASH_CANON_INVHASH, m=200, key_hasher=_lib.HASHER_NOHASH,
                            dtype=np.float64, out_device_ptr=None):
        """SuperMinHash signature per sequence -> (nseq, m) array of f32 / f64 (get_hsketch())."""
        dtype = np.dtype(dtype)
        if dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise ValueError("SuperMinHash signatures are f32 or f64")
        if out_device_ptr is not None:
            check(self.lib.kmu_sketch_superminhash(self.ctx, batch.handle, k, kmer_type, hash_kind, m, key_hasher,
                                                   dtype.itemsize, C.c_void_p(out_device_ptr), 1))
            return None
        out = np.zeros((len(batch), m), dtype=dtype)
        check(self.lib.kmu_sketch_superminhash(self.ctx, batch.handle, k, kmer_type, hash_kind, m, key_hasher,
                                               dtype.itemsize, _p(out), 0))
        return out

    def sketch_superminhash_whole(self, batch, k, kmer_type, hash_kind=HASH_CANON_INVHASH, m=200,
                                  key_hasher=_lib.HASHER_NOHASH, dtype=np.float64):
        """ONE SuperMinHash signature for the whole batch (SuperHashSketch::sketch_compressedkmer_seqs)."""
        dtype = np.dtype(dtype)
        out = np.zeros(m, dtype=dtype)
        check(self.lib.kmu_sketch_superminhash_whole(self.ctx, batch.handle, k, kmer_type, hash_kind, m, key_hasher,
                                                     dtype.itemsize, _p(out), 0))
        return out

    def sketch_setsketch(self, batch, k, kmer_type, hash_kind=HASH_CANON_INVHASH, params=None, dtype=np.uint16,
                         whole=False, out_device_ptr=None):
        """SetSketch registers (HyperLogLogSketch): (nseq, m) array, or (m,) when whole=True (one sketch for the
        batch, sketch_compressedkmer_seqs).  params = (b, m, a, q) or None for SetSketchParams::default()."""
        dtype = np.dtype(dtype)
        if dtype not in (np.dtype(np.uint16), np.dtype(np.uint32), np.dtype(np.uint64)):
            raise ValueError("SetSketch registers are u16, u32 or u64")
        prm = _lib.KmuSetSketchParams(*(params if params is not None else (1.001, 4096, 20.0, 65534)))
        m = int(prm.m)
        if out_device_ptr is not None:
            check(self.lib.kmu_sketch_setsketch(self.ctx, batch.handle, k, kmer_type, hash_kind, C.byref(prm),
                                                dtype.itemsize, int(bool(whole)), C.c_void_p(out_device_ptr), 1))
            return None
        out = np.zeros((1 if whole else len(batch), m), dtype=dtype)
        check(self.lib.kmu_sketch_setsketch(self.ctx, batch.handle, k, kmer_type, hash_kind, C.byref(prm),
                                            dtype.itemsize, int(bool(whole)), _p(out), 0))
        return out[0] if whole else out

    def sketch_pmh3a_whole(self, batch, k, kmer_type, hash_kind=HASH_CANON_INVHASH, m=200):
        """ONE ProbMinHash3a signature for the whole batch (ProbHash3aSketch::sketch_compressedkmer_seqs)."""
        out = np.zeros(m, dtype=val_dtype(kmer_type))
        check(self.lib.kmu_sketch_pmh3a_whole(self.ctx, batch.handle, k, kmer_type, hash_kind, m, _p(out), 0))
        return out

    def sketch_pmh3a_groups(self, batch, group_sizes, k, kmer_type, hash_kind=HASH_CANON_INVHASH, m=200):
        """One whole-file ProbMinHash3a signature per group of consecutive sequences, all groups in one call
        (kmu_sketch_pmh3a_groups).  -> (ngroups, m)"""
        gs = _as_u64(group_sizes)
        out = np.zeros((len(gs), m), dtype=val_dtype(kmer_type))
        check(self.lib.kmu_sketch_pmh3a_groups(self.ctx, batch.handle, _p(gs, u64p), len(gs), k, kmer_type, hash_kind, m,
                                               _p(out), 0))
        return out

    def pmh3a_counter_slots(self, counter, hash_kind, m, bound):
        """Partial ProbMinHash3a registers of a counting table (kmu_pmh3a_counter_slots): -> (h bits u64[m], key u64[m]);
        h bits == the largest f64 where no point fell below `bound`."""
        out = np.zeros((m, 2), dtype=np.uint64)
        check(self.lib.kmu_pmh3a_counter_slots(self.ctx, counter._h, hash_kind, m, float(bound), _p(out), 0))
        return out[:, 0].copy(), out[:, 1].copy()

    def pmh3a_weighted(self, keys, weights, m):
        """ProbMinHash3a::hash_weigthed_hashmap on explicit (key, weight) arrays; keys u32 or u64."""
        keys = np.ascontiguousarray(keys)
        if keys.dtype not in (np.dtype(np.uint32), np.dtype(np.uint64)):
            raise ValueError("keys must be uint32 or uint64")
        w = np.ascontiguousarray(weights, dtype=np.float64)
        out = np.zeros(m, dtype=keys.dtype)
        check(self.lib.kmu_pmh3a_weighted(self.ctx, _p(keys), _p(w), len(keys), keys.dtype.itemsize, m, _p(out)))
        return out

    def sketch_pmh3a_host(self, packed, byte_off, nbases, k, kmer_type, hash_kind, m, out):
        """One-shot: host packed buffer in, host signatures out (H2D + kernels + D2H)."""
        off = _as_u64(byte_off)
        nb = _as_u64(nbases)
        if isinstance(packed, np.ndarray):
            ptr, nbytes = packed.ctypes.data, packed.nbytes
        else:
            ptr, nbytes = packed
        optr = out.ctypes.data if isinstance(out, np.ndarray) else out
        check(self.lib.kmu_sketch_pmh3a_host(self.ctx, C.c_void_p(ptr), nbytes, _p(off, u64p), _p(nb, u64p), len(nb),
                                             k, kmer_type, hash_kind, m, C.c_void_p(optr)))
        return out

    def sketch_pmh3a_host_ptrs(self, seq_addrs, nbases, k, kmer_type, hash_kind, m, out):
        """One-shot from nseq SEPARATE host allocations (the `&[&Sequence]` form): seq_addrs = uint64 array of addresses."""
        addrs = np.ascontiguousarray(seq_addrs, dtype=np.uint64)
        nb = _as_u64(nbases)
        optr = out.ctypes.data if isinstance(out, np.ndarray) else out
        check(self.lib.kmu_sketch_pmh3a_host_ptrs(self.ctx, _p(addrs), _p(nb, u64p), len(nb), k, kmer_type, hash_kind, m,
                                                  C.c_void_p(optr)))
        return out

    def signature_jaccard(self, sig_a, sig_b):
        """(na, nb) matrix of Jaccard estimates = fraction of equal slots (compute_probminhash_jaccard,
        seqsketchjaccard.rs:86-108).  sig_a (na, m), sig_b (nb, m), same dtype (any 2 / 4 / 8 byte slot type)."""
        a = np.ascontiguousarray(sig_a)
        b = np.ascontiguousarray(sig_b)
        if a.ndim != 2 or b.ndim != 2 or a.shape[1] != b.shape[1] or a.dtype != b.dtype:
            raise ValueError("signatures must be 2-D arrays of one dtype and one sketch size")
        out = np.zeros((a.shape[0], b.shape[0]), dtype=np.float64)
        check(self.lib.kmu_signature_jaccard(self.ctx, _p(a), a.shape[0], _p(b), b.shape[0], a.shape[1], a.dtype.itemsize,
                                             _p(out), 0))
        return out

    def jaccard_index_probminhash3a(self, batch_a, batch_b, k, kmer_type, hash_kind=HASH_CANON_INVHASH, m=200):
        """SeqSketcher::jaccard_index_probminhash3a (seqsketchjaccard.rs:423-495) generalised to many-vs-many:
        sketch both batches, compare.  -> (len(batch_a), len(batch_b)) f64"""
        return self.signature_jaccard(self.sketch_pmh3a(batch_a, k, kmer_type, hash_kind, m),
                                      self.sketch_pmh3a(batch_b, k, kmer_type, hash_kind, m))

    # ---- counting ---------------------------------------------------------------------------
    def counter(self, k, kmer_type, capacity, count_bits=8):
        return KmerCounter(self, k, kmer_type, capacity, count_bits)

    def count_partition(self, batch, k, kmer_type, nparts, canonical=True, out_device_ptr=None):
        """All (canonical) compressed k-mer values of the batch bucketed by DispatchableT::dispatch
        (kmercount.rs:382-420) -> (kmers part-major, part_counts[nparts])."""
        counts = np.zeros(nparts, dtype=np.uint64)
        if out_device_ptr is not None:
            check(self.lib.kmu_count_partition(self.ctx, batch.handle, k, kmer_type, int(bool(canonical)), nparts,
                                               C.c_void_p(out_device_ptr), _p(counts, u64p), 1))
            return None, counts
        out = np.zeros(max(batch.kmer_count(k), 1), dtype=val_dtype(kmer_type))
        check(self.lib.kmu_count_partition(self.ctx, batch.handle, k, kmer_type, int(bool(canonical)), nparts, _p(out),
                                           _p(counts, u64p), 0))
        return out[: batch.kmer_count(k)], counts


    # ---- peer-to-peer exchange (one box, CUDA IPC over NVLink) ----------------------------------
    def count_partition_counts(self, batch, k, kmer_type, nparts, canonical=True):
        counts = np.zeros(nparts, dtype=np.uint64)
        check(self.lib.kmu_count_partition_counts(self.ctx, batch.handle, k, kmer_type, int(bool(canonical)), nparts,
                                                  _p(counts, u64p)))
        return counts

    def count_partition_scatter(self, batch, k, kmer_type, nparts, dests, dest_offsets, canonical=True):
        """dests: nparts device pointers (ints); dest_offsets: element offset of this rank's bucket in each of them"""
        ptrs = (C.c_void_p * nparts)(*[int(d) for d in dests])
        offs = _as_u64(dest_offsets)
        check(self.lib.kmu_count_partition_scatter(self.ctx, batch.handle, k, kmer_type, int(bool(canonical)), nparts,
                                                   ptrs, _p(offs, u64p)))

    def ipc_alloc(self, nbytes):
        """-> (device pointer, 64-byte IPC handle) of a buffer other processes of the box can map"""
        ptr = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        check(self.lib.kmu_ipc_alloc(self.ctx, int(nbytes), C.byref(ptr), handle))
        return ptr.value, bytes(handle)

    def ipc_free(self, ptr):
        check(self.lib.kmu_ipc_free(self.ctx, C.c_void_p(ptr)))

    def ipc_open(self, handle):
        ptr = C.c_void_p()
        buf = (C.c_uint8 * 64)(*handle)
        check(self.lib.kmu_ipc_open(self.ctx, buf, C.byref(ptr)))
        return ptr.value

    def ipc_close(self, ptr):
        check(self.lib.kmu_ipc_close(self.ctx, C.c_void_p(ptr)))


class KmerCounter:
    """Exact k-mer multiplicity table in HBM behind the KmerCountT interface
    (src/base/kmercount.rs:48-59): insert_kmer / get_count / get_nb_distinct / get_nb_unique."""

    def __init__(self, engine, k, kmer_type, capacity, count_bits=8):
        self.engine = engine
        self.k = k
        self.kmer_type = kmer_type
        self.count_bits = count_bits
        h = C.c_void_p()
        check(engine.lib.kmu_count_create(engine.ctx, k, kmer_type, count_bits, int(capacity), C.byref(h)))
        self._h = h

    @property
    def dtype(self):
        return val_dtype(self.kmer_type)

    def capacity(self):
        return int(self.engine.lib.kmu_count_capacity(self._h))

    def insert_seqs(self, batch, canonical=True):
        """count_kmer / count_kmer_threaded_one_to_many (kmercount.rs:293-362, 881-974)."""
        check(self.engine.lib.kmu_count_insert_seqs(self.engine.ctx, self._h, batch.handle, int(bool(canonical))))

    def insert_kmers(self, kmers=None, device_ptr=None, n=None):
        """KmerCountT::insert_kmer for an array of compressed k-mer values (host array or device pointer)."""
        if device_ptr is not None:
            check(self.engine.lib.kmu_count_insert_kmers(self.engine.ctx, self._h, C.c_void_p(device_ptr), int(n), 1))
            return
        a = np.ascontiguousarray(kmers, dtype=self.dtype)
        check(self.engine.lib.kmu_count_insert_kmers(self.engine.ctx, self._h, _p(a), len(a), 0))

    def exchange_regions(self, nowners):
        """Regions this table is cut into for an exchange among `nowners` ranks (kmu_count_exchange_geometry)."""
        n = C.c_uint32()
        check(self.engine.lib.kmu_count_exchange_geometry(self._h, int(nowners), C.byref(n)))
        return int(n.value)

    def exchange_scatter(self, batch, nowners, self_rank, slab_cap, dests, canonical=True):
        """ONE kernel: canonical k-mers of `batch` bucketed by (owner, table region) and stored straight into the owners'
        receive buffers `dests` (device pointers, peers over NVLink).  -> (sent_counts[nowners, nregions], overflowed)"""
        nreg = self.exchange_regions(nowners)
        ptrs = (C.c_void_p * nowners)(*[int(d) for d in dests])
        sent = np.zeros(nowners * nreg, dtype=np.uint64)
        ovf = C.c_int32()
        check(self.engine.lib.kmu_count_exchange_scatter(self.engine.ctx, batch.handle, self._h, int(bool(canonical)), int(nowners),
                                                         int(self_rank), int(slab_cap), ptrs, _p(sent, u64p), C.byref(ovf)))
        return sent.reshape(nowners, nreg), bool(ovf.value)

    def insert_slabs(self, slabs_ptr, slab_cap, counts):
        """Insert a receive buffer of [region][sender] slabs; counts[sender, region] keys each (kmu_count_insert_slabs)."""
        cnt = np.ascontiguousarray(counts, dtype=np.uint64)
        check(self.engine.lib.kmu_count_insert_slabs(self.engine.ctx, self._h, C.c_void_p(int(slabs_ptr)), int(slab_cap),
                                                     int(cnt.shape[0]), _p(cnt.reshape(-1), u64p)))

    def get_count(self, kmers):
        a = np.ascontiguousarray(kmers, dtype=self.dtype)
        out = np.zeros(len(a), dtype=np.uint32)
        check(self.engine.lib.kmu_count_query(self.engine.ctx, self._h, _p(a), len(a), _p(out), 0))
        return out

    def stats(self):
        """-> dict(nb_distinct, nb_unique, nb_inserted, hist) ; hist[c] = #k-mers with min(multiplicity, 255) == c"""
        d, u, t = C.c_uint64(), C.c_uint64(), C.c_uint64()
        hist = np.zeros(256, dtype=np.uint64)
        check(self.engine.lib.kmu_count_stats(self.engine.ctx, self._h, C.byref(d), C.byref(u), C.byref(t),
                                              _p(hist, u64p)))
        return {"nb_distinct": d.value, "nb_unique": u.value, "nb_inserted": t.value, "hist": hist}

    def get_nb_distinct(self):
        return self.stats()["nb_distinct"]

    def get_nb_unique(self):
        return self.stats()["nb_unique"]

    def export(self, min_count=1):
        """-> (kmers, counts) of every k-mer with multiplicity >= min_count, unordered."""
        st = self.stats()
        cap = int(st["hist"][max(min_count, 0):].sum()) if min_count < 256 else st["nb_distinct"]
        kmers = np.zeros(max(cap, 1), dtype=self.dtype)
        counts = np.zeros(max(cap, 1), dtype=np.uint32)
        n = C.c_uint64()
        check(self.engine.lib.kmu_count_export(self.engine.ctx, self._h, min_count, _p(kmers), _p(counts), cap,
                                               C.byref(n)))
        return kmers[: n.value], counts[: n.value]

    def dump_multiple(self, path, count_bytes=2):
        """threaded_dump_kmer_counter (kmercount.rs:653-791): file of the k-mers seen at least twice with their counts."""
        import os
        n = C.c_uint64()
        check(self.engine.lib.kmu_count_dump_multiple(self.engine.ctx, self._h, os.fsencode(path), count_bytes, C.byref(n)))
        return n.value

    def destroy(self):
        if self._h is not None:
            self.engine.lib.kmu_count_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


_DEFAULT = {}


def default_engine(device=0):
    """Process-wide engine per device (created on first use)."""
    if device not in _DEFAULT:
        _DEFAULT[device] = Engine(device)
    return _DEFAULT[device]


__all__ = ["Engine", "SeqBatch", "KmerCounter", "default_engine", "val_dtype", "KMER32", "KMER16B32", "KMER64", "KMERAA32",
           "KMERAA64"]
