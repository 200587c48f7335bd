"""Host-side feeders and writers around the GPU path: FASTA/FASTQ packs (src/io.rs, datasketcher's readblockseq),
signature dumps (seqsketchjaccard.rs:382-414, 572-712), block signature dumps (seqblocksketch.rs:59-65, 172-226),
sketch parameter JSON (sketcharg.rs:79-138), and the datasketcher loop that ties them to the kernels."""
import ctypes as C
import json
import os

import numpy as np

from . import _lib
from ._lib import check, u64p


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


class FastxReader:
    """Packs of accepted reads; a record with any non-ACGT character is dropped (io.rs:41-48)."""

    def __init__(self, path, pack_bases=64 << 20):
        self.lib = _lib.load_library()
        h = C.c_void_p()
        check(self.lib.kmu_fastx_open(os.fsencode(path), C.byref(h)))
        self._h = h
        self.buf = np.zeros(pack_bases, dtype=np.uint8)

    def next_pack(self, max_seqs=10000):
        """-> list of bytes (ASCII reads), empty at end of file.  10000 is datasketcher's pack (datasketcher.rs:244)."""
        buf, off = self.next_pack_raw(max_seqs)
        return [buf[int(off[i]): int(off[i + 1])].tobytes() for i in range(len(off) - 1)]

    def next_pack_raw(self, max_seqs=10000):
        """-> (ascii buffer view, offsets[n + 1]) without splitting into Python objects"""
        off = np.zeros(max_seqs + 1, dtype=np.uint64)
        n = C.c_uint64()
        check(self.lib.kmu_fastx_next_pack(self._h, max_seqs, _p(self.buf), self.buf.nbytes, _p(off, u64p), C.byref(n)))
        return self.buf, off[: n.value + 1]

    def stats(self):
        v = [C.c_uint64() for _ in range(4)]
        self.lib.kmu_fastx_stats(self._h, *[C.byref(x) for x in v])
        return dict(zip(("nb_read", "nb_bad_read", "nb_bases", "nb_bad_bases"), (x.value for x in v)))

    def close(self):
        if self._h is not None:
            self.lib.kmu_fastx_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class IngestReader:
    """Multi-threaded feeder (kmu_ingest_*): packs of accepted reads in file order, parsed by `nthreads` host threads into
    pinned buffers while the caller works on the previous pack."""

    def __init__(self, path, nthreads=0, block_bytes=64 << 20):
        self.lib = _lib.load_library()
        h = C.c_void_p()
        check(self.lib.kmu_ingest_open(os.fsencode(path), int(nthreads), int(block_bytes), C.byref(h)))
        self._h = h

    def next(self):
        """-> (ascii address, offsets array view [n + 1], n, token) or None at end of file; call release(token) when the
        pack has been uploaded"""
        a, o, tok = C.c_void_p(), C.c_void_p(), C.c_void_p()
        n = C.c_uint64()
        check(self.lib.kmu_ingest_next(self._h, C.byref(a), C.byref(o), C.byref(n), C.byref(tok)))
        if n.value == 0:
            return None
        off = np.ctypeslib.as_array(C.cast(o, C.POINTER(C.c_uint64)), shape=(n.value + 1,))
        return a.value, off, int(n.value), tok

    def release(self, token):
        check(self.lib.kmu_ingest_release(self._h, token))

    def stats(self):
        v = [C.c_uint64() for _ in range(4)]
        self.lib.kmu_ingest_stats(self._h, *[C.byref(x) for x in v])
        return dict(zip(("nb_read", "nb_bad_read", "nb_bases", "nb_bad_bases"), (x.value for x in v)))

    def close(self):
        if self._h is not None:
            self.lib.kmu_ingest_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def datasketcher_mt(engine, fastx_path, dump_path=None, kmer_size=8, sketch_size=200, nthreads=0, block_bytes=64 << 20,
                    sink=None):
    """The datasketcher loop (src/bin/datasketcher.rs:236-300) on the multi-threaded feeder: the host threads parse the
    following packs while the GPU packs (kmu_seqbatch_from_ascii from pinned memory) and sketches the current one.
    Signatures go to the dump file and / or to sink(sig) in file order.  -> dict(reads, bases, seconds, stats)"""
    import time

    from ._lib import HASH_CANON_INVHASH, KMER32
    from .engine import SeqBatch
    t0 = time.perf_counter()
    n_done = bases = npacks = 0
    t_wait = t_upload = t_sketch = 0.0
    out = SignatureDump(dump_path, sketch_size, kmer_size) if dump_path else None
    with IngestReader(fastx_path, nthreads, block_bytes) as rd:
        t_open = time.perf_counter() - t0
        while True:
            ta = time.perf_counter()
            pack = rd.next()
            tb = time.perf_counter()
            t_wait += tb - ta
            if pack is None:
                break
            addr, off, n, tok = pack
            h = C.c_void_p()
            check(engine.lib.kmu_seqbatch_from_ascii(engine.ctx, C.c_void_p(addr), _p(off, u64p), n, 0, None, C.byref(h)))
            rd.release(tok)  # the pack is on the device: its buffer goes back to the parsers
            tc = time.perf_counter()
            t_upload += tc - tb
            batch = SeqBatch(engine, h)
            sig = engine.sketch_pmh3a(batch, kmer_size, KMER32, HASH_CANON_INVHASH, sketch_size)
            bases += batch.total_bases
            batch.destroy()
            t_sketch += time.perf_counter() - tc
            if out:
                out.write(sig)
            if sink:
                sink(sig)
            n_done += n
            npacks += 1
        stats = rd.stats()
    if out:
        out.close()
    return {"reads": n_done, "bases": bases, "seconds": time.perf_counter() - t0, "stats": stats, "packs": npacks,
            "open_s": t_open, "wait_for_parser_s": t_wait, "upload_and_pack_s": t_upload, "sketch_s": t_sketch}


class SignatureDump:
    """SeqSketcher::create_signature_dump + dump_signatures_block_u32."""

    def __init__(self, path, sketch_size, kmer_size):
        self.lib = _lib.load_library()
        h = C.c_void_p()
        check(self.lib.kmu_sigdump_create(os.fsencode(path), sketch_size, kmer_size, C.byref(h)))
        self._h = h
        self.sketch_size = sketch_size

    def write(self, sig):
        sig = np.ascontiguousarray(sig, dtype=np.uint32)
        assert sig.ndim == 2 and sig.shape[1] == self.sketch_size
        check(self.lib.kmu_sigdump_write(self._h, _p(sig), sig.shape[0]))

    def close(self):
        if self._h is not None:
            check(self.lib.kmu_sigdump_close(self._h))
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class BlockSignatureDump:
    """BlockSeqSketcher::create_signature_dump + dump_blocks."""

    def __init__(self, path, sketch_size, kmer_size, block_size):
        self.lib = _lib.load_library()
        h = C.c_void_p()
        check(self.lib.kmu_blockdump_create(os.fsencode(path), sketch_size, kmer_size, block_size, C.byref(h)))
        self._h = h
        self.sketch_size = sketch_size

    def write(self, sig, numseq, numblock):
        sig = np.ascontiguousarray(sig, dtype=np.uint32)
        ns = np.ascontiguousarray(numseq, dtype=np.uint32)
        nb = np.ascontiguousarray(numblock, dtype=np.uint32)
        assert sig.shape == (len(ns), self.sketch_size) and len(nb) == len(ns)
        check(self.lib.kmu_blockdump_write(self._h, _p(sig), _p(ns), _p(nb), len(ns)))

    def close(self):
        if self._h is not None:
            check(self.lib.kmu_sigdump_close(self._h))
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def read_signature_dump(path, first=0, count=None):
    """SigSketchFileReader: -> (header dict, signatures[count, sketch_size] u32)"""
    lib = _lib.load_library()
    ss, sk, ks, n = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint64()
    check(lib.kmu_sigdump_read(os.fsencode(path), C.byref(ss), C.byref(sk), C.byref(ks), C.byref(n), None, 0, 0))
    if count is None:
        count = n.value - first
    sig = np.zeros((count, sk.value), dtype=np.uint32)
    if count:
        check(lib.kmu_sigdump_read(os.fsencode(path), None, None, None, None, _p(sig), first, count))
    return {"sig_size": ss.value, "sketch_size": sk.value, "kmer_size": ks.value, "nb_signatures": n.value}, sig


def reload_multiple_kmers(path):
    """KmerCountReload::load_multiple_kmers_from_file (kmercount.rs:1209-1351) -> dict(kmer_size, count_bytes,
    nb_declared, kmers (u64 words), counts (u32))"""
    lib = _lib.load_library()
    ks, cb, nd, n = C.c_uint32(), C.c_uint32(), C.c_uint64(), C.c_uint64()
    check(lib.kmu_count_reload_multiple(os.fsencode(path), C.byref(ks), C.byref(cb), C.byref(nd), None, None, 0, C.byref(n)))
    kmers = np.zeros(max(n.value, 1), dtype=np.uint64)
    counts = np.zeros(max(n.value, 1), dtype=np.uint32)
    check(lib.kmu_count_reload_multiple(os.fsencode(path), None, None, None, _p(kmers, C.POINTER(C.c_uint64)),
                                        _p(counts, C.POINTER(C.c_uint32)), n.value, C.byref(n)))
    return {"kmer_size": ks.value, "count_bytes": cb.value, "nb_declared": nd.value, "kmers": kmers[: n.value],
            "counts": counts[: n.value]}


def dump_sketch_params(dirpath, kmer_size, sketch_size, algo="PROB3A", data_t="DNA"):
    """SeqSketcherParams::dump_json (sketcharg.rs:79-106): `sketchparams_dump.json` in a directory."""
    path = os.path.join(dirpath, "sketchparams_dump.json")
    with open(path, "w") as f:
        json.dump({"kmer_size": kmer_size, "sketch_size": sketch_size, "algo": algo, "data_t": data_t}, f)
    return path


def reload_sketch_params(dirpath):
    """SeqSketcherParams::reload_json (sketcharg.rs:109-138)"""
    with open(os.path.join(dirpath, "sketchparams_dump.json")) as f:
        d = json.load(f)
    for key in ("kmer_size", "sketch_size", "algo", "data_t"):
        if key not in d:
            raise ValueError(f"SeqSketcherParams reload: missing {key}")
    return d


def datasketcher(engine, fastx_path, dump_path, kmer_size=8, sketch_size=200, pack=10000, block_size=0):
    """The loop of src/bin/datasketcher.rs:236-300: read packs of `pack` accepted reads, sketch each read with
    ProbMinHash3a (Kmer32bit, canonical + int32_hash, :222-226) on the GPU, append the signatures to the dump.
    block_size > 0 sketches blocks of k-mers instead (BlockSeqSketcher, pack 5000).  -> number of reads sketched"""
    from ._lib import HASH_CANON_INVHASH, KMER32
    n_done = 0
    with FastxReader(fastx_path) as rd:
        if block_size:
            out = BlockSignatureDump(dump_path, sketch_size, kmer_size, block_size)
        else:
            out = SignatureDump(dump_path, sketch_size, kmer_size)
        with out:
            while True:
                reads = rd.next_pack(pack)
                if not reads:
                    break
                batch, _ = engine.batch_from_ascii(reads)
                if block_size:
                    sig, numseq, numblock = engine.blocksketch(batch, kmer_size, sketch_size, block_size)
                    out.write(sig, numseq + np.uint32(n_done), numblock)
                else:
                    out.write(engine.sketch_pmh3a(batch, kmer_size, KMER32, HASH_CANON_INVHASH, sketch_size))
                batch.destroy()
                n_done += len(reads)
    return n_done
