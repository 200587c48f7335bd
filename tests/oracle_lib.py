"""ctypes binding of the CPU oracle (oracle/libkmer_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(kmerutils_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_SO = os.path.join(ORACLE_DIR, "libkmer_oracle.so")

KMER32, KMER16B32, KMER64, KMERAA32, KMERAA64 = 0, 1, 2, 3, 4
HASH_IDENTITY_RAW, HASH_MASKED_VALUE, HASH_CANON_INVHASH, HASH_CANON_RAW, HASH_INVHASH = 0, 1, 2, 3, 4

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
f64p = C.POINTER(C.c_double)


def build_oracle(force=False):
    src = [os.path.join(ORACLE_DIR, f) for f in ("kmer_oracle.cpp", "kmer_oracle.hpp", "Makefile", "det_math.hpp",
                                                 "zig_exp_tables.h")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return _SO


def _ptr(a, t):
    return a.ctypes.data_as(t)


class Oracle:
    def __init__(self):
        build_oracle()
        L = self.L = C.CDLL(_SO)
        L.orc_pack_2bit.restype = C.c_int64
        L.orc_pack_2bit.argtypes = [u8p, C.c_uint64, u8p]
        L.orc_encode_and_add_2bit.restype = C.c_uint64
        L.orc_encode_and_add_2bit.argtypes = [u8p, C.c_uint64, u8p]
        L.orc_count_non_acgt.restype = C.c_uint64
        L.orc_count_non_acgt.argtypes = [u8p, C.c_uint64]
        L.orc_get_base.restype = C.c_uint8
        L.orc_get_base.argtypes = [u8p, C.c_uint64]
        L.orc_unpack_2bit.argtypes = [u8p, C.c_uint64, u8p]
        L.orc_seq_revcomp_2bit.argtypes = [u8p, C.c_uint64, u8p]
        for name in ("orc_kmer_build",):
            getattr(L, name).restype = C.c_uint64
            getattr(L, name).argtypes = [C.c_uint64, C.c_int, C.c_int]
        L.orc_kmer_push.restype = C.c_uint64
        L.orc_kmer_push.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_uint8]
        L.orc_kmer_revcomp.restype = C.c_uint64
        L.orc_kmer_revcomp.argtypes = [C.c_uint64, C.c_int, C.c_int]
        L.orc_kmer_cmp.restype = C.c_int
        L.orc_kmer_cmp.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int]
        L.orc_kmer_compressed_value.restype = C.c_uint64
        L.orc_kmer_compressed_value.argtypes = [C.c_uint64, C.c_int, C.c_int]
        L.orc_generate_kmers.restype = C.c_uint64
        L.orc_generate_kmers.argtypes = [u8p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int, u64p]
        L.orc_apply_hash.restype = C.c_uint64
        L.orc_apply_hash.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int]
        L.orc_int32_hash.restype = C.c_uint32
        L.orc_int32_hash.argtypes = [C.c_uint32]
        L.orc_int64_hash.restype = C.c_uint64
        L.orc_int64_hash.argtypes = [C.c_uint64]
        L.orc_nthash_init.restype = C.c_uint64
        L.orc_nthash_init.argtypes = [C.c_uint64, C.c_int, C.c_int]
        L.orc_nthash_canonical_init.restype = C.c_int
        L.orc_nthash_canonical_init.argtypes = [C.c_uint64, C.c_int, C.c_int, u64p, u64p, u64p]
        L.orc_nthash_mult.argtypes = [C.c_uint64, C.c_int, u64p, C.c_int]
        L.orc_nthash_cycle.restype = C.c_uint64
        L.orc_nthash_cycle.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_uint64, C.c_uint8]
        L.orc_nthash_canonical_cycle.restype = C.c_int
        L.orc_nthash_canonical_cycle.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_uint8, u64p, u64p, u64p]
        L.orc_nohash_seed.restype = C.c_uint64
        L.orc_nohash_seed.argtypes = [C.c_uint64, C.c_int]
        L.orc_fnv1a_seed.restype = C.c_uint64
        L.orc_fnv1a_seed.argtypes = [C.c_uint64, C.c_int]
        L.orc_xoshiro_seed.argtypes = [C.c_uint64, u64p]
        L.orc_xoshiro_next.restype = C.c_uint64
        L.orc_xoshiro_next.argtypes = [u64p]
        L.orc_pmh3a_weighted.argtypes = [u64p, f64p, C.c_uint64, C.c_uint32, C.c_int, u64p]
        L.orc_sketch_pmh3a_seq.argtypes = [u8p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_uint32, u64p]
        L.orc_sketch_pmh3a_batch.argtypes = [u8p, u64p, u64p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_uint32,
                                             C.c_void_p, C.c_int, C.c_int]
        L.orc_sketch_pmh3a_seqs.argtypes = [u8p, u64p, u64p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_uint32, u64p]
        L.orc_blocksketch_seq.restype = C.c_uint64
        L.orc_blocksketch_seq.argtypes = [u8p, C.c_uint64, C.c_int, C.c_uint32, C.c_uint64, u32p, C.c_uint64]
        L.orc_jaccard_equal_fraction.restype = C.c_double
        L.orc_jaccard_equal_fraction.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int]
        L.orc_sketch_superminhash.argtypes = [u8p, u64p, u64p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int,
                                              C.c_int, C.c_void_p]
        L.orc_sketch_superminhash_batch.argtypes = [u8p, u64p, u64p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_uint32,
                                                    C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orc_sketch_setsketch.argtypes = [u8p, u64p, u64p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_double, C.c_uint64,
                                           C.c_double, C.c_uint64, C.c_int, C.c_void_p]
        L.orc_sketch_setsketch_batch.argtypes = [u8p, u64p, u64p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_double,
                                                 C.c_uint64, C.c_double, C.c_uint64, C.c_int, C.c_void_p, C.c_int]
        L.orc_det_log.restype = C.c_double
        L.orc_det_log.argtypes = [C.c_double]
        L.orc_det_exp.restype = C.c_double
        L.orc_det_exp.argtypes = [C.c_double]
        L.orc_det_expm1.restype = C.c_double
        L.orc_det_expm1.argtypes = [C.c_double]
        L.orc_libm_divergence.argtypes = [C.c_uint64, C.c_uint64, C.c_double, C.c_uint64, C.c_double, C.c_uint64, C.c_uint32,
                                          C.c_uint32, C.c_int, u64p]
        L.orc_exp1_from_seed.restype = C.c_double
        L.orc_exp1_from_seed.argtypes = [C.c_uint64, C.c_int]
        L.orc_count_kmers.restype = C.c_uint64
        L.orc_count_kmers.argtypes = [u8p, u64p, u64p, C.c_uint64, C.c_int, C.c_int, C.c_int, u64p, u64p, C.c_uint64]
        L.orc_dispatch.restype = C.c_uint64
        L.orc_dispatch.argtypes = [C.c_uint64, C.c_int, C.c_uint64]
        L.orc_synth_packed.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, u8p]
        L.orc_synth_ascii.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, u8p]
        L.orc_synth_aa.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, u8p]
        L.orc_aa_filter.restype = C.c_uint64
        L.orc_aa_filter.argtypes = [u8p, C.c_uint64, u8p]
        L.orc_sample_read.argtypes = [u8p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, u8p]
        L.orc_hardware_threads.restype = C.c_int

    # ---- sequences -------------------------------------------------------------
    def pack_2bit(self, ascii_bytes):
        a = np.frombuffer(bytes(ascii_bytes), dtype=np.uint8)
        out = np.zeros((len(a) + 3) // 4, dtype=np.uint8)
        n = self.L.orc_pack_2bit(_ptr(a, u8p), len(a), _ptr(out, u8p))
        if n < 0:
            raise ValueError("pattern not a code in alphabet_2b")
        return out

    def encode_and_add(self, ascii_bytes):
        a = np.frombuffer(bytes(ascii_bytes), dtype=np.uint8)
        out = np.zeros((len(a) + 3) // 4 + 1, dtype=np.uint8)
        kept = self.L.orc_encode_and_add_2bit(_ptr(a, u8p), len(a), _ptr(out, u8p))
        return out[: (kept + 3) // 4].copy(), int(kept)

    def count_non_acgt(self, ascii_bytes):
        a = np.frombuffer(bytes(ascii_bytes), dtype=np.uint8)
        return int(self.L.orc_count_non_acgt(_ptr(a, u8p), len(a)))

    def unpack_2bit(self, packed, nbases):
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        out = np.zeros(nbases, dtype=np.uint8)
        self.L.orc_unpack_2bit(_ptr(packed, u8p), nbases, _ptr(out, u8p))
        return out.tobytes()

    def seq_revcomp(self, packed, nbases):
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        out = np.zeros_like(packed)
        self.L.orc_seq_revcomp_2bit(_ptr(packed, u8p), nbases, _ptr(out, u8p))
        return out

    # ---- k-mers ------------------------------------------------------------------
    def generate_kmers(self, packed, nbases, k, ktype, begin=0, end=None):
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        end = nbases if end is None else end
        out = np.zeros(max(nbases, 1), dtype=np.uint64)
        n = self.L.orc_generate_kmers(_ptr(packed, u8p), nbases, begin, end, k, ktype, _ptr(out, u64p))
        if n == 2**64 - 1:
            raise ValueError("kmer size not supported by kmer type")
        if n == 2**64 - 2:
            raise ValueError("encode: not a code in alphabet for amino acid")
        return out[:n].copy()

    def apply_hash(self, words, k, ktype, kind):
        return np.array([self.L.orc_apply_hash(int(w), k, ktype, kind) for w in words], dtype=np.uint64)

    def nthash_canonical(self, word, k, ktype):
        f, r, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        s = self.L.orc_nthash_canonical_init(int(word), k, ktype, C.byref(f), C.byref(r), C.byref(c))
        return f.value, r.value, c.value, s

    def nthash_mult(self, h0, k, n):
        out = np.zeros(n, dtype=np.uint64)
        self.L.orc_nthash_mult(int(h0), k, _ptr(out, u64p), n)
        return out

    # ---- sketches ----------------------------------------------------------------
    def pmh3a_weighted(self, keys, weights, m, key_bytes):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        weights = np.ascontiguousarray(weights, dtype=np.float64)
        sig = np.zeros(m, dtype=np.uint64)
        self.L.orc_pmh3a_weighted(_ptr(keys, u64p), _ptr(weights, f64p), len(keys), m, key_bytes, _ptr(sig, u64p))
        return sig

    def sketch_pmh3a_seq(self, packed, nbases, k, ktype, kind, m):
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        sig = np.zeros(m, dtype=np.uint64)
        self.L.orc_sketch_pmh3a_seq(_ptr(packed, u8p), nbases, k, ktype, kind, m, _ptr(sig, u64p))
        return sig

    def sketch_pmh3a_batch(self, packed, byte_off, nbases, k, ktype, kind, m, nthreads=0):
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        byte_off = np.ascontiguousarray(byte_off, dtype=np.uint64)
        nbases = np.ascontiguousarray(nbases, dtype=np.uint64)
        nseq = len(nbases)
        sig_bytes = 4 if ktype in (KMER32, KMER16B32, KMERAA32) else 8
        sig = np.zeros((nseq, m), dtype=np.uint32 if sig_bytes == 4 else np.uint64)
        if nthreads <= 0:
            nthreads = self.hardware_threads()
        self.L.orc_sketch_pmh3a_batch(_ptr(packed, u8p), _ptr(byte_off, u64p), _ptr(nbases, u64p), nseq, k, ktype,
                                      kind, m, sig.ctypes.data_as(C.c_void_p), sig_bytes, nthreads)
        return sig

    def sketch_pmh3a_seqs(self, packed, byte_off, nbases, k, ktype, kind, m):
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        byte_off = np.ascontiguousarray(byte_off, dtype=np.uint64)
        nbases = np.ascontiguousarray(nbases, dtype=np.uint64)
        sig = np.zeros(m, dtype=np.uint64)
        self.L.orc_sketch_pmh3a_seqs(_ptr(packed, u8p), _ptr(byte_off, u64p), _ptr(nbases, u64p), len(nbases), k,
                                     ktype, kind, m, _ptr(sig, u64p))
        return sig

    def blocksketch_seq(self, packed, nbases, k, m, block_size):
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        nb = (nbases + block_size - 1) // block_size
        sig = np.zeros((max(nb, 1), m), dtype=np.uint32)
        n = self.L.orc_blocksketch_seq(_ptr(packed, u8p), nbases, k, m, block_size, _ptr(sig, u32p), nb)
        return sig[:n]

    def jaccard(self, a, b):
        a = np.ascontiguousarray(a)
        b = np.ascontiguousarray(b)
        assert a.dtype == b.dtype and a.shape == b.shape
        return float(self.L.orc_jaccard_equal_fraction(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p),
                                                       a.shape[-1], a.dtype.itemsize))

    def sketch_superminhash_batch(self, packed, byte_off, nbases, k, ktype, kind, m, hasher=0, dtype=np.float64,
                                  nthreads=0):
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        byte_off = np.ascontiguousarray(byte_off, dtype=np.uint64)
        nbases = np.ascontiguousarray(nbases, dtype=np.uint64)
        out = np.zeros((len(nbases), m), dtype=dtype)
        if nthreads <= 0:
            nthreads = self.hardware_threads()
        self.L.orc_sketch_superminhash_batch(_ptr(packed, u8p), _ptr(byte_off, u64p), _ptr(nbases, u64p), len(nbases), k,
                                             ktype, kind, m, hasher, out.dtype.itemsize,
                                             out.ctypes.data_as(C.c_void_p), nthreads)
        return out

    def sketch_superminhash_seqs(self, packed, byte_off, nbases, k, ktype, kind, m, hasher=0, dtype=np.float64):
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        byte_off = np.ascontiguousarray(byte_off, dtype=np.uint64)
        nbases = np.ascontiguousarray(nbases, dtype=np.uint64)
        out = np.zeros(m, dtype=dtype)
        self.L.orc_sketch_superminhash(_ptr(packed, u8p), _ptr(byte_off, u64p), _ptr(nbases, u64p), len(nbases), k, ktype,
                                       kind, m, hasher, out.dtype.itemsize, out.ctypes.data_as(C.c_void_p))
        return out

    def sketch_setsketch_batch(self, packed, byte_off, nbases, k, ktype, kind, params=(1.001, 4096, 20.0, 65534),
                               dtype=np.uint16, nthreads=0):
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        byte_off = np.ascontiguousarray(byte_off, dtype=np.uint64)
        nbases = np.ascontiguousarray(nbases, dtype=np.uint64)
        b, m, a, q = params
        out = np.zeros((len(nbases), m), dtype=dtype)
        if nthreads <= 0:
            nthreads = self.hardware_threads()
        self.L.orc_sketch_setsketch_batch(_ptr(packed, u8p), _ptr(byte_off, u64p), _ptr(nbases, u64p), len(nbases), k,
                                          ktype, kind, b, m, a, q, out.dtype.itemsize, out.ctypes.data_as(C.c_void_p),
                                          nthreads)
        return out

    def sketch_setsketch_seqs(self, packed, byte_off, nbases, k, ktype, kind, params=(1.001, 4096, 20.0, 65534),
                              dtype=np.uint16):
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        byte_off = np.ascontiguousarray(byte_off, dtype=np.uint64)
        nbases = np.ascontiguousarray(nbases, dtype=np.uint64)
        b, m, a, q = params
        out = np.zeros(m, dtype=dtype)
        self.L.orc_sketch_setsketch(_ptr(packed, u8p), _ptr(byte_off, u64p), _ptr(nbases, u64p), len(nbases), k, ktype,
                                    kind, b, m, a, q, out.dtype.itemsize, out.ctypes.data_as(C.c_void_p))
        return out

    # ---- counting ----------------------------------------------------------------
    def count_kmers(self, packed, byte_off, nbases, k, ktype, canonical=True):
        """-> (distinct compressed canonical values ascending, multiplicities)"""
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        byte_off = np.ascontiguousarray(byte_off, dtype=np.uint64)
        nbases = np.ascontiguousarray(nbases, dtype=np.uint64)
        cap = int(np.maximum(nbases.astype(np.int64) - k + 1, 0).sum()) + 1
        keys = np.zeros(cap, dtype=np.uint64)
        cnts = np.zeros(cap, dtype=np.uint64)
        n = self.L.orc_count_kmers(_ptr(packed, u8p), _ptr(byte_off, u64p), _ptr(nbases, u64p), len(nbases), k, ktype,
                                   int(bool(canonical)), _ptr(keys, u64p), _ptr(cnts, u64p), cap)
        if n == 2**64 - 1:
            raise ValueError("kmer size not supported by kmer type")
        return keys[:n].copy(), cnts[:n].copy()

    def dispatch(self, values, ktype, nb_receiver):
        return np.array([self.L.orc_dispatch(int(v), ktype, nb_receiver) for v in values], dtype=np.uint64)

    # ---- synthetic ---------------------------------------------------------------
    def synth_packed(self, seed, first_base, nbases):
        out = np.zeros((nbases + 3) // 4, dtype=np.uint8)
        self.L.orc_synth_packed(seed, first_base, nbases, _ptr(out, u8p))
        return out

    def synth_ascii(self, seed, first_base, nbases):
        out = np.zeros(nbases, dtype=np.uint8)
        self.L.orc_synth_ascii(seed, first_base, nbases, _ptr(out, u8p))
        return out.tobytes()

    def synth_aa(self, seed, first_res, nres):
        out = np.zeros(nres, dtype=np.uint8)
        self.L.orc_synth_aa(seed, first_res, nres, _ptr(out, u8p))
        return out.tobytes()

    def aa_filter(self, ascii_bytes):
        a = np.frombuffer(bytes(ascii_bytes), dtype=np.uint8)
        out = np.zeros(len(a) + 1, dtype=np.uint8)
        n = self.L.orc_aa_filter(_ptr(a, u8p), len(a), _ptr(out, u8p))
        return out[:n].tobytes()

    def sample_reads(self, genome_packed, glen, seed, first_read, nreads, read_len=150, err_ppm=5000):
        g = np.ascontiguousarray(genome_packed, dtype=np.uint8)
        out = []
        buf = np.zeros(read_len, dtype=np.uint8)
        for r in range(first_read, first_read + nreads):
            self.L.orc_sample_read(_ptr(g, u8p), glen, seed, r, read_len, err_ppm, _ptr(buf, u8p))
            out.append(buf.tobytes())
        return out

    def libm_divergence(self, nkeys, seed=1, params=(1.001, 4096, 20.0, 65534), points=4, m_pmh=2, nthreads=0):
        """Deviations of the deterministic ln / exp / expm1 from the platform libm on arguments drawn as the sketchers draw
        them -> dict of the 9 counters of orc_libm_divergence."""
        out = np.zeros(9, dtype=np.uint64)
        b, m, a, q = params
        self.L.orc_libm_divergence(int(nkeys), int(seed), float(b), int(m), float(a), int(q), int(points), int(m_pmh), int(nthreads),
                                   _ptr(out, u64p))
        names = ["ln_evals", "ln_bits_differ", "keys_with_different_registers", "zig_evals", "zig_bits_differ", "zig_wedge_decision_differs",
                 "expm1_evals", "expm1_bits_differ", "expm1_decision_differs"]
        return {n: int(v) for n, v in zip(names, out)}

    def hardware_threads(self):
        return int(self.L.orc_hardware_threads())


_ORACLE = None


def get_oracle():
    global _ORACLE
    if _ORACLE is None:
        _ORACLE = Oracle()
    return _ORACLE
