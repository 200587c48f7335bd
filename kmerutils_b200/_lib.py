"""ctypes binding of the C ABI declared in include/kmerutils_b200.h.

The CUDA library is the product: there is no CPU fallback here.  Loading fails loudly
when libkmerutils_b200.so is missing, and every compute call fails with KmuError when no
B200 is present.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libkmerutils_b200.so")

KMU_OK, KMU_EINVAL, KMU_ECUDA, KMU_ENOMEM, KMU_EOVERFLOW = 0, 1, 2, 3, 4

KMER32, KMER16B32, KMER64, KMERAA32, KMERAA64 = 0, 1, 2, 3, 4
HASH_IDENTITY_RAW, HASH_MASKED_VALUE, HASH_CANON_INVHASH, HASH_CANON_RAW, HASH_INVHASH = 0, 1, 2, 3, 4
HASHER_NOHASH, HASHER_FNV = 0, 1

u8p = C.POINTER(C.c_uint8)
u64p = C.POINTER(C.c_uint64)
vpp = C.POINTER(C.c_void_p)


class KmuTimes(C.Structure):
    _fields_ = [("kernel_ms", C.c_float), ("h2d_ms", C.c_float), ("d2h_ms", C.c_float),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("launches", C.c_uint64),
                ("host_ms", C.c_float)]


class KmuLaunchRec(C.Structure):
    _fields_ = [("mode", C.c_int32), ("table_global", C.c_int32), ("team_warps", C.c_uint32),
                ("teams_per_cta", C.c_uint32), ("grid", C.c_uint32), ("block", C.c_uint32),
                ("smem_bytes", C.c_uint32), ("nseq", C.c_uint64), ("nbases", C.c_uint64), ("nk_max", C.c_uint64),
                ("ms", C.c_float), ("counter_idx", C.c_uint32), ("phase_clocks", C.c_uint64 * 8)]


class KmuSetSketchParams(C.Structure):
    _fields_ = [("b", C.c_double), ("m", C.c_uint64), ("a", C.c_double), ("q", C.c_uint64)]


class KmuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"kmerutils_b200 error {code}: {msg}")
        self.code = code


class KmuInvalid(KmuError, ValueError):
    """The reference panics on this input (bad k for the k-mer type, m < 2, non-ACGT, ...)."""


# every exported symbol of include/kmerutils_b200.h : (restype, argtypes)
SIGNATURES = {
    "kmu_ctx_create": (C.c_int32, [C.c_int32, vpp]),
    "kmu_ctx_destroy": (None, [C.c_void_p]),
    "kmu_last_error": (C.c_char_p, []),
    "kmu_version": (C.c_char_p, []),
    "kmu_launch_count": (C.c_uint64, [C.c_void_p]),
    "kmu_ctx_stream": (C.c_void_p, [C.c_void_p]),
    "kmu_ctx_sync": (C.c_int32, [C.c_void_p]),
    "kmu_seqbatch_from_ptrs": (C.c_int32, [C.c_void_p, C.POINTER(C.c_void_p), u64p, C.c_uint64, vpp]),
    "kmu_seqbatch_from_packed": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, u64p, u64p, C.c_uint64, vpp]),
    "kmu_seqbatch_from_ascii": (C.c_int32, [C.c_void_p, C.c_void_p, u64p, C.c_uint64, C.c_int32, u64p, vpp]),
    "kmu_seqbatch_synth": (C.c_int32, [C.c_void_p, C.c_uint64, u64p, C.c_uint64, vpp]),
    "kmu_seqbatch_from_aa": (C.c_int32, [C.c_void_p, C.c_void_p, u64p, C.c_uint64, C.c_int32, u64p, vpp]),
    "kmu_seqbatch_synth_aa": (C.c_int32, [C.c_void_p, C.c_uint64, u64p, C.c_uint64, vpp]),
    "kmu_seqbatch_alphabet": (C.c_int32, [C.c_void_p]),
    "kmu_seqbatch_sample_reads": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32,
                                              C.c_uint32, vpp]),
    "kmu_seqbatch_slices": (C.c_int32, [C.c_void_p, C.c_void_p, u64p, u64p, u64p, C.c_uint64, vpp]),
    "kmu_seqbatch_view": (C.c_int32, [C.c_void_p, C.c_uint64, C.c_uint64, vpp]),
    "kmu_seqbatch_destroy": (None, [C.c_void_p]),
    "kmu_seqbatch_nseq": (C.c_uint64, [C.c_void_p]),
    "kmu_seqbatch_total_bases": (C.c_uint64, [C.c_void_p]),
    "kmu_seqbatch_packed_bytes": (C.c_uint64, [C.c_void_p]),
    "kmu_seqbatch_download": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, u64p, u64p]),
    "kmu_kmer_count": (C.c_uint64, [C.c_void_p, C.c_uint32]),
    "kmu_generate_kmers": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int32, C.c_int32, C.c_void_p, u64p,
                                       C.c_int32]),
    "kmu_nthash_canonical": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                         C.c_int32]),
    "kmu_sketch_pmh3a": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int32, C.c_int32, C.c_uint32, C.c_void_p,
                                     C.c_int32]),
    "kmu_sketch_pmh3a_host": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, u64p, u64p, C.c_uint64, C.c_uint32,
                                          C.c_int32, C.c_int32, C.c_uint32, C.c_void_p]),
    "kmu_sketch_pmh3a_host_ptrs": (C.c_int32, [C.c_void_p, C.c_void_p, u64p, C.c_uint64, C.c_uint32, C.c_int32, C.c_int32,
                                               C.c_uint32, C.c_void_p]),
    "kmu_sketch_pmh3a_whole": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int32, C.c_int32, C.c_uint32, C.c_void_p,
                                           C.c_int32]),
    "kmu_sketch_pmh3a_groups": (C.c_int32, [C.c_void_p, C.c_void_p, u64p, C.c_uint64, C.c_uint32, C.c_int32, C.c_int32, C.c_uint32,
                                            C.c_void_p, C.c_int32]),
    "kmu_pmh3a_counter_slots": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_uint32, C.c_double, C.c_void_p, C.c_int32]),
    "kmu_pmh3a_weighted": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int32, C.c_uint32, C.c_void_p]),
    "kmu_sketch_superminhash": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int32, C.c_int32, C.c_uint32,
                                            C.c_int32, C.c_int32, C.c_void_p, C.c_int32]),
    "kmu_sketch_setsketch": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int32, C.c_int32,
                                         C.POINTER(KmuSetSketchParams), C.c_int32, C.c_int32, C.c_void_p, C.c_int32]),
    "kmu_signature_jaccard": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32,
                                          C.c_void_p, C.c_int32]),
    "kmu_sketch_superminhash_whole": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int32, C.c_int32, C.c_uint32,
                                                  C.c_int32, C.c_int32, C.c_void_p, C.c_int32]),
    "kmu_count_create": (C.c_int32, [C.c_void_p, C.c_uint32, C.c_int32, C.c_uint32, C.c_uint64, vpp]),
    "kmu_count_destroy": (None, [C.c_void_p]),
    "kmu_count_capacity": (C.c_uint64, [C.c_void_p]),
    "kmu_count_insert_seqs": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "kmu_count_insert_kmers": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int32]),
    "kmu_count_query": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int32]),
    "kmu_count_stats": (C.c_int32, [C.c_void_p, C.c_void_p, u64p, u64p, u64p, u64p]),
    "kmu_count_export": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, u64p]),
    "kmu_count_dump_multiple": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_char_p, C.c_int32, u64p]),
    "kmu_count_reload_multiple": (C.c_int32, [C.c_char_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), u64p, u64p,
                                              C.POINTER(C.c_uint32), C.c_uint64, u64p]),
    "kmu_count_partition": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int32, C.c_int32, C.c_uint32,
                                        C.c_void_p, u64p, C.c_int32]),
    "kmu_fastx_open": (C.c_int32, [C.c_char_p, vpp]),
    "kmu_fastx_close": (None, [C.c_void_p]),
    "kmu_fastx_next_pack": (C.c_int32, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, u64p, u64p]),
    "kmu_fastx_stats": (None, [C.c_void_p, u64p, u64p, u64p, u64p]),
    "kmu_ingest_open": (C.c_int32, [C.c_char_p, C.c_uint32, C.c_uint64, vpp]),
    "kmu_ingest_next": (C.c_int32, [C.c_void_p, vpp, vpp, u64p, vpp]),
    "kmu_ingest_release": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "kmu_ingest_stats": (None, [C.c_void_p, u64p, u64p, u64p, u64p]),
    "kmu_ingest_close": (None, [C.c_void_p]),
    "kmu_sigdump_create": (C.c_int32, [C.c_char_p, C.c_uint32, C.c_uint32, vpp]),
    "kmu_sigdump_write": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "kmu_blockdump_create": (C.c_int32, [C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, vpp]),
    "kmu_blockdump_write": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]),
    "kmu_sigdump_close": (C.c_int32, [C.c_void_p]),
    "kmu_sigdump_read": (C.c_int32, [C.c_char_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                     u64p, C.c_void_p, C.c_uint64, C.c_uint64]),
    "kmu_count_partition_counts": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int32, C.c_int32, C.c_uint32, u64p]),
    "kmu_count_partition_scatter": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int32, C.c_int32, C.c_uint32,
                                                C.POINTER(C.c_void_p), u64p]),
    "kmu_count_exchange_geometry": (C.c_int32, [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]),
    "kmu_count_exchange_scatter": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_uint32, C.c_uint32,
                                               C.c_uint64, C.POINTER(C.c_void_p), u64p, C.POINTER(C.c_int32)]),
    "kmu_count_insert_slabs": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, u64p]),
    "kmu_ipc_alloc": (C.c_int32, [C.c_void_p, C.c_uint64, vpp, C.c_void_p]),
    "kmu_ipc_free": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "kmu_ipc_open": (C.c_int32, [C.c_void_p, C.c_void_p, vpp]),
    "kmu_ipc_close": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "kmu_last_times": (C.c_int32, [C.c_void_p, C.POINTER(KmuTimes)]),
    "kmu_ctx_set_profiling": (C.c_int32, [C.c_void_p, C.c_int32]),
    "kmu_last_launch_profile": (C.c_uint32, [C.c_void_p, C.POINTER(KmuLaunchRec), C.c_uint32]),
}

_LIB = None


def load_library():
    """Load libkmerutils_b200.so (built by kmerutils_b200/build.py); raise if it is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m kmerutils_b200.build` "
            "(this package has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(rc):
    if rc != KMU_OK:
        msg = load_library().kmu_last_error().decode("utf-8", "replace")
        raise (KmuInvalid if rc == KMU_EINVAL else KmuError)(rc, msg)
