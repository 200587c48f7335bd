"""kmerutils_b200 -- B200-native k-mer engine behind the kmerutils hot-path interface.

The arithmetic lives in hand-written sm_100a CUDA kernels (kmerutils_b200/csrc/) behind the C
ABI of include/kmerutils_b200.h; this package is the host-side mirror of the reference's
operator interface.  There is no CPU fallback: importing works anywhere (so that the symbol
table can be checked), every compute call needs a B200.
"""
from ._lib import (HASHER_FNV, HASHER_NOHASH, HASH_CANON_INVHASH, HASH_CANON_RAW, HASH_IDENTITY_RAW, HASH_INVHASH, HASH_MASKED_VALUE, KMER16B32,
                   KMER32, KMER64, KMERAA32, KMERAA64, KmuError, KmuInvalid, load_library)
from .engine import Engine, KmerCounter, SeqBatch, default_engine, val_dtype

__version__ = "0.1.0"
