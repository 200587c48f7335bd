"""Writes tests/golden/oracle_fixtures.json: outputs of the CPU oracle (oracle/kmer_oracle.cpp) on small fixed inputs.

What these fixtures are: a regression pin of the ORACLE (and, through tests/test_golden_gpu.py, of the CUDA path) --
any later change to either that moves a single bit of a signature fails a test.  What they are not: outputs of the Rust
crate.  The reference cannot be run in this environment (no Rust toolchain, sketch arithmetic in the un-vendored
`probminhash` crate), so for signatures the oracle itself stays "parity unpinned" (DESIGN.md 3).

Run from the repo root:  python tests/golden/make_oracle_fixtures.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import get_oracle, KMER32, KMER16B32, KMER64, KMERAA64  # noqa: E402

HASH_IDENTITY_RAW, HASH_MASKED_VALUE, HASH_CANON_INVHASH, HASH_CANON_RAW = 0, 1, 2, 3
S80 = b"TCAAAGGGAAACATTCAAAATCAGTATGCGCCCGTTCAGTTACGTATTGCTCTCGCTAATGAGATGGGCTGGGTACAGAG"
LENGTHS = [1000, 37, 8, 150, 5000, 2500]   # synthetic reads, SplitMix64 stream of seed 1 (kmu_seqbatch_synth layout)
SEED = 1


def layout(nb):
    sizes = ((nb + 3) // 4 + 15) // 16 * 16
    off = np.zeros(len(nb), dtype=np.uint64)
    off[1:] = np.cumsum(sizes)[:-1]
    return off, int(sizes.sum())


def synth_batch(orc):
    nb = np.array(LENGTHS, dtype=np.uint64)
    off, total = layout(nb)
    packed = np.zeros(total + 64, dtype=np.uint8)
    first = 0
    for i, L in enumerate(LENGTHS):
        packed[int(off[i]): int(off[i]) + (L + 3) // 4] = orc.synth_packed(SEED, first, L)
        first += L
    return packed, off, nb


def hexrows(a):
    a = np.ascontiguousarray(a)
    return [row.tobytes().hex() for row in a.reshape(a.shape[0], -1)]


def make():
    orc = get_oracle()
    packed, off, nb = synth_batch(orc)
    out = {"_comment": "oracle outputs (little-endian bytes as hex); made by tests/golden/make_oracle_fixtures.py -- pins the "
                       "oracle and the CUDA path against drift, NOT a vector of the Rust crate",
           "seed": SEED, "lengths": LENGTHS, "s80": S80.decode()}
    s80p = orc.pack_2bit(S80)
    out["s80_kmers"] = {
        "k8_kmer32_canon_invhash": orc.apply_hash(orc.generate_kmers(s80p, 80, 8, KMER32), 8, KMER32, HASH_CANON_INVHASH).astype(np.uint32).tobytes().hex(),
        "k16_kmer16b32_raw": orc.generate_kmers(s80p, 80, 16, KMER16B32).astype(np.uint32).tobytes().hex(),
        "k31_kmer64_canon": orc.apply_hash(orc.generate_kmers(s80p, 80, 31, KMER64), 31, KMER64, HASH_CANON_RAW).tobytes().hex(),
    }
    nth = []
    for w in orc.generate_kmers(s80p, 80, 16, KMER16B32)[:8]:
        f, r, c, s = orc.nthash_canonical(w, 16, KMER16B32)
        nth.append([f"{c:016x}", int(s)])
    out["s80_nthash_k16_first8"] = nth
    out["pmh3a"] = {
        "k8_kmer32_m64": hexrows(orc.sketch_pmh3a_batch(packed, off, nb, 8, KMER32, HASH_CANON_INVHASH, 64)),
        "k16_kmer16b32_m64": hexrows(orc.sketch_pmh3a_batch(packed, off, nb, 16, KMER16B32, HASH_CANON_INVHASH, 64)),
        "k21_kmer64_m64": hexrows(orc.sketch_pmh3a_batch(packed, off, nb, 21, KMER64, HASH_CANON_INVHASH, 64)),
        "whole_k8_kmer32_m64": orc.sketch_pmh3a_seqs(packed, off, nb, 8, KMER32, HASH_CANON_INVHASH, 64).astype(np.uint32).tobytes().hex(),
    }
    out["superminhash"] = {
        "k8_kmer32_m64_f64_nohash": hexrows(orc.sketch_superminhash_batch(packed, off, nb, 8, KMER32, HASH_CANON_INVHASH, 64, 0, np.float64)),
        "k16_kmer16b32_m64_f32_fnv": hexrows(orc.sketch_superminhash_batch(packed, off, nb, 16, KMER16B32, HASH_CANON_INVHASH, 64, 1, np.float32)),
    }
    prm = (1.001, 64, 20.0, 65534)
    out["setsketch"] = {
        "params": list(prm),
        "k8_kmer32_u16": hexrows(orc.sketch_setsketch_batch(packed, off, nb, 8, KMER32, HASH_CANON_INVHASH, prm, np.uint16)),
        "whole_k21_kmer64_u16": orc.sketch_setsketch_seqs(packed, off, nb, 21, KMER64, HASH_CANON_INVHASH, prm, np.uint16).tobytes().hex(),
    }
    keys, cnts = orc.count_kmers(packed, off, nb, 31, KMER64, True)
    keys8, cnts8 = orc.count_kmers(packed, off, nb, 8, KMER32, True)
    out["count"] = {
        "k31": {"nb_distinct": int(len(keys)), "nb_unique": int((cnts == 1).sum()), "xor_keys": f"{int(np.bitwise_xor.reduce(keys)):016x}"},
        "k8": {"nb_distinct": int(len(keys8)), "nb_unique": int((cnts8 == 1).sum()), "max_count": int(cnts8.max()),
               "sum_count_times_key_mod64": f"{int((keys8 * cnts8).sum(dtype=np.uint64)):016x}"},
    }
    return out


if __name__ == "__main__":
    with open(os.path.join(HERE, "oracle_fixtures.json"), "w") as f:
        json.dump(make(), f, indent=1)
    print("written")
