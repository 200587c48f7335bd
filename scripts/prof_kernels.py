#!/usr/bin/env python
"""Small fixed workload that launches each secondary kernel once (for ncu captures)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kmerutils_b200 as kb  # noqa: E402
from kmerutils_b200 import workloads  # noqa: E402


def main():
    import torch
    eng = kb.Engine(0)
    dev = torch.device("cuda", 0)
    nb = workloads.c2_lengths()[:60000]
    batch = eng.batch_synth(2, nb)
    out32 = torch.empty(batch.kmer_count(8), dtype=torch.int32, device=dev)
    out64 = torch.empty(batch.kmer_count(31), dtype=torch.int64, device=dev)
    for _ in range(2):
        kb._lib.check(eng.lib.kmu_generate_kmers(eng.ctx, batch.handle, 8, kb.KMER32, kb.HASH_CANON_INVHASH, out32.data_ptr(), None, 1))
        kb._lib.check(eng.lib.kmu_generate_kmers(eng.ctx, batch.handle, 31, kb.KMER64, kb.HASH_CANON_RAW, out64.data_ptr(), None, 1))
        kb._lib.check(eng.lib.kmu_nthash_canonical(eng.ctx, batch.handle, 31, 1, out64.data_ptr(), None, 1))
    genome = eng.batch_synth(3, np.array([100_000_000], dtype=np.uint64))
    reads = eng.batch_sample_reads(genome, 3, 0, 4_000_000, 150, 5000)
    ctr = eng.counter(31, kb.KMER64, capacity=int(4_000_000 * 120 * 0.45))
    ctr.insert_seqs(reads, canonical=True)
    ctr.insert_seqs(reads, canonical=True)
    print("ok", ctr.stats()["nb_distinct"])


if __name__ == "__main__":
    main()
