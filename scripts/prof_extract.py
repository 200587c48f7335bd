#!/usr/bin/env python
"""The three materialising kernels on the C2 reads, twice each (for ncu captures: -k regex:run_kernel)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kmerutils_b200 as kb  # noqa: E402
from kmerutils_b200 import workloads  # noqa: E402


def main():
    import torch
    eng = kb.Engine(0)
    dev = torch.device("cuda", 0)
    nb = workloads.c2_lengths()[:200000]
    batch = eng.batch_synth(2, nb)
    out32 = torch.empty(batch.kmer_count(8), dtype=torch.int32, device=dev)
    out64 = torch.empty(batch.kmer_count(31), dtype=torch.int64, device=dev)
    for _ in range(2):
        kb._lib.check(eng.lib.kmu_generate_kmers(eng.ctx, batch.handle, 8, kb.KMER32, kb.HASH_CANON_INVHASH, out32.data_ptr(), None, 1))
        kb._lib.check(eng.lib.kmu_generate_kmers(eng.ctx, batch.handle, 31, kb.KMER64, kb.HASH_CANON_RAW, out64.data_ptr(), None, 1))
        kb._lib.check(eng.lib.kmu_nthash_canonical(eng.ctx, batch.handle, 31, 1, out64.data_ptr(), None, 1))
    torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()
