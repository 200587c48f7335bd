#!/usr/bin/env python
"""Is the per-rank skew of the 8-GPU counting insertion (ranks 0 and 6 fast, 1 and 7 slow) a property of the KEYS an owner
gets?  One GPU plays all eight owners in turn: the C3 reads are bucketed by owner (kmu_count_exchange_scatter into eight local
buffers), then every owner's segment is inserted into a fresh table of the same size and the call is timed."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import kmerutils_b200 as kb  # noqa: E402
from kmerutils_b200.dist import exchange_slab_cap  # noqa: E402


def main():
    import torch
    eng = kb.Engine(0)
    k, nown = 31, 8
    reads = int(os.environ.get("SKEW_READS", 26666667))
    genome = eng.batch_synth(3, np.array([100_000_000], dtype=np.uint64))
    batch = eng.batch_sample_reads(genome, 3, 0, reads, 150, 5000)
    nk = batch.kmer_count(k)
    probe = eng.counter(k, kb.KMER64, capacity=1 << 29, count_bits=8)
    nreg = probe.exchange_regions(nown)
    slab_cap = exchange_slab_cap(nk, nown, nreg)
    bufs = [torch.empty(nreg * slab_cap, dtype=torch.int64, device="cuda:0") for _ in range(nown)]  # only sender 0's slab of every buffer is used
    sent, ovf = probe.exchange_scatter(batch, nown, 0, slab_cap, [b.data_ptr() for b in bufs], True)
    assert not ovf
    probe.destroy()
    print("keys per owner:", [int(x) for x in np.asarray(sent).reshape(nown, nreg).sum(axis=1)])
    for rep in range(2):
        times = []
        for o in range(nown):
            ctr = eng.counter(k, kb.KMER64, capacity=1 << 29, count_bits=8)
            cnt = np.zeros((nown, nreg), dtype=np.uint64)
            cnt[0] = np.asarray(sent).reshape(nown, nreg)[o]  # this GPU was sender 0 of every owner's buffer
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ctr.insert_slabs(bufs[o].data_ptr(), slab_cap, cnt)
            torch.cuda.synchronize()
            times.append((time.perf_counter() - t0) * 1e3)
            st = ctr.stats()
            ctr.destroy()
        print("rep", rep, "insert ms per owner:", [round(t, 1) for t in times], "distinct of the last:", st["nb_distinct"])


if __name__ == "__main__":
    main()
