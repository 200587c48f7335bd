#!/usr/bin/env python
"""Per API call: wall time between synchronisations against the library's own kernel time (ev[0]..ev[1] inside the call).
The difference is host work inside the call (metadata loops, copies, synchronisations)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kmerutils_b200 as kb  # noqa: E402
from kmerutils_b200 import workloads  # noqa: E402


def timeit(eng, name, fn, n=5):
    fn()
    eng.sync()
    t0 = time.perf_counter()
    k = 0.0
    for _ in range(n):
        fn()
        k += eng.last_times()["kernel_ms"]
    eng.sync()
    wall = (time.perf_counter() - t0) * 1e3 / n
    print(f"{name:48s} wall {wall:8.3f} ms   kernels {k / n:8.3f} ms   host {wall - k / n:7.3f} ms", flush=True)


def main():
    import torch
    eng = kb.Engine(0)
    dev = torch.device("cuda", 0)
    # C1
    b1 = eng.batch_synth(1, workloads.c1_lengths())
    s1 = torch.empty((len(b1), 200), dtype=torch.int32, device=dev)
    timeit(eng, "C1 pmh3a 1000 x 1 kb", lambda: eng.sketch_pmh3a(b1, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 200, out_device_ptr=s1.data_ptr()), 20)
    # C2 subset
    nb = workloads.c2_lengths()
    b2 = eng.batch_synth(2, nb)
    s2 = torch.empty((len(b2), 200), dtype=torch.int32, device=dev)
    timeit(eng, "C2 pmh3a 746333 reads", lambda: eng.sketch_pmh3a(b2, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 200, out_device_ptr=s2.data_ptr()))
    s2f = torch.empty((len(b2), 200), dtype=torch.float64, device=dev)
    timeit(eng, "C2 superminhash per read", lambda: eng.sketch_superminhash(b2, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 200, out_device_ptr=s2f.data_ptr()))
    b2.destroy()
    del s2, s2f
    # C5b
    rng = np.random.default_rng(7)
    nres = rng.integers(50, 600, 20000).astype(np.uint64)
    b5 = eng.batch_synth_aa(5, nres)
    timeit(eng, "C5b pmh3a 20000 proteins", lambda: eng.sketch_pmh3a(b5, 12, kb.KMERAA64, kb.HASH_IDENTITY_RAW, 400), 10)
    # C4
    g = eng.batch_synth(4, np.full(148, 5_000_000, dtype=np.uint64))
    sg = torch.empty((148, 12000), dtype=torch.float64, device=dev)
    timeit(eng, "C4 superminhash 148 genomes", lambda: eng.sketch_superminhash(g, 16, kb.KMER16B32, kb.HASH_CANON_INVHASH, 12000, out_device_ptr=sg.data_ptr()))
    timeit(eng, "C4 pmh3a groups 148 genomes", lambda: eng.sketch_pmh3a_groups(g, np.ones(148, dtype=np.uint64), 16, kb.KMER16B32, kb.HASH_CANON_INVHASH, 12000), 3)


if __name__ == "__main__":
    main()
