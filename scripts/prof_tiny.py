"""Short reads (300 bases, k = 8, m = 200) through kmu_sketch_pmh3a: the team kernel's table mode."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kmerutils_b200 as kb  # noqa: E402

eng = kb.Engine(0)
n = int(os.environ.get("NREADS", "100000"))
L = int(os.environ.get("LEN", "300"))
nb = np.full(n, L, dtype=np.uint64)
b = eng.batch_synth(5, nb)
import torch
out = torch.empty((n, 200), dtype=torch.int32, device="cuda:0")
for i in range(3):
    eng.sketch_pmh3a(b, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 200, out_device_ptr=out.data_ptr())
    eng.sync()
    t = eng.last_times()
print(n, L, t["kernel_ms"], "ms", n * L / t["kernel_ms"] / 1e6, "Gbases/s")
