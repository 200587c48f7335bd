"""Per-read SetSketch on a C2-like sample (run under ncu --metrics gpu__time_duration.sum for per-launch times)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kmerutils_b200 as kb  # noqa: E402
from kmerutils_b200 import workloads  # noqa: E402

eng = kb.Engine(0)
nb = workloads.c2_lengths()[:120000]
batch = eng.batch_synth(2, nb)
hll = torch.empty((len(nb), 256), dtype=torch.int16, device="cuda:0")
for _ in range(2):
    eng.sketch_setsketch(batch, 8, kb.KMER32, kb.HASH_CANON_INVHASH, (1.001, 256, 20.0, 65534), np.uint16, out_device_ptr=hll.data_ptr())
    eng.sync()
    print(eng.last_times()["kernel_ms"], "ms for", int(nb.sum()), "bases")
