"""One call of kmu_sketch_pmh3a_host on the C2 workload (run under ncu --metrics gpu__time_duration.sum for the
per-launch times of every chunk)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kmerutils_b200 as kb  # noqa: E402
from kmerutils_b200 import workloads  # noqa: E402

eng = kb.Engine(0)
nb = workloads.c2_lengths()
batch = eng.batch_synth(2, nb)
packed, off, _ = batch.download()
pin_in = torch.empty(len(packed) + 64, dtype=torch.uint8).pin_memory()
pin_in[: len(packed)] = torch.from_numpy(packed)
out = torch.empty((len(nb), 200), dtype=torch.int32).pin_memory()
for _ in range(int(os.environ.get("CALLS", "1"))):
    eng.sketch_pmh3a_host((pin_in.data_ptr(), pin_in.numel()), off, nb, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 200, out.data_ptr())
    print("host_ms %.2f" % eng.last_times()["host_ms"], end="  ")
print()
print(eng.last_times())
