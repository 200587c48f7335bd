"""Launch-level profile of kmu_sketch_pmh3a_groups (run under ncu --metrics gpu__time_duration.sum)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kmerutils_b200 as kb

eng = kb.Engine(0)
ng = 3
nb = np.full(ng, 5_000_000, dtype=np.uint64)
b = eng.batch_synth(4, nb)
eng.sketch_pmh3a_groups(b, np.ones(ng, dtype=np.uint64), 16, kb.KMER16B32, kb.HASH_CANON_INVHASH, 12000)
eng.sync()
