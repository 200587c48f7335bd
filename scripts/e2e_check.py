import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import kmerutils_b200 as kb
from kmerutils_b200 import workloads
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
eng = kb.Engine(0)
nb = workloads.c2_lengths()[:n]
batch = eng.batch_synth(2, nb)
ref = eng.sketch_pmh3a(batch, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 200)
packed, off, _ = batch.download()
packed = np.concatenate([packed, np.zeros(64, np.uint8)])
out = np.zeros((len(nb), 200), dtype=np.uint32)
for it in range(2):
    out[:] = 0
    eng.sketch_pmh3a_host(packed, off, nb, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 200, out)
    print("iter", it, "equal:", np.array_equal(out, ref), eng.last_times())
