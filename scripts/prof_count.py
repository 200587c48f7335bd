import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kmerutils_b200 as kb
eng = kb.Engine(0)
genome = eng.batch_synth(3, np.array([100_000_000], dtype=np.uint64))
reads = eng.batch_sample_reads(genome, 3, 0, 8_000_000, 150, 5000)
ctr = eng.counter(31, kb.KMER64, capacity=int(8_000_000 * 120 * 0.45))
ctr.insert_seqs(reads, canonical=True)
print(eng.last_times())
