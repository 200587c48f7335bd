"""Full-size parity on BASELINE.json's headline configuration (C2: 746 333 ONT-like reads, 4.38 Gbases, k = 8,
ProbMinHash3a m = 200): every signature of the GPU path equals the oracle's, including the few dozen reads that go
through the redo launch (failed speculation / wrapped counters).  The oracle needs ~15 s on 16 host threads."""
import hashlib

import numpy as np
import pytest

import kmerutils_b200 as kb
from kmerutils_b200 import workloads

pytestmark = pytest.mark.gpu


def test_c2_all_signatures_match_oracle(engine, oracle):
    nb = workloads.c2_lengths()
    assert len(nb) == 746_333 and int(nb.sum()) == 4_380_000_000
    batch = engine.batch_synth(2, nb)
    sig = engine.sketch_pmh3a(batch, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 200)
    packed, off, _ = batch.download()  # the device's own synthetic reads: same bytes as orc_synth_packed (checked elsewhere)
    batch.destroy()
    packed = np.concatenate([packed, np.zeros(64, np.uint8)])
    want = oracle.sketch_pmh3a_batch(packed, off, nb, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 200)
    bad = np.nonzero((sig != want).any(axis=1))[0]
    assert len(bad) == 0, f"{len(bad)} reads differ, first {bad[:5]} (lengths {nb[bad[:5]]})"
    # checksum of checksums, for the record (same value from the oracle and the GPU)
    assert hashlib.sha256(sig.tobytes()).hexdigest() == hashlib.sha256(want.tobytes()).hexdigest()
    # size-independent properties: a signature slot is 0 or the hash of a canonical 8-mer word of that read
    assert sig.shape == (746_333, 200)
    shortest = int(np.argmin(nb))
    kmers, koff = engine.generate_kmers(engine.batch_from_packed(packed[int(off[shortest]):int(off[shortest]) + 64],
                                                                 np.zeros(1, np.uint64), nb[shortest:shortest + 1]),
                                        8, kb.KMER32, kb.HASH_CANON_INVHASH)
    assert set(sig[shortest].tolist()) <= set(kmers.tolist()) | {0}
