"""Jaccard estimates between signatures (compute_probminhash_jaccard, seqsketchjaccard.rs:86-108, 423-495)."""
import numpy as np
import pytest

import kmerutils_b200 as kb
from test_pmh3a_gpu import S80

pytestmark = pytest.mark.gpu


# 12 000 slots: gsearch's sketch size (4 rows of u32 / 2 rows of u64 per CTA in shared memory); 40 000 slots of u64 do
# not fit shared memory at all and are read through the caches
@pytest.mark.parametrize("dtype,m", [(np.uint32, 200), (np.uint64, 64), (np.uint16, 4096), (np.float64, 333), (np.float32, 7),
                                     (np.uint32, 12000), (np.float64, 12000), (np.uint64, 40000)])
def test_signature_jaccard_matrix(engine, oracle, dtype, m):
    rng = np.random.default_rng(m)
    na, nb = 37, 101
    base = rng.integers(0, 50, (nb, m))
    b = base.astype(dtype)
    a = base[rng.integers(0, nb, na)].copy()
    a[rng.random(a.shape) < 0.4] = 77  # break about 40 % of the slots
    a = a.astype(dtype)
    got = engine.signature_jaccard(a, b)
    want = (a[:, None, :] == b[None, :, :]).mean(axis=2)
    assert np.array_equal(got, want)
    assert got[0, 0] == oracle.jaccard(a[0], b[0])  # the oracle's compute_probminhash_jaccard


def test_jaccard_index_probminhash3a_reference_inequalities(engine):
    # seqsketchjaccard.rs:742-851 through the one-vs-many entry point: k = 5, m = 4000
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    a, _ = engine.batch_from_ascii([S80])
    others, _ = engine.batch_from_ascii([S80[:40], S80.translate(comp)[::-1], S80])
    j = engine.jaccard_index_probminhash3a(a, others, 5, kb.KMER32, kb.HASH_CANON_INVHASH, 4000)
    assert j.shape == (1, 3)
    assert j[0, 0] >= 0.75 * (40 - 5) / (80 - 5) and j[0, 1] >= 1.0 and j[0, 2] == 1.0
    j_id = engine.jaccard_index_probminhash3a(a, others, 5, kb.KMER32, kb.HASH_IDENTITY_RAW, 4000)
    assert j_id[0, 1] <= 0.1
