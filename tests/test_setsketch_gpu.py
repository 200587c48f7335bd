"""GPU parity of SetSketch (HyperLogLogSketch, setsketchert.rs:648-896) against the oracle: bit-exact registers."""
import numpy as np
import pytest

import kmerutils_b200 as kb
from test_aa_gpu import aa_oracle_batch
from test_pmh3a_gpu import oracle_batch

pytestmark = pytest.mark.gpu

SMALL = (1.001, 256, 20.0, 65534)
DEFAULT = (1.001, 4096, 20.0, 65534)


def check_batch(engine, oracle, seed, nb, k, ktype, kind, params, dtype):
    nb = np.asarray(nb, dtype=np.uint64)
    batch = engine.batch_synth(seed, nb)
    packed, off = oracle_batch(oracle, seed, nb)
    got = engine.sketch_setsketch(batch, k, ktype, kind, params, dtype)
    want = oracle.sketch_setsketch_batch(packed, off, nb, k, ktype, kind, params, dtype)
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert len(bad) == 0, f"sequences {bad[:10]} (lengths {nb[bad[:10]]}) differ"
    return batch, packed, off, got


@pytest.mark.parametrize("dtype", [np.uint16, np.uint32, np.uint64])
def test_setsketch_per_sequence_small_m(engine, oracle, dtype):
    rng = np.random.default_rng(3)
    # exact path (few k-mers relative to m), speculative path, multi-warp teams
    nb = np.concatenate([[1, 20, 21, 22, 100, 1000, 4000, 4200, 5000, 9000, 30000, 100000, 400000],
                         rng.integers(21, 20000, 40)])
    check_batch(engine, oracle, 5, nb, 21, kb.KMER64, kb.HASH_CANON_INVHASH, SMALL, dtype)


def test_setsketch_failed_speculation_paths(engine, oracle):
    """Sequences whose distinct k-mer count is far below what the speculative level assumes: tandem repeats (the
    speculation fails, the redo at the cautious level fails too, the exact path finishes), a half-repetitive read (the
    redo succeeds) and small-key-space reads (k = 6: 2080 canonical keys whatever the length)."""
    rng = np.random.default_rng(17)
    unit = bytes(rng.choice(list(b"ACGT"), 53).astype(np.uint8))
    rnd = bytes(rng.choice(list(b"ACGT"), 60000).astype(np.uint8))
    seqs = [unit * 2000, unit * 300, rnd[:30000] + unit * 600, rnd, (unit * 40 + rnd[:500]) * 20]
    batch, _ = engine.batch_from_ascii(seqs)
    packed, off, nb = batch.download()
    packed = np.concatenate([packed, np.zeros(64, np.uint8)])
    for k, ktype, prm in [(21, kb.KMER64, SMALL), (6, kb.KMER32, SMALL), (12, kb.KMER32, (1.001, 64, 20.0, 65534))]:
        got = engine.sketch_setsketch(batch, k, ktype, kb.HASH_CANON_INVHASH, prm, np.uint16)
        want = oracle.sketch_setsketch_batch(packed, off, nb, k, ktype, kb.HASH_CANON_INVHASH, prm, np.uint16)
        bad = np.nonzero((got != want).any(axis=1))[0]
        assert len(bad) == 0, f"k={k}: sequences {bad} differ"


def test_setsketch_default_params(engine, oracle):
    nb = np.array([300, 70000, 200000, 1500000], dtype=np.uint64)
    check_batch(engine, oracle, 6, nb, 16, kb.KMER16B32, kb.HASH_CANON_INVHASH, DEFAULT, np.uint16)


@pytest.mark.parametrize("params", [(2.0, 64, 20.0, 62), (1.2, 1000, 5.0, 200), (1.001, 37, 20.0, 65534)])
def test_setsketch_other_params(engine, oracle, params):
    rng = np.random.default_rng(9)
    nb = np.concatenate([[5, 12, 13, 50000], rng.integers(12, 9000, 30)])
    check_batch(engine, oracle, 7, nb, 12, kb.KMER32, kb.HASH_CANON_INVHASH, params, np.uint32)


def test_setsketch_whole_file(engine, oracle):
    # HyperLogLogSketch::sketch_compressedkmer_seqs: one sketch over all sequences = max-merge of the blocks
    rng = np.random.default_rng(13)
    nb = np.concatenate([[600000, 250000, 10], rng.integers(21, 30000, 50)]).astype(np.uint64)
    batch = engine.batch_synth(8, nb)
    packed, off = oracle_batch(oracle, 8, nb)
    got = engine.sketch_setsketch(batch, 21, kb.KMER64, kb.HASH_CANON_INVHASH, SMALL, np.uint16, whole=True)
    want = oracle.sketch_setsketch_seqs(packed, off, nb, 21, kb.KMER64, kb.HASH_CANON_INVHASH, SMALL, np.uint16)
    assert np.array_equal(got, want)
    per_seq = engine.sketch_setsketch(batch, 21, kb.KMER64, kb.HASH_CANON_INVHASH, SMALL, np.uint16)
    assert np.array_equal(per_seq.max(axis=0), got)  # SetSketcher::merge (setsketchert.rs:876-882)
    # small whole batch: exact path
    few = engine.batch_synth(8, nb[2:20])
    got2 = engine.sketch_setsketch(few, 21, kb.KMER64, kb.HASH_CANON_INVHASH, DEFAULT, np.uint16, whole=True)
    p2, o2 = oracle_batch(oracle, 8, nb[2:20])
    # (same stream seed but the batch starts at base 0 again: compare with the oracle on the same layout)
    want2 = oracle.sketch_setsketch_seqs(p2, o2, nb[2:20], 21, kb.KMER64, kb.HASH_CANON_INVHASH, DEFAULT, np.uint16)
    assert np.array_equal(got2, want2)


def test_setsketch_amino_acids(engine, oracle):
    rng = np.random.default_rng(21)
    nres = np.concatenate([[3, 12, 13, 40000], np.clip(np.rint(np.exp(rng.normal(5.6, 0.6, 40))), 50, 5000)]).astype(np.uint64)
    batch = engine.batch_synth_aa(31, nres)
    buf, off = aa_oracle_batch(oracle, 31, nres)
    got = engine.sketch_setsketch(batch, 12, kb.KMERAA64, kb.HASH_MASKED_VALUE, SMALL, np.uint16)
    want = oracle.sketch_setsketch_batch(buf, off, nres, 12, kb.KMERAA64, kb.HASH_MASKED_VALUE, SMALL, np.uint16)
    assert np.array_equal(got, want)
    whole = engine.sketch_setsketch(batch, 12, kb.KMERAA64, kb.HASH_MASKED_VALUE, SMALL, np.uint16, whole=True)
    assert np.array_equal(whole, want.max(axis=0))


def test_setsketch_bad_arguments(engine):
    b = engine.batch_synth(1, np.array([1000], dtype=np.uint64))
    with pytest.raises(kb.KmuInvalid):
        engine.sketch_setsketch(b, 21, kb.KMER64, params=(1.0, 4096, 20.0, 65534))
    with pytest.raises(kb.KmuInvalid):
        engine.sketch_setsketch(b, 21, kb.KMER64, params=(1.001, 4096, 20.0, 70000), dtype=np.uint16)
    with pytest.raises(kb.KmuInvalid):
        engine.sketch_setsketch(b, 21, kb.KMER64, params=(1.001, 1 << 20, 20.0, 65534))
