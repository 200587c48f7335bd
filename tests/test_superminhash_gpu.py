"""GPU parity of SuperMinHash against the oracle (bit-exact f32 / f64 signatures) and the reference's
statistical tests (seqsketchjaccard.rs:947-1005) on the GPU output."""
import numpy as np
import pytest

import kmerutils_b200 as kb
from test_pmh3a_gpu import S80, oracle_batch

pytestmark = pytest.mark.gpu


def check_batch(engine, oracle, seed, nb, k, ktype, kind, m, hasher, dtype):
    nb = np.asarray(nb, dtype=np.uint64)
    batch = engine.batch_synth(seed, nb)
    packed, off = oracle_batch(oracle, seed, nb)
    got = engine.sketch_superminhash(batch, k, ktype, kind, m, hasher, dtype)
    want = oracle.sketch_superminhash_batch(packed, off, nb, k, ktype, kind, m, hasher, dtype)
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert len(bad) == 0, f"sequences {bad[:10]} (lengths {nb[bad[:10]]}) differ"
    return got


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("hasher", [kb.HASHER_NOHASH, kb.HASHER_FNV])
def test_superminhash_reads_k8(engine, oracle, dtype, hasher):
    rng = np.random.default_rng(11)
    # every regime: no k-mer, a handful (exact path), a_spec > 0, a_spec == 0, multi-warp teams
    nb = np.concatenate([[1, 7, 8, 9, 20, 50, 100, 150, 207, 300, 500, 800, 1200, 2000, 3000, 5000, 9000, 20000, 70000],
                         rng.integers(8, 4000, 60)])
    check_batch(engine, oracle, 21, nb, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 200, hasher, dtype)


@pytest.mark.parametrize("k,ktype,kind,m", [(16, kb.KMER16B32, kb.HASH_CANON_INVHASH, 1000),
                                            (21, kb.KMER64, kb.HASH_CANON_INVHASH, 128),
                                            (31, kb.KMER64, kb.HASH_IDENTITY_RAW, 50),
                                            (11, kb.KMER32, kb.HASH_MASKED_VALUE, 7),
                                            (5, kb.KMER32, kb.HASH_INVHASH, 1)])
def test_superminhash_types(engine, oracle, k, ktype, kind, m):
    rng = np.random.default_rng(k)
    nb = np.concatenate([[k - 1, k, k + 1, 40000], rng.integers(k, 6000, 40)])
    check_batch(engine, oracle, 30 + k, nb, k, ktype, kind, m, kb.HASHER_NOHASH, np.float64)


def test_superminhash_genome_sized(engine, oracle):
    # gsearch shape in small: k = 16 Kmer16b32bit, m = 12000 f64, one team of 32 warps per sequence
    nb = np.array([1_200_000, 400_000, 150_000, 3000], dtype=np.uint64)
    got = check_batch(engine, oracle, 41, nb, 16, kb.KMER16B32, kb.HASH_CANON_INVHASH, 12000, kb.HASHER_NOHASH, np.float64)
    assert (got[0] < 1.0).all()  # many more k-mers than slots: every slot was hit at j = 0


def test_superminhash_reference_inequalities(engine):
    # seqsketchjaccard.rs:947-1005: k = 16 Kmer16b32bit, f64 signatures; jaccard = fraction of equal slots
    def sig(seq, kind):
        b, _ = engine.batch_from_ascii([seq])
        return engine.sketch_superminhash(b, 16, kb.KMER16B32, kind, 50, kb.HASHER_FNV, np.float64)[0]
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    rc = S80.translate(comp)[::-1]
    a, half, arc = sig(S80, kb.HASH_CANON_INVHASH), sig(S80[:40], kb.HASH_CANON_INVHASH), sig(rc, kb.HASH_CANON_INVHASH)
    assert np.mean(a == half) >= 0.75 * (40 - 16) / (80 - 16)  # :992-993
    assert np.mean(a == arc) >= 1.0  # canonical hash: same k-mer set, same signature
    ia, irc = sig(S80, kb.HASH_IDENTITY_RAW), sig(rc, kb.HASH_IDENTITY_RAW)
    assert np.mean(ia == irc) <= 0.1  # :1003-1004


def test_superminhash_merge_is_min(engine, oracle):
    # SuperHashSketch::sketch_compressedkmer_seqs (setsketchert.rs:299-335): one sketch over several
    # sequences == element-wise minimum of the per-sequence sketches
    nb = np.array([5000, 1200, 300, 40, 9000], dtype=np.uint64)
    batch = engine.batch_synth(9, nb)
    packed, off = oracle_batch(oracle, 9, nb)
    per_seq = engine.sketch_superminhash(batch, 12, kb.KMER32, kb.HASH_CANON_INVHASH, 300)
    whole = oracle.sketch_superminhash_seqs(packed, off, nb, 12, kb.KMER32, kb.HASH_CANON_INVHASH, 300)
    assert np.array_equal(per_seq.min(axis=0), whole)


def test_superminhash_value_cut_paths(engine, oracle):
    """Long sequences take the value cut (an item whose first value is not below 4 m ln(1e4 m) / nk is dropped after half
    a seeding): random ones pass its verification, tandem repeats and half-repetitive reads (far fewer distinct k-mers than
    the cut assumes) fail it and are redone on the exact path; f32 and f64, both key hashers."""
    rng = np.random.default_rng(19)
    unit = bytes(rng.choice(list(b"ACGT"), 53).astype(np.uint8))
    rnd = bytes(rng.choice(list(b"ACGT"), 120000).astype(np.uint8))
    seqs = [rnd, unit * 2000, rnd[:40000] + unit * 1000, rnd[:20000], (unit * 40 + rnd[:500]) * 30]
    batch, _ = engine.batch_from_ascii(seqs)
    packed, off, nb = batch.download()
    packed = np.concatenate([packed, np.zeros(64, np.uint8)])
    for k, ktype, m, dtype, hasher in [(21, kb.KMER64, 64, np.float64, kb.HASHER_NOHASH), (16, kb.KMER16B32, 100, np.float32, kb.HASHER_FNV),
                                       (12, kb.KMER32, 37, np.float64, kb.HASHER_FNV)]:
        got = engine.sketch_superminhash(batch, k, ktype, kb.HASH_CANON_INVHASH, m, hasher, dtype)
        want = oracle.sketch_superminhash_batch(packed, off, nb, k, ktype, kb.HASH_CANON_INVHASH, m, hasher, dtype)
        bad = np.nonzero((got.view(np.uint8).reshape(len(seqs), -1) != want.view(np.uint8).reshape(len(seqs), -1)).any(axis=1))[0]
        assert len(bad) == 0, f"k={k}: sequences {bad} differ"


def test_superminhash_bad_arguments(engine):
    b = engine.batch_synth(1, np.array([100], dtype=np.uint64))
    with pytest.raises(kb.KmuInvalid):
        engine.sketch_superminhash(b, 15, kb.KMER32)
    with pytest.raises(kb.KmuInvalid):
        engine.sketch_superminhash(b, 8, kb.KMER32, m=0)
    with pytest.raises(kb.KmuInvalid):
        engine.sketch_superminhash(b, 8, kb.KMER32, m=100000)  # 800 kB of slots: does not fit one SM


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("nb", [[600000, 250000, 10, 3000, 77], [500, 300, 20], [40], [5, 3]])
def test_superminhash_whole_file(engine, oracle, nb, dtype):
    # SuperHashSketch::sketch_compressedkmer_seqs (setsketchert.rs:299-335): one sketch over all contigs;
    # large inputs take the whole-batch kernel, small ones the per-sequence path merged by minimum
    nb = np.array(nb, dtype=np.uint64)
    batch = engine.batch_synth(61, nb)
    packed, off = oracle_batch(oracle, 61, nb)
    got = engine.sketch_superminhash_whole(batch, 12, kb.KMER32, kb.HASH_CANON_INVHASH, 300, kb.HASHER_NOHASH, dtype)
    want = oracle.sketch_superminhash_seqs(packed, off, nb, 12, kb.KMER32, kb.HASH_CANON_INVHASH, 300, 0, dtype)
    assert np.array_equal(got, want)


def test_superminhash_whole_genome_m12000(engine, oracle):
    nb = np.array([900000, 400000, 200000], dtype=np.uint64)  # a 1.5 Mb "genome" in three contigs
    batch = engine.batch_synth(62, nb)
    packed, off = oracle_batch(oracle, 62, nb)
    got = engine.sketch_superminhash_whole(batch, 16, kb.KMER16B32, kb.HASH_CANON_INVHASH, 12000)
    want = oracle.sketch_superminhash_seqs(packed, off, nb, 16, kb.KMER16B32, kb.HASH_CANON_INVHASH, 12000)
    assert np.array_equal(got, want)
