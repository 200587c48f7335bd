"""GPU parity for the amino-acid path (src/aautils/kmeraa.rs, src/aautils/setsketchert.rs): 5-bit k-mers,
ProbMinHash3a and SuperMinHash on proteins, against the oracle and the reference's own tests."""
import numpy as np
import pytest

import kmerutils_b200 as kb
from test_oracle_kat import AA_STR1, AA_STR2, GOLD, aa_to_str

pytestmark = pytest.mark.gpu


def aa_oracle_batch(oracle, seed, nres):
    """the synthetic proteins of kmu_seqbatch_synth_aa as one ASCII buffer + byte offsets"""
    nres = np.asarray(nres, dtype=np.uint64)
    off = np.zeros(len(nres), dtype=np.uint64)
    if len(nres) > 1:
        off[1:] = np.cumsum(nres)[:-1]
    first = off.copy()
    buf = b"".join(oracle.synth_aa(seed, int(f), int(n)) for f, n in zip(first, nres))
    return np.frombuffer(buf + b"\0" * 16, dtype=np.uint8), off


def oracle_kmers(oracle, buf, off, nres, k, ktype, kind):
    out = []
    for o, n in zip(off, nres):
        n = int(n)
        if n >= k:
            w = oracle.generate_kmers(buf[int(o): int(o) + n], n, k, ktype)
            out.append(oracle.apply_hash(w, k, ktype, kind))
    return np.concatenate(out) if out else np.zeros(0, np.uint64)


@pytest.mark.parametrize("k,ktype", [(1, kb.KMERAA32), (4, kb.KMERAA32), (6, kb.KMERAA32), (5, kb.KMERAA64),
                                     (8, kb.KMERAA64), (12, kb.KMERAA64)])
@pytest.mark.parametrize("kind", [kb.HASH_IDENTITY_RAW, kb.HASH_MASKED_VALUE, kb.HASH_INVHASH])
def test_aa_generate_kmers(engine, oracle, k, ktype, kind):
    rng = np.random.default_rng(k)
    nres = np.concatenate([np.arange(1, 40), rng.integers(40, 3000, 30)]).astype(np.uint64)
    batch = engine.batch_synth_aa(50 + k, nres)
    buf, off = aa_oracle_batch(oracle, 50 + k, nres)
    got, koff = engine.generate_kmers(batch, k, ktype, kind)
    want = oracle_kmers(oracle, buf, off, nres, k, ktype, kind)
    assert len(got) == len(want) == int(koff[-1])
    assert np.array_equal(got.astype(np.uint64), want)


def test_aa_reference_vectors(engine):
    prot = GOLD["aa"]["protein"].encode()
    batch, bad = engine.batch_from_aa([prot, prot[:32]])
    assert not bad.any()
    got, koff = engine.generate_kmers(batch, 4, kb.KMERAA32)
    # kmeraa.rs:920-958: range 3..10 gives QIEL IELI ELIK LIKL = k-mers 3..6 of the sequence
    assert [aa_to_str(v, 4) for v in got[3:7]] == GOLD["aa"]["range_3_10_4mers"]
    got8, koff8 = engine.generate_kmers(batch, 8, kb.KMERAA64)
    assert aa_to_str(got8[-1], 8) == GOLD["aa"]["last_8mer"]  # kmeraa.rs:998-1021
    assert int(koff8[1]) == len(prot) - 7 and len(got8) - int(koff8[1]) == 25


def test_aa_ingest_filtering(engine, oracle):
    seqs = [b"MTEQIELIKLYS", b"MTxEQ*IBLKZ", b"", b"ACDEFGHIKLMNPQRSTVWY"]
    with pytest.raises(kb.KmuInvalid):  # Alphabet::encode panics on x, *, B, Z (kmeraa.rs:106)
        engine.batch_from_aa(seqs)
    batch, bad = engine.batch_from_aa(seqs, drop_invalid=True)  # SequenceAA::new_filtered (kmeraa.rs:447-456)
    assert bad.tolist() == [0, 4, 0, 0]
    codes, off, n = batch.download()
    assert n.tolist() == [12, 7, 0, 20]
    letters = "?ACDEFGHIKLMNP?QRSTVWY"
    for i, s_ in enumerate(seqs):
        kept = oracle.aa_filter(s_)
        assert "".join(letters[c] for c in codes[int(off[i]): int(off[i]) + int(n[i])]) == kept.decode()


def test_aa_rejects_canonical_and_dna_mixups(engine):
    aa = engine.batch_synth_aa(1, np.array([100], dtype=np.uint64))
    dna = engine.batch_synth(1, np.array([100], dtype=np.uint64))
    with pytest.raises(kb.KmuInvalid):  # no reverse complement for amino acids (kmeraa.rs:185-187)
        engine.generate_kmers(aa, 5, kb.KMERAA64, kb.HASH_CANON_INVHASH)
    with pytest.raises(kb.KmuInvalid):
        engine.generate_kmers(aa, 8, kb.KMER32)
    with pytest.raises(kb.KmuInvalid):
        engine.generate_kmers(dna, 5, kb.KMERAA64)
    with pytest.raises(kb.KmuInvalid):
        engine.generate_kmers(aa, 7, kb.KMERAA32)
    with pytest.raises(kb.KmuInvalid):
        engine.sketch_pmh3a(aa, 13, kb.KMERAA64, kb.HASH_MASKED_VALUE, 100)


@pytest.mark.parametrize("k,ktype,kind,m", [(5, kb.KMERAA64, kb.HASH_MASKED_VALUE, 400),
                                            (12, kb.KMERAA64, kb.HASH_MASKED_VALUE, 400),
                                            (6, kb.KMERAA32, kb.HASH_INVHASH, 64),
                                            (3, kb.KMERAA32, kb.HASH_IDENTITY_RAW, 200)])
def test_aa_pmh3a_parity(engine, oracle, k, ktype, kind, m):
    rng = np.random.default_rng(100 + k)
    # proteome-like lengths plus the edge cases
    nres = np.concatenate([[1, k - 1 if k > 1 else 1, k, k + 1, 50, 5000, 40000],
                           np.clip(np.rint(np.exp(rng.normal(5.6, 0.6, 80))), 50, 5000)]).astype(np.uint64)
    batch = engine.batch_synth_aa(60 + k, nres)
    buf, off = aa_oracle_batch(oracle, 60 + k, nres)
    got = engine.sketch_pmh3a(batch, k, ktype, kind, m)
    want = oracle.sketch_pmh3a_batch(buf, off, nres, k, ktype, kind, m)
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert len(bad) == 0, f"proteins {bad[:10]} (lengths {nres[bad[:10]]}) differ"


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_aa_superminhash_parity(engine, oracle, dtype):
    rng = np.random.default_rng(7)
    nres = np.concatenate([[1, 4, 5, 6, 30, 20000], rng.integers(5, 3000, 50)]).astype(np.uint64)
    batch = engine.batch_synth_aa(70, nres)
    buf, off = aa_oracle_batch(oracle, 70, nres)
    got = engine.sketch_superminhash(batch, 5, kb.KMERAA64, kb.HASH_MASKED_VALUE, 300, kb.HASHER_NOHASH, dtype)
    want = oracle.sketch_superminhash_batch(buf, off, nres, 5, kb.KMERAA64, kb.HASH_MASKED_VALUE, 300, 0, dtype)
    assert np.array_equal(got, want)


def test_aa_reference_sketch_inequalities(engine):
    # aautils/setsketchert.rs:1217-1391: |J - 0.5| < 0.1 for ProbMinHash3a (64 and 32 bit k-mers) and SuperMinHash
    batch, _ = engine.batch_from_aa([AA_STR1, AA_STR2])
    s64 = engine.sketch_pmh3a(batch, 5, kb.KMERAA64, kb.HASH_MASKED_VALUE, 400)
    s32 = engine.sketch_pmh3a(batch, 5, kb.KMERAA32, kb.HASH_MASKED_VALUE, 800)
    smh = engine.sketch_superminhash(batch, 5, kb.KMERAA64, kb.HASH_MASKED_VALUE, 800)
    for s_ in (s64, s32, smh):
        assert abs(float(np.mean(s_[0] == s_[1])) - 0.5) < 0.1
