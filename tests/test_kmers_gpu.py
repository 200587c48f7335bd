"""GPU parity for the materialising kernels: 2-bit packing, k-mer generation + hash closures, ntHash."""
import numpy as np
import pytest

import kmerutils_b200 as kb
from test_pmh3a_gpu import S80, oracle_batch

pytestmark = pytest.mark.gpu

TYPES = [(3, kb.KMER32), (8, kb.KMER32), (11, kb.KMER32), (14, kb.KMER32), (16, kb.KMER16B32), (5, kb.KMER64),
         (16, kb.KMER64), (21, kb.KMER64), (31, kb.KMER64), (32, kb.KMER64)]


def oracle_kmers(oracle, packed, off, nbases, k, ktype, kind):
    out = []
    for i, L in enumerate(nbases):
        L = int(L)
        p = packed[int(off[i]): int(off[i]) + (L + 3) // 4 + 8]
        w = oracle.generate_kmers(p, L, k, ktype) if L >= k else np.zeros(0, np.uint64)
        out.append(oracle.apply_hash(w, k, ktype, kind) if len(w) else w)
    return np.concatenate(out) if out else np.zeros(0, np.uint64)


@pytest.mark.parametrize("k,ktype", TYPES)
def test_generate_kmers_identity(engine, oracle, k, ktype):
    rng = np.random.default_rng(k)
    nb = np.concatenate([np.arange(0, 70), rng.integers(70, 1500, 40)]).astype(np.uint64)
    nb[0] = 1  # empty sequences are rejected by the reference (set_range fails)
    batch = engine.batch_synth(100 + k, nb)
    packed, off = oracle_batch(oracle, 100 + k, nb)
    got, koff = engine.generate_kmers(batch, k, ktype, kb.HASH_IDENTITY_RAW)
    want = oracle_kmers(oracle, packed, off, nb, k, ktype, kb.HASH_IDENTITY_RAW)
    assert len(got) == len(want) == int(koff[-1])
    assert np.array_equal(got.astype(np.uint64), want)
    exp_off = np.concatenate([[0], np.cumsum(np.maximum(nb.astype(np.int64) - k + 1, 0))])
    assert np.array_equal(koff.astype(np.int64), exp_off)


@pytest.mark.parametrize("kind", [kb.HASH_MASKED_VALUE, kb.HASH_CANON_INVHASH, kb.HASH_CANON_RAW, kb.HASH_INVHASH])
@pytest.mark.parametrize("k,ktype", [(8, kb.KMER32), (16, kb.KMER16B32), (31, kb.KMER64)])
def test_generate_kmers_hash_kinds(engine, oracle, kind, k, ktype):
    nb = np.array([80, 33, 1000, 15, 31, 32, 257], dtype=np.uint64)
    batch = engine.batch_synth(7, nb)
    packed, off = oracle_batch(oracle, 7, nb)
    got, _ = engine.generate_kmers(batch, k, ktype, kind)
    want = oracle_kmers(oracle, packed, off, nb, k, ktype, kind)
    assert np.array_equal(got.astype(np.uint64), want)


RUN_CASES = [(3, kb.KMER32), (7, kb.KMER32), (8, kb.KMER32), (9, kb.KMER32), (14, kb.KMER32), (16, kb.KMER16B32),
             (5, kb.KMER64), (21, kb.KMER64), (31, kb.KMER64), (32, kb.KMER64)]


@pytest.mark.parametrize("kind", [kb.HASH_IDENTITY_RAW, kb.HASH_MASKED_VALUE, kb.HASH_CANON_INVHASH, kb.HASH_CANON_RAW,
                                  kb.HASH_INVHASH])
@pytest.mark.parametrize("k,ktype", RUN_CASES)
def test_generate_kmers_run_form(engine, oracle, kind, k, ktype):
    """The run form (8 u32 / 4 u64 consecutive k-mers per lane, one 256-bit store): sequences that cross several 2 KB
    groups, segments that start at every output alignment, ragged heads and tails, sequences shorter than a run."""
    nb = np.array([9000, 5, 20011, 8191, 8193, 33, 33002, 1, 70000, 8200, 17, 4099], dtype=np.uint64)
    batch = engine.batch_synth(31 + k, nb)
    packed, off = oracle_batch(oracle, 31 + k, nb)
    got, _ = engine.generate_kmers(batch, k, ktype, kind)
    want = oracle_kmers(oracle, packed, off, nb, k, ktype, kind)
    assert len(got) == len(want)
    bad = np.flatnonzero(got.astype(np.uint64) != want)
    assert len(bad) == 0, f"first mismatch at k-mer {bad[0]}: {int(got[bad[0]]):#x} != {int(want[bad[0]]):#x}"


def test_reference_kmer_vectors(engine, oracle):
    # kmergenerator.rs:596-699 : 16-mers of the 80-base test string; first three words are in the source comments
    batch, bad = engine.batch_from_ascii([S80])
    got, _ = engine.generate_kmers(batch, 16, kb.KMER16B32)
    assert [hex(x) for x in got[:3]] == ["0xd02a013d", "0x40a804f4", "0x2a013d0"]
    assert len(got) == 65
    # every k-mer decompresses to seq[i..i+k]
    for i, w in enumerate(got):
        s = "".join("ACGT"[(int(w) >> (2 * (15 - j))) & 3] for j in range(16))
        assert s == S80[i:i + 16].decode()
    # Kmer32bit k=8: the word carries k in its top 4 bits (kmer32bit.rs:212-216)
    got8, _ = engine.generate_kmers(batch, 8, kb.KMER32)
    assert hex(got8[0]) == "0x8000d02a"
    canon, _ = engine.generate_kmers(batch, 8, kb.KMER32, kb.HASH_CANON_RAW)
    assert hex(canon[0]) == "0x800057f8"


def test_bad_kmer_sizes(engine):
    batch = engine.batch_synth(1, [100])
    for k, ktype in [(15, kb.KMER32), (0, kb.KMER32), (12, kb.KMER16B32), (33, kb.KMER64)]:
        with pytest.raises(kb.KmuInvalid):
            engine.generate_kmers(batch, k, ktype)
        with pytest.raises(kb.KmuInvalid):
            engine.sketch_pmh3a(batch, k, ktype, kb.HASH_CANON_INVHASH, 10)
    with pytest.raises(kb.KmuInvalid):
        engine.sketch_pmh3a(batch, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 1)  # ProbMinHash3a::new asserts m >= 2


@pytest.mark.parametrize("k", [1, 4, 8, 15, 16, 17, 24, 31, 32])
@pytest.mark.parametrize("n_multi", [1, 4])
def test_nthash(engine, oracle, k, n_multi):
    nb = np.array([80, 40, 513, 16, 31, 32, 33, 1200], dtype=np.uint64)
    batch = engine.batch_synth(11, nb)
    packed, off = oracle_batch(oracle, 11, nb)
    h, strand = engine.nthash_canonical(batch, k, n_multi)
    words = oracle_kmers(oracle, packed, off, nb, k, kb.KMER64, kb.HASH_IDENTITY_RAW)
    assert len(words) == len(h)
    for i in np.linspace(0, len(words) - 1, min(len(words), 400)).astype(int):
        f, r, c, s = oracle.nthash_canonical(words[i], k, kb.KMER64)
        assert int(h[i, 0]) == c and int(strand[i]) == s
        assert np.array_equal(h[i], oracle.nthash_mult(c, k, n_multi))
    # exhaustive on hash 0 (vectorised through the oracle is slow; sample above, full here for one sequence)
    first = int(np.maximum(nb[0] - k + 1, 0))
    for i in range(first):
        assert int(h[i, 0]) == oracle.nthash_canonical(words[i], k, kb.KMER64)[2]


@pytest.mark.parametrize("k", [1, 8, 15, 16, 17, 31, 32])
def test_nthash_run_form(engine, oracle, k):
    """n_multi = 1 takes the run form (per-lane stretches, 256-bit stores): long and ragged sequences, sampled against the
    oracle, and every value against the tile kernel (n_multi = 2, same first hash)."""
    nb = np.array([9000, 5, 20011, 8191, 8193, 33, 33002, 1, 70000, 8200, 17, 4099, 64, 95], dtype=np.uint64)
    batch = engine.batch_synth(77 + k, nb)
    packed, off = oracle_batch(oracle, 77 + k, nb)
    h1, s1 = engine.nthash_canonical(batch, k, 1)
    h2, s2 = engine.nthash_canonical(batch, k, 2)
    assert np.array_equal(h1[:, 0], h2[:, 0]) and np.array_equal(s1, s2)
    h0, _ = engine.nthash_canonical(batch, k, 1, want_strand=False)
    assert np.array_equal(h0, h1)
    words = oracle_kmers(oracle, packed, off, nb, k, kb.KMER64, kb.HASH_IDENTITY_RAW)
    assert len(words) == len(h1)
    rng = np.random.default_rng(k)
    for i in np.concatenate([rng.integers(0, len(words), 600), np.arange(40), np.arange(len(words) - 40, len(words))]):
        f, r, c, s = oracle.nthash_canonical(words[i], k, kb.KMER64)
        assert int(h1[i, 0]) == c and int(s1[i]) == s


def test_nthash_survey_vectors(engine):
    # SURVEY.md Appendix C (derived from kmer.rs:74-94): k=16 windows 0..2 of the 80-base test string
    batch, _ = engine.batch_from_ascii([S80])
    h, strand = engine.nthash_canonical(batch, 16, 4)
    assert hex(h[0, 0]) == "0x684a2ec1114d51c5" and strand[0] == 1
    assert hex(h[1, 0]) == "0xe9a4f48606ebe72" and strand[1] == 1
    assert hex(h[2, 0]) == "0x76f05375b83658db" and strand[2] == 0
    assert [hex(x) for x in h[0]] == ["0x684a2ec1114d51c5", "0x9f8aa4cb1f1ddcf8", "0x7d4d399050cea95", "0x701f02551303a00d"]
    h8, s8 = engine.nthash_canonical(batch, 8, 1)
    assert [hex(x) for x in h8[:3, 0]] == ["0x935533199c1dfb81", "0x4f6868cb4fb9a55e", "0x319aaf47aa9e02f9"]
    assert list(s8[:3]) == [0, 0, 0]


def test_pack_ascii(engine, oracle):
    rng = np.random.default_rng(3)
    seqs = [b"ACGTC", b"TCNGCAGTTGGATCCC", b"acgtACGTnnNN", S80, b"A", b"", b"TTTT",
            bytes(rng.choice(list(b"ACGTacgtNRY-"), 5000).astype(np.uint8))]
    # strict: any invalid character fails like Alphabet2b::encode (alphabet.rs:125)
    with pytest.raises(kb.KmuInvalid):
        engine.batch_from_ascii(seqs)
    good = [s for s in seqs if oracle.count_non_acgt(s) == 0]
    batch, bad = engine.batch_from_ascii(good)
    assert bad.sum() == 0
    packed, off, nb = batch.download()
    for i, s in enumerate(good):
        assert nb[i] == len(s)
        want = oracle.pack_2bit(s)
        assert np.array_equal(packed[int(off[i]): int(off[i]) + len(want)], want)
    assert packed[int(off[0]): int(off[0]) + 2].tobytes() == b"\x1b\x40"  # sequence.rs:844-848
    # drop_invalid: Sequence::encode_and_add (sequence.rs:388-451), "TCNGCAGTTGGATCCC" -> "TCGCAGTTGGATCCC"
    batch, bad = engine.batch_from_ascii(seqs, drop_invalid=True)
    packed, off, nb = batch.download()
    for i, s in enumerate(seqs):
        want, kept = oracle.encode_and_add(s)
        assert bad[i] == oracle.count_non_acgt(s)
        assert nb[i] == kept
        assert np.array_equal(packed[int(off[i]): int(off[i]) + len(want)], want)
    assert oracle.unpack_2bit(packed[int(off[1]):], int(nb[1])) == b"TCGCAGTTGGATCCC"


def test_batch_from_packed_layouts(engine, oracle):
    # unaligned offsets are re-laid out; aligned ones are copied as is; both give the same batch
    nb = np.array([5, 80, 1000, 3, 64, 17], dtype=np.uint64)
    parts = [oracle.synth_packed(9, 1000 * i, int(L)) for i, L in enumerate(nb)]
    tight = np.concatenate(parts)
    tight_off = np.concatenate([[0], np.cumsum([len(p) for p in parts])[:-1]]).astype(np.uint64)
    b1 = engine.batch_from_packed(tight, tight_off, nb)
    b2 = engine.batch_from_sequences(parts, nb)
    p1, o1, n1 = b1.download()
    p2, o2, n2 = b2.download()
    assert np.array_equal(p1, p2) and np.array_equal(o1, o2) and np.array_equal(n1, nb)
    b3 = engine.batch_from_packed(p1, o1, nb)
    p3, _, _ = b3.download()
    assert np.array_equal(p3, p1)
    k1, _ = engine.generate_kmers(b1, 5, kb.KMER32)
    k3, _ = engine.generate_kmers(b3, 5, kb.KMER32)
    assert np.array_equal(k1, k3)
