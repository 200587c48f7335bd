"""GPU parity: ProbMinHash3a signatures from the CUDA path == the CPU oracle, bit for bit."""
import numpy as np
import pytest

import kmerutils_b200 as kb

pytestmark = pytest.mark.gpu

S80 = b"TCAAAGGGAAACATTCAAAATCAGTATGCGCCCGTTCAGTTACGTATTGCTCTCGCTAATGAGATGGGCTGGGTACAGAG"


def layout(nbases):
    nb = np.asarray(nbases, dtype=np.uint64)
    sizes = ((nb + 3) // 4 + 15) // 16 * 16
    off = np.zeros(len(nb), dtype=np.uint64)
    off[1:] = np.cumsum(sizes)[:-1]
    return off, int(sizes.sum())


def oracle_batch(oracle, seed, nbases):
    """Same synthetic reads as kmu_seqbatch_synth: one SplitMix64 stream, reads back to back."""
    off, total = layout(nbases)
    packed = np.zeros(total + 64, dtype=np.uint8)
    first = 0
    for i, L in enumerate(nbases):
        L = int(L)
        packed[int(off[i]): int(off[i]) + (L + 3) // 4] = oracle.synth_packed(seed, first, L)
        first += L
    return packed, off


def check_config(engine, oracle, seed, nbases, k, ktype, kind, m):
    nbases = np.asarray(nbases, dtype=np.uint64)
    batch = engine.batch_synth(seed, nbases)
    packed, off = oracle_batch(oracle, seed, nbases)
    dl_packed, dl_off, dl_nb = batch.download()
    assert np.array_equal(dl_off, off)
    assert np.array_equal(dl_packed, packed[: len(dl_packed)])
    got = engine.sketch_pmh3a(batch, k, ktype, kind, m)
    want = oracle.sketch_pmh3a_batch(packed, off, nbases, k, ktype, kind, m)
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert len(bad) == 0, f"{len(bad)} of {len(nbases)} signatures differ, first: seq {bad[:5]} len {nbases[bad[:5]]}"
    batch.destroy()


def test_c1_config(engine, oracle):
    # BASELINE config 1: 1000 reads x 1000 b, k=8 Kmer32bit, canonical + int32_hash, m=200
    check_config(engine, oracle, 1, [1000] * 1000, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 200)


def test_ragged_lengths_k8(engine, oracle):
    rng = np.random.default_rng(7)
    nb = np.concatenate([
        np.arange(1, 40), rng.integers(40, 300, 200), rng.integers(300, 5000, 200), rng.integers(5000, 40000, 40),
        [70000, 131072, 250000]])
    check_config(engine, oracle, 2, nb, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 200)


@pytest.mark.parametrize("k,ktype", [(3, kb.KMER32), (5, kb.KMER32), (11, kb.KMER32), (14, kb.KMER32),
                                     (16, kb.KMER16B32), (12, kb.KMER64), (21, kb.KMER64), (31, kb.KMER64),
                                     (32, kb.KMER64)])
def test_kmer_types(engine, oracle, k, ktype):
    rng = np.random.default_rng(k)
    nb = np.concatenate([rng.integers(1, 200, 60), rng.integers(200, 3000, 60), rng.integers(3000, 30000, 12)])
    check_config(engine, oracle, 10 + k, nb, k, ktype, kb.HASH_CANON_INVHASH, 64)


@pytest.mark.parametrize("kind", [kb.HASH_IDENTITY_RAW, kb.HASH_MASKED_VALUE, kb.HASH_CANON_RAW, kb.HASH_INVHASH])
@pytest.mark.parametrize("k,ktype", [(8, kb.KMER32), (16, kb.KMER16B32), (21, kb.KMER64)])
def test_hash_kinds(engine, oracle, kind, k, ktype):
    rng = np.random.default_rng(kind * 100 + k)
    nb = rng.integers(20, 4000, 80)
    check_config(engine, oracle, 50 + kind, nb, k, ktype, kind, 50)


@pytest.mark.parametrize("m", [2, 3, 17, 255, 256, 257, 1000, 4000, 12000])
def test_sketch_sizes(engine, oracle, m):
    rng = np.random.default_rng(m)
    nb = rng.integers(100, 20000, 24)
    check_config(engine, oracle, 90, nb, 16, kb.KMER16B32, kb.HASH_CANON_INVHASH, m)


def test_low_complexity_and_u16_wrap(engine, oracle):
    # homopolymers / short tandem repeats: one k-mer with a huge multiplicity, including
    # > 65535 occurrences (wraps the u16 histogram counters -> table redo path)
    seqs = [b"A" * 100000, b"AC" * 40000, b"ACGT" * 30, b"T" * 70000 + S80 * 20, S80, b"G" * 9]
    batch, bad = engine.batch_from_ascii(seqs)
    assert bad.sum() == 0
    packed, off, nb = batch.download()
    for k, ktype in [(8, kb.KMER32), (16, kb.KMER16B32)]:
        got = engine.sketch_pmh3a(batch, k, ktype, kb.HASH_CANON_INVHASH, 200)
        want = oracle.sketch_pmh3a_batch(np.concatenate([packed, np.zeros(64, np.uint8)]), off, nb, k, ktype,
                                         kb.HASH_CANON_INVHASH, 200)
        assert np.array_equal(got, want)


def test_reference_statistical_tests(engine, oracle):
    # seqsketchjaccard.rs:742-851 : k=5, m=4000 ; J(a, a[0..40]) >= 0.75 * (40-k)/(80-k) ; J(a, revcomp a) >= 1
    p80 = oracle.pack_2bit(S80)
    rc = oracle.seq_revcomp(p80, 80)
    batch = engine.batch_from_sequences([p80, oracle.pack_2bit(S80[:40]), rc], [80, 40, 80])
    sig = engine.sketch_pmh3a(batch, 5, kb.KMER32, kb.HASH_CANON_INVHASH, 4000)
    j_half = oracle.jaccard(sig[0], sig[1])
    assert j_half >= 0.75 * (40 - 5) / (80 - 5)
    assert oracle.jaccard(sig[0], sig[2]) >= 1.0
    # identity hash: reverse complement shares (almost) nothing (<= 0.1)
    sig_id = engine.sketch_pmh3a(batch, 5, kb.KMER32, kb.HASH_IDENTITY_RAW, 4000)
    assert oracle.jaccard(sig_id[0], sig_id[2]) <= 0.1


@pytest.mark.parametrize("chunk_bytes", [None, 4096, 100000])
@pytest.mark.parametrize("ragged", [False, True])
def test_host_pipeline(engine, oracle, chunk_bytes, ragged, monkeypatch):
    # kmu_sketch_pmh3a_host: host buffers in, host signatures out, chunked over three streams.  Same
    # signatures whatever the chunking, for a buffer in the batch layout and for a tightly packed one.
    rng = np.random.default_rng(17)
    nb = np.concatenate([[1, 7, 8, 30000], rng.integers(8, 6000, 300)]).astype(np.uint64)
    packed, off = oracle_batch(oracle, 33, nb)
    want = oracle.sketch_pmh3a_batch(packed, off, nb, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 200)
    if ragged:  # sequences back to back without the 16-byte alignment: re-laid out through the pinned staging buffer
        sizes = (nb + np.uint64(3)) // np.uint64(4)
        roff = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.uint64)
        tight = np.zeros(int(sizes.sum()) + 16, dtype=np.uint8)
        for o, r, s_ in zip(off, roff, sizes):
            tight[int(r): int(r + s_)] = packed[int(o): int(o + s_)]
        packed, off = tight, roff
    if chunk_bytes:
        monkeypatch.setenv("KMU_HOST_CHUNK_BYTES", str(chunk_bytes))
    out = np.zeros((len(nb), 200), dtype=np.uint32)
    engine.sketch_pmh3a_host(packed, off, nb, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 200, out)
    assert np.array_equal(out, want)
    t = engine.last_times()
    assert t["d2h_bytes"] == out.nbytes and t["h2d_bytes"] > 0 and t["host_ms"] > 0


@pytest.mark.parametrize("chunk_bytes", [None, 4096])
def test_host_pipeline_from_separate_sequences(engine, oracle, chunk_bytes, monkeypatch):
    # kmu_sketch_pmh3a_host_ptrs: every sequence its own host allocation (the `&[&Sequence]` of the Rust entry points,
    # setsketchert.rs:70-79), gathered into pinned staging memory chunk by chunk.  Same signatures.
    rng = np.random.default_rng(19)
    nb = np.concatenate([[1, 7, 8, 40000], rng.integers(8, 7000, 6000)]).astype(np.uint64)
    packed, off = oracle_batch(oracle, 35, nb)
    want = oracle.sketch_pmh3a_batch(packed, off, nb, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 200)
    seqs = [np.array(packed[int(o): int(o) + (int(n) + 3) // 4], dtype=np.uint8, copy=True) for o, n in zip(off, nb)]
    addrs = np.array([a.ctypes.data for a in seqs], dtype=np.uint64)
    if chunk_bytes:
        monkeypatch.setenv("KMU_HOST_CHUNK_BYTES", str(chunk_bytes))
    out = np.zeros((len(nb), 200), dtype=np.uint32)
    engine.sketch_pmh3a_host_ptrs(addrs, nb, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 200, out)
    assert np.array_equal(out, want)
    with pytest.raises(kb.KmuInvalid):
        bad = addrs.copy()
        bad[3] = 0
        engine.sketch_pmh3a_host_ptrs(bad, nb, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 200, out)


def test_host_pipeline_with_redo_sequences(engine, oracle, monkeypatch):
    # sequences whose u8 histogram counters wrap are redone by a second launch per chunk: with several chunks in
    # flight every chunk parity has its own counters / redo list
    rng = np.random.default_rng(23)
    seqs = []
    for i in range(60):
        if i % 7 == 3:
            seqs.append(b"AC" * int(rng.integers(400, 3000)))  # one k-mer seen > 255 times
        else:
            seqs.append(oracle.synth_ascii(55, int(rng.integers(0, 1 << 20)), int(rng.integers(50, 4000))))
    packed = [oracle.pack_2bit(s) for s in seqs]
    nb = np.array([len(s) for s in seqs], dtype=np.uint64)
    sizes = np.array([(len(p) + 15) // 16 * 16 for p in packed], dtype=np.uint64)
    off = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.uint64)
    buf = np.zeros(int(sizes.sum()) + 64, dtype=np.uint8)
    for o, p in zip(off, packed):
        buf[int(o): int(o) + len(p)] = p
    want = oracle.sketch_pmh3a_batch(buf, off, nb, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 200)
    monkeypatch.setenv("KMU_HOST_CHUNK_BYTES", "3000")
    out = np.zeros((len(nb), 200), dtype=np.uint32)
    for _ in range(2):
        out[:] = 0
        engine.sketch_pmh3a_host(buf, off, nb, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 200, out)
        assert np.array_equal(out, want)


@pytest.mark.parametrize("k,m,kind", [(8, 200, kb.HASH_CANON_INVHASH), (8, 512, kb.HASH_CANON_INVHASH),
                                      (8, 513, kb.HASH_CANON_INVHASH), (7, 100, kb.HASH_IDENTITY_RAW),
                                      (6, 31, kb.HASH_CANON_RAW), (4, 16, kb.HASH_INVHASH)])
def test_one_pass_kernel_long_reads(engine, oracle, k, m, kind):
    # sequences with >= 2048 k-mers over <= 4^8 keys take the one-pass kernel (kmu_pmh3a_direct.cu); lengths straddle
    # its lower bound, small k gives counts far above 255 (u8 wrap -> flagged -> general kernel)
    rng = np.random.default_rng(1000 + k)
    nb = np.concatenate([[2047 + k - 1, 2048 + k - 1, 2049 + k - 1], rng.integers(2000, 2200, 40),
                         rng.integers(2200, 12000, 150), rng.integers(12000, 90000, 30), [300000]])
    check_config(engine, oracle, 300 + k, nb, k, kb.KMER32, kind, m)


def test_one_pass_kernel_repeats(engine, oracle, monkeypatch):
    # tandem repeats inside random context: a few keys with counts 2..300 next to thousands of singletons
    rng = np.random.default_rng(5)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    seqs = []
    for i in range(60):
        parts = []
        for _ in range(int(rng.integers(2, 8))):
            parts.append(acgt[rng.integers(0, 4, int(rng.integers(500, 4000)))].tobytes())
            unit = acgt[rng.integers(0, 4, int(rng.integers(1, 40)))].tobytes()
            parts.append(unit * int(rng.integers(2, 120)))
        seqs.append(b"".join(parts))
    batch, bad = engine.batch_from_ascii(seqs)
    assert bad.sum() == 0
    packed, off, nb = batch.download()
    want = oracle.sketch_pmh3a_batch(np.concatenate([packed, np.zeros(64, np.uint8)]), off, nb, 8, kb.KMER32,
                                     kb.HASH_CANON_INVHASH, 200)
    got = engine.sketch_pmh3a(batch, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 200)
    assert np.array_equal(got, want)
    monkeypatch.setenv("KMU_NO_DIRECT", "1")  # the general kernel alone gives the same
    got2 = engine.sketch_pmh3a(batch, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 200)
    assert np.array_equal(got2, want)
