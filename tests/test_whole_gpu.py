"""Whole-file ProbMinHash3a (sketch_compressedkmer_seqs), explicit weighted sets, slices and block sketches."""
import numpy as np
import pytest

import kmerutils_b200 as kb
from test_pmh3a_gpu import S80, oracle_batch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k,ktype,kind,m", [(16, kb.KMER16B32, kb.HASH_CANON_INVHASH, 12000),
                                            (8, kb.KMER32, kb.HASH_CANON_INVHASH, 200),
                                            (21, kb.KMER64, kb.HASH_CANON_INVHASH, 1000),
                                            (31, kb.KMER64, kb.HASH_IDENTITY_RAW, 64),
                                            (12, kb.KMER32, kb.HASH_MASKED_VALUE, 15000)])
def test_pmh3a_whole_file(engine, oracle, k, ktype, kind, m):
    # ProbHash3aSketch::sketch_compressedkmer_seqs (setsketchert.rs:160-202): contigs of one genome -> one signature
    rng = np.random.default_rng(k + m)
    nb = np.concatenate([[400000, 90000, 5, k], rng.integers(k, 20000, 30)]).astype(np.uint64)
    batch = engine.batch_synth(60 + k, nb)
    packed, off = oracle_batch(oracle, 60 + k, nb)
    got = engine.sketch_pmh3a_whole(batch, k, ktype, kind, m)
    want = oracle.sketch_pmh3a_seqs(packed, off, nb, k, ktype, kind, m)
    assert np.array_equal(got.astype(np.uint64), want)


def test_pmh3a_whole_small_inputs(engine, oracle):
    # fewer distinct k-mers than slots: the bound has to grow until every slot is filled; no k-mer at all: zeros
    for nb in ([30], [9, 8], [7, 3]):
        nb = np.array(nb, dtype=np.uint64)
        batch = engine.batch_synth(5, nb)
        packed, off = oracle_batch(oracle, 5, nb)
        got = engine.sketch_pmh3a_whole(batch, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 200)
        want = oracle.sketch_pmh3a_seqs(packed, off, nb, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 200)
        assert np.array_equal(got.astype(np.uint64), want)


@pytest.mark.parametrize("dtype", [np.uint32, np.uint64])
def test_pmh3a_weighted_set(engine, oracle, dtype):
    # ProbMinHash3a::hash_weigthed_hashmap on explicit (key, weight) pairs, f64 weights (seqblocksketch.rs:121-138)
    rng = np.random.default_rng(4)
    for n, m in [(1, 50), (7, 50), (300, 200), (50000, 200), (200000, 12000)]:
        keys = np.unique(rng.integers(0, np.iinfo(dtype).max, n, dtype=np.uint64).astype(dtype))
        w = rng.integers(1, 40, len(keys)).astype(np.float64)
        if n == 300:
            w = w + 0.5  # non-integer weights
        got = engine.pmh3a_weighted(keys, w, m)
        want = oracle.pmh3a_weighted(keys.astype(np.uint64), w, m, np.dtype(dtype).itemsize)
        assert np.array_equal(got.astype(np.uint64), want)
    with pytest.raises(kb.KmuInvalid):
        engine.pmh3a_weighted(np.array([1, 2], dtype=dtype), np.array([1.0, 0.0]), 10)


@pytest.mark.parametrize("m", [2, 3])
def test_pmh3a_tiny_sketch_stress(engine, oracle, m):
    """m = 2 and 3 (lambda = ln 2, ln 1.5): the sketch sizes that send the most draws through the rejection branch of
    ExpRestricted01, down to its last test `y c1 lambda <= expm1(lambda (1 - x))` -- evaluated with the SAME deterministic
    expm1 on the device and in the oracle (kmu_detmath.cuh == oracle/det_math.hpp).  1.2e7 weighted items."""
    rng = np.random.default_rng(100 + m)
    keys = np.unique(rng.integers(0, np.iinfo(np.uint64).max, 12_000_000, dtype=np.uint64))
    w = rng.integers(1, 6, len(keys)).astype(np.float64)
    got = engine.pmh3a_weighted(keys, w, m)
    want = oracle.pmh3a_weighted(keys, w, m, 8)
    assert np.array_equal(got.astype(np.uint64), want)
    # many small sets as well: every one of them decides its slots among a handful of items
    for rep in range(200):
        kk = np.unique(rng.integers(0, 1 << 32, rng.integers(1, 40), dtype=np.uint64).astype(np.uint32))
        ww = rng.integers(1, 4, len(kk)).astype(np.float64)
        assert np.array_equal(engine.pmh3a_weighted(kk, ww, m).astype(np.uint64), oracle.pmh3a_weighted(kk.astype(np.uint64), ww, m, 4))


def test_slices(engine, oracle):
    rng = np.random.default_rng(8)
    nb = np.array([1000, 37, 5000, 16, 64], dtype=np.uint64)
    batch = engine.batch_synth(12, nb)
    packed, off = oracle_batch(oracle, 12, nb)
    seqs = [oracle.unpack_2bit(packed[int(o): int(o) + (int(n) + 3) // 4], int(n)) for o, n in zip(off, nb)]
    idx = rng.integers(0, len(nb), 200).astype(np.uint64)
    begin = np.array([rng.integers(0, int(nb[i]) + 1) for i in idx], dtype=np.uint64)
    end = np.array([b + rng.integers(0, 300) for b in begin], dtype=np.uint64)  # may overrun: clamped
    sl = engine.batch_slices(batch, idx, begin, end)
    sp, so, sn = sl.download()
    for j in range(len(idx)):
        want = seqs[int(idx[j])][int(begin[j]): int(end[j])]
        assert int(sn[j]) == len(want)
        nbytes = (len(want) + 3) // 4
        got = sp[int(so[j]): int(so[j]) + nbytes]
        assert oracle.unpack_2bit(got, len(want)) == want
        if len(want) % 4:  # the tail is padded with 'A' (00) like Sequence::new (sequence.rs:66-71)
            assert got[-1] & ((1 << (8 - 2 * (len(want) % 4))) - 1) == 0
    # amino-acid batches slice too
    aa = engine.batch_synth_aa(3, np.array([500, 20], dtype=np.uint64))
    codes, aoff, _ = aa.download()
    sa = engine.batch_slices(aa, [0, 1, 0], [10, 0, 490], [30, 25, 600])
    c2, o2, n2 = sa.download()
    assert n2.tolist() == [20, 20, 10]
    assert np.array_equal(c2[int(o2[0]): int(o2[0]) + 20], codes[10:30])
    assert np.array_equal(c2[int(o2[2]): int(o2[2]) + 10], codes[490:500])


@pytest.mark.parametrize("block_size", [50, 64, 1000])
def test_block_sketch(engine, oracle, block_size):
    # BlockSeqSketcher::blocksketch_sequences (seqblocksketch.rs:97-167), k = 8 Kmer32bit, canonical + int32_hash
    nb = np.array([1000, 207, 64, 50, 5, 3333], dtype=np.uint64)
    batch = engine.batch_synth(19, nb)
    packed, off = oracle_batch(oracle, 19, nb)
    sig, numseq, numblock = engine.blocksketch(batch, 8, 40, block_size)
    row = 0
    for s, L in enumerate(nb):
        want = oracle.blocksketch_seq(packed[int(off[s]):], int(L), 8, 40, block_size)
        n = len(want)
        assert n == (int(L) + block_size - 1) // block_size
        assert np.array_equal(sig[row:row + n], want)
        assert (numseq[row:row + n] == s).all() and numblock[row:row + n].tolist() == list(range(n))
        row += n
    assert row == len(sig)


def test_block_sketch_reference_test(engine):
    # seqblocksketch.rs:458-496: a sequence compared with itself block by block has distance 1 (all slots equal)
    batch, _ = engine.batch_from_ascii([S80, S80])
    sig, numseq, numblock = engine.blocksketch(batch, 3, 10, 20)
    a, b = sig[numseq == 0], sig[numseq == 1]
    assert a.shape == b.shape == (4, 10) and np.array_equal(a, b)


def test_views_and_groups(engine, oracle):
    # genomes made of several contigs in one batch: one whole-file signature per genome through views
    rng = np.random.default_rng(31)
    groups = [3, 1, 5, 2]
    nb = np.concatenate([rng.integers(2000, 60000, g) for g in groups]).astype(np.uint64)
    batch = engine.batch_synth(44, nb)
    packed, off = oracle_batch(oracle, 44, nb)
    first = 0
    sm = engine.sketch_groups(batch, groups, "superminhash", 16, kb.KMER16B32, m=500)
    pm = engine.sketch_groups(batch, groups, "pmh3a", 16, kb.KMER16B32, m=500)
    hl = engine.sketch_groups(batch, groups, "setsketch", 16, kb.KMER16B32, params=(1.001, 256, 20.0, 65534))
    for gi, g in enumerate(groups):
        sl = slice(first, first + g)
        o, n = off[sl], nb[sl]
        assert np.array_equal(sm[gi], oracle.sketch_superminhash_seqs(packed, o, n, 16, kb.KMER16B32, kb.HASH_CANON_INVHASH, 500))
        assert np.array_equal(pm[gi].astype(np.uint64), oracle.sketch_pmh3a_seqs(packed, o, n, 16, kb.KMER16B32, kb.HASH_CANON_INVHASH, 500))
        assert np.array_equal(hl[gi], oracle.sketch_setsketch_seqs(packed, o, n, 16, kb.KMER16B32, kb.HASH_CANON_INVHASH,
                                                                    (1.001, 256, 20.0, 65534)))
        first += g
    # a view behaves like a batch for the per-sequence calls too
    view = engine.batch_view(batch, 4, 5)
    assert len(view) == 5 and view.total_bases == int(nb[4:9].sum())
    got = engine.sketch_pmh3a(view, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 64)
    want = oracle.sketch_pmh3a_batch(packed, off[4:9], nb[4:9], 8, kb.KMER32, kb.HASH_CANON_INVHASH, 64)
    assert np.array_equal(got, want)
    kmers, _ = engine.generate_kmers(view, 21, kb.KMER64, kb.HASH_CANON_RAW)
    assert len(kmers) == int(np.maximum(nb[4:9].astype(np.int64) - 20, 0).sum())
    with pytest.raises(kb.KmuInvalid):
        engine.batch_view(batch, 9, 5)


@pytest.mark.parametrize("k,ktype", [(16, kb.KMER16B32), (21, kb.KMER64), (8, kb.KMER32)])
def test_pmh3a_groups_one_call(engine, oracle, k, ktype):
    # kmu_sketch_pmh3a_groups == one kmu_sketch_pmh3a_whole per group; includes a group too short for a k-mer (zero
    # signature), a one-contig genome and a highly repetitive genome (few distinct k-mers: its item bound fails and the
    # group goes through the full procedure)
    rng = np.random.default_rng(77 + k)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    unit = acgt[rng.integers(0, 4, 300)].tobytes()
    seqs = [acgt[rng.integers(0, 4, int(n))].tobytes() for n in (5000, 800, 12000)]          # genome 0
    seqs += [b"ACG"]                                                                          # genome 1: no k-mer (k > 3)
    seqs += [acgt[rng.integers(0, 4, 40000)].tobytes()]                                       # genome 2
    seqs += [unit * 60, unit * 45]                                                            # genome 3: tandem repeats
    groups = [3, 1, 1, 2]
    batch, bad = engine.batch_from_ascii(seqs)
    assert bad.sum() == 0
    packed, off, nb = batch.download()
    packed = np.concatenate([packed, np.zeros(64, np.uint8)])
    got = engine.sketch_pmh3a_groups(batch, groups, k, ktype, kb.HASH_CANON_INVHASH, 300)
    first = 0
    for gi, g in enumerate(groups):
        sl = slice(first, first + g)
        if k > 3 and gi == 1:
            assert not got[gi].any()
        else:
            want = oracle.sketch_pmh3a_seqs(packed, off[sl], nb[sl], k, ktype, kb.HASH_CANON_INVHASH, 300)
            assert np.array_equal(got[gi].astype(np.uint64), want), f"group {gi}"
        first += g
    with pytest.raises(kb.KmuInvalid):
        engine.sketch_pmh3a_groups(batch, [3, 1], k, ktype, kb.HASH_CANON_INVHASH, 300)


def test_pmh3a_counter_slots_match_whole_file(engine, oracle):
    # the building block of the multi-GPU whole-file sketch: registers (h, key) of a counting table == the whole-file
    # signature once the bound covers every slot; a bound that is too small leaves slots at the f64 maximum
    nb = np.array([30000, 12000, 700], dtype=np.uint64)
    batch = engine.batch_synth(91, nb)
    k, ktype, m = 21, kb.KMER64, 300
    want = engine.sketch_pmh3a_whole(batch, k, ktype, kb.HASH_CANON_INVHASH, m)
    counter = engine.counter(k, ktype, int(nb.sum()), 32)
    counter.insert_seqs(batch, canonical=True)
    h, keys = engine.pmh3a_counter_slots(counter, kb.HASH_CANON_INVHASH, m, 1.0)  # every item: all its points below 1
    assert np.array_equal(keys, want.astype(np.uint64))
    hv = h.view(np.float64)
    assert (hv > 0).all() and hv.max() < 1.0
    h2, _ = engine.pmh3a_counter_slots(counter, kb.HASH_CANON_INVHASH, m, hv.min() * 0.5)
    assert (h2.view(np.float64) == np.finfo(np.float64).max).all()
    with pytest.raises(kb.KmuInvalid):
        engine.pmh3a_counter_slots(counter, kb.HASH_CANON_INVHASH, 1, 1.0)
    with pytest.raises(kb.KmuInvalid):  # a bound no sketch can need (it would emit 1e300 points per key)
        engine.pmh3a_counter_slots(counter, kb.HASH_CANON_INVHASH, m, 1e300)
    counter.destroy()


@pytest.mark.parametrize("k,ktype,m", [(16, kb.KMER16B32, 2000), (21, kb.KMER64, 1500)])
def test_pmh3a_groups_prefiltered_insertion(engine, oracle, k, ktype, m):
    """Genome-sized groups take the prefiltered insertion (only the k-mers that can place a point below the bound are counted,
    plus every k-mer seen twice): a random genome in three contigs, a genome with a long segment repeated many times (its
    repeated k-mers matter through their weights), a genome with a low-complexity stretch -- against the oracle's whole-file
    signatures, and against the unfiltered path."""
    import os
    rng = np.random.default_rng(500 + k)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)

    def rnd(n):
        return acgt[rng.integers(0, 4, int(n))].tobytes()

    seg = rnd(60000)
    seqs = [rnd(700000), rnd(500000), rnd(300000)]                      # genome 0: 1.5 Mb, three contigs
    seqs += [rnd(400000) + seg * 12 + rnd(300000) + seg * 7]            # genome 1: a 60 kb segment 19 times
    seqs += [rnd(900000) + b"ACACACAC" * 20000 + rnd(400000)]           # genome 2: 160 kb of a dinucleotide repeat
    groups = [3, 1, 1]
    batch, bad = engine.batch_from_ascii(seqs)
    assert bad.sum() == 0
    packed, off, nb = batch.download()
    packed = np.concatenate([packed, np.zeros(64, np.uint8)])
    got = engine.sketch_pmh3a_groups(batch, groups, k, ktype, kb.HASH_CANON_INVHASH, m)
    os.environ["KMU_GROUP_NO_PREFILTER"] = "1"
    try:
        plain = engine.sketch_pmh3a_groups(batch, groups, k, ktype, kb.HASH_CANON_INVHASH, m)
    finally:
        del os.environ["KMU_GROUP_NO_PREFILTER"]
    assert np.array_equal(got, plain)
    first = 0
    for gi, g in enumerate(groups):
        sl = slice(first, first + g)
        want = oracle.sketch_pmh3a_seqs(packed, off[sl], nb[sl], k, ktype, kb.HASH_CANON_INVHASH, m)
        assert np.array_equal(got[gi].astype(np.uint64), want), f"group {gi}"
        first += g
    batch.destroy()
