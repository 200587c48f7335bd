"""GPU parity of the counting table against the oracle's exact multiset counts
(KmerCounter semantics with zero filter false positives, kmercount.rs:241-288)."""
import numpy as np
import pytest

import kmerutils_b200 as kb
from test_pmh3a_gpu import oracle_batch

pytestmark = pytest.mark.gpu


def genome_reads(oracle, seed, genome_len, nreads, read_len, rng):
    """reads drawn from one random genome, random strand: overlapping reads give multiplicities > 1"""
    g = oracle.synth_ascii(seed, 0, genome_len)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    reads = []
    for _ in range(nreads):
        s = int(rng.integers(0, genome_len - read_len + 1))
        r = g[s:s + read_len]
        if rng.integers(0, 2):
            r = r.translate(comp)[::-1]
        reads.append(r)
    return reads


@pytest.mark.parametrize("k,ktype", [(31, kb.KMER64), (21, kb.KMER64), (32, kb.KMER64), (16, kb.KMER16B32),
                                     (12, kb.KMER32), (5, kb.KMER32)])
def test_count_parity(engine, oracle, k, ktype):
    rng = np.random.default_rng(k)
    reads = genome_reads(oracle, 30 + k, 20000, 1500, 150, rng)
    reads += [b"ACGT" * 3, b"A" * 400, b"ACGTTGCA" * 40, b"C"]  # short, homopolymer, periodic
    batch, _ = engine.batch_from_ascii(reads)
    packed, off, nb = batch.download()
    keys, cnts = oracle.count_kmers(packed, off, nb, k, ktype, True)
    ctr = engine.counter(k, ktype, capacity=len(keys), count_bits=8)
    ctr.insert_seqs(batch, canonical=True)
    st = ctr.stats()
    assert st["nb_distinct"] == len(keys)
    assert st["nb_unique"] == int((cnts == 1).sum())
    assert st["nb_inserted"] == int(cnts.sum()) == batch.kmer_count(k)
    got = ctr.get_count(keys)
    assert np.array_equal(got, np.minimum(cnts, 255).astype(np.uint32))  # 8-bit saturation (kmercount.rs:1615)
    # k-mers never inserted count 0 (kmercount.rs:1612)
    absent = np.setdiff1d(rng.integers(0, 1 << min(2 * k, 62), 2000).astype(np.uint64), keys)
    assert not ctr.get_count(absent).any()
    want_hist = np.bincount(np.minimum(cnts, 255).astype(np.int64), minlength=256)
    assert np.array_equal(st["hist"].astype(np.int64), want_hist)
    # multiple k-mers, as dumped by threaded_dump_kmer_counter (kmercount.rs:653-791)
    mk, mc = ctr.export(min_count=2)
    order = np.argsort(mk)
    sel = cnts >= 2
    assert np.array_equal(mk[order].astype(np.uint64), keys[sel])
    assert np.array_equal(mc[order].astype(np.uint64), cnts[sel])
    ctr.destroy()


def test_count_reference_semantics(engine):
    # kmercount.rs:1523-1575: k0 inserted once, k1 twice among many others -> counts 1 and 2
    rng = np.random.default_rng(5)
    others = rng.integers(0, 1 << 32, 1_000_000, dtype=np.uint64).astype(np.uint32)
    k0, k1 = np.uint32(0x12345678), np.uint32(0x0BADF00D)
    others = others[(others != k0) & (others != k1)]
    ctr = engine.counter(16, kb.KMER16B32, capacity=1_100_000, count_bits=8)
    ctr.insert_kmers(np.array([k0, k1], dtype=np.uint32))
    ctr.insert_kmers(others)
    ctr.insert_kmers(np.array([k1], dtype=np.uint32))
    c = ctr.get_count(np.array([k0, k1], dtype=np.uint32))
    vals, mult = np.unique(others, return_counts=True)
    assert c.tolist() == [1, 2]
    st = ctr.stats()
    assert st["nb_distinct"] == len(vals) + 2
    assert st["nb_unique"] == int((mult == 1).sum()) + 1
    ctr.destroy()


def test_count_saturation(engine):
    # kmercount.rs:1579-1621: keys of the upper half inserted many times -> 255 with 8-bit counters,
    # the lower half stays 0
    upper = (np.arange(1000, dtype=np.uint64) + (1 << 31)).astype(np.uint32)
    ctr = engine.counter(16, kb.KMER16B32, capacity=4096, count_bits=8)
    for _ in range(3):
        ctr.insert_kmers(np.repeat(upper, 100))
    assert (ctr.get_count(upper) == 255).all()
    assert not ctr.get_count(np.arange(1000, dtype=np.uint32)).any()
    ctr16 = engine.counter(16, kb.KMER16B32, capacity=4096, count_bits=16)
    ctr16.insert_kmers(np.repeat(upper, 300))
    assert (ctr16.get_count(upper) == 300).all()
    ctr.destroy()
    ctr16.destroy()


def test_count_sentinel_key_and_overflow(engine):
    # the u64 value ~0 (32 T's, only reachable non-canonically) is the table's empty mark: counted apart
    ctr = engine.counter(32, kb.KMER64, capacity=1024)
    allT = np.array([2**64 - 1] * 3 + [5, 5, 7], dtype=np.uint64)
    ctr.insert_kmers(allT)
    assert ctr.get_count(np.array([2**64 - 1, 5, 7, 9], dtype=np.uint64)).tolist() == [3, 2, 1, 0]
    st = ctr.stats()
    assert (st["nb_distinct"], st["nb_unique"], st["nb_inserted"]) == (3, 1, 6)
    mk, mc = ctr.export(min_count=2)
    assert sorted(zip(mk.tolist(), mc.tolist())) == [(5, 2), (2**64 - 1, 3)]
    ctr.destroy()
    small = engine.counter(31, kb.KMER64, capacity=16)
    with pytest.raises(kb.KmuError) as ei:
        small.insert_kmers(np.arange(1, 5000, dtype=np.uint64))
    assert ei.value.code == 4  # KMU_EOVERFLOW: never degrades silently
    small.destroy()


def test_count_bad_arguments(engine):
    with pytest.raises(kb.KmuInvalid):
        engine.counter(15, kb.KMER32, 100)  # Kmer32bit holds k <= 14
    with pytest.raises(kb.KmuInvalid):
        engine.counter(31, kb.KMER64, 100, count_bits=0)


@pytest.mark.parametrize("k,ktype,nparts", [(31, kb.KMER64, 8), (16, kb.KMER16B32, 2), (11, kb.KMER32, 5)])
def test_partition_by_owner(engine, oracle, k, ktype, nparts):
    rng = np.random.default_rng(nparts)
    nb = rng.integers(1, 600, 300).astype(np.uint64)
    batch = engine.batch_synth(77, nb)
    packed, off = oracle_batch(oracle, 77, nb)
    kmers, counts = engine.count_partition(batch, k, ktype, nparts, canonical=True)
    assert int(counts.sum()) == len(kmers) == batch.kmer_count(k)
    # every bucket holds exactly the k-mers DispatchableT::dispatch sends to that receiver
    keys, cnts = oracle.count_kmers(packed, off, nb, k, ktype, True)
    start = 0
    seen = []
    for p in range(nparts):
        part = kmers[start:start + int(counts[p])].astype(np.uint64)
        start += int(counts[p])
        if len(part):
            uniq = np.unique(part)
            assert (oracle.dispatch(uniq[:200], ktype, nparts) == p).all()
        seen.append(part)
    allk, mult = np.unique(np.concatenate(seen), return_counts=True)
    assert np.array_equal(allk, keys) and np.array_equal(mult.astype(np.uint64), cnts)
    # receive side: inserting the buckets rank by rank gives the same table as inserting the sequences
    ctr = engine.counter(k, ktype, capacity=len(keys))
    for part in seen:
        ctr.insert_kmers(part.astype(kb.val_dtype(ktype)))
    assert np.array_equal(ctr.get_count(keys), np.minimum(cnts, 255).astype(np.uint32))
    ctr.destroy()


def test_sample_reads_and_sharded_count_single_rank(engine, oracle):
    # the synthetic short reads of config C3 are the same on the device and in the oracle
    glen = 50_000
    genome = engine.batch_synth(3, np.array([glen], dtype=np.uint64))
    gp, _ = oracle_batch(oracle, 3, np.array([glen], dtype=np.uint64))
    reads = engine.batch_sample_reads(genome, 3, first_read=1000, nreads=400, read_len=150, err_ppm=5000)
    packed, off, nb = reads.download()
    want = oracle.sample_reads(gp, glen, 3, 1000, 400, 150, 5000)
    for i in (0, 1, 2, 57, 399):
        assert oracle.unpack_2bit(packed[int(off[i]): int(off[i]) + 38], 150) == want[i]
    # count through the multi-GPU driver with one rank: partition -> (no exchange) -> insert
    from kmerutils_b200 import dist as kd
    keys, cnts = oracle.count_kmers(packed, off, nb, 31, kb.KMER64, True)
    counter, st = kd.count_sharded(engine, reads, 31, kb.KMER64, capacity_per_rank=len(keys))
    assert (st["nb_distinct"], st["nb_unique"], st["nb_inserted"]) == (len(keys), int((cnts == 1).sum()), int(cnts.sum()))
    assert np.array_equal(kd.query_sharded(engine, counter, keys, kb.KMER64), np.minimum(cnts, 255).astype(np.uint32))
    counter.destroy()


def test_count_two_phase_large_table(engine, oracle, monkeypatch):
    # a table far larger than L2 takes a large batch in two phases by default (k-mers grouped by table region first,
    # kmu_count_part.cu); KMU_COUNT_DIRECT=1 forces direct insertion: same table either way
    glen = 300_000
    genome = engine.batch_synth(9, np.array([glen], dtype=np.uint64))
    reads = engine.batch_sample_reads(genome, 9, 0, 40_000, 150, 5000)
    packed, off, nb = reads.download()
    keys, cnts = oracle.count_kmers(packed, off, nb, 31, kb.KMER64, True)
    sel = slice(None, None, 7)
    ctr = engine.counter(31, kb.KMER64, capacity=9_000_000)
    assert ctr.capacity() * 16 >= 256 << 20
    ctr.insert_seqs(reads, canonical=True)
    assert engine.last_times()["launches"] == 2  # partition + regioned insertion
    st = ctr.stats()
    assert (st["nb_distinct"], st["nb_unique"], st["nb_inserted"]) == (len(keys), int((cnts == 1).sum()), int(cnts.sum()))
    assert np.array_equal(ctr.get_count(keys[sel]), np.minimum(cnts[sel], 255).astype(np.uint32))
    monkeypatch.setenv("KMU_COUNT_DIRECT", "1")
    ctr2 = engine.counter(31, kb.KMER64, capacity=9_000_000)
    ctr2.insert_seqs(reads, canonical=True)
    assert engine.last_times()["launches"] == 1  # direct path
    assert ctr2.stats()["nb_distinct"] == len(keys)
    assert np.array_equal(ctr2.get_count(keys[sel]), ctr.get_count(keys[sel]))
    ctr.destroy()
    ctr2.destroy()


def test_p2p_exchange_single_rank(engine, oracle):
    # the peer-to-peer form of the exchange with one rank: the bucket is written through a device pointer table into
    # an IPC-exportable buffer, then inserted
    from kmerutils_b200 import dist as kd
    rng = np.random.default_rng(3)
    nb = rng.integers(31, 900, 400).astype(np.uint64)
    batch = engine.batch_synth(88, nb)
    packed, off = oracle_batch(oracle, 88, nb)
    keys, cnts = oracle.count_kmers(packed, off, nb, 31, kb.KMER64, True)
    xchg = kd.P2PExchange(engine)
    ctr = engine.counter(31, kb.KMER64, capacity=len(keys))
    for _ in range(2):  # second round reuses the buffer
        n = kd.count_sharded_p2p(engine, batch, 31, kb.KMER64, ctr, xchg)
        assert n == batch.kmer_count(31)
    assert np.array_equal(ctr.get_count(keys), np.minimum(2 * cnts, 255).astype(np.uint32))
    # four owners on one GPU: buckets land at the right offsets of four buffers
    counts = engine.count_partition_counts(batch, 31, kb.KMER64, 4)
    import torch
    bufs = [torch.zeros(int(c) + 8, dtype=torch.int64, device="cuda:0") for c in counts]
    engine.count_partition_scatter(batch, 31, kb.KMER64, 4, [b.data_ptr() for b in bufs], [3, 0, 5, 1])
    for p, (b, o) in enumerate(zip(bufs, [3, 0, 5, 1])):
        got = b.cpu().numpy().astype(np.uint64)[o:o + int(counts[p])]
        assert (oracle.dispatch(np.unique(got)[:100], kb.KMER64, 4) == p).all()
    allk = np.concatenate([b.cpu().numpy().astype(np.uint64)[o:o + int(c)] for b, o, c in zip(bufs, [3, 0, 5, 1], counts)])
    u, m = np.unique(allk, return_counts=True)
    assert np.array_equal(u, keys) and np.array_equal(m.astype(np.uint64), cnts)
    ctr.destroy()
    xchg.close()


@pytest.mark.parametrize("k,ktype", [(31, kb.KMER64), (16, kb.KMER16B32), (12, kb.KMER32)])
def test_multiple_kmer_dump_format(engine, oracle, tmp_path, k, ktype):
    import struct
    rng = np.random.default_rng(k)
    reads = genome_reads(oracle, 5, 8000, 600, 150, rng)
    batch, _ = engine.batch_from_ascii(reads)
    packed, off, nb = batch.download()
    keys, cnts = oracle.count_kmers(packed, off, nb, k, ktype, True)
    ctr = engine.counter(k, ktype, capacity=len(keys), count_bits=8)
    ctr.insert_seqs(batch)
    path = str(tmp_path / "x.multi_kmer.bin")
    n = ctr.dump_multiple(path, count_bytes=2)
    raw = open(path, "rb").read()
    magic, ksz, cb, nk = struct.unpack("<IBBQ", raw[:14])  # kmercount.rs:139-145
    assert (magic, ksz, cb, nk) == (0xcea2bbff, k, 2, n) and n == int((cnts >= 2).sum())
    recsz = (9 if ktype == kb.KMER64 else 4) + 2
    assert len(raw) == 14 + n * recsz
    got = {}
    for i in range(n):
        r = raw[14 + i * recsz: 14 + (i + 1) * recsz]
        if ktype == kb.KMER64:  # Kmer64bit::dump = u8 k + u64 value (kmer64bit.rs:98-104)
            assert r[0] == k
            key = struct.unpack("<Q", r[1:9])[0]
        else:
            key = struct.unpack("<I", r[:4])[0]
            if ktype == kb.KMER32:  # the word carries k in its top four bits
                assert key >> 28 == k
                key &= 0x0FFFFFFF
        got[key] = struct.unpack("<H", r[-2:])[0]
    want = {int(a): min(int(c), 255) for a, c in zip(keys, cnts) if c >= 2}
    assert got == want
    # and back through the reloader (KmerCountReload::load_multiple_kmers_from_file)
    from kmerutils_b200 import io as kio
    d = kio.reload_multiple_kmers(path)
    assert (d["kmer_size"], d["count_bytes"], d["nb_declared"]) == (k, 2, n)
    mask = 0x0FFFFFFF if ktype == kb.KMER32 else 0xFFFFFFFFFFFFFFFF
    assert {int(a) & mask: int(c) for a, c in zip(d["kmers"], d["counts"])} == want
    ctr.destroy()
