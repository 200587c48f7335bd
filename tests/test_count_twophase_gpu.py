"""Two-phase insertion (partition by table region + regioned insertion, kmu_count_part.cu) against the oracle's exact
multiset counts and against direct insertion: same table contents whatever the path.  The environment knobs
KMU_COUNT_REGION_KB / KMU_COUNT_TWO_PHASE_MIN_KEYS push small tables through the regioned path (many regions, several
chunks, the slab-overflow fallback)."""
import os

import numpy as np
import pytest

import kmerutils_b200 as kb
from test_count_gpu import genome_reads

pytestmark = pytest.mark.gpu


class knobs:
    def __init__(self, **kw):
        self.kw = {k: str(v) for k, v in kw.items()}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kw}
        os.environ.update(self.kw)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def table_contents(ctr):
    k, c = ctr.export(min_count=1)
    o = np.argsort(k, kind="stable")
    return k[o], c[o]


@pytest.mark.parametrize("k,ktype,region_kb,level1", [(31, kb.KMER64, 64, 512), (21, kb.KMER64, 16, 512), (16, kb.KMER16B32, 32, 512),
                                                      (12, kb.KMER32, 16, 512), (32, kb.KMER64, 64, 512),
                                                      # two levels: 4 coarse regions cut into fine ones before the insertion
                                                      (31, kb.KMER64, 16, 4), (16, kb.KMER16B32, 16, 2), (21, kb.KMER64, 32, 1)])
def test_two_phase_matches_oracle_and_direct(engine, oracle, k, ktype, region_kb, level1):
    rng = np.random.default_rng(100 + k)
    reads = genome_reads(oracle, 60 + k, 30000, 3000, 150, rng)
    reads += [b"ACGT" * 3, b"ACGTTGCA" * 40, b"C", oracle.synth_ascii(9, 0, 70000)]  # short, periodic, one long sequence
    batch, _ = engine.batch_from_ascii(reads)
    packed, off, nb = batch.download()
    keys, cnts = oracle.count_kmers(packed, off, nb, k, ktype, True)
    with knobs(KMU_COUNT_REGION_KB=region_kb, KMU_COUNT_TWO_PHASE_MIN_KEYS=1, KMU_COUNT_LEVEL1_BUCKETS=level1):
        ctr = engine.counter(k, ktype, capacity=max(len(keys), 1 << 16), count_bits=8)
        l0 = engine.launch_count()
        ctr.insert_seqs(batch, canonical=True)
        assert engine.launch_count() - l0 >= (2 if level1 >= 512 else 3)  # partition(s) + regioned insertion, not the direct kernel
        st = ctr.stats()
        assert st["nb_distinct"] == len(keys)
        assert st["nb_unique"] == int((cnts == 1).sum())
        assert st["nb_inserted"] == int(cnts.sum()) == batch.kmer_count(k)
        assert np.array_equal(ctr.get_count(keys), np.minimum(cnts, 255).astype(np.uint32))
        # a second batch on top (key array form): every count doubles
        allk, _ = engine.generate_kmers(batch, k, ktype, kb.HASH_CANON_RAW)
        if ktype == kb.KMER32:
            allk = allk & np.uint32(0x0FFFFFFF)  # the table is keyed by get_compressed_value() (kmer32bit.rs:173-178)
        ctr.insert_kmers(allk)
        assert np.array_equal(ctr.get_count(keys), np.minimum(2 * cnts, 255).astype(np.uint32))
        two = table_contents(ctr)
        ctr.destroy()
    with knobs(KMU_COUNT_DIRECT=1):
        ctr = engine.counter(k, ktype, capacity=max(len(keys), 1 << 16), count_bits=8)
        l0 = engine.launch_count()
        ctr.insert_seqs(batch, canonical=True)
        assert engine.launch_count() - l0 == 1
        ctr.insert_kmers(allk)
        direct = table_contents(ctr)
        ctr.destroy()
    assert np.array_equal(two[0], direct[0]) and np.array_equal(two[1], direct[1])
    batch.destroy()


def test_two_phase_chunks_and_overflow_fallback(engine, oracle):
    """Several chunks (tiny slab budget) and a batch whose k-mers all fall into one bucket (a homopolymer run): the
    slab of that bucket overflows, the chunk is redone by direct insertion, the counts stay exact."""
    k, ktype = 31, kb.KMER64
    rng = np.random.default_rng(7)
    reads = genome_reads(oracle, 77, 50000, 4000, 150, rng) + [b"A" * 300000, b"ACGT" * 20000]
    batch, _ = engine.batch_from_ascii(reads)
    packed, off, nb = batch.download()
    keys, cnts = oracle.count_kmers(packed, off, nb, k, ktype, True)
    # level1 = 1: the level-1 partition cannot overflow (one slab), the homopolymer overflows a FINE slab of level 2
    for slab_mb, level1 in ((2, 512), (64, 512), (64, 4), (64, 1)):
        with knobs(KMU_COUNT_REGION_KB=64, KMU_COUNT_TWO_PHASE_MIN_KEYS=1, KMU_COUNT_SLAB_MB=slab_mb, KMU_COUNT_LEVEL1_BUCKETS=level1):
            ctr = engine.counter(k, ktype, capacity=1 << 17, count_bits=32)
            ctr.insert_seqs(batch, canonical=True)
            st = ctr.stats()
            assert st["nb_distinct"] == len(keys) and st["nb_inserted"] == int(cnts.sum())
            assert np.array_equal(ctr.get_count(keys).astype(np.uint64), cnts.astype(np.uint64))
            ctr.destroy()
    batch.destroy()


def test_two_phase_default_geometry(engine):
    """The default geometry (64 MB regions) on a table of 256 MB: 2 M reads-worth of random 31-mers, conservation of the
    inserted k-mers and agreement with direct insertion on the table statistics."""
    nb = np.full(20000, 150, dtype=np.uint64)
    batch = engine.batch_synth(41, nb)
    res = []
    for direct in (False, True):
        with knobs(**({"KMU_COUNT_DIRECT": 1} if direct else {})):
            ctr = engine.counter(31, kb.KMER64, capacity=8_000_000, count_bits=8)  # 2^24 slots * 16 B = 256 MB
            for _ in range(2):
                ctr.insert_seqs(batch, canonical=True)
            st = ctr.stats()
            res.append((st["nb_distinct"], st["nb_unique"], st["nb_inserted"], st["hist"].tolist()))
            ctr.destroy()
    assert res[0] == res[1]
    assert res[0][2] == 2 * batch.kmer_count(31)
    batch.destroy()


def test_exchange_scatter_single_rank(engine, oracle):
    """The fused exchange kernel with one owner writing into its own buffer, then the slab insertion: the path every
    rank runs in a multi-GPU round (tests/dist_check.py covers two ranks)."""
    import torch
    k, ktype = 31, kb.KMER64
    rng = np.random.default_rng(3)
    reads = genome_reads(oracle, 12, 40000, 5000, 150, rng)
    batch, _ = engine.batch_from_ascii(reads)
    packed, off, nb = batch.download()
    keys, cnts = oracle.count_kmers(packed, off, nb, k, ktype, True)
    with knobs(KMU_COUNT_REGION_KB=64):
        ctr = engine.counter(k, ktype, capacity=1 << 18, count_bits=8)
        nreg = ctr.exchange_regions(1)
        assert nreg > 1
        from kmerutils_b200.dist import exchange_slab_cap
        slab_cap = exchange_slab_cap(batch.kmer_count(k), 1, nreg)
        buf = torch.empty(nreg * slab_cap, dtype=torch.int64, device="cuda:0")
        sent, ovf = ctr.exchange_scatter(batch, 1, 0, slab_cap, [buf.data_ptr()], True)
        assert not ovf and int(sent.sum()) == batch.kmer_count(k)
        ctr.insert_slabs(buf.data_ptr(), slab_cap, sent.reshape(1, nreg))
        st = ctr.stats()
        assert st["nb_distinct"] == len(keys) and st["nb_unique"] == int((cnts == 1).sum())
        assert np.array_equal(ctr.get_count(keys), np.minimum(cnts, 255).astype(np.uint32))
        ctr.destroy()
    batch.destroy()


def test_exchange_two_owners_on_one_gpu(engine, oracle):
    """The multi-owner exchange on ONE GPU: two read sets play ranks 0 and 1, their scatter kernels bucket by (owner, region)
    or by owner only and store into both owners' receive buffers (all local here, peers over NVLink in a job), and every
    owner inserts the two senders' slabs (kmu_count_insert_slabs with two senders; the owner-only form partitions them by
    region first) -- the path every rank of a multi-GPU round runs, checked against the oracle's counts of the union."""
    import torch
    from kmerutils_b200.dist import exchange_slab_cap
    k, ktype = 31, kb.KMER64
    rng = np.random.default_rng(5)
    reads = genome_reads(oracle, 13, 60000, 6000, 150, rng)
    half = len(reads) // 2
    batches = [engine.batch_from_ascii(reads[:half])[0], engine.batch_from_ascii(reads[half:])[0]]
    whole, _ = engine.batch_from_ascii(reads)
    packed, off, nb = whole.download()
    keys, cnts = oracle.count_kmers(packed, off, nb, k, ktype, True)
    owner = (oracle.apply_hash(keys, k, ktype, kb.HASH_INVHASH) % np.uint64(2)).astype(np.int64)  # DispatchableT::dispatch
    # (owner, region) buckets in one pass and the plain regioned insertion at the receiver (up to 2048 buckets) / buckets by
    # owner only, the receiver partitions the senders' segments by region / by owner only and direct insertion (small job)
    for region_kb, min_keys, bucket_limit in ((64, 1, 2048), (64, 1, 1), (65536, 1 << 40, 1)):
        with knobs(KMU_COUNT_REGION_KB=region_kb, KMU_COUNT_TWO_PHASE_MIN_KEYS=min_keys, KMU_COUNT_EXCHANGE_BUCKETS=bucket_limit):
            ctrs = [engine.counter(k, ktype, capacity=1 << 18, count_bits=8) for _ in range(2)]
            nreg = ctrs[0].exchange_regions(2)
            assert (nreg > 1) == (bucket_limit > 1)
            slab_cap = exchange_slab_cap(max(b.kmer_count(k) for b in batches), 2, nreg)
            bufs = [torch.empty(2 * nreg * slab_cap, dtype=torch.int64, device="cuda:0") for _ in range(2)]
            sent = []
            for r in range(2):
                s_r, ovf = ctrs[r].exchange_scatter(batches[r], 2, r, slab_cap, [b_.data_ptr() for b_ in bufs], True)
                assert not ovf and int(s_r.sum()) == batches[r].kmer_count(k)
                sent.append(np.asarray(s_r).reshape(2, nreg))
            for o in range(2):
                ctrs[o].insert_slabs(bufs[o].data_ptr(), slab_cap, np.stack([sent[0][o], sent[1][o]]).astype(np.uint64))
            for o in range(2):
                mine = owner == o
                st = ctrs[o].stats()
                assert st["nb_distinct"] == int(mine.sum()) and st["nb_unique"] == int((cnts[mine] == 1).sum())
                assert np.array_equal(ctrs[o].get_count(keys[mine]), np.minimum(cnts[mine], 255).astype(np.uint32))
                assert not ctrs[o].get_count(keys[~mine]).any()
                ctrs[o].destroy()
    for b_ in batches + [whole]:
        b_.destroy()


@pytest.mark.parametrize("k,ktype", [(8, kb.KMER32), (16, kb.KMER16B32), (31, kb.KMER64)])
def test_two_phase_ragged_tiles(engine, oracle, k, ktype):
    """The partition kernel cuts a 1 KB tile of packed bytes into sequence segments and those into 128-position chunks:
    tiles with more than 32 sequences (16-byte records), sequences shorter than k and empty ones in between (segments
    without chunks), segments of one k-mer, and sequences spanning many tiles, in one batch."""
    rng = np.random.default_rng(7000 + k)
    reads = []
    for i in range(6000):
        m = i % 7
        if m == 0:
            n = 0
        elif m == 1:
            n = int(rng.integers(1, k))          # too short for a k-mer
        elif m == 2:
            n = k                                # exactly one k-mer
        elif m == 3:
            n = int(rng.integers(k, 64))         # one 16-byte record
        elif m == 4:
            n = int(rng.integers(120, 140))      # one chunk, nearly full
        elif m == 5:
            n = int(rng.integers(128, 600))      # several chunks
        else:
            n = int(rng.integers(4000, 9000))    # several tiles
        reads.append(oracle.synth_ascii(9000 + i, 0, n) if n else b"")
    batch, _ = engine.batch_from_ascii(reads)
    packed, off, nb = batch.download()
    keys, cnts = oracle.count_kmers(packed, off, nb, k, ktype, True)
    with knobs(KMU_COUNT_REGION_KB=16, KMU_COUNT_TWO_PHASE_MIN_KEYS=1, KMU_COUNT_LEVEL1_BUCKETS=512):
        ctr = engine.counter(k, ktype, capacity=max(len(keys), 1 << 16), count_bits=16)
        l0 = engine.launch_count()
        ctr.insert_seqs(batch, canonical=True)
        assert engine.launch_count() - l0 >= 2
        st = ctr.stats()
        assert st["nb_inserted"] == int(cnts.sum()) == batch.kmer_count(k)
        assert st["nb_distinct"] == len(keys)
        tk, tc = table_contents(ctr)
        o = np.argsort(keys, kind="stable")
        assert np.array_equal(tk, keys[o]) and np.array_equal(tc, np.minimum(cnts[o], 65535))
