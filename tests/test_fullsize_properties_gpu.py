"""BASELINE.json's configurations C3 / C4 / C5 at sizes the CPU oracle cannot replay in seconds, checked through
size-independent properties of the domain (the parity tests proper compare with the oracle on smaller inputs):

* canonical k-mers make every result strand-symmetric: a sequence and its reverse complement hold the same multiset of
  canonical k-mers, hence bit-identical signatures and counts (the reference's own test `J(a, revcomp a) >= 1`,
  src/sketching/seqsketchjaccard.rs:785, at genome size);
* SetSketch / SuperMinHash registers merge by element-wise max / min (setsketchert.rs:876-882): sketch(A u B) equals
  merge(sketch(A), sketch(B)), and sketching the same data twice changes nothing (idempotence);
* a counting table conserves mass: sum_c c * hist[c] equals the number of k-mers inserted while nothing saturates, and
  inserting everything a second time doubles every count;
* every signature slot is the hash of a k-mer that occurs in the input.
"""
import numpy as np
import pytest

import kmerutils_b200 as kb
from test_aa_gpu import aa_oracle_batch

pytestmark = pytest.mark.gpu


def revcomp_batch(engine, oracle, batch):
    """the reverse complements of the sequences of `batch`, in the same order (host side, test infrastructure)"""
    packed, off, nb = batch.download()
    rows = [oracle.seq_revcomp(packed[int(o): int(o) + (int(n) + 3) // 4], int(n)) for o, n in zip(off, nb)]
    return engine.batch_from_sequences(rows, nb)


# ---- C3: Illumina-like 150-base reads from a 100 Mb genome, canonical 31-mers, 8-bit counters ---------------------
def test_c3_counting_conservation_and_doubling(engine):
    genome = engine.batch_synth(3, np.array([100_000_000], dtype=np.uint64))
    reads_per_round, rounds, k = 8_000_000, 4, 31  # 4.8 Gbases: a quarter of C3 (the full set is 17 such rounds)
    kmers_per_round = reads_per_round * (150 - k + 1)
    ctr = engine.counter(k, kb.KMER64, capacity=900_000_000)
    for r in range(rounds):
        reads = engine.batch_sample_reads(genome, 3, r * reads_per_round, reads_per_round, 150, 5000)
        ctr.insert_seqs(reads, canonical=True)
        reads.destroy()
    st = ctr.stats()
    hist = st["hist"].astype(np.int64)
    c = np.arange(256, dtype=np.int64)
    assert st["nb_inserted"] == rounds * kmers_per_round
    assert hist[255] == 0, "coverage 48x: no 31-mer of a random genome reaches 255"
    assert int((hist * c).sum()) == rounds * kmers_per_round          # conservation of mass
    assert int(hist[1:].sum()) == st["nb_distinct"] and int(hist[1]) == st["nb_unique"]
    # substitution errors create the unique k-mers; the genome's own k-mers sit near the coverage
    assert st["nb_unique"] > 0.3 * st["nb_distinct"]
    gk, _ = engine.generate_kmers(engine.batch_synth(3, np.array([100_000], dtype=np.uint64)), k, kb.KMER64, kb.HASH_CANON_RAW)
    cov = ctr.get_count(gk).astype(np.float64)  # the first 100 kb of the genome (same SplitMix64 stream, same seed)
    expect = rounds * reads_per_round * (150 - k + 1) / 100_000_000 * 0.995 ** k
    assert abs(cov.mean() - expect) < 0.1 * expect, (cov.mean(), expect)
    # the same reads once more: every multiplicity doubles
    for r in range(rounds):
        reads = engine.batch_sample_reads(genome, 3, r * reads_per_round, reads_per_round, 150, 5000)
        ctr.insert_seqs(reads, canonical=True)
        reads.destroy()
    st2 = ctr.stats()
    hist2 = st2["hist"].astype(np.int64)
    assert st2["nb_distinct"] == st["nb_distinct"] and st2["nb_unique"] == 0
    top = 127
    assert np.array_equal(hist2[2:2 * top + 1:2], hist[1:top + 1]) and not hist2[1:2 * top:2].any()
    ctr.destroy()
    genome.destroy()


def test_c3_strand_symmetry(engine, oracle):
    genome = engine.batch_synth(33, np.array([2_000_000], dtype=np.uint64))
    reads = engine.batch_sample_reads(genome, 33, 0, 200_000, 150, 5000)
    rc = revcomp_batch(engine, oracle, reads)
    a = engine.counter(31, kb.KMER64, capacity=30_000_000)
    b = engine.counter(31, kb.KMER64, capacity=30_000_000)
    a.insert_seqs(reads, canonical=True)
    b.insert_seqs(rc, canonical=True)
    sa, sb = a.stats(), b.stats()
    assert sa["nb_distinct"] == sb["nb_distinct"] and sa["nb_unique"] == sb["nb_unique"]
    assert np.array_equal(sa["hist"], sb["hist"])
    ka, ca = a.export(1)
    assert np.array_equal(b.get_count(ka), ca)
    for x in (a, b, reads, rc, genome):
        x.destroy()


# ---- C4: 5 Mb genomes, k = 16 Kmer16b32bit, ProbMinHash3a and SuperMinHash with 12 000 slots -------------------------
def test_c4_genome_sketches_strand_symmetric_and_consistent(engine, oracle, monkeypatch):
    ngen, glen, k, m = 6, 5_000_000, 16, 12_000
    genomes = engine.batch_synth(4, np.full(ngen, glen, dtype=np.uint64))
    rc = revcomp_batch(engine, oracle, genomes)
    ones = np.ones(ngen, dtype=np.uint64)
    sig = engine.sketch_pmh3a_groups(genomes, ones, k, kb.KMER16B32, kb.HASH_CANON_INVHASH, m)
    sig_rc = engine.sketch_pmh3a_groups(rc, ones, k, kb.KMER16B32, kb.HASH_CANON_INVHASH, m)
    assert sig.shape == (ngen, m) and np.array_equal(sig, sig_rc)
    # the per-sequence entry point and the whole-file one agree on a one-contig genome
    view = engine.batch_view(genomes, 2, 1)
    assert np.array_equal(engine.sketch_pmh3a_whole(view, k, kb.KMER16B32, kb.HASH_CANON_INVHASH, m).astype(np.uint32), sig[2])
    per_seq = engine.sketch_pmh3a(view, k, kb.KMER16B32, kb.HASH_CANON_INVHASH, m)  # routed to the whole-file procedure
    assert np.array_equal(per_seq[0], sig[2])
    monkeypatch.setenv("KMU_PMH3A_TEAM_ONLY", "1")  # the team kernel (one SM per genome) gives the same
    per_seq = engine.sketch_pmh3a(view, k, kb.KMER16B32, kb.HASH_CANON_INVHASH, m)
    monkeypatch.delenv("KMU_PMH3A_TEAM_ONLY")
    assert np.array_equal(per_seq[0], sig[2])
    # every slot names a k-mer of its genome
    kmers, _ = engine.generate_kmers(view, k, kb.KMER16B32, kb.HASH_CANON_INVHASH)
    assert np.isin(sig[2], kmers).all()
    view.destroy()
    # different random genomes share (almost) nothing; a genome against itself is 1
    jac = engine.signature_jaccard(sig, sig)
    assert np.allclose(np.diag(jac), 1.0) and (jac[~np.eye(ngen, dtype=bool)] < 0.01).all()
    # SuperMinHash, f64, NoHashHasher: strand symmetry, and the whole-file signature of a genome cut into contigs is
    # the element-wise minimum of the contigs' signatures
    smh = engine.sketch_superminhash(genomes, k, kb.KMER16B32, kb.HASH_CANON_INVHASH, m)
    smh_rc = engine.sketch_superminhash(rc, k, kb.KMER16B32, kb.HASH_CANON_INVHASH, m)
    assert np.array_equal(smh.view(np.uint64), smh_rc.view(np.uint64))
    cuts = np.linspace(0, glen, 51).astype(np.uint64)
    contigs = engine.batch_slices(genomes, np.zeros(50, dtype=np.uint64), cuts[:-1], cuts[1:])
    whole = engine.sketch_superminhash_whole(contigs, k, kb.KMER16B32, kb.HASH_CANON_INVHASH, m)
    parts = engine.sketch_superminhash(contigs, k, kb.KMER16B32, kb.HASH_CANON_INVHASH, m)
    assert np.array_equal(whole.view(np.uint64), parts.min(axis=0).view(np.uint64))
    for x in (contigs, rc, genomes):
        x.destroy()


# ---- C5a: 24 chromosomes of 50..200 Mb (3.0 Gbases), k = 21 Kmer64bit, SetSketch default parameters ------------------
def test_c5_setsketch_whole_file_merge_and_idempotence(engine):
    nb = np.linspace(50e6, 200e6, 24)
    nb = np.rint(nb * (3.0e9 / nb.sum())).astype(np.uint64)
    chrom = engine.batch_synth(5, nb)
    k = 21
    whole = engine.sketch_setsketch(chrom, k, kb.KMER64, kb.HASH_CANON_INVHASH, None, np.uint16, whole=True)
    assert whole.shape == (4096,) and whole.min() > 0
    first, last = engine.batch_view(chrom, 0, 12), engine.batch_view(chrom, 12, 12)
    a = engine.sketch_setsketch(first, k, kb.KMER64, kb.HASH_CANON_INVHASH, None, np.uint16, whole=True)
    b = engine.sketch_setsketch(last, k, kb.KMER64, kb.HASH_CANON_INVHASH, None, np.uint16, whole=True)
    assert np.array_equal(np.maximum(a, b), whole)                     # SetSketcher::merge
    per_seq = engine.sketch_setsketch(chrom, k, kb.KMER64, kb.HASH_CANON_INVHASH, None, np.uint16)
    assert np.array_equal(per_seq.max(axis=0), whole)                  # 24-way merge of the per-sequence sketches
    # idempotence: the first half once more adds nothing
    twice = engine.batch_slices(chrom, np.concatenate([np.arange(24), np.arange(12)]).astype(np.uint64),
                                np.zeros(36, dtype=np.uint64), np.concatenate([nb, nb[:12]]))
    assert np.array_equal(engine.sketch_setsketch(twice, k, kb.KMER64, kb.HASH_CANON_INVHASH, None, np.uint16, whole=True), whole)
    # cardinality from the registers (Ertl 2021, eq. 12 with b -> 1: n ~ m / (a * sum b^-K)): within 6 % (3.8 sigma at m = 4096) of the truth
    bb, aa = 1.001, 20.0
    est = 4096 * (1 - 1 / bb) / (aa * np.log(bb) * np.sum(bb ** (-whole.astype(np.float64))))
    truth = float((nb - np.uint64(k - 1)).sum())
    assert abs(est - truth) < 0.06 * truth, (est, truth)
    for x in (twice, first, last, chrom):
        x.destroy()


# ---- C5b: proteome, 20 000 proteins, amino-acid 12-mers, ProbMinHash3a 400 slots: full size against the oracle ------
def test_c5_proteome_full_size(engine, oracle):
    rng = np.random.default_rng(5)
    nres = np.clip(np.rint(np.exp(rng.normal(5.6, 0.6, 20_000))), 50, 5000).astype(np.uint64)
    batch = engine.batch_synth_aa(5, nres)
    buf, off = aa_oracle_batch(oracle, 5, nres)
    got = engine.sketch_pmh3a(batch, 12, kb.KMERAA64, kb.HASH_MASKED_VALUE, 400)
    want = oracle.sketch_pmh3a_batch(buf, off, nres, 12, kb.KMERAA64, kb.HASH_MASKED_VALUE, 400)
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert len(bad) == 0, f"{len(bad)} proteins differ, first {bad[:5]}"
    batch.destroy()
