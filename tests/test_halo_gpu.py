"""Long sequences cut into chunks with a k - 1 halo (SURVEY 5 "long sequences", config C5a; reference analogue: the block split
+ SetSketcher::merge of HyperLogLogSketch::sketch_compressedkmer_seqs, setsketchert.rs:811-895): chunk c of a sequence is
bases [c L / n, (c + 1) L / n + k - 1), so every k-mer of the sequence starts in exactly one chunk; the chunk sketches merge
by element-wise max (SetSketch) / min (SuperMinHash) into the sketch of the whole sequence.  This is what every rank does in
bench.py's C5a at N > 1 (the merge there is an NCCL allreduce; tests/dist_check.py runs it on two GPUs)."""
import numpy as np
import pytest

import kmerutils_b200 as kb
from test_pmh3a_gpu import oracle_batch

pytestmark = pytest.mark.gpu


def chunks_with_halo(engine, batch, nb, nchunks, k):
    idx = np.repeat(np.arange(len(nb), dtype=np.uint64), nchunks)
    c = np.tile(np.arange(nchunks, dtype=np.uint64), len(nb))
    L = nb[idx.astype(np.int64)]
    begin = (L * c) // np.uint64(nchunks)
    end = np.minimum(L, (L * (c + np.uint64(1))) // np.uint64(nchunks) + np.uint64(k - 1))
    return [engine.batch_slices(batch, idx[c == r], begin[c == r], end[c == r]) for r in range(nchunks)]


@pytest.mark.parametrize("nchunks", [2, 3, 8])
def test_halo_chunks_merge_to_the_whole_sequence_sketch(engine, oracle, nchunks):
    k = 21
    nb = np.array([100_000_000, 37, 1_234_567], dtype=np.uint64)  # a 100 Mb sequence, one shorter than k, a mid-sized one
    batch = engine.batch_synth(5, nb)
    prm = (1.001, 4096, 20.0, 65534)
    whole_ssk = engine.sketch_setsketch(batch, k, kb.KMER64, kb.HASH_CANON_INVHASH, prm, np.uint16, whole=True)
    whole_smh = engine.sketch_superminhash_whole(batch, k, kb.KMER64, kb.HASH_CANON_INVHASH, 1024)
    parts = chunks_with_halo(engine, batch, nb, nchunks, k)
    assert sum(p.kmer_count(k) for p in parts) == batch.kmer_count(k)  # every k-mer starts in exactly one chunk
    ssk = np.stack([engine.sketch_setsketch(p, k, kb.KMER64, kb.HASH_CANON_INVHASH, prm, np.uint16, whole=True) for p in parts])
    smh = np.stack([engine.sketch_superminhash_whole(p, k, kb.KMER64, kb.HASH_CANON_INVHASH, 1024) for p in parts])
    assert np.array_equal(ssk.max(axis=0), whole_ssk)
    assert np.array_equal(smh.min(axis=0), whole_smh)
    for p in parts:
        p.destroy()
    if nchunks == 2:  # once: the single-sequence sketch of the oracle itself (100 M k-mers on the CPU)
        packed, off = oracle_batch(oracle, 5, nb)
        want = oracle.sketch_setsketch_seqs(packed, off, nb, k, kb.KMER64, kb.HASH_CANON_INVHASH, prm)
        assert np.array_equal(whole_ssk, want)
    batch.destroy()
