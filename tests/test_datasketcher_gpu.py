"""End to end: FASTQ file -> packs of accepted reads -> GPU ProbMinHash3a -> signature dump (datasketcher.rs:236-300)."""
import numpy as np
import pytest

import kmerutils_b200 as kb
from kmerutils_b200 import io as kio

pytestmark = pytest.mark.gpu


def test_datasketcher_end_to_end(engine, oracle, tmp_path):
    rng = np.random.default_rng(2)
    reads = [oracle.synth_ascii(77, int(rng.integers(0, 1 << 30)), int(n)) for n in rng.integers(150, 4000, 230)]
    reads[5] = reads[5][:100] + b"N" + reads[5][101:]   # dropped (datasketcher.rs:367-371)
    reads[17] = reads[17].lower()                        # lower case is valid (alphabet.rs:157-159)
    fq = tmp_path / "reads.fastq"
    with open(fq, "wb") as f:
        for i, r in enumerate(reads):
            f.write(b"@read%d\n" % i + r + b"\n+\n" + b"I" * len(r) + b"\n")
    dump = str(tmp_path / "sig.bin")
    n = kio.datasketcher(engine, str(fq), dump, kmer_size=8, sketch_size=200, pack=64)
    assert n == len(reads) - 1
    hdr, sig = kio.read_signature_dump(dump)
    assert hdr == {"sig_size": 4, "sketch_size": 200, "kmer_size": 8, "nb_signatures": n}
    kept = [r for i, r in enumerate(reads) if i != 5]
    for i in (0, 4, 5, 16, 100, n - 1):
        r = kept[i].upper()
        want = oracle.sketch_pmh3a_seq(oracle.pack_2bit(r), len(r), 8, kb.KMER32, kb.HASH_CANON_INVHASH, 200)
        assert np.array_equal(sig[i].astype(np.uint64), want)
    # block mode (datasketcher -b): BlockSeqSketcher dump
    bdump = str(tmp_path / "blocks.bin")
    nb = kio.datasketcher(engine, str(fq), bdump, kmer_size=8, sketch_size=40, pack=50, block_size=500)
    assert nb == n
    raw = np.frombuffer(open(bdump, "rb").read()[17:], dtype="<u4")
    # first sequence: numseq 0, ceil(L / 500) blocks
    assert raw[0] == 0 and raw[1] == (len(kept[0]) + 499) // 500
    want0 = oracle.blocksketch_seq(oracle.pack_2bit(kept[0].upper()), len(kept[0]), 8, 40, 500)
    assert np.array_equal(raw[4:44], want0[0])
