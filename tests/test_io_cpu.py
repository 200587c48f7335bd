"""Host-side feeders / writers (CPU): FASTA/FASTQ packs with the reference's non-ACGT read rejection
(src/io.rs:12-72, datasketcher.rs:358-388), signature dump formats, params JSON."""
import struct

import numpy as np
import pytest

import kmerutils_b200 as kb
from kmerutils_b200 import io as kio


def write(tmp_path, name, text):
    p = tmp_path / name
    p.write_bytes(text)
    return str(p)


def test_fastq_packs_and_rejection(tmp_path):
    recs = [(b"r1", b"ACGTACGTAC"), (b"r2", b"ACGTNACGT"), (b"r3", b"acgtTTGA"), (b"r4", b"GGGG"), (b"r5", b"ACRT")]
    text = b"".join(b"@" + n + b" desc\n" + s + b"\n+\n" + b"I" * len(s) + b"\n" for n, s in recs)
    with kio.FastxReader(write(tmp_path, "a.fastq", text)) as rd:
        first = rd.next_pack(2)
        rest = rd.next_pack(10)
        assert rd.next_pack(10) == []
        st = rd.stats()
    assert first == [b"ACGTACGTAC", b"acgtTTGA"] and rest == [b"GGGG"]  # reads with N / R are dropped whole
    assert st == {"nb_read": 5, "nb_bad_read": 2, "nb_bases": 35, "nb_bad_bases": 2}


def test_reload_multiple_kmer_dump(tmp_path):
    """KmerCountReload::load_multiple_kmers_from_file (kmercount.rs:1209-1351) on hand-written dumps: 4-byte k-mer words
    with 1- or 2-byte counts, and the Kmer64bit form (u8 k + u64 value)."""
    recs = [(0xd02a013d, 2), (0x40a804f4, 300), (0x02a013d0, 7)]
    body = b"".join(struct.pack("<IH", k, c) for k, c in recs)
    p = write(tmp_path, "a.multi_kmer.bin", struct.pack("<IBBQ", 0xcea2bbff, 16, 2, 3) + body)
    d = kio.reload_multiple_kmers(p)
    assert (d["kmer_size"], d["count_bytes"], d["nb_declared"]) == (16, 2, 3)
    assert d["kmers"].tolist() == [k for k, _ in recs] and d["counts"].tolist() == [c for _, c in recs]
    body = b"".join(struct.pack("<IB", k, min(c, 255)) for k, c in recs)
    d = kio.reload_multiple_kmers(write(tmp_path, "b.bin", struct.pack("<IBBQ", 0xcea2bbff, 16, 1, 99) + body))
    assert d["counts"].tolist() == [2, 255, 7] and d["nb_declared"] == 99  # the records count, not the header
    body = b"".join(struct.pack("<BQH", 31, v, c) for v, c in [(0x123456789abcdef, 5), (42, 2)])
    d = kio.reload_multiple_kmers(write(tmp_path, "c.bin", struct.pack("<IBBQ", 0xcea2bbff, 31, 2, 2) + body))
    assert d["kmers"].tolist() == [0x123456789abcdef, 42] and d["counts"].tolist() == [5, 2]
    with pytest.raises(kb.KmuError):
        kio.reload_multiple_kmers(write(tmp_path, "d.bin", struct.pack("<IBBQ", 0xceabeadd, 16, 2, 0)))
    with pytest.raises(kb.KmuError):
        kio.reload_multiple_kmers(str(tmp_path / "missing.bin"))


def test_gzip_input(tmp_path):
    """needletail (src/io.rs:20-24) reads .gz transparently: so does the feeder (zlib), with the same packs and counts"""
    import gzip
    rng = np.random.default_rng(4)
    recs = [bytes(rng.choice(list(b"ACGT"), int(n)).astype(np.uint8)) for n in rng.integers(50, 3000, 400)]
    recs[7] = recs[7][:20] + b"N" + recs[7][21:]
    text = b"".join(b"@r%d\n" % i + s + b"\n+\n" + b"I" * len(s) + b"\n" for i, s in enumerate(recs))
    plain = write(tmp_path, "reads.fastq", text)
    gz = str(tmp_path / "reads.fastq.gz")
    with gzip.open(gz, "wb") as f:
        f.write(text)
    out = []
    for path in (plain, gz):
        with kio.FastxReader(path) as rd:
            got = []
            while True:
                pack = rd.next_pack(64)
                if not pack:
                    break
                got += pack
            out.append((got, rd.stats()))
    assert out[0] == out[1]
    assert out[1][0] == recs[:7] + recs[8:] and out[1][1]["nb_bad_read"] == 1
    # a truncated gzip stream is an error, not a short file
    bad = write(tmp_path, "cut.fastq.gz", open(gz, "rb").read()[:-200])
    with pytest.raises(kb.KmuError):
        with kio.FastxReader(bad) as rd:
            while rd.next_pack(64):
                pass


def test_fasta_multiline_crlf_and_quality_traps(tmp_path):
    fasta = b">s1 first\r\nACGT\r\nAC\r\n\r\n>s2\nTTTT\n>s3\nNNNN\n>s4\nGATTACA"
    with kio.FastxReader(write(tmp_path, "b.fa", fasta)) as rd:
        assert rd.next_pack() == [b"ACGTAC", b"TTTT", b"GATTACA"]
    # a quality line may start with '@' or '>' : the reader counts quality characters, it does not look for headers
    fq = b"@q1\nACGT\n+\n@>II\n@q2\nGGCC\n+q2\n>@@@\n"
    with kio.FastxReader(write(tmp_path, "c.fq", fq)) as rd:
        assert rd.next_pack() == [b"ACGT", b"GGCC"]
    with pytest.raises(kb.KmuError):
        kio.FastxReader(str(tmp_path / "missing.fq"))
    with kio.FastxReader(write(tmp_path, "d.fq", b"garbage\n")) as rd:
        with pytest.raises(kb.KmuInvalid):
            rd.next_pack()


def test_pack_buffer_boundary(tmp_path):
    reads = [b"A" * 30, b"C" * 30, b"G" * 30]
    text = b"".join(b">x\n" + r + b"\n" for r in reads)
    rd = kio.FastxReader(write(tmp_path, "e.fa", text), pack_bases=64)
    assert rd.next_pack() == reads[:2]  # the third read does not fit: it opens the next pack
    assert rd.next_pack() == reads[2:]
    assert rd.next_pack() == []
    assert rd.stats()["nb_read"] == 3
    rd.close()


def test_signature_dump_format(tmp_path):
    path = str(tmp_path / "sig.bin")
    rng = np.random.default_rng(0)
    a = rng.integers(0, 2**32, (3, 5), dtype=np.uint64).astype(np.uint32)
    b = rng.integers(0, 2**32, (2, 5), dtype=np.uint64).astype(np.uint32)
    with kio.SignatureDump(path, 5, 8) as d:
        d.write(a)
        d.write(b)
    raw = open(path, "rb").read()
    # u32 0xceabeadd | u32 sig_size = 4 | u32 sketch_size | u32 kmer_size, little endian (seqsketchjaccard.rs:402-409)
    assert struct.unpack("<4I", raw[:16]) == (0xceabeadd, 4, 5, 8)
    assert raw[16:] == np.concatenate([a, b]).astype("<u4").tobytes()
    hdr, sig = kio.read_signature_dump(path)
    assert hdr == {"sig_size": 4, "sketch_size": 5, "kmer_size": 8, "nb_signatures": 5}
    assert np.array_equal(sig, np.concatenate([a, b]))
    _, part = kio.read_signature_dump(path, first=1, count=3)
    assert np.array_equal(part, np.concatenate([a, b])[1:4])
    open(path, "r+b").write(b"\0\0\0\0")
    with pytest.raises(kb.KmuInvalid):  # "file is not a dump of signature"
        kio.read_signature_dump(path)


def test_block_dump_format(tmp_path):
    path = str(tmp_path / "blocks.bin")
    sig = np.arange(5 * 3, dtype=np.uint32).reshape(5, 3)
    numseq = np.array([7, 7, 7, 9, 9], dtype=np.uint32)
    numblock = np.array([0, 1, 2, 0, 1], dtype=np.uint32)
    with kio.BlockSignatureDump(path, 3, 8, 100) as d:
        d.write(sig, numseq, numblock)
    raw = open(path, "rb").read()
    # 17-byte header: sig_size is ONE byte in the writer (seqblocksketch.rs:216-224)
    assert struct.unpack("<IBIII", raw[:17]) == (0xceabbadd, 4, 3, 8, 100)
    body = np.frombuffer(raw[17:], dtype="<u4")
    want = [7, 3, 7, 0, 0, 1, 2, 7, 1, 3, 4, 5, 7, 2, 6, 7, 8, 9, 2, 9, 0, 9, 10, 11, 9, 1, 12, 13, 14]
    assert body.tolist() == want


def test_sketch_params_json(tmp_path):
    kio.dump_sketch_params(str(tmp_path), 8, 200, "PROB3A", "DNA")
    assert kio.reload_sketch_params(str(tmp_path)) == {"kmer_size": 8, "sketch_size": 200, "algo": "PROB3A", "data_t": "DNA"}


def _all_reads_single(path):
    out = []
    with kio.FastxReader(path) as rd:
        while True:
            pack = rd.next_pack(1000)
            if not pack:
                break
            out += pack
        return out, rd.stats()


def _all_reads_mt(path, nthreads, block_bytes):
    import ctypes as C
    out = []
    with kio.IngestReader(path, nthreads, block_bytes) as rd:
        while True:
            pack = rd.next()
            if pack is None:
                break
            addr, off, n, tok = pack
            buf = (C.c_uint8 * int(off[n])).from_address(addr)
            raw = bytes(buf)
            out += [raw[int(off[i]): int(off[i + 1])] for i in range(n)]
            rd.release(tok)
        return out, rd.stats()


@pytest.mark.parametrize("fmt", ["fastq", "fasta", "fastq.gz"])
def test_multithreaded_feeder_matches_the_serial_reader(tmp_path, fmt):
    """kmu_ingest_* (reader thread + parser threads + ordered packs) against kmu_fastx_*: the same accepted reads in the same
    order and the same statistics, for blocks far smaller than the file (boundaries fall everywhere: inside headers,
    sequences, quality lines that start with '@')."""
    import gzip
    rng = np.random.default_rng(11)
    recs = []
    for i in range(3000):
        n = int(rng.integers(1, 400))
        s = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), n).tobytes()
        if i % 17 == 0:
            s = s[: n // 2] + b"N" + s[n // 2:]       # dropped read
        if i % 23 == 0:
            s = s.lower()
        recs.append((b"read%d some description" % i, s))
    if fmt.startswith("fastq"):
        quals = [rng.choice(np.frombuffer(b"@>+IIIIFF#", dtype=np.uint8), len(s)).tobytes() for _, s in recs]
        text = b"".join(b"@" + h + b"\n" + s + b"\n+" + (h if i % 5 == 0 else b"") + b"\n" + q + b"\n"
                        for i, ((h, s), q) in enumerate(zip(recs, quals)))
    else:
        text = b"".join(b">" + h + b"\n" + b"\n".join(s[j: j + 60] for j in range(0, len(s), 60)) + b"\n" for h, s in recs)
    path = str(tmp_path / ("x." + fmt))
    if fmt.endswith(".gz"):
        with gzip.open(path, "wb") as f:
            f.write(text)
    else:
        (tmp_path / ("x." + fmt)).write_bytes(text)
    want, want_st = _all_reads_single(path)
    assert len(want) == sum(1 for _, s in recs if b"N" not in s.upper())
    for nthreads, block in ((1, 1 << 16), (4, 1 << 16), (8, 1 << 20)):
        got, st = _all_reads_mt(path, nthreads, block)
        assert got == want
        assert st == want_st


def test_multithreaded_feeder_long_records_and_errors(tmp_path):
    # one FASTA record larger than several blocks, then short ones; a FASTQ whose sequence spans two lines is refused
    rng = np.random.default_rng(12)
    big = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 700_000).tobytes()
    text = b">big\n" + b"\n".join(big[j: j + 80] for j in range(0, len(big), 80)) + b"\n>s2\nACGT\n>s3\nGGNA\n>s4\nTT\n"
    (tmp_path / "big.fa").write_bytes(text)
    got, st = _all_reads_mt(str(tmp_path / "big.fa"), 3, 1 << 16)
    assert got == [big, b"ACGT", b"TT"] and st["nb_bad_read"] == 1 and st["nb_read"] == 4
    (tmp_path / "ml.fq").write_bytes(b"@r1\nACGT\nACGT\n+\nIIIIIIII\n" * 50)
    with pytest.raises(kb.KmuError):
        _all_reads_mt(str(tmp_path / "ml.fq"), 2, 1 << 16)
    with pytest.raises(kb.KmuError):
        kio.IngestReader(str(tmp_path / "missing.fq"))
    (tmp_path / "empty.fq").write_bytes(b"")
    assert _all_reads_mt(str(tmp_path / "empty.fq"), 2, 1 << 16)[0] == []
