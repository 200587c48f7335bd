#!/usr/bin/env python
"""Host ingest: FASTQ / FASTQ.gz -> multi-threaded parser -> pinned packs -> kmu_seqbatch_from_ascii ->
ProbMinHash3a, the loop of src/bin/datasketcher.rs:236-300.  Prints one JSON line per input: Gbases/s of the parser alone,
of the whole loop, of the serial reader (kmu_fastx_*), and the fraction of the PCIe H2D rate the ASCII upload uses.

  python scripts/bench_ingest.py [--mbases 1000] [--read-len 6000] [--threads 0]"""
import argparse
import gzip
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kmerutils_b200 as kb  # noqa: E402
from kmerutils_b200 import io as kio  # noqa: E402


def write_fastq(path, nreads, L, seed, frac_n=0.01):
    rng = np.random.default_rng(seed)
    hdr = np.frombuffer(b"@read/0000000 len=xxxx\n", dtype=np.uint8)
    row = np.empty((nreads, len(hdr) + L + 3 + L + 1), dtype=np.uint8)
    row[:, : len(hdr)] = hdr
    row[:, len(hdr): len(hdr) + L] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, (nreads, L), dtype=np.uint8)]
    bad = rng.random(nreads) < frac_n  # reads holding an N: dropped by the feeder (io.rs:41-48)
    row[bad, len(hdr) + L // 2] = ord("N")
    row[:, len(hdr) + L: len(hdr) + L + 3] = np.frombuffer(b"\n+\n", dtype=np.uint8)
    row[:, len(hdr) + L + 3: -1] = ord("I")
    row[:, -1] = ord("\n")
    row.tofile(path)
    return int(bad.sum())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mbases", type=int, default=1000)
    ap.add_argument("--read-len", type=int, default=6000)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--gz-mbases", type=int, default=200)
    args = ap.parse_args()
    tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    eng = kb.Engine(0)
    try:
        for kind, mb in (("fastq", args.mbases), ("fastq.gz", args.gz_mbases)):
            nreads = mb * 1_000_000 // args.read_len
            plain = os.path.join(tmp, "reads.fastq")
            nbad = write_fastq(plain, nreads, args.read_len, 7)
            path = plain
            if kind.endswith(".gz"):
                path = plain + ".gz"
                with open(plain, "rb") as fi, gzip.open(path, "wb", compresslevel=1) as fo:
                    shutil.copyfileobj(fi, fo, 1 << 24)
            fsize = os.path.getsize(path)
            bases = nreads * args.read_len
            # (a) parser alone
            t0 = time.perf_counter()
            got = 0
            with kio.IngestReader(path, args.threads) as rd:
                t_open = time.perf_counter() - t0
                while True:
                    p = rd.next()
                    if p is None:
                        break
                    got += int(p[1][p[2]])
                    rd.release(p[3])
                st = rd.stats()
            t_parse = time.perf_counter() - t0
            assert st["nb_read"] == nreads and st["nb_bad_read"] == nbad and got == (nreads - nbad) * args.read_len
            # (b) the whole loop: parse + H2D + 2-bit pack + sketch + signatures back
            kio.datasketcher_mt(eng, path, None, 8, 200, args.threads)  # warm-up (allocations)
            r = kio.datasketcher_mt(eng, path, None, 8, 200, args.threads)
            assert r["reads"] == nreads - nbad
            # (c) the serial reader feeding the same loop
            t0 = time.perf_counter()
            n = 0
            with kio.FastxReader(path, pack_bases=256 << 20) as srd:
                while True:
                    buf, off = srd.next_pack_raw(20000)
                    if len(off) <= 1:
                        break
                    n += len(off) - 1
            t_serial = time.perf_counter() - t0
            assert n == nreads - nbad
            ascii_gbs = got / r["seconds"] / 1e9
            print(json.dumps({
                "input": kind, "file_bytes": fsize, "reads": nreads, "read_len": args.read_len, "bases": bases,
                "host_threads": args.threads or os.cpu_count(),
                "parse_only_gbases_s": round(bases / t_parse / 1e9, 3), "parse_only_file_gbs": round(fsize / t_parse / 1e9, 3), "parse_open_s": round(t_open, 3),
                "loop_gbases_s": round(bases / r["seconds"] / 1e9, 3), "loop_seconds": round(r["seconds"], 3),
                "loop_gbases_s_without_open": round(bases / max(r["seconds"] - r["open_s"], 1e-9) / 1e9, 3),
                "parse_only_gbases_s_without_open": round(bases / max(t_parse - t_open, 1e-9) / 1e9, 3),
                "serial_reader_gbases_s": round(bases / t_serial / 1e9, 3),
                "loop_phases_s": {q: round(r[q], 3) for q in ("open_s", "wait_for_parser_s", "upload_and_pack_s", "sketch_s")},
                "packs": r["packs"],
                "ascii_h2d_gbs": round(ascii_gbs, 2), "pcie_fraction_of_55gbs": round(ascii_gbs / 55.0, 3),
                "limit": "host parsing" if kind == "fastq" else "serial gzip inflation (zlib)"}), flush=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
        eng.close()


if __name__ == "__main__":
    main()
