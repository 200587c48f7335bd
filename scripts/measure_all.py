#!/usr/bin/env python
"""Device-timed throughput of every kernel family on its BASELINE.json shape (reduced counts, same per-unit
sizes), one JSON line each: Gbases/s, algorithmic GB/s (SURVEY 8d bytes per base) and fraction of the measured
HBM peak.  Not the headline bench (that is bench.py on C2); these lines feed DESIGN.md section 5."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kmerutils_b200 as kb  # noqa: E402
from kmerutils_b200 import workloads  # noqa: E402

PEAK = 6514.8
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def line(name, bases, ms, bytes_per_base, **kw):
    gb = bases / (ms * 1e-3) / 1e9
    out = {"kernel": name, "bases": int(bases), "ms": round(ms, 3), "gbases_s": round(gb, 2),
           "alg_bytes_per_base": round(bytes_per_base, 4), "alg_gbs": round(gb * bytes_per_base, 1),
           "hbm_frac": round(gb * bytes_per_base / PEAK, 4)}
    out.update(kw)
    print(json.dumps(out), flush=True)


def timed(eng, fn, reps=3):
    fn()
    best = 1e30
    for _ in range(reps):
        fn()
        best = min(best, eng.last_times()["kernel_ms"])
    return best


def main():
    import torch
    eng = kb.Engine(0)
    dev = torch.device("cuda", 0)
    # ---- C2 reads: extraction / ntHash (materialising kernels) ----
    nb = workloads.c2_lengths()[:120000]
    batch = eng.batch_synth(2, nb)
    bases = int(nb.sum())
    nk8 = batch.kmer_count(8)
    out32 = torch.empty(nk8, dtype=torch.int32, device=dev)
    ms = timed(eng, lambda: kb.check(eng.lib.kmu_generate_kmers(eng.ctx, batch.handle, 8, kb.KMER32, kb.HASH_CANON_INVHASH,
                                                               out32.data_ptr(), None, 1)) if False else
               kb._lib.check(eng.lib.kmu_generate_kmers(eng.ctx, batch.handle, 8, kb.KMER32, kb.HASH_CANON_INVHASH,
                                                        out32.data_ptr(), None, 1)))
    line("generate_kmers k=8 Kmer32bit canon+int32_hash", bases, ms, 0.25 + 4.0 * nk8 / bases)
    nk31 = batch.kmer_count(31)
    out64 = torch.empty(nk31, dtype=torch.int64, device=dev)
    ms = timed(eng, lambda: kb._lib.check(eng.lib.kmu_generate_kmers(eng.ctx, batch.handle, 31, kb.KMER64, kb.HASH_CANON_RAW,
                                                                     out64.data_ptr(), None, 1)))
    line("generate_kmers k=31 Kmer64bit canonical", bases, ms, 0.25 + 8.0 * nk31 / bases)
    ms = timed(eng, lambda: kb._lib.check(eng.lib.kmu_nthash_canonical(eng.ctx, batch.handle, 31, 1, out64.data_ptr(), None, 1)))
    line("nthash_canonical k=31", bases, ms, 0.25 + 8.0 * nk31 / bases)
    del out32, out64
    if "--extract-only" in sys.argv:
        return
    # ---- C2 per-read sketches: SuperMinHash / SetSketch beside ProbMinHash3a ----
    sig = torch.empty((len(nb), 200), dtype=torch.float64, device=dev)
    ms = timed(eng, lambda: eng.sketch_superminhash(batch, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 200, out_device_ptr=sig.data_ptr()))
    line("superminhash per read k=8 m=200 f64", bases, ms, 0.25 + 1600.0 * len(nb) / bases)
    sig32 = torch.empty((len(nb), 200), dtype=torch.int32, device=dev)
    ms = timed(eng, lambda: eng.sketch_pmh3a(batch, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 200, out_device_ptr=sig32.data_ptr()))
    line("probminhash3a per read k=8 m=200", bases, ms, 0.25 + 800.0 * len(nb) / bases)
    hll = torch.empty((len(nb), 256), dtype=torch.int16, device=dev)
    ms = timed(eng, lambda: eng.sketch_setsketch(batch, 8, kb.KMER32, kb.HASH_CANON_INVHASH, (1.001, 256, 20.0, 65534), np.uint16,
                                                 out_device_ptr=hll.data_ptr()))
    line("setsketch per read k=8 m=256 u16", bases, ms, 0.25 + 512.0 * len(nb) / bases)
    del hll
    batch.destroy()
    del sig, sig32
    # ---- C3: counting, 150-base reads from a 100 Mb genome, k = 31 ----
    genome = eng.batch_synth(3, np.array([100_000_000], dtype=np.uint64))
    nreads = 8_000_000
    reads = eng.batch_sample_reads(genome, 3, 0, nreads, 150, 5000)
    rb = nreads * 150
    ctr = eng.counter(31, kb.KMER64, capacity=int(nreads * 120 * 0.45), count_bits=8)
    t0 = time.perf_counter()
    ctr.insert_seqs(reads, canonical=True)
    ms = eng.last_times()["kernel_ms"]
    st = ctr.stats()
    line("count_insert_seqs k=31 (cold table)", rb, ms, 0.25 + 32.0 * 120 / 150, nb_distinct=st["nb_distinct"],
         nb_unique=st["nb_unique"], table_slots=ctr.capacity(), wall_ms=round((time.perf_counter() - t0) * 1e3, 1))
    ctr.insert_seqs(reads, canonical=True)
    line("count_insert_seqs k=31 (all keys present)", rb, eng.last_times()["kernel_ms"], 0.25 + 32.0 * 120 / 150)
    send = torch.empty(reads.kmer_count(31), dtype=torch.int64, device=dev)
    ms = timed(eng, lambda: eng.count_partition(reads, 31, kb.KMER64, 8, True, out_device_ptr=send.data_ptr()))
    line("count_partition k=31 8 owners", rb, ms, 0.25 + 8.0 * 120 / 150)
    ctr.destroy()
    reads.destroy()
    del send
    # ---- C4: genomes of 5 Mb, k = 16 Kmer16b32bit, m = 12000 ----
    ng = 32
    gnb = np.full(ng, 5_000_000, dtype=np.uint64)
    gb_ = eng.batch_synth(4, gnb)
    gbases = int(gnb.sum())
    sig = torch.empty((ng, 12000), dtype=torch.float64, device=dev)
    ms = timed(eng, lambda: eng.sketch_superminhash(gb_, 16, kb.KMER16B32, kb.HASH_CANON_INVHASH, 12000, out_device_ptr=sig.data_ptr()), 2)
    line("superminhash per genome k=16 m=12000 f64", gbases, ms, 0.25 + 96000.0 / 5e6)
    sig32 = torch.empty((ng, 12000), dtype=torch.int32, device=dev)
    ms = timed(eng, lambda: eng.sketch_pmh3a(gb_, 16, kb.KMER16B32, kb.HASH_CANON_INVHASH, 12000, out_device_ptr=sig32.data_ptr()), 1)
    line("probminhash3a per genome k=16 m=12000 (per-sequence entry point, routed to the whole-file procedure)", gbases, ms, 0.25 + 48000.0 / 5e6)
    os.environ["KMU_PMH3A_TEAM_ONLY"] = "1"
    ms = timed(eng, lambda: eng.sketch_pmh3a(gb_, 16, kb.KMER16B32, kb.HASH_CANON_INVHASH, 12000, out_device_ptr=sig32.data_ptr()), 1)
    del os.environ["KMU_PMH3A_TEAM_ONLY"]
    line("probminhash3a per genome (team kernel, 32 of 148 SMs busy) k=16 m=12000", gbases, ms, 0.25 + 48000.0 / 5e6)
    one = eng.batch_synth(4, gnb[:1])
    ms = timed(eng, lambda: eng.sketch_pmh3a_whole(one, 16, kb.KMER16B32, kb.HASH_CANON_INVHASH, 12000), 2)
    line("probminhash3a whole-file (table + item kernel), one 5 Mb genome", 5_000_000, ms, 0.25 + 48000.0 / 5e6)
    ms = timed(eng, lambda: eng.sketch_pmh3a_groups(gb_, np.ones(ng, dtype=np.uint64), 16, kb.KMER16B32, kb.HASH_CANON_INVHASH, 12000), 2)
    line("probminhash3a whole-file, 32 genomes of 5 Mb in one call (kmu_sketch_pmh3a_groups)", gbases, ms, 0.25 + 48000.0 / 5e6)
    hll = torch.empty((ng, 4096), dtype=torch.int16, device=dev)
    ms = timed(eng, lambda: eng.sketch_setsketch(gb_, 16, kb.KMER16B32, kb.HASH_CANON_INVHASH, None, np.uint16, out_device_ptr=hll.data_ptr()), 2)
    line("setsketch per genome k=16 m=4096 u16", gbases, ms, 0.25 + 8192.0 / 5e6)
    gb_.destroy()
    one.destroy()
    del sig, sig32, hll
    # ---- C5a: whole-file SetSketch over long sequences, k = 21 ----
    cnb = np.linspace(50e6, 200e6, 8).astype(np.uint64)
    cb = eng.batch_synth(5, cnb)
    ms = timed(eng, lambda: eng.sketch_setsketch(cb, 21, kb.KMER64, kb.HASH_CANON_INVHASH, None, np.uint16, whole=True), 2)
    line("setsketch whole-file k=21 m=4096 (1 Gbase)", int(cnb.sum()), ms, 0.25)
    ms = timed(eng, lambda: eng.sketch_superminhash_whole(cb, 21, kb.KMER64, kb.HASH_CANON_INVHASH, 4096), 2)
    line("superminhash whole-file k=21 m=4096 f64 (1 Gbase)", int(cnb.sum()), ms, 0.25)
    os.environ["KMU_SMH_NO_CUT"] = "1"
    ms = timed(eng, lambda: eng.sketch_superminhash_whole(cb, 21, kb.KMER64, kb.HASH_CANON_INVHASH, 4096), 2)
    del os.environ["KMU_SMH_NO_CUT"]
    line("superminhash whole-file k=21 m=4096 f64 (1 Gbase), per-thread-chunk kernel without the value cut", int(cnb.sum()), ms, 0.25)
    cb.destroy()
    # ---- C5b: proteome, AA k = 12, ProbMinHash3a m = 400 ----
    rng = np.random.default_rng(5)
    pl = np.clip(np.rint(np.exp(rng.normal(5.6, 0.6, 200000))), 50, 5000).astype(np.uint64)
    pb = eng.batch_synth_aa(5, pl)
    psig = torch.empty((len(pl), 400), dtype=torch.int64, device=dev)
    ms = timed(eng, lambda: eng.sketch_pmh3a(pb, 12, kb.KMERAA64, kb.HASH_MASKED_VALUE, 400, out_device_ptr=psig.data_ptr()), 2)
    line("probminhash3a proteome AA k=12 m=400 (residues)", int(pl.sum()), ms, 1.0 + 3200.0 * len(pl) / int(pl.sum()))
    # one SetSketch for the whole proteome (sketch_compressedkmeraa_seqs: what a proteome comparison uses), default parameters
    ms = timed(eng, lambda: eng.sketch_setsketch(pb, 12, kb.KMERAA64, kb.HASH_MASKED_VALUE, None, np.uint16, whole=True), 2)
    line("setsketch whole proteome AA k=12 m=4096 u16 (residues)", int(pl.sum()), ms, 1.0)
    pb.destroy()
    # SetSketch per protein, default parameters (m = 4096 registers for ~270 12-mers: every item places every point,
    # the exact path; 2 000 proteins)
    pl2 = pl[:2000]
    pb2 = eng.batch_synth_aa(5, pl2)
    phll = torch.empty((len(pl2), 4096), dtype=torch.int16, device=dev)
    ms = timed(eng, lambda: eng.sketch_setsketch(pb2, 12, kb.KMERAA64, kb.HASH_MASKED_VALUE, None, np.uint16, out_device_ptr=phll.data_ptr()), 1)
    line("setsketch proteome AA k=12 m=4096 u16 (residues)", int(pl2.sum()), ms, 1.0 + 8192.0 * len(pl2) / int(pl2.sum()))
    pb2.destroy()
    eng.close()


if __name__ == "__main__":
    main()
