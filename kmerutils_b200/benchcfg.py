"""The BASELINE.json shapes beside the headline one, as bench.py's `configs` array: each runner
returns one dict {workload, value, unit, ms_per_step, n_gpus, scaling, roofline{...}, ...}.

Timing: CUDA events on the engine's stream around the timed steps (max over ranks); inputs smaller than L2 are timed
step by step with an L2 flush (a 512 MB fill) between steps.  Algorithmic bytes per base are those of SURVEY.md 8(d).
Only the steps of the path that really exchange data run a collective: the counting exchange (one fused kernel that
stores into the peers over NVLink, or the NCCL all-to-all) and the register merge of the whole-file sketches
(NCCL allreduce max / min) -- both INSIDE the timed region.
"""
import math
import os
import time

import numpy as np

from . import dist as kd
from . import workloads


class Timer:
    def __init__(self, eng, local_rank, world):
        import torch
        self.torch = torch
        self.eng = eng
        self.dev = torch.device("cuda", local_rank)
        self.world = world
        self.ext = torch.cuda.ExternalStream(eng.stream(), device=self.dev)
        self._flush = None

    def barrier(self):
        self.torch.cuda.synchronize(self.dev)
        self.eng.sync()
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()

    def max_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        import torch.distributed as dist
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def flush_l2(self):
        if self._flush is None:
            self._flush = self.torch.empty(512 << 20, dtype=self.torch.uint8, device=self.dev)
        self._flush.fill_(1)
        self.torch.cuda.synchronize(self.dev)

    def run(self, fn, steps, warmup, flush=False):
        """-> ms per step (max over ranks).  fn() enqueues (or runs) one step; host work between the kernels of a step
        is inside the timed region."""
        torch = self.torch
        for _ in range(max(warmup, 0)):
            fn()
        self.barrier()
        if not flush:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(self.ext)
            for _ in range(steps):
                fn()
            ev1.record(self.ext)
            self.barrier()
            ms = ev0.elapsed_time(ev1) / steps
        else:
            tot = 0.0
            for _ in range(steps):
                self.flush_l2()
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record(self.ext)
                fn()
                ev1.record(self.ext)
                self.barrier()
                tot += ev0.elapsed_time(ev1)
            ms = tot / steps
        return self.max_over_ranks(ms)


def _entry(workload, bases_job, ms, bytes_per_base, peak, n_gpus, scaling, kernel, l2, **kw):
    gb = bases_job / (ms * 1e-3) / 1e9
    alg = bases_job * bytes_per_base
    out = {"workload": workload, "value": gb, "unit": "Gbases/s", "ms_per_step": ms, "n_gpus": n_gpus, "scaling": scaling,
           "bases_per_step": int(bases_job), "l2": l2,
           "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9 / n_gpus, "peak": peak, "unit": "GB/s per GPU",
                        "frac": alg / (ms * 1e-3) / 1e9 / n_gpus / peak, "algorithmic_bytes": alg / n_gpus,
                        "algorithmic_bytes_per_base": bytes_per_base, "kernel": kernel, "traffic": None}}
    out.update(kw)
    return out


def run_c1(kb, eng, T, rank, world, steps, warmup, peak):
    """C1: 1000 reads x 1000 b, k = 8 Kmer32bit, ProbMinHash3a m = 200 (the reference's own CPU-runnable case).  Every rank
    sketches the same full C1 (it is 250 KB of packed bases): replicas, value = one replica's rate."""
    torch = T.torch
    nb = workloads.c1_lengths()
    batch = eng.batch_synth(1, nb)
    sig = torch.empty((len(nb), 200), dtype=torch.int32, device=T.dev)
    ms = T.run(lambda: eng.sketch_pmh3a(batch, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 200, out_device_ptr=sig.data_ptr()), steps, warmup,
               flush=True)
    bases = int(nb.sum())
    batch.destroy()
    return _entry("C1: 1000 reads x 1000 b, k=8 Kmer32bit canonical+int32_hash, ProbMinHash3a m=200 (full config)", bases, ms,
                  0.25 + 800.0 / 1000.0, peak, 1, "replicas", "pmh3a_direct_kernel / pmh3a_sketch_kernel",
                  "input 250 KB: L2 flushed between timed steps", launch_bound=True)


def run_extract(kb, eng, T, rank, world, steps, warmup, peak, batch, bases):
    """K2 / K3 alone on the C2 reads of this rank: generate_kmers k = 8 (u32) and k = 31 (u64), canonical ntHash k = 31."""
    torch = T.torch
    out = []
    nk8, nk31 = batch.kmer_count(8), batch.kmer_count(31)
    buf = torch.empty(max(nk8 * 4, nk31 * 8), dtype=torch.uint8, device=T.dev)
    lib, ctx = eng.lib, eng.ctx
    cases = [
        ("extract k=8: KmerGenerator<Kmer32bit> + canonical + int32_hash, all k-mers materialised (u32)",
         lambda: kb._lib.check(lib.kmu_generate_kmers(ctx, batch.handle, 8, kb.KMER32, kb.HASH_CANON_INVHASH, buf.data_ptr(), None, 1)),
         0.25 + 4.0 * nk8 / bases, "generate_kmers_run_kernel<u32>"),
        ("extract k=31: KmerGenerator<Kmer64bit> + canonical, all k-mers materialised (u64)",
         lambda: kb._lib.check(lib.kmu_generate_kmers(ctx, batch.handle, 31, kb.KMER64, kb.HASH_CANON_RAW, buf.data_ptr(), None, 1)),
         0.25 + 8.0 * nk31 / bases, "generate_kmers_run_kernel<u64>"),
        ("ntHash k=31: canonical hash of every k-mer (u64)",
         lambda: kb._lib.check(lib.kmu_nthash_canonical(ctx, batch.handle, 31, 1, buf.data_ptr(), None, 1)),
         0.25 + 8.0 * nk31 / bases, "nthash_run_kernel"),
    ]
    for name, fn, bpb, kernel in cases:
        # the first call on a batch for a given k also builds and uploads the per-sequence output offsets (host loop over the
        # lengths + one copy; kept with the batch afterwards, like the processing order of the sketch kernels): timed apart
        T.barrier()
        t0 = time.perf_counter()
        fn()
        T.barrier()
        first_ms = (time.perf_counter() - t0) * 1e3
        ms = T.run(fn, steps, warmup)
        out.append(_entry(name + " on the C2 reads", bases * world, ms, bpb, peak, world, "weak", kernel,
                          "inputs (1.1 GB packed) and outputs (17-35 GB) larger than L2",
                          first_call_ms=T.max_over_ranks(first_ms),
                          note="every step recomputes and rewrites all k-mers; only the output offsets of the batch (metadata) are kept between calls"))
    del buf
    return out


def run_c2_strong(kb, eng, T, rank, world, steps, warmup, peak):
    """C2 as BASELINE.json words it: the 746 333 reads sharded over the GPUs (contiguous ranges balanced by bases), no collective."""
    torch = T.torch
    nbases = workloads.c2_lengths()
    lo, hi = kd.shard_by_bases(nbases, world)[rank]
    # the synthetic stream is indexed by base position: a shard is generated as its own batch from its own seed
    batch = eng.batch_synth(1000 + rank, nbases[lo:hi])
    sig = torch.empty((hi - lo, 200), dtype=torch.int32, device=T.dev)
    ms = T.run(lambda: eng.sketch_pmh3a(batch, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 200, out_device_ptr=sig.data_ptr()), steps, warmup)
    total = int(nbases.sum())
    batch.destroy()
    return _entry("C2 strong scaling: the 746333 reads / 4.38 Gbases sharded over the GPUs, k=8, ProbMinHash3a m=200", total, ms,
                  0.25 + 800.0 * len(nbases) / total, peak, world, "strong", "pmh3a_direct_kernel",
                  "inputs larger than L2", reads_this_rank=hi - lo)


def run_c3(kb, eng, T, rank, world, steps, warmup, peak, reads_per_step=26_666_667, exchange="fused"):
    """C3: 150-base reads drawn from a 100 Mb genome (random strand, 0.5 % substitutions), k = 31 canonical Kmer64bit,
    exact multiplicities.  Weak scaling: every rank counts reads_per_step reads (4 Gbases) per step; at N > 1 the
    k-mers go to their owner (intNN_hash % N, DispatchableT) INSIDE the timed region -- `fused`: one kernel buckets and
    stores into the peers' buffers over NVLink; `nccl`: partition + NCCL all-to-all."""
    import torch.distributed as dist
    torch = T.torch
    k, read_len = 31, 150
    # three rounds of 3.2 G k-mers (one warm-up, two timed): ~1.4 G distinct keys, a table of 2^32 slots (64 GB); a fourth
    # round would push the requested capacity past 2^31 keys and double the table (and the sweep every step pays)
    steps = max(1, min(steps, 2))
    warmup = max(1, min(warmup, 1))
    nrounds = steps + warmup
    genome = eng.batch_synth(3, np.array([100_000_000], dtype=np.uint64))
    nk_round = reads_per_step * (read_len - k + 1)
    # distinct keys a rank ends up owning: its share of the genome's k-mers + of the error k-mers of all ranks
    expect = 100e6 / world + nrounds * nk_round * 0.17
    counter = eng.counter(k, kb.KMER64, capacity=int(expect * 0.98), count_bits=8)
    batches = [eng.batch_sample_reads(genome, 3, (r * world + rank) * reads_per_step, reads_per_step, read_len, 5000)
               for r in range(nrounds)]
    xchg = kd.P2PExchange(eng) if (world > 1 and exchange == "fused") else None
    send = torch.empty(nk_round, dtype=torch.int64, device=T.dev) if (world > 1 and exchange == "nccl") else None
    phases = {"scatter_ms": 0.0, "share_counts_ms": 0.0, "insert_ms": 0.0}
    sent_bytes = [0]
    it = [0]

    def step():
        if it[0] == warmup:  # the phases of the timed rounds only (the warm-up round allocates the slabs and maps the peers)
            for key in phases:
                phases[key] = 0.0
            sent_bytes[0] = 0
        t_step = time.perf_counter()
        reads = batches[it[0]]
        it[0] += 1
        if world == 1:
            counter.insert_seqs(reads, canonical=True)
            phases["insert_ms"] += eng.last_times()["kernel_ms"]
        elif exchange == "fused":
            _, sent = kd.count_round_fused(eng, reads, counter, xchg, nk_round, phases=phases)
            sent_bytes[0] += sent
        else:
            _, counts = eng.count_partition(reads, k, kb.KMER64, world, True, out_device_ptr=send.data_ptr())
            phases["scatter_ms"] += eng.last_times()["kernel_ms"]
            t0 = time.perf_counter()
            recv, _ = kd.exchange_kmers(send[:nk_round], counts)
            torch.cuda.synchronize(T.dev)
            phases["share_counts_ms"] += (time.perf_counter() - t0) * 1e3
            counter.insert_kmers(device_ptr=recv.data_ptr(), n=recv.numel())
            phases["insert_ms"] += eng.last_times()["kernel_ms"]
            sent_bytes[0] += 8 * int(sum(int(c) for i, c in enumerate(counts) if i != rank))
        phases["step_wall_ms"] = phases.get("step_wall_ms", 0.0) + (time.perf_counter() - t_step) * 1e3

    def timed_step():
        step()

    # every rank samples its own GPU while the rounds run (the ranks do the same work: is a spread of their times the GPUs'?)
    import threading
    samples = {"sm": [], "mem": [], "w": [], "reasons": 0, "t_gpu": [], "t_mem": []}
    stop = threading.Event()

    def sampler():
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(int(os.environ.get("LOCAL_RANK", 0)))
            while not stop.is_set():
                if it[0] > warmup:
                    samples["sm"].append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                    samples["mem"].append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM))
                    samples["w"].append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
                    samples["reasons"] |= int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                    try:  # temperatures: HBM refreshes more often when hot, which random accesses feel first
                        samples["t_gpu"].append(pynvml.nvmlDeviceGetTemperature(h, pynvml.NVML_TEMPERATURE_GPU))
                        fv = pynvml.nvmlDeviceGetFieldValues(h, [pynvml.NVML_FI_DEV_MEMORY_TEMP])[0]
                        if fv.nvmlReturn == 0:
                            samples["t_mem"].append(int(fv.value.uiVal))
                    except Exception:
                        pass
                stop.wait(0.005)
        except Exception:
            pass

    th = threading.Thread(target=sampler, daemon=True) if world > 1 else None
    if th:
        th.start()
    # warm-up rounds (allocations, IPC mapping), then the timed ones
    ms = T.run(timed_step, steps, warmup)
    stop.set()
    if th:
        th.join(timeout=1.0)
    # every rank's insertion time and SM clock right after the timed rounds (the ranks do the same work: a spread is the GPUs')
    by_rank = None
    if world > 1:
        mhz = -1.0
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(int(os.environ.get("LOCAL_RANK", 0)))
            mhz = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            pass
        med = lambda v: float(np.median(v)) if v else -1.0  # noqa: E731
        mine = torch.tensor([phases["insert_ms"] / steps, phases["scatter_ms"] / steps, mhz, med(samples["sm"]), min(samples["sm"] or [-1]),
                             med(samples["mem"]), med(samples["w"]), max(samples["w"] or [-1]), float(samples["reasons"]),
                             float(len(samples["sm"])), max(samples["t_gpu"] or [-1]), max(samples["t_mem"] or [-1])],
                            dtype=torch.float64, device=T.dev)
        allr = torch.empty((world, mine.numel()), dtype=torch.float64, device=T.dev)
        dist.all_gather_into_tensor(allr, mine)
        allr = allr.cpu().numpy()
        by_rank = {"insert_ms": [round(float(x), 1) for x in allr[:, 0]], "scatter_ms": [round(float(x), 1) for x in allr[:, 1]],
                   "sm_mhz_after": [float(x) for x in allr[:, 2]],
                   "during_the_rounds": {"sm_mhz_median": [float(x) for x in allr[:, 3]], "sm_mhz_min": [float(x) for x in allr[:, 4]],
                                         "mem_mhz_median": [float(x) for x in allr[:, 5]], "power_w_median": [round(float(x)) for x in allr[:, 6]],
                                         "power_w_max": [round(float(x)) for x in allr[:, 7]],
                                         "clock_event_reasons_or": [int(x) for x in allr[:, 8]], "samples": [int(x) for x in allr[:, 9]],
                                         "gpu_temp_c_max": [int(x) for x in allr[:, 10]], "hbm_temp_c_max": [int(x) for x in allr[:, 11]]}}
    # per step, slowest and fastest rank of every phase (the step itself is the max over ranks, barrier waits included)
    phases_min = {}
    for key in sorted(phases):
        phases[key] = phases[key] / steps
        phases_min[key] = -T.max_over_ranks(-phases[key])
        phases[key] = T.max_over_ranks(phases[key])
    st = counter.stats()
    tot = kd.allreduce_sum([st["nb_distinct"], st["nb_unique"], st["nb_inserted"]], T.dev)
    ok = tot[2] == nrounds * nk_round * world
    bases = reads_per_step * read_len * world
    for b_ in batches:
        b_.destroy()
    genome.destroy()
    table_slots = counter.capacity()
    counter.destroy()
    if xchg:
        xchg.close()
    lim = max((q for q in phases if q not in ('step_wall_ms', 'end_barrier_ms')), key=lambda q: phases[q])
    nk_frac = (read_len - k + 1) / read_len
    return _entry("C3: 150 b reads from a 100 Mb genome (0.5 % substitutions), k=31 canonical Kmer64bit exact counting, "
                  f"{reads_per_step} reads / {reads_per_step * read_len / 1e9:.2f} Gbases per GPU per step", bases, ms,
                  0.25 + 32.0 * nk_frac, peak, world, "weak",
                  "count_part_kernel (partition by owner x table region in shared memory) + count_insert_slabs_kernel "
                  "(regioned insertion, updates hit L2)", "table (64 GB+) and inputs far larger than L2",
                  exchange=("none" if world == 1 else exchange), collective_in_timed_region=world > 1,
                  phases_ms_per_step=phases, phases_ms_per_step_fastest_rank=phases_min, by_rank=by_rank, limiting_phase=lim,
                  nvlink_bytes_sent_per_gpu_per_step=int(sent_bytes[0] // max(steps, 1)),
                  nb_distinct=tot[0], nb_unique=tot[1], nb_inserted=tot[2], conservation_ok=bool(ok), table_slots_per_gpu=table_slots,
                  steps=steps, warmup=warmup)


def run_c4(kb, eng, T, rank, world, steps, warmup, peak, ngenomes=148):
    """C4 (gsearch shape): genomes of 5 Mb, k = 16 Kmer16b32bit, canonical + int32_hash; ProbMinHash3a m = 12 000 (whole-genome
    signature, kmu_sketch_pmh3a_groups) and SuperMinHash m = 12 000 f64.  Genomes shard over the GPUs, no collective:
    every rank sketches ngenomes genomes (one per SM)."""
    torch = T.torch
    steps = max(1, min(steps, 3))
    warmup = max(1, min(warmup, 1))
    gnb = np.full(ngenomes, 5_000_000, dtype=np.uint64)
    batch = eng.batch_synth(4 + 1000 * rank, gnb)
    bases = int(gnb.sum())
    out = []
    groups = np.ones(ngenomes, dtype=np.uint64)
    ms = T.run(lambda: eng.sketch_pmh3a_groups(batch, groups, 16, kb.KMER16B32, kb.HASH_CANON_INVHASH, 12000), steps, warmup)
    out.append(_entry(f"C4 ProbMinHash3a: {ngenomes} genomes x 5 Mb per GPU, k=16 Kmer16b32bit, m=12000 (u32), one call", bases * world, ms,
                      0.25 + 48000.0 / 5e6, peak, world, "weak", "count_insert_seqs_kernel<u32> (per-genome table in L2) + pmh3a_items_kernel",
                      "inputs (185 MB packed per GPU) larger than L2", includes="7 MB of signatures copied to the host per step"))
    sig = torch.empty((ngenomes, 12000), dtype=torch.float64, device=T.dev)
    ms = T.run(lambda: eng.sketch_superminhash(batch, 16, kb.KMER16B32, kb.HASH_CANON_INVHASH, 12000, out_device_ptr=sig.data_ptr()),
               steps, warmup)
    out.append(_entry(f"C4 SuperMinHash: {ngenomes} genomes x 5 Mb per GPU, k=16, m=12000 f64, NoHashHasher", bases * world, ms,
                      0.25 + 96000.0 / 5e6, peak, world, "weak", "smh_fast_kernel", "inputs larger than L2"))
    batch.destroy()
    return out


def c5a_lengths():
    nb = np.linspace(50e6, 200e6, 24)
    nb = np.rint(nb * (3.0e9 / nb.sum())).astype(np.uint64)
    return nb


def run_c5a(kb, eng, T, rank, world, steps, warmup, peak):
    """C5a: 24 sequences of 50-200 Mb (3.0 Gbases), k = 21 Kmer64bit canonical + int64_hash, ONE SetSketch (m = 4096 u16, default
    parameters) for the set, and ONE SuperMinHash (m = 4096 f64).  Strong scaling: every sequence is cut into N chunks with a
    k - 1 halo, rank r sketches chunk r of every sequence, the registers merge with an NCCL allreduce max / min inside the
    timed region (SetSketcher::merge, setsketchert.rs:876-882)."""
    torch = T.torch
    steps = max(1, min(steps, 3))
    warmup = max(1, min(warmup, 1))
    k = 21
    nb = c5a_lengths()
    full = eng.batch_synth(5, nb)
    if world > 1:
        idx = np.arange(len(nb), dtype=np.uint64)
        begin = (nb * np.uint64(rank)) // np.uint64(world)
        end = np.minimum(nb, (nb * np.uint64(rank + 1)) // np.uint64(world) + np.uint64(k - 1))
        mine = eng.batch_slices(full, idx, begin, end)
        full.destroy()
    else:
        mine = full
    total = int(nb.sum())
    out = []
    regs = torch.empty(4096, dtype=torch.int16, device=T.dev)

    def setsketch_step():
        eng.sketch_setsketch(mine, k, kb.KMER64, kb.HASH_CANON_INVHASH, None, np.uint16, whole=True, out_device_ptr=regs.data_ptr())
        if world > 1:
            import torch.distributed as dist
            eng.sync()
            wide = regs.to(torch.int32) & 0xFFFF  # NCCL has no u16: widen (SetSketcher::merge is an element-wise max)
            dist.all_reduce(wide, op=dist.ReduceOp.MAX)
            regs.copy_(wide.to(torch.int16))

    ms = T.run(setsketch_step, steps, warmup)
    out.append(_entry("C5a SetSketch: 24 sequences of 50-200 Mb (3.0 Gbases), k=21 Kmer64bit, whole-set registers m=4096 u16"
                      + (", chunks with k-1 halo over the GPUs + NCCL allreduce-max" if world > 1 else ""), total, ms, 0.25, peak,
                      world, "strong", "ssk_whole_warp_kernel", "inputs (750 MB packed) larger than L2",
                      collective_in_timed_region=world > 1, collective="allreduce max of 4096 registers (widened to i32)" if world > 1 else None))
    smh = torch.empty(4096, dtype=torch.float64, device=T.dev)
    hsig = np.zeros(4096, dtype=np.float64)

    def smh_step():
        h = eng.sketch_superminhash_whole(mine, k, kb.KMER64, kb.HASH_CANON_INVHASH, 4096)
        if world > 1:
            smh.copy_(torch.from_numpy(h))
            kd.merge_registers(smh, "min")

    ms = T.run(smh_step, steps, warmup)
    out.append(_entry("C5a SuperMinHash: the same 3.0 Gbases, k=21, whole-set signature m=4096 f64"
                      + (", chunks with k-1 halo + NCCL allreduce-min" if world > 1 else ""), total, ms, 0.25, peak, world, "strong",
                      "smh_whole_warp_kernel", "inputs larger than L2", collective_in_timed_region=world > 1))
    mine.destroy()
    del hsig
    return out


def c5b_lengths(n=20000):
    rng = np.random.default_rng(5)
    return np.clip(np.rint(np.exp(rng.normal(5.6, 0.6, n))), 50, 5000).astype(np.uint64)


def run_c5b(kb, eng, T, rank, world, steps, warmup, peak):
    """C5b: proteome of 20 000 proteins, amino-acid k = 12 KmerAA64bit (5 bits per residue), MASKED_VALUE closure: ProbMinHash3a
    m = 400 per protein and ONE SetSketch (m = 4096) for the proteome.  Proteins shard over the GPUs (every rank its own
    proteome), no collective."""
    torch = T.torch
    pl = c5b_lengths()
    batch = eng.batch_synth_aa(5 + 1000 * rank, pl)
    res = int(pl.sum())
    out = []
    sig = torch.empty((len(pl), 400), dtype=torch.int64, device=T.dev)
    ms = T.run(lambda: eng.sketch_pmh3a(batch, 12, kb.KMERAA64, kb.HASH_MASKED_VALUE, 400, out_device_ptr=sig.data_ptr()), steps, warmup,
               flush=True)
    e = _entry("C5b ProbMinHash3a: proteome of 20000 proteins per GPU, AA k=12 KmerAA64bit, m=400 per protein", res * world, ms,
               1.0 + 3200.0 * len(pl) / res, peak, world, "weak", "pmh3a_sketch_kernel<u64, table>",
               "input 7 MB: L2 flushed between timed steps")
    e["unit"] = "Gresidues/s"
    out.append(e)
    regs = torch.empty(4096, dtype=torch.int16, device=T.dev)
    ms = T.run(lambda: eng.sketch_setsketch(batch, 12, kb.KMERAA64, kb.HASH_MASKED_VALUE, None, np.uint16, whole=True,
                                            out_device_ptr=regs.data_ptr()), steps, warmup, flush=True)
    e = _entry("C5b SetSketch: one sketch for the proteome (sketch_compressedkmeraa_seqs), AA k=12, m=4096 u16", res * world, ms, 1.0,
               peak, world, "weak", "ssk_whole_kernel", "input 7 MB: L2 flushed between timed steps")
    e["unit"] = "Gresidues/s"
    out.append(e)
    batch.destroy()
    return out
