#!/usr/bin/env python
"""bench.py -- headline benchmark of the kmerutils hot path on B200.

Workload (BASELINE.json configs[1], the README benchmark shape): 746 333 ONT-like synthetic reads,
4.38 Gbases, k = 8 (Kmer32bit), canonical + int32_hash, ProbMinHash3a with 200 slots per read.
One "step" = one pass of extraction + counting + sketching over the whole batch.

  python bench.py --gpus N --steps K --warmup W          (N > 1: launched under torchrun)
  python bench.py --impl reference ...                   CPU arm: the oracle port on all host cores

Prints ONE JSON line (rank 0).  `value` is device-timed with the packed reads resident in HBM and
the signatures left in HBM; `e2e` goes through the host-buffer C-ABI call (pinned host buffers,
H2D + kernels + D2H inside the timed region).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from kmerutils_b200 import workloads  # noqa: E402

K = 8
M = 200
KMER_TYPE = 0      # Kmer32bit
HASH_KIND = 2      # canonical + int32_hash (datasketcher.rs:222-226)
WORKLOAD = "C2: 746333 ONT-like synthetic reads / 4.38 Gbases, k=8 Kmer32bit, canonical+int32_hash, ProbMinHash3a 200 slots/read"


def measured_peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def bind_to_gpu_numa_node(local_rank):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned memory is allocated: pinned pages
    are then first-touched on that node and the H2D copies do not cross the socket interconnect.  -> description (or why not)"""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0" % (dom, bus, dev)
        node = int(open(path + "/numa_node").read().strip())
        if node < 0:
            return {"numa_node": node, "bound": False, "why": "the platform reports no NUMA node for the GPU"}
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"numa_node": node, "bound": False, "why": "no allowed CPU on that node"}
        os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "bound": True, "cpus": len(cpus)}
    except Exception as exc:  # containers without /sys: run unbound
        return {"bound": False, "why": f"{type(exc).__name__}: {exc}"}


def cpu_sample_rate(orc, nbases, seed, target_s, nthreads):
    """Time the oracle port (one task per read over nthreads, like rayon in the reference) on a
    bounded prefix of the workload sized for about target_s seconds."""
    from test_pmh3a_gpu import oracle_batch

    def run(n_reads):
        nb = nbases[:n_reads]
        packed, off = oracle_batch(orc, seed, nb)
        t0 = time.perf_counter()
        orc.sketch_pmh3a_batch(packed, off, nb, K, KMER_TYPE, HASH_KIND, M, nthreads)
        return time.perf_counter() - t0, int(nb.sum())

    n = min(len(nbases), 64 * nthreads)
    dt, bases = run(n)
    rate = bases / max(dt, 1e-6)
    want_bases = rate * target_s
    mean_len = float(nbases.mean())
    n2 = int(min(len(nbases), max(n, want_bases / mean_len)))
    dt, bases = run(n2)
    return bases / dt / 1e9, n2, bases, dt


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    from oracle_lib import get_oracle
    orc = get_oracle()
    nthreads = orc.hardware_threads()
    nbases = workloads.c2_lengths()
    from test_pmh3a_gpu import oracle_batch
    # per-step bounded sample: sized from a pilot so that a step takes ~3 s
    pilot_rate, _, _, _ = cpu_sample_rate(orc, nbases, 2, 1.0, nthreads)
    n_reads = int(min(len(nbases), max(64, pilot_rate * 1e9 * 3.0 / float(nbases.mean()))))
    nb = nbases[:n_reads]
    packed, off = oracle_batch(orc, 2, nb)
    bases = int(nb.sum())
    for _ in range(args.warmup):
        orc.sketch_pmh3a_batch(packed, off, nb, K, KMER_TYPE, HASH_KIND, M, nthreads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.sketch_pmh3a_batch(packed, off, nb, K, KMER_TYPE, HASH_KIND, M, nthreads)
    dt = (time.perf_counter() - t0) / args.steps
    val = bases / dt / 1e9
    sample = f"first {n_reads} reads of the workload ({bases} bases) per step"
    out = {
        "impl": "reference", "metric": "Gbases/s k-mer sketch (extract + count + ProbMinHash3a)", "value": val,
        "unit": "Gbases/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32+f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "k": K, "sketch_size": M},
        "cpu_baseline": {"value": val, "unit": "Gbases/s", "cores": nthreads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU oracle port of the Rust reference (no Rust toolchain in this image); README.md:45 publishes "
                "0.0859 Gbases/s for the same shape on an 8-core laptop",
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--reads", type=int, default=0, help="debug: use only the first N reads of the workload")
    ap.add_argument("--profile", action="store_true", help="print the per-launch profile to stderr")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--configs", default="all",
                    help="the other BASELINE shapes reported in `configs`: all | none | comma list of "
                         "extract,c2strong,c1,c3,c4,c5a,c5b")
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"],
                    help="C3 at N > 1: fused = one kernel buckets and stores into the peers over NVLink; nccl = partition + all-to-all")
    ap.add_argument("--c3-reads", type=int, default=26_666_667, help="C3: reads of 150 b per GPU per step (default 4 Gbases)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import kmerutils_b200 as kb

    assert args.warmup >= 0 and args.steps >= 1
    torch.cuda.set_device(local_rank)
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its version banner (and anything NCCL_DEBUG asks for) to stdout when the first communicator is made:
        # keep stdout for the one JSON line, send the library's chatter to stderr
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    eng = kb.Engine(local_rank)
    nbases = workloads.c2_lengths()
    if args.reads:
        nbases = nbases[: args.reads]
    nseq = len(nbases)
    total_bases = int(nbases.sum())
    # weak scaling: every rank sketches its own C2-sized batch (reads are independent units: no collective)
    seed = 2 + rank
    batch = eng.batch_synth(seed, nbases)
    sig_dev = torch.empty((nseq, M), dtype=torch.int32, device=f"cuda:{local_rank}")
    ext = torch.cuda.ExternalStream(eng.stream(), device=f"cuda:{local_rank}")

    def step():
        eng.sketch_pmh3a(batch, K, KMER_TYPE, HASH_KIND, M, out_device_ptr=sig_dev.data_ptr())

    def barrier():
        torch.cuda.synchronize()
        eng.sync()
        if world > 1:
            dist.barrier()

    for _ in range(max(args.warmup, 0)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    ev0.record(ext)
    for _ in range(args.steps):
        step()
        kernel_ms.append(eng.last_times()["kernel_ms"])
    ev1.record(ext)
    barrier()
    clocks = sampler.stop()
    launches = eng.launch_count() - launches0
    elapsed_ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = total_bases * world / (ms_per_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (pmh3a_sketch_kernel): per-launch events ------------
    eng.set_profiling(True)
    step()
    eng.sync()
    prof = eng.last_launch_profile()
    eng.set_profiling(False)
    peak, peak_kind = measured_peak_hbm()
    dom = max(prof, key=lambda r: r["ms"]) if prof else None
    roofline = None
    if dom:
        # algorithmic bytes of one launch: packed bases in (0.25 B/base) + signatures out (m * 4 B/read)
        alg_bytes = dom["nbases"] * 0.25 + dom["nseq"] * M * 4
        achieved = alg_bytes / (dom["ms"] * 1e-3) / 1e9
        # DRAM traffic is NOT measured in this run (ncu is not running): `traffic` stays null; the figure of the committed
        # ncu --set full capture of the same launch (same reads and bases) is quoted apart, labelled as such
        precaptured = None
        for name in ("r2_traffic.json", "r1_traffic.json"):
            try:
                with open(os.path.join(ROOT, "profiles", name)) as f:
                    tr = json.load(f)
                if tr["launch_reads"] == dom["nseq"] and tr["launch_bases"] == dom["nbases"]:
                    precaptured = {"dram_bytes_per_launch": tr["dram_bytes_per_launch"], "source": "profiles/" + name,
                                   "note": "ncu --set full capture committed with the repo, not measured in this run"}
                    if "issue" in tr:  # the kernel's own bound (instruction issue), from the same capture
                        precaptured["issue"] = tr["issue"]
                    break
            except Exception:
                pass
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": None, "traffic_precaptured": precaptured,
                    "algorithmic_bytes": alg_bytes, "peak_kind": peak_kind,
                    "kernel": ("pmh3a_direct_kernel (one pass, %d threads per read)" % dom["block"] if dom["mode"] == 2 else
                               "pmh3a_sketch_kernel<u32,%s> team_warps=%d" % ("hist" if dom["mode"] == 0 else "table", dom["team_warps"])),
                    "launch_ms": dom["ms"], "launch_bases": dom["nbases"], "launch_reads": dom["nseq"],
                    "share_of_step": dom["ms"] / max(sum(r["ms"] for r in prof), 1e-9),
                    "note": "latency/instruction-bound shared-memory histogram + per-occurrence offers (two shared atomics and one L2 memo lookup per k-mer), see DESIGN.md"}
    if args.profile:
        for r in prof:
            sys.stderr.write(json.dumps(r) + "\n")

    # ---- e2e: host buffers through the one-shot C-ABI call -------------------------------------
    e2e = None
    if not args.no_e2e:
        packed_host, off_host, _ = batch.download()
        pin_in = torch.empty(len(packed_host) + 64, dtype=torch.uint8).pin_memory()
        pin_in[: len(packed_host)] = torch.from_numpy(packed_host)
        pin_out = torch.empty((nseq, M), dtype=torch.int32).pin_memory()

        def e2e_step():
            eng.sketch_pmh3a_host((pin_in.data_ptr(), pin_in.numel()), off_host, nbases, K, KMER_TYPE, HASH_KIND, M,
                                  pin_out.data_ptr())

        def timed_calls(fn, n):
            """n calls between two barriers -> (seconds per call of this rank, max over ranks)"""
            barrier()
            t0 = time.perf_counter()
            for _ in range(n):
                fn()
            mine = (time.perf_counter() - t0) / n
            barrier()
            worst = mine
            if world > 1:
                t = torch.tensor([mine], dtype=torch.float64, device=f"cuda:{local_rank}")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                worst = float(t.item())
            return mine, worst

        e2e_step()
        n_e2e = max(1, args.steps)
        mine_s, e2e_s = timed_calls(e2e_step, n_e2e)
        tm = eng.last_times()
        e2e = {"value": total_bases * world / e2e_s / 1e9, "unit": "Gbases/s",
               "h2d_bytes_per_step": int(tm["h2d_bytes"]), "d2h_bytes_per_step": int(tm["d2h_bytes"]),
               "ms_per_step": e2e_s * 1e3, "ms_per_step_this_rank": mine_s * 1e3, "steps": n_e2e,
               "h2d_ms": tm["h2d_ms"], "kernel_ms": tm["kernel_ms"], "d2h_ms": tm["d2h_ms"], "host_ms": tm["host_ms"],
               "numa": numa,
               "api": "kmu_sketch_pmh3a_host (pinned host buffers in the batch layout, 3-stream chunked pipeline)"}
        # the signatures the host got must be the ones left in HBM by the device-resident path
        if not torch.equal(pin_out, sig_dev.cpu()):
            raise SystemExit("e2e signatures differ from the device-resident run")

        # the host ceiling: the same bytes with NOTHING but the copies, all ranks at once -- what the box's PCIe / host
        # memory gives N GPUs together; e2e cannot beat max(copy time, kernel time)
        cs, cs2 = torch.cuda.Stream(device=f"cuda:{local_rank}"), torch.cuda.Stream(device=f"cuda:{local_rank}")
        dbuf = torch.empty(pin_in.numel(), dtype=torch.uint8, device=f"cuda:{local_rank}")

        def copies():  # the two directions at the same time, as the pipeline runs them
            with torch.cuda.stream(cs):
                dbuf.copy_(pin_in, non_blocking=True)
            with torch.cuda.stream(cs2):
                pin_out.copy_(sig_dev, non_blocking=True)
            cs.synchronize()
            cs2.synchronize()

        copies()
        _, copy_s = timed_calls(copies, 3)
        bytes_moved = pin_in.numel() + pin_out.numel() * 4
        e2e["host_ceiling"] = {"copies_only_ms": copy_s * 1e3, "gbs_per_gpu_both_directions": bytes_moved / copy_s / 1e9,
                               "gbs_all_gpus": bytes_moved * world / copy_s / 1e9,
                               "bound_ms": max(copy_s * 1e3, ms_per_step),
                               "e2e_over_bound": e2e_s * 1e3 / max(copy_s * 1e3, ms_per_step),
                               "note": "H2D of the packed reads and D2H of the signatures alone, the two directions on two streams, "
                                       "every rank at the same time (max over ranks); bound_ms = max(copies, device-resident step)"}
        del dbuf

        # the same call from SEPARATE per-sequence host allocations (the `&[&Sequence]` the Rust entry points get): one numpy
        # array per read, gathered into pinned staging memory by the library while the previous chunk is sketched
        sizes = (nbases + np.uint64(3)) // np.uint64(4)
        seqs = [packed_host[int(o): int(o) + int(sz)].copy() for o, sz in zip(off_host, sizes)]
        addrs = np.array([a.ctypes.data for a in seqs], dtype=np.uint64)
        del packed_host
        pin_out.zero_()

        def e2e_ptrs_step():
            eng.sketch_pmh3a_host_ptrs(addrs, nbases, K, KMER_TYPE, HASH_KIND, M, pin_out.data_ptr())

        e2e_ptrs_step()
        _, ptrs_s = timed_calls(e2e_ptrs_step, min(n_e2e, 3))
        if not torch.equal(pin_out, sig_dev.cpu()):
            raise SystemExit("e2e (separate sequences) signatures differ from the device-resident run")
        tm = eng.last_times()
        e2e["from_separate_sequences"] = {"value": total_bases * world / ptrs_s / 1e9, "unit": "Gbases/s", "ms_per_step": ptrs_s * 1e3,
                                          "host_ms": tm["host_ms"], "api": "kmu_sketch_pmh3a_host_ptrs (one host allocation per read, "
                                          "gathered by up to 16 host threads into pinned staging memory, chunk by chunk)"}
        del seqs, addrs

    # ---- a step on a FRESH batch: includes the per-batch build of the processing order that later steps reuse ------
    eng.sync()
    fresh = eng.batch_synth(seed, nbases)
    ev0.record(ext)
    eng.sketch_pmh3a(fresh, K, KMER_TYPE, HASH_KIND, M, out_device_ptr=sig_dev.data_ptr())
    ev1.record(ext)
    barrier()
    fresh_ms = ev0.elapsed_time(ev1)
    fresh.destroy()

    # ---- the other BASELINE shapes -----------------------------------------------------------------------------
    configs = []
    want = set(("extract,c2strong,c1,c3,c4,c5a,c5b" if args.configs == "all" else args.configs).split(",")) - {"none", ""}
    if want:
        from kmerutils_b200 import benchcfg
        T = benchcfg.Timer(eng, local_rank, world)

        def guarded(name, fn):
            try:
                r = fn()
                configs.extend(r if isinstance(r, list) else [r])
            except Exception as exc:  # a failed shape must not take the headline line with it
                configs.append({"workload": name, "error": f"{type(exc).__name__}: {exc}"})
                if world > 1:
                    raise
            torch.cuda.empty_cache()

        if "extract" in want:
            guarded("extract", lambda: benchcfg.run_extract(kb, eng, T, rank, world, args.steps, args.warmup, peak, batch, total_bases))
    batch.destroy()
    del sig_dev
    torch.cuda.empty_cache()
    if want:
        if "c2strong" in want and world > 1:
            guarded("c2strong", lambda: benchcfg.run_c2_strong(kb, eng, T, rank, world, args.steps, args.warmup, peak))
        if "c1" in want:
            guarded("c1", lambda: benchcfg.run_c1(kb, eng, T, rank, world, args.steps, args.warmup, peak))
        if "c3" in want:
            guarded("c3", lambda: benchcfg.run_c3(kb, eng, T, rank, world, args.steps, args.warmup, peak, args.c3_reads, args.exchange))
        if "c4" in want:
            guarded("c4", lambda: benchcfg.run_c4(kb, eng, T, rank, world, args.steps, args.warmup, peak))
        if "c5a" in want:
            guarded("c5a", lambda: benchcfg.run_c5a(kb, eng, T, rank, world, args.steps, args.warmup, peak))
        if "c5b" in want:
            guarded("c5b", lambda: benchcfg.run_c5b(kb, eng, T, rank, world, args.steps, args.warmup, peak))

    # ---- CPU baseline (rank 0, N = 1 only): oracle port on a bounded sample ----------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        os.sched_setaffinity(0, all_cpus)  # the CPU arm uses every host core, not only the GPU's NUMA node
        from oracle_lib import get_oracle
        orc = get_oracle()
        nthreads = orc.hardware_threads()
        rate, n_reads, bases, dt = cpu_sample_rate(orc, nbases, seed, 12.0, nthreads)
        cpu = {"value": rate, "unit": "Gbases/s", "cores": nthreads, "kind": "port",
               "sample": f"first {n_reads} reads of the workload ({bases} bases, {dt:.1f} s)"}

    if rank == 0:
        out = {
            "metric": "Gbases/s k-mer sketch (extract + count + ProbMinHash3a)", "value": value, "unit": "Gbases/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32+f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "k": K, "sketch_size": M, "reads_per_gpu": nseq,
                       "bases_per_gpu": total_bases, "l2": "inputs (1.1 GB packed) larger than L2, no flush needed",
                       "parallelism": f"reads sharded, {world} independent shard(s), no collective"},
            "kernel_ms_per_step": float(np.mean(kernel_ms)),
            "fresh_batch_step_ms": fresh_ms,
            "configs": configs,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "published_reference": {"value": 0.0859, "unit": "Gbases/s", "hardware": "8-core i7 laptop",
                                    "source": "README.md:45"},
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
