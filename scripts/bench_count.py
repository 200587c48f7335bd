#!/usr/bin/env python
"""Counting benchmark on config C3's shape (150-base reads from a 100 Mb genome, 0.5 % substitutions, k = 31 canonical
Kmer64bit), weak scaling: every rank counts `--reads` reads per round for `--rounds` rounds.
  single GPU : python scripts/bench_count.py
  N GPUs     : python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
               scripts/bench_count.py [--exchange nccl|p2p]
Prints one JSON line (rank 0): job-wide Gbases/s, per-phase milliseconds, bytes crossing NVLink per round."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kmerutils_b200 as kb  # noqa: E402
from kmerutils_b200 import dist as kd  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=8_000_000)
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--exchange", default="fused", choices=["fused", "p2p", "nccl"])
    ap.add_argument("--genome", type=int, default=100_000_000)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    eng = kb.Engine(local)
    k = 31
    genome = eng.batch_synth(3, np.array([args.genome], dtype=np.uint64))
    nk_round = args.reads * (150 - k + 1)
    # distinct k-mers this rank will own over all rounds: genome k-mers / world + the error k-mers it receives
    cap = int(args.genome / world * 1.1 + args.rounds * nk_round * 0.22)
    counter = eng.counter(k, kb.KMER64, capacity=cap)
    xchg = kd.P2PExchange(eng) if args.exchange in ("p2p", "fused") else None
    send = torch.empty(nk_round, dtype=torch.int64, device=dev) if args.exchange == "nccl" else None
    t_part = t_xchg = t_ins = 0.0
    sent_bytes = 0

    def barrier():
        torch.cuda.synchronize()
        eng.sync()
        if world > 1:
            dist.barrier()

    # the read batches of all rounds are resident in HBM before the clock starts (round 0 is the warm-up)
    batches = [eng.batch_sample_reads(genome, 3, (r * world + rank) * args.reads, args.reads, 150, 5000)
               for r in range(args.rounds + 1)]
    barrier()
    t0 = time.perf_counter()
    for r in range(args.rounds + 1):  # round 0 is the warm-up (allocations, IPC mapping)
        if r == 1:
            barrier()
            t0 = time.perf_counter()
            t_part = t_xchg = t_ins = 0.0
            sent_bytes = 0
        reads = batches[r]
        if world == 1:
            a = time.perf_counter()
            counter.insert_seqs(reads, canonical=True)
            t_ins += time.perf_counter() - a
        elif args.exchange == "fused":
            a = time.perf_counter()
            _, sent = kd.count_round_fused(eng, reads, counter, xchg, nk_round)
            t_ins += time.perf_counter() - a
            sent_bytes += sent
        elif args.exchange == "p2p":
            a = time.perf_counter()
            counts = eng.count_partition_counts(reads, k, kb.KMER64, world, True)
            mat = [None] * world
            dist.all_gather_object(mat, [int(c) for c in counts])
            recv_total = sum(mat[s][rank] for s in range(world))
            dests = xchg.ensure(recv_total * 8)
            offsets = [sum(mat[s][p] for s in range(rank)) for p in range(world)]
            b = time.perf_counter()
            eng.count_partition_scatter(reads, k, kb.KMER64, world, dests, offsets, True)
            dist.barrier()
            c = time.perf_counter()
            counter.insert_kmers(device_ptr=xchg.local_ptr, n=recv_total)
            dist.barrier()
            d = time.perf_counter()
            t_part += b - a
            t_xchg += c - b
            t_ins += d - c
            sent_bytes += 8 * sum(int(x) for i, x in enumerate(counts) if i != rank)
        else:
            a = time.perf_counter()
            _, counts = eng.count_partition(reads, k, kb.KMER64, world, True, out_device_ptr=send.data_ptr())
            b = time.perf_counter()
            recv, _ = kd.exchange_kmers(send[:nk_round], counts)
            torch.cuda.synchronize()
            c = time.perf_counter()
            counter.insert_kmers(device_ptr=recv.data_ptr(), n=recv.numel())
            d = time.perf_counter()
            t_part += b - a
            t_xchg += c - b
            t_ins += d - c
            sent_bytes += 8 * sum(int(x) for i, x in enumerate(counts) if i != rank)
    barrier()
    elapsed = time.perf_counter() - t0
    for b_ in batches:
        b_.destroy()
    st = counter.stats()
    tot = kd.allreduce_sum([st["nb_distinct"], st["nb_unique"], st["nb_inserted"]], dev)
    if world > 1:
        t = torch.tensor([elapsed], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t.item())
    if rank == 0:
        bases = args.rounds * args.reads * 150 * world
        print(json.dumps({
            "bench": "count C3 shape", "n_gpus": world, "exchange": args.exchange if world > 1 else "none",
            "reads_per_rank_per_round": args.reads, "rounds": args.rounds, "gbases_s": round(bases / elapsed / 1e9, 2),
            "ms_per_round": round(elapsed / args.rounds * 1e3, 1),
            "ms_partition_counts": round(t_part / args.rounds * 1e3, 1), "ms_scatter_or_alltoall": round(t_xchg / args.rounds * 1e3, 1),
            "ms_insert": round(t_ins / args.rounds * 1e3, 1),
            "nvlink_bytes_sent_per_rank_per_round": sent_bytes // max(args.rounds, 1),
            "nb_distinct": tot[0], "nb_unique": tot[1], "nb_inserted": tot[2],
            "includes": "partition, exchange, insertion (reads resident in HBM); wall clock between barriers, max over ranks"}),
            flush=True)
    counter.destroy()
    if xchg:
        xchg.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
