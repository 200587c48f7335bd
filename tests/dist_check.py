#!/usr/bin/env python
"""Multi-GPU check, run under torchrun on N GPUs of one box:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py
Counting over NCCL all-to-all (owners = DispatchableT::dispatch) and SetSketch / SuperMinHash register merges, each
compared on rank 0 with the CPU oracle run on the whole input."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import kmerutils_b200 as kb  # noqa: E402
from kmerutils_b200 import dist as kd  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = kb.Engine(local)
    glen, nreads, k = 200_000, 40_000, 31
    genome = eng.batch_synth(3, np.array([glen], dtype=np.uint64))
    per = nreads // world
    first = rank * per
    mine = eng.batch_sample_reads(genome, 3, first, per if rank < world - 1 else nreads - first, 150, 5000)
    counter, st = kd.count_sharded(eng, mine, k, kb.KMER64, capacity_per_rank=int(nreads * 150 * 1.2 / world) + 1024)
    # the same exchange without NCCL on the data path: partition kernel stores into peer buffers (CUDA IPC / NVLink)
    xchg = kd.P2PExchange(eng)
    counter2 = eng.counter(k, kb.KMER64, int(nreads * 150 * 1.2 / world) + 1024)
    half = mine  # two rounds over the same reads would double the counts: split the shard in two slices instead
    nloc = len(mine)
    idx = np.arange(nloc, dtype=np.uint64)
    for lo, hi in ((0, nloc // 2), (nloc // 2, nloc)):
        part = eng.batch_slices(mine, idx[lo:hi], np.zeros(hi - lo, np.uint64), np.full(hi - lo, 150, np.uint64))
        kd.count_sharded_p2p(eng, part, k, kb.KMER64, counter2, xchg)
        part.destroy()
    st2 = counter2.stats()
    tot2 = kd.allreduce_sum([st2["nb_distinct"], st2["nb_unique"], st2["nb_inserted"]], torch.device("cuda", local))
    # fused exchange, one walk: (owner, table region) buckets stored into the peers' slabs, regioned insertion
    # (small regions so that this small table is cut into many)
    os.environ["KMU_COUNT_REGION_KB"] = "64"
    counter3 = eng.counter(k, kb.KMER64, int(nreads * 150 * 1.2 / world) + 1024)
    nk_bound = (nreads - (world - 1) * per) * (150 - k + 1)
    for lo, hi in ((0, nloc // 2), (nloc // 2, nloc)):
        part = eng.batch_slices(mine, idx[lo:hi], np.zeros(hi - lo, np.uint64), np.full(hi - lo, 150, np.uint64))
        kd.count_round_fused(eng, part, counter3, xchg, nk_bound)
        part.destroy()
    os.environ.pop("KMU_COUNT_REGION_KB")
    st3 = counter3.stats()
    tot3 = kd.allreduce_sum([st3["nb_distinct"], st3["nb_unique"], st3["nb_inserted"]], torch.device("cuda", local))
    hll_local = eng.sketch_setsketch(mine, 21, kb.KMER64, kb.HASH_CANON_INVHASH, (1.001, 256, 20.0, 65534), np.uint16, whole=True)
    hll = kd.merge_registers(torch.from_numpy(hll_local.astype(np.int32)).cuda(), "max").cpu().numpy().astype(np.uint16)
    smh_local = eng.sketch_superminhash(mine, 21, kb.KMER64, kb.HASH_CANON_INVHASH, 256).min(axis=0)
    smh = kd.merge_registers(torch.from_numpy(smh_local).cuda(), "min").cpu().numpy()
    # one long sequence cut into `world` chunks with a k - 1 halo (bench.py C5a at N > 1): rank r sketches chunk r, the
    # registers merge with the allreduce
    hl = np.array([3_000_000], dtype=np.uint64)
    hfull = eng.batch_synth(21, hl)
    hb = (hl * np.uint64(rank)) // np.uint64(world)
    he = np.minimum(hl, (hl * np.uint64(rank + 1)) // np.uint64(world) + np.uint64(20))
    hmine = eng.batch_slices(hfull, np.zeros(1, np.uint64), hb, he)
    halo_local = eng.sketch_setsketch(hmine, 21, kb.KMER64, kb.HASH_CANON_INVHASH, (1.001, 256, 20.0, 65534), np.uint16, whole=True)
    halo_hll = kd.merge_registers(torch.from_numpy(halo_local.astype(np.int32)).cuda(), "max").cpu().numpy().astype(np.uint16)
    halo_whole = eng.sketch_setsketch(hfull, 21, kb.KMER64, kb.HASH_CANON_INVHASH, (1.001, 256, 20.0, 65534), np.uint16, whole=True)
    # whole-file ProbMinHash3a over the shards: owners count, sketch their keys, registers merge (min h, then key)
    pmh = {kk: kd.pmh3a_whole_sharded(eng, mine, kk, kt, kb.HASH_CANON_INVHASH, 500)
           for kk, kt in ((21, kb.KMER64), (16, kb.KMER16B32))}
    ok = True
    if rank == 0:
        from oracle_lib import get_oracle
        from test_pmh3a_gpu import oracle_batch
        orc = get_oracle()
        gp, _ = oracle_batch(orc, 3, np.array([glen], dtype=np.uint64))
        reads = orc.sample_reads(gp, glen, 3, 0, nreads, 150, 5000)
        packed = [orc.pack_2bit(r) for r in reads]
        buf = np.zeros(48 * nreads + 16, np.uint8)
        for i, p in enumerate(packed):
            buf[48 * i: 48 * i + len(p)] = p
        off = np.arange(nreads, dtype=np.uint64) * 48
        nb = np.full(nreads, 150, dtype=np.uint64)
        keys, cnts = orc.count_kmers(buf, off, nb, k, kb.KMER64, True)
        want = (len(keys), int((cnts == 1).sum()), int(cnts.sum()))
        got = (st["nb_distinct"], st["nb_unique"], st["nb_inserted"])
        ok &= got == want
        print(f"[dist_check] world={world} counting stats {got} want {want}", flush=True)
        ok &= tuple(tot2) == want
        print(f"[dist_check] peer-to-peer exchange (no data-path collective) stats {tuple(tot2)} ok={tuple(tot2) == want}", flush=True)
        ok &= tuple(tot3) == want
        print(f"[dist_check] fused one-walk exchange + regioned insertion stats {tuple(tot3)} ok={tuple(tot3) == want}", flush=True)
        want_hll = orc.sketch_setsketch_seqs(buf, off, nb, 21, kb.KMER64, kb.HASH_CANON_INVHASH, (1.001, 256, 20.0, 65534))
        ok &= bool(np.array_equal(hll, want_hll))
        want_smh = orc.sketch_superminhash_seqs(buf, off, nb, 21, kb.KMER64, kb.HASH_CANON_INVHASH, 256)
        ok &= bool(np.array_equal(smh, want_smh))
        print(f"[dist_check] setsketch merge ok={np.array_equal(hll, want_hll)} superminhash merge ok={np.array_equal(smh, want_smh)}", flush=True)
        hp_, ho_ = oracle_batch(orc, 21, hl)
        want_halo = orc.sketch_setsketch_seqs(hp_, ho_, hl, 21, kb.KMER64, kb.HASH_CANON_INVHASH, (1.001, 256, 20.0, 65534))
        halo_ok = bool(np.array_equal(halo_hll, want_halo) and np.array_equal(halo_whole, want_halo))
        ok &= halo_ok
        print(f"[dist_check] one 3 Mb sequence in {world} chunks with a k-1 halo, allreduce-max vs the oracle's single-sequence sketch: ok={halo_ok}", flush=True)
        for kk, kt in ((21, kb.KMER64), (16, kb.KMER16B32)):
            want_pmh = orc.sketch_pmh3a_seqs(buf, off, nb, kk, kt, kb.HASH_CANON_INVHASH, 500)
            same = bool(np.array_equal(pmh[kk].astype(np.uint64), want_pmh))
            ok &= same
            print(f"[dist_check] whole-file probminhash3a over {world} ranks, k={kk}: ok={same}", flush=True)
        probe = keys[:: max(1, len(keys) // 5000)]
        probe_want = np.minimum(cnts[:: max(1, len(keys) // 5000)], 255).astype(np.uint32)
    else:
        probe = None
    # every rank must take part in the query collective with the same probe set
    obj = [probe]
    dist.broadcast_object_list(obj, src=0)
    res = kd.query_sharded(eng, counter, obj[0], kb.KMER64)
    if rank == 0:
        ok &= bool(np.array_equal(res, probe_want))
        print(f"[dist_check] sharded get_count ok={np.array_equal(res, probe_want)}", flush=True)
        print("[dist_check] PASS" if ok else "[dist_check] FAIL", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    counter.destroy()
    counter2.destroy()
    counter3.destroy()
    xchg.close()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
