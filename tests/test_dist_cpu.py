"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: sharding, the k-mer exchange of the
counting path (DispatchableT::dispatch owners) and register merges.  The k-mers / registers come from the
CPU oracle here; on the GPU box the same helpers move the CUDA path's buffers over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as ol
from kmerutils_b200 import dist as kd

K = 21


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def make_reads(oracle, nreads=60, genome=4000, read_len=150):
    g = oracle.synth_ascii(123, 0, genome)
    rng = np.random.default_rng(1)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    reads = []
    for _ in range(nreads):
        s = int(rng.integers(0, genome - read_len + 1))
        r = g[s:s + read_len]
        reads.append(r.translate(comp)[::-1] if rng.integers(0, 2) else r)
    return reads


def pack_batch(oracle, reads):
    packed = [oracle.pack_2bit(r) for r in reads]
    off, cur = [], 0
    for p in packed:
        off.append(cur)
        cur += (len(p) + 15) // 16 * 16
    buf = np.zeros(cur + 16, np.uint8)
    for o, p in zip(off, packed):
        buf[o:o + len(p)] = p
    return buf, np.array(off, np.uint64), np.array([len(r) for r in reads], np.uint64)


def canonical_kmers(oracle, reads):
    out = []
    for r in reads:
        w = oracle.generate_kmers(oracle.pack_2bit(r), len(r), K, ol.KMER64)
        out.append(oracle.apply_hash(w, K, ol.KMER64, ol.HASH_CANON_RAW))
    return np.concatenate(out)


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        oracle = ol.get_oracle()
        reads = make_reads(oracle)
        nb = np.array([len(r) for r in reads], np.uint64)
        lo, hi = kd.shard_by_bases(nb, world)[rank]
        mine = reads[lo:hi]
        # ---- counting: bucket by owner, exchange, count what this rank owns ----
        kmers = canonical_kmers(oracle, mine)
        owner = oracle.dispatch(kmers, ol.KMER64, world)
        order = np.argsort(owner, kind="stable")
        counts = np.bincount(owner.astype(np.int64), minlength=world)
        send = torch.from_numpy(kmers[order].astype(np.int64))
        recv, recv_counts = kd.exchange_kmers(send, counts)
        got = recv.numpy().astype(np.uint64)
        assert len(got) == sum(recv_counts)
        assert (oracle.dispatch(np.unique(got), ol.KMER64, world) == rank).all()  # only k-mers this rank owns
        keys, mult = np.unique(got, return_counts=True)
        # the reference result: count everything in one place, keep this rank's owners
        buf, off, nb_all = pack_batch(oracle, reads)
        all_keys, all_mult = oracle.count_kmers(buf, off, nb_all, K, ol.KMER64, True)
        sel = oracle.dispatch(all_keys, ol.KMER64, world) == rank
        assert np.array_equal(keys, all_keys[sel]) and np.array_equal(mult.astype(np.uint64), all_mult[sel])
        tot = kd.allreduce_sum([len(keys), int((mult == 1).sum()), int(mult.sum())], "cpu")
        assert tot == [len(all_keys), int((all_mult == 1).sum()), int(all_mult.sum())]
        # ---- whole-file registers: per-rank sketch of the shard, merged ----
        pbuf, poff, pnb = pack_batch(oracle, mine)
        prm = (1.001, 128, 20.0, 65534)
        local_hll = oracle.sketch_setsketch_seqs(pbuf, poff, pnb, 12, ol.KMER32, ol.HASH_CANON_INVHASH, prm)
        merged = kd.merge_registers(torch.from_numpy(local_hll.astype(np.int32)).to(torch.int32), "max").numpy()
        want_hll = oracle.sketch_setsketch_seqs(buf, off, nb_all, 12, ol.KMER32, ol.HASH_CANON_INVHASH, prm)
        assert np.array_equal(merged.astype(np.uint16), want_hll)
        u16 = kd.merge_registers(torch.from_numpy(local_hll.view(np.int16)).view(torch.uint16), "max")
        assert np.array_equal(u16.view(torch.int16).numpy().view(np.uint16), want_hll)  # widened for the collective
        local_smh = oracle.sketch_superminhash_seqs(pbuf, poff, pnb, 12, ol.KMER32, ol.HASH_CANON_INVHASH, 64)
        merged_smh = kd.merge_registers(torch.from_numpy(local_smh), "min").numpy()
        want_smh = oracle.sketch_superminhash_seqs(buf, off, nb_all, 12, ol.KMER32, ol.HASH_CANON_INVHASH, 64)
        assert np.array_equal(merged_smh, want_smh)
        # ---- per-read signatures: no collective on the data path, rows gathered in input order ----
        sig = oracle.sketch_pmh3a_batch(pbuf, poff, pnb, 8, ol.KMER32, ol.HASH_CANON_INVHASH, 32, 1)
        rows = kd.gather_rows(torch.from_numpy(sig.astype(np.int64)))
        want_sig = oracle.sketch_pmh3a_batch(buf, off, nb_all, 8, ol.KMER32, ol.HASH_CANON_INVHASH, 32, 1)
        assert np.array_equal(rows.numpy().astype(np.uint32), want_sig)
        # ---- ProbMinHash3a registers: per-slot minimum of (h, key) over the ranks ----
        rng = np.random.default_rng(99)  # same stream on every rank: all ranks' registers are known everywhere
        hs = rng.random((world, 64)) + 1e-9
        ks = rng.integers(0, 2**64, (world, 64), dtype=np.uint64)  # keys beyond 2^63 too
        hs[1, :8] = hs[0, :8]  # equal h on both ranks: the smaller key wins
        hs[0, 8:12] = np.finfo(np.float64).max  # slots one rank never filled
        mh, mk = kd.merge_pmh3a_registers(hs[rank].view(np.uint64), ks[rank], "cpu")
        for j in range(64):
            best = min(range(world), key=lambda r: (hs[r, j], int(ks[r, j])))
            assert mh[j] == hs[best, j].view(np.uint64) and mk[j] == ks[best, j], j
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_world2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_shard_by_bases():
    nb = np.array([10, 10, 10, 10, 100, 1, 1, 1, 57], dtype=np.uint64)
    for world in (1, 2, 3, 4, 8):
        sh = kd.shard_by_bases(nb, world)
        assert sh[0][0] == 0 and sh[-1][1] == len(nb)
        assert all(sh[i][1] == sh[i + 1][0] for i in range(world - 1))
        assert all(a <= b for a, b in sh)
    two = kd.shard_by_bases(nb, 2)
    left = int(nb[two[0][0]:two[0][1]].sum())
    assert abs(left - 100) <= 100  # the 100-base read decides the cut
    assert kd.shard_by_bases(np.zeros(0, np.uint64), 4) == [(0, 0)] * 4
    assert kd.round_robin(10, 4, 1).tolist() == [1, 5, 9]
