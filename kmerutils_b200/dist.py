"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the GPU box, gloo in
the CPU tests).  Only the steps of the path that really exchange data use a collective:

* per-read / per-genome sketches shard by sequence -- NO collective (`shard_by_bases`, `round_robin`);
* counting: the reference hands every k-mer to the thread `intNN_hash(kmer) % N`
  (DispatchableT::dispatch, src/base/kmercount.rs:382-420); here rank = thread, the hand-off is an
  all-to-all of the buckets `kmu_count_partition` builds (`exchange_kmers`, `count_sharded`);
* whole-file SetSketch / SuperMinHash registers are mergeable (SetSketcher::merge,
  src/sketching/setsketchert.rs:876-882): allreduce max / min (`merge_registers`).

The helpers take torch tensors on whatever device the process group's backend needs, so the same code
runs under gloo on CPU tensors (tests/test_dist_cpu.py) and under NCCL on device tensors.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_by_bases(nbases, world):
    """Contiguous ranges of sequences with about equal numbers of bases: [(start, end)] * world."""
    nb = np.asarray(nbases, dtype=np.uint64)
    cum = np.concatenate([[0], np.cumsum(nb, dtype=np.uint64)]).astype(np.float64)
    total = cum[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        cuts.append(int(np.searchsorted(cum, target, side="left")))
    cuts.append(len(nb))
    cuts = np.maximum.accumulate(np.minimum(cuts, len(nb)))
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(world)]


def round_robin(n, world, rank):
    """Indices of the units (genomes) rank `rank` owns."""
    return np.arange(rank, n, world, dtype=np.int64)


def _world(group=None):
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def exchange_kmers(keys, part_counts, group=None):
    """All-to-all of k-mer buckets.  keys: 1-D tensor, bucket p (destined to rank p) at offset
    sum(part_counts[:p]); part_counts: sequence of world ints.  Returns (received keys, recv_counts):
    the k-mers this rank owns, grouped by sending rank."""
    rank, world = _world(group)
    counts = [int(c) for c in part_counts]
    assert len(counts) == world and sum(counts) == keys.numel()
    if world == 1:
        return keys, counts
    send_counts = torch.tensor(counts, dtype=torch.int64, device=keys.device)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    recv = [int(c) for c in recv_counts.cpu()]
    out = torch.empty(sum(recv), dtype=keys.dtype, device=keys.device)
    dist.all_to_all_single(out, keys, output_split_sizes=recv, input_split_sizes=counts, group=group)
    return out, recv


def allreduce_sum(values, device, group=None):
    t = torch.tensor([int(v) for v in values], dtype=torch.int64, device=device)
    if _world(group)[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return [int(v) for v in t.cpu()]


_WIDEN = {torch.uint16: torch.int32, torch.uint32: torch.int64, torch.uint64: torch.int64}


def merge_registers(regs, op, group=None):
    """Element-wise max (SetSketch) or min (SuperMinHash) of the per-rank registers, in place semantics:
    returns a tensor of the input dtype.  NCCL has no unsigned 16/32-bit types: those are widened."""
    assert op in ("max", "min")
    if _world(group)[1] == 1:
        return regs
    wide = _WIDEN.get(regs.dtype)
    t = regs.to(wide) if wide is not None else regs.clone()
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.MIN, group=group)
    return t.to(regs.dtype) if wide is not None else t


def gather_rows(local_rows, group=None):
    """Concatenate per-rank row blocks (signatures of contiguous shards) in rank order on every rank."""
    rank, world = _world(group)
    if world == 1:
        return local_rows
    n = torch.tensor([local_rows.shape[0]], dtype=torch.int64, device=local_rows.device)
    sizes = [torch.empty_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    width = local_rows.shape[1]
    pad = torch.zeros((max(sizes), width), dtype=local_rows.dtype, device=local_rows.device)
    pad[: local_rows.shape[0]] = local_rows
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)


def count_sharded(engine, batch, k, kmer_type, capacity_per_rank, count_bits=8, canonical=True, group=None):
    """count_kmer_threaded_one_to_many across ranks (kmercount.rs:881-974): every rank extracts the
    canonical k-mers of ITS sequences, buckets them by owner on the GPU, the buckets cross NVLink in
    one all-to-all and each rank inserts what it owns.  Returns (counter, stats) where stats are the
    job-wide nb_distinct / nb_unique / nb_inserted (allreduce of the per-rank table statistics)."""
    import kmerutils_b200 as kb

    rank, world = _world(group)
    dev = torch.device("cuda", engine.device)
    tdtype = torch.int64 if kb.val_dtype(kmer_type) == np.uint64 else torch.int32
    n = batch.kmer_count(k)
    send = torch.empty(max(n, 1), dtype=tdtype, device=dev)
    _, counts = engine.count_partition(batch, k, kmer_type, world, canonical, out_device_ptr=send.data_ptr())
    engine.sync()
    torch.cuda.current_stream(dev).synchronize()
    recv, _ = exchange_kmers(send[:n], counts, group)
    torch.cuda.current_stream(dev).synchronize()
    counter = engine.counter(k, kmer_type, capacity_per_rank, count_bits)
    if recv.numel():
        counter.insert_kmers(device_ptr=recv.data_ptr(), n=recv.numel())
    st = counter.stats()
    tot = allreduce_sum([st["nb_distinct"], st["nb_unique"], st["nb_inserted"]], dev, group)
    return counter, {"nb_distinct": tot[0], "nb_unique": tot[1], "nb_inserted": tot[2], "local": st}


def query_sharded(engine, counter, kmers, kmer_type, group=None):
    """get_count for k-mers held by any rank: every rank asks its own table for the k-mers it owns
    (dispatch) and the answers are summed (a k-mer lives on exactly one rank)."""
    rank, world = _world(group)
    kmers = np.ascontiguousarray(kmers)
    local = counter.get_count(kmers).astype(np.int64)
    if world == 1:
        return local.astype(np.uint32)
    t = torch.from_numpy(local).to(torch.device("cuda", engine.device))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy().astype(np.uint32)


class P2PExchange:
    """Receive buffers of all ranks of one box, mapped into each other with CUDA IPC (grow-only).  The counting
    exchange then needs no collective on the data path: every rank's partition kernel stores its buckets straight
    into the owners' buffers over NVLink (`count_sharded_p2p`)."""

    def __init__(self, engine, group=None):
        self.engine = engine
        self.group = group
        self.rank, self.world = _world(group)
        self.local_ptr, self.local_handle, self.local_bytes, self.generation = None, None, 0, 0
        self.peer_ptrs = [None] * self.world
        self.peer_gen = [-1] * self.world

    def ensure(self, recv_bytes, symmetric=False):
        """Collective.  Makes this rank's receive buffer at least recv_bytes large and (re)maps every peer buffer that
        changed.  -> list of `world` device pointers valid in this process (own buffer at index rank).
        symmetric=True: every rank asks for the same size in the same call, so a rank whose buffer is large enough
        knows that all are and returns without any collective."""
        if symmetric and self.local_ptr is not None and recv_bytes <= self.local_bytes and all(
                p is not None for p in self.peer_ptrs):
            return list(self.peer_ptrs)
        if self.local_ptr is None or recv_bytes > self.local_bytes:
            if self.local_ptr is not None:
                # peers still map the old buffer: they drop it below, after the barrier implied by the gather
                old = self.local_ptr
            else:
                old = None
            size = max(int(recv_bytes) if symmetric else int(recv_bytes * 1.25), 1 << 20)
            self.local_ptr, self.local_handle = self.engine.ipc_alloc(size)
            self.local_bytes = size
            self.generation += 1
            self._old = old
        else:
            self._old = None
        if self.world == 1:
            if self._old is not None:
                self.engine.ipc_free(self._old)
            return [self.local_ptr]
        info = [None] * self.world
        dist.all_gather_object(info, (self.local_handle, self.generation), group=self.group)
        for r, (handle, gen) in enumerate(info):
            if r == self.rank:
                self.peer_ptrs[r] = self.local_ptr
                continue
            if gen != self.peer_gen[r]:
                if self.peer_ptrs[r] is not None:
                    self.engine.ipc_close(self.peer_ptrs[r])
                self.peer_ptrs[r] = self.engine.ipc_open(handle)
                self.peer_gen[r] = gen
        dist.barrier(group=self.group)  # every peer has dropped its mapping of a replaced buffer
        if self._old is not None:
            self.engine.ipc_free(self._old)
        return list(self.peer_ptrs)

    def close(self):
        for r, p in enumerate(self.peer_ptrs):
            if p is not None and r != self.rank:
                self.engine.ipc_close(p)
        self.peer_ptrs = [None] * self.world
        if self.world > 1:
            dist.barrier(group=self.group)
        if self.local_ptr is not None:
            self.engine.ipc_free(self.local_ptr)
            self.local_ptr = None


def count_sharded_p2p(engine, batch, k, kmer_type, counter, xchg, canonical=True, group=None):
    """One round of the counting exchange without a data-path collective: the partition kernel of every rank writes
    its buckets directly into the owners' receive buffers (peer memory over NVLink), then every rank inserts what
    landed in its buffer.  `counter` is this rank's table, `xchg` a P2PExchange.  -> number of k-mers received"""
    import kmerutils_b200 as kb

    rank, world = _world(group)
    esz = 8 if kb.val_dtype(kmer_type) == np.uint64 else 4
    counts = engine.count_partition_counts(batch, k, kmer_type, world, canonical)
    if world > 1:
        mat = [None] * world
        dist.all_gather_object(mat, [int(c) for c in counts], group=group)
    else:
        mat = [[int(c) for c in counts]]
    recv_total = sum(mat[s][rank] for s in range(world))
    dests = xchg.ensure(recv_total * esz)
    offsets = [sum(mat[s][p] for s in range(rank)) for p in range(world)]
    engine.count_partition_scatter(batch, k, kmer_type, world, dests, offsets, canonical)  # returns when the kernel is done
    if world > 1:
        dist.barrier(group=group)  # every sender's stores have landed
    if recv_total:
        counter.insert_kmers(device_ptr=xchg.local_ptr, n=recv_total)
    if world > 1:
        dist.barrier(group=group)  # the buffers may be overwritten by the next round
    return recv_total


def exchange_slab_cap(nk_max_per_rank, world, nregions):
    """Keys per slab for a round in which no rank sends more than nk_max_per_rank k-mers: the expected share of a
    (sender, owner, region) bucket + 64 sqrt(share) + slack (8 sigma for k-mers repeated 64 times on average: a key seen
    c times puts c entries into one bucket)."""
    mean = nk_max_per_rank / float(world * nregions)
    return int(mean + 64.0 * mean ** 0.5 + 1024)


def count_round_fused(engine, batch, counter, xchg, nk_bound, canonical=True, group=None, phases=None):
    """One round of multi-GPU counting with the fused exchange (kmu_count_exchange_scatter): a single kernel per rank
    extracts the canonical k-mers, buckets them by owner (intNN_hash % world) and stores them into the owners' receive
    buffers over NVLink; the ranks then share their bucket counts (a few KB) and every rank partitions what it received by
    region of its table and inserts region after region (kmu_count_insert_slabs).  `nk_bound`: no rank sends more k-mers than this in
    a round (fixes the slab size; the same on every rank).  Every rank's `counter` was created with the same arguments.
    -> (keys received, bytes sent to peers)"""
    import kmerutils_b200 as kb

    rank, world = _world(group)
    esz = 8 if kb.val_dtype(counter.kmer_type) == np.uint64 else 4
    nreg = counter.exchange_regions(world)
    slab_cap = exchange_slab_cap(nk_bound, world, nreg)
    dests = xchg.ensure(nreg * world * slab_cap * esz, symmetric=True)
    import time as _time
    sent, overflowed = counter.exchange_scatter(batch, world, rank, slab_cap, dests, canonical)  # returns when the kernel is done
    if phases is not None:
        phases["scatter_ms"] = phases.get("scatter_ms", 0.0) + engine.last_times()["kernel_ms"]
    t_share = _time.perf_counter()
    if world > 1:
        dev = torch.device("cuda", engine.device)
        mine = torch.from_numpy(np.append(sent.reshape(-1).astype(np.int64), int(overflowed))).to(dev)
        allc = torch.empty((world, mine.numel()), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allc, mine, group=group)  # also the barrier: every sender's stores have landed
        allc = allc.cpu().numpy()
        bad = bool(allc[:, -1].any())
        counts = allc[:, :-1].reshape(world, world, nreg)[:, rank, :]
    else:
        bad = overflowed
        counts = sent.reshape(1, 1, nreg)[:, 0, :]
    if bad:
        raise RuntimeError("count_round_fused: a slab overflowed (one k-mer repeated millions of times?); use count_sharded")
    if phases is not None:
        phases["share_counts_ms"] = phases.get("share_counts_ms", 0.0) + (_time.perf_counter() - t_share) * 1e3
    counter.insert_slabs(xchg.local_ptr, slab_cap, counts.astype(np.uint64))
    if phases is not None:
        phases["insert_ms"] = phases.get("insert_ms", 0.0) + engine.last_times()["kernel_ms"]
    if world > 1:
        t_bar = _time.perf_counter()
        dist.barrier(group=group)  # the buffers may be overwritten by the next round
        if phases is not None:  # what this rank waits for the slowest one
            phases["end_barrier_ms"] = phases.get("end_barrier_ms", 0.0) + (_time.perf_counter() - t_bar) * 1e3
    sent_peer = int(sent.sum() - sent[rank].sum()) * esz
    return int(counts.sum()), sent_peer


def merge_pmh3a_registers(hbits, keys, device, group=None):
    """ProbMinHash3a registers are per-slot minima of (h, key): across ranks an allreduce-min on h (positive doubles
    order like their bit patterns) and, among the ranks that hold that minimum, the smallest key.  hbits, keys: u64
    arrays of m entries (numpy).  -> (hbits, keys) of the merged sketch, identical on every rank."""
    h = np.ascontiguousarray(hbits, dtype=np.uint64)
    kk = np.ascontiguousarray(keys, dtype=np.uint64)
    if _world(group)[1] == 1:
        return h, kk
    th = torch.from_numpy(h.view(np.int64).copy()).to(device)  # h > 0: the sign bit is clear
    tmin = th.clone()
    dist.all_reduce(tmin, op=dist.ReduceOp.MIN, group=group)
    # keys as signed integers that order like the unsigned ones; ranks without the minimum stay out of the way
    flipped = torch.from_numpy((kk ^ np.uint64(1 << 63)).view(np.int64).copy()).to(device)
    cand = torch.where(th == tmin, flipped, torch.full_like(flipped, torch.iinfo(torch.int64).max))
    dist.all_reduce(cand, op=dist.ReduceOp.MIN, group=group)
    out_h = tmin.cpu().numpy().view(np.uint64)
    out_k = cand.cpu().numpy().view(np.uint64) ^ np.uint64(1 << 63)
    return out_h, out_k


def pmh3a_whole_sharded(engine, batch, k, kmer_type, hash_kind, m, group=None):
    """ProbHash3aSketch::sketch_compressedkmer_seqs (setsketchert.rs:160-202) for a file spread over the ranks: the
    k-mers go to the rank that owns them (count_sharded: every key's multiplicity is complete on one GPU), every rank
    sketches its keys into partial registers, the registers are merged (merge_pmh3a_registers).  Items are cut at a
    bound; the merged maximum must stay below it, else the bound grows and the registers are rebuilt -- the same
    verification as on one GPU.  -> signature (m values), identical on every rank"""
    import math

    import kmerutils_b200 as kb

    rank, world = _world(group)
    dev = torch.device("cuda", engine.device)
    canonical = hash_kind in (kb.HASH_CANON_INVHASH, kb.HASH_CANON_RAW)
    nk_local = batch.kmer_count(k)
    nk_total = allreduce_sum([nk_local], dev, group)[0]
    cap = max(1024, int(nk_total * 1.5 / world) + 1024)
    counter, stats = count_sharded(engine, batch, k, kmer_type, cap, count_bits=32, canonical=canonical, group=group)
    distinct = stats["nb_distinct"]
    sig_dtype = kb.val_dtype(kmer_type)
    if distinct == 0:
        counter.destroy()
        return np.zeros(m, dtype=sig_dtype)
    bound_max = m * (math.log(m) + 40.0)  # one key alone fills every slot below this (kmu_pmh3a_counter_slots refuses more)
    bound = min(bound_max, m / distinct * math.log(m / 1e-4))
    while True:
        h, keys = engine.pmh3a_counter_slots(counter, hash_kind, m, bound)
        h, keys = merge_pmh3a_registers(h, keys, dev, group)
        top = float(h.view(np.float64).max())
        if top < bound:
            counter.destroy()
            return keys.astype(sig_dtype)
        if bound >= bound_max:
            counter.destroy()
            raise RuntimeError("ProbMinHash3a sharded sketch did not converge")
        bound = min(bound_max, bound * 4.0)
