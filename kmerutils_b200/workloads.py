"""Synthetic workloads of BASELINE.json: read-length models only.

The bases themselves come from the counter-based SplitMix64 stream implemented on the device
(kmu_seqbatch_synth) and in the oracle (orc_synth_packed), so both sides see identical input.
"""
import numpy as np


def c1_lengths():
    """config 1: 1000 reads x 1000 b (1 Mbase), seed 1."""
    return np.full(1000, 1000, dtype=np.uint64)


def c2_lengths(n_reads=746_333, total_bases=4_380_000_000, seed=2):
    """config 2 (README benchmark shape): ONT-like log-normal read lengths,
    round(exp(N(8.35, 0.85))) clamped to [200, 250000], rescaled so that the sum is total_bases."""
    rng = np.random.default_rng(seed)
    L = np.exp(rng.normal(8.35, 0.85, n_reads))
    L = np.clip(np.rint(L), 200, 250_000)
    L = np.clip(np.rint(L * (total_bases / L.sum())), 200, 250_000).astype(np.int64)
    # absorb the rounding residue in the longest reads, one base each
    diff = int(total_bases - L.sum())
    idx = np.argsort(-L)[: abs(diff)]
    L[idx] += 1 if diff > 0 else -1
    return L.astype(np.uint64)


def batch_layout(nbases):
    """byte offsets of the 16-byte aligned batch layout (include/kmerutils_b200.h) and total bytes."""
    nb = np.asarray(nbases, dtype=np.uint64)
    sizes = ((nb + np.uint64(3)) // np.uint64(4) + np.uint64(15)) // np.uint64(16) * np.uint64(16)
    off = np.zeros(len(nb), dtype=np.uint64)
    if len(nb) > 1:
        off[1:] = np.cumsum(sizes)[:-1]
    return off, int(sizes.sum())
