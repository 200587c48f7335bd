"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): tests/dist_check.py under torchrun -- counting over the
NCCL all-to-all and over the fused peer-to-peer exchange, register merges, sharded whole-file ProbMinHash3a and sharded
queries, each compared with the CPU oracle on the whole input."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_dist_check_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU on this box")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "dist_check.py")],
                       capture_output=True, text=True, timeout=280, cwd=ROOT)
    assert r.returncode == 0 and "[dist_check] PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
