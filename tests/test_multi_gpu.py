"""The N-rank path on real GPUs: runs tests/multi_gpu_check.py under torchrun when the box has at least two devices
(pattern / entries / SpMV with halo exchange / Jacobi-, block-Jacobi- and multigrid-preconditioned CG / products /
error norms / estimators / p = 2 against the oracle on every rank).  Skipped on single-GPU boxes; the host-side
partition logic of the same path is covered on CPU by tests/test_multi_gloo.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_ranks_against_the_oracle(gpu):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("multi-GPU check ok") == 3
