"""The N-rank path on real GPUs: runs tests/multi_gpu_check.py under torchrun when the box has at least two devices
(pattern / entries / SpMV with halo exchange / Jacobi-, block-Jacobi- and multigrid-preconditioned CG / products /
error norms / estimators / p = 2 against the oracle on every rank).  Skipped on single-GPU boxes; the host-side
partition logic of the same path is covered on CPU by tests/test_multi_gloo.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    env = dict(os.environ)
    env.update(env_extra)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("multi-GPU check ok") == 4
    return [line for line in r.stdout.splitlines() if line.startswith("cg.mg iterations")]


@pytest.mark.gpu
def test_two_ranks_against_the_oracle(gpu):
    """default: peer-memory SpMV for the Jacobi solve, strip-distributed multigrid with as many distributed levels as the
    strips allow; then every other mode of the multigrid (1 and 2 distributed levels, replicated V-cycle) and the NCCL
    halo exchange - all against the same oracle solutions, with the same iteration counts"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    base = _run({})
    for env in ({"HDD_MG_DIST_LEVELS": "1"}, {"HDD_MG_DIST_LEVELS": "2"}, {"HDD_MG_DISTRIBUTED": "0", "HDD_P2P": "0"}):
        assert _run(env) == base, env
