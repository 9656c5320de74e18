"""GPU tier: one mesh over several GPUs (include/mof_b200.h, mof_dist_*). The row-partitioned flow solve must give the
single-GPU result on every rank. A world of 1 runs the same code path on any box; the 2-rank case needs two GPUs."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(cmd):
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert p.returncode == 0 and lines, p.stdout[-2000:] + p.stderr[-2000:]
    return json.loads(lines[-1])


def _check(d, world):
    assert d["world"] == world and d["ranks_agree"] and d["repeatable"]
    assert d["flow_rel_diff_vs_single_gpu"] < 1e-6 and d["colour_max_diff"] < 1e-3
    assert d["last_flow_residual"] <= 1.01e-8


def test_partitioned_path_with_one_rank():
    d = _run([sys.executable, "tests/dist_worker.py", "6", "2"])
    _check(d, 1)
    assert d["halo_entries_rank0"] == 0


def test_partitioned_path_on_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = 29700 + os.getpid() % 200
    d = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port),
              "tests/dist_worker.py", "7", "2"])
    _check(d, 2)
    assert d["halo_entries_rank0"] > 0
