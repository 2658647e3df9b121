"""Real multi-GPU check of the sharded path (copy-engine all-gather, CUDA-IPC peer-memory push, NCCL): needs
>= 2 GPUs on the box, otherwise skipped.  Runs scripts/dist_check.py under torchrun for several layouts; every
rank compares its shard with the single-GPU propagation of the same graph."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _run(nproc, grid, mode, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "scripts", "dist_check.py"), grid, mode]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("PASS") == nproc and "FAIL" not in out.stdout


@pytest.mark.parametrize("grid,mode", [("2x1", "copy"), ("2x1", "push"), ("2x1", "nccl"), ("1x2", "push")])
def test_sharded_propagation_two_gpus(grid, mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, grid, mode, 29600 + hash((grid, mode)) % 200)


@pytest.mark.parametrize("grid", ["4x1", "2x2"])
def test_sharded_propagation_four_gpus(grid):
    if torch.cuda.device_count() < 4:
        pytest.skip("needs 4 GPUs")
    _run(4, grid, "push", 29800 + hash(grid) % 100)
