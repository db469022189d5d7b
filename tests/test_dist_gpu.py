"""GPU tier, needs >= 2 GPUs (``gpurun --gpus 2``): the row-partitioned engine
with both halo back ends against the single-GPU engine, bit for bit."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_bit_identical():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 4 if n >= 4 else 2
    cmd = [
        sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
        "--master-addr", "127.0.0.1", "--master-port", "29517",
        os.path.join(ROOT, "tests", "dist_workers.py"),
    ]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    print(r.stdout[-4000:], r.stderr[-4000:])
    assert r.returncode == 0 and "DIST_GPU_OK" in r.stdout
