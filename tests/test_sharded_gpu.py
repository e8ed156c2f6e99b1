"""Particle-sharded forward on 2+ GPUs == single-GPU forward, bit for bit (both result exchanges), inside the -m gpu
suite: launches tools/check_sharded.py under torchrun.  Skips on a box with a single GPU."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_sharded_equals_single_gpu_bit_exact():
    world = 2 if torch.cuda.device_count() < 4 else 4
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tools", "check_sharded.py")]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    print(r.stdout[-4000:])
    assert r.returncode == 0, r.stderr[-4000:]
    assert "bit-exact: False" not in r.stdout and "bit-exact: True" in r.stdout
