"""-m gpu, needs >= 2 GPUs on the box: the native NCCL ring + ring drivers, one
process per GPU under torchrun, reference protocol (tests/ring_check.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_ring_parity_all_visible_gpus():
    n = min(torch.cuda.device_count(), 8)
    n = 1 << (n.bit_length() - 1)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "ring_check.py")]
    env = dict(os.environ)
    if n >= 4:  # hierarchical (double) ring over NCCL as well: intra-node rings of 2 (and 4 at W = 8)
        env["RING_CHECK_DOUBLE"] = "2,4" if n >= 8 else "2"
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    sys.stdout.write(res.stdout[-4000:])
    sys.stderr.write(res.stderr[-4000:])
    assert res.returncode == 0
