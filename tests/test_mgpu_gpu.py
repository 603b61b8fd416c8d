"""Launches tests/mgpu_check.py under torchrun when the box has >= 2 GPUs (one process per GPU, NCCL): the tile-sharded
step, the autograd-level sharded rasterizer and the distributed mapper against single-GPU runs, with both exchange
paths (NVLink peer memory and NCCL reduce-scatter)."""
import os
import subprocess
import sys

import pytest
import torch

from util import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
def test_sharded_step_matches_single_gpu(exchange):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    env = dict(os.environ, EGS_EXCHANGE=exchange)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", "29517" if exchange == "peer" else "29518",
           os.path.join(ROOT, "tests", "mgpu_check.py"), "C5", "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for level in ("raw", "autograd", "mapper"):
        assert "MGPU_CHECK OK " + level in r.stdout, r.stdout[-3000:]
