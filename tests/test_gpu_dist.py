"""GPU, world_size 2 over NCCL: the sharded hot path (SIFT per rank, descriptor
all-gather, dealt pair tasks, strip blend + gather) must reproduce the one-GPU
match lists and mosaic bit for bit.  Skipped on a single-GPU box."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("bands", [0, 3])
def test_two_rank_nccl_equals_one_gpu(bands):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(29517 + bands), str(ROOT / "tools" / "run_dist.py"), "small", "2", str(bands)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["matches_identical"] and res["mosaic_identical"], res
