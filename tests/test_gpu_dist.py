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


def test_cpp_host_drives_c1_c2_on_two_gpus(tmp_path):
    """No Python in the data path: a C++ program (tests/adaptor/comm_test.cc) with two host threads,
    one context + one pano_comm per GPU, runs SIFT on the owned images, pano_comm_allgather_features,
    its half of the pair list, a canvas strip and pano_comm_allgather_dev, and compares with one GPU."""
    import os
    import struct

    import numpy as np
    import torch

    from openpano_b200 import synth
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    binary = ROOT / "oracle" / "comm_test"
    if not binary.exists():
        pytest.skip("oracle/comm_test not built")
    imgs, org = synth.make_stack(5, 320, 240, 100, 53, rows=2, step_y=90)
    items, geom = synth.translation_blend_setup(org, 320, 240)
    path = tmp_path / "stack.bin"
    with open(path, "wb") as f:
        f.write(struct.pack("<3i", len(imgs), 320, 240))
        for im in imgs:
            f.write(np.ascontiguousarray(im, np.float32).tobytes())
        for it in items:
            f.write(struct.pack("<4i", *it[:4]))
            f.write(struct.pack("<9d", *it[4]))
        f.write(struct.pack("<3d", geom["res_x"], geom["proj_min_x"], geom["proj_min_y"]))
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = f"{ROOT / 'openpano_b200'}:" + env.get("LD_LIBRARY_PATH", "")
    out = subprocess.run([str(binary), str(path)], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert "COMM TEST OK" in out.stdout
