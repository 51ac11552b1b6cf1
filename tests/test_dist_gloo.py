"""CPU, world_size 2 over gloo: the multi-GPU plumbing of openpano_b200.parallel
(image sharding, descriptor all-gather, LPT pair dealing, result gather) must
give exactly the single-process result.  The compute backend here is the oracle
(test infrastructure); on GPUs the same functions run with EngineBackend."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


class OracleBackend:
    def __init__(self):
        from tests.checker import get_checker
        self.orc = get_checker("orc")

    def detect(self, imgs):
        return [self.orc.sift_detect(im) for im in imgs]

    def match_pairs(self, descs, pairs):
        return [self.orc.match(descs[i], descs[j]) for i, j in pairs]


def _worker(rank, world, port, outdir):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from openpano_b200 import synth
    from openpano_b200.parallel import distributed_features_and_matches, shard_images
    from openpano_b200.stitcher import all_pairs

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        imgs, _ = synth.make_stack(5, 200, 150, 60, 91)
        pairs = all_pairs(len(imgs))
        mine = {k: imgs[k] for k in shard_images(len(imgs), world, rank)}
        coors, descs, matches = distributed_features_and_matches(OracleBackend(), mine, len(imgs), pairs)
        np.savez(Path(outdir) / f"rank{rank}.npz", n=np.array([len(d) for d in descs]),
                 desc_sum=np.array([float(d.astype(np.float64).sum()) for d in descs]),
                 coor0=coors[0], desc_last=descs[-1])
        if rank == 0:
            np.savez(Path(outdir) / "matches.npz", **{f"m{t}": m for t, m in enumerate(matches)})
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_pipeline_equals_single_process(orc, tmp_path):
    import torch.multiprocessing as mp
    from openpano_b200 import synth
    from openpano_b200.stitcher import all_pairs

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    imgs, _ = synth.make_stack(5, 200, 150, 60, 91)
    feats = [orc.sift_detect(im) for im in imgs]
    pairs = all_pairs(len(imgs))
    want = [orc.match(feats[i][1], feats[j][1]) for i, j in pairs]
    got = np.load(tmp_path / "matches.npz")
    for t, m in enumerate(want):
        assert np.array_equal(got[f"m{t}"], m), pairs[t]
    assert sum(len(m) for m in want) > 20
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        assert list(z["n"]) == [len(f[1]) for f in feats]
        assert np.array_equal(z["coor0"], feats[0][0])          # f64 coordinates survive the f32-word transport
        assert np.array_equal(z["desc_last"], feats[-1][1])
        assert np.allclose(z["desc_sum"], [float(f[1].astype(np.float64).sum()) for f in feats])


def test_sharding_and_dealing_logic():
    from openpano_b200.parallel import deal_pairs, shard_images
    from openpano_b200.stitcher import all_pairs, ordered_pairs

    assert shard_images(13, 8, 0) == [0, 8] and shard_images(13, 8, 5) == [5]
    assert sorted(sum((shard_images(38, 8, r) for r in range(8)), [])) == list(range(38))
    pairs = all_pairs(38)
    assert len(pairs) == 703                                   # stitcher.cc:98-100
    counts = [1000 + 37 * (k % 11) for k in range(38)]
    dealt = deal_pairs(pairs, counts, 8)
    assert sorted(sum(dealt, [])) == list(range(703))          # every task exactly once
    loads = [sum(counts[pairs[t][0]] * counts[pairs[t][1]] for t in d) for d in dealt]
    assert max(loads) / min(loads) < 1.02                      # LPT balances within 2 %
    assert deal_pairs(pairs, counts, 8) == dealt               # deterministic
    assert ordered_pairs(4) == [(0, 1), (1, 2), (2, 3), (3, 0)]  # stitcher.cc:121-122 wraps
    assert deal_pairs([], [], 4) == [[], [], [], []]
