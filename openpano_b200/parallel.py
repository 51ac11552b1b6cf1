"""Multi-GPU sharding of the hot path: one process per GPU.

The reference is a single process whose parallel axes are independent units —
images in calc_feature (stitcherbase.cc:14), image pairs in pairwise_match
(stitcher.cc:106).  Across GPUs the same axes are sharded:

  SIFT           image k -> rank k mod G                       no collective
  exchange (C1)  every pair needs both descriptor sets         all-gather of
                 (matcher.cc:96-101)                            {count, desc, coor}
  matching       the task list of stitcher.cc:98-100 / :121-122 dealt by
                 descending N_i*N_j (longest-processing-time)   no collective
  results        match pairs are tiny                           gather to rank 0

Collectives go through torch.distributed (NCCL on GPUs, gloo in the CPU tests).
The compute is a *backend* object so that the same plumbing is exercised on CPU
(tests pass an oracle-backed backend) and on GPUs (EngineBackend below).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


# ----------------------------------------------------------------------------- pure sharding logic
def shard_images(n_images: int, world: int, rank: int) -> List[int]:
    """Image k is owned by rank k mod world."""
    return [k for k in range(n_images) if k % world == rank]


def deal_pairs(pairs: Sequence[Tuple[int, int]], counts: Sequence[int], world: int) -> List[List[int]]:
    """Longest-processing-time assignment of pair tasks by cost N_i*N_j.
    Returns, per rank, the indices into `pairs` it matches (deterministic)."""
    order = sorted(range(len(pairs)), key=lambda t: (-counts[pairs[t][0]] * counts[pairs[t][1]], t))
    load = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for t in order:
        r = min(range(world), key=lambda q: (load[q], q))
        out[r].append(t)
        load[r] += counts[pairs[t][0]] * counts[pairs[t][1]] + 1
    for lst in out:
        lst.sort()
    return out


# ----------------------------------------------------------------------------- collectives
def _dist():
    import torch.distributed as dist
    return dist


def all_gather_features(local: dict, n_images: int, device=None):
    """C1: all-gather of the per-image descriptor blocks.

    local: {image index: (coor float64 [n,2], desc float32 [n,128])} owned by this
    rank.  Returns the full lists (coors, descs) for images 0..n_images-1 on every
    rank.  Variable sizes: counts first, then one padded all_gather."""
    import torch
    dist = _dist()
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = device if device is not None else torch.device("cpu")
    counts = torch.zeros(n_images, dtype=torch.int64, device=dev)
    for k, (_, d) in local.items():
        counts[k] = len(d)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    counts_h = [int(c) for c in counts.cpu()]
    # per-rank payload: images it owns, concatenated in image order, padded to the largest rank payload
    owned = [shard_images(n_images, world, r) for r in range(world)]
    rows = [sum(counts_h[k] for k in owned[r]) for r in range(world)]
    pad = max(max(rows), 1)
    mine = torch.zeros((pad, 128 + 4), dtype=torch.float32, device=dev)   # 128 desc + 2 f64 coords as 4 f32 words
    off = 0
    for k in owned[rank]:
        c, d = local[k]
        n = len(d)
        if n:
            mine[off:off + n, :128] = torch.from_numpy(np.ascontiguousarray(d, np.float32)).to(dev)
            cw = torch.from_numpy(np.ascontiguousarray(c, np.float64).view(np.float32).reshape(n, 4)).to(dev)
            mine[off:off + n, 128:] = cw
        off += n
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    coors, descs = [None] * n_images, [None] * n_images
    for r in range(world):
        g = gathered[r].cpu().numpy()
        off = 0
        for k in owned[r]:
            n = counts_h[k]
            descs[k] = np.ascontiguousarray(g[off:off + n, :128])
            coors[k] = np.ascontiguousarray(g[off:off + n, 128:]).view(np.float64).reshape(n, 2)
            off += n
    return coors, descs


def gather_matches(my_tasks: Sequence[int], my_results: Sequence[np.ndarray], n_tasks: int):
    """Gather per-pair match arrays to rank 0 (returns the full list there, None elsewhere)."""
    dist = _dist()
    payload = [(int(t), np.ascontiguousarray(m, np.int32)) for t, m in zip(my_tasks, my_results)]
    out = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(payload, out, dst=0)
    if dist.get_rank() != 0:
        return None
    full = [None] * n_tasks
    for part in out:
        for t, m in part:
            full[t] = m
    return full


# ----------------------------------------------------------------------------- driver
def distributed_features_and_matches(backend, images: dict, n_images: int, pairs, device=None):
    """images: {image index: HxWx3 float32} for the images this rank owns
    (shard_images).  Every rank returns (coors, descs) for ALL images; rank 0 also
    gets the per-pair match arrays in task order (others get None)."""
    dist = _dist()
    world, rank = dist.get_world_size(), dist.get_rank()
    mine = sorted(images)
    assert mine == shard_images(n_images, world, rank), "images must follow shard_images()"
    feats = backend.detect([images[k] for k in mine])
    local = {k: f for k, f in zip(mine, feats)}
    coors, descs = all_gather_features(local, n_images, device)
    counts = [len(d) for d in descs]
    tasks = deal_pairs(pairs, counts, world)[rank]
    results = backend.match_pairs(descs, [pairs[t] for t in tasks])
    matches = gather_matches(tasks, results, len(pairs))
    return coors, descs, matches


class EngineBackend:
    """Compute backend on one GPU through the C ABI (openpano_b200.capi.Engine)."""

    def __init__(self, engine, params=None):
        from ._abi import default_params
        self.eng = engine
        self.params = params or default_params()

    def detect(self, imgs):
        if not imgs:
            return []
        fs = self.eng.sift_detect_batch(imgs, self.params)
        try:
            return [fs.download(i) for i in range(len(imgs))]
        finally:
            fs.free()

    def match_pairs(self, descs, pairs):
        if not pairs:
            return []
        used = sorted({i for p in pairs for i in p})
        remap = {k: q for q, k in enumerate(used)}
        fs = self.eng.featureset_upload([descs[k] for k in used])
        try:
            return self.eng.match_pairs(fs, [(remap[i], remap[j]) for i, j in pairs], self.params)
        finally:
            fs.free()
