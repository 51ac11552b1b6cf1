#!/usr/bin/env python
"""Runs the other BASELINE.json configs through the engine on one GPU and prints a
JSON line per config (stage times, counts, sanity checks).  These are parity /
capacity cases, not the bench line (bench.py measures configs[1]).

  python tools/run_configs.py cmu0 unordered38 uav64 sweep
"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from openpano_b200 import synth  # noqa: E402
from openpano_b200._abi import default_params  # noqa: E402
from openpano_b200.capi import Engine  # noqa: E402
from openpano_b200.stitcher import Stitcher, all_pairs, ordered_pairs  # noqa: E402


def timed(eng, fn):
    eng.sync()
    t = time.perf_counter()
    r = fn()
    eng.sync()
    return r, (time.perf_counter() - t) * 1e3


def run_stack(eng, name, cfg_name, pairs_fn, bands, params, max_output=None, n=None, check_pairs=2):
    t0 = time.perf_counter()
    imgs, org = synth.config_stack(cfg_name, n=n)
    gen_s = time.perf_counter() - t0
    h, w = imgs[0].shape[:2]
    items, geom = synth.translation_blend_setup(org, w, h, max_output)
    pairs = pairs_fn(len(imgs))
    st = Stitcher(eng, params)
    ow, oh = max(it[2] for it in items), max(it[3] for it in items)
    hp = [im.ctypes.data for im in imgs]
    shapes = [im.shape[:2] for im in imgs]
    st.upload(hp, shapes, (ow, oh))
    eng.sync()
    ptrs = st.image_ptrs()
    res = {"config": name, "images": len(imgs), "wh": [w, h], "pairs": len(pairs), "canvas": [ow, oh],
           "gen_s": round(gen_s, 1)}
    for rep in range(2):            # second pass = warm allocator
        fs, t_sift = timed(eng, lambda: eng.sift_detect_batch_ptr(ptrs, [s[1] for s in shapes], [s[0] for s in shapes],
                                                                  params, device=True))
        counts = [fs.count(i) for i in range(len(imgs))]
        m, t_match = timed(eng, lambda: eng.match_pairs(fs, pairs, params))
        _, t_lin = timed(eng, lambda: eng.blend_dev(ptrs, shapes, items, geom, st._d_out, ow, oh, 0, params))
        t_mb = None
        if bands:
            _, t_mb = timed(eng, lambda: eng.blend_dev(ptrs, shapes, items, geom, st._d_out, ow, oh, bands, params))
        if rep == 0:
            fs.free()
    res.update(features=int(sum(counts)), feat_min=int(min(counts)), feat_max=int(max(counts)),
               matches=int(sum(len(x) for x in m)), exact_rows=None,
               ms={"sift": round(t_sift, 3), "match": round(t_match, 3), "linear_blend": round(t_lin, 3),
                   "multiband": None if t_mb is None else round(t_mb, 3)})
    mpx = sum(s[0] * s[1] for s in shapes) / 1e6
    tot = t_sift + t_match + (t_mb if t_mb is not None else t_lin)
    res["mpx_per_s_device_resident"] = round(mpx / (tot / 1e3), 1)
    # spot-check a few pairs against the oracle (exact fp32 rule)
    try:
        from tests.checker import get_checker
        orc = get_checker("orc")
        ok = True
        for (i, j), got in list(zip(pairs, m))[:check_pairs]:
            di, dj = fs.download(i)[1], fs.download(j)[1]
            ok = ok and np.array_equal(got, orc.match(di, dj, params))
        res["oracle_pairs_ok"] = bool(ok)
    except Exception as ex:  # checker optional
        res["oracle_pairs_ok"] = repr(ex)
    out = np.empty((oh, ow, 3), np.float32)
    eng.dev_download(out, st._d_out)
    res["canvas_covered"] = round(float((out[..., 0] >= 0).mean()), 4)
    fs.free()
    st.close()
    print(json.dumps(res), flush=True)


def run_sweep(eng, sizes):
    params = default_params()
    for n in sizes:
        rng = np.random.RandomState(4)
        a = synth.rootsift_like(n, 4)
        perm = rng.permutation(n)
        b = a[perm].copy()
        half = n // 2
        b[:half] += rng.randn(half, 128).astype(np.float32) * 10.0          # ~50 % keep a true mutual match
        b[half:] = synth.rootsift_like(n - half, 5)
        fs = eng.featureset_upload([a, b])
        for rep in range(2):
            tot, t = timed(eng, lambda: eng.match_pairs_dev(fs, [(0, 1)], params))
        fs.free()
        print(json.dumps({"config": "sweep", "n": n, "matches": tot, "ms": round(t, 3),
                          "exact_rows": eng.match_last_exact_rows(),
                          "tflops_algorithmic": round(2.0 * n * n * 128 / (t * 1e-3) / 1e12, 2)}), flush=True)


def main():
    which = sys.argv[1:] or ["cmu0", "unordered38"]
    eng = Engine(0)
    if "cmu0" in which:
        run_stack(eng, "1: 8x600x400 (cylinder-mode set, hot-path stages only)", "cmu0_8x600x400", ordered_pairs, 0,
                  default_params(ordered_input=1))
    if "unordered38" in which:
        run_stack(eng, "3: 38x1300x867 unordered, all pairs", "unordered_38x1300x867", all_pairs, 0, default_params(),
                  check_pairs=3)
    if "uav64" in which:
        run_stack(eng, "5: 64x4000x3000 UAV, multiband 5, MAX_OUTPUT_SIZE 8000", "uav_64x4000x3000",
                  lambda n: [(i, i + 1) for i in range(n - 1)], 5, default_params(multiband=5, lazy_read=0),
                  max_output=8000, check_pairs=1)
    if "uav16" in which:
        run_stack(eng, "5 (reduced): 16x4000x3000 UAV, multiband 5", "uav_64x4000x3000",
                  lambda n: [(i, i + 1) for i in range(n - 1)], 5, default_params(multiband=5, lazy_read=0),
                  max_output=8000, n=16, check_pairs=1)
    if "sweep" in which:
        run_sweep(eng, [10000, 50000, 100000])
    if "sweep500k" in which:
        run_sweep(eng, [500000])
    eng.close()


if __name__ == "__main__":
    main()
