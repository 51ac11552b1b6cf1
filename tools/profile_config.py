#!/usr/bin/env python
"""Event-timed kernel breakdown of one stage on one BASELINE config.
  python tools/profile_config.py unordered38 match | uav16 multiband"""
import json, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from openpano_b200 import synth
from openpano_b200._abi import default_params
from openpano_b200.capi import Engine
from openpano_b200.stitcher import Stitcher, all_pairs

cfg, stage = sys.argv[1], sys.argv[2]
eng = Engine(0)
if cfg == "unordered38":
    imgs, org = synth.config_stack("unordered_38x1300x867"); params = default_params(); pairs = all_pairs(38); mo = None
else:
    imgs, org = synth.config_stack("uav_64x4000x3000", n=16); params = default_params(multiband=5, lazy_read=0)
    pairs = [(i, i + 1) for i in range(15)]; mo = 8000
h, w = imgs[0].shape[:2]
items, geom = synth.translation_blend_setup(org, w, h, mo)
ow, oh = max(it[2] for it in items), max(it[3] for it in items)
st = Stitcher(eng, params)
shapes = [im.shape[:2] for im in imgs]
st.upload([im.ctypes.data for im in imgs], shapes, (ow, oh)); eng.sync()
ptrs = st.image_ptrs()
fs = eng.sift_detect_batch_ptr(ptrs, [s[1] for s in shapes], [s[0] for s in shapes], params, device=True)
def work():
    if stage == "match":
        return eng.match_pairs_dev(fs, pairs, params)
    return eng.blend_dev(ptrs, shapes, items, geom, st._d_out, ow, oh, 5 if stage == "multiband" else 0, params)
work(); eng.sync()
t = time.perf_counter(); r = work(); eng.sync(); wall = (time.perf_counter() - t) * 1e3
eng.profile(True); eng.profile_reset(); work(); prof = eng.profile_read(); eng.profile(False)
print(json.dumps({"config": cfg, "stage": stage, "wall_ms": round(wall, 3), "result": r if isinstance(r, int) else None,
                  "exact_rows": eng.match_last_exact_rows(),
                  "kernels": {k: [v[0], round(v[1], 4)] for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}}))
