#!/usr/bin/env python
"""A/B helper: device-resident step time and top kernels of the bench workload with the
library named by PANO_B200_LIB (default: the in-tree build).
  PANO_B200_LIB=openpano_b200/_variants/x.so python tools/ab_value.py [steps]"""
import os, sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import bench
from openpano_b200.capi import Engine
from openpano_b200.stitcher import Stitcher

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
imgs, pairs, items, geom, params, mpx, pix = bench.make_workload(0, 0)
shapes = [im.shape[:2] for im in imgs]
ow, oh = max(it[2] for it in items), max(it[3] for it in items)
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    eng = Engine(0, stream.cuda_stream)
    st = Stitcher(eng, params)
    host = [torch.from_numpy(im).pin_memory() for im in imgs]
    st.upload([t.data_ptr() for t in host], shapes, (ow, oh)); eng.sync()

    def step():
        f, _ = st.run_device(pairs, items, geom, 0, want_matches=False)
        f.free()
    for _ in range(5): step()
    eng.sync()
    best = 1e9
    for trial in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps): step()
        e1.record(stream); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / steps)
    eng.profile(True); eng.profile_reset()
    for _ in range(5): step()
    prof = eng.profile_read(); eng.profile(False)
    top = sorted(prof.items(), key=lambda kv: -kv[1][1])[:6]
    print(os.environ.get("PANO_B200_LIB", "default"), f"{best:.3f} ms/step", {k: round(v[1] / 5, 4) for k, v in top}, flush=True)
    st.close(); eng.close()
