"""Diagnostic: per-step time distribution of the device-resident loop and host-side split."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import bench
from openpano_b200.capi import Engine
from openpano_b200.stitcher import Stitcher
imgs, pairs, items, geom, params, mpx = bench.make_workload(0, 0)
shapes = [im.shape[:2] for im in imgs]
ow, oh = max(it[2] for it in items), max(it[3] for it in items)
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    eng = Engine(0, stream.cuda_stream)
    st = Stitcher(eng, params)
    host = [torch.from_numpy(im).pin_memory() for im in imgs]
    st.upload([t.data_ptr() for t in host], shapes, (ow, oh)); eng.sync()
    T = time.perf_counter
    def step(split=None):
        t0 = T(); ptrs = st.image_ptrs()
        fs = eng.sift_detect_batch_ptr(ptrs, [s[1] for s in shapes], [s[0] for s in shapes], params, device=True); t1 = T()
        eng.match_pairs_dev(fs, pairs, params); t2 = T()
        eng.blend_dev(ptrs, shapes, items, geom, st._d_out, ow, oh, 0, params); t3 = T()
        fs.free(); t4 = T()
        if split is not None: split.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3))
    for _ in range(5): step()
    eng.sync()
    for trial in range(3):
        sp = []; t0 = T()
        for _ in range(50): step(sp)
        eng.sync(); dt = (T() - t0) / 50
        a = np.array(sp) * 1e3
        print(f"trial {trial}: {dt*1e3:.3f} ms/step | host ms median: sift {np.median(a[:,0]):.3f} match {np.median(a[:,1]):.3f} blend {np.median(a[:,2]):.3f} free {np.median(a[:,3]):.3f} | match max {a[:,1].max():.2f} p90 {np.percentile(a[:,1],90):.2f}")
