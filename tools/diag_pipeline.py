"""Diagnostic: host-side timeline of PipelinedStitcher."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import bench
from openpano_b200.capi import Engine
from openpano_b200.stitcher import PipelinedStitcher

imgs, pairs, items, geom, params, mpx = bench.make_workload(0, 0)
shapes = [im.shape[:2] for im in imgs]
ow, oh = max(it[2] for it in items), max(it[3] for it in items)
host = [torch.from_numpy(im).pin_memory() for im in imgs]
ptrs = [t.data_ptr() for t in host]
outs = [torch.empty((oh, ow, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
ps = PipelinedStitcher(0, params, depth=2)
T = time.perf_counter

def timed_run(k, out_ptr):
    s = ps.slots[k]; sh = s["shapes"]; p = [s["imgs"] + o for o in s["offs"]]
    t0 = T(); ps.cmp.event_wait(s["ev_up"])
    fs = ps.cmp.sift_detect_batch_ptr(p, [q[1] for q in sh], [q[0] for q in sh], params, device=True); t1 = T()
    m = ps.cmp.match_pairs(fs, pairs, params); t2 = T()
    ps.cmp.event_wait(s["ev_dn"]); ps.cmp.blend_dev(p, sh, items, geom, s["out"], ow, oh, 0, params); ps.cmp.event_record(s["ev_cmp"]); t3 = T()
    fs.free(); ps.dn.event_wait(s["ev_cmp"]); ps.dn.dev_download_async(out_ptr, s["out"], ow * oh * 12); ps.dn.event_record(s["ev_dn"]); s["busy"] = True; t4 = T()
    return (k, m), (t1 - t0, t2 - t1, t3 - t2, t4 - t3)

for trial in range(2):
    slot = ps.stage(ptrs, shapes, (ow, oh)); pending = None
    t_all = T()
    for i in range(8):
        a = T(); nxt = ps.stage(ptrs, shapes, (ow, oh)); b = T()
        job, parts = timed_run(slot, outs[i & 1].data_ptr()); c = T()
        if pending: ps.wait(pending)
        d = T()
        if trial: print(f"i={i} stage {1e3*(b-a):6.2f}  run {1e3*(c-b):6.2f} [sift {1e3*parts[0]:5.2f} match {1e3*parts[1]:5.2f} blend {1e3*parts[2]:5.2f} dn {1e3*parts[3]:5.2f}] wait {1e3*(d-c):5.2f}")
        pending, slot = job, nxt
    ps.wait(pending)
    print("avg per job ms", 1e3 * (T() - t_all) / 8)
ps.close()
