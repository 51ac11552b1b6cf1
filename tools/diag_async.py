"""Diagnostic: are the async copies really asynchronous and do they overlap?"""
import sys, time, ctypes
sys.path.insert(0, '.')
import numpy as np, torch
from openpano_b200.capi import Engine
N = 260 * 1000 * 1000
up, cmp_, dn = Engine(0), Engine(0), Engine(0)
d_a = cmp_.dev_alloc(N); d_b = cmp_.dev_alloc(N); cmp_.sync()
t_pin = torch.empty(N, dtype=torch.uint8).pin_memory()
h_own = Engine.host_alloc(N)
h_own2 = Engine.host_alloc(N)
for name, ptr in (("torch-pinned", t_pin.data_ptr()), ("pano_host_alloc", h_own)):
    for rep in range(3):
        t0 = time.perf_counter(); up.dev_upload_async(d_a, ptr, N); t1 = time.perf_counter(); up.sync(); t2 = time.perf_counter()
    print(f"{name}: enqueue {1e3*(t1-t0):.3f} ms, complete {1e3*(t2-t0):.3f} ms -> {N/(t2-t0)/1e9:.1f} GB/s")
# overlap H2D with D2H
t0 = time.perf_counter(); up.dev_upload_async(d_a, h_own, N); dn.dev_download_async(h_own2, d_b, N); up.sync(); dn.sync(); t2 = time.perf_counter()
print(f"H2D+D2H concurrently: {1e3*(t2-t0):.3f} ms")
t0 = time.perf_counter(); dn.dev_download_async(h_own2, d_b, N); dn.sync(); t2 = time.perf_counter()
print(f"D2H alone: {1e3*(t2-t0):.3f} ms -> {N/(t2-t0)/1e9:.1f} GB/s")
