#!/usr/bin/env python
"""Prints the interesting numbers of a bench.py JSON line (development helper)."""
import json
import sys

d = json.load(open(sys.argv[1]))
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), "e2e ms", round(d["e2e"]["ms_per_step"], 3))
print("roofline", d["roofline"])
print("cpu", d["cpu_baseline"])
for k, v in list(d["kernels"].items())[:14]:
    print(f'{k:22s} {v.get("avg_ms", 0):.4f} ms x{v["launches_per_step"]:.0f} share {v["share"]:.3f} frac {v.get("frac")}')
for k, v in (d.get("configs") or {}).items():
    print("=====", k)
    if "error" in v:
        print(v)
        continue
    if "sizes" in v:
        for n, s in v["sizes"].items():
            print(n, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in s.items()})
        print(v["cpu_baseline"])
        continue
    print({kk: v[kk] for kk in ("images", "pairs", "bands", "canvas_wh", "gen_s", "ms_device", "value", "features", "matches",
                                "match_rows_rescanned_exactly")})
    print("e2e", v["e2e"]["ms"], "roofline", v["roofline"]["kernel"], v["roofline"]["frac"], "parity", v["parity_sample"])
    print("cpu", v["cpu_baseline"])
    for kk, vv in v["kernels"].items():
        print(f'   {kk:22s} {vv["ms_per_step"]:.4f} share {vv["share"]:.3f} frac {vv.get("frac")}')
if d.get("sharded"):
    print(json.dumps(d["sharded"], indent=1))
