#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares of one kernel in an ncu report.
  python tools/ncu_lines.py report.ncu-rep kernel_substring [top_n]"""
import collections, csv, subprocess, sys
rep, kname = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "-k", f"regex:{kname}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr = None, None
agg = collections.defaultdict(lambda: [0, 0, 0, ""])
seen_kernel = 0
for r in rows:
    if not r: continue
    if r[0] == "Function Name":
        seen_kernel += 1
        continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; ie = hdr.index("Instructions Executed"); sm = hdr.index("# Samples"); ti = hdr.index("Thread Instructions Executed"); continue
    if hdr is None or len(r) <= ie: continue
    try: n, s, t = int(r[ie]), int(r[sm]), int(r[ti])
    except ValueError: continue
    if not r[0]: continue          # SASS rows carry no line number: count CUDA rows only
    a = agg[(cur, r[0])]; a[0] += n; a[1] += s; a[2] += t; a[3] = r[1]
tot = sum(a[0] for a in agg.values()) or 1; tots = sum(a[1] for a in agg.values()) or 1
print(f"total warp-instr {tot}  samples {tots}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{a[0]/tot*100:5.1f}% inst {a[1]/tots*100:5.1f}% smp act={a[2]/max(a[0],1):4.1f} | {k[0]}:{k[1]} {a[3].strip()[:100]}")
