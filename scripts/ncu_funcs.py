#!/usr/bin/env python
"""Instruction / stall-sample share per source FUNCTION of an .ncu-rep (needs -lineinfo). usage: ncu_funcs.py rep file.cu"""
import csv, subprocess, sys, collections, io, re
rep, cu = sys.argv[1], sys.argv[2]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
per_line = collections.defaultdict(lambda: [0.0, 0.0])
hd = None
for r in rows:
    if "Warp Stall Sampling (All Samples)" in r and "Line No" in r:
        hd = r
        li, ci, ii = hd.index("Line No"), hd.index("Warp Stall Sampling (All Samples)"), hd.index("Instructions Executed")
        continue
    if hd is None or len(r) < len(hd) or not r[li].strip(): continue
    try: per_line[int(r[li])][0] += float(r[ii]); per_line[int(r[li])][1] += float(r[ci])
    except ValueError: pass
starts = []
for n, line in enumerate(open(cu), 1):
    if re.match(r"^(template|__device__|__global__|static|int |void |extern)", line) and "(" in line and not line.rstrip().endswith(";"):
        m = re.findall(r"(\w+)\s*\(", line)
        m = [x for x in m if x not in ("__launch_bounds__", "__align__")]
        if m: starts.append((n, m[0] if not line.startswith("__global__") else m[-1] if len(m) == 1 else m[0]))
agg = collections.defaultdict(lambda: [0.0, 0.0])
for ln, (ins, sm) in per_line.items():
    name = "?"
    for n, nm in starts:
        if n <= ln: name = nm
    agg[name][0] += ins; agg[name][1] += sm
T = sum(v[0] for v in agg.values()) or 1; S = sum(v[1] for v in agg.values()) or 1
for k, v in sorted(agg.items(), key=lambda x: -x[1][0]): print(f"{k:28s} inst {100*v[0]/T:5.1f}%  samples {100*v[1]/S:5.1f}%")
