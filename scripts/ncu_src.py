#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics + the hottest source lines (needs -lineinfo). usage: ncu_src.py rep [topN]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, v = rows[0], rows[-1]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
for k in want:
    if k in h: print(f"{k:80s} {v[h.index(k)]}  [{rows[1][h.index(k)]}]")
for i, k in enumerate(h):
    if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
        try:
            if float(v[i]) > 0.15: print(f"{k:80s} {v[i]}")
        except ValueError: pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
hd = None
for r in rows:
    if "Warp Stall Sampling (All Samples)" in r and "Line No" in r:
        hd = r
        li, si, ci = hd.index("Line No"), hd.index("Source"), hd.index("Warp Stall Sampling (All Samples)")
        bi, ii = hd.index("stall_barrier"), hd.index("Instructions Executed")
        continue
    if hd is None or len(r) < len(hd): continue
    try: sm = float(r[ci]); ba = float(r[bi]); ie = float(r[ii])
    except ValueError: continue
    if not r[li].strip(): continue  # SASS rows repeat their CUDA line's counts
    k = (r[li], r[si].strip())
    agg[k][0] += sm; agg[k][1] += ba; agg[k][2] += ie
tot = sum(v[0] for v in agg.values()) or 1
totb = sum(v[1] for v in agg.values())
print(f"--- samples total {tot:.0f}, of which barrier stalls {totb:.0f}; top lines by non-barrier samples, then by barrier")
for (ln, txt), (sm, ba, ie) in sorted(agg.items(), key=lambda x: -(x[1][0] - x[1][1]))[:top]:
    print(f"{100*(sm-ba)/tot:5.1f}% nb {100*ba/tot:5.1f}% bar inst={ie:12.0f} L{ln:>4s} {txt[:120]}")
toti = sum(v[2] for v in agg.values()) or 1
print(f"--- top lines by warp instructions executed (total {toti:.3g})")
for (ln, txt), (sm, ba, ie) in sorted(agg.items(), key=lambda x: -x[1][2])[:top]:
    print(f"{100*ie/toti:5.1f}% inst  {100*sm/tot:5.1f}% samp L{ln:>4s} {txt[:110]}")
print("--- top lines by barrier-stall samples")
for (ln, txt), (sm, ba, ie) in sorted(agg.items(), key=lambda x: -x[1][1])[:8]:
    print(f"{100*ba/tot:5.1f}% bar L{ln:>4s} {txt[:120]}")
