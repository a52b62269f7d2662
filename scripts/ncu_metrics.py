#!/usr/bin/env python
"""Headline metrics of .ncu-rep captures -> one JSON (per kernel launch).  usage: ncu_metrics.py out.json rep [rep ...]"""
import csv, io, json, subprocess, sys
WANT = {
    "gpu__time_duration.sum": "duration",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__occupancy_limit_registers": "occ_limit_regs",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "lts__t_bytes.sum": "l2_bytes",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pipe_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "usecond": 1, "msecond": 1e3, "nsecond": 1e-3}
out = {}
for rep in sys.argv[2:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        continue
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[h.index("Kernel Name")]
        d = {"capture": rep.split("/")[-1]}
        for k, short in WANT.items():
            if k in h:
                i = h.index(k)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                v *= UNIT.get(units[i], 1)
                d[short + ("_us" if short == "duration" else "_bytes" if short.startswith(("dram_r", "dram_w", "l2_")) else "")] = v
        stalls = {}
        for i, k in enumerate(h):
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v > 0.15:
                    stalls[k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")] = round(v, 2)
        d["stall_cycles_per_issue"] = stalls
        key = name
        n = 2
        while key in out:
            key = f"{name} #{n}"
            n += 1
        out[key] = d
json.dump(out, open(sys.argv[1], "w"), indent=1)
for k, d in out.items():
    print(f"{k[:60]:60s} {d.get('duration_us', 0):10.1f} us  dram {(d.get('dram_read_bytes', 0) + d.get('dram_write_bytes', 0)) / 1e6:9.1f} MB  issue {d.get('issue_active_pct', 0):5.1f}%  fp64 {d.get('fp64_pipe_pct', 0):5.1f}%  regs {d.get('registers', 0):.0f}")
