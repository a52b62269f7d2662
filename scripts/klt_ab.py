#!/usr/bin/env python
"""A/B timing of the KLT kernel variants on the C2 shape: python scripts/klt_ab.py [frames] [mode ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "structure-from-motion-3d-reconstruction_b200"))
import sfmgpu

nfr = int(sys.argv[1]) if len(sys.argv) > 1 else 120
modes = [int(m) for m in sys.argv[2:]] or [1, 10, 11, 12, 13, 14]
ctx = sfmgpu.Context(0)
cfg = sfmgpu.lkcfg(max_tracks=2000, pyr_levels=3)
frames = ctx.frames(1920, 1080, nfr, 3)
pairs = ctx.pairs(nfr - 1, 2000)
frames.synth(0, nfr, 20261018, 0)
frames.build_pyramid(0, nfr)
for m in modes:
    ctx.klt_set_mode(m)
    best = None
    for rep in range(4):
        ctx.profile(True)
        pairs.run(frames, 0, nfr - 1, cfg)
        st = ctx.stage_times()
        ctx.profile(False)
        best = st["klt"] if best is None else min(best, st["klt"])
    n, k, it = pairs.totals()
    print(f"mode {m:3d}: klt {best:8.3f} ms  {n / best / 1e3:8.2f} M tracks/s  ({n} tracks, {k} kept, {it} LK iterations)", flush=True)
ctx.klt_set_mode(0)
