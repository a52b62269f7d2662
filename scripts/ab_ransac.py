#!/usr/bin/env python
"""A/B timing of the RANSAC scoring kernel on the C4 shape (65,536 x 10,000) and a two-view-sized batch; run once per library
build: SFMGPU_LIB=<path> python scripts/ab_ransac.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "structure-from-motion-3d-reconstruction_b200"))
sys.path.insert(0, ROOT)
import sfmgpu
from bench import c4_points, RS_N, RS_H, temple_K

ctx = sfmgpu.Context(0)
xi, xj = c4_points(RS_N)
idx8 = ctx.ransac_sample(RS_N, RS_H * 8).reshape(RS_H, 8)
ctx.ransac_hypotheses(xi, xj, idx8, fetch=False)
for _ in range(3):
    ctx.ransac_score_resident(1e-3, fetch=False)
ctx.sync()
ctx.timer_start()
for _ in range(20):
    ctx.ransac_score_resident(1e-3, fetch=False)
ms = ctx.timer_stop() / 20
bh, bn = ctx.ransac_score_resident(1e-3)
print(f"{os.environ.get('SFMGPU_LIB', 'default')}: C4 scoring {ms:.3f} ms = {RS_H * RS_N / ms / 1e6:.1f} G hyp*pts/s, winner {bh} with {bn}")
# batched: 256 pairs x 4000 hypotheses x 7000 points (C3-like)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import two_view_scene
P, n, cap = 256, 7000, 8000
pi, pj = two_view_scene(n, seed=5)
pairs = ctx.pairs(P, cap)
pairs.set_matches([pi] * P, [pj] * P)
for rep in range(2):
    ctx.sync()
    t0 = time.perf_counter()
    pairs.ransac(temple_K(), 4000, 2e-3, 80, 120)
    ctx.sync()
    dt = time.perf_counter() - t0
print(f"  batched stage {P} x 4000 x {n}: {dt * 1e3:.1f} ms")
st, bhh, inl, E, R, t = pairs.ransac_download(3)
print("  pair 3:", st, bhh, len(inl))
# solver alone: C4 hypotheses, then a C2-like batch (999 pairs x 4000 hypotheses x 1750 points)
ctx.sync()
ctx.timer_start()
for _ in range(5):
    ctx.ransac_hypotheses(xi, xj, idx8, fetch=False)
print(f"  solver {RS_H} hypotheses: {ctx.timer_stop() / 5:.3f} ms")
P, n, cap = 999, 1750, 2000
pi, pj = two_view_scene(n, seed=6)
pairs2 = ctx.pairs(P, cap)
pairs2.set_matches([pi] * P, [pj] * P)
for rep in range(2):
    ctx.sync()
    t0 = time.perf_counter()
    pairs2.ransac(temple_K(), 4000, 2e-3, 80, 120)
    ctx.sync()
    dt = time.perf_counter() - t0
print(f"  batched stage {P} x 4000 x {n}: {dt * 1e3:.1f} ms")
