"""Where a single find_E_ransac-sized call spends its time (2200 points, 2500 hypotheses)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "structure-from-motion-3d-reconstruction_b200")); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sfmgpu, oracle
from conftest import TEMPLE_K, two_view_scene
chk, _ = oracle.best()
ctx = sfmgpu.Context(0)
pi, pj = two_view_scene(2200, seed=2200)
xi, xj = chk.norm_points(TEMPLE_K, pi), chk.norm_points(TEMPLE_K, pj)
idx = ctx.ransac_sample(2200, 2500 * 8).reshape(2500, 8)
def t(f, reps=30):
    f(); ctx.sync()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    ctx.sync()
    return (time.perf_counter() - t0) / reps * 1e3
for mode in (3, 1, 0):
    ctx.solver_set_mode(mode)
    print("mode", mode, "solve_score %.3f ms" % t(lambda: ctx.ransac_solve_score(xi, xj, idx, 1e-3)))
    print("   hypotheses resident (upload + solve, no sync) %.3f ms" % t(lambda: ctx.ransac_hypotheses(xi, xj, idx, fetch=False)))
    ctx.ransac_hypotheses(xi, xj, idx, fetch=False)
    print("   score_resident (count, argmax, winner, mask, 8-byte read-back) %.3f ms" % t(lambda: ctx.ransac_score_resident(1e-3)))
    print("   score_resident without read-back %.3f ms" % t(lambda: ctx.ransac_score_resident(1e-3, fetch=False)))

print("sync only %.3f ms" % t(lambda: ctx.sync()))
