#!/usr/bin/env python
"""Side measurements for the configurations bench.py does not time (BASELINE.json C3, C4 sweep, C5, tracker mode) on ONE
GPU, bounded samples.  Prints one JSON object; the round's copy is profiles/r1_configs.json.  Not a bench line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "structure-from-motion-3d-reconstruction_b200"))
sys.path.insert(0, ROOT)
import sfmgpu  # noqa: E402
from sfmgpu import sched  # noqa: E402
import bench  # noqa: E402  (C4 scene helpers)

out = {}
ctx = sfmgpu.Context(0)
SECTIONS = set(sys.argv[1:]) or {"c1", "c3", "c4", "c5"}

# ---- C3: 4K, 8000 corners per frame, pair mode, resident ------------------------------------------------------------------
W, H, NF, MC = 3840, 2160, (120 if "c3" in SECTIONS else 3), 8000
cfg = sfmgpu.lkcfg(max_tracks=MC, pyr_levels=3)
frames = ctx.frames(W, H, NF, 3)
pairs = ctx.pairs(NF - 1, MC)
frames.synth(0, NF, 20261018, 0)
for _ in range(2):
    frames.build_pyramid(0, NF)
    pairs.run(frames, 0, NF - 1, cfg)
ctx.sync()
ctx.timer_start()
REP = 3
for _ in range(REP):
    frames.build_pyramid(0, NF)
    pairs.run(frames, 0, NF - 1, cfg)
ms = ctx.timer_stop() / REP
nt, nk, nit = pairs.totals()
ctx.profile(True)
pairs.run(frames, 0, NF - 1, cfg)
st = ctx.stage_times()
ctx.profile(False)
out["C3_4k_pairs"] = {"frames": NF, "pairs": NF - 1, "max_corners": MC, "ms_per_step": ms, "ms_per_pair": ms / (NF - 1),
                      "feature_tracks_per_s": nt / (ms * 1e-3), "tracks": nt, "kept_fraction": nk / max(nt, 1), "stages_ms": st}
del pairs, frames

# ---- C4: scoring sweep, 10 000 correspondences ---------------------------------------------------------------------------
xi, xj = bench.c4_points(10000)
sweep = {}
for Hh in ((1024, 2048, 4096, 8192, 16384, 32768, 65536) if "c4" in SECTIONS else (1024,)):
    E = bench.synthetic_hypotheses(Hh)
    ctx.ransac_upload(xi, xj, E)
    for _ in range(3):
        ctx.ransac_score_resident(1e-3, fetch=False)
    ctx.sync()
    ctx.timer_start()
    for _ in range(20):
        ctx.ransac_score_resident(1e-3, fetch=False)
    t = ctx.timer_stop() / 20
    sweep[str(Hh)] = {"ms": t, "hyp_pts_per_s": Hh * 10000 / (t * 1e-3)}
out["C4_sweep_10k_points"] = sweep

# ---- C5 / tracker mode: stateful KLTTracker twin, 1080p, 2000 tracks, host-fed frames ---------------------------------------
from sfmgpu import synth  # noqa: E402
NSEQ, LEN = 8, (24 if "c5" in SECTIONS else 4)
seqs = [[synth.frame(20261018 + s, t, 1920, 1080) for t in range(LEN)] for s in range(NSEQ)]
kw = dict(max_tracks=2000, min_tracks=818)
trk = ctx.tracker(**kw)
for img in seqs[0][:3]:  # warm-up (allocations)
    trk.step(img)
trk.reset(seqs[0][0])
t0 = time.perf_counter()
for img in seqs[0][1:]:
    trk.step(img)
one = time.perf_counter() - t0
n_tr = trk.totals()[0]
out["tracker_one_sequence_1080p"] = {"frames": LEN, "s": one, "steps_per_s": (LEN - 1) / one, "feature_tracks_per_s": n_tr / one}
del trk
for mode in (1, 2):  # KLT kernel forced: 1 warp per feature, 2 lane per feature (quadratic forms)
    ctx.klt_set_mode(mode)
    trk = ctx.tracker(**kw)
    for img in seqs[0][:3]:
        trk.step(img)
    trk.reset(seqs[0][0])
    t0 = time.perf_counter()
    for img in seqs[0][1:]:
        trk.step(img)
    dt = time.perf_counter() - t0
    out[f"tracker_one_sequence_klt_mode_{mode}"] = {"steps_per_s": (LEN - 1) / dt}
    del trk
ctx.klt_set_mode(0)
t0 = time.perf_counter()
res = sched.run_sequences(seqs, 0, kw, max_workers=NSEQ, lockstep=False)
par = time.perf_counter() - t0
out["C5_8_sequences_threads"] = {"sequences": NSEQ, "frames_each": LEN, "s": par, "steps_per_s": NSEQ * (LEN - 1) / par,
                                 "speedup_vs_one_by_one": one * NSEQ / par}
for nseq in (8, 32):
    # frames of step t for all sequences, staged in PINNED host memory (pageable input is limited to ~8 GB/s by the driver's
    # staging copy); the 32-sequence case reuses the 8 generated sequences four times
    stack = ctx.pinned_empty((LEN, nseq, 1080, 1920), np.uint8)
    for t in range(LEN):
        for s in range(nseq):
            stack[t, s] = seqs[s % NSEQ][t]
    mt = ctx.multitracker(nseq, 1920, 1080, **kw)
    for t in range(3):
        mt.step(stack[t])
    mt.close()
    mt = ctx.multitracker(nseq, 1920, 1080, **kw)
    mt.step(stack[0])
    t0 = time.perf_counter()
    for t in range(1, LEN):
        mt.step(stack[t])
    lock = time.perf_counter() - t0
    out[f"C5_{nseq}_sequences_lockstep"] = {"sequences": nseq, "frames_each": LEN, "s": lock, "steps_per_s": nseq * (LEN - 1) / lock,
                                            "feature_tracks_per_s": mt.totals()[0] / lock,
                                            "speedup_vs_one_by_one": one * nseq / lock,
                                            "h2d_gb_per_s": nseq * (LEN - 1) * 1920 * 1080 / lock / 1e9}
    mt.close()
    mt = ctx.multitracker(nseq, 1920, 1080, **kw)
    mt.prefetch(stack[0])
    mt.step(None, next_imgs=stack[1])
    t0 = time.perf_counter()
    for t in range(1, LEN):
        mt.step(None, next_imgs=stack[t + 1] if t + 1 < LEN else None)
    piped = time.perf_counter() - t0
    out[f"C5_{nseq}_sequences_lockstep"]["pipelined_steps_per_s"] = nseq * (LEN - 1) / piped
    out[f"C5_{nseq}_sequences_lockstep"]["pipelined_feature_tracks_per_s"] = mt.totals()[0] / piped
    mt.close()
    del stack
# ---- C1 shape: the reference's own loop (tracker.step + find_E_ransac per frame) through the C++ shim, 640 x 480 -------------------
if "c1" in SECTIONS:
    import ctypes as C
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import shimlib
    import oracle
    from conftest import TEMPLE_K
    shim = shimlib.load()
    chk, kind = oracle.best()
    NFR1, W1, H1 = 47, 640, 480
    fr1 = [synth.frame(20261018, t, W1, H1) for t in range(NFR1)]
    Kf = np.ascontiguousarray(TEMPLE_K.reshape(9))

    def run_shim(device_solver):
        shim.shim_set_device_solver(1 if device_solver else 0)
        t = shim.shim_tracker_create(2200, 900, 0.01, 8, 3, 5, 10, 1.0)
        prev, cur, ids = np.zeros((2300, 2)), np.zeros((2300, 2)), np.zeros(2300, np.int32)
        R, tt, il, kk = np.zeros(9), np.zeros(3), np.zeros(2300, np.int32), C.c_int(0)
        t_trk = t_e = 0.0
        n_e = 0
        for img in fr1:
            t0 = time.perf_counter()
            n = shimlib.ck(shim, shim.shim_tracker_step(t, img, W1, H1, prev, cur, ids, 2300))
            t1 = time.perf_counter()
            if n >= 8:
                shimlib.ck(shim, shim.shim_find_E_ransac(Kf, np.ascontiguousarray(prev[:n]), np.ascontiguousarray(cur[:n]), n, 2500, 1e-3, 60,
                                                         R, tt, il, C.byref(kk)))
                n_e += 1
            t2 = time.perf_counter()
            t_trk += t1 - t0
            t_e += t2 - t1
        shim.shim_tracker_destroy(t)
        shim.shim_set_device_solver(0)
        return {"tracker_step_ms_per_frame": 1e3 * t_trk / NFR1, "find_E_ransac_ms_per_call": 1e3 * t_e / max(n_e, 1),
                "front_end_ms_per_frame": 1e3 * (t_trk + t_e) / NFR1}

    run_shim(False)  # warm-up (allocations, first launches)
    c1 = {"frames": NFR1, "shape": [W1, H1], "max_tracks": 2200, "ransac": [2500, 1e-3, 60],
          "gpu_host_solver": run_shim(False), "gpu_device_solver": run_shim(True)}
    # the reference itself on ONE host core (it is single-threaded), first frames only
    trk = chk.tracker(max_tracks=2200, min_tracks=900)
    t_trk = t_e = 0.0
    n_e = 0
    NREF = 6
    for img in fr1[:NREF]:
        t0 = time.perf_counter()
        p, c, i = trk.step(img)
        t1 = time.perf_counter()
        if len(i) >= 8:
            chk.find_E_ransac(TEMPLE_K, p, c, 2500, 1e-3, 60)
            n_e += 1
        t2 = time.perf_counter()
        t_trk += t1 - t0
        t_e += t2 - t1
    c1["cpu_reference_one_core"] = {"kind": kind, "frames": NREF, "tracker_step_ms_per_frame": 1e3 * t_trk / NREF,
                                    "find_E_ransac_ms_per_call": 1e3 * t_e / max(n_e, 1),
                                    "front_end_ms_per_frame": 1e3 * (t_trk + t_e) / NREF}
    out["C1_shape_dropin_loop"] = c1

ctx.close()
print(json.dumps(out, indent=1))
