#!/usr/bin/env python
"""bench.py — the front-end hot path on synthetic sequences, one JSON line on stdout.

Workload at N GPUs (BASELINE.json configs[1], "C2"): every rank owns ONE synthetic 1080p 1,000-frame sequence
(seed 20261018 + rank, integer generator of sfmgpu/synth.py) and runs the stateless two-view front end
(cpp/src/templering_sfm.cpp:1836-1857) on all 999 frame pairs: 3-level pyramid for every frame, Shi-Tomasi +
exact std::sort + greedy NMS (2000 corners/frame), forward+backward pyramidal LK with the fb test.  One "step" =
one pass over the whole sequence.  Sequences are independent, so ranks shard with no data-path collective
(weak scaling, SURVEY.md §8e mode 2); tracks and counts are gathered to rank 0 with NCCL at the end of a step.

metric  : KLT feature-tracks/s (1 feature-track = fwd + bwd track_one + fb test of one corner);
          the RANSAC half of BASELINE.json's metric (hyp x pts / s, config C4) is reported under "ransac".
value   : inputs resident in HBM, CUDA events on the library's stream, max over ranks.
e2e     : same step through the C ABI with HOST (pinned) frames: H2D of the sequence and D2H of all tracks inside
          the timed region.
--impl reference : the reference's own CPU front end (oracle/_ref = the unmodified TU compiled where it lies,
          else the oracle port) on all host cores, bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "structure-from-motion-3d-reconstruction_b200"))

W, H, NFRAMES, MAX_CORNERS, LEVELS = 1920, 1080, 1000, 2000, 3
SEED0 = 20261018
METRIC = "KLT feature-tracks/sec (+ RANSAC hyp*pts/sec under 'ransac')"
UNIT = "feature-tracks/s"
B_KLT = 2092.0      # algorithmic bytes per feature-track
F_KLT_IT = 5045.0   # algorithmic FP64 flop per LK iteration
F_RS = 35.0         # flop per hyp x pt
RS_N, RS_H = 10000, 65536


def workload_cfg(n_gpus, nframes):
    return {
        "workload": f"{'C2' if W == 1920 else 'C3'}: synthetic {W}x{H} {nframes}-frame sequence per GPU, pair-mode front end "
                    f"(pyramid L=3 + Shi-Tomasi/NMS {MAX_CORNERS} corners/frame + fwd/bwd KLT r=5 iters=10 + fb<1.0)",
        "frames_per_gpu": nframes, "pairs_per_gpu": nframes - 1, "width": W, "height": H, "max_corners": MAX_CORNERS,
        "pyr_levels": LEVELS, "sharding": f"sequence-per-rank x{n_gpus} (no data-path collective; NCCL gather of results)",
        "l2_policy": "inputs (2.07 GB of frames per GPU) exceed the 126 MB L2; no explicit flush needed",
        "ransac_workload": f"C4: {RS_H} hypotheses x {RS_N} correspondences, thr 1e-3",
    }


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.p, self.lines = index, None, []

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def bind_near_gpu(index):
    """Pin this rank (and therefore its first-touch pinned allocations) to the CPUs NVML reports as local to its GPU: with
    8 ranks uploading 2 GB per step each, remote-socket host memory is the first bottleneck.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {i * 64 + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def profile_metrics():
    """ncu-derived numbers of the committed captures (profiles/r1_metrics.json, C2 shape): static evidence, not re-measured."""
    p = os.path.join(ROOT, "profiles", "r1_metrics.json")
    try:
        with open(p) as fh:
            m = json.load(fh)
    except Exception:
        return {}
    out = {}
    for name, d in m.items():
        for short in ("klt_quad_kernel", "score_tile_kernel<2>", "radix_sort_frame_kernel", "nms_kernel", "ransac_count_kernel", "pyr_down_kernel"):
            if short in name and short not in out:
                out[short] = d
    return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json (of measured)"
        except Exception:
            pass
    return 6650.0, "B200_PROFILING.md fallback (of fallback)"


def synthetic_hypotheses(H, seed=5):
    """Essential matrices [t]x R around the C4 scene's true motion (scoring cost does not depend on the values)."""
    rng = np.random.default_rng(seed)
    w = np.array([0.02, -0.15, 0.01]) + rng.normal(0, 0.05, (H, 3))
    t = np.array([0.2, 0.01, 0.03]) + rng.normal(0, 0.05, (H, 3))
    th = np.linalg.norm(w, axis=1, keepdims=True)
    k = w / th
    Kx = np.zeros((H, 3, 3))
    Kx[:, 0, 1], Kx[:, 0, 2], Kx[:, 1, 0], Kx[:, 1, 2], Kx[:, 2, 0], Kx[:, 2, 1] = -k[:, 2], k[:, 1], k[:, 2], -k[:, 0], -k[:, 1], k[:, 0]
    R = np.eye(3) + np.sin(th)[:, :, None] * Kx + (1 - np.cos(th))[:, :, None] * (Kx @ Kx)
    t = t / np.linalg.norm(t, axis=1, keepdims=True)
    Tx = np.zeros((H, 3, 3))
    Tx[:, 0, 1], Tx[:, 0, 2], Tx[:, 1, 0], Tx[:, 1, 2], Tx[:, 2, 0], Tx[:, 2, 1] = -t[:, 2], t[:, 1], t[:, 2], -t[:, 0], -t[:, 1], t[:, 0]
    return np.ascontiguousarray((Tx @ R).reshape(H, 9))


def c4_points(n):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import TEMPLE_K, two_view_scene
    pi, pj = two_view_scene(n)
    Kinv = np.linalg.inv(TEMPLE_K)
    def norm(p):
        hp = np.concatenate([p, np.ones((len(p), 1))], 1) @ Kinv.T
        return np.ascontiguousarray(hp[:, :2] / hp[:, 2:3])
    return norm(pi), norm(pj)


# ---------------------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """The reference's CPU front end on all host cores (rank 0 only), bounded sample of C2."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    chk, kind = oracle.best()
    gen = oracle.port()
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, 64))
    nfr = threads + 1  # one pair per thread and step
    frames = gen.synth_frames(SEED0, 0, nfr, W, H, threads=threads)
    times, tracks = [], 0
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        tracks, kept = chk.pair_frontend_mt(frames, MAX_CORNERS, threads)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times)) if times else float("nan")
    val = tracks / (ms * 1e-3) if times else 0.0
    # RANSAC scoring half, same threads
    xi, xj = c4_points(RS_N)
    Hs = 64 * threads
    E = synthetic_hypotheses(Hs)
    t0 = time.perf_counter()
    chk.ransac_score_mt(xi, xj, E, 1e-3, threads)
    rs = Hs * RS_N / (time.perf_counter() - t0)
    sample = f"{threads} pairs of the C2 sequence per step (one per host thread), {tracks} feature-tracks; RANSAC {Hs} x {RS_N}"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_cfg(args.gpus, NFRAMES),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample,
                         "ransac_hyp_pts_per_s": rs},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def claim_stdout():
    """Route everything libraries print on fd 1 (e.g. NCCL's version banner) to stderr; the JSON line alone goes to
    the real stdout through emit()."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=NFRAMES, help="frames per GPU (default: the C2 sequence length)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pipe", type=int, default=-1, help="pairs per sub-chunk of the two-lane pipeline (-1: library default, 0: off)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3"],
                    help="c2 (default, the bench line): 1080p, 2000 corners; c3: 4K, 8000 corners (side measurement, use --frames <= 300)")
    ap.add_argument("--klt-mode", type=int, default=0, help="sfmgpu_klt_set_mode value (A/B timing of kernel variants)")
    ap.add_argument("--chunk", type=int, default=0, help="frames per chunk of the streaming e2e call (0: a quarter of the sequence)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.workload == "c3":
        global W, H, MAX_CORNERS
        W, H, MAX_CORNERS = 3840, 2160, 8000

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import sfmgpu

    torch.cuda.set_device(local)
    near_cpus = bind_near_gpu(local) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = sfmgpu.Context(local)  # raises without the CUDA library / device: there is no CPU fallback
    if args.pipe >= 0:
        ctx.pipeline_set(args.pipe)
    if args.klt_mode:
        ctx.klt_set_mode(args.klt_mode)
    nfr, npairs = args.frames, args.frames - 1
    cfg = sfmgpu.lkcfg(max_tracks=MAX_CORNERS, pyr_levels=LEVELS)
    frames = ctx.frames(W, H, nfr, LEVELS)
    pairs = ctx.pairs(npairs, MAX_CORNERS)
    frames.synth(0, nfr, SEED0 + rank, 0)  # this rank's sequence, generated in HBM
    ctx.sync()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    def step_resident():
        frames.build_pyramid(0, nfr)
        pairs.run(frames, 0, npairs, cfg)

    # ---- value: inputs resident in HBM ---------------------------------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
    ctx.sync()
    tot = pairs.totals()  # raises if a frame overflowed the candidate capacity
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launches()
    ctx.timer_start()
    for _ in range(args.steps):
        step_resident()
    ms_total = ctx.timer_stop()
    launches = ctx.launches() - l0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    n_tracks, n_kept, n_it = pairs.totals()

    # ---- per-stage times (separate, untimed-for-value pass) + FP64 peak ---------------------------------------------
    ctx.profile(True)
    ctx.timer_start()
    frames.build_pyramid(0, nfr)
    pyr_ms = ctx.timer_stop()
    pairs.run(frames, 0, npairs, cfg)
    st = ctx.stage_times()
    ctx.profile(False)
    fp64_peak = ctx.fp64_peak()

    # ---- e2e: host (pinned) frames in, all tracks out, copies inside the timed region --------------------------------
    host = ctx.pinned_empty((nfr, H, W), np.uint8)
    for k in range(0, nfr, 50):  # fill the pinned buffer with this rank's sequence (device generator == numpy twin)
        for j in range(k, min(k + 50, nfr)):
            host[j] = frames.download(j, 0)
    li = ctx.pinned_empty((npairs, MAX_CORNERS, 2), np.float64)
    lj = ctx.pinned_empty((npairs, MAX_CORNERS, 2), np.float64)
    nk = ctx.pinned_empty((npairs,), np.int32)
    nc = ctx.pinned_empty((npairs,), np.int32)
    h2d = host.nbytes
    d2h = li.nbytes + lj.nbytes + nk.nbytes + nc.nbytes

    if world > 1:
        # multi-GPU: every rank's tracks go to rank 0 over NCCL (padded gather of the device-resident results),
        # rank 0 then reads everything back to pinned host memory
        v_li, v_lj, v_nk, v_nc = pairs.torch_views(npairs)
        if rank == 0:
            g_li = [torch.empty_like(v_li) for _ in range(world)]
            g_lj = [torch.empty_like(v_lj) for _ in range(world)]
            g_nk = [torch.empty_like(v_nk) for _ in range(world)]
            g_nc = [torch.empty_like(v_nc) for _ in range(world)]
            h_li = torch.empty((world,) + tuple(v_li.shape), dtype=v_li.dtype, pin_memory=True)
            h_lj = torch.empty((world,) + tuple(v_lj.shape), dtype=v_lj.dtype, pin_memory=True)
            h_nk = torch.empty((world, npairs), dtype=torch.int32, pin_memory=True)
            h_nc = torch.empty((world, npairs), dtype=torch.int32, pin_memory=True)
            d2h = h_li.numel() * 8 + h_lj.numel() * 8 + h_nk.numel() * 4 + h_nc.numel() * 4
        else:
            g_li = g_lj = g_nk = g_nc = None
            d2h = 0

    def step_e2e():
        if world == 1:
            # the library's streaming call: H2D of chunk c+1 || pyramid + corners + KLT of chunk c || D2H of chunk c-1
            pairs.run_host(frames, host, cfg, li, lj, nk, nc, chunk=args.chunk)
            return
        # same streaming call (chunked H2D || pyramids + corners + KLT), results stay on the device for the NCCL gather
        t_a = time.perf_counter()
        pairs.run_host(frames, host, cfg, None, None, None, None, chunk=args.chunk)
        ctx.sync()  # results are written on the library's streams; NCCL runs on torch's
        t_b = time.perf_counter()
        dist.gather(v_nk, g_nk, dst=0)
        dist.gather(v_nc, g_nc, dst=0)
        dist.gather(v_li, g_li, dst=0)
        dist.gather(v_lj, g_lj, dst=0)
        torch.cuda.synchronize()
        t_c = time.perf_counter()
        if rank == 0:
            for r in range(world):
                h_li[r].copy_(g_li[r], non_blocking=True)
                h_lj[r].copy_(g_lj[r], non_blocking=True)
                h_nk[r].copy_(g_nk[r], non_blocking=True)
                h_nc[r].copy_(g_nc[r], non_blocking=True)
        torch.cuda.synchronize()
        e2e_parts.update(stream_call_ms=(t_b - t_a) * 1e3, nccl_gather_ms=(t_c - t_b) * 1e3, d2h_rank0_ms=(time.perf_counter() - t_c) * 1e3)

    e2e_parts = {}
    # the PCIe floor of the e2e step: the same frames, upload only
    frames.upload_ptr(0, nfr, host.ctypes.data)
    ctx.sync()
    ctx.timer_start()
    frames.upload_ptr(0, nfr, host.ctypes.data)
    h2d_only_ms = ctx.timer_stop()

    e2e_steps = max(1, min(args.steps, 3))
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    ctx.timer_start()
    for _ in range(e2e_steps):
        step_e2e()
    e2e_ms_dev = ctx.timer_stop()
    e2e_wall = (time.perf_counter() - t0) * 1e3
    e2e_ms = max(e2e_ms_dev, e2e_wall) / e2e_steps
    if world == 1:
        assert int(nc.sum()) == n_tracks and int(nk.sum()) == n_kept, "e2e results differ from the resident run"
    elif rank == 0:
        assert int(h_nc[0].sum()) == n_tracks and int(h_nk[0].sum()) == n_kept, "gathered results differ from the resident run"

    # ---- RANSAC half of the metric (C4), rank-local ---------------------------------------------------------------------
    xi, xj = c4_points(RS_N)
    E = synthetic_hypotheses(RS_H)
    ctx.ransac_upload(xi, xj, E)
    for _ in range(3):
        ctx.ransac_score_resident(1e-3, fetch=False)
    ctx.sync()
    ctx.timer_start()
    RS_REP = 10
    for _ in range(RS_REP):
        ctx.ransac_score_resident(1e-3, fetch=False)
    rs_ms = ctx.timer_stop() / RS_REP
    t0 = time.perf_counter()
    counts, bh, inl = ctx.ransac_score(xi, xj, E, 1e-3)
    rs_e2e_ms = (time.perf_counter() - t0) * 1e3

    # ---- the reference's own two-view entry point through the C++ shim: host solver (bit-identical hypotheses) vs the
    # opt-in device solver, TempleRing-sized call (2500 iterations x 2200 correspondences, sfm.cpp:1739)
    find_e = None
    if rank == 0:
        try:
            import ctypes as C
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import shimlib
            from conftest import TEMPLE_K, two_view_scene
            shim = shimlib.load()
            pi, pj = two_view_scene(2200, seed=2200)
            Kf = np.ascontiguousarray(TEMPLE_K.reshape(9))
            Rr, tt, il, kk = np.zeros(9), np.zeros(3), np.zeros(2200, np.int32), C.c_int(0)
            find_e = {"iters": 2500, "points": 2200}
            for name, flag in (("host_solver_ms", 0), ("device_solver_ms", 1)):
                shim.shim_set_device_solver(flag)
                shim.shim_find_E_ransac(Kf, pi, pj, 2200, 2500, 1e-3, 60, Rr, tt, il, C.byref(kk))
                t0 = time.perf_counter()
                for _ in range(3):
                    shim.shim_find_E_ransac(Kf, pi, pj, 2200, 2500, 1e-3, 60, Rr, tt, il, C.byref(kk))
                find_e[name] = (time.perf_counter() - t0) / 3 * 1e3
                find_e[name.replace("_ms", "_inliers")] = int(kk.value)
            shim.shim_set_device_solver(0)
        except Exception as ex:  # the shim is optional for the bench line
            find_e = {"error": str(ex)[:200]}

    # ---- reduce over ranks: max time, summed work; gather per-pair counts to rank 0 with NCCL --------------------------------
    ms_step = ms_total / args.steps
    vals = torch.tensor([ms_step, e2e_ms, rs_ms, pyr_ms, st["klt"], st["corner_score"], st["corner_select"]], device="cuda",
                        dtype=torch.float64)
    work = torch.tensor([n_tracks, n_kept, n_it], device="cuda", dtype=torch.int64)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
        d2h_t = torch.tensor([d2h], device="cuda", dtype=torch.int64)
        dist.all_reduce(d2h_t, op=dist.ReduceOp.SUM)
        d2h = int(d2h_t.item())
    ms_step, e2e_ms, rs_ms, pyr_ms, klt_ms, cs_ms, sel_ms = [float(v) for v in vals.tolist()]
    tracks_all, kept_all, it_all = [int(v) for v in work.tolist()]

    if rank == 0:
        hbm_peak, peak_src = measured_peaks()
        value = tracks_all / (ms_step * 1e-3)
        e2e_val = tracks_all / (e2e_ms * 1e-3)
        per_rank_tracks = tracks_all / world
        prof = profile_metrics()
        kq = prof.get("klt_quad_kernel", {})
        traffic = (kq.get("dram_read_bytes", 0) + kq.get("dram_write_bytes", 0)) or None
        klt_kernels = "klt_quad_kernel (+ klt_lane_kernel<...,masked> / klt_kernel<5,true> on border features)"
        klt_bytes = B_KLT * per_rank_tracks
        achieved = klt_bytes / (klt_ms * 1e-3) / 1e9
        pyr_bytes = nfr * W * H * sum(0.25 ** l for l in range(LEVELS))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_cfg(world, nfr),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": int(d2h) * (world if world == 1 else 1),
                    "gather": "none (1 GPU)" if world == 1 else f"NCCL gather of tracks + counts to rank 0, {world} ranks",
                    "cpus_bound_near_gpu": near_cpus, "rank0_parts_ms": e2e_parts or None,
                    "ms_per_step": e2e_ms, "steps": e2e_steps,
                    "h2d_only_ms": h2d_only_ms, "h2d_only_gb_per_s": h2d / (h2d_only_ms * 1e-3) / 1e9,
                    "note": "upload of the frames alone takes h2d_only_ms on rank 0: the end-to-end step is PCIe-bound"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": klt_kernels, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic,
                         "traffic_source": "dram__bytes_read+write of one klt_quad_kernel launch, ncu --set full on 998 C2 pairs (profiles/r1_metrics.json)",
                         "peak_source": peak_src,
                         "note": "KLT is bound by instruction issue (integer-valued FP32 matrix build + FP64 iterations), not HBM "
                                 "(SURVEY.md §8d: ~140 flop/B): see roofline_issue"},
            # The reference formulation costs 5,045 FP64 flop per LK iteration; the kernel evaluates the same sums as quadratic
            # forms over exact integer matrices, so "reference flops per second" may exceed the FP64 peak - it is a speed
            # figure, not a utilisation.  Utilisation = issue slots (ncu, committed capture).
            "roofline_issue": {"kernel": "klt_quad_kernel", "bound": "instruction issue",
                               "issue_active_frac": (kq.get("issue_active_pct") or 0) / 100.0 or None,
                               "fp64_pipe_frac": (kq.get("fp64_pipe_pct") or 0) / 100.0 or None,
                               "source": "profiles/r1_metrics.json (ncu smsp__issue_active / sm__pipe_fp64_cycles_active, not re-measured here)",
                               "reference_formulation_tflops": F_KLT_IT * (it_all / world) / (klt_ms * 1e-3) / 1e12,
                               "fp64_peak_tflops": fp64_peak, "fp64_peak_source": "in-run DFMA micro-benchmark (sfmgpu_fp64_peak)",
                               "flop_per_lk_iteration_reference": F_KLT_IT, "lk_iterations": it_all // world},
            "stages_ms": {"pyramid": pyr_ms, "corner_score": cs_ms, "corner_select": sel_ms, "klt": klt_ms, "compact": st["compact"]},
            "stage_rooflines": {
                "pyramid": {"bound": "hbm", "achieved": pyr_bytes / (pyr_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                            "frac": pyr_bytes / (pyr_ms * 1e-3) / 1e9 / hbm_peak},
                "corner_score": {"bound": "hbm", "achieved": npairs * W * H / (cs_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": npairs * W * H / (cs_ms * 1e-3) / 1e9 / hbm_peak, "mpx_per_s": npairs * W * H / (cs_ms * 1e-3) / 1e6},
            },
            "kept_fraction": kept_all / max(tracks_all, 1),
            "ransac": {"value": RS_H * RS_N / (rs_ms * 1e-3), "unit": "hyp*pts/s", "ms": rs_ms, "hypotheses": RS_H, "points": RS_N,
                       "e2e_value": RS_H * RS_N / (rs_e2e_ms * 1e-3), "best_h": int(bh), "best_inliers": int(len(inl)),
                       # 35 flop per pair in the reference formulation; the kernel screens in FP32 and runs the FP64
                       # arithmetic only for undecided pairs, so this is a speed figure relative to the DFMA peak
                       "reference_formulation_tflops_over_fp64_peak": F_RS * RS_H * RS_N / (rs_ms * 1e-3) / 1e12 / fp64_peak,
                       "issue_active_frac": (prof.get("ransac_count_kernel", {}).get("issue_active_pct") or 0) / 100.0 or None,
                       "hypotheses_source": "synthetic [t]x R around the C4 motion (scoring cost is value-independent)",
                       "find_E_ransac": find_e},
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(host)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(host_frames):
    """The reference's CPU path on this box's host cores, bounded sample of the same frames (reported, not the target)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    chk, kind = oracle.best()
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, 64))
    sample = np.ascontiguousarray(host_frames[:threads + 1])
    t0 = time.perf_counter()
    tracks, kept = chk.pair_frontend_mt(sample, MAX_CORNERS, threads)
    dt = time.perf_counter() - t0
    xi, xj = c4_points(RS_N)
    Hs = 64 * threads
    E = synthetic_hypotheses(Hs)
    t1 = time.perf_counter()
    chk.ransac_score_mt(xi, xj, E, 1e-3, threads)
    rs = Hs * RS_N / (time.perf_counter() - t1)
    t2 = time.perf_counter()
    one_tracks, _ = chk.pair_frontend_mt(sample[:2], MAX_CORNERS, 1)  # the reference is single-threaded: one pair on one core
    one_core = one_tracks / (time.perf_counter() - t2)
    return {"value": tracks / dt, "unit": UNIT, "cores": threads, "kind": kind, "one_core_value": one_core,
            "sample": f"first {threads} pairs of the same sequence, one per host thread ({tracks} feature-tracks in {dt:.2f} s wall)",
            "ransac_hyp_pts_per_s": rs}


if __name__ == "__main__":
    main()
