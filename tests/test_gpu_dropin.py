"""Whole-pipeline drop-in check: the reference's own CLI (`templering_sfm`, unmodified) against the same main()
with the front-end definitions replaced by host/sfmgpu_shim.hpp (oracle/build_dropin.sh applies the patch of
INTEGRATION.md to a scratch copy).  Both run on synthetic PGM datasets laid out like TempleRing; the keyframe
centres and pose-graph edges they write must agree.  This stands in for the "downstream ATE within 1 %" criterion
(the TempleRing dataset itself is not shipped): identical keyframes and centres imply identical ATE.

What can be compared.  The reference's `main` triangulates a new keyframe's points with `kfs[idl].pose` BEFORE that
keyframe is pushed (sfm.cpp:1809 reads one element past the vector, :1815 pushes it; AddressSanitizer: heap-buffer-overflow),
so its map points - and, through bundle adjustment, its camera centres - depend on heap contents, not on the front end:
the unmodified binary returns NaN centres on every synthetic set here, the drop-in binary (other allocations) finite ones.
Keyframe decisions, map-point COUNTS and pose-graph edges do not depend on the map; with BA switched off through the
reference's own config file (`cpp.ba.iters = 0`) neither do the centres (chained `find_E_ransac` poses + pose graph).

Two datasets:
* value-noise frames (planar scene, `sfmgpu.synth`): the default pipeline incl. BA - progress lines and
  `posegraph_edges.csv` must agree; then BA off - `keyframes_camera_centers.csv` must agree too;
* a ray-cast 3-D ring scene with true poses in the par file (`tests/ring_dataset.py`), BA off: centres finite and equal,
  and the reference's own `ate_keyframes` tool scores both runs against the ground truth - Sim(3) and SE(3) ATE must
  agree within 1 %.
"""
import csv
import math
import os
import re
import subprocess

import numpy as np
import pytest

from sfmgpu import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "templering_sfm_ref")
GPU = os.path.join(ROOT, "oracle", "_ref", "templering_sfm_gpu")
ATE = os.path.join(ROOT, "oracle", "_ref", "ate_keyframes")
W, H, NFR = 640, 480, 14


def write_dataset(root):
    os.makedirs(os.path.join(root, "templeRing"), exist_ok=True)
    os.makedirs(os.path.join(root, "templeRing_pgm"), exist_ok=True)
    K = [1520.4, 0, 302.32, 0, 1525.9, 246.87, 0, 0, 1]
    par, ang = [str(NFR)], []
    for i in range(NFR):
        name = f"templeR{i + 1:04d}"
        img = synth.frame(3, 8 * i, W, H)  # stride 8: 0.5 px of true motion per frame -> parallax >= 18 px, 13 keyframes, BA runs
        with open(os.path.join(root, "templeRing_pgm", name + ".pgm"), "wb") as f:
            f.write(f"P5\n{W} {H}\n255\n".encode())
            f.write(img.tobytes())
        vals = K + [1, 0, 0, 0, 1, 0, 0, 0, 1] + [0.01 * i, 0, 0]
        par.append(name + ".png " + " ".join(repr(float(v)) for v in vals))
        ang.append(f"{10.0} {float(i)} {name}.png")
    open(os.path.join(root, "templeRing", "templeR_par.txt"), "w").write("\n".join(par) + "\n")
    open(os.path.join(root, "templeRing", "templeR_ang.txt"), "w").write("\n".join(ang) + "\n")


def read_csv(path):
    rows = list(csv.reader(open(path)))
    return rows[0], rows[1:]


def run(binary, root, out, cwd, frames=NFR):
    r = subprocess.run([binary, root, out, str(frames)], cwd=cwd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def compare_csv(tmp_path, name, tol=1e-6):
    """Rows of <name> in out_ref / out_gpu: equal text, or numbers within tol (relative); NaNs must coincide.
    Returns (rows, worst difference, number of NaN fields)."""
    h1, r1 = read_csv(str(tmp_path / "out_ref" / name))
    h2, r2 = read_csv(str(tmp_path / "out_gpu" / name))
    assert h1 == h2 and len(r1) == len(r2), name
    worst, nans = 0.0, 0
    for a, b in zip(r1, r2):
        assert len(a) == len(b), name
        for x, y in zip(a, b):
            try:
                fx, fy = float(x), float(y)
            except ValueError:
                assert x == y
                continue
            if math.isnan(fx) or math.isnan(fy):
                assert math.isnan(fx) and math.isnan(fy), (name, a, b)
                nans += 1
                continue
            worst = max(worst, abs(fx - fy) / max(1.0, abs(fx)))
    print(f"{name}: {len(r1)} rows, worst relative difference {worst:.3e}, {nans} NaN fields (coinciding)")
    assert worst <= tol, name
    return len(r1), worst, nans


@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(GPU)), reason="drop-in binaries not built (no reference sources at build time)")
def test_cli_outputs_agree(tmp_path):
    root = str(tmp_path / "data")
    write_dataset(root)
    so_ref = run(REF, root, str(tmp_path / "out_ref"), str(tmp_path))
    so_gpu = run(GPU, root, str(tmp_path / "out_gpu"), str(tmp_path))
    # same progress lines: keyframe and map-point counts per frame
    assert [l for l in so_ref.splitlines() if l.startswith("frame ")] == [l for l in so_gpu.splitlines() if l.startswith("frame ")]
    compare_csv(tmp_path, "posegraph_edges.csv")
    assert len(read_csv(str(tmp_path / "out_ref" / "keyframes_camera_centers.csv"))[1]) >= 2
    # BA off (the reference reads ./config.json from its working directory, sfm.cpp:1614): the centres are comparable
    (tmp_path / "config.json").write_text(NO_BA)
    nfr = 8
    so_ref = run(REF, root, str(tmp_path / "out_ref"), str(tmp_path), nfr)
    so_gpu = run(GPU, root, str(tmp_path / "out_gpu"), str(tmp_path), nfr)
    assert [l for l in so_ref.splitlines() if l.startswith("frame ")] == [l for l in so_gpu.splitlines() if l.startswith("frame ")]
    nkf, _, nans = compare_csv(tmp_path, "keyframes_camera_centers.csv")
    compare_csv(tmp_path, "posegraph_edges.csv")
    assert nkf >= 4 and nans == 0, (nkf, nans)


NO_BA = '{"cpp": {"ba": {"iters": 0}}}\n'


def ate(par, keyframes, count, mode):
    r = subprocess.run([ATE, "--par", par, "--keyframes", keyframes, "--count", str(count), mode], capture_output=True, text=True,
                       timeout=120)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
    m = re.search(r"ATE_RMSE:\s*(\S+)", r.stdout)
    assert m, r.stdout
    return float(m.group(1))


RING_FRAMES, RING_STEP_DEG = 12, 0.18  # 0.18 deg per frame: ~0.6 px of true flow, >= 18 px tracked (a keyframe per frame)


@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(GPU) and os.path.exists(ATE)),
                    reason="drop-in binaries not built (no reference sources at build time)")
def test_ring_ate_within_one_percent(tmp_path):
    import ring_dataset

    root = str(tmp_path / "ring")
    ring_dataset.write_dataset(root, RING_FRAMES, RING_STEP_DEG)
    (tmp_path / "config.json").write_text(NO_BA)  # BA off (see the module docstring)
    so_ref = run(REF, root, str(tmp_path / "out_ref"), str(tmp_path), RING_FRAMES)
    so_gpu = run(GPU, root, str(tmp_path / "out_gpu"), str(tmp_path), RING_FRAMES)
    assert [l for l in so_ref.splitlines() if l.startswith("frame ")] == [l for l in so_gpu.splitlines() if l.startswith("frame ")]
    nkf, _, nans = compare_csv(tmp_path, "keyframes_camera_centers.csv")
    compare_csv(tmp_path, "posegraph_edges.csv")
    assert nkf >= 6 and nans == 0, (nkf, nans)
    par = os.path.join(root, "templeRing", "templeR_par.txt")
    for mode in ("--sim3", "--se3"):
        a_ref = ate(par, str(tmp_path / "out_ref" / "keyframes_camera_centers.csv"), nkf, mode)
        a_gpu = ate(par, str(tmp_path / "out_gpu" / "keyframes_camera_centers.csv"), nkf, mode)
        print(f"ATE {mode}: reference {a_ref:.6g}, GPU front end {a_gpu:.6g} over {nkf} keyframes")
        assert math.isfinite(a_ref) and math.isfinite(a_gpu)
        assert abs(a_gpu - a_ref) <= 0.01 * a_ref, (mode, a_ref, a_gpu)
