"""CPU checks of the 3-D ring stand-in for TempleRing (tests/ring_dataset.py): the par-file poses are consistent with the
rendered frames, and the reference's own CLI + `ate_keyframes` (compiled in place by oracle/build_dropin.sh) give a finite
trajectory error on it.  The GPU comparison proper is tests/test_gpu_dropin.py::test_ring_ate_within_one_percent."""
import math
import os
import re
import subprocess

import numpy as np
import pytest

import ring_dataset

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "templering_sfm_ref")
ATE = os.path.join(ROOT, "oracle", "_ref", "ate_keyframes")


def test_poses_are_consistent_with_the_frames():
    K = ring_dataset.K_TEMPLE
    for deg in (0.0, 2.0, 40.0):
        R, t, C = ring_dataset.camera(np.deg2rad(deg))
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and np.linalg.det(R) > 0.999999
        assert np.allclose(-R.T @ t, C, atol=1e-12)  # Middlebury convention: C = -R^T t (ate_keyframes.cpp camera_center_world)
        x = K @ t  # the sphere's centre (the world origin) in the image
        assert np.allclose(x[:2] / x[2], K[:2, 2], atol=1e-9)  # the cameras look at it
    # a small rotation moves the sphere's texture one way and the wall's the other (depth on both sides of the centre)
    a = ring_dataset.render(0.0).astype(np.int32)
    b = ring_dataset.render(np.deg2rad(0.18)).astype(np.int32)
    assert a.std() > 40 and 0 < np.abs(a - b).mean() < 30


@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(ATE)), reason="reference binaries not built (no reference sources)")
def test_reference_cli_scores_finite_ate(tmp_path):
    n = 7
    root = str(tmp_path / "ring")
    ring_dataset.write_dataset(root, n, 0.18)
    (tmp_path / "config.json").write_text('{"cpp": {"ba": {"iters": 0}}}\n')
    out = str(tmp_path / "out")
    r = subprocess.run([REF, root, out, str(n)], cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    rows = open(os.path.join(out, "keyframes_camera_centers.csv")).read().strip().splitlines()[1:]
    assert len(rows) >= 4 and "nan" not in "".join(rows).lower(), rows
    q = subprocess.run([ATE, "--par", os.path.join(root, "templeRing", "templeR_par.txt"), "--keyframes",
                        os.path.join(out, "keyframes_camera_centers.csv"), "--count", str(len(rows)), "--sim3"],
                       capture_output=True, text=True, timeout=60)
    assert q.returncode == 0, q.stdout + q.stderr
    v = float(re.search(r"ATE_RMSE:\s*(\S+)", q.stdout).group(1))
    assert math.isfinite(v) and v < 0.05, q.stdout  # metres; the ring has a radius of 0.6 m
