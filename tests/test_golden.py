"""The restatement against the committed golden vectors (tests/golden/frontend_golden.npz, produced from the
compiled reference by tests/golden/gen_golden.py).  Runs anywhere, no /root/reference needed."""
import os

import numpy as np
import pytest

from conftest import TEMPLE_K, two_view_scene
from sfmgpu import synth

W, H, SEED = 320, 240, 20261018


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "frontend_golden.npz"))


def frames(n=5):
    return [synth.frame(SEED, t, W, H) for t in range(n)]


def test_generator_is_stable(g):
    assert [synth.fnv1a64(f.tobytes()) for f in frames()] == g["frame_fnv"].tolist()


def test_pyramid(g, port):
    p = port.build_pyr(frames(1)[0], 3)
    assert np.array_equal(p[1], g["pyr_l1"]) and np.array_equal(p[2], g["pyr_l2"])


def test_corners(g, port):
    f0 = frames(1)[0]
    assert np.array_equal(port.shi_tomasi(f0, 400), g["corners"])
    assert np.array_equal(port.shi_tomasi(f0 >> 4 << 4, 400), g["corners_ties"])
    assert np.array_equal(port.shi_tomasi(synth.kat_image(W, H), 600), g["corners_kat"])


def test_klt(g, port):
    f = frames(2)
    p1, pb = port.klt_track(f[0], f[1], g["corners"].astype(np.float64)[:200])
    assert np.array_equal(p1, g["klt_p1"]) and np.array_equal(pb, g["klt_pb"])


def test_tracker(g, port):
    trk = port.tracker(max_tracks=150, min_tracks=120)
    for i, t in enumerate([0, 1, 2, 40, 41]):
        prev, cur, ids = trk.step(synth.frame(SEED, t, W, H))
        xy, tid = trk.tracks()
        assert np.array_equal(cur, g[f"trk{i}_cur"]) and np.array_equal(ids, g[f"trk{i}_ids"])
        assert np.array_equal(xy, g[f"trk{i}_tracks"]) and np.array_equal(tid, g[f"trk{i}_tids"])


def test_ransac(g, port):
    pi, pj = two_view_scene(300)
    xi, xj = port.norm_points(TEMPLE_K, pi), port.norm_points(TEMPLE_K, pj)
    assert np.array_equal(xi, g["rs_xi"]) and np.array_equal(xj, g["rs_xj"])
    E, idx = port.ransac_hypotheses(xi, xj, 40)
    assert np.array_equal(idx, g["rs_idx"]) and np.array_equal(E, g["rs_E"])
    counts, bh, inl = port.ransac_score(xi, xj, E, 1e-3)
    assert np.array_equal(counts, g["rs_counts"]) and bh == g["rs_best"][0] and np.array_equal(inl, g["rs_inl"])
    R, t, finl = port.find_E_ransac(TEMPLE_K, pi, pj, 40, 1e-3, 60)
    assert np.array_equal(R, g["fe_R"]) and np.array_equal(t, g["fe_t"]) and np.array_equal(finl, g["fe_inl"])


def test_pair_frontend(g, port):
    f = frames(4)
    li, lj, nc = port.pair_frontend(f[2], f[3], 300)
    assert nc == g["pair_nc"][0] and np.array_equal(li, g["pair_li"]) and np.array_equal(lj, g["pair_lj"])
