"""Pins the restatement (oracle/sfm_oracle.cpp) against the reference itself (oracle/_ref, the unmodified TU).

The reference ships no tests or golden vectors; these cases are therefore the pin.  CPU only.
"""
import numpy as np
import pytest

from conftest import TEMPLE_K, two_view_scene, triangulation_scene
from sfmgpu import synth


def _images():
    rng = np.random.default_rng(3)
    kat = synth.kat_image(320, 200)
    return {
        "kat": kat,
        "synth": synth.frame(11, 0, 320, 240),
        "ties": (synth.frame(12, 0, 200, 160) >> 4 << 4),  # coarse grey levels -> many tied scores
        "noise": rng.integers(0, 256, (97, 131), dtype=np.uint8),  # odd sizes
        "flat": np.full((24, 40), 77, np.uint8),  # thr = 0: every pixel is a candidate
        "tiny": rng.integers(0, 256, (4, 9), dtype=np.uint8),  # h < 5: all-zero score map
        "one_bright": np.pad(np.full((1, 1), 255, np.uint8), 20),
    }


@pytest.mark.parametrize("name", list(_images()))
def test_pyramid(name, port, ref):
    img = _images()[name]
    for levels in (1, 2, 3, 4):
        a, b = ref.build_pyr(img, levels), port.build_pyr(img, levels)
        assert len(a) == len(b) == levels
        for x, y in zip(a, b):
            assert x.shape == y.shape and np.array_equal(x, y)


@pytest.mark.parametrize("name", list(_images()))
@pytest.mark.parametrize("mc,q,d", [(2200, 0.01, 8), (50, 0.2, 3), (0, 0.01, 8), (300, 0.0, 1), (100, 0.01, 0)])
def test_corners(name, mc, q, d, port, ref):
    img = _images()[name]
    assert np.array_equal(ref.shi_tomasi(img, mc, q, d), port.shi_tomasi(img, mc, q, d))


def test_kat_corners_match_survey(ref):
    # first six corners of the 640x480 known-answer image
    c = ref.shi_tomasi(synth.kat_image(), 2200)
    assert len(c) == 2200
    assert c[:6].astype(int).tolist() == [[545, 159], [305, 367], [255, 256], [239, 49], [49, 111], [289, 159]]


def test_klt(port, ref):
    rng = np.random.default_rng(5)
    f0, f1 = synth.frame(21, 0, 320, 240), synth.frame(21, 3, 320, 240)
    pts = np.concatenate([port.shi_tomasi(f0, 150), rng.uniform(-8, 330, (80, 2)) * [1, 0.75],
                          [[0.0, 0.0], [318.0, 238.0], [319.5, 100.25], [127.99999999999999, 64.0]]])
    for lv, r, it in [(3, 5, 10), (1, 3, 4), (4, 2, 7)]:
        a = ref.klt_track(f0, f1, pts, lv, r, it)
        b = port.klt_track(f0, f1, pts, lv, r, it)
        assert np.array_equal(a[0], b[0], equal_nan=True) and np.array_equal(a[1], b[1], equal_nan=True)


def test_tracker_sequence(port, ref):
    kw = dict(max_tracks=120, min_tracks=100, quality=0.01, min_distance=8, levels=3, radius=5, iters=10, fb=1.0)
    ta, tb = ref.tracker(**kw), port.tracker(**kw)
    for t in [0, 1, 2, 40, 41, 42, 43]:  # the jump 2 -> 40 kills tracks and triggers the replenish rule
        img = synth.frame(31, t, 256, 192)
        ra, rb = ta.step(img), tb.step(img)
        for x, y in zip(ra, rb):
            assert np.array_equal(x, y, equal_nan=True)
        for x, y in zip(ta.tracks(), tb.tracks()):
            assert np.array_equal(x, y, equal_nan=True)


def test_rng_kat(port, ref):
    # SURVEY.md §8c(1)
    assert ref.rng_draws(100, 16).tolist() == [92, 89, 31, 13, 18, 3, 20, 82, 56, 53, 59, 95, 96, 46, 65, 92]
    for n in (8, 100, 2200, 10000):
        assert np.array_equal(ref.rng_draws(n, 64), port.rng_draws(n, 64))


def test_sampson_kat(port, ref):
    E = [0, -0.1, 0.3, 0.12, 0, -0.9, -0.28, 0.91, 0.01]
    assert ref.sampson(E, (0.1, -0.05), (0.11, -0.04)) == 1.7519471550715924e-05
    assert port.sampson(E, (0.1, -0.05), (0.11, -0.04)) == 1.7519471550715924e-05


def test_hypotheses_and_scores(port, ref):
    pi, pj = two_view_scene(500)
    xi, xj = ref.norm_points(TEMPLE_K, pi), ref.norm_points(TEMPLE_K, pj)
    assert np.array_equal(xi, port.norm_points(TEMPLE_K, pi))
    Ea, ia = ref.ransac_hypotheses(xi, xj, 60)
    Eb, ib = port.ransac_hypotheses(xi, xj, 60)
    assert np.array_equal(ia, ib) and np.array_equal(Ea, Eb)
    for thr in (1e-3, 1e-6):
        a, b = ref.ransac_score(xi, xj, Ea, thr), port.ransac_score(xi, xj, Ea, thr)
        assert np.array_equal(a[0], b[0]) and a[1] == b[1] and np.array_equal(a[2], b[2])


@pytest.mark.parametrize("n,iters,thr,mi", [(400, 120, 1e-3, 60), (400, 60, 1e-9, 80), (7, 50, 1e-3, 1), (30, 40, 2e-3, 5)])
def test_find_E_ransac(n, iters, thr, mi, port, ref):
    pi, pj = two_view_scene(n, seed=n)
    a, b = ref.find_E_ransac(TEMPLE_K, pi, pj, iters, thr, mi), port.find_E_ransac(TEMPLE_K, pi, pj, iters, thr, mi)
    assert (a is None) == (b is None)
    if a is not None:
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


def test_pair_frontend(port, ref):
    f0, f1 = synth.frame(41, 7, 256, 192), synth.frame(41, 8, 256, 192)
    a, b = ref.pair_frontend(f0, f1, 150), port.pair_frontend(f0, f1, 150)
    assert a[2] == b[2] and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("w,h", [(640, 480), (1920, 1080), (33, 70), (20, 15), (64, 64)])
def test_loop_descriptor(w, h, port, ref):
    """global_desc_32 :1100-1122 and the candidate search :1823-1831: restatement == compiled reference, bit for bit."""
    from sfmgpu import synth
    imgs = [synth.frame(9, t, w, h) for t in (0, 5, 30)]
    da, db = [port.global_desc32(i) for i in imgs], [ref.global_desc32(i) for i in imgs]
    assert all(np.array_equal(a, b) for a, b in zip(da, db))
    ra, rb = port.desc_search(np.stack(da[:2]), da[2]), ref.desc_search(np.stack(db[:2]), db[2])
    assert ra[0] == rb[0] and ra[1] == rb[1] and np.array_equal(ra[2], rb[2])


def test_triangulate_dlt(port, ref):
    """triangulate_dlt :1477-1516: restatement == compiled reference, bit for bit; and close to the true points."""
    poses, ia, ib, ui, uj, X = triangulation_scene(400)
    a, b = ref.triangulate_dlt(TEMPLE_K, poses, ia, ib, ui, uj), port.triangulate_dlt(TEMPLE_K, poses, ia, ib, ui, uj)
    assert np.array_equal(a, b)
    assert np.median(np.linalg.norm(a - X, axis=1)) < 0.05


def test_two_view_unit(port, ref):
    """The whole two-view unit (:1836-1857: front end, size guard, find_E_ransac) per pair: restatement == compiled
    reference, bit for bit; and equal to the single-pair entry points."""
    fr = np.stack([synth.frame(20261018, t, 320, 240) for t in range(4)])
    ta, a = ref.two_view_mt(fr, 300, 3, K=TEMPLE_K, rs_iters=150, rs_thr=2e-3, rs_min_inliers=80, min_points=120)
    tb, b = port.two_view_mt(fr, 300, 2, K=TEMPLE_K, rs_iters=150, rs_thr=2e-3, rs_min_inliers=80, min_points=120)
    assert ta == tb and all(np.array_equal(a[k], b[k]) for k in a)
    for p in range(3):
        li, lj, nc = ref.pair_frontend(fr[p], fr[p + 1], 300)
        assert nc == a["n_corners"][p] and np.array_equal(li, a["li"][p, :len(li)]) and np.array_equal(lj, a["lj"][p, :len(lj)])
        if len(li) < 120:
            assert a["status"][p] == 0
            continue
        r = ref.find_E_ransac(TEMPLE_K, li, lj, 150, 2e-3, 80)
        assert (r is None) == (a["status"][p] == 1)
        if r is not None:
            assert np.array_equal(r[0].reshape(9), a["R"][p]) and np.array_equal(r[1], a["t"][p])
            assert np.array_equal(r[2], a["inliers"][p, :a["n_inl"][p]])
    assert {0, 2} <= set(a["status"].tolist())
