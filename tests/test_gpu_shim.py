"""The C++ drop-in shim (host/sfmgpu_shim.hpp: build_pyr, shi_tomasi, KLTTracker, find_E_ransac with the
reference's own names and semantics) driven end to end on the GPU, against the CPU checker."""
import ctypes as C

import numpy as np
import pytest

import shimlib
from conftest import TEMPLE_K, two_view_scene
from sfmgpu import synth

pytestmark = pytest.mark.gpu
KLT_TOL = 1e-3
W, H, SEED = 320, 240, 20261018


@pytest.fixture(scope="module")
def shim():
    return shimlib.load()


def test_build_pyr_and_corners(shim, checker):
    img = synth.frame(SEED, 0, W, H)
    out = np.zeros(W * H, np.uint8)
    shimlib.ck(shim, shim.shim_build_pyr(img, W, H, 3, out))
    want = checker.build_pyr(img, 3)
    assert np.array_equal(out[:160 * 120].reshape(120, 160), want[1])
    assert np.array_equal(out[160 * 120:160 * 120 + 80 * 60].reshape(60, 80), want[2])
    xy = np.zeros((400, 2))
    n = shimlib.ck(shim, shim.shim_shi_tomasi(img, W, H, 400, 0.01, 8, xy, 400))
    assert np.array_equal(xy[:n], checker.shi_tomasi(img, 400))


@pytest.mark.parametrize("batched", [0, 1])
def test_loop_closure_block(shim, checker, batched):
    # cpp/src/templering_sfm.cpp:1836-1857 written against the shim, per-point and batched
    f0, f1 = synth.frame(SEED, 2, W, H), synth.frame(SEED, 3, W, H)
    li, lj, nc = np.zeros((300, 2)), np.zeros((300, 2)), C.c_int(0)
    k = shimlib.ck(shim, shim.shim_pair_frontend(f0, f1, W, H, 300, 0.01, 8, 3, 5, 10, 1.0, batched, li, lj, C.byref(nc)))
    wl, wj, wnc = checker.pair_frontend(f0, f1, 300)
    assert nc.value == wnc and k == len(wl) and np.array_equal(li[:k], wl)
    assert np.abs(lj[:k] - wj).max() <= KLT_TOL


def test_track_one_public_speculation_is_bit_identical(shim):
    """The per-point loop of :1845-1849 through the shim: answered from ONE speculative batch (default) it returns exactly
    what 2 x n single launches return, for two different pairs in a row (the table is rebuilt per pair) and a repeated one."""
    res = {}
    for spec in (1, 0):
        shim.shim_set_speculate(spec)
        try:
            out = []
            for (a, b) in [(2, 3), (7, 8), (2, 3)]:
                f0, f1 = synth.frame(SEED, a, W, H), synth.frame(SEED, b, W, H)
                li, lj, nc = np.zeros((300, 2)), np.zeros((300, 2)), C.c_int(0)
                k = shimlib.ck(shim, shim.shim_pair_frontend(f0, f1, W, H, 300, 0.01, 8, 3, 5, 10, 1.0, 0, li, lj, C.byref(nc)))
                out.append((k, nc.value, li[:k].copy(), lj[:k].copy()))
            res[spec] = out
        finally:
            shim.shim_set_speculate(1)
    for x, y in zip(res[1], res[0]):
        assert x[0] == y[0] and x[1] == y[1] and np.array_equal(x[2], y[2])
        assert np.array_equal(x[3].view(np.uint64), y[3].view(np.uint64))


def test_tracker_class(shim, checker):
    kw = dict(max_tracks=150, min_tracks=120, quality=0.01, min_distance=8, levels=3, radius=5, iters=10, fb=1.0)
    want = checker.tracker(**kw)
    t = shim.shim_tracker_create(150, 120, 0.01, 8, 3, 5, 10, 1.0)
    assert t
    for fr in [0, 1, 2, 40, 41]:
        img = synth.frame(SEED, fr, W, H)
        prev, cur, ids = np.zeros((200, 2)), np.zeros((200, 2)), np.zeros(200, np.int32)
        n = shimlib.ck(shim, shim.shim_tracker_step(t, img, W, H, prev, cur, ids, 200))
        wp, wc, wi = want.step(img)
        assert n == len(wi) and np.array_equal(ids[:n], wi)
        assert n == 0 or (np.abs(prev[:n] - wp).max() <= KLT_TOL and np.abs(cur[:n] - wc).max() <= KLT_TOL)
        xy, tid = np.zeros((200, 2)), np.zeros(200, np.int32)
        m = shim.shim_tracker_tracks(t, xy, tid, 200)
        wxy, wid = want.tracks()
        assert m == len(wid) and np.array_equal(tid[:m], wid) and np.abs(xy[:m] - wxy).max() <= KLT_TOL
    shim.shim_tracker_destroy(t)


def test_multitracker_class(shim, checker):
    """MultiKLTTracker (C++ face of sfmgpu_multitracker): every sequence equals the reference's KLTTracker run on it."""
    S, cap = 3, 200
    wants = [checker.tracker(max_tracks=150, min_tracks=120) for _ in range(S)]
    t = shim.shim_multitracker_create(S, W, H, 150, 120)
    assert t
    for fr in range(5):
        imgs = np.stack([synth.frame(SEED + 7 * s, fr, W, H) for s in range(S)])
        prev, cur, ids = np.zeros((S, cap, 2)), np.zeros((S, cap, 2)), np.zeros((S, cap), np.int32)
        n, txy, tid, tn = np.zeros(S, np.int32), np.zeros((S, cap, 2)), np.zeros((S, cap), np.int32), np.zeros(S, np.int32)
        shimlib.ck(shim, shim.shim_multitracker_step(t, imgs, S, W, H, prev, cur, ids, n, cap, txy, tid, tn))
        for s in range(S):
            wp, wc, wi = wants[s].step(imgs[s])
            k = int(n[s])
            assert k == len(wi) and np.array_equal(ids[s, :k], wi), (fr, s)
            assert k == 0 or (np.abs(prev[s, :k] - wp).max() <= KLT_TOL and np.abs(cur[s, :k] - wc).max() <= KLT_TOL)
            wxy, wid = wants[s].tracks()
            m = int(tn[s])
            assert m == len(wid) and np.array_equal(tid[s, :m], wid) and np.abs(txy[s, :m] - wxy).max() <= KLT_TOL
    shim.shim_multitracker_destroy(t)


def test_find_E_ransac_device_solver(shim, checker):
    """Opt-in device solver behind the same find_E_ransac: same inlier set, pose within 1e-9 of the reference's."""
    n, iters, thr, mi = 2200, 2500, 1e-3, 60
    pi, pj = two_view_scene(n, seed=n)
    K = np.ascontiguousarray(TEMPLE_K.reshape(9))
    R, t, inl, k = np.zeros(9), np.zeros(3), np.zeros(n, np.int32), C.c_int(0)
    shim.shim_set_device_solver(1)
    try:
        ok = shimlib.ck(shim, shim.shim_find_E_ransac(K, pi, pj, n, iters, thr, mi, R, t, inl, C.byref(k)))
    finally:
        shim.shim_set_device_solver(0)
    want = checker.find_E_ransac(TEMPLE_K, pi, pj, iters, thr, mi)
    assert ok == 1 and want is not None
    assert np.array_equal(inl[:k.value], want[2])
    assert np.abs(R.reshape(3, 3) - want[0]).max() < 1e-9 and np.abs(t - want[1]).max() < 1e-9


@pytest.mark.parametrize("n,iters,thr,mi", [(400, 120, 1e-3, 60), (400, 60, 1e-9, 80), (7, 50, 1e-3, 1), (2200, 250, 1e-3, 60)])
def test_find_E_ransac(shim, checker, n, iters, thr, mi):
    pi, pj = two_view_scene(n, seed=n)
    K = np.ascontiguousarray(TEMPLE_K.reshape(9))
    R, t, inl, k = np.zeros(9), np.zeros(3), np.zeros(max(n, 1), np.int32), C.c_int(0)
    ok = shimlib.ck(shim, shim.shim_find_E_ransac(K, pi, pj, n, iters, thr, mi, R, t, inl, C.byref(k)))
    want = checker.find_E_ransac(TEMPLE_K, pi, pj, iters, thr, mi)
    assert (ok == 1) == (want is not None)
    if want is not None:
        assert np.array_equal(inl[:k.value], want[2])              # inlier mask: bit-exact
        assert np.array_equal(R.reshape(3, 3), want[0]) and np.array_equal(t, want[1])
