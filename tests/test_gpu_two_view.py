"""GPU parity tests of the RANSAC stage of the two-view unit, batched over pairs (cpp/src/templering_sfm.cpp:1855-1857 =
find_E_ransac :640-761 per pair), through the C ABI.

Bars: index octets of the seeded sampler, inlier counts, winner and inlier lists BIT-EXACT (for the same hypotheses);
with the caller's (reference) hypotheses the whole result equals find_E_ransac's, R and t bit-identical through the host
tail and within 1e-8 through the device tail; with the device solver the same status / winner count / inlier list on the
test scenes (hypotheses agree to ~1e-9)."""
import ctypes as C

import numpy as np
import pytest

from conftest import TEMPLE_K, two_view_scene
from sfmgpu import synth
import sfmgpu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,count", [(8, 64), (100, 4000), (2200, 20000), (10000, 32000), (7919, 32000), (1500000000, 20000),
                                     (2147483647, 5000), (3, 100), (1, 10)])
def test_device_sampler_matches_libstdcpp(ctx, checker, n, count):
    """std::mt19937(12345) + std::uniform_int_distribution<int>(0, n-1), :657-665.  n = 1.5e9 rejects 30 % of the raw
    values (Lemire's rejection loop): the device's ordered compaction must skip exactly those."""
    assert np.array_equal(ctx.ransac_sample(n, count), checker.rng_draws(n, count))


def _scenes():
    """Correspondence sets of different sizes: big, tiny (< 8: nullopt before any RNG use), below the caller's guard, mostly
    outliers (winner below min_inliers), exactly the capacity."""
    specs = [(2200, 777, 0.3), (5, 1, 0.0), (100, 2, 0.2), (900, 3, 0.97), (2500, 4, 0.5), (8, 5, 0.0), (640, 6, 0.1), (0, 7, 0.0)]
    out = []
    for n, seed, frac in specs:
        pi, pj = two_view_scene(max(n, 1), seed=seed, outlier_frac=frac)
        out.append((pi[:n], pj[:n]))
    return out


def _reference_results(checker, K, scenes, iters, thr, min_inl, min_pts):
    want = []
    for pi, pj in scenes:
        if len(pi) < min_pts:
            want.append((0, None))
            continue
        r = checker.find_E_ransac(K, pi, pj, iters, thr, min_inl)
        want.append((1, None) if r is None else (2, r))
    return want


@pytest.mark.parametrize("iters,thr,min_inl,min_pts", [(400, 1e-3, 60, 120), (257, 1e-4, 80, 0), (64, 2e-3, 700, 9)])
def test_pairs_ransac_with_reference_hypotheses(ctx, checker, iters, thr, min_inl, min_pts):
    """The reference's own hypotheses (seeded sampling + eight_point_E) scored per pair in one batch: status, winner count,
    inlier list bit-identical to find_E_ransac; R, t bit-identical through the host tail, ~1e-9 through the device tail."""
    import shimlib
    shim = shimlib.load()
    K = TEMPLE_K
    scenes = _scenes()
    cap = 2500
    pairs = ctx.pairs(len(scenes), cap)
    pairs.set_matches([s[0] for s in scenes], [s[1] for s in scenes])
    E_host = np.zeros((len(scenes), iters, 9))
    norm = []
    for k, (pi, pj) in enumerate(scenes):
        xi, xj = checker.norm_points(K, pi), checker.norm_points(K, pj)
        norm.append((xi, xj))
        if len(pi) >= 8:
            E_host[k] = checker.ransac_hypotheses(xi, xj, iters)[0]
    pairs.ransac(K, iters, thr, min_inl, min_pts, E_host=E_host)
    want = _reference_results(checker, K, scenes, iters, thr, min_inl, min_pts)
    seen = set()
    for k, (wst, wr) in enumerate(want):
        st, bh, inl, E, R, t = pairs.ransac_download(k)
        assert st == wst, (k, st, wst)
        seen.add(st)
        if wst != 2:
            continue
        wR, wt, winl = wr
        assert np.array_equal(inl, winl), k
        assert np.array_equal(E.reshape(9), E_host[k, bh]), k
        assert np.abs(R - wR).max() < 1e-8 and np.abs(t - wt).max() < 1e-8, (k, np.abs(R - wR).max())
        # host tail on the device's winner: bit-identical R, t
        xi, xj = norm[k]
        hR, ht = np.zeros(9), np.zeros(3)
        shim.shim_host_recover_pose(np.ascontiguousarray(E.reshape(9)), np.ascontiguousarray(xi), np.ascontiguousarray(xj),
                                    np.ascontiguousarray(inl), len(inl), hR, ht)
        assert np.array_equal(hR.reshape(3, 3), wR) and np.array_equal(ht, wt), k
    assert seen >= ({0, 2} if min_pts == 120 else {1, 2})  # skipped, no pose and pose all occur across the configurations


@pytest.mark.parametrize("solver_mode", [3, 1, 0])
@pytest.mark.parametrize("iters,thr,min_inl,min_pts", [(400, 1e-3, 60, 120), (1000, 2e-3, 80, 120), (4000, 2e-3, 80, 120)])
def test_pairs_ransac_device_solver(ctx, checker, iters, thr, min_inl, min_pts, solver_mode):
    """Everything on the device (sampler, 8-point solver, scoring, pose): same status, winner count and inlier list as
    find_E_ransac on these scenes; R, t within 1e-6 (the device solver's hypotheses agree to ~1e-9, not bit for bit).
    solver_mode 3: counts from the screening solver, winner re-solved by the Jacobi emulation (what mode 1, the default,
    does for launches of more than 75,776 hypotheses); 0: the emulation for every hypothesis."""
    K = TEMPLE_K
    scenes = _scenes()
    ctx.solver_set_mode(solver_mode)
    pairs = ctx.pairs(len(scenes), 2500)
    pairs.set_matches([s[0] for s in scenes], [s[1] for s in scenes])
    try:
        pairs.ransac(K, iters, thr, min_inl, min_pts)
    finally:
        ctx.solver_set_mode(1)
    want = _reference_results(checker, K, scenes, iters, thr, min_inl, min_pts)
    st_all, bn_all = np.zeros(len(scenes), np.int32), np.zeros(len(scenes), np.int32)
    inl_all = np.zeros((len(scenes), 2500), np.int32)
    R_all, t_all = np.zeros((len(scenes), 9)), np.zeros((len(scenes), 3))
    pairs.ransac_download_all(st_all, bn_all, inl_all, R_all, t_all)
    for k, (wst, wr) in enumerate(want):
        st, bh, inl, E, R, t = pairs.ransac_download(k)
        assert st == wst == st_all[k], (k, st, wst)
        if wst != 2:
            continue
        wR, wt, winl = wr
        assert np.array_equal(inl, winl) and bn_all[k] == len(winl) and np.array_equal(inl_all[k, :len(winl)], winl), k
        assert np.abs(R - wR).max() < 1e-6 and np.abs(t - wt).max() < 1e-6, (k, np.abs(R - wR).max())
        assert np.array_equal(R_all[k], R.reshape(9)) and np.array_equal(t_all[k], t)
        # the winning hypothesis against the reference solver's for the same octet
        xi, xj = checker.norm_points(K, scenes[k][0]), checker.norm_points(K, scenes[k][1])
        Eref = checker.ransac_hypotheses(xi, xj, bh + 1)[0][bh]
        d = min(np.abs(E.reshape(9) - Eref).max(), np.abs(E.reshape(9) + Eref).max())
        assert d < 1e-7, (k, d)


def test_device_solver_hypotheses_close_to_reference(ctx, checker):
    """The shared-memory 8-point kernel against the reference's eight_point_E for the same octets (sign-insensitive)."""
    pi, pj = two_view_scene(2200, seed=11)
    xi, xj = checker.norm_points(TEMPLE_K, pi), checker.norm_points(TEMPLE_K, pj)
    H = 3000
    Eref, idx = checker.ransac_hypotheses(xi, xj, H)
    E = ctx.ransac_hypotheses(xi, xj, idx)
    d = np.minimum(np.abs(E - Eref).max(1), np.abs(E + Eref).max(1))
    print(f"device solver: median deviation {np.median(d):.2e}, 99 % {np.quantile(d, 0.99):.2e}, max {d.max():.2e}")
    # hypotheses with a repeated index (sampling with replacement) or a near-degenerate octet have an arbitrary basis
    distinct = np.array([len(set(r)) == 8 for r in idx])
    assert np.quantile(d[distinct], 0.9) < 1e-8
    ctx.ransac_hypotheses(xi, xj, idx, fetch=False)
    bh, bn = ctx.ransac_score_resident(1e-3)
    counts, inl = ctx.ransac_download(H, len(xi))
    wc, wbh, winl = checker.ransac_score(xi, xj, Eref, 1e-3)
    assert bh == wbh and np.array_equal(inl[:bn], winl)
    assert (counts != wc).mean() < 0.02


@pytest.mark.parametrize("n,seed,frac,sigma,thr", [(2200, 11, 0.3, 0.3, 1e-3), (1000, 5, 0.5, 1.0, 2e-3), (3000, 9, 0.1, 0.1, 1e-4),
                                                  (640, 6, 0.1, 0.3, 1e-5), (2500, 4, 0.5, 0.3, 2e-3)])
def test_screening_solver_counts(ctx, checker, n, seed, frac, sigma, thr):
    """Solver mode 3 (= mode 1 for launches beyond one wave): the counts of the screening solver (unit null vector of the design matrix by Householder QR) against
    the reference solver's counts for the same octets.  Octets of eight distinct points: equal counts except where a
    point's error sits within the reference iteration's own error of the threshold (< 0.5 % of the hypotheses, by
    at most a few points); octets with a repeated index have a two-dimensional null space and are left to the Jacobi
    emulation in both modes (which member comes back is a property of the iteration: the emulation reproduces it for
    ~80 % of them).  Winner, its count and inlier list as find_E_ransac's."""
    pi, pj = two_view_scene(n, seed=seed, outlier_frac=frac, sigma=sigma)
    xi, xj = checker.norm_points(TEMPLE_K, pi), checker.norm_points(TEMPLE_K, pj)
    H = 3000
    Eref, idx = checker.ransac_hypotheses(xi, xj, H)
    wc, wbh, winl = checker.ransac_score(xi, xj, Eref, thr)
    distinct = np.array([len(set(r)) == 8 for r in idx])
    res, rep_counts = {}, {}
    for mode in (3, 0):
        ctx.solver_set_mode(mode)
        try:
            bh, bn, E, inl = ctx.ransac_solve_score(xi, xj, idx, thr)
        finally:
            ctx.solver_set_mode(1)
        counts, _ = ctx.ransac_download(H, n)
        diff = counts != wc
        res[mode] = (diff[distinct].mean(), np.abs(counts - wc)[distinct].max(), diff[~distinct].mean())
        rep_counts[mode] = counts[~distinct].copy()
        assert bh == wbh and bn == len(winl) and np.array_equal(inl, winl), mode
        d = min(np.abs(E - Eref[wbh]).max(), np.abs(E + Eref[wbh]).max())
        assert d < 1e-7, (mode, d)  # the winner's hypothesis is the emulation's in both modes
        assert diff[distinct].mean() < 0.005 and np.abs(counts - wc)[distinct].max() <= 4, (mode, res[mode])
    assert np.array_equal(rep_counts[0], rep_counts[3])
    print(f"counts != reference (distinct octets / max |diff| / repeated-index octets): screening {res[3]}, emulation {res[0]}")


@pytest.mark.parametrize("n,seed", [(2200, 11), (8, 5), (40, 3)])
def test_warp_emulation_bit_identical(ctx, checker, n, seed):
    """The warp-per-hypothesis form of the Jacobi emulation (winners, repeated-index octets of single calls) returns the
    thread-per-hypothesis kernel's hypotheses bit for bit - octets with repeated indices included (n = 8: all of them)."""
    pi, pj = two_view_scene(n, seed=seed, outlier_frac=0.2)
    xi, xj = checker.norm_points(TEMPLE_K, pi), checker.norm_points(TEMPLE_K, pj)
    _, idx = checker.ransac_hypotheses(xi, xj, 1500)
    got = {}
    for mode in (0, 2):
        ctx.solver_set_mode(mode)
        try:
            got[mode] = ctx.ransac_hypotheses(xi, xj, idx)
        finally:
            ctx.solver_set_mode(1)
    assert np.array_equal(got[0].view(np.uint64), got[2].view(np.uint64))


def test_solve_score_edge_cases(ctx, checker):
    """No hypothesis (H = 0), eight points, a threshold nothing passes: winner -1, zero matrix, empty list (as :667-677)."""
    pi, pj = two_view_scene(8, seed=5, outlier_frac=0.0)
    xi, xj = checker.norm_points(TEMPLE_K, pi), checker.norm_points(TEMPLE_K, pj)
    bh, bn, E, inl = ctx.ransac_solve_score(xi, xj, np.zeros((0, 8), np.int32), 1e-3)
    assert bh == -1 and bn == 0 and not E.any() and len(inl) == 0
    _, idx = checker.ransac_hypotheses(xi, xj, 50)
    bh, bn, E, inl = ctx.ransac_solve_score(xi, xj, idx, -1.0)
    assert bh == -1 and bn == 0 and not E.any() and len(inl) == 0
    bh, bn, E, inl = ctx.ransac_solve_score(xi, xj, idx, 1e-3)
    Eref, _ = checker.ransac_hypotheses(xi, xj, 50)
    wc, wbh, winl = checker.ransac_score(xi, xj, Eref, 1e-3)
    assert bh == wbh and np.array_equal(inl, winl)
    # octets sampled on the device (idx8 == NULL): the same call
    bh2, bn2, E2, inl2 = ctx.ransac_solve_score(xi, xj, None, 1e-3, iters=50)
    assert bh2 == bh and bn2 == bn and np.array_equal(E2, E) and np.array_equal(inl2, inl)


@pytest.mark.parametrize("n,H", [(2200, 2500), (150, 4000), (9, 300)])
def test_solve_score_device_sampling(ctx, checker, n, H):
    """sfmgpu_ransac_solve_score with idx8 == NULL samples the reference's octets itself (:657-665): same winner, count,
    hypothesis and inlier list as with the reference's octets passed in."""
    pi, pj = two_view_scene(n, seed=n, outlier_frac=0.3)
    xi, xj = checker.norm_points(TEMPLE_K, pi), checker.norm_points(TEMPLE_K, pj)
    _, idx = checker.ransac_hypotheses(xi, xj, H)
    a = ctx.ransac_solve_score(xi, xj, idx, 1e-3)
    b = ctx.ransac_solve_score(xi, xj, None, 1e-3, iters=H)
    assert a[0] == b[0] and a[1] == b[1] and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
    Eref, _ = checker.ransac_hypotheses(xi, xj, H)
    wc, wbh, winl = checker.ransac_score(xi, xj, Eref, 1e-3)
    assert b[0] == wbh and np.array_equal(b[3], winl)


@pytest.mark.parametrize("streaming", [False, True])
def test_pair_frontend_with_ransac_stage(ctx, checker, streaming):
    """The stage inside sfmgpu_pair_frontend / sfmgpu_pair_frontend_host: same results as find_E_ransac on the survivors
    of every pair (device solver: status, count, inlier list), and as the explicit call."""
    W, H = 320, 240
    imgs = [synth.frame(20261018, t, W, H) for t in range(6)]
    imgs[3] = np.full((H, W), 90, np.uint8)  # a flat frame: overflow redo path with the stage on (pairs 2 and 3)
    cfg = sfmgpu.lkcfg(max_tracks=300)
    K = TEMPLE_K
    iters, thr, min_inl, min_pts = 300, 2e-3, 80, 120
    pairs = ctx.pairs(5, 300)
    pairs.set_ransac(K, iters, thr, min_inl, min_pts)
    st, bn = np.full(5, -7, np.int32), np.full(5, -7, np.int32)
    inl = np.zeros((5, 300), np.int32)
    R, t = np.zeros((5, 9)), np.zeros((5, 3))
    li, lj = np.zeros((5, 300, 2)), np.zeros((5, 300, 2))
    nk, nc = np.zeros(5, np.int32), np.zeros(5, np.int32)
    if streaming:
        pairs.ransac_host_outputs(st, bn, inl, R, t)
        f = ctx.frames(W, H, 6, 3)
        pairs.run_host(f, np.stack(imgs), cfg, li, lj, nk, nc, chunk=2)
    else:
        f = ctx.frames(W, H, 6, 3)
        f.upload(0, np.stack(imgs))
        f.build_pyramid()
        pairs.run(f, 0, 5, cfg)
        pairs.download_all(li, lj, nk, nc)
        pairs.ransac_download_all(st, bn, inl, R, t)
    for p in range(5):
        a, b = li[p, :nk[p]], lj[p, :nk[p]]
        wl, wj, _ = checker.pair_frontend(imgs[p], imgs[p + 1], 300)
        assert np.array_equal(a, wl)
        if len(a) < min_pts:
            assert st[p] == 0, p
            continue
        r = checker.find_E_ransac(K, a, b, iters, thr, min_inl)  # on the GPU's own survivors (lj agrees to ~1e-11)
        if r is None:
            assert st[p] == 1, p
            continue
        assert st[p] == 2 and bn[p] == len(r[2]) and np.array_equal(inl[p, :bn[p]], r[2]), p
        assert np.abs(R[p].reshape(3, 3) - r[0]).max() < 1e-6 and np.abs(t[p] - r[1]).max() < 1e-6, p
    # the explicit call on the same batch returns the same
    pairs.set_ransac(None)
    pairs.ransac(K, iters, thr, min_inl, min_pts)
    st2, bn2, inl2 = np.zeros(5, np.int32), np.zeros(5, np.int32), np.zeros((5, 300), np.int32)
    R2, t2 = np.zeros((5, 9)), np.zeros((5, 3))
    pairs.ransac_download_all(st2, bn2, inl2, R2, t2)
    assert np.array_equal(st, st2) and np.array_equal(bn, bn2) and np.array_equal(R, R2) and np.array_equal(t, t2)
    for p in range(5):
        if st[p] == 2:
            assert np.array_equal(inl[p, :bn[p]], inl2[p, :bn2[p]])


@pytest.mark.parametrize("solver_mode", [1, 3, 0])
@pytest.mark.parametrize("iters,thr,min_pts", [(4000, 2e-3, 120), (300, 2e-3, 120), (4000, 1e-6, 9), (4000, 6e-7, 9)])
def test_pairs_ransac_early_stop_is_exact(ctx, checker, iters, thr, min_pts, solver_mode):
    """The early stop of the batched stage (a pair leaves the launch set once one of its first 128 hypotheses explains all
    of its points: nothing later can win, :673) against the same stage with every hypothesis solved and scored: status,
    winner, count, inlier list, E, R, t bit-identical on a batch that mixes outlier-free sets with contaminated, tiny and
    skipped ones.  thr 2e-3 (the reference's): hypothesis 0 explains every outlier-free set; 1e-6: the first full hypothesis
    sits at 0 ... 51; 6e-7: at 1, 103 (stopped), 155, 224, 355 (beyond the probe: scored to the end, the full count
    still wins) or nowhere.  At the reference's threshold also against find_E_ransac."""
    K = TEMPLE_K
    min_inl = 80
    scenes = _scenes()
    for n, seed in ((500, 21), (2500, 22), (130, 23), (1200, 24), (60, 25), (300, 26)):  # no outliers, 0.3 px noise
        pi, pj = two_view_scene(n, seed=seed, outlier_frac=0.0)
        scenes.insert(len(scenes) // 2, (pi, pj))
    P = len(scenes)
    res = {}
    ctx.solver_set_mode(solver_mode)
    try:
        for early in (1, 0):
            ctx.ransac_set_early_stop(early)
            pairs = ctx.pairs(P, 2500)
            pairs.set_matches([s[0] for s in scenes], [s[1] for s in scenes])
            pairs.ransac_early()
            pairs.ransac(K, iters, thr, min_inl, min_pts)
            stopped = pairs.ransac_early()
            st, bn = np.zeros(P, np.int32), np.zeros(P, np.int32)
            inl, R, t = np.zeros((P, 2500), np.int32), np.zeros((P, 9)), np.zeros((P, 3))
            pairs.ransac_download_all(st, bn, inl, R, t)
            one = [pairs.ransac_download(k) for k in range(P)]
            res[early] = (stopped, st, bn, inl, R, t, one)
    finally:
        ctx.ransac_set_early_stop(1)
        ctx.solver_set_mode(1)
    s1, st1, bn1, inl1, R1, t1, one1 = res[1]
    s0, st0, bn0, inl0, R0, t0, one0 = res[0]
    assert s0 == 0
    assert np.array_equal(st1, st0) and np.array_equal(bn1, bn0)
    okp = st1 == 2  # R, t of pairs without a pose are not written (the two runs use different pairs objects)
    assert okp.sum() >= 6 and np.array_equal(R1[okp], R0[okp]) and np.array_equal(t1[okp], t0[okp])
    full = 0
    for k in range(P):
        a, b = one1[k], one0[k]
        assert a[0] == b[0], k
        if st1[k] == 2:
            assert a[1] == b[1] and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3]), k
            assert np.array_equal(a[4], b[4]) and np.array_equal(a[5], b[5]), k
            assert np.array_equal(inl1[k, :bn1[k]], inl0[k, :bn0[k]]), k
            full += int(bn1[k] == len(scenes[k][0]) and a[1] < 128)
    print(f"iters {iters} thr {thr} mode {solver_mode}: {s1} of {P} pairs stopped early, {full} valid poses with a full count in the probe")
    assert full >= 1 and 1 <= s1 <= P, (full, s1)
    if solver_mode == 0:  # exact counts of the emulation's hypotheses decide both the stop and the final count
        assert full <= s1, (full, s1)
    if thr == 2e-3:
        assert full >= 4
        want = _reference_results(checker, K, scenes, iters, thr, min_inl, min_pts)
        for k, (wst, wr) in enumerate(want):
            assert st1[k] == wst, k
            if wst == 2:
                assert bn1[k] == len(wr[2]) and np.array_equal(inl1[k, :bn1[k]], wr[2]), k
