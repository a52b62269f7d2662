"""CPU check of the error bounds behind the FP32 screen of csrc/ransac.cu (ransac_count_kernel): the kernel's FP32
arithmetic is emulated (float32 storage, FMA = exact double product + one rounding), the exact values come from
80-bit long doubles, and the distance must stay inside the bounds the kernel uses:
    |num_f - num| <= 32u Esum P' Pm,   |den_f - den| <= 64u (Emax Pm)^2 + 8u den_f,   u = 2^-24,
and the single-residual test built on them (|r_f| > B_f, then the sign of r_f) must never contradict the reference's
FP64 decision fl(n2 / den) < thr, including thresholds placed a few ppm beside the computed errors."""
import numpy as np
import pytest

U = np.float64(2.0 ** -24)
f32 = np.float32


def fma(a, b, c):
    # a*b is exact in double (24+24 bits); the sum is rounded once to double (2^-53) and once to float
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


@pytest.mark.parametrize("scale_pts,scale_e,seed", [(1.0, 1.0, 0), (0.3, 1.0, 1), (1500.0, 1.0, 2), (1.0, 1e-6, 3), (1e4, 1e3, 4), (1e-3, 1.0, 5)])
def test_fp32_screen_bounds_hold(scale_pts, scale_e, seed):
    rng = np.random.default_rng(seed)
    n = 400_000
    x, y, xp, yp = [rng.uniform(-1, 1, n) * scale_pts for _ in range(4)]
    # near-epipolar pairs (small numerators, the cancellation case) mixed with random ones
    E = rng.normal(0, 1, (n, 9)) * scale_e
    half = n // 2
    ex64 = E[:half, 0] * x[:half] + E[:half, 1] * y[:half] + E[:half, 2]
    ey64 = E[:half, 3] * x[:half] + E[:half, 4] * y[:half] + E[:half, 5]
    ez64 = E[:half, 6] * x[:half] + E[:half, 7] * y[:half] + E[:half, 8]
    with np.errstate(divide="ignore", invalid="ignore"):
        ypn = -(xp[:half] * ex64 + ez64) / ey64  # makes x'^T E x ~ 0
    ok = np.isfinite(ypn) & (np.abs(ypn) <= 10 * scale_pts)
    yp[:half][ok] = ypn[ok]

    L = np.longdouble
    El = E.astype(L)
    xl, yl, xpl, ypl = x.astype(L), y.astype(L), xp.astype(L), yp.astype(L)
    ex = El[:, 0] * xl + El[:, 1] * yl + El[:, 2]
    ey = El[:, 3] * xl + El[:, 4] * yl + El[:, 5]
    ez = El[:, 6] * xl + El[:, 7] * yl + El[:, 8]
    tx = El[:, 0] * xpl + El[:, 3] * ypl + El[:, 6]
    ty = El[:, 1] * xpl + El[:, 4] * ypl + El[:, 7]
    num = np.abs(xpl * ex + ypl * ey + ez)
    den = ex * ex + ey * ey + tx * tx + ty * ty + L(1e-12)

    Ef = E.astype(f32)
    xf, yf, xpf, ypf = x.astype(f32), y.astype(f32), xp.astype(f32), yp.astype(f32)
    exf = fma(Ef[:, 0], xf, fma(Ef[:, 1], yf, Ef[:, 2]))
    eyf = fma(Ef[:, 3], xf, fma(Ef[:, 4], yf, Ef[:, 5]))
    ezf = fma(Ef[:, 6], xf, fma(Ef[:, 7], yf, Ef[:, 8]))
    txf = fma(Ef[:, 0], xpf, fma(Ef[:, 3], ypf, Ef[:, 6]))
    tyf = fma(Ef[:, 1], xpf, fma(Ef[:, 4], ypf, Ef[:, 7]))
    numf = np.abs(fma(xpf, exf, fma(ypf, eyf, ezf)))
    denf = fma(tyf, tyf, fma(txf, txf, fma(eyf, eyf, fma(exf, exf, np.full(n, 1e-12, f32)))))

    A = np.abs(E)
    esum = A.sum(1)
    emax = np.max(np.stack([A[:, 0:3].sum(1), A[:, 3:6].sum(1), A[:, 6:9].sum(1), A[:, [0, 3, 6]].sum(1), A[:, [1, 4, 7]].sum(1)]), 0)
    P = np.maximum(np.maximum(np.abs(x), np.abs(y)), 1.0)
    Pp = np.maximum(np.maximum(np.abs(xp), np.abs(yp)), 1.0)
    Pm = np.maximum(P, Pp)
    dn = 32 * U * esum * Pp * Pm
    dd = 64 * U * (emax * Pm) ** 2 + 8 * U * denf.astype(np.float64)
    err_n = np.abs(numf.astype(L) - num).astype(np.float64)
    err_d = np.abs(denf.astype(L) - den).astype(np.float64)
    assert np.all(err_n <= dn), (float((err_n / dn).max()),)
    assert np.all(err_d <= dd), (float((err_d / dd).max()),)
    # the slack the derivation promises (21u / 48u of 32u / 64u): the bounds are not hanging by a thread
    assert (err_n / dn).max() < 0.7 and (err_d / dd).max() < 0.8


def _ru(x):
    """round a float64 array up to float32 (like __double2float_ru)"""
    f = x.astype(f32)
    low = f.astype(np.float64) < x
    return np.where(low, np.nextafter(f, f32(np.inf)), f).astype(f32)


@pytest.mark.parametrize("scale_pts,scale_e,seed", [(1.0, 1.0, 10), (0.3, 1.0, 11), (1500.0, 1.0, 12), (1.0, 1e-6, 13), (30.0, 1e3, 14),
                                                    (1e-3, 1.0, 15)])
def test_residual_screen_never_contradicts_the_reference(scale_pts, scale_e, seed):
    rng = np.random.default_rng(seed)
    n = 300_000
    x, y, xp, yp = [rng.uniform(-1, 1, n) * scale_pts for _ in range(4)]
    E = rng.normal(0, 1, (n, 9)) * scale_e
    # the reference's FP64 arithmetic in its own operation order (sampson_err, templering_sfm.cpp:629-638)
    ex = (E[:, 0] * x + E[:, 1] * y) + E[:, 2]
    ey = (E[:, 3] * x + E[:, 4] * y) + E[:, 5]
    ez = (E[:, 6] * x + E[:, 7] * y) + E[:, 8]
    tx = (E[:, 0] * xp + E[:, 3] * yp) + E[:, 6]
    ty = (E[:, 1] * xp + E[:, 4] * yp) + E[:, 7]
    num = (xp * ex + yp * ey) + ez
    den = (((ex * ex + ey * ey) + tx * tx) + ty * ty) + 1e-12
    e64 = (num * num) / den

    Ef = E.astype(f32)
    xf, yf, xpf, ypf = x.astype(f32), y.astype(f32), xp.astype(f32), yp.astype(f32)
    exf = fma(Ef[:, 0], xf, fma(Ef[:, 1], yf, Ef[:, 2]))
    eyf = fma(Ef[:, 3], xf, fma(Ef[:, 4], yf, Ef[:, 5]))
    ezf = fma(Ef[:, 6], xf, fma(Ef[:, 7], yf, Ef[:, 8]))
    txf = fma(Ef[:, 0], xpf, fma(Ef[:, 3], ypf, Ef[:, 6]))
    tyf = fma(Ef[:, 1], xpf, fma(Ef[:, 4], ypf, Ef[:, 7]))
    nv = fma(xpf, exf, fma(ypf, eyf, ezf))
    denf = fma(tyf, tyf, fma(txf, txf, fma(eyf, eyf, fma(exf, exf, np.full(n, 1e-12, f32)))))

    A = np.abs(E)
    esum = (A[:, 0] + A[:, 1] + A[:, 2]) + (A[:, 3] + A[:, 4] + A[:, 5]) + (A[:, 6] + A[:, 7] + A[:, 8])
    emax = np.max(np.stack([A[:, 0:3].sum(1), A[:, 3:6].sum(1), A[:, 6:9].sum(1), A[:, [0, 3, 6]].sum(1), A[:, [1, 4, 7]].sum(1)]), 0)
    Pm = np.maximum(np.maximum(np.maximum(np.abs(x), np.abs(y)), np.maximum(np.abs(xp), np.abs(yp))), 1.0)
    z = _ru(Pm * Pm * 1.000002)
    cn = np.maximum(_ru(esum * (32.0 * U * 1.000001)), f32(1e-30))
    c1 = (cn + cn).astype(f32)
    c2 = _ru(cn.astype(np.float64) * cn.astype(np.float64) * 1.000001)
    kappa = f32((2e-6 + 16.0 * U) * 1.00001)
    undecided, total = 0, 0
    for rel in (0.0, 1e-9, -1e-9, 3e-7, -3e-7, 1.5e-6, -1.5e-6, 4e-6, -4e-6, 3e-5, -3e-5, 1e-3, -1e-3, 0.05, -0.05, 3.0, -0.7):
        thr = e64 * (1.0 + rel)
        ok = (thr > 1e-18) & (thr < 1e18)
        thr_f = thr.astype(f32)
        cq = _ru(thr * (1.0 + 2e-6) * (64.0 * U * 1.000001) * emax * emax)
        with np.errstate(over="ignore", invalid="ignore"):
            td = (thr_f.astype(np.float64) * denf.astype(np.float64)).astype(f32)
            r = fma(nv, nv, -td)
            sm = fma(nv, nv, td)
            i1 = fma(c2, z, cq)
            i2 = fma(c1, np.abs(nv), i1)
            ks = (kappa.astype(np.float64) * sm.astype(np.float64)).astype(f32)
            bb = fma(z, i2, ks)
            decided = (np.abs(r) > bb) & ok
        want = e64 < thr
        bad = decided & ((r < 0) != want)
        assert not bad.any(), (rel, int(bad.sum()), float(e64[bad][0]), float(thr[bad][0]))
        if abs(rel) >= 1e-3:
            undecided += int((~decided & ok).sum())
            total += int(ok.sum())
    # the screen has to be worth it: pairs a per-mille or more away from the threshold are (almost) always decided
    assert undecided <= 0.02 * total, (undecided, total)
