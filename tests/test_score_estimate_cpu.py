"""CPU check of the FP32 estimate behind the corner-score screen (csrc/corner_score.cu: est_u, EST_MARGIN).

Every pixel gets u~ = (a+b) - sqrt.approx(fma(2c, 2c, (a-b)^2)) in FP32 on the exact integer window sums
a = sum Gx^2, b = sum Gy^2, c = sum GxGy (each <= 25 * 255^2 = 1,625,625 < 2^24, so a, b, c, a+b, a-b, 2c are exact
floats); only pixels with u~ >= float_rd(8 thr) - EST_MARGIN see the exact FP64 expression
u = (a+b) - sqrt((a-b)^2 + 4c^2) (= 8 lmin of sfm.cpp:266-270), which alone decides.  The screen is rigorous iff
|u~ - u| < EST_MARGIN.  Derivation: (a-b)^2 and the FMA carry one rounding each (relative 2^-24 each on D, half of that
on the root), sqrt.approx.ftz.f32 has a maximum relative error of 2^-23 (PTX ISA), the root is <= a+b <= 3.25e6
(Cauchy-Schwarz: c^2 <= ab), and the final subtraction rounds once more: |u~ - u| <= 3.25e6 * 1.8e-7 + 0.125 < 0.72.
This test emulates the arithmetic (float32 storage, FMA = exact double product-sum + one rounding, the approximate root
perturbed by +-2^-23 in both directions) on window sums of real images, random admissible triples and the extreme corners
of the domain."""
import numpy as np

f32 = np.float32
EST_MARGIN = 2.0  # csrc/corner_score.cu
AMAX = 25 * 255 * 255


def est_u_worst(a, b, c):
    """Largest |u~ - u| over the two extreme outcomes of sqrt.approx, for integer arrays a, b, c."""
    af, bf, cf = a.astype(f32), b.astype(f32), c.astype(f32)
    assert np.all(af.astype(np.int64) == a) and np.all(bf.astype(np.int64) == b) and np.all(cf.astype(np.int64) == c)
    d = af - bf  # exact
    c2 = cf + cf  # exact
    dd = d * d  # one rounding
    D = (c2.astype(np.float64) * c2.astype(np.float64) + dd.astype(np.float64)).astype(f32)  # FMA: one rounding
    root = np.sqrt(D.astype(np.float64))
    t = af + bf  # exact (< 2^24)
    u_exact = (a + b).astype(np.float64) - np.sqrt(((a - b).astype(np.float64)) ** 2 + (2.0 * c.astype(np.float64)) ** 2)
    worst = np.zeros(len(a))
    for rel in (-(2.0 ** -23), 0.0, 2.0 ** -23):
        s = (root * (1.0 + rel)).astype(f32)
        u = (t - s).astype(np.float64)  # float32 subtraction, value held in double
        worst = np.maximum(worst, np.abs(u - u_exact))
    return worst


def window_sums(img):
    im = img.astype(np.int64)
    p = np.pad(im, 1, mode="edge")
    gx = p[1:-1, 2:] - p[1:-1, :-2]
    gy = p[2:, 1:-1] - p[:-2, 1:-1]

    def box(v):
        h, w = v.shape
        out = np.zeros((h - 4, w - 4), np.int64)
        for dy in range(5):
            for dx in range(5):
                out += v[dy:dy + h - 4, dx:dx + w - 4]
        return out

    return box(gx * gx).ravel(), box(gy * gy).ravel(), box(gx * gy).ravel()


def test_estimate_error_on_image_window_sums():
    rng = np.random.default_rng(3)
    worst = 0.0
    imgs = [rng.integers(0, 256, (96, 128)), (rng.integers(0, 2, (96, 128)) * 255),  # noise, saturated noise
            np.add.outer(np.arange(96) * 5 % 256, np.arange(128) * 7 % 256) % 256,  # ramps with wrap-around edges
            ((np.add.outer(np.arange(96) // 3, np.arange(128) // 3) % 2) * 255)]  # checkerboard: the largest sums
    for im in imgs:
        a, b, c = window_sums(im)
        assert a.max() <= AMAX and b.max() <= AMAX and np.all(c * c <= a * b)
        worst = max(worst, est_u_worst(a, b, c).max())
    print(f"worst |u~ - u| on image window sums: {worst:.4f}")
    assert worst < 0.72 < EST_MARGIN


def test_estimate_error_on_admissible_triples_and_extremes():
    rng = np.random.default_rng(4)
    n = 2_000_000
    a = rng.integers(0, AMAX + 1, n)
    b = rng.integers(0, AMAX + 1, n)
    # half of the sample close to the Cauchy-Schwarz boundary c^2 = ab (root ~ a+b: the largest absolute error),
    # a quarter with a ~ b (cancellation in a-b), the rest anywhere inside
    lim = np.floor(np.sqrt(a.astype(np.float64) * b)).astype(np.int64)
    lim -= (lim * lim > a * b)
    frac = np.where(rng.random(n) < 0.5, 1.0 - rng.random(n) * 1e-3, rng.random(n))
    c = (np.floor(lim * frac).astype(np.int64)) * rng.choice([-1, 1], n)
    q = n // 4
    b[:q] = np.clip(a[:q] + rng.integers(-3, 4, q), 0, AMAX)
    lim_q = np.floor(np.sqrt(a[:q].astype(np.float64) * b[:q])).astype(np.int64)
    lim_q -= (lim_q * lim_q > a[:q] * b[:q])
    c[:q] = np.clip(c[:q], -lim_q, lim_q)
    ext = np.array([(AMAX, AMAX, AMAX), (AMAX, AMAX, -AMAX), (AMAX, AMAX, 0), (AMAX, 0, 0), (0, AMAX, 0), (0, 0, 0), (1, 1, 1),
                    (AMAX, AMAX - 1, AMAX - 1), (AMAX, 1, 1275), (AMAX - 7, AMAX, -(AMAX - 8))], np.int64)
    a = np.concatenate([a, ext[:, 0]])
    b = np.concatenate([b, ext[:, 1]])
    c = np.concatenate([c, ext[:, 2]])
    assert np.all(c * c <= a * b)
    w = est_u_worst(a, b, c)
    print(f"worst |u~ - u| on {len(a)} admissible triples: {w.max():.4f}")
    assert w.max() < 0.72 < EST_MARGIN
