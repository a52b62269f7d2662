"""CPU check of the algebra behind klt_quad_kernel (csrc/klt_lane.cu): with ONE fractional offset (fx, fy) for the whole
window, the five sums of lk_step (cpp/src/templering_sfm.cpp:431-448) are quadratic forms w^T M w over INTEGER 4x4
matrices built from tap differences — including the kernel's way of getting the ten entries of a family from five product
sums over the 12 x 12 difference grid (same position, right / lower neighbour, the two diagonals) minus boundary strips.
The direct evaluation (bilinear samples, central differences, 121-term sums) is the reference's formulation."""
import numpy as np
import pytest

ENTRIES = [(0, 0), (0, 1), (0, 2), (0, 3), (1, 1), (1, 2), (1, 3), (2, 2), (2, 3), (3, 3)]


def direct_sums(t1, t0, fx, fy):
    """t1, t0: 14 x 14 tap tiles (rows, cols) of I1 / I0.  Returns (sum gx2^2, sum gx2 gy2, sum gy2^2, sum gx2 e, sum gy2 e)
    with gx2 = 2 Ix, gy2 = 2 Iy, e = I0 - I1 at the same location, over the 11 x 11 window."""
    def grid(t):  # 13 x 13 bilinear samples
        h = t[:, :-1] + (t[:, 1:] - t[:, :-1]) * fx
        return h[:-1] + (h[1:] - h[:-1]) * fy
    s1, s0 = grid(t1.astype(np.float64)), grid(t0.astype(np.float64))
    gx = s1[1:12, 2:13] - s1[1:12, 0:11]
    gy = s1[2:13, 1:12] - s1[0:11, 1:12]
    e = s0[1:12, 1:12] - s1[1:12, 1:12]
    return np.array([(gx * gx).sum(), (gx * gy).sum(), (gy * gy).sum(), (gx * e).sum(), (gy * e).sum()])


def family(A, B, sym):
    """Ten entries of sum_{window} v v'^T (upper triangle; mixed families symmetrised) the kernel's way.  A, B: 14 x 14 integer
    difference images, valid on rows / cols 1..12; window pixel (r, c), r, c = 1..11, uses (r,c), (r,c+1), (r+1,c), (r+1,c+1)."""
    g = lambda X, r0, r1, c0, c1: X[r0:r1 + 1, c0:c1 + 1].astype(np.int64)
    def P(dr, dc, r0, r1, c0, c1):  # sum over the index range of A[r][c] B[r+dr][c+dc] (+ B[r][c] A[r+dr][c+dc] if mixed and shifted)
        s = (g(A, r0, r1, c0, c1) * g(B, r0 + dr, r1 + dr, c0 + dc, c1 + dc)).sum()
        if not sym and (dr or dc):
            s += (g(B, r0, r1, c0, c1) * g(A, r0 + dr, r1 + dr, c0 + dc, c1 + dc)).sum()
        return s
    T00, R1, R12 = P(0, 0, 1, 12, 1, 12), P(0, 0, 1, 1, 1, 12), P(0, 0, 12, 12, 1, 12)
    C1, C12 = P(0, 0, 1, 12, 1, 1), P(0, 0, 1, 12, 12, 12)
    K11, K1c, Kc1, Kcc = P(0, 0, 1, 1, 1, 1), P(0, 0, 1, 1, 12, 12), P(0, 0, 12, 12, 1, 1), P(0, 0, 12, 12, 12, 12)
    T01, R1_01, R12_01 = P(0, 1, 1, 12, 1, 11), P(0, 1, 1, 1, 1, 11), P(0, 1, 12, 12, 1, 11)
    T02, C1_02, C12_02 = P(1, 0, 1, 11, 1, 12), P(1, 0, 1, 11, 1, 1), P(1, 0, 1, 11, 12, 12)
    M03 = P(1, 1, 1, 11, 1, 11)
    M12 = (g(A, 1, 11, 2, 12) * g(B, 2, 12, 1, 11)).sum() + (0 if sym else (g(B, 1, 11, 2, 12) * g(A, 2, 12, 1, 11)).sum())
    return [T00 - R12 - C12 + Kcc, T01 - R12_01, T02 - C12_02, M03, T00 - R12 - C1 + Kc1, M12, T02 - C1_02,
            T00 - R1 - C12 + K1c, T01 - R1_01, T00 - R1 - C1 + K11]


def family_direct(A, B, sym):
    out = []
    off = [(0, 0), (0, 1), (1, 0), (1, 1)]  # k -> (dr, dc)
    for k, l in ENTRIES:
        s = 0
        for r in range(1, 12):
            for c in range(1, 12):
                a = [int(A[r + dr, c + dc]) for dr, dc in off]
                b = [int(B[r + dr, c + dc]) for dr, dc in off]
                s += a[k] * b[l] + (a[l] * b[k] if (not sym and k != l) else 0)
        out.append(s)
    return out


def quad_eval(M, sym, fx, fy):
    w = [(1 - fx) * (1 - fy), fx * (1 - fy), (1 - fx) * fy, fx * fy]
    return sum(float(m) * w[k] * w[l] * (2.0 if (sym and k != l) else 1.0) for m, (k, l) in zip(M, ENTRIES))


@pytest.mark.parametrize("seed", range(6))
def test_quadratic_forms_equal_the_window_sums(seed):
    rng = np.random.default_rng(seed)
    t1 = rng.integers(0, 256, (14, 14)).astype(np.int64)
    t0 = np.clip(t1 + rng.integers(-40, 41, (14, 14)), 0, 255) if seed % 2 else rng.integers(0, 256, (14, 14)).astype(np.int64)
    DX, DY, DE = np.zeros((14, 14), np.int64), np.zeros((14, 14), np.int64), t0 - t1
    DX[:, 1:13] = t1[:, 2:14] - t1[:, 0:12]
    DY[1:13, :] = t1[2:14, :] - t1[0:12, :]
    fams = {"xx": (DX, DX, True), "yy": (DY, DY, True), "xy": (DX, DY, False), "xe": (DX, DE, False), "ye": (DY, DE, False)}
    M = {}
    for name, (A, B, sym) in fams.items():
        M[name] = family(A, B, sym)
        if seed < 2:  # the strip bookkeeping against the plain 121-pixel accumulation
            assert M[name] == family_direct(A, B, sym), name
        assert max(abs(int(v)) for v in M[name]) < 2 ** 24  # what makes the FP32 build exact
    for fx, fy in [(0.0, 0.0), (0.25, 0.5), (0.999, 0.001), tuple(rng.uniform(0, 1, 2))]:
        want = direct_sums(t1, t0, fx, fy)
        got = np.array([quad_eval(M["xx"], True, fx, fy), quad_eval(M["xy"], False, fx, fy), quad_eval(M["yy"], True, fx, fy),
                        quad_eval(M["xe"], False, fx, fy), quad_eval(M["ye"], False, fx, fy)])
        assert np.allclose(got, want, rtol=1e-12, atol=1e-6), (fx, fy, got, want)
