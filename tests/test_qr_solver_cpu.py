"""CPU check of the screening solver's algebra (csrc/solver.cu: qr_null_vector): the unit vector orthogonal to the eight
rows of the design matrix, obtained as Q e_9 from a left-looking Householder QR of A^T, IS the vector the reference's
eight_point_E asks for (the eigenvector of the smallest eigenvalue of A^T A, cpp/src/templering_sfm.cpp:609-627 with
jacobi_eig_sym, cpp/include/linalg.hpp:133-201) - and it satisfies A e = 0 more accurately than the reference's own
result, which stops at 120 rotations / an absolute 1e-12.  The restatement below mirrors the kernel statement by
statement (without FMA contraction, which only makes the kernel more accurate)."""
import numpy as np
import pytest

from conftest import TEMPLE_K, two_view_scene


def design(xi, xj, octet):
    x, y, xp, yp = xi[octet, 0], xi[octet, 1], xj[octet, 0], xj[octet, 1]
    return np.stack([xp * x, xp * y, xp, yp * x, yp * y, yp, x, y, np.ones(8)], 1)


def qr_null_vector(A):
    v = np.zeros((8, 9))
    beta = np.zeros(8)
    for j in range(8):
        col = A[j].copy()
        for k in range(j):
            s = -beta[k] * np.dot(v[k, k:], col[k:])
            col[k:] += s * v[k, k:]
        nrm = np.sqrt(np.dot(col[j:], col[j:]))
        vj = col[j] + np.copysign(nrm, col[j])
        den = nrm * abs(vj)
        beta[j] = 1.0 / den if den > 0 else 0.0
        v[j, j] = vj
        v[j, j + 1:] = col[j + 1:]
    q = np.zeros(9)
    q[8] = 1.0
    for k in range(7, -1, -1):
        s = -beta[k] * np.dot(v[k, k:], q[k:])
        q[k:] += s * v[k, k:]
    return q


@pytest.mark.parametrize("n,seed,frac,sigma", [(2200, 777, 0.3, 0.3), (1000, 5, 0.5, 1.0), (3000, 9, 0.1, 0.1)])
def test_qr_null_vector_is_the_eight_point_solution(checker, n, seed, frac, sigma):
    pi, pj = two_view_scene(n, seed, frac, sigma)
    xi, xj = checker.norm_points(TEMPLE_K, pi), checker.norm_points(TEMPLE_K, pj)
    H = 600
    Eref, idx = checker.ransac_hypotheses(xi, xj, H)
    dev, res_qr, res_svd = [], [], []
    for h in range(H):
        if len(set(idx[h])) < 8:
            continue  # two-dimensional null space: no unique answer
        A = design(xi, xj, idx[h])
        e = qr_null_vector(A)
        assert abs(np.linalg.norm(e) - 1.0) < 1e-14
        sv = np.linalg.svd(A)
        e_svd = sv[2][-1]
        res_qr.append(np.abs(A @ e).max())
        res_svd.append(np.abs(A @ e_svd).max())
        dev.append(min(np.abs(e - e_svd).max(), np.abs(e + e_svd).max()) * sv[1][7] / sv[1][0])
    # orthogonal to the rows to rounding, like LAPACK's null vector; equal to it up to sign and conditioning
    assert max(res_qr) < 5e-15 and max(res_qr) < 4 * max(res_svd) + 1e-15
    assert max(dev) < 1e-14


def test_rank2_of_qr_vector_close_to_reference_hypothesis(checker):
    """After enforce_rank2 (:595-607, numpy SVD here) the QR vector gives the reference's hypothesis up to sign and the
    reference iteration's own error; its Sampson counts on the scene equal the reference's on octets of distinct points."""
    pi, pj = two_view_scene(2200, 777, 0.3, 0.3)
    xi, xj = checker.norm_points(TEMPLE_K, pi), checker.norm_points(TEMPLE_K, pj)
    H = 500
    Eref, idx = checker.ransac_hypotheses(xi, xj, H)
    D = np.zeros((H, 9))
    for h in range(H):
        e = qr_null_vector(design(xi, xj, idx[h])).reshape(3, 3)
        U, S, Vt = np.linalg.svd(e)
        D[h] = (U @ np.diag([S[0], S[1], 0.0]) @ Vt).reshape(9)
    distinct = np.array([len(set(r)) == 8 for r in idx])
    d = np.minimum(np.abs(D - Eref).max(1), np.abs(D + Eref).max(1))[distinct]
    assert np.median(d) < 1e-8 and np.quantile(d, 0.99) < 1e-4
    c_ref = checker.ransac_score(xi, xj, Eref, 1e-3)[0]
    c_qr = checker.ransac_score(xi, xj, D, 1e-3)[0]
    assert (c_ref != c_qr)[distinct].mean() < 0.005
