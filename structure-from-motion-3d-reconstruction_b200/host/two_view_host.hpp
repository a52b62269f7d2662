// two_view_host.hpp — host-side pieces of find_E_ransac that stay on the CPU in this round:
// K^-1 normalisation, the seeded 8-point minimal solver (8x9 -> AtA -> Jacobi -> rank-2 projection) and the pose
// recovery tail (E -> R,t by SVD + cheirality vote).  The hypotheses they produce are scored on the GPU.
//
// Behavioural contract (reference cpp/src/templering_sfm.cpp, cpp/include/linalg.hpp): invert_K :471-486,
// norm_point :498-501, AtA_from_A :503-517, svd3 :537-593, enforce_rank2 :595-607, eight_point_E :609-627,
// pose recovery :680-760, jacobi_eig_sym linalg.hpp:133-201.  The operation order follows the reference so
// that, compiled for baseline x86-64 without -ffast-math (no FMA contraction), hypotheses are bit-identical to
// the reference's ("same seeded hypotheses", BASELINE.json north_star); tests/test_shim_host.py checks that
// against the compiled reference.
//
// Plain arrays, no allocation in the inner loops (the reference allocates std::vectors per call).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>

namespace sfmgpu_host {

struct P3 {
  double x, y, z;
};

inline double det3(const double* a) {
  return a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
}

// false when |det K| < 1e-12 (the reference throws "Singular K")
inline bool invert_K(const double* K, double* o) {
  const double d = det3(K);
  if (std::fabs(d) < 1e-12) return false;
  o[0] = (K[4] * K[8] - K[5] * K[7]) / d;
  o[1] = -(K[1] * K[8] - K[2] * K[7]) / d;
  o[2] = (K[1] * K[5] - K[2] * K[4]) / d;
  o[3] = -(K[3] * K[8] - K[5] * K[6]) / d;
  o[4] = (K[0] * K[8] - K[2] * K[6]) / d;
  o[5] = -(K[0] * K[5] - K[2] * K[3]) / d;
  o[6] = (K[3] * K[7] - K[4] * K[6]) / d;
  o[7] = -(K[0] * K[7] - K[1] * K[6]) / d;
  o[8] = (K[0] * K[4] - K[1] * K[3]) / d;
  return true;
}

inline void norm_point(const double* Ki, double u, double v, double* xy) {
  const double a = Ki[0] * u + Ki[1] * v + Ki[2] * 1.0;
  const double b = Ki[3] * u + Ki[4] * v + Ki[5] * 1.0;
  const double c = Ki[6] * u + Ki[7] * v + Ki[8] * 1.0;
  xy[0] = a / c;
  xy[1] = b / c;
}

// Jacobi eigen-decomposition of a symmetric N x N matrix (N <= 9): largest off-diagonal pivot, angle from atan2,
// rows then columns, at most `sweeps` rotations; eigenvalues ascending in w, eigenvectors in the columns of V.
template <int N>
inline void jacobi(const double* A_in, int sweeps, double* w, double* V) {
  double A[N * N], Q[N * N];
  std::memcpy(A, A_in, sizeof A);
  for (int i = 0; i < N * N; i++) Q[i] = 0.0;
  for (int i = 0; i < N; i++) Q[i * N + i] = 1.0;
  for (int it = 0; it < sweeps; it++) {
    int p = 0, q = 1;
    double big = 0;
    for (int i = 0; i < N; i++)
      for (int j = i + 1; j < N; j++) {
        const double v = std::fabs(A[i * N + j]);
        if (v > big) {
          big = v;
          p = i;
          q = j;
        }
      }
    if (big < 1e-12) break;
    const double phi = 0.5 * std::atan2(2.0 * A[p * N + q], A[q * N + q] - A[p * N + p]);
    const double c = std::cos(phi), s = std::sin(phi);
    for (int k = 0; k < N; k++) {
      const double u = A[p * N + k], v = A[q * N + k];
      A[p * N + k] = c * u - s * v;
      A[q * N + k] = s * u + c * v;
    }
    for (int k = 0; k < N; k++) {
      const double u = A[k * N + p], v = A[k * N + q];
      A[k * N + p] = c * u - s * v;
      A[k * N + q] = s * u + c * v;
    }
    A[p * N + q] = 0.0;
    A[q * N + p] = 0.0;
    for (int k = 0; k < N; k++) {
      const double u = Q[k * N + p], v = Q[k * N + q];
      Q[k * N + p] = c * u - s * v;
      Q[k * N + q] = s * u + c * v;
    }
  }
  double d[N];
  int perm[N];
  for (int i = 0; i < N; i++) {
    d[i] = A[i * N + i];
    perm[i] = i;
  }
  std::sort(perm, perm + N, [&](int i, int j) { return d[i] < d[j]; });  // same libstdc++ sort as the reference
  for (int c = 0; c < N; c++) {
    w[c] = d[perm[c]];
    for (int r = 0; r < N; r++) V[r * N + c] = Q[r * N + perm[c]];
  }
}

template <int ROWS, int COLS>
inline void gram(const double* A, double* M) {
  for (int i = 0; i < COLS; i++)
    for (int j = i; j < COLS; j++) {
      double s = 0;
      for (int r = 0; r < ROWS; r++) s += A[r * COLS + i] * A[r * COLS + j];
      M[i * COLS + j] = s;
      M[j * COLS + i] = s;
    }
}

inline P3 unit3(P3 v) {
  const double n = std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
  if (!std::isfinite(n) || n < 1e-12) return P3{0, 0, 0};
  return P3{v.x / n, v.y / n, v.z / n};
}
inline P3 mulv(const double* A, P3 v) {
  return P3{A[0] * v.x + A[1] * v.y + A[2] * v.z, A[3] * v.x + A[4] * v.y + A[5] * v.z, A[6] * v.x + A[7] * v.y + A[8] * v.z};
}
inline void mulm(const double* A, const double* B, double* C) {
  double T[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += A[3 * r + k] * B[3 * k + c];
      T[3 * r + c] = s;
    }
  std::memcpy(C, T, sizeof T);
}
inline void transp(const double* A, double* T) {
  double t[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) t[3 * r + c] = A[3 * c + r];
  std::memcpy(T, t, sizeof t);
}

// A = U diag(s) V^T through eig(A^T A); singular values descending; U re-orthonormalised, u2 = u0 x u1.
inline void svd3(const double* A, double* U, double* s, double* V) {
  double At[9], G[9], w[3], Ve[9];
  transp(A, At);
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) {
      double t = 0;
      for (int k = 0; k < 3; k++) t += At[3 * r + k] * A[3 * k + c];
      G[3 * r + c] = t;
    }
  jacobi<3>(G, 80, w, Ve);
  const double sv[3] = {std::sqrt(std::max(0.0, w[0])), std::sqrt(std::max(0.0, w[1])), std::sqrt(std::max(0.0, w[2]))};
  int ord[3] = {0, 1, 2};
  std::sort(ord, ord + 3, [&](int i, int j) { return sv[i] > sv[j]; });
  for (int c = 0; c < 3; c++) {
    s[c] = sv[ord[c]];
    for (int r = 0; r < 3; r++) V[3 * r + c] = Ve[3 * r + ord[c]];
  }
  P3 u[3];
  for (int c = 0; c < 3; c++) {
    P3 t = mulv(A, P3{V[c], V[3 + c], V[6 + c]});
    u[c] = s[c] > 1e-12 ? P3{t.x / s[c], t.y / s[c], t.z / s[c]} : unit3(t);
  }
  const P3 u0 = unit3(u[0]);
  const double d01 = u0.x * u[1].x + u0.y * u[1].y + u0.z * u[1].z;
  const P3 u1 = unit3(P3{u[1].x - d01 * u0.x, u[1].y - d01 * u0.y, u[1].z - d01 * u0.z});
  const P3 u2 = unit3(P3{u0.y * u1.z - u0.z * u1.y, u0.z * u1.x - u0.x * u1.z, u0.x * u1.y - u0.y * u1.x});
  U[0] = u0.x; U[3] = u0.y; U[6] = u0.z;
  U[1] = u1.x; U[4] = u1.y; U[7] = u1.z;
  U[2] = u2.x; U[5] = u2.y; U[8] = u2.z;
}

// Un-normalised 8-point on NORMALISED points xi/xj (interleaved x,y) for the 8 sampled indices; rank-2 projected.
inline void eight_point_E(const double* xi, const double* xj, const int* idx8, double* E) {
  double A[72], G[81], w[9], V[81];
  for (int r = 0; r < 8; r++) {
    const int i = idx8[r];
    const double x = xi[2 * i], y = xi[2 * i + 1], xp = xj[2 * i], yp = xj[2 * i + 1];
    const double row[9] = {xp * x, xp * y, xp, yp * x, yp * y, yp, x, y, 1.0};
    for (int c = 0; c < 9; c++) A[r * 9 + c] = row[c];
  }
  gram<8, 9>(A, G);
  jacobi<9>(G, 120, w, V);
  double E0[9], U[9], s[3], Vv[9], US[9], Vt[9];
  for (int r = 0; r < 9; r++) E0[r] = V[r * 9 + 0];
  svd3(E0, U, s, Vv);
  const double S[9] = {s[0], 0, 0, 0, s[1], 0, 0, 0, 0.0};
  mulm(U, S, US);
  transp(Vv, Vt);
  mulm(US, Vt, E);
}

// Pose recovery from the winning E and its inliers (first 20 vote on cheirality); R row-major, t unit.
inline void recover_pose(const double* bestE, const double* xi, const double* xj, const int* inliers, int n_inl, double* R,
                         double* t) {
  double U[9], s[3], V[9], Vt[9], Wt[9], R1[9], R2[9], tmp[9];
  svd3(bestE, U, s, V);
  transp(V, Vt);
  const double W[9] = {0, -1, 0, 1, 0, 0, 0, 0, 1};
  transp(W, Wt);
  mulm(U, W, tmp);
  mulm(tmp, Vt, R1);
  mulm(U, Wt, tmp);
  mulm(tmp, Vt, R2);
  if (det3(R1) < 0)
    for (double& v : R1) v = -v;
  if (det3(R2) < 0)
    for (double& v : R2) v = -v;
  const P3 tt = unit3(P3{U[2], U[5], U[8]});
  auto cheirality = [&](const double* Rm, P3 tv) {
    int ok = 0;
    const int M = std::min(n_inl, 20);
    for (int k = 0; k < M; k++) {
      const int i = inliers[k];
      const double x = xi[2 * i], y = xi[2 * i + 1], xp = xj[2 * i], yp = xj[2 * i + 1];
      const double A[16] = {-1, 0, x, 0, 0, -1, y, 0,
                            xp * Rm[6] - Rm[0], xp * Rm[7] - Rm[1], xp * Rm[8] - Rm[2], xp * tv.z - tv.x,
                            yp * Rm[6] - Rm[3], yp * Rm[7] - Rm[4], yp * Rm[8] - Rm[5], yp * tv.z - tv.y};
      double G[16], w[4], Ve[16];
      gram<4, 4>(A, G);
      jacobi<4>(G, 80, w, Ve);
      const double ww = Ve[12];
      const P3 X{Ve[0] / ww, Ve[4] / ww, Ve[8] / ww};
      const double z2 = mulv(Rm, X).z + tv.z;
      if (X.z > 0 && z2 > 0) ok++;
    }
    return ok;
  };
  const double* Rs[4] = {R1, R1, R2, R2};
  const P3 ts[4] = {tt, P3{-tt.x, -tt.y, -tt.z}, tt, P3{-tt.x, -tt.y, -tt.z}};
  int bi = 0, bok = -1;
  for (int i = 0; i < 4; i++) {
    const int ok = cheirality(Rs[i], ts[i]);
    if (ok > bok) {
      bok = ok;
      bi = i;
    }
  }
  std::memcpy(R, Rs[bi], 9 * sizeof(double));
  t[0] = ts[bi].x;
  t[1] = ts[bi].y;
  t[2] = ts[bi].z;
}

// triangulate_dlt (:1477-1516): poses are camera-to-world (R row-major, camera centre C); ui / uj in pixels.
inline void triangulate_dlt(const double* K, const double* Ri, const double* Ci, const double* Rj, const double* Cj, const double* ui,
                            const double* uj, double* X) {
  double Ki[9];
  invert_K(K, Ki);  // the reference throws on a singular K before it gets here
  double xi[2], xj[2];
  norm_point(Ki, ui[0], ui[1], xi);
  norm_point(Ki, uj[0], uj[1], xj);
  double A[16];
  auto rows = [&](const double* R, const double* C, const double* x, double* out) {
    double Rw[9];
    transp(R, Rw);  // world -> camera (:164-167)
    const P3 m = mulv(Rw, P3{C[0], C[1], C[2]});
    const P3 t{-m.x, -m.y, -m.z};
    for (int which = 0; which < 2; which++) {
      const double sgn = which == 0 ? x[0] : x[1];
      double* o = out + 4 * which;
      o[0] = sgn * Rw[6] - Rw[3 * which + 0];
      o[1] = sgn * Rw[7] - Rw[3 * which + 1];
      o[2] = sgn * Rw[8] - Rw[3 * which + 2];
      o[3] = sgn * t.z - (which == 0 ? t.x : t.y);
    }
  };
  rows(Ri, Ci, xi, A);
  rows(Rj, Cj, xj, A + 8);
  double G[16], w[4], V[16];
  gram<4, 4>(A, G);
  jacobi<4>(G, 80, w, V);
  const double ww = V[12];
  X[0] = V[0] / ww;
  X[1] = V[4] / ww;
  X[2] = V[8] / ww;
}

}  // namespace sfmgpu_host
