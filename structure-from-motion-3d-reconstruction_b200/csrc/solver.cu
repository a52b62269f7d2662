// solver.cu — the 8-point minimal solver on the device, one thread per hypothesis (OPT-IN).
//
// Replaces (reference cpp/src/templering_sfm.cpp) AtA_from_A :503-517, svd3 :537-593, enforce_rank2 :595-607,
// eight_point_E :609-627 and jacobi_eig_sym (cpp/include/linalg.hpp:133-201) for the hypotheses of find_E_ransac.
// The seeded sampling (:657-665, std::mt19937 + std::uniform_int_distribution) stays on the host and is passed in
// as index octets, so the SAME eight correspondences feed every hypothesis.
//
// NOT bit-identical to the host solver (host/two_view_host.hpp, which is): the Jacobi rotation angle goes through
// atan2 / cos / sin, and CUDA's double-precision functions differ from glibc's in the last bits.  Formulas, operation
// order, pivot rule (largest off-diagonal, first in raster order), rotation caps (120 / 80) and the 1e-12 stop are
// the reference's; measured agreement of the hypotheses is ~1e-12 relative (up to the eigenvector's sign, which the
// Sampson error ignores) and the tests require the same winner and inlier set on their scenes.  The reference path
// of the drop-in (bit-identical hypotheses from the host, scored on the device) remains the default.
#include <string.h>

#include "common.cuh"

namespace {

template <int N>
__device__ void jacobi_dev(double* A, double* Q, int sweeps) {
  for (int i = 0; i < N * N; i++) Q[i] = 0.0;
  for (int i = 0; i < N; i++) Q[i * N + i] = 1.0;
  for (int it = 0; it < sweeps; it++) {
    int p = 0, q = 1;
    double big = 0;
    for (int i = 0; i < N; i++)
      for (int j = i + 1; j < N; j++) {
        const double v = fabs(A[i * N + j]);
        if (v > big) {
          big = v;
          p = i;
          q = j;
        }
      }
    if (big < 1e-12) break;
    const double phi = 0.5 * atan2(2.0 * A[p * N + q], A[q * N + q] - A[p * N + p]);
    double s, c;
    sincos(phi, &s, &c);
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
}

__device__ __forceinline__ void unit3(double& x, double& y, double& z) {
  const double n = sqrt(x * x + y * y + z * z);
  if (!isfinite(n) || n < 1e-12) {
    x = y = z = 0.0;
  } else {
    x /= n;
    y /= n;
    z /= n;
  }
}

// ---- 8-point solver: one thread per hypothesis, working set in shared memory ------------------------------------------------
// The 9x9 Gram matrix is symmetric and the reference's row-then-column rotation keeps it EXACTLY symmetric outside the
// 2x2 pivot block (entry (k,p) gets only the column update c*a_kp - s*a_kq, entry (p,k) only the row update
// c*a_pk - s*a_qk: same operands, same operations), so the packed upper triangle (45 doubles) carries the whole state.
// It lives in shared memory, thread-interleaved (element e of thread t at e*TPB + t): every lane owns its own pair of
// banks, so the data-dependent pivot indices never conflict.  The eigenvector matrix is not accumulated: the rotations
// are recorded (c, s, p, q) and applied backwards to the unit vector of the smallest eigenvalue - 4 instead of 54
// multiply-adds per rotation (V = J1 J2 ... Jn, column m = J1 (J2 (... (Jn e_m)))).
// The rotation angle phi = atan2(2 a_pq, a_qq - a_pp) / 2 is evaluated algebraically (half-angle formulas, 1 sqrt + 1 div +
// 1 sqrt + 1 div) instead of atan2 / sincos: cos and sin agree with libm's to a few ulp, which is the same class of
// deviation CUDA's own trigonometry has against glibc's.
#ifndef SV_TPB_V
#define SV_TPB_V 128
#endif
#ifndef SV_MINB_V
#define SV_MINB_V 4
#endif
constexpr int SV_TPB = SV_TPB_V, SV_MINB = SV_MINB_V;  // 128 x 4: 16 warps per SM in 184 KB of shared memory, <= 128 registers
constexpr int SV_TRI = 45;
constexpr int SV_ROT9 = 120, SV_ROT3 = 80;  // the reference's rotation caps (:622, :551)

__device__ __forceinline__ int tri_idx(int i, int j) { return (i * (17 - i)) / 2 + j; }  // i <= j < 9

__device__ __forceinline__ void half_angle(double y, double x, double& c, double& s) {
  const double r = sqrt(x * x + y * y);
  if (!(r > 0.0)) {  // atan2(0, 0) = 0
    c = 1.0;
    s = 0.0;
    return;
  }
  const double ct = x / r;
  if (x >= 0.0) {
    c = sqrt(0.5 * (1.0 + ct));
    s = (y / r) / (2.0 * c);
  } else {
    const double sa = sqrt(0.5 * (1.0 - ct));
    s = copysign(sa, y);
    c = (fabs(y) / r) / (2.0 * sa);
  }
}

// One rotation of a 3x3 symmetric problem held in registers, pivot (P, Q) static: rows, columns, zero, V (the
// reference's order, linalg.hpp:164-188).
template <int P, int Q>
__device__ __forceinline__ void rot3(double* A, double* V) {
  double c, s;
  half_angle(2.0 * A[P * 3 + Q], A[Q * 3 + Q] - A[P * 3 + P], c, s);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const double u = A[P * 3 + k], v = A[Q * 3 + k];
    A[P * 3 + k] = c * u - s * v;
    A[Q * 3 + k] = s * u + c * v;
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const double u = A[k * 3 + P], v = A[k * 3 + Q];
    A[k * 3 + P] = c * u - s * v;
    A[k * 3 + Q] = s * u + c * v;
  }
  A[P * 3 + Q] = 0.0;
  A[Q * 3 + P] = 0.0;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const double u = V[k * 3 + P], v = V[k * 3 + Q];
    V[k * 3 + P] = c * u - s * v;
    V[k * 3 + Q] = s * u + c * v;
  }
}

__device__ __forceinline__ void jacobi3(double* A, double* V) {
#pragma unroll
  for (int i = 0; i < 9; i++) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
  for (int it = 0; it < SV_ROT3; it++) {
    const double a01 = fabs(A[1]), a02 = fabs(A[2]), a12 = fabs(A[5]);
    int piv = 0;
    double big = 0.0;
    if (a01 > big) { big = a01; piv = 0; }
    if (a02 > big) { big = a02; piv = 1; }
    if (a12 > big) { big = a12; piv = 2; }
    if (big < 1e-12) break;
    if (piv == 0) rot3<0, 1>(A, V);
    else if (piv == 1) rot3<0, 2>(A, V);
    else rot3<1, 2>(A, V);
  }
}

// svd3 (:537-593) of E0 (row-major): singular values descending in s, V (columns), the first two columns of the
// re-orthonormalised U (the third never contributes to U diag(s0, s1, 0) V^T; u2 = u0 x u1 is returned for the pose tail).
__device__ __forceinline__ void svd3_dev(const double* E0, double* U, double* s, double* V) {
  double G3[9], V3[9];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) {
      double t = 0;
#pragma unroll
      for (int k = 0; k < 3; k++) t += E0[3 * k + r] * E0[3 * k + c];
      G3[3 * r + c] = t;
    }
  jacobi3(G3, V3);
  // jacobi's ascending order first (stable on the diagonal), then the descending order of sqrt(max(0, w))
  int asc[3] = {0, 1, 2};
#pragma unroll
  for (int i = 1; i < 3; i++)
    for (int j = i; j > 0 && G3[asc[j] * 4] < G3[asc[j - 1] * 4]; j--) {
      const int t = asc[j];
      asc[j] = asc[j - 1];
      asc[j - 1] = t;
    }
  double sv[3];
#pragma unroll
  for (int c = 0; c < 3; c++) sv[c] = sqrt(fmax(0.0, G3[asc[c] * 4]));
  int ord[3] = {0, 1, 2};
#pragma unroll
  for (int i = 1; i < 3; i++)
    for (int j = i; j > 0 && sv[ord[j]] > sv[ord[j - 1]]; j--) {
      const int t = ord[j];
      ord[j] = ord[j - 1];
      ord[j - 1] = t;
    }
#pragma unroll
  for (int c = 0; c < 3; c++) {
    s[c] = sv[ord[c]];
    const int src = asc[ord[c]];
#pragma unroll
    for (int r = 0; r < 3; r++) V[3 * r + c] = V3[3 * r + src];
  }
  double u[3][3];
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const double vx = V[c], vy = V[3 + c], vz = V[6 + c];
    double tx = E0[0] * vx + E0[1] * vy + E0[2] * vz, ty = E0[3] * vx + E0[4] * vy + E0[5] * vz,
           tz = E0[6] * vx + E0[7] * vy + E0[8] * vz;
    if (s[c] > 1e-12) {
      tx /= s[c];
      ty /= s[c];
      tz /= s[c];
    } else {
      unit3(tx, ty, tz);
    }
    u[c][0] = tx; u[c][1] = ty; u[c][2] = tz;
  }
  double u0x = u[0][0], u0y = u[0][1], u0z = u[0][2];
  unit3(u0x, u0y, u0z);
  const double d01 = u0x * u[1][0] + u0y * u[1][1] + u0z * u[1][2];
  double u1x = u[1][0] - d01 * u0x, u1y = u[1][1] - d01 * u0y, u1z = u[1][2] - d01 * u0z;
  unit3(u1x, u1y, u1z);
  double u2x = u0y * u1z - u0z * u1y, u2y = u0z * u1x - u0x * u1z, u2z = u0x * u1y - u0y * u1x;
  unit3(u2x, u2y, u2z);
  U[0] = u0x; U[3] = u0y; U[6] = u0z;
  U[1] = u1x; U[4] = u1y; U[7] = u1z;
  U[2] = u2x; U[5] = u2y; U[8] = u2z;
}

// enforce_rank2 (:595-607) of the 3x3 matrix E0 (row-major): E = U diag(s0, s1, 0) V^T
__device__ __forceinline__ void rank2_project(const double* E0, double* __restrict__ E) {
  double U[9], sv[3], V[9];
  svd3_dev(E0, U, sv, V);
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) {
      // (U S)(r,k) = U(r,k) s_k ; row-by-column products in the reference's order (k = 0, 1, 2 with a zero third term)
      double acc = 0;
      acc += (U[3 * r] * sv[0]) * V[3 * c + 0];
      acc += (U[3 * r + 1] * sv[1]) * V[3 * c + 1];
      acc += 0.0 * V[3 * c + 2];
      E[3 * r + c] = acc;
    }
}

// eight_point_E (:609-627) for one index octet; sA = this thread's shared-memory column (stride TPB doubles).
template <int TPB>
__device__ void eight_point_solve(const double2* __restrict__ xi, const double2* __restrict__ xj, const int* __restrict__ idx8, int n,
                                  double* sA, double* __restrict__ E) {
  {
    // Gram matrix of the 8 x 9 design matrix, every entry summed over the rows in order (AtA_from_A :503-517)
    double g[SV_TRI];
#pragma unroll
    for (int e = 0; e < SV_TRI; e++) g[e] = 0.0;
#pragma unroll 1
    for (int r = 0; r < 8; r++) {
      int i = idx8[r];
      i = i < 0 ? 0 : (i >= n ? n - 1 : i);
      const double2 a = xi[i], b = xj[i];
      const double row[9] = {b.x * a.x, b.x * a.y, b.x, b.y * a.x, b.y * a.y, b.y, a.x, a.y, 1.0};
#pragma unroll
      for (int p = 0; p < 9; p++)
#pragma unroll
        for (int q = p; q < 9; q++) g[(p * (17 - p)) / 2 + q] += row[p] * row[q];
    }
#pragma unroll
    for (int e = 0; e < SV_TRI; e++) sA[e * TPB] = g[e];
  }
  double rc[SV_ROT9], rs[SV_ROT9];
  unsigned char rpq[SV_ROT9];
  int nrot = 0;
  for (; nrot < SV_ROT9; nrot++) {
    // largest off-diagonal, first in raster order (strict >); the pivot travels as one code (p * 16 + q)
    int code = 1;
    double big = 0.0;
#pragma unroll
    for (int i = 0; i < 9; i++)
#pragma unroll
      for (int j = i + 1; j < 9; j++) {
        const double v = fabs(sA[((i * (17 - i)) / 2 + j) * TPB]);
        if (v > big) {
          big = v;
          code = i * 16 + j;
        }
      }
    if (big < 1e-12) break;
    const int p = code >> 4, q = code & 15;
    const int ipp = tri_idx(p, p), iqq = tri_idx(q, q), ipq = tri_idx(p, q);
    const double app = sA[ipp * TPB], aqq = sA[iqq * TPB], apq = sA[ipq * TPB];
    double c, s;
    half_angle(2.0 * apq, aqq - app, c, s);
    rc[nrot] = c;
    rs[nrot] = s;
    rpq[nrot] = (unsigned char)code;
    for (int k = 0; k < 9; k++) {
      if (k == p || k == q) continue;
      const int ikp = k < p ? tri_idx(k, p) : tri_idx(p, k), ikq = k < q ? tri_idx(k, q) : tri_idx(q, k);
      const double u = sA[ikp * TPB], v = sA[ikq * TPB];
      sA[ikp * TPB] = c * u - s * v;
      sA[ikq * TPB] = s * u + c * v;
    }
    // the pivot block sees both updates (rows first, then columns)
    const double tpp = c * app - s * apq, tpq = c * apq - s * aqq, tqp = s * app + c * apq, tqq = s * apq + c * aqq;
    sA[ipp * TPB] = c * tpp - s * tpq;
    sA[iqq * TPB] = s * tqp + c * tqq;
    sA[ipq * TPB] = 0.0;
  }
  // eigenvalues ascending: the smallest diagonal entry (first one on ties, like the stable sort of linalg.hpp:190-193)
  int m = 0;
  {
    double best = sA[0];
#pragma unroll
    for (int i = 1; i < 9; i++) {
      const double d = sA[((i * (17 - i)) / 2 + i) * TPB];
      if (d < best) {
        best = d;
        m = i;
      }
    }
  }
  // column m of V = J1 J2 ... Jn, rotations applied backwards to e_m (elements 0..8 of the thread's column reused)
#pragma unroll
  for (int i = 0; i < 9; i++) sA[i * TPB] = i == m ? 1.0 : 0.0;
  for (int r = nrot - 1; r >= 0; r--) {
    const int p = rpq[r] >> 4, q = rpq[r] & 15;
    const double c = rc[r], s = rs[r];
    const double u = sA[p * TPB], v = sA[q * TPB];
    sA[p * TPB] = c * u + s * v;
    sA[q * TPB] = c * v - s * u;
  }
  double E0[9];
#pragma unroll
  for (int i = 0; i < 9; i++) E0[i] = sA[i * TPB];
  rank2_project(E0, E);
}

// Hypothesis h of pair blockIdx.y: octet idx8[pair][h][8], points xi/xj[pair * stride ...], n = npts[pair] (npts == nullptr:
// n_single); pairs with fewer than 8 points produce nothing.
__global__ void __launch_bounds__(SV_TPB, SV_MINB) eight_point_kernel(const double2* __restrict__ xi, const double2* __restrict__ xj,
                                                               size_t pt_stride, const int* __restrict__ npts, int n_single,
                                                               const int* __restrict__ idx8, int H, int h0, int h1,
                                                               double* __restrict__ Eout) {
  extern __shared__ double sv_smem[];
  const int pair = blockIdx.y, hyp = h0 + blockIdx.x * SV_TPB + threadIdx.x;  // hypotheses [h0, h1) of every pair
  const int n = npts ? npts[pair] : n_single;
  if (hyp >= h1 || n < 8) return;
  const size_t ho = (size_t)pair * H + hyp;
  eight_point_solve<SV_TPB>(xi + (size_t)pair * pt_stride, xj + (size_t)pair * pt_stride, idx8 + ho * 8, n, sv_smem + threadIdx.x,
                            Eout + ho * 9);
}

// ---- the same emulation, one WARP per hypothesis (latency: the winner of a pair, the few repeated-index octets of a single call) ----
// Bit-identical to eight_point_solve: every matrix entry sees the same operations on the same operands, only spread over
// lanes.  Per rotation: the 36 off-diagonal entries are searched by 32 lanes (two entries for lanes 0..12) and reduced with
// the sequential scan's rule (largest magnitude, first in raster order; NaN never wins); every lane evaluates the rotation;
// lanes 0..8 update their row's two entries, lane 9 the pivot block, lane 10 records the rotation.  ~550 cycles per
// rotation instead of ~1,500 for the single thread.
constexpr int WJ_WARPS = 4;  // hypotheses per block
struct WarpJacobi {
  double a[SV_TRI];
  double rows[8][9];
  double rc[SV_ROT9], rs[SV_ROT9];
  double vec[9];
  unsigned char rpq[SV_ROT9];
};

__device__ void eight_point_solve_warp(const double2* __restrict__ xi, const double2* __restrict__ xj, const int* __restrict__ idx8, int n,
                                       WarpJacobi& w, double* __restrict__ E, int lane) {
  const unsigned FULL = 0xffffffffu;
  if (lane < 8) {
    int i = idx8[lane];
    i = i < 0 ? 0 : (i >= n ? n - 1 : i);
    const double2 a = xi[i], b = xj[i];
    const double row[9] = {b.x * a.x, b.x * a.y, b.x, b.y * a.x, b.y * a.y, b.y, a.x, a.y, 1.0};
#pragma unroll
    for (int c = 0; c < 9; c++) w.rows[lane][c] = row[c];
  }
  __syncwarp();
  // this lane's packed entries e0 = lane, e1 = lane + 32 (< 45) and their (row, column)
  int ei[2], ej[2];
#pragma unroll
  for (int t = 0; t < 2; t++) {
    const int e = lane + 32 * t;
    int i = 0, base = 0;
    while (i < 8 && e >= base + (9 - i)) {
      base += 9 - i;
      i++;
    }
    ei[t] = i;
    ej[t] = i + (e - base);
  }
#pragma unroll
  for (int t = 0; t < 2; t++) {
    const int e = lane + 32 * t;
    if (e < SV_TRI) {
      double g = 0.0;
      for (int r = 0; r < 8; r++) g += w.rows[r][ei[t]] * w.rows[r][ej[t]];
      w.a[e] = g;
    }
  }
  __syncwarp();
  int nrot = 0;
  for (; nrot < SV_ROT9; nrot++) {
    double best = -1.0;
    int bcode = 1, be = 1 << 20;
#pragma unroll
    for (int t = 0; t < 2; t++) {
      const int e = lane + 32 * t;
      if (e < SV_TRI && ei[t] != ej[t]) {
        double v = fabs(w.a[e]);
        v = v == v ? v : -1.0;
        if (v > best) {  // e0 < e1: the earlier entry keeps a tie
          best = v;
          bcode = ei[t] * 16 + ej[t];
          be = e;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(FULL, best, o);
      const int oc = __shfl_xor_sync(FULL, bcode, o), oe = __shfl_xor_sync(FULL, be, o);
      if (ov > best || (ov == best && oe < be)) {
        best = ov;
        bcode = oc;
        be = oe;
      }
    }
    if (!(best > 0.0) || best < 1e-12) break;  // big starts at 0 and only a strictly larger entry replaces it
    const int p = bcode >> 4, q = bcode & 15;
    const int ipp = tri_idx(p, p), iqq = tri_idx(q, q), ipq = tri_idx(p, q);
    const double app = w.a[ipp], aqq = w.a[iqq], apq = w.a[ipq];
    double c, s;
    half_angle(2.0 * apq, aqq - app, c, s);
    double u = 0.0, v = 0.0;
    int ikp = 0, ikq = 0;
    const bool rowlane = lane < 9 && lane != p && lane != q;
    if (rowlane) {
      const int k = lane;
      ikp = k < p ? tri_idx(k, p) : tri_idx(p, k);
      ikq = k < q ? tri_idx(k, q) : tri_idx(q, k);
      u = w.a[ikp];
      v = w.a[ikq];
    }
    __syncwarp();
    if (rowlane) {
      w.a[ikp] = c * u - s * v;
      w.a[ikq] = s * u + c * v;
    } else if (lane == 9) {
      // the pivot block sees both updates (rows first, then columns)
      const double tpp = c * app - s * apq, tpq = c * apq - s * aqq, tqp = s * app + c * apq, tqq = s * apq + c * aqq;
      w.a[ipp] = c * tpp - s * tpq;
      w.a[iqq] = s * tqp + c * tqq;
      w.a[ipq] = 0.0;
    } else if (lane == 10) {
      w.rc[nrot] = c;
      w.rs[nrot] = s;
      w.rpq[nrot] = (unsigned char)bcode;
    }
    __syncwarp();
  }
  if (lane == 0) {
    int m = 0;
    double bestd = w.a[0];
    for (int i = 1; i < 9; i++) {
      const double d = w.a[tri_idx(i, i)];
      if (d < bestd) {
        bestd = d;
        m = i;
      }
    }
    for (int i = 0; i < 9; i++) w.vec[i] = i == m ? 1.0 : 0.0;
    for (int r = nrot - 1; r >= 0; r--) {
      const int p = w.rpq[r] >> 4, q = w.rpq[r] & 15;
      const double c = w.rc[r], s = w.rs[r];
      const double u = w.vec[p], v = w.vec[q];
      w.vec[p] = c * u + s * v;
      w.vec[q] = c * v - s * u;
    }
    double E0[9];
    for (int i = 0; i < 9; i++) E0[i] = w.vec[i];
    rank2_project(E0, E);
  }
  __syncwarp();
}

// Items: best == nullptr: list[0 .. min(*list_count, max_items)); best != nullptr: the winners of npairs sets.
__global__ void __launch_bounds__(WJ_WARPS * 32) eight_point_warp_kernel(const double2* __restrict__ xi, const double2* __restrict__ xj,
                                                                        size_t pt_stride, const int* __restrict__ npts, int n_single,
                                                                        const int* __restrict__ idx8, int H, const int* __restrict__ list_count,
                                                                        const int* __restrict__ list, int max_items,
                                                                        const int* __restrict__ best, int npairs, double* __restrict__ E) {
  __shared__ WarpJacobi wj[WJ_WARPS];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int total = best ? npairs : *list_count;
  if (!best && total > max_items) total = max_items;
  for (int k = blockIdx.x * WJ_WARPS + wid; k < total; k += gridDim.x * WJ_WARPS) {
    size_t ho;
    int pair;
    if (best) {
      pair = k;
      const int bh = best[2 * pair];
      if (bh < 0 || bh >= H) continue;
      ho = (size_t)pair * H + bh;
    } else {
      ho = (size_t)list[k];
      pair = (int)(ho / (size_t)H);
    }
    const int n = npts ? npts[pair] : n_single;
    if (n < 8) continue;
    eight_point_solve_warp(xi + (size_t)pair * pt_stride, xj + (size_t)pair * pt_stride, idx8 + ho * 8, n, wj[wid], E + ho * 9, lane);
  }
}

// ---- direct null vector: the SCREENING solver of the batched RANSAC stage -----------------------------------------------------
// eight_point_E asks for the unit vector e minimising ||A e|| (A: 8 x 9, one row per correspondence) and gets it as the
// eigenvector of the smallest eigenvalue of A^T A from a Jacobi iteration (<= 120 rotations, absolute 1e-12 stop).  For
// eight points A has an exact null vector, and that is what the iteration converges to.  This kernel computes the same
// vector directly: Householder QR of A^T (9 x 8), column by column (left-looking: the finished reflectors are applied to
// one new column at a time, so only the reflectors - 44 doubles - and one column are live), and e = Q e_9 = H_1 ... H_8 e_9,
// the unit vector orthogonal to every row of A.  ~1,000 FMAs in registers instead of ~50,000 instructions of
// shared-memory Jacobi.  Orthogonal transformations throughout: e is accurate to ~eps * cond(A), the Jacobi result to
// ~eps * cond(A)^2 + 1e-12, so the two agree to the JACOBI result's own error (measured on the test scenes: median 1e-10,
// 99 % below 5e-6 relative, against 1e-13 / 1e-8 for the emulation above), up to the sign of e, which neither the Sampson
// error nor the pose tail sees.  Octets with a repeated index (sampling is with replacement, ~1 % of the hypotheses) have
// a two-dimensional null space: every implementation returns an arbitrary member of it.
// Used for COUNTING only: the winner of every correspondence set is re-solved by the Jacobi emulation (eight_point_winner_
// kernel) before its inlier list, count and pose are produced, so what leaves the stage is the emulation's hypothesis.
// FMAs are used on purpose here (nothing of the reference's operation order is being reproduced).
constexpr int QR_TPB = 128;

__device__ __forceinline__ void qr_null_vector(const double2* __restrict__ xi, const double2* __restrict__ xj, const int* __restrict__ idx8,
                                               int n, double* __restrict__ e) {
  double v[8][9];  // reflector k lives in v[k][k..8]
  double beta[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    int i = idx8[j];
    i = i < 0 ? 0 : (i >= n ? n - 1 : i);
    const double2 a = xi[i], b = xj[i];
    double col[9] = {b.x * a.x, b.x * a.y, b.x, b.y * a.x, b.y * a.y, b.y, a.x, a.y, 1.0};
#pragma unroll
    for (int k = 0; k < j; k++) {
      double s = 0.0;
#pragma unroll
      for (int c = k; c < 9; c++) s = __fma_rn(v[k][c], col[c], s);
      s *= -beta[k];
#pragma unroll
      for (int c = k; c < 9; c++) col[c] = __fma_rn(s, v[k][c], col[c]);
    }
    double sig = 0.0;
#pragma unroll
    for (int c = j; c < 9; c++) sig = __fma_rn(col[c], col[c], sig);
    const double nrm = sqrt(sig);
    const double vj = col[j] + copysign(nrm, col[j]);
    const double den = nrm * fabs(vj);  // = v^T v / 2
    beta[j] = den > 0.0 ? 1.0 / den : 0.0;  // a zero column needs no reflection
    v[j][j] = vj;
#pragma unroll
    for (int c = j + 1; c < 9; c++) v[j][c] = col[c];
  }
  double q[9] = {0, 0, 0, 0, 0, 0, 0, 0, 1.0};
#pragma unroll
  for (int k = 7; k >= 0; k--) {
    double s = 0.0;
#pragma unroll
    for (int c = k; c < 9; c++) s = __fma_rn(v[k][c], q[c], s);
    s *= -beta[k];
#pragma unroll
    for (int c = k; c < 9; c++) q[c] = __fma_rn(s, v[k][c], q[c]);
  }
#pragma unroll
  for (int c = 0; c < 9; c++) e[c] = q[c];
}

__global__ void __launch_bounds__(QR_TPB, 4) eight_point_qr_kernel(const double2* __restrict__ xi, const double2* __restrict__ xj,
                                                                   size_t pt_stride, const int* __restrict__ npts, int n_single,
                                                                   const int* __restrict__ idx8, int H, int h0, int h1,
                                                                   double* __restrict__ Eout, int* __restrict__ rep_count,
                                                                   int* __restrict__ rep_list) {
  const int pair = blockIdx.y, hyp = h0 + blockIdx.x * QR_TPB + threadIdx.x;  // hypotheses [h0, h1) of every pair
  const int n = npts ? npts[pair] : n_single;
  if (hyp >= h1 || n < 8) return;
  const size_t ho = (size_t)pair * H + hyp;
  const int* oct = idx8 + ho * 8;
  // a repeated index: two-dimensional null space, which member comes back is a property of the Jacobi iteration - those
  // octets (about 28 / n of them) are left to the emulation (eight_point_list_kernel)
  int id[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int i = oct[j];
    id[j] = i < 0 ? 0 : (i >= n ? n - 1 : i);
  }
  bool rep = false;
#pragma unroll
  for (int a = 0; a < 8; a++)
#pragma unroll
    for (int b = a + 1; b < 8; b++) rep |= id[a] == id[b];
  if (rep) {
    rep_list[atomicAdd(rep_count, 1)] = (int)ho;
    return;
  }
  double e[9];
  qr_null_vector(xi + (size_t)pair * pt_stride, xj + (size_t)pair * pt_stride, oct, n, e);
  rank2_project(e, Eout + ho * 9);
}

// The Jacobi emulation for the hypotheses on a device-side list (flat indices pair * H + hyp), resident grid.
__global__ void __launch_bounds__(SV_TPB, SV_MINB) eight_point_list_kernel(const double2* __restrict__ xi, const double2* __restrict__ xj,
                                                                    size_t pt_stride, const int* __restrict__ npts, int n_single,
                                                                    const int* __restrict__ idx8, int H, const int* __restrict__ rep_count,
                                                                    const int* __restrict__ rep_list, int skip, double* __restrict__ Eout) {
  extern __shared__ double sv_smem[];
  const int total = *rep_count;
  for (int k = skip + blockIdx.x * SV_TPB + threadIdx.x; k < total; k += gridDim.x * SV_TPB) {
    const size_t ho = (size_t)rep_list[k];
    const int pair = (int)(ho / (size_t)H);
    const int n = npts ? npts[pair] : n_single;
    eight_point_solve<SV_TPB>(xi + (size_t)pair * pt_stride, xj + (size_t)pair * pt_stride, idx8 + ho * 8, n, sv_smem + threadIdx.x,
                              Eout + ho * 9);
  }
}

__global__ void iota_list_kernel(int* count, int* list, int total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *count = total;
  if (i < total) list[i] = i;
}

// Batched triangulate_dlt (:1477-1516), one thread per track.  Same formulas and order as the host
// restatement (host/two_view_host.hpp: triangulate_dlt); the 4x4 Jacobi uses CUDA's trig, hence ~1e-10 relative, not
// bit-identical.
__global__ void __launch_bounds__(128) triangulate_kernel(const double* __restrict__ Kinv, const double* __restrict__ poses,
                                                         const int* __restrict__ ia, const int* __restrict__ ib,
                                                         const double2* __restrict__ ui, const double2* __restrict__ uj, int n, int P,
                                                         double* __restrict__ X) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  double A[16], G[16], Q[16];
  for (int cam = 0; cam < 2; cam++) {
    int pi = cam == 0 ? ia[k] : ib[k];
    pi = pi < 0 ? 0 : (pi >= P ? P - 1 : pi);
    const double* pose = poses + 12 * (size_t)pi;
    const double2 u = cam == 0 ? ui[k] : uj[k];
    const double a = Kinv[0] * u.x + Kinv[1] * u.y + Kinv[2] * 1.0;
    const double b = Kinv[3] * u.x + Kinv[4] * u.y + Kinv[5] * 1.0;
    const double c = Kinv[6] * u.x + Kinv[7] * u.y + Kinv[8] * 1.0;
    const double x = a / c, y = b / c;
    double Rw[9];
    for (int r = 0; r < 3; r++)
      for (int cc = 0; cc < 3; cc++) Rw[3 * r + cc] = pose[3 * cc + r];
    const double tx = -(Rw[0] * pose[9] + Rw[1] * pose[10] + Rw[2] * pose[11]);
    const double ty = -(Rw[3] * pose[9] + Rw[4] * pose[10] + Rw[5] * pose[11]);
    const double tz = -(Rw[6] * pose[9] + Rw[7] * pose[10] + Rw[8] * pose[11]);
    double* o = A + 8 * cam;
    o[0] = x * Rw[6] - Rw[0]; o[1] = x * Rw[7] - Rw[1]; o[2] = x * Rw[8] - Rw[2]; o[3] = x * tz - tx;
    o[4] = y * Rw[6] - Rw[3]; o[5] = y * Rw[7] - Rw[4]; o[6] = y * Rw[8] - Rw[5]; o[7] = y * tz - ty;
  }
  for (int i = 0; i < 4; i++)
    for (int j = i; j < 4; j++) {
      double s = 0;
      for (int r = 0; r < 4; r++) s += A[r * 4 + i] * A[r * 4 + j];
      G[i * 4 + j] = s;
      G[j * 4 + i] = s;
    }
  jacobi_dev<4>(G, Q, 80);
  int m = 0;
  for (int i = 1; i < 4; i++)
    if (G[i * 5] < G[m * 5]) m = i;
  const double ww = Q[12 + m];
  X[3 * (size_t)k] = Q[m] / ww;
  X[3 * (size_t)k + 1] = Q[4 + m] / ww;
  X[3 * (size_t)k + 2] = Q[8 + m] / ww;
}

}  // namespace

// ---- pose recovery, batched (find_E_ransac tail :680-760): one block per correspondence set --------------------------------
// E -> (R1, R2) x (+t, -t) by svd3, cheirality vote over the first min(best_n, 20) inliers (two-view DLT through a 4x4
// Jacobi each), first candidate with the strictly largest vote wins.  Device trigonometry-free Jacobi (half-angle form)
// for svd3, CUDA trig in jacobi_dev<4>: R, t agree with the host tail to ~1e-10, not bit for bit.
// status[pair] must be 2 (a pose is wanted); E = bestE [npairs][9]; writes R [npairs][9], t [npairs][3].
__global__ void __launch_bounds__(96) pose_kernel(const double2* __restrict__ xi, const double2* __restrict__ xj, size_t pt_stride,
                                                  const int* __restrict__ status, const int* __restrict__ best,
                                                  const int* __restrict__ inl, const double* __restrict__ bestE, double* __restrict__ Rout,
                                                  double* __restrict__ tout) {
  __shared__ double sR[2][9];
  __shared__ double st[3];
  __shared__ int votes[4];
  const int pair = blockIdx.x, tid = threadIdx.x;
  if (status[pair] != 2) return;
  xi += (size_t)pair * pt_stride;
  xj += (size_t)pair * pt_stride;
  inl += (size_t)pair * pt_stride;
  if (tid < 4) votes[tid] = 0;
  if (tid == 0) {
    double E[9], U[9], sv[3], V[9];
    for (int i = 0; i < 9; i++) E[i] = bestE[(size_t)pair * 9 + i];
    svd3_dev(E, U, sv, V);
    // R1 = (U W) V^T, R2 = (U W^T) V^T with W = [0 -1 0; 1 0 0; 0 0 1]: U W = [u1, -u0, u2], U W^T = [-u1, u0, u2] (columns)
    for (int which = 0; which < 2; which++) {
      const double sg = which == 0 ? 1.0 : -1.0;
      double UW[9];
      for (int r = 0; r < 3; r++) {
        UW[3 * r + 0] = sg * U[3 * r + 1];
        UW[3 * r + 1] = -sg * U[3 * r + 0];
        UW[3 * r + 2] = U[3 * r + 2];
      }
      double R[9];
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
          double acc = 0;
          for (int k = 0; k < 3; k++) acc += UW[3 * r + k] * V[3 * c + k];  // V^T(k, c) = V(c, k)
          R[3 * r + c] = acc;
        }
      const double det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) + R[2] * (R[3] * R[7] - R[4] * R[6]);
      for (int i = 0; i < 9; i++) sR[which][i] = det < 0 ? -R[i] : R[i];
    }
    double tx = U[2], ty = U[5], tz = U[8];
    unit3(tx, ty, tz);
    st[0] = tx; st[1] = ty; st[2] = tz;
  }
  __syncthreads();
  const int bn = best[2 * pair + 1];
  const int M = bn < 20 ? bn : 20;
  if (tid < 4 * M) {
    const int cand = tid / M, k = tid - cand * M;
    const double* R = sR[cand >> 1];
    const double sg = (cand & 1) ? -1.0 : 1.0;
    const double t0 = sg * st[0], t1 = sg * st[1], t2 = sg * st[2];
    const int i = inl[k];
    const double2 x = xi[i], xp = xj[i];
    double A[16], G[16], Q[16];
    A[0] = -1; A[1] = 0; A[2] = x.x; A[3] = 0;
    A[4] = 0; A[5] = -1; A[6] = x.y; A[7] = 0;
    A[8] = xp.x * R[6] - R[0]; A[9] = xp.x * R[7] - R[1]; A[10] = xp.x * R[8] - R[2]; A[11] = xp.x * t2 - t0;
    A[12] = xp.y * R[6] - R[3]; A[13] = xp.y * R[7] - R[4]; A[14] = xp.y * R[8] - R[5]; A[15] = xp.y * t2 - t1;
    for (int a = 0; a < 4; a++)
      for (int b = a; b < 4; b++) {
        double acc = 0;
        for (int r = 0; r < 4; r++) acc += A[r * 4 + a] * A[r * 4 + b];
        G[a * 4 + b] = acc;
        G[b * 4 + a] = acc;
      }
    jacobi_dev<4>(G, Q, 80);
    int m = 0;
    for (int a = 1; a < 4; a++)
      if (G[a * 5] < G[m * 5]) m = a;
    const double w = Q[12 + m];
    const double X0 = Q[m] / w, X1 = Q[4 + m] / w, X2 = Q[8 + m] / w;
    const double z2 = (R[6] * X0 + R[7] * X1 + R[8] * X2) + t2;
    if (X2 > 0 && z2 > 0) atomicAdd(&votes[cand], 1);
  }
  __syncthreads();
  if (tid == 0) {
    int bi = 0, bok = -1;
    for (int c = 0; c < 4; c++)
      if (votes[c] > bok) {
        bok = votes[c];
        bi = c;
      }
    const double sg = (bi & 1) ? -1.0 : 1.0;
    for (int i = 0; i < 9; i++) Rout[(size_t)pair * 9 + i] = sR[bi >> 1][i];
    for (int i = 0; i < 3; i++) tout[(size_t)pair * 3 + i] = sg * st[i];
  }
}

int sfm_pose_batched(sfmgpu_ctx* ctx, const double2* xi, const double2* xj, size_t pt_stride, int npairs, const int* status,
                     const int* best, const int* inl, const double* bestE, double* R, double* t) {
  if (npairs <= 0) return 0;
  SFM_LAUNCH(ctx, pose_kernel, npairs, 96, 0, xi, xj, pt_stride, status, best, inl, bestE, R, t);
  return 0;
}

// Whether a launch of npairs x H hypotheses is counted with the screening solver (solver mode 1).  Up to one wave of the
// emulation kernel (n_sm x 4 blocks x 128 hypotheses: 75,776 on B200) the emulation for EVERY octet is the faster choice: its
// time is one thread's latency (~150 us) either way, and the screening path adds the list and winner launches behind it
// (measured, 2,500 hypotheses x 2,200 points: 0.31 ms against 0.42 ms for the whole solve + score call).
bool sfm_solver_screens(const sfmgpu_ctx* ctx, int npairs, int H) {
  if (ctx->solver_mode == 3) return true;  // tests: always
  return ctx->solver_mode == 1 && (long long)npairs * H > (long long)ctx->n_sm * SV_MINB * SV_TPB;
}

// Hypotheses for `npairs` correspondence sets in one launch: xi/xj + pair * pt_stride, npts[pair] points (npts may be null:
// n_single for all), idx8 [npairs][H][8], Eout [npairs][H][9].  screen != 0: the direct null-vector solver (counting only,
// see eight_point_qr_kernel; the caller re-solves the winners with sfm_eight_point_winners).
int sfm_eight_point_batched(sfmgpu_ctx* ctx, const double2* xi, const double2* xj, size_t pt_stride, const int* npts, int n_single,
                            int npairs, const int* idx8, int H, double* Eout, int screen) {
  if (npairs <= 0 || H <= 0) return 0;
  return sfm_eight_point_range(ctx, xi, xj, pt_stride, npts, n_single, npairs, idx8, H, 0, H, Eout, screen && sfm_solver_screens(ctx, npairs, H));
}

// The same for the hypotheses [h0, h1) of every set (idx8 / Eout keep their [npairs][H] layout); `screen` is taken as given
// (the caller decides once per launch set, so that the parts of a set are solved the same way).  Solver mode 2 (tests)
// supports the full range only.
int sfm_eight_point_range(sfmgpu_ctx* ctx, const double2* xi, const double2* xj, size_t pt_stride, const int* npts, int n_single,
                          int npairs, const int* idx8, int H, int h0, int h1, double* Eout, int screen) {
  if (npairs <= 0 || H <= 0 || h1 <= h0) return 0;
  if (h0 < 0 || h1 > H) return sfm_fail(ctx, SFMGPU_E_ARG, "eight_point: hypothesis range [%d, %d) outside [0, %d)", h0, h1, H);
  const int Hr = h1 - h0;
  if (screen) {
    if ((long long)npairs * H > 0x7fffffffll) return sfm_fail(ctx, SFMGPU_E_ARG, "eight_point: %d x %d hypotheses in one launch", npairs, H);
    SFM_TRY(sfm_reserve(ctx, ctx->sv_list, ((size_t)npairs * H + 64) * sizeof(int)));
    int* rep_count = (int*)ctx->sv_list.p;
    int* rep_list = rep_count + 64;
    SFM_CUDA(ctx, cudaMemsetAsync(rep_count, 0, sizeof(int), ctx->stream));
    SFM_LAUNCH(ctx, eight_point_qr_kernel, dim3(sfm_cdiv(Hr, QR_TPB), npairs), QR_TPB, 0, xi, xj, pt_stride, npts, n_single, idx8, H, h0, h1,
               Eout, rep_count, rep_list);
    static const int list_cfg = sfm_next_cfg_id();
    const size_t lsmem = (size_t)SV_TRI * SV_TPB * sizeof(double);
    SFM_SMEM_OPTIN(ctx, list_cfg, eight_point_list_kernel, lsmem);
    const long long want = ((long long)npairs * Hr + SV_TPB - 1) / SV_TPB, cap = (long long)ctx->n_sm * SV_MINB;
    // a single correspondence set has a few dozen such octets: the first 512 go to the warp-per-hypothesis kernel (latency)
    const int warp_items = npairs == 1 ? 512 : 0;
    if (warp_items)
      SFM_LAUNCH(ctx, eight_point_warp_kernel, warp_items / WJ_WARPS, WJ_WARPS * 32, 0, xi, xj, pt_stride, npts, n_single, idx8, H,
                 (const int*)rep_count, (const int*)rep_list, warp_items, (const int*)nullptr, 0, Eout);
    SFM_LAUNCH(ctx, eight_point_list_kernel, (unsigned)(want < cap ? want : cap), SV_TPB, lsmem, xi, xj, pt_stride, npts, n_single, idx8, H,
               (const int*)rep_count, (const int*)rep_list, warp_items, Eout);
    return 0;
  }
  if (screen == 0 && ctx->solver_mode == 2) {  // tests: every octet through the warp-per-hypothesis emulation
    if (h0 != 0 || h1 != H) return sfm_fail(ctx, SFMGPU_E_ARG, "eight_point: solver mode 2 solves whole sets only");
    if ((long long)npairs * H > 0x7fffffffll) return sfm_fail(ctx, SFMGPU_E_ARG, "eight_point: %d x %d hypotheses in one launch", npairs, H);
    SFM_TRY(sfm_reserve(ctx, ctx->sv_list, ((size_t)npairs * H + 64) * sizeof(int)));
    int* rep_count = (int*)ctx->sv_list.p;
    int* rep_list = rep_count + 64;
    const int total = npairs * H;
    SFM_LAUNCH(ctx, iota_list_kernel, sfm_cdiv(total, 256), 256, 0, rep_count, rep_list, total);
    SFM_LAUNCH(ctx, eight_point_warp_kernel, (unsigned)(ctx->n_sm * 8), WJ_WARPS * 32, 0, xi, xj, pt_stride, npts, n_single, idx8, H,
               (const int*)rep_count, (const int*)rep_list, total, (const int*)nullptr, 0, Eout);
    return 0;
  }
  static const int cfg_id = sfm_next_cfg_id();
  const size_t smem = (size_t)SV_TRI * SV_TPB * sizeof(double);
  SFM_SMEM_OPTIN(ctx, cfg_id, eight_point_kernel, smem);
  SFM_LAUNCH(ctx, eight_point_kernel, dim3(sfm_cdiv(Hr, SV_TPB), npairs), SV_TPB, smem, xi, xj, pt_stride, npts, n_single, idx8, H, h0, h1, Eout);
  return 0;
}

// E[pair][best[2 * pair]] := the Jacobi emulation's hypothesis for that octet (no-op for sets without a winner).
int sfm_eight_point_winners(sfmgpu_ctx* ctx, const double2* xi, const double2* xj, size_t pt_stride, const int* npts, int n_single,
                            int npairs, const int* idx8, int H, const int* best, double* E) {
  if (npairs <= 0 || H <= 0) return 0;
  SFM_LAUNCH(ctx, eight_point_warp_kernel, sfm_cdiv(npairs, WJ_WARPS), WJ_WARPS * 32, 0, xi, xj, pt_stride, npts, n_single, idx8, H,
             (const int*)nullptr, (const int*)nullptr, 0, best, npairs, E);
  return 0;
}

extern "C" int sfmgpu_triangulate_dlt(sfmgpu_ctx* ctx, const double* K, const double* poses, int P, const int32_t* ia, const int32_t* ib,
                                      const double* ui_xy, const double* uj_xy, int n, double* X_out) {
  SFM_ENTER(ctx);
  if (!ctx || n < 0 || P < 0) return sfm_fail(ctx, SFMGPU_E_ARG, "triangulate_dlt: bad sizes");
  if (n == 0) return 0;
  if (!K || !poses || P < 1 || !ia || !ib || !ui_xy || !uj_xy || !X_out) return sfm_fail(ctx, SFMGPU_E_ARG, "triangulate_dlt: null pointer");
  // K^-1 on the host, exactly as invert_K :471-486 (adjugate / determinant)
  const double d = K[0] * (K[4] * K[8] - K[5] * K[7]) - K[1] * (K[3] * K[8] - K[5] * K[6]) + K[2] * (K[3] * K[7] - K[4] * K[6]);
  if (fabs(d) < 1e-12) return sfm_fail(ctx, SFMGPU_E_ARG, "triangulate_dlt: singular K");
  const double Ki[9] = {(K[4] * K[8] - K[5] * K[7]) / d,  -(K[1] * K[8] - K[2] * K[7]) / d, (K[1] * K[5] - K[2] * K[4]) / d,
                        -(K[3] * K[8] - K[5] * K[6]) / d, (K[0] * K[8] - K[2] * K[6]) / d,  -(K[0] * K[5] - K[2] * K[3]) / d,
                        (K[3] * K[7] - K[4] * K[6]) / d,  -(K[0] * K[7] - K[1] * K[6]) / d, (K[0] * K[4] - K[1] * K[3]) / d};
  const size_t bK = 128, bP = ((size_t)P * 96 + 255) & ~(size_t)255, bI = ((size_t)n * 4 + 255) & ~(size_t)255,
               bU = ((size_t)n * 16 + 255) & ~(size_t)255, bX = (size_t)n * 24;
  SFM_TRY(sfm_reserve(ctx, ctx->cs_work, bK + bP + 2 * bI + 2 * bU + bX + 256));
  char* base = (char*)ctx->cs_work.p;
  double* dK = (double*)base;
  double* dP = (double*)(base + bK);
  int* dia = (int*)(base + bK + bP);
  int* dib = (int*)(base + bK + bP + bI);
  double2* dui = (double2*)(base + bK + bP + 2 * bI);
  double2* duj = (double2*)(base + bK + bP + 2 * bI + bU);
  double* dX = (double*)(base + bK + bP + 2 * bI + 2 * bU);
  SFM_CUDA(ctx, cudaMemcpyAsync(dK, Ki, sizeof Ki, cudaMemcpyHostToDevice, ctx->stream));
  SFM_CUDA(ctx, cudaMemcpyAsync(dP, poses, (size_t)P * 96, cudaMemcpyHostToDevice, ctx->stream));
  SFM_CUDA(ctx, cudaMemcpyAsync(dia, ia, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
  SFM_CUDA(ctx, cudaMemcpyAsync(dib, ib, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
  SFM_CUDA(ctx, cudaMemcpyAsync(dui, ui_xy, (size_t)n * 16, cudaMemcpyHostToDevice, ctx->stream));
  SFM_CUDA(ctx, cudaMemcpyAsync(duj, uj_xy, (size_t)n * 16, cudaMemcpyHostToDevice, ctx->stream));
  SFM_LAUNCH(ctx, triangulate_kernel, sfm_cdiv(n, 128), 128, 0, (const double*)dK, (const double*)dP, (const int*)dia, (const int*)dib,
             (const double2*)dui, (const double2*)duj, n, P, dX);
  SFM_CUDA(ctx, cudaMemcpyAsync(X_out, dX, (size_t)n * 24, cudaMemcpyDeviceToHost, ctx->stream));
  SFM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

// staged: the caller synchronises the stream before it returns, so the points may go through the context's pinned staging
// area (one memcpy + truly asynchronous copies instead of two staged pageable copies)
static int hypotheses_resident(sfmgpu_ctx* ctx, const double* xi_xy, const double* xj_xy, int n, const int32_t* idx8, int H, int screen,
                               bool staged = false) {
  if (staged && n > 0) {
    const size_t pb = (size_t)n * 16;
    SFM_TRY(sfm_pinned(ctx, 2 * pb + 256));
    memcpy(ctx->pinned, xi_xy, pb);
    memcpy((char*)ctx->pinned + pb, xj_xy, pb);
    xi_xy = (const double*)ctx->pinned;
    xj_xy = (const double*)((char*)ctx->pinned + pb);
  }
  SFM_TRY(sfmgpu_ransac_upload(ctx, xi_xy, xj_xy, n, nullptr, 0));  // points resident, room for 0 hypotheses
  SFM_TRY(sfm_reserve(ctx, ctx->rs_E, (size_t)(H + 1) * 72));
  SFM_TRY(sfm_reserve(ctx, ctx->rs_counts, (size_t)(H + 1) * 4));
  SFM_TRY(sfm_reserve(ctx, ctx->rs_idx8, (size_t)(H + 1) * 32 + 64));
  ctx->rs_H = H;
  ctx->rs_screened = false;
  if (H == 0) return 0;
  if (idx8) {
    SFM_CUDA(ctx, cudaMemcpyAsync(ctx->rs_idx8.p, idx8, (size_t)H * 32, cudaMemcpyHostToDevice, ctx->stream));
  } else {  // the reference's seeded sampling on the device; the flag word sits behind the octets
    SFM_TRY(sfm_sample_octets(ctx, n, H, (int*)ctx->rs_idx8.p, (int*)ctx->rs_idx8.p + (size_t)H * 8 + 8));
  }
  SFM_TRY(sfm_eight_point_batched(ctx, (const double2*)ctx->rs_xi.p, (const double2*)ctx->rs_xj.p, 0, nullptr, n, 1,
                                  (const int*)ctx->rs_idx8.p, H, (double*)ctx->rs_E.p, screen));
  ctx->rs_screened = screen != 0 && sfm_solver_screens(ctx, 1, H);
  return 0;
}

// (winner, count, sampler flag, -, E of the winner or the zero matrix) in one block: one read-back instead of two round trips
struct SolveScoreResult {
  int bh, bn, flag, pad;
  double E[9];
};

__global__ void solve_score_pack_kernel(const int* __restrict__ best, const int* __restrict__ flag, const double* __restrict__ E,
                                        SolveScoreResult* __restrict__ out) {
  const int i = threadIdx.x;
  const int bh = best[0];
  if (i == 0) {
    out->bh = bh;
    out->bn = best[1];
    out->flag = flag ? *flag : 0;
    out->pad = 0;
  }
  if (i < 9) out->E[i] = bh >= 0 ? E[9 * (size_t)bh + i] : 0.0;
}

extern "C" int sfmgpu_ransac_hypotheses(sfmgpu_ctx* ctx, const double* xi_xy, const double* xj_xy, int n, const int32_t* idx8, int H,
                                        double* E_out) {
  SFM_ENTER(ctx);
  if (!ctx || n < 0 || H < 0) return sfm_fail(ctx, SFMGPU_E_ARG, "ransac_hypotheses: bad sizes");
  if (n < 1 && H > 0) return sfm_fail(ctx, SFMGPU_E_ARG, "ransac_hypotheses: no correspondences");
  if ((n > 0 && (!xi_xy || !xj_xy)) || (H > 0 && !idx8)) return sfm_fail(ctx, SFMGPU_E_ARG, "ransac_hypotheses: null pointer");
  // the caller wants the hypotheses themselves: the Jacobi emulation for every octet; resident only (E_out == NULL: they
  // are there to be scored): the context's solver mode, the winner is re-solved when scored
  SFM_TRY(hypotheses_resident(ctx, xi_xy, xj_xy, n, idx8, H, E_out ? 0 : 1));
  if (E_out && H > 0) {
    SFM_CUDA(ctx, cudaMemcpyAsync(E_out, ctx->rs_E.p, (size_t)H * 72, cudaMemcpyDeviceToHost, ctx->stream));
    SFM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return 0;
}

extern "C" int sfmgpu_ransac_solve_score(sfmgpu_ctx* ctx, const double* xi_xy, const double* xj_xy, int n, const int32_t* idx8, int H,
                                         double thr, int* best_h, int* best_n, double* best_E, int32_t* best_inl) {
  SFM_ENTER(ctx);
  if (!ctx || n < 0 || H < 0) return sfm_fail(ctx, SFMGPU_E_ARG, "ransac_solve_score: bad sizes");
  if (n < 1 && H > 0) return sfm_fail(ctx, SFMGPU_E_ARG, "ransac_solve_score: no correspondences");
  if (n > 0 && (!xi_xy || !xj_xy)) return sfm_fail(ctx, SFMGPU_E_ARG, "ransac_solve_score: null pointer");
  if (!idx8 && H > 0 && n < 1) return sfm_fail(ctx, SFMGPU_E_ARG, "ransac_solve_score: no correspondences to sample from");
  const size_t pb = (size_t)n * 16, res_off = (2 * pb + 255) & ~(size_t)255;
  SFM_TRY(sfm_pinned(ctx, res_off + sizeof(SolveScoreResult) + (size_t)n * 4 + 256));  // staging of the points + the results
  SFM_TRY(hypotheses_resident(ctx, xi_xy, xj_xy, n, idx8, H, 1, true));
  SFM_TRY(sfmgpu_ransac_score_resident(ctx, thr, nullptr, nullptr));
  // winner, count (of the re-solved winner when the counts were screened), its E, the sampler's flag: one block, one sync
  SFM_TRY(sfm_reserve(ctx, ctx->misc, 256));
  SolveScoreResult* d_res = (SolveScoreResult*)ctx->misc.p;
  SolveScoreResult* h_res = (SolveScoreResult*)((char*)ctx->pinned + res_off);
  int* h_inl = (int*)(h_res + 1);
  SFM_LAUNCH(ctx, solve_score_pack_kernel, 1, 32, 0, (const int*)ctx->rs_best.p,
             (!idx8 && H > 0) ? (const int*)ctx->rs_idx8.p + (size_t)H * 8 + 8 : (const int*)nullptr, (const double*)ctx->rs_E.p, d_res);
  SFM_CUDA(ctx, cudaMemcpyAsync(h_res, d_res, sizeof(SolveScoreResult), cudaMemcpyDeviceToHost, ctx->stream));
  if (best_inl && n > 0) SFM_CUDA(ctx, cudaMemcpyAsync(h_inl, ctx->rs_inl.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  SFM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (h_res->flag) return sfm_fail(ctx, SFMGPU_E_CAPACITY, "ransac_solve_score: the random stream was too short (internal)");
  const int bh = h_res->bh, bn = h_res->bn;
  if (best_h) *best_h = bh;
  if (best_n) *best_n = bn;
  if (best_E) memcpy(best_E, h_res->E, sizeof h_res->E);
  if (best_inl && bh >= 0 && bn > 0) memcpy(best_inl, h_inl, (size_t)bn * 4);
  return 0;
}

extern "C" int sfmgpu_solver_set_mode(sfmgpu_ctx* ctx, int mode) {
  if (!ctx || mode < 0 || mode > 3) return sfm_fail(ctx, SFMGPU_E_ARG, "solver_set_mode: mode must be 0 ... 3");
  ctx->solver_mode = mode;
  return 0;
}
