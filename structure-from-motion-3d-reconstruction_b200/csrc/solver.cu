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

// eight_point_E for hypothesis h: rows of A from the 8 sampled correspondences, AtA, smallest eigenvector, rank 2.
__global__ void __launch_bounds__(64) eight_point_kernel(const double2* __restrict__ xi, const double2* __restrict__ xj,
                                                        const int* __restrict__ idx8, int H, int n, double* __restrict__ Eout) {
  const int hyp = blockIdx.x * blockDim.x + threadIdx.x;
  if (hyp >= H) return;
  double A[72], G[81], Q[81];
  for (int r = 0; r < 8; r++) {
    int i = idx8[(size_t)hyp * 8 + r];
    i = i < 0 ? 0 : (i >= n ? n - 1 : i);
    const double2 a = xi[i], b = xj[i];
    const double x = a.x, y = a.y, xp = b.x, yp = b.y;
    double* row = A + r * 9;
    row[0] = xp * x; row[1] = xp * y; row[2] = xp;
    row[3] = yp * x; row[4] = yp * y; row[5] = yp;
    row[6] = x;      row[7] = y;      row[8] = 1.0;
  }
  for (int i = 0; i < 9; i++)
    for (int j = i; j < 9; j++) {
      double s = 0;
      for (int r = 0; r < 8; r++) s += A[r * 9 + i] * A[r * 9 + j];
      G[i * 9 + j] = s;
      G[j * 9 + i] = s;
    }
  jacobi_dev<9>(G, Q, 120);
  // eigenvalues ascending: column of the smallest diagonal entry (first one on ties, like a stable sort)
  int m = 0;
  for (int i = 1; i < 9; i++)
    if (G[i * 9 + i] < G[m * 9 + m]) m = i;
  double E0[9];
  for (int r = 0; r < 9; r++) E0[r] = Q[r * 9 + m];

  // svd3(E0) through eig(E0^T E0); singular values descending (stable), U re-orthonormalised, u2 = u0 x u1
  double G3[9], V3[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) {
      double t = 0;
      for (int k = 0; k < 3; k++) t += E0[3 * k + r] * E0[3 * k + c];
      G3[3 * r + c] = t;
    }
  jacobi_dev<3>(G3, V3, 80);
  // jacobi's ascending order first (stable on the diagonal), then the descending order of sqrt(max(0, w))
  int asc[3] = {0, 1, 2};
  for (int i = 1; i < 3; i++)
    for (int j = i; j > 0 && G3[asc[j] * 4] < G3[asc[j - 1] * 4]; j--) {
      const int t = asc[j];
      asc[j] = asc[j - 1];
      asc[j - 1] = t;
    }
  double sv[3];
  for (int c = 0; c < 3; c++) sv[c] = sqrt(fmax(0.0, G3[asc[c] * 4]));
  int ord[3] = {0, 1, 2};
  for (int i = 1; i < 3; i++)
    for (int j = i; j > 0 && sv[ord[j]] > sv[ord[j - 1]]; j--) {
      const int t = ord[j];
      ord[j] = ord[j - 1];
      ord[j - 1] = t;
    }
  double s[3], V[9];
  for (int c = 0; c < 3; c++) {
    s[c] = sv[ord[c]];
    for (int r = 0; r < 3; r++) V[3 * r + c] = V3[3 * r + asc[ord[c]]];
  }
  double u[3][3];
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
  // E = U diag(s0, s1, 0) V^T: the third column of U never contributes
  double* E = Eout + (size_t)hyp * 9;
  const double U0[3] = {u0x, u0y, u0z}, U1[3] = {u1x, u1y, u1z};
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) {
      // (U S)(r,k) = U(r,k) s_k ; row-by-column products in the reference's order (k = 0, 1, 2 with a zero third term)
      double acc = 0;
      acc += (U0[r] * s[0]) * V[3 * c + 0];
      acc += (U1[r] * s[1]) * V[3 * c + 1];
      acc += 0.0 * V[3 * c + 2];
      E[3 * r + c] = acc;
    }
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

extern "C" int sfmgpu_ransac_hypotheses(sfmgpu_ctx* ctx, const double* xi_xy, const double* xj_xy, int n, const int32_t* idx8, int H,
                                        double* E_out) {
  SFM_ENTER(ctx);
  if (!ctx || n < 0 || H < 0) return sfm_fail(ctx, SFMGPU_E_ARG, "ransac_hypotheses: bad sizes");
  if (n < 1 && H > 0) return sfm_fail(ctx, SFMGPU_E_ARG, "ransac_hypotheses: no correspondences");
  if ((n > 0 && (!xi_xy || !xj_xy)) || (H > 0 && !idx8)) return sfm_fail(ctx, SFMGPU_E_ARG, "ransac_hypotheses: null pointer");
  SFM_TRY(sfmgpu_ransac_upload(ctx, xi_xy, xj_xy, n, nullptr, 0));  // points resident, room for 0 hypotheses
  SFM_TRY(sfm_reserve(ctx, ctx->rs_E, (size_t)(H + 1) * 72));
  SFM_TRY(sfm_reserve(ctx, ctx->rs_counts, (size_t)(H + 1) * 4));
  SFM_TRY(sfm_reserve(ctx, ctx->rs_idx8, (size_t)(H + 1) * 32));
  ctx->rs_H = H;
  if (H == 0) return 0;
  SFM_CUDA(ctx, cudaMemcpyAsync(ctx->rs_idx8.p, idx8, (size_t)H * 32, cudaMemcpyHostToDevice, ctx->stream));
  SFM_LAUNCH(ctx, eight_point_kernel, sfm_cdiv(H, 64), 64, 0, (const double2*)ctx->rs_xi.p, (const double2*)ctx->rs_xj.p,
             (const int*)ctx->rs_idx8.p, H, n, (double*)ctx->rs_E.p);
  if (E_out) {
    SFM_CUDA(ctx, cudaMemcpyAsync(E_out, ctx->rs_E.p, (size_t)H * 72, cudaMemcpyDeviceToHost, ctx->stream));
    SFM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return 0;
}
