// oracle/sfm_oracle.cpp — TEST INFRASTRUCTURE, not product code.
//
// CPU restatement ("port") of the reference front end, written from the reference's behaviour:
//   cpp/src/templering_sfm.cpp :183-198 (bilinear sampler), :200-232 (pyramid), :237-302 (Shi-Tomasi +
//   greedy NMS), :323-466 (KLT tracker), :471-501 (K^-1 normalisation), :503-627 (8-point solver),
//   :629-638 (Sampson error), :646-761 (RANSAC + pose recovery), :1836-1857 (two-view front end),
//   cpp/include/linalg.hpp :133-201 (Jacobi eigen-solver).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
//
// Parity pinning: the reference ships no tests or golden vectors, so this restatement is
// pinned against the reference ITSELF: oracle/_ref/libsfmref.so (the unmodified reference TU compiled where
// it lies) on the cases in tests/test_oracle_vs_ref.py, and against tests/golden/*.npz generated from
// that library by tests/golden/gen_golden.py.
//
// Third-party behaviour relied on (same as the reference): libstdc++ 13 std::sort tie permutation,
// std::mt19937 + std::uniform_int_distribution<int>, libm sqrt/hypot/floor/atan2/cos/sin.
// Build WITHOUT -ffast-math / -march=native (no FMA contraction), see oracle/Makefile.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

namespace orc {

struct Img {
  int w = 0, h = 0;
  const uint8_t* p = nullptr;
  int at(int x, int y) const { return p[(size_t)y * w + x]; }
};

// ---- sampler (:183-198): any out-of-range tap zeroes the whole sample -----------------------------
static inline double bilerp(const Img& im, double x, double y) {
  const int x0 = (int)std::floor(x), y0 = (int)std::floor(y);
  if (x0 < 0 || y0 < 0 || x0 + 1 >= im.w || y0 + 1 >= im.h) return 0.0;
  const double fx = x - x0, fy = y - y0;
  const double a = im.at(x0, y0), b = im.at(x0 + 1, y0), c = im.at(x0, y0 + 1), d = im.at(x0 + 1, y0 + 1);
  const double top = a * (1 - fx) + b * fx;
  const double bot = c * (1 - fx) + d * fx;
  return top * (1 - fy) + bot * fy;
}

// ---- pyramid (:200-232): truncating 2x2 box; floor(w/2) x floor(h/2) -------------------------------
static void halve(const uint8_t* src, int w, int h, uint8_t* dst) {
  const int ow = w / 2, oh = h / 2;
  for (int y = 0; y < oh; y++)
    for (int x = 0; x < ow; x++) {
      const uint8_t* r0 = src + (size_t)(2 * y) * w + 2 * x;
      const uint8_t* r1 = r0 + w;
      dst[(size_t)y * ow + x] = (uint8_t)((r0[0] + r0[1] + r1[0] + r1[1]) / 4);
    }
}

struct Pyr {
  std::vector<std::vector<uint8_t>> store;  // levels 1..
  std::vector<Img> lvl;
};
static Pyr make_pyr(const uint8_t* pix, int w, int h, int levels) {
  Pyr P;
  P.lvl.push_back(Img{w, h, pix});
  P.store.reserve(levels);
  for (int l = 1; l < levels; l++) {
    const Img s = P.lvl.back();
    P.store.emplace_back((size_t)(s.w / 2) * (s.h / 2));
    halve(s.p, s.w, s.h, P.store.back().data());
    P.lvl.push_back(Img{s.w / 2, s.h / 2, P.store.back().data()});
  }
  return P;
}

// ---- Shi-Tomasi score (:242-272) --------------------------------------------------------------------
// Gradients are central differences with clamped taps, window radius 2, border (2 px) stays 0.
static void score_map(const Img& im, std::vector<double>& score) {
  const int w = im.w, h = im.h;
  score.assign((size_t)w * h, 0.0);
  for (int y = 2; y < h - 2; y++)
    for (int x = 2; x < w - 2; x++) {
      double sxx = 0, sxy = 0, syy = 0;
      for (int v = y - 2; v <= y + 2; v++)
        for (int u = x - 2; u <= x + 2; u++) {
          const int ul = u > 0 ? u - 1 : 0, ur = u < w - 1 ? u + 1 : w - 1;
          const int vu = v > 0 ? v - 1 : 0, vd = v < h - 1 ? v + 1 : h - 1;
          const double gx = 0.5 * ((double)im.at(ur, v) - (double)im.at(ul, v));
          const double gy = 0.5 * ((double)im.at(u, vd) - (double)im.at(u, vu));
          sxx += gx * gx;
          sxy += gx * gy;
          syy += gy * gy;
        }
      const double tr = sxx + syy;
      const double det = sxx * syy - sxy * sxy;
      const double disc = std::max(0.0, tr * tr - 4.0 * det);
      score[(size_t)y * w + x] = 0.5 * (tr - std::sqrt(disc));
    }
}

struct Cand {
  int x, y;
  double s;
};

// Candidates in raster order (:274-285), then libstdc++ std::sort by score descending (:286).
static void candidates(const Img& im, double quality, std::vector<Cand>& c, double* max_out, bool sorted) {
  std::vector<double> score;
  score_map(im, score);
  double mx = score.empty() ? 0.0 : *std::max_element(score.begin(), score.end());
  const double thr = mx * quality;
  if (max_out) *max_out = mx;
  c.clear();
  for (int y = 0; y < im.h; y++)
    for (int x = 0; x < im.w; x++) {
      const double s = score[(size_t)y * im.w + x];
      if (s >= thr) c.push_back(Cand{x, y, s});
    }
  if (sorted) std::sort(c.begin(), c.end(), [](const Cand& a, const Cand& b) { return a.s > b.s; });
}

// Greedy min-distance selection (:288-300); the cap is tested after the push.
static std::vector<double> corners(const Img& im, int max_corners, double quality, int min_dist) {
  std::vector<Cand> c;
  candidates(im, quality, c, nullptr, true);
  std::vector<double> out;  // x,y interleaved
  const double d2 = (double)min_dist * min_dist;
  for (const Cand& k : c) {
    bool clash = false;
    for (size_t j = 0; j < out.size(); j += 2) {
      const double dx = out[j] - k.x, dy = out[j + 1] - k.y;
      if (dx * dx + dy * dy < d2) {
        clash = true;
        break;
      }
    }
    if (clash) continue;
    out.push_back((double)k.x);
    out.push_back((double)k.y);
    if ((int)(out.size() / 2) >= max_corners) break;
  }
  return out;
}

// ---- KLT (:402-460) -----------------------------------------------------------------------------------
struct LK {
  int levels = 3, radius = 5, iters = 10;
};

static void lk_update(const Img& I0, const Img& I1, int r, double x, double y, double& sx, double& sy) {
  double a00 = 0, a01 = 0, a11 = 0, b0 = 0, b1 = 0;
  for (int dy = -r; dy <= r; dy++)
    for (int dx = -r; dx <= r; dx++) {
      const double xx = x + dx, yy = y + dy;
      const double ix = 0.5 * (bilerp(I1, xx + 1, yy) - bilerp(I1, xx - 1, yy));
      const double iy = 0.5 * (bilerp(I1, xx, yy + 1) - bilerp(I1, xx, yy - 1));
      const double e = bilerp(I0, xx, yy) - bilerp(I1, xx, yy);
      a00 += ix * ix;
      a01 += ix * iy;
      a11 += iy * iy;
      b0 += ix * e;
      b1 += iy * e;
    }
  const double det = a00 * a11 - a01 * a01;
  if (std::fabs(det) < 1e-9) {
    sx = 0;
    sy = 0;
    return;
  }
  const double i00 = a11 / det, i01 = -a01 / det, i11 = a00 / det;
  sx = i00 * b0 + i01 * b1;
  sy = i01 * b0 + i11 * b1;
}

static void track(const Pyr& A, const Pyr& B, const LK& c, double px, double py, double& ox, double& oy,
                  int* iters_done = nullptr) {
  int n_it = 0;
  for (int l = c.levels - 1; l >= 0; l--) {
    const double sc = 1.0 / (1 << l);
    const double lx = px * sc, ly = py * sc;
    double dx = 0, dy = 0;
    for (int it = 0; it < c.iters; it++) {
      double sx, sy;
      lk_update(A.lvl[l], B.lvl[l], c.radius, lx + dx, ly + dy, sx, sy);
      n_it++;
      dx += sx;
      dy += sy;
      if (std::hypot(sx, sy) < 1e-3) break;
    }
    px = (lx + dx) * (1 << l);
    py = (ly + dy) * (1 << l);
  }
  ox = px;
  oy = py;
  if (iters_done) *iters_done += n_it;
}

// ---- stateful tracker (:323-391) ------------------------------------------------------------------------
struct Tracker {
  int max_tracks, min_tracks, min_dist;
  double quality, fb;
  LK lk;
  int w = 0, h = 0;
  std::vector<uint8_t> prev;
  std::vector<int> ids;
  std::vector<double> xy;
  int next_id = 0;

  void reset(const uint8_t* pix, int W, int H) {
    w = W;
    h = H;
    prev.assign(pix, pix + (size_t)W * H);
    ids.clear();
    xy = corners(Img{W, H, pix}, max_tracks, quality, min_dist);
    for (size_t i = 0; i < xy.size() / 2; i++) ids.push_back(next_id++);
  }

  int step(const uint8_t* pix, int W, int H, double* prev_xy, double* cur_xy, int* out_ids, int cap) {
    if (w == 0 || ids.empty()) {
      reset(pix, W, H);
      return 0;
    }
    Pyr P0 = make_pyr(prev.data(), w, h, lk.levels);
    Pyr P1 = make_pyr(pix, W, H, lk.levels);
    std::vector<int> kid;
    std::vector<double> kxy;
    int n = 0;
    for (size_t i = 0; i < ids.size(); i++) {
      const double x0 = xy[2 * i], y0 = xy[2 * i + 1];
      double x1, y1, xb, yb;
      track(P0, P1, lk, x0, y0, x1, y1);
      track(P1, P0, lk, x1, y1, xb, yb);
      const double d = std::hypot(xb - x0, yb - y0);
      if (d >= fb) continue;  // NaN is kept, as in the reference
      kid.push_back(ids[i]);
      kxy.push_back(x1);
      kxy.push_back(y1);
      if (n < cap) {
        prev_xy[2 * n] = x0;
        prev_xy[2 * n + 1] = y0;
        cur_xy[2 * n] = x1;
        cur_xy[2 * n + 1] = y1;
        out_ids[n] = ids[i];
      }
      n++;
    }
    w = W;
    h = H;
    prev.assign(pix, pix + (size_t)W * H);
    ids.swap(kid);
    xy.swap(kxy);
    if ((int)ids.size() < min_tracks) {
      const int need = max_tracks - (int)ids.size();
      std::vector<double> fresh = corners(Img{W, H, pix}, need * 3, quality, min_dist);
      const double d2 = (double)min_dist * min_dist;
      for (size_t k = 0; k < fresh.size(); k += 2) {
        bool clash = false;
        for (size_t j = 0; j < xy.size(); j += 2) {
          const double dx = xy[j] - fresh[k], dy = xy[j + 1] - fresh[k + 1];
          if (dx * dx + dy * dy < d2) {
            clash = true;
            break;
          }
        }
        if (clash) continue;
        ids.push_back(next_id++);
        xy.push_back(fresh[k]);
        xy.push_back(fresh[k + 1]);
        if ((int)ids.size() >= max_tracks) break;
      }
    }
    return n;
  }
};

// ---- geometry (:471-638, linalg.hpp:133-201) ---------------------------------------------------------
typedef double M3[9];

static double det3(const double* a) {
  return a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
}
static bool kinv(const double* K, double* o) {
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
static void normalise(const double* Ki, double u, double v, double& x, double& y) {
  const double a = Ki[0] * u + Ki[1] * v + Ki[2] * 1.0;
  const double b = Ki[3] * u + Ki[4] * v + Ki[5] * 1.0;
  const double c = Ki[6] * u + Ki[7] * v + Ki[8] * 1.0;
  x = a / c;
  y = b / c;
}

// Cyclic-free Jacobi: largest off-diagonal pivot, angle from atan2, rows then columns, ascending output.
static void jacobi(std::vector<double> A, int N, int sweeps, std::vector<double>& w, std::vector<double>& V) {
  V.assign((size_t)N * N, 0.0);
  for (int i = 0; i < N; i++) V[i * N + i] = 1.0;
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
      const double u = V[k * N + p], v = V[k * N + q];
      V[k * N + p] = c * u - s * v;
      V[k * N + q] = s * u + c * v;
    }
  }
  std::vector<double> d(N);
  std::vector<int> perm(N);
  for (int i = 0; i < N; i++) {
    d[i] = A[i * N + i];
    perm[i] = i;
  }
  std::sort(perm.begin(), perm.end(), [&](int i, int j) { return d[i] < d[j]; });
  w.resize(N);
  std::vector<double> V2((size_t)N * N);
  for (int c = 0; c < N; c++) {
    w[c] = d[perm[c]];
    for (int r = 0; r < N; r++) V2[r * N + c] = V[r * N + perm[c]];
  }
  V.swap(V2);
}

static void gram(const double* A, int rows, int cols, std::vector<double>& M) {
  M.assign((size_t)cols * cols, 0.0);
  for (int i = 0; i < cols; i++)
    for (int j = i; j < cols; j++) {
      double s = 0;
      for (int r = 0; r < rows; r++) s += A[r * cols + i] * A[r * cols + j];
      M[i * cols + j] = s;
      M[j * cols + i] = s;
    }
}

struct V3 {
  double x, y, z;
};
static V3 unit3(V3 v) {
  const double n = std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
  if (!std::isfinite(n) || n < 1e-12) return V3{0, 0, 0};
  return V3{v.x / n, v.y / n, v.z / n};
}
static V3 mulv(const double* A, V3 v) {
  return V3{A[0] * v.x + A[1] * v.y + A[2] * v.z, A[3] * v.x + A[4] * v.y + A[5] * v.z,
            A[6] * v.x + A[7] * v.y + A[8] * v.z};
}
static void mulm(const double* A, const double* B, double* C) {
  double T[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += A[3 * r + k] * B[3 * k + c];
      T[3 * r + c] = s;
    }
  std::memcpy(C, T, sizeof T);
}
static void transp(const double* A, double* T) {
  double t[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) t[3 * r + c] = A[3 * c + r];
  std::memcpy(T, t, sizeof t);
}

// svd3 (:537-593): V, s from eig(AtA) sorted descending; U = A V / s then Gram-Schmidt with u2 = u0 x u1.
static void svd3(const double* A, double* U, double* s, double* V) {
  double At[9];
  transp(A, At);
  std::vector<double> G(9);
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) {
      double t = 0;
      for (int k = 0; k < 3; k++) t += At[3 * r + k] * A[3 * k + c];
      G[3 * r + c] = t;
    }
  std::vector<double> w, Ve;
  jacobi(G, 3, 80, w, Ve);
  double sv[3] = {std::sqrt(std::max(0.0, w[0])), std::sqrt(std::max(0.0, w[1])), std::sqrt(std::max(0.0, w[2]))};
  int ord[3] = {0, 1, 2};
  std::sort(ord, ord + 3, [&](int i, int j) { return sv[i] > sv[j]; });
  for (int c = 0; c < 3; c++) {
    s[c] = sv[ord[c]];
    for (int r = 0; r < 3; r++) V[3 * r + c] = Ve[3 * r + ord[c]];
  }
  V3 u[3];
  for (int c = 0; c < 3; c++) {
    V3 vc{V[c], V[3 + c], V[6 + c]};
    V3 t = mulv(A, vc);
    if (s[c] > 1e-12)
      t = V3{t.x / s[c], t.y / s[c], t.z / s[c]};
    else
      t = unit3(t);
    u[c] = t;
  }
  V3 u0 = unit3(u[0]);
  const double d01 = u0.x * u[1].x + u0.y * u[1].y + u0.z * u[1].z;
  V3 u1{u[1].x - d01 * u0.x, u[1].y - d01 * u0.y, u[1].z - d01 * u0.z};
  u1 = unit3(u1);
  V3 u2{u0.y * u1.z - u0.z * u1.y, u0.z * u1.x - u0.x * u1.z, u0.x * u1.y - u0.y * u1.x};
  u2 = unit3(u2);
  U[0] = u0.x; U[3] = u0.y; U[6] = u0.z;
  U[1] = u1.x; U[4] = u1.y; U[7] = u1.z;
  U[2] = u2.x; U[5] = u2.y; U[8] = u2.z;
}

// eight_point_E (:609-627) + enforce_rank2 (:595-607).
static void eight_point(const double* xi, const double* xj, const int* idx8, double* E) {
  double A[72];
  for (int r = 0; r < 8; r++) {
    const int i = idx8[r];
    const double x = xi[2 * i], y = xi[2 * i + 1], xp = xj[2 * i], yp = xj[2 * i + 1];
    const double row[9] = {xp * x, xp * y, xp, yp * x, yp * y, yp, x, y, 1.0};
    for (int c = 0; c < 9; c++) A[r * 9 + c] = row[c];
  }
  std::vector<double> G, w, V;
  gram(A, 8, 9, G);
  jacobi(G, 9, 120, w, V);
  double E0[9];
  for (int r = 0; r < 9; r++) E0[r] = V[r * 9 + 0];
  double U[9], s[3], Vv[9];
  svd3(E0, U, s, Vv);
  double S[9] = {s[0], 0, 0, 0, s[1], 0, 0, 0, 0.0};
  double US[9], Vt[9];
  mulm(U, S, US);
  transp(Vv, Vt);
  mulm(US, Vt, E);
}

// sampson_err (:629-638)
static inline double sampson(const double* E, double x, double y, double xp, double yp) {
  const double ex = E[0] * x + E[1] * y + E[2] * 1.0;
  const double ey = E[3] * x + E[4] * y + E[5] * 1.0;
  const double ez = E[6] * x + E[7] * y + E[8] * 1.0;
  const double tx = E[0] * xp + E[3] * yp + E[6] * 1.0;
  const double ty = E[1] * xp + E[4] * yp + E[7] * 1.0;
  const double num = xp * ex + yp * ey + 1.0 * ez;
  const double den = ex * ex + ey * ey + tx * tx + ty * ty + 1e-12;
  return (num * num) / den;
}

}  // namespace orc

using namespace orc;

extern "C" {

int orc_build_pyr(const uint8_t* pix, int w, int h, int levels, uint8_t* out) {
  Pyr P = make_pyr(pix, w, h, levels);
  size_t off = 0;
  for (int l = 1; l < levels; l++) {
    std::memcpy(out + off, P.store[l - 1].data(), P.store[l - 1].size());
    off += P.store[l - 1].size();
  }
  return 0;
}

int orc_score_map(const uint8_t* pix, int w, int h, double* score) {
  std::vector<double> s;
  score_map(Img{w, h, pix}, s);
  std::memcpy(score, s.data(), s.size() * sizeof(double));
  return 0;
}

// Candidate list (x, y, score); sorted=0 raster order (:280-285), sorted=1 after std::sort (:286).
int orc_candidates(const uint8_t* pix, int w, int h, double quality, int sorted, int* xy, double* s, int cap,
                   double* max_score) {
  std::vector<Cand> c;
  candidates(Img{w, h, pix}, quality, c, max_score, sorted != 0);
  for (size_t i = 0; i < c.size() && (int)i < cap; i++) {
    xy[2 * i] = c[i].x;
    xy[2 * i + 1] = c[i].y;
    s[i] = c[i].s;
  }
  return (int)c.size();
}

// The libstdc++ std::sort permutation for keys sorted descending: perm[i] = original index at position i.
int orc_sort_perm_desc(const double* keys, int n, int* perm) {
  struct KI {
    double k;
    int i;
  };
  std::vector<KI> v(n);
  for (int i = 0; i < n; i++) v[i] = KI{keys[i], i};
  std::sort(v.begin(), v.end(), [](const KI& a, const KI& b) { return a.k > b.k; });
  for (int i = 0; i < n; i++) perm[i] = v[i].i;
  return 0;
}

int orc_shi_tomasi(const uint8_t* pix, int w, int h, int max_corners, double quality, int min_dist, double* xy_out,
                   int cap) {
  std::vector<double> c = corners(Img{w, h, pix}, max_corners, quality, min_dist);
  const int n = (int)(c.size() / 2);
  for (int i = 0; i < n && i < cap; i++) {
    xy_out[2 * i] = c[2 * i];
    xy_out[2 * i + 1] = c[2 * i + 1];
  }
  return n;
}

int orc_klt_track(const uint8_t* im0, const uint8_t* im1, int w, int h, int levels, int radius, int iters,
                  const double* p0, int n, double* p1, double* p0b) {
  Pyr A = make_pyr(im0, w, h, levels), B = make_pyr(im1, w, h, levels);
  LK c{levels, radius, iters};
  for (int i = 0; i < n; i++) {
    track(A, B, c, p0[2 * i], p0[2 * i + 1], p1[2 * i], p1[2 * i + 1]);
    track(B, A, c, p1[2 * i], p1[2 * i + 1], p0b[2 * i], p0b[2 * i + 1]);
  }
  return 0;
}

// Same, also reporting the LK iterations executed per point (fwd + bwd) for flop accounting.
int orc_klt_track_count(const uint8_t* im0, const uint8_t* im1, int w, int h, int levels, int radius, int iters,
                        const double* p0, int n, double* p1, double* p0b, int* n_it) {
  Pyr A = make_pyr(im0, w, h, levels), B = make_pyr(im1, w, h, levels);
  LK c{levels, radius, iters};
  for (int i = 0; i < n; i++) {
    n_it[i] = 0;
    track(A, B, c, p0[2 * i], p0[2 * i + 1], p1[2 * i], p1[2 * i + 1], &n_it[i]);
    track(B, A, c, p1[2 * i], p1[2 * i + 1], p0b[2 * i], p0b[2 * i + 1], &n_it[i]);
  }
  return 0;
}

void* orc_tracker_create(int max_tracks, int min_tracks, double quality, int min_distance, int levels, int radius,
                         int iters, double fb) {
  Tracker* t = new Tracker();
  t->max_tracks = max_tracks;
  t->min_tracks = min_tracks;
  t->quality = quality;
  t->min_dist = min_distance;
  t->fb = fb;
  t->lk = LK{levels, radius, iters};
  return t;
}
void orc_tracker_destroy(void* t) { delete (Tracker*)t; }
void orc_tracker_reset(void* t, const uint8_t* pix, int w, int h) { ((Tracker*)t)->reset(pix, w, h); }
int orc_tracker_step(void* t, const uint8_t* pix, int w, int h, double* prev_xy, double* cur_xy, int* ids, int cap) {
  return ((Tracker*)t)->step(pix, w, h, prev_xy, cur_xy, ids, cap);
}
int orc_tracker_tracks(void* t, double* xy, int* ids, int cap) {
  Tracker* T = (Tracker*)t;
  const int n = (int)T->ids.size();
  for (int i = 0; i < n && i < cap; i++) {
    xy[2 * i] = T->xy[2 * i];
    xy[2 * i + 1] = T->xy[2 * i + 1];
    ids[i] = T->ids[i];
  }
  return n;
}

int orc_norm_points(const double* K, const double* p, int n, double* out) {
  double Ki[9];
  if (!kinv(K, Ki)) return -1;
  for (int i = 0; i < n; i++) normalise(Ki, p[2 * i], p[2 * i + 1], out[2 * i], out[2 * i + 1]);
  return 0;
}

double orc_sampson(const double* E, double x, double y, double xp, double yp) { return sampson(E, x, y, xp, yp); }

int orc_rng_draws(int n, int count, int* out) {
  std::mt19937 rng(12345);
  std::uniform_int_distribution<int> uni(0, n - 1);
  for (int i = 0; i < count; i++) out[i] = uni(rng);
  return 0;
}

int orc_ransac_hypotheses(const double* xi, const double* xj, int n, int iters, double* E_out, int* idx_out) {
  std::mt19937 rng(12345);
  std::uniform_int_distribution<int> uni(0, n - 1);
  int idx8[8];
  for (int it = 0; it < iters; it++) {
    for (int k = 0; k < 8; k++) idx8[k] = uni(rng);
    if (idx_out) std::memcpy(idx_out + 8 * it, idx8, sizeof idx8);
    eight_point(xi, xj, idx8, E_out + 9 * it);
  }
  return 0;
}

int orc_ransac_score(const double* xi, const double* xj, int n, const double* E, int H, double thr, int* counts,
                     int* best_h, int* best_inl, int* best_n) {
  int bh = -1, bn = 0;
  for (int h = 0; h < H; h++) {
    int c = 0;
    for (int i = 0; i < n; i++)
      if (sampson(E + 9 * h, xi[2 * i], xi[2 * i + 1], xj[2 * i], xj[2 * i + 1]) < thr) c++;
    counts[h] = c;
    if (c > bn) {
      bn = c;
      bh = h;
    }
  }
  *best_h = bh;
  *best_n = bn;
  if (bh >= 0 && best_inl) {
    int k = 0;
    for (int i = 0; i < n; i++)
      if (sampson(E + 9 * bh, xi[2 * i], xi[2 * i + 1], xj[2 * i], xj[2 * i + 1]) < thr) best_inl[k++] = i;
  }
  return 0;
}

// find_E_ransac (:646-761): returns 1 on success, 0 for "nullopt".
int orc_find_E_ransac(const double* K, const double* pi, const double* pj, int n, int iters, double thr,
                      int min_inliers, double* R, double* t, int* inliers, int* n_inl) {
  *n_inl = 0;
  if (n < 8) return 0;
  double Ki[9];
  if (!kinv(K, Ki)) return -1;
  std::vector<double> xi(2 * (size_t)n), xj(2 * (size_t)n);
  for (int i = 0; i < n; i++) {
    normalise(Ki, pi[2 * i], pi[2 * i + 1], xi[2 * i], xi[2 * i + 1]);
    normalise(Ki, pj[2 * i], pj[2 * i + 1], xj[2 * i], xj[2 * i + 1]);
  }
  std::mt19937 rng(12345);
  std::uniform_int_distribution<int> uni(0, n - 1);
  double bestE[9] = {0};
  std::vector<int> best, cur;
  int idx8[8];
  for (int it = 0; it < iters; it++) {
    for (int k = 0; k < 8; k++) idx8[k] = uni(rng);
    double E[9];
    eight_point(xi.data(), xj.data(), idx8, E);
    cur.clear();
    for (int i = 0; i < n; i++)
      if (sampson(E, xi[2 * i], xi[2 * i + 1], xj[2 * i], xj[2 * i + 1]) < thr) cur.push_back(i);
    if (cur.size() > best.size()) {
      best.swap(cur);
      std::memcpy(bestE, E, sizeof E);
    }
  }
  if ((int)best.size() < min_inliers) return 0;

  // pose recovery (:680-760)
  double U[9], s[3], V[9], Vt[9];
  svd3(bestE, U, s, V);
  transp(V, Vt);
  const double W[9] = {0, -1, 0, 1, 0, 0, 0, 0, 1};
  double Wt[9], R1[9], R2[9], tmp[9];
  transp(W, Wt);
  mulm(U, W, tmp);
  mulm(tmp, Vt, R1);
  mulm(U, Wt, tmp);
  mulm(tmp, Vt, R2);
  if (det3(R1) < 0)
    for (double& v : R1) v = -v;
  if (det3(R2) < 0)
    for (double& v : R2) v = -v;
  V3 tt = unit3(V3{U[2], U[5], U[8]});

  auto cheir = [&](const double* Rm, V3 tv) {
    int ok = 0;
    const int M = std::min((int)best.size(), 20);
    for (int k = 0; k < M; k++) {
      const int i = best[k];
      const double x = xi[2 * i], y = xi[2 * i + 1], xp = xj[2 * i], yp = xj[2 * i + 1];
      double A[16] = {-1, 0, x, 0, 0, -1, y, 0,
                      xp * Rm[6] - Rm[0], xp * Rm[7] - Rm[1], xp * Rm[8] - Rm[2], xp * tv.z - tv.x,
                      yp * Rm[6] - Rm[3], yp * Rm[7] - Rm[4], yp * Rm[8] - Rm[5], yp * tv.z - tv.y};
      std::vector<double> G, w, Ve;
      gram(A, 4, 4, G);
      jacobi(G, 4, 80, w, Ve);
      const double ww = Ve[12];
      V3 X{Ve[0] / ww, Ve[4] / ww, Ve[8] / ww};
      V3 X2 = mulv(Rm, X);
      const double z2 = X2.z + tv.z;
      if (X.z > 0 && z2 > 0) ok++;
    }
    return ok;
  };
  const double* Rs[4] = {R1, R1, R2, R2};
  const V3 ts[4] = {tt, V3{-tt.x, -tt.y, -tt.z}, tt, V3{-tt.x, -tt.y, -tt.z}};
  int bi = 0, bok = -1;
  for (int i = 0; i < 4; i++) {
    const int ok = cheir(Rs[i], ts[i]);
    if (ok > bok) {
      bok = ok;
      bi = i;
    }
  }
  std::memcpy(R, Rs[bi], 9 * sizeof(double));
  t[0] = ts[bi].x;
  t[1] = ts[bi].y;
  t[2] = ts[bi].z;
  *n_inl = (int)best.size();
  for (size_t i = 0; i < best.size(); i++) inliers[i] = best[i];
  return 1;
}

int orc_pair_frontend(const uint8_t* im0, const uint8_t* im1, int w, int h, int max_corners, double quality,
                      int min_dist, int levels, int radius, int iters, double fb_thresh, double* li, double* lj,
                      int* n_corners) {
  std::vector<double> c = corners(Img{w, h, im0}, max_corners, quality, min_dist);
  Pyr A = make_pyr(im0, w, h, levels), B = make_pyr(im1, w, h, levels);
  LK lk{levels, radius, iters};
  int k = 0;
  for (size_t i = 0; i < c.size(); i += 2) {
    double x1, y1, xb, yb;
    track(A, B, lk, c[i], c[i + 1], x1, y1);
    track(B, A, lk, x1, y1, xb, yb);
    if (std::hypot(xb - c[i], yb - c[i + 1]) >= fb_thresh) continue;
    li[2 * k] = c[i];
    li[2 * k + 1] = c[i + 1];
    lj[2 * k] = x1;
    lj[2 * k + 1] = y1;
    k++;
  }
  if (n_corners) *n_corners = (int)(c.size() / 2);
  return k;
}

long orc_pair_frontend_mt(const uint8_t* frames, int nframes, int w, int h, int max_corners, double quality,
                          int min_dist, int levels, int radius, int iters, double fb_thresh, int threads,
                          long* kept_total) {
  const int npairs = nframes - 1;
  std::vector<long> tracks(threads, 0), kept(threads, 0);
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++)
    pool.emplace_back([&, t]() {
      std::vector<double> li((size_t)2 * max_corners), lj((size_t)2 * max_corners);
      for (int p = t; p < npairs; p += threads) {
        int nc = 0;
        const int k = orc_pair_frontend(frames + (size_t)p * w * h, frames + (size_t)(p + 1) * w * h, w, h,
                                        max_corners, quality, min_dist, levels, radius, iters, fb_thresh, li.data(),
                                        lj.data(), &nc);
        tracks[t] += nc;
        kept[t] += k;
      }
    });
  for (auto& th : pool) th.join();
  long tt = 0, kk = 0;
  for (int t = 0; t < threads; t++) {
    tt += tracks[t];
    kk += kept[t];
  }
  if (kept_total) *kept_total = kk;
  return tt;
}


// The whole two-view unit (:1836-1857) for `npairs` independent pairs (frames[i], frames[i+1]) on `threads` host threads:
// detect + fwd/bwd track + fb filter, then `if (li.size() >= min_points) find_E_ransac(K, li, lj, rs_iters, rs_thr,
// rs_min_inliers)`.  Per pair: n_corners, n_kept, li / lj [max_corners][2], status (0 skipped, 1 nullopt, 2 pose),
// n_inl, inliers [max_corners], R [9], t [3].  Any per-pair output may be null.  Returns the corners processed.
long orc_two_view_mt(const uint8_t* frames, int nframes, int w, int h, int max_corners, double quality, int min_dist,
                     int levels, int radius, int iters, double fb_thresh, const double* K, int rs_iters, double rs_thr,
                     int rs_min_inliers, int min_points, int threads, int* n_corners, int* n_kept, double* li_out,
                     double* lj_out, int* status, int* n_inl, int* inliers, double* R_out, double* t_out) {
  const int npairs = nframes - 1;
  const int cap = max_corners < 1 ? 1 : max_corners;
  std::vector<long> tracks(threads, 0);
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++) {
    pool.emplace_back([&, t]() {
      std::vector<double> li((size_t)2 * cap), lj((size_t)2 * cap);
      std::vector<int> inl((size_t)cap);
      for (int p = t; p < npairs; p += threads) {
        int nc = 0;
        const int k = orc_pair_frontend(frames + (size_t)p * w * h, frames + (size_t)(p + 1) * w * h, w, h, max_corners,
                                        quality, min_dist, levels, radius, iters, fb_thresh, li.data(), lj.data(), &nc);
        tracks[t] += nc;
        int st = 0, ni = 0;
        double R[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, tt[3] = {0, 0, 0};
        if (K && k >= min_points) st = orc_find_E_ransac(K, li.data(), lj.data(), k, rs_iters, rs_thr, rs_min_inliers, R, tt, inl.data(), &ni) ? 2 : 1;
        if (n_corners) n_corners[p] = nc;
        if (n_kept) n_kept[p] = k;
        if (li_out) std::copy(li.begin(), li.begin() + 2 * (size_t)k, li_out + (size_t)p * 2 * cap);
        if (lj_out) std::copy(lj.begin(), lj.begin() + 2 * (size_t)k, lj_out + (size_t)p * 2 * cap);
        if (status) status[p] = st;
        if (n_inl) n_inl[p] = st == 2 ? ni : 0;
        if (inliers && st == 2) std::copy(inl.begin(), inl.begin() + ni, inliers + (size_t)p * cap);
        if (R_out) std::copy(R, R + 9, R_out + (size_t)p * 9);
        if (t_out) std::copy(tt, tt + 3, t_out + (size_t)p * 3);
      }
    });
  }
  for (auto& th : pool) th.join();
  long total = 0;
  for (int t = 0; t < threads; t++) total += tracks[t];
  return total;
}

int orc_ransac_score_mt(const double* xi, const double* xj, int n, const double* E, int H, double thr, int* counts,
                        int threads) {
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++)
    pool.emplace_back([=]() {
      for (int h = t; h < H; h += threads) {
        int c = 0;
        for (int i = 0; i < n; i++)
          if (sampson(E + 9 * h, xi[2 * i], xi[2 * i + 1], xj[2 * i], xj[2 * i + 1]) < thr) c++;
        counts[h] = c;
      }
    });
  for (auto& th : pool) th.join();
  return 0;
}

// Integer synthetic generator, twin of sfmgpu/synth.py (inputs for the CPU arms without touching GPU code).
static inline uint32_t h32(uint32_t a) {
  a ^= a >> 16; a *= 0x7FEB352Du; a ^= a >> 15; a *= 0x846CA68Bu; a ^= a >> 16;
  return a;
}
static inline uint32_t lat(uint32_t s, uint32_t ix, uint32_t iy) { return h32(ix + h32(iy + s)) >> 24; }
static inline uint32_t vnoise(uint32_t s, uint32_t X, uint32_t Y, int lc) {
  const uint32_t ix = X >> (lc + 8), iy = Y >> (lc + 8), fx = (X >> lc) & 255u, fy = (Y >> lc) & 255u;
  const uint32_t top = lat(s, ix, iy) * (256u - fx) + lat(s, ix + 1, iy) * fx;
  const uint32_t bot = lat(s, ix, iy + 1) * (256u - fx) + lat(s, ix + 1, iy + 1) * fx;
  return (top * (256u - fy) + bot * fy) >> 16;
}
static inline uint32_t blk(uint32_t s, uint32_t i, uint32_t j) { return (lat(s, i, j) & 3u) == 0u ? 255u : 0u; }

int orc_synth_frames(uint32_t seed, int t0, int nframes, int w, int h, uint8_t* out, int threads) {
  std::vector<std::thread> pool;
  for (int th = 0; th < threads; th++)
    pool.emplace_back([=]() {
      const uint32_t s0 = h32(seed * 4u), s1 = h32(seed * 4u + 1u), s2 = h32(seed * 4u + 2u);
      for (int k = th; k < nframes; k += threads) {
        const int t = t0 + k, m = ((t % 128) + 128) % 128, tri = m < 64 ? m : 128 - m;
        uint8_t* dst = out + (size_t)k * w * h;
        for (int y = 0; y < h; y++) {
          const uint32_t Y = (uint32_t)y * 256u + (uint32_t)((1 << 20) - 8 * tri);
          for (int x = 0; x < w; x++) {
            const uint32_t X = (uint32_t)x * 256u + (uint32_t)((1 << 20) + 16 * tri);
            const uint32_t ix = X >> 13, iy = Y >> 13, fx = X & 8191u, fy = Y & 8191u;
            const uint32_t w1x = fx > 8192u - 256u ? fx - (8192u - 256u) : 0u, w0x = 256u - w1x;
            const uint32_t w1y = fy > 8192u - 256u ? fy - (8192u - 256u) : 0u, w0y = 256u - w1y;
            const uint32_t B = (blk(s0, ix, iy) * w0x * w0y + blk(s0, ix + 1, iy) * w1x * w0y +
                                blk(s0, ix, iy + 1) * w0x * w1y + blk(s0, ix + 1, iy + 1) * w1x * w1y) >> 16;
            dst[(size_t)y * w + x] = (uint8_t)((4u * B + 3u * vnoise(s1, X, Y, 4) + vnoise(s2, X, Y, 2)) >> 3);
          }
        }
      }
    });
  for (auto& t : pool) t.join();
  return 0;
}

// ---- batched triangulate_dlt (:1477-1516): poses = 9 doubles R (camera->world) + 3 doubles centre each ------------------
int orc_triangulate_dlt(const double* K, const double* poses, int P, const int* ia, const int* ib, const double* ui, const double* uj,
                        int n, double* X) {
  (void)P;
  double Ki[9];
  if (!kinv(K, Ki)) return -1;
  for (int k = 0; k < n; k++) {
    double A[16];
    for (int cam = 0; cam < 2; cam++) {
      const double* pose = poses + 12 * (cam == 0 ? ia[k] : ib[k]);
      const double* u = cam == 0 ? ui + 2 * k : uj + 2 * k;
      double x, y;
      normalise(Ki, u[0], u[1], x, y);
      double Rw[9];
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) Rw[3 * r + c] = pose[3 * c + r];
      const double tx = -(Rw[0] * pose[9] + Rw[1] * pose[10] + Rw[2] * pose[11]);
      const double ty = -(Rw[3] * pose[9] + Rw[4] * pose[10] + Rw[5] * pose[11]);
      const double tz = -(Rw[6] * pose[9] + Rw[7] * pose[10] + Rw[8] * pose[11]);
      double* o = A + 8 * cam;
      o[0] = x * Rw[6] - Rw[0]; o[1] = x * Rw[7] - Rw[1]; o[2] = x * Rw[8] - Rw[2]; o[3] = x * tz - tx;
      o[4] = y * Rw[6] - Rw[3]; o[5] = y * Rw[7] - Rw[4]; o[6] = y * Rw[8] - Rw[5]; o[7] = y * tz - ty;
    }
    std::vector<double> G, w, V;
    gram(A, 4, 4, G);
    jacobi(G, 4, 80, w, V);
    const double ww = V[12];
    X[3 * k] = V[0] / ww;
    X[3 * k + 1] = V[4] / ww;
    X[3 * k + 2] = V[8] / ww;
  }
  return 0;
}

// ---- loop-closure descriptor (:1100-1122) and candidate search (:1124-1129, :1823-1831) ----------------------------
// Halve until both sides are <= 32, nearest-sample to 32x32 (std::round), float values, mean in double, subtract
// (float)mean in float, squared norm accumulated in double in raster order, scale by 1/sqrt(n2 + 1e-12) in double and
// round to float.
int orc_global_desc32(const uint8_t* pix, int w, int h, float* out) {
  std::vector<uint8_t> a(pix, pix + (size_t)w * h), b;
  int cw = w, ch = h;
  while (cw > 32 || ch > 32) {
    b.assign((size_t)(cw / 2) * (ch / 2), 0);
    halve(a.data(), cw, ch, b.data());
    a.swap(b);
    cw /= 2;
    ch /= 2;
  }
  if (cw < 1 || ch < 1) return -1;
  double mean = 0.0;
  for (int y = 0; y < 32; y++)
    for (int x = 0; x < 32; x++) {
      const int sx = std::min(cw - 1, (int)std::round((double)x * (cw - 1) / 31.0));
      const int sy = std::min(ch - 1, (int)std::round((double)y * (ch - 1) / 31.0));
      const float val = (float)a[(size_t)sy * cw + sx];
      out[y * 32 + x] = val;
      mean += val;
    }
  mean /= (32.0 * 32.0);
  double n2 = 0.0;
  for (int i = 0; i < 1024; i++) {
    out[i] = (float)(out[i] - (float)mean);
    n2 += (double)out[i] * (double)out[i];
  }
  const double invn = 1.0 / std::sqrt(n2 + 1e-12);
  for (int i = 0; i < 1024; i++) out[i] = (float)(out[i] * invn);
  return 1024;
}

int orc_desc_search(const float* descs, int n_search, const float* query, float* scores, int* best_id, float* best_score) {
  int bid = -1;
  float bs = 0.0f;
  for (int kk = 0; kk < n_search; ++kk) {
    const float* a = descs + (size_t)kk * 1024;
    float s = 0.0f;
    for (int i = 0; i < 1024; i++) s += a[i] * query[i];
    if (scores) scores[kk] = s;
    if (s > bs) {
      bs = s;
      bid = kk;
    }
  }
  *best_id = bid;
  *best_score = bs;
  return 0;
}

const char* orc_kind() { return "port"; }
}
