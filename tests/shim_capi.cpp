// shim_capi.cpp — flat C wrappers around host/sfmgpu_shim.hpp so that the tests can drive the C++ drop-in
// (same entry points the reference's main() calls) through ctypes.  Test plumbing, built into libsfmshim.so.
#define SFMGPU_SHIM_STANDALONE
#include "sfmgpu_shim.hpp"

#include <cstring>

namespace {
GrayImage wrap(const uint8_t* pix, int w, int h) {
  GrayImage g;
  g.w = w;
  g.h = h;
  g.pix.assign(pix, pix + (size_t)w * h);
  return g;
}
thread_local std::string g_err;
template <class F>
int guarded(F&& f) {
  try {
    return f();
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1000;
  }
}
}  // namespace

extern "C" {

const char* shim_last_error() { return g_err.c_str(); }

int shim_build_pyr(const uint8_t* pix, int w, int h, int levels, uint8_t* out) {
  return guarded([&] {
    Pyramid p = build_pyr(wrap(pix, w, h), levels);
    size_t off = 0;
    for (int l = 1; l < levels; l++) {
      std::memcpy(out + off, p.lvl[l].pix.data(), p.lvl[l].pix.size());
      off += p.lvl[l].pix.size();
    }
    return 0;
  });
}

int shim_shi_tomasi(const uint8_t* pix, int w, int h, int max_corners, double quality, int min_dist, double* xy_out, int cap) {
  return guarded([&] {
    auto pts = shi_tomasi(wrap(pix, w, h), max_corners, quality, min_dist);
    const int n = (int)pts.size();
    for (int i = 0; i < n && i < cap; i++) {
      xy_out[2 * i] = pts[i].x;
      xy_out[2 * i + 1] = pts[i].y;
    }
    return n;
  });
}

// The loop-closure block :1836-1857 written against the shim exactly as the reference writes it (per-point
// track_one_public calls) when batched == 0, or with track_pairs() when batched == 1.
int shim_pair_frontend(const uint8_t* im0, const uint8_t* im1, int w, int h, int max_corners, double quality, int min_dist,
                       int levels, int radius, int iters, double fb_thresh, int batched, double* li, double* lj, int* n_corners) {
  return guarded([&] {
    LKConfig lc;
    lc.max_tracks = max_corners;
    lc.min_tracks = 0;
    lc.quality = quality;
    lc.min_distance = min_dist;
    lc.pyr_levels = levels;
    lc.win_radius = radius;
    lc.iters = iters;
    lc.fb_thresh = fb_thresh;
    KLTTracker tmp(lc);
    GrayImage g0 = wrap(im0, w, h), g1 = wrap(im1, w, h);
    const auto pts0 = shi_tomasi(g0, lc.max_tracks, lc.quality, lc.min_distance);
    Pyramid pyr0 = build_pyr(g0, lc.pyr_levels);
    Pyramid pyr1 = build_pyr(g1, lc.pyr_levels);
    std::vector<Vec2> p1v, pbv;
    if (batched) track_pairs(pyr0, pyr1, pts0, lc.win_radius, lc.iters, p1v, &pbv);
    int k = 0;
    for (size_t q = 0; q < pts0.size(); q++) {
      const Vec2 p0 = pts0[q];
      const Vec2 p1 = batched ? p1v[q] : tmp.track_one_public(pyr0, pyr1, p0);
      const Vec2 p0b = batched ? pbv[q] : tmp.track_one_public(pyr1, pyr0, p1);
      const double fb = std::hypot(p0b.x - p0.x, p0b.y - p0.y);
      if (fb >= lc.fb_thresh) continue;
      li[2 * k] = p0.x;
      li[2 * k + 1] = p0.y;
      lj[2 * k] = p1.x;
      lj[2 * k + 1] = p1.y;
      k++;
    }
    if (n_corners) *n_corners = (int)pts0.size();
    return k;
  });
}

void* shim_tracker_create(int max_tracks, int min_tracks, double quality, int min_distance, int levels, int radius, int iters,
                          double fb) {
  try {
    LKConfig c;
    c.max_tracks = max_tracks;
    c.min_tracks = min_tracks;
    c.quality = quality;
    c.min_distance = min_distance;
    c.pyr_levels = levels;
    c.win_radius = radius;
    c.iters = iters;
    c.fb_thresh = fb;
    return new KLTTracker(c);
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
void shim_tracker_destroy(void* t) { delete (KLTTracker*)t; }

// MultiKLTTracker: pix = [n_sequences][h][w]; outputs [n_sequences][cap][2] / [n_sequences][cap] / [n_sequences]
void* shim_multitracker_create(int n_sequences, int w, int h, int max_tracks, int min_tracks) {
  try {
    LKConfig c;
    c.max_tracks = max_tracks;
    c.min_tracks = min_tracks;
    return new MultiKLTTracker(c, n_sequences, w, h);
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
void shim_multitracker_destroy(void* t) { delete (MultiKLTTracker*)t; }
int shim_multitracker_step(void* t, const uint8_t* pix, int n_sequences, int w, int h, double* prev_xy, double* cur_xy, int* ids,
                           int* n_out, int cap, double* tracks_xy, int* tracks_ids, int* tracks_n) {
  return guarded([&] {
    std::vector<GrayImage> imgs;
    for (int s = 0; s < n_sequences; s++) imgs.push_back(wrap(pix + (size_t)s * w * h, w, h));
    std::vector<const GrayImage*> ptrs;
    for (auto& im : imgs) ptrs.push_back(&im);
    auto* mt = (MultiKLTTracker*)t;
    auto out = mt->step(ptrs);
    for (int s = 0; s < n_sequences; s++) {
      const int n = (int)out[s].ids.size();
      n_out[s] = n;
      for (int i = 0; i < n && i < cap; i++) {
        const size_t o = (size_t)s * cap + i;
        prev_xy[2 * o] = out[s].prev_pts[i].x;
        prev_xy[2 * o + 1] = out[s].prev_pts[i].y;
        cur_xy[2 * o] = out[s].cur_pts[i].x;
        cur_xy[2 * o + 1] = out[s].cur_pts[i].y;
        ids[o] = out[s].ids[i];
      }
      const auto tr = mt->tracks(s);
      tracks_n[s] = (int)tr.size();
      for (int i = 0; i < (int)tr.size() && i < cap; i++) {
        const size_t o = (size_t)s * cap + i;
        tracks_xy[2 * o] = tr[i].p.x;
        tracks_xy[2 * o + 1] = tr[i].p.y;
        tracks_ids[o] = tr[i].id;
      }
    }
    return 0;
  });
}
int shim_tracker_step(void* t, const uint8_t* pix, int w, int h, double* prev_xy, double* cur_xy, int* ids, int cap) {
  return guarded([&] {
    auto out = ((KLTTracker*)t)->step(wrap(pix, w, h));
    const int n = (int)out.ids.size();
    for (int i = 0; i < n && i < cap; i++) {
      prev_xy[2 * i] = out.prev_pts[i].x;
      prev_xy[2 * i + 1] = out.prev_pts[i].y;
      cur_xy[2 * i] = out.cur_pts[i].x;
      cur_xy[2 * i + 1] = out.cur_pts[i].y;
      ids[i] = out.ids[i];
    }
    return n;
  });
}
int shim_tracker_tracks(void* t, double* xy, int* ids, int cap) {
  const auto& tr = ((KLTTracker*)t)->tracks();
  const int n = (int)tr.size();
  for (int i = 0; i < n && i < cap; i++) {
    xy[2 * i] = tr[i].p.x;
    xy[2 * i + 1] = tr[i].p.y;
    ids[i] = tr[i].id;
  }
  return n;
}

// 1 = RelPose, 0 = std::nullopt, <0 = exception
int shim_find_E_ransac(const double* K, const double* pi, const double* pj, int n, int iters, double thr, int min_inliers,
                       double* R, double* t, int* inliers, int* n_inl) {
  return guarded([&] {
    Mat33 Km;
    for (int i = 0; i < 9; i++) Km.a[i] = K[i];
    std::vector<Vec2> a((size_t)n), b((size_t)n);
    for (int i = 0; i < n; i++) {
      a[i] = Vec2{pi[2 * i], pi[2 * i + 1]};
      b[i] = Vec2{pj[2 * i], pj[2 * i + 1]};
    }
    auto r = find_E_ransac(Km, a, b, iters, thr, min_inliers);
    *n_inl = 0;
    if (!r) return 0;
    for (int i = 0; i < 9; i++) R[i] = r->R_ji.a[i];
    t[0] = r->t_ji.x;
    t[1] = r->t_ji.y;
    t[2] = r->t_ji.z;
    *n_inl = (int)r->inliers.size();
    for (int i = 0; i < *n_inl; i++) inliers[i] = r->inliers[i];
    return 1;
  });
}

// host single-track triangulation (bit-identical to the reference), n tracks in a loop
int shim_host_triangulate(const double* K, const double* poses, const int* ia, const int* ib, const double* ui, const double* uj, int n,
                          double* X) {
  for (int k = 0; k < n; k++) {
    const double *pi = poses + 12 * ia[k], *pj = poses + 12 * ib[k];
    sfmgpu_host::triangulate_dlt(K, pi, pi + 9, pj, pj + 9, ui + 2 * k, uj + 2 * k, X + 3 * k);
  }
  return 0;
}

// opt-in device solver for find_E_ransac
void shim_set_device_solver(int on) { sfmgpu_shim::set_device_solver(on != 0); }
void shim_set_speculate(int on) { sfmgpu_shim::set_speculate(on != 0); }

// ---- host-only pieces (no GPU needed): checked on the CPU against the compiled reference ----------------------
int shim_host_norm_points(const double* K, const double* p, int n, double* out) {
  double Ki[9];
  if (!sfmgpu_host::invert_K(K, Ki)) return -1;
  for (int i = 0; i < n; i++) sfmgpu_host::norm_point(Ki, p[2 * i], p[2 * i + 1], out + 2 * i);
  return 0;
}
int shim_host_hypotheses(const double* xi, const double* xj, int n, int iters, double* E_out) {
  std::mt19937 rng(12345);
  std::uniform_int_distribution<int> uni(0, n - 1);
  int idx8[8];
  for (int it = 0; it < iters; it++) {
    for (int k = 0; k < 8; k++) idx8[k] = uni(rng);
    sfmgpu_host::eight_point_E(xi, xj, idx8, E_out + 9 * it);
  }
  return 0;
}
int shim_host_recover_pose(const double* E, const double* xi, const double* xj, const int* inliers, int n_inl, double* R,
                           double* t) {
  sfmgpu_host::recover_pose(E, xi, xj, inliers, n_inl, R, t);
  return 0;
}
}
