// oracle/ref_harness.cpp — TEST INFRASTRUCTURE, not product code.
//
// Compiles the UNMODIFIED reference translation unit where it lies
// (${REF_DIR}/cpp/src/templering_sfm.cpp, REF_DIR defaults to /root/reference) and exposes
// its file-static hot-path functions through a flat C ABI (prefix ref_).  No reference
// source is copied: the TU is pulled in by #include with `main` renamed.  The resulting
// shared object is written to oracle/_ref/ (git-ignored, travels to the GPU box).
//
// Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference arm may load it.
//
// Symbols reached (reference file:line, relative to cpp/src/templering_sfm.cpp):
//   build_pyr :224-232, shi_tomasi :237-302, KLTTracker :323-466 (reset/step/track_one_public),
//   norm_point/invert_K :471-501, eight_point_E :609-627, sampson_err :629-638,
//   find_E_ransac :646-761, stateless two-view front end :1836-1857 (lifted, see ref_pair_frontend),
//   global_desc_32 :1100-1122, dot_desc :1124-1129, loop-candidate search :1823-1831 (lifted, see ref_desc_search),
//   triangulate_dlt :1477-1516.
#define main ref_main_unused
#include "cpp/src/templering_sfm.cpp"
#undef main

#include <cstring>
#include <thread>

namespace {
GrayImage wrap(const uint8_t* pix, int w, int h) {
  GrayImage g;
  g.w = w;
  g.h = h;
  g.pix.assign(pix, pix + (size_t)w * h);
  return g;
}
LKConfig mkcfg(int max_tracks, int min_tracks, double quality, int min_distance, int levels, int radius,
               int iters, double fb) {
  LKConfig c;
  c.max_tracks = max_tracks;
  c.min_tracks = min_tracks;
  c.quality = quality;
  c.min_distance = min_distance;
  c.pyr_levels = levels;
  c.win_radius = radius;
  c.iters = iters;
  c.fb_thresh = fb;
  return c;
}
Mat33 mat(const double* k) {
  Mat33 m;
  for (int i = 0; i < 9; i++) m.a[i] = k[i];
  return m;
}
}  // namespace

extern "C" {

// Pyramid levels 1..levels-1 concatenated into out (level 0 is the input itself).
int ref_build_pyr(const uint8_t* pix, int w, int h, int levels, uint8_t* out) {
  Pyramid p = build_pyr(wrap(pix, w, h), levels);
  size_t off = 0;
  for (int l = 1; l < levels; l++) {
    std::memcpy(out + off, p.lvl[l].pix.data(), p.lvl[l].pix.size());
    off += p.lvl[l].pix.size();
  }
  return 0;
}

// Returns the number of corners; xy_out holds up to cap (x,y) pairs.
int ref_shi_tomasi(const uint8_t* pix, int w, int h, int max_corners, double quality, int min_dist,
                   double* xy_out, int cap) {
  auto pts = shi_tomasi(wrap(pix, w, h), max_corners, quality, min_dist);
  int n = (int)pts.size();
  for (int i = 0; i < n && i < cap; i++) {
    xy_out[2 * i] = pts[i].x;
    xy_out[2 * i + 1] = pts[i].y;
  }
  return n;
}

// Forward + backward track_one for n points (KLTTracker::track_one_public).
int ref_klt_track(const uint8_t* im0, const uint8_t* im1, int w, int h, int levels, int radius, int iters,
                  const double* p0, int n, double* p1, double* p0b) {
  KLTTracker trk(mkcfg(0, 0, 0.01, 8, levels, radius, iters, 1.0));
  Pyramid a = build_pyr(wrap(im0, w, h), levels);
  Pyramid b = build_pyr(wrap(im1, w, h), levels);
  for (int i = 0; i < n; i++) {
    Vec2 q = trk.track_one_public(a, b, Vec2{p0[2 * i], p0[2 * i + 1]});
    Vec2 r = trk.track_one_public(b, a, q);
    p1[2 * i] = q.x;
    p1[2 * i + 1] = q.y;
    p0b[2 * i] = r.x;
    p0b[2 * i + 1] = r.y;
  }
  return 0;
}

// Stateful tracker.
void* ref_tracker_create(int max_tracks, int min_tracks, double quality, int min_distance, int levels,
                         int radius, int iters, double fb) {
  return new KLTTracker(mkcfg(max_tracks, min_tracks, quality, min_distance, levels, radius, iters, fb));
}
void ref_tracker_destroy(void* t) { delete (KLTTracker*)t; }
void ref_tracker_reset(void* t, const uint8_t* pix, int w, int h) { ((KLTTracker*)t)->reset(wrap(pix, w, h)); }
// Returns number of survivors; arrays sized cap.
int ref_tracker_step(void* t, const uint8_t* pix, int w, int h, double* prev_xy, double* cur_xy, int* ids,
                     int cap) {
  auto out = ((KLTTracker*)t)->step(wrap(pix, w, h));
  int n = (int)out.ids.size();
  for (int i = 0; i < n && i < cap; i++) {
    prev_xy[2 * i] = out.prev_pts[i].x;
    prev_xy[2 * i + 1] = out.prev_pts[i].y;
    cur_xy[2 * i] = out.cur_pts[i].x;
    cur_xy[2 * i + 1] = out.cur_pts[i].y;
    ids[i] = out.ids[i];
  }
  return n;
}
int ref_tracker_tracks(void* t, double* xy, int* ids, int cap) {
  const auto& tr = ((KLTTracker*)t)->tracks();
  int n = (int)tr.size();
  for (int i = 0; i < n && i < cap; i++) {
    xy[2 * i] = tr[i].p.x;
    xy[2 * i + 1] = tr[i].p.y;
    ids[i] = tr[i].id;
  }
  return n;
}

// K^-1 normalisation used at :649-655.
int ref_norm_points(const double* K, const double* p, int n, double* out) {
  const Mat33 Kinv = invert_K(mat(K));
  for (int i = 0; i < n; i++) {
    Vec2 q = norm_point(Kinv, Vec2{p[2 * i], p[2 * i + 1]});
    out[2 * i] = q.x;
    out[2 * i + 1] = q.y;
  }
  return 0;
}

double ref_sampson(const double* E, double x, double y, double xp, double yp) {
  return sampson_err(mat(E), Vec2{x, y}, Vec2{xp, yp});
}

// First `count` draws of the RANSAC index stream (:657-665) for n correspondences.
int ref_rng_draws(int n, int count, int* out) {
  std::mt19937 rng(12345);
  std::uniform_int_distribution<int> uni(0, n - 1);
  for (int i = 0; i < count; i++) out[i] = uni(rng);
  return 0;
}

// The reference's own seeded hypotheses: sampling + eight_point_E on NORMALISED points (:657-666).
int ref_ransac_hypotheses(const double* xi, const double* xj, int n, int iters, double* E_out, int* idx_out) {
  std::vector<Vec2> a(n), b(n);
  for (int i = 0; i < n; i++) {
    a[i] = Vec2{xi[2 * i], xi[2 * i + 1]};
    b[i] = Vec2{xj[2 * i], xj[2 * i + 1]};
  }
  std::mt19937 rng(12345);
  std::uniform_int_distribution<int> uni(0, n - 1);
  std::vector<int> idx8(8);
  for (int it = 0; it < iters; it++) {
    for (int k = 0; k < 8; k++) idx8[k] = uni(rng);
    if (idx_out)
      for (int k = 0; k < 8; k++) idx_out[8 * it + k] = idx8[k];
    const Mat33 E = eight_point_E(a, b, idx8);
    for (int k = 0; k < 9; k++) E_out[9 * it + k] = E.a[k];
  }
  return 0;
}

// Scoring loop :667-676 for given hypotheses on normalised points (uses the reference sampson_err).
// counts[H]; best_h = first hypothesis with the strictly largest count (-1 if every count is 0).
int ref_ransac_score(const double* xi, const double* xj, int n, const double* E, int H, double thr, int* counts,
                     int* best_h, int* best_inl, int* best_n) {
  int bh = -1, bn = 0;
  for (int h = 0; h < H; h++) {
    const Mat33 Eh = mat(E + 9 * h);
    int c = 0;
    for (int i = 0; i < n; i++) {
      const double e = sampson_err(Eh, Vec2{xi[2 * i], xi[2 * i + 1]}, Vec2{xj[2 * i], xj[2 * i + 1]});
      if (e < thr) c++;
    }
    counts[h] = c;
    if (c > bn) {
      bn = c;
      bh = h;
    }
  }
  *best_h = bh;
  *best_n = bn;
  if (bh >= 0 && best_inl) {
    const Mat33 Eh = mat(E + 9 * bh);
    int k = 0;
    for (int i = 0; i < n; i++) {
      const double e = sampson_err(Eh, Vec2{xi[2 * i], xi[2 * i + 1]}, Vec2{xj[2 * i], xj[2 * i + 1]});
      if (e < thr) best_inl[k++] = i;
    }
  }
  return 0;
}

// find_E_ransac :646-761 on pixel coordinates.  Returns 1 and fills R (9), t (3), inliers on success,
// 0 for std::nullopt.
int ref_find_E_ransac(const double* K, const double* pi, const double* pj, int n, int iters, double thr,
                      int min_inliers, double* R, double* t, int* inliers, int* n_inl) {
  std::vector<Vec2> a(n), b(n);
  for (int i = 0; i < n; i++) {
    a[i] = Vec2{pi[2 * i], pi[2 * i + 1]};
    b[i] = Vec2{pj[2 * i], pj[2 * i + 1]};
  }
  auto r = find_E_ransac(mat(K), a, b, iters, thr, min_inliers);
  if (!r) {
    *n_inl = 0;
    return 0;
  }
  for (int i = 0; i < 9; i++) R[i] = r->R_ji.a[i];
  t[0] = r->t_ji.x;
  t[1] = r->t_ji.y;
  t[2] = r->t_ji.z;
  *n_inl = (int)r->inliers.size();
  for (int i = 0; i < *n_inl; i++) inliers[i] = r->inliers[i];
  return 1;
}

// Stateless two-view front end lifted from :1836-1857 (detect on im0, fwd/bwd track, fb filter).
// Returns the number of surviving correspondences; li/lj sized max_corners.
int ref_pair_frontend(const uint8_t* im0, const uint8_t* im1, int w, int h, int max_corners, double quality,
                      int min_dist, int levels, int radius, int iters, double fb_thresh, double* li, double* lj,
                      int* n_corners) {
  KLTTracker tmp(mkcfg(max_corners, 0, quality, min_dist, levels, radius, iters, fb_thresh));
  GrayImage g0 = wrap(im0, w, h), g1 = wrap(im1, w, h);
  const auto pts0 = shi_tomasi(g0, max_corners, quality, min_dist);
  Pyramid pyr0 = build_pyr(g0, levels);
  Pyramid pyr1 = build_pyr(g1, levels);
  int k = 0;
  for (const auto& p0 : pts0) {
    const Vec2 p1 = tmp.track_one_public(pyr0, pyr1, p0);
    const Vec2 p0b = tmp.track_one_public(pyr1, pyr0, p1);
    const double fb = std::hypot(p0b.x - p0.x, p0b.y - p0.y);
    if (fb >= fb_thresh) continue;
    li[2 * k] = p0.x;
    li[2 * k + 1] = p0.y;
    lj[2 * k] = p1.x;
    lj[2 * k + 1] = p1.y;
    k++;
  }
  if (n_corners) *n_corners = (int)pts0.size();
  return k;
}

// Multi-threaded CPU baseline: `npairs` independent pairs (frames[i], frames[i+1]) spread over `threads`
// host threads.  The reference itself is single-threaded; pairs are independent (:1836-1857), so this is
// the "all host cores" arm of BASELINE.md §3.  Returns total input tracks (corners) processed.
long ref_pair_frontend_mt(const uint8_t* frames, int nframes, int w, int h, int max_corners, double quality,
                          int min_dist, int levels, int radius, int iters, double fb_thresh, int threads,
                          long* kept_total) {
  const int npairs = nframes - 1;
  std::vector<long> tracks(threads, 0), kept(threads, 0);
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++) {
    pool.emplace_back([&, t]() {
      std::vector<double> li((size_t)2 * max_corners), lj((size_t)2 * max_corners);
      for (int p = t; p < npairs; p += threads) {
        int nc = 0;
        int k = ref_pair_frontend(frames + (size_t)p * w * h, frames + (size_t)(p + 1) * w * h, w, h, max_corners,
                                  quality, min_dist, levels, radius, iters, fb_thresh, li.data(), lj.data(), &nc);
        tracks[t] += nc;
        kept[t] += k;
      }
    });
  }
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
long ref_two_view_mt(const uint8_t* frames, int nframes, int w, int h, int max_corners, double quality, int min_dist,
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
        const int k = ref_pair_frontend(frames + (size_t)p * w * h, frames + (size_t)(p + 1) * w * h, w, h, max_corners,
                                        quality, min_dist, levels, radius, iters, fb_thresh, li.data(), lj.data(), &nc);
        tracks[t] += nc;
        int st = 0, ni = 0;
        double R[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, tt[3] = {0, 0, 0};
        if (K && k >= min_points) st = ref_find_E_ransac(K, li.data(), lj.data(), k, rs_iters, rs_thr, rs_min_inliers, R, tt, inl.data(), &ni) ? 2 : 1;
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

// Multi-threaded scoring baseline: hypotheses split across threads.
int ref_ransac_score_mt(const double* xi, const double* xj, int n, const double* E, int H, double thr, int* counts,
                        int threads) {
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++) {
    pool.emplace_back([=]() {
      for (int h = t; h < H; h += threads) {
        const Mat33 Eh = mat(E + 9 * h);
        int c = 0;
        for (int i = 0; i < n; i++) {
          const double e = sampson_err(Eh, Vec2{xi[2 * i], xi[2 * i + 1]}, Vec2{xj[2 * i], xj[2 * i + 1]});
          if (e < thr) c++;
        }
        counts[h] = c;
      }
    });
  }
  for (auto& th : pool) th.join();
  return 0;
}

const char* ref_kind() { return "reference"; }

// Batched triangulate_dlt (:1477-1516): poses are PoseCW (:157-168) as 9 doubles R (camera->world, row-major) + 3 doubles
// camera centre; track k observes ui[k] in pose ia[k] and uj[k] in pose ib[k].
int ref_triangulate_dlt(const double* K, const double* poses, int P, const int* ia, const int* ib, const double* ui, const double* uj,
                        int n, double* X) {
  std::vector<PoseCW> ps((size_t)P);
  for (int p = 0; p < P; p++) {
    for (int k = 0; k < 9; k++) ps[p].R.a[k] = poses[12 * p + k];
    ps[p].t = Vec3{poses[12 * p + 9], poses[12 * p + 10], poses[12 * p + 11]};
  }
  const Mat33 Km = mat(K);
  for (int k = 0; k < n; k++) {
    const Vec3 x = triangulate_dlt(Km, ps[ia[k]], ps[ib[k]], Vec2{ui[2 * k], ui[2 * k + 1]}, Vec2{uj[2 * k], uj[2 * k + 1]});
    X[3 * k] = x.x;
    X[3 * k + 1] = x.y;
    X[3 * k + 2] = x.z;
  }
  return 0;
}

// Loop-closure descriptor (:1100-1122) of one image: 1024 floats.
int ref_global_desc32(const uint8_t* pix, int w, int h, float* out) {
  const std::vector<float> v = global_desc_32(wrap(pix, w, h));
  std::memcpy(out, v.data(), v.size() * sizeof(float));
  return (int)v.size();
}

// The candidate search of :1823-1831 over the first n_search descriptors: best_id (-1: none), best_score.
int ref_desc_search(const float* descs, int n_search, const float* query, float* scores, int* best_id, float* best_score) {
  const std::vector<float> q(query, query + 1024);
  int bid = -1;
  float bs = 0.0f;
  for (int kk = 0; kk < n_search; ++kk) {
    const std::vector<float> a(descs + (size_t)kk * 1024, descs + (size_t)(kk + 1) * 1024);
    const float s = dot_desc(a, q);
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
}
