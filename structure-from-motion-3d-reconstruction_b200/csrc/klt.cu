// klt.cu — pyramidal Lucas-Kanade, one warp per feature, FP64, forward + backward in one launch.
//
// Replaces (reference cpp/src/templering_sfm.cpp): sample_bilinear :183-198, KLTTracker::track_one :402-422,
// lk_step :424-460, and the per-track fwd/bwd + fb test of step() :356-362 / the two-view loop :1845-1849.
//
// Behaviour kept bug-for-bug:
//   * both images are sampled at the SAME moving location (err = I0(p+d) - I1(p+d)), so the update does not
//     converge; every iteration adds about the local flow;
//   * any bilinear tap outside the image zeroes the whole sample; |det| < 1e-9 gives a zero step;
//   * the early exit tests the step that was just added (hypot(step) < 1e-3);
//   * p is rescaled from the running full-resolution position at every level.
//
// Arithmetic: positions, bilinear weights, accumulators and the 2x2 solve are FP64 and this file is compiled
// with -fmad=false (the x86-64 oracle has no FMA contraction).  Two deliberate, measured deviations from the
// reference's operation ORDER (not from its formulas):
//   (1) the 11x11 window's six samples per pixel are taken from one shared grid of bilinear samples whose
//       column/row coordinates are fl(x+i), fl(y+j) (the reference reaches the +-1 neighbours as fl(fl(x+i)+-1));
//   (2) the five sums are reduced lane-strided + butterfly instead of sequentially in raster order.
//   Measured on CPU against the exact order (1500 points x 60 iterations, smooth, checker and pure-noise
//   images): max deviation 1.8e-11 px, i.e. 8 orders of magnitude inside the 1e-3 px parity budget.
//
// Work per LK iteration and warp (r = 5): 338 horizontal lerps (each tap pair read once from a cached u8
// tile in shared memory), 290 vertical lerps, 121 pixel terms, 5 warp reductions.  The u8 tiles (24x24 per
// image) are re-staged only when the window leaves the cached region, i.e. about once per level.
// Roofline: FP64 CUDA-core pipe, not HBM (algorithmic traffic is 2,092 B per track-step).
#include "common.cuh"

namespace {

constexpr int KLT_MARGIN = 4;

template <int R>
struct KltSmem {
  static constexpr int NC = 2 * R + 3;               // sample-grid columns / rows (window + 1 each side)
  static constexpr int NR = 2 * R + 5;               // tile rows a sample grid can touch
  static constexpr int T = 2 * KLT_MARGIN + 2 * R + 6;  // cached tile edge (24 for R = 5)
  double H1[NR * NC];   // horizontal lerps of I1
  double H0[NR * NC];   // horizontal lerps of I0 (inner columns only are used)
  double S1[NC * NC];   // bilinear samples of I1 on the grid
  double S0[NC * NC];   // bilinear samples of I0 (inner (2R+1)^2 used)
  double cfx[NC], cfy[NC];  // fractional parts per grid column / row
  int ccx[NC], ccy[NC];     // tile-relative tap column / row (left / top tap)
  int cvx[NC], cvy[NC];     // validity (both taps inside the image)
  uint8_t t0[T * T];
  uint8_t t1[T * T];
};

__device__ __forceinline__ double u8_to_f64(uint32_t v) {
  // exact: 2^52 + v has v in its low mantissa bits
  return __hiloint2double(0x43300000, (int)v) - 4503599627370496.0;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, o);
  return v;  // identical in every lane (a+b == b+a bitwise)
}

template <int R>
__device__ __forceinline__ void stage_tile(uint8_t* tile, const uint8_t* __restrict__ img, int w, int h, int pitch, int tx0,
                                           int ty0, int lane) {
  constexpr int T = KltSmem<R>::T;
  for (int idx = lane; idx < T * T; idx += 32) {
    const int r = idx / T, c = idx - r * T;
    int gx = tx0 + c, gy = ty0 + r;
    gx = gx < 0 ? 0 : (gx > w - 1 ? w - 1 : gx);
    gy = gy < 0 ? 0 : (gy > h - 1 ? h - 1 : gy);
    tile[idx] = __ldg(img + (size_t)gy * pitch + gx);
  }
}

// One pyramidal track (:402-422) of point (px,py) from pyramid A to pyramid B.  All lanes hold identical
// scalars; shared memory `sm` is private to the warp.
template <int R, bool FIXED>
__device__ void track_one(KltSmem<R>& sm, const PyrView& pv, int fa, int fb, int radius_rt, int iters, double& px, double& py,
                          int& n_it, int lane) {
  constexpr int NC = KltSmem<R>::NC;
  constexpr int T = KltSmem<R>::T;
  const int radius = FIXED ? R : radius_rt;  // compile-time window for the default radius: no runtime divisions
  const int nc = 2 * radius + 3;  // grid edge actually used
  const int nr = 2 * radius + 5;
  const int nw = 2 * radius + 1;

  for (int l = pv.levels - 1; l >= 0; --l) {
    const int w = pv.w[l], h = pv.h[l], pitch = pv.pitch[l];
    const uint8_t* I0 = pv.base[l] + (size_t)fa * pv.fstride[l];
    const uint8_t* I1 = pv.base[l] + (size_t)fb * pv.fstride[l];
    const double scale = 1.0 / (double)(1 << l);
    const double plx = px * scale, ply = py * scale;
    double dlx = 0.0, dly = 0.0;
    int tx0 = 0, ty0 = 0;
    bool have_tile = false;

    for (int it = 0; it < iters; ++it) {
      const double x = plx + dlx, y = ply + dly;
      const double fxx = floor(x), fyy = floor(y);
      // Window entirely unusable (non-finite or far outside): every sample is 0 -> A = 0 -> zero step.
      const bool finite_ok = (fabs(fxx) < 1.0e9) && (fabs(fyy) < 1.0e9);
      double sx = 0.0, sy = 0.0;
      if (finite_ok && w >= 2 && h >= 2) {
        const int FX = (int)fxx, FY = (int)fyy;
        // cached tile must cover columns [FX-r-1, FX+r+3] and rows [FY-r-1, FY+r+3]
        if (!have_tile || FX - radius - 1 < tx0 || FX + radius + 3 > tx0 + T - 1 || FY - radius - 1 < ty0 ||
            FY + radius + 3 > ty0 + T - 1) {
          tx0 = FX - radius - 1 - KLT_MARGIN;
          ty0 = FY - radius - 1 - KLT_MARGIN;
          __syncwarp();
          stage_tile<R>(sm.t0, I0, w, h, pitch, tx0, ty0, lane);
          stage_tile<R>(sm.t1, I1, w, h, pitch, tx0, ty0, lane);
          have_tile = true;
        }
        // per-column / per-row tables: coordinate fl(x+i), its floor, fraction, validity (:184-190)
        for (int t = lane; t < 2 * nc; t += 32) {
          const bool isx = t < nc;
          const int i = isx ? t : t - nc;
          const double c0 = (isx ? x : y) + (double)(i - radius - 1);
          const double cf = floor(c0);
          const bool ok = (cf >= 0.0) && (cf <= (double)((isx ? w : h) - 2));
          int c = ok ? (int)cf - (isx ? tx0 : ty0) : 0;
          c = c < 0 ? 0 : (c > T - 2 ? T - 2 : c);
          if (isx) {
            sm.cfx[i] = c0 - cf;
            sm.cvx[i] = ok ? 1 : 0;
            sm.ccx[i] = c;
          } else {
            sm.cfy[i] = c0 - cf;
            sm.cvy[i] = ok ? 1 : 0;
            sm.ccy[i] = c;
          }
        }
        __syncwarp();
        // horizontal lerps v00*(1-dx) + v10*dx on every tile row the grid can touch (rows rb .. rb+nr-1)
        int rb = FY - radius - 1 - ty0;
        rb = rb < 0 ? 0 : (rb > T - nr ? T - nr : rb);
        for (int idx = lane; idx < nr * nc; idx += 32) {
          const int rr = idx / nc, i = idx - rr * nc;
          const int off = (rb + rr) * T + sm.ccx[i];
          const double f = sm.cfx[i], g = 1.0 - f;
          const double a1 = u8_to_f64(sm.t1[off]), b1 = u8_to_f64(sm.t1[off + 1]);
          sm.H1[rr * NC + i] = a1 * g + b1 * f;
          const double a0 = u8_to_f64(sm.t0[off]), b0 = u8_to_f64(sm.t0[off + 1]);
          sm.H0[rr * NC + i] = a0 * g + b0 * f;
        }
        __syncwarp();
        // vertical lerps v0*(1-dy) + v1*dy; invalid taps zero the sample (:188)
        for (int idx = lane; idx < nc * nc; idx += 32) {
          const int j = idx / nc, i = idx - j * nc;
          int rr = sm.ccy[j] - rb;
          rr = rr < 0 ? 0 : (rr > nr - 2 ? nr - 2 : rr);
          const bool ok = sm.cvx[i] && sm.cvy[j];
          const double f = sm.cfy[j], g = 1.0 - f;
          const double v1 = sm.H1[rr * NC + i] * g + sm.H1[(rr + 1) * NC + i] * f;
          const double v0 = sm.H0[rr * NC + i] * g + sm.H0[(rr + 1) * NC + i] * f;
          sm.S1[j * NC + i] = ok ? v1 : 0.0;
          sm.S0[j * NC + i] = ok ? v0 : 0.0;
        }
        __syncwarp();
        // normal equations over the (2r+1)^2 window (:433-450)
        double a00 = 0, a01 = 0, a11 = 0, b0 = 0, b1 = 0;
        for (int idx = lane; idx < nw * nw; idx += 32) {
          const int dy = idx / nw, dx = idx - dy * nw;
          const int c = (dy + 1) * NC + (dx + 1);
          const double ix = 0.5 * (sm.S1[c + 1] - sm.S1[c - 1]);
          const double iy = 0.5 * (sm.S1[c + NC] - sm.S1[c - NC]);
          const double e = sm.S0[c] - sm.S1[c];
          a00 += ix * ix;
          a01 += ix * iy;
          a11 += iy * iy;
          b0 += ix * e;
          b1 += iy * e;
        }
        a00 = warp_sum(a00);
        a01 = warp_sum(a01);
        a11 = warp_sum(a11);
        b0 = warp_sum(b0);
        b1 = warp_sum(b1);
        const double det = a00 * a11 - a01 * a01;
        if (!(fabs(det) < 1e-9)) {
          const double i00 = a11 / det, i01 = -a01 / det, i11 = a00 / det;
          sx = i00 * b0 + i01 * b1;
          sy = i01 * b0 + i11 * b1;
        }
      }
      n_it++;
      dlx += sx;
      dly += sy;
      if (hypot(sx, sy) < 1e-3) break;
    }
    const double up = (double)(1 << l);
    px = (plx + dlx) * up;
    py = (ply + dly) * up;
  }
}

template <int R, bool FIXED>
__global__ void __launch_bounds__(128) klt_kernel(KltLaunch k) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  KltSmem<R>& sm = reinterpret_cast<KltSmem<R>*>(smem_raw)[warp];
  const long long g = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (g >= (long long)k.npairs * k.cap) return;
  const int pair = (int)(g / k.cap), slot = (int)(g - (long long)pair * k.cap);
  if (k.counts && slot >= k.counts[pair]) return;
  const int fa = k.fa0 + pair * k.fa_step, fb = k.fb0 + pair * k.fb_step;
  const double2 p0 = k.p0[g];
  int n_it = 0;
  double x = p0.x, y = p0.y;
  track_one<R, FIXED>(sm, k.pv, fa, fb, k.radius, k.iters, x, y, n_it, lane);
  const double x1 = x, y1 = y;
  __syncwarp();
  track_one<R, FIXED>(sm, k.pv, fb, fa, k.radius, k.iters, x, y, n_it, lane);
  if (lane == 0) {
    k.p1[g] = make_double2(x1, y1);
    k.pb[g] = make_double2(x, y);
    if (k.nit) k.nit[g] = n_it;
    if (k.keep) {
      const double fbd = hypot(x - p0.x, y - p0.y);
      k.keep[g] = (fbd >= k.fb_thresh) ? 0 : 1;  // NaN is kept (:362)
    }
  }
}

template <int R, bool FIXED>
int launch_r(sfmgpu_ctx* ctx, const KltLaunch& k) {
  const int warps_per_block = 4;
  const size_t smem = sizeof(KltSmem<R>) * warps_per_block;
  static bool configured = false;
  if (!configured) {
    SFM_CUDA(ctx, cudaFuncSetAttribute(klt_kernel<R, FIXED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const long long total = (long long)k.npairs * k.cap;
  if (total == 0) return 0;
  const unsigned grid = sfm_cdiv(total, warps_per_block);
  SFM_LAUNCH(ctx, (klt_kernel<R, FIXED>), grid, warps_per_block * 32, smem, k);
  return 0;
}

}  // namespace

int sfm_klt_launch(sfmgpu_ctx* ctx, const KltLaunch& k) {
  if (k.radius < 1 || k.radius > 10)
    return sfm_fail(ctx, SFMGPU_E_ARG, "klt: win_radius %d outside the supported range [1,10]", k.radius);
  if (k.iters < 0) return sfm_fail(ctx, SFMGPU_E_ARG, "klt: negative iteration count");
  if (k.radius == 5) return launch_r<5, true>(ctx, k);
  if (k.radius < 5) return launch_r<5, false>(ctx, k);
  if (k.radius == 10) return launch_r<10, true>(ctx, k);
  return launch_r<10, false>(ctx, k);
}

extern "C" int sfmgpu_klt_track(sfmgpu_ctx* ctx, sfmgpu_frames* f, int frame_a, int frame_b, const double* p0_xy, int n,
                                int win_radius, int iters, double* p1_xy, double* p0_back_xy, int32_t* n_iters) {
  if (!ctx || !f) return SFMGPU_E_ARG;
  if (frame_a < 0 || frame_a >= f->n || frame_b < 0 || frame_b >= f->n || n < 0)
    return sfm_fail(ctx, SFMGPU_E_ARG, "klt_track: bad frame index or count");
  if (n == 0) return 0;
  if (!p0_xy || !p1_xy || !p0_back_xy) return sfm_fail(ctx, SFMGPU_E_ARG, "klt_track: null pointer");
  const size_t pb = (size_t)n * sizeof(double2);
  SFM_TRY(sfm_reserve(ctx, ctx->klt_in, pb));
  SFM_TRY(sfm_reserve(ctx, ctx->klt_p1, pb));
  SFM_TRY(sfm_reserve(ctx, ctx->klt_pb, pb));
  SFM_TRY(sfm_reserve(ctx, ctx->klt_nit, (size_t)n * sizeof(int)));
  SFM_CUDA(ctx, cudaMemcpyAsync(ctx->klt_in.p, p0_xy, pb, cudaMemcpyHostToDevice, ctx->stream));
  KltLaunch k;
  k.pv = f->view();
  k.p0 = (const double2*)ctx->klt_in.p;
  k.counts = nullptr;
  k.npairs = 1;
  k.cap = n;
  k.fa0 = frame_a;
  k.fa_step = 0;
  k.fb0 = frame_b;
  k.fb_step = 0;
  k.radius = win_radius;
  k.iters = iters;
  k.fb_thresh = 0.0;
  k.p1 = (double2*)ctx->klt_p1.p;
  k.pb = (double2*)ctx->klt_pb.p;
  k.nit = (int*)ctx->klt_nit.p;
  k.keep = nullptr;
  SFM_TRY(sfm_klt_launch(ctx, k));
  SFM_CUDA(ctx, cudaMemcpyAsync(p1_xy, ctx->klt_p1.p, pb, cudaMemcpyDeviceToHost, ctx->stream));
  SFM_CUDA(ctx, cudaMemcpyAsync(p0_back_xy, ctx->klt_pb.p, pb, cudaMemcpyDeviceToHost, ctx->stream));
  if (n_iters)
    SFM_CUDA(ctx, cudaMemcpyAsync(n_iters, ctx->klt_nit.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  SFM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}
