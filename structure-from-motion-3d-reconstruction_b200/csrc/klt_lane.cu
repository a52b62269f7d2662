// klt_lane.cu — pyramidal Lucas-Kanade for LARGE batches: one LANE per feature (32 features per warp).
//
// Same reference functions as klt.cu (cpp/src/templering_sfm.cpp: sample_bilinear :183-198, track_one :402-422,
// lk_step :424-460, fwd/bwd + fb test :356-362 / :1845-1849) and the same bug-for-bug behaviour; what changes is
// the mapping.  Measured on B200 (scripts/ubench_pipes.cu): a DFMA/DADD holds the issue port of its
// sub-partition for ~2.3 cycles and does NOT overlap with ALU/LSU issue, so the cost of an LK iteration is
// (2.3 x FP64 instructions + 1 x everything else).  The warp-per-feature kernel of klt.cu spends 77 % of its issue
// slots on "everything else" (byte loads, conversions, shuffles, predication, the warp reduction).  Here a lane
// owns a whole window, so there are no shuffles and no reduction, each tile byte is loaded and converted once, and
// the FP64 instructions are ~75 % of the issue slots.
//
//   * Tile: the 14 x 14 u8 taps of both images around the window, 16-byte rows, private to the lane in shared
//     memory (464 B per lane, lane stride chosen so that the per-row LDS.128 is bank-conflict free).  The tile is
//     re-staged only when the INTEGER window position changes; staging is cooperative (28 lanes fetch the 28 rows
//     of one feature with 16-byte loads and realign them), so global traffic stays sector-efficient.
//   * u8 -> double: one PRMT (byte extract) + one I2F.F64 per tile byte.  The conversion pipe runs beside the
//     FP64 pipe (measured: a DFMA + I2F.F64 pair costs what the I2F alone costs), so conversions only cost their
//     issue slot.  (Dropping the byte into the mantissa of a constant needs a second ALU instruction to zero the low
//     word of every register pair and would move the lerp roundings to ulp(4096).)
//   * One fractional offset (fx, fy) = position - floor(position) is shared by the whole window (the reference
//     re-rounds x + dx per tap, which jitters each sample position by <= ulp(x)/2 ~ 1e-13 px).
//   * Only INTERIOR windows are handled (every tap of the 14 x 14 region, plus the reference's round-up tap, inside
//     the image).  A feature whose window touches the border, or whose position is not finite, is appended to a
//     "deferred" list and recomputed from scratch by the exact warp-per-feature kernel of klt.cu (which implements
//     the out-of-bounds rule :188 tap by tap).
#include "common.cuh"

namespace {

constexpr int LR = 5;                        // window radius handled here (the reference default, LKConfig :313)
constexpr int LN = 2 * LR + 4;               // 14 tap rows / columns
constexpr int LNC = 2 * LR + 3;              // 13 grid rows / columns (window + 1 each side)
constexpr int LIMG = LN * 16;                // bytes of one staged image tile
constexpr int LSTRIDE = 2 * LIMG + 16;       // bytes per lane: 116 words, == 20 (mod 32) -> LDS.128 conflict free
constexpr int LWARPS = 1;                    // warps per block (no block-level cooperation)

__device__ __forceinline__ double off_byte(uint32_t w, int k, uint32_t) {
  // PRMT extracts the byte, the conversion pipe (I2F.F64, overlaps with FP64 issue) widens it.  Inline PTX keeps
  // the compiler from rewriting (double)c - (double)a as a second conversion of the integer difference.
  const uint32_t b = __byte_perm(w, 0, 0x4440 | k);
  double d;
  asm("cvt.rn.f64.u32 %0, %1;" : "=d"(d) : "r"(b));
  return d;
}

// 14 consecutive tap bytes of one tile row as doubles
__device__ __forceinline__ void row_bytes(const uint4 q, uint32_t c4096, double (&v)[LN]) {
  v[0] = off_byte(q.x, 0, c4096);
  v[1] = off_byte(q.x, 1, c4096);
  v[2] = off_byte(q.x, 2, c4096);
  v[3] = off_byte(q.x, 3, c4096);
  v[4] = off_byte(q.y, 0, c4096);
  v[5] = off_byte(q.y, 1, c4096);
  v[6] = off_byte(q.y, 2, c4096);
  v[7] = off_byte(q.y, 3, c4096);
  v[8] = off_byte(q.z, 0, c4096);
  v[9] = off_byte(q.z, 1, c4096);
  v[10] = off_byte(q.z, 2, c4096);
  v[11] = off_byte(q.z, 3, c4096);
  v[12] = off_byte(q.w, 0, c4096);
  v[13] = off_byte(q.w, 1, c4096);
}

// The five sums of lk_step (:431-448) for one interior window.  tile: this lane's staged taps (image B = I1 first,
// then image A = I0).  Returns 4*A and 2*b like klt.cu (the 0.5 factors are folded into the solve).
__device__ __forceinline__ void window_sums(const unsigned char* tile, double fx, double fy, uint32_t c4096, double& a00,
                                            double& a01, double& a11, double& b0, double& b1) {
  const uint4* t1 = reinterpret_cast<const uint4*>(tile);
  const uint4* t0 = reinterpret_cast<const uint4*>(tile + LIMG);
  double h1p[LNC], sa[LNC], sb[LNC], sc[LNC];
  double h0p[LNC];
  a00 = a01 = a11 = b0 = b1 = 0.0;
#pragma unroll
  for (int v = 0; v < LN; v++) {
    double px[LN];
    double h1n[LNC];
    row_bytes(t1[v], c4096, px);
#pragma unroll
    for (int i = 0; i < LNC; i++) h1n[i] = __fma_rn(fx, px[i + 1] - px[i], px[i]);
    if (v >= 1) {
#pragma unroll
      for (int i = 0; i < LNC; i++) {
        sa[i] = sb[i];
        sb[i] = sc[i];
        sc[i] = __fma_rn(fy, h1n[i] - h1p[i], h1p[i]);  // grid row v-1 of I1
      }
    }
#pragma unroll
    for (int i = 0; i < LNC; i++) h1p[i] = h1n[i];
    if (v >= 2) {
      // I0 tap row v-1 (grid rows 1..11 need tap rows 1..12)
      double h0n[LNC];
      row_bytes(t0[v - 1], c4096, px);
#pragma unroll
      for (int i = 1; i <= LNC - 2; i++) h0n[i] = __fma_rn(fx, px[i + 1] - px[i], px[i]);
      if (v >= 3) {
        // pixel row g = v-2: sa = S1[g-1], sb = S1[g], sc = S1[g+1]; S0[g] from tap rows g (h0p) and g+1 (h0n)
#pragma unroll
        for (int i = 1; i <= LNC - 2; i++) {
          const double s0 = __fma_rn(fy, h0n[i] - h0p[i], h0p[i]);
          const double gx2 = sb[i + 1] - sb[i - 1];  // 2*Ix (:439)
          const double gy2 = sc[i] - sa[i];          // 2*Iy (:440)
          const double e = s0 - sb[i];               // I0 - I1 at the same location (:441-442)
          a00 = __fma_rn(gx2, gx2, a00);
          a01 = __fma_rn(gx2, gy2, a01);
          a11 = __fma_rn(gy2, gy2, a11);
          b0 = __fma_rn(gx2, e, b0);
          b1 = __fma_rn(gy2, e, b1);
        }
      }
#pragma unroll
      for (int i = 1; i <= LNC - 2; i++) h0p[i] = h0n[i];
    }
  }
}

// Cooperative staging of ONE feature's tiles: lanes 0..27 each fetch one 14-byte row segment (image = t / 14, row =
// t % 14) starting at byte column x0 with two aligned 16-byte loads, realign, and store one 16-byte smem row.
__device__ __forceinline__ void stage_feature(unsigned char* tile_s, const uint8_t* __restrict__ imgA, const uint8_t* __restrict__ imgB,
                                              int pitch, int x0, int y0, int lane) {
  if (lane < 2 * LN) {
    const int img = lane >= LN ? 1 : 0, row = lane - img * LN;
    const uint8_t* rowp = (img ? imgA : imgB) + (size_t)(y0 + row) * pitch;  // tile order: I1 (= image B) first
    const int xa = x0 & ~15, o = x0 & 15;
    uint32_t W[8];
    const uint4 lo = __ldg(reinterpret_cast<const uint4*>(rowp + xa));
    W[0] = lo.x; W[1] = lo.y; W[2] = lo.z; W[3] = lo.w;
    uint4 hi = make_uint4(0, 0, 0, 0);
    if (o > 2) hi = __ldg(reinterpret_cast<const uint4*>(rowp + xa + 16));  // needed bytes reach the next chunk
    W[4] = hi.x; W[5] = hi.y; W[6] = hi.z; W[7] = hi.w;
    const int q = o >> 2, sh = (o & 3) * 8;
    if (q & 2) { W[0] = W[2]; W[1] = W[3]; W[2] = W[4]; W[3] = W[5]; W[4] = W[6]; W[5] = W[7]; }
    if (q & 1) { W[0] = W[1]; W[1] = W[2]; W[2] = W[3]; W[3] = W[4]; W[4] = W[5]; }
    uint4 out;
    out.x = __funnelshift_r(W[0], W[1], sh);
    out.y = __funnelshift_r(W[1], W[2], sh);
    out.z = __funnelshift_r(W[2], W[3], sh);
    out.w = __funnelshift_r(W[3], W[4], sh);
    *reinterpret_cast<uint4*>(tile_s + img * LIMG + row * 16) = out;
  }
}

__global__ void __launch_bounds__(32 * LWARPS) klt_lane_kernel(KltLaunch k, int* __restrict__ defer_count, int* __restrict__ defer_list) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const unsigned FULL = 0xffffffffu;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* wtile = smem_raw + (size_t)warp * 32 * LSTRIDE;
  unsigned char* tile = wtile + lane * LSTRIDE;
  const long long total = (long long)k.npairs * k.cap;
  const long long g = ((long long)blockIdx.x * LWARPS + warp) * 32 + lane;
  const int pair = g < total ? (int)(g / k.cap) : 0;
  const int slot = (int)(g - (long long)pair * k.cap);
  const bool valid = g < total && (!k.counts || slot < k.counts[pair]);
  uint32_t c4096 = 0x40B00000u;
  asm volatile("" : "+r"(c4096));  // keep it in a register (PRMT takes register operands)

  double2 p0 = make_double2(0.0, 0.0);
  if (valid) p0 = k.p0[g];
  double px = p0.x, py = p0.y, x1 = 0.0, y1 = 0.0;
  int n_it = 0;
  bool alive = valid;
  const int ndir = k.pb ? 2 : 1;
  const int frA = k.fa0 + pair * k.fa_step, frB = k.fb0 + pair * k.fb_step;

  for (int dir = 0; dir < ndir; dir++) {
    const int fa = dir ? frB : frA, fb = dir ? frA : frB;  // track from image fa to image fb
    for (int l = k.pv.levels - 1; l >= 0; --l) {
      const int w = k.pv.w[l], h = k.pv.h[l], pitch = k.pv.pitch[l];
      const uint8_t* lbase = k.pv.base[l];
      const size_t fstride = k.pv.fstride[l];
      const double scale = 1.0 / (double)(1 << l);
      const double plx = px * scale, ply = py * scale;
      double dlx = 0.0, dly = 0.0;
      bool done = !alive;
      int tFX = INT_MIN, tFY = INT_MIN;
      for (int it = 0; it < k.iters; ++it) {
        if (!__any_sync(FULL, !done)) break;
        const double x = plx + dlx, y = ply + dly;
        const double fxx = floor(x), fyy = floor(y);
        int FX = 0, FY = 0;
        bool act = !done;
        if (act) {
          const bool finite_ok = (fabs(fxx) < 1.0e9) && (fabs(fyy) < 1.0e9);
          FX = finite_ok ? (int)fxx : 0;
          FY = finite_ok ? (int)fyy : 0;
          const bool interior =
              finite_ok && FX - LR - 1 >= 0 && FX + LR + 3 <= w - 1 && FY - LR - 1 >= 0 && FY + LR + 3 <= h - 1;
          if (!interior) {  // border / non-finite: the exact kernel redoes this feature
            alive = false;
            done = true;
            act = false;
          }
        }
        unsigned need = __ballot_sync(FULL, act && (FX != tFX || FY != tFY));
        if (need) {
          __syncwarp();
          while (need) {
            const int s = __ffs(need) - 1;
            need &= need - 1;
            const int sFX = __shfl_sync(FULL, FX, s), sFY = __shfl_sync(FULL, FY, s);
            const int sfa = __shfl_sync(FULL, fa, s), sfb = __shfl_sync(FULL, fb, s);
            stage_feature(wtile + s * LSTRIDE, lbase + (size_t)sfa * fstride, lbase + (size_t)sfb * fstride, pitch, sFX - LR - 1,
                          sFY - LR - 1, lane);
          }
          __syncwarp();
        }
        if (act) {
          tFX = FX;
          tFY = FY;
          double a00, a01, a11, b0, b1;
          window_sums(tile, x - fxx, y - fyy, c4096, a00, a01, a11, b0, b1);
          double sx = 0.0, sy = 0.0;
          const double det = a00 * a11 - a01 * a01;
          if (!(fabs(det) < 16.0 * 1e-9)) {  // |det| < 1e-9 of :452 on the 16x scaled determinant
            const double rd = 2.0 / det;
            sx = (a11 * b0 - a01 * b1) * rd;
            sy = (a00 * b1 - a01 * b0) * rd;
          }
          n_it++;
          dlx += sx;
          dly += sy;
          if (sx * sx + sy * sy < 1e-6) done = true;  // hypot(step) < 1e-3 (:416), tested on the step just added
        }
      }
      const double up = (double)(1 << l);
      px = (plx + dlx) * up;
      py = (ply + dly) * up;
    }
    if (dir == 0) {
      x1 = px;
      y1 = py;
    }
  }
  if (alive) {
    k.p1[g] = make_double2(x1, y1);
    if (k.pb) k.pb[g] = make_double2(px, py);
    if (k.nit) k.nit[g] = n_it;
    if (k.keep) {
      const double fbd = hypot(px - p0.x, py - p0.y);
      k.keep[g] = (fbd >= k.fb_thresh) ? 0 : 1;  // NaN is kept (:362)
    }
  }
  // deferred features: one atomic per warp, list order is irrelevant (results are written by slot)
  const unsigned dm = __ballot_sync(FULL, valid && !alive);
  if (dm) {
    int base = 0;
    if (lane == 0) base = atomicAdd(defer_count, __popc(dm));
    base = __shfl_sync(FULL, base, 0);
    if (valid && !alive) defer_list[base + __popc(dm & ((1u << lane) - 1u))] = (int)g;
  }
}

}  // namespace

// Runs the lane kernel over all slots; features it cannot handle are appended to defer_list (device), count in
// *defer_count (device, must be zeroed by the caller on the same stream).
int sfm_klt_lane_launch(sfmgpu_ctx* ctx, const KltLaunch& k, int* defer_count, int* defer_list) {
  const size_t smem = (size_t)LWARPS * 32 * LSTRIDE;
  static bool configured = false;
  if (!configured) {
    SFM_CUDA(ctx, cudaFuncSetAttribute(klt_lane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const long long total = (long long)k.npairs * k.cap;
  if (total == 0) return 0;
  const unsigned grid = sfm_cdiv(total, 32 * LWARPS);
  SFM_LAUNCH(ctx, klt_lane_kernel, grid, 32 * LWARPS, smem, k, defer_count, defer_list);
  return 0;
}
