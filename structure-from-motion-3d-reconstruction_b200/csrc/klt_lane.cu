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

__device__ __forceinline__ double byte_f64(uint32_t w, int k) {
  // PRMT extracts the byte, the conversion pipe (I2F.F64, overlaps with FP64 issue) widens it.  Inline PTX keeps
  // the compiler from rewriting (double)c - (double)a as a second conversion of the integer difference.
  const uint32_t b = __byte_perm(w, 0, 0x4440 | k);
  double d;
  asm("cvt.rn.f64.u32 %0, %1;" : "=d"(d) : "r"(b));
  return d;
}

// 14 consecutive tap bytes of one tile row as doubles
__device__ __forceinline__ void row_bytes(const uint4 q, double (&v)[LN]) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < LN; i++) v[i] = byte_f64(w[i >> 2], i & 3);
}

// Horizontal lerps a + (b - a) * fx of one tile row (grid columns I0..I1-1 use taps i, i+1)
template <int I0, int I1>
__device__ __forceinline__ void hlerp_row(const unsigned char* rowp, double fx, double (&h)[LNC]) {
  double px[LN];
  row_bytes(*reinterpret_cast<const uint4*>(rowp), px);
#pragma unroll
  for (int i = I0; i < I1; i++) h[i] = __fma_rn(fx, px[i + 1] - px[i], px[i]);
}

// One tap row of the main window walk, straight-line.  v = tap row (3..13, warp-uniform, runtime).  Computes the
// horizontal lerps of I1 row v, completes grid row v-1 of I1 into `snew`, and accumulates pixel row g = v-2 from
// sold = S1[g-1], smid = S1[g], snew = S1[g+1] and S0[g] (I0 tap rows g = h0p and g+1 = tap row v-1).
// ACC: accumulator sets (2 = even / odd columns: shorter dependent DFMA chains).
template <int ACC>
__device__ __forceinline__ void walk_row(const unsigned char* tile, int v, double fx, double fy, double (&h1p)[LNC], double (&h0p)[LNC],
                                         const double (&sold)[LNC], const double (&smid)[LNC], double (&snew)[LNC],
                                         double (&acc)[ACC][5]) {
  double px[LN];
  row_bytes(*reinterpret_cast<const uint4*>(tile + v * 16), px);
#pragma unroll
  for (int i = 0; i < LNC; i++) {
    const double h = __fma_rn(fx, px[i + 1] - px[i], px[i]);
    snew[i] = __fma_rn(fy, h - h1p[i], h1p[i]);  // grid row v-1 of I1
    h1p[i] = h;
  }
  row_bytes(*reinterpret_cast<const uint4*>(tile + LIMG + (v - 1) * 16), px);
#pragma unroll
  for (int i = 1; i <= LNC - 2; i++) {
    const double h = __fma_rn(fx, px[i + 1] - px[i], px[i]);
    const double s0 = __fma_rn(fy, h - h0p[i], h0p[i]);
    h0p[i] = h;
    const double gx2 = smid[i + 1] - smid[i - 1];  // 2*Ix (:439)
    const double gy2 = snew[i] - sold[i];          // 2*Iy (:440)
    const double e = s0 - smid[i];                 // I0 - I1 at the same location (:441-442)
    double(&a)[5] = acc[i % ACC];
    a[0] = __fma_rn(gx2, gx2, a[0]);
    a[1] = __fma_rn(gx2, gy2, a[1]);
    a[2] = __fma_rn(gy2, gy2, a[2]);
    a[3] = __fma_rn(gx2, e, a[3]);
    a[4] = __fma_rn(gy2, e, a[4]);
  }
}

// The five sums of lk_step (:431-448) for one interior window.  tile: this lane's staged taps (image B = I1 first,
// then image A = I0).  Returns 4*A and 2*b like klt.cu (the 0.5 factors are folded into the solve).
//
// Tap rows 0..2 only feed lerps (no pixel row is complete yet) and are peeled.  Rows 3..13 are a ROLLED loop over
// groups of three rows (the three S1 row buffers rotate by argument order), so the hot code is ~16 KB and stays in
// the 32 KB instruction cache.  Fully unrolled, the 14 rows are 53 KB and the kernel stalls on instruction fetch
// (measured: 'no_instruction' was the top stall reason).  The 15th row of the last group does not exist and is skipped.
template <int ACC>
__device__ __forceinline__ void window_sums(const unsigned char* tile, double fx, double fy, double& a00, double& a01, double& a11,
                                            double& b0, double& b1) {
  double h1p[LNC], h0p[LNC], s0[LNC], s1[LNC], s2[LNC], t[LNC];
  double acc[ACC][5];
#pragma unroll
  for (int j = 0; j < ACC; j++)
#pragma unroll
    for (int q = 0; q < 5; q++) acc[j][q] = 0.0;
  hlerp_row<0, LNC>(tile, fx, h1p);  // I1 tap row 0
  hlerp_row<0, LNC>(tile + 16, fx, t);  // I1 tap row 1 -> grid row 0
#pragma unroll
  for (int i = 0; i < LNC; i++) {
    s0[i] = __fma_rn(fy, t[i] - h1p[i], h1p[i]);
    h1p[i] = t[i];
  }
  hlerp_row<0, LNC>(tile + 32, fx, t);  // I1 tap row 2 -> grid row 1
#pragma unroll
  for (int i = 0; i < LNC; i++) {
    s1[i] = __fma_rn(fy, t[i] - h1p[i], h1p[i]);
    h1p[i] = t[i];
  }
  hlerp_row<1, LNC - 1>(tile + LIMG + 16, fx, h0p);  // I0 tap row 1
  // step v writes grid row v-1 into buffer (v-1) % 3 and reads grid rows v-3, v-2 from the other two
#pragma unroll 1
  for (int v = 3; v < LN; v += 3) {
    walk_row<ACC>(tile, v, fx, fy, h1p, h0p, s0, s1, s2, acc);      // writes s2
    walk_row<ACC>(tile, v + 1, fx, fy, h1p, h0p, s1, s2, s0, acc);  // writes s0
    if (v + 2 < LN) walk_row<ACC>(tile, v + 2, fx, fy, h1p, h0p, s2, s0, s1, acc);  // writes s1
  }
  a00 = acc[0][0]; a01 = acc[0][1]; a11 = acc[0][2]; b0 = acc[0][3]; b1 = acc[0][4];
#pragma unroll
  for (int j = 1; j < ACC; j++) {
    a00 += acc[j][0]; a01 += acc[j][1]; a11 += acc[j][2]; b0 += acc[j][3]; b1 += acc[j][4];
  }
}

// Cooperative staging of one feature's tiles, split in two halves so that the loads of several features are in
// flight together: lanes 0..27 each fetch one 14-byte row segment (image = t / 14, row = t % 14) starting at byte
// column x0 with two aligned 16-byte loads (stage_load), then realign and store one 16-byte smem row (stage_store).
struct StageRegs {
  uint4 lo, hi;
};

__device__ __forceinline__ void stage_load(StageRegs& r, const uint8_t* __restrict__ imgA, const uint8_t* __restrict__ imgB, int pitch,
                                           int x0, int y0, int lane) {
  const int img = lane >= LN ? 1 : 0, row = lane - img * LN;
  const uint8_t* rowp = (img ? imgA : imgB) + (size_t)(y0 + row) * pitch;  // tile order: I1 (= image B) first
  const int xa = x0 & ~15, o = x0 & 15;
  r.lo = __ldg(reinterpret_cast<const uint4*>(rowp + xa));
  r.hi = make_uint4(0, 0, 0, 0);
  if (o > 2) r.hi = __ldg(reinterpret_cast<const uint4*>(rowp + xa + 16));  // needed bytes reach the next chunk
}

__device__ __forceinline__ void stage_store(const StageRegs& r, unsigned char* tile_s, int x0, int lane) {
  const int img = lane >= LN ? 1 : 0, row = lane - img * LN;
  const int o = x0 & 15, q = o >> 2, sh = (o & 3) * 8;
  uint32_t W[8] = {r.lo.x, r.lo.y, r.lo.z, r.lo.w, r.hi.x, r.hi.y, r.hi.z, r.hi.w};
  if (q & 2) { W[0] = W[2]; W[1] = W[3]; W[2] = W[4]; W[3] = W[5]; W[4] = W[6]; W[5] = W[7]; }
  if (q & 1) { W[0] = W[1]; W[1] = W[2]; W[2] = W[3]; W[3] = W[4]; W[4] = W[5]; }
  uint4 out;
  out.x = __funnelshift_r(W[0], W[1], sh);
  out.y = __funnelshift_r(W[1], W[2], sh);
  out.z = __funnelshift_r(W[2], W[3], sh);
  out.w = __funnelshift_r(W[3], W[4], sh);
  *reinterpret_cast<uint4*>(tile_s + img * LIMG + row * 16) = out;
}

template <int ACC, int MINB, int LSTAGE>
__global__ void __launch_bounds__(32 * LWARPS, MINB) klt_lane_kernel(KltLaunch k, int* __restrict__ defer_count, int* __restrict__ defer_list) {
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

  double2 p0 = make_double2(0.0, 0.0);
  if (valid) p0 = k.p0[g];
  double px = p0.x, py = p0.y, x1 = 0.0, y1 = 0.0;
  int n_it = 0;
  bool alive = valid;
  const int ndir = k.pb ? 2 : 1;

  for (int dir = 0; dir < ndir; dir++) {
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
            // up to LSTAGE features per round: all their loads are issued before the first realign / store
            int ss[LSTAGE], sx0[LSTAGE];
            StageRegs sr[LSTAGE];
#pragma unroll
            for (int j = 0; j < LSTAGE; j++) {
              ss[j] = need ? __ffs(need) - 1 : -1;
              need &= need - 1;
              const int src = ss[j] < 0 ? 0 : ss[j];
              const int sFX = __shfl_sync(FULL, FX, src), sFY = __shfl_sync(FULL, FY, src), spair = __shfl_sync(FULL, pair, src);
              const int sA = k.fa0 + spair * k.fa_step, sB = k.fb0 + spair * k.fb_step;
              sx0[j] = sFX - LR - 1;
              if (ss[j] >= 0 && lane < 2 * LN)
                stage_load(sr[j], lbase + (size_t)(dir ? sB : sA) * fstride, lbase + (size_t)(dir ? sA : sB) * fstride, pitch, sx0[j],
                           sFY - LR - 1, lane);
            }
#pragma unroll
            for (int j = 0; j < LSTAGE; j++)
              if (ss[j] >= 0 && lane < 2 * LN) stage_store(sr[j], wtile + ss[j] * LSTRIDE, sx0[j], lane);
          }
          __syncwarp();
        }
        if (act) {
          tFX = FX;
          tFY = FY;
          double a00, a01, a11, b0, b1;
          window_sums<ACC>(tile, x - fxx, y - fyy, a00, a01, a11, b0, b1);
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

template <int ACC, int MINB, int LSTAGE>
static int lane_launch(sfmgpu_ctx* ctx, const KltLaunch& k, int* defer_count, int* defer_list) {
  const size_t smem = (size_t)LWARPS * 32 * LSTRIDE;
  static bool configured = false;
  if (!configured) {
    SFM_CUDA(ctx, cudaFuncSetAttribute(klt_lane_kernel<ACC, MINB, LSTAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const long long total = (long long)k.npairs * k.cap;
  if (total == 0) return 0;
  const unsigned grid = sfm_cdiv(total, 32 * LWARPS);
  SFM_LAUNCH(ctx, (klt_lane_kernel<ACC, MINB, LSTAGE>), grid, 32 * LWARPS, smem, k, defer_count, defer_list);
  return 0;
}

// Runs the lane kernel over all slots; features it cannot handle are appended to defer_list (device), count in
// *defer_count (device, must be zeroed by the caller on the same stream).  variant: tuning builds (A/B runs).
int sfm_klt_lane_launch(sfmgpu_ctx* ctx, const KltLaunch& k, int* defer_count, int* defer_list, int variant) {
  // Measured on B200, C2 shape (scripts/klt_ab.py, KLT stage incl. the deferred pass, 238k tracks): <1,12,4> 4.43 ms
  // (168 registers, spills), <1,8,4> 3.57 ms, <2,8,4> 3.66 ms, <2,8,8> 3.97 ms: the schedule wants registers, not warps.
  switch (variant) {
    case 1: return lane_launch<1, 12, 4>(ctx, k, defer_count, defer_list);
    case 2: return lane_launch<2, 8, 4>(ctx, k, defer_count, defer_list);
    default: return lane_launch<1, 8, 4>(ctx, k, defer_count, defer_list);
  }
}
