// corner_score.cu — Shi-Tomasi min-eigenvalue score, global maximum, and the raster-ordered candidate list.
//
// Replaces (reference cpp/src/templering_sfm.cpp) shi_tomasi :237-285: clamped central-difference gradients
// (:242-249), 5x5 structure tensor (:252-264), lmin = 0.5*(tr - sqrt(max(0, tr^2 - 4 det))) (:266-270), zero
// 2-px border (:253-254), thr = max*quality (:274-275), candidates s >= thr in raster order (:280-285).
//
// Exactness.  With Gx = I[x+1]-I[x-1], Gy likewise (integers in [-255,255]; the reference uses gx = Gx/2), let
// a = sum Gx^2, b = sum Gy^2, c = sum GxGy over the window (int32 <= 1.63e6).  Every intermediate of the
// reference up to disc is a dyadic rational below 2^53, hence exact in double, and
//        lmin = 0.125 * fl( (a+b) - fl(sqrt( (a-b)^2 + 4c^2 )) )
// bit for bit (scaling by powers of two commutes with rounding).  The kernels therefore keep integer sums and
// evaluate that single expression in FP64 (IEEE sqrt) — only for pixels that pass an exact INTEGER screen
// D <= (T - theta)^2 against the running threshold, so the FP64 pipe sees a few percent of the pixels.
//
// Kernels (tile = 64 x 28 output pixels, 256 threads, taps staged once in shared memory with clamping):
//   score_tile_kernel<0>  per-frame maximum of u = 8*lmin  (atomicMax on the double's bit pattern)
//   score_tile_kernel<1>  candidate bitmap (one ballot word per 32 pixels) + unordered (pixel, score) list
//   bitmap_scan_kernel    exclusive prefix of the bitmap popcounts -> raster rank of every candidate
//   order_kernel          scatter the unordered list to raster order (the order std::sort starts from)
// Roofline: the structure tensor costs ~40 integer instructions per pixel against 1 B/pixel of traffic, so
// this stage is bound by the integer/shared-memory pipes, not HBM.
#include "common.cuh"
#include "corner_work.cuh"

namespace {

constexpr int TW = 64, TH = 28;
constexpr int TAP_W = TW + 6, TAP_H = TH + 6, TAP_S = 76;  // 19-word row stride: conflict-free column walks
constexpr int HS_ROWS = TH + 4, HS_S = 65;                 // horizontal sums, padded rows

__device__ __forceinline__ double i2d(int v) {
  // exact int32 -> double without the slow conversion pipe: (2^52 + 2^31 + v) - (2^52 + 2^31)
  return __hiloint2double(0x43300000, v ^ 0x80000000) - 4503601774854144.0;
}

// u = 8*lmin, bit-exact with the reference's lmin/0.125 (see header).
__device__ __forceinline__ double exact_u(int a, int b, int c) {
  const double dd = i2d(a - b), c2 = i2d(2 * c);
  const double D = dd * dd + c2 * c2;
  return i2d(a + b) - sqrt(D);
}

// Exact integer screen: can u reach theta_i (an integer <= theta - 1)?
__device__ __forceinline__ bool may_reach(int a, int b, int c, int theta_i) {
  if (theta_i <= 0) return true;
  const int T = a + b;
  if (T < theta_i) return false;
  const long long d = a - b, cc = c, m = T - theta_i;
  return d * d + 4 * cc * cc <= m * m;
}

template <int MODE>
__global__ void __launch_bounds__(256) score_tile_kernel(const uint8_t* __restrict__ img, int w, int h, int pitch, size_t fstride,
                                                        int frame0, CornerWorkView wv, double quality) {
  __shared__ __align__(16) uint8_t taps[TAP_H * TAP_S];
  __shared__ int hxx[HS_ROWS * HS_S], hxy[HS_ROWS * HS_S], hyy[HS_ROWS * HS_S];
  __shared__ double wmax[8];
  // MODE 1: pixels that pass the integer screen are queued (local index) and scored densely afterwards, so the
  // FP64 sqrt path runs with full warps instead of once per warp-row that holds a single candidate
  __shared__ unsigned short queue[MODE == 1 ? TW * TH : 1];
  __shared__ unsigned tile_bm[MODE == 1 ? TH * 2 : 1];
  __shared__ int qn;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int fr = blockIdx.z;
  if (MODE == 1) {
    if (tid < TH * 2) tile_bm[tid] = 0;
    if (tid == 0) qn = 0;
  }
  const uint8_t* im = img + (size_t)(frame0 + fr) * fstride;
  const int X0 = blockIdx.x * TW, Y0 = blockIdx.y * TH;

  // phase 0: taps of image(X0-3 .., Y0-3 ..) with clamped coordinates (= the reference's clamped gradient taps).
  // Shared rows start at image column X0-4 (4-aligned), so tap tx lives at byte tx+1.  Tiles whose halo lies
  // inside the image are staged with aligned 32-bit loads (all loads issued before the stores); border tiles
  // take the byte path, which implements the clamp.
  if (X0 >= 4 && X0 + 68 <= w && Y0 >= 3 && Y0 + TAP_H - 3 <= h) {
    constexpr int WPT = (TAP_H * 18 + 255) / 256;
    uint32_t v[WPT];
    const uint8_t* src = im + (size_t)(Y0 - 3) * pitch + (X0 - 4);
#pragma unroll
    for (int k = 0; k < WPT; k++) {
      const int idx = tid + 256 * k, r = idx / 18, c = idx - r * 18;
      if (idx < TAP_H * 18) v[k] = __ldg(reinterpret_cast<const uint32_t*>(src + (size_t)r * pitch) + c);
    }
#pragma unroll
    for (int k = 0; k < WPT; k++) {
      const int idx = tid + 256 * k, r = idx / 18, c = idx - r * 18;
      if (idx < TAP_H * 18) reinterpret_cast<uint32_t*>(taps)[r * (TAP_S / 4) + c] = v[k];
    }
  } else {
    for (int idx = tid; idx < TAP_H * TAP_W; idx += 256) {
      const int ty = idx / TAP_W, tx = idx - ty * TAP_W;
      int gx = X0 - 3 + tx, gy = Y0 - 3 + ty;
      gx = gx < 0 ? 0 : (gx > w - 1 ? w - 1 : gx);
      gy = gy < 0 ? 0 : (gy > h - 1 ? h - 1 : gy);
      taps[ty * TAP_S + tx + 1] = __ldg(im + (size_t)gy * pitch + gx);
    }
  }
  __syncthreads();

  // phase 1: horizontal 5-sums of the gradient products.  warp = column group (8 outputs), lane = row.
  {
    const int r = lane, cb = warp * 8;
    const uint8_t* tm = taps + r * TAP_S + cb + 1;
    const uint8_t* tc = tm + TAP_S;
    const uint8_t* tp = tc + TAP_S;
    int pxx[12], pxy[12], pyy[12];
#pragma unroll
    for (int k = 0; k < 12; k++) {
      const int gx = (int)tc[k + 2] - (int)tc[k];
      const int gy = (int)tp[k + 1] - (int)tm[k + 1];
      pxx[k] = gx * gx;
      pxy[k] = gx * gy;
      pyy[k] = gy * gy;
    }
    int sxx = pxx[0] + pxx[1] + pxx[2] + pxx[3] + pxx[4];
    int sxy = pxy[0] + pxy[1] + pxy[2] + pxy[3] + pxy[4];
    int syy = pyy[0] + pyy[1] + pyy[2] + pyy[3] + pyy[4];
    int* oxx = hxx + r * HS_S + cb;
    int* oxy = hxy + r * HS_S + cb;
    int* oyy = hyy + r * HS_S + cb;
    oxx[0] = sxx;
    oxy[0] = sxy;
    oyy[0] = syy;
#pragma unroll
    for (int j = 1; j < 8; j++) {
      sxx += pxx[j + 4] - pxx[j - 1];
      sxy += pxy[j + 4] - pxy[j - 1];
      syy += pyy[j + 4] - pyy[j - 1];
      oxx[j] = sxx;
      oxy[j] = sxy;
      oyy[j] = syy;
    }
  }
  __syncthreads();

  // phase 2: vertical 5-sums, score, max / candidate test.  thread = column, 7 consecutive rows.
  const int oc = tid & 63, q = tid >> 6;
  const int x = X0 + oc;
  const int* cxx = hxx + (q * 7) * HS_S + oc;
  const int* cxy = hxy + (q * 7) * HS_S + oc;
  const int* cyy = hyy + (q * 7) * HS_S + oc;
  int a = cxx[0] + cxx[HS_S] + cxx[2 * HS_S] + cxx[3 * HS_S];
  int c = cxy[0] + cxy[HS_S] + cxy[2 * HS_S] + cxy[3 * HS_S];
  int b = cyy[0] + cyy[HS_S] + cyy[2 * HS_S] + cyy[3 * HS_S];

  unsigned long long* maxbits = wv.maxbits + fr;
  double thr8 = 0.0, best = 0.0;
  int theta_i;
  if (MODE == 0) {
    best = __longlong_as_double(*(volatile unsigned long long*)maxbits);  // what earlier blocks already found
    theta_i = (int)best - 1;
  } else {
    const double maxv = 0.125 * __longlong_as_double(*maxbits);
    thr8 = 8.0 * (maxv * quality);  // thr = maxv*quality (:275); s >= thr  <=>  u >= 8*thr
    theta_i = thr8 < 2.0e9 ? (int)thr8 - 1 : 2000000000;
  }
  const bool col_in = x < w, col_interior = x >= 2 && x < w - 2;

#pragma unroll
  for (int j = 0; j < 7; j++) {
    a += cxx[(j + 4) * HS_S];
    c += cxy[(j + 4) * HS_S];
    b += cyy[(j + 4) * HS_S];
    const int y = Y0 + q * 7 + j;
    const bool interior = col_interior && y >= 2 && y < h - 2;
    if (MODE == 0) {
      if (interior && may_reach(a, b, c, theta_i)) {
        const double u = exact_u(a, b, c);
        if (u > best) {
          best = u;
          theta_i = (int)u - 1;
        }
      }
    } else {
      if (col_in && y < h) {
        const int loc = (q * 7 + j) * TW + oc;
        if (interior) {
          if (may_reach(a, b, c, theta_i)) queue[atomicAdd(&qn, 1)] = (unsigned short)loc;
        } else if (0.0 >= thr8) {
          queue[atomicAdd(&qn, 1)] = (unsigned short)(loc | 0x8000);  // border score is exactly 0 (:240, :253-254)
        }
      }
    }
    a -= cxx[j * HS_S];
    c -= cxy[j * HS_S];
    b -= cyy[j * HS_S];
  }

  if (MODE == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0) wmax[warp] = best;
    __syncthreads();
    if (tid == 0) {
      double m = wmax[0];
      for (int k = 1; k < 8; k++) m = fmax(m, wmax[k]);
      const unsigned long long mb = (unsigned long long)__double_as_longlong(m);
      if (mb > *(volatile unsigned long long*)maxbits) atomicMax(maxbits, mb);  // u >= 0: bit order == value order
    }
  } else {
    __syncthreads();
    const int nq = qn;
    for (int e0 = 0; e0 < nq; e0 += 256) {  // uniform trip count: the ballots below are warp-complete
      const int e = e0 + tid;
      bool cand = false;
      double u = 0.0;
      int row = 0, col = 0;
      if (e < nq) {
        const int id = queue[e], loc = id & 0x7fff;
        row = loc / TW;
        col = loc - row * TW;
        cand = true;
        if (!(id & 0x8000)) {
          const int* pxx = hxx + row * HS_S + col;
          const int* pxy = hxy + row * HS_S + col;
          const int* pyy = hyy + row * HS_S + col;
          const int a = pxx[0] + pxx[HS_S] + pxx[2 * HS_S] + pxx[3 * HS_S] + pxx[4 * HS_S];
          const int c = pxy[0] + pxy[HS_S] + pxy[2 * HS_S] + pxy[3 * HS_S] + pxy[4 * HS_S];
          const int b = pyy[0] + pyy[HS_S] + pyy[2 * HS_S] + pyy[3 * HS_S] + pyy[4 * HS_S];
          u = exact_u(a, b, c);
          cand = u >= thr8;
        }
      }
      const unsigned m = __ballot_sync(0xffffffffu, cand);
      if (m) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(wv.ncand + fr, (unsigned)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (cand) {
          atomicOr(&tile_bm[row * 2 + (col >> 5)], 1u << (col & 31));
          const unsigned slot = base + __popc(m & ((1u << lane) - 1u));
          if (slot < (unsigned)wv.cand_cap) {
            wv.tmp_idx[(size_t)fr * wv.cand_cap + slot] = (unsigned)(Y0 + row) * (unsigned)w + (unsigned)(X0 + col);
            wv.tmp_key[(size_t)fr * wv.cand_cap + slot] = (unsigned long long)__double_as_longlong(0.125 * u);
          }
        }
      }
    }
    __syncthreads();
    if (tid < TH * 2) {
      const int y = Y0 + (tid >> 1), xw = X0 + 32 * (tid & 1);
      if (y < h && xw < w) wv.bitmap[(size_t)fr * wv.words_per_frame + (size_t)y * wv.wpr + (xw >> 5)] = tile_bm[tid];
    }
  }
}

// Exclusive prefix sum of popc(bitmap word) over one frame (one block per frame).
__global__ void __launch_bounds__(1024) bitmap_scan_kernel(CornerWorkView wv) {
  __shared__ unsigned wsum[32];
  const int fr = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t nwords = wv.words_per_frame;
  const unsigned* bm = wv.bitmap + (size_t)fr * nwords;
  unsigned* off = wv.wordoff + (size_t)fr * nwords;
  const size_t chunk = (nwords + 1023) / 1024;
  const size_t beg = (size_t)tid * chunk, end = beg + chunk < nwords ? beg + chunk : nwords;
  unsigned s = 0;
  for (size_t i = beg; i < end; i++) s += __popc(bm[i]);
  // block exclusive scan of s
  unsigned v = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  if (lane == 31) wsum[warp] = v;
  __syncthreads();
  if (warp == 0) {
    unsigned t = wsum[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned u = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += u;
    }
    wsum[lane] = t;
  }
  __syncthreads();
  unsigned run = v - s + (warp ? wsum[warp - 1] : 0u);
  for (size_t i = beg; i < end; i++) {
    off[i] = run;
    run += __popc(bm[i]);
  }
  if (tid == 1023) wv.ntotal[fr] = run;  // chunk boundaries are monotone: the last thread ends at the total
}

// Scatter the unordered candidate list into raster order.
__global__ void __launch_bounds__(256) order_kernel(CornerWorkView wv, int w) {
  const int fr = blockIdx.y;
  const unsigned n = min(wv.ncand[fr], (unsigned)wv.cand_cap);
  const size_t cb = (size_t)fr * wv.cand_cap, wb = (size_t)fr * wv.words_per_frame;
  for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const unsigned pix = wv.tmp_idx[cb + e];
    const unsigned y = pix / (unsigned)w, x = pix - y * (unsigned)w;
    const size_t word = wb + (size_t)y * wv.wpr + (x >> 5);
    const unsigned rank = wv.wordoff[word] + __popc(wv.bitmap[word] & ((1u << (x & 31)) - 1u));
    if (rank < (unsigned)wv.cand_cap) {
      wv.key[cb + rank] = wv.tmp_key[cb + e];
      wv.idx[cb + rank] = pix;
    }
  }
}

}  // namespace

// ---- work-area carving -------------------------------------------------------------------------------------
size_t sfm_corner_work_bytes(int w, int h, int nframes, int cand_cap) {
  CornerWorkView v;
  return corner_work_carve(v, nullptr, w, h, nframes, cand_cap);
}

// Runs max -> candidates -> scan -> order for frames [first, first+count); leaves raster-ordered (key, idx)
// and counts in the work area.
int sfm_corner_candidates_batch(sfmgpu_ctx* ctx, const sfmgpu_frames* f, int first, int count, double quality,
                                const CornerWorkView& wv) {
  SFM_CUDA(ctx, cudaMemsetAsync(wv.maxbits, 0, sizeof(unsigned long long) * count, ctx->stream));
  SFM_CUDA(ctx, cudaMemsetAsync(wv.ncand, 0, sizeof(unsigned) * count, ctx->stream));
  dim3 grid(sfm_cdiv(f->w, TW), sfm_cdiv(f->h, TH), count);
  SFM_LAUNCH(ctx, score_tile_kernel<0>, grid, 256, 0, f->lvl[0], f->w, f->h, f->pitch[0], f->fstride[0], first, wv, quality);
  SFM_LAUNCH(ctx, score_tile_kernel<1>, grid, 256, 0, f->lvl[0], f->w, f->h, f->pitch[0], f->fstride[0], first, wv, quality);
  SFM_LAUNCH(ctx, bitmap_scan_kernel, count, 1024, 0, wv);
  SFM_LAUNCH(ctx, order_kernel, dim3(16, count), 256, 0, wv, f->w);
  return 0;
}

int sfm_candidates_single(sfmgpu_ctx* ctx, const sfmgpu_frames* f, int frame, double quality, int32_t* xy, double* score,
                          int cap, int* n_out, double* max_score) {
  const int cand_cap = f->w * f->h;
  const size_t bytes = sfm_corner_work_bytes(f->w, f->h, 1, cand_cap);
  SFM_TRY(sfm_reserve(ctx, ctx->cs_work, bytes));
  CornerWorkView wv;
  corner_work_carve(wv, ctx->cs_work.p, f->w, f->h, 1, cand_cap);
  SFM_TRY(sfm_corner_candidates_batch(ctx, f, frame, 1, quality, wv));
  unsigned long long mb = 0;
  unsigned n = 0;
  SFM_CUDA(ctx, cudaMemcpyAsync(&mb, wv.maxbits, 8, cudaMemcpyDeviceToHost, ctx->stream));
  SFM_CUDA(ctx, cudaMemcpyAsync(&n, wv.ntotal, 4, cudaMemcpyDeviceToHost, ctx->stream));
  SFM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (max_score) {
    double u;
    memcpy(&u, &mb, 8);
    *max_score = 0.125 * u;
  }
  if (n_out) *n_out = (int)n;
  if ((int)n > cap) return sfm_fail(ctx, SFMGPU_E_CAPACITY, "corner_candidates: %u candidates, room for %d", n, cap);
  if (n == 0) return 0;
  std::vector<unsigned> idx(n);
  SFM_CUDA(ctx, cudaMemcpyAsync(idx.data(), wv.idx, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (score) SFM_CUDA(ctx, cudaMemcpyAsync(score, wv.key, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  SFM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (xy)
    for (unsigned i = 0; i < n; i++) {
      xy[2 * i] = (int32_t)(idx[i] % (unsigned)f->w);
      xy[2 * i + 1] = (int32_t)(idx[i] / (unsigned)f->w);
    }
  return 0;
}

extern "C" int sfmgpu_corner_candidates(sfmgpu_ctx* ctx, sfmgpu_frames* f, int frame, double quality, int32_t* xy,
                                        double* score, int cap, int* n_out, double* max_score) {
  if (!ctx || !f) return SFMGPU_E_ARG;
  if (frame < 0 || frame >= f->n || cap < 0) return sfm_fail(ctx, SFMGPU_E_ARG, "corner_candidates: bad frame/cap");
  return sfm_candidates_single(ctx, f, frame, quality, xy, score, cap, n_out, max_score);
}
