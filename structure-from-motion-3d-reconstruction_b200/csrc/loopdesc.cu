// loopdesc.cu — loop-closure descriptor and candidate search.
//
// Replaces (reference cpp/src/templering_sfm.cpp) global_desc_32 :1100-1122 (repeated downsample2 until both sides
// are <= 32, nearest sampling to 32x32, mean removal, L2 normalisation), dot_desc :1124-1129 and the candidate loop
// :1823-1831 (first strictly larger score wins, best_score starts at 0).
//
// Bit-exact: the box filter is integer; the mean is a sum of integers (exact in double in any order); the squared
// norm IS order dependent, so one thread accumulates it in raster order like the reference; the dot products are
// sequential float multiply-add pairs WITHOUT contraction (__fmul_rn / __fadd_rn), one thread per stored descriptor.
#include "common.cuh"

namespace {

// dst(x, y) = (s00 + s10 + s01 + s11) / 4, floor(w/2) x floor(h/2), tight rows; grid.z = image
__global__ void __launch_bounds__(256) halve_kernel(const uint8_t* __restrict__ src, int sw, int sh, size_t spitch, size_t sstride,
                                                   uint8_t* __restrict__ dst, size_t dstride) {
  const int ow = sw / 2, oh = sh / 2;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= ow || y >= oh) return;
  const uint8_t* r0 = src + (size_t)blockIdx.z * sstride + (size_t)(2 * y) * spitch + 2 * x;
  const uint8_t* r1 = r0 + spitch;
  dst[(size_t)blockIdx.z * dstride + (size_t)y * ow + x] = (uint8_t)(((int)r0[0] + r0[1] + r1[0] + r1[1]) / 4);
}

// One block of 1024 threads per image: nearest sampling, mean, centring, norm, scaling.
__global__ void __launch_bounds__(1024) desc_finish_kernel(const uint8_t* __restrict__ src, int cw, int ch, size_t spitch, size_t sstride,
                                                          float* __restrict__ out) {
  __shared__ float v[1024];
  __shared__ double s_mean, s_invn;
  const int t = threadIdx.x, x = t & 31, y = t >> 5;
  const uint8_t* im = src + (size_t)blockIdx.x * sstride;
  int sx = (int)round((double)x * (double)(cw - 1) / 31.0), sy = (int)round((double)y * (double)(ch - 1) / 31.0);
  sx = sx < cw - 1 ? sx : cw - 1;
  sy = sy < ch - 1 ? sy : ch - 1;
  v[t] = (float)im[(size_t)sy * spitch + sx];
  __syncthreads();
  if (t == 0) {
    double m = 0.0;
    for (int i = 0; i < 1024; i++) m += v[i];  // integers: exact
    s_mean = m / (32.0 * 32.0);
  }
  __syncthreads();
  const float c = __fsub_rn(v[t], (float)s_mean);
  __syncthreads();
  v[t] = c;
  __syncthreads();
  if (t == 0) {
    double n2 = 0.0;
    for (int i = 0; i < 1024; i++) n2 = __dadd_rn(n2, __dmul_rn((double)v[i], (double)v[i]));  // raster order (:1118)
    s_invn = 1.0 / sqrt(n2 + 1e-12);
  }
  __syncthreads();
  out[(size_t)blockIdx.x * 1024 + t] = (float)__dmul_rn((double)c, s_invn);
}

__global__ void __launch_bounds__(128) desc_dot_kernel(const float* __restrict__ descs, int n, const float* __restrict__ query,
                                                      float* __restrict__ scores) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const float* a = descs + (size_t)k * 1024;
  float s = 0.0f;
  for (int i = 0; i < 1024; i++) s = __fadd_rn(s, __fmul_rn(a[i], query[i]));
  scores[k] = s;
}

}  // namespace

extern "C" {

int sfmgpu_global_desc32(sfmgpu_ctx* ctx, sfmgpu_frames* f, int first, int count, float* desc_out) {
  SFM_ENTER(ctx);
  if (!ctx || !f) return SFMGPU_E_ARG;
  if (first < 0 || count < 0 || first + count > f->n) return sfm_fail(ctx, SFMGPU_E_ARG, "global_desc32: bad frame range");
  if (count == 0) return 0;
  if (!desc_out) return sfm_fail(ctx, SFMGPU_E_ARG, "global_desc32: null output");
  // ping-pong scratch for the halving chain (tight rows) + the descriptors
  const size_t half = (size_t)(f->w / 2) * (f->h / 2), quarter = (size_t)(f->w / 4) * (f->h / 4);
  const size_t bytes = (half + quarter + 64) * count + (size_t)count * 4096 + 512;
  SFM_TRY(sfm_reserve(ctx, ctx->cs_work, bytes));
  uint8_t* buf[2] = {(uint8_t*)ctx->cs_work.p, (uint8_t*)ctx->cs_work.p + (half + 32) * count};
  float* d_desc = (float*)(((uintptr_t)(buf[1] + (quarter + 32) * count) + 255) & ~(uintptr_t)255);
  const uint8_t* src = f->lvl[0] + (size_t)first * f->fstride[0];
  int cw = f->w, ch = f->h;
  size_t spitch = f->pitch[0], sstride = f->fstride[0];
  int which = 0;
  while (cw > 32 || ch > 32) {
    const int ow = cw / 2, oh = ch / 2;
    if (ow < 1 || oh < 1) return sfm_fail(ctx, SFMGPU_E_ARG, "global_desc32: image degenerates to %dx%d", ow, oh);
    const size_t dstride = (size_t)ow * oh;
    SFM_LAUNCH(ctx, halve_kernel, dim3(sfm_cdiv(ow, 32), sfm_cdiv(oh, 8), count), 256, 0, src, cw, ch, spitch, sstride, buf[which],
               dstride);
    src = buf[which];
    which ^= 1;
    cw = ow;
    ch = oh;
    spitch = (size_t)cw;
    sstride = dstride;
  }
  SFM_LAUNCH(ctx, desc_finish_kernel, count, 1024, 0, src, cw, ch, spitch, sstride, d_desc);
  SFM_CUDA(ctx, cudaMemcpyAsync(desc_out, d_desc, (size_t)count * 4096, cudaMemcpyDeviceToHost, ctx->stream));
  SFM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

int sfmgpu_desc_search(sfmgpu_ctx* ctx, const float* descs, int n_search, const float* query, float* scores, int* best_id,
                       float* best_score) {
  SFM_ENTER(ctx);
  if (!ctx || n_search < 0 || !query || (n_search > 0 && !descs)) return sfm_fail(ctx, SFMGPU_E_ARG, "desc_search: bad arguments");
  int bid = -1;
  float bs = 0.0f;
  if (n_search > 0) {
    SFM_TRY(sfm_reserve(ctx, ctx->cs_work, (size_t)(n_search + 1) * 4096 + (size_t)n_search * 4 + 512));
    float* d_descs = (float*)ctx->cs_work.p;
    float* d_query = d_descs + (size_t)n_search * 1024;
    float* d_scores = d_query + 1024;
    SFM_CUDA(ctx, cudaMemcpyAsync(d_descs, descs, (size_t)n_search * 4096, cudaMemcpyHostToDevice, ctx->stream));
    SFM_CUDA(ctx, cudaMemcpyAsync(d_query, query, 4096, cudaMemcpyHostToDevice, ctx->stream));
    SFM_LAUNCH(ctx, desc_dot_kernel, sfm_cdiv(n_search, 128), 128, 0, d_descs, n_search, d_query, d_scores);
    SFM_TRY(sfm_pinned(ctx, (size_t)n_search * 4));
    float* h = (float*)ctx->pinned;
    SFM_CUDA(ctx, cudaMemcpyAsync(h, d_scores, (size_t)n_search * 4, cudaMemcpyDeviceToHost, ctx->stream));
    SFM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < n_search; k++) {  // :1827-1830: first strictly larger score wins
      if (scores) scores[k] = h[k];
      if (h[k] > bs) {
        bs = h[k];
        bid = k;
      }
    }
  }
  if (best_id) *best_id = bid;
  if (best_score) *best_score = bs;
  return 0;
}

}  // extern "C"
