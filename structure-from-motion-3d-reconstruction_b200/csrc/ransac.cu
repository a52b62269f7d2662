// ransac.cu — essential-matrix hypothesis scoring: Sampson error + exact inlier counts, winner and its mask.
//
// Replaces (reference cpp/src/templering_sfm.cpp): sampson_err :629-638 and the scoring loop of
// find_E_ransac :667-676 (inlier iff e < thr; winner = FIRST hypothesis with the strictly largest count;
// its inlier indices ascending).  Hypotheses are produced by the caller (the reference's own seeded sampling
// :657-665 + eight_point_E) so "same seeded hypotheses" holds by construction.
//
// Exactness: every count is bit-exact.  The Sampson numerator/denominator are evaluated in FP64 in the
// reference's operation order (this file is compiled with -fmad=false).  The reference's test
// fl(n2/den) < thr is decided without the division whenever n2 is outside a 2^-50 relative band around
// thr*den (proof in DESIGN.md §RANSAC); inside the band the literal division is executed.
//
// Mapping: a thread owns PTS correspondences in registers; a block walks a chunk of hypotheses whose nine
// coefficients are staged in shared memory; per-hypothesis block counts are reduced with ballots + shared
// atomics and added to counts[h] with one global atomic per (block, hypothesis).  Operand traffic is
// negligible (~4000 flop/B at 64k x 10k): the bound is the FP64 CUDA-core pipe, tensor cores are not used
// (no dense contraction).
#include <math.h>

#include <type_traits>

#include "common.cuh"

namespace {

// block size x hypotheses staged per block iteration: one instantiation for single correspondence sets (C4 shape), one for the
// batched stage (a few thousand points per pair: 128-thread blocks waste less of a pair's last block)
constexpr int RS_THREADS_1 = 256, RS_HCHUNK_1 = 128, RS_THREADS_B = 128, RS_HCHUNK_B = 128;
#ifndef RS_PTS_V
#define RS_PTS_V 2
#endif
constexpr int RS_PTS = RS_PTS_V;  // points per thread, packed two by two (FFMA2)
// (Measured, threads x chunk: RANSAC stage per 999 C2 pairs / per 399 C3 pairs / C4 rate: 256 x 64: 9.93 ms / 14.53 ms / 834 G;
//  128 x 128: 9.69 / 14.02 / 806; 128 x 64: 9.72 / 14.11 / 810; 256 x 128: 9.84 / 14.34 / 848; 256 x 32: 10.12 / 14.82 / 839;
//  512 x 64: 10.89 / 15.48 / 816 - within 4 % of each other: 128 x 128 for the batched stage, 256 x 128 for single sets.)

__device__ __forceinline__ float2 B2(float s) { return make_float2(s, s); }  // scalar broadcast operand of a packed instruction

// Reference-order Sampson pieces; returns n2 = num*num and den.
__device__ __forceinline__ void sampson_parts(const double* __restrict__ E, double x, double y, double xp, double yp, double& n2,
                                              double& den) {
  const double ex = (E[0] * x + E[1] * y) + E[2];  // E*[x y 1]; the *1.0 of the reference is exact
  const double ey = (E[3] * x + E[4] * y) + E[5];
  const double ez = (E[6] * x + E[7] * y) + E[8];
  const double tx = (E[0] * xp + E[3] * yp) + E[6];  // E^T*[xp yp 1], z row unused
  const double ty = (E[1] * xp + E[4] * yp) + E[7];
  const double num = (xp * ex + yp * ey) + ez;
  den = (((ex * ex + ey * ey) + tx * tx) + ty * ty) + 1e-12;
  n2 = num * num;
}

__device__ __forceinline__ bool is_inlier(double n2, double den, double thr, double thr_lo, double thr_hi) {
  if (n2 < thr_lo * den) return true;    // n2/den < thr(1-2^-51)  =>  fl(n2/den) < thr
  if (n2 > thr_hi * den) return false;   // n2/den > thr(1+2^-51)  =>  fl(n2/den) >= thr
  return (n2 / den) < thr;               // inside the band: the reference's literal test
}

// ---- FP32 screen -------------------------------------------------------------------------------------------------------
// Most pairs are far from the threshold: the same quantities are evaluated in FP32 (FMA) together with a rigorous bound of
// their distance to the EXACT real values, and only pairs the bound cannot decide run the FP64 code above.  With
// u = 2^-24, P = max(|x|,|y|,1), P' = max(|x'|,|y'|,1), Pm = max(P,P'), Esum = sum |E_ij|, Emax = the largest absolute
// row / column sum of E that enters ex, ey, ez, tx, ty:
//   |v_f - v| <= 6u Emax Pm               for v in {ex, ey, ez, tx, ty}   (inputs rounded to FP32, two FMAs: <= 4.1u S)
//   |num_f - num| <= dn = 32u Esum P' Pm  (three such terms scaled by |x'|, |y'|, 1 plus two FMA roundings: <= 21u ...)
//   |den_f - den| <= dd = 64u (Emax Pm)^2 + 8u den_f   (v^2 error <= dv(2|v| + dv), |v| <= Emax Pm; four FMA roundings)
// The FP64 reference's own value differs from the exact one by < 1e-13 relative to the same magnitudes, which the slack
// between 21u and 32u (resp. 48u and 64u) covers.  With N = |num_f|, D = den_f and eps = 2e-6:
//   e < thr   is certain when (N + dn)^2 < thr (1 - eps) (D - dd),   e >= thr   when max(N - dn, 0)^2 > thr (1 + eps) (D + dd).
// Both follow from ONE residual test (expand the squares; (N - dn)^2 >= N^2 - 2 N dn):
//   r = N^2 - thr D,     |r| > B,     B >= dn (2N + dn) + thr (1 + eps) 64u (Emax Pm)^2 + (eps + 8u (1 + eps)) thr D,
// the sign of r then being the answer.  The kernel evaluates r_f = fma(N, N, -td), td = fl(thr_f D), s_f = fma(N, N, td):
// |r_f - r| <= 3.1u s, thr D <= (1 + 2.1u) s, so with z >= Pm^2 >= P' Pm per point and, per hypothesis, c1 = 2 cn,
// c2 >= cn^2 (cn >= 32u Esum), cq >= thr (1 + eps) 64u Emax^2, kappa = eps + 16u (>= eps + 11.2u, the rest is slack for
// the <= 4 FP32 roundings of the bound's own arithmetic, as is the 1e-6 every constant is scaled up by):
//   B_f = fma(z, fma(c1, N, fma(c2, z, cq)), kappa s_f).
// NaNs, infinities and overflowing magnitudes make the comparison false and take the FP64 path; thresholds outside
// (1e-18, 1e18) disable the screen (every pair takes the FP64 path).  tests/test_ransac_bound.py emulates this arithmetic
// on the CPU against 80-bit values.
struct __align__(16) RsHyp {  // 48 bytes: read with three 128-bit shared-memory loads
  float e[9];
  float c1;  // 2 cn, cn = 32u * Esum (rounded up, floored at 1e-30: subnormal hypotheses)
  float c2;  // cn^2 (rounded up)
  float cq;  // thr (1 + eps) * 64u * Emax^2 (rounded up)
};

// blockIdx.z = correspondence set ("pair") of a batch: points at xi/xj + pair * pt_stride, npts[pair] of them (npts == nullptr:
// n_single), hypotheses E + pair * H * 9, counts + pair * H.  Sets with fewer than 8 points are not scored (:648).
template <int RS_THREADS, int RS_HCHUNK>
__global__ void __launch_bounds__(RS_THREADS) ransac_count_kernel(const double2* __restrict__ xi, const double2* __restrict__ xj,
                                                                 size_t pt_stride, const int* __restrict__ npts, int n_single,
                                                                 const double* __restrict__ E, int H, int h0, int h1,
                                                                 int h_per_block, double thr, double thr_lo, double thr_hi,
                                                                 float thr_f, float kappa_f, int* __restrict__ counts) {
  __shared__ double sE[RS_HCHUNK * 9];
  __shared__ RsHyp sH[RS_HCHUNK];
  // per-warp counts of the chunk, four hypotheses per word (a warp counts at most 2 x 32 per hypothesis: one byte each):
  // one warp reduction and one plain store per FOUR hypotheses, summed once per chunk
  __shared__ unsigned sC[RS_THREADS / 32][RS_HCHUNK / 4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pair = blockIdx.z;
  const int n = npts ? npts[pair] : n_single;
  if ((npts && n < 8) || (int)(blockIdx.x * RS_THREADS * RS_PTS) >= n) return;  // block-uniform
  xi += (size_t)pair * pt_stride;
  xj += (size_t)pair * pt_stride;
  E += (size_t)pair * H * 9;
  counts += (size_t)pair * H;
  const int p0 = (blockIdx.x * RS_THREADS + tid) * RS_PTS;
  double x[RS_PTS], y[RS_PTS], xp[RS_PTS], yp[RS_PTS];
  float xf[RS_PTS], yf[RS_PTS], xpf[RS_PTS], ypf[RS_PTS], zf[RS_PTS];
  unsigned vmask = 0;  // bit k: point k of this thread exists
#pragma unroll
  for (int k = 0; k < RS_PTS; k++) {
    const bool valid = p0 + k < n;
    vmask |= valid ? (1u << k) : 0u;
    const double2 a = valid ? xi[p0 + k] : make_double2(0, 0);
    const double2 b = valid ? xj[p0 + k] : make_double2(0, 0);
    x[k] = a.x;
    y[k] = a.y;
    xp[k] = b.x;
    yp[k] = b.y;
    xf[k] = (float)a.x;
    yf[k] = (float)a.y;
    xpf[k] = (float)b.x;
    ypf[k] = (float)b.y;
    const double Pm = fmax(fmax(fmax(fabs(a.x), fabs(a.y)), fmax(fabs(b.x), fabs(b.y))), 1.0);
    zf[k] = __double2float_ru(Pm * Pm * 1.000002);  // NaN / inf propagate and end in the FP64 path
  }
  static_assert(RS_PTS % 2 == 0, "the FP32 screen packs the thread's points two by two into float2");
  constexpr int NP = RS_PTS / 2;
  float2 X[NP], Y[NP], XP[NP], YP[NP], Z[NP];
#pragma unroll
  for (int g = 0; g < NP; g++) {
    X[g] = make_float2(xf[2 * g], xf[2 * g + 1]);
    Y[g] = make_float2(yf[2 * g], yf[2 * g + 1]);
    XP[g] = make_float2(xpf[2 * g], xpf[2 * g + 1]);
    YP[g] = make_float2(ypf[2 * g], ypf[2 * g + 1]);
    Z[g] = make_float2(zf[2 * g], zf[2 * g + 1]);
  }
  // only the last block of a set has threads without points: the common case counts without masking
  const bool tail = (int)((blockIdx.x + 1) * RS_THREADS * RS_PTS) > n;
  const int h_begin = h0 + blockIdx.y * h_per_block;  // hypotheses [h0, h1) of the set; E / counts keep their [H] layout
  const int h_end = min(h1, h_begin + h_per_block);
  for (int hc = h_begin; hc < h_end; hc += RS_HCHUNK) {
    const int nh = min(RS_HCHUNK, h_end - hc);
    __syncthreads();
    for (int i = tid; i < nh * 9; i += RS_THREADS) {
      const double v = E[(size_t)hc * 9 + i];
      sE[i] = v;
      sH[i / 9].e[i % 9] = (float)v;
    }
    __syncthreads();
    if (tid < nh) {
      const double* e = sE + tid * 9;
      double a[9];
#pragma unroll
      for (int i = 0; i < 9; i++) a[i] = fabs(e[i]);
      const double r0 = a[0] + a[1] + a[2], r1 = a[3] + a[4] + a[5], r2 = a[6] + a[7] + a[8];
      const double c0 = a[0] + a[3] + a[6], c1 = a[1] + a[4] + a[7];
      const double emax = fmax(fmax(fmax(r0, r1), fmax(r2, c0)), c1), esum = r0 + r1 + r2;
      const float cn = fmaxf(__double2float_ru(esum * (32.0 * 5.9604644775390625e-8 * 1.000001)), 1e-30f);
      sH[tid].c1 = cn + cn;
      sH[tid].c2 = __double2float_ru((double)cn * (double)cn * 1.000001);
      sH[tid].cq = __double2float_ru(thr * (1.0 + 2e-6) * (64.0 * 5.9604644775390625e-8 * 1.000001) * emax * emax);
    }
    __syncthreads();
    auto chunk = [&](auto tail_tag) {
      constexpr bool TAIL = decltype(tail_tag)::value;
      for (int h4 = 0; h4 < nh; h4 += 4) {  // (entries past nh hold stale hypotheses: computed, never read back)
        unsigned cpack = 0;
#pragma unroll
        for (int hk = 0; hk < 4; hk++) {
          const int h = h4 + hk;
          const float4 q0 = reinterpret_cast<const float4*>(&sH[h])[0], q1 = reinterpret_cast<const float4*>(&sH[h])[1],
                       q2 = reinterpret_cast<const float4*>(&sH[h])[2];
          RsHyp hy;
          hy.e[0] = q0.x; hy.e[1] = q0.y; hy.e[2] = q0.z; hy.e[3] = q0.w;
          hy.e[4] = q1.x; hy.e[5] = q1.y; hy.e[6] = q1.z; hy.e[7] = q1.w;
          hy.e[8] = q2.x; hy.c1 = q2.y; hy.c2 = q2.z; hy.cq = q2.w;
          // (Tried and dropped: the denominator chain and the residual pair as scalar FFMAs - an FFMA2 holds the FP32 pipe ~2.4
          // cycles against 2 x 1 for two FFMAs, and issue slots are free at 62 % - 838 vs 834 G hyp*pts/s on C4: noise.)
          // the thread's points two at a time: FFMA2 / FMUL2 / FADD2 (sm_100a packed FP32, one issue slot for two points'
          // worth of work; a scalar coefficient is a broadcast operand).  Every half is the correctly rounded FP32 operation
          // the bound derivation assumes.
          int c = 0;
          unsigned undecided = 0;
#pragma unroll
          for (int g = 0; g < NP; g++) {
            const float2 ex = __ffma2_rn(B2(hy.e[0]), X[g], __ffma2_rn(B2(hy.e[1]), Y[g], B2(hy.e[2])));
            const float2 ey = __ffma2_rn(B2(hy.e[3]), X[g], __ffma2_rn(B2(hy.e[4]), Y[g], B2(hy.e[5])));
            const float2 ez = __ffma2_rn(B2(hy.e[6]), X[g], __ffma2_rn(B2(hy.e[7]), Y[g], B2(hy.e[8])));
            const float2 tx = __ffma2_rn(B2(hy.e[0]), XP[g], __ffma2_rn(B2(hy.e[3]), YP[g], B2(hy.e[6])));
            const float2 ty = __ffma2_rn(B2(hy.e[1]), XP[g], __ffma2_rn(B2(hy.e[4]), YP[g], B2(hy.e[7])));
            const float2 nv = __ffma2_rn(XP[g], ex, __ffma2_rn(YP[g], ey, ez));
            const float2 den = __ffma2_rn(ty, ty, __ffma2_rn(tx, tx, __ffma2_rn(ey, ey, __ffma2_rn(ex, ex, B2(1e-12f)))));
            const float2 td = __fmul2_rn(B2(thr_f), den);
            const float2 r = __ffma2_rn(nv, nv, make_float2(-td.x, -td.y));
            const float2 sm = __ffma2_rn(nv, nv, td);
            const float2 i1 = __ffma2_rn(B2(hy.c2), Z[g], B2(hy.cq));
            const float2 i2 = __ffma2_rn(B2(hy.c1), make_float2(fabsf(nv.x), fabsf(nv.y)), i1);
            const float2 bb = __ffma2_rn(Z[g], i2, __fmul2_rn(B2(kappa_f), sm));
            const bool dec0 = fabsf(r.x) > bb.x, dec1 = fabsf(r.y) > bb.y;  // false for NaN / inf: the FP64 path decides
            const bool in0 = dec0 && r.x < 0.f, in1 = dec1 && r.y < 0.f;
            if (TAIL) {
              c += ((in0 && (vmask >> (2 * g)) & 1u) ? 1 : 0) + ((in1 && (vmask >> (2 * g + 1)) & 1u) ? 1 : 0);
              undecided |= ((!dec0 ? 1u : 0u) | (!dec1 ? 2u : 0u)) << (2 * g);
            } else {
              c += (in0 ? 1 : 0) + (in1 ? 1 : 0);
              undecided |= ((!dec0 ? 1u : 0u) | (!dec1 ? 2u : 0u)) << (2 * g);
            }
          }
          if (TAIL) undecided &= vmask;
          if (__any_sync(0xffffffffu, undecided != 0)) {  // rare: the reference's FP64 arithmetic decides
            const double* e = sE + h * 9;
#pragma unroll
            for (int k = 0; k < RS_PTS; k++)
              if (undecided & (1u << k)) {
                double n2, den;
                sampson_parts(e, x[k], y[k], xp[k], yp[k], n2, den);
                c += is_inlier(n2, den, thr, thr_lo, thr_hi) ? 1 : 0;
              }
          }
          cpack += (unsigned)c << (8 * hk);
        }
        cpack = __reduce_add_sync(0xffffffffu, cpack);
        if (lane == 0) sC[warp][h4 >> 2] = cpack;
      }
    };
    if (tail) chunk(std::true_type{});
    else chunk(std::false_type{});
    __syncthreads();
    if (tid < nh) {
      int t = 0;
#pragma unroll
      for (int wq = 0; wq < RS_THREADS / 32; wq++) t += (int)((sC[wq][tid >> 2] >> (8 * (tid & 3))) & 0xffu);
      if (t) atomicAdd(&counts[hc + tid], t);
    }
  }
}

// Winner: largest count, lowest index on ties; -1 when every count is 0 (best_inl starts empty, :673).
__global__ void __launch_bounds__(1024) ransac_argmax_kernel(const int* __restrict__ counts, int H, int* __restrict__ best) {
  __shared__ unsigned long long sbest[32];
  counts += (size_t)blockIdx.x * H;  // one block per correspondence set
  best += 2 * blockIdx.x;
  // key = count << 32 | (0xFFFFFFFF - h): max key = max count then min h
  unsigned long long key = 0;
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    const unsigned long long k = ((unsigned long long)(unsigned)counts[h] << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)h);
    key = k > key ? k : key;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
    key = other > key ? other : key;
  }
  if ((threadIdx.x & 31) == 0) sbest[threadIdx.x >> 5] = key;
  __syncthreads();
  if (threadIdx.x < 32) {
    key = threadIdx.x < (blockDim.x >> 5) ? sbest[threadIdx.x] : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
      key = other > key ? other : key;
    }
    if (threadIdx.x == 0) {
      const int cnt = (int)(key >> 32);
      best[0] = cnt > 0 ? (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull)) : -1;
      best[1] = cnt;
    }
  }
}

// Inlier indices of the winner, ascending (:669-672), by one block with an ordered ballot compaction.
__global__ void __launch_bounds__(1024) ransac_mask_kernel(const double2* __restrict__ xi, const double2* __restrict__ xj,
                                                          size_t pt_stride, const int* __restrict__ npts, int n_single,
                                                          const double* __restrict__ E, int H, int* __restrict__ best, double thr,
                                                          int* __restrict__ inl, int recount) {
  __shared__ int swarp[32];
  __shared__ int sbase;
  const int pair = blockIdx.x;  // one block per correspondence set
  const int n = npts ? npts[pair] : n_single;
  xi += (size_t)pair * pt_stride;
  xj += (size_t)pair * pt_stride;
  E += (size_t)pair * H * 9;
  inl += (size_t)pair * pt_stride;
  best += 2 * pair;
  const int bh = best[0];
  if (bh < 0 || (npts && n < 8)) return;
  const double* e = E + (size_t)bh * 9;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) sbase = 0;
  __syncthreads();
  for (int start = 0; start < n; start += blockDim.x) {
    const int i = start + tid;
    bool in = false;
    if (i < n) {
      double n2, den;
      const double2 a = xi[i], b = xj[i];
      sampson_parts(e, a.x, a.y, b.x, b.y, n2, den);
      in = (n2 / den) < thr;  // literal
    }
    const unsigned m = __ballot_sync(0xffffffffu, in);
    if (lane == 0) swarp[warp] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
    for (int wq = 0; wq < (int)(blockDim.x >> 5); wq++) {
      const int c = swarp[wq];
      before += wq < warp ? c : 0;
      total += c;
    }
    const int base = sbase;
    if (in) inl[base + before + __popc(m & ((1u << lane) - 1u))] = i;
    __syncthreads();
    if (tid == 0) sbase = base + total;
    __syncthreads();
  }
  // the winner was re-solved after the counting pass (screening solver): its count is that of the hypothesis listed here
  if (recount && tid == 0) best[1] = sbase;
}

}  // namespace

// Scoring loop (:667-676) for `npairs` correspondence sets in one launch set: counts [npairs][H], best [npairs][2] =
// (winner or -1, its count), inl [npairs][pt_stride] the winner's inlier indices ascending.  n_max bounds the points of a
// set (grid size); npts (device, may be null: n_max points in every set) holds the actual numbers.
// refine_idx8 != nullptr: E holds SCREENING hypotheses (solver.cu, eight_point_qr_kernel) of the octets refine_idx8
// [npairs][H][8]: they decide the winner; the winner is then solved again by the Jacobi emulation (written over its slot
// of E) and its inlier list and count (best[1]) are those of that hypothesis.
int sfm_ransac_score_batched(sfmgpu_ctx* ctx, const double2* xi, const double2* xj, size_t pt_stride, const int* npts, int n_max,
                             int npairs, double* E, int H, double thr, int* counts, int* best, int* inl, const int* refine_idx8) {
  if (npairs <= 0) return 0;
  SFM_CUDA(ctx, cudaMemsetAsync(counts, 0, (size_t)npairs * (H > 0 ? H : 1) * sizeof(int), ctx->stream));
  SFM_TRY(sfm_ransac_count_range(ctx, xi, xj, pt_stride, npts, n_max, npairs, E, H, 0, H, thr, counts));
  return sfm_ransac_finish(ctx, xi, xj, pt_stride, npts, n_max, npairs, E, H, thr, counts, best, inl, refine_idx8);
}

// Counts of the hypotheses [h0, h1) of every set, ADDED to counts [npairs][H] (the caller zeroes them); sets with
// npts[pair] < 8 are skipped.
int sfm_ransac_count_range(sfmgpu_ctx* ctx, const double2* xi, const double2* xj, size_t pt_stride, const int* npts, int n_max,
                           int npairs, const double* E, int H, int h0, int h1, double thr, int* counts) {
  const int n = n_max, Hr = h1 - h0;
  if (npairs <= 0 || H <= 0 || n <= 0 || Hr <= 0) return 0;
  if (h0 < 0 || h1 > H) return sfm_fail(ctx, SFMGPU_E_ARG, "ransac: hypothesis range [%d, %d) outside [0, %d)", h0, h1, H);
  const bool batched = npairs > 1;
  const int threads = batched ? RS_THREADS_B : RS_THREADS_1, hchunk = batched ? RS_HCHUNK_B : RS_HCHUNK_1;
  const unsigned gx = sfm_cdiv(n, threads * RS_PTS);
  // enough blocks to fill the machine several times over, whole chunks of hypotheses per block
  long long want_y = ((long long)ctx->n_sm * 8 + (long long)gx * npairs - 1) / ((long long)gx * npairs);
  long long hpb = (Hr + want_y - 1) / want_y;
  hpb = ((hpb + hchunk - 1) / hchunk) * hchunk;
  const unsigned gy = sfm_cdiv(Hr, hpb);
  // thr_lo / thr_hi bracket thr by 2^-50 relative (see is_inlier)
  const double eps = 8.8817841970012523e-16;  // 2^-50
  // FP32 screen (see RsHyp): thr as float, kappa = (eps + 16u) with slack.  A threshold the screen cannot handle
  // (<= 0, non-finite, outside (1e-18, 1e18)) disables it: td = NaN leaves every pair undecided, i.e. to the FP64 path.
  float thr_f = (float)thr;
  const float kappa_f = (float)((2e-6 + 16.0 * 5.9604644775390625e-8) * 1.00001);
  if (!(thr > 1e-18 && thr < 1e18)) thr_f = NAN;
  for (int z0 = 0; z0 < npairs; z0 += 65535) {  // grid.z limit
    const int nz = npairs - z0 < 65535 ? npairs - z0 : 65535;
    if (batched)
      SFM_LAUNCH(ctx, (ransac_count_kernel<RS_THREADS_B, RS_HCHUNK_B>), dim3(gx, gy, nz), RS_THREADS_B, 0, xi + (size_t)z0 * pt_stride,
                 xj + (size_t)z0 * pt_stride, pt_stride, npts ? npts + z0 : nullptr, n, E + (size_t)z0 * H * 9, H, h0, h1, (int)hpb, thr,
                 thr * (1.0 - eps), thr * (1.0 + eps), thr_f, kappa_f, counts + (size_t)z0 * H);
    else
      SFM_LAUNCH(ctx, (ransac_count_kernel<RS_THREADS_1, RS_HCHUNK_1>), dim3(gx, gy, nz), RS_THREADS_1, 0, xi + (size_t)z0 * pt_stride,
                 xj + (size_t)z0 * pt_stride, pt_stride, npts ? npts + z0 : nullptr, n, E + (size_t)z0 * H * 9, H, h0, h1, (int)hpb, thr,
                 thr * (1.0 - eps), thr * (1.0 + eps), thr_f, kappa_f, counts + (size_t)z0 * H);
  }
  return 0;
}

// Winner of every set from counts [npairs][H] (largest count, lowest index), the screening winner re-solved (refine_idx8),
// its inlier list.
int sfm_ransac_finish(sfmgpu_ctx* ctx, const double2* xi, const double2* xj, size_t pt_stride, const int* npts, int n_max, int npairs,
                      double* E, int H, double thr, const int* counts, int* best, int* inl, const int* refine_idx8) {
  const int n = n_max;
  if (npairs <= 0) return 0;
  SFM_LAUNCH(ctx, ransac_argmax_kernel, npairs, 1024, 0, counts, H, best);
  if (H > 0 && n > 0) {
    if (refine_idx8) SFM_TRY(sfm_eight_point_winners(ctx, xi, xj, pt_stride, npts, n, npairs, refine_idx8, H, (const int*)best, E));
    SFM_LAUNCH(ctx, ransac_mask_kernel, npairs, 1024, 0, xi, xj, pt_stride, npts, n, (const double*)E, H, best, thr, inl,
               refine_idx8 ? 1 : 0);
  }
  return 0;
}

namespace {

int score_resident(sfmgpu_ctx* ctx, double thr) {
  return sfm_ransac_score_batched(ctx, (const double2*)ctx->rs_xi.p, (const double2*)ctx->rs_xj.p, 0, nullptr, ctx->rs_n, 1,
                                  (double*)ctx->rs_E.p, ctx->rs_H, thr, (int*)ctx->rs_counts.p, (int*)ctx->rs_best.p,
                                  (int*)ctx->rs_inl.p, ctx->rs_screened ? (const int*)ctx->rs_idx8.p : nullptr);
}

}  // namespace

extern "C" {

int sfmgpu_ransac_upload(sfmgpu_ctx* ctx, const double* xi_xy, const double* xj_xy, int n, const double* E, int H) {
  SFM_ENTER(ctx);
  if (!ctx || n < 0 || H < 0) return sfm_fail(ctx, SFMGPU_E_ARG, "ransac_upload: bad sizes");
  if ((n > 0 && (!xi_xy || !xj_xy)) || (H > 0 && !E)) return sfm_fail(ctx, SFMGPU_E_ARG, "ransac_upload: null pointer");
  SFM_TRY(sfm_reserve(ctx, ctx->rs_xi, (size_t)(n + 1) * 16));
  SFM_TRY(sfm_reserve(ctx, ctx->rs_xj, (size_t)(n + 1) * 16));
  SFM_TRY(sfm_reserve(ctx, ctx->rs_E, (size_t)(H + 1) * 72));
  SFM_TRY(sfm_reserve(ctx, ctx->rs_counts, (size_t)(H + 1) * 4));
  SFM_TRY(sfm_reserve(ctx, ctx->rs_inl, (size_t)(n + 1) * 4));
  SFM_TRY(sfm_reserve(ctx, ctx->rs_best, 16));
  if (n > 0) {
    SFM_CUDA(ctx, cudaMemcpyAsync(ctx->rs_xi.p, xi_xy, (size_t)n * 16, cudaMemcpyHostToDevice, ctx->stream));
    SFM_CUDA(ctx, cudaMemcpyAsync(ctx->rs_xj.p, xj_xy, (size_t)n * 16, cudaMemcpyHostToDevice, ctx->stream));
  }
  if (H > 0) SFM_CUDA(ctx, cudaMemcpyAsync(ctx->rs_E.p, E, (size_t)H * 72, cudaMemcpyHostToDevice, ctx->stream));
  ctx->rs_n = n;
  ctx->rs_H = H;
  ctx->rs_screened = false;
  return 0;
}

int sfmgpu_ransac_score_resident(sfmgpu_ctx* ctx, double thr, int* best_h, int* best_n) {
  SFM_ENTER(ctx);
  if (!ctx) return SFMGPU_E_ARG;
  if (!ctx->rs_best.p) return sfm_fail(ctx, SFMGPU_E_STATE, "ransac_score_resident: nothing uploaded");
  SFM_TRY(score_resident(ctx, thr));
  if (best_h || best_n) {
    int hb[2];
    SFM_CUDA(ctx, cudaMemcpyAsync(hb, ctx->rs_best.p, sizeof hb, cudaMemcpyDeviceToHost, ctx->stream));
    SFM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (best_h) *best_h = hb[0];
    if (best_n) *best_n = hb[1];
  }
  return 0;
}

int sfmgpu_ransac_download(sfmgpu_ctx* ctx, int32_t* counts, int32_t* best_inl, int cap_inl) {
  SFM_ENTER(ctx);
  if (!ctx) return SFMGPU_E_ARG;
  if (!ctx->rs_best.p) return sfm_fail(ctx, SFMGPU_E_STATE, "ransac_download: nothing scored");
  int hb[2];
  SFM_CUDA(ctx, cudaMemcpyAsync(hb, ctx->rs_best.p, sizeof hb, cudaMemcpyDeviceToHost, ctx->stream));
  if (counts && ctx->rs_H > 0)
    SFM_CUDA(ctx, cudaMemcpyAsync(counts, ctx->rs_counts.p, (size_t)ctx->rs_H * 4, cudaMemcpyDeviceToHost, ctx->stream));
  SFM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (best_inl && hb[0] >= 0) {
    if (cap_inl < hb[1]) return sfm_fail(ctx, SFMGPU_E_CAPACITY, "ransac_download: %d inliers, room for %d", hb[1], cap_inl);
    SFM_CUDA(ctx, cudaMemcpyAsync(best_inl, ctx->rs_inl.p, (size_t)hb[1] * 4, cudaMemcpyDeviceToHost, ctx->stream));
    SFM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return 0;
}

int sfmgpu_ransac_score(sfmgpu_ctx* ctx, const double* xi_xy, const double* xj_xy, int n, const double* E, int H, double thr,
                        int32_t* counts, int* best_h, int32_t* best_inl, int* best_n) {
  SFM_ENTER(ctx);
  if (!ctx) return SFMGPU_E_ARG;
  SFM_TRY(sfmgpu_ransac_upload(ctx, xi_xy, xj_xy, n, E, H));
  int bh = -1, bn = 0;
  SFM_TRY(sfmgpu_ransac_score_resident(ctx, thr, &bh, &bn));
  if (best_h) *best_h = bh;
  if (best_n) *best_n = bn;
  return sfmgpu_ransac_download(ctx, counts, best_inl, n);
}

}  // extern "C"
