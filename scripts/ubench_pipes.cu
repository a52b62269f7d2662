// ubench_pipes.cu — issue-rate probe for the instruction mixes the KLT kernel is built from (sm_100a).
// One block of 512 threads on one SM (4 warps per SM sub-partition); every warp runs ITER iterations of an
// unrolled body of 8 independent chains; rate = warp-instructions per cycle per sub-partition (clock64).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_pipes ubench_pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITER = 4096;

#define BODY8(stmt) stmt(0) stmt(1) stmt(2) stmt(3) stmt(4) stmt(5) stmt(6) stmt(7)

template <int MODE>
__global__ void probe(double* out, long long* cyc, int seed) {
  double d[8];
  int n[8];
  float f[8];
  for (int i = 0; i < 8; i++) {
    d[i] = 1.0 + 1e-9 * (threadIdx.x + i + seed);
    n[i] = threadIdx.x * 8 + i + seed;
    f[i] = 1.0f + 1e-3f * (threadIdx.x + i + seed);
  }
  const double m = 1.0 + 1e-12 * seed, c = 1e-13 * seed;
  const float mf = 1.0f + 1e-6f * seed, cf = 1e-7f * seed;
  const int hi = 0x43300000 + (seed & 0);
  unsigned long long q[8];
  for (int i = 0; i < 8; i++) {
    const float2 v = make_float2(f[i], f[i] * 1.5f);
    q[i] = *reinterpret_cast<const unsigned long long*>(&v);
  }
  const float2 mv = make_float2(mf, mf), cv = make_float2(cf, cf);
  const unsigned long long mq = *reinterpret_cast<const unsigned long long*>(&mv), cq = *reinterpret_cast<const unsigned long long*>(&cv);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; it++) {
    if (MODE == 0) {  // DFMA
#define S(i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(m), "d"(c));
      BODY8(S)
#undef S
    } else if (MODE == 1) {  // DADD
#define S(i) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(c));
      BODY8(S)
#undef S
    } else if (MODE == 2) {  // I2F.F64.S32 + LOP (int consumer)
#define S(i) { double t; asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(t) : "r"(n[i])); n[i] ^= __double2hiint(t); }
      BODY8(S)
#undef S
    } else if (MODE == 3) {  // I2F.F64.U16
#define S(i) { double t; unsigned short u = (unsigned short)n[i]; asm volatile("cvt.rn.f64.u16 %0, %1;" : "=d"(t) : "h"(u)); n[i] ^= __double2hiint(t); }
      BODY8(S)
#undef S
    } else if (MODE == 4) {  // DFMA + I2F.F64.S32 1:1 (shared pipe or not?)
#define S(i) { double t; asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(t) : "r"(n[i])); n[i] ^= __double2hiint(t); \
               asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(m), "d"(c)); }
      BODY8(S)
#undef S
    } else if (MODE == 5) {  // magic conversion: pair (lo = int, hi = 0x43300000) then DADD -2^52
#define S(i) { double t = __hiloint2double(hi, n[i] & 255); asm volatile("add.rn.f64 %0, %1, %2;" : "=d"(t) : "d"(t), "d"(-4503599627370496.0)); n[i] += __double2loint(t); }
      BODY8(S)
#undef S
    } else if (MODE == 6) {  // FFMA
#define S(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(mf), "f"(cf));
      BODY8(S)
#undef S
    } else if (MODE == 7) {  // F2F.F64.F32 + LOP
#define S(i) { double t; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(t) : "f"(f[i])); n[i] ^= __double2hiint(t); f[i] = __int_as_float((n[i] & 0x007fffff) | 0x3f800000); }
      BODY8(S)
#undef S
    } else if (MODE == 8) {  // DFMA + FFMA 1:1 (do FP32 instructions issue in the FP64 shadow?)
#define S(i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(m), "d"(c)); \
             asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(mf), "f"(cf));
      BODY8(S)
#undef S
    } else if (MODE == 9) {  // DFMA + 2 x FFMA
#define S(i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(m), "d"(c)); \
             asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(mf), "f"(cf)); \
             asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(cf), "f"(mf));
      BODY8(S)
#undef S
    } else if (MODE == 10) {  // I2F.F32.S32 + LOP
#define S(i) { float t; asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(t) : "r"(n[i])); n[i] ^= __float_as_int(t); }
      BODY8(S)
#undef S
    } else if (MODE == 12) {  // FFMA2 (fma.rn.f32x2: two FP32 FMAs per lane and instruction, sm_100+)
#define S(i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(q[i]) : "l"(mq), "l"(cq));
      BODY8(S)
#undef S
    } else if (MODE == 13) {  // FFMA2 + LOP3 1:1 (does the packed FMA free issue slots for the integer pipe?)
#define S(i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(q[i]) : "l"(mq), "l"(cq)); \
             asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(n[i]) : "r"(seed), "r"(hi));
      BODY8(S)
#undef S
    } else if (MODE == 14) {  // FFMA + LOP3 1:1
#define S(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(mf), "f"(cf)); \
             asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(n[i]) : "r"(seed), "r"(hi));
      BODY8(S)
#undef S
    } else if (MODE == 15) {  // FFMA2 + 2 LOP3
#define S(i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(q[i]) : "l"(mq), "l"(cq)); \
             asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(n[i]) : "r"(seed), "r"(hi)); \
             asm volatile("lop3.b32 %0, %0, %1, %2, 0x69;" : "+r"(n[i]) : "r"(hi), "r"(seed));
      BODY8(S)
#undef S
    } else if (MODE == 16) {  // FADD2
#define S(i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(q[i]) : "l"(cq));
      BODY8(S)
#undef S
    } else if (MODE == 11) {  // DMUL
#define S(i) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(m));
      BODY8(S)
#undef S
    }
  }
  const long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < 8; i++) s += d[i] + n[i] + f[i] + (double)(q[i] & 0xffff);
  out[threadIdx.x] = s;
  if ((threadIdx.x & 31) == 0) cyc[threadIdx.x >> 5] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_chain, double* out, long long* cyc) {
  probe<MODE><<<1, 512>>>(out, cyc, 1);
  cudaDeviceSynchronize();
  probe<MODE><<<1, 512>>>(out, cyc, 1);
  cudaDeviceSynchronize();
  long long h[16];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < 16; i++) mx = h[i] > mx ? h[i] : mx;
  const double inst = 4.0 * ITER * 8 * per_chain;  // warp-instructions of interest per sub-partition
  printf("%-40s %8.3f cycles per warp-instruction per sub-partition (%.3f inst/clk)\n", name, mx / inst, inst / mx);
}

int main() {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 512 * sizeof(double));
  cudaMalloc(&cyc, 16 * sizeof(long long));
  run<0>("DFMA", 1, out, cyc);
  run<1>("DADD", 1, out, cyc);
  run<11>("DMUL", 1, out, cyc);
  run<2>("I2F.F64.S32 (+LOP)", 1, out, cyc);
  run<3>("I2F.F64.U16 (+LOP)", 1, out, cyc);
  run<4>("DFMA + I2F.F64.S32 pair", 1, out, cyc);
  run<5>("magic cvt: DADD (+MOV,LOP,IADD)", 1, out, cyc);
  run<6>("FFMA", 1, out, cyc);
  run<7>("F2F.F64.F32 (+LOP x3)", 1, out, cyc);
  run<8>("DFMA + FFMA pair", 1, out, cyc);
  run<9>("DFMA + 2 FFMA triple", 1, out, cyc);
  run<10>("I2F.F32.S32 (+LOP)", 1, out, cyc);
  run<12>("FFMA2 (f32x2)", 1, out, cyc);
  run<16>("FADD2 (f32x2)", 1, out, cyc);
  run<14>("FFMA + LOP3 pair", 1, out, cyc);
  run<13>("FFMA2 + LOP3 pair", 1, out, cyc);
  run<15>("FFMA2 + 2 LOP3 triple", 1, out, cyc);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
