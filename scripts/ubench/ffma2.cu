// Microbenchmark: issue throughput of FFMA vs FFMA2 (fma.rn.f32x2), FADD2, HFMA2.BF16, F2FP pack, FMNMX on sm_100a.
// Each variant runs ITER x 32 independent-chain instructions per thread; reports warp-instructions / cycle / SM.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
#define ITER 4096
template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, long long* cyc, float seed) {
  float a[16];
  unsigned long long p[16];
  uint32_t h[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { a[i] = seed + i + threadIdx.x; p[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(seed); h[i] = __float_as_uint(a[i]); }
  const float m = seed * 0.5f, c = seed * 0.25f;
  const unsigned long long m2 = ((unsigned long long)__float_as_uint(m) << 32) | __float_as_uint(m), c2 = ((unsigned long long)__float_as_uint(c) << 32) | __float_as_uint(c);
  const uint32_t mh = __float_as_uint(m);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) a[i] = fmaf(a[i], m, c);
      if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(m2), "l"(c2));
      if (MODE == 2) asm volatile("add.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(c2));
      if (MODE == 3) asm volatile("fma.rn.bf16x2 %0, %0, %1, %1;" : "+r"(h[i]) : "r"(mh));
      if (MODE == 4) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(a[i]), "f"(a[(i + 1) & 15]));
      if (MODE == 5) a[i] = fmaxf(a[i], m);
      if (MODE == 6) { a[i] = fmaf(a[i], m, c); h[i] = max(h[i], mh) ^ 0x55u; }   // fma + alu mix
      if (MODE == 7) asm volatile("mul.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(m2));
      if (MODE == 8) a[i] = a[i] + c;
      if (MODE == 9) asm volatile("redux.sync.max.f32 %0, %0, 0xffffffff;" : "+f"(a[i]));
      if (MODE == 10) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 16);
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i] + __uint_as_float((uint32_t)p[i]) + __uint_as_float((uint32_t)(p[i] >> 32)) + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, int per_iter) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  k<MODE><<<148, 512>>>(out, cyc, 1.0f);
  k<MODE><<<148, 512>>>(out, cyc, 1.0f);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  double winstr = (double)ITER * 16 * per_iter * 16;   // 16 warps per CTA
  printf("%-28s cycles %.0f  warp-instr/cycle/SM %.3f  (per SMSP %.3f)\n", name, c, winstr / c, winstr / c / 4);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("FFMA", 1); run<1>("FFMA2 (fma.rn.f32x2)", 1); run<2>("FADD2", 1); run<7>("FMUL2", 1); run<8>("FADD", 1);
  run<3>("HFMA2.BF16", 1); run<4>("F2FP.BF16 pack", 1); run<5>("FMNMX", 1); run<6>("FFMA + 2 ALU mix", 3); run<9>("CREDUX.MAX.F32", 1); run<10>("SHFL.BFLY", 1);
  return 0;
}
