// Microbenchmark: per-SMSP issue rate of the instruction classes the matcher epilogues are built from (sm_100a).
// 16 independent dependency chains per thread, 16 warps per SM; results feed back into the chain so nothing is dead.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 2048
template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, long long* cyc, float seed) {
  float a[16]; uint32_t h[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { a[i] = -seed * (0.01f * i + 0.001f * threadIdx.x); h[i] = __float_as_uint(a[i]) | 0x00010001u; }
  __shared__ float4 sm[512];
  sm[threadIdx.x] = make_float4(seed, seed, seed, seed);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "+r"(h[i]) : "f"(a[i]), "f"(__uint_as_float(h[i])));
      if (MODE == 3) asm volatile("max.bf16x2 %0, %0, %1;" : "+r"(h[i]) : "r"(h[(i + 1) & 15]));
      if (MODE == 4) asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(h[i]) : "r"(h[(i + 1) & 15]));
      if (MODE == 5) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(h[i]) : "r"(h[(i + 1) & 15]), "r"(h[(i + 2) & 15]));
      if (MODE == 6) asm volatile("add.s32 %0, %0, %1;" : "+r"(h[i]) : "r"(h[(i + 1) & 15]));
      if (MODE == 7) asm volatile("shl.b32 %0, %0, 1;" : "+r"(h[i]));
      if (MODE == 8) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 1) & 15]));
      if (MODE == 9) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 10) { float4 v = sm[(threadIdx.x + (h[i] & 1)) & 511]; a[i] += v.x; }            // LDS.128 + FADD
      if (MODE == 11) { sm[threadIdx.x] = make_float4(a[i], a[i], a[i], a[i]); }                    // STS.128
      if (MODE == 12) asm volatile("fma.rn.bf16x2 %0, %0, %1, %1;" : "+r"(h[i]) : "r"(h[(i + 1) & 15]));
      if (MODE == 13) asm volatile("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "+r"(h[i]) : "f"(a[i]), "f"(__uint_as_float(h[i])));
      if (MODE == 14) asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(h[i]) : "r"(h[(i + 1) & 15]));
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + sm[(threadIdx.x + 1) & 511].x;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  k<MODE><<<148, 512>>>(out, cyc, 1.0f);
  k<MODE><<<148, 512>>>(out, cyc, 1.0f);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  double winstr = (double)ITER * 16 * 16;
  printf("%-34s warp-instr/cycle/SMSP %.3f  (cycles per warp-instr per SMSP %.2f)\n", name, winstr / c / 4, 4 * c / winstr);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("MUFU.EX2 f32"); run<1>("ex2.bf16x2 (2 MUFU + PRMT)"); run<9>("MUFU.RSQ f32"); run<2>("F2FP.BF16 pack"); run<13>("F2FP.RELU.BF16 pack");
  run<3>("HMNMX2.BF16"); run<12>("HFMA2.BF16"); run<4>("PRMT"); run<5>("LOP3"); run<6>("IADD3"); run<7>("SHL"); run<14>("IMAD"); run<8>("FMNMX");
  run<10>("LDS.128 + FADD"); run<11>("STS.128");
  return 0;
}
