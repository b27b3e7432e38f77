// Microbenchmark: do redux.sync (f32 max / s32 add) and shfl.sync share the shared-memory data pipe with LDS traffic?
// One CTA per SM, 9 warps: warps 1..4 run a conflict-free LDS.128 stream (4 wavefronts per instruction: pipe-bound),
// warps 5..8 run NX iterations of 8 independent X operations (X = credux.max.f32 | redux.add.s32 | shfl.bfly | fadd as control).
// Times: LDS alone, X alone, both.  If "both" ~ max(LDS, X) the unit behind X does not sit on the LSU / shared-memory pipe.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define NLDS 32768
#define NX 16384
template <int X>
__global__ void __launch_bounds__(288) k(int do_lds, int do_x, float* out, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  long long t0 = clock64();
  float acc = 0.f;
  if (warp >= 1 && warp <= 4 && do_lds) {
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem) + ((warp - 1) * 128 + lane) * 16;
    float4 a0 = make_float4(0, 0, 0, 0), a1 = a0, a2 = a0, a3 = a0;
    for (int i = 0; i < NLDS; i += 4) {
      float4 v0, v1, v2, v3;
      const uint32_t ad = base ^ ((i & 4) << 10);
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v0.x), "=f"(v0.y), "=f"(v0.z), "=f"(v0.w) : "r"(ad));
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v1.x), "=f"(v1.y), "=f"(v1.z), "=f"(v1.w) : "r"(ad + 512));
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v2.x), "=f"(v2.y), "=f"(v2.z), "=f"(v2.w) : "r"(ad + 1024));
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v3.x), "=f"(v3.y), "=f"(v3.z), "=f"(v3.w) : "r"(ad + 1536));
      a0.x += v0.x; a1.y += v1.y; a2.z += v2.z; a3.w += v3.w;
    }
    acc = a0.x + a1.y + a2.z + a3.w;
  } else if (warp >= 5 && do_x) {
    float f[8]; int n[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { f[j] = (float)(lane * 8 + j) * 1e-3f; n[j] = lane + j; }
    for (int i = 0; i < NX; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (X == 0) asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(f[j]) : "f"(f[j] + 1e-3f));
        if (X == 1) asm volatile("redux.sync.add.s32 %0, %1, 0xffffffff;" : "=r"(n[j]) : "r"(n[j] & 0xff));
        if (X == 2) asm volatile("shfl.sync.bfly.b32 %0, %1, 16, 0x1f, 0xffffffff;" : "=f"(f[j]) : "f"(f[j] + 1e-3f));
        if (X == 3) asm volatile("add.f32 %0, %1, %2;" : "=f"(f[j]) : "f"(f[j]), "f"(1e-3f));
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += f[j] + (float)n[j];
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __shared__ long long tm[9];
  if (lane == 0) tm[warp] = t1 - t0;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long a = 0, b = 0;
    for (int w = 1; w <= 4; ++w) a = tm[w] > a ? tm[w] : a;
    for (int w = 5; w <= 8; ++w) b = tm[w] > b ? tm[w] : b;
    cyc[blockIdx.x * 2] = a; cyc[blockIdx.x * 2 + 1] = b;
  }
}
template <int X>
void run(const char* name, int do_lds, int do_x) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 288 * 4); cudaMalloc(&cyc, 148 * 16);
  cudaFuncSetAttribute(k<X>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  k<X><<<148, 288, 65536>>>(do_lds, do_x, out, cyc);
  k<X><<<148, 288, 65536>>>(do_lds, do_x, out, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[296]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double cl = 0, cx = 0; for (int i = 0; i < 148; ++i) { cl += h[2 * i]; cx += h[2 * i + 1]; } cl /= 148; cx /= 148;
  printf("%-34s LDS warps %8.0f cyc (%.2f cyc / LDS.128 / SM)   X warps %8.0f cyc (%.2f cyc / op / warp)   %s\n", name, cl,
         do_lds ? cl / (4.0 * NLDS) : 0.0, cx, do_x ? cx / (8.0 * NX) : 0.0, cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("LDS alone", 1, 0);
  run<0>("credux.max.f32 alone", 0, 1); run<0>("credux.max.f32 + LDS", 1, 1);
  run<1>("redux.add.s32 alone", 0, 1);  run<1>("redux.add.s32 + LDS", 1, 1);
  run<2>("shfl.bfly alone", 0, 1);      run<2>("shfl.bfly + LDS", 1, 1);
  run<3>("fadd (control) alone", 0, 1); run<3>("fadd (control) + LDS", 1, 1);
  return 0;
}
