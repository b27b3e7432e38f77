// Microbenchmark: do tcgen05.mma shared-memory operand reads and LSU shared-memory accesses share one data pipe?
// One CTA per SM: warp 0 issues NMMA back-to-back SS-mode MMAs (M=128, N=128, K=16, bf16: 8 KB of operands per 64-cycle
// instruction = 128 B/clk), warps 1..8 run conflict-free LDS.128 loops.  Times: MMA alone, LDS alone, both together.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../point-cloud-reid_b200/csrc/tc_common.cuh"
#define NMMA 4096
#define NLDS 65536
__global__ void __launch_bounds__(288) k(int do_mma, int do_lds, int n_cols, float* out, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  if (warp == 0) { tc::tmem_alloc(&tbase, 256); tc::tmem_relinquish(); }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  long long t0 = clock64();
  float acc = 0.f;
  if (warp == 0) {
    if (do_mma) {
      const uint32_t idesc = tc::instr_desc(128, n_cols, tc::FMT_BF16, tc::MAJOR_K, tc::MAJOR_K);
      const uint64_t ad = tc::smem_desc(tc::smem_u32(smem), 2048, 128, tc::LAYOUT_NONE);
      const uint64_t bd = tc::smem_desc(tc::smem_u32(smem + 16384), n_cols * 16, 128, tc::LAYOUT_NONE);
      if (tc::elect_one()) {
        for (int i = 0; i < NMMA; ++i) tc::umma_f16(tbase, ad + (uint64_t)((i & 3) * 256), bd + (uint64_t)((i & 3) * (n_cols * 2)), idesc, 1u);
        tc::umma_commit(&bar);
      }
      __syncwarp();
      tc::mbar_wait(&bar, 0);
    }
  } else if (do_lds) {
    const uint32_t base = tc::smem_u32(smem + 32768) + ((warp - 1) * 128 + lane) * 16;
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
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __shared__ long long tm[9];
  if (lane == 0) tm[warp] = t1 - t0;
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) { cyc[blockIdx.x * 2] = tm[0]; long long m = 0; for (int w = 1; w < 9; ++w) m = tm[w] > m ? tm[w] : m; cyc[blockIdx.x * 2 + 1] = m; }
  if (warp == 0) tc::tmem_dealloc(tbase, 256);
}
void run(const char* name, int do_mma, int do_lds, int n_cols) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 288 * 4); cudaMalloc(&cyc, 148 * 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  k<<<148, 288, 65536>>>(do_mma, do_lds, n_cols, out, cyc);
  k<<<148, 288, 65536>>>(do_mma, do_lds, n_cols, out, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[296]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double cm = 0, cl = 0; for (int i = 0; i < 148; ++i) { cm += h[2 * i]; cl += h[2 * i + 1]; } cm /= 148; cl /= 148;
  const double mma_bytes = (128.0 * 16 * 2 + n_cols * 16.0 * 2) * NMMA, lds_bytes = 8.0 * NLDS * 512;
  printf("%-28s N=%3d  MMA warp %.0f cyc (%.1f cyc/MMA, %.0f B/clk)   LDS warps %.0f cyc (%.0f B/clk)   %s\n", name, n_cols, cm, cm / NMMA,
         do_mma ? mma_bytes / cm : 0.0, cl, do_lds ? lds_bytes / cl : 0.0, cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int n : {64, 128, 256}) { run("MMA alone", 1, 0, n); }
  run("LDS alone", 0, 1, 128);
  for (int n : {64, 128, 256}) { run("MMA + LDS", 1, 1, n); }
  return 0;
}
