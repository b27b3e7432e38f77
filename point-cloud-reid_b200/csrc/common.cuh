// Shared device helpers for the sm_100a kernels of pcreid-b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PCREID_OK 0
#define PCREID_ERR_ARG 1
#define PCREID_ERR_LAUNCH 2
#define PCREID_ERR_UNSUPPORTED 3

#define FULL_MASK 0xffffffffu

static inline int pcreid_launch_status() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? PCREID_OK : PCREID_ERR_LAUNCH;
}

__host__ __device__ static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Order preserving map float -> uint32 (ascending float == ascending unsigned). -0.0 is canonicalised to
// +0.0 first so that, as in torch.sort, the two compare equal.
__device__ __forceinline__ uint32_t f32_to_ordered(float f) {
  f = f + 0.0f;
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_f32(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return __uint_as_float(u);
}

__device__ __forceinline__ uint32_t warp_min_u32(uint32_t v) { return __reduce_min_sync(FULL_MASK, v); }
__device__ __forceinline__ uint32_t warp_max_u32(uint32_t v) { return __reduce_max_sync(FULL_MASK, v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL_MASK, v, o));
  return v;
}

// cp.async helpers (LDGSTS): 16-byte global -> shared copies without register staging.
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// streaming 128-bit global store (outputs that are written once and not re-read by this kernel)
__device__ __forceinline__ void st_cs_f4(float* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};\n" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}

// activation codes shared by host and device
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY02 = 2, ACT_ELU1 = 3 };

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_LEAKY02) return v > 0.f ? v : 0.2f * v;
  if (act == ACT_ELU1) return v > 0.f ? v + 1.f : expf(v);   // elu(v)+1 == exp(v) for v<=0
  return v;
}
