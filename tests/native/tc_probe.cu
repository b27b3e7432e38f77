// tcgen05 probe: one 128 x N x K GEMM per launch in each operand configuration the fused kernels use, so
// that descriptor encodings / TMEM layouts are validated on hardware against a plain fp32 reference
// (tests/test_gpu_tc.py) before they are relied on.
//   mode 0: bf16, A and B K-major in shared memory   (no-swizzle canonical layout [k/8][row][8])
//   mode 1: bf16, A and B MN-major in shared memory  (same physical image, roles of row / k swapped)
//   mode 2: tf32, A and B K-major in shared memory   ([k/4][row][4] fp32)
//   mode 3: bf16, A from tensor memory (tcgen05.st by the row-owning threads), B K-major in shared memory
//   mode 4: tf32, A and B MN-major in shared memory  ([mn/4][k][4] fp32; global holds A^T (K x 128), B^T (K x N))
//   mode 5: tf32, A from tensor memory (one fp32 element per 32-bit column), B K-major in shared memory
//   modes 6 / 7 / 8: modes 0 / 1 / 3 with fp16 instead of bf16 operands (the parity_tc mode of the fused matcher)
// Test-only library (tests/native/libpcreid_tcprobe.so, built by __graft_entry__.build()); not part of the product .so.
// D (128 x N, fp32, row-major) = A (128 x K) * B (N x K)^T.
#include "../../include/pcreid.h"
#include "../../point-cloud-reid_b200/csrc/common.cuh"
#include "../../point-cloud-reid_b200/csrc/tc_common.cuh"

namespace {

__global__ void __launch_bounds__(128) tc_probe_kernel(int mode, int N, int K, const void* __restrict__ Ag,
                                                       const void* __restrict__ Bg, float* __restrict__ D) {
  const int fmt16 = mode >= 6 ? tc::FMT_F16 : tc::FMT_BF16;
  if (mode >= 6) mode = mode == 6 ? 0 : (mode == 7 ? 1 : 3);
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int esz = (mode == 2 || mode == 4 || mode == 5) ? 4 : 2;   // bytes per element
  const int cpe = 16 / esz;                     // elements per 16-byte chunk
  uint8_t* As = smem;
  uint8_t* Bs = smem + (size_t)128 * K * esz;

  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) {
    tc::tmem_alloc(&tmem_base_s, 512);
    tc::tmem_relinquish();
  }
  // ---- stage operands in the canonical no-swizzle layouts
  if (mode == 0 || mode == 2 || mode == 3 || mode == 5) {
    // K-major: element (r, k) at (k/cpe)*(R*16) + r*16 + (k%cpe)*esz     (LBO = R*16, SBO = 128)
    if (mode != 3 && mode != 5)
      for (int i = tid; i < 128 * K; i += 128) {
        int r = i / K, k = i % K;
        size_t off = (size_t)(k / cpe) * (128 * 16) + (size_t)r * 16 + (size_t)(k % cpe) * esz;
        if (esz == 2) *reinterpret_cast<uint16_t*>(As + off) = reinterpret_cast<const uint16_t*>(Ag)[i];
        else *reinterpret_cast<uint32_t*>(As + off) = reinterpret_cast<const uint32_t*>(Ag)[i];
      }
    for (int i = tid; i < N * K; i += 128) {
      int r = i / K, k = i % K;
      size_t off = (size_t)(k / cpe) * ((size_t)N * 16) + (size_t)r * 16 + (size_t)(k % cpe) * esz;
      if (esz == 2) *reinterpret_cast<uint16_t*>(Bs + off) = reinterpret_cast<const uint16_t*>(Bg)[i];
      else *reinterpret_cast<uint32_t*>(Bs + off) = reinterpret_cast<const uint32_t*>(Bg)[i];
    }
  } else if (mode == 4) {
    // tf32 MN-major: element (mn, k) at (mn/4)*(K*16) + k*16 + (mn%4)*4
    for (int i = tid; i < 128 * K; i += 128) {
      int k = i / 128, m = i % 128;
      *reinterpret_cast<uint32_t*>(As + (size_t)(m / 4) * ((size_t)K * 16) + (size_t)k * 16 + (size_t)(m % 4) * 4) =
          reinterpret_cast<const uint32_t*>(Ag)[i];
    }
    for (int i = tid; i < N * K; i += 128) {
      int k = i / N, n = i % N;
      *reinterpret_cast<uint32_t*>(Bs + (size_t)(n / 4) * ((size_t)K * 16) + (size_t)k * 16 + (size_t)(n % 4) * 4) =
          reinterpret_cast<const uint32_t*>(Bg)[i];
    }
  } else {
    // MN-major: global holds A^T (K x 128) and B^T (K x N); element (mn, k) at (mn/8)*(K*16) + k*16 + (mn%8)*2
    for (int i = tid; i < 128 * K; i += 128) {
      int k = i / 128, m = i % 128;
      size_t off = (size_t)(m / 8) * ((size_t)K * 16) + (size_t)k * 16 + (size_t)(m % 8) * 2;
      *reinterpret_cast<uint16_t*>(As + off) = reinterpret_cast<const uint16_t*>(Ag)[i];
    }
    for (int i = tid; i < N * K; i += 128) {
      int k = i / N, n = i % N;
      size_t off = (size_t)(n / 8) * ((size_t)K * 16) + (size_t)k * 16 + (size_t)(n % 8) * 2;
      *reinterpret_cast<uint16_t*>(Bs + off) = reinterpret_cast<const uint16_t*>(Bg)[i];
    }
  }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tmem_acc = tmem;                 // columns [0, N)
  const uint32_t tmem_a = tmem + 256;             // columns [256, 256 + K/2) for mode 3
  if (mode == 3) {
    // thread = row: write its K bf16 values, two per 32-bit column
    const uint16_t* arow = reinterpret_cast<const uint16_t*>(Ag) + (size_t)tid * K;
    for (int c0 = 0; c0 < K / 2; c0 += 8) {
      uint32_t v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (uint32_t)arow[2 * (c0 + j)] | ((uint32_t)arow[2 * (c0 + j) + 1] << 16);
      tc::tmem_st8(tmem_a + ((uint32_t)(warp * 32) << 16) + c0, v);
    }
    tc::tmem_st_wait();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
  }
  if (mode == 5) {
    // thread = row: its K fp32 values, one per column
    const uint32_t* arow = reinterpret_cast<const uint32_t*>(Ag) + (size_t)tid * K;
    for (int c0 = 0; c0 < K; c0 += 8) {
      uint32_t v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = arow[c0 + j];
      tc::tmem_st8(tmem_a + ((uint32_t)(warp * 32) << 16) + c0, v);
    }
    tc::tmem_st_wait();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
  }
  if (tid == 0) {
    const uint32_t a0 = tc::smem_u32(As), b0 = tc::smem_u32(Bs);
    if (mode == 0 || mode == 3) {
      const uint32_t idesc = tc::instr_desc(128, N, fmt16, tc::MAJOR_K, tc::MAJOR_K);
      for (int ks = 0; ks < K / 16; ++ks) {
        uint64_t bd = tc::smem_desc(b0 + ks * 2 * (N * 16), N * 16, 128, tc::LAYOUT_NONE);
        if (mode == 0) {
          uint64_t ad = tc::smem_desc(a0 + ks * 2 * (128 * 16), 128 * 16, 128, tc::LAYOUT_NONE);
          tc::umma_f16(tmem_acc, ad, bd, idesc, ks > 0);
        } else {
          tc::umma_f16_ts(tmem_acc, tmem_a + ks * 8, bd, idesc, ks > 0);
        }
      }
    } else if (mode == 1) {
      const uint32_t idesc = tc::instr_desc(128, N, fmt16, tc::MAJOR_MN, tc::MAJOR_MN);
      for (int ks = 0; ks < K / 16; ++ks) {
        uint64_t ad = tc::smem_desc(a0 + ks * 2 * 128, 128, K * 16, tc::LAYOUT_NONE);
        uint64_t bd = tc::smem_desc(b0 + ks * 2 * 128, 128, K * 16, tc::LAYOUT_NONE);
        tc::umma_f16(tmem_acc, ad, bd, idesc, ks > 0);
      }
    } else if (mode == 4) {
      const uint32_t idesc = tc::instr_desc(128, N, tc::FMT_TF32, tc::MAJOR_MN, tc::MAJOR_MN);
      for (int ks = 0; ks < K / 8; ++ks) {   // K step of 8 = one k-group of 8 x 16 B
        uint64_t ad = tc::smem_desc(a0 + ks * 128, 128, K * 16, tc::LAYOUT_NONE);
        uint64_t bd = tc::smem_desc(b0 + ks * 128, 128, K * 16, tc::LAYOUT_NONE);
        tc::umma_tf32(tmem_acc, ad, bd, idesc, ks > 0);
      }
    } else {
      const uint32_t idesc = tc::instr_desc(128, N, tc::FMT_TF32, tc::MAJOR_K, tc::MAJOR_K);
      for (int ks = 0; ks < K / 8; ++ks) {
        uint64_t ad = tc::smem_desc(a0 + ks * 2 * (128 * 16), 128 * 16, 128, tc::LAYOUT_NONE);
        uint64_t bd = tc::smem_desc(b0 + ks * 2 * (N * 16), N * 16, 128, tc::LAYOUT_NONE);
        if (mode == 5) tc::umma_tf32_ts(tmem_acc, tmem_a + ks * 8, bd, idesc, ks > 0);
        else tc::umma_tf32(tmem_acc, ad, bd, idesc, ks > 0);
      }
    }
    tc::umma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    tc::tmem_ld8(tmem_acc + ((uint32_t)(warp * 32) << 16) + c0, v);
    tc::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

}  // namespace

extern "C" int pcreid_tc_probe(int mode, int n, int k, const void* a, const void* b, float* d, void* stream) {
  if (!a || !b || !d || mode < 0 || mode > 8) return PCREID_ERR_ARG;
  if (n < 16 || n > 256 || n % 16 || k < 16 || k > 256 || k % 16) return PCREID_ERR_UNSUPPORTED;
  const int esz = (mode == 2 || mode == 4 || mode == 5) ? 4 : 2;
  size_t smem = (size_t)(128 + n) * k * esz;
  if (smem > 200 * 1024) return PCREID_ERR_UNSUPPORTED;
  cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  tc_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(mode, n, k, a, b, d);
  return pcreid_launch_status();
}
