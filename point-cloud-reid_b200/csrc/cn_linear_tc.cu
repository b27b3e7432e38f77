// Tensor-core version of cn_linear (fast mode): every 1x1 conv / Linear over channel-major (B, C, N) tensors as a
// tcgen05 kind::tf32 GEMM with fp32 accumulation in TMEM.
//
//   Y[b, co, n] = act( sum_k W1[k, co] X1[b, k, n] + sum_k W2[k, co] X2[b, k, n] + bias[co] (+R) ) (+R)
//
// Operands are staged K-major ([k/4][row][4] fp32, the no-swizzle canonical layout validated by tc_probe mode 2): the
// loader reads 4 channel rows per 16-byte chunk (each read coalesced across the 128 points / output channels of the
// tile) and transposes in registers, so channel-major activations (B, K, N) and k-major weights (K, CO) are consumed
// as they lie in HBM, with no alignment requirements; the tensor core reads the fp32 bits as tf32.  (The MN-major
// no-swizzle tf32 layout, which would allow a pure cp.async copy, does not produce a GEMM on this part: probe mode 4.)
// CTA = 128 points x 128 output channels, K streamed in chunks of 32 through a 2-stage ring (global -> registers while
// the previous chunk's MMAs run -> shared memory -> fence.proxy.async -> tcgen05.mma, per-stage mbarrier commits),
// epilogue by the row-owning threads (tcgen05.ld) with fused bias / activation / residual and coalesced
// channel-major stores.  32 KB x 2 of shared memory and 128 TMEM columns per CTA -> three CTAs per SM overlap.
#include "../../include/pcreid.h"
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int KC = 32;                       // K chunk
constexpr int STAGE_A = 32 * KC * 16;        // 128 rows / 4 per group x KC x 16 B = 16 KB
constexpr int STAGE_BYTES = 2 * STAGE_A;     // A + B

// chunk registers <- src[(k0 + 4c + j) * ld + mn] for c < 8, j < 4 (zero outside k < K, mn < mn_end)
__device__ __forceinline__ void fetch_operand(float (&r)[32], const float* __restrict__ src, int ld, int K, int k0, int mn, int mn_end) {
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int kk = k0 + i;
    r[i] = (kk < K && mn < mn_end) ? __ldg(src + (size_t)kk * ld + mn) : 0.f;
  }
}
// [k/4][row][16 B] <- chunk registers
__device__ __forceinline__ void stash_operand(uint8_t* dst, const float (&r)[32], int row) {
#pragma unroll
  for (int c = 0; c < 8; ++c)
    *reinterpret_cast<float4*>(dst + c * 2048 + row * 16) = make_float4(r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]);
}

__global__ void __launch_bounds__(128) cn_linear_tc_kernel(const pcreid_linear_args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t done[2], accbar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float bias_s[128];
  const int b = blockIdx.z, n0 = blockIdx.x * 128, co0 = blockIdx.y * 128;
  const int tid = threadIdx.x;
  const int warp_u = (int)tc::uniform(tid >> 5);
  if (tid == 0) { tc::mbar_init(&done[0], 1); tc::mbar_init(&done[1], 1); tc::mbar_init(&accbar, 1); tc::fence_mbar_init(); }
  if (warp_u == 0) { tc::tmem_alloc(&tmem_base_s, 128); tc::tmem_relinquish(); }
  bias_s[tid] = (a.bias && co0 + tid < a.CO) ? a.bias[co0 + tid] : 0.f;
  const float* X1 = a.X1 + (size_t)b * a.x1_bs;
  const float* W1 = a.W1 + (size_t)b * a.w1_bs;
  const float* X2 = a.K2 > 0 ? a.X2 + (size_t)b * a.x2_bs : nullptr;
  const float* W2 = a.K2 > 0 ? a.W2 + (size_t)b * a.w2_bs : nullptr;
  const int nch1 = (a.K1 + KC - 1) / KC, nch2 = a.K2 > 0 ? (a.K2 + KC - 1) / KC : 0, nch = nch1 + nch2;
  float ra[32], rb[32];
  auto fetch = [&](int c) {
    if (c < nch1) {
      fetch_operand(ra, X1, a.ldx1, a.K1, c * KC, n0 + tid, a.rows);
      fetch_operand(rb, W1, a.CO, a.K1, c * KC, co0 + tid, a.CO);
    } else {
      fetch_operand(ra, X2, a.ldx2, a.K2, (c - nch1) * KC, n0 + tid, a.rows);
      fetch_operand(rb, W2, a.CO, a.K2, (c - nch1) * KC, co0 + tid, a.CO);
    }
  };
  auto stash = [&](int stage) {
    stash_operand(smem + stage * STAGE_BYTES, ra, tid);
    stash_operand(smem + stage * STAGE_BYTES + STAGE_A, rb, tid);
  };
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tc::uniform(tmem_base_s);
  const uint32_t idesc = tc::instr_desc(128, 128, tc::FMT_TF32, tc::MAJOR_K, tc::MAJOR_K);
  const uint32_t s0 = tc::smem_u32(smem);

  fetch(0);
  stash(0);
  for (int c = 0; c < nch; ++c) {
    if (c + 1 < nch) fetch(c + 1);                     // global loads of the next chunk fly while this chunk's MMAs are issued
    tc::fence_async_smem();
    __syncthreads();
    if (warp_u == 0) {
      if (tc::elect_one()) {
        tc::tc_fence_after();
        const uint32_t sa = s0 + (c & 1) * STAGE_BYTES, sb = sa + STAGE_A;
        const uint64_t ad = tc::smem_desc(sa, 2048, 128, tc::LAYOUT_NONE);
        const uint64_t bd = tc::smem_desc(sb, 2048, 128, tc::LAYOUT_NONE);
#pragma unroll
        for (int ks = 0; ks < KC / 8; ++ks)            // K = 8 per MMA = two 16-byte k-chunks = 4096 B
          tc::umma_tf32(tmem, ad + (uint64_t)(ks * 256), bd + (uint64_t)(ks * 256), idesc, (c > 0 || ks > 0) ? 1u : 0u);
        tc::umma_commit(&done[c & 1]);
        if (c == nch - 1) tc::umma_commit(&accbar);
      }
      __syncwarp();
    }
    if (c + 1 < nch) {
      const int s1 = (c + 1) & 1;
      if (c + 1 >= 2) tc::mbar_wait(&done[s1], (uint32_t)((((c + 1) >> 1) - 1) & 1));   // MMAs of chunk c-1 released the stage
      stash(s1);
    }
  }
  tc::mbar_wait(&accbar, 0);
  tc::tc_fence_after();
  // ---- epilogue: thread == point row
  const int n = n0 + tid;
  const uint32_t tl = tmem + ((uint32_t)((tid >> 5) * 32) << 16);
  float* Y = a.Y + (size_t)b * a.y_bs;
  const float* R = a.R ? a.R + (size_t)b * a.r_bs : nullptr;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    uint32_t r[16];
    tc::tmem_ld16(tl + 16 * q, r);
    tc::tmem_ld_wait();
    if (n < a.rows) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int co = co0 + 16 * q + j;
        if (co < a.CO) {
          float v = __uint_as_float(r[j]) + bias_s[16 * q + j];
          const float rr = R ? R[(size_t)co * a.ldr + n] : 0.f;
          if (R && !a.res_after_act) v += rr;
          v = apply_act(v, a.act);
          if (R && a.res_after_act) v += rr;
          if (a.y_pm) Y[(size_t)n * a.ldy + co] = v;
          else Y[(size_t)co * a.ldy + n] = v;
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp_u == 0) tc::tmem_dealloc(tmem, 128);
}

}  // namespace

extern "C" int pcreid_cn_linear_tc(const pcreid_linear_args* p, void* stream) {
  if (!p) return PCREID_ERR_ARG;
  const pcreid_linear_args& a = *p;
  if (a.B <= 0 || a.rows <= 0 || a.CO <= 0) return PCREID_OK;
  if (a.K1 <= 0 || !a.X1 || !a.W1 || !a.Y) return PCREID_ERR_ARG;
  if (a.K2 > 0 && (!a.X2 || !a.W2)) return PCREID_ERR_ARG;
  // shapes this kernel was built for; everything else stays on the FFMA kernel (pcreid_cn_linear)
  if (a.x1_map || a.x2_map || a.w1_map || a.r_map || a.x1_pm || a.x2_pm) return PCREID_ERR_UNSUPPORTED;
  if (a.K1 < 8 || a.B > 65535) return PCREID_ERR_UNSUPPORTED;       // tiny K (xyz inputs) stays on the FFMA kernel
  const int smem = 2 * STAGE_BYTES;
  cudaFuncSetAttribute(cn_linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  dim3 grid((a.rows + 127) / 128, (a.CO + 127) / 128, a.B);
  cn_linear_tc_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(a);
  return pcreid_launch_status();
}
