// cn_linear on the tensor cores, second generation (fast mode, K >= 256): warp-specialised persistent tf32 GEMM.
//
//   Y[b, co, n] = act( sum_k W1[k, co] X1[b, k, n] + sum_k W2[k, co] X2[b, k, n] + bias[co] (+R) ) (+R)
//
// The first generation (cn_linear_tc.cu) has one set of 128 threads load, stage, issue and drain a tile in turn and
// reaches ~100 TFLOP/s on the 512 x 1024 / 1024 x 512 heads of DGCNN / PointNet (21 of 40 ms of the DGCNN encoder).
// Here a CTA of 10 warps runs three pipelines over a persistent tile loop (tile = 128 points x 128 output channels):
//   warps 4-7  activation loaders: channel-major (B, K, N) rows are read coalesced along the points, 32 channels deep,
//              transposed in registers into the K-major tf32 operand image [k/4][row][4] (the loads of chunk g+1 are
//              in flight while chunk g is stored);
//   warp  9    weight loader: the weights arrive as pre-built operand images [k/4][co][4]; one thread streams the
//              128-column slice of each 32-channel chunk with 1-D bulk TMA copies (cp.async.bulk + complete_tx);
//   warp  8    MMA issuer: tcgen05.mma kind::tf32 per ready stage, tcgen05.commit releases the stage; the accumulator
//              (128 TMEM columns) is double buffered, so the epilogue of tile i overlaps the main loop of tile i+1;
//   warps 0-3  epilogue (thread == point row == TMEM lane): bias / activation / residual, coalesced channel-major stores.
// Three stages of 32 KB per CTA, two CTAs per SM.
#include "../../include/pcreid.h"
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int ST = 3;
constexpr int KC = 32;
constexpr int STAGE_A = 128 * KC * 4;        // 16 KB
constexpr int STAGE_BYTES = 2 * STAGE_A;     // A + B
constexpr int NTHR2 = 320;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}

struct Lin2Args {
  pcreid_linear_args a;
  const float *W1img, *W2img;      // [K/4][CO][4]
  int tiles_n, tiles_c, total_tiles;
};

__global__ void __launch_bounds__(NTHR2, 2) cn_linear_tc2_kernel(const __grid_constant__ Lin2Args p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full[ST], empty[ST], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  const pcreid_linear_args& a = p.a;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < ST; ++i) { tc::mbar_init(&full[i], 129); tc::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&acc_full[i], 1); tc::mbar_init(&acc_empty[i], 128); }
    tc::fence_mbar_init();
  }
  if (warp == 0) { tc::tmem_alloc(&tmem_base_s, 256); tc::tmem_relinquish(); }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const int nch1 = (a.K1 + KC - 1) / KC, nch2 = a.K2 > 0 ? (a.K2 + KC - 1) / KC : 0, nch = nch1 + nch2;
  const int per_obj = p.tiles_n * p.tiles_c;

  if (warp >= 4 && warp < 8) {
    // ============================================================ activation loaders
    const int r = tid - 128;
    int g = 0;
    float cur[32], nxt[32];
    auto fetch = [&](float (&v)[32], int tile, int c) {
      const int b = tile / per_obj, n0 = ((tile % per_obj) / p.tiles_c) * 128;
      const bool second = c >= nch1;
      const float* X = second ? a.X2 + (size_t)b * a.x2_bs : a.X1 + (size_t)b * a.x1_bs;
      const int ld = second ? a.ldx2 : a.ldx1, K = second ? a.K2 : a.K1, k0 = (second ? c - nch1 : c) * KC;
      const int n = n0 + r;
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = (k0 + i < K && n < a.rows) ? __ldg(X + (size_t)(k0 + i) * ld + n) : 0.f;
    };
    int tile = blockIdx.x;
    if (tile < p.total_tiles) fetch(cur, tile, 0);
    for (; tile < p.total_tiles; tile += gridDim.x) {
      for (int c = 0; c < nch; ++c, ++g) {
        // next chunk's loads fly while this chunk waits for its stage and is stored
        int ntile = tile, nc = c + 1;
        if (nc == nch) { nc = 0; ntile = tile + gridDim.x; }
        const bool has_next = ntile < p.total_tiles;
        if (has_next) fetch(nxt, ntile, nc);
        const int s = g % ST, use = g / ST;
        if (use > 0) tc::mbar_wait(&empty[s], (uint32_t)((use - 1) & 1));
        uint8_t* As = smem + s * STAGE_BYTES;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(As + j * 2048 + r * 16) = make_float4(cur[4 * j], cur[4 * j + 1], cur[4 * j + 2], cur[4 * j + 3]);
        tc::fence_async_smem();
        mbar_arrive(&full[s]);
        if (has_next) {
#pragma unroll
          for (int i = 0; i < 32; ++i) cur[i] = nxt[i];
        }
      }
    }
  } else if (warp == 9) {
    // ============================================================ weight loader (one thread)
    if (lane == 0) {
      int g = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int co0 = (tile % p.tiles_c) * 128;
        const int nvalid = min(128, a.CO - co0);
        for (int c = 0; c < nch; ++c, ++g) {
          const bool second = c >= nch1;
          const float* W = second ? p.W2img : p.W1img;
          const int K = second ? a.K2 : a.K1, k0 = (second ? c - nch1 : c) * KC;
          const int planes = min(KC, K - k0) / 4;
          const int s = g % ST, use = g / ST;
          if (use > 0) tc::mbar_wait(&empty[s], (uint32_t)((use - 1) & 1));
          const uint32_t Bs = tc::smem_u32(smem + s * STAGE_BYTES + STAGE_A);
          mbar_expect_tx(&full[s], (uint32_t)(planes * nvalid * 16));
          for (int j = 0; j < planes; ++j)
            bulk_copy(Bs + j * 2048, W + ((size_t)(k0 / 4 + j) * a.CO + co0) * 4, (uint32_t)(nvalid * 16), &full[s]);
        }
      }
    }
  } else if (warp == 8) {
    // ============================================================ MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc = tc::instr_desc(128, 128, tc::FMT_TF32, tc::MAJOR_K, tc::MAJOR_K);
      int g = 0, ti = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++ti) {
        const int buf = ti & 1, au = ti >> 1;
        if (au > 0) tc::mbar_wait(&acc_empty[buf], (uint32_t)((au - 1) & 1));
        tc::tc_fence_after();
        const uint32_t d = tmem + (uint32_t)(buf * 128);
        for (int c = 0; c < nch; ++c, ++g) {
          const bool second = c >= nch1;
          const int K = second ? a.K2 : a.K1, k0 = (second ? c - nch1 : c) * KC;
          const int ksteps = min(KC, K - k0) / 8;
          const int s = g % ST, use = g / ST;
          tc::mbar_wait(&full[s], (uint32_t)(use & 1));
          tc::tc_fence_after();
          const uint32_t sa = tc::smem_u32(smem + s * STAGE_BYTES), sb = sa + STAGE_A;
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t ad = tc::smem_desc(sa + ks * 4096, 2048, 128, tc::LAYOUT_NONE);
            const uint64_t bd = tc::smem_desc(sb + ks * 4096, 2048, 128, tc::LAYOUT_NONE);
            tc::umma_tf32(d, ad, bd, idesc, (c > 0 || ks > 0) ? 1u : 0u);
          }
          tc::umma_commit(&empty[s]);
        }
        tc::umma_commit(&acc_full[buf]);
      }
    }
  } else {
    // ============================================================ epilogue (thread == point row)
    int ti = 0;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++ti) {
      const int b = tile / per_obj, n0 = ((tile % per_obj) / p.tiles_c) * 128, co0 = (tile % p.tiles_c) * 128;
      const int buf = ti & 1, au = ti >> 1;
      const int n = n0 + tid;
      float* Y = a.Y + (size_t)b * a.y_bs;
      const float* R = a.R ? a.R + (size_t)b * a.r_bs : nullptr;
      tc::mbar_wait(&acc_full[buf], (uint32_t)(au & 1));
      tc::tc_fence_after();
      const uint32_t tl = tmem + (uint32_t)(buf * 128) + lane_off;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        uint32_t rg[16];
        tc::tmem_ld16(tl + 16 * q, rg);
        tc::tmem_ld_wait();
        if (n < a.rows) {
          float rr[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int co = co0 + 16 * q + j;
            rr[j] = (R && co < a.CO) ? __ldg(R + (size_t)co * a.ldr + n) : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int co = co0 + 16 * q + j;
            if (co < a.CO) {
              float v = __uint_as_float(rg[j]) + (a.bias ? __ldg(a.bias + co) : 0.f);
              if (R && !a.res_after_act) v += rr[j];
              v = apply_act(v, a.act);
              if (R && a.res_after_act) v += rr[j];
              if (a.y_pm) Y[(size_t)n * a.ldy + co] = v;
              else Y[(size_t)co * a.ldy + n] = v;
            }
          }
        }
      }
      tc::tc_fence_before();
      mbar_arrive(&acc_empty[buf]);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

}  // namespace

extern "C" int pcreid_cn_linear_tc2(const pcreid_linear_args* pa, const float* W1img, const float* W2img, int n_sms, void* stream) {
  if (!pa) return PCREID_ERR_ARG;
  const pcreid_linear_args& a = *pa;
  if (a.B <= 0 || a.rows <= 0 || a.CO <= 0) return PCREID_OK;
  if (a.K1 <= 0 || !a.X1 || !W1img || !a.Y) return PCREID_ERR_ARG;
  if (a.K2 > 0 && (!a.X2 || !W2img)) return PCREID_ERR_ARG;
  // shapes this kernel was built for; everything else stays on pcreid_cn_linear_tc / pcreid_cn_linear
  if (a.x1_map || a.x2_map || a.w1_map || a.r_map || a.x1_pm || a.x2_pm || a.w1_bs || a.w2_bs) return PCREID_ERR_UNSUPPORTED;
  if (a.K1 % 8 || a.K2 % 8) return PCREID_ERR_UNSUPPORTED;
  Lin2Args p;
  p.a = a; p.W1img = W1img; p.W2img = W2img;
  p.tiles_n = (a.rows + 127) / 128;
  p.tiles_c = (a.CO + 127) / 128;
  const long long total = (long long)a.B * p.tiles_n * p.tiles_c;
  if (total > 0x7fffffffLL) return PCREID_ERR_UNSUPPORTED;
  p.total_tiles = (int)total;
  if (n_sms <= 0) n_sms = 148;
  const int smem = ST * STAGE_BYTES;
  cudaFuncSetAttribute(cn_linear_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int grid = total < 2LL * n_sms ? (int)total : 2 * n_sms;
  cn_linear_tc2_kernel<<<grid, NTHR2, smem, (cudaStream_t)stream>>>(p);
  return pcreid_launch_status();
}
