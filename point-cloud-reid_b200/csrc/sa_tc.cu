// Set-abstraction edge MLP on the 5th-gen tensor cores (tcgen05, kind::tf32, fp32 accumulate in TMEM) -- the
// "fast" mode counterpart of sa_edge_mlp_kernel (rowops.cu).
//
//   out[b,c,s] = max_j relu(W3 relu(W2 relu(P1[b,idx[b,s,j],:] + Cc[b,s,:]) + b2) + b3)[c]      (P1, Cc point-major)
// (PointNetSetAbstractionEdgeSA.forward, mmdet3d/models/pointnet2_utils.py:333-357, first conv factorised per point /
// per centre on the host side, eval BatchNorm folded).
//
// One persistent CTA per SM (8 warps = 4 TMEM lane quadrants x 2 column halves).  A tile is the 128-row GEMM M
// dimension = floor(128/k) centres x k neighbours.  Per tile: gather + ReLU straight into the fp32 operand image
// [c/4][row][4] (no-swizzle K-major), GEMM (K = N = C) -> bias + ReLU epilogue back into the operand image ->
// second GEMM -> bias + ReLU -> shared-memory transpose -> max over the k rows of each centre.  The grouped
// (B, C, S, k) tensor of the reference never exists; both weight matrices stay resident in shared memory.
#include "../../include/pcreid.h"
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int NT_ = 256;

template <int C>
__global__ void __launch_bounds__(NT_, (C == 128 ? 1 : (C == 64 ? 3 : 4))) sa_edge_mlp_tc_kernel(int N, int S, int k, int cpt, int tiles_per_obj, int total_tiles,
                                                               const float* __restrict__ P1, const float* __restrict__ Cc,
                                                               const int* __restrict__ idx, const float* __restrict__ W2img,
                                                               const float* __restrict__ b2, const float* __restrict__ W3img,
                                                               const float* __restrict__ b3, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float bias_s[2][C];
  constexpr int WBYTES = C * C * 4;            // one weight image [C/4][C][4]
  constexpr int ABYTES = C * 128 * 4;          // activation image [C/4][128][4] == transpose buffer [C][128]
  constexpr int TCOLS = C < 32 ? 32 : C;
  uint8_t* W2s = smem;
  uint8_t* W3s = smem + WBYTES;
  uint8_t* As = smem + 2 * WBYTES;
  float* Tf = reinterpret_cast<float*>(As);
  const int t = threadIdx.x, warp = t >> 5;
  const int row = 32 * (warp & 3) + (t & 31), h = warp >> 2;
  if (t == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  if (warp == 0) { tc::tmem_alloc(&tmem_base_s, TCOLS); tc::tmem_relinquish(); }
  for (int i = t * 16; i < WBYTES; i += NT_ * 16) {
    cp_async16(W2s + i, reinterpret_cast<const uint8_t*>(W2img) + i);
    cp_async16(W3s + i, reinterpret_cast<const uint8_t*>(W3img) + i);
  }
  cp_async_commit();
  for (int i = t; i < C; i += NT_) { bias_s[0][i] = b2[i]; bias_s[1][i] = b3[i]; }
  cp_async_wait<0>();
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t sA = tc::smem_u32(As), sW2 = tc::smem_u32(W2s), sW3 = tc::smem_u32(W3s);
  const uint32_t idesc = tc::instr_desc(128, C, tc::FMT_TF32, tc::MAJOR_K, tc::MAJOR_K);
  uint32_t par = 0;
  constexpr int CH = C / 2;                    // channels (columns) per thread

  // neighbour index of this thread's row, fetched one tile ahead so that the dependent row gather of the next tile does
  // not start with an exposed L2 round trip
  auto fetch_idx = [&](int tile_) {
    const int b_ = tile_ / tiles_per_obj, s0_ = (tile_ % tiles_per_obj) * cpt;
    const int nedge_ = min(cpt, S - s0_) * k;
    return row < nedge_ ? __ldg(idx + ((size_t)b_ * S + s0_ + row / k) * k + (row % k)) : -1;
  };
  int src_next = blockIdx.x < total_tiles ? fetch_idx(blockIdx.x) : -1;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int b = tile / tiles_per_obj, s0 = (tile % tiles_per_obj) * cpt;
    const int ncen = min(cpt, S - s0), nedge = ncen * k;
    const int src_cur = src_next;
    if (tile + (int)gridDim.x < total_tiles) src_next = fetch_idx(tile + gridDim.x);
    // P1 / Cc are point-major here ((B, N, C) / (B, S, C)): a gathered neighbour is one contiguous C-vector
    const float* Pb = P1 + (size_t)b * N * C;
    const float* Cb = Cc + (size_t)b * S * C;
    // ---- gather + relu(P1 + Cc) -> operand image (128-bit loads; chunk of 4 channels == one 16-byte image chunk)
    {
      int src = -1, cen = 0;
      if (row < nedge) { cen = s0 + row / k; src = src_cur; }
      const float4* prow = reinterpret_cast<const float4*>(Pb + (size_t)max(src, 0) * C) + h * (CH / 4);
      const float4* crow = reinterpret_cast<const float4*>(Cb + (size_t)cen * C) + h * (CH / 4);
#pragma unroll
      for (int c4 = 0; c4 < CH / 4; ++c4) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (src >= 0) {
          const float4 p = __ldg(prow + c4), q = __ldg(crow + c4);
          v = make_float4(fmaxf(p.x + q.x, 0.f), fmaxf(p.y + q.y, 0.f), fmaxf(p.z + q.z, 0.f), fmaxf(p.w + q.w, 0.f));
        }
        *reinterpret_cast<float4*>(As + (h * (CH / 4) + c4) * 2048 + row * 16) = v;
      }
    }
    for (int layer = 0; layer < 2; ++layer) {
      tc::fence_async_smem();
      tc::tc_fence_before();
      __syncthreads();
      tc::tc_fence_after();
      if (t == 0) {
        const uint32_t sW = layer == 0 ? sW2 : sW3;
        for (int ks = 0; ks < C / 8; ++ks) {
          const uint64_t ad = tc::smem_desc(sA + ks * 4096, 2048, 128, tc::LAYOUT_NONE);
          const uint64_t bd = tc::smem_desc(sW + ks * 2 * (C * 16), C * 16, 128, tc::LAYOUT_NONE);
          tc::umma_tf32(tmem, ad, bd, idesc, ks > 0 ? 1u : 0u);
        }
        tc::umma_commit(&bar);
      }
      tc::mbar_wait(&bar, par);
      par ^= 1u;
      tc::tc_fence_after();
      // ---- epilogue: + bias, ReLU; layer 0 -> operand image, layer 1 -> rotated transpose buffer
#pragma unroll
      for (int q = 0; q < CH / 16; ++q) {
        uint32_t r[16];
        const int c0 = h * CH + 16 * q;
        tc::tmem_ld16(tlane + c0, r);
        tc::tmem_ld_wait();
        if (layer == 0) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 v;
            v.x = fmaxf(__uint_as_float(r[j + 0]) + bias_s[0][c0 + j + 0], 0.f);
            v.y = fmaxf(__uint_as_float(r[j + 1]) + bias_s[0][c0 + j + 1], 0.f);
            v.z = fmaxf(__uint_as_float(r[j + 2]) + bias_s[0][c0 + j + 2], 0.f);
            v.w = fmaxf(__uint_as_float(r[j + 3]) + bias_s[0][c0 + j + 3], 0.f);
            *reinterpret_cast<float4*>(As + ((c0 + j) / 4) * 2048 + row * 16) = v;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int c = c0 + j;
            Tf[c * 128 + ((row + c) & 127)] = fmaxf(__uint_as_float(r[j]) + bias_s[1][c], 0.f);
          }
        }
      }
    }
    tc::tc_fence_before();
    __syncthreads();
    // ---- max over the k edges of each centre
    for (int i = t; i < C * ncen; i += NT_) {
      const int c = i / ncen, cl = i % ncen;
      float mx = 0.f;                                   // values are post-ReLU (>= 0)
      for (int j = 0; j < k; ++j) mx = fmaxf(mx, Tf[c * 128 + ((cl * k + j + c) & 127)]);
      out[((size_t)b * C + c) * S + s0 + cl] = mx;
    }
    __syncthreads();                                    // the next tile's gather overwrites the buffer
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, TCOLS);
}

template <int C>
int launch(int B, int N, int S, int k, const float* P1, const float* Cc, const int* idx, const float* W2img, const float* b2,
           const float* W3img, const float* b3, float* out, int n_ctas, cudaStream_t st) {
  const int cpt = 128 / k, tiles_per_obj = (S + cpt - 1) / cpt;
  const long long total = (long long)B * tiles_per_obj;
  if (total > 0x7fffffffLL) return PCREID_ERR_UNSUPPORTED;
  const int smem = 2 * C * C * 4 + C * 128 * 4;
  cudaFuncSetAttribute(sa_edge_mlp_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int ctas_per_sm = C == 128 ? 1 : (C == 64 ? 3 : 4);      // shared memory (192 / 64 / 24 KB) and registers bound residency
  int grid = (n_ctas > 0 ? n_ctas : 148) * ctas_per_sm;
  if (grid > total) grid = (int)total;
  sa_edge_mlp_tc_kernel<C><<<grid, NT_, smem, st>>>(N, S, k, cpt, tiles_per_obj, (int)total, P1, Cc, idx, W2img, b2, W3img, b3, out);
  return pcreid_launch_status();
}

}  // namespace

extern "C" int pcreid_sa_edge_mlp_tc(int B, int C, int N, int S, int k, const float* P1, const float* Cc, const int* idx,
                                     const float* W2img, const float* b2, const float* W3img, const float* b3, float* out,
                                     int n_ctas, void* stream) {
  if (B <= 0 || S <= 0) return PCREID_OK;
  if (!P1 || !Cc || !idx || !W2img || !b2 || !W3img || !b3 || !out || k <= 0 || N <= 0) return PCREID_ERR_ARG;
  if (k > 128) return PCREID_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  switch (C) {
    case 32: return launch<32>(B, N, S, k, P1, Cc, idx, W2img, b2, W3img, b3, out, n_ctas, st);
    case 64: return launch<64>(B, N, S, k, P1, Cc, idx, W2img, b2, W3img, b3, out, n_ctas, st);
    case 128: return launch<128>(B, N, S, k, P1, Cc, idx, W2img, b2, W3img, b3, out, n_ctas, st);
    default: return PCREID_ERR_UNSUPPORTED;
  }
}
