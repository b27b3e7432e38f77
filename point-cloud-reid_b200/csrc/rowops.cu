// fp32 "parity mode" building blocks of the encoder / match path (sm_100a SIMT):
//   cn_linear      -- every 1x1 conv / nn.Linear of the path as a smem-tiled GEMM over channel-major
//                     (B, C, N) tensors: 128 points x 64|128 output channels per CTA, 8x4|8x8 register
//                     micro-tiles, cp.async double-buffered operand chunks, fused bias/activation/residual
//   cn_groupnorm   -- LayerNorm / GroupNorm per point with fused residual + activation
//   linattn_kv / linattn_scale -- the two reductions of linear attention
//   cn_pool / cn_chanmax       -- pooling
//   sa_edge_mlp    -- fused gather + 2-layer shared MLP + max-over-k of a set-abstraction layer
//   edge_gather_max-- EdgeConv neighbour max
//   pair_concat_head -- fused all-pairs 'concat' match head
// All fp32 FFMA with fp32 accumulation: this is the mode that meets the 1e-4 logit parity gate; the
// tcgen05 kernels (pair_tc.cu) are the throughput mode.
#include "../../include/pcreid.h"
#include "common.cuh"

namespace {

constexpr int TR = 128;   // points (rows) per CTA tile
constexpr int KC = 16;    // reduction chunk
constexpr int NTHR = 256;

template <int TN>
struct TileCols { static constexpr int value = 16 * TN; };

// acc[8][TN] += Xs[kk][rows] * Ws[kk][cols] for kk < KC.  Thread (tx, ty): rows {tx*4+i, 64+tx*4+i},
// cols {ty*4+j} (TN=4) or {ty*4+j, 64+ty*4+j} (TN=8) -> conflict-free 128-bit LDS.
template <int TN>
__device__ __forceinline__ void fma_chunk(const float* __restrict__ Xs, const float* __restrict__ Ws, int tx, int ty,
                                          float (&acc)[8][TN]) {
  constexpr int TC = TileCols<TN>::value;
#pragma unroll
  for (int kk = 0; kk < KC; ++kk) {
    const float4 a0 = *reinterpret_cast<const float4*>(Xs + kk * TR + tx * 4);
    const float4 a1 = *reinterpret_cast<const float4*>(Xs + kk * TR + 64 + tx * 4);
    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    float w[TN];
    {
      const float4 w0 = *reinterpret_cast<const float4*>(Ws + kk * TC + ty * 4);
      w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w;
      if (TN == 8) {
        const float4 w1 = *reinterpret_cast<const float4*>(Ws + kk * TC + 64 + ty * 4);
        w[TN - 4] = w1.x; w[TN - 3] = w1.y; w[TN - 2] = w1.z; w[TN - 1] = w1.w;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
  }
}

template <int TN>
__device__ __forceinline__ int col_of(int ty, int j) {
  return (TN == 8 && j >= 4) ? 64 + ty * 4 + (j - 4) : ty * 4 + j;
}
__device__ __forceinline__ int row_of(int tx, int i) { return (i >= 4 ? 64 : 0) + tx * 4 + (i & 3); }

// stage a [KC][TC] chunk of a k-major weight matrix W[K][CO] (rows k0.., cols co0..) into shared memory
template <int TN>
__device__ __forceinline__ void load_w_chunk(float* __restrict__ Ws, const float* __restrict__ W, int K, int CO, int k0,
                                             int co0, bool vec) {
  constexpr int TC = TileCols<TN>::value;
  if (vec) {
    for (int i = threadIdx.x; i < KC * TC / 4; i += NTHR) {
      int kk = i / (TC / 4), c4 = (i % (TC / 4)) * 4;
      float* dst = Ws + kk * TC + c4;
      int k = k0 + kk, co = co0 + c4;
      if (k < K && co < CO) cp_async16(dst, W + (size_t)k * CO + co);
      else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else {
    for (int i = threadIdx.x; i < KC * TC; i += NTHR) {
      int kk = i / TC, c = i % TC;
      int k = k0 + kk, co = co0 + c;
      Ws[kk * TC + c] = (k < K && co < CO) ? __ldg(W + (size_t)k * CO + co) : 0.f;
    }
  }
}

// stage a [KC][TR] chunk of an activation tensor (channel-major or point-major) into shared memory
__device__ __forceinline__ void load_x_chunk(float* __restrict__ Xs, const float* __restrict__ X, int K, int ld, int pm,
                                             int rows, int k0, int n0, bool vec) {
  if (!pm && vec) {
    for (int i = threadIdx.x; i < KC * TR / 4; i += NTHR) {
      int kk = i / (TR / 4), r4 = (i % (TR / 4)) * 4;
      float* dst = Xs + kk * TR + r4;
      int k = k0 + kk, n = n0 + r4;
      if (k < K && n < rows) cp_async16(dst, X + (size_t)k * ld + n);   // vec => rows % 4 == 0
      else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else if (!pm) {
    for (int i = threadIdx.x; i < KC * TR; i += NTHR) {
      int kk = i / TR, r = i % TR;
      int k = k0 + kk, n = n0 + r;
      Xs[kk * TR + r] = (k < K && n < rows) ? __ldg(X + (size_t)k * ld + n) : 0.f;
    }
  } else {
    for (int i = threadIdx.x; i < KC * TR; i += NTHR) {
      int r = i / KC, kk = i % KC;
      int k = k0 + kk, n = n0 + r;
      Xs[kk * TR + r] = (k < K && n < rows) ? __ldg(X + (size_t)n * ld + k) : 0.f;
    }
  }
}

// packed variant: the 128-row tile holds TR / rows whole objects (rows < TR, TR % rows == 0, rows % 4 == 0): tile row r is
// point r % rows of object b0 + r / rows.  Channel-major, 16-byte aligned sources only.
__device__ __forceinline__ void load_x_chunk_packed(float* __restrict__ Xs, const float* __restrict__ X, long long bs, int K, int ld,
                                                    int rows, int k0, int b0, int B) {
  for (int i = threadIdx.x; i < KC * TR / 4; i += NTHR) {
    int kk = i / (TR / 4), r4 = (i % (TR / 4)) * 4;
    float* dst = Xs + kk * TR + r4;
    int k = k0 + kk, ob = b0 + r4 / rows, n = r4 % rows;
    if (k < K && ob < B) cp_async16(dst, X + (size_t)ob * bs + (size_t)k * ld + n);
    else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ------------------------------------------------------------------------------------------------
// cn_linear
// ------------------------------------------------------------------------------------------------
// PACK: objects with fewer than 128 rows share a tile (shared weights, channel-major, no gather maps): the deep levels of the
// encoders (64 / 32 points per object) otherwise run the 128-row tile 50 % / 75 % empty
template <int TN, bool PACK = false>
__global__ void __launch_bounds__(NTHR) cn_linear_kernel(const pcreid_linear_args a) {
  constexpr int TC = TileCols<TN>::value;
  __shared__ __align__(16) float Xs[2][KC * TR];
  __shared__ __align__(16) float Ws[2][KC * TC];
  const int b = PACK ? blockIdx.z * (TR / a.rows) : blockIdx.z, n0 = blockIdx.x * TR, co0 = blockIdx.y * TC;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  if (PACK) {
    const float* W1 = a.W1;
    const float* W2 = a.K2 > 0 ? a.W2 : nullptr;
    const int nch1 = ceil_div(a.K1, KC), nch2 = a.K2 > 0 ? ceil_div(a.K2, KC) : 0, nch = nch1 + nch2;
    auto load = [&](int stage, int ch) {
      if (ch < nch1) {
        load_x_chunk_packed(Xs[stage], a.X1, a.x1_bs, a.K1, a.ldx1, a.rows, ch * KC, b, a.B);
        load_w_chunk<TN>(Ws[stage], W1, a.K1, a.CO, ch * KC, co0, true);
      } else {
        load_x_chunk_packed(Xs[stage], a.X2, a.x2_bs, a.K2, a.ldx2, a.rows, (ch - nch1) * KC, b, a.B);
        load_w_chunk<TN>(Ws[stage], W2, a.K2, a.CO, (ch - nch1) * KC, co0, true);
      }
    };
    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    load(0, 0);
    cp_async_commit();
    for (int ch = 0; ch < nch; ++ch) {
      if (ch + 1 < nch) load((ch + 1) & 1, ch + 1);
      cp_async_commit();
      cp_async_wait<1>();
      __syncthreads();
      fma_chunk<TN>(Xs[ch & 1], Ws[ch & 1], tx, ty, acc);
      __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int co = co0 + col_of<TN>(ty, j);
      if (co >= a.CO) continue;
      const float bv = a.bias ? __ldg(a.bias + co) : 0.f;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = h * 64 + tx * 4, ob = b + r / a.rows, n = r % a.rows;
        if (ob >= a.B) continue;
        float4 rr = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.R) rr = *reinterpret_cast<const float4*>(a.R + (size_t)ob * a.r_bs + (size_t)co * a.ldr + n);
        const float rv[4] = {rr.x, rr.y, rr.z, rr.w};
        float v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float x = acc[h * 4 + i][j] + bv;
          if (a.R && !a.res_after_act) x += rv[i];
          x = apply_act(x, a.act);
          if (a.R && a.res_after_act) x += rv[i];
          v[i] = x;
        }
        *reinterpret_cast<float4*>(a.Y + (size_t)ob * a.y_bs + (size_t)co * a.ldy + n) = make_float4(v[0], v[1], v[2], v[3]);
      }
    }
    return;
  }

  const int xb1 = a.x1_map ? a.x1_map[b] : b;
  const float* X1 = a.X1 + (size_t)xb1 * a.x1_bs;
  const int wb1 = a.w1_map ? a.w1_map[b] : b;
  const float* W1 = a.W1 + (size_t)wb1 * a.w1_bs;
  const float* X2 = nullptr;
  const float* W2 = nullptr;
  if (a.K2 > 0) {
    const int xb2 = a.x2_map ? a.x2_map[b] : b;
    X2 = a.X2 + (size_t)xb2 * a.x2_bs;
    W2 = a.W2 + (size_t)b * a.w2_bs;
  }
  const bool vx1 = !a.x1_pm && (a.ldx1 % 4 == 0) && (a.rows % 4 == 0) && aligned16(X1);
  const bool vx2 = a.K2 > 0 && !a.x2_pm && (a.ldx2 % 4 == 0) && (a.rows % 4 == 0) && aligned16(X2);
  const bool vw1 = (a.CO % 4 == 0) && aligned16(W1);
  const bool vw2 = a.K2 > 0 && (a.CO % 4 == 0) && aligned16(W2);

  const int nch1 = ceil_div(a.K1, KC), nch2 = a.K2 > 0 ? ceil_div(a.K2, KC) : 0, nch = nch1 + nch2;
  auto load = [&](int stage, int ch) {
    if (ch < nch1) {
      load_x_chunk(Xs[stage], X1, a.K1, a.ldx1, a.x1_pm, a.rows, ch * KC, n0, vx1);
      load_w_chunk<TN>(Ws[stage], W1, a.K1, a.CO, ch * KC, co0, vw1);
    } else {
      load_x_chunk(Xs[stage], X2, a.K2, a.ldx2, a.x2_pm, a.rows, (ch - nch1) * KC, n0, vx2);
      load_w_chunk<TN>(Ws[stage], W2, a.K2, a.CO, (ch - nch1) * KC, co0, vw2);
    }
  };

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  load(0, 0);
  cp_async_commit();
  for (int ch = 0; ch < nch; ++ch) {
    if (ch + 1 < nch) load((ch + 1) & 1, ch + 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    fma_chunk<TN>(Xs[ch & 1], Ws[ch & 1], tx, ty, acc);
    __syncthreads();
  }

  // epilogue
  float* Y = a.Y + (size_t)b * a.y_bs;
  const float* R = nullptr;
  if (a.R) R = a.R + (size_t)(a.r_map ? a.r_map[b] : b) * a.r_bs;
  const bool vy = !a.y_pm && (a.ldy % 4 == 0) && (a.rows % 4 == 0) && aligned16(Y) && (!R || ((a.ldr % 4 == 0) && aligned16(R)));
  if (a.y_pm && !R && (a.ldy % 4 == 0) && (a.CO % 4 == 0) && aligned16(Y)) {
    // point-major output without residual: each thread writes its 4 (or 2 x 4) consecutive channels of a row as 128-bit stores
#pragma unroll
    for (int jb = 0; jb < TN; jb += 4) {
      const int co = co0 + col_of<TN>(ty, jb);
      if (co >= a.CO) continue;
      float bv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = a.bias ? __ldg(a.bias + co + j) : 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int n = n0 + (i >= 4 ? 64 : 0) + tx * 4 + (i & 3);
        if (n >= a.rows) continue;
        float4 v;
        v.x = apply_act(acc[i][jb + 0] + bv[0], a.act);
        v.y = apply_act(acc[i][jb + 1] + bv[1], a.act);
        v.z = apply_act(acc[i][jb + 2] + bv[2], a.act);
        v.w = apply_act(acc[i][jb + 3] + bv[3], a.act);
        *reinterpret_cast<float4*>(Y + (size_t)n * a.ldy + co) = v;
      }
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int co = co0 + col_of<TN>(ty, j);
    if (co >= a.CO) continue;
    const float bv = a.bias ? __ldg(a.bias + co) : 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = n0 + h * 64 + tx * 4;
      if (n >= a.rows) continue;
      float v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = acc[h * 4 + i][j] + bv;
      if (vy) {
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (R) r = *reinterpret_cast<const float4*>(R + (size_t)co * a.ldr + n);
        const float rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (R && !a.res_after_act) v[i] += rr[i];
          v[i] = apply_act(v[i], a.act);
          if (R && a.res_after_act) v[i] += rr[i];
        }
        *reinterpret_cast<float4*>(Y + (size_t)co * a.ldy + n) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (n + i >= a.rows) break;
          float x = v[i];
          float r = R ? R[(size_t)co * a.ldr + n + i] : 0.f;
          if (R && !a.res_after_act) x += r;
          x = apply_act(x, a.act);
          if (R && a.res_after_act) x += r;
          if (a.y_pm) Y[(size_t)(n + i) * a.ldy + co] = x;
          else Y[(size_t)co * a.ldy + n + i] = x;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// cn_groupnorm: one thread per (object, point); channels strided by ld (coalesced across points)
// ------------------------------------------------------------------------------------------------
// CG > 0: channels per group known at compile time -> the group is read ONCE into registers (the generic CG = 0 path reads it
// three times: sum, centred squares, normalise); same operations in the same order, bit-identical results.
template <int CG>
__global__ void __launch_bounds__(128) cn_groupnorm_kernel(const pcreid_norm_args a) {
  // one thread per (object, point), flattened so that objects with few points still fill the CTAs
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)a.B * a.rows) return;
  const int b = (int)(idx / a.rows), n = (int)(idx % a.rows);
  const float* X = a.X + (size_t)b * a.x_bs + n;
  const float* R = a.R ? a.R + (size_t)(a.r_map ? a.r_map[b] : b) * a.r_bs + n : nullptr;
  float* Y = a.Y + (size_t)b * a.y_bs + n;
  const int cg = CG > 0 ? CG : a.C / a.G;
  for (int g = 0; g < a.G; ++g) {
    if (CG > 0) {
      float x[CG > 0 ? CG : 1];
#pragma unroll
      for (int i = 0; i < CG; ++i) x[i] = X[(size_t)(g * CG + i) * a.ldx];
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < CG; ++i) s += x[i];
      const float mean = s / (float)CG;
      float v = 0.f;
#pragma unroll
      for (int i = 0; i < CG; ++i) { const float d = x[i] - mean; v = fmaf(d, d, v); }
      const float rstd = rsqrtf(v / (float)CG + 1e-5f);
#pragma unroll
      for (int i = 0; i < CG; ++i) {
        const int c = g * CG + i;
        float y = (x[i] - mean) * rstd * __ldg(a.gamma + c) + __ldg(a.beta + c);
        if (R) y += R[(size_t)c * a.ldr];
        Y[(size_t)c * a.ldy] = apply_act(y, a.act);
      }
      continue;
    }
    float s = 0.f;
    for (int c = g * cg; c < (g + 1) * cg; ++c) s += X[(size_t)c * a.ldx];
    const float mean = s / (float)cg;
    float v = 0.f;
    for (int c = g * cg; c < (g + 1) * cg; ++c) {
      float d = X[(size_t)c * a.ldx] - mean;
      v = fmaf(d, d, v);
    }
    const float rstd = rsqrtf(v / (float)cg + 1e-5f);
    for (int c = g * cg; c < (g + 1) * cg; ++c) {
      float y = (X[(size_t)c * a.ldx] - mean) * rstd * __ldg(a.gamma + c) + __ldg(a.beta + c);
      if (R) y += R[(size_t)c * a.ldr];
      Y[(size_t)c * a.ldy] = apply_act(y, a.act);
    }
  }
}

// Tail of the match head: GroupNorm + shortcut + ReLU + dot with the final Linear's weight row, one thread per pair (column);
// the same operations in the same order as cn_groupnorm_kernel followed by a sequential fma chain over the channels.
template <int CG>
__global__ void __launch_bounds__(128) gn_res_relu_dot_kernel(int rows, int C, int G, const float* __restrict__ X, int ldx,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              const float* __restrict__ R, int ldr, const float* __restrict__ w,
                                                              float bias, float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= rows) return;
  const int cg = CG > 0 ? CG : C / G;
  float acc = 0.f;
  for (int g = 0; g < G; ++g) {
    if (CG > 0) {
      float x[CG > 0 ? CG : 1];
#pragma unroll
      for (int i = 0; i < CG; ++i) x[i] = X[(size_t)(g * CG + i) * ldx + n];
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < CG; ++i) s += x[i];
      const float mean = s / (float)CG;
      float v = 0.f;
#pragma unroll
      for (int i = 0; i < CG; ++i) { const float d = x[i] - mean; v = fmaf(d, d, v); }
      const float rstd = rsqrtf(v / (float)CG + 1e-5f);
#pragma unroll
      for (int i = 0; i < CG; ++i) {
        const int c = g * CG + i;
        const float y = (x[i] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c) + R[(size_t)c * ldr + n];
        acc = fmaf(fmaxf(y, 0.f), __ldg(w + c), acc);
      }
      continue;
    }
    float s = 0.f;
    for (int c = g * cg; c < (g + 1) * cg; ++c) s += X[(size_t)c * ldx + n];
    const float mean = s / (float)cg;
    float v = 0.f;
    for (int c = g * cg; c < (g + 1) * cg; ++c) { const float d = X[(size_t)c * ldx + n] - mean; v = fmaf(d, d, v); }
    const float rstd = rsqrtf(v / (float)cg + 1e-5f);
    for (int c = g * cg; c < (g + 1) * cg; ++c) {
      const float y = (X[(size_t)c * ldx + n] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c) + R[(size_t)c * ldr + n];
      acc = fmaf(fmaxf(y, 0.f), __ldg(w + c), acc);
    }
  }
  out[n] = acc + bias;
}

// ------------------------------------------------------------------------------------------------
// linear attention reductions
// ------------------------------------------------------------------------------------------------
constexpr int KV_SC_DEFAULT = 128;   // points per staged chunk (head dim <= 64)
// grid (H, B); Wkv[b][(h*D+i)*d + h*D+j] = sum_s elu1(k[h*D+i][s]) * (v[h*D+j][s] / S); D <= 64
// head dims above 64 (the mul=2 / mul=4 model variants: D up to 256) split the D x D outputs of a head over blockIdx.z in
// blocks of 4096 and stage shorter point chunks (KV_SC_ template parameter) so that the tiles still fit shared memory
template <int KV_SC>
__global__ void __launch_bounds__(256) linattn_kv_kernel(int S, int d, int H, const float* __restrict__ K, long long k_bs,
                                                         int ldk, const float* __restrict__ V, long long v_bs, int ldv,
                                                         float* __restrict__ Wkv, float* __restrict__ ksum) {
  extern __shared__ float sm[];
  const int D = d / H, h = blockIdx.x, b = blockIdx.y, e0 = blockIdx.z * 4096;
  float* Ks = sm;                        // [D][KV_SC+1]
  float* Vs = sm + D * (KV_SC + 1);      // [D][KV_SC+1]
  const float* Kb = K + (size_t)b * k_bs + (size_t)h * D * ldk;
  const float* Vb = V + (size_t)b * v_bs + (size_t)h * D * ldv;
  const int nout = D * D;
  float acc[16];
#pragma unroll
  for (int o = 0; o < 16; ++o) acc[o] = 0.f;
  float ks = 0.f;
  const float invS = (float)S;
  for (int s0 = 0; s0 < S; s0 += KV_SC) {
    const int sc = min(KV_SC, S - s0);
    for (int i = threadIdx.x; i < D * KV_SC; i += blockDim.x) {
      int c = i / KV_SC, s = i % KV_SC;
      float kv = 0.f, vv = 0.f;
      if (s < sc) {
        kv = apply_act(Kb[(size_t)c * ldk + s0 + s], ACT_ELU1);
        vv = Vb[(size_t)c * ldv + s0 + s] / invS;
      }
      Ks[c * (KV_SC + 1) + s] = kv;
      Vs[c * (KV_SC + 1) + s] = vv;
    }
    __syncthreads();
#pragma unroll
    for (int o = 0; o < 16; ++o) {
      int e = e0 + threadIdx.x + o * 256;
      if (e < nout) {
        const float* kr = Ks + (e / D) * (KV_SC + 1);
        const float* vr = Vs + (e % D) * (KV_SC + 1);
        float s = acc[o];
        for (int t = 0; t < sc; ++t) s = fmaf(kr[t], vr[t], s);
        acc[o] = s;
      }
    }
    if (blockIdx.z == 0 && threadIdx.x < D) {
      const float* kr = Ks + threadIdx.x * (KV_SC + 1);
      for (int t = 0; t < sc; ++t) ks += kr[t];
    }
    __syncthreads();
  }
  float* Wb = Wkv + (size_t)b * d * d;
#pragma unroll
  for (int o = 0; o < 16; ++o) {
    int e = e0 + threadIdx.x + o * 256;
    if (e < nout) Wb[(size_t)(h * D + e / D) * d + h * D + (e % D)] = acc[o];
  }
  if (blockIdx.z == 0 && threadIdx.x < D) ksum[(size_t)b * d + h * D + threadIdx.x] = ks;
}

__global__ void __launch_bounds__(128) linattn_scale_kernel(int rows, int d, int H, int S, const float* __restrict__ Q,
                                                            long long q_bs, int ldq, const int* __restrict__ q_map,
                                                            const float* __restrict__ ksum, const int* __restrict__ ksum_map,
                                                            float* __restrict__ Qs, long long qs_bs, int ldqs) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (n >= rows) return;
  const float* q = Q + (size_t)(q_map ? q_map[b] : b) * q_bs + n;
  const float* ks = ksum + (size_t)(ksum_map ? ksum_map[b] : b) * d;
  float* o = Qs + (size_t)b * qs_bs + n;
  const int D = d / H;
  for (int h = 0; h < H; ++h) {
    float dot = 0.f;
    for (int c = h * D; c < (h + 1) * D; ++c) dot = fmaf(apply_act(q[(size_t)c * ldq], ACT_ELU1), __ldg(ks + c), dot);
    const float z = (1.f / (dot + 1e-6f)) * (float)S;
    for (int c = h * D; c < (h + 1) * D; ++c) o[(size_t)c * ldqs] = apply_act(q[(size_t)c * ldq], ACT_ELU1) * z;
  }
}

// ------------------------------------------------------------------------------------------------
// local_self_attention (mmdet3d/models/attention.py:221-296): every point attends its knum nearest neighbours in
// feature space with the linear-attention kernel, one query per point:
//   w_jh = (elu(q_h)+1) . (elu(k_jh)+1),   out_h = sum_j w_jh v_jh / (sum_j w_jh + 1e-6)
// (LinearAttention with L = 1, S = knum: the reference's v / S and x S cancel).  One warp per point; lanes own C/32
// contiguous channels, so a head is a contiguous lane range and its dot products reduce with xor shuffles.
// qkv point-major (B, N, 3C) rows [q | k | v] (pre-activation), idx (B, N, knum), out point-major (B, N, C).
// ------------------------------------------------------------------------------------------------
template <int CPL>
__global__ void __launch_bounds__(256) local_linattn_kernel(long long rows, int N, int C, int H, int knum, const float* __restrict__ qkv,
                                                            const int* __restrict__ idx, float* __restrict__ out) {
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const long long b = r / N;
  const float* base = qkv + (size_t)b * N * 3 * C;
  const float* qrow = qkv + (size_t)r * 3 * C + lane * CPL;
  const int lph = 32 / H;                                    // lanes per head
  float qf[CPL], acc[CPL];
#pragma unroll
  for (int i = 0; i < CPL; ++i) { qf[i] = apply_act(qrow[i], ACT_ELU1); acc[i] = 0.f; }
  float wsum = 0.f;
  const int* ir = idx + (size_t)r * knum;
  for (int j = 0; j < knum; ++j) {
    const float* nrow = base + (size_t)__ldg(ir + j) * 3 * C + lane * CPL;
    float w = 0.f;
#pragma unroll
    for (int i = 0; i < CPL; ++i) w = fmaf(qf[i], apply_act(__ldg(nrow + C + i), ACT_ELU1), w);
    for (int o = lph >> 1; o > 0; o >>= 1) w += __shfl_xor_sync(FULL_MASK, w, o);
    wsum += w;
#pragma unroll
    for (int i = 0; i < CPL; ++i) acc[i] = fmaf(w, __ldg(nrow + 2 * C + i), acc[i]);
  }
  const float z = 1.f / (wsum + 1e-6f);
  float* o = out + (size_t)r * C + lane * CPL;
#pragma unroll
  for (int i = 0; i < CPL; ++i) o[i] = acc[i] * z;
}

// ------------------------------------------------------------------------------------------------
// unfused SA edge path for channel counts above the fused kernel's tile (C > 128: the mul=2 / mul=4 model variants):
//   edge_build : H1[b,c,s*k+j] = relu(P1[b,c,idx[b,s,j]] + Cc[b,c,s])       (then two cn_linear calls)
//   seg_max    : out[b,c,s]    = max_j X[b,c,s*k+j]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) edge_build_kernel(int C, int N, int S, int k, const float* __restrict__ P1,
                                                         const float* __restrict__ Cc, const int* __restrict__ idx,
                                                         float* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.z;
  const int E = S * k;
  if (e >= E) return;
  const int s = e / k, src = idx[(size_t)b * E + e];
  const int c0 = blockIdx.y * 16;
#pragma unroll 4
  for (int c = c0; c < min(c0 + 16, C); ++c)
    out[((size_t)b * C + c) * E + e] = fmaxf(__ldg(P1 + ((size_t)b * C + c) * N + src) + __ldg(Cc + ((size_t)b * C + c) * S + s), 0.f);
}

__global__ void __launch_bounds__(256) seg_max_kernel(long long rows, int k, const float* __restrict__ x, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  const float* r = x + i * k;
  float mx = r[0];
  for (int j = 1; j < k; ++j) mx = fmaxf(mx, r[j]);
  out[i] = mx;
}

// pool_mod='avg' of the mmdet3d SA modules (point_sa_module.py:158-160): mean over the k samples of a group
__global__ void __launch_bounds__(256) seg_mean_kernel(long long rows, int k, const float* __restrict__ x, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  const float* r = x + i * k;
  float s = r[0];
  for (int j = 1; j < k; ++j) s += r[j];
  out[i] = s / (float)k;
}

// ------------------------------------------------------------------------------------------------
// pooling
// ------------------------------------------------------------------------------------------------
// one warp per (b, c)
__global__ void __launch_bounds__(256) cn_pool_kernel(int B, int C, int rows1, const float* __restrict__ X1, long long x1_bs,
                                                      int ldx1, int rows2, const float* __restrict__ X2, long long x2_bs,
                                                      int ldx2, int mode, float* __restrict__ out, long long ob, long long oc) {
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (w >= B * C) return;
  const int b = w / C, c = w % C;
  float mx = -INFINITY, sm = 0.f;
  const float* r1 = X1 + (size_t)b * x1_bs + (size_t)c * ldx1;
  for (int n = lane; n < rows1; n += 32) { float v = r1[n]; mx = fmaxf(mx, v); sm += v; }
  if (X2) {
    const float* r2 = X2 + (size_t)b * x2_bs + (size_t)c * ldx2;
    for (int n = lane; n < rows2; n += 32) { float v = r2[n]; mx = fmaxf(mx, v); sm += v; }
  }
  mx = warp_max(mx);
  sm = warp_sum(sm);
  if (lane == 0) {
    out[(size_t)b * ob + (size_t)c * oc] = mx;
    if (mode == 0) out[(size_t)b * ob + (size_t)(C + c) * oc] = sm / (float)(rows1 + (X2 ? rows2 : 0));
  }
}

__global__ void __launch_bounds__(128) cn_chanmax_kernel(int C, int rows, const float* __restrict__ X, long long x_bs, int ldx,
                                                         float* __restrict__ out, long long ob, long long on) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (n >= rows) return;
  const float* x = X + (size_t)b * x_bs + n;
  float mx = -INFINITY;
  for (int c = 0; c < C; ++c) mx = fmaxf(mx, x[(size_t)c * ldx]);
  out[(size_t)b * ob + (size_t)n * on] = mx;
}

// ------------------------------------------------------------------------------------------------
// fused SA edge MLP: gather + relu(P1+Cc) -> conv2 -> relu -> conv3 -> relu -> max over k
// ------------------------------------------------------------------------------------------------
// one CTA per (object, group of CPT centres); edges of the group are the 128-row GEMM tile.
template <int TN>
__global__ void __launch_bounds__(NTHR) sa_edge_mlp_kernel(int C, int N, int S, int k, int cpt, const float* __restrict__ P1,
                                                           const float* __restrict__ Cc, const int* __restrict__ idx,
                                                           const float* __restrict__ W2, const float* __restrict__ b2,
                                                           const float* __restrict__ W3, const float* __restrict__ b3,
                                                           float* __restrict__ out) {
  constexpr int TC = TileCols<TN>::value;
  extern __shared__ __align__(16) float smf[];
  float* A = smf;                     // [C][TR]
  float* Bm = A + (size_t)C * TR;     // [C][TR]
  float* Ws = Bm + (size_t)C * TR;    // [2][KC][TC]
  const int b = blockIdx.y, s0 = blockIdx.x * cpt;
  const int ncen = min(cpt, S - s0), nedge = ncen * k;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float* Pb = P1 + (size_t)b * C * N;
  const float* Cb = Cc + (size_t)b * C * S;

  // build the h1 tile
  {
    const int e = threadIdx.x & (TR - 1);
    int src = -1, cen = 0;
    if (e < nedge) { cen = s0 + e / k; src = idx[((size_t)b * S + cen) * k + (e % k)]; }
    for (int c = threadIdx.x / TR; c < C; c += NTHR / TR) {
      float v = 0.f;
      if (src >= 0) v = fmaxf(__ldg(Pb + (size_t)c * N + src) + __ldg(Cb + (size_t)c * S + cen), 0.f);
      A[c * TR + e] = v;
    }
  }
  const bool vw = (C % 4 == 0) && aligned16(W2) && aligned16(W3);
  const int nch = ceil_div(C, KC);
  for (int layer = 0; layer < 2; ++layer) {
    const float* W = layer == 0 ? W2 : W3;
    const float* bias = layer == 0 ? b2 : b3;
    const float* Xin = layer == 0 ? A : Bm;
    float* Xout = layer == 0 ? Bm : A;
    for (int co0 = 0; co0 < C; co0 += TC) {
      float acc[8][TN];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
      __syncthreads();   // Xin complete / previous Ws consumers done
      load_w_chunk<TN>(Ws, W, C, C, 0, co0, vw);
      cp_async_commit();
      for (int ch = 0; ch < nch; ++ch) {
        if (ch + 1 < nch) load_w_chunk<TN>(Ws + ((ch + 1) & 1) * KC * TC, W, C, C, (ch + 1) * KC, co0, vw);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        fma_chunk<TN>(Xin + (size_t)ch * KC * TR, Ws + (ch & 1) * KC * TC, tx, ty, acc);
        __syncthreads();
      }
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int co = co0 + col_of<TN>(ty, j);
        if (co >= C) continue;
        const float bv = __ldg(bias + co);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float4 v;
          v.x = fmaxf(acc[h * 4 + 0][j] + bv, 0.f);
          v.y = fmaxf(acc[h * 4 + 1][j] + bv, 0.f);
          v.z = fmaxf(acc[h * 4 + 2][j] + bv, 0.f);
          v.w = fmaxf(acc[h * 4 + 3][j] + bv, 0.f);
          *reinterpret_cast<float4*>(Xout + (size_t)co * TR + h * 64 + tx * 4) = v;
        }
      }
    }
  }
  __syncthreads();
  // max over the k edges of each centre (result of layer 3 lives in A)
  for (int i = threadIdx.x; i < C * ncen; i += NTHR) {
    const int c = i / ncen, cl = i % ncen;
    const float* r = A + (size_t)c * TR + cl * k;
    float mx = r[0];
    for (int j = 1; j < k; ++j) mx = fmaxf(mx, r[j]);
    out[((size_t)b * C + c) * S + s0 + cl] = mx;
  }
}

// EdgeConv: out[b,c,i] = act(max_j P[b,c,idx[b,i,j]] + Q[b,c,i])
__global__ void __launch_bounds__(256) edge_gather_max_kernel(int C, int N, int k, const float* __restrict__ P,
                                                              const float* __restrict__ Q, const int* __restrict__ idx,
                                                              int act, float* __restrict__ out, long long o_bs, int ldo) {
  extern __shared__ int sidx[];   // [128][k]
  const int b = blockIdx.y, i0 = blockIdx.x * 128;
  const int ni = min(128, N - i0);
  for (int t = threadIdx.x; t < ni * k; t += blockDim.x) sidx[t] = idx[((size_t)b * N + i0) * k + t];
  __syncthreads();
  const int il = threadIdx.x & 127;
  if (il >= ni) return;
  const int* id = sidx + il * k;
  for (int c = threadIdx.x >> 7; c < C; c += 2) {
    const float* row = P + ((size_t)b * C + c) * N;
    float mx = -INFINITY;
    for (int j = 0; j < k; ++j) mx = fmaxf(mx, __ldg(row + id[j]));
    out[(size_t)b * o_bs + (size_t)c * ldo + i0 + il] = apply_act(mx + Q[((size_t)b * C + c) * N + i0 + il], act);
  }
}

// ------------------------------------------------------------------------------------------------
// fused all-pairs 'concat' head.  One warp per pair, a CTA handles 8 tracks x 8 detections... kept
// simple: CTA = (track t, 32 detections); W2 (Hd x Hd, k-major) streamed through shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int PH_DT = 32;   // detections per CTA
template <int HD, int CG>
__global__ void __launch_bounds__(256) pair_concat_head_kernel(int T, int D, int E, const float* __restrict__ A,
                                                               const float* __restrict__ Bv, const float* __restrict__ Et,
                                                               const float* __restrict__ Ed, const float* __restrict__ W2,
                                                               const float* __restrict__ g1, const float* __restrict__ be1,
                                                               const float* __restrict__ g2, const float* __restrict__ be2,
                                                               const float* __restrict__ w, float b0,
                                                               const unsigned char* __restrict__ mask, float* __restrict__ out) {
  // CTA = (track t, 32 detections); lane = pair, warp wq owns CPW = HD/8 consecutive channels (whole groups).
  constexpr int CPW = HD / 8;
  static_assert(CPW % CG == 0 && CPW % 4 == 0, "groups must not straddle warps");
  __shared__ float H1[HD][PH_DT + 1];
  __shared__ float red[8][PH_DT];
  const int t = blockIdx.y, d0 = blockIdx.x * PH_DT;
  const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
  const int d = d0 + lane;
  const bool valid = d < D;
  float x[CPW];
#pragma unroll
  for (int i = 0; i < CPW; ++i) {
    const int c = wq * CPW + i;
    x[i] = valid ? __ldg(A + (size_t)t * HD + c) + __ldg(Bv + (size_t)d * HD + c) : 0.f;
  }
#pragma unroll
  for (int g0 = 0; g0 < CPW; g0 += CG) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CG; ++i) s += x[g0 + i];
    const float mean = s / (float)CG;
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < CG; ++i) { float dd = x[g0 + i] - mean; v = fmaf(dd, dd, v); }
    const float rstd = rsqrtf(v / (float)CG + 1e-5f);
#pragma unroll
    for (int i = 0; i < CG; ++i) {
      const int c = wq * CPW + g0 + i;
      x[g0 + i] = fmaxf((x[g0 + i] - mean) * rstd * __ldg(g1 + c) + __ldg(be1 + c), 0.f);
    }
  }
#pragma unroll
  for (int i = 0; i < CPW; ++i) H1[wq * CPW + i][lane] = x[i];
  __syncthreads();
  float y[CPW];
#pragma unroll
  for (int i = 0; i < CPW; ++i) y[i] = 0.f;
  for (int kk = 0; kk < HD; ++kk) {
    const float hv = H1[kk][lane];
    const float4* wr = reinterpret_cast<const float4*>(W2 + (size_t)kk * HD + wq * CPW);
#pragma unroll
    for (int i = 0; i < CPW / 4; ++i) {
      const float4 w4 = __ldg(wr + i);
      y[4 * i + 0] = fmaf(hv, w4.x, y[4 * i + 0]);
      y[4 * i + 1] = fmaf(hv, w4.y, y[4 * i + 1]);
      y[4 * i + 2] = fmaf(hv, w4.z, y[4 * i + 2]);
      y[4 * i + 3] = fmaf(hv, w4.w, y[4 * i + 3]);
    }
  }
  float part = 0.f;
#pragma unroll
  for (int g0 = 0; g0 < CPW; g0 += CG) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CG; ++i) s += y[g0 + i];
    const float mean = s / (float)CG;
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < CG; ++i) { float dd = y[g0 + i] - mean; v = fmaf(dd, dd, v); }
    const float rstd = rsqrtf(v / (float)CG + 1e-5f);
#pragma unroll
    for (int i = 0; i < CG; ++i) {
      const int c = wq * CPW + g0 + i;
      float r = 0.f;
      if (valid) r = c < E ? __ldg(Et + (size_t)t * E + c) : __ldg(Ed + (size_t)d * E + (c - E));
      const float o = fmaxf((y[g0 + i] - mean) * rstd * __ldg(g2 + c) + __ldg(be2 + c) + r, 0.f);
      part = fmaf(o, __ldg(w + c), part);
    }
  }
  red[wq][lane] = part;
  __syncthreads();
  if (wq == 0 && valid) {
    float s = b0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][lane];
    if (mask && !mask[(size_t)t * D + d]) s = 0.f;
    out[(size_t)t * D + d] = s;
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int pcreid_abi_version(void) { return 1; }

int pcreid_cn_linear(const pcreid_linear_args* p, void* stream) {
  if (!p) return PCREID_ERR_ARG;
  pcreid_linear_args a = *p;
  if (a.B <= 0 || a.rows <= 0 || a.CO <= 0) return PCREID_OK;
  if (a.K1 <= 0 || !a.X1 || !a.W1 || !a.Y) return PCREID_ERR_ARG;
  if (a.K2 > 0 && (!a.X2 || !a.W2)) return PCREID_ERR_ARG;
  if (a.B > 65535 * 32) return PCREID_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  // grid.z is limited to 65535: fold objects in slices
  const int zmax = 65535;
  for (int b0 = 0; b0 < a.B; b0 += zmax) {
    pcreid_linear_args s = a;
    s.B = (a.B - b0 < zmax) ? a.B - b0 : zmax;
    if (b0) {
      if (s.x1_map) s.x1_map += b0; else s.X1 += (size_t)b0 * s.x1_bs;
      if (s.K2 > 0) { if (s.x2_map) s.x2_map += b0; else s.X2 += (size_t)b0 * s.x2_bs; s.W2 += (size_t)b0 * s.w2_bs; }
      if (s.w1_map) s.w1_map += b0; else s.W1 += (size_t)b0 * s.w1_bs;
      if (s.R) { if (s.r_map) s.r_map += b0; else s.R += (size_t)b0 * s.r_bs; }
      s.Y += (size_t)b0 * s.y_bs;
    }
    auto a16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    const bool pack = s.rows < TR && TR % s.rows == 0 && s.rows % 4 == 0 && s.B > 1 && !s.x1_map && !s.x2_map && !s.w1_map && !s.r_map &&
                      !s.x1_pm && !s.x2_pm && !s.y_pm && s.w1_bs == 0 && (s.K2 == 0 || s.w2_bs == 0) && s.CO % 4 == 0 &&
                      a16(s.X1) && s.ldx1 % 4 == 0 && s.x1_bs % 4 == 0 && a16(s.W1) && a16(s.Y) && s.ldy % 4 == 0 && s.y_bs % 4 == 0 &&
                      (s.K2 == 0 || (a16(s.X2) && s.ldx2 % 4 == 0 && s.x2_bs % 4 == 0 && a16(s.W2))) &&
                      (!s.R || (a16(s.R) && s.ldr % 4 == 0 && s.r_bs % 4 == 0));
    if (pack) {
      const int nz = ceil_div(s.B, TR / s.rows);
      if (a.CO > 64) cn_linear_kernel<8, true><<<dim3(1, ceil_div(a.CO, 128), nz), NTHR, 0, st>>>(s);
      else cn_linear_kernel<4, true><<<dim3(1, 1, nz), NTHR, 0, st>>>(s);
    } else if (a.CO > 64) {
      dim3 grid(ceil_div(a.rows, TR), ceil_div(a.CO, 128), s.B);
      cn_linear_kernel<8><<<grid, NTHR, 0, st>>>(s);
    } else {
      dim3 grid(ceil_div(a.rows, TR), 1, s.B);
      cn_linear_kernel<4><<<grid, NTHR, 0, st>>>(s);
    }
  }
  return pcreid_launch_status();
}

int pcreid_cn_groupnorm(const pcreid_norm_args* p, void* stream) {
  if (!p) return PCREID_ERR_ARG;
  pcreid_norm_args a = *p;
  if (a.B <= 0 || a.rows <= 0 || a.C <= 0) return PCREID_OK;
  if (!a.X || !a.Y || !a.gamma || !a.beta || a.G <= 0 || a.C % a.G) return PCREID_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int ymax = 65535;
  for (int b0 = 0; b0 < a.B; b0 += ymax) {
    pcreid_norm_args s = a;
    s.B = (a.B - b0 < ymax) ? a.B - b0 : ymax;
    s.X += (size_t)b0 * s.x_bs;
    s.Y += (size_t)b0 * s.y_bs;
    if (s.R) { if (s.r_map) s.r_map += b0; else s.R += (size_t)b0 * s.r_bs; }
    const unsigned grid = (unsigned)(((long long)s.B * a.rows + 127) / 128);
    switch (a.C / a.G) {
      case 4: cn_groupnorm_kernel<4><<<grid, 128, 0, st>>>(s); break;
      case 8: cn_groupnorm_kernel<8><<<grid, 128, 0, st>>>(s); break;
      case 16: cn_groupnorm_kernel<16><<<grid, 128, 0, st>>>(s); break;
      case 32: cn_groupnorm_kernel<32><<<grid, 128, 0, st>>>(s); break;
      case 64: cn_groupnorm_kernel<64><<<grid, 128, 0, st>>>(s); break;
      case 128: cn_groupnorm_kernel<128><<<grid, 128, 0, st>>>(s); break;
      default: cn_groupnorm_kernel<0><<<grid, 128, 0, st>>>(s); break;
    }
  }
  return pcreid_launch_status();
}

int pcreid_gn_res_relu_dot(int rows, int C, int G, const float* X, int ldx, const float* gamma, const float* beta, const float* R,
                           int ldr, const float* w, float bias, float* out, void* stream) {
  if (rows <= 0) return PCREID_OK;
  if (!X || !gamma || !beta || !R || !w || !out || C <= 0 || G <= 0 || C % G) return PCREID_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((rows + 127) / 128);
  switch (C / G) {
    case 8: gn_res_relu_dot_kernel<8><<<grid, 128, 0, st>>>(rows, C, G, X, ldx, gamma, beta, R, ldr, w, bias, out); break;
    case 16: gn_res_relu_dot_kernel<16><<<grid, 128, 0, st>>>(rows, C, G, X, ldx, gamma, beta, R, ldr, w, bias, out); break;
    case 32: gn_res_relu_dot_kernel<32><<<grid, 128, 0, st>>>(rows, C, G, X, ldx, gamma, beta, R, ldr, w, bias, out); break;
    default: gn_res_relu_dot_kernel<0><<<grid, 128, 0, st>>>(rows, C, G, X, ldx, gamma, beta, R, ldr, w, bias, out); break;
  }
  return pcreid_launch_status();
}

int pcreid_linattn_kv(int B, int S, int d, int H, const float* K, long long k_bs, int ldk, const float* V, long long v_bs,
                      int ldv, float* Wkv, float* ksum, void* stream) {
  if (B <= 0) return PCREID_OK;
  if (!K || !V || !Wkv || !ksum || S <= 0 || H <= 0 || d % H) return PCREID_ERR_ARG;
  const int D = d / H;
  if (D > 256) return PCREID_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(Wkv, 0, (size_t)B * d * d * sizeof(float), st);
  const int nz = ceil_div(D * D, 4096);
  for (int b0 = 0; b0 < B; b0 += 65535) {
    int nb = B - b0 < 65535 ? B - b0 : 65535;
    if (D <= 64) {
      size_t smem = (size_t)2 * D * (KV_SC_DEFAULT + 1) * sizeof(float);
      if (smem > 48 * 1024) cudaFuncSetAttribute(linattn_kv_kernel<KV_SC_DEFAULT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      linattn_kv_kernel<KV_SC_DEFAULT><<<dim3(H, nb, nz), 256, smem, st>>>(S, d, H, K + (size_t)b0 * k_bs, k_bs, ldk, V + (size_t)b0 * v_bs,
                                                                       v_bs, ldv, Wkv + (size_t)b0 * d * d, ksum + (size_t)b0 * d);
    } else {
      size_t smem = (size_t)2 * D * (32 + 1) * sizeof(float);
      if (smem > 48 * 1024) cudaFuncSetAttribute(linattn_kv_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      linattn_kv_kernel<32><<<dim3(H, nb, nz), 256, smem, st>>>(S, d, H, K + (size_t)b0 * k_bs, k_bs, ldk, V + (size_t)b0 * v_bs, v_bs, ldv,
                                                                Wkv + (size_t)b0 * d * d, ksum + (size_t)b0 * d);
    }
  }
  return pcreid_launch_status();
}

int pcreid_linattn_scale(int B, int rows, int d, int H, int S, const float* Q, long long q_bs, int ldq, const int* q_map,
                         const float* ksum, const int* ksum_map, float* Qs, long long qs_bs, int ldqs, void* stream) {
  if (B <= 0 || rows <= 0) return PCREID_OK;
  if (!Q || !ksum || !Qs || H <= 0 || d % H) return PCREID_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  for (int b0 = 0; b0 < B; b0 += 65535) {
    int nb = B - b0 < 65535 ? B - b0 : 65535;
    linattn_scale_kernel<<<dim3(ceil_div(rows, 128), nb), 128, 0, st>>>(
        rows, d, H, S, q_map ? Q : Q + (size_t)b0 * q_bs, q_bs, ldq, q_map ? q_map + b0 : nullptr,
        ksum_map ? ksum : ksum + (size_t)b0 * d, ksum_map ? ksum_map + b0 : nullptr, Qs + (size_t)b0 * qs_bs, qs_bs, ldqs);
  }
  return pcreid_launch_status();
}

int pcreid_local_linattn(int B, int N, int C, int H, int knum, const float* qkv, const int* idx, float* out, void* stream) {
  if (B <= 0 || N <= 0) return PCREID_OK;
  if (!qkv || !idx || !out || knum <= 0 || H <= 0) return PCREID_ERR_ARG;
  if (C % 32 || C > 128 || (H != 1 && H != 2 && H != 4) || C % H) return PCREID_ERR_UNSUPPORTED;
  const long long rows = (long long)B * N;
  const unsigned grid = (unsigned)((rows + 7) / 8);
  cudaStream_t st = (cudaStream_t)stream;
  switch (C / 32) {
    case 1: local_linattn_kernel<1><<<grid, 256, 0, st>>>(rows, N, C, H, knum, qkv, idx, out); break;
    case 2: local_linattn_kernel<2><<<grid, 256, 0, st>>>(rows, N, C, H, knum, qkv, idx, out); break;
    case 3: local_linattn_kernel<3><<<grid, 256, 0, st>>>(rows, N, C, H, knum, qkv, idx, out); break;
    default: local_linattn_kernel<4><<<grid, 256, 0, st>>>(rows, N, C, H, knum, qkv, idx, out); break;
  }
  return pcreid_launch_status();
}

int pcreid_cn_pool(int B, int C, int rows1, const float* X1, long long x1_bs, int ldx1, int rows2, const float* X2,
                   long long x2_bs, int ldx2, int mode, float* out, long long ob, long long oc, void* stream) {
  if (B <= 0 || C <= 0) return PCREID_OK;
  if (!X1 || !out || rows1 <= 0) return PCREID_ERR_ARG;
  long long nw = (long long)B * C;
  cn_pool_kernel<<<(unsigned)((nw + 7) / 8), 256, 0, (cudaStream_t)stream>>>(B, C, rows1, X1, x1_bs, ldx1, rows2, X2, x2_bs,
                                                                             ldx2, mode, out, ob, oc);
  return pcreid_launch_status();
}

int pcreid_cn_chanmax(int B, int C, int rows, const float* X, long long x_bs, int ldx, float* out, long long ob, long long on,
                      void* stream) {
  if (B <= 0 || rows <= 0) return PCREID_OK;
  if (!X || !out || C <= 0) return PCREID_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  for (int b0 = 0; b0 < B; b0 += 65535) {
    int nb = B - b0 < 65535 ? B - b0 : 65535;
    cn_chanmax_kernel<<<dim3(ceil_div(rows, 128), nb), 128, 0, st>>>(C, rows, X + (size_t)b0 * x_bs, x_bs, ldx,
                                                                     out + (size_t)b0 * ob, ob, on);
  }
  return pcreid_launch_status();
}

int pcreid_sa_edge_mlp(int B, int C, int N, int S, int k, const float* P1, const float* Cc, const int* idx, const float* W2,
                       const float* b2, const float* W3, const float* b3, float* out, void* stream) {
  if (B <= 0 || S <= 0) return PCREID_OK;
  if (!P1 || !Cc || !idx || !W2 || !b2 || !W3 || !b3 || !out || k <= 0 || N <= 0) return PCREID_ERR_ARG;
  if (k > TR || C > 128 || C % 16) return PCREID_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int cpt = TR / k;
  if (B > 65535) return PCREID_ERR_UNSUPPORTED;
  dim3 grid(ceil_div(S, cpt), B);
  if (C > 64) {
    size_t smem = ((size_t)2 * C * TR + 2 * KC * 128) * sizeof(float);
    cudaFuncSetAttribute(sa_edge_mlp_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sa_edge_mlp_kernel<8><<<grid, NTHR, smem, st>>>(C, N, S, k, cpt, P1, Cc, idx, W2, b2, W3, b3, out);
  } else {
    size_t smem = ((size_t)2 * C * TR + 2 * KC * 64) * sizeof(float);
    cudaFuncSetAttribute(sa_edge_mlp_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sa_edge_mlp_kernel<4><<<grid, NTHR, smem, st>>>(C, N, S, k, cpt, P1, Cc, idx, W2, b2, W3, b3, out);
  }
  return pcreid_launch_status();
}

int pcreid_edge_build(int B, int C, int N, int S, int k, const float* P1, const float* Cc, const int* idx, float* out, void* stream) {
  if (B <= 0 || S <= 0) return PCREID_OK;
  if (!P1 || !Cc || !idx || !out || C <= 0 || k <= 0 || N <= 0) return PCREID_ERR_ARG;
  if (B > 65535 || ceil_div(C, 16) > 65535) return PCREID_ERR_UNSUPPORTED;
  edge_build_kernel<<<dim3(ceil_div(S * k, 256), ceil_div(C, 16), B), 256, 0, (cudaStream_t)stream>>>(C, N, S, k, P1, Cc, idx, out);
  return pcreid_launch_status();
}

int pcreid_seg_max(long long rows, int k, const float* x, float* out, void* stream) {
  if (rows <= 0) return PCREID_OK;
  if (!x || !out || k <= 0) return PCREID_ERR_ARG;
  seg_max_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rows, k, x, out);
  return pcreid_launch_status();
}

int pcreid_seg_mean(long long rows, int k, const float* x, float* out, void* stream) {
  if (rows <= 0) return PCREID_OK;
  if (!x || !out || k <= 0) return PCREID_ERR_ARG;
  seg_mean_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rows, k, x, out);
  return pcreid_launch_status();
}

int pcreid_edge_gather_max(int B, int C, int N, int k, const float* P, const float* Q, const int* idx, int act, float* out,
                           long long o_bs, int ldo, void* stream) {
  if (B <= 0 || N <= 0) return PCREID_OK;
  if (!P || !Q || !idx || !out || C <= 0 || k <= 0) return PCREID_ERR_ARG;
  if (B > 65535 || (size_t)128 * k * 4 > 48 * 1024) return PCREID_ERR_UNSUPPORTED;
  edge_gather_max_kernel<<<dim3(ceil_div(N, 128), B), 256, (size_t)128 * k * sizeof(int), (cudaStream_t)stream>>>(
      C, N, k, P, Q, idx, act, out, o_bs, ldo);
  return pcreid_launch_status();
}

int pcreid_pair_concat_head(int T, int D, int E, int G, const float* A, const float* Bv, const float* Et, const float* Ed,
                            const float* W2, const float* g1, const float* be1, const float* g2, const float* be2,
                            const float* w, float b0, const unsigned char* mask, float* out, void* stream) {
  if (T <= 0 || D <= 0) return PCREID_OK;
  if (!A || !Bv || !Et || !Ed || !W2 || !g1 || !be1 || !g2 || !be2 || !w || !out) return PCREID_ERR_ARG;
  const int HD = 2 * E;
  if (G <= 0 || HD % G) return PCREID_ERR_ARG;
  const int cg = HD / G;
  if (T > 65535) return PCREID_ERR_UNSUPPORTED;
  dim3 grid(ceil_div(D, PH_DT), T);
  cudaStream_t st = (cudaStream_t)stream;
  if ((reinterpret_cast<uintptr_t>(W2) & 15) != 0) return PCREID_ERR_ARG;
  if (HD == 256 && cg == 8)
    pair_concat_head_kernel<256, 8><<<grid, 256, 0, st>>>(T, D, E, A, Bv, Et, Ed, W2, g1, be1, g2, be2, w, b0, mask, out);
  else if (HD == 256 && cg == 16)
    pair_concat_head_kernel<256, 16><<<grid, 256, 0, st>>>(T, D, E, A, Bv, Et, Ed, W2, g1, be1, g2, be2, w, b0, mask, out);
  else if (HD == 128 && cg == 8)
    pair_concat_head_kernel<128, 8><<<grid, 256, 0, st>>>(T, D, E, A, Bv, Et, Ed, W2, g1, be1, g2, be2, w, b0, mask, out);
  else if (HD == 128 && cg == 16)
    pair_concat_head_kernel<128, 16><<<grid, 256, 0, st>>>(T, D, E, A, Bv, Et, Ed, W2, g1, be1, g2, be2, w, b0, mask, out);
  else
    return PCREID_ERR_UNSUPPORTED;
  return pcreid_launch_status();
}

}  // extern "C"
