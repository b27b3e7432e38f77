// Linear-attention blocks of the 'Point Transformer' encoder on the 5th-gen tensor cores (tcgen05 kind::tf32, fp32
// accumulation in TMEM) -- the "fast" mode counterpart of the cn_linear / cn_groupnorm / linattn_scale chains that
// models/pointnet2_utils.py issues for Self_Attention and FP_SA (mmdet3d/models/pointnet2_utils.py:55-114, 362-437).
//
// One block = three launches:
//
//   attn_front : key-side rows.  pos = Wp2 relu(Wp0 xyz + bp0) + bp2 ;  fp = feat + pos ;
//                out[:, 0:NFP] = Wfp fp ; out[:, NFP:NFP+NF] = Wf feat            (q|k|v of Self_Attention, v|k of FP_SA)
//                Rows are independent, so a tile is 128 consecutive rows of the flattened (object, point) axis.
//   kv_img     : per object and head  KV_h = sum_s (elu(k_s)+1) (x) v_s  written as a ready-to-use tcgen05 operand
//                image, Ksum = sum_s (elu(k_s)+1)   (register-tiled FFMA reduction over the points)
//   attn_back  : query-side rows.  q (given, or Wq feat1) -> (elu+1) -> . KV_h per head -> x 1/(Q.Ksum) -> merge ->
//                LayerNorm1 -> relu(W0 [feat1 ; msg]) -> W2 -> LayerNorm2 (+ feat1) -> out
//                (the reference's v / S and x S cancel and are not applied)
//
// An attn_back CTA owns 128 rows (thread == row == TMEM lane): up to 128 rows of one object, or 2 / 4 whole objects
// of 64 / 32 rows (the per-object KV operands then go to separate TMEM column blocks and a row reads its own block).  Activations live in shared memory as fp32
// K-major operand images [k/4][128][4] (the no-swizzle canonical layout, read by the tensor core as tf32) and never
// travel to HBM between the stages of a block.  Weights are prepared once on the host as operand images
// [k/4][n][4], concatenated in the order the kernel consumes them, and streamed global -> shared memory through a
// three-slot ring of 1-D bulk TMA copies (cp.async.bulk + mbarrier complete_tx) issued by the thread that also issues
// the MMAs; slots are released by tcgen05.commit, so the copies of the next stage's weights fly during the current
// stage's epilogue.  All shapes are run-time parameters (multiples of 16 output channels / 8 input channels, at most
// 128 model channels); residency is bounded by shared memory (40 ... 224 KB per CTA).
#include "../../include/pcreid.h"
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int NTH = 128;
constexpr int RING = 3;
constexpr int MAXCH = 56;
constexpr int MAX_SLOT = 16384;

struct ChunkTab {
  int n;
  int slot_bytes;
  uint32_t off[MAXCH];      // byte offset in the weight blob, or in this object's M image
  uint32_t bytes[MAXCH];
  uint8_t per_obj[MAXCH];
};

__device__ __forceinline__ void expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}

// state of the weight pipeline; lives in the registers of thread 0
struct Pipe {
  uint64_t *full, *empty;
  uint32_t ring_s;
  const uint8_t *blob, *obj;
  const ChunkTab* tab;
  int loaded, used;
};

__device__ __forceinline__ void pipe_prefetch(Pipe& p) {
  while (p.loaded < p.tab->n && p.loaded < p.used + RING) {
    const int s = p.loaded % RING, n = p.loaded / RING;
    if (n > 0) tc::mbar_wait(&p.empty[s], (uint32_t)((n - 1) & 1));      // MMAs of the slot's previous tenant are done
    const uint8_t* src = (p.tab->per_obj[p.loaded] ? p.obj : p.blob) + p.tab->off[p.loaded];
    const uint32_t nb = p.tab->bytes[p.loaded];
    expect_tx(&p.full[s], nb);
    bulk_g2s(p.ring_s + (uint32_t)(s * p.tab->slot_bytes), src, nb, &p.full[s]);
    ++p.loaded;
  }
}

// thread 0:  D[:, 0:N) (+)= A_img[:, 0:K) . W^T   with W's image chunks arriving through the ring (tmem_d = first column)
__device__ __forceinline__ void pipe_gemm(Pipe& p, uint32_t tmem_d, uint32_t a_img_s, int K, int N, bool accumulate) {
  int kk = 0;
  while (kk < K) {
    pipe_prefetch(p);
    const int s = p.used % RING, n = p.used / RING;
    const int kc = (int)(p.tab->bytes[p.used] / (uint32_t)(N * 4));
    tc::mbar_wait(&p.full[s], (uint32_t)(n & 1));
    tc::tc_fence_after();
    const uint32_t slot = p.ring_s + (uint32_t)(s * p.tab->slot_bytes);
    for (int n0 = 0; n0 < N; n0 += 256) {
      const int nb = min(256, N - n0);
      const uint32_t idesc = tc::instr_desc(128, nb, tc::FMT_TF32, tc::MAJOR_K, tc::MAJOR_K);
      for (int ks = 0; ks < kc / 8; ++ks) {
        const uint64_t ad = tc::smem_desc(a_img_s + (uint32_t)(((kk >> 3) + ks) * 4096), 2048, 128, tc::LAYOUT_NONE);
        const uint64_t bd = tc::smem_desc(slot + (uint32_t)(ks * 2 * N * 16 + n0 * 16), (uint32_t)(N * 16), 128, tc::LAYOUT_NONE);
        tc::umma_tf32(tmem_d + (uint32_t)n0, ad, bd, idesc, (accumulate || kk > 0 || ks > 0) ? 1u : 0u);
      }
    }
    tc::umma_commit(&p.empty[s]);
    ++p.used;
    kk += kc;
  }
}

// all threads: operand writes visible to the tensor core, accumulator reads retired, then one barrier
__device__ __forceinline__ void stage_sync() {
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
}

__device__ __forceinline__ float elu1(float v) { return v > 0.f ? v + 1.f : __expf(v); }

// channel-major source -> operand image, zero outside (c < C, valid); Xr points at (channel 0, this thread's row), ld is the
// channel stride.  32 channels are fetched per batch so that the loads of a batch are all in flight before the first
// shared-memory store needs its data.
// RNA: the image is a pure MMA operand -> values are rounded to tf32 here (tc::tf32_rna) instead of truncated by the tensor core.
template <bool RNA>
__device__ __forceinline__ void load_image_cm(uint8_t* img, const float* __restrict__ Xr, int ld, int C, int CP, bool valid, int tid) {
  for (int c0 = 0; c0 < CP; c0 += 32) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = (valid && c0 + i < C) ? __ldg(Xr + (size_t)(c0 + i) * ld) : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (c0 + 4 * j < CP)
        *reinterpret_cast<float4*>(img + ((c0 >> 2) + j) * 2048 + tid * 16) =
            RNA ? tc::tf32_rna4(make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3])) : make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
}
// point-major source (x = this thread's row of C floats) -> operand image
template <bool RNA>
__device__ __forceinline__ void load_image_pm(uint8_t* img, const float* __restrict__ x, int C, int CP, bool valid, int tid) {
  for (int c0 = 0; c0 < CP; c0 += 32) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = (valid && c0 + i < C) ? __ldg(x + c0 + i) : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (c0 + 4 * j < CP)
        *reinterpret_cast<float4*>(img + ((c0 >> 2) + j) * 2048 + tid * 16) =
            RNA ? tc::tf32_rna4(make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3])) : make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
}

// LayerNorm statistics of this thread's row over TMEM columns [0, n): mean, 1/sqrt(var + 1e-5) (biased variance)
__device__ __forceinline__ void row_stats(uint32_t tl, int n, float& mean, float& rstd) {
  float s = 0.f;
  for (int q = 0; q < n / 16; ++q) {
    uint32_t r[16];
    tc::tmem_ld16(tl + 16 * q, r);
    tc::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) s += __uint_as_float(r[j]);
  }
  mean = s / (float)n;
  float v = 0.f;
  for (int q = 0; q < n / 16; ++q) {
    uint32_t r[16];
    tc::tmem_ld16(tl + 16 * q, r);
    tc::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float d = __uint_as_float(r[j]) - mean;
      v = fmaf(d, d, v);
    }
  }
  rstd = rsqrtf(v / (float)n + 1e-5f);
}

__host__ __device__ inline int tmem_cols(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}

// ------------------------------------------------------------------------------------------------------------------
struct FrontArgs {
  int B, S, C2, DP, NFP, NF;
  const float* xyz;
  const float* feat;
  long long f_bs;
  int ldf;
  const float *wp0, *bp0, *bp2;
  const uint8_t* blob;
  float* out;
  long long o_bs;
  int ldo;
  ChunkTab tab;
};

__global__ void __launch_bounds__(NTH) attn_front_kernel(const __grid_constant__ FrontArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full[RING], empty[RING], stagebar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const long long flat = (long long)blockIdx.x * 128 + tid;          // row of the flattened (object, point) axis
  const bool valid = flat < (long long)a.B * a.S;
  const int b = valid ? (int)(flat / a.S) : 0, row = valid ? (int)(flat % a.S) : 0;
  const int CA = a.DP > a.C2 ? a.DP : a.C2;
  uint8_t* imgF = smem;                        // feat                     [C2/4][128][4]
  uint8_t* imgA = imgF + a.C2 * 512;           // pos hidden, then feat+pos [CA/4][128][4]
  uint8_t* ring = imgA + CA * 512;
  float* wp0_s = reinterpret_cast<float*>(ring + RING * a.tab.slot_bytes);   // [3][DP]
  float* bp0_s = wp0_s + 3 * a.DP;
  float* bp2_s = bp0_s + a.DP;
  const int ncols = a.NFP + a.NF;
  if (tid == 0) {
    for (int i = 0; i < RING; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    tc::mbar_init(&stagebar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) { tc::tmem_alloc(&tmem_base_s, (uint32_t)tmem_cols(max(ncols, a.C2))); tc::tmem_relinquish(); }
  for (int i = tid; i < 3 * a.DP; i += NTH) wp0_s[i] = a.wp0[i];
  for (int i = tid; i < a.DP; i += NTH) bp0_s[i] = a.bp0[i];
  for (int i = tid; i < a.C2; i += NTH) bp2_s[i] = a.bp2[i];
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  Pipe p;
  if (tid == 0) {
    p.full = full; p.empty = empty; p.ring_s = tc::smem_u32(ring); p.blob = a.blob; p.obj = nullptr; p.tab = &a.tab;
    p.loaded = 0; p.used = 0;
    pipe_prefetch(p);
  }
  const uint32_t tmem = tmem_base_s;
  const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
  uint32_t par = 0;
  // ---- operand images: feat, relu(Wp0 xyz + bp0)
  load_image_cm<false>(imgF, a.feat + (size_t)b * a.f_bs + row, a.ldf, a.C2, a.C2, valid, tid);   // exact: feat + pos reads it back
  {
    float x = 0.f, y = 0.f, z = 0.f;
    if (valid) {
      const float* pp = a.xyz + ((size_t)b * a.S + row) * 3;
      x = __ldg(pp); y = __ldg(pp + 1); z = __ldg(pp + 2);
    }
    for (int c4 = 0; c4 < a.DP / 4; ++c4) {
      float h[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = 4 * c4 + j;
        h[j] = fmaxf(fmaf(x, wp0_s[c], fmaf(y, wp0_s[a.DP + c], fmaf(z, wp0_s[2 * a.DP + c], bp0_s[c]))), 0.f);
      }
      *reinterpret_cast<float4*>(imgA + c4 * 2048 + tid * 16) = tc::tf32_rna4(make_float4(h[0], h[1], h[2], h[3]));
    }
  }
  stage_sync();
  // ---- GEMM 1: pos = hid . Wp2^T
  if (tid == 0) {
    pipe_gemm(p, tmem, tc::smem_u32(imgA), a.DP, a.C2, false);
    tc::umma_commit(&stagebar);
    pipe_prefetch(p);
  }
  __syncwarp();
  tc::mbar_wait(&stagebar, par);
  par ^= 1u;
  tc::tc_fence_after();
  for (int q = 0; q < a.C2 / 16; ++q) {
    uint32_t r[16];
    tc::tmem_ld16(tl + 16 * q, r);
    tc::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const int c = 16 * q + j;
      const float4 f = *reinterpret_cast<const float4*>(imgF + (c / 4) * 2048 + tid * 16);
      float4 v;
      v.x = __uint_as_float(r[j + 0]) + bp2_s[c + 0] + f.x;
      v.y = __uint_as_float(r[j + 1]) + bp2_s[c + 1] + f.y;
      v.z = __uint_as_float(r[j + 2]) + bp2_s[c + 2] + f.z;
      v.w = __uint_as_float(r[j + 3]) + bp2_s[c + 3] + f.w;
      *reinterpret_cast<float4*>(imgA + (c / 4) * 2048 + tid * 16) = tc::tf32_rna4(v);
      // the exact feat value has been consumed: from here on imgF is only the A operand of the feat projections
      *reinterpret_cast<float4*>(imgF + (c / 4) * 2048 + tid * 16) = tc::tf32_rna4(f);
    }
  }
  stage_sync();
  // ---- GEMM 2: projections of feat+pos and of feat
  if (tid == 0) {
    pipe_gemm(p, tmem, tc::smem_u32(imgA), a.C2, a.NFP, false);
    if (a.NF > 0) pipe_gemm(p, tmem + (uint32_t)a.NFP, tc::smem_u32(imgF), a.C2, a.NF, false);
    tc::umma_commit(&stagebar);
  }
  __syncwarp();
  tc::mbar_wait(&stagebar, par);
  par ^= 1u;
  tc::tc_fence_after();
  float* O = a.out + (size_t)b * a.o_bs;
  for (int q = 0; q < ncols / 16; ++q) {
    uint32_t r[16];
    tc::tmem_ld16(tl + 16 * q, r);
    tc::tmem_ld_wait();
    if (valid) {
#pragma unroll
      for (int j = 0; j < 16; ++j) O[(size_t)(16 * q + j) * a.ldo + row] = __uint_as_float(r[j]);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, (uint32_t)tmem_cols(max(ncols, a.C2)));
}

// ------------------------------------------------------------------------------------------------------------------
// KV summaries of linear attention as tcgen05 operand images (one CTA per object):
//   kvimg[b][h][(i/4)*dh + j][i%4] = sum_s (elu(k[h*dh+i][s]) + 1) * v[h*dh+j][s]          (B operand: N = j, K = i)
//   ksum[b][c]                     = sum_s (elu(k[c][s]) + 1)
// Points stream through shared memory in chunks of 64; a thread owns a 4 x 4 block of one head's dh x dh outputs and
// reads 4 + 4 128-bit rows per 4 points (64 FMA per 8 LDS.128); when there are fewer blocks than threads the points of
// a chunk are split over thread groups and the partial sums meet in shared memory.
constexpr int KV_SC = 64, KV_SCP = KV_SC + 4;
template <int TPT>
__global__ void __launch_bounds__(256, 3) linattn_kv_img_kernel(int S, int d, int H, const float* __restrict__ K, long long k_bs, int ldk,
                                                             const float* __restrict__ V, long long v_bs, int ldv,
                                                             float* __restrict__ kvimg, float* __restrict__ ksum) {
  extern __shared__ __align__(16) float kvs[];
  float* Ks = kvs;                        // [d][KV_SCP]  elu(k)+1
  float* Vs = kvs + d * KV_SCP;           // [d][KV_SCP]
  const int b = blockIdx.x, tid = threadIdx.x, dh = d / H, nb = dh / 4;
  const int ntiles = H * nb * nb;         // 4 x 4 output blocks: 512 (d = 128), 128 (64), 32 (32)
  const int G = ntiles >= 256 ? 1 : 256 / ntiles;      // point groups
  const float* Kb = K + (size_t)b * k_bs;
  const float* Vb = V + (size_t)b * v_bs;
  float acc[TPT][4][4];                                // TPT = 4 x 4 blocks per thread
#pragma unroll
  for (int u = 0; u < TPT; ++u)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[u][i][j] = 0.f;
  float ks = 0.f;
  const int g = G > 1 ? tid / ntiles : 0;
  for (int s0 = 0; s0 < S; s0 += KV_SC) {
    for (int i0 = tid; i0 < d * KV_SC; i0 += 256 * 8) {      // 16 loads in flight per thread before the first store
      float kq[8], vq[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int i = i0 + e * 256, c = i / KV_SC, s = i % KV_SC;
        const bool in = i < d * KV_SC && s0 + s < S;
        kq[e] = in ? __ldg(Kb + (size_t)c * ldk + s0 + s) : 0.f;
        vq[e] = in ? __ldg(Vb + (size_t)c * ldv + s0 + s) : 0.f;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int i = i0 + e * 256, c = i / KV_SC, s = i % KV_SC;
        if (i < d * KV_SC) {
          Ks[c * KV_SCP + s] = s0 + s < S ? elu1(kq[e]) : 0.f;
          Vs[c * KV_SCP + s] = vq[e];
        }
      }
    }
    __syncthreads();
    if (tid < d) {
      const float4* kr = reinterpret_cast<const float4*>(Ks + tid * KV_SCP);
#pragma unroll 4
      for (int s = 0; s < KV_SC / 4; ++s) { const float4 x = kr[s]; ks += (x.x + x.y) + (x.z + x.w); }
    }
#pragma unroll
    for (int u = 0; u < TPT; ++u) {
      const int tile = (G > 1 ? tid % ntiles : tid) + u * 256;
      if (tile < ntiles) {
        const int h = tile / (nb * nb), ti = (tile % (nb * nb)) / nb, tj = tile % nb;
        const float* kr = Ks + (h * dh + 4 * ti) * KV_SCP;
        const float* vr = Vs + (h * dh + 4 * tj) * KV_SCP;
        for (int s = 4 * g; s < KV_SC; s += 4 * G) {
          float4 kq[4], vq[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) kq[i] = *reinterpret_cast<const float4*>(kr + i * KV_SCP + s);
#pragma unroll
          for (int j = 0; j < 4; ++j) vq[j] = *reinterpret_cast<const float4*>(vr + j * KV_SCP + s);
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
              acc[u][i][j] = fmaf(kq[i].x, vq[j].x, fmaf(kq[i].y, vq[j].y, fmaf(kq[i].z, vq[j].z, fmaf(kq[i].w, vq[j].w, acc[u][i][j]))));
        }
      }
    }
    __syncthreads();
  }
  if (tid < d) ksum[(size_t)b * d + tid] = ks;
  if (G > 1) {
    // partial sums of the point groups meet in shared memory: red[g][tile][16]
    float* red = kvs;
    const int tile = tid % ntiles;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4*>(red + ((size_t)(g * ntiles + tile) * 16) + 4 * i) = make_float4(acc[0][i][0], acc[0][i][1], acc[0][i][2], acc[0][i][3]);
    __syncthreads();
    if (g == 0) {
      for (int gg = 1; gg < G; ++gg)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 x = *reinterpret_cast<const float4*>(red + ((size_t)(gg * ntiles + tile) * 16) + 4 * i);
          acc[0][i][0] += x.x; acc[0][i][1] += x.y; acc[0][i][2] += x.z; acc[0][i][3] += x.w;
        }
    }
  }
  float* img = kvimg + (size_t)b * d * dh;
#pragma unroll
  for (int u = 0; u < TPT; ++u) {
    const int tile = (G > 1 ? tid % ntiles : tid) + u * 256;
    if (tile < ntiles && g == 0) {
      const int h = tile / (nb * nb), ti = (tile % (nb * nb)) / nb, tj = tile % nb;
      float* ih = img + (size_t)h * dh * dh;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<float4*>(ih + ((size_t)ti * dh + 4 * tj + j) * 4) = tc::tf32_rna4(make_float4(acc[u][0][j], acc[u][1][j], acc[u][2][j], acc[u][3][j]));
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
struct BackArgs {
  int B, rows, D, H, C1, C1P, CO, opt;       // opt objects of `rows` rows per tile (opt * rows == 128) or opt == 1
  int qpre, res, f1_pm;
  const float* feat1;
  long long f1_bs;
  int ldf1;
  const float* q;
  long long q_bs;
  int ldq;
  const float* ksum;
  const uint8_t* kvimg;
  const float *g1, *b1, *g2, *b2;
  const uint8_t* blob;
  float* out;
  long long o_bs;
  int ldo;
  ChunkTab tab;
};

__global__ void __launch_bounds__(NTH) attn_back_kernel(const __grid_constant__ BackArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full[RING], empty[RING], stagebar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int D = a.D, D2 = 2 * a.D, dh = a.D / a.H;
  int b0, ob, row;
  bool valid;
  if (a.opt > 1) {
    b0 = blockIdx.x * a.opt; ob = tid / a.rows; row = tid % a.rows;
    valid = b0 + ob < a.B;
  } else {
    const int tpo = (a.rows + 127) / 128;
    b0 = blockIdx.x / tpo; ob = 0; row = (blockIdx.x % tpo) * 128 + tid;
    valid = row < a.rows;
  }
  const int b = valid ? b0 + ob : b0;
  const int img_ch = (a.C1P + D) > D2 ? (a.C1P + D) : D2;
  uint8_t* imgF = smem;                       // feat1              [C1P/4][128][4]
  uint8_t* imgQ = imgF + a.C1P * 512;         // elu(q)+1, then the attention output, then msg [D/4][128][4]
  uint8_t* imgH = smem;                       // mlp hidden         [2D/4][128][4]  (overlays imgF | imgQ once both are consumed)
  uint8_t* ring = smem + img_ch * 512;
  float* ksum_s = reinterpret_cast<float*>(ring + RING * a.tab.slot_bytes);    // [opt][D]
  float* g1_s = ksum_s + a.opt * D;
  float* b1_s = g1_s + D;
  float* g2_s = b1_s + D;          // [CO]
  float* b2_s = g2_s + a.CO;
  const int tcols = tmem_cols(max(max(D2, a.CO), a.opt * D));
  if (tid == 0) {
    for (int i = 0; i < RING; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    tc::mbar_init(&stagebar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) { tc::tmem_alloc(&tmem_base_s, (uint32_t)tcols); tc::tmem_relinquish(); }
  for (int i = tid; i < a.opt * D; i += NTH) ksum_s[i] = (b0 + i / D < a.B) ? a.ksum[(size_t)(b0 + i / D) * D + i % D] : 0.f;
  for (int i = tid; i < D; i += NTH) { g1_s[i] = a.g1[i]; b1_s[i] = a.b1[i]; }
  for (int i = tid; i < a.CO; i += NTH) { g2_s[i] = a.g2[i]; b2_s[i] = a.b2[i]; }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  Pipe p;
  if (tid == 0) {
    p.full = full; p.empty = empty; p.ring_s = tc::smem_u32(ring); p.blob = a.blob;
    p.obj = a.kvimg + (size_t)b0 * D * dh * 4; p.tab = &a.tab;
    p.loaded = 0; p.used = 0;
    pipe_prefetch(p);
  }
  const uint32_t tmem = tmem_base_s;
  const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
  uint32_t par = 0;
  const float* F1 = a.feat1 + (size_t)b * a.f1_bs;
  if (a.f1_pm) load_image_pm<true>(imgF, F1 + (size_t)row * a.C1, a.C1, a.C1P, valid, tid);     // residual is re-read from global
  else load_image_cm<true>(imgF, F1 + row, a.ldf1, a.C1, a.C1P, valid, tid);
  const float* ksr = ksum_s + ob * D;

  float z[4] = {0.f, 0.f, 0.f, 0.f};        // per-head 1 / (Q.Ksum + eps)   (H <= 4)
  if (a.qpre) {
    // ---- q from HBM: elu+1 -> operand image, per-head dot with Ksum
    const float* Q = a.q + (size_t)b * a.q_bs + row;
    float dot[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c0 = 0; c0 < D; c0 += 16) {              // 16 channels per batch (one head: dh is a multiple of 16)
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = valid ? __ldg(Q + (size_t)(c0 + i) * a.ldq) : 0.f;
      const int h = c0 / dh;
      float dd = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        v[i] = elu1(v[i]);
        dd = fmaf(v[i], ksr[c0 + i], dd);
      }
#pragma unroll
      for (int hh = 0; hh < 4; ++hh) dot[hh] += (hh == h) ? dd : 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<float4*>(imgQ + ((c0 >> 2) + j) * 2048 + tid * 16) = tc::tf32_rna4(make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
    }
#pragma unroll
    for (int hh = 0; hh < 4; ++hh) z[hh] = 1.f / (dot[hh] + 1e-6f);
    stage_sync();
  } else {
    // ---- GEMM 0: q = feat1 . Wq^T, then the same out of TMEM
    stage_sync();
    if (tid == 0) {
      pipe_gemm(p, tmem, tc::smem_u32(imgF), a.C1P, D, false);
      tc::umma_commit(&stagebar);
      pipe_prefetch(p);
    }
    __syncwarp();
    tc::mbar_wait(&stagebar, par);
    par ^= 1u;
    tc::tc_fence_after();
    float dot[4] = {0.f, 0.f, 0.f, 0.f};
    for (int q = 0; q < D / 16; ++q) {
      uint32_t r[16];
      tc::tmem_ld16(tl + 16 * q, r);
      tc::tmem_ld_wait();
      const int h = (16 * q) / dh;
      float v[16];
      float dd = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        v[j] = elu1(__uint_as_float(r[j]));
        dd = fmaf(v[j], ksr[16 * q + j], dd);
      }
#pragma unroll
      for (int hh = 0; hh < 4; ++hh) dot[hh] += (hh == h) ? dd : 0.f;
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<float4*>(imgQ + ((16 * q + j) / 4) * 2048 + tid * 16) = tc::tf32_rna4(make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
    }
#pragma unroll
    for (int hh = 0; hh < 4; ++hh) z[hh] = 1.f / (dot[hh] + 1e-6f);
    stage_sync();
  }
  // ---- GEMM 1a: (elu(q)+1)_h . KV_h for every object of the tile and every head (K = N = dh), column block o*D + h*dh
  if (tid == 0) {
    for (int o = 0; o < a.opt; ++o)
      for (int h = 0; h < a.H; ++h)
        pipe_gemm(p, tmem + (uint32_t)(o * D + h * dh), tc::smem_u32(imgQ) + (uint32_t)((h * dh / 8) * 4096), dh, dh, false);
    tc::umma_commit(&stagebar);
    pipe_prefetch(p);
  }
  __syncwarp();
  tc::mbar_wait(&stagebar, par);
  par ^= 1u;
  tc::tc_fence_after();
  for (int q = 0; q < D / 16; ++q) {           // this row's object block, scaled by 1 / (Q.Ksum) of its head
    uint32_t r[16];
    tc::tmem_ld16(tl + (uint32_t)(ob * D + 16 * q), r);
    tc::tmem_ld_wait();
    const int h = (16 * q) / dh;
    float zz = z[0];
#pragma unroll
    for (int hh = 1; hh < 4; ++hh) zz = (hh == h) ? z[hh] : zz;
#pragma unroll
    for (int j = 0; j < 16; j += 4)
      *reinterpret_cast<float4*>(imgQ + ((16 * q + j) / 4) * 2048 + tid * 16) =
          tc::tf32_rna4(make_float4(__uint_as_float(r[j]) * zz, __uint_as_float(r[j + 1]) * zz, __uint_as_float(r[j + 2]) * zz, __uint_as_float(r[j + 3]) * zz));
  }
  stage_sync();
  // ---- GEMM 1b: merge projection, LayerNorm1
  if (tid == 0) {
    pipe_gemm(p, tmem, tc::smem_u32(imgQ), D, D, false);
    tc::umma_commit(&stagebar);
    pipe_prefetch(p);
  }
  __syncwarp();
  tc::mbar_wait(&stagebar, par);
  par ^= 1u;
  tc::tc_fence_after();
  {
    float mean, rstd;
    row_stats(tl, D, mean, rstd);
    for (int q = 0; q < D / 16; ++q) {
      uint32_t r[16];
      tc::tmem_ld16(tl + 16 * q, r);
      tc::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const int c = 16 * q + j;
        float4 v;
        v.x = (__uint_as_float(r[j + 0]) - mean) * rstd * g1_s[c + 0] + b1_s[c + 0];
        v.y = (__uint_as_float(r[j + 1]) - mean) * rstd * g1_s[c + 1] + b1_s[c + 1];
        v.z = (__uint_as_float(r[j + 2]) - mean) * rstd * g1_s[c + 2] + b1_s[c + 2];
        v.w = (__uint_as_float(r[j + 3]) - mean) * rstd * g1_s[c + 3] + b1_s[c + 3];
        *reinterpret_cast<float4*>(imgQ + (c / 4) * 2048 + tid * 16) = tc::tf32_rna4(v);
      }
    }
  }
  stage_sync();
  // ---- GEMM 2: hidden = relu(W0a feat1 + W0b msg)
  if (tid == 0) {
    pipe_gemm(p, tmem, tc::smem_u32(imgF), a.C1P, D2, false);
    pipe_gemm(p, tmem, tc::smem_u32(imgQ), D, D2, true);
    tc::umma_commit(&stagebar);
    pipe_prefetch(p);
  }
  __syncwarp();
  tc::mbar_wait(&stagebar, par);
  par ^= 1u;
  tc::tc_fence_after();
  for (int q = 0; q < D2 / 16; ++q) {
    uint32_t r[16];
    tc::tmem_ld16(tl + 16 * q, r);
    tc::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      float4 v;
      v.x = fmaxf(__uint_as_float(r[j + 0]), 0.f);
      v.y = fmaxf(__uint_as_float(r[j + 1]), 0.f);
      v.z = fmaxf(__uint_as_float(r[j + 2]), 0.f);
      v.w = fmaxf(__uint_as_float(r[j + 3]), 0.f);
      *reinterpret_cast<float4*>(imgH + ((16 * q + j) / 4) * 2048 + tid * 16) = tc::tf32_rna4(v);
    }
  }
  stage_sync();
  // ---- GEMM 3: W2 hidden, LayerNorm2 (+ feat1)
  if (tid == 0) {
    pipe_gemm(p, tmem, tc::smem_u32(imgH), D2, a.CO, false);
    tc::umma_commit(&stagebar);
  }
  __syncwarp();
  tc::mbar_wait(&stagebar, par);
  par ^= 1u;
  tc::tc_fence_after();
  {
    float mean, rstd;
    row_stats(tl, a.CO, mean, rstd);
    float* O = a.out + (size_t)b * a.o_bs;
    for (int q = 0; q < a.CO / 16; ++q) {
      uint32_t r[16];
      tc::tmem_ld16(tl + 16 * q, r);
      tc::tmem_ld_wait();
      if (valid) {
        float rr[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int c = 16 * q + j;
          rr[j] = !a.res ? 0.f : (a.f1_pm ? __ldg(F1 + (size_t)row * a.C1 + c) : __ldg(F1 + (size_t)c * a.ldf1 + row));
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int c = 16 * q + j;
          O[(size_t)c * a.ldo + row] = (__uint_as_float(r[j]) - mean) * rstd * g2_s[c] + b2_s[c] + rr[j];
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, (uint32_t)tcols);
}

// ---- host side -----------------------------------------------------------------------------------------------------
struct TabBuilder {
  ChunkTab t;
  uint32_t blob_off = 0;
  bool ok = true;
  int cap;
  // slots of 16 KB when the operand images leave room for one CTA per SM anyway, 8 KB otherwise (2 ... 5 CTAs per SM)
  explicit TabBuilder(int img_bytes) : cap(img_bytes > 96 * 1024 ? MAX_SLOT : 8192) { t.n = 0; t.slot_bytes = 0; }
  // one GEMM operand W (K x N image); per_obj parts live at obj_off in the tile's per-object operand array
  void part(int K, int N, bool per_obj, uint32_t obj_off = 0) {
    int kc = K < 32 ? K : 32;
    while (kc > 8 && kc * N * 4 > cap) kc >>= 1;
    if (kc * N * 4 > MAX_SLOT || (K % 8) || (N % 16)) { ok = false; return; }
    uint32_t off = per_obj ? obj_off : blob_off;
    for (int kk = 0; kk < K; kk += kc) {
      const int k = (K - kk) < kc ? (K - kk) : kc;
      if (t.n >= MAXCH) { ok = false; return; }
      t.off[t.n] = off;
      t.bytes[t.n] = (uint32_t)(k * N * 4);
      t.per_obj[t.n] = per_obj ? 1 : 0;
      if ((int)t.bytes[t.n] > t.slot_bytes) t.slot_bytes = (int)t.bytes[t.n];
      off += t.bytes[t.n];
      ++t.n;
    }
    if (!per_obj) blob_off = off;
  }
  void finish() { t.slot_bytes = (t.slot_bytes + 1023) & ~1023; }
};

}  // namespace

extern "C" {

// bytes of the weight blob pcreid_attn_front expects: images of Wp2 (DP -> C2), Wfp (C2 -> NFP), Wf (C2 -> NF)
int pcreid_attn_front_blob_bytes(int C2, int DP, int NFP, int NF) { return 4 * (DP * C2 + C2 * NFP + C2 * NF); }

int pcreid_attn_front(int B, int S, int C2, int DP, int NFP, int NF, const float* xyz, const float* feat, long long f_bs, int ldf,
                      const float* wp0, const float* bp0, const float* bp2, const void* blob, float* out, long long o_bs, int ldo,
                      void* stream) {
  if (B <= 0 || S <= 0) return PCREID_OK;
  if (!xyz || !feat || !wp0 || !bp0 || !bp2 || !blob || !out) return PCREID_ERR_ARG;
  if (C2 % 16 || DP % 16 || NFP % 16 || NF % 16 || NFP <= 0 || C2 > 128 || DP > 128 || NFP + NF > 512 || ((long long)B * S + 127) / 128 > 0x7fffffffLL)
    return PCREID_ERR_UNSUPPORTED;
  FrontArgs a;
  a.B = B; a.S = S; a.C2 = C2; a.DP = DP; a.NFP = NFP; a.NF = NF;
  a.xyz = xyz; a.feat = feat; a.f_bs = f_bs; a.ldf = ldf; a.wp0 = wp0; a.bp0 = bp0; a.bp2 = bp2;
  a.blob = static_cast<const uint8_t*>(blob); a.out = out; a.o_bs = o_bs; a.ldo = ldo;
  TabBuilder tb((C2 + (DP > C2 ? DP : C2)) * 512);
  tb.part(DP, C2, false);
  tb.part(C2, NFP, false);
  if (NF > 0) tb.part(C2, NF, false);
  tb.finish();
  if (!tb.ok) return PCREID_ERR_UNSUPPORTED;
  a.tab = tb.t;
  const int CA = DP > C2 ? DP : C2;
  const int smem = (C2 + CA) * 512 + RING * a.tab.slot_bytes + (4 * DP + C2) * 4;
  if (smem > 227 * 1024) return PCREID_ERR_UNSUPPORTED;
  cudaFuncSetAttribute(attn_front_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  attn_front_kernel<<<(unsigned)(((long long)B * S + 127) / 128), NTH, smem, (cudaStream_t)stream>>>(a);
  return pcreid_launch_status();
}

int pcreid_linattn_kv_img(int B, int S, int d, int H, const float* K, long long k_bs, int ldk, const float* V, long long v_bs, int ldv,
                          float* kvimg, float* ksum, void* stream) {
  if (B <= 0) return PCREID_OK;
  if (!K || !V || !kvimg || !ksum || H <= 0 || S <= 0) return PCREID_ERR_ARG;
  if (d % H || (d / H) % 4 || d > 128 || d < 16) return PCREID_ERR_UNSUPPORTED;
  const int ntiles = H * (d / H / 4) * (d / H / 4);
  if (ntiles > 512 || (ntiles < 256 && 256 % ntiles)) return PCREID_ERR_UNSUPPORTED;
  const int G = ntiles >= 256 ? 1 : 256 / ntiles;
  int smem = 2 * d * KV_SCP * 4;
  if (G > 1 && G * ntiles * 16 * 4 > smem) smem = G * ntiles * 16 * 4;
  if (ntiles > 256) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(linattn_kv_img_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    linattn_kv_img_kernel<2><<<B, 256, smem, (cudaStream_t)stream>>>(S, d, H, K, k_bs, ldk, V, v_bs, ldv, kvimg, ksum);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(linattn_kv_img_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    linattn_kv_img_kernel<1><<<B, 256, smem, (cudaStream_t)stream>>>(S, d, H, K, k_bs, ldk, V, v_bs, ldv, kvimg, ksum);
  }
  return pcreid_launch_status();
}

// objects per attn_back tile for `rows` query rows per object
int pcreid_attn_back_objects_per_tile(int rows, int D) {
  if (rows >= 128 || rows <= 0 || 128 % rows) return 1;
  const int opt = 128 / rows;
  return (opt <= 4 && opt * D <= 512) ? opt : 1;
}

// blob: [Wq (C1P -> D) unless q is given][Wm (D -> D)][W0a (C1P -> 2D)][W0b (D -> 2D)][W2 (2D -> CO)]
int pcreid_attn_back_blob_bytes(int D, int C1P, int CO, int qpre) {
  return 4 * ((qpre ? 0 : C1P * D) + D * D + C1P * 2 * D + D * 2 * D + 2 * D * CO);
}

int pcreid_attn_back(int B, int rows, int D, int H, int C1, int CO, int res, int f1_pm, const float* feat1, long long f1_bs,
                     int ldf1, const float* q, long long q_bs, int ldq, const float* ksum, const float* kvimg, const float* g1,
                     const float* b1, const float* g2, const float* b2, const void* blob, float* out, long long o_bs, int ldo,
                     void* stream) {
  if (B <= 0 || rows <= 0) return PCREID_OK;
  if (!feat1 || !ksum || !kvimg || !g1 || !b1 || !g2 || !b2 || !blob || !out || H <= 0) return PCREID_ERR_ARG;
  const int C1P = (C1 + 7) & ~7;
  if (D % 16 || D > 128 || H > 4 || D % H || (D / H) % 16 || CO % 16 || CO > 2 * D || C1 <= 0 || C1P > 128) return PCREID_ERR_UNSUPPORTED;
  if (res && C1 != CO) return PCREID_ERR_ARG;
  const int dh = D / H;
  BackArgs a;
  a.B = B; a.rows = rows; a.D = D; a.H = H; a.C1 = C1; a.C1P = C1P; a.CO = CO;
  a.opt = pcreid_attn_back_objects_per_tile(rows, D);
  a.qpre = q ? 1 : 0; a.res = res; a.f1_pm = f1_pm;
  a.feat1 = feat1; a.f1_bs = f1_bs; a.ldf1 = ldf1; a.q = q; a.q_bs = q_bs; a.ldq = ldq; a.ksum = ksum;
  a.kvimg = reinterpret_cast<const uint8_t*>(kvimg); a.g1 = g1; a.b1 = b1; a.g2 = g2; a.b2 = b2;
  a.blob = static_cast<const uint8_t*>(blob); a.out = out; a.o_bs = o_bs; a.ldo = ldo;
  const long long tiles = a.opt > 1 ? ((long long)B + a.opt - 1) / a.opt : (long long)B * ((rows + 127) / 128);
  if (tiles > 0x7fffffffLL) return PCREID_ERR_UNSUPPORTED;
  TabBuilder tb(((C1P + D) > 2 * D ? (C1P + D) : 2 * D) * 512);
  if (!q) tb.part(C1P, D, false);
  for (int o = 0; o < a.opt; ++o)
    for (int h = 0; h < H; ++h) tb.part(dh, dh, true, (uint32_t)((o * H + h) * dh * dh * 4));
  tb.part(D, D, false);
  tb.part(C1P, 2 * D, false);
  tb.part(D, 2 * D, false);
  tb.part(2 * D, CO, false);
  tb.finish();
  if (!tb.ok) return PCREID_ERR_UNSUPPORTED;
  a.tab = tb.t;
  const int img_ch = (C1P + D) > 2 * D ? (C1P + D) : 2 * D;
  const int smem = img_ch * 512 + RING * a.tab.slot_bytes + ((a.opt + 2) * D + 2 * CO) * 4;
  if (smem > 227 * 1024 - 256) return PCREID_ERR_UNSUPPORTED;
  cudaFuncSetAttribute(attn_back_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  attn_back_kernel<<<(unsigned)tiles, NTH, smem, (cudaStream_t)stream>>>(a);
  return pcreid_launch_status();
}

}  // extern "C"
