// All-pairs 'concat' match head on the 5th-gen tensor cores ("fast" mode of pair_concat_head_kernel, rowops.cu).
//
//   logit[t,d] = w . relu( GN2(W2 relu(GN1(A[t] + Bv[d]))) + [e_t ; e_d] ) + b0
// (ReIDNet.match_forward 'concat', mmdet3d/models/ReIDNet.py:415-419, 455-458; LinearRes + Linear,
// lanegcn_nets.py:228-241; config reid_pts_point-transformer_baseline.py: LinearRes(256, 256, GN ng=32) + Linear(256, 1)).
// The first Linear is hoisted per object on the host side (A = W1[:, :E] e_t, Bv = W1[:, E:] e_d), so the T x D x 2E
// pair tensor never exists; what is left per pair is GroupNorm1 + ReLU, the 256 x 256 second Linear, GroupNorm2 +
// residual + ReLU and the final dot product.
//
// One persistent CTA per SM, two groups of 4 warps; a group owns one tile of 128 pairs = 4 tracks x 32 detections
// (warp <-> track, lane <-> detection, thread = pair = TMEM lane).  The detection block's Bv / e_d rows stay resident in
// shared memory (channel-major, conflict-free) while the CTA walks over the tracks, so L2 only sees the 1.5 KB of a track.
// Per tile: the prologue computes relu(GN1(A[t] + Bv[d])) in fp32 registers and writes it, packed to bf16, STRAIGHT INTO
// TMEM as the A operand (tcgen05.st: no shared memory, no proxy fence) -> two tcgen05 GEMMs (N = 128 output channels
// each, K = 256, W2 resident in shared memory as one bf16 K-major image) -> epilogue per half by the row-owning thread:
// GroupNorm2 over groups of 8 channels, + residual, ReLU, dot with w.  TMEM: 2 groups x (128 operand + 128 accumulator
// columns).  The tensor pipe of one group runs behind the fp32 prologue / epilogue of the other.
#include "../../include/pcreid.h"
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int HD = 256, EE = 128, CG = 8;      // hidden width, embedding width per object, channels per GroupNorm group
constexpr int DB = 32, TQ = 4, GT = 128, NG = 2;
constexpr int W2_BYTES = HD * HD * 2;          // bf16 K-major image [k/8][256 n][8]
constexpr int BV_OFF = W2_BYTES, ED_OFF = BV_OFF + HD * DB * 4, PAR_OFF = ED_OFF + EE * DB * 4, GRP_OFF = PAR_OFF + 5 * HD * 4;
constexpr int GRP_BYTES = TQ * HD * 4 + TQ * EE * 4;
constexpr int SMEM_BYTES = GRP_OFF + NG * GRP_BYTES;    // 197 632 B
constexpr float GN_EPS = 1e-5f;

struct Args {
  int T, D, n_db, n_tq;
  const float *A, *Bv, *Et, *Ed;
  const uint8_t* W2img;
  const float *g1, *be1, *g2, *be2, *w;
  float b0;
  const uint8_t* mask;
  float* out;
};

__device__ __forceinline__ float u2f(uint32_t u) { return __uint_as_float(u); }

// GroupNorm over 8 values + affine + optional residual + ReLU, in place
__device__ __forceinline__ void gn8(float (&x)[8], const float4& ga, const float4& gb, const float4& ba, const float4& bb) {
  float s = ((x[0] + x[1]) + (x[2] + x[3])) + ((x[4] + x[5]) + (x[6] + x[7]));
  const float mean = s * 0.125f;
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] -= mean; v = fmaf(x[i], x[i], v); }
  const float rstd = rsqrtf(v * 0.125f + GN_EPS);
  const float g[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
  const float b[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i] * rstd, g[i], b[i]);
}

__global__ void __launch_bounds__(NG * GT, 1) pair_concat_head_tc_kernel(const Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[NG];
  __shared__ uint32_t tmem_base_s;
  float* Bvs = reinterpret_cast<float*>(smem + BV_OFF);          // [256 c][32 d]
  float* Eds = reinterpret_cast<float*>(smem + ED_OFF);          // [128 c][32 d]
  const float4* par = reinterpret_cast<const float4*>(smem + PAR_OFF);   // g1 | be1 | g2 | be2 | w, 64 float4 each
  if (threadIdx.x == 0) {
    for (int i = 0; i < NG; ++i) tc::mbar_init(&bars[i], 1);
    tc::fence_mbar_init();
  }
  if (threadIdx.x < 32) { tc::tmem_alloc(&tmem_base_s, 512); tc::tmem_relinquish(); }
  for (int i = threadIdx.x * 16; i < W2_BYTES; i += NG * GT * 16) cp_async16(smem + i, a.W2img + i);
  cp_async_commit();
  {
    float* p = reinterpret_cast<float*>(smem + PAR_OFF);
    const float* src[5] = {a.g1, a.be1, a.g2, a.be2, a.w};
    for (int i = threadIdx.x; i < 5 * HD; i += NG * GT) p[i] = src[i / HD][i % HD];
  }
  cp_async_wait<0>();
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();

  const int warp_u = (int)tc::uniform(threadIdx.x >> 5);
  const int gid = warp_u / 4, wg = warp_u % 4, lane = threadIdx.x & 31, gt = threadIdx.x % GT;
  const uint32_t tbase = tc::uniform(tmem_base_s) + gid * 256;
  const uint32_t tA = tbase + ((uint32_t)(wg * 32) << 16), tD = tA + 128;      // this warp's lane quadrant
  uint64_t* bar = bars + gid;
  uint32_t par_phase = 0;
  float* At = reinterpret_cast<float*>(smem + GRP_OFF + gid * GRP_BYTES);       // [4 tracks][256]
  float* Etg = At + TQ * HD;                                                    // [4 tracks][128]
  const uint32_t sW2 = tc::smem_u32(smem);
  const uint32_t idesc = tc::instr_desc(128, 128, tc::FMT_BF16, tc::MAJOR_K, tc::MAJOR_K);
  // B operand of output half h: the same image entered at row 128 h (K-major [k/8][256 rows][8]: LBO = 4096, SBO = 128)
  uint64_t bdesc[2];
  bdesc[0] = tc::smem_desc(sW2, 4096, 128, tc::LAYOUT_NONE);
  bdesc[1] = tc::smem_desc(sW2 + 128 * 16, 4096, 128, tc::LAYOUT_NONE);
  constexpr uint32_t KSTEP = (2 * HD * 16) >> 4;

  const long long total = (long long)a.n_db * a.n_tq;
  const long long lo = total * blockIdx.x / gridDim.x, hi = total * (blockIdx.x + 1) / gridDim.x;
  for (int db = (int)(lo / a.n_tq); db <= (int)((hi - 1) / a.n_tq) && lo < hi; ++db) {
    const long long s0 = max(lo, (long long)db * a.n_tq), s1 = min(hi, (long long)(db + 1) * a.n_tq);
    __syncthreads();                              // every group is done with the previous detection block
    for (int i = threadIdx.x; i < HD * DB; i += NG * GT) {
      const int d = db * DB + (i % DB), c = i / DB;
      Bvs[i] = d < a.D ? __ldg(a.Bv + (size_t)d * HD + c) : 0.f;
    }
    for (int i = threadIdx.x; i < EE * DB; i += NG * GT) {
      const int d = db * DB + (i % DB), c = i / DB;
      Eds[i] = d < a.D ? __ldg(a.Ed + (size_t)d * EE + c) : 0.f;
    }
    __syncthreads();
    const int d = db * DB + lane;
    for (long long tile = s0 + gid; tile < s1; tile += NG) {
      const int tq = (int)(tile - (long long)db * a.n_tq), t = tq * TQ + wg;
      tc::bar_sync(1 + gid, GT);                  // previous tile's reads of At / Etg are finished
      for (int i = gt; i < TQ * HD / 4; i += GT) {
        const int tt = tq * TQ + i / (HD / 4);
        reinterpret_cast<float4*>(At)[i] = tt < a.T ? __ldg(reinterpret_cast<const float4*>(a.A + (size_t)tt * HD) + i % (HD / 4))
                                                    : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      for (int i = gt; i < TQ * EE / 4; i += GT) {
        const int tt = tq * TQ + i / (EE / 4);
        reinterpret_cast<float4*>(Etg)[i] = tt < a.T ? __ldg(reinterpret_cast<const float4*>(a.Et + (size_t)tt * EE) + i % (EE / 4))
                                                     : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      tc::bar_sync(1 + gid, GT);
      // ---- prologue: H1 = relu(GN1(A[t] + Bv[d])) -> bf16 A operand in TMEM columns [0, 128) of the group
      const float4* At4 = reinterpret_cast<const float4*>(At + wg * HD);
#pragma unroll 2
      for (int it = 0; it < 16; ++it) {           // 16 channels = 2 GroupNorm groups = 8 packed columns per iteration
        uint32_t wds[8];
#pragma unroll
        for (int gg = 0; gg < 2; ++gg) {
          const int c0 = 16 * it + 8 * gg;
          const float4 a0 = At4[c0 / 4], a1 = At4[c0 / 4 + 1];
          float x[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) x[i] += Bvs[(c0 + i) * DB + lane];
          gn8(x, par[c0 / 4], par[c0 / 4 + 1], par[64 + c0 / 4], par[64 + c0 / 4 + 1]);
#pragma unroll
          for (int j = 0; j < 4; ++j) wds[4 * gg + j] = tc::pack_bf16_relu(x[2 * j], x[2 * j + 1]);
        }
        tc::tmem_st8(tA + 8 * it, wds);
      }
      tc::tmem_st_wait();
      float part = 0.f;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        tc::tc_fence_before();
        tc::bar_sync(1 + gid, GT);                // operand complete (half 0) / accumulator drained by every thread (half 1)
        tc::tc_fence_after();
        if (wg == 0) {
          if (tc::elect_one()) {
#pragma unroll
            for (int ks = 0; ks < HD / 16; ++ks)
              tc::umma_f16_ts(tbase + 128, tbase + 8 * ks, bdesc[half] + (uint64_t)(ks * KSTEP), idesc, ks > 0 ? 1u : 0u);
            tc::umma_commit(bar);
          }
          __syncwarp();
        }
        tc::mbar_wait(bar, par_phase);
        par_phase ^= 1u;
        tc::tc_fence_after();
        // ---- epilogue of this half: GN2 (groups of 8) + residual + ReLU + dot with w
#pragma unroll 2
        for (int q = 0; q < 8; ++q) {
          uint32_t r[16];
          tc::tmem_ld16(tD + 16 * q, r);
          tc::tmem_ld_wait();
#pragma unroll
          for (int gg = 0; gg < 2; ++gg) {
            const int cl = 16 * q + 8 * gg, c0 = 128 * half + cl;       // channel inside the half / overall
            float x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = u2f(r[8 * gg + i]);
            gn8(x, par[128 + c0 / 4], par[128 + c0 / 4 + 1], par[192 + c0 / 4], par[192 + c0 / 4 + 1]);
            const float4 w0 = par[256 + c0 / 4], w1 = par[256 + c0 / 4 + 1];
            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
            if (half == 0) {                      // residual: channels [0, E) come from the track, [E, 2E) from the detection
              const float4 e0 = reinterpret_cast<const float4*>(Etg + wg * EE)[cl / 4], e1 = reinterpret_cast<const float4*>(Etg + wg * EE)[cl / 4 + 1];
              const float ev[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
              for (int i = 0; i < 8; ++i) part = fmaf(fmaxf(x[i] + ev[i], 0.f), wv[i], part);
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) part = fmaf(fmaxf(x[i] + Eds[(cl + i) * DB + lane], 0.f), wv[i], part);
            }
          }
        }
      }
      tc::tc_fence_before();                      // accumulator reads done before the next tile's GEMM (ordered by its barriers)
      if (t < a.T && d < a.D) {
        float v = part + a.b0;
        if (a.mask && !a.mask[(size_t)t * a.D + d]) v = 0.f;
        a.out[(size_t)t * a.D + d] = v;
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(tmem_base_s, 512);
}

}  // namespace

extern "C" int pcreid_pair_concat_head_tc(int T, int D, int E, int G, const float* A, const float* Bv, const float* Et, const float* Ed,
                                          const void* W2img, const float* g1, const float* be1, const float* g2, const float* be2,
                                          const float* w, float b0, const unsigned char* mask, float* out, int n_ctas, void* stream) {
  if (T <= 0 || D <= 0) return PCREID_OK;
  if (!A || !Bv || !Et || !Ed || !W2img || !g1 || !be1 || !g2 || !be2 || !w || !out) return PCREID_ERR_ARG;
  if (E != EE || G * CG != HD) return PCREID_ERR_UNSUPPORTED;      // LinearRes(256, 256, GN 32 groups): the shipped 'concat' head
  Args a{T, D, (D + DB - 1) / DB, (T + TQ - 1) / TQ, A, Bv, Et, Ed, (const uint8_t*)W2img, g1, be1, g2, be2, w, b0, mask, out};
  long long total = (long long)a.n_db * a.n_tq;
  int grid = n_ctas > 0 ? n_ctas : 148;
  if (grid > total) grid = (int)total;
  cudaFuncSetAttribute(pair_concat_head_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  pair_concat_head_tc_kernel<<<grid, NG * GT, SMEM_BYTES, (cudaStream_t)stream>>>(a);
  return pcreid_launch_status();
}
