// Fused all-pairs `xcorr_eff` match on the 5th-gen tensor cores (tcgen05 + TMEM), bf16 operands, fp32
// accumulation, fp32 LayerNorm / linear-attention normalisation.  "fast" mode of the match head.
//
// Reference arithmetic: ReIDNet.xcorr_eff (mmdet3d/models/ReIDNet.py:231-247) = 2 x corss_attention each way
// (mmdet3d/models/attention.py:192-219) + get_pooled_feats (ReIDNet.py:526-534).  The reference gathers
// feat[pairs[:,0]] / feat[pairs[:,1]] and runs ~40 torch ops per pair batch; here a pair never leaves the chip
// between layers:
//
//   phase 1 (pair_p1_kernel), unit = (pair, direction), per 128-point tile of the search object:
//     G1  [Qf1_i | .] x MK1_j      -> per-head Q.KV.merge and the two Q.Ksum dots in ONE N=144 GEMM
//         epilogue: z = 1/(dot+eps), merged = z0 D0 + z1 D1, LayerNorm1 -> X (bf16, smem operand image)
//     G2  X x W0b^T (+ U_i = W0a h_i, precomputed per object) , ReLU -> Hd
//     G3  Hd x W2^T, LayerNorm2, + h_i  -> a   (stage-1 output, kept on chip as the next A operand, also spilled
//                                              once as bf16 for phase 2)
//     G4  a x [Wk2;Wv2]^T -> Kf = elu+1, V (+ Wv2 pos_j)                -> operand image
//     G5  [Kf|V]^T x [V|1]  accumulated over the tiles in TMEM          -> KV (64x64) and Ksum of the template
//     G6  blockdiag(KV) x Wm2^T                                         -> B7 = stage-2 attention operand (global)
//   phase 2 (pair_p2_kernel), unit = (pair, direction), per tile:
//     G4' a x Wq2^T -> Qf2 ; G7 Qf2 x B7 (as G1) ; G8 [a|X] x W0^T, ReLU ; G9 Hd x W2^T, LayerNorm2, + a
//     epilogue: per-channel max / sum over the points (smem transpose), accumulated over the tiles
//   pool_finish: combine the two directions -> pooled (128) -> match head.
//
// Every GEMM is a tcgen05.mma (M=128, K=16 per instruction) issued by one thread per 128-thread group, operands in
// shared memory in the no-swizzle canonical layouts validated by tc_probe.cu, accumulators in TMEM, epilogues by the
// row-owning threads via tcgen05.ld.  A CTA holds two independent groups that share one copy of the weights, so the
// tensor pipe of one group overlaps the epilogue of the other.
#include "pair_common.cuh"

namespace {

// A group = 8 warps working on one 128-point tile: thread (row, h) owns row `row` of the tile (TMEM lane) and the
// column half `h` of every accumulator, so the epilogue work of a tile is spread over 256 threads.
struct Group {
  int t, row, h, gid;        // thread in group, tile row (== TMEM lane), column half, group in CTA (gid is warp-uniform)
  bool issuer;               // warp 0 of the group (warp-uniform): one elected lane issues the MMAs
  uint32_t tmem;             // TMEM base of the group (lane 0, first column)
  uint32_t tlane;            // tmem + (lane base of this warp << 16)
  uint64_t* bar;
  uint64_t* bar2;
  float2* xch;               // [2][128] LayerNorm partial exchange between the two column halves
  uint32_t par, par2;
  __device__ __forceinline__ void sync() { tc::bar_sync(1 + gid, GT); }
  // smem operands written by the group -> visible to the tensor core; returns after the group barrier
  __device__ __forceinline__ void publish() {
    tc::fence_async_smem();
    tc::tc_fence_before();
    sync();
    tc::tc_fence_after();
  }
  __device__ __forceinline__ void wait() { tc::mbar_wait(bar, par); par ^= 1u; tc::tc_fence_after(); }
  __device__ __forceinline__ void wait2() { tc::mbar_wait(bar2, par2); par2 ^= 1u; tc::tc_fence_after(); }
  // LayerNorm statistics over 64 channels from the two 32-channel halves (one-pass: E[x^2] - mean^2, fp32)
  __device__ __forceinline__ void ln_stats(float s, float ss, float& mean, float& rstd) {
    xch[h * 128 + row] = make_float2(s, ss);
    sync();
    const float2 o = xch[(1 - h) * 128 + row];
    mean = (s + o.x) * (1.f / 64.f);
    const float var = fmaxf((ss + o.y) * (1.f / 64.f) - mean * mean, 0.f);
    rstd = rsqrtf(var + LN_EPS);
  }
};

// ---- epilogues ------------------------------------------------------------------------------------------------
// Each starts with its global side loads (if any), THEN waits for the MMA, so the L2 latency hides behind the GEMM.

// LayerNorm affine + bf16 pack of 32 channels (cb .. cb+31) into chunks 4h .. 4h+3 of an operand image.
// y = x * (rstd * gamma) + (beta - mean * rstd * gamma): two FMAs per element, gamma/beta read as float4 from smem.
__device__ __forceinline__ void ln_apply_store(const float (&x)[32], float mean, float rstd, const float* __restrict__ ln, int cb,
                                               uint8_t* dst_row) {
  const float4* g4 = reinterpret_cast<const float4*>(ln + cb);
  const float4* b4 = reinterpret_cast<const float4*>(ln + 64 + cb);
  const float nm = -mean * rstd;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 ga = g4[2 * c], gb = g4[2 * c + 1], ba = b4[2 * c], bb = b4[2 * c + 1];
    const float gam[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
    const float bet[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float y0 = fmaf(fmaf(x[c * 8 + 2 * j], rstd, nm), gam[2 * j], bet[2 * j]);
      const float y1 = fmaf(fmaf(x[c * 8 + 2 * j + 1], rstd, nm), gam[2 * j + 1], bet[2 * j + 1]);
      w[j] = tc::pack_bf16(y0, y1);
    }
    *reinterpret_cast<uint4*>(dst_row + c * 2048) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// attention GEMM (N=144) -> z-normalised, head-merged message -> LayerNorm -> bf16 operand image (32 channels / thread)
__device__ __forceinline__ void epi_attn_ln(Group& g, const float* __restrict__ ln, uint8_t* dst) {
  float m[32];
  uint32_t d8[8];
  tc::tmem_ld8(g.tlane + 128, d8);
  const int cb = 32 * g.h;
  float s = 0.f, ss = 0.f, s2 = 0.f, ss2 = 0.f;
  float z0 = 0.f, z1 = 0.f;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t r0[16], r1[16];
    tc::tmem_ld16(g.tlane + cb + 16 * half, r0);
    tc::tmem_ld16(g.tlane + 64 + cb + 16 * half, r1);
    tc::tmem_ld_wait();
    if (half == 0) {
      z0 = 1.f / (__uint_as_float(d8[0]) + ATT_EPS);
      z1 = 1.f / (__uint_as_float(d8[1]) + ATT_EPS);
    }
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      const float a = fmaf(z0, __uint_as_float(r0[j]), z1 * __uint_as_float(r1[j]));
      const float b = fmaf(z0, __uint_as_float(r0[j + 1]), z1 * __uint_as_float(r1[j + 1]));
      m[16 * half + j] = a;
      m[16 * half + j + 1] = b;
      s += a; ss = fmaf(a, a, ss);
      s2 += b; ss2 = fmaf(b, b, ss2);
    }
  }
  float mean, rstd;
  g.ln_stats(s + s2, ss + ss2, mean, rstd);
  ln_apply_store(m, mean, rstd, ln, cb, dst + (4 * g.h) * 2048 + g.row * 16);
}

// acc[128] (+ side registers) -> ReLU -> bf16 operand image (64 channels / thread: chunks 8h .. 8h+7)
template <bool HAS_SIDE>
__device__ __forceinline__ void epi_relu128(Group& g, const uint4 (&sd)[8], uint8_t* dst) {
  uint8_t* drow = dst + (8 * g.h) * 2048 + g.row * 16;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t r[16];
    tc::tmem_ld16(g.tlane + 64 * g.h + 16 * q, r);
    tc::tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const uint4 s4 = HAS_SIDE ? sd[2 * q + c] : make_uint4(0, 0, 0, 0);
      const uint32_t sw[4] = {s4.x, s4.y, s4.z, s4.w};
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a = __uint_as_float(r[c * 8 + 2 * j]), b = __uint_as_float(r[c * 8 + 2 * j + 1]);
        if (HAS_SIDE) { a += bf_lo(sw[j]); b += bf_hi(sw[j]); }
        w[j] = tc::pack_bf16(fmaxf(a, 0.f), fmaxf(b, 0.f));
      }
      *reinterpret_cast<uint4*>(drow + (2 * q + c) * 2048) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

// acc[64] -> LayerNorm -> + residual registers (bf16 chunks 4h .. 4h+3 of the row) -> o[32] fp32 (channels 32h .. 32h+31)
__device__ __forceinline__ void epi_ln_res(Group& g, const float* __restrict__ ln, const uint4 (&rs)[4], float (&o)[32]) {
  const int cb = 32 * g.h;
  float s = 0.f, ss = 0.f, s2 = 0.f, ss2 = 0.f;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t r[16];
    tc::tmem_ld16(g.tlane + cb + 16 * half, r);
    tc::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      const float a = __uint_as_float(r[j]), b = __uint_as_float(r[j + 1]);
      o[16 * half + j] = a;
      o[16 * half + j + 1] = b;
      s += a; ss = fmaf(a, a, ss);
      s2 += b; ss2 = fmaf(b, b, ss2);
    }
  }
  float mean, rstd;
  g.ln_stats(s + s2, ss + ss2, mean, rstd);
  const float4* g4 = reinterpret_cast<const float4*>(ln + cb);
  const float4* b4 = reinterpret_cast<const float4*>(ln + 64 + cb);
  const float nm = -mean * rstd;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 ga = g4[2 * c], gb = g4[2 * c + 1], ba = b4[2 * c], bb = b4[2 * c + 1];
    const float gam[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
    const float bet[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
    const uint32_t rw[4] = {rs[c].x, rs[c].y, rs[c].z, rs[c].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = c * 8 + 2 * j;
      o[k] = fmaf(fmaf(o[k], rstd, nm), gam[2 * j], bet[2 * j]) + bf_lo(rw[j]);
      o[k + 1] = fmaf(fmaf(o[k + 1], rstd, nm), gam[2 * j + 1], bet[2 * j + 1]) + bf_hi(rw[j]);
    }
  }
}

// o[32] (channels 32h ..) -> chunks 4h .. 4h+3 of a 64-channel operand image
__device__ __forceinline__ void store_image32(const float (&o)[32], uint8_t* dst, int row, int h) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = tc::pack_bf16(o[c * 8 + 2 * j], o[c * 8 + 2 * j + 1]);
    *reinterpret_cast<uint4*>(dst + (4 * h + c) * 2048 + row * 16) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// 32 accumulator columns starting at col -> (elu+1 | + side) -> 4 chunks starting at chunk0 of dst
template <bool ELU, bool HAS_SIDE>
__device__ __forceinline__ void feat32(Group& g, int col, const uint4* sd, uint8_t* dst, int chunk0) {
  uint8_t* drow = dst + chunk0 * 2048 + g.row * 16;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    uint32_t r[16];
    tc::tmem_ld16(g.tlane + col + 16 * q, r);
    tc::tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t sw[4] = {0, 0, 0, 0};
      if (HAS_SIDE) { sw[0] = sd[2 * q + c].x; sw[1] = sd[2 * q + c].y; sw[2] = sd[2 * q + c].z; sw[3] = sd[2 * q + c].w; }
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a = __uint_as_float(r[c * 8 + 2 * j]), b = __uint_as_float(r[c * 8 + 2 * j + 1]);
        if (HAS_SIDE) { a += bf_lo(sw[j]); b += bf_hi(sw[j]); }
        if (ELU) { a = elu1(a); b = elu1(b); }
        w[j] = tc::pack_bf16(a, b);
      }
      *reinterpret_cast<uint4*>(drow + (2 * q + c) * 2048) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

__device__ __forceinline__ void group_setup(Group& g, uint64_t* bars, uint32_t tmem_base, uint8_t* xch) {
  const int warp_u = (int)tc::uniform(threadIdx.x >> 5);      // warp index, known uniform to the compiler
  g.gid = warp_u / (GT / 32);
  g.t = threadIdx.x % GT;
  const int warp = warp_u % (GT / 32);
  g.issuer = warp == 0;
  tmem_base = tc::uniform(tmem_base);
  g.row = 32 * (warp & 3) + (g.t & 31);
  g.h = warp >> 2;
  g.tmem = tmem_base + g.gid * 256;
  g.tlane = g.tmem + ((uint32_t)((warp & 3) * 32) << 16);
  g.bar = bars + 2 * g.gid;
  g.bar2 = bars + 2 * g.gid + 1;
  g.xch = reinterpret_cast<float2*>(xch);
  g.par = 0;
  g.par2 = 0;
}

// ---------------------------------------------------------------------------------------------------------------
// phase 1
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(2 * GT, 1) pair_p1_kernel(const P1Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[4];
  __shared__ uint32_t tmem_base_s;
  uint8_t* Wsm = smem;
  const float* ln1 = reinterpret_cast<const float*>(Wsm + P1_LN);         // gamma[64] | beta[64] of cross_stage1.norm1
  const float* ln2 = ln1 + 128;                                           // cross_stage1.norm2
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) tc::mbar_init(&bars[i], 1);
    tc::fence_mbar_init();
  }
  if (threadIdx.x < 32) { tc::tmem_alloc(&tmem_base_s, 512); tc::tmem_relinquish(); }
  copy_to_smem(Wsm, a.W, P1_WBYTES, threadIdx.x, 2 * GT);
  cp_async_commit();
  cp_async_wait<0>();
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  Group g;
  uint8_t* G = smem + P1_WBYTES + tc::uniform(threadIdx.x / GT) * P1_GBYTES;
  group_setup(g, bars, tmem_base_s, G + P1_XCH);
  uint8_t* QXa = G + P1_QXA;
  uint8_t* HdKV = G + P1_HDKV;
  uint8_t* MK1 = G + P1_MK1;
  if (g.t < 128) {   // the two constant chunks appended to V: column 64 of B == 1 for every point
    uint4* ones = reinterpret_cast<uint4*>(G + P1_ONES);
    ones[g.t] = make_uint4(0x00003f80u, 0, 0, 0);      // bf16 1.0 in element 0 of the chunk
    ones[128 + g.t] = make_uint4(0, 0, 0, 0);
  }
  const uint32_t sQXa = tc::smem_u32(QXa), sHd = tc::smem_u32(HdKV), sMK1 = tc::smem_u32(MK1), sW = tc::smem_u32(Wsm);
  const uint32_t id144 = tc::instr_desc(128, NB7, tc::FMT_BF16, tc::MAJOR_K, tc::MAJOR_MN);
  const uint32_t id128 = tc::instr_desc(128, 128, tc::FMT_BF16, tc::MAJOR_K, tc::MAJOR_K);
  const uint32_t id64 = tc::instr_desc(128, 64, tc::FMT_BF16, tc::MAJOR_K, tc::MAJOR_K);
  const uint32_t idkv = tc::instr_desc(128, 80, tc::FMT_BF16, tc::MAJOR_MN, tc::MAJOR_MN);
  const Opnd oQXa = A_IMG(sQXa), oHd = A_IMG(sHd), oMK1 = B7_IMG(sMK1), oW0b = W_IMG(sW + P1_W0B, 128), oW2 = W_IMG(sW + P1_W2, 64),
             oWkv = W_IMG(sW + P1_WKV, 128), oWm = W_IMG(sW + P1_WM, 64), oKfV = opnd(sHd, 128u, 2048u, 256u),
             oVones = opnd(sHd + 8 * 2048, 128u, 2048u, 256u);

  const int ngroups = gridDim.x * 2, gg = blockIdx.x * 2 + g.gid;
  const int u0 = (int)((long long)a.n_units * gg / ngroups), u1 = (int)((long long)a.n_units * (gg + 1) / ngroups);
  int cur_templ = -1;
  // register prefetch of the next tile's query image (16 KB / 256 threads = 4 x 16 B each)
  uint4 pre[4];
  if (u0 < u1) {
    const uint4* src = reinterpret_cast<const uint4*>(a.QF1 + (size_t)a.u_search[u0] * a.NT * IMG);
#pragma unroll
    for (int i = 0; i < 4; ++i) pre[i] = __ldg(src + g.t + i * GT);
  }
  int so_next = u0 < u1 ? a.u_search[u0] : 0, te_next = u0 < u1 ? a.u_templ[u0] : 0, slot_next = u0 < u1 ? a.u_slot[u0] : 0;
  for (int u = u0; u < u1; ++u) {
    const int so = so_next, te = te_next, slot = slot_next;
    if (u + 1 < u1) { so_next = a.u_search[u + 1]; te_next = a.u_templ[u + 1]; slot_next = a.u_slot[u + 1]; }
    for (int tile = 0; tile < a.NT; ++tile) {
      const size_t ti = (size_t)so * a.NT + tile;
      if (tile > 0) g.wait2();                                            // previous tile's KV GEMM still reads HdKV
      // ---- stage operands of G1
#pragma unroll
      for (int i = 0; i < 4; ++i) reinterpret_cast<uint4*>(QXa)[g.t + i * GT] = pre[i];
      if (te != cur_templ) {
        copy_to_smem(MK1, a.MK1 + (size_t)te * B7_BYTES, B7_BYTES, g.t, GT);
        cp_async_commit();
        cp_async_wait<0>();
        cur_templ = te;
      }
      g.publish();
      if (g.issuer) { if (tc::elect_one()) { issue_gemm<4>(g.tmem, oQXa, oMK1, id144, false); tc::umma_commit(g.bar); } __syncwarp(); }
      {   // prefetch the next (unit, tile) query image while the tensor core works
        int nu = u, nt = tile + 1;
        if (nt == a.NT) { nu = u + 1; nt = 0; }
        if (nu < u1) {
          const uint4* src = reinterpret_cast<const uint4*>(a.QF1 + ((size_t)(nt == 0 ? so_next : so) * a.NT + nt) * IMG);
#pragma unroll
          for (int i = 0; i < 4; ++i) pre[i] = __ldg(src + g.t + i * GT);
        }
      }
      uint4 sdU[8], sdH[4], sdPV[8];
      g.wait();
      load_side<8>(sdU, a.U + ti * 2 * IMG, 8 * g.h, g.row);              // consumed one stage later (after G2)
      epi_attn_ln(g, ln1, QXa);                                           // X
      g.publish();
      if (g.issuer) { if (tc::elect_one()) { issue_gemm<4>(g.tmem, oQXa, oW0b, id128, false); tc::umma_commit(g.bar); } __syncwarp(); }
      g.wait();
      load_side<4>(sdH, a.H + ti * IMG, 4 * g.h, g.row);                  // consumed after G3
      epi_relu128<true>(g, sdU, HdKV);                                    // Hd = relu(X W0b^T + W0a h)
      g.publish();
      if (g.issuer) { if (tc::elect_one()) { issue_gemm<8>(g.tmem, oHd, oW2, id64, false); tc::umma_commit(g.bar); } __syncwarp(); }
      g.wait();
      if (g.h == 1) load_side<8>(sdPV, a.PV + ti * IMG, 0, g.row);        // consumed after G4
      {
        float o[32];
        epi_ln_res(g, ln2, sdH, o);                                       // a = h + LN2(.)
        store_image32(o, QXa, g.row, g.h);
        store_image32(o, a.A_out + (((size_t)slot * 2 + a.role) * a.NT + tile) * IMG, g.row, g.h);
      }
      g.publish();
      if (g.issuer) { if (tc::elect_one()) { issue_gemm<4>(g.tmem, oQXa, oWkv, id128, false); tc::umma_commit(g.bar); } __syncwarp(); }
      g.wait();
      if (g.h == 0) {   // column half 0: Kf = elu(k)+1 -> chunks 0..7 ; column half 1: V = v + Wv pos -> chunks 8..15
        feat32<true, false>(g, 0, nullptr, HdKV, 0);
        feat32<true, false>(g, 32, nullptr, HdKV, 4);
      } else {
        feat32<false, true>(g, 64, sdPV, HdKV, 8);
        feat32<false, true>(g, 96, sdPV + 4, HdKV, 12);
      }
      g.publish();
      if (g.issuer) {   // KV += [Kf|V]^T [V|1]   (M = 128 channels, N = 80, K = 128 points)
        if (tc::elect_one()) {
          issue_gemm<8>(g.tmem + KV_COL, oKfV, oVones, idkv, tile > 0);
          tc::umma_commit(g.bar2);
        }
        __syncwarp();
      }
    }
    g.wait2();
    // ---- B7 = stage-2 attention operand of this (pair, direction) as template.  Rows d < 64 of the KV accumulator.
    {
      float kv[32];
      float ksum = 0.f;
      if (g.row < 64) {
        uint32_t r[32];
        tc::tmem_ld32(g.tlane + KV_COL + 32 * g.h, r);
        uint32_t r8[8];
        tc::tmem_ld8(g.tlane + KV_COL + 64, r8);
        tc::tmem_ld_wait();
        ksum = __uint_as_float(r8[0]);
        const bool keep = (g.row >> 5) == g.h;                            // block diagonal: head of row d == head of columns
#pragma unroll
        for (int j = 0; j < 32; ++j) kv[j] = keep ? __uint_as_float(r[j]) : 0.f;
        store_image32(kv, QXa, g.row, g.h);
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(QXa + (4 * g.h + c) * 2048 + g.row * 16) = make_uint4(0, 0, 0, 0);
      }
      g.publish();
      if (g.issuer) { if (tc::elect_one()) { issue_gemm<4>(g.tmem, oQXa, oWm, id64, false); tc::umma_commit(g.bar); } __syncwarp(); }
      g.wait();
      if (g.row < 64) {
        uint32_t r[32];
        tc::tmem_ld32(g.tlane + 32 * g.h, r);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) kv[j] = __uint_as_float(r[j]);
        write_b7_part(kv, 4 * g.h, ksum, g.h == 0, g.row, a.B7_out + ((size_t)slot * 2 + a.role) * B7_BYTES);
      }
      tc::tc_fence_before();
      g.sync();
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(tmem_base_s, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// phase 2
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(2 * GT, 1) pair_p2_kernel(const P2Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[4];
  __shared__ uint32_t tmem_base_s;
  __shared__ float comb[2][2][4][64];
  uint8_t* Wsm = smem;
  const float* ln1 = reinterpret_cast<const float*>(Wsm + P2_LN);         // cross_stage2.norm1
  const float* ln2 = ln1 + 128;                                           // cross_stage2.norm2
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) tc::mbar_init(&bars[i], 1);
    tc::fence_mbar_init();
  }
  if (threadIdx.x < 32) { tc::tmem_alloc(&tmem_base_s, 512); tc::tmem_relinquish(); }
  copy_to_smem(Wsm, a.W, P2_WBYTES, threadIdx.x, 2 * GT);
  cp_async_commit();
  cp_async_wait<0>();
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  Group g;
  uint8_t* G = smem + P2_WBYTES + tc::uniform(threadIdx.x / GT) * P2_GBYTES;
  group_setup(g, bars, tmem_base_s, G + P2_XCH);
  uint8_t* R1 = G + P2_R1;
  float* R1f = reinterpret_cast<float*>(R1);
  const uint32_t sR1 = tc::smem_u32(R1), sW = tc::smem_u32(Wsm);
  const uint32_t id144 = tc::instr_desc(128, NB7, tc::FMT_BF16, tc::MAJOR_K, tc::MAJOR_MN);
  const uint32_t id128 = tc::instr_desc(128, 128, tc::FMT_BF16, tc::MAJOR_K, tc::MAJOR_K);
  const uint32_t id64 = tc::instr_desc(128, 64, tc::FMT_BF16, tc::MAJOR_K, tc::MAJOR_K);
  const Opnd oR1 = A_IMG(sR1), oQf = A_IMG(sR1 + IMG), oWq = W_IMG(sW + P2_WQ, 64), oW0 = W_IMG(sW + P2_W0, 128),
             oW2 = W_IMG(sW + P2_W2, 64);
  const Opnd oB7[2] = {B7_IMG(tc::smem_u32(G + P2_B7)), B7_IMG(tc::smem_u32(G + P2_B7 + B7_BYTES))};

  const int ngroups = gridDim.x * 2, gg = blockIdx.x * 2 + g.gid;
  const int u0 = (int)((long long)a.n_units * gg / ngroups), u1 = (int)((long long)a.n_units * (gg + 1) / ngroups);
  const int pc = g.t & 63, pq = g.t >> 6;                                 // pooling: channel, row quarter
  // prefetch: the next tile's `a` image travels through registers (16 KB / 256 threads = 4 x 16 B each), the next
  // unit's attention operand through cp.async into the second B7 buffer, both while the current tile computes
  auto a_image_slot = [&](int slot_, int tile) { return a.A_in + (((size_t)slot_ * 2 + a.role) * a.NT + tile) * IMG; };
  auto a_image = [&](int u, int tile) { return a_image_slot(a.u_slot[u], tile); };
  uint4 pre[4];
  if (u0 < u1) {
    const uint4* src = reinterpret_cast<const uint4*>(a_image(u0, 0));
#pragma unroll
    for (int i = 0; i < 4; ++i) pre[i] = ldg_early(src + g.t + i * GT);
    copy_to_smem(G + P2_B7, a.B7_in + ((size_t)a.u_slot[u0] * 2 + (1 - a.role)) * B7_BYTES, B7_BYTES, g.t, GT);
  }
  cp_async_commit();
  int b7buf = 0;
  int ntr = 0;
  int slot_next = u0 < u1 ? a.u_slot[u0] : 0;
  for (int u = u0; u < u1; ++u) {
    const int slot = slot_next;
    if (u + 1 < u1) slot_next = a.u_slot[u + 1];                          // loaded a whole unit before it is needed
    float pmax = -INFINITY, psum = 0.f;
    for (int tile = 0; tile < a.NT; ++tile) {
      const uint8_t* a_img = a_image_slot(slot, tile);
      trace_mark(ntr, 100);
      cp_async_wait<0>();                                                 // this thread's share of B7 has landed
      g.sync();                                                           // previous tile's pooling reads of R1 are done
      trace_mark(ntr, 101);
#pragma unroll
      for (int i = 0; i < 4; ++i) reinterpret_cast<uint4*>(R1)[g.t + i * GT] = pre[i];
      g.publish();
      {   // prefetch the next tile (and, at the end of the unit, the next unit's attention operand)
        int nu = u, nt = tile + 1;
        if (nt == a.NT) { nu = u + 1; nt = 0; }
        if (nu < u1) {
          const int nslot = nt == 0 ? slot_next : slot;
          const uint4* src = reinterpret_cast<const uint4*>(a_image_slot(nslot, nt));
#pragma unroll
          for (int i = 0; i < 4; ++i) pre[i] = ldg_early(src + g.t + i * GT);
          if (nt == 0)
            copy_to_smem(G + P2_B7 + (1 - b7buf) * B7_BYTES, a.B7_in + ((size_t)nslot * 2 + (1 - a.role)) * B7_BYTES,
                         B7_BYTES, g.t, GT);
        }
        cp_async_commit();
      }
      trace_mark(ntr, 102);
      if (g.issuer) { if (tc::elect_one()) { issue_gemm<4>(g.tmem, oR1, oWq, id64, false); tc::umma_commit(g.bar); } __syncwarp(); }
      trace_mark(ntr, 103);
      uint4 sdA[4];
      g.wait();
      trace_mark(ntr, 104);
      feat32<true, false>(g, 32 * g.h, nullptr, R1 + IMG, 4 * g.h);       // Qf = elu(q)+1
      trace_mark(ntr, 105);
      g.publish();
      trace_mark(ntr, 106);
      if (g.issuer) { if (tc::elect_one()) { issue_gemm<4>(g.tmem, oQf, oB7[b7buf], id144, false); tc::umma_commit(g.bar); } __syncwarp(); }
      trace_mark(ntr, 107);
      g.wait();
      trace_mark(ntr, 108);
      epi_attn_ln(g, ln1, R1 + IMG);                                      // X next to a: [a | X] is the K=128 operand
      trace_mark(ntr, 109);
      g.publish();
      trace_mark(ntr, 110);
      if (g.issuer) { if (tc::elect_one()) { issue_gemm<8>(g.tmem, oR1, oW0, id128, false); tc::umma_commit(g.bar); } __syncwarp(); }
      trace_mark(ntr, 111);
      g.wait();
      trace_mark(ntr, 112);
      load_side<4>(sdA, a_img, 4 * g.h, g.row);                           // residual a, consumed after G9
      {
        uint4 none[8];
        epi_relu128<false>(g, none, R1);                                  // Hd over [a | X]
      }
      trace_mark(ntr, 113);
      g.publish();
      trace_mark(ntr, 114);
      if (g.issuer) { if (tc::elect_one()) { issue_gemm<8>(g.tmem, oR1, oW2, id64, false); tc::umma_commit(g.bar); } __syncwarp(); }
      trace_mark(ntr, 115);
      g.wait();
      trace_mark(ntr, 116);
      {
        float o[32];
        epi_ln_res(g, ln2, sdA, o);                                       // o = a + LN2(.)
        // transpose through shared memory (R1 is free: G9 has completed) with a rotation that keeps both the
        // row-wise writes and the channel-wise reads bank-conflict free
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const int ch = 32 * g.h + c;
          R1f[ch * 128 + ((g.row + ch) & 127)] = o[c];
        }
      }
      trace_mark(ntr, 117);
      g.sync();
      trace_mark(ntr, 118);
#pragma unroll 8
      for (int i = 0; i < 32; ++i) {
        const float v = R1f[pc * 128 + ((pq * 32 + i + pc) & 127)];
        pmax = fmaxf(pmax, v);
        psum += v;
      }
    }
    comb[g.gid][0][pq][pc] = pmax;
    comb[g.gid][1][pq][pc] = psum;
    g.sync();
    if (g.t < 64) {
      float* out = a.pool_part + ((size_t)slot * 2 + a.role) * 128;
      out[g.t] = fmaxf(fmaxf(comb[g.gid][0][0][g.t], comb[g.gid][0][1][g.t]), fmaxf(comb[g.gid][0][2][g.t], comb[g.gid][0][3][g.t]));
      out[64 + g.t] = (comb[g.gid][1][0][g.t] + comb[g.gid][1][1][g.t]) + (comb[g.gid][1][2][g.t] + comb[g.gid][1][3][g.t]);
    }
    b7buf ^= 1;
  }
  cp_async_wait<0>();
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(tmem_base_s, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// phase 2, three-tile variant: 3 groups of 4 warps per CTA (one thread per tile row, both column halves), so that
// three independent tiles are in flight per SM.  Same arithmetic as pair_p2_kernel.
// ---------------------------------------------------------------------------------------------------------------
constexpr int P2X_R1 = 0, P2X_B7 = 2 * IMG, P2X_GBYTES = P2X_B7 + B7_BYTES;      // 51200 B per group

__global__ void __launch_bounds__(NGX * GX, 1) pair_p2x_kernel(const P2Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[NGX];
  __shared__ uint32_t tmem_base_s;
  __shared__ float comb[NGX][2][4][64];
  uint8_t* Wsm = smem;
  const float* ln1 = reinterpret_cast<const float*>(Wsm + P2_LN);
  const float* ln2 = ln1 + 128;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NGX; ++i) tc::mbar_init(&bars[i], 1);
    tc::fence_mbar_init();
  }
  if (threadIdx.x < 32) { tc::tmem_alloc(&tmem_base_s, 512); tc::tmem_relinquish(); }
  copy_to_smem(Wsm, a.W, P2_WBYTES, threadIdx.x, NGX * GX);
  cp_async_commit();
  cp_async_wait<0>();
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  GroupX g;
  {
    const int warp_u = (int)tc::uniform(threadIdx.x >> 5);
    g.gid = warp_u / 4;
    g.t = threadIdx.x % GX;
    g.issuer = (warp_u % 4) == 0;
    g.tmem = tc::uniform(tmem_base_s) + g.gid * 160;
    g.tlane = g.tmem + ((uint32_t)((warp_u % 4) * 32) << 16);
    g.bar = bars + g.gid;
    g.par = 0;
  }
  uint8_t* G = smem + P2_WBYTES + g.gid * P2X_GBYTES;
  uint8_t* R1 = G + P2X_R1;
  uint8_t* B7 = G + P2X_B7;
  float* R1f = reinterpret_cast<float*>(R1);
  const uint32_t sR1 = tc::smem_u32(R1), sW = tc::smem_u32(Wsm);
  const uint32_t id144 = tc::instr_desc(128, NB7, tc::FMT_BF16, tc::MAJOR_K, tc::MAJOR_MN);
  const uint32_t id128 = tc::instr_desc(128, 128, tc::FMT_BF16, tc::MAJOR_K, tc::MAJOR_K);
  const uint32_t id64 = tc::instr_desc(128, 64, tc::FMT_BF16, tc::MAJOR_K, tc::MAJOR_K);
  const Opnd oR1 = A_IMG(sR1), oQf = A_IMG(sR1 + IMG), oWq = W_IMG(sW + P2_WQ, 64), oW0 = W_IMG(sW + P2_W0, 128),
             oW2 = W_IMG(sW + P2_W2, 64), oB7 = B7_IMG(tc::smem_u32(B7));
  const int row = g.t;
  uint8_t* arow = R1 + row * 16;            // this thread's row inside the operand images

  const int ngroups = gridDim.x * NGX, gg = blockIdx.x * NGX + g.gid;
  const int u0 = (int)((long long)a.n_units * gg / ngroups), u1 = (int)((long long)a.n_units * (gg + 1) / ngroups);
  const int pc = g.t & 31, pq = g.t >> 5;                                 // pooling: channel within a 32-channel pass, row quarter
  float* Tb = reinterpret_cast<float*>(R1 + IMG);                         // 16 KB transpose buffer: [32 ch][128 rows] fp32, rotated
  int slot_next = u0 < u1 ? a.u_slot[u0] : 0;
  if (u0 < u1) {   // first tile of this group: nothing to hide the loads behind
    copy_to_smem(R1, a.A_in + ((size_t)slot_next * 2 + a.role) * a.NT * IMG, IMG, g.t, GX);
    copy_to_smem(B7, a.B7_in + ((size_t)slot_next * 2 + (1 - a.role)) * B7_BYTES, B7_BYTES, g.t, GX);
  }
  cp_async_commit();
  for (int u = u0; u < u1; ++u) {
    const int slot = slot_next;
    if (u + 1 < u1) slot_next = a.u_slot[u + 1];
    float pmax0 = -INFINITY, psum0 = 0.f, pmax1 = -INFINITY, psum1 = 0.f;  // channels pc and 32 + pc
    for (int tile = 0; tile < a.NT; ++tile) {
      const uint8_t* a_img = a.A_in + (((size_t)slot * 2 + a.role) * a.NT + tile) * IMG;
      cp_async_wait<0>();                                                 // `a` image (and B7 at a unit start) prefetched earlier
      g.publish();
      if (g.issuer) { if (tc::elect_one()) { issue_gemm<4>(g.tmem, oR1, oWq, id64, false); tc::umma_commit(g.bar); } __syncwarp(); }
      g.wait();
      {   // Qf = elu(q)+1 -> second half of R1
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t r[16];
          tc::tmem_ld16(g.tlane + 16 * q, r);
          tc::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              w[j] = tc::bf2_elu1(tc::pack_bf16(__uint_as_float(r[c * 8 + 2 * j]), __uint_as_float(r[c * 8 + 2 * j + 1])));
            *reinterpret_cast<uint4*>(arow + IMG + (2 * q + c) * 2048) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
      g.publish();
      if (g.issuer) { if (tc::elect_one()) { issue_gemm<4>(g.tmem, oQf, oB7, id144, false); tc::umma_commit(g.bar); } __syncwarp(); }
      g.wait();
      if (tile + 1 == a.NT && u + 1 < u1) {   // last attention GEMM of the unit is done: next unit's B7 streams in
        copy_to_smem(B7, a.B7_in + ((size_t)slot_next * 2 + (1 - a.role)) * B7_BYTES, B7_BYTES, g.t, GX);
        cp_async_commit();
      }
      {   // attention epilogue: z-normalise, merge heads, LayerNorm1 -> X (second half of R1)
        uint32_t d8[8];
        tc::tmem_ld8(g.tlane + 128, d8);
        tc::tmem_ld_wait();
        const float z0 = 1.f / (__uint_as_float(d8[0]) + ATT_EPS), z1 = 1.f / (__uint_as_float(d8[1]) + ATT_EPS);
        float m0[32], m1[32];
        float s = 0.f, ss = 0.f;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float (&m)[32] = hh == 0 ? m0 : m1;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t r0[16], r1[16];
            tc::tmem_ld16(g.tlane + 32 * hh + 16 * half, r0);
            tc::tmem_ld16(g.tlane + 64 + 32 * hh + 16 * half, r1);
            tc::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float v = fmaf(z0, __uint_as_float(r0[j]), z1 * __uint_as_float(r1[j]));
              m[16 * half + j] = v;
              s += v;
              ss = fmaf(v, v, ss);
            }
          }
        }
        const float mean = s * (1.f / 64.f);
        const float rstd = rsqrtf(fmaxf(ss * (1.f / 64.f) - mean * mean, 0.f) + LN_EPS);
        ln_apply_store(m0, mean, rstd, ln1, 0, arow + IMG);
        ln_apply_store(m1, mean, rstd, ln1, 32, arow + IMG + 4 * 2048);
      }
      g.publish();
      if (g.issuer) { if (tc::elect_one()) { issue_gemm<8>(g.tmem, oR1, oW0, id128, false); tc::umma_commit(g.bar); } __syncwarp(); }
      uint4 sdA[8];
      g.wait();
      load_side<8>(sdA, a_img, 0, row);                                   // residual a (this tile), consumed after G9
      {   // G8 has consumed [a | X]: the next tile's `a` image streams into R1 behind the rest of this tile
        int nu = u, nt = tile + 1;
        if (nt == a.NT) { nu = u + 1; nt = 0; }
        if (nu < u1) {
          copy_to_smem(R1, a.A_in + (((size_t)(nt == 0 ? slot_next : slot) * 2 + a.role) * a.NT + nt) * IMG, IMG, g.t, GX);
          cp_async_commit();
        }
      }
      {   // Hd = relu(acc) -> bf16, written back in place to TMEM columns [0, 64): the A operand of G9 (no shared memory)
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          uint32_t r[16];
          tc::tmem_ld16(g.tlane + 16 * q, r);
          tc::tmem_ld_wait();
          uint32_t w[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) w[j] = tc::bf2_max(tc::pack_bf16(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1])), 0u);
          tc::tmem_st8(g.tlane + 8 * q, w);
        }
        tc::tmem_st_wait();
      }
      tc::tc_fence_before();
      g.sync();
      tc::tc_fence_after();
      if (g.issuer) {
        if (tc::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            tc::umma_f16_ts(g.tmem + 64, g.tmem + 8 * ks, oW2.desc + (uint64_t)(ks * oW2.kstep), id64, ks > 0 ? 1u : 0u);
          tc::umma_commit(g.bar);
        }
        __syncwarp();
      }
      g.wait();
      {   // o = a + LN2(acc); pooled over the points through a 16 KB rotated transpose buffer, 32 channels per pass
        float o0[32], o1[32];
        float s = 0.f, ss = 0.f;
        ld32_stats(g.tlane + 64, o0, s, ss);
        ld32_stats(g.tlane + 96, o1, s, ss);
        const float mean = s * (1.f / 64.f);
        const float rstd = rsqrtf(fmaxf(ss * (1.f / 64.f) - mean * mean, 0.f) + LN_EPS);
        const float nm = -mean * rstd;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float (&o)[32] = hh == 0 ? o0 : o1;
          if (hh == 1) g.sync();                                          // pass-0 reads are done before pass 1 overwrites
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 rs = sdA[4 * hh + c];
            const uint32_t rw[4] = {rs.x, rs.y, rs.z, rs.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int k = c * 8 + 2 * j, ch = 32 * hh + k;
              const float y0 = fmaf(fmaf(o[k], rstd, nm), ln2[ch], ln2[64 + ch]) + bf_lo(rw[j]);
              const float y1 = fmaf(fmaf(o[k + 1], rstd, nm), ln2[ch + 1], ln2[64 + ch + 1]) + bf_hi(rw[j]);
              Tb[k * 128 + ((row + k) & 127)] = y0;
              Tb[(k + 1) * 128 + ((row + k + 1) & 127)] = y1;
            }
          }
          g.sync();
          float mx = -INFINITY, sm = 0.f;
#pragma unroll 8
          for (int i = 0; i < 32; ++i) {
            const float v = Tb[pc * 128 + ((pq * 32 + i + pc) & 127)];
            mx = fmaxf(mx, v);
            sm += v;
          }
          if (hh == 0) { pmax0 = fmaxf(pmax0, mx); psum0 += sm; } else { pmax1 = fmaxf(pmax1, mx); psum1 += sm; }
        }
      }
      tc::tc_fence_before();
      g.sync();                                                           // transpose buffer (= Qf region) is rewritten by the next tile
    }
    comb[g.gid][0][pq][pc] = pmax0;
    comb[g.gid][0][pq][32 + pc] = pmax1;
    comb[g.gid][1][pq][pc] = psum0;
    comb[g.gid][1][pq][32 + pc] = psum1;
    g.sync();
    if (g.t < 64) {
      float* out = a.pool_part + ((size_t)slot * 2 + a.role) * 128;
      out[g.t] = fmaxf(fmaxf(comb[g.gid][0][0][g.t], comb[g.gid][0][1][g.t]), fmaxf(comb[g.gid][0][2][g.t], comb[g.gid][0][3][g.t]));
      out[64 + g.t] = (comb[g.gid][1][0][g.t] + comb[g.gid][1][1][g.t]) + (comb[g.gid][1][2][g.t] + comb[g.gid][1][3][g.t]);
    }
  }
  cp_async_wait<0>();
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(tmem_base_s, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// phase 1, three-tile variants: the monolithic phase-1 kernel needs 72 KB of shared memory and 224 TMEM columns per
// group, which caps it at two tiles in flight.  Split in two kernels that each fit three 4-warp groups per SM:
//   pair_p1a_kernel : G1 (attention, smem operands) -> LN1 -> G2 (+U, ReLU; result written back to TMEM in place as the
//                     bf16 A operand of G3: tcgen05.st, no shared memory) -> G3 -> LN2 + h -> a (global, bf16 image)
//   pair_p1b_kernel : a -> G4k -> Kf, G4v -> V (+Wv pos) -> G5 (KV accumulated in TMEM over the tiles) -> G6 -> B7 (global)
// ---------------------------------------------------------------------------------------------------------------
constexpr int P1A_W0B = 0, P1A_W2 = 16384, P1A_LN = 32768, P1A_WBYTES = 32768 + 1024;
constexpr int P1A_QXA = 0, P1A_MK1 = IMG, P1A_GBYTES = IMG + B7_BYTES;                       // 34816 B per group
constexpr int P1B_WKV = 0, P1B_WM = 16384, P1B_WBYTES = 24576;
constexpr int P1B_AIMG = 0, P1B_KFV = IMG, P1B_GBYTES = IMG + 2 * IMG + ONES_BYTES;          // 53248 B per group

__global__ void __launch_bounds__(NGX * GX, 1) pair_p1a_kernel(const P1Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[NGX];
  __shared__ uint32_t tmem_base_s;
  uint8_t* Wsm = smem;
  const float* ln1 = reinterpret_cast<const float*>(Wsm + P1A_LN);
  const float* ln2 = ln1 + 128;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NGX; ++i) tc::mbar_init(&bars[i], 1);
    tc::fence_mbar_init();
  }
  if (threadIdx.x < 32) { tc::tmem_alloc(&tmem_base_s, 512); tc::tmem_relinquish(); }
  copy_to_smem(Wsm, a.W, P1A_WBYTES, threadIdx.x, NGX * GX);
  cp_async_commit();
  cp_async_wait<0>();
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  GroupX g;
  groupx_setup(g, bars, tmem_base_s);
  uint8_t* G = smem + P1A_WBYTES + g.gid * P1A_GBYTES;
  uint8_t* QXa = G + P1A_QXA;
  uint8_t* MK1 = G + P1A_MK1;
  const uint32_t sQXa = tc::smem_u32(QXa), sW = tc::smem_u32(Wsm);
  const uint32_t id144 = tc::instr_desc(128, NB7, tc::FMT_BF16, tc::MAJOR_K, tc::MAJOR_MN);
  const uint32_t id128 = tc::instr_desc(128, 128, tc::FMT_BF16, tc::MAJOR_K, tc::MAJOR_K);
  const uint32_t id64 = tc::instr_desc(128, 64, tc::FMT_BF16, tc::MAJOR_K, tc::MAJOR_K);
  const Opnd oQXa = A_IMG(sQXa), oMK1 = B7_IMG(tc::smem_u32(MK1)), oW0b = W_IMG(sW + P1A_W0B, 128), oW2 = W_IMG(sW + P1A_W2, 64);
  const int row = g.t;
  uint8_t* xrow = QXa + row * 16;

  const int ngroups = gridDim.x * NGX, gg = blockIdx.x * NGX + g.gid;
  const int u0 = (int)((long long)a.n_units * gg / ngroups), u1 = (int)((long long)a.n_units * (gg + 1) / ngroups);
  int cur_templ = -1;
  bool prefetched = false;
  int so_next = u0 < u1 ? a.u_search[u0] : 0, te_next = u0 < u1 ? a.u_templ[u0] : 0, slot_next = u0 < u1 ? a.u_slot[u0] : 0;
  for (int u = u0; u < u1; ++u) {
    const int so = so_next, te = te_next, slot = slot_next;
    if (u + 1 < u1) { so_next = a.u_search[u + 1]; te_next = a.u_templ[u + 1]; slot_next = a.u_slot[u + 1]; }
    for (int tile = 0; tile < a.NT; ++tile) {
      const size_t ti = (size_t)so * a.NT + tile;
      if (!prefetched) copy_to_smem(QXa, a.QF1 + ti * IMG, IMG, g.t, GX);  // first tile of this group only
      prefetched = false;
      if (te != cur_templ) { copy_to_smem(MK1, a.MK1 + (size_t)te * B7_BYTES, B7_BYTES, g.t, GX); cur_templ = te; }
      cp_async_commit();
      cp_async_wait<0>();
      g.publish();
      if (g.issuer) { if (tc::elect_one()) { issue_gemm<4>(g.tmem, oQXa, oMK1, id144, false); tc::umma_commit(g.bar); } __syncwarp(); }
      g.wait();
      {   // z-normalise, merge heads, LayerNorm1 -> X (over the query image)
        uint32_t d8[8];
        tc::tmem_ld8(g.tlane + 128, d8);
        tc::tmem_ld_wait();
        const float z0 = 1.f / (__uint_as_float(d8[0]) + ATT_EPS), z1 = 1.f / (__uint_as_float(d8[1]) + ATT_EPS);
        float m0[32], m1[32];
        float s = 0.f, ss = 0.f;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float (&m)[32] = hh == 0 ? m0 : m1;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t r0[16], r1[16];
            tc::tmem_ld16(g.tlane + 32 * hh + 16 * half, r0);
            tc::tmem_ld16(g.tlane + 64 + 32 * hh + 16 * half, r1);
            tc::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float v = fmaf(z0, __uint_as_float(r0[j]), z1 * __uint_as_float(r1[j]));
              m[16 * half + j] = v;
              s += v;
              ss = fmaf(v, v, ss);
            }
          }
        }
        const float mean = s * (1.f / 64.f);
        const float rstd = rsqrtf(fmaxf(ss * (1.f / 64.f) - mean * mean, 0.f) + LN_EPS);
        ln_apply_store(m0, mean, rstd, ln1, 0, xrow);
        ln_apply_store(m1, mean, rstd, ln1, 32, xrow + 4 * 2048);
      }
      g.publish();
      if (g.issuer) { if (tc::elect_one()) { issue_gemm<4>(g.tmem, oQXa, oW0b, id128, false); tc::umma_commit(g.bar); } __syncwarp(); }
      {   // Hd = relu(acc + U) -> bf16, written back IN PLACE to TMEM columns [0, 64): the A operand of G3.
          // One thread owns one lane, and the packed write [8q, 8q+8) never passes the unread columns >= 16(q+1).
        uint4 sdU[16];
        load_side<16>(sdU, a.U + ti * 2 * IMG, 0, row);
        g.wait();
        {   // G2 has consumed X: the query image of the next (unit, tile) streams into QXa behind the rest of this tile
          int nu = u, nt = tile + 1;
          if (nt == a.NT) { nu = u + 1; nt = 0; }
          if (nu < u1) {
            copy_to_smem(QXa, a.QF1 + ((size_t)(nt == 0 ? so_next : so) * a.NT + nt) * IMG, IMG, g.t, GX);
            cp_async_commit();
            prefetched = true;
          }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          uint32_t r[16];
          tc::tmem_ld16(g.tlane + 16 * q, r);
          tc::tmem_ld_wait();
          uint32_t w[8];
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const uint4 s4 = sdU[2 * q + c];
            const uint32_t sw[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
              w[c * 4 + j] = tc::bf2_max(tc::bf2_add(tc::pack_bf16(__uint_as_float(r[c * 8 + 2 * j]), __uint_as_float(r[c * 8 + 2 * j + 1])), sw[j]), 0u);
          }
          tc::tmem_st8(g.tlane + 8 * q, w);
        }
        tc::tmem_st_wait();
      }
      tc::tc_fence_before();
      g.sync();
      tc::tc_fence_after();
      if (g.issuer) {
        if (tc::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            tc::umma_f16_ts(g.tmem + 64, g.tmem + 8 * ks, oW2.desc + (uint64_t)(ks * oW2.kstep), id64, ks > 0 ? 1u : 0u);
          tc::umma_commit(g.bar);
        }
        __syncwarp();
      }
      {   // a = h + LN2(acc) -> global bf16 image (stage-1 output)
        uint4 sdH[8];
        load_side<8>(sdH, a.H + ti * IMG, 0, row);
        g.wait();
        float o0[32], o1[32];
        float s = 0.f, ss = 0.f;
        ld32_stats(g.tlane + 64, o0, s, ss);
        ld32_stats(g.tlane + 96, o1, s, ss);
        const float mean = s * (1.f / 64.f);
        const float rstd = rsqrtf(fmaxf(ss * (1.f / 64.f) - mean * mean, 0.f) + LN_EPS);
        const float nm = -mean * rstd;
        uint8_t* orow = a.A_out + (((size_t)slot * 2 + a.role) * a.NT + tile) * IMG + row * 16;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float (&o)[32] = hh == 0 ? o0 : o1;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 rs = sdH[4 * hh + c];
            const uint32_t rw[4] = {rs.x, rs.y, rs.z, rs.w};
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int k = c * 8 + 2 * j, ch = 32 * hh + k;
              const float y0 = fmaf(fmaf(o[k], rstd, nm), ln2[ch], ln2[64 + ch]) + bf_lo(rw[j]);
              const float y1 = fmaf(fmaf(o[k + 1], rstd, nm), ln2[ch + 1], ln2[64 + ch + 1]) + bf_hi(rw[j]);
              w[j] = tc::pack_bf16(y0, y1);
            }
            *reinterpret_cast<uint4*>(orow + (4 * hh + c) * 2048) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
      tc::tc_fence_before();                                              // TMEM reads done before the next tile's G1 overwrites
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(tmem_base_s, 512);
}

__global__ void __launch_bounds__(NGX * GX, 1) pair_p1b_kernel(const P1Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[2 * NGX];
  __shared__ uint32_t tmem_base_s;
  uint8_t* Wsm = smem;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * NGX; ++i) tc::mbar_init(&bars[i], 1);
    tc::fence_mbar_init();
  }
  if (threadIdx.x < 32) { tc::tmem_alloc(&tmem_base_s, 512); tc::tmem_relinquish(); }
  copy_to_smem(Wsm, a.W, P1B_WBYTES, threadIdx.x, NGX * GX);
  cp_async_commit();
  cp_async_wait<0>();
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  GroupX g;
  groupx_setup(g, bars, tmem_base_s);
  uint64_t* bar2 = bars + NGX + g.gid;
  uint32_t par2 = 0;
  uint8_t* G = smem + P1B_WBYTES + g.gid * P1B_GBYTES;
  uint8_t* Aimg = G + P1B_AIMG;
  uint8_t* KfV = G + P1B_KFV;
  {
    uint4* ones = reinterpret_cast<uint4*>(KfV + 2 * IMG);
    ones[g.t] = make_uint4(0x00003f80u, 0, 0, 0);
    ones[128 + g.t] = make_uint4(0, 0, 0, 0);
  }
  const uint32_t sA = tc::smem_u32(Aimg), sKfV = tc::smem_u32(KfV), sW = tc::smem_u32(Wsm);
  const uint32_t id64 = tc::instr_desc(128, 64, tc::FMT_BF16, tc::MAJOR_K, tc::MAJOR_K);
  const uint32_t idkv = tc::instr_desc(128, 80, tc::FMT_BF16, tc::MAJOR_MN, tc::MAJOR_MN);
  // Wkv image is [k/8][128 rows: Wk 0..63 | Wv 64..127][8]: an N=64 operand is the same image entered at row 0 / row 64
  const Opnd oA = A_IMG(sA), oWk = W_IMG(sW + P1B_WKV, 128), oWv = W_IMG(sW + P1B_WKV + 64 * 16, 128), oWm = W_IMG(sW + P1B_WM, 64),
             oKfV = opnd(sKfV, 128u, 2048u, 256u), oVones = opnd(sKfV + 8 * 2048, 128u, 2048u, 256u);
  const int row = g.t;
  const uint32_t KVC = 64;                                                // TMEM columns [64, 144): KV / Ksum accumulator

  const int ngroups = gridDim.x * NGX, gg = blockIdx.x * NGX + g.gid;
  const int u0 = (int)((long long)a.n_units * gg / ngroups), u1 = (int)((long long)a.n_units * (gg + 1) / ngroups);
  int so_next = u0 < u1 ? a.u_search[u0] : 0, slot_next = u0 < u1 ? a.u_slot[u0] : 0;
  bool prefetched = false;
  for (int u = u0; u < u1; ++u) {
    const int so = so_next, slot = slot_next;
    if (u + 1 < u1) { so_next = a.u_search[u + 1]; slot_next = a.u_slot[u + 1]; }
    for (int tile = 0; tile < a.NT; ++tile) {
      const size_t ti = (size_t)so * a.NT + tile;
      if (!prefetched) copy_to_smem(Aimg, a.A_out + (((size_t)slot * 2 + a.role) * a.NT + tile) * IMG, IMG, g.t, GX);
      prefetched = false;
      cp_async_commit();
      {   // L2 prefetch of the image after this one (the shared-memory copy can only start once both projections have read Aimg)
        int ns = slot, nt = tile + 1;
        if (nt == a.NT) { ns = slot_next; nt = 0; }
        if (nt != 0 || u + 1 < u1) prefetch_l2_16k(a.A_out + (((size_t)ns * 2 + a.role) * a.NT + nt) * IMG, g.t);
      }
      cp_async_wait<0>();
      g.publish();
      if (g.issuer) { if (tc::elect_one()) { issue_gemm<4>(g.tmem, oA, oWk, id64, false); tc::umma_commit(g.bar); } __syncwarp(); }
      g.wait();
      if (tile > 0) { tc::mbar_wait(bar2, par2); par2 ^= 1u; tc::tc_fence_after(); }   // previous KV GEMM still reads KfV
      {   // Kf = elu(k)+1 -> chunks 0..7 ; zero for padding rows (point index >= npts): they must not enter KV / Ksum
        const bool keep = tile * 128 + row < a.npts;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t r[16];
          tc::tmem_ld16(g.tlane + 16 * q, r);
          tc::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              w[j] = keep ? tc::bf2_elu1(tc::pack_bf16(__uint_as_float(r[c * 8 + 2 * j]), __uint_as_float(r[c * 8 + 2 * j + 1]))) : 0u;
            *reinterpret_cast<uint4*>(KfV + (2 * q + c) * 2048 + row * 16) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
      tc::tc_fence_before();
      g.sync();                                                           // everybody has read k before v overwrites the columns
      tc::tc_fence_after();
      if (g.issuer) { if (tc::elect_one()) { issue_gemm<4>(g.tmem, oA, oWv, id64, false); tc::umma_commit(g.bar); } __syncwarp(); }
      {   // V = v + Wv pos -> chunks 8..15
        uint4 sdPV[8];
        load_side<8>(sdPV, a.PV + ti * IMG, 0, row);
        g.wait();
        if (tile + 1 < a.NT) {   // both projections have consumed `a`: stream the unit's next tile in behind the V epilogue + KV GEMM
          copy_to_smem(Aimg, a.A_out + (((size_t)slot * 2 + a.role) * a.NT + tile + 1) * IMG, IMG, g.t, GX);
          cp_async_commit();
          prefetched = true;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t r[16];
          tc::tmem_ld16(g.tlane + 16 * q, r);
          tc::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const uint4 s4 = sdPV[2 * q + c];
            const uint32_t sw[4] = {s4.x, s4.y, s4.z, s4.w};
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              w[j] = tc::bf2_add(tc::pack_bf16(__uint_as_float(r[c * 8 + 2 * j]), __uint_as_float(r[c * 8 + 2 * j + 1])), sw[j]);
            *reinterpret_cast<uint4*>(KfV + (8 + 2 * q + c) * 2048 + row * 16) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
      g.publish();
      if (g.issuer) {   // KV += [Kf|V]^T [V|1]
        if (tc::elect_one()) { issue_gemm<8>(g.tmem + KVC, oKfV, oVones, idkv, tile > 0); tc::umma_commit(bar2); }
        __syncwarp();
      }
    }
    tc::mbar_wait(bar2, par2);
    par2 ^= 1u;
    tc::tc_fence_after();
    {   // B7 = [head-split blockdiag(KV) Wm^T | Ksum dots] of this (pair, direction) as template
      float kv[64];
      float ksum = 0.f;
      if (row < 64) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t r[16];
          tc::tmem_ld16(g.tlane + KVC + 16 * q, r);
          tc::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) kv[16 * q + j] = __uint_as_float(r[j]);
        }
        uint32_t r8[8];
        tc::tmem_ld8(g.tlane + KVC + 64, r8);
        tc::tmem_ld_wait();
        ksum = __uint_as_float(r8[0]);
        const int hd = row >> 5;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t w[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) w[j] = ((c >> 2) == hd) ? tc::pack_bf16(kv[c * 8 + 2 * j], kv[c * 8 + 2 * j + 1]) : 0u;
          *reinterpret_cast<uint4*>(Aimg + c * 2048 + row * 16) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(Aimg + c * 2048 + row * 16) = make_uint4(0, 0, 0, 0);
      }
      g.publish();
      if (g.issuer) { if (tc::elect_one()) { issue_gemm<4>(g.tmem, oA, oWm, id64, false); tc::umma_commit(g.bar); } __syncwarp(); }
      g.wait();
      if (u + 1 < u1) {   // G6 has consumed the operand buffer: the next unit's first tile streams in behind the B7 write-out
        copy_to_smem(Aimg, a.A_out + ((size_t)slot_next * 2 + a.role) * a.NT * IMG, IMG, g.t, GX);
        cp_async_commit();
        prefetched = true;
      }
      if (row < 64) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t r[16];
          tc::tmem_ld16(g.tlane + 16 * q, r);
          tc::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) kv[16 * q + j] = __uint_as_float(r[j]);
        }
        write_b7_row(kv, ksum, row, a.B7_out + ((size_t)slot * 2 + a.role) * B7_BYTES);
      }
      tc::tc_fence_before();
      g.sync();
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(tmem_base_s, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// packing kernels (fp32 per-object tensors of the parity path -> bf16 operand images)
// ---------------------------------------------------------------------------------------------------------------
// src (B, C, N) channel-major fp32 -> dst [B][N/128][C/8][128][8] bf16, optional elu+1
__global__ void __launch_bounds__(256) pack_image_kernel(int B, int C, int N, const float* __restrict__ src, long long s_bs,
                                                         int lds, int act, uint8_t* __restrict__ dst) {
  const int nt = (N + 127) / 128, nch = C / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int row = (int)(idx % 128);
  const int chunk = (int)((idx / 128) % nch);
  const int tile = (int)((idx / (128LL * nch)) % nt);
  const long long b = idx / (128LL * nch * nt);
  if (b >= B) return;
  const float* s = src + b * s_bs + (size_t)(chunk * 8) * lds + tile * 128 + row;
  uint32_t w[4] = {0u, 0u, 0u, 0u};                   // rows beyond N (last tile of a ragged object) are zero padding
  if (tile * 128 + row < N) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float x0 = s[(size_t)(2 * j) * lds], x1 = s[(size_t)(2 * j + 1) * lds];
      if (act == ACT_ELU1) { x0 = x0 > 0.f ? x0 + 1.f : expf(x0); x1 = x1 > 0.f ? x1 + 1.f : expf(x1); }
      w[j] = tc::pack_bf16(x0, x1);
    }
  }
  *reinterpret_cast<uint4*>(dst + (((size_t)b * nt + tile) * nch + chunk) * 2048 + row * 16) = make_uint4(w[0], w[1], w[2], w[3]);
}

// M (B, 64 d, 64 out) fp32 [= blockdiag(KV) Wm^T rows, before head masking] + ksum (B, 64) -> attention operand images
__global__ void __launch_bounds__(64) pack_b7_kernel(const float* __restrict__ M, const float* __restrict__ ksum,
                                                     uint8_t* __restrict__ dst) {
  const int b = blockIdx.x, d = threadIdx.x;
  float m[64];
#pragma unroll
  for (int j = 0; j < 64; ++j) m[j] = M[((size_t)b * 64 + d) * 64 + j];
  write_b7_row(m, ksum[(size_t)b * 64 + d], d, dst + (size_t)b * B7_BYTES);
}

// pool_part (P, 2, 128) -> pooled^T (128, P): max over both directions | mean over the 2*npts points
__global__ void __launch_bounds__(256) pool_finish_kernel(int P, int npts, const float* __restrict__ part, float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 128LL * P) return;
  const int c = (int)(idx / P), p = (int)(idx % P);
  const float x0 = part[((size_t)p * 2) * 128 + c], x1 = part[((size_t)p * 2 + 1) * 128 + c];
  out[idx] = c < 64 ? fmaxf(x0, x1) : (x0 + x1) / (float)(2 * npts);
}

}  // namespace

extern "C" {

int pcreid_pair_tc_set_trace(void* dev_buffer) {   /* debug: device buffer of >= 2048 int64, or NULL to disable */
  long long* p = (long long*)dev_buffer;
  return cudaMemcpyToSymbol(g_trace, &p, sizeof(p)) == cudaSuccess ? PCREID_OK : PCREID_ERR_LAUNCH;
}

int pcreid_pair_tc_smem_bytes(int phase) { return phase == 1 ? P1_WBYTES + 2 * P1_GBYTES : P2_WBYTES + 2 * P2_GBYTES; }

int pcreid_pack_image(int B, int C, int N, const float* src, long long s_bs, int lds, int act, void* dst, void* stream) {
  if (B <= 0) return PCREID_OK;
  if (!src || !dst || C % 8 || N <= 0) return PCREID_ERR_ARG;
  const long long per = 128LL * (C / 8) * ((N + 127) / 128);
  pack_image_kernel<<<(unsigned)((per * B + 255) / 256), 256, 0, (cudaStream_t)stream>>>(B, C, N, src, s_bs, lds, act, (uint8_t*)dst);
  return pcreid_launch_status();
}

int pcreid_pack_b7(int B, const float* M, const float* ksum, void* dst, void* stream) {
  if (B <= 0) return PCREID_OK;
  if (!M || !ksum || !dst) return PCREID_ERR_ARG;
  pack_b7_kernel<<<B, 64, 0, (cudaStream_t)stream>>>(M, ksum, (uint8_t*)dst);
  return pcreid_launch_status();
}

int pcreid_pool_finish(int P, int npts, const float* part, float* out, void* stream) {
  if (P <= 0) return PCREID_OK;
  if (!part || !out) return PCREID_ERR_ARG;
  pool_finish_kernel<<<(unsigned)((128LL * P + 255) / 256), 256, 0, (cudaStream_t)stream>>>(P, npts, part, out);
  return pcreid_launch_status();
}

int pcreid_pair_p1(int n_units, int NT, int role, const int* u_search, const int* u_templ, const int* u_slot, const void* QF1,
                   const void* U, const void* H, const void* PV, const void* MK1, const void* W, void* A_out, void* B7_out,
                   int n_ctas, void* stream) {
  if (n_units <= 0) return PCREID_OK;
  if (!u_search || !u_templ || !u_slot || !QF1 || !U || !H || !PV || !MK1 || !W || !A_out || !B7_out || NT <= 0) return PCREID_ERR_ARG;
  P1Args a{n_units, NT, role, 128 * NT, u_search, u_templ, u_slot, (const uint8_t*)QF1, (const uint8_t*)U, (const uint8_t*)H,
           (const uint8_t*)PV, (const uint8_t*)MK1, (const uint8_t*)W, (uint8_t*)A_out, (uint8_t*)B7_out};
  const int smem = P1_WBYTES + 2 * P1_GBYTES;
  cudaFuncSetAttribute(pair_p1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int grid = n_ctas > 0 ? n_ctas : 148;
  if (grid * 2 > n_units) grid = (n_units + 1) / 2;
  pair_p1_kernel<<<grid, 2 * GT, smem, (cudaStream_t)stream>>>(a);
  return pcreid_launch_status();
}

int pcreid_pair_p1ab(int which, int n_units, int NT, int role, const int* u_search, const int* u_templ, const int* u_slot,
                     const void* QF1, const void* U, const void* H, const void* PV, const void* MK1, const void* W, void* A_out,
                     void* B7_out, int n_ctas, void* stream) {
  if (n_units <= 0) return PCREID_OK;
  if (!u_search || !u_templ || !u_slot || !W || !A_out || !B7_out || NT <= 0 || (which != 0 && which != 1)) return PCREID_ERR_ARG;
  if (which == 0 && (!QF1 || !U || !H || !MK1)) return PCREID_ERR_ARG;
  if (which == 1 && !PV) return PCREID_ERR_ARG;
  P1Args a{n_units, NT, role, 128 * NT, u_search, u_templ, u_slot, (const uint8_t*)QF1, (const uint8_t*)U, (const uint8_t*)H,
           (const uint8_t*)PV, (const uint8_t*)MK1, (const uint8_t*)W, (uint8_t*)A_out, (uint8_t*)B7_out};
  int grid = n_ctas > 0 ? n_ctas : 148;
  if (grid * NGX > n_units) grid = (n_units + NGX - 1) / NGX;
  if (which == 0) {
    const int smem = P1A_WBYTES + NGX * P1A_GBYTES;
    cudaFuncSetAttribute(pair_p1a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    pair_p1a_kernel<<<grid, NGX * GX, smem, (cudaStream_t)stream>>>(a);
  } else {
    const int smem = P1B_WBYTES + NGX * P1B_GBYTES;
    cudaFuncSetAttribute(pair_p1b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    pair_p1b_kernel<<<grid, NGX * GX, smem, (cudaStream_t)stream>>>(a);
  }
  return pcreid_launch_status();
}

int pcreid_pair_p1b_n(int n_units, int npts, int role, const int* u_search, const int* u_templ, const int* u_slot, const void* PV,
                      const void* W, void* A_out, void* B7_out, int n_ctas, void* stream) {
  if (n_units <= 0) return PCREID_OK;
  if (!u_search || !u_templ || !u_slot || !PV || !W || !A_out || !B7_out || npts <= 0) return PCREID_ERR_ARG;
  const int NT = (npts + 127) / 128;
  P1Args a{n_units, NT, role, npts, u_search, u_templ, u_slot, nullptr, nullptr, nullptr, (const uint8_t*)PV, nullptr,
           (const uint8_t*)W, (uint8_t*)A_out, (uint8_t*)B7_out};
  int grid = n_ctas > 0 ? n_ctas : 148;
  if (grid * NGX > n_units) grid = (n_units + NGX - 1) / NGX;
  const int smem = P1B_WBYTES + NGX * P1B_GBYTES;
  cudaFuncSetAttribute(pair_p1b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  pair_p1b_kernel<<<grid, NGX * GX, smem, (cudaStream_t)stream>>>(a);
  return pcreid_launch_status();
}

int pcreid_pair_p2(int n_units, int NT, int role, const int* u_slot, const void* A_in, const void* B7_in, const void* W,
                   float* pool_part, int n_ctas, void* stream) {
  if (n_units <= 0) return PCREID_OK;
  if (!u_slot || !A_in || !B7_in || !W || !pool_part || NT <= 0) return PCREID_ERR_ARG;
  P2Args a{n_units, NT, role, 128 * NT, u_slot, (const uint8_t*)A_in, (const uint8_t*)B7_in, (const uint8_t*)W, pool_part};
  if (n_ctas < 0) {   // three-tile variant (3 groups x 4 warps per CTA)
    const int smemx = P2_WBYTES + NGX * P2X_GBYTES;
    cudaFuncSetAttribute(pair_p2x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smemx);
    int gridx = -n_ctas;
    if (gridx * NGX > n_units) gridx = (n_units + NGX - 1) / NGX;
    pair_p2x_kernel<<<gridx, NGX * GX, smemx, (cudaStream_t)stream>>>(a);
    return pcreid_launch_status();
  }
  const int smem = P2_WBYTES + 2 * P2_GBYTES;
  cudaFuncSetAttribute(pair_p2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int grid = n_ctas > 0 ? n_ctas : 148;
  if (grid * 2 > n_units) grid = (n_units + 1) / 2;
  pair_p2_kernel<<<grid, 2 * GT, smem, (cudaStream_t)stream>>>(a);
  return pcreid_launch_status();
}

}  // extern "C"
