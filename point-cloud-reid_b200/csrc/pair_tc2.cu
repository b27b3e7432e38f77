// Fused all-pairs `xcorr_eff` match, second generation of the phase-1a / phase-2 kernels (pair_tc.cu holds the first).
//
// Same GEMM chain and operand images as pair_p1a_kernel / pair_p2x_kernel; what changed is the SIMT side.  ncu on the
// first generation (profiles/r01_ncu_pair_kernels.md) showed 8.8 k warp-instructions per 128-point tile in phase 2
// against a 1.2 k-cycle MMA floor: the kernels were bound by epilogue instruction issue.  The epilogues here do the same
// arithmetic (reference: corss_attention.forward, mmdet3d/models/attention.py:192-219; LinearAttention :14-47;
// get_pooled_feats, ReIDNet.py:526-534) with less than half the instructions:
//   * LayerNorm means are removed algebraically: the producer weights (merge, mlp[2]) are centred over their output
//     channels on the host, so every LayerNorm input already has zero mean and only sum(x^2) is needed;
//   * LayerNorm1's affine is folded into the next GEMM (gamma scales the columns of W0b; W0b.beta arrives through one
//     extra K=16 step against a constant ones chunk in stage 2, and through the per-object term U in stage 1);
//   * the per-head normalisation 1/(Q.Ksum+eps) uses LayerNorm's scale invariance: LN(z0 D0 + z1 D1) =
//     LN'(D0 + (z1/z0) D1) with eps' = eps (dot0+eps_att)^2 -- one FMA per element instead of a multiply and an FMA;
//   * ReLU is fused into the fp32->bf16x2 conversion (F2FP.RELU) or the packed bias add (HFMA2.RELU);
//   * elu(x)+1 takes 4 packed instructions per two elements (projection weights pre-scaled by 1/bf16(ln 2));
//   * LayerNorm2's beta and the residual share one pre-added image in stage 1, and in stage 2 beta is added after the
//     pooling (max / mean commute with a per-channel constant): pcreid_pool_finish2;
//   * the max / sum pooling transposes through shared memory with 128-bit accesses (XOR-swizzled rows) and keeps the
//     running partials in registers across the tiles of a unit;
//   * accumulators are read with 32-column tcgen05.ld and one wait per batch.
#include "pair_common.cuh"
#include <stdlib.h>

namespace {

// weights blob of phase 1a (bytes): [W0b.diag(g1) | W0a | W0b.beta1 - W0a.beta2 | 0] image (N=128, K=144) | centred W2 image
// (N=64, K=128) | LN2 gamma (64 fp32).  The search object's own term W0a.h rides on the tensor core as four more K steps of G2
// against the (h + beta2) image that the residual needs anyway, instead of a per-object precomputed image U that every tile had
// to read (32 KB) and add in the epilogue (the constant W0a.beta2 this adds is taken back out through the bias K step).
constexpr int Q1A_W0 = 0, Q1A_W2 = 18 * 2048, Q1A_LN = Q1A_W2 + 16384, Q1A_WBYTES = Q1A_LN + 256;   // 53504
constexpr int Q1A_ONES = Q1A_WBYTES;                                                         // 4 KB: A chunk pair, k = 0 is 1.0
constexpr int Q1A_QXA = 0, Q1A_H = IMG, Q1A_MK1 = 2 * IMG, Q1A_GBYTES = 2 * IMG + B7_BYTES;  // 43008 B per group
// weights blob of phase 2: Wq/bf16(ln2) image (N=64, K=64) | [W0a | W0b.diag(g1) | W0b.beta1 | 0] image (N=128, K=144) |
// centred W2 image (N=64, K=128) | LN2 gamma (64 fp32)
constexpr int Q2_WQ = 0, Q2_W0 = 8192, Q2_W2 = Q2_W0 + 18 * 2048, Q2_LN = Q2_W2 + 16384, Q2_WBYTES = Q2_LN + 256;   // 61696
constexpr int Q2_ONES = Q2_WBYTES;                                                           // 4 KB: A chunk pair, k = 0 is 1.0
constexpr int Q2_R1 = 0, Q2_B7 = 2 * IMG, Q2_GBYTES = Q2_B7 + B7_BYTES;                      // 43008 B per group

__device__ __forceinline__ float u2f(uint32_t u) { return __uint_as_float(u); }

// The three groups use 3 x 160 of the CTA's 512 tensor-memory columns; columns [480, 512) hold ONE constant A operand for all of
// them: element k = 0 of every row is 1.0 (the K = 16 step that adds the folded LayerNorm1 bias).  Written once by group 0
// (thread == lane); the caller's __syncthreads publishes it.
template <class F>
__device__ __forceinline__ void tmem_ones(const GroupX& g) {
  if (g.gid == 0) {
    uint32_t w[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) w[j] = 0u;
    w[0] = F::ONE_LO;
    tc::tmem_st32(g.tlane + 480, w);
    tc::tmem_st_wait();
  }
  tc::tc_fence_before();
}

// one thread: the attention GEMM of a tile -- head h's queries (A K-steps 2h, 2h + 1 of the query image) against head h's image
// of the template operand (32 k-rows, N = 80) -> accumulator columns [80 h, 80 h + 80)
__device__ __forceinline__ void issue_attn(uint32_t tmem_d, const Opnd& Q, const Opnd& T, uint32_t id80) {
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
      tc::umma_f16(tmem_d + NB7H * h, Q.desc + (uint64_t)((2 * h + ks) * Q.kstep), T.desc + (uint64_t)((B7_HEAD >> 4) * h + ks * T.kstep),
                   id80, ks > 0 ? 1u : 0u);
}

// Attention accumulator (columns [0,64) head 0 | 64 its Q.Ksum dot, [80,144) head 1 | 144 its dot) -> LayerNorm1-normalised
// merged message WITHOUT affine (folded into the next GEMM), packed to bf16 into this thread's row of an operand image.
// XT: X' goes to tensor-memory columns [128, 160) of this thread's lane (the TMEM-sourced A operand of the next GEMM's X' K-steps)
// instead of the shared-memory image: the attention accumulator's columns there (head 1's tail and dot) are read before they are
// overwritten, and a lane is only ever touched by its own thread.  Saves the 16 KB image store and the tensor core's 4 x 4 KB operand reads per
// tile on the shared-memory pipe, the most loaded unit of these kernels.
template <class F, bool XT = false>
__device__ __forceinline__ void epi_attn_norm(uint32_t tl, uint8_t* dst_row, float att_eps) {
  uint32_t d8[8], e8[8], a0[32], a1[32], b0[32], b1[32];
  tc::tmem_ld8(tl + 64, d8);
  tc::tmem_ld8(tl + 144, e8);
  tc::tmem_ld32(tl, a0);
  tc::tmem_ld32(tl + 80, a1);
  tc::tmem_ld_wait();
  tc::tmem_ld32(tl + 32, b0);                                   // in flight while the first half is processed
  tc::tmem_ld32(tl + 112, b1);
  const float d0 = u2f(d8[0]) + att_eps, d1 = u2f(e8[0]) + att_eps;
  const float r = __fdividef(d0, d1);
  float ss0 = 0.f, ss1 = 0.f, ss2 = 0.f, ss3 = 0.f;
  float m0[32], m1[32];
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    m0[j] = fmaf(r, u2f(a1[j]), u2f(a0[j]));
    m0[j + 1] = fmaf(r, u2f(a1[j + 1]), u2f(a0[j + 1]));
    ss0 = fmaf(m0[j], m0[j], ss0);
    ss1 = fmaf(m0[j + 1], m0[j + 1], ss1);
  }
  tc::tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    m1[j] = fmaf(r, u2f(b1[j]), u2f(b0[j]));
    m1[j + 1] = fmaf(r, u2f(b1[j + 1]), u2f(b0[j + 1]));
    ss2 = fmaf(m1[j], m1[j], ss2);
    ss3 = fmaf(m1[j + 1], m1[j + 1], ss3);
  }
  const float rstd = rsqrtf(((ss0 + ss1) + (ss2 + ss3)) * (1.f / 64.f) + LN_EPS * d0 * d0);
  if constexpr (XT) {
    uint32_t w[32];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      w[j] = F::pack(m0[2 * j] * rstd, m0[2 * j + 1] * rstd);
      w[16 + j] = F::pack(m1[2 * j] * rstd, m1[2 * j + 1] * rstd);
    }
    tc::tmem_st32(tl + 128, w);
    tc::tmem_st_wait();
  } else {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t w[4], v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        w[j] = F::pack(m0[c * 8 + 2 * j] * rstd, m0[c * 8 + 2 * j + 1] * rstd);
        v[j] = F::pack(m1[c * 8 + 2 * j] * rstd, m1[c * 8 + 2 * j + 1] * rstd);
      }
      *reinterpret_cast<uint4*>(dst_row + c * 2048) = make_uint4(w[0], w[1], w[2], w[3]);
      *reinterpret_cast<uint4*>(dst_row + (4 + c) * 2048) = make_uint4(v[0], v[1], v[2], v[3]);
    }
  }
}

// 64 accumulator columns at tl -> fp32 registers + sum of squares (the producer weights are centred: zero mean)
__device__ __forceinline__ float ld64_sumsq(uint32_t tl, uint32_t (&x0)[32], uint32_t (&x1)[32]) {
  tc::tmem_ld32(tl, x0);
  tc::tmem_ld32(tl + 32, x1);
  tc::tmem_ld_wait();
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    s0 = fmaf(u2f(x0[j]), u2f(x0[j]), s0);
    s1 = fmaf(u2f(x0[j + 1]), u2f(x0[j + 1]), s1);
    s2 = fmaf(u2f(x1[j]), u2f(x1[j]), s2);
    s3 = fmaf(u2f(x1[j + 1]), u2f(x1[j + 1]), s3);
  }
  return (s0 + s1) + (s2 + s3);
}

// ---------------------------------------------------------------------------------------------------------------
// phase 1a: G1 attention (Qf1_i x MK1_j) -> LN1 -> G2 = [X' | h + beta2 | 1] W0'^T, ReLU (TMEM-resident) -> G3 -> LN2 + (h + beta2) -> a
// ---------------------------------------------------------------------------------------------------------------
template <class F, bool XT>
__global__ void __launch_bounds__(NGX * GX, 1) pair_p1a2_kernel(const P1Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[NGX];
  __shared__ uint32_t tmem_base_s;
  uint8_t* Wsm = smem;
  const float4* g2v = reinterpret_cast<const float4*>(Wsm + Q1A_LN);
  if (threadIdx.x == 0) {
    for (int i = 0; i < NGX; ++i) tc::mbar_init(&bars[i], 1);
    tc::fence_mbar_init();
  }
  if (threadIdx.x < 32) { tc::tmem_alloc(&tmem_base_s, 512); tc::tmem_relinquish(); }
  copy_to_smem(Wsm, a.W, Q1A_WBYTES, threadIdx.x, NGX * GX);
  cp_async_commit();
  if (threadIdx.x < 256)   // constant A chunk pair of the bias K-step: element k = 0 of every row is 1.0
    reinterpret_cast<uint4*>(smem + Q1A_ONES)[threadIdx.x] = threadIdx.x < 128 ? make_uint4(F::ONE_LO, 0, 0, 0) : make_uint4(0, 0, 0, 0);
  cp_async_wait<0>();
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  GroupX g;
  groupx_setup(g, bars, tmem_base_s);
  const uint32_t t_ones = tc::uniform(tmem_base_s) + 480;        // XT: the bias K-step's constant A operand lives in the spare columns
  if constexpr (XT) { tmem_ones<F>(g); __syncthreads(); tc::tc_fence_after(); }
  uint8_t* G = smem + Q1A_ONES + 4096 + g.gid * Q1A_GBYTES;
  uint8_t* QXa = G + Q1A_QXA;
  uint8_t* Hs = G + Q1A_H;
  uint8_t* MK1 = G + Q1A_MK1;
  const uint32_t sQXa = tc::smem_u32(QXa), sW = tc::smem_u32(Wsm);
  const uint32_t id80 = tc::instr_desc(128, NB7H, F::FMT, tc::MAJOR_K, tc::MAJOR_MN);
  const uint32_t id128 = tc::instr_desc(128, 128, F::FMT, tc::MAJOR_K, tc::MAJOR_K);
  const uint32_t id64 = tc::instr_desc(128, 64, F::FMT, tc::MAJOR_K, tc::MAJOR_K);
  // QXa and Hs are adjacent: one K-major A operand of 8 K steps [X' | h + beta2]
  const Opnd oQXa = A_IMG(sQXa), oOnes = A_IMG(tc::smem_u32(smem + Q1A_ONES)), oMK1 = B7_IMG(tc::smem_u32(MK1)),
             oW0 = W_IMG(sW + Q1A_W0, 128), oW2 = W_IMG(sW + Q1A_W2, 64);
  const int row = g.t;
  uint8_t* xrow = QXa + row * 16;

  const int ngroups = gridDim.x * NGX, gg = blockIdx.x * NGX + g.gid;
  const int u0 = (int)((long long)a.n_units * gg / ngroups), u1 = (int)((long long)a.n_units * (gg + 1) / ngroups);
  int cur_templ = -1;
  bool prefetched = false;
  int so_next = u0 < u1 ? a.u_search[u0] : 0, te_next = u0 < u1 ? a.u_templ[u0] : 0, slot_next = u0 < u1 ? a.u_slot[u0] : 0;
  // attention GEMM of a tile (query image ti against the template operand te).  Called for tile n+1 as soon as tile n's
  // last accumulator has been read into registers, so that it runs behind tile n's LayerNorm2 epilogue.
  auto start_tile = [&](size_t ti_, int te_) {
    if (!prefetched) {                                                      // first tile of this group only
      copy_to_smem(QXa, a.QF1 + ti_ * IMG, IMG, g.t, GX);
      copy_to_smem(Hs, a.H + ti_ * IMG, IMG, g.t, GX);
    }
    prefetched = false;
    if (te_ != cur_templ) { copy_to_smem(MK1, a.MK1 + (size_t)te_ * B7_BYTES, B7_BYTES, g.t, GX); cur_templ = te_; }
    cp_async_commit();
    cp_async_wait<0>();
    g.publish();
    if (g.issuer) { if (tc::elect_one()) { issue_attn(g.tmem, oQXa, oMK1, id80); tc::umma_commit(g.bar); } __syncwarp(); }
  };
  for (int u = u0; u < u1; ++u) {
    const int so = so_next, te = te_next, slot = slot_next;
    if (u + 1 < u1) { so_next = a.u_search[u + 1]; te_next = a.u_templ[u + 1]; slot_next = a.u_slot[u + 1]; }
    for (int tile = 0; tile < a.NT; ++tile) {
      const size_t ti = (size_t)so * a.NT + tile;
      if (u == u0 && tile == 0) start_tile(ti, te);                       // every later tile was started by its predecessor
      g.wait();
      epi_attn_norm<F, XT>(g.tlane, xrow, a.att_eps);                                   // X' over the query image / into TMEM
      g.publish();
      if (g.issuer) {
        if (tc::elect_one()) {
          if constexpr (XT) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)                                                           // X' K-steps: A from tensor memory
              tc::umma_f16_ts(g.tmem, g.tmem + 128 + 8 * ks, oW0.desc + (uint64_t)(ks * oW0.kstep), id128, ks > 0 ? 1u : 0u);
#pragma unroll
            for (int ks = 4; ks < 8; ++ks)                                                           // h + beta2 K-steps: shared memory
              tc::umma_f16(g.tmem, oQXa.desc + (uint64_t)(ks * oQXa.kstep), oW0.desc + (uint64_t)(ks * oW0.kstep), id128, 1u);
          } else {
            issue_gemm<8>(g.tmem, oQXa, oW0, id128, false);                                          // [X' | h + beta2]
          }
          if constexpr (XT) tc::umma_f16_ts(g.tmem, t_ones, oW0.desc + (uint64_t)(8 * oW0.kstep), id128, 1u);
          else tc::umma_f16(g.tmem, oOnes.desc, oW0.desc + (uint64_t)(8 * oW0.kstep), id128, 1u);    // + W0b.beta1 - W0a.beta2
          tc::umma_commit(g.bar);
        }
        __syncwarp();
      }
      g.wait();
      {   // G2 has consumed [X' | h]: the images of the next (unit, tile) stream into QXa / Hs behind the rest of this tile
        int nu = u, nt = tile + 1;
        if (nt == a.NT) { nu = u + 1; nt = 0; }
        if (nu < u1) {
          const size_t tn = (size_t)(nt == 0 ? so_next : so) * a.NT + nt;
          copy_to_smem(QXa, a.QF1 + tn * IMG, IMG, g.t, GX);
          copy_to_smem(Hs, a.H + tn * IMG, IMG, g.t, GX);
          cp_async_commit();
          prefetched = true;
        }
      }
#pragma unroll
      for (int b = 0; b < 2; ++b) {   // Hd = relu(acc) -> 16 bit, written back IN PLACE to TMEM columns [0, 64): the A operand of G3
        uint32_t r0[32], r1[32], w[32];
        tc::tmem_ld32(g.tlane + 64 * b, r0);
        tc::tmem_ld32(g.tlane + 64 * b + 32, r1);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          w[j] = F::pack_relu(u2f(r0[2 * j]), u2f(r0[2 * j + 1]));
          w[16 + j] = F::pack_relu(u2f(r1[2 * j]), u2f(r1[2 * j + 1]));
        }
        tc::tmem_st32(g.tlane + 32 * b, w);
      }
      tc::tmem_st_wait();
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
      {   // a = (h + beta2) + gamma2 * acc * rstd -> global 16-bit image (stage-1 output)
        uint4 sdH[8];
        load_side<8>(sdH, a.H + ti * IMG, 0, row);
        g.wait();
        uint32_t x0[32], x1[32];
        const float rstd = rsqrtf(ld64_sumsq(g.tlane + 64, x0, x1) * (1.f / 64.f) + LN_EPS);
        if (tile + 1 < a.NT) start_tile(ti + 1, te);                      // the accumulator now lives in registers:
        else if (u + 1 < u1) start_tile((size_t)so_next * a.NT, te_next);  // the next G1 may overwrite the columns
        uint8_t* orow = a.A_out + (((size_t)slot * 2 + a.role) * a.NT + tile) * IMG + row * 16;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t (&x)[32] = hh == 0 ? x0 : x1;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 rs = sdH[4 * hh + c];
            const float4 ga = g2v[8 * hh + 2 * c], gb = g2v[8 * hh + 2 * c + 1];
            uint32_t w[4];
            w[0] = F::pack(fmaf(u2f(x[c * 8 + 0]) * rstd, ga.x, F::lo(rs.x)), fmaf(u2f(x[c * 8 + 1]) * rstd, ga.y, F::hi(rs.x)));
            w[1] = F::pack(fmaf(u2f(x[c * 8 + 2]) * rstd, ga.z, F::lo(rs.y)), fmaf(u2f(x[c * 8 + 3]) * rstd, ga.w, F::hi(rs.y)));
            w[2] = F::pack(fmaf(u2f(x[c * 8 + 4]) * rstd, gb.x, F::lo(rs.z)), fmaf(u2f(x[c * 8 + 5]) * rstd, gb.y, F::hi(rs.z)));
            w[3] = F::pack(fmaf(u2f(x[c * 8 + 6]) * rstd, gb.z, F::lo(rs.w)), fmaf(u2f(x[c * 8 + 7]) * rstd, gb.w, F::hi(rs.w)));
            *reinterpret_cast<uint4*>(orow + (4 * hh + c) * 2048) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
      tc::tc_fence_before();                                              // TMEM reads done before the next tile's G1 overwrites
    }
  }
  cp_async_wait<0>();
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(tmem_base_s, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// phase 2: q projection -> attention against B7 -> LN1 -> [a | X' | 1] W0'^T, ReLU (TMEM-resident) -> W2 -> LN2 + a -> pooling
// ---------------------------------------------------------------------------------------------------------------
template <class F, bool XT>
__global__ void __launch_bounds__(NGX * GX, 1) pair_p2y_kernel(const P2Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[NGX];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float comb[NGX][4][2][64];
  uint8_t* Wsm = smem;
  const float4* g2v = reinterpret_cast<const float4*>(Wsm + Q2_LN);
  if (threadIdx.x == 0) {
    for (int i = 0; i < NGX; ++i) tc::mbar_init(&bars[i], 1);
    tc::fence_mbar_init();
  }
  if (threadIdx.x < 32) { tc::tmem_alloc(&tmem_base_s, 512); tc::tmem_relinquish(); }
  copy_to_smem(Wsm, a.W, Q2_WBYTES, threadIdx.x, NGX * GX);
  cp_async_commit();
  if (threadIdx.x < 256)   // constant A chunk pair of the bias K-step: element k = 0 of every row is 1.0
    reinterpret_cast<uint4*>(smem + Q2_ONES)[threadIdx.x] = threadIdx.x < 128 ? make_uint4(F::ONE_LO, 0, 0, 0) : make_uint4(0, 0, 0, 0);
  cp_async_wait<0>();
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  GroupX g;
  groupx_setup(g, bars, tmem_base_s);
  const uint32_t t_ones = tc::uniform(tmem_base_s) + 480;        // XT: the bias K-step's constant A operand lives in the spare columns
  if constexpr (XT) { tmem_ones<F>(g); __syncthreads(); tc::tc_fence_after(); }
  const int warp_in_group = (int)tc::uniform(threadIdx.x >> 5) % 4;
  uint8_t* G = smem + Q2_ONES + 4096 + g.gid * Q2_GBYTES;
  uint8_t* R1 = G + Q2_R1;
  uint8_t* B7 = G + Q2_B7;
  const uint32_t sR1 = tc::smem_u32(R1), sW = tc::smem_u32(Wsm);
  const uint32_t id80 = tc::instr_desc(128, NB7H, F::FMT, tc::MAJOR_K, tc::MAJOR_MN);
  const uint32_t id128 = tc::instr_desc(128, 128, F::FMT, tc::MAJOR_K, tc::MAJOR_K);
  const uint32_t id64 = tc::instr_desc(128, 64, F::FMT, tc::MAJOR_K, tc::MAJOR_K);
  const Opnd oR1 = A_IMG(sR1), oQf = A_IMG(sR1 + IMG), oOnes = A_IMG(tc::smem_u32(smem + Q2_ONES)), oWq = W_IMG(sW + Q2_WQ, 64),
             oW0 = W_IMG(sW + Q2_W0, 128), oW2 = W_IMG(sW + Q2_W2, 64), oB7 = B7_IMG(tc::smem_u32(B7));
  const int row = g.t;
  uint8_t* arow = R1 + row * 16;            // this thread's row inside the operand images
  float4* Tb = reinterpret_cast<float4*>(R1 + IMG);     // 16 KB transpose buffer: [128 rows][8 x float4], XOR-swizzled
  const int jg = g.t & 7, rsub = g.t >> 3;  // pooling: channel quad, row block of 8

  const int ngroups = gridDim.x * NGX, gg = blockIdx.x * NGX + g.gid;
  const int u0 = (int)((long long)a.n_units * gg / ngroups), u1 = (int)((long long)a.n_units * (gg + 1) / ngroups);
  int slot_next = u0 < u1 ? a.u_slot[u0] : 0;
  if (u0 < u1) {   // first tile of this group: nothing to hide the loads behind
    copy_to_smem(R1, a.A_in + ((size_t)slot_next * 2 + a.role) * a.NT * IMG, IMG, g.t, GX);
    copy_to_smem(B7, a.B7_in + ((size_t)slot_next * 2 + (1 - a.role)) * B7_BYTES, B7_BYTES, g.t, GX);
  }
  cp_async_commit();
  // q projection of a tile: its `a` image (and, at a unit start, B7) was prefetched by cp.async.  Called for tile n+1
  // right after tile n's last GEMM has completed, so that the projection runs behind tile n's LayerNorm2 + pooling
  // epilogue instead of in front of an idle group (it writes TMEM columns [0, 64): the dead hidden layer of tile n).
  auto start_tile = [&]() {
    cp_async_wait<0>();
    g.publish();
    if (g.issuer) { if (tc::elect_one()) { issue_gemm<4>(g.tmem, oR1, oWq, id64, false); tc::umma_commit(g.bar); } __syncwarp(); }
  };
  for (int u = u0; u < u1; ++u) {
    const int slot = slot_next;
    if (u + 1 < u1) slot_next = a.u_slot[u + 1];
    float4 pmx[2], psm[2];                                                // channels 32 hh + 4 jg .. + 3, rows of this thread's blocks
    pmx[0] = pmx[1] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    psm[0] = psm[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int tile = 0; tile < a.NT; ++tile) {
      const uint8_t* a_img = a.A_in + (((size_t)slot * 2 + a.role) * a.NT + tile) * IMG;
      if (u == u0 && tile == 0) start_tile();                             // every later tile was started by its predecessor
      else g.sync();              // the Qf image below overwrites the transpose buffer: every warp has finished its pooling reads
      {   // L2 prefetch of the next tile's `a` image (its shared-memory copy is issued after this tile's G8)
        int ns = slot, nt = tile + 1;
        if (nt == a.NT) { ns = slot_next; nt = 0; }
        if (nt != 0 || u + 1 < u1) prefetch_l2_16k(a.A_in + (((size_t)ns * 2 + a.role) * a.NT + nt) * IMG, g.t);
        if (tile == 0 && u + 1 < u1 && g.t < B7_BYTES / 128) prefetch_l2_16k(a.B7_in + ((size_t)slot_next * 2 + (1 - a.role)) * B7_BYTES, g.t);
      }
      g.wait();
      {   // Qf = elu(q)+1 -> second half of R1
        uint32_t r0[32], r1[32];
        tc::tmem_ld32(g.tlane, r0);
        tc::tmem_ld32(g.tlane + 32, r1);
        tc::tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t w[4], v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            w[j] = F::elu1_scaled(u2f(r0[c * 8 + 2 * j]), u2f(r0[c * 8 + 2 * j + 1]));
            v[j] = F::elu1_scaled(u2f(r1[c * 8 + 2 * j]), u2f(r1[c * 8 + 2 * j + 1]));
          }
          *reinterpret_cast<uint4*>(arow + IMG + c * 2048) = make_uint4(w[0], w[1], w[2], w[3]);
          *reinterpret_cast<uint4*>(arow + IMG + (4 + c) * 2048) = make_uint4(v[0], v[1], v[2], v[3]);
        }
      }
      g.publish();
      if (g.issuer) { if (tc::elect_one()) { issue_attn(g.tmem, oQf, oB7, id80); tc::umma_commit(g.bar); } __syncwarp(); }
      g.wait();
      if (tile + 1 == a.NT && u + 1 < u1) {   // last attention GEMM of the unit is done: next unit's B7 streams in
        copy_to_smem(B7, a.B7_in + ((size_t)slot_next * 2 + (1 - a.role)) * B7_BYTES, B7_BYTES, g.t, GX);
        cp_async_commit();
      }
      epi_attn_norm<F, XT>(g.tlane, arow + IMG, a.att_eps);                             // X' (XT: in tensor memory): [a | X' | 1] is the K = 144 operand
      g.publish();
      if (g.issuer) {
        if (tc::elect_one()) {
          if constexpr (XT) {
            issue_gemm<4>(g.tmem, oR1, oW0, id128, false);                                          // a K-steps: shared memory
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)                                                          // X' K-steps: A from tensor memory
              tc::umma_f16_ts(g.tmem, g.tmem + 128 + 8 * ks, oW0.desc + (uint64_t)((4 + ks) * oW0.kstep), id128, 1u);
          } else {
            issue_gemm<8>(g.tmem, oR1, oW0, id128, false);
          }
          if constexpr (XT) tc::umma_f16_ts(g.tmem, t_ones, oW0.desc + (uint64_t)(8 * oW0.kstep), id128, 1u);
          else tc::umma_f16(g.tmem, oOnes.desc, oW0.desc + (uint64_t)(8 * oW0.kstep), id128, 1u);   // + W0b.beta1
          tc::umma_commit(g.bar);
        }
        __syncwarp();
      }
      uint4 sdA[8];
      load_side<8>(sdA, a_img, 0, row);                                   // residual a (this tile), consumed after G9
      g.wait();
      {   // G8 has consumed [a | X']: the next tile's `a` image streams into R1 behind the rest of this tile
        int nu = u, nt = tile + 1;
        if (nt == a.NT) { nu = u + 1; nt = 0; }
        if (nu < u1) {
          copy_to_smem(R1, a.A_in + (((size_t)(nt == 0 ? slot_next : slot) * 2 + a.role) * a.NT + nt) * IMG, IMG, g.t, GX);
          cp_async_commit();
        }
      }
#pragma unroll
      for (int b = 0; b < 2; ++b) {   // Hd = relu(acc) -> bf16 in place in TMEM columns [0, 64): the A operand of G9
        uint32_t r0[32], r1[32], w[32];
        tc::tmem_ld32(g.tlane + 64 * b, r0);
        tc::tmem_ld32(g.tlane + 64 * b + 32, r1);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          w[j] = F::pack_relu(u2f(r0[2 * j]), u2f(r0[2 * j + 1]));
          w[16 + j] = F::pack_relu(u2f(r1[2 * j]), u2f(r1[2 * j + 1]));
        }
        tc::tmem_st32(g.tlane + 32 * b, w);
      }
      tc::tmem_st_wait();
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
      {   // o - beta2 = a + gamma2 * acc * rstd ; pooled over the points, 32 channels per pass
        uint32_t x0[32], x1[32];
        const float rstd = rsqrtf(ld64_sumsq(g.tlane + 64, x0, x1) * (1.f / 64.f) + LN_EPS);
        if (tile + 1 < a.NT || u + 1 < u1) start_tile();                  // the accumulator now lives in registers
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t (&x)[32] = hh == 0 ? x0 : x1;
          if (hh == 1) __syncwarp();                                      // pass-0 reads are done before pass 1 overwrites (a warp only
                                                                          // ever touches the 32 transpose-buffer rows it owns)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 rs = sdA[4 * hh + c];
            const float4 ga = g2v[8 * hh + 2 * c], gb = g2v[8 * hh + 2 * c + 1];
            float4 oa, ob;
            oa.x = fmaf(u2f(x[c * 8 + 0]) * rstd, ga.x, F::lo(rs.x));
            oa.y = fmaf(u2f(x[c * 8 + 1]) * rstd, ga.y, F::hi(rs.x));
            oa.z = fmaf(u2f(x[c * 8 + 2]) * rstd, ga.z, F::lo(rs.y));
            oa.w = fmaf(u2f(x[c * 8 + 3]) * rstd, ga.w, F::hi(rs.y));
            ob.x = fmaf(u2f(x[c * 8 + 4]) * rstd, gb.x, F::lo(rs.z));
            ob.y = fmaf(u2f(x[c * 8 + 5]) * rstd, gb.y, F::hi(rs.z));
            ob.z = fmaf(u2f(x[c * 8 + 6]) * rstd, gb.z, F::lo(rs.w));
            ob.w = fmaf(u2f(x[c * 8 + 7]) * rstd, gb.w, F::hi(rs.w));
            Tb[row * 8 + ((2 * c) ^ (row & 7))] = oa;
            Tb[row * 8 + ((2 * c + 1) ^ (row & 7))] = ob;
          }
          __syncwarp();
          if ((tile + 1) * 128 <= a.npts) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 v = Tb[(rsub * 8 + i) * 8 + (jg ^ i)];
              pmx[hh].x = fmaxf(pmx[hh].x, v.x); pmx[hh].y = fmaxf(pmx[hh].y, v.y);
              pmx[hh].z = fmaxf(pmx[hh].z, v.z); pmx[hh].w = fmaxf(pmx[hh].w, v.w);
              psm[hh].x += v.x; psm[hh].y += v.y; psm[hh].z += v.z; psm[hh].w += v.w;
            }
          } else {   // ragged last tile: rows beyond the object's point count are padding and stay out of the pooling
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (tile * 128 + rsub * 8 + i < a.npts) {
                const float4 v = Tb[(rsub * 8 + i) * 8 + (jg ^ i)];
                pmx[hh].x = fmaxf(pmx[hh].x, v.x); pmx[hh].y = fmaxf(pmx[hh].y, v.y);
                pmx[hh].z = fmaxf(pmx[hh].z, v.z); pmx[hh].w = fmaxf(pmx[hh].w, v.w);
                psm[hh].x += v.x; psm[hh].y += v.y; psm[hh].z += v.z; psm[hh].w += v.w;
              }
            }
          }
        }
      }
      tc::tc_fence_before();      // the next tile's publish() orders these TMEM / transpose-buffer reads before its writes
    }
    // unit done: combine the row blocks (lanes with equal jg inside a warp, then the 4 warps through shared memory)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
      for (int o = 8; o <= 16; o <<= 1) {
        pmx[hh].x = fmaxf(pmx[hh].x, __shfl_xor_sync(FULL_MASK, pmx[hh].x, o));
        pmx[hh].y = fmaxf(pmx[hh].y, __shfl_xor_sync(FULL_MASK, pmx[hh].y, o));
        pmx[hh].z = fmaxf(pmx[hh].z, __shfl_xor_sync(FULL_MASK, pmx[hh].z, o));
        pmx[hh].w = fmaxf(pmx[hh].w, __shfl_xor_sync(FULL_MASK, pmx[hh].w, o));
        psm[hh].x += __shfl_xor_sync(FULL_MASK, psm[hh].x, o);
        psm[hh].y += __shfl_xor_sync(FULL_MASK, psm[hh].y, o);
        psm[hh].z += __shfl_xor_sync(FULL_MASK, psm[hh].z, o);
        psm[hh].w += __shfl_xor_sync(FULL_MASK, psm[hh].w, o);
      }
      if ((g.t & 31) < 8) {
        *reinterpret_cast<float4*>(&comb[g.gid][warp_in_group][0][32 * hh + 4 * jg]) = pmx[hh];
        *reinterpret_cast<float4*>(&comb[g.gid][warp_in_group][1][32 * hh + 4 * jg]) = psm[hh];
      }
    }
    g.sync();
    if (g.t < 64) {
      float* out = a.pool_part + ((size_t)slot * 2 + a.role) * 128;
      out[g.t] = fmaxf(fmaxf(comb[g.gid][0][0][g.t], comb[g.gid][1][0][g.t]), fmaxf(comb[g.gid][2][0][g.t], comb[g.gid][3][0][g.t]));
      out[64 + g.t] = (comb[g.gid][0][1][g.t] + comb[g.gid][1][1][g.t]) + (comb[g.gid][2][1][g.t] + comb[g.gid][3][1][g.t]);
    }
  }
  cp_async_wait<0>();
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(tmem_base_s, 512);
}

// pool_part (P, 2, 128) [max | sum of (o - beta2)] -> pooled^T (128, P): max over both directions | mean over the 2*npts points
// One CTA = 32 pairs x 128 channels through a padded shared-memory tile: rows of `part` are read coalesced (a warp reads one
// pair's 128 floats per direction), columns of the (128, P) output are written coalesced (32 consecutive pairs per channel).
__global__ void __launch_bounds__(256) pool_finish2_kernel(int P, int npts, const float* __restrict__ part, const float* __restrict__ bias,
                                                           float* __restrict__ out) {
  __shared__ float tile[32][129];
  const int p0 = blockIdx.x * 32, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < 32; r += 8) {
    const int p = p0 + r;
    if (p >= P) break;
    const float* row = part + (size_t)p * 256;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = lane + 32 * j;
      const float x0 = row[c], x1 = row[128 + c];
      const float b = bias ? bias[c & 63] : 0.f;
      tile[r][c] = (c < 64 ? fmaxf(x0, x1) : (x0 + x1) / (float)(2 * npts)) + b;
    }
  }
  __syncthreads();
  if (p0 + lane < P)
    for (int c = warp; c < 128; c += 8) out[(size_t)c * P + p0 + lane] = tile[lane][c];
}

// src (B, C, N) channel-major fp32 + per-channel bias -> dst [B][N/128][C/8][128][8] bf16
template <class F>
__global__ void __launch_bounds__(256) pack_image_bias_kernel(int B, int C, int N, const float* __restrict__ src, long long s_bs,
                                                              int lds, const float* __restrict__ bias, uint8_t* __restrict__ dst) {
  const int nt = (N + 127) / 128, nch = C / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int row = (int)(idx % 128);
  const int chunk = (int)((idx / 128) % nch);
  const int tile = (int)((idx / (128LL * nch)) % nt);
  const long long b = idx / (128LL * nch * nt);
  if (b >= B) return;
  const float* s = src + b * s_bs + (size_t)(chunk * 8) * lds + tile * 128 + row;
  uint32_t w[4] = {0u, 0u, 0u, 0u};                   // zero padding beyond N
  if (tile * 128 + row < N) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      w[j] = F::pack(s[(size_t)(2 * j) * lds] + bias[chunk * 8 + 2 * j], s[(size_t)(2 * j + 1) * lds] + bias[chunk * 8 + 2 * j + 1]);
  }
  *reinterpret_cast<uint4*>(dst + (((size_t)b * nt + tile) * nch + chunk) * 2048 + row * 16) = make_uint4(w[0], w[1], w[2], w[3]);
}

}  // namespace

// The LayerNorm1 output X' of phase 1a (bit 0) / phase 2 (bit 1) goes to tensor memory instead of its shared-memory operand image:
// bit-identical results, phase 1a -8 %, phase 2 -2 % (profiles/r02_x_tmem_ab.json).  A/B: PCREID_X_TMEM=0 restores the images.
// (Tried on top and dropped: phase 2's queries in tensor memory as well, with the Q.Ksum dots as 64 SIMT FMAs per row -- the
// extra issue slots cost more than the 32 KB per tile saved on the shared-memory pipe: +1-2 % on phase 2.)
static int x_tmem_mask() {
  static const int m = [] { const char* e = getenv("PCREID_X_TMEM"); return e ? atoi(e) : 3; }();
  return m;
}

template <class F, bool XT>
static int launch_p1a2_x(const P1Args& a, int grid, cudaStream_t st) {
  const int smem = Q1A_ONES + 4096 + NGX * Q1A_GBYTES;
  cudaFuncSetAttribute(pair_p1a2_kernel<F, XT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  pair_p1a2_kernel<F, XT><<<grid, NGX * GX, smem, st>>>(a);
  return pcreid_launch_status();
}
// Tried and dropped (profiles/r02_unit_order_ab.json): a phase-1a loop that keeps the SEARCH tile resident and streams the 10 KB
// template operands past it through a double buffer (37 -> 10 KB of cp.async fill per tile): 8 % slower than this per-unit loop,
// whose units run in order of equal TEMPLATE -- the fills are not what loads the shared-memory pipe.
template <class F>
static int launch_p1a2(const P1Args& a, int grid, cudaStream_t st) {
  return (x_tmem_mask() & 1) ? launch_p1a2_x<F, true>(a, grid, st) : launch_p1a2_x<F, false>(a, grid, st);
}

template <class F, bool XT>
static int launch_p2y_x(const P2Args& a, int grid, cudaStream_t st) {
  const int smem = Q2_ONES + 4096 + NGX * Q2_GBYTES;
  cudaFuncSetAttribute(pair_p2y_kernel<F, XT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  pair_p2y_kernel<F, XT><<<grid, NGX * GX, smem, st>>>(a);
  return pcreid_launch_status();
}
template <class F>
static int launch_p2y(const P2Args& a, int grid, cudaStream_t st) {
  return (x_tmem_mask() & 2) ? launch_p2y_x<F, true>(a, grid, st) : launch_p2y_x<F, false>(a, grid, st);
}

extern "C" {

int pcreid_pack_image_bias(int B, int C, int N, const float* src, long long s_bs, int lds, const float* bias, int fmt, void* dst,
                           void* stream) {
  if (B <= 0) return PCREID_OK;
  if (!src || !dst || !bias || C % 8 || N <= 0 || (fmt != PCREID_FMT_BF16 && fmt != PCREID_FMT_F16)) return PCREID_ERR_ARG;
  const long long per = 128LL * (C / 8) * ((N + 127) / 128);
  const unsigned grid = (unsigned)((per * B + 255) / 256);
  if (fmt == PCREID_FMT_F16)
    pack_image_bias_kernel<tc::OpF16><<<grid, 256, 0, (cudaStream_t)stream>>>(B, C, N, src, s_bs, lds, bias, (uint8_t*)dst);
  else
    pack_image_bias_kernel<tc::OpBF16><<<grid, 256, 0, (cudaStream_t)stream>>>(B, C, N, src, s_bs, lds, bias, (uint8_t*)dst);
  return pcreid_launch_status();
}

int pcreid_pool_finish2(int P, int npts, const float* part, const float* bias, float* out, void* stream) {
  if (P <= 0) return PCREID_OK;
  if (!part || !out) return PCREID_ERR_ARG;
  pool_finish2_kernel<<<(unsigned)((P + 31) / 32), 256, 0, (cudaStream_t)stream>>>(P, npts, part, bias, out);
  return pcreid_launch_status();
}

int pcreid_pair_p1a2(int n_units, int npts, int role, int fmt, float att_eps, const int* u_search, const int* u_templ, const int* u_slot,
                     const void* QF1, const void* H, const void* MK1, const void* W, void* A_out, int n_ctas, void* stream) {
  if (n_units <= 0) return PCREID_OK;
  if (!u_search || !u_templ || !u_slot || !QF1 || !H || !MK1 || !W || !A_out || npts <= 0 ||
      (fmt != PCREID_FMT_BF16 && fmt != PCREID_FMT_F16))
    return PCREID_ERR_ARG;
  const int NT = (npts + 127) / 128;
  P1Args a{n_units, NT, role, npts, att_eps, 1.f, u_search, u_templ, u_slot, (const uint8_t*)QF1, (const uint8_t*)H,
           nullptr, (const uint8_t*)MK1, (const uint8_t*)W, (uint8_t*)A_out, nullptr};
  int grid = n_ctas > 0 ? n_ctas : 148;
  if (grid * NGX > n_units) grid = (n_units + NGX - 1) / NGX;
  return fmt == PCREID_FMT_F16 ? launch_p1a2<tc::OpF16>(a, grid, (cudaStream_t)stream) : launch_p1a2<tc::OpBF16>(a, grid, (cudaStream_t)stream);
}

int pcreid_pair_p2y(int n_units, int npts, int role, int fmt, float att_eps, const int* u_slot, const void* A_in, const void* B7_in,
                    const void* W, float* pool_part, int n_ctas, void* stream) {
  if (n_units <= 0) return PCREID_OK;
  if (!u_slot || !A_in || !B7_in || !W || !pool_part || npts <= 0 || (fmt != PCREID_FMT_BF16 && fmt != PCREID_FMT_F16)) return PCREID_ERR_ARG;
  const int NT = (npts + 127) / 128;
  P2Args a{n_units, NT, role, npts, att_eps, u_slot, (const uint8_t*)A_in, (const uint8_t*)B7_in, (const uint8_t*)W, pool_part};
  int grid = n_ctas > 0 ? n_ctas : 148;
  if (grid * NGX > n_units) grid = (n_units + NGX - 1) / NGX;
  return fmt == PCREID_FMT_F16 ? launch_p2y<tc::OpF16>(a, grid, (cudaStream_t)stream) : launch_p2y<tc::OpBF16>(a, grid, (cudaStream_t)stream);
}

}  // extern "C"
