// Fused all-pairs `xcorr_eff` match on the 5th-gen tensor cores (tcgen05 + TMEM), 16-bit operands (bf16 in "fast" mode,
// fp16 in "parity_tc" mode: tc_common.cuh OpBF16 / OpF16), fp32 accumulation, fp32 LayerNorm / linear-attention
// normalisation.  This file: phase 1b (key/value summaries of the stage-1 outputs) and the per-object packing kernels;
// phase 1a and phase 2 live in pair_tc2.cu.
//
// Reference arithmetic: ReIDNet.xcorr_eff (mmdet3d/models/ReIDNet.py:231-247) = 2 x corss_attention each way
// (mmdet3d/models/attention.py:192-219) + get_pooled_feats (ReIDNet.py:526-534).  The reference gathers
// feat[pairs[:,0]] / feat[pairs[:,1]] and runs ~40 torch ops per pair batch; here, per (pair, direction) unit and
// 128-point tile of the search object:
//
//   phase 1a (pair_p1a2_kernel):
//     G1  Qf1_i x MK1_j            -> per-head Q.KV.merge and the two Q.Ksum dots in ONE N=144 GEMM
//         epilogue: head merge + LayerNorm1 -> X (smem operand image)
//     G2  X x W0b^T (+ U_i = W0a h_i, precomputed per object), ReLU -> Hd (TMEM-resident A operand)
//     G3  Hd x W2^T, LayerNorm2, + h_i  -> a   (stage-1 output, spilled once as a 16-bit operand image)
//   phase 1b (pair_p1b_kernel, this file):
//     G4  a x [Wk2;Wv2]^T -> Kf = elu+1, V (+ Wv2 pos_j)                -> operand image
//     G5  [Kf|V]^T x [V|1]  accumulated over the tiles in TMEM          -> KV (64x64) and Ksum of the template
//     G6  blockdiag(KV) x Wm2^T                                         -> B7 = stage-2 attention operand (global)
//   phase 2 (pair_p2y_kernel):
//     G4' a x Wq2^T -> Qf2 ; G7 Qf2 x B7 (as G1) ; G8 [a|X|1] x W0'^T, ReLU ; G9 Hd x W2^T, LayerNorm2, + a
//     epilogue: per-channel max / sum over the points (smem transpose), accumulated over the tiles
//   pool_finish2: combine the two directions -> pooled (128) -> match head.
//
// Every GEMM is a tcgen05.mma (M=128, K=16 per instruction) issued by one thread per 128-thread group, operands in
// shared memory in the no-swizzle canonical layouts, accumulators in TMEM, epilogues by the row-owning threads via
// tcgen05.ld.  A CTA holds three independent groups that share one copy of the weights, so the tensor pipe of one group
// overlaps the epilogues of the other two.
#include "pair_common.cuh"
#include <stdlib.h>

namespace {

constexpr int P1B_WKV = 0, P1B_WM = 16384, P1B_WBYTES = 24576;
constexpr int P1B_AIMG = 0, P1B_KFV = IMG, P1B_GBYTES = IMG + 2 * IMG + ONES_BYTES;          // 53248 B per group

// CS = column split: CS threads share a tile row (= TMEM lane), each on 64 / CS of the accumulator columns, so a group is
// 4 CS warps (warps w and w + 4 sit on the same TMEM lane quadrant).  The epilogues of this kernel are purely elementwise
// (no row statistics), so the split needs no exchange between the threads of a row.
template <class F, int CS>
__global__ void __launch_bounds__(NGX * GX * CS, 1) pair_p1b_kernel(const P1Args a) {
  constexpr int GT = GX * CS, NTHR = NGX * GT;                           // threads per group / per CTA
  constexpr int QB = 4 / CS, CB = 8 / CS;                                // 16-column batches / 8-column chunks per thread
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[2 * NGX];
  __shared__ uint32_t tmem_base_s;
  uint8_t* Wsm = smem;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * NGX; ++i) tc::mbar_init(&bars[i], 1);
    tc::fence_mbar_init();
  }
  if (threadIdx.x < 32) { tc::tmem_alloc(&tmem_base_s, 512); tc::tmem_relinquish(); }
  copy_to_smem(Wsm, a.W, P1B_WBYTES, threadIdx.x, NTHR);
  cp_async_commit();
  cp_async_wait<0>();
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const int warp_u = (int)tc::uniform(threadIdx.x >> 5);
  const int gid = warp_u / (4 * CS), wg = warp_u % (4 * CS), gt = threadIdx.x % GT;
  const int half = wg / 4;                                               // which part of the columns this thread owns
  const int row = (wg % 4) * 32 + (threadIdx.x & 31);                    // tile row == TMEM lane
  const bool issuer = wg == 0;
  const uint32_t tmem = tc::uniform(tmem_base_s) + gid * 160;
  const uint32_t tlane = tmem + ((uint32_t)((wg % 4) * 32) << 16);
  uint64_t* bar = bars + gid;
  uint32_t par = 0;
  auto gsync = [&]() { tc::bar_sync(1 + gid, GT); };
  auto publish = [&]() { tc::fence_async_smem(); tc::tc_fence_before(); gsync(); tc::tc_fence_after(); };
  auto gwait = [&]() { tc::mbar_wait(bar, par); par ^= 1u; tc::tc_fence_after(); };
  uint64_t* bar2 = bars + NGX + gid;
  uint32_t par2 = 0;
  uint8_t* G = smem + P1B_WBYTES + gid * P1B_GBYTES;
  uint8_t* Aimg = G + P1B_AIMG;
  uint8_t* KfV = G + P1B_KFV;
  {
    uint4* ones = reinterpret_cast<uint4*>(KfV + 2 * IMG);
    for (int i = gt; i < 256; i += GT) ones[i] = i < 128 ? make_uint4(F::ONE_LO, 0, 0, 0) : make_uint4(0, 0, 0, 0);
  }
  const uint32_t sA = tc::smem_u32(Aimg), sKfV = tc::smem_u32(KfV), sW = tc::smem_u32(Wsm);
  const uint32_t id64 = tc::instr_desc(128, 64, F::FMT, tc::MAJOR_K, tc::MAJOR_K);
  const uint32_t idkv = tc::instr_desc(128, 80, F::FMT, tc::MAJOR_MN, tc::MAJOR_MN);
  // Wkv image is [k/8][128 rows: Wk 0..63 | Wv 64..127][8]: an N=64 operand is the same image entered at row 0 / row 64
  const Opnd oA = A_IMG(sA), oWk = W_IMG(sW + P1B_WKV, 128), oWv = W_IMG(sW + P1B_WKV + 64 * 16, 128), oWm = W_IMG(sW + P1B_WM, 64),
             oKfV = opnd(sKfV, 128u, 2048u, 256u), oVones = opnd(sKfV + 8 * 2048, 128u, 2048u, 256u);
  const uint32_t KVC = 64;                                                // TMEM columns [64, 144): KV / Ksum accumulator

  const int ngroups = gridDim.x * NGX, gg = blockIdx.x * NGX + gid;
  const int u0 = (int)((long long)a.n_units * gg / ngroups), u1 = (int)((long long)a.n_units * (gg + 1) / ngroups);
  int so_next = u0 < u1 ? a.u_search[u0] : 0, slot_next = u0 < u1 ? a.u_slot[u0] : 0;
  bool prefetched = false;
  for (int u = u0; u < u1; ++u) {
    const int so = so_next, slot = slot_next;
    if (u + 1 < u1) { so_next = a.u_search[u + 1]; slot_next = a.u_slot[u + 1]; }
    for (int tile = 0; tile < a.NT; ++tile) {
      const size_t ti = (size_t)so * a.NT + tile;
      if (!prefetched) copy_to_smem(Aimg, a.A_out + (((size_t)slot * 2 + a.role) * a.NT + tile) * IMG, IMG, gt, GT);
      prefetched = false;
      cp_async_commit();
      {   // L2 prefetch of the image after this one (the shared-memory copy can only start once both projections have read Aimg)
        int ns = slot, nt = tile + 1;
        if (nt == a.NT) { ns = slot_next; nt = 0; }
        if ((nt != 0 || u + 1 < u1) && gt < 128) prefetch_l2_16k(a.A_out + (((size_t)ns * 2 + a.role) * a.NT + nt) * IMG, gt);
      }
      cp_async_wait<0>();
      publish();
      if (issuer) { if (tc::elect_one()) { issue_gemm<4>(tmem, oA, oWk, id64, false); tc::umma_commit(bar); } __syncwarp(); }
      gwait();
      if (tile > 0) { tc::mbar_wait(bar2, par2); par2 ^= 1u; tc::tc_fence_after(); }   // previous KV GEMM still reads KfV
      {   // Kf = elu(k)+1 -> chunks 0..7 ; zero for padding rows (point index >= npts): they must not enter KV / Ksum
        const bool keep = tile * 128 + row < a.npts;
#pragma unroll
        for (int qq = 0; qq < QB; ++qq) {
          const int q = half * QB + qq;
          uint32_t r[16];
          tc::tmem_ld16(tlane + 16 * q, r);
          tc::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              w[j] = keep ? F::elu1(__uint_as_float(r[c * 8 + 2 * j]), __uint_as_float(r[c * 8 + 2 * j + 1])) : 0u;
            *reinterpret_cast<uint4*>(KfV + (2 * q + c) * 2048 + row * 16) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
      tc::tc_fence_before();
      gsync();                                                            // everybody has read k before v overwrites the columns
      tc::tc_fence_after();
      if (issuer) { if (tc::elect_one()) { issue_gemm<4>(tmem, oA, oWv, id64, false); tc::umma_commit(bar); } __syncwarp(); }
      {   // V = v + Wv pos -> chunks 8..15
        uint4 sdPV[CB];
        load_side<CB>(sdPV, a.PV + ti * IMG, half * CB, row);
        gwait();
        if (tile + 1 < a.NT) {   // both projections have consumed `a`: stream the unit's next tile in behind the V epilogue + KV GEMM
          copy_to_smem(Aimg, a.A_out + (((size_t)slot * 2 + a.role) * a.NT + tile + 1) * IMG, IMG, gt, GT);
          cp_async_commit();
          prefetched = true;
        }
#pragma unroll
        for (int qq = 0; qq < QB; ++qq) {
          const int q = half * QB + qq;
          uint32_t r[16];
          tc::tmem_ld16(tlane + 16 * q, r);
          tc::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const uint4 s4 = sdPV[2 * qq + c];
            const uint32_t sw[4] = {s4.x, s4.y, s4.z, s4.w};
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              w[j] = F::add(__uint_as_float(r[c * 8 + 2 * j]), __uint_as_float(r[c * 8 + 2 * j + 1]), sw[j]);
            *reinterpret_cast<uint4*>(KfV + (8 + 2 * q + c) * 2048 + row * 16) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
      publish();
      if (issuer) {   // KV += [Kf|V]^T [V|1]
        if (tc::elect_one()) { issue_gemm<8>(tmem + KVC, oKfV, oVones, idkv, tile > 0); tc::umma_commit(bar2); }
        __syncwarp();
      }
    }
    tc::mbar_wait(bar2, par2);
    par2 ^= 1u;
    tc::tc_fence_after();
    {   // B7 = [head-split blockdiag(KV) Wm^T | Ksum dots] of this (pair, direction) as template
      float kv[64 / CS];
      float ksum = 0.f;
      if (row < 64) {
#pragma unroll
        for (int qq = 0; qq < QB; ++qq) {
          uint32_t r[16];
          tc::tmem_ld16(tlane + KVC + 16 * (half * QB + qq), r);
          tc::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) kv[16 * qq + j] = __uint_as_float(r[j]) * a.kv_scale;
        }
        uint32_t r8[8];
        tc::tmem_ld8(tlane + KVC + 64, r8);
        tc::tmem_ld_wait();
        ksum = __uint_as_float(r8[0]) * a.kv_scale;
        const int hd = row >> 5;
#pragma unroll
        for (int cc = 0; cc < CB; ++cc) {
          const int c = half * CB + cc;
          uint32_t w[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) w[j] = ((c >> 2) == hd) ? F::pack(kv[cc * 8 + 2 * j], kv[cc * 8 + 2 * j + 1]) : 0u;
          *reinterpret_cast<uint4*>(Aimg + c * 2048 + row * 16) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      } else {
#pragma unroll
        for (int cc = 0; cc < CB; ++cc) *reinterpret_cast<uint4*>(Aimg + (half * CB + cc) * 2048 + row * 16) = make_uint4(0, 0, 0, 0);
      }
      publish();
      if (issuer) { if (tc::elect_one()) { issue_gemm<4>(tmem, oA, oWm, id64, false); tc::umma_commit(bar); } __syncwarp(); }
      gwait();
      if (u + 1 < u1) {   // G6 has consumed the operand buffer: the next unit's first tile streams in behind the B7 write-out
        copy_to_smem(Aimg, a.A_out + ((size_t)slot_next * 2 + a.role) * a.NT * IMG, IMG, gt, GT);
        cp_async_commit();
        prefetched = true;
      }
      if (row < 64) {
#pragma unroll
        for (int qq = 0; qq < QB; ++qq) {
          uint32_t r[16];
          tc::tmem_ld16(tlane + 16 * (half * QB + qq), r);
          tc::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) kv[16 * qq + j] = __uint_as_float(r[j]);
        }
        uint8_t* dst = a.B7_out + ((size_t)slot * 2 + a.role) * B7_BYTES;
        if constexpr (CS == 1) write_b7_row<F>(kv, ksum, row, dst);
        else write_b7_part<F>(kv, 4 * half, ksum, half == 0, row, dst);
      }
      tc::tc_fence_before();
      gsync();
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(tmem_base_s, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// phase 1b, second cut: the stage-1 output `a` never enters shared memory.  Each thread loads ITS row of the tile's image from
// global memory (8 x 16 B, a whole tile ahead) and writes it into tensor-memory columns [136, 168) of its lane: the TS-mode A
// operand of both projections (before: a cp.async image that the tensor core read twice).  The key/value reduction runs as an
// M = 64 GEMM (A = the Kf half of the image only; the old M = 128 form also computed V^T V into lanes nobody read) with N = 72
// accumulator columns, which is what makes room: 64 (k / v) + 72 (KV | Ksum) + 32 (`a`) = 168 columns per group (3 x 168 <= 512).
// An M = 64 accumulator lives in lanes 0-15 / 32-47 / 64-79 / 96-111 (row i -> lane (i % 16) + 32 (i / 16)): the B7 build reads
// it with the low 16 lanes of each warp.  Same sums in the same order as pair_p1b_kernel: bit-identical.
// Per tile this takes 16 KB of cp.async fill, 2 x 16 KB of projection A-operand reads and 16 KB of key/value A-operand reads off
// the shared-memory pipe (1366 -> ~850 wavefronts).
// ---------------------------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(NGX * GX, 1) pair_p1b2_kernel(const P1Args a) {
  constexpr int TG = 168, KVC = 64, AC = 136;                            // TMEM columns per group; KV accumulator; `a` operand
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[2 * NGX];
  __shared__ uint32_t tmem_base_s;
  __shared__ float ksum_s[NGX][64];
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
  const int warp_u = (int)tc::uniform(threadIdx.x >> 5);
  const int gid = warp_u / 4, wq = warp_u % 4, row = threadIdx.x % GX, lane = threadIdx.x & 31;
  const bool issuer = wq == 0;
  const uint32_t tmem = tc::uniform(tmem_base_s) + gid * TG;
  const uint32_t tlane = tmem + ((uint32_t)(wq * 32) << 16);
  uint64_t* bar = bars + gid;
  uint32_t par = 0;
  auto gsync = [&]() { tc::bar_sync(1 + gid, GX); };
  auto publish = [&]() { tc::fence_async_smem(); tc::tc_fence_before(); gsync(); tc::tc_fence_after(); };
  auto gwait = [&]() { tc::mbar_wait(bar, par); par ^= 1u; tc::tc_fence_after(); };
  uint64_t* bar2 = bars + NGX + gid;
  uint32_t par2 = 0;
  uint8_t* G = smem + P1B_WBYTES + gid * P1B_GBYTES;
  uint8_t* Aimg = G + P1B_AIMG;                                          // only the B7 build's blockdiag(KV) operand lives here now
  uint8_t* KfV = G + P1B_KFV;
  {
    uint4* ones = reinterpret_cast<uint4*>(KfV + 2 * IMG);
    ones[row] = make_uint4(F::ONE_LO, 0, 0, 0);
    ones[128 + row] = make_uint4(0, 0, 0, 0);
  }
  const uint32_t sA = tc::smem_u32(Aimg), sKfV = tc::smem_u32(KfV), sW = tc::smem_u32(Wsm);
  const uint32_t id64 = tc::instr_desc(128, 64, F::FMT, tc::MAJOR_K, tc::MAJOR_K);
  const uint32_t idkv = tc::instr_desc(64, 72, F::FMT, tc::MAJOR_MN, tc::MAJOR_MN);
  const Opnd oA = A_IMG(sA), oWk = W_IMG(sW + P1B_WKV, 128), oWv = W_IMG(sW + P1B_WKV + 64 * 16, 128), oWm = W_IMG(sW + P1B_WM, 64),
             oKf = opnd(sKfV, 128u, 2048u, 256u), oVones = opnd(sKfV + 8 * 2048, 128u, 2048u, 256u);

  const int ngroups = gridDim.x * NGX, gg = blockIdx.x * NGX + gid;
  const int u0 = (int)((long long)a.n_units * gg / ngroups), u1 = (int)((long long)a.n_units * (gg + 1) / ngroups);
  int so_next = u0 < u1 ? a.u_search[u0] : 0, slot_next = u0 < u1 ? a.u_slot[u0] : 0;
  // this thread's row of an `a` tile image: chunk c (channels 8c .. 8c+7) = words 4c .. 4c+3 of the TMEM operand
  uint32_t areg[32];
  auto load_row = [&](const uint8_t* img) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint4 v = ldg_early(img + c * 2048 + row * 16);
      areg[4 * c] = v.x; areg[4 * c + 1] = v.y; areg[4 * c + 2] = v.z; areg[4 * c + 3] = v.w;
    }
  };
  if (u0 < u1) load_row(a.A_out + ((size_t)slot_next * 2 + a.role) * a.NT * IMG);
  for (int u = u0; u < u1; ++u) {
    const int so = so_next, slot = slot_next;
    if (u + 1 < u1) { so_next = a.u_search[u + 1]; slot_next = a.u_slot[u + 1]; }
    for (int tile = 0; tile < a.NT; ++tile) {
      const size_t ti = (size_t)so * a.NT + tile;
      // the previous tile's v projection (the last reader of the operand columns) was waited for: overwrite them
      tc::tmem_st32(tlane + AC, areg);
      tc::tmem_st_wait();
      tc::tc_fence_before();
      gsync();
      tc::tc_fence_after();
      if (issuer) {
        if (tc::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) tc::umma_f16_ts(tmem, tmem + AC + 8 * ks, oWk.desc + (uint64_t)(ks * oWk.kstep), id64, ks > 0 ? 1u : 0u);
          tc::umma_commit(bar);
        }
        __syncwarp();
      }
      {   // the next tile's row -> registers now (consumed a whole tile later), the one after that -> L2
        int ns = slot, nt = tile + 1;
        if (nt == a.NT) { ns = slot_next; nt = 0; }
        if (nt != 0 || u + 1 < u1) {
          load_row(a.A_out + (((size_t)ns * 2 + a.role) * a.NT + nt) * IMG);
          int ns2 = ns, nt2 = nt + 1;
          if (nt2 == a.NT) { nt2 = 0; ns2 = -1; }                        // (the unit after next is not known yet: only within a unit)
          if (ns2 >= 0) prefetch_l2_16k(a.A_out + (((size_t)ns2 * 2 + a.role) * a.NT + nt2) * IMG, row);
        }
      }
      gwait();
      if (tile > 0) { tc::mbar_wait(bar2, par2); par2 ^= 1u; tc::tc_fence_after(); }   // previous KV GEMM still reads KfV
      {   // Kf = elu(k)+1 -> chunks 0..7 ; zero for padding rows (point index >= npts): they must not enter KV / Ksum
        const bool keep = tile * 128 + row < a.npts;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t r[16];
          tc::tmem_ld16(tlane + 16 * q, r);
          tc::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              w[j] = keep ? F::elu1(__uint_as_float(r[c * 8 + 2 * j]), __uint_as_float(r[c * 8 + 2 * j + 1])) : 0u;
            *reinterpret_cast<uint4*>(KfV + (2 * q + c) * 2048 + row * 16) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
      tc::tc_fence_before();
      gsync();                                                            // everybody has read k before v overwrites the columns
      tc::tc_fence_after();
      if (issuer) {
        if (tc::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) tc::umma_f16_ts(tmem, tmem + AC + 8 * ks, oWv.desc + (uint64_t)(ks * oWv.kstep), id64, ks > 0 ? 1u : 0u);
          tc::umma_commit(bar);
        }
        __syncwarp();
      }
      {   // V = v + Wv pos -> chunks 8..15
        uint4 sdPV[8];
        load_side<8>(sdPV, a.PV + ti * IMG, 0, row);
        gwait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t r[16];
          tc::tmem_ld16(tlane + 16 * q, r);
          tc::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const uint4 s4 = sdPV[2 * q + c];
            const uint32_t sw[4] = {s4.x, s4.y, s4.z, s4.w};
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              w[j] = F::add(__uint_as_float(r[c * 8 + 2 * j]), __uint_as_float(r[c * 8 + 2 * j + 1]), sw[j]);
            *reinterpret_cast<uint4*>(KfV + (8 + 2 * q + c) * 2048 + row * 16) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
      publish();
      if (issuer) {   // KV += Kf^T [V|1]   (M = 64: the Kf half of the image)
        if (tc::elect_one()) { issue_gemm<8>(tmem + KVC, oKf, oVones, idkv, tile > 0); tc::umma_commit(bar2); }
        __syncwarp();
      }
    }
    tc::mbar_wait(bar2, par2);
    par2 ^= 1u;
    tc::tc_fence_after();
    {   // B7 = [per-head blockdiag(KV) Wm^T | Ksum] of this (pair, direction) as template
      {   // accumulator row i = 16 wq + l sits in lane 32 wq + l (l < 16): the low half of every warp turns it into operand row i
        float kv[64];
        uint32_t r8[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t r[16];
          tc::tmem_ld16(tlane + KVC + 16 * q, r);
          tc::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) kv[16 * q + j] = __uint_as_float(r[j]) * a.kv_scale;
        }
        tc::tmem_ld8(tlane + KVC + 64, r8);
        tc::tmem_ld_wait();
        if (lane < 16) {
          const int i = 16 * wq + lane, hd = i >> 5;
          ksum_s[gid][i] = __uint_as_float(r8[0]) * a.kv_scale;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) w[j] = ((c >> 2) == hd) ? F::pack(kv[c * 8 + 2 * j], kv[c * 8 + 2 * j + 1]) : 0u;
            *reinterpret_cast<uint4*>(Aimg + c * 2048 + i * 16) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
      if (row >= 64) {
#pragma unroll
        for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(Aimg + c * 2048 + row * 16) = make_uint4(0, 0, 0, 0);
      }
      publish();
      if (issuer) { if (tc::elect_one()) { issue_gemm<4>(tmem, oA, oWm, id64, false); tc::umma_commit(bar); } __syncwarp(); }
      gwait();
      if (row < 64) {
        float kv[64];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t r[16];
          tc::tmem_ld16(tlane + 16 * q, r);
          tc::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) kv[16 * q + j] = __uint_as_float(r[j]);
        }
        write_b7_row<F>(kv, ksum_s[gid][row], row, a.B7_out + ((size_t)slot * 2 + a.role) * B7_BYTES);
      }
      tc::tc_fence_before();
      gsync();
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(tmem_base_s, 512);
}

// Tried on top of this kernel and dropped (profiles/r02_p1b2_ab.json): pipelining the projections BEHIND the epilogues -- read an
// accumulator into registers, issue the next GEMM over its columns (v(n) after k(n) is in registers, k(n+1) after v(n) is, with the
// Kf image double-buffered through Aimg and a(n+2) prefetched), then do the elu / add arithmetic while the tensor core works.
// Bit-identical, 155-168 registers, and 14-17 % SLOWER (1.23-1.27 ms against 1.05-1.09): the tile chain of this kernel is not
// what bounds it either.

// ---------------------------------------------------------------------------------------------------------------
// packing kernels (fp32 per-object tensors of the parity path -> bf16 operand images)
// ---------------------------------------------------------------------------------------------------------------
// src (B, C, N) channel-major fp32 -> dst [B][N/128][C/8][128][8] bf16, optional elu+1
template <class F>
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
      w[j] = F::pack(x0, x1);
    }
  }
  *reinterpret_cast<uint4*>(dst + (((size_t)b * nt + tile) * nch + chunk) * 2048 + row * 16) = make_uint4(w[0], w[1], w[2], w[3]);
}

// M (B, 64 d, 64 out) fp32 [= blockdiag(KV) Wm^T rows, before head masking] + ksum (B, 64) -> attention operand images
template <class F>
__global__ void __launch_bounds__(64) pack_b7_kernel(const float* __restrict__ M, const float* __restrict__ ksum,
                                                     uint8_t* __restrict__ dst) {
  const int b = blockIdx.x, d = threadIdx.x;
  float m[64];
#pragma unroll
  for (int j = 0; j < 64; ++j) m[j] = M[((size_t)b * 64 + d) * 64 + j];
  write_b7_row<F>(m, ksum[(size_t)b * 64 + d], d, dst + (size_t)b * B7_BYTES);
}

}  // namespace

template <class F, int CS>
static int launch_p1b_cs(const P1Args& a, int grid, cudaStream_t st) {
  const int smem = P1B_WBYTES + NGX * P1B_GBYTES;
  cudaFuncSetAttribute(pair_p1b_kernel<F, CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  pair_p1b_kernel<F, CS><<<grid, NGX * GX * CS, smem, st>>>(a);
  return pcreid_launch_status();
}
template <class F>
static int launch_p1b(const P1Args& a, int grid, cudaStream_t st) {
  // default: pair_p1b2_kernel (`a` in tensor memory, M = 64 key/value GEMM: -3 %, profiles/r02_p1b2_ab.json).
  // A/B: PCREID_P1B=1 -> the shared-memory form (pair_p1b_kernel); with it PCREID_P1B_SPLIT=2 -> two threads per tile row
  static const int gen = [] { const char* e = getenv("PCREID_P1B"); return e ? atoi(e) : 2; }();
  static const int cs = [] { const char* e = getenv("PCREID_P1B_SPLIT"); return e ? atoi(e) : 1; }();
  if (gen == 2) {
    const int smem = P1B_WBYTES + NGX * P1B_GBYTES;
    cudaFuncSetAttribute(pair_p1b2_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    pair_p1b2_kernel<F><<<grid, NGX * GX, smem, st>>>(a);
    return pcreid_launch_status();
  }
  return cs == 2 ? launch_p1b_cs<F, 2>(a, grid, st) : launch_p1b_cs<F, 1>(a, grid, st);
}

extern "C" {

int pcreid_pack_image(int B, int C, int N, const float* src, long long s_bs, int lds, int act, int fmt, void* dst, void* stream) {
  if (B <= 0) return PCREID_OK;
  if (!src || !dst || C % 8 || N <= 0 || (fmt != PCREID_FMT_BF16 && fmt != PCREID_FMT_F16)) return PCREID_ERR_ARG;
  const long long per = 128LL * (C / 8) * ((N + 127) / 128);
  const unsigned grid = (unsigned)((per * B + 255) / 256);
  if (fmt == PCREID_FMT_F16)
    pack_image_kernel<tc::OpF16><<<grid, 256, 0, (cudaStream_t)stream>>>(B, C, N, src, s_bs, lds, act, (uint8_t*)dst);
  else
    pack_image_kernel<tc::OpBF16><<<grid, 256, 0, (cudaStream_t)stream>>>(B, C, N, src, s_bs, lds, act, (uint8_t*)dst);
  return pcreid_launch_status();
}

int pcreid_pack_b7(int B, const float* M, const float* ksum, int fmt, void* dst, void* stream) {
  if (B <= 0) return PCREID_OK;
  if (!M || !ksum || !dst || (fmt != PCREID_FMT_BF16 && fmt != PCREID_FMT_F16)) return PCREID_ERR_ARG;
  if (fmt == PCREID_FMT_F16) pack_b7_kernel<tc::OpF16><<<B, 64, 0, (cudaStream_t)stream>>>(M, ksum, (uint8_t*)dst);
  else pack_b7_kernel<tc::OpBF16><<<B, 64, 0, (cudaStream_t)stream>>>(M, ksum, (uint8_t*)dst);
  return pcreid_launch_status();
}

int pcreid_pair_p1b_n(int n_units, int npts, int role, int fmt, float kv_scale, const int* u_search, const int* u_templ,
                      const int* u_slot, const void* PV, const void* W, void* A_out, void* B7_out, int n_ctas, void* stream) {
  if (n_units <= 0) return PCREID_OK;
  if (!u_search || !u_templ || !u_slot || !PV || !W || !A_out || !B7_out || npts <= 0 ||
      (fmt != PCREID_FMT_BF16 && fmt != PCREID_FMT_F16))
    return PCREID_ERR_ARG;
  const int NT = (npts + 127) / 128;
  P1Args a{n_units, NT, role, npts, 0.f, kv_scale, u_search, u_templ, u_slot, nullptr, nullptr, (const uint8_t*)PV, nullptr,
           (const uint8_t*)W, (uint8_t*)A_out, (uint8_t*)B7_out};
  int grid = n_ctas > 0 ? n_ctas : 148;
  if (grid * NGX > n_units) grid = (n_units + NGX - 1) / NGX;
  return fmt == PCREID_FMT_F16 ? launch_p1b<tc::OpF16>(a, grid, (cudaStream_t)stream) : launch_p1b<tc::OpBF16>(a, grid, (cudaStream_t)stream);
}

}  // extern "C"
