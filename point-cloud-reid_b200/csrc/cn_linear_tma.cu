// cn_linear on the tensor cores, third generation: operands staged by TMA tensor maps (cp.async.bulk.tensor) straight
// from the tensors as they lie in HBM -- no packed weight images, no register transposes.
//
//   Y[b, co, n] = act( sum_k W1[k, co] X1[b, k, n] + sum_k W2[k, co] X2[b, k, n] + bias[co] (+R) ) (+R)
//
// = every nn.Linear / 1x1 conv of the encoders and heads (DGCNN EdgeConv projections dgcnn_orig.py:129-152, PointNet
// shared MLPs pointnet.py:88-127, attention projections attention.py:192-219, LinearRes heads lanegcn_nets.py:206-221).
//
// Both operands are MN-major for the tensor core exactly as they are stored:
//   activations (B, K, N) channel-major: the point axis (M of the MMA) is contiguous, the channel axis (K) strided;
//   weights     (K, CO)   k-major:       the output channel axis (N of the MMA) is contiguous.
// MN-major 32-bit operands have ONE swizzled shared-memory layout on tcgen05: 128-byte rows whose four 32-byte chunks are
// XOR-ed with the row index mod 4 (descriptor layout type SWIZZLE_128B_BASE32B, cute Swizzle<2,5,2>; the plain 128-byte
// swizzle yields zeros) -- the TMA swizzle mode CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B writes exactly that.  A 3-D tensor map
// (points | channels | objects) with a 32 x 32 x 1 box lands one [32 k][32 points] fp32 block as eight such atoms (4 k-rows
// x 128 B each) stacked along K: four boxes make the 128-point A tile of a stage, TN/32 boxes of the weight map its B tile.  Out-of-range points, channels (K not a
// multiple of 32, K = 3 inputs) and output channels are zero-filled by the TMA unit, so there is no tail code.
// Object gather maps (x1_map / x2_map / w1_map: the all-pairs matcher's track / detection indices) are just the third
// box coordinate.
//
// One persistent CTA of 10 warps per SM, tiles ordered (object, point tile, channel tile) with the channel tile fastest so that
// concurrently running CTAs share an activation tile through L2:
//   warp 8 (one lane)  producer: arms full[s] with the stage's byte count, issues the box copies (6 stages of 32 KB);
//   warp 9 (one lane)  MMA issuer: tcgen05.mma kind::tf32 M=128 N=TN K=8 per 8 k-rows, tcgen05.commit -> empty[s];
//                      four accumulator buffers of 128 TMEM columns: tcgen05.commit -> acc_full[buf];
//   warps 0-7          epilogue, two groups of four warps (thread == point row == TMEM lane; group g owns the 32-channel
//                      slabs g, g+2 of the tile): tcgen05.ld -> bias / activation / residual (/ tf32 rounding for a following
//                      GEMM) -> [32 channels][128 points] staging tile in shared memory (conflict-free: lanes = consecutive
//                      points) -> ONE TMA tensor store per slab (cp.async.bulk.tensor ... global.shared::cta), which also
//                      clips ragged point / channel tails.  The first version of this epilogue stored from registers
//                      (thread per row, 128 scalar STG with per-element predicates): 75 warp-instructions per store, 0.9 TB/s of
//                      output regardless of K (profiles/r02_cn_linear_tma.md).
// The tensor core reads only the upper 19 bits of an fp32 operand (truncation); with `tf32_maps` the tensor maps are
// encoded as TFLOAT32, which makes the TMA unit round the data to tf32 (nearest) on its way into shared memory.
//
// X3 variant (pcreid_cn_linear_tma_x3, the fp32-grade mode): every operand is split x = hi + lo with hi = the 19 bits the
// tensor core reads anyway (so the raw fp32 tile IS the hi operand) and lo = x - hi (exact; 13 significant bits, of which the
// tensor core keeps 11).  The weights' lo part comes from the host as a second tensor, the activations' lo tile is written by
// four converter warps into a second shared-memory tile once the stage has landed, and each K = 8 step issues three MMAs
// lo.hi + hi.lo + hi.hi into the same accumulator (small terms first).  Products of 11-bit significands are exact in the
// fp32 accumulator; the dropped lo.lo term is 2^-20 relative: the result matches the fp32 FFMA kernel to ~1e-6 relative at a third of
// the tf32 tensor rate -- still far above the HBM roofline of the K <= 256 contractions it is used for.
#include <cuda.h>

#include "../../include/pcreid.h"
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int ST = 6;                     // pipeline stages (4 with 256-channel tiles)
constexpr int NACC = 4;                   // accumulator buffers of 128 TMEM columns (2 of 256 with 256-channel tiles)
constexpr int KC = 32;                    // channels (K) per stage
constexpr int BOX_BYTES = 32 * KC * 4;    // one 32 x 32 fp32 box = 4 KB = eight swizzle atoms stacked along K
constexpr int STAGE_A = 4 * BOX_BYTES;    // 128 points
constexpr int STAGE_BYTES = 2 * STAGE_A;  // A + up to 128 output channels (48 KB stages with 256-channel tiles)
constexpr int RING_BYTES = ST * STAGE_BYTES;
constexpr int SLAB_BYTES = 32 * 128 * 4;  // epilogue staging tile: 32 channels x 128 points
constexpr int SMEM_BYTES = RING_BYTES + 2 * SLAB_BYTES + 1024;
constexpr int NTHR = 320;
constexpr int NTHR_X3 = 448;               // + four converter warps
constexpr int ST_X3 = 3;                  // X3 stages: A | A_lo | W | W_lo = 64 KB

struct TmaLinArgs {
  pcreid_linear_args a;
  int tiles_n, tiles_c, total_tiles, TN, round_out;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_box3(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(tc::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_store3(const CUtensorMap* map, int c0, int c1, int c2, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2),
               "r"(src)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_map(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

template <int ACT>
__device__ __forceinline__ float act_t(float v) {
  if (ACT == ACT_RELU) return fmaxf(v, 0.f);
  if (ACT == ACT_LEAKY02) return fmaxf(v, 0.2f * v);
  if (ACT == ACT_ELU1) return v > 0.f ? v + 1.f : __expf(v);
  return v;
}

// 32 accumulator columns of this thread's row -> epilogue arithmetic -> column `row` of the [32][128] staging slab
template <int ACT, int RES>      // RES: 0 none, 1 before the activation, 2 after it
__device__ __forceinline__ void slab_to_smem(uint32_t taddr, float* slab, int row, const float* __restrict__ bias, const float* __restrict__ R,
                                             int ldr, int co0, int CO, bool row_ok, bool round_out) {
  uint32_t rg[32];
  tc::tmem_ld32(taddr, rg);
  float rr[32];
  if (RES) {
#pragma unroll
    for (int j = 0; j < 32; ++j) rr[j] = (row_ok && co0 + j < CO) ? __ldg(R + (size_t)(co0 + j) * ldr) : 0.f;
  }
  tc::tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float v = __uint_as_float(rg[j]);
    if (bias) v += (co0 + j < CO) ? __ldg(bias + co0 + j) : 0.f;
    if (RES == 1) v += rr[j];
    v = act_t<ACT>(v);
    if (RES == 2) v += rr[j];
    if (round_out) v = tc::tf32_rna(v);
    slab[j * 128 + row] = v;
  }
}

template <int ACT, bool X3>
__global__ void __launch_bounds__(X3 ? NTHR_X3 : NTHR, 1)
cn_linear_tma_kernel(const __grid_constant__ CUtensorMap mX1, const __grid_constant__ CUtensorMap mX2, const __grid_constant__ CUtensorMap mW1,
                     const __grid_constant__ CUtensorMap mW2, const __grid_constant__ CUtensorMap mW1lo, const __grid_constant__ CUtensorMap mW2lo,
                     const __grid_constant__ CUtensorMap mY, const __grid_constant__ TmaLinArgs p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full[ST], empty[ST], conv[ST], acc_full[NACC], acc_empty[NACC];
  __shared__ uint32_t tmem_base_s;
  // the swizzle pattern is a function of the shared-memory address bits: the stages start on a 1 KB boundary
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const pcreid_linear_args& a = p.a;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < ST; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); tc::mbar_init(&conv[i], 128); }
    for (int i = 0; i < NACC; ++i) { tc::mbar_init(&acc_full[i], 1); tc::mbar_init(&acc_empty[i], 256); }
    tc::fence_mbar_init();
  }
  if (warp == 0) { tc::tmem_alloc(&tmem_base_s, 512); tc::tmem_relinquish(); }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const int nch1 = (a.K1 + KC - 1) / KC, nch2 = a.K2 > 0 ? (a.K2 + KC - 1) / KC : 0, nch = nch1 + nch2;
  const int per_obj = p.tiles_n * p.tiles_c;
  const int TN = p.TN;
  // 256-channel tiles (large K and CO: halves the activation re-reads from L2): 4 stages of 48 KB, 2 accumulators of 256 columns
  const int nst = X3 ? ST_X3 : (TN > 128 ? 4 : ST), nacc = TN > 128 ? 2 : NACC, acc_cols = TN > 128 ? 256 : 128;
  const int stage_bytes = X3 ? 4 * STAGE_A : (TN > 128 ? 3 * STAGE_A : STAGE_BYTES);
  // X3 stage: A at 0, A_lo at STAGE_A, W at 2 STAGE_A, W_lo at 3 STAGE_A; else A at 0, W at STAGE_A
  const uint32_t off_b = X3 ? 2 * STAGE_A : STAGE_A;

  if (warp == 8) {
    // ============================================================ producer (one thread)
    if (lane == 0) {
      prefetch_map(&mX1); prefetch_map(&mW1);
      if (nch2) { prefetch_map(&mX2); prefetch_map(&mW2); }
      int g = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int b = tile / per_obj, n0 = ((tile % per_obj) / p.tiles_c) * 128, co0 = (tile % p.tiles_c) * TN;
        const int bx1 = a.x1_bs ? (a.x1_map ? __ldg(a.x1_map + b) : b) : 0;
        const int bx2 = (nch2 && a.x2_bs) ? (a.x2_map ? __ldg(a.x2_map + b) : b) : 0;
        const int bw1 = a.w1_bs ? (a.w1_map ? __ldg(a.w1_map + b) : b) : 0;
        const int bw2 = a.w2_bs ? b : 0;
        const int nbox_a = min(4, (a.rows - n0 + 31) / 32), nbox_b = min(TN / 32, (a.CO - co0 + 31) / 32);
        for (int c = 0; c < nch; ++c, ++g) {
          const bool second = c >= nch1;
          const int k0 = (second ? c - nch1 : c) * KC;
          const int s = g % nst, use = g / nst;
          if (use > 0) tc::mbar_wait(&empty[s], (uint32_t)((use - 1) & 1));
          const uint32_t sa = tc::smem_u32(smem + s * stage_bytes), sb = sa + off_b;
          mbar_expect_tx(&full[s], (uint32_t)((nbox_a + (X3 ? 2 : 1) * nbox_b) * BOX_BYTES));
          const CUtensorMap* mx = second ? &mX2 : &mX1;
          const CUtensorMap* mw = second ? &mW2 : &mW1;
          for (int i = 0; i < nbox_a; ++i) tma_box3(sa + i * BOX_BYTES, mx, n0 + 32 * i, k0, second ? bx2 : bx1, &full[s]);
          for (int j = 0; j < nbox_b; ++j) tma_box3(sb + j * BOX_BYTES, mw, co0 + 32 * j, k0, second ? bw2 : bw1, &full[s]);
          if (X3) {
            const CUtensorMap* ml = second ? &mW2lo : &mW1lo;
            for (int j = 0; j < nbox_b; ++j) tma_box3(sb + STAGE_A + j * BOX_BYTES, ml, co0 + 32 * j, k0, second ? bw2 : bw1, &full[s]);
          }
        }
      }
    }
  } else if (warp == 9) {
    // ============================================================ MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc = tc::instr_desc(128, TN, tc::FMT_TF32, tc::MAJOR_MN, tc::MAJOR_MN);
      int g = 0, ti = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++ti) {
        const int buf = ti % nacc, au = ti / nacc;
        if (au > 0) tc::mbar_wait(&acc_empty[buf], (uint32_t)((au - 1) & 1));
        tc::tc_fence_after();
        const uint32_t d = tmem + (uint32_t)(buf * acc_cols);
        for (int c = 0; c < nch; ++c, ++g) {
          const bool second = c >= nch1;
          const int K = second ? a.K2 : a.K1, k0 = (second ? c - nch1 : c) * KC;
          const int ksteps = (min(KC, K - k0) + 7) / 8;
          const int s = g % nst, use = g / nst;
          tc::mbar_wait(X3 ? &conv[s] : &full[s], (uint32_t)(use & 1));
          tc::tc_fence_after();
          const uint32_t sa = tc::smem_u32(smem + s * stage_bytes), sb = sa + off_b;
          // MN-major SWIZZLE_128B_BASE32B: leading byte offset = distance between 32-element blocks along M / N (one box),
          // stride byte offset = distance between groups of 4 k-rows (one 512 B atom); a K = 8 step spans two atoms
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t ad = tc::smem_desc(sa + ks * 1024, BOX_BYTES, 512, tc::LAYOUT_SW128_BASE32B);
            const uint64_t bd = tc::smem_desc(sb + ks * 1024, BOX_BYTES, 512, tc::LAYOUT_SW128_BASE32B);
            if (X3) {   // lo.hi + hi.lo first, hi.hi last
              const uint64_t ald = tc::smem_desc(sa + STAGE_A + ks * 1024, BOX_BYTES, 512, tc::LAYOUT_SW128_BASE32B);
              const uint64_t bld = tc::smem_desc(sb + STAGE_A + ks * 1024, BOX_BYTES, 512, tc::LAYOUT_SW128_BASE32B);
              tc::umma_tf32(d, ald, bd, idesc, (c > 0 || ks > 0) ? 1u : 0u);
              tc::umma_tf32(d, ad, bld, idesc, 1u);
              tc::umma_tf32(d, ad, bd, idesc, 1u);
            } else {
              tc::umma_tf32(d, ad, bd, idesc, (c > 0 || ks > 0) ? 1u : 0u);
            }
          }
          tc::umma_commit(&empty[s]);
        }
        tc::umma_commit(&acc_full[buf]);
      }
    }
  } else if (X3 && warp >= 10) {
    // ============================================================ converters: A_lo = A - (the 19 bits the tensor core reads of A)
    const int t = tid - NTHR;
    int g = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int n0 = ((tile % per_obj) / p.tiles_c) * 128;
      const int nvec = min(4, (a.rows - n0 + 31) / 32) * (BOX_BYTES / 16);
      for (int c = 0; c < nch; ++c, ++g) {
        const int s = g % nst, use = g / nst;
        tc::mbar_wait(&full[s], (uint32_t)(use & 1));
        const float4* A = reinterpret_cast<const float4*>(smem + s * stage_bytes);
        float4* Alo = reinterpret_cast<float4*>(smem + s * stage_bytes + STAGE_A);
        for (int v = t; v < nvec; v += 128) {
          const float4 x = A[v];
          float4 l;
          l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
          l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
          l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
          l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
          Alo[v] = l;
        }
        tc::fence_async_smem();
        mbar_arrive(&conv[s]);
      }
    }
  } else if (warp < 8) {
    // ============================================================ epilogue: group = warp / 4 owns slabs group, group + 2
    const int grp = warp >> 2, row = tid & 127;
    float* slab = reinterpret_cast<float*>(smem + RING_BYTES + grp * SLAB_BYTES);
    const uint32_t slab_s = tc::smem_u32(slab);
    const bool leader = row == 0;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const int nslab_t = TN / 32;
    int ti = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++ti) {
      const int b = tile / per_obj, n0 = ((tile % per_obj) / p.tiles_c) * 128, co0 = (tile % p.tiles_c) * TN;
      const int buf = ti % nacc, au = ti / nacc;
      const int n = n0 + row;
      const float* R = a.R ? a.R + (size_t)(a.r_map ? __ldg(a.r_map + b) : b) * a.r_bs + n : nullptr;
      tc::mbar_wait(&acc_full[buf], (uint32_t)(au & 1));
      tc::tc_fence_after();
      const uint32_t tl = tmem + (uint32_t)(buf * acc_cols) + lane_off;
      for (int sl = grp; sl < nslab_t; sl += 2) {
        const int c0 = co0 + 32 * sl;
        if (c0 >= a.CO) break;
        if (leader) bulk_wait_read0();            // the previous store out of this staging tile has read it
        tc::bar_sync(1 + grp, 128);
        if (!R) slab_to_smem<ACT, 0>(tl + 32 * sl, slab, row, a.bias, nullptr, 0, c0, a.CO, true, p.round_out);
        else if (!a.res_after_act) slab_to_smem<ACT, 1>(tl + 32 * sl, slab, row, a.bias, R, a.ldr, c0, a.CO, n < a.rows, p.round_out);
        else slab_to_smem<ACT, 2>(tl + 32 * sl, slab, row, a.bias, R, a.ldr, c0, a.CO, n < a.rows, p.round_out);
        tc::fence_async_smem();
        tc::bar_sync(1 + grp, 128);
        if (leader) { tma_store3(&mY, n0, c0, b, slab_s); bulk_commit(); }
      }
      tc::tc_fence_before();
      mbar_arrive(&acc_empty[buf]);
    }
    if (leader) bulk_wait0();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// ---- host side: tensor maps through the driver entry point (no link-time dependency on libcuda) ----------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// (inner | rows | objects) fp32 tensor, strides in elements, zero fill out of range.  Operand maps: box 32 x 32 x 1 with 32-byte
// chunks swizzled within 128-byte rows; the output map: box 128 x 32 x 1, no swizzle (the epilogue's staging slab).
bool make_map(CUtensorMap* m, const float* base, long long inner, long long rows, long long ld, long long objs, long long obj_stride,
              bool tf32, bool output) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld & 3) || ld <= 0 || inner <= 0 || rows <= 0) return false;
  if (objs <= 1 || obj_stride == 0) { objs = 1; obj_stride = ld * rows; }
  if (obj_stride & 3) return false;
  // the third extent is a plain upper bound (object maps may address any object of the source tensor)
  const cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)objs};
  const cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)obj_stride * 4};
  const cuuint32_t box[3] = {output ? 128u : 32u, (cuuint32_t)KC, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(m, tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box,
             estr, CU_TENSOR_MAP_INTERLEAVE_NONE, output ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int ACT, bool X3>
int launch_tma(const CUtensorMap* m, const TmaLinArgs& p, int grid, cudaStream_t st) {
  const int smem = X3 ? ST_X3 * 4 * STAGE_A + 2 * SLAB_BYTES + 1024 : SMEM_BYTES;
  cudaFuncSetAttribute(cn_linear_tma_kernel<ACT, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cn_linear_tma_kernel<ACT, X3><<<grid, X3 ? NTHR_X3 : NTHR, smem, st>>>(m[0], m[1], m[2], m[3], m[4], m[5], m[6], p);
  return pcreid_launch_status();
}

template <bool X3>
int run_tma(const pcreid_linear_args* pa, const float* W1lo, const float* W2lo, long long x1_objs, long long x2_objs, long long w1_objs,
            int flags, int n_sms, void* stream) {
  if (!pa) return PCREID_ERR_ARG;
  const pcreid_linear_args& a = *pa;
  if (a.B <= 0 || a.rows <= 0 || a.CO <= 0) return PCREID_OK;
  if (a.K1 <= 0 || !a.X1 || !a.W1 || !a.Y) return PCREID_ERR_ARG;
  if (a.K2 > 0 && (!a.X2 || !a.W2)) return PCREID_ERR_ARG;
  if (X3 && (!W1lo || (a.K2 > 0 && !W2lo))) return PCREID_ERR_ARG;
  // point-major inputs would be K-major operands (another descriptor family), point-major outputs another staging layout: they
  // stay on pcreid_cn_linear
  // (the TMA unit clips the innermost extent in 16-byte units: rows and CO must be multiples of 4)
  if (a.x1_pm || (a.K2 > 0 && a.x2_pm) || a.y_pm || a.CO < 32 || (a.CO & 3) || (a.rows & 3)) return PCREID_ERR_UNSUPPORTED;
  if (a.act < ACT_NONE || a.act > ACT_ELU1) return PCREID_ERR_ARG;
  if ((a.x1_map && x1_objs <= 0) || (a.K2 > 0 && a.x2_map && x2_objs <= 0) || (a.w1_map && w1_objs <= 0)) return PCREID_ERR_ARG;
  const bool tf32 = !X3 && (flags & PCREID_TMA_TF32_MAPS);      // X3 needs the raw fp32 tile: its hi part is what the tensor core reads
  TmaLinArgs p;
  p.a = a;
  p.round_out = (!X3 && (flags & PCREID_TMA_ROUND_OUT)) ? 1 : 0;
  p.TN = a.CO >= 128 ? ((!X3 && a.CO >= 256 && a.K1 + a.K2 >= 256 && !(flags & PCREID_TMA_TILE128)) ? 256 : 128) : ((a.CO + 31) / 32) * 32;
  p.tiles_n = (a.rows + 127) / 128;
  p.tiles_c = (a.CO + p.TN - 1) / p.TN;
  const long long total = (long long)a.B * p.tiles_n * p.tiles_c;
  if (total > 0x7fffffffLL) return PCREID_ERR_UNSUPPORTED;
  p.total_tiles = (int)total;
  alignas(64) CUtensorMap m[7];      // X1, X2, W1, W2, W1lo, W2lo, Y
  const long long wobjs = a.w1_map ? w1_objs : a.B;
  if (!make_map(&m[0], a.X1, a.rows, a.K1, a.ldx1, a.x1_map ? x1_objs : a.B, a.x1_bs, tf32, false)) return PCREID_ERR_UNSUPPORTED;
  if (!make_map(&m[2], a.W1, a.CO, a.K1, a.CO, wobjs, a.w1_bs, false, false)) return PCREID_ERR_UNSUPPORTED;
  m[4] = m[2];
  if (X3 && !make_map(&m[4], W1lo, a.CO, a.K1, a.CO, wobjs, a.w1_bs, false, false)) return PCREID_ERR_UNSUPPORTED;
  if (a.K2 > 0) {
    if (!make_map(&m[1], a.X2, a.rows, a.K2, a.ldx2, a.x2_map ? x2_objs : a.B, a.x2_bs, tf32, false)) return PCREID_ERR_UNSUPPORTED;
    if (!make_map(&m[3], a.W2, a.CO, a.K2, a.CO, a.B, a.w2_bs, false, false)) return PCREID_ERR_UNSUPPORTED;
    m[5] = m[3];
    if (X3 && !make_map(&m[5], W2lo, a.CO, a.K2, a.CO, a.B, a.w2_bs, false, false)) return PCREID_ERR_UNSUPPORTED;
  } else {
    m[1] = m[0];
    m[3] = m[2];
    m[5] = m[4];
  }
  if (!make_map(&m[6], a.Y, a.rows, a.CO, a.ldy, a.B, a.y_bs, false, true)) return PCREID_ERR_UNSUPPORTED;
  if (n_sms <= 0) n_sms = 148;
  const int grid = total < n_sms ? (int)total : n_sms;
  cudaStream_t st = (cudaStream_t)stream;
  switch (a.act) {
    case ACT_RELU: return launch_tma<ACT_RELU, X3>(m, p, grid, st);
    case ACT_LEAKY02: return launch_tma<ACT_LEAKY02, X3>(m, p, grid, st);
    case ACT_ELU1: return launch_tma<ACT_ELU1, X3>(m, p, grid, st);
    default: return launch_tma<ACT_NONE, X3>(m, p, grid, st);
  }
}

}  // namespace

extern "C" int pcreid_cn_linear_tma(const pcreid_linear_args* pa, long long x1_objs, long long x2_objs, long long w1_objs, int flags,
                                    int n_sms, void* stream) {
  return run_tma<false>(pa, nullptr, nullptr, x1_objs, x2_objs, w1_objs, flags, n_sms, stream);
}

extern "C" int pcreid_cn_linear_tma_x3(const pcreid_linear_args* pa, const float* W1lo, const float* W2lo, long long x1_objs,
                                       long long x2_objs, long long w1_objs, int n_sms, void* stream) {
  return run_tma<true>(pa, W1lo, W2lo, x1_objs, x2_objs, w1_objs, 0, n_sms, stream);
}
