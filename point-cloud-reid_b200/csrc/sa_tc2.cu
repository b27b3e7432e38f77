// Set-abstraction edge MLP, second generation: warp-specialised and pipelined, both GEMM A operands in tensor memory.
//
//   out[b,c,s] = max_j relu(W3 relu(W2 relu(P1[b,idx[b,s,j],:] + Cc[b,s,:]) + b2) + b3)[c]
// (PointNetSetAbstractionEdgeSA.forward, mmdet3d/models/pointnet2_utils.py:333-357; first conv factorised per point /
// per centre on the host, eval BatchNorm folded; same contract as sa_edge_mlp_tc_kernel in sa_tc.cu.)
//
// ncu on the first generation showed a fully serial tile (gather from L2 -> smem image -> MMA -> epilogue -> smem image
// -> MMA -> epilogue -> smem transpose -> max): 7.5 us per 128-edge tile at C = 128 against 1.1 us of tensor-pipe time.
// This version removes every shared-memory activation image:
//   * producer warps (4..7, thread == edge row == TMEM lane) gather relu(P1 + Cc) and write it straight into tensor
//     memory (tcgen05.st) as the A operand of GEMM 1 (kind::tf32, A from TMEM, W2 resident in shared memory);
//   * consumer warps (0..3) apply bias + ReLU to the accumulator IN PLACE (tcgen05.ld / tcgen05.st), which makes it the
//     A operand of GEMM 2; the second accumulator is reduced over the k edges of a centre with warp-level
//     redux.sync.max.f32 (no transpose buffer) and the per-warp partials are combined through 4 KB of shared memory;
//   * the object's per-point term P1 (N x C fp32) is staged once per object in shared memory (rows padded by 16 B), so
//     the gather is shared-memory -> tensor-memory and never waits on L2; Cc rows of the tile are prefetched with
//     cp.async before the producer waits for the A region;
//   * accumulator 1 is double buffered: the gather + GEMM 1 of tile t+1 overlap epilogue 1 / GEMM 2 / epilogue 2 of tile t;
//   * instruction diet (ncu: the first cut of this kernel was issue-bound, 10.4 k warp-instructions per tile at C = 128):
//     b2 is preloaded into accumulator 1 by the producers (GEMM 1 accumulates onto it: epilogue 1 is one FMNMX per
//     element), b3 + ReLU are applied AFTER the max over the edges (both commute with max: epilogue 2 is the redux only),
//     the gather is compiled per address space, TMEM addresses stay on the uniform datapath;
//   * at C = 128 (one CTA per SM) every role runs two warps per TMEM lane quadrant, each on half of the columns.
//   * tiles are DENSE: tile t of an object is its edges [128 t, 128 t + 128) of the S k (row = edge, centre = edge / k), not
//     128 / k whole centres -- at k = 48 every tile carries 128 valid rows instead of 96 (25 % fewer tiles for the same
//     producer / epilogue cost per tile).  A centre whose edges straddle two tiles leaves its partial maximum in a
//     double-buffered shared-memory carry row; unit boundaries fall on tiles that start a centre.
// TMEM columns: A1 | D1[0] | D1[1] | D2 = 4 C (512 at C = 128).  A CTA walks whole objects (or contiguous tile ranges
// of an object when there are fewer objects than CTAs).
#include "../../include/pcreid.h"
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int MAXSEG = 3;      // centres a warp's 32 rows can belong to (k >= 16)

struct Sa2Args {
  int N, S, k, cmax, tpo, upo, tpu, n_units;     // max centres a tile touches, tiles per object, units per object, tiles per unit
  int p1_smem;
  const float* P1;      // (B, N, C) point-major
  const float* Cc;      // (B, S, C) point-major
  const int* idx;       // (B, S, k)
  const float *W2img, *W3img;
  float* out;
  long long o_bs;
  int o_cs, o_ss;       // element (b, c, s) at out + b*o_bs + c*o_cs + s*o_ss
  const float *b2, *b3;
};

__device__ __forceinline__ float redux_max(float v, uint32_t mask) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, %2;" : "=f"(r) : "f"(v), "r"(mask));
  return r;
}

template <int C, int CS, bool P1S>
__global__ void __launch_bounds__(256 * CS, (C == 32 ? 4 : (C == 64 ? (CS == 2 ? 2 : 2) : 1))) sa_edge_mlp_tc2_kernel(const __grid_constant__ Sa2Args a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t mma1_done[2], mma2_done[2];
  __shared__ uint32_t tmem_base_s;
  constexpr int WBYTES = C * C * 4;
  constexpr int PSTR = C + 4;                    // padded P1 row (floats)
  constexpr int RT = 128 * CS;                   // threads per role
  constexpr int CW = C / CS;                     // columns per thread
  const int N = a.N, S = a.S, k = a.k, cmax = a.cmax, E = a.S * a.k;
  uint8_t* W2s = smem;
  uint8_t* W3s = smem + WBYTES;
  float* part = reinterpret_cast<float*>(smem + 2 * WBYTES);          // [4 quadrants][MAXSEG][C]
  float* cc_s = part + 4 * MAXSEG * C;                                // [2][cmax][C]
  float* b2_s = cc_s + 2 * cmax * C;                                  // [C]
  float* b3_s = b2_s + C;                                             // [C]
  float* carry = b3_s + C;                                            // [2][C] partial max of the centre that straddles into the next tile
  float* p1_s = carry + 2 * C;                                        // [N][PSTR] when P1S
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const bool producer = warp >= 4 * CS;
  const int rt = producer ? t - RT : t;                               // thread index inside the role
  const int quad = warp & 3;                                          // TMEM lane quadrant this warp may access
  const int row = quad * 32 + lane;                                   // edge row == TMEM lane
  const int cb = ((warp >> 2) % CS) * CW;                             // first column of this thread
  if (t == 0) {
    tc::mbar_init(&mma1_done[0], 1); tc::mbar_init(&mma1_done[1], 1); tc::mbar_init(&mma2_done[0], 1); tc::mbar_init(&mma2_done[1], 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) { tc::tmem_alloc(&tmem_base_s, 4 * C); tc::tmem_relinquish(); }
  for (int i = t * 16; i < WBYTES; i += 256 * CS * 16) {
    cp_async16(W2s + i, reinterpret_cast<const uint8_t*>(a.W2img) + i);
    cp_async16(W3s + i, reinterpret_cast<const uint8_t*>(a.W3img) + i);
  }
  cp_async_commit();
  for (int i = t; i < C; i += 256 * CS) { b2_s[i] = a.b2[i]; b3_s[i] = a.b3[i]; }
  cp_async_wait<0>();
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tc::uniform(tmem_base_s);
  const uint32_t lane_off = tc::uniform((uint32_t)(quad * 32) << 16);
  const uint32_t tA1 = tmem, tD1 = tmem + C, tD2 = tmem + 3 * C;      // D1[i] = tD1 + i*C
  const uint32_t idesc = tc::instr_desc(128, C, tc::FMT_TF32, tc::MAJOR_K, tc::MAJOR_K);
  const uint32_t sW2 = tc::smem_u32(W2s), sW3 = tc::smem_u32(W3s);

  int tcount = 0;                                // tiles processed by this CTA (barrier parities)
  if (producer) {
    // ================================================= producers: gather -> TMEM A1, bias -> D1, issue GEMM 1
    for (int u = blockIdx.x; u < a.n_units; u += gridDim.x) {
      const int b = u / a.upo, tl0 = (u % a.upo) * a.tpu, tl1 = min(tl0 + a.tpu, a.tpo);
      const float* Pb = a.P1 + (size_t)b * N * C;
      const float* Cb = a.Cc + (size_t)b * S * C;
      const int* Ib = a.idx + (size_t)b * S * k;
      if (P1S) {
        tc::bar_sync(1, RT);                     // every producer finished gathering the previous object
        for (int i = rt; i < N * (C / 4); i += RT) {
          const int r = i / (C / 4), c4 = i % (C / 4);
          cp_async16(p1_s + r * PSTR + 4 * c4, Pb + (size_t)r * C + 4 * c4);
        }
        cp_async_commit();
      }
      int src_next = tl0 * 128 + row < E ? __ldg(Ib + tl0 * 128 + row) : -1;      // idx rows are the object's edges in order
      for (int tl = tl0; tl < tl1; ++tl, ++tcount) {
        const int e0 = tl * 128, nedge = min(128, E - e0), c_lo = e0 / k, ncen = (e0 + nedge - 1) / k - c_lo + 1;
        const int src = src_next;
        if (tl + 1 < tl1) src_next = e0 + 128 + row < E ? __ldg(Ib + e0 + 128 + row) : -1;
        // centre rows of this tile -> shared memory (double buffered by tile parity)
        float* ccb = cc_s + (tcount & 1) * cmax * C;
        for (int i = rt; i < ncen * (C / 4); i += RT) cp_async16(ccb + 4 * i, Cb + (size_t)c_lo * C + 4 * i);
        cp_async_commit();
        // D1[tcount & 1] was the A operand of GEMM 2 of tile tcount-2 (same barrier slot, previous phase): preload b2
        const uint32_t d1 = tD1 + (uint32_t)((tcount & 1) * C);
        if (tcount >= 2) tc::mbar_wait(&mma2_done[tcount & 1], (uint32_t)((((tcount - 2) >> 1)) & 1));
        tc::tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < CW; c0 += 16) {
          uint32_t r[16];
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 bv = *reinterpret_cast<const float4*>(b2_s + cb + c0 + j);
            r[j] = __float_as_uint(bv.x); r[j + 1] = __float_as_uint(bv.y); r[j + 2] = __float_as_uint(bv.z); r[j + 3] = __float_as_uint(bv.w);
          }
          tc::tmem_st16(d1 + lane_off + cb + c0, r);
        }
        // A1 is free once GEMM 1 of the previous tile has read it
        if (tcount > 0) tc::mbar_wait(&mma1_done[(tcount - 1) & 1], (uint32_t)(((tcount - 1) >> 1) & 1));
        cp_async_wait<0>();
        tc::bar_sync(1, RT);
        tc::tc_fence_after();
        const bool valid = row < nedge;
        const float* crow = ccb + (valid ? (e0 + row) / k - c_lo : 0) * C + cb;
        const float* prow = (P1S ? p1_s + (size_t)max(src, 0) * PSTR : Pb + (size_t)max(src, 0) * C) + cb;
#pragma unroll
        for (int c0 = 0; c0 < CW; c0 += 16) {
          uint32_t r[16];
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) {
              const float4 p = P1S ? *reinterpret_cast<const float4*>(prow + c0 + j) : __ldg(reinterpret_cast<const float4*>(prow + c0 + j));
              const float4 q = *reinterpret_cast<const float4*>(crow + c0 + j);
              v = make_float4(fmaxf(p.x + q.x, 0.f), fmaxf(p.y + q.y, 0.f), fmaxf(p.z + q.z, 0.f), fmaxf(p.w + q.w, 0.f));
            }
            v = tc::tf32_rna4(v);                                        // A operand of GEMM 1: round, do not let the MMA truncate
            r[j] = __float_as_uint(v.x); r[j + 1] = __float_as_uint(v.y); r[j + 2] = __float_as_uint(v.z); r[j + 3] = __float_as_uint(v.w);
          }
          tc::tmem_st16(tA1 + lane_off + cb + c0, r);
        }
        tc::tmem_st_wait();
        tc::tc_fence_before();
        tc::bar_sync(1, RT);
        if (rt == 0) {
          tc::tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < C / 8; ++ks) {
            const uint64_t bd = tc::smem_desc(sW2 + ks * 2 * (C * 16), C * 16, 128, tc::LAYOUT_NONE);
            tc::umma_tf32_ts(d1, tA1 + ks * 8, bd, idesc, 1u);            // accumulates onto the preloaded bias
          }
          tc::umma_commit(&mma1_done[tcount & 1]);
        }
        __syncwarp();
      }
    }
  } else {
    // ================================================= consumers: epilogue 1 (in place), GEMM 2, epilogue 2 = max over k
    for (int u = blockIdx.x; u < a.n_units; u += gridDim.x) {
      const int b = u / a.upo, tl0 = (u % a.upo) * a.tpu, tl1 = min(tl0 + a.tpu, a.tpo);
      float* Ob = a.out + (size_t)b * a.o_bs;
      for (int tl = tl0; tl < tl1; ++tl, ++tcount) {
        const int e0 = tl * 128, nedge = min(128, E - e0), c_lo = e0 / k, ncen = (e0 + nedge - 1) / k - c_lo + 1;
        const bool valid = row < nedge;
        const int cl = valid ? (e0 + row) / k : -1;
        const uint32_t d1 = tD1 + (uint32_t)((tcount & 1) * C);
        tc::mbar_wait(&mma1_done[tcount & 1], (uint32_t)((tcount >> 1) & 1));
        tc::tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < CW; c0 += 16) {
          uint32_t r[16];
          tc::tmem_ld16(d1 + lane_off + cb + c0, r);
          tc::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(tc::tf32_rna(fmaxf(__uint_as_float(r[j]), 0.f)));   // A operand of GEMM 2
          tc::tmem_st16(d1 + lane_off + cb + c0, r);
        }
        tc::tmem_st_wait();
        tc::tc_fence_before();
        tc::bar_sync(2, RT);
        if (t == 0) {
          tc::tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < C / 8; ++ks) {
            const uint64_t bd = tc::smem_desc(sW3 + ks * 2 * (C * 16), C * 16, 128, tc::LAYOUT_NONE);
            tc::umma_tf32_ts(tD2, d1 + ks * 8, bd, idesc, ks > 0 ? 1u : 0u);
          }
          tc::umma_commit(&mma2_done[tcount & 1]);
        }
        __syncwarp();
        // segments (centres) of this warp's 32 rows
        const int cfirst = __shfl_sync(FULL_MASK, cl, 0);                 // rows are ordered: lane 0 has the lowest centre
        uint32_t segmask[MAXSEG];
#pragma unroll
        for (int sgi = 0; sgi < MAXSEG; ++sgi) segmask[sgi] = __ballot_sync(FULL_MASK, valid && cl == cfirst + sgi);
        tc::mbar_wait(&mma2_done[tcount & 1], (uint32_t)((tcount >> 1) & 1));
        tc::tc_fence_after();
        float* pw = part + quad * MAXSEG * C + cb;
#pragma unroll
        for (int c0 = 0; c0 < CW; c0 += 16) {
          uint32_t r[16];
          tc::tmem_ld16(tD2 + lane_off + cb + c0, r);
          tc::tmem_ld_wait();
#pragma unroll
          for (int sgi = 0; sgi < MAXSEG; ++sgi) {
            const uint32_t m = segmask[sgi];
            if (m == 0u) continue;                                      // warp-uniform
            if ((m >> lane) & 1u) {                                     // padding rows never take part
              float mx[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) mx[j] = redux_max(__uint_as_float(r[j]), m);
              if (lane == __ffs(m) - 1) {
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                  *reinterpret_cast<float4*>(pw + sgi * C + c0 + j) = make_float4(mx[j], mx[j + 1], mx[j + 2], mx[j + 3]);
              }
            }
          }
        }
        tc::tc_fence_before();
        tc::bar_sync(2, RT);
        // combine the per-warp partials of each centre (rows [cl*k, (cl+1)*k) -> quadrants w0..w1); bias and ReLU commute
        // with the max over the edges, so they are applied here, once per (centre, channel)
        const float* cin = carry + ((tcount & 1) ^ 1) * C;                // written by the previous tile of this unit
        float* cout = carry + (tcount & 1) * C;
        for (int o = t; o < ncen * C; o += RT) {
          const int c = a.o_cs == 1 ? o % C : o / ncen, ce = c_lo + (a.o_cs == 1 ? o / C : o % ncen);
          const int r0 = max(ce * k, e0) - e0, r1 = min((ce + 1) * k, e0 + nedge) - 1 - e0;      // the centre's rows inside this tile
          float mx = ce * k < e0 ? cin[c] : -INFINITY;                    // its edges in the previous tile
          for (int w = r0 >> 5; w <= (r1 >> 5); ++w) mx = fmaxf(mx, part[(w * MAXSEG + (ce - (e0 + 32 * w) / k)) * C + c]);
          if ((ce + 1) * k > e0 + 128) cout[c] = mx;                      // continues in the next tile (same unit: boundaries are aligned)
          else Ob[(size_t)c * a.o_cs + (size_t)ce * a.o_ss] = fmaxf(mx + b3_s[c], 0.f);
        }
        // the next tile's partials are written only after its own bar_sync(2) in epilogue 1, i.e. after every thread
        // has finished this loop
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 4 * C);
}

template <int C, int CS>
int launch2(int B, int N, int S, int k, const float* P1, const float* Cc, const int* idx, const float* W2img, const float* b2,
            const float* W3img, const float* b3, float* out, int out_pm, int n_sms, cudaStream_t st) {
  Sa2Args a;
  a.N = N; a.S = S; a.k = k; a.cmax = 127 / k + 2; a.tpo = (int)(((long long)S * k + 127) / 128);
  a.P1 = P1; a.Cc = Cc; a.idx = idx; a.W2img = W2img; a.W3img = W3img; a.b2 = b2; a.b3 = b3; a.out = out;
  a.o_bs = (long long)C * S;
  a.o_cs = out_pm ? 1 : S;
  a.o_ss = out_pm ? C : 1;
  if (n_sms <= 0) n_sms = 148;
  const int base = 2 * C * C * 4 + (4 * MAXSEG * C + 2 * a.cmax * C + 4 * C) * 4;
  const int p1_bytes = N * (C + 4) * 4;
  a.p1_smem = base + p1_bytes <= 226 * 1024 ? 1 : 0;
  const int smem = base + (a.p1_smem ? p1_bytes : 0);
  if (smem > 226 * 1024) return PCREID_ERR_UNSUPPORTED;
  int per_sm = (228 * 1024) / (smem + 1024);
  if (per_sm > 512 / (4 * C)) per_sm = 512 / (4 * C);          // tensor-memory columns
  if (per_sm > 2048 / (256 * CS)) per_sm = 2048 / (256 * CS);
  if (per_sm < 1) per_sm = 1;
  const int max_ctas = n_sms * per_sm;
  // whole objects per unit when there are enough of them; otherwise contiguous tile ranges of an object
  a.upo = 1;
  if (B < 2 * max_ctas) {
    a.upo = (2 * max_ctas + B - 1) / B;
    if (a.upo > a.tpo) a.upo = a.tpo;
  }
  a.tpu = (a.tpo + a.upo - 1) / a.upo;
  int gcd = 128, r = k;                                            // a unit starts on a tile that starts a centre: 128 t % k == 0
  while (r) { const int q = gcd % r; gcd = r; r = q; }
  const int period = k / gcd;
  a.tpu = (a.tpu + period - 1) / period * period;
  a.upo = (a.tpo + a.tpu - 1) / a.tpu;
  const long long units = (long long)B * a.upo;
  if (units > 0x7fffffffLL) return PCREID_ERR_UNSUPPORTED;
  a.n_units = (int)units;
  const int grid = units < max_ctas ? (int)units : max_ctas;
  if (a.p1_smem) {
    cudaFuncSetAttribute(sa_edge_mlp_tc2_kernel<C, CS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    sa_edge_mlp_tc2_kernel<C, CS, true><<<grid, 256 * CS, smem, st>>>(a);
  } else {
    cudaFuncSetAttribute(sa_edge_mlp_tc2_kernel<C, CS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    sa_edge_mlp_tc2_kernel<C, CS, false><<<grid, 256 * CS, smem, st>>>(a);
  }
  return pcreid_launch_status();
}

}  // namespace

extern "C" int pcreid_sa_edge_mlp_tc2(int B, int C, int N, int S, int k, const float* P1, const float* Cc, const int* idx,
                                      const float* W2img, const float* b2, const float* W3img, const float* b3, float* out,
                                      int out_pm, int n_sms, void* stream) {
  if (B <= 0 || S <= 0) return PCREID_OK;
  if (!P1 || !Cc || !idx || !W2img || !b2 || !W3img || !b3 || !out || k <= 0 || N <= 0) return PCREID_ERR_ARG;
  if (k > 128 || k < 16) return PCREID_ERR_UNSUPPORTED;        // a warp's 32 rows may span at most 3 centres
  cudaStream_t st = (cudaStream_t)stream;
  switch (C) {
    case 32: return launch2<32, 1>(B, N, S, k, P1, Cc, idx, W2img, b2, W3img, b3, out, out_pm, n_sms, st);
    case 64: return launch2<64, 2>(B, N, S, k, P1, Cc, idx, W2img, b2, W3img, b3, out, out_pm, n_sms, st);
    case 128: return launch2<128, 2>(B, N, S, k, P1, Cc, idx, W2img, b2, W3img, b3, out, out_pm, n_sms, st);
    default: return PCREID_ERR_UNSUPPORTED;
  }
}
