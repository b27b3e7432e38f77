// Shared pieces of the fused tensor-core matcher kernels (pair_tc.cu, pair_tc2.cu): operand-image geometry, kernel
// argument blocks, descriptor helpers, the 4-warp group abstraction.  Everything lives in an anonymous namespace: each
// translation unit gets its own copy.
#pragma once
#include "../../include/pcreid.h"
#include "common.cuh"
#include "tc_common.cuh"

namespace {


constexpr int IMG = 16384;              // bytes of a 128 x 64 bf16 operand image  [k/8][row][8]
// Attention operand of a template (stage 1: MK1 per object, stage 2: B7 per (pair, direction)), one image PER HEAD:
//   head h: [n/8 (10 chunks)][k - 32 h (32 rows)][8]   columns 0..63 = (blockdiag(KV) Wm^T)[k][:], column 64 = Ksum[k], 65..79 = 0
// Head h's queries (A K-steps 2h, 2h+1) meet only their own 32 k-rows: two N = 80, K = 32 GEMMs into accumulator columns
// [0, 80) and [80, 160) instead of one N = 144, K = 64 GEMM against an image that was half structural zeros (18 KB -> 10 KB
// per operand in HBM and shared memory, 18.4 -> 10 KB of tensor-core operand reads per tile; the sums are bit-identical).
constexpr int NB7H = 80;                // columns per head: 64 merged output channels | Q.Ksum dot | pad
constexpr int B7_HEAD = (NB7H / 8) * 32 * 16;   // 5120 B per head
constexpr int B7_BYTES = 2 * B7_HEAD;           // 10240
constexpr int ONES_BYTES = 2 * 2048;    // two extra 8-column chunks appended to V: column 64 == 1 (Ksum), rest 0
constexpr float LN_EPS = 1e-5f;

struct P1Args {
  int n_units, NT, role;
  int npts;                              // true points per object (<= 128 NT; rows beyond it are zero padding)
  float att_eps;                         // LinearAttention eps times the scale the key/value sums were stored with
  float kv_scale;                        // phase 1b: scale applied to KV / Ksum before they become the stage-2 operand
  const int *u_search, *u_templ, *u_slot;
  const uint8_t *QF1, *H, *PV;          // search-side per-object images: [obj][NT][IMG]
  const uint8_t* MK1;                    // template-side per-object stage-1 attention operand [obj][B7_BYTES]
  const uint8_t* W;                      // P1 weights blob
  uint8_t* A_out;                        // [slot][2][NT][IMG]   stage-1 outputs (bf16 operand images)
  uint8_t* B7_out;                       // [slot][2][B7_BYTES]  stage-2 attention operands
};
struct P2Args {
  int n_units, NT, role;
  int npts;
  float att_eps;
  const int* u_slot;
  const uint8_t* A_in;                   // == A_out of phase 1
  const uint8_t* B7_in;                  // == B7_out of phase 1
  const uint8_t* W;                      // P2 weights blob
  float* pool_part;                      // [slot][2][128]: max (64) | sum (64) over the search object's points
};

__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x + 1.f : __expf(x); }

// 16-byte read-only global load that the compiler may NOT sink to its first use (plain __ldg of data consumed a whole
// stage later was being moved next to the consumer, which exposed the full L2 latency there)
__device__ __forceinline__ uint4 ldg_early(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// one 128-byte line per thread pulled into L2 (no registers, no shared memory): used a whole tile ahead of the cp.async
// that needs the data, so that the later copy pays an L2 hit instead of a DRAM round trip
__device__ __forceinline__ void prefetch_l2_16k(const uint8_t* img, int t) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(img + (size_t)t * 128));
}

__device__ __forceinline__ void copy_to_smem(uint8_t* dst, const uint8_t* __restrict__ src, int bytes, int t, int nthr) {
  for (int i = t * 16; i < bytes; i += nthr * 16) cp_async16(dst + i, src + i);
}

// Descriptors are built once per operand buffer; stepping along K only adds (bytes >> 4) to the 14-bit start-address
// field (shared memory is < 256 KB, so the field never carries).  Building them per MMA cost ~130 cycles of dependent
// 64-bit arithmetic on the single issuing thread (measured with the cycle trace) -- a quarter of the tile time.
struct Opnd {
  uint64_t desc;     // descriptor at k = 0
  uint32_t kstep;    // (bytes per K=16 step) >> 4
};
__device__ __forceinline__ Opnd opnd(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t kstep_bytes) {
  Opnd o;
  o.desc = tc::smem_desc(addr, lbo, sbo, tc::LAYOUT_NONE);
  o.kstep = kstep_bytes >> 4;
  return o;
}
// operand geometry (bytes): K-major activation image (128 rows), K-major weight image (N rows), MN-major attention operand
#define A_IMG(addr) opnd((addr), 2048u, 128u, 4096u)
#define W_IMG(addr, N) opnd((addr), (uint32_t)((N)*16), 128u, (uint32_t)(2 * (N)*16))
#define B7_IMG(addr) opnd((addr), 128u, 512u, 256u)      /* one head: MN-major, 32 k-rows per 8-column chunk */

// one thread: D[tmem_d] (+)= A x B^T over KSTEPS K=16 steps
template <int KSTEPS>
__device__ __forceinline__ void issue_gemm(uint32_t tmem_d, const Opnd& A, const Opnd& B, uint32_t idesc, bool accumulate) {
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks)
    tc::umma_f16(tmem_d, A.desc + (uint64_t)(ks * A.kstep), B.desc + (uint64_t)(ks * B.kstep), idesc, (accumulate || ks > 0) ? 1u : 0u);
}

// 8 x 16 B of a side image row (chunks c0 .. c0+7) -> registers (issued early so the L2 latency hides behind a stage)
template <int NCH>
__device__ __forceinline__ void load_side(uint4 (&sd)[NCH], const uint8_t* __restrict__ img, int chunk0, int row) {
#pragma unroll
  for (int c = 0; c < NCH; ++c) sd[c] = __ldg(reinterpret_cast<const uint4*>(img + (chunk0 + c) * 2048 + row * 16));
}

// chunks [c_lo, c_lo + 4) of row d (= k) of the stage operand B7 / MK1: M32 holds M[d][8*c_lo .. 8*c_lo + 32); the row lives in
// the image of head d >> 5 only
template <class F>
__device__ __forceinline__ void write_b7_part(const float (&M32)[32], int c_lo, float ksum, bool tail, int d, uint8_t* dst) {
  uint8_t* base = dst + (d >> 5) * B7_HEAD + (d & 31) * 16;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = F::pack(M32[c * 8 + 2 * j], M32[c * 8 + 2 * j + 1]);
    *reinterpret_cast<uint4*>(base + (c_lo + c) * 512) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  if (tail) {
    *reinterpret_cast<uint4*>(base + 8 * 512) = make_uint4(F::pack(ksum, 0.f), 0, 0, 0);
    *reinterpret_cast<uint4*>(base + 9 * 512) = make_uint4(0, 0, 0, 0);
  }
}

constexpr int GX = 128, NGX = 3;     // three-tile variants: 3 groups of 4 warps per CTA, one thread per tile row
struct GroupX {
  int t, gid;
  bool issuer;
  uint32_t tmem, tlane;
  uint64_t* bar;
  uint32_t par;
  __device__ __forceinline__ void sync() { tc::bar_sync(1 + gid, GX); }
  __device__ __forceinline__ void publish() { tc::fence_async_smem(); tc::tc_fence_before(); sync(); tc::tc_fence_after(); }
  __device__ __forceinline__ void wait() { tc::mbar_wait(bar, par); par ^= 1u; tc::tc_fence_after(); }
};

// 32 accumulator columns -> fp32 registers, with running sum / sum of squares
__device__ __forceinline__ void ld32_stats(uint32_t taddr, float (&o)[32], float& s, float& ss) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t r[16];
    tc::tmem_ld16(taddr + 16 * half, r);
    tc::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float v = __uint_as_float(r[j]);
      o[16 * half + j] = v;
      s += v;
      ss = fmaf(v, v, ss);
    }
  }
}

// row d of the stage operand (used by the per-object packer): all 64 columns at once
template <class F>
__device__ __forceinline__ void write_b7_row(const float (&M)[64], float ksum, int d, uint8_t* dst) {
  float lo[32], hi[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) { lo[j] = M[j]; hi[j] = M[32 + j]; }
  write_b7_part<F>(lo, 0, ksum, true, d, dst);
  write_b7_part<F>(hi, 4, ksum, false, d, dst);
}

__device__ __forceinline__ void groupx_setup(GroupX& g, uint64_t* bars, uint32_t tmem_base) {
  const int warp_u = (int)tc::uniform(threadIdx.x >> 5);
  g.gid = warp_u / 4;
  g.t = threadIdx.x % GX;
  g.issuer = (warp_u % 4) == 0;
  g.tmem = tc::uniform(tmem_base) + g.gid * 160;
  g.tlane = g.tmem + ((uint32_t)((warp_u % 4) * 32) << 16);
  g.bar = bars + g.gid;
  g.par = 0;
}

}  // namespace
