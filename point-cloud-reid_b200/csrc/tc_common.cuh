// tcgen05 / TMEM / mbarrier primitives for sm_100a (inline PTX; no CUTLASS dependency).
// Encodings follow the PTX ISA tables as mirrored in cute/arch/mma_sm100_desc.hpp (SmemDescriptor,
// InstrDescriptor) of the CUTLASS header tree vendored in this image.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  const uint32_t a = smem_u32(bar);
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!ok);
}

// ---- fences ---------------------------------------------------------------------------------
// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM allocation (one full warp executes these) --------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- descriptors ----------------------------------------------------------------------------
enum { LAYOUT_NONE = 0, LAYOUT_SW128_BASE32B = 1, LAYOUT_SW128 = 2, LAYOUT_SW64 = 4, LAYOUT_SW32 = 6 };
// shared-memory matrix descriptor: start address, leading / stride byte offsets (all >> 4), version 1 (Blackwell)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout & 7) << 61;
  return d;
}
enum { FMT_F16 = 0, FMT_BF16 = 1, FMT_TF32 = 2 };
enum { MAJOR_K = 0, MAJOR_MN = 1 };
// instruction descriptor for kind::f16 / kind::tf32 with fp32 accumulation
__host__ __device__ constexpr uint32_t instr_desc(int M, int N, int fmt, int a_major, int b_major) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)a_major << 15) | ((uint32_t)b_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- MMA issue (ONE thread) --------------------------------------------------------------------
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from tensor memory
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMEM <-> registers: 32 lanes x 32 bit, x8 / x16 / x32 columns per thread ------------------------
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// 32 columns as four back-to-back x8 loads: measured 23.6 cycles (4 x x8 + wait) against 53 cycles for one .x32 + wait
// (scripts/ubench/tmem_bw.cu), i.e. the wide shapes are not faster than the narrow ones issued in a row.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
#pragma unroll
  for (int q = 0; q < 4; ++q)
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[8 * q]), "=r"(r[8 * q + 1]), "=r"(r[8 * q + 2]), "=r"(r[8 * q + 3]), "=r"(r[8 * q + 4]), "=r"(r[8 * q + 5]),
                   "=r"(r[8 * q + 6]), "=r"(r[8 * q + 7])
                 : "r"(taddr + 8 * q));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
               "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
               : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);   // .x = lo (low 16 bits), .y = hi
  return *reinterpret_cast<uint32_t*>(&v);
}

// ---- packed bf16x2 arithmetic for epilogues whose result is rounded to bf16 anyway (2 elements per instruction) ----
__device__ __forceinline__ uint32_t bf2_add(uint32_t a, uint32_t b) { uint32_t r; asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t bf2_mul(uint32_t a, uint32_t b) { uint32_t r; asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t bf2_max(uint32_t a, uint32_t b) { uint32_t r; asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t bf2_min(uint32_t a, uint32_t b) { uint32_t r; asm("min.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t bf2_ex2(uint32_t a) { uint32_t r; asm("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(r) : "r"(a)); return r; }
// elu(x)+1 = max(x,0) + exp(min(x,0)) on a packed pair (one MUFU for two elements)
__device__ __forceinline__ uint32_t bf2_elu1(uint32_t x) {
  const uint32_t LOG2E = 0x3fb93fb9u;   // bf16(1.4427) in both halves
  return bf2_add(bf2_max(x, 0u), bf2_ex2(bf2_mul(bf2_min(x, 0u), LOG2E)));
}

__device__ __forceinline__ uint32_t bf2_fma(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t bf2_fma_relu(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm("fma.rn.relu.bf16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
// fp32 pair -> relu -> packed bf16x2 in ONE instruction (F2FP.RELU)
__device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi) { uint32_t r; asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
// elu(x)+1 on a packed pair whose producer weights were pre-scaled by 1/bf16(ln 2): x' = x / LN2B, so
//   max(x',0) * LN2B + 2^min(x',0) = max(x,0) + exp(min(x,0) * ln2/LN2B)       (4 instructions, one MUFU, per two elements)
__device__ __forceinline__ uint32_t bf2_elu1s(uint32_t x) {
  const uint32_t LN2B = 0x3f313f31u;    // bf16(0.69140625) in both halves
  return bf2_fma(bf2_max(x, 0u), LN2B, bf2_ex2(bf2_min(x, 0u)));
}

// ---- operand element formats of the kind::f16 MMAs ----------------------------------------------------------------
// The fused matcher kernels are templated on one of these.  Both are 16-bit operands with fp32 accumulation, i.e. the same
// tensor-pipe rate and the same operand bytes; they differ in where the 16 bits go:
//   OpBF16: 8-bit significand (rel. rounding error 2^-9), fp32 exponent range           -> "fast" mode, |dlogit| <= 3e-2
//   OpF16 : 11-bit significand (rel. rounding error 2^-12, the significand of tf32), 5-bit exponent -> "parity_tc" mode.
//           Every operand of the matcher is O(1) by construction (LayerNorm outputs, elu+1 features, key/value sums
//           pre-scaled by 1/points); conversions saturate to +-65504 instead of producing inf.
// Epilogue arithmetic whose result is rounded to the operand format anyway runs packed (2 elements per instruction) for
// OpBF16; OpF16 keeps it in fp32 and rounds ONCE (the parity mode pays ~10 % more epilogue instructions for that).
struct OpBF16 {
  static constexpr int FMT = FMT_BF16;
  static constexpr uint32_t ONE_LO = 0x00003f80u;     // packed pair (1.0, 0.0)
  static constexpr bool PACKED_MATH = true;
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) { return pack_bf16(lo, hi); }
  static __device__ __forceinline__ uint32_t pack_relu(float lo, float hi) { return pack_bf16_relu(lo, hi); }
  static __device__ __forceinline__ float lo(uint32_t w) { return __uint_as_float(w << 16); }
  static __device__ __forceinline__ float hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
  // relu(acc + side) -> packed
  static __device__ __forceinline__ uint32_t add_relu(float a, float b, uint32_t side) {
    return bf2_fma_relu(pack_bf16(a, b), 0x3f803f80u, side);
  }
  // acc + side -> packed
  static __device__ __forceinline__ uint32_t add(float a, float b, uint32_t side) { return bf2_add(pack_bf16(a, b), side); }
  // elu(x)+1 of accumulators produced by weights pre-scaled by 1/bf16(ln 2) -> packed
  static __device__ __forceinline__ uint32_t elu1_scaled(float a, float b) { return bf2_elu1s(pack_bf16(a, b)); }
  // elu(x)+1 of plain accumulators -> packed
  static __device__ __forceinline__ uint32_t elu1(float a, float b) { return bf2_elu1(pack_bf16(a, b)); }
};
// packed f16x2 arithmetic (same instruction shapes as the bf16x2 helpers above)
__device__ __forceinline__ uint32_t h2_add(uint32_t a, uint32_t b) { uint32_t r; asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t h2_mul(uint32_t a, uint32_t b) { uint32_t r; asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t h2_max(uint32_t a, uint32_t b) { uint32_t r; asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t h2_min(uint32_t a, uint32_t b) { uint32_t r; asm("min.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t h2_ex2(uint32_t a) { uint32_t r; asm("ex2.approx.f16x2 %0, %1;" : "=r"(r) : "r"(a)); return r; }
__device__ __forceinline__ uint32_t h2_fma(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }

struct OpF16 {
  static constexpr int FMT = FMT_F16;
  static constexpr uint32_t ONE_LO = 0x00003c00u;
#ifdef PCREID_F16_PACKED
  static constexpr bool PACKED_MATH = true;
#else
  static constexpr bool PACKED_MATH = false;
#endif
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
    uint32_t r; asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r;
  }
  static __device__ __forceinline__ uint32_t pack_relu(float lo, float hi) {
    uint32_t r; asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r;
  }
  static __device__ __forceinline__ float lo(uint32_t w) {
    float f; asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %1;\n\tcvt.f32.f16 %0, l;\n\t}\n" : "=f"(f) : "r"(w)); return f;
  }
  static __device__ __forceinline__ float hi(uint32_t w) {
    float f; asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %1;\n\tcvt.f32.f16 %0, h;\n\t}\n" : "=f"(f) : "r"(w)); return f;
  }
  static __device__ __forceinline__ uint32_t add_relu(float a, float b, uint32_t side) { return pack_relu(a + lo(side), b + hi(side)); }
#ifdef PCREID_F16_PACKED
  static __device__ __forceinline__ uint32_t add(float a, float b, uint32_t side) { return h2_add(pack(a, b), side); }
#else
  static __device__ __forceinline__ uint32_t add(float a, float b, uint32_t side) { return pack(a + lo(side), b + hi(side)); }
#endif
  // weights pre-scaled by 1/ln 2 (fp32-exact constant here: the multiply happens in fp32)
  static __device__ __forceinline__ float elu1s_f(float x) {
    float e; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(x, 0.f)));
    return fmaf(fmaxf(x, 0.f), 0.69314718056f, e);
  }
#ifdef PCREID_F16_PACKED
  // weights pre-scaled by 1/f16(ln 2) on the host: max(x',0) * LN2H + 2^min(x',0)
  static __device__ __forceinline__ uint32_t elu1_scaled(float a, float b) {
    const uint32_t x = pack(a, b);
    return h2_fma(h2_max(x, 0u), 0x398c398cu, h2_ex2(h2_min(x, 0u)));
  }
#else
  static __device__ __forceinline__ uint32_t elu1_scaled(float a, float b) { return pack(elu1s_f(a), elu1s_f(b)); }
#endif
  // elu(x)+1 = max(x,0) + exp(min(x,0)), branch-free: 5 instructions per element (the select form `x > 0 ? x + 1 : __expf(x)`
  // compiled to divergent regions around a non-ftz exp with denormal scaling: 12+ instructions and BSSY/BSYNC pairs)
  static __device__ __forceinline__ float elu1_f(float x) {
    float e; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(x, 0.f) * 1.4426950408889634f));
    return fmaxf(x, 0.f) + e;
  }
#ifdef PCREID_F16_PACKED
  static __device__ __forceinline__ uint32_t elu1(float a, float b) {
    const uint32_t x = pack(a, b);
    return h2_add(h2_max(x, 0u), h2_ex2(h2_mul(h2_min(x, 0u), 0x3dc53dc5u)));
  }
#else
  static __device__ __forceinline__ uint32_t elu1(float a, float b) { return pack(elu1_f(a), elu1_f(b)); }
#endif
};

// fp32 -> tf32 with round-to-nearest (ties away), returned in an fp32 container.  tcgen05.mma kind::tf32 reads 32-bit operands
// and IGNORES the low 13 mantissa bits (truncation: error up to 2^-10 relative, biased towards zero); operands that are
// pre-rounded here carry half that error and no bias.  Used by the encoder kernels for every activation they write as a
// pure MMA operand (the weights are pre-rounded on the host: kernels.py tf32_image).
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ float4 tf32_rna4(float4 v) { return make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w)); }

// warp-uniform helpers: values produced through these are known to be warp-uniform by the compiler, so descriptor
// arithmetic and tcgen05.mma operands stay on the uniform datapath (no per-MMA R2UR moves)
__device__ __forceinline__ uint32_t uniform(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
  return pred != 0;
}

// named barrier for a sub-group of the CTA (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

}  // namespace tc
