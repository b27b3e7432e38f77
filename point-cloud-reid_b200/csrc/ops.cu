// Point ops for sm_100a: kNN, furthest point sampling, ball query, group / gather.
//
// These replace the reference's scalar SIMT kernels (one thread per query, heap in local memory,
// FPS distances in global memory) with warp-cooperative kernels:
//   knn          one warp per query, distances staged in shared memory as order-preserving keys,
//                k rounds of warp arg-min (two REDUX per round), canonical (d, idx) ascending order;
//                exact-tie queries of the mmdet3d op fall back to an in-warp emulation of the
//                reference heap so indices stay bit-identical   (ref: ops/knn/src/knn_cuda.cu:58-94)
//   fps          one warp / CTA per object, coordinates in shared memory in tie-priority order, running
//                min-distances in registers, arg-max = REDUX.MAX over value bits + REDUX.MIN over slots
//                (fps_rank_kernel; fps_kernel keeps everything in registers for the with-dist form)
//                                                  (ref: ops/furthest_point_sample/src/furthest_point_sample_cuda.cu:25-141)
//   ball_query   one warp per query, ballot + prefix popcount keeps index order
//                                                  (ref: ops/ball_query/src/ball_query_cuda.cu:11-54)
//   group/gather indices read once per position (not once per channel), 128-bit streaming stores
//                                                  (ref: ops/group_points/src/group_points_cuda.cu:56-79,
//                                                        ops/gather_points/src/gather_points_cuda.cu:8-26)
#include "common.cuh"
#include <math.h>
#include <stdlib.h>
#include <string.h>

// ------------------------------------------------------------------------------------------------
// distance arithmetic (must match the reference bit for bit; never let nvcc re-contract)
// ------------------------------------------------------------------------------------------------
// mode 0: mmdet3d CUDA ops.  nvcc -fmad=true contracts dx*dx+dy*dy+dz*dz to fma(dz,dz,fma(dx,dx,dy*dy)).
__device__ __forceinline__ float dist_direct(float ax, float ay, float az, float bx, float by, float bz) {
  float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  float t = __fmul_rn(dy, dy);
  t = __fmaf_rn(dx, dx, t);
  return __fmaf_rn(dz, dz, t);
}
// torch-path FPS (models/pointnet2_utils.py:116-137): dist = torch.sum((xyz - centroid) ** 2, -1) = (dx*dx + dy*dy) + dz*dz,
// separate multiplies and adds (no contraction)
__device__ __forceinline__ float dist_sum_sq(float ax, float ay, float az, float bx, float by, float bz) {
  float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}
// mode 1: torch path square_distance (models/pointnet2_utils.py:169-188):
//   m = q.p as a sequential fma chain, |v|^2 = (x*x + y*y) + z*z, d = (-2*m + |q|^2) + |p|^2
__device__ __forceinline__ float sqnorm3(float x, float y, float z) {
  return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}
__device__ __forceinline__ float dist_expand(float qx, float qy, float qz, float qn, float px, float py, float pz) {
  float m = __fmul_rn(qx, px);
  m = __fmaf_rn(qy, py, m);
  m = __fmaf_rn(qz, pz, m);
  float pn = sqnorm3(px, py, pz);
  return __fadd_rn(__fadd_rn(__fmul_rn(-2.f, m), qn), pn);
}

// ------------------------------------------------------------------------------------------------
// kNN
// ------------------------------------------------------------------------------------------------
__device__ void heap_reheap(float* dist, int* idx, int k) {
  int root = 0, child = 1;
  while (child < k) {
    if (child + 1 < k && dist[child + 1] > dist[child]) child++;
    if (dist[root] > dist[child]) return;
    float tf = dist[root]; dist[root] = dist[child]; dist[child] = tf;
    int ti = idx[root]; idx[root] = idx[child]; idx[child] = ti;
    root = child;
    child = root * 2 + 1;
  }
}

// the reference's sequential max-heap over all candidates + heap_sort (knn_cuda.cu:26-94), run by ONE lane for the rare
// queries with an exact tie among the selected k / at the k-th boundary (which tied candidates survive, and in which
// order, depends on the heap's internal structure).
template <int MODE>
__device__ void heap_replay(const float* __restrict__ P, int N, int k, float qx, float qy, float qz, float qn, float* hd, int* hi,
                            int* oi, float* od, long long sk) {
  for (int i = 0; i < k; ++i) { hd[i] = 1e10f; hi[i] = 0; }
  for (int i = 0; i < N; ++i) {
    float px = __ldg(P + i * 3), py = __ldg(P + i * 3 + 1), pz = __ldg(P + i * 3 + 2);
    float d = MODE == 0 ? dist_direct(qx, qy, qz, px, py, pz) : dist_expand(qx, qy, qz, qn, px, py, pz);
    if (d < hd[0]) { hd[0] = d; hi[0] = i; heap_reheap(hd, hi, k); }
  }
  for (int i = k - 1; i > 0; --i) {   // heap_sort (knn_cuda.cu:45-54)
    float tf = hd[0]; hd[0] = hd[i]; hd[i] = tf;
    int ti = hi[0]; hi[0] = hi[i]; hi[i] = ti;
    heap_reheap(hd, hi, i);
  }
  for (int i = 0; i < k; ++i) {
    oi[(size_t)i * sk] = hi[i];
    if (od) od[(size_t)i * sk] = hd[i];
  }
}

// MODE 0: direct-form distance, MODE 1: expansion-form distance.
// idx/dist2 element (b, q, j) lives at  b*M*k + q*sq + j*sk   (sq=k, sk=1 -> [B,M,k]; sq=1, sk=M -> [B,k,M]).
template <int MODE>
__global__ void __launch_bounds__(256) knn_kernel(int N, int M, int k, const float* __restrict__ xyz,
                                                  const float* __restrict__ qxyz, int* __restrict__ idx,
                                                  float* __restrict__ dist2, long long sq, long long sk,
                                                  int heap_ties, int npad) {
  extern __shared__ uint32_t smem_u32[];
  const int warps = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * warps + warp;
  const int b = blockIdx.y;
  if (q >= M) return;   // warp-uniform; only warp-level sync below

  uint32_t* keys = smem_u32 + (size_t)warp * npad;
  const float* P = xyz + (size_t)b * N * 3;
  const float* Q = qxyz + ((size_t)b * M + q) * 3;
  const float qx = Q[0], qy = Q[1], qz = Q[2];
  const float qn = sqnorm3(qx, qy, qz);
  int* oi = idx + (size_t)b * M * k + (size_t)q * sq;
  float* od = dist2 ? dist2 + (size_t)b * M * k + (size_t)q * sq : nullptr;

  uint32_t lmin = 0xffffffffu;
  int lidx = lane;
  for (int i = lane; i < N; i += 32) {
    float px = __ldg(P + i * 3), py = __ldg(P + i * 3 + 1), pz = __ldg(P + i * 3 + 2);
    float d = MODE == 0 ? dist_direct(qx, qy, qz, px, py, pz) : dist_expand(qx, qy, qz, qn, px, py, pz);
    uint32_t key = f32_to_ordered(d);
    keys[i] = key;
    if (key < lmin) { lmin = key; lidx = i; }
  }
  __syncwarp();

  const int nsel = k < N ? k : N;
  const int rounds = (heap_ties && N > k) ? k + 1 : nsel;
  bool slow = false;
  uint32_t prev = 0;
  for (int r = 0; r < rounds; ++r) {
    uint32_t dmin = warp_min_u32(lmin);
    uint32_t win = warp_min_u32(lmin == dmin ? (uint32_t)lidx : 0xffffffffu);
    if (r > 0 && dmin == prev) slow = true;
    prev = dmin;
    if (r < nsel) {
      float dv = ordered_to_f32(dmin);
      if (heap_ties && !(dv < 1e10f)) slow = true;   // reference heap never admits d2 >= 1e10
      if (lane == 0) {
        oi[(size_t)r * sk] = (int)win;
        if (od) od[(size_t)r * sk] = dv;
      }
    }
    if (lane == (int)(win & 31u)) {
      keys[win] = 0xffffffffu;
      lmin = 0xffffffffu;
      lidx = lane;
      for (int i = lane; i < N; i += 32) {
        uint32_t key = keys[i];
        if (key < lmin) { lmin = key; lidx = i; }
      }
    }
    __syncwarp();
  }
  if (heap_ties) {
    // unfilled slots of the reference heap: (idx 0, dist 1e10) at the tail (knn_cuda.cu:74-77)
    for (int r = nsel + lane; r < k; r += 32) {
      oi[(size_t)r * sk] = 0;
      if (od) od[(size_t)r * sk] = 1e10f;
    }
    __syncwarp();
    if (slow) {
      // exact tie among the selected k / at the k-th boundary: which tied candidates survive and their
      // order depend on the heap's internal structure -> replay the reference heap (one lane).
      float* hd = reinterpret_cast<float*>(smem_u32 + (size_t)warps * npad) + (size_t)warp * 2 * k;
      int* hi = reinterpret_cast<int*>(hd + k);
      if (lane == 0) heap_replay<MODE>(P, N, k, qx, qy, qz, qn, hd, hi, oi, od, sk);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Ordered kNN by SELECT + SORT (N <= 1024, k <= min(N, 128)): instead of k rounds of warp arg-min over all N keys, the
// k-th smallest key is found by the early-exit radix search of knn_set_kernel (below), the k members are collected by
// ballot / prefix popcount into a warp-private list of (key << 32 | index) words, and the list is rank-sorted (k words,
// each lane counts the smaller words for its own) into the canonical ascending (distance, index) order -- the same output
// as knn_kernel, bit for bit.  heap_ties (the mmdet3d op): a query with equal keys among its k members, a tie at the k-th
// boundary or a distance >= 1e10 is replayed through the reference heap exactly as in knn_kernel.
// ------------------------------------------------------------------------------------------------
template <int MODE, int KPL, bool STAGED>
__global__ void __launch_bounds__(256) knn_sel_kernel(int N, int M, int k, int qpw, const float* __restrict__ xyz,
                                                      const float* __restrict__ qxyz, int* __restrict__ idx,
                                                      float* __restrict__ dist2, long long sq, long long sk, int heap_ties) {
  extern __shared__ unsigned long long sel_smem[];            // [8][kpad] member words, then [8][2k] heap scratch
  __shared__ float4 cand[STAGED ? KPL * 32 : 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int q0 = (blockIdx.x * 8 + warp) * qpw;
  const int kpad = (k + 31) & ~31;
  const float* P = xyz + (size_t)b * N * 3;
  if (STAGED) {
    for (int i = threadIdx.x; i < N; i += 256) cand[i] = make_float4(__ldg(P + i * 3), __ldg(P + i * 3 + 1), __ldg(P + i * 3 + 2), 0.f);
    __syncthreads();
  }
  if (q0 >= M) return;
  unsigned long long* sel = sel_smem + (size_t)warp * kpad;
  float px[STAGED ? 1 : KPL], py[STAGED ? 1 : KPL], pz[STAGED ? 1 : KPL];
  if (!STAGED) {
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
      const int i = lane + 32 * j;
      px[j] = py[j] = pz[j] = 0.f;
      if (i < N) { px[j] = __ldg(P + i * 3); py[j] = __ldg(P + i * 3 + 1); pz[j] = __ldg(P + i * 3 + 2); }
    }
  }
  const uint32_t lt_mask = (1u << lane) - 1u;
  const int qend = min(q0 + qpw, M);
  for (int q = q0; q < qend; ++q) {
    const float* Q = qxyz + ((size_t)b * M + q) * 3;
    const float qx = __ldg(Q), qy = __ldg(Q + 1), qz = __ldg(Q + 2);
    const float qn = sqnorm3(qx, qy, qz);
    uint32_t key[KPL];
    uint32_t lmin = 0xffffffffu, lmax = 0u;
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
      key[j] = 0xffffffffu;
      if (lane + 32 * j < N) {
        float cx, cy, cz;
        if (STAGED) { const float4 c = cand[lane + 32 * j]; cx = c.x; cy = c.y; cz = c.z; }
        else { cx = px[j]; cy = py[j]; cz = pz[j]; }
        key[j] = f32_to_ordered(MODE == 0 ? dist_direct(qx, qy, qz, cx, cy, cz) : dist_expand(qx, qy, qz, qn, cx, cy, cz));
        lmin = min(lmin, key[j]);
        lmax = max(lmax, key[j]);
      }
    }
    const uint32_t kmin = __reduce_min_sync(FULL_MASK, lmin), kmax = __reduce_max_sync(FULL_MASK, lmax);
    uint32_t T = kmin;
    bool exact = false;
    if (kmin != kmax) {
      const int top = 31 - __clz(kmin ^ kmax);
      T = top == 31 ? 0u : (kmin & ~((2u << top) - 1u));
      for (int bit = top; bit >= 0; --bit) {
        const uint32_t t = T | (1u << bit);
        int c = 0;
#pragma unroll
        for (int j = 0; j < KPL; ++j) c += key[j] < t ? 1 : 0;
        c = __reduce_add_sync(FULL_MASK, c);
        if (c == k) { T = t; exact = true; break; }
        if (c < k) T = t;
      }
    }
    int base = 0;
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
      const bool in = key[j] < T;
      const uint32_t bal = __ballot_sync(FULL_MASK, in);
      if (in) sel[base + __popc(bal & lt_mask)] = ((unsigned long long)key[j] << 32) | (uint32_t)(lane + 32 * j);
      base += __popc(bal);
    }
    if (!exact) {     // keys equal to the k-th smallest, in ascending index order, until k members
#pragma unroll
      for (int j = 0; j < KPL; ++j) {
        const bool eq = key[j] == T && lane + 32 * j < N;
        const uint32_t bal = __ballot_sync(FULL_MASK, eq);
        const int pos = base + __popc(bal & lt_mask);
        if (eq && pos < k) sel[pos] = ((unsigned long long)key[j] << 32) | (uint32_t)(lane + 32 * j);
        base += __popc(bal);
      }
    }
    const bool boundary_tie = base > k;          // more keys equal to the k-th one than slots left for them
    __syncwarp();
    int* oi = idx + (size_t)b * M * k + (size_t)q * sq;
    float* od = dist2 ? dist2 + (size_t)b * M * k + (size_t)q * sq : nullptr;
    bool bad = false;
    for (int e = lane; e < k; e += 32) {
      const unsigned long long v = sel[e];
      int rank = 0;
      bool dup = false;
      for (int j = 0; j < k; ++j) {
        const unsigned long long w = sel[j];
        rank += w < v ? 1 : 0;
        dup |= (j != e) && ((uint32_t)(w >> 32) == (uint32_t)(v >> 32));
      }
      const float dv = ordered_to_f32((uint32_t)(v >> 32));
      bad |= dup || !(dv < 1e10f);
      oi[(size_t)rank * sk] = (int)(uint32_t)v;
      if (od) od[(size_t)rank * sk] = dv;
    }
    if (heap_ties) {
      const bool slow = __any_sync(FULL_MASK, bad) || boundary_tie;
      __syncwarp();
      if (slow && lane == 0) {
        float* hd = reinterpret_cast<float*>(sel_smem + (size_t)8 * kpad) + (size_t)warp * 2 * k;
        heap_replay<MODE>(P, N, k, qx, qy, qz, qn, hd, reinterpret_cast<int*>(hd + k), oi, od, sk);
      }
    }
    __syncwarp();      // the member list is reused by the next query
  }
}

template <int MODE, int KPL, bool STAGED>
static int knn_sel_launch(int b, int n, int m, int k, const float* xyz, const float* q, int* idx, float* dist2, long long sq,
                          long long sk, int heap_ties, cudaStream_t st) {
  const int qpw = m >= 64 ? 8 : (m >= 16 ? 2 : 1);
  const int kpad = (k + 31) & ~31;
  const size_t smem = (size_t)8 * kpad * 8 + (heap_ties ? (size_t)8 * 2 * k * 4 : 0);
  knn_sel_kernel<MODE, KPL, STAGED><<<dim3(ceil_div(m, 8 * qpw), b), 256, smem, st>>>(n, m, k, qpw, xyz, q, idx, dist2, sq, sk,
                                                                                      heap_ties);
  return pcreid_launch_status();
}

template <int MODE>
static int knn_sel_dispatch(int b, int n, int m, int k, const float* xyz, const float* q, int* idx, float* dist2, long long sq,
                            long long sk, int heap_ties, cudaStream_t st) {
  if (n <= 128) return knn_sel_launch<MODE, 4, false>(b, n, m, k, xyz, q, idx, dist2, sq, sk, heap_ties, st);
  if (n <= 256) return knn_sel_launch<MODE, 8, false>(b, n, m, k, xyz, q, idx, dist2, sq, sk, heap_ties, st);
  if (n <= 512) return knn_sel_launch<MODE, 16, true>(b, n, m, k, xyz, q, idx, dist2, sq, sk, heap_ties, st);
  return knn_sel_launch<MODE, 32, true>(b, n, m, k, xyz, q, idx, dist2, sq, sk, heap_ties, st);
}

static int knn_launch(int mode, int b, int n, int m, int k, const float* xyz, const float* qxyz, int* idx,
                      float* dist2, long long sq, long long sk, int heap_ties, cudaStream_t st) {
  if (b <= 0 || m <= 0 || k <= 0) return PCREID_OK;
  if (n <= 0 || !xyz || !qxyz || !idx) return PCREID_ERR_ARG;
  if (heap_ties && k > 100) return PCREID_ERR_ARG;   // reference limit (knn.py:30)
  if (!heap_ties && k > n) return PCREID_ERR_ARG;
  static const bool use_rounds = getenv("PCREID_KNN_ORDERED") && !strcmp(getenv("PCREID_KNN_ORDERED"), "rounds");   // A/B knob
  if (!use_rounds && n <= 1024 && k <= n && k <= 128 && b <= 65535)
    return mode == 0 ? knn_sel_dispatch<0>(b, n, m, k, xyz, qxyz, idx, dist2, sq, sk, heap_ties, st)
                     : knn_sel_dispatch<1>(b, n, m, k, xyz, qxyz, idx, dist2, sq, sk, heap_ties, st);
  int npad = (n + 31) & ~31;
  int warps = 8;
  const size_t budget = 200 * 1024;
  auto need = [&](int w) { return (size_t)w * npad * 4 + (heap_ties ? (size_t)w * 2 * k * 4 : 0); };
  while (warps > 1 && need(warps) > budget) warps >>= 1;
  if (need(warps) > budget) return PCREID_ERR_UNSUPPORTED;
  if (m < warps) { while (warps > 1 && warps / 2 >= m) warps >>= 1; }
  size_t smem = need(warps);
  dim3 grid(ceil_div(m, warps), b);
  if (mode == 0) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(knn_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    knn_kernel<0><<<grid, warps * 32, smem, st>>>(n, m, k, xyz, qxyz, idx, dist2, sq, sk, heap_ties, npad);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(knn_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    knn_kernel<1><<<grid, warps * 32, smem, st>>>(n, m, k, xyz, qxyz, idx, dist2, sq, sk, heap_ties, npad);
  }
  return pcreid_launch_status();
}

// ------------------------------------------------------------------------------------------------
// kNN as an unordered SET (fast-mode encoder): the set-abstraction layers max-pool over the k neighbours, so only
// the membership of the k nearest points matters, with the ordered kernel's tie rule at the k-th boundary (equal
// distances: lower index first).  One warp per query, candidates (coordinates, squared norms) in registers and
// reused for QPW consecutive queries; distances become order-preserving 32-bit keys; the k-th smallest key is found
// by a most-significant-bit-first radix search over the bits in which the keys of this query differ (one REDUX.ADD
// per bit instead of k rounds of two REDUX.MIN + rescan), then the members are emitted by ballot / prefix popcount.
// Same distance arithmetic as knn_kernel<1> (torch path, pointnet2_utils.py:169-216).
// ------------------------------------------------------------------------------------------------
// STAGED (large N): the candidates live in shared memory (one float4 per point, loaded once per CTA) instead of 3 x KPL
// registers per lane, so that 32 keys per lane still leave room for full occupancy.
template <int KPL, bool STAGED>
__global__ void __launch_bounds__(256) knn_set_kernel(int N, int M, int k, int qpw, const float* __restrict__ xyz,
                                                      const float* __restrict__ qxyz, int* __restrict__ idx) {
  __shared__ float4 cand[STAGED ? KPL * 32 : 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int q0 = (blockIdx.x * 8 + warp) * qpw;
  const float* P = xyz + (size_t)b * N * 3;
  if (STAGED) {
    for (int i = threadIdx.x; i < N; i += 256) cand[i] = make_float4(__ldg(P + i * 3), __ldg(P + i * 3 + 1), __ldg(P + i * 3 + 2), 0.f);
    __syncthreads();
  }
  if (q0 >= M) return;
  float px[STAGED ? 1 : KPL], py[STAGED ? 1 : KPL], pz[STAGED ? 1 : KPL];
  if (!STAGED) {
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
      const int i = lane + 32 * j;
      px[j] = py[j] = pz[j] = 0.f;
      if (i < N) { px[j] = __ldg(P + i * 3); py[j] = __ldg(P + i * 3 + 1); pz[j] = __ldg(P + i * 3 + 2); }
    }
  }
  const uint32_t lt_mask = (1u << lane) - 1u;
  const int qend = min(q0 + qpw, M);
  for (int q = q0; q < qend; ++q) {
    const float* Q = qxyz + ((size_t)b * M + q) * 3;
    const float qx = __ldg(Q), qy = __ldg(Q + 1), qz = __ldg(Q + 2);
    const float qn = sqnorm3(qx, qy, qz);
    uint32_t key[KPL];
    uint32_t lmin = 0xffffffffu, lmax = 0u;
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
      key[j] = 0xffffffffu;
      if (lane + 32 * j < N) {
        if (STAGED) {
          const float4 c = cand[lane + 32 * j];
          key[j] = f32_to_ordered(dist_expand(qx, qy, qz, qn, c.x, c.y, c.z));
        } else {
          key[j] = f32_to_ordered(dist_expand(qx, qy, qz, qn, px[j], py[j], pz[j]));
        }
        lmin = min(lmin, key[j]);
        lmax = max(lmax, key[j]);
      }
    }
    const uint32_t kmin = __reduce_min_sync(FULL_MASK, lmin), kmax = __reduce_max_sync(FULL_MASK, lmax);
    // T = k-th smallest key: largest T with count(key < T) < k; bits above the highest differing bit are common
    uint32_t T = kmin;
    bool exact = false;         // a threshold with exactly k keys below it: the member set is known, no tie to break
    if (kmin != kmax) {
      const int top = 31 - __clz(kmin ^ kmax);
      T = top == 31 ? 0u : (kmin & ~((2u << top) - 1u));
      for (int bit = top; bit >= 0; --bit) {
        const uint32_t t = T | (1u << bit);
        int c = 0;
#pragma unroll
        for (int j = 0; j < KPL; ++j) c += key[j] < t ? 1 : 0;
        c = __reduce_add_sync(FULL_MASK, c);
        // distinct keys: the search interval isolates the gap between the k-th and (k+1)-th key after ~log2(N) bits
        if (c == k) { T = t; exact = true; break; }
        if (c < k) T = t;
      }
    }
    int* oi = idx + ((size_t)b * M + q) * k;
    int base = 0;
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
      const bool in = key[j] < T;
      const uint32_t bal = __ballot_sync(FULL_MASK, in);
      if (in) oi[base + __popc(bal & lt_mask)] = lane + 32 * j;
      base += __popc(bal);
    }
    if (exact) continue;
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
      const bool eq = key[j] == T && lane + 32 * j < N;
      const uint32_t bal = __ballot_sync(FULL_MASK, eq);
      const int pos = base + __popc(bal & lt_mask);
      if (eq && pos < k) oi[pos] = lane + 32 * j;
      base += __popc(bal);
    }
  }
}

template <int KPL, bool STAGED>
static int knn_set_launch(int b, int n, int m, int k, const float* xyz, const float* q, int* idx, cudaStream_t st) {
  const int qpw = m >= 64 ? 8 : (m >= 16 ? 2 : 1);
  knn_set_kernel<KPL, STAGED><<<dim3(ceil_div(m, 8 * qpw), b), 256, 0, st>>>(n, m, k, qpw, xyz, q, idx);
  return pcreid_launch_status();
}

// ------------------------------------------------------------------------------------------------
// DGCNN kNN in feature space (models/dgcnn_orig.py:22-28)
//   m_ij  = sequential fma chain over channels (what torch.matmul does on the CPU oracle)
//   xx_i  = sum_c x_ci^2, ATen cascade: sequential inside blocks of 16 channels, block sums added in order
//   pd_ij = ((-xx_j) - (-2 m_ij)) - xx_i ; k largest, lower index first on ties
// CTA = 8 warps x QPW queries; the object's features stream through shared memory in 128-point tiles.
// ------------------------------------------------------------------------------------------------
constexpr int KF_JT = 128;
template <int QPW>
__global__ void __launch_bounds__(256) knn_feature_kernel(int C, int N, int k, const float* __restrict__ x,
                                                          long long x_bs, int* __restrict__ idx, int npad) {
  extern __shared__ uint32_t smem_u32[];
  constexpr int QPC = 8 * QPW;
  float* xs = reinterpret_cast<float*>(smem_u32);              // [C][KF_JT]
  float* xx = xs + (size_t)C * KF_JT;                          // [npad]
  float* qv = xx + npad;                                       // [QPC][C]
  uint32_t* keys = reinterpret_cast<uint32_t*>(qv + (size_t)QPC * C);   // [QPC][npad]
  const int b = blockIdx.y, q0 = blockIdx.x * QPC;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* xb = x + (size_t)b * x_bs;

  for (int j = threadIdx.x; j < N; j += 256) {
    float tot = 0.f;
    for (int c0 = 0; c0 < C; c0 += 16) {
      float blk = 0.f;
      const int ce = min(c0 + 16, C);
      for (int c = c0; c < ce; ++c) { float v = xb[(size_t)c * N + j]; blk = __fadd_rn(blk, __fmul_rn(v, v)); }
      tot = c0 == 0 ? blk : __fadd_rn(tot, blk);
    }
    xx[j] = tot;
  }
  for (int t = threadIdx.x; t < QPC * C; t += 256) {
    int ql = t / C, c = t % C;
    qv[t] = (q0 + ql < N) ? xb[(size_t)c * N + q0 + ql] : 0.f;
  }
  for (int j0 = 0; j0 < N; j0 += KF_JT) {
    __syncthreads();
    for (int t = threadIdx.x; t < C * KF_JT; t += 256) {
      int c = t / KF_JT, j = t % KF_JT;
      xs[t] = (j0 + j < N) ? xb[(size_t)c * N + j0 + j] : 0.f;
    }
    __syncthreads();
    float acc[QPW][4];
#pragma unroll
    for (int a = 0; a < QPW; ++a)
#pragma unroll
      for (int t = 0; t < 4; ++t) acc[a][t] = 0.f;
    for (int c = 0; c < C; ++c) {
      float pv[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) pv[t] = xs[c * KF_JT + lane + 32 * t];
#pragma unroll
      for (int a = 0; a < QPW; ++a) {
        const float qc = qv[(warp * QPW + a) * C + c];
#pragma unroll
        for (int t = 0; t < 4; ++t) acc[a][t] = __fmaf_rn(qc, pv[t], acc[a][t]);
      }
    }
#pragma unroll
    for (int a = 0; a < QPW; ++a) {
      const int ql = warp * QPW + a;
      if (q0 + ql >= N) continue;
      const float xi = xx[q0 + ql];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int j = j0 + lane + 32 * t;
        if (j < N) {
          float inner = __fmul_rn(-2.f, acc[a][t]);
          float pd = __fsub_rn(__fsub_rn(-xx[j], inner), xi);   // (-xx (B,1,N)) - inner - xx^T
          keys[(size_t)ql * npad + j] = f32_to_ordered(-pd);
        }
      }
    }
  }
  __syncwarp();
  for (int a = 0; a < QPW; ++a) {
    const int ql = warp * QPW + a, q = q0 + ql;
    if (q >= N) continue;
    uint32_t* kq = keys + (size_t)ql * npad;
    uint32_t lmin = 0xffffffffu;
    int lidx = lane;
    for (int i = lane; i < N; i += 32) { uint32_t key = kq[i]; if (key < lmin) { lmin = key; lidx = i; } }
    int* oi = idx + ((size_t)b * N + q) * k;
    for (int r = 0; r < k; ++r) {
      uint32_t dmin = warp_min_u32(lmin);
      uint32_t win = warp_min_u32(lmin == dmin ? (uint32_t)lidx : 0xffffffffu);
      if (lane == 0) oi[r] = (int)win;
      if (lane == (int)(win & 31u)) {
        kq[win] = 0xffffffffu;
        lmin = 0xffffffffu;
        lidx = lane;
        for (int i = lane; i < N; i += 32) { uint32_t key = kq[i]; if (key < lmin) { lmin = key; lidx = i; } }
      }
      __syncwarp();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// furthest point sampling
// ------------------------------------------------------------------------------------------------
// Tie rule of the reference (first strict '>' per thread over k = tid, tid+bs, ..; shared-memory tree that
// keeps the lower slot on equal values): among tied maxima the winner minimises (bitreverse(k mod bs), k).
__device__ __forceinline__ uint32_t fps_tie32(int k, int log2bs) {
  if (log2bs == 0) return (uint32_t)k;
  uint32_t kmod = (uint32_t)k & ((1u << log2bs) - 1u);
  uint32_t rev = __brev(kmod) >> (32 - log2bs);
  return (rev << (32 - log2bs)) | ((uint32_t)k >> log2bs);
}
__device__ __forceinline__ int fps_untie32(uint32_t t, int log2bs) {
  if (log2bs == 0) return (int)t;
  uint32_t rev = t >> (32 - log2bs);
  uint32_t kmod = __brev(rev) >> (32 - log2bs);
  uint32_t hi = t & ((1u << (32 - log2bs)) - 1u);
  return (int)((hi << log2bs) | kmod);
}

template <int PPT, int MAXT>
__global__ void __launch_bounds__(MAXT) fps_kernel(int N, int M, int log2bs, const float* __restrict__ data,
                                                   float* __restrict__ temp, int* __restrict__ idxs, int with_dist,
                                                   const int* __restrict__ start) {
  // with_dist: 0 = xyz, the op's distance fma(dz,dz,fma(dx,dx,dy*dy)); 1 = precomputed (b,n,n) distances; 2 = xyz, the torch
  // path's (dx*dx + dy*dy) + dz*dz (then log2bs = 0: ties go to the lowest index, and the first sample is start[b])
  if (M <= 0) return;
  __shared__ uint32_t s_hi[2][32], s_lo[2][32];
  const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = T >> 5;
  const int b = blockIdx.x;
  const float* ds = data + (size_t)b * N * (with_dist == 1 ? (size_t)N : 3);
  float* tp = temp ? temp + (size_t)b * N : nullptr;
  int* out = idxs + (size_t)b * M;

  float x[PPT], y[PPT], z[PPT], t[PPT];
  uint32_t tie[PPT];
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    int k = tid + T * j;
    x[j] = y[j] = z[j] = 0.f;
    t[j] = 1e10f;
    tie[j] = 0;
    if (k < N) {
      if (with_dist != 1) { x[j] = ds[k * 3]; y[j] = ds[k * 3 + 1]; z[j] = ds[k * 3 + 2]; }
      if (tp) t[j] = tp[k];
      tie[j] = ~fps_tie32(k, log2bs);
    }
  }
  int old = start ? start[b] : 0;
  if (tid == 0) out[0] = old;
  for (int s = 1; s < M; ++s) {
    float x1 = 0.f, y1 = 0.f, z1 = 0.f;
    if (with_dist != 1) { x1 = __ldg(ds + old * 3); y1 = __ldg(ds + old * 3 + 1); z1 = __ldg(ds + old * 3 + 2); }
    uint32_t bhi = 0, blo = 0;
    bool have = false;
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      int k = tid + T * j;
      if (k < N) {
        float d = with_dist == 1 ? __ldg(ds + (size_t)old * N + k)
                                 : (with_dist == 2 ? dist_sum_sq(x[j], y[j], z[j], x1, y1, z1) : dist_direct(x[j], y[j], z[j], x1, y1, z1));
        float d2 = fminf(d, t[j]);
        t[j] = d2;
        uint32_t hi = f32_to_ordered(d2);
        if (!have || hi > bhi || (hi == bhi && tie[j] > blo)) { bhi = hi; blo = tie[j]; have = true; }
      }
    }
    uint32_t whi = warp_max_u32(bhi);
    uint32_t wlo = warp_max_u32((have && bhi == whi) ? blo : 0u);
    if (nwarp > 1) {
      const int buf = s & 1;
      if (lane == 0) { s_hi[buf][warp] = whi; s_lo[buf][warp] = wlo; }
      __syncthreads();
      uint32_t h = lane < nwarp ? s_hi[buf][lane] : 0u;
      uint32_t l = lane < nwarp ? s_lo[buf][lane] : 0u;
      whi = warp_max_u32(h);
      wlo = warp_max_u32((lane < nwarp && h == whi) ? l : 0u);
    }
    old = fps_untie32(~wlo, log2bs);
    if (tid == 0) out[s] = old;
  }
  if (tp) {
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      int k = tid + T * j;
      if (k < N) tp[k] = t[j];
    }
  }
}

// Second FPS kernel (xyz inputs, modes 0 / 2): coordinates live in shared memory in TIE-PRIORITY order and only the running
// min-distances stay in registers, so a 1024-point object costs ~70 registers of ONE warp (14+ objects resident per SM; the
// register-resident kernel above needs 220 and runs 2048 objects in two waves).
//   slot s = rev(k mod bs) * Q + k div bs   (Q = ceil(N / bs); slots whose k >= N are padding with t = -1: never selected)
// ascends exactly with the reference's tie priority (fps_tie32), and a thread scans its slots 4 (T j + tid) + c in ascending
// order with a strict '>', so the block winner is  min slot over { slots whose distance equals the block maximum } : one
// REDUX.MAX over the value bits (distances are >= +0: the int order is the float order) and one REDUX.MIN over the slots --
// no per-point tie key, no lexicographic compare.  The next centroid's coordinates are a shared-memory broadcast read.
struct FpsSlots {
  int log2bs, Q;
  __device__ __forceinline__ int to_k(int s) const {
    const int rv = s / Q, q = s - rv * Q;
    return log2bs ? (q << log2bs) + (int)(__brev((uint32_t)rv) >> (32 - log2bs)) : q;
  }
  __device__ __forceinline__ bool valid(int s, int N) const { return (s / Q) < (1 << log2bs) && to_k(s) < N; }
  __device__ __forceinline__ int to_slot(int k) const {
    if (!log2bs) return k;
    const uint32_t kmod = (uint32_t)k & ((1u << log2bs) - 1u);
    return (int)(__brev(kmod) >> (32 - log2bs)) * Q + (k >> log2bs);
  }
};

constexpr int fps_rank_maxt(int ppt) { return ppt >= 32 ? 256 : ppt >= 16 ? 512 : 1024; }   // keeps t[PPT] out of local memory

template <int PPT, int MODE>
__global__ void __launch_bounds__(fps_rank_maxt(PPT)) fps_rank_kernel(int N, int M, int nsl, FpsSlots sl, const float* __restrict__ data,
                                                        float* __restrict__ temp, int* __restrict__ idxs,
                                                        const int* __restrict__ start) {
  extern __shared__ __align__(16) float fps_sm[];
  __shared__ int s_val[2][32], s_slot[2][32];
  const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = T >> 5;
  const int b = blockIdx.x;
  const float* ds = data + (size_t)b * N * 3;
  float* tp = temp ? temp + (size_t)b * N : nullptr;
  int* out = idxs + (size_t)b * M;
  float *xs = fps_sm, *ys = fps_sm + nsl, *zs = fps_sm + 2 * nsl;
  for (int s = tid; s < nsl; s += T) { xs[s] = 0.f; ys[s] = 0.f; zs[s] = 0.f; }
  __syncthreads();
  for (int i = tid; i < 3 * N; i += T) {                       // coalesced read, scattered into the slot order
    const int k = i / 3, c = i - 3 * k;
    fps_sm[c * nsl + sl.to_slot(k)] = ds[i];
  }
  float t[PPT];
#pragma unroll
  for (int j = 0; j < PPT / 4; ++j)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int s = 4 * (T * j + tid) + c;
      t[4 * j + c] = sl.valid(s, N) ? (tp ? tp[sl.to_k(s)] : 1e10f) : -1.f;
    }
  int old_slot = sl.to_slot(start ? start[b] : 0);
  if (tid == 0) out[0] = start ? start[b] : 0;
  __syncthreads();
  const float4 *x4 = reinterpret_cast<const float4*>(xs), *y4 = reinterpret_cast<const float4*>(ys),
               *z4 = reinterpret_cast<const float4*>(zs);
  for (int s = 1; s < M; ++s) {
    const float x1 = xs[old_slot], y1 = ys[old_slot], z1 = zs[old_slot];
    float best = -2.f;
    int bslot = 0;
#pragma unroll
    for (int j = 0; j < PPT / 4; ++j) {
      const int g = T * j + tid;
      const float4 X = x4[g], Y = y4[g], Z = z4[g];
      const float xx[4] = {X.x, X.y, X.z, X.w}, yy[4] = {Y.x, Y.y, Y.z, Y.w}, zz[4] = {Z.x, Z.y, Z.z, Z.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float d = MODE == 2 ? dist_sum_sq(xx[c], yy[c], zz[c], x1, y1, z1) : dist_direct(xx[c], yy[c], zz[c], x1, y1, z1);
        const float d2 = fminf(d, t[4 * j + c]);
        t[4 * j + c] = d2;
        if (d2 > best) { best = d2; bslot = 4 * g + c; }
      }
    }
    int wv = __reduce_max_sync(FULL_MASK, __float_as_int(best));
    uint32_t ws = __reduce_min_sync(FULL_MASK, __float_as_int(best) == wv ? (uint32_t)bslot : 0xffffffffu);
    if (nwarp > 1) {
      const int buf = s & 1;
      if (lane == 0) { s_val[buf][warp] = wv; s_slot[buf][warp] = (int)ws; }
      __syncthreads();
      const int v = lane < nwarp ? s_val[buf][lane] : (int)0x80000000;
      const uint32_t l = lane < nwarp ? (uint32_t)s_slot[buf][lane] : 0xffffffffu;
      wv = __reduce_max_sync(FULL_MASK, v);
      ws = __reduce_min_sync(FULL_MASK, v == wv ? l : 0xffffffffu);
    }
    old_slot = (int)ws;
    if (tid == 0) out[s] = sl.to_k(old_slot);
  }
  if (tp) {
#pragma unroll
    for (int j = 0; j < PPT / 4; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int s = 4 * (T * j + tid) + c;
        if (sl.valid(s, N)) tp[sl.to_k(s)] = t[4 * j + c];
      }
  }
}

static int fps_block_size_ref(int n) {   // opt_n_threads, furthest_point_sample_cuda.cu:11-15
  int pow_2 = (int)(std::log(static_cast<double>(n)) / std::log(2.0));
  int t = 1 << pow_2;
  if (t > 1024) t = 1024;
  if (t < 1) t = 1;
  return t;
}

// shared-memory / rank-order kernel: returns false when the shape is outside its range (the register kernel serves it)
static bool fps_rank_launch(int b, int n, int m, int log2bs, const float* data, float* temp, int* idxs, int with_dist,
                            cudaStream_t st, const int* start) {
  static const int knob = [] { const char* e = getenv("PCREID_FPS_PPT"); return e ? atoi(e) : 0; }();   // A/B: 0 = heuristic, -1 = off
  if (knob < 0) return false;
  const int bs = 1 << log2bs;                                 // 1 on the torch path: slots are the point indices
  const int Q = ceil_div(n, bs);
  const long long raw = (long long)bs * Q;
  if (raw > 8192) return false;
  // slots per thread: 32 (a warp per <= 1024-point object, no barriers; the fewest warps per large object -- the two-level
  // arg-max chain, not the distance arithmetic, bounds a step) unless few small objects leave the SMs empty
  // (profiles/r02_fps_sweep.jsonl)
  int want = knob > 0 ? knob : ((b >= 296 || raw >= 4096) ? 32 : 8);
  int W = 1;
  while (W < 32 && (long long)W * 32 * want < raw) W <<= 1;
  int ppt = 4;
  while ((long long)W * 32 * ppt < raw) ppt <<= 1;
  if (ppt > 32 || W * 32 > fps_rank_maxt(ppt)) return false;
  const int T = 32 * W, nsl = T * ppt;
  const size_t smem = (size_t)nsl * 12;
  FpsSlots sl{log2bs, Q};
#define FPS_RANK_CASE(P)                                                                                                   \
  case P:                                                                                                                  \
    if (with_dist == 2) {                                                                                                  \
      cudaFuncSetAttribute(fps_rank_kernel<P, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                 \
      fps_rank_kernel<P, 2><<<b, T, smem, st>>>(n, m, nsl, sl, data, temp, idxs, start);                                   \
    } else {                                                                                                               \
      cudaFuncSetAttribute(fps_rank_kernel<P, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                 \
      fps_rank_kernel<P, 0><<<b, T, smem, st>>>(n, m, nsl, sl, data, temp, idxs, start);                                   \
    }                                                                                                                      \
    break;
  switch (ppt) {
    FPS_RANK_CASE(4) FPS_RANK_CASE(8) FPS_RANK_CASE(16) FPS_RANK_CASE(32)
  }
#undef FPS_RANK_CASE
  return true;
}

static int fps_launch(int b, int n, int m, const float* data, float* temp, int* idxs, int with_dist, cudaStream_t st,
                      const int* start = nullptr) {
  if (b <= 0 || m <= 0) return PCREID_OK;
  if (n <= 0 || !data || !idxs) return PCREID_ERR_ARG;
  int bs = fps_block_size_ref(n), log2bs = 0;
  while ((1 << log2bs) < bs) ++log2bs;
  if (with_dist == 2) log2bs = 0;                           // torch.max: the first index among tied maxima
  if (with_dist != 1 && fps_rank_launch(b, n, m, log2bs, data, temp, idxs, with_dist, st, start)) return pcreid_launch_status();
  int T;
  if (b >= 592 && n <= 1024 && with_dist != 1) T = 32;      // many small objects: one warp each, no barriers
  else {
    T = 32;
    while (T < 1024 && T * 4 < n) T <<= 1;                  // ~4 points per thread, latency bound otherwise
  }
  int ppt = ceil_div(n, T);
  int p2 = 1;
  while (p2 < ppt) p2 <<= 1;
  if (p2 > 32) return PCREID_ERR_UNSUPPORTED;               // n > 32768
#define FPS_CASE(P)                                                                              \
  case P:                                                                                        \
    if (T == 32) fps_kernel<P, 32><<<b, T, 0, st>>>(n, m, log2bs, data, temp, idxs, with_dist, start);   \
    else fps_kernel<P, 1024><<<b, T, 0, st>>>(n, m, log2bs, data, temp, idxs, with_dist, start);         \
    break;
  switch (p2) {
    FPS_CASE(1) FPS_CASE(2) FPS_CASE(4) FPS_CASE(8) FPS_CASE(16) FPS_CASE(32)
  }
#undef FPS_CASE
  return pcreid_launch_status();
}

// ------------------------------------------------------------------------------------------------
// ball query
// ------------------------------------------------------------------------------------------------
// MODE 0: the mmdet3d op (ball_query_cuda.cu:11-54): direct-form distance, hit = d2 == 0 || (min_r2 <= d2 < max_r2), rows
//         without a hit keep the caller's zeros.
// MODE 1: the torch-path query_ball_point of the ReID backbone (models/pointnet2_utils.py:218-240, use_knn=False):
//         expansion-form distance of square_distance (op for op, as knn_kernel<1>), hit = !(d2 > r2), the first nsample hits
//         in index order, padded with the first hit; a row without any hit is filled with N exactly like the reference's
//         sort-based formulation (which then fails in index_points).
template <int MODE>
__global__ void __launch_bounds__(256) ball_query_kernel(int N, int M, int k, float min_r2, float max_r2,
                                                         const float* __restrict__ qxyz, const float* __restrict__ xyz,
                                                         int* __restrict__ idx) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + warp, b = blockIdx.y;
  if (q >= M) return;
  const float* P = xyz + (size_t)b * N * 3;
  const float* Q = qxyz + ((size_t)b * M + q) * 3;
  const float qx = Q[0], qy = Q[1], qz = Q[2];
  const float qn = MODE == 1 ? sqnorm3(qx, qy, qz) : 0.f;
  int* o = idx + ((size_t)b * M + q) * k;
  int cnt = 0;
  for (int base = 0; base < N && cnt < k; base += 32) {
    int i = base + lane;
    bool hit = false;
    if (i < N) {
      const float px = __ldg(P + i * 3), py = __ldg(P + i * 3 + 1), pz = __ldg(P + i * 3 + 2);
      if (MODE == 0) {
        float d2 = dist_direct(qx, qy, qz, px, py, pz);
        hit = (d2 == 0.f) || (d2 >= min_r2 && d2 < max_r2);
      } else {
        hit = !(dist_expand(qx, qy, qz, qn, px, py, pz) > max_r2);
      }
    }
    uint32_t mask = __ballot_sync(FULL_MASK, hit);
    if (mask == 0) continue;
    if (cnt == 0) {   // first hit is broadcast to every slot (ball_query_cuda.cu:44-48 / pointnet2_utils.py:237-239)
      int first = base + __ffs(mask) - 1;
      for (int l = lane; l < k; l += 32) o[l] = first;
      __syncwarp();
    }
    int pos = cnt + __popc(mask & ((1u << lane) - 1u));
    if (hit && pos < k) o[pos] = i;
    cnt += __popc(mask);
  }
  if (MODE == 1 && cnt == 0)
    for (int l = lane; l < k; l += 32) o[l] = N;
}

// Thread-per-query variant (the default for >= 64 queries per object).  The warp-per-query kernel above spends a ballot +
// popcount + dependent count update per 32-point chunk and keeps 31 of 32 lanes waiting on that chain; here every lane walks
// its own query over the object's points, which are staged once per CTA in shared memory as float4 (x, y, z, |p|^2) and read
// as warp-wide broadcasts (one wavefront per point and warp).  Hits go to a shared-memory staging tile [slot][query] (row
// stride 129 words: conflict-free for the per-thread writes and for the transposed read-out), padded with the first hit as
// the reference does, and leave as fully coalesced 128-byte rows.  Same distances, same predicate, same order as MODE 0 / 1.
constexpr int BQ_T = 128;          // queries per CTA
constexpr int BQ_TILE = 2048;      // points staged per pass (32 KB)
template <int MODE>
__global__ void __launch_bounds__(BQ_T) ball_query_tq_kernel(int N, int M, int k, float min_r2, float max_r2,
                                                             const float* __restrict__ qxyz, const float* __restrict__ xyz,
                                                             int* __restrict__ idx) {
  extern __shared__ __align__(16) uint8_t bq_smem[];
  float4* pts = reinterpret_cast<float4*>(bq_smem);
  int* stage = reinterpret_cast<int*>(bq_smem + (size_t)min(N, BQ_TILE) * sizeof(float4));      // [k][BQ_T + 1]
  const int t = threadIdx.x, b = blockIdx.y, q0 = blockIdx.x * BQ_T, q = q0 + t;
  const bool live = q < M;
  const float* P = xyz + (size_t)b * N * 3;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (live) {
    const float* Q = qxyz + ((size_t)b * M + q) * 3;
    qx = Q[0]; qy = Q[1]; qz = Q[2];
  }
  const float qn = MODE == 1 ? sqnorm3(qx, qy, qz) : 0.f;
  int cnt = 0, first = MODE == 1 ? N : 0;
  int* col = stage + t;
  for (int base = 0; base < N; base += BQ_TILE) {
    const int nt = min(BQ_TILE, N - base);
    __syncthreads();
    for (int i = t; i < nt; i += BQ_T) {
      const float x = __ldg(P + (size_t)(base + i) * 3), y = __ldg(P + (size_t)(base + i) * 3 + 1), z = __ldg(P + (size_t)(base + i) * 3 + 2);
      pts[i] = make_float4(x, y, z, 0.f);
    }
    __syncthreads();
    if (!live || cnt >= k) continue;
    for (int i = 0; i < nt; ++i) {
      const float4 p = pts[i];
      bool hit;
      if (MODE == 0) {
        const float d2 = dist_direct(qx, qy, qz, p.x, p.y, p.z);
        hit = (d2 == 0.f) || (d2 >= min_r2 && d2 < max_r2);
      } else {
        hit = !(dist_expand(qx, qy, qz, qn, p.x, p.y, p.z) > max_r2);
      }
      if (hit) {
        if (cnt == 0) first = base + i;
        col[cnt * (BQ_T + 1)] = base + i;
        if (++cnt >= k) break;
      }
    }
  }
  // remaining slots: the first hit (MODE 0 without any hit: the caller's zeros; MODE 1 without any hit: N)
  for (int s = cnt; s < k; ++s) col[s * (BQ_T + 1)] = first;
  __syncthreads();
  const int nq = min(BQ_T, M - q0);
  int* o = idx + ((size_t)b * M + q0) * k;
  for (int i = t; i < nq * k; i += BQ_T) o[i] = stage[(i % k) * (BQ_T + 1) + i / k];
}

// ------------------------------------------------------------------------------------------------
// group_points / gather_points:  out[b, c, p] = points[b, c, idx[b, p]],  p in [0, P)  (P = S*k or M)
// ------------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256) group_kernel(int C, int N, int P, const float* __restrict__ points,
                                                    const int* __restrict__ idx, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int p0 = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (p0 >= P) return;
  const int* ib = idx + (size_t)b * P + p0;
  int id[VEC];
  if (VEC == 4) {
    int4 v = *reinterpret_cast<const int4*>(ib);
    id[0] = v.x; id[1 % VEC] = v.y; id[2 % VEC] = v.z; id[3 % VEC] = v.w;
  } else {
    id[0] = ib[0];
  }
  const float* pb = points + (size_t)b * C * N;
  float* ob = out + (size_t)b * C * P + p0;
#pragma unroll 4
  for (int c = 0; c < C; ++c) {
    const float* row = pb + (size_t)c * N;
    if (VEC == 4) {
      float4 v = make_float4(__ldg(row + id[0]), __ldg(row + id[1 % VEC]), __ldg(row + id[2 % VEC]), __ldg(row + id[3 % VEC]));
      st_cs_f4(ob + (size_t)c * P, v);
    } else {
      ob[(size_t)c * P] = __ldg(row + id[0]);
    }
  }
}

static int group_launch(int b, int c, int n, int p, const float* points, const int* idx, float* out, cudaStream_t st) {
  if (b <= 0 || c <= 0 || p <= 0) return PCREID_OK;
  if (!points || !idx || !out || n <= 0) return PCREID_ERR_ARG;
  bool vec = (p % 4 == 0) && ((reinterpret_cast<uintptr_t>(idx) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  if (vec) {
    dim3 grid(ceil_div(p / 4, 256), b);
    group_kernel<4><<<grid, 256, 0, st>>>(c, n, p, points, idx, out);
  } else {
    dim3 grid(ceil_div(p, 256), b);
    group_kernel<1><<<grid, 256, 0, st>>>(c, n, p, points, idx, out);
  }
  return pcreid_launch_status();
}

// ------------------------------------------------------------------------------------------------
// query_group: the body of QueryAndGroup.forward (mmdet3d/ops/group_points/group_points.py:93-118) in ONE pass:
//   out[b, 0:3, s, j]   = (xyz[b, idx[b,s,j], :] - center[b, s, :]) (/ radius when normalize_xyz)      (use_xyz)
//   out[b, c0+c, s, j]  = features[b, c, idx[b,s,j]]                                                    (features given)
//   gxyz[b, 0:3, s, j]  = xyz[b, idx[b,s,j], :]                                                         (return_grouped_xyz)
// The reference runs two grouping kernels, a transpose, a subtraction, a division and a cat over the (B, C+3, S, k)
// tensor; here the index is read once per position and every output element is written exactly once.  Same fp32
// arithmetic (one subtraction, one IEEE division), so results are bit-identical.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) query_group_kernel(int C, int N, int S, int k, const float* __restrict__ xyz,
                                                          const float* __restrict__ center, const float* __restrict__ feats,
                                                          const int* __restrict__ idx, int use_xyz, float inv_div, float* __restrict__ out,
                                                          float* __restrict__ gxyz) {
  const int b = blockIdx.z, e = blockIdx.x * blockDim.x + threadIdx.x;
  const int E = S * k;
  if (e >= E) return;
  const int s = e / k, src = idx[(size_t)b * E + e];
  const int c_xyz = use_xyz ? 3 : 0, Ctot = c_xyz + (feats ? C : 0);
  float* ob = out + (size_t)b * Ctot * E + e;
  if (blockIdx.y == 0 && (use_xyz || gxyz)) {
    const float* p = xyz + ((size_t)b * N + src) * 3;
    const float* q = center + ((size_t)b * S + s) * 3;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float v = __ldg(p + a);
      if (gxyz) gxyz[((size_t)b * 3 + a) * E + e] = v;
      if (use_xyz) {
        float d = v - __ldg(q + a);
        if (inv_div != 0.f) d = __fdiv_rn(d, inv_div);
        ob[(size_t)a * E] = d;
      }
    }
  }
  if (feats) {
    const int c0 = blockIdx.y * 16;
    const float* fb = feats + (size_t)b * C * N + src;
#pragma unroll 4
    for (int c = c0; c < min(c0 + 16, C); ++c) ob[(size_t)(c_xyz + c) * E] = __ldg(fb + (size_t)c * N);
  }
}

// ------------------------------------------------------------------------------------------------
// C ABI (declared in include/pcreid.h)
// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
// three_nn / three_interpolate (mmdet3d/ops/interpolate/src/three_nn_cuda.cu:11-66, three_interpolate_cuda.cu:11-38)
//   three_nn: the three nearest `known` points of every `unknown` point, ascending (d2, index) -- the reference's
//   sequential scan with strict `<` inserts ties after the earlier index.  d2 = fma(dz,dz, fma(dx,dx, dy*dy)) (the
//   contraction nvcc applies to the reference expression, SASS-verified); slots that no point fills keep
//   (index 0, d2 = float(1e40) = +inf).  The known points stream through shared memory in tiles that all 256 queries
//   of the CTA scan with broadcast reads, instead of every thread walking global memory on its own.
//   three_interpolate: out[b,c,n] = fma(w2,p2, fma(w0,p0, w1*p1)) (same contraction); index / weight triples are read
//   once per point and reused over a block of channels, not once per (channel, point).
// ------------------------------------------------------------------------------------------------
constexpr int TNN_TILE = 1024;
__global__ void __launch_bounds__(256) three_nn_kernel(int n, int m, const float* __restrict__ unknown, const float* __restrict__ known,
                                                       float* __restrict__ dist2, int* __restrict__ idx) {
  __shared__ float ks[TNN_TILE * 3];
  const int b = blockIdx.y, q = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = q < n;
  float ux = 0.f, uy = 0.f, uz = 0.f;
  if (live) {
    const float* u = unknown + ((size_t)b * n + q) * 3;
    ux = u[0]; uy = u[1]; uz = u[2];
  }
  // the reference keeps its running best in doubles initialised to 1e40 and inserts on strict `<`: every finite float
  // beats an empty slot, +inf and NaN never enter, an empty slot is written back as float(1e40) = +inf
  const float INF = __int_as_float(0x7f800000);
  float b1 = INF, b2 = INF, b3 = INF;
  int i1 = 0, i2 = 0, i3 = 0;
  for (int k0 = 0; k0 < m; k0 += TNN_TILE) {
    const int kt = min(TNN_TILE, m - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < kt * 3; i += blockDim.x) ks[i] = known[((size_t)b * m + k0) * 3 + i];
    __syncthreads();
    if (!live) continue;
    for (int k = 0; k < kt; ++k) {
      const float d = dist_direct(ux, uy, uz, ks[3 * k], ks[3 * k + 1], ks[3 * k + 2]);
      if (d < b1) {
        b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k0 + k;
      } else if (d < b2) {
        b3 = b2; i3 = i2; b2 = d; i2 = k0 + k;
      } else if (d < b3) {
        b3 = d; i3 = k0 + k;
      }
    }
  }
  if (live) {
    float* od = dist2 + ((size_t)b * n + q) * 3;
    int* oi = idx + ((size_t)b * n + q) * 3;
    od[0] = b1; od[1] = b2; od[2] = b3;
    oi[0] = i1; oi[1] = i2; oi[2] = i3;
  }
}

__global__ void __launch_bounds__(256) three_interpolate_kernel(int c, int m, int n, const float* __restrict__ points,
                                                                const int* __restrict__ idx, const float* __restrict__ weight,
                                                                float* __restrict__ out) {
  const int b = blockIdx.z, p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int* ip = idx + ((size_t)b * n + p) * 3;
  const float* wp = weight + ((size_t)b * n + p) * 3;
  const int i0 = ip[0], i1 = ip[1], i2 = ip[2];
  const float w0 = wp[0], w1 = wp[1], w2 = wp[2];
  const int c0 = blockIdx.y * 16;
#pragma unroll 4
  for (int ch = c0; ch < min(c0 + 16, c); ++ch) {
    const float* pr = points + ((size_t)b * c + ch) * m;
    const float t = __fmul_rn(w1, __ldg(pr + i1));
    out[((size_t)b * c + ch) * n + p] = __fmaf_rn(w2, __ldg(pr + i2), __fmaf_rn(w0, __ldg(pr + i0), t));
  }
}

// ------------------------------------------------------------------------------------------------------------------
// pairwise squared feature distance (calc_square_dist, ops/furthest_point_sample/utils.py:4-31): the N x M matrix the
// F-FPS / FS samplers hand to furthest_point_sample_with_dist.  a (B, N, C), b (B, M, C) point-major; out (B, N, M).
//   dot = fma chain over c ascending;  |a|^2, |b|^2 likewise;  d = fma(-2, dot, |a|^2 + |b|^2);  norm: sqrt(d) / C
// 64 x 64 output tile per CTA (16 x 16 threads, 4 x 4 outputs each), channels staged through shared memory 16 at a time;
// output-write bound (4 N M bytes per object against 4 C (N + M) read).
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pairwise_sqdist_kernel(int N, int M, int C, const float* __restrict__ A,
                                                              const float* __restrict__ Bm, float* __restrict__ out, int norm) {
  __shared__ float As[16][65], Bs[16][65];
  const int b = blockIdx.z, i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float* a = A + (size_t)b * N * C;
  const float* bb = Bm + (size_t)b * M * C;
  float dot[4][4], a2[4], b2[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    a2[r] = 0.f; b2[r] = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) dot[r][q] = 0.f;
  }
  for (int c0 = 0; c0 < C; c0 += 16) {
    for (int e = threadIdx.x; e < 64 * 16; e += 256) {
      const int r = e >> 4, c = e & 15;                         // consecutive threads read consecutive channels of a row
      As[c][r] = (i0 + r < N && c0 + c < C) ? __ldg(a + (size_t)(i0 + r) * C + c0 + c) : 0.f;
      Bs[c][r] = (j0 + r < M && c0 + c < C) ? __ldg(bb + (size_t)(j0 + r) * C + c0 + c) : 0.f;
    }
    __syncthreads();
    const int cn = min(16, C - c0);
    for (int c = 0; c < cn; ++c) {
      float av[4], bv[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) { av[r] = As[c][ty + 16 * r]; bv[r] = Bs[c][tx + 16 * r]; }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        a2[r] = __fmaf_rn(av[r], av[r], a2[r]);
        b2[r] = __fmaf_rn(bv[r], bv[r], b2[r]);
#pragma unroll
        for (int q = 0; q < 4; ++q) dot[r][q] = __fmaf_rn(av[r], bv[q], dot[r][q]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty + 16 * r;
    if (i >= N) continue;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = j0 + tx + 16 * q;
      if (j >= M) continue;
      float d = __fmaf_rn(-2.f, dot[r][q], __fadd_rn(a2[r], b2[q]));
      if (norm) d = __fdiv_rn(__fsqrt_rn(d), (float)C);
      out[((size_t)b * N + i) * M + j] = d;
    }
  }
}

extern "C" {

int pcreid_pairwise_sqdist(int B, int N, int M, int C, const float* a, const float* b, float* out, int norm, void* stream) {
  if (B <= 0 || N <= 0 || M <= 0) return PCREID_OK;
  if (!a || !b || !out || C <= 0) return PCREID_ERR_ARG;
  if (B > 65535 || ceil_div(N, 64) > 65535) return PCREID_ERR_UNSUPPORTED;
  pairwise_sqdist_kernel<<<dim3(ceil_div(M, 64), ceil_div(N, 64), B), 256, 0, (cudaStream_t)stream>>>(N, M, C, a, b, out, norm);
  return pcreid_launch_status();
}

int pcreid_fps(int b, int n, int m, const float* xyz, float* temp, int* idx, void* stream) {
  return fps_launch(b, n, m, xyz, temp, idx, 0, (cudaStream_t)stream);
}
int pcreid_fps_with_dist(int b, int n, int m, const float* dist, float* temp, int* idx, void* stream) {
  return fps_launch(b, n, m, dist, temp, idx, 1, (cudaStream_t)stream);
}
int pcreid_fps_block_size(int n) { return n > 0 ? fps_block_size_ref(n) : 1; }
int pcreid_fps_torch(int b, int n, int m, const float* xyz, const int* start, int* idx, void* stream) {
  if (b > 0 && m > 0 && !start) return PCREID_ERR_ARG;
  return fps_launch(b, n, m, xyz, nullptr, idx, 2, (cudaStream_t)stream, start);
}

// mmdet3d op: idx/dist2 laid out [b, m, k] exactly like knn_kernel_launcher's outputs.
int pcreid_knn(int b, int n, int m, int k, const float* xyz, const float* new_xyz, int* idx, float* dist2, void* stream) {
  return knn_launch(0, b, n, m, k, xyz, new_xyz, idx, dist2, k, 1, 1, (cudaStream_t)stream);
}
// same search, output already transposed to [b, k, m] (what knn.py:62 returns after .transpose(2,1).contiguous())
int pcreid_knn_t(int b, int n, int m, int k, const float* xyz, const float* new_xyz, int* idx, float* dist2, void* stream) {
  return knn_launch(0, b, n, m, k, xyz, new_xyz, idx, dist2, 1, m, 1, (cudaStream_t)stream);
}
// torch-path kNN of the ReID backbones (pointnet2_utils.py:205-216): expansion-form distance, canonical
// (d, idx) order, idx [b, m, k] int32.
int pcreid_knn_point(int b, int n, int m, int k, const float* xyz, const float* new_xyz, int* idx, void* stream) {
  return knn_launch(1, b, n, m, k, xyz, new_xyz, idx, nullptr, k, 1, 0, (cudaStream_t)stream);
}

int pcreid_knn_point_set(int b, int n, int m, int k, const float* xyz, const float* new_xyz, int* idx, void* stream) {
  if (b <= 0 || m <= 0 || k <= 0) return PCREID_OK;
  if (!xyz || !new_xyz || !idx || n <= 0) return PCREID_ERR_ARG;
  if (k > n || n > 1024 || b > 65535) return PCREID_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  static const int staged_from = getenv("PCREID_KNN_STAGED_FROM") ? atoi(getenv("PCREID_KNN_STAGED_FROM")) : 257;   // A/B knob
  if (n <= 128) return knn_set_launch<4, false>(b, n, m, k, xyz, new_xyz, idx, st);
  if (n <= 256) return knn_set_launch<8, false>(b, n, m, k, xyz, new_xyz, idx, st);
  if (n <= 512) return n >= staged_from ? knn_set_launch<16, true>(b, n, m, k, xyz, new_xyz, idx, st)
                                        : knn_set_launch<16, false>(b, n, m, k, xyz, new_xyz, idx, st);
  return n >= staged_from ? knn_set_launch<32, true>(b, n, m, k, xyz, new_xyz, idx, st)
                          : knn_set_launch<32, false>(b, n, m, k, xyz, new_xyz, idx, st);
}

int pcreid_knn_feature(int b, int c, int n, int k, const float* x, long long x_bs, int* idx, void* stream) {
  if (b <= 0 || n <= 0 || k <= 0) return PCREID_OK;
  if (!x || !idx || c <= 0 || k > n) return PCREID_ERR_ARG;
  if (b > 65535) return PCREID_ERR_UNSUPPORTED;
  const int npad = (n + 31) & ~31;
  auto need = [&](int qpc) { return ((size_t)c * KF_JT + npad + (size_t)qpc * c + (size_t)qpc * npad) * 4; };
  cudaStream_t st = (cudaStream_t)stream;
  const size_t budget = 200 * 1024;
  if (need(32) <= budget) {
    cudaFuncSetAttribute(knn_feature_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need(32));
    knn_feature_kernel<4><<<dim3(ceil_div(n, 32), b), 256, need(32), st>>>(c, n, k, x, x_bs, idx, npad);
  } else if (need(16) <= budget) {
    cudaFuncSetAttribute(knn_feature_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need(16));
    knn_feature_kernel<2><<<dim3(ceil_div(n, 16), b), 256, need(16), st>>>(c, n, k, x, x_bs, idx, npad);
  } else if (need(8) <= budget) {
    cudaFuncSetAttribute(knn_feature_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need(8));
    knn_feature_kernel<1><<<dim3(ceil_div(n, 8), b), 256, need(8), st>>>(c, n, k, x, x_bs, idx, npad);
  } else {
    return PCREID_ERR_UNSUPPORTED;
  }
  return pcreid_launch_status();
}

int pcreid_ball_query(int b, int n, int m, float min_radius, float max_radius, int nsample, const float* new_xyz,
                      const float* xyz, int* idx, void* stream) {
  if (b <= 0 || m <= 0 || nsample <= 0) return PCREID_OK;
  if (n <= 0 || !new_xyz || !xyz || !idx) return PCREID_ERR_ARG;
  if (b > 65535) return PCREID_ERR_UNSUPPORTED;
  const size_t smem = (size_t)(n < BQ_TILE ? n : BQ_TILE) * sizeof(float4) + (size_t)nsample * (BQ_T + 1) * sizeof(int);
  if (m >= 64 && smem <= 96 * 1024) {
    cudaFuncSetAttribute(ball_query_tq_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    ball_query_tq_kernel<0><<<dim3(ceil_div(m, BQ_T), b), BQ_T, smem, (cudaStream_t)stream>>>(n, m, nsample, min_radius * min_radius,
                                                                                          max_radius * max_radius, new_xyz, xyz, idx);
    return pcreid_launch_status();
  }
  dim3 grid(ceil_div(m, 8), b);
  ball_query_kernel<0><<<grid, 256, 0, (cudaStream_t)stream>>>(n, m, nsample, min_radius * min_radius,
                                                               max_radius * max_radius, new_xyz, xyz, idx);
  return pcreid_launch_status();
}

// torch-path query_ball_point (models/pointnet2_utils.py:218-240): r2 = fp32(radius ** 2) as torch compares it.
int pcreid_query_ball_point(int b, int n, int m, float r2, int nsample, const float* new_xyz, const float* xyz, int* idx,
                            void* stream) {
  if (b <= 0 || m <= 0 || nsample <= 0) return PCREID_OK;
  if (n <= 0 || !new_xyz || !xyz || !idx) return PCREID_ERR_ARG;
  if (b > 65535) return PCREID_ERR_UNSUPPORTED;
  const size_t smem = (size_t)(n < BQ_TILE ? n : BQ_TILE) * sizeof(float4) + (size_t)nsample * (BQ_T + 1) * sizeof(int);
  if (m >= 64 && smem <= 96 * 1024) {
    cudaFuncSetAttribute(ball_query_tq_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    ball_query_tq_kernel<1><<<dim3(ceil_div(m, BQ_T), b), BQ_T, smem, (cudaStream_t)stream>>>(n, m, nsample, 0.f, r2, new_xyz, xyz, idx);
    return pcreid_launch_status();
  }
  ball_query_kernel<1><<<dim3(ceil_div(m, 8), b), 256, 0, (cudaStream_t)stream>>>(n, m, nsample, 0.f, r2, new_xyz, xyz, idx);
  return pcreid_launch_status();
}

int pcreid_group_points(int b, int c, int n, int npoints, int nsample, const float* points, const int* idx, float* out,
                        void* stream) {
  return group_launch(b, c, n, npoints * nsample, points, idx, out, (cudaStream_t)stream);
}
int pcreid_gather_points(int b, int c, int n, int npoints, const float* points, const int* idx, float* out, void* stream) {
  return group_launch(b, c, n, npoints, points, idx, out, (cudaStream_t)stream);
}

int pcreid_query_group(int b, int c, int n, int npoints, int nsample, const float* xyz, const float* center_xyz, const float* features,
                       const int* idx, int use_xyz, float divide_by, float* out, float* grouped_xyz, void* stream) {
  if (b <= 0 || npoints <= 0 || nsample <= 0) return PCREID_OK;
  if (!xyz || !center_xyz || !idx || !out || n <= 0 || (!use_xyz && !features) || (features && c <= 0)) return PCREID_ERR_ARG;
  if (b > 65535 || ceil_div(c, 16) > 65535) return PCREID_ERR_UNSUPPORTED;
  const int cy = features ? ceil_div(c, 16) : 1;
  query_group_kernel<<<dim3(ceil_div(npoints * nsample, 256), cy, b), 256, 0, (cudaStream_t)stream>>>(
      c, n, npoints, nsample, xyz, center_xyz, features, idx, use_xyz, divide_by, out, grouped_xyz);
  return pcreid_launch_status();
}

int pcreid_three_nn(int b, int n, int m, const float* unknown, const float* known, float* dist2, int* idx, void* stream) {
  if (b <= 0 || n <= 0) return PCREID_OK;
  if (!unknown || !dist2 || !idx || m < 0 || (m > 0 && !known)) return PCREID_ERR_ARG;
  if (b > 65535) return PCREID_ERR_UNSUPPORTED;
  three_nn_kernel<<<dim3(ceil_div(n, 256), b), 256, 0, (cudaStream_t)stream>>>(n, m, unknown, known, dist2, idx);
  return pcreid_launch_status();
}

int pcreid_three_interpolate(int b, int c, int m, int n, const float* points, const int* idx, const float* weight, float* out,
                             void* stream) {
  if (b <= 0 || c <= 0 || n <= 0) return PCREID_OK;
  if (!points || !idx || !weight || !out || m <= 0) return PCREID_ERR_ARG;
  if (b > 65535 || ceil_div(c, 16) > 65535) return PCREID_ERR_UNSUPPORTED;
  three_interpolate_kernel<<<dim3(ceil_div(n, 256), ceil_div(c, 16), b), 256, 0, (cudaStream_t)stream>>>(c, m, n, points, idx, weight, out);
  return pcreid_launch_status();
}

}  // extern "C"
