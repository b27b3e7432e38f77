// Crop -> centre -> resample front-end of the ReID encoder: from one LiDAR sweep and a set of boxes
// straight to the (B, N, 3) encoder input.
//
// Reference (mmdet3d/models/trackers/deprecated/pc_utils.py:31-96):
//   interpolate_per_frame: DepthInstance3DBoxes(bboxes, origin=(.5,.5,.5)).points_in_boxes(points)  -> per-box crops in
//     point order (core/bbox/structures/depth_box3d.py:256-282 -> ops/roiaware_pool3d/src/points_in_boxes_cuda.cu:24-105),
//     padded to the longest crop, times inverse(affine(Rz(-(-yaw)), centre))   -> box-frame coordinates;
//   get_input_batch: torch.randint(high=length) per box -> gather with replacement to N points, zeros for empty boxes.
// The reference materialises a (P, B) int mask, a python list of B crops, a (B, Lmax, 3) padded batch and a
// (B, Lmax, 4) homogeneous copy.  Here: one pass writes a bit mask (1 bit per (box, point)) and per-tile counts; after
// an exclusive scan of the counts the second pass turns every requested sample rank directly into a point (binary
// search over the tile prefix, select-the-r-th-set-bit inside the tile) and writes it centred.  Nothing of size
// B x Lmax exists.
//
// In-box test, arithmetic of the reference kernel after the Depth -> LiDAR change of frame that
// DepthInstance3DBoxes.points_in_boxes applies (SASS of the reference file compiled for sm_100a):
//   point (y, -x, z); box centre (by, -bx, fl(bz - dz/2)), w = dy, l = dx, h = dz
//   cz = float(double(z') + double(h) * 0.5);  reject if double(|pz - cz|) > double(h) * 0.5
//   a = float(double(yaw) + pi/2); c = cosf(a), s = sinf(a);  sx = py - by, sy = (-px) - (-bx)
//   lx = fma(sx, c, -(sy * s)), ly = fma(sy, c, sx * s);  inside iff -l/2 < lx < l/2 and -w/2 < ly < w/2  (doubles)
#include "../../include/pcreid.h"
#include "common.cuh"

namespace {

constexpr int CT = 1024;          // points per tile = 32 mask words
constexpr int BB = 32;            // boxes per CTA of the mask pass

struct BoxC {                     // per-box constants of the in-box test
  float cx, cy, cz, c, s;
  double hl, hw, hh;
};

__device__ __forceinline__ BoxC box_consts(const float* __restrict__ b) {
  BoxC k;
  const float bx = b[0], by = b[1], bz = b[2], dx = b[3], dy = b[4], dz = b[5], yaw = b[6];
  const float zb = __fadd_rn(bz, __fmul_rn(dz, -0.5f));            // centre origin -> bottom centre (base_box3d.py:61-64)
  k.cx = by;
  k.cy = -bx;
  k.hh = (double)dz * 0.5;
  k.cz = (float)((double)zb + k.hh);
  const float a = (float)((double)yaw + 1.57079632679489661923);
  k.c = cosf(a);
  k.s = sinf(a);
  k.hl = (double)dx * 0.5;
  k.hw = (double)dy * 0.5;
  return k;
}

__device__ __forceinline__ bool in_box(const BoxC& k, float px, float py, float pz) {
  const float xl = py, yl = -px;
  if ((double)fabsf(__fsub_rn(pz, k.cz)) > k.hh) return false;
  const float sx = __fsub_rn(xl, k.cx), sy = __fsub_rn(yl, k.cy);
  const float lx = __fmaf_rn(sx, k.c, -__fmul_rn(sy, k.s));
  const float ly = __fmaf_rn(sy, k.c, __fmul_rn(sx, k.s));
  return ((double)lx > -k.hl) & ((double)lx < k.hl) & ((double)ly > -k.hw) & ((double)ly < k.hw);
}

// pass 1: mask[b][tile][32] (bit i of word w <-> point tile*1024 + 32 w + i), counts[b][tile]
__global__ void __launch_bounds__(256) crop_mask_kernel(int P, int B, int ntiles, const float* __restrict__ pts, int ps,
                                                        const float* __restrict__ boxes, uint32_t* __restrict__ mask,
                                                        int* __restrict__ counts) {
  __shared__ BoxC bc[BB];
  __shared__ int cnt[BB];
  const int tile = blockIdx.x, b0 = blockIdx.y * BB;
  const int nb = min(BB, B - b0);
  if (threadIdx.x < nb) { bc[threadIdx.x] = box_consts(boxes + (size_t)(b0 + threadIdx.x) * 7); cnt[threadIdx.x] = 0; }
  float px[4], py[4], pz[4];
  bool live[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int p = tile * CT + j * 256 + threadIdx.x;
    live[j] = p < P;
    const float* q = pts + (size_t)(live[j] ? p : 0) * ps;
    px[j] = q[0]; py[j] = q[1]; pz[j] = q[2];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = 0; i < nb; ++i) {
    const BoxC k = bc[i];
    int c = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t w = __ballot_sync(FULL_MASK, live[j] && in_box(k, px[j], py[j], pz[j]));
      if (lane == 0) {
        mask[((size_t)(b0 + i) * ntiles + tile) * 32 + j * 8 + warp] = w;
        c += __popc(w);
      }
    }
    if (lane == 0 && c) atomicAdd(&cnt[i], c);
  }
  __syncthreads();
  if (threadIdx.x < nb) counts[(size_t)(b0 + threadIdx.x) * ntiles + tile] = cnt[threadIdx.x];
}

// pass 2: out[b][n] = Rz(yaw)-frame coordinates of the sample_rank[b][n]-th in-box point (zeros for an empty box)
//   prefix[b][0..ntiles] exclusive scan of counts
__global__ void __launch_bounds__(128) crop_gather_kernel(int P, int N, int ntiles, const float* __restrict__ pts, int ps,
                                                          const float* __restrict__ boxes, const uint32_t* __restrict__ mask,
                                                          const int* __restrict__ prefix, const long long* __restrict__ rank,
                                                          float* __restrict__ out) {
  const int b = blockIdx.y, n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float* o = out + ((size_t)b * N + n) * 3;
  const int* pf = prefix + (size_t)b * (ntiles + 1);
  const int len = pf[ntiles];
  if (len <= 0) { o[0] = 0.f; o[1] = 0.f; o[2] = 0.f; return; }
  long long r = rank[(size_t)b * N + n];
  r = r < 0 ? 0 : (r >= len ? len - 1 : r);
  int lo = 0, hi = ntiles;                       // largest tile with prefix[tile] <= r
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (pf[mid] <= r) lo = mid; else hi = mid;
  }
  int rem = (int)r - pf[lo];
  const uint32_t* mw = mask + ((size_t)b * ntiles + lo) * 32;
  int word = 0;
  uint32_t w = mw[0];
  while (rem >= __popc(w)) { rem -= __popc(w); w = mw[++word]; }     // mask words are stored as [pass j][warp]: word = 8 j + warp
  for (int i = 0; i < rem; ++i) w &= w - 1;                          // drop the rem lowest set bits
  const int bit = __ffs(w) - 1;
  const int p = lo * CT + (word >> 3) * 256 + (word & 7) * 32 + bit;
  const float* q = pts + (size_t)p * ps;
  const float* bx = boxes + (size_t)b * 7;
  float s, c;
  sincosf(bx[6], &s, &c);
  const float dx = q[0] - bx[0], dy = q[1] - bx[1];
  o[0] = c * dx - s * dy;                        // inverse of [Rz(-yaw) | centre] applied to the point (pc_utils.py:62-76)
  o[1] = s * dx + c * dy;
  o[2] = q[2] - bx[2];
}

}  // namespace

extern "C" {

int pcreid_crop_tiles(int P) { return (P + CT - 1) / CT; }

int pcreid_crop_mask(int P, int B, const float* pts, int pts_stride, const float* boxes, void* mask, int* counts, void* stream) {
  if (B <= 0 || P <= 0) return PCREID_OK;
  if (!pts || !boxes || !mask || !counts || pts_stride < 3) return PCREID_ERR_ARG;
  const int ntiles = (P + CT - 1) / CT;
  if ((B + BB - 1) / BB > 65535) return PCREID_ERR_UNSUPPORTED;
  crop_mask_kernel<<<dim3(ntiles, (B + BB - 1) / BB), 256, 0, (cudaStream_t)stream>>>(P, B, ntiles, pts, pts_stride, boxes,
                                                                                     (uint32_t*)mask, counts);
  return pcreid_launch_status();
}

int pcreid_crop_gather(int P, int B, int N, const float* pts, int pts_stride, const float* boxes, const void* mask,
                       const int* prefix, const long long* rank, float* out, void* stream) {
  if (B <= 0 || N <= 0) return PCREID_OK;
  if (!pts || !boxes || !mask || !prefix || !rank || !out || pts_stride < 3 || P <= 0) return PCREID_ERR_ARG;
  if (B > 65535) return PCREID_ERR_UNSUPPORTED;
  const int ntiles = (P + CT - 1) / CT;
  crop_gather_kernel<<<dim3((N + 127) / 128, B), 128, 0, (cudaStream_t)stream>>>(P, N, ntiles, pts, pts_stride, boxes,
                                                                                 (const uint32_t*)mask, prefix, rank, out);
  return pcreid_launch_status();
}

}  // extern "C"
