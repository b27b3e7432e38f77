/* TEST INFRASTRUCTURE ONLY -- plain-C CPU restatement of the five mmdet3d point ops the reference
 * ships as CUDA kernels (no CPU implementation exists upstream).  Used by tests/ and smoke() as the
 * checker, never by the product.  Build:  gcc -O2 -ffp-contract=off -shared -fPIC (oracle/Makefile).
 *
 * Each function follows the cited kernel statement-for-statement in *behaviour* (thread-serialised),
 * including its tie rules:
 *   oracle_fps            mmdet3d/ops/furthest_point_sample/src/furthest_point_sample_cuda.cu:25-141
 *   oracle_fps_with_dist  same file :213-331
 *   oracle_knn            mmdet3d/ops/knn/src/knn_cuda.cu:26-94   (max-heap, strict '<' replace, heap sort)
 *   oracle_ball_query     mmdet3d/ops/ball_query/src/ball_query_cuda.cu:11-54
 *   oracle_group_points   mmdet3d/ops/group_points/src/group_points_cuda.cu:56-79
 *   oracle_gather_points  mmdet3d/ops/gather_points/src/gather_points_cuda.cu:8-26
 *
 * Distance arithmetic: nvcc's default -fmad=true contracts  dx*dx + dy*dy + dz*dz  to
 *   fma(dz,dz, fma(dx,dx, fl(dy*dy)))   (SASS-verified for sm_100a / sm_80 with nvcc 12.9);
 * reproduced here with explicit fmaf and contraction disabled.  On the GPU box the restatement is
 * additionally cross-checked against the unmodified reference .cu compiled into oracle/_ref.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline float sqdist(float ax, float ay, float az, float bx, float by, float bz) {
  float dx = ax - bx, dy = ay - by, dz = az - bz;
  float t = dy * dy;
  t = fmaf(dx, dx, t);
  return fmaf(dz, dz, t);
}

/* opt_n_threads (furthest_point_sample_cuda.cu:11-15) */
int oracle_fps_block_size(int n) {
  int pow_2 = (int)(log((double)n) / log(2.0));
  int t = 1 << pow_2;
  if (t > 1024) t = 1024;
  if (t < 1) t = 1;
  return t;
}

static void tree_argmax(float *dists, int *dists_i, int bs) {
  /* __update(idx1, idx2): keeps the lower slot on equal values (L17-23) */
  for (int s = bs / 2; s >= 1; s >>= 1) {
    for (int tid = 0; tid < s; ++tid) {
      float v1 = dists[tid], v2 = dists[tid + s];
      int i1 = dists_i[tid], i2 = dists_i[tid + s];
      dists[tid] = v1 > v2 ? v1 : v2;       /* max(v1, v2) */
      dists_i[tid] = v2 > v1 ? i2 : i1;
    }
  }
}

static void fps_impl(int b, int n, int m, const float *dataset, float *temp, int *idxs, int with_dist) {
  if (m <= 0) return;
  int bs = oracle_fps_block_size(n);
  float *dists = (float *)malloc(sizeof(float) * bs);
  int *dists_i = (int *)malloc(sizeof(int) * bs);
  for (int bi = 0; bi < b; ++bi) {
    const float *ds = dataset + (size_t)bi * n * (with_dist ? n : 3);
    float *tp = temp + (size_t)bi * n;
    int *out = idxs + (size_t)bi * m;
    int old = 0;
    out[0] = 0;
    for (int j = 1; j < m; ++j) {
      float x1 = 0, y1 = 0, z1 = 0;
      if (!with_dist) { x1 = ds[old * 3 + 0]; y1 = ds[old * 3 + 1]; z1 = ds[old * 3 + 2]; }
      for (int tid = 0; tid < bs; ++tid) {
        int besti = 0;
        float best = -1.f;
        for (int k = tid; k < n; k += bs) {
          float d;
          if (with_dist) d = ds[(size_t)old * n + k];
          else d = sqdist(ds[k * 3 + 0], ds[k * 3 + 1], ds[k * 3 + 2], x1, y1, z1);
          float d2 = d < tp[k] ? d : tp[k];   /* min(d, temp[k]) */
          tp[k] = d2;
          besti = d2 > best ? k : besti;
          best = d2 > best ? d2 : best;
        }
        dists[tid] = best;
        dists_i[tid] = besti;
      }
      tree_argmax(dists, dists_i, bs);
      old = dists_i[0];
      out[j] = old;
    }
  }
  free(dists);
  free(dists_i);
}

void oracle_fps(int b, int n, int m, const float *xyz, float *temp, int *idx) { fps_impl(b, n, m, xyz, temp, idx, 0); }
void oracle_fps_with_dist(int b, int n, int m, const float *dist, float *temp, int *idx) { fps_impl(b, n, m, dist, temp, idx, 1); }

/* ---- calc_square_dist (ops/furthest_point_sample/utils.py:4-31) in the arithmetic of pcreid_pairwise_sqdist:
 * fma chains over the channels in ascending order, d = fma(-2, dot, |a|^2 + |b|^2); norm: sqrt(d) / c.  The reference
 * computes the same quantity with torch.sum / torch.matmul (library reduction order), so it agrees to rounding only;
 * tests/test_ops_oracle.py checks that both give the same F-FPS indices on the seeded inputs. ---- */
void oracle_pairwise_sqdist(int b, int n, int m, int c, const float *a, const float *bm, float *out, int norm) {
  for (int bi = 0; bi < b; ++bi)
    for (int i = 0; i < n; ++i) {
      const float *ai = a + ((size_t)bi * n + i) * c;
      float a2 = 0.f;
      for (int ch = 0; ch < c; ++ch) a2 = fmaf(ai[ch], ai[ch], a2);
      for (int j = 0; j < m; ++j) {
        const float *bj = bm + ((size_t)bi * m + j) * c;
        float b2 = 0.f, dot = 0.f;
        for (int ch = 0; ch < c; ++ch) { b2 = fmaf(bj[ch], bj[ch], b2); dot = fmaf(ai[ch], bj[ch], dot); }
        float d = fmaf(-2.f, dot, a2 + b2);
        if (norm) d = sqrtf(d) / (float)c;
        out[((size_t)bi * n + i) * m + j] = d;
      }
    }
}

/* ---- kNN: knn_cuda.cu:26-94 ---- */
static void reheap(float *dist, int *idx, int k) {
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

static void heap_sort(float *dist, int *idx, int k) {
  for (int i = k - 1; i > 0; --i) {
    float tf = dist[0]; dist[0] = dist[i]; dist[i] = tf;
    int ti = idx[0]; idx[0] = idx[i]; idx[i] = ti;
    reheap(dist, idx, i);
  }
}

/* xyz (b,n,3), new_xyz (b,m,3) -> idx (b,m,nsample), dist2 (b,m,nsample); nsample <= 100 */
void oracle_knn(int b, int n, int m, int nsample, const float *xyz, const float *new_xyz, int *idx, float *dist2) {
  float best_dist[100];
  int best_idx[100];
  for (int bi = 0; bi < b; ++bi)
    for (int q = 0; q < m; ++q) {
      const float *c = new_xyz + ((size_t)bi * m + q) * 3;
      const float *p = xyz + (size_t)bi * n * 3;
      for (int i = 0; i < nsample; ++i) { best_dist[i] = 1e10f; best_idx[i] = 0; }
      for (int i = 0; i < n; ++i) {
        float d2 = sqdist(c[0], c[1], c[2], p[i * 3], p[i * 3 + 1], p[i * 3 + 2]);
        if (d2 < best_dist[0]) {
          best_dist[0] = d2;
          best_idx[0] = i;
          reheap(best_dist, best_idx, nsample);
        }
      }
      heap_sort(best_dist, best_idx, nsample);
      int *oi = idx + ((size_t)bi * m + q) * nsample;
      float *od = dist2 + ((size_t)bi * m + q) * nsample;
      for (int i = 0; i < nsample; ++i) { oi[i] = best_idx[i]; od[i] = best_dist[i]; }
    }
}

/* ---- ball query: ball_query_cuda.cu:11-54; idx must be pre-zeroed by the caller (ball_query.py:41) ---- */
void oracle_ball_query(int b, int n, int m, float min_radius, float max_radius, int nsample,
                       const float *new_xyz, const float *xyz, int *idx) {
  float max_r2 = max_radius * max_radius, min_r2 = min_radius * min_radius;
  for (int bi = 0; bi < b; ++bi)
    for (int q = 0; q < m; ++q) {
      const float *c = new_xyz + ((size_t)bi * m + q) * 3;
      const float *p = xyz + (size_t)bi * n * 3;
      int *o = idx + ((size_t)bi * m + q) * nsample;
      int cnt = 0;
      for (int k = 0; k < n; ++k) {
        float d2 = sqdist(c[0], c[1], c[2], p[k * 3], p[k * 3 + 1], p[k * 3 + 2]);
        if (d2 == 0 || (d2 >= min_r2 && d2 < max_r2)) {
          if (cnt == 0) for (int l = 0; l < nsample; ++l) o[l] = k;
          o[cnt] = k;
          ++cnt;
          if (cnt >= nsample) break;
        }
      }
    }
}

/* ---- group / gather ---- */
void oracle_group_points(int b, int c, int n, int npoints, int nsample, const float *points, const int *idx, float *out) {
  for (int bi = 0; bi < b; ++bi)
    for (int ci = 0; ci < c; ++ci)
      for (int s = 0; s < npoints; ++s)
        for (int j = 0; j < nsample; ++j)
          out[(((size_t)bi * c + ci) * npoints + s) * nsample + j] =
              points[((size_t)bi * c + ci) * n + idx[((size_t)bi * npoints + s) * nsample + j]];
}

void oracle_gather_points(int b, int c, int n, int m, const float *points, const int *idx, float *out) {
  for (int bi = 0; bi < b; ++bi)
    for (int ci = 0; ci < c; ++ci)
      for (int s = 0; s < m; ++s)
        out[((size_t)bi * c + ci) * m + s] = points[((size_t)bi * c + ci) * n + idx[(size_t)bi * m + s]];
}

/* three_nn (mmdet3d/ops/interpolate/src/three_nn_cuda.cu:11-66): sequential scan, running best three in DOUBLES initialised
 * to 1e40, strict `<`; dist2 written back as float (an empty slot becomes +inf), same contracted distance as kNN. */
void oracle_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx) {
  for (int bi = 0; bi < b; ++bi)
    for (int p = 0; p < n; ++p) {
      const float *u = unknown + ((size_t)bi * n + p) * 3;
      double best1 = 1e40, best2 = 1e40, best3 = 1e40;
      int besti1 = 0, besti2 = 0, besti3 = 0;
      for (int k = 0; k < m; ++k) {
        const float *q = known + ((size_t)bi * m + k) * 3;
        float d = sqdist(u[0], u[1], u[2], q[0], q[1], q[2]);
        if (d < best1) { best3 = best2; besti3 = besti2; best2 = best1; besti2 = besti1; best1 = d; besti1 = k; }
        else if (d < best2) { best3 = best2; besti3 = besti2; best2 = d; besti2 = k; }
        else if (d < best3) { best3 = d; besti3 = k; }
      }
      float *od = dist2 + ((size_t)bi * n + p) * 3;
      int *oi = idx + ((size_t)bi * n + p) * 3;
      od[0] = (float)best1; od[1] = (float)best2; od[2] = (float)best3;
      oi[0] = besti1; oi[1] = besti2; oi[2] = besti3;
    }
}

/* three_interpolate (three_interpolate_cuda.cu:11-38): w0*p0 + w1*p1 + w2*p2, contracted by nvcc to
 * fma(w2,p2, fma(w0,p0, fl(w1*p1))) (SASS of the reference file built for sm_100a with nvcc 12.9). */
void oracle_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx, const float *weight, float *out) {
  for (int bi = 0; bi < b; ++bi)
    for (int ci = 0; ci < c; ++ci)
      for (int p = 0; p < n; ++p) {
        const int *ip = idx + ((size_t)bi * n + p) * 3;
        const float *w = weight + ((size_t)bi * n + p) * 3;
        const float *pr = points + ((size_t)bi * c + ci) * m;
        float t = w[1] * pr[ip[1]];
        t = fmaf(w[0], pr[ip[0]], t);
        out[((size_t)bi * c + ci) * n + p] = fmaf(w[2], pr[ip[2]], t);
      }
}

/* ---- canonical arithmetic of the torch-path distances (what the CUDA kNN kernels reproduce) ----
 * square_distance (models/pointnet2_utils.py:169-188) as torch/MKL evaluates it on the CPU oracle:
 *   m = fma chain over (x,y,z); |v|^2 = (x*x + y*y) + z*z; d = (-2*m + |q|^2) + |p|^2             */
void oracle_sqdist_expand(int b, int n, int m, const float *xyz, const float *new_xyz, float *out) {
  for (int bi = 0; bi < b; ++bi)
    for (int q = 0; q < m; ++q) {
      const float *c = new_xyz + ((size_t)bi * m + q) * 3;
      float qn = (c[0] * c[0] + c[1] * c[1]) + c[2] * c[2];
      for (int i = 0; i < n; ++i) {
        const float *p = xyz + ((size_t)bi * n + i) * 3;
        float mm = c[0] * p[0];
        mm = fmaf(c[1], p[1], mm);
        mm = fmaf(c[2], p[2], mm);
        float pn = (p[0] * p[0] + p[1] * p[1]) + p[2] * p[2];
        out[((size_t)bi * m + q) * n + i] = (-2.f * mm + qn) + pn;
      }
    }
}

/* DGCNN pairwise "distance" (models/dgcnn_orig.py:22-25): x (b,c,n);
 *   m_ij sequential fma chain over channels; xx_i = cascade sum (blocks of 16 channels, block sums in order);
 *   pd_ij = ((-xx_j) - (-2 m_ij)) - xx_i                                                           */
void oracle_dgcnn_pd(int b, int c, int n, const float *x, float *out) {
  float *xx = (float *)malloc(sizeof(float) * n);
  for (int bi = 0; bi < b; ++bi) {
    const float *xb = x + (size_t)bi * c * n;
    for (int j = 0; j < n; ++j) {
      float tot = 0.f;
      for (int c0 = 0; c0 < c; c0 += 16) {
        float blk = 0.f;
        for (int cc = c0; cc < c0 + 16 && cc < c; ++cc) { float v = xb[(size_t)cc * n + j]; blk = blk + v * v; }
        tot = c0 == 0 ? blk : tot + blk;
      }
      xx[j] = tot;
    }
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        float acc = 0.f;
        for (int cc = 0; cc < c; ++cc) acc = fmaf(xb[(size_t)cc * n + i], xb[(size_t)cc * n + j], acc);
        float inner = -2.f * acc;
        out[((size_t)bi * n + i) * n + j] = ((-xx[j]) - inner) - xx[i];   /* xx is (B,1,N): broadcasts over rows first */
      }
  }
  free(xx);
}
