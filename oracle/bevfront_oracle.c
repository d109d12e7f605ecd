/*
 * bevfront_oracle.c -- TEST INFRASTRUCTURE ONLY. NOT PART OF THE PRODUCT PATH.
 *
 * Plain-C, CPU restatement of the reference algorithms on the BEV front-end hot path of
 * lhn0323/BEVFUSION-3D_object_detection (paths below are relative to /root/reference).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library, and
 * only as the checker / the timed CPU arm.  The CUDA product never links or calls it.
 *
 * Parity pinning (see DESIGN.md "Oracle"):
 *   - hard voxelize   : pinned against the reference numba voxelizer
 *                       (mmdet3d/models/task_modules/voxel/voxel_generator.py:219-289), the reference's
 *                       golden vector (tests/test_models/test_task_modules/test_voxel/test_voxel_generator.py:7-20)
 *                       and the reference C++ CPU op compiled into oracle/_ref (cubic grids as shipped,
 *                       any grid with the lookup-table shape fix).
 *   - dynamic voxelize: pinned against oracle/_ref dynamic_voxelize_cpu for in-range rows; the GPU
 *                       partial-write rows restate voxelization_cuda.cu:25-61 (no CPU reference exists).
 *   - bev_pool / aux  : pinned against fixtures generated from the reference Python (depth_lss.py
 *                       bev_pool_aux/get_geometry, ops/bev_pool/bev_pool.py QuickCumsum).
 *   - dynamic scatter : parity unpinned by the reference (no CPU implementation is bound, no tests);
 *                       pinned against torch.unique + index_reduce in tests.
 *   - sparse conv     : parity unpinned by the reference (arithmetic lives in third-party spconv>=2.3 /
 *                       cumm, absent from the tree); pinned against torch.nn.functional.conv3d on
 *                       densified grids in tests.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* grid size: voxelization_cpu.cpp:120-123, voxelization_cuda.cu:257-259, voxelize.py:108-112 */
static int grid_dim(float lo, float hi, float vs) { return (int)roundf((hi - lo) / vs); }

ORACLE_API void oracle_grid_size(const float *vs, const float *rng, int *grid) {
  for (int j = 0; j < 3; ++j) grid[j] = grid_dim(rng[j], rng[3 + j], vs[j]);
}

ORACLE_API int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/*
 * Per-point voxel coordinate.  fp32 subtract, IEEE fp32 divide, floor.
 * gpu_partial != 0 : GPU contract (voxelization_cuda.cu:25-61): the kernel returns at the first failing
 *                    axis, so a row that fails on x is (-1, <untouched>, <untouched>), on y (-1,-1,<untouched>),
 *                    on z (-1,-1,-1).  "Untouched" = whatever the caller put there (zeros in voxelize.py:42).
 * gpu_partial == 0 : CPU contract (voxelization_cpu.cpp:8-43): failing rows are (-1,-1,-1).
 */
ORACLE_API void oracle_dynamic_voxelize(const float *pts, int n, int c, const float *vs, const float *rng,
                                        int *coors, int gpu_partial) {
  int grid[3];
  oracle_grid_size(vs, rng, grid);
  for (int i = 0; i < n; ++i) {
    const float *p = pts + (size_t)i * c;
    int *o = coors + (size_t)i * 3;
    int cc[3];
    int fail_axis = -1;
    for (int j = 0; j < 3; ++j) {
      volatile float q = (p[j] - rng[j]) / vs[j];
      int v = (int)floorf(q);
      if (v < 0 || v >= grid[j]) { fail_axis = j; break; }
      cc[j] = v;
    }
    if (fail_axis < 0) {
      o[0] = cc[0]; o[1] = cc[1]; o[2] = cc[2];
    } else if (gpu_partial) {
      for (int j = 0; j <= fail_axis; ++j) o[j] = -1;
    } else {
      o[0] = o[1] = o[2] = -1;
    }
  }
}

/*
 * Hard voxelization, deterministic semantics.
 * Follows voxelization_cpu.cpp:46-101 (serial walk, dense coor->voxel lookup; the table is shaped
 * [gx][gy][gz], i.e. WITH the fix for the shipped {gz,gy,gx} allocation at :129-130) which is the same
 * algorithm as voxel_generator.py:219-289 and as the GPU deterministic path
 * voxelization_cuda.cu:106-180 (rank among same-coordinate predecessors + serial first-appearance ids).
 * voxels/coors/npv must be zero-filled by the caller (voxelize.py:51-53).  Returns voxel_num.
 */
ORACLE_API int oracle_hard_voxelize(const float *pts, int n, int c, const float *vs, const float *rng,
                                    int max_points, int max_voxels, float *voxels, int *coors, int *npv) {
  int grid[3];
  oracle_grid_size(vs, rng, grid);
  size_t cells = (size_t)grid[0] * grid[1] * grid[2];
  int *table = (int *)malloc(cells * sizeof(int));
  if (!table) return -1;
  memset(table, 0xff, cells * sizeof(int));
  int voxel_num = 0;
  for (int i = 0; i < n; ++i) {
    const float *p = pts + (size_t)i * c;
    int cc[3];
    int failed = 0;
    for (int j = 0; j < 3; ++j) {
      volatile float q = (p[j] - rng[j]) / vs[j];
      int v = (int)floorf(q);
      if (v < 0 || v >= grid[j]) { failed = 1; break; }
      cc[j] = v;
    }
    if (failed) continue;
    size_t cell = ((size_t)cc[0] * grid[1] + cc[1]) * grid[2] + cc[2];
    int vid = table[cell];
    if (vid == -1) {
      vid = voxel_num;
      if (max_voxels != -1 && voxel_num >= max_voxels) continue;
      voxel_num += 1;
      table[cell] = vid;
      coors[(size_t)vid * 3 + 0] = cc[0];
      coors[(size_t)vid * 3 + 1] = cc[1];
      coors[(size_t)vid * 3 + 2] = cc[2];
    }
    int num = npv[vid];
    if (max_points == -1 || num < max_points) {
      memcpy(voxels + ((size_t)vid * max_points + num) * c, p, sizeof(float) * c);
      npv[vid] = num + 1;
    }
  }
  free(table);
  return voxel_num;
}

/* BEVFusion.voxelize mean reduce: bevfusion.py:251-253  feats.sum(dim=1) / sizes  (fp32, slot order). */
ORACLE_API void oracle_voxel_mean(const float *voxels, const int *npv, int m, int max_points, int c, float *out) {
  for (int v = 0; v < m; ++v) {
    for (int k = 0; k < c; ++k) {
      float s = 0.f;
      for (int t = 0; t < max_points; ++t) s += voxels[((size_t)v * max_points + t) * c + k];
      out[(size_t)v * c + k] = s / (float)npv[v];
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * Dynamic scatter.  scatter_points_cuda.cu:183-239:
 *   rows with any negative component -> all -1; unique rows sorted lexicographically (at::unique_dim,
 *   sorted=true); the leading (-1,..) row is stripped; coors_map[i] = row of point i (or -1);
 *   reduce_count = points per row; reduce max / sum / mean(= sum / count).
 * reduce_type: 0 SUM, 1 MEAN, 2 MAX (voxelization.h:4).
 * out_coors [n, ndim], coors_map [n], reduce_count [n], reduced [n, c] are sized for the worst case.
 * Returns M.  Summation here is in point order, fp32 (the GPU reference uses unordered atomics).
 * ------------------------------------------------------------------------------------------------ */
static int g_ndim_cmp;
static const int *g_rows_cmp;
static int cmp_rows(const void *a, const void *b) {
  const int *ra = g_rows_cmp + (size_t)(*(const int *)a) * g_ndim_cmp;
  const int *rb = g_rows_cmp + (size_t)(*(const int *)b) * g_ndim_cmp;
  for (int j = 0; j < g_ndim_cmp; ++j) {
    if (ra[j] < rb[j]) return -1;
    if (ra[j] > rb[j]) return 1;
  }
  return 0;
}

ORACLE_API int oracle_dynamic_scatter(const float *feats, const int *coors, int n, int c, int ndim,
                                      int reduce_type, float *reduced, int *out_coors, int *coors_map,
                                      int *reduce_count) {
  if (n == 0) return 0;
  int *clean = (int *)malloc((size_t)n * ndim * sizeof(int));
  int *order = (int *)malloc((size_t)n * sizeof(int));
  for (int i = 0; i < n; ++i) {
    int neg = 0;
    for (int j = 0; j < ndim; ++j) neg |= (coors[(size_t)i * ndim + j] < 0);
    for (int j = 0; j < ndim; ++j) clean[(size_t)i * ndim + j] = neg ? -1 : coors[(size_t)i * ndim + j];
    order[i] = i;
  }
  g_ndim_cmp = ndim;
  g_rows_cmp = clean;
  qsort(order, n, sizeof(int), cmp_rows);
  int m = 0;
  for (int t = 0; t < n; ++t) {
    int i = order[t];
    const int *row = clean + (size_t)i * ndim;
    if (row[0] < 0) { coors_map[i] = -1; continue; }
    int is_new = 1;
    if (m > 0) {
      is_new = 0;
      for (int j = 0; j < ndim; ++j) is_new |= (out_coors[(size_t)(m - 1) * ndim + j] != row[j]);
    }
    if (is_new) {
      memcpy(out_coors + (size_t)m * ndim, row, sizeof(int) * ndim);
      reduce_count[m] = 0;
      m += 1;
    }
    coors_map[i] = m - 1;
    reduce_count[m - 1] += 1;
  }
  for (size_t t = 0; t < (size_t)m * c; ++t) reduced[t] = (reduce_type == 2) ? -INFINITY : 0.f;
  for (int i = 0; i < n; ++i) {
    int r = coors_map[i];
    if (r < 0) continue;
    for (int k = 0; k < c; ++k) {
      float v = feats[(size_t)i * c + k];
      float *d = reduced + (size_t)r * c + k;
      if (reduce_type == 2) *d = fmaxf(*d, v); else *d += v;
    }
  }
  if (reduce_type == 1)
    for (int r = 0; r < m; ++r)
      for (int k = 0; k < c; ++k) reduced[(size_t)r * c + k] /= (float)reduce_count[r];
  free(clean);
  free(order);
  return m;
}

/* backward: scatter_points_cuda.cu:106-179, 241-308.  grad_feats is zero-filled here (:258). */
ORACLE_API void oracle_dynamic_scatter_backward(float *grad_feats, const float *grad_reduced, const float *feats,
                                                const float *reduced, const int *coors_map,
                                                const int *reduce_count, int n, int m, int c, int reduce_type) {
  memset(grad_feats, 0, (size_t)n * c * sizeof(float));
  if (n == 0 || m == 0) return;
  if (reduce_type != 2) {
    for (int i = 0; i < n; ++i) {
      int r = coors_map[i];
      if (r < 0) continue;
      for (int k = 0; k < c; ++k) {
        float g = grad_reduced[(size_t)r * c + k];
        grad_feats[(size_t)i * c + k] = (reduce_type == 1) ? g / (float)reduce_count[r] : g;
      }
    }
  } else {
    int *from = (int *)malloc((size_t)m * c * sizeof(int));
    for (size_t t = 0; t < (size_t)m * c; ++t) from[t] = n;
    for (int i = 0; i < n; ++i) {
      int r = coors_map[i];
      if (r < 0) continue;
      for (int k = 0; k < c; ++k)
        if (feats[(size_t)i * c + k] == reduced[(size_t)r * c + k] && i < from[(size_t)r * c + k])
          from[(size_t)r * c + k] = i;
    }
    for (int r = 0; r < m; ++r)
      for (int k = 0; k < c; ++k) {
        int i = from[(size_t)r * c + k];
        if (i < n) grad_feats[(size_t)i * c + k] = grad_reduced[(size_t)r * c + k];
      }
    free(from);
  }
}

/* ------------------------------------------------------------------------------------------------
 * bev_pool.  bev_pool_cuda.cu:20-42 (forward), :61-84 (backward).
 *   out[g3*d*h*w*c + g2*h*w*c + g0*w*c + g1*c + ch] = sum_{i<len} x[(start+i)*c + ch]   (fp32, in order)
 * geom row of the interval's FIRST point is used; out must be zero-filled (bev_pool.cpp:40).
 * ------------------------------------------------------------------------------------------------ */
ORACLE_API void oracle_bev_pool_forward(const float *x, const int *geom, const int *lengths, const int *starts,
                                        int n_int, int c, int b, int d, int h, int w, float *out) {
  (void)b;
#pragma omp parallel for schedule(dynamic, 64)
  for (int t = 0; t < n_int; ++t) {
    int s = starts[t], len = lengths[t];
    const int *g = geom + (size_t)s * 4;
    float *o = out + (((size_t)g[3] * d + g[2]) * h + g[0]) * (size_t)w * c + (size_t)g[1] * c;
    for (int ch = 0; ch < c; ++ch) {
      float acc = 0.f;
      for (int i = 0; i < len; ++i) acc += x[(size_t)(s + i) * c + ch];
      o[ch] = acc;
    }
  }
}

ORACLE_API void oracle_bev_pool_backward(const float *out_grad, const int *geom, const int *lengths,
                                         const int *starts, int n_int, int c, int b, int d, int h, int w,
                                         float *x_grad) {
  (void)b;
#pragma omp parallel for schedule(dynamic, 64)
  for (int t = 0; t < n_int; ++t) {
    int s = starts[t], len = lengths[t];
    const int *g = geom + (size_t)s * 4;
    const float *o = out_grad + (((size_t)g[3] * d + g[2]) * h + g[0]) * (size_t)w * c + (size_t)g[1] * c;
    for (int i = 0; i < len; ++i) memcpy(x_grad + (size_t)(s + i) * c, o, sizeof(float) * c);
  }
}

/*
 * bev_pool_aux: depth_lss.py:118-176.
 *   idx = ((geom - (bx - dx/2)) / dx).long()   -- fp32 arithmetic, TRUNCATION toward zero (:129)
 *   keep 0<=x<nx0, 0<=y<nx1, 0<=z<nx2 (:147-154)
 *   rank = x*(W*D*B) + y*(D*B) + z*B + b with D,H,W = nx2,nx0,nx1 (:158,169); sort by rank.
 * The reference argsort is unstable; this oracle fixes the order as STABLE (ties keep frustum order).
 * geom [B, per_batch, 3] fp32 flattened; outputs sized nprime: geom_out [nk,4] (x,y,z,b) i64,
 * kept[nprime] u8, ranks[nk] i64, indices[nk] i64 (positions into the kept-compacted array).
 * Returns nk.
 */
typedef struct { int64_t rank; int64_t pos; } rank_pos_t;
static int cmp_rank_pos(const void *a, const void *b) {
  const rank_pos_t *x = (const rank_pos_t *)a, *y = (const rank_pos_t *)b;
  if (x->rank != y->rank) return x->rank < y->rank ? -1 : 1;
  return x->pos < y->pos ? -1 : (x->pos > y->pos);
}

ORACLE_API int64_t oracle_bev_pool_aux(const float *geom, int64_t nprime, int B, const float *bx,
                                       const float *dx, const int64_t *nx, int64_t *geom_out,
                                       uint8_t *kept, int64_t *ranks, int64_t *indices) {
  int64_t per_b = nprime / B;
  float lo[3];
  for (int j = 0; j < 3; ++j) {
    volatile float half = dx[j] / 2.0f;
    volatile float l = bx[j] - half;
    lo[j] = l;
  }
  int64_t *q = (int64_t *)malloc((size_t)nprime * 4 * sizeof(int64_t));
  int64_t nk = 0;
  for (int64_t i = 0; i < nprime; ++i) {
    int64_t v[3];
    for (int j = 0; j < 3; ++j) {
      volatile float t = geom[i * 3 + j] - lo[j];
      volatile float u = t / dx[j];
      v[j] = (int64_t)u; /* truncation */
    }
    int ok = v[0] >= 0 && v[0] < nx[0] && v[1] >= 0 && v[1] < nx[1] && v[2] >= 0 && v[2] < nx[2];
    kept[i] = (uint8_t)ok;
    if (ok) {
      q[nk * 4 + 0] = v[0]; q[nk * 4 + 1] = v[1]; q[nk * 4 + 2] = v[2]; q[nk * 4 + 3] = i / per_b;
      nk += 1;
    }
  }
  int64_t D = nx[2], W = nx[1];
  rank_pos_t *rp = (rank_pos_t *)malloc((size_t)(nk > 0 ? nk : 1) * sizeof(rank_pos_t));
  for (int64_t t = 0; t < nk; ++t) {
    rp[t].rank = q[t * 4 + 0] * (W * D * B) + q[t * 4 + 1] * (D * B) + q[t * 4 + 2] * B + q[t * 4 + 3];
    rp[t].pos = t;
  }
  qsort(rp, (size_t)nk, sizeof(rank_pos_t), cmp_rank_pos);
  for (int64_t t = 0; t < nk; ++t) {
    ranks[t] = rp[t].rank;
    indices[t] = rp[t].pos;
    memcpy(geom_out + t * 4, q + rp[t].pos * 4, 4 * sizeof(int64_t));
  }
  free(rp);
  free(q);
  return nk;
}

/*
 * The reference's view-transform data path around bev_pool, restated end to end (used as the CPU arm and
 * as the checker for the fused CUDA kernel):
 *   depth_lss.py:723-725  x = depth[:,None] * ctx[:,:,None]  -> [BN, C, D, fH, fW] -> permute to [.., D,fH,fW,C]
 *   depth_lss.py:184-192  reshape(N', C); x[kept]; x[indices]
 *   bev_pool.py:146-172   intervals from ranks, K1, permute(0,4,1,2,3)
 *   depth_lss.py:202      cat(unbind(dim=2), 1)  ->  [B, C*nz, nx, ny]
 * Inputs: depth [BN, D, fH, fW], ctx [BN, C, fH, fW] (NCHW, as the depthnet emits them);
 * src [nk] = frustum-point index (into N') of each sorted kept point; geom4 [nk,4] (x,y,z,b) int32;
 * starts/lengths [n_int].  Output out_bczxy [B, C, nz, nx, ny] fp32 (zero-filled here).
 */
ORACLE_API void oracle_bev_pool_fused(const float *depth, const float *ctx, const int64_t *src,
                                      const int *geom4, const int *starts, const int *lengths, int n_int,
                                      int BN, int C, int D, int fH, int fW, int B, int nz, int nxx, int nyy,
                                      float *out_bczxy) {
  (void)BN;
  size_t plane = (size_t)fH * fW;
  memset(out_bczxy, 0, (size_t)B * C * nz * nxx * nyy * sizeof(float));
#pragma omp parallel for schedule(dynamic, 64)
  for (int t = 0; t < n_int; ++t) {
    int s = starts[t], len = lengths[t];
    const int *g = geom4 + (size_t)s * 4;
    for (int ch = 0; ch < C; ++ch) {
      float acc = 0.f;
      for (int i = 0; i < len; ++i) {
        int64_t p = src[s + i];
        int64_t pix = p % (int64_t)plane;
        int64_t bn = p / ((int64_t)plane * D);
        float dv = depth[p];
        float cv = ctx[((size_t)bn * C + ch) * plane + pix];
        acc += dv * cv;
      }
      out_bczxy[((((size_t)g[3] * C + ch) * nz + g[2]) * nxx + g[0]) * nyy + g[1]] = acc;
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * Sparse convolution (spconv >= 2.3 semantics restated; third-party, not in /root/reference; call
 * sites: mmdet3d/models/layers/sparse_block.py:201-217, projects/BEVFusion/bevfusion/sparse_encoder.py:131-147,
 * projects/SparseConvolution/sparse_functional.py:118-137, 287-314).
 *   indices [n,4] int32 = (b, x, y, z); spatial dims (X, Y, Z); kernel taps row-major over (kx, ky, kz),
 *   the same order as the weight's (kD,kH,kW) axes in W[Cout, kD, kH, kW, Cin].
 *   Cross-correlation: out[o] = sum_k W[:,k,:] . in[o*stride - pad + k*dil].
 *   SubM: out sites == in sites in input order, kernel centred (pad = (k/2)*dil), stride 1.
 *   Strided: out sites = every o with at least one active input under some tap; ORDER = ascending linear
 *   index ((b*OX + x)*OY + y)*OZ + z  (our canonical order; spconv's own order is an implementation detail).
 * ------------------------------------------------------------------------------------------------ */
typedef struct { int64_t *keys; int *vals; size_t cap; } hmap_t;
static uint64_t mix64(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33; return k;
}
static void hmap_init(hmap_t *h, size_t n) {
  size_t cap = 16;
  while (cap < 2 * n + 1) cap <<= 1;
  h->cap = cap;
  h->keys = (int64_t *)malloc(cap * sizeof(int64_t));
  h->vals = (int *)malloc(cap * sizeof(int));
  for (size_t i = 0; i < cap; ++i) h->keys[i] = -1;
}
static void hmap_free(hmap_t *h) { free(h->keys); free(h->vals); }
static void hmap_put(hmap_t *h, int64_t k, int v) {
  size_t s = mix64((uint64_t)k) & (h->cap - 1);
  while (h->keys[s] != -1 && h->keys[s] != k) s = (s + 1) & (h->cap - 1);
  h->keys[s] = k; h->vals[s] = v;
}
static int hmap_get(const hmap_t *h, int64_t k) {
  size_t s = mix64((uint64_t)k) & (h->cap - 1);
  while (h->keys[s] != -1) {
    if (h->keys[s] == k) return h->vals[s];
    s = (s + 1) & (h->cap - 1);
  }
  return -1;
}

ORACLE_API void oracle_spconv_out_shape(const int *shape, const int *k, const int *s, const int *p, const int *d,
                                        int *out_shape) {
  for (int j = 0; j < 3; ++j) out_shape[j] = (shape[j] + 2 * p[j] - d[j] * (k[j] - 1) - 1) / s[j] + 1;
}

static int cmp_i64(const void *a, const void *b) {
  int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
  return x < y ? -1 : (x > y);
}

/* strided conv: compute the sorted set of output sites. out_idx sized n_in*kv*4 worst case. returns n_out */
ORACLE_API int oracle_spconv_out_sites(const int *in_idx, int n_in, const int *shape, const int *k, const int *s,
                                       const int *p, const int *d, int *out_idx) {
  int os[3];
  oracle_spconv_out_shape(shape, k, s, p, d, os);
  int kv = k[0] * k[1] * k[2];
  int64_t *cand = (int64_t *)malloc((size_t)n_in * kv * sizeof(int64_t));
  size_t nc = 0;
  for (int i = 0; i < n_in; ++i) {
    const int *c = in_idx + (size_t)i * 4;
    for (int kx = 0; kx < k[0]; ++kx) {
      int ox = c[1] + p[0] - kx * d[0];
      if (ox < 0 || ox % s[0]) continue;
      ox /= s[0];
      if (ox >= os[0]) continue;
      for (int ky = 0; ky < k[1]; ++ky) {
        int oy = c[2] + p[1] - ky * d[1];
        if (oy < 0 || oy % s[1]) continue;
        oy /= s[1];
        if (oy >= os[1]) continue;
        for (int kz = 0; kz < k[2]; ++kz) {
          int oz = c[3] + p[2] - kz * d[2];
          if (oz < 0 || oz % s[2]) continue;
          oz /= s[2];
          if (oz >= os[2]) continue;
          cand[nc++] = (((int64_t)c[0] * os[0] + ox) * os[1] + oy) * os[2] + oz;
        }
      }
    }
  }
  qsort(cand, nc, sizeof(int64_t), cmp_i64);
  int n_out = 0;
  for (size_t t = 0; t < nc; ++t) {
    if (t > 0 && cand[t] == cand[t - 1]) continue;
    int64_t v = cand[t];
    int *o = out_idx + (size_t)n_out * 4;
    o[3] = (int)(v % os[2]); v /= os[2];
    o[2] = (int)(v % os[1]); v /= os[1];
    o[1] = (int)(v % os[0]); v /= os[0];
    o[0] = (int)v;
    n_out += 1;
  }
  free(cand);
  return n_out;
}

/* rulebook: pair_fwd[kv, n_out] = input row feeding output j under tap k, or -1. */
ORACLE_API void oracle_spconv_rulebook(const int *in_idx, int n_in, const int *out_idx, int n_out,
                                       const int *shape, const int *k, const int *s, const int *p, const int *d,
                                       int subm, int *pair_fwd) {
  hmap_t h;
  hmap_init(&h, (size_t)n_in);
  for (int i = 0; i < n_in; ++i) {
    const int *c = in_idx + (size_t)i * 4;
    hmap_put(&h, (((int64_t)c[0] * shape[0] + c[1]) * shape[1] + c[2]) * shape[2] + c[3], i);
  }
  int pad[3], st[3];
  for (int j = 0; j < 3; ++j) {
    pad[j] = subm ? (k[j] / 2) * d[j] : p[j];
    st[j] = subm ? 1 : s[j];
  }
  int kv = k[0] * k[1] * k[2];
  (void)kv;
#pragma omp parallel for schedule(static)
  for (int j = 0; j < n_out; ++j) {
    const int *o = out_idx + (size_t)j * 4;
    int t = 0;
    for (int kx = 0; kx < k[0]; ++kx)
      for (int ky = 0; ky < k[1]; ++ky)
        for (int kz = 0; kz < k[2]; ++kz, ++t) {
          int ix = o[1] * st[0] - pad[0] + kx * d[0];
          int iy = o[2] * st[1] - pad[1] + ky * d[1];
          int iz = o[3] * st[2] - pad[2] + kz * d[2];
          int r = -1;
          if (ix >= 0 && ix < shape[0] && iy >= 0 && iy < shape[1] && iz >= 0 && iz < shape[2])
            r = hmap_get(&h, (((int64_t)o[0] * shape[0] + ix) * shape[1] + iy) * shape[2] + iz);
          pair_fwd[(size_t)t * n_out + j] = r;
        }
  }
  hmap_free(&h);
}

/*
 * gather-GEMM-scatter: out[j, co] = bias[co] + sum_k sum_ci feats[pair[k,j], ci] * W[co, k, ci].
 * acc_double != 0 accumulates in fp64 (the checker's "true" value); otherwise fp32 in (k, ci) order.
 */
ORACLE_API void oracle_spconv_gemm(const float *feats, const float *weight, const float *bias,
                                   const int *pair_fwd, int n_out, int kv, int cin, int cout, int acc_double,
                                   float *out) {
#pragma omp parallel for schedule(dynamic, 256)
  for (int j = 0; j < n_out; ++j) {
    for (int co = 0; co < cout; ++co) {
      double accd = bias ? (double)bias[co] : 0.0;
      float accf = bias ? bias[co] : 0.f;
      for (int t = 0; t < kv; ++t) {
        int r = pair_fwd[(size_t)t * n_out + j];
        if (r < 0) continue;
        const float *f = feats + (size_t)r * cin;
        const float *w = weight + ((size_t)co * kv + t) * cin;
        if (acc_double) {
          for (int ci = 0; ci < cin; ++ci) accd += (double)f[ci] * (double)w[ci];
        } else {
          for (int ci = 0; ci < cin; ++ci) accf += f[ci] * w[ci];
        }
      }
      out[(size_t)j * cout + co] = acc_double ? (float)accd : accf;
    }
  }
}

/* SparseConvTensor.dense(): [B, C, X, Y, Z] fp32, zero-filled here. */
ORACLE_API void oracle_sparse_to_dense(const float *feats, const int *idx, int n, int c, int B, const int *shape,
                                       float *dense) {
  size_t vol = (size_t)shape[0] * shape[1] * shape[2];
  memset(dense, 0, (size_t)B * c * vol * sizeof(float));
  for (int i = 0; i < n; ++i) {
    const int *q = idx + (size_t)i * 4;
    size_t cell = ((size_t)q[1] * shape[1] + q[2]) * shape[2] + q[3];
    for (int ch = 0; ch < c; ++ch) dense[((size_t)q[0] * c + ch) * vol + cell] = feats[(size_t)i * c + ch];
  }
}

/* eval-mode BatchNorm1d (+ optional residual) + ReLU on [n, c] rows: sparse_block.py:137-154. */
ORACLE_API void oracle_bn_relu(float *x, const float *residual, int n, int c, const float *gamma,
                               const float *beta, const float *mean, const float *var, float eps, int relu) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i)
    for (int ch = 0; ch < c; ++ch) {
      float v = x[(size_t)i * c + ch];
      v = (v - mean[ch]) / sqrtf(var[ch] + eps) * gamma[ch] + beta[ch];
      if (residual) v += residual[(size_t)i * c + ch];
      if (relu && v < 0.f) v = 0.f;
      x[(size_t)i * c + ch] = v;
    }
}

/* ------------------------------------------------------------------------------------------------ *
 * LiDAR depth image + depth histogram (upstream of bev_pool)
 *   pinned against fixtures generated by running the reference's BaseDepthTransform.forward /
 *   DepthLSSTransform.get_cam_feats on CPU (tests/golden/make_golden.py depth_prep()).
 * ------------------------------------------------------------------------------------------------ */

/* Roundings as the reference's CPU run performs them (checked against tests/golden/depth_prep.npz): the batched
 * per-camera products accumulate with fma in k order, the single [3,3] x [3,N] product rounds every term.  (The file is
 * compiled with -ffp-contract=off, so a * b + c below is two roundings.) */
static float o_dot3(float a0, float a1, float a2, float x, float y, float z) {
  return fmaf(a2, z, fmaf(a1, y, a0 * x));
}
static float o_dot3_rounded(float a0, float a1, float a2, float x, float y, float z) {
  const float s = a0 * x + a1 * y;
  return s + a2 * z;
}

/* projects/BEVFusion/bevfusion/depth_lss.py:372-420, one sample.  depth [n_cams, H, W] fully written.
 * Duplicates: the sequential scatter_ (:417) leaves the LAST (camera, point) pair in nonzero order, i.e. the largest
 * point index of each pixel. */
ORACLE_API void oracle_lidar_depth_image(const float *points, int n, int c, const float *laug_t,
                                         const float *laug_inv_r, const float *l2i_all, const float *iaug_all,
                                         int n_cams, int H, int W, float *depth) {
  memset(depth, 0, sizeof(float) * (size_t)n_cams * H * W);
  for (int cam = 0; cam < n_cams; ++cam) {
    const float *l2i = l2i_all + cam * 16, *ia = iaug_all + cam * 16;
    for (int i = 0; i < n; ++i) {
      const float *p = points + (size_t)i * c;
      const float x0 = p[0] - laug_t[0], y0 = p[1] - laug_t[1], z0 = p[2] - laug_t[2];          /* :379 */
      const float x1 = o_dot3_rounded(laug_inv_r[0], laug_inv_r[1], laug_inv_r[2], x0, y0, z0);          /* :380 */
      const float y1 = o_dot3_rounded(laug_inv_r[3], laug_inv_r[4], laug_inv_r[5], x0, y0, z0);
      const float z1 = o_dot3_rounded(laug_inv_r[6], laug_inv_r[7], laug_inv_r[8], x0, y0, z0);
      float x2 = o_dot3(l2i[0], l2i[1], l2i[2], x1, y1, z1) + l2i[3];                            /* :383-384 */
      float y2 = o_dot3(l2i[4], l2i[5], l2i[6], x1, y1, z1) + l2i[7];
      const float z2 = o_dot3(l2i[8], l2i[9], l2i[10], x1, y1, z1) + l2i[11];
      const float zc = fminf(fmaxf(z2, 1e-5f), 1e5f);                                            /* :387 (dist aliases it) */
      x2 = x2 / zc;                                                                              /* :388 */
      y2 = y2 / zc;
      const float x3 = o_dot3(ia[0], ia[1], ia[2], x2, y2, zc) + ia[3];                          /* :391-392 */
      const float y3 = o_dot3(ia[4], ia[5], ia[6], x2, y2, zc) + ia[7];
      if (!(y3 < (float)H && y3 >= 0.f && x3 < (float)W && x3 >= 0.f)) continue;                 /* :399-404 */
      depth[((size_t)cam * H + (int)y3) * W + (int)x3] = zc;                                     /* :411-417 */
    }
  }
}

/* depth_lss.py:632-661.  counts / distr: [bn, fH, fW, D] */
ORACLE_API void oracle_depth_histogram(const float *depth, int bn, int H, int W, int fH, int fW, int D, float d0,
                                       float d1, float dd, float *counts, float *distr) {
  const size_t total = (size_t)bn * fH * fW * D;
  memset(counts, 0, sizeof(float) * total);
  const float dmax = (float)((double)d1 - 0.5 * (double)dd), half = (float)(0.5 * (double)dd);
  const int ph = H / fH, pw = W / fW;
  for (int cam = 0; cam < bn; ++cam)
    for (int r = 0; r < H; ++r)
      for (int q = 0; q < W; ++q) {
        const float d = depth[((size_t)cam * H + r) * W + q];
        const float t = (fminf(fmaxf(d, d0), dmax) + half - d0) / dd;
        const long long flat = ((long long)cam * fH * fW + (long long)(r / ph) * fW + q / pw) * D + (long long)t;
        if (flat >= 0 && (size_t)flat < total) counts[flat] += 1.f; /* bin D spills into the next cell's bin 0 */
      }
  for (size_t cell = 0; cell < (size_t)bn * fH * fW; ++cell) {
    float *h = counts + cell * D;
    h[0] = 0.f;
    float s = 0.f;
    for (int b = 0; b < D; ++b) s += h[b];
    for (int b = 0; b < D; ++b) distr[cell * D + b] = h[b] / (s + 1e-8f);
  }
}
