/*
 * bevfront_b200.h -- C ABI of the B200-native BEV front-end library (libbevfront_b200.so).
 *
 * Drop-in boundary for the three data-parallel operators of BEVFusion's BEV feature construction:
 * voxelization, bev_pool and sparse 3-D convolution.  Every entry point is `extern "C"`, takes plain
 * device pointers + sizes + a CUDA stream (as void*), launches asynchronously on that stream and returns
 * an int status (BEVF_OK == 0).  No torch / ATen types cross this boundary.  Each function cites the
 * reference interface (paths relative to the reference repo root) it replaces; the bindings a maintainer
 * adds on the reference side are shown in INTEGRATION.md.
 *
 * All pointers are device pointers unless a name ends in `_host`.  Tensors are dense, row-major, fp32 /
 * int32 unless stated.  Workspaces are caller-owned scratch of at least the size the matching
 * `*_workspace_bytes` call returns; they need no initialisation.
 */
#ifndef BEVFRONT_B200_H_
#define BEVFRONT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BEVF_OK 0
#define BEVF_ERR_INVALID_ARGUMENT 1
#define BEVF_ERR_CUDA 2
#define BEVF_ERR_UNSUPPORTED 3
#define BEVF_ERR_WORKSPACE 4

#define BEVF_REDUCE_SUM 0 /* voxelization.h:4  typedef enum { SUM = 0, MEAN = 1, MAX = 2 } reduce_t */
#define BEVF_REDUCE_MEAN 1
#define BEVF_REDUCE_MAX 2

/* ABI version (bumped when a signature changes) and the last error message of the calling thread. */
int bevf_abi_version(void);
const char *bevf_last_error(void);
/* compute capability of the device the library was compiled for: 100 (sm_100a). */
int bevf_compiled_arch(void);
/* number of CUDA kernels this library has launched in this process so far (bench.py's gpu_launches). */
long long bevf_launch_count(void);

/* ------------------------------------------------------------------------------------------------ *
 * Voxelization
 * ------------------------------------------------------------------------------------------------ */

/* grid_size[j] = round((range[3+j]-range[j]) / voxel_size[j]) in fp32
 * (ops/voxel/src/voxelization_cuda.cu:257-259, ops/voxel/voxelize.py:108-112).  Host-only helper. */
int bevf_voxel_grid_size(const float *voxel_size_host, const float *coors_range_host, int *grid_size_host);

/*
 * dynamic_voxelize  (ops/voxel/src/voxelization.h:83-95 -> dynamic_voxelize_gpu, voxelization_cuda.cu:485-528,
 * kernel :25-61; pybind name `dynamic_voxelize`, voxelization.cpp:8).
 * coors[N,3] int32 (x,y,z), written in place.  Out-of-range rows reproduce the reference kernel's partial
 * writes: x fails -> coors[i,0] = -1 only; y fails -> [i,0:2] = -1; z fails -> all three -1 (the caller
 * zero-fills, voxelize.py:42).  Only NDim == 3 is supported (the reference kernel hard-codes it too).
 */
int bevf_dynamic_voxelize(const float *points, int num_points, int num_features, int *coors,
                          const float *voxel_size_host, const float *coors_range_host, int ndim, void *stream);

/*
 * hard_voxelize  (voxelization.h:58-81; deterministic: hard_voxelize_gpu voxelization_cuda.cu:231-373;
 * non-deterministic: :375-483; pybind name `hard_voxelize`, voxelization.cpp:7).
 *
 * Semantics (bit-exact with the deterministic reference): voxel id = order of first appearance in point
 * order; a point's slot = its rank among the points of the same voxel in point order; points with rank >=
 * max_points and voxels with id >= max_voxels (and their points) are dropped;
 * num_points_per_voxel = number of kept points.  voxels[max_voxels,max_points,C], coors[max_voxels,3],
 * num_points_per_voxel[max_voxels] follow the reference contract: CALLER zero-fills them
 * (voxelize.py:51-53) unless `zero_fill` != 0, in which case the library writes the zero padding of the
 * first voxel_num voxels itself and the buffers may be uninitialised.
 * deterministic == 0 is accepted and returns the same (deterministic) result, which is one of the
 * outcomes the reference's racing kernels can produce.
 * voxel_num is written to *voxel_num_dev (device int32).  The launch is asynchronous; the reference's
 * `int` return value is obtained by copying that word to the host (bevf_hard_voxelize_sync does it).
 */
size_t bevf_hard_voxelize_workspace_bytes(int num_points, int max_points, int max_voxels);
int bevf_hard_voxelize(const float *points, int num_points, int num_features, float *voxels, int *coors,
                       int *num_points_per_voxel, const float *voxel_size_host, const float *coors_range_host,
                       int max_points, int max_voxels, int ndim, int deterministic, int zero_fill,
                       void *workspace, size_t workspace_bytes, int *voxel_num_dev, void *stream);
/* same, then stream-synchronises and returns voxel_num through *voxel_num_host (the reference return value). */
int bevf_hard_voxelize_sync(const float *points, int num_points, int num_features, float *voxels, int *coors,
                            int *num_points_per_voxel, const float *voxel_size_host,
                            const float *coors_range_host, int max_points, int max_voxels, int ndim,
                            int deterministic, int zero_fill, void *workspace, size_t workspace_bytes,
                            int *voxel_num_dev, void *stream, int *voxel_num_host);

/*
 * Extension: hard voxelization fused with BEVFusion.voxelize's mean reduce and batch-index
 * pad (projects/BEVFusion/bevfusion/bevfusion.py:227-255 with voxelize_reduce=True):
 *   feats[M,C] = sum over kept points / count, coords[M,4] = (batch_idx, x, y, z), sizes[M] = count.
 * Same voxel order / drop rules as bevf_hard_voxelize.  The padded [max_voxels,max_points,C] tensor is
 * never materialised.  Outputs are written at row offset *row_offset_dev (device int32, may be NULL = 0)
 * so consecutive samples of a batch append without a host sync; *voxel_num_dev receives this sample's M
 * and, when row_offset_dev is not NULL, *row_offset_dev += M afterwards.
 */
int bevf_voxelize_mean(const float *points, int num_points, int num_features, float *feats, int *coords4,
                       int *sizes, const float *voxel_size_host, const float *coors_range_host, int max_points,
                       int max_voxels, int batch_idx, void *workspace, size_t workspace_bytes,
                       int *voxel_num_dev, int *row_offset_dev, void *stream);

/*
 * dynamic_point_to_voxel_forward / _backward  (voxelization.h:107-138 -> scatter_points_cuda.cu:183-308;
 * pybind names voxelization.cpp:9-10).
 * forward: rows of coors[N,ndim] with any negative component are dropped; unique rows are emitted in
 * ascending lexicographic order (at::unique_dim sorted=true); reduced[M,C] = max / sum / mean of the
 * points of each row; coors_map[N] = row of each point or -1; reduce_count[M].  Output buffers are
 * sized for N rows; M is written to *num_out_dev.  ndim in 1..4.
 * Two-step call (the reference synchronises inside unique_dim as well):
 *   1. bevf_dynamic_scatter_extents -> extents_dev[4] = per-column (max over valid rows) + 1; copy to host
 *   2. bevf_dynamic_scatter_workspace_bytes(n, ndim, extents_host) and bevf_dynamic_scatter_forward(...)
 * The product of the extents (the coordinate bounding box) must be <= 2^34 cells.
 */
int bevf_dynamic_scatter_extents(const int *coors, int num_points, int ndim, int *extents_dev, void *stream);
size_t bevf_dynamic_scatter_workspace_bytes(int num_points, int ndim, const int *extents_host);
int bevf_dynamic_scatter_forward(const float *feats, const int *coors, int num_points, int num_features,
                                 int ndim, const int *extents_host, int reduce_type, float *reduced_feats,
                                 int *out_coors, int *coors_map, int *reduce_count, int *num_out_dev,
                                 void *workspace, size_t workspace_bytes, void *stream);
/* grad_feats[N,C] is fully written (zero where no gradient flows).  workspace: M*C int32 for MAX. */
int bevf_dynamic_scatter_backward(float *grad_feats, const float *grad_reduced_feats, const float *feats,
                                  const float *reduced_feats, const int *coors_map, const int *reduce_count,
                                  int num_points, int num_reduced, int num_features, int reduce_type,
                                  void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------ *
 * bev_pool
 * ------------------------------------------------------------------------------------------------ */

/*
 * bev_pool_forward  (ops/bev_pool/src/bev_pool.cpp:22-47 -> bev_pool_cuda.cu:20-42, launch :86-91;
 * pybind name `bev_pool_forward`, bev_pool.cpp:90).  Argument order keeps the reference's
 * (lengths before starts).
 *   out[b,d,h,w,c]: for interval t with first row s = interval_starts[t], g = geom_feats[s,:] = (x,y,z,batch):
 *   out[g3, g2, g0, g1, :] = sum_{i < interval_lengths[t]} x[s+i, :]
 * Preconditions as in the reference: rows sorted so equal-rank rows are contiguous; all indices in range.
 * `out` is fully written: cells no interval maps to are zero (the reference returns torch::zeros + K1).
 * workspace: bevf_bev_pool_workspace_bytes(n, c).
 */
size_t bevf_bev_pool_workspace_bytes(int n, int c);
int bevf_bev_pool_forward(const float *x, const int *geom_feats, const int *interval_lengths,
                          const int *interval_starts, int n, int c, int n_intervals, int b, int d, int h, int w,
                          float *out, void *workspace, size_t workspace_bytes, void *stream);
/*
 * bev_pool_backward  (bev_pool.cpp:60-87 -> bev_pool_cuda.cu:61-84, launch :93-97; pybind :92).
 *   x_grad[s+i, :] = out_grad[g3, g2, g0, g1, :]; rows not covered by any interval are zero.
 */
int bevf_bev_pool_backward(const float *out_grad, const int *geom_feats, const int *interval_lengths,
                           const int *interval_starts, int n, int c, int n_intervals, int b, int d, int h, int w,
                           float *x_grad, void *workspace, size_t workspace_bytes, void *stream);

/*
 * Fused view-transform pooling (the north-star form; replaces the tensor chain
 * depth_lss.py:723-725 outer product -> :184-192 reshape / x[kept] / x[indices] -> bev_pool.py:146-172 ->
 * depth_lss.py:202 collapse-Z, none of which is materialised):
 *   out[b, c*nz + z, x, y] = sum over the interval's points p of depth[p] * ctx[pix(p), c]
 * depth      [BN, D, fH, fW] fp32 (softmax output, frustum order: p = ((bn*D + d)*fH + h)*fW + w)
 * ctx_nhwc   [BN, fH, fW, C] fp32 (context features, channels-last)
 * src        [nk] int32: frustum index p of each kept point, sorted by rank (= kept.nonzero()[indices])
 * interval_starts [n_int+1] int32 (CSR offsets into src; last = nk)
 * interval_cell   [n_int] int32: (b*nz + z)*nx*ny + x*ny + y, strictly increasing
 * out        [B, C*nz, nx, ny] fp32, fully written (zeros where no interval maps).
 * C must be a multiple of 4 and <= 256.
 */
int bevf_bev_pool_fused_forward(const float *depth, const float *ctx_nhwc, const int *src,
                                const int *interval_starts, const int *interval_cell, int n_intervals, int nk,
                                int bn, int d, int fh, int fw, int c, int b, int nz, int nx, int ny, float *out,
                                void *stream);
/*
 * Same result from the ray-major "run" tables (see bev_pool.cu): a run = points of one (camera, depth bin, column)
 * with consecutive rows h and one BEV cell.
 *   run_p0 [n_runs] int32 frustum index of the run's first point, runs ordered by (camera, w, d, h0);
 *   run_len [n_runs]; col_run_starts [BN*fW + 1]: CSR of the runs of each pixel column (camera*fW + w);
 *   cell_run_starts [n_int+1] / cell_run_ids [n_runs]: CSR of the runs of each occupied cell
 *   (interval_cell [n_int], ascending); partial [n_runs, C] fp32 scratch.
 * Two launches; `out` fully written; deterministic (no atomics).
 */
int bevf_bev_pool_num_tiles(int b, int nz, int nx, int ny);
/* tile_starts [bevf_bev_pool_num_tiles() + 1]: first interval of every output tile (per calibration; optional). */
int bevf_bev_pool_tile_starts(const int *interval_cell, int n_intervals, int b, int nz, int nx, int ny, int *tile_starts,
                              void *stream);
int bevf_bev_pool_fused_forward_runs(const float *depth, const float *ctx_nhwc, const int *run_p0, const int *run_len,
                                     int n_runs, const int *col_run_starts, const int *cell_run_starts,
                                     const int *cell_run_ids,
                                     const int *interval_cell, const int *tile_starts, int n_intervals, int bn, int d, int fh, int fw, int c,
                                     int b, int nz, int nx, int ny, float *partial, float *out, void *stream);
/*
 * Second-generation fused forward (csrc/bev_pool_v2.cu): same result, two launches, reads depth [BN, D, fH, fW] and the
 * context features in their NATIVE NCHW layout [BN, C, fH, fW] (no channels-last pre-pass).  run_pos[q] = position of run
 * q in cell order (the inverse permutation of cell_run_ids); `partial` [n_runs, C] fp32 scratch.  Needs fW % 4 == 0,
 * C % 8 == 0, C <= 256 (BEVF_ERR_UNSUPPORTED otherwise: use the _runs form).  Replaces the data path of
 * depth_lss.py:699-725 + :179-204 + bev_pool_cuda.cu:20-42.
 */
int bevf_bev_pool_fused_forward_v2(const float *depth, const float *ctx_nchw, const int *run_p0, const int *run_len,
                                   const int *run_pos, int n_runs, const int *col_run_starts, const int *cell_run_starts,
                                   const int *interval_cell, const int *tile_starts, int n_int, int bn, int d, int fh, int fw,
                                   int c, int b, int nz, int nx, int ny, float *partial, float *out, void *stream);
/*
 * Fused backward.  Given the output gradient in channels-last form out_grad_nhwc [B*nz*nx*ny, C]
 * (row = (b*nz+z)*nx*ny + x*ny + y; bevf_nchw_to_nhwc(out_grad, ., B, C, nz*nx*ny) produces it from
 * [B, C*nz, nx, ny]):
 *   d_depth[p]      = sum_c out_grad[cell(p), c] * ctx[pix(p), c]      (0 for points not kept)
 *   d_ctx[pix, c]   = sum over kept p at that pixel of depth[p] * out_grad[cell(p), c]
 * cell_of_point [BN*D*fH*fW] int32: output row of each frustum point or -1.
 * d_depth [BN,D,fH,fW], d_ctx_nhwc [BN,fH,fW,C] are fully written.
 */
int bevf_bev_pool_fused_backward(const float *out_grad_nhwc, const float *depth, const float *ctx_nhwc,
                                 const int *cell_of_point, int bn, int d, int fh, int fw, int c, float *d_depth,
                                 float *d_ctx_nhwc, void *stream);

/* NCHW <-> NHWC transposes for the context features (tiny, tiled through shared memory). */
int bevf_nchw_to_nhwc(const float *src, float *dst, int n, int c, int hw, void *stream);
int bevf_nhwc_to_nchw(const float *src, float *dst, int n, int c, int hw, void *stream);

/*
 * Device-side construction of every pooling table from the frustum geometry: what
 * BaseViewTransform.bev_pool_aux (projects/BEVFusion/bevfusion/depth_lss.py:118-176) plus the interval construction
 * of bev_pool() (ops/bev_pool/bev_pool.py:158-166) compute with ~20 torch kernels, an int64 argsort and several host
 * round trips.  geom: [BN, D, fH, fW, 3] fp32 LiDAR-frame xyz (get_geometry's output, B = b samples of BN / b
 * cameras); bx / dx: HOST float[3] first-cell centre and cell size; nx, ny, nz cell counts.
 * Outputs (device int32; the caller sizes them for the worst case N' = BN*D*fH*fW points, n_cells = b*nz*nx*ny):
 *   cell_of_point [N']            output cell (b*nz + z)*nx*ny + x*ny + y of every frustum point, -1 = outside
 *   src [N']                      frustum indices of the kept points ordered by cell, ties in frustum order (first nk)
 *   interval_starts [min(N',n_cells)+1], interval_cell [min(N',n_cells)]     CSR over src, one interval per hit cell
 *   tile_starts [bevf_bev_pool_num_tiles()+1]
 *   run_p0, run_len, cell_run_ids [N'], col_run_starts [BN*fW+1], cell_run_starts [min(N',n_cells)+1]
 *                                 the ray-major run tables of bevf_bev_pool_fused_forward_runs
 *   counts [3]                    nk, n_intervals, n_runs -- written on the device; nothing here waits for the host
 */
size_t bevf_bev_pool_tables_workspace_bytes(long long nprime, int b, int nz, int nx, int ny);
int bevf_bev_pool_build_tables(const float *geom, int bn, int d, int fh, int fw, int b, const float *bx,
                               const float *dx, int nx, int ny, int nz, int *cell_of_point, int *src,
                               int *interval_starts, int *interval_cell, int *tile_starts, int *run_p0, int *run_len,
                               int *col_run_starts, int *cell_run_ids, int *cell_run_starts, int *counts,
                               void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------ *
 * Upstream of bev_pool: LiDAR depth image and its per-feature-cell histogram
 * ------------------------------------------------------------------------------------------------ */

/*
 * BaseDepthTransform.forward's per-sample loop (projects/BEVFusion/bevfusion/depth_lss.py:372-420) for ONE sample:
 * every point is taken back through the LiDAR augmentation (p - lidar_aug_trans, then lidar_aug_inv_rot [3,3]),
 * into each camera (lidar2image [n_cams,4,4]), divided by its clamped depth, through the image augmentation
 * (img_aug [n_cams,4,4]); points landing inside the H x W image write their camera-frame z into
 * depth [n_cams, H, W] (zero elsewhere; fully written).  Where several points hit one pixel the LARGEST point index
 * wins (a sequential scatter_; the reference's CUDA scatter_ keeps an arbitrary one).  owner_ws: n_cams*H*W int32.
 * All matrices row-major fp32 on the device.
 */
int bevf_lidar_depth_image(const float *points, int num_points, int num_features, const float *lidar_aug_trans,
                           const float *lidar_aug_inv_rot, const float *lidar2image, const float *img_aug, int n_cams,
                           int h, int w, float *depth, int *owner_ws, void *stream);
/*
 * The histogram block of DepthLSSTransform.get_cam_feats (depth_lss.py:632-661): counts [BN, fH, fW, D] = number of
 * pixels of each (h/fH) x (w/fW) cell whose depth falls in bin trunc((clamp(d, d0, d1 - dd/2) + dd/2 - d0) / dd), bin 0
 * (no depth) cleared; distr = counts / (sum + 1e-8).  (A pixel in bin D, d >= d1 - dd/2, is dropped: in the reference it
 * lands in the next cell's bin 0, which is cleared as well.)
 */
int bevf_depth_histogram(const float *depth, int bn, int h, int w, int fh, int fw, int d, float d0, float d1, float dd,
                         float *counts, float *distr, void *stream);

/* ------------------------------------------------------------------------------------------------ *
 * Sparse 3-D convolution (SubMConv3d / SparseConv3d of spconv >= 2.3, the third-party dependency the
 * reference builds its encoder from: mmdet3d/models/layers/spconv/overwrite_spconv/write_spconv2.py:21-38,
 * mmdet3d/models/layers/sparse_block.py:201-217, projects/BEVFusion/bevfusion/sparse_encoder.py:131-147;
 * functional boundary documented by projects/SparseConvolution/sparse_functional.py:70-162, 220-320).
 *
 * Row counts: every function takes the row count as a host int `n`; where a `const int *n_dev` follows it, a non-NULL
 * pointer makes the kernels use min(*n_dev, n) read on the device (n is then the buffer capacity the grid is sized
 * for), so a whole frame can be enqueued -- or captured into a CUDA graph -- without a host synchronisation.
 * Conventions: indices [n,4] int32 = (batch, x, y, z), 16-byte aligned; spatial shape (X, Y, Z); kernel taps
 * row-major over (kx, ky, kz) = the (kD,kH,kW) axes of the weight W[Cout, kD, kH, kW, Cin]; cross-correlation
 * out[o] = sum_k W[:,k,:] . in[o*stride - pad + k*dil]; SubM: out sites == in sites (same order), kernel
 * centred; strided conv: out sites in ASCENDING LINEAR ORDER ((b*OX + x)*OY + y)*OZ + z.
 * pair_fwd [kv, ld] int32: input row feeding output j under tap k, or -1 (spconv's PairFwd layout).
 * ------------------------------------------------------------------------------------------------ */

/* out = (in + 2p - d(k-1) - 1)/s + 1 per axis.  Host-only helper. */
int bevf_spconv_out_shape(const int *shape_host, const int *ksize_host, const int *stride_host,
                          const int *padding_host, const int *dilation_host, int *out_shape_host);

/*
 * Coordinate index of one sparse level: occupancy bitmap over (batch, X, Y, Z) + popcount prefix.
 * bevf_spconv_index_build marks `indices`, scans, and, if perm != NULL, writes perm[rank] = row
 * (rank = position in ascending linear order).  Pass perm for tensors whose rows are not sorted (the
 * voxelizer output); levels produced by bevf_spconv_strided_sites are sorted and need none.
 * bevf_spconv_index_error_flag returns a device int inside index_mem: 0 ok, 1 coordinate outside the
 * grid, 2 duplicate coordinate.
 */
size_t bevf_spconv_index_bytes(int batch, const int *shape_host);
int bevf_spconv_index_build(const int *indices, int n, const int *n_dev, int batch, const int *shape_host,
                            void *index_mem, size_t index_bytes, int *perm, void *stream);
const int *bevf_spconv_index_error_flag(void *index_mem, size_t index_bytes, int batch, const int *shape_host);

/* SubM rulebook: pair_fwd[k, j] = row of the input at indices[j] + (k - ksize/2) * dilation, or -1. */
int bevf_spconv_subm_rulebook(const int *indices, int n, const int *n_dev, int batch, const int *shape_host,
                              const int *ksize_host,
                              const int *dilation_host, const void *index_mem, size_t index_bytes,
                              const int *perm, int *pair_fwd, int ld, void *stream);

/*
 * Strided (regular) sparse conv, step 1: output sites.  Builds the OUTPUT level's coordinate index in
 * out_index_mem (sized with bevf_spconv_index_bytes(batch, out_shape)), writes out_indices[min(n_out,cap),4]
 * in ascending linear order and n_out to *n_out_dev.  cap = rows available in out_indices
 * (n_in * min(kv, prod(ceil(k/s))) always suffices).
 */
int bevf_spconv_strided_sites(const int *in_indices, int n_in, const int *n_in_dev, int batch,
                              const int *in_shape_host,
                              const int *ksize_host, const int *stride_host, const int *padding_host,
                              const int *dilation_host, void *out_index_mem, size_t out_index_bytes,
                              int *out_indices, int cap, int *n_out_dev, void *stream);
/* step 2: pair_fwd[k, j] = row of the input at out[j]*stride - pad + k*dil, or -1.  n_out may be given on the
 * host, or (n_out_dev != NULL) read from the device with rows >= *n_out_dev left untouched. */
/* The output sites of a CHAIN of strided convolutions (dilation 1), all levels at once, from the coordinates of the level
 * below the first one (any row order, e.g. the voxelizer's): level l's sites are the union over the level-0 sites of the
 * box each reaches on level l, and that box follows per axis from its box one level down -- so one pass over the
 * level-0 coordinates marks every level's bitmap, and the rank scans and site lists of all levels are one launch each
 * (same results as nlev calls of bevf_spconv_strided_sites, level after level).  ksizes / strides / paddings: [nlev, 3];
 * index_mems / index_bytes / out_indices / caps / n_out_devs: one entry per level (host arrays). */
int bevf_spconv_strided_sites_chain(const int *coords0, int n0, const int *n0_dev, int batch, const int *shape0, int nlev,
                                    const int *ksizes, const int *strides, const int *paddings, void *const *index_mems,
                                    const size_t *index_bytes, int *const *out_indices, const int *caps,
                                    int *const *n_out_devs, void *stream);
int bevf_spconv_strided_rulebook(const int *out_indices, int n_out, const int *n_out_dev, int batch,
                                 const int *in_shape_host, const int *ksize_host, const int *stride_host,
                                 const int *padding_host, const int *dilation_host, const void *in_index_mem,
                                 size_t in_index_bytes, const int *in_perm, int *pair_fwd, int ld, void *stream);

/* SparseConvTensor.dense(): dense[B, C, X, Y, Z] (bev_layout == 0) or the BEVFusionSparseEncoder tail
 * (sparse_encoder.py:147-151) dense().permute(0,1,4,2,3).view(B, C*Z, X, Y) (bev_layout != 0).  Fully written. */
int bevf_sparse_to_dense(const float *feats, const int *indices, int n, const int *n_dev, int c, int batch,
                         const int *shape_host, float *dense, int bev_layout, void *stream);
/* The transpose of bevf_sparse_to_dense (its backward in training): feats[i, :] = the dense tensor at site i, for
 * either layout.  feats [n, c] is fully written for the first n (or *n_dev) rows. */
/* The BEV tail of BEVFusionSparseEncoder (dense() -> permute(0, 1, 4, 2, 3) -> view(N, C*D, H, W), sparse_encoder.py:147-151)
 * driven from the output side: bev[b, ch*Z + z, x, y] for EVERY cell, zeros included (no memset of the map needed), from the
 * rows of a level whose coordinate index (index_mem: bevf_spconv_index_build / _strided_sites[_chain]) maps cell -> row,
 * i.e. rows in ascending cell order.  Same values as bevf_sparse_to_dense(..., bev_layout = 1).  c % 4 == 0. */
int bevf_sparse_to_bev_indexed(const float *feats, int n_cap, const int *n_dev, int c, int batch, const int *shape,
                               const void *index_mem, size_t index_bytes, float *bev, void *stream);
int bevf_dense_to_sparse(const float *dense, const int *indices, int n, const int *n_dev, int c, int batch,
                         const int *shape_host, int bev_layout, float *feats, void *stream);
/* Rows re-ordered by perm (rank -> row, from bevf_spconv_index_build): out_indices[r] = indices[perm[r]]; features
 * copied as fp32 (out_f32 [n, c], optional) and / or as bf16 zero-padded to cin_pad columns (out_bf16, optional). */
int bevf_spconv_permute_rows(const float *feats, const int *indices, const int *perm, int n, const int *n_dev, int c,
                             int cin_pad, float *out_f32, void *out_bf16, int *out_indices, void *stream);

/*
 * Gather-GEMM-scatter (replaces ConvGemmOps.implicit_gemm).  out[j, :] = epilogue(sum_k feats[pair_fwd[k,j], :]
 * @ W[k]) with the optional fused epilogue v = acc + bias; v = v*bn_scale + bn_shift; v += residual[j];
 * v = max(v, 0)  (what mmdet3d/models/layers/sparse_block.py:137-154 applies after every conv in eval mode).
 * n_out rows are computed; with n_out_dev != NULL the row count is read from the device (<= ld) and the host
 * n_out is only a hint (expected rows, may be 0 = unknown) for the tile-shape choice.
 *
 * fp32 (parity path, FFMA): weights repacked once to [kv, Cin, Cout] with bevf_spconv_pack_weight_f32 from the
 * spconv-2.x parameter layout [Cout, kD, kH, kW, Cin].
 */
int bevf_spconv_pack_weight_f32(const float *weight_okc, float *weight_kio, int kv, int cin, int cout, void *stream);
int bevf_spconv_gemm_f32(const float *feats, const float *weight_kio, const int *pair_fwd, int ld, int n_out,
                         const int *n_out_dev, int kv, int cin, int cout, const float *bias, const float *bn_scale,
                         const float *bn_shift, const float *residual, int relu, float *out, void *stream);
/*
 * Backward of the sparse convolution (training).  bevf_spconv_pair_bwd inverts a rulebook: pair_bwd[k, i] = j where
 * pair_fwd[k, j] = i (-1 elsewhere; [kv, ld_in], fully written) -- an input row feeds at most one output row per tap.
 * The data gradient is then the forward gather-GEMM over pair_bwd with the weights transposed
 * (d_feats = gemm(d_out, pack(W^T), pair_bwd): bevf_spconv_gemm_f32 / _bf16), which is what spconv's implicit-GEMM
 * backward computes (reference call site projects/SparseConvolution/sparse_functional.py:287-314, ConvGemmOps).
 * bevf_spconv_wgrad_f32: d_weight[Cout, kv, Cin] = sum_j d_out[j, :]^T (x) feats[pair_fwd[k, j], :] in fp32 (the
 * buffer is zeroed first; row ranges are reduced with fp32 atomics, so the last bits depend on scheduling).
 */
int bevf_spconv_pair_bwd(const int *pair_fwd, int ld, int n_out, const int *n_out_dev, int kv, int *pair_bwd, int ld_in,
                         int n_in, void *stream);
int bevf_spconv_wgrad_f32(const float *feats, const float *d_out, const int *pair_fwd, int ld, int n_out,
                          const int *n_out_dev, int kv, int cin, int cout, float *d_weight_okc, void *stream);
/*
 * Same weight gradient on the tensor cores (tcgen05.mma with both operands MN-major, fp32 accumulation in TMEM):
 * feats bf16 [n_in, cin_pad] (zero padded, bevf_spconv_cast_bf16), d_out bf16 [n_out, cout]; cin_pad and cout in
 * {16, 32, 64, 128}.  d_weight_okc fp32 [Cout, kv, Cin] is zeroed first and reduced over row ranges with fp32 atomics.
 * What spconv's implicit-GEMM backward-weight kernels compute (projects/SparseConvolution/sparse_functional.py:287-314).
 */
int bevf_spconv_wgrad_bf16(const void *feats_bf16, const void *d_out_bf16, const int *pair_fwd, int ld, int n_out, int kv,
                           int cin, int cin_pad, int cout, float *d_weight_okc, void *stream);

/*
 * bf16 tensor-core path (tcgen05.mma, fp32 accumulation in TMEM): features bf16 [n_in, cin_pad] (cin_pad =
 * bevf_spconv_tc_cin_pad(cin) in {16,32,64,128}, zero padded: bevf_spconv_cast_bf16), weights packed per tap into
 * the UMMA K-major core-matrix image (bevf_spconv_pack_weight_bf16, kv*cout*cin_pad bf16).  Cout in
 * {16,32,64,128}.  Writes fp32 (out_f32) and/or bf16 (out_bf16, feeds the next layer without a cast pass).  The
 * residual is read as fp32 [n_out, cout] (residual) or as the bf16 copy a previous layer wrote (residual_bf16).
 */
int bevf_spconv_tc_cin_pad(int cin);
int bevf_spconv_tc_supported(int cin, int cout);
/* Which gather-GEMM kernel bevf_spconv_gemm_bf16 launches: 0 = operand tiles in shared memory ("SS" tcgen05.mma)
 * everywhere; 2 = operand rows in tensor memory ("TS" form, TMA-swizzled halo, global loads where the row range does not
 * fit) for the channel pairs it is instantiated for -- 16->16, 32->32, 64->64, 128->128, 16->32, 32->64, 64->128 -- and
 * the SS kernel for the rest; 0 and 2 are bit-identical (same bf16 products, same fp32 accumulation order).
 * 1 (default) = like 2, but the narrow layers (cin_pad <= 32, cout <= 64: 8 % of the BEVFusion encoder's flops) run on
 * the register-gather kernel (mma.sync fragments loaded straight from L2, spconv_rg.cu): same products, fp32 sums in a
 * different order.  variant < 0 only queries.  Returns the previous setting; process-wide (also env BEVFRONT_TC_TS). */
int bevf_spconv_tc_variant(int variant);
/* Persistent CTAs per tensor-core gather-GEMM launch: 0 (default) = one per SM; a smaller number leaves SMs to the launches
 * of another stream (two frames in flight on half the SMs each).  Returns the previous value; a negative argument only
 * queries.  Process-wide; read at launch (i.e. at graph capture). */
int bevf_spconv_tc_max_ctas(int max_ctas);
int bevf_spconv_cast_bf16(const float *src, void *dst_bf16, int n, int cin, int cin_pad, const int *n_dev,
                          void *stream);
int bevf_spconv_pack_weight_bf16(const float *weight_okc, void *weight_packed, int kv, int cin, int cout,
                                 void *stream);
/* Bytes of the packed-weight buffer bevf_spconv_pack_weight_bf16 writes: the UMMA image (kv*cout*cin_pad bf16) and,
 * for the narrow layers (cin_pad <= 32, cout <= 64) that run on the register-gather kernel, a second image of the same
 * size in mma.sync fragment order behind it. */
long long bevf_spconv_packed_weight_bytes(int kv, int cin, int cout);
int bevf_spconv_gemm_bf16(const void *feats_bf16, int n_in, const void *weight_packed, const int *pair_fwd, int ld, int n_out,
                          const int *n_out_dev, int kv, int cin_pad, int cout, const float *bias,
                          const float *bn_scale, const float *bn_shift, const float *residual,
                          const void *residual_bf16, int relu, float *out_f32, void *out_bf16, void *stream);

/*
 * Train-mode BatchNorm1d over the n rows of a sparse tensor, fused with the ReLU / residual add around it and with the
 * bf16 operand copy the next tensor-core conv reads (training split, BASELINE configs[2]).  Reference: the separate
 * norm / relu / += identity ops of mmdet3d/models/layers/sparse_block.py:137-154 on torch.nn.BatchNorm1d (batch
 * statistics, running statistics updated with `momentum`, unbiased running variance).
 *   forward : y = relu?((x - mean) * invstd * gamma + beta [+ residual]); save_mean / save_invstd [c] for the backward;
 *             y_bf16 optional.  sums_ws: 2*c doubles of scratch.  c % 4 == 0, c <= 256.
 *   backward: dz = relu ? dy * (y > 0) : dy; dx (and optionally dx_bf16), d_residual = dz, dgamma, dbeta.
 */
int bevf_bn_train_forward(const float *x, const float *residual, const float *gamma, const float *beta, int n, int c,
                          float eps, float momentum, int relu, float *running_mean, float *running_var, float *save_mean,
                          float *save_invstd, double *sums_ws, float *y, void *y_bf16, void *stream);
int bevf_bn_train_backward(const float *x, const float *dy, const float *y, const float *gamma, const float *save_mean,
                           const float *save_invstd, int n, int c, int relu, double *sums_ws, float *dx, void *dx_bf16,
                           float *d_residual, float *dgamma, float *dbeta, void *stream);

/*
 * Lossless "occupied rows" form of the two BEV maps (extension; the reference keeps its maps on the device for the fuser,
 * bevfusion.py:300-330 -- this is for a caller that wants them on the host).  `dst` may be device memory or PINNED HOST
 * memory (device-visible at its own address under unified addressing): the kernels then store straight over the host
 * link and only the occupied rows travel.
 *   bevf_pack_sparse_rows: the active rows of a sparse tensor (BEVFusionSparseEncoder's last level before dense(),
 *     projects/BEVFusion/bevfusion/sparse_encoder.py:147-151):
 *       dst = int32 header {n, c, cap, 0} | int32 indices [cap, 4] | fp32 rows [cap, c]     (the first n rows are written)
 *     n is read on the device (min(*n_dev, cap)); c % 4 == 0; bevf_pack_sparse_rows_bytes(cap, c) = bytes of dst.
 *   bevf_pack_cells: columns of a dense [B, nz*C, nx, ny] map at the given cells (cell = (b*nz + z)*nx*ny + x*ny + y, e.g.
 *     the `interval_cell` table of bev_pool: the camera map is zero everywhere else): dst[ch, i] = value of channel ch at
 *     cells[i]; pitch (floats, multiple of 4, >= n_cells rounded up to 4) between channels.
 * max_blocks > 0 caps the grid (the CTAs are 64 threads so that they fit next to other resident kernels).
 */
size_t bevf_pack_sparse_rows_bytes(int cap, int c);
int bevf_pack_sparse_rows(const float *feats, const int *indices, int cap, const int *n_dev, int c, void *dst,
                          size_t dst_bytes, int max_blocks, void *stream);
int bevf_pack_cells(const float *dense, const int *cells, int n_cells, int c, int nz, int plane, float *dst, int pitch,
                    int max_blocks, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* BEVFRONT_B200_H_ */
