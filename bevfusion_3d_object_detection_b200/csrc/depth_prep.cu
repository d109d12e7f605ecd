// depth_prep.cu -- the step upstream of bev_pool: LiDAR points -> per-camera sparse depth image, and
// the per-feature-cell depth-bin histogram of that image.
//
// Reference: BaseDepthTransform.forward, projects/BEVFusion/bevfusion/depth_lss.py:372-420 (a per-sample Python loop
// of 3x3 matmuls, a boolean mask, nonzero, two gathers and a scatter_ over n_cams * N candidates), and the histogram
// block of DepthLSSTransform.get_cam_feats, depth_lss.py:632-661 (index arithmetic over B*N*H*W pixels + scatter_add_).
// Here: one pass over (camera, point) pairs claims pixels, one pass over pixels writes the depth; one warp per feature
// cell builds its histogram in shared memory.
//
// Arithmetic follows the reference op by op in fp32 (sub, 3-term dot products, add, clamp, IEEE divide, truncation)
// with the roundings of the reference's own CPU run, so pixel / bin assignment and the depth values are bit-identical
// to the fixtures generated from it (tests/golden/depth_prep.npz): the [3,3] x [3,N] product of the inverse
// augmentation (:380) rounds every product and sum (mm on a transposed operand), the batched per-camera products
// (:383, :391) accumulate with fma in k order (bmm).
#include "common.cuh"

namespace {

struct Mat34 {
  float m[12];  // rows 0..2 of a 4x4: [r00 r01 r02 t0 | r10 r11 r12 t1 | r20 r21 r22 t2]
};

__device__ __forceinline__ float dot3(float a0, float a1, float a2, float x, float y, float z) {
  return __fmaf_rn(a2, z, __fmaf_rn(a1, y, __fmul_rn(a0, x)));
}
__device__ __forceinline__ float dot3_rounded(float a0, float a1, float a2, float x, float y, float z) {
  return __fadd_rn(__fadd_rn(__fmul_rn(a0, x), __fmul_rn(a1, y)), __fmul_rn(a2, z));
}

struct Proj {
  float dist;  // camera-frame z AFTER the clamp: `dist` is a view of the row the reference clamps in place (:386-387)
  float py, px;
  bool on_img;
};

// depth_lss.py:379-404 for one point and one camera
__device__ __forceinline__ Proj project(const float *__restrict__ p, const float *__restrict__ laug_t,
                                        const float *__restrict__ laug_inv_r, const float *__restrict__ l2i,
                                        const float *__restrict__ iaug, int H, int W) {
  const float x0 = __fsub_rn(p[0], laug_t[0]), y0 = __fsub_rn(p[1], laug_t[1]), z0 = __fsub_rn(p[2], laug_t[2]);
  const float x1 = dot3_rounded(laug_inv_r[0], laug_inv_r[1], laug_inv_r[2], x0, y0, z0);
  const float y1 = dot3_rounded(laug_inv_r[3], laug_inv_r[4], laug_inv_r[5], x0, y0, z0);
  const float z1 = dot3_rounded(laug_inv_r[6], laug_inv_r[7], laug_inv_r[8], x0, y0, z0);
  float x2 = __fadd_rn(dot3(l2i[0], l2i[1], l2i[2], x1, y1, z1), l2i[3]);
  float y2 = __fadd_rn(dot3(l2i[4], l2i[5], l2i[6], x1, y1, z1), l2i[7]);
  const float z2 = __fadd_rn(dot3(l2i[8], l2i[9], l2i[10], x1, y1, z1), l2i[11]);
  Proj r;
  const float zc = fminf(fmaxf(z2, 1e-5f), 1e5f);
  r.dist = zc;
  x2 = __fdiv_rn(x2, zc);
  y2 = __fdiv_rn(y2, zc);
  const float x3 = __fadd_rn(dot3(iaug[0], iaug[1], iaug[2], x2, y2, zc), iaug[3]);
  const float y3 = __fadd_rn(dot3(iaug[4], iaug[5], iaug[6], x2, y2, zc), iaug[7]);
  r.py = y3;
  r.px = x3;
  r.on_img = (y3 < (float)H) && (y3 >= 0.f) && (x3 < (float)W) && (x3 >= 0.f);
  return r;
}

// pass 1: every (camera, point) pair that lands on the image claims its pixel with the point index (largest wins)
__global__ void __launch_bounds__(256)
    depth_claim_kernel(const float *__restrict__ points, int n, int c, const float *__restrict__ laug_t,
                       const float *__restrict__ laug_inv_r, const float *__restrict__ l2i,
                       const float *__restrict__ iaug, int n_cams, int H, int W, int *__restrict__ owner) {
  const long long total = (long long)n * n_cams;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int cam = (int)(t / n), i = (int)(t % n);
    const Proj r = project(points + (size_t)i * c, laug_t, laug_inv_r, l2i + cam * 16, iaug + cam * 16, H, W);
    if (!r.on_img) continue;
    const int iy = (int)r.py, ix = (int)r.px;  // .long(): truncation (values are >= 0 here)
    atomicMax(owner + ((size_t)cam * H + iy) * W + ix, i);
  }
}

// pass 2: depth[pixel] = camera-frame z of the owning point (0 where nobody landed)
__global__ void __launch_bounds__(256)
    depth_write_kernel(const float *__restrict__ points, int c, const float *__restrict__ laug_t,
                       const float *__restrict__ laug_inv_r, const float *__restrict__ l2i,
                       const float *__restrict__ iaug, int n_cams, int H, int W, const int *__restrict__ owner,
                       float *__restrict__ depth) {
  const long long total = (long long)n_cams * H * W;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int o = owner[t];
    float d = 0.f;
    if (o >= 0) {
      const int cam = (int)(t / ((long long)H * W));
      d = project(points + (size_t)o * c, laug_t, laug_inv_r, l2i + cam * 16, iaug + cam * 16, H, W).dist;
    }
    depth[t] = d;
  }
}

// one warp per feature cell: bins of its (H/fH) x (W/fW) pixels, bin 0 cleared, normalised copy
__global__ void __launch_bounds__(256)
    depth_hist_kernel(const float *__restrict__ depth, int bn, int H, int W, int fH, int fW, int D, float d0,
                      float dmax, float half_dd, float dd, float *__restrict__ counts, float *__restrict__ distr) {
  extern __shared__ int hist[];  // [warps][D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cell = blockIdx.x * (blockDim.x >> 5) + warp;
  int *h = hist + warp * D;
  for (int b = lane; b < D; b += 32) h[b] = 0;
  __syncwarp();
  const int cells = bn * fH * fW;
  if (cell < cells) {
    const int ph = H / fH, pw = W / fW;
    const int ci = cell % fW, cj = (cell / fW) % fH, cam = cell / (fW * fH);
    const float *base = depth + ((size_t)cam * H + (size_t)cj * ph) * W + (size_t)ci * pw;
    for (int e = lane; e < ph * pw; e += 32) {
      const float d = base[(size_t)(e / pw) * W + e % pw];
      // depth_lss.py:640-645: (clamp(d, d0, d1 - 0.5 dd) + 0.5 dd - d0) / dd, truncated
      const float t = __fdiv_rn(__fsub_rn(__fadd_rn(fminf(fmaxf(d, d0), dmax), half_dd), d0), dd);
      const int bin = (int)t;
      if (bin >= 0 && bin < D) atomicAdd(h + bin, 1);  // bin == D (d >= d1 - dd/2) lands in the NEXT cell's bin 0 in
    }                                                   // the reference, which is cleared there: dropped here
    __syncwarp();
    if (lane == 0) h[0] = 0;                            // depth_lss.py:655 counts_3d[..., 0] = 0
    __syncwarp();
    int s = 0;
    for (int b = lane; b < D; b += 32) s += h[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float denom = __fadd_rn((float)s, 1e-8f);
    for (int b = lane; b < D; b += 32) {
      const float cnt = (float)h[b];
      counts[(size_t)cell * D + b] = cnt;
      distr[(size_t)cell * D + b] = __fdiv_rn(cnt, denom);
    }
  }
}

}  // namespace

BEVF_API int bevf_lidar_depth_image(const float *points, int n, int c, const float *lidar_aug_trans,
                                    const float *lidar_aug_inv_rot, const float *lidar2image, const float *img_aug,
                                    int n_cams, int H, int W, float *depth, int *owner_ws, void *stream) {
  BEVF_CHECK_ARG(n >= 0 && c >= 3 && n_cams > 0 && H > 0 && W > 0, "bad sizes");
  BEVF_CHECK_ARG(depth && owner_ws && lidar_aug_trans && lidar_aug_inv_rot && lidar2image && img_aug, "NULL tensor");
  BEVF_CHECK_ARG(n == 0 || points, "points is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t pixels = (size_t)n_cams * H * W;
  BEVF_CHECK_CUDA(cudaMemsetAsync(owner_ws, 0xFF, pixels * sizeof(int), st));
  const int grid = bevf::kNumSMs * 8;
  if (n > 0) {
    depth_claim_kernel<<<grid, 256, 0, st>>>(points, n, c, lidar_aug_trans, lidar_aug_inv_rot, lidar2image, img_aug,
                                             n_cams, H, W, owner_ws);
    BEVF_CHECK_LAUNCH();
  }
  depth_write_kernel<<<grid, 256, 0, st>>>(points, c, lidar_aug_trans, lidar_aug_inv_rot, lidar2image, img_aug, n_cams,
                                           H, W, owner_ws, depth);
  BEVF_CHECK_LAUNCH();
  return BEVF_OK;
}

BEVF_API int bevf_depth_histogram(const float *depth, int bn, int H, int W, int fH, int fW, int D, float d0, float d1,
                                  float dd, float *counts, float *distr, void *stream) {
  BEVF_CHECK_ARG(bn > 0 && H > 0 && W > 0 && fH > 0 && fW > 0 && D > 0, "bad sizes");
  BEVF_CHECK_ARG(H % fH == 0 && W % fW == 0, "image %dx%d is not a multiple of the feature map %dx%d", H, W, fH, fW);
  BEVF_CHECK_ARG(depth && counts && distr, "NULL tensor");
  BEVF_CHECK_ARG(D <= 1024, "too many depth bins (%d)", D);
  const int warps = 8;
  const int cells = bn * fH * fW;
  // the scalars are formed exactly as the reference forms them: python floats (double) rounded to fp32 when they
  // meet the tensor
  const float dmax = (float)((double)d1 - 0.5 * (double)dd);
  const float half_dd = (float)(0.5 * (double)dd);
  depth_hist_kernel<<<bevf::ceil_div(cells, warps), warps * 32, (size_t)warps * D * sizeof(int), (cudaStream_t)stream>>>(
      depth, bn, H, W, fH, fW, D, d0, dmax, half_dd, dd, counts, distr);
  BEVF_CHECK_LAUNCH();
  return BEVF_OK;
}
