// bn_train.cu -- train-mode BatchNorm1d over the rows of a sparse tensor, fused with what surrounds it in the encoder
// (training, BASELINE configs[2]):
//
//   forward   y = relu( (x - mean) * invstd * gamma + beta  [+ residual] )      mean / var over the n active rows
//             written as fp32 AND as the bf16 operand copy of the next sparse conv (no separate cast pass)
//   backward  dz = dy * (y > 0);  dx = gamma * invstd * (dz - mean(dz) - xhat * mean(dz * xhat));  dgamma, dbeta;
//             dx also as the bf16 operand of the conv's data / weight gradient; d_residual = dz
//
// The reference runs these as separate ops (mmdet3d/models/layers/sparse_block.py:137-154: conv -> norm -> relu ->
// conv -> norm -> += identity -> relu with torch.nn.BatchNorm1d on `.features`): per layer 2 + 1 + 1 kernels forward
// and 1 + 2 backward plus the fp32 -> bf16 casts of the tensor-core path, ~9 passes over [n, C]; here 2 + 2 launches,
// ~3.5 passes.  Statistics are accumulated in fp32 per thread / fp64 across threads (order-of-summation noise far below
// fp32 rounding of the result).  Channel counts: multiples of 4 up to 256.
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

constexpr int kBnThreads = 256;

// per-channel sums of up to two quantities over the rows: thread (r, q) owns channels 4q..4q+3 of rows r, r + R, ...
template <bool BWD>
__global__ void __launch_bounds__(kBnThreads)
    bn_reduce_kernel(const float *__restrict__ x, const float *__restrict__ dy, const float *__restrict__ y,
                     const float *__restrict__ mean, const float *__restrict__ invstd, int n, int C, int relu,
                     double *__restrict__ sums) {   // sums [2][C]: fwd (sum x, sum x^2); bwd (sum dz, sum dz * xhat)
  extern __shared__ float red[];   // [R][2][C]
  const int L = C >> 2;
  const int R = kBnThreads / L;
  const int r = threadIdx.x / L, q = threadIdx.x - r * L;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
  float4 mu = a, is = a;
  if (BWD && r < R) {
    mu = *reinterpret_cast<const float4 *>(mean + 4 * q);
    is = *reinterpret_cast<const float4 *>(invstd + 4 * q);
  }
  if (r < R) {
    for (long long i = (long long)blockIdx.x * R + r; i < n; i += (long long)gridDim.x * R) {
      const float4 v = __ldg(reinterpret_cast<const float4 *>(x + i * C) + q);
      if (!BWD) {
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        b.x = fmaf(v.x, v.x, b.x); b.y = fmaf(v.y, v.y, b.y); b.z = fmaf(v.z, v.z, b.z); b.w = fmaf(v.w, v.w, b.w);
      } else {
        float4 g = __ldg(reinterpret_cast<const float4 *>(dy + i * C) + q);
        if (relu) {
          const float4 o = __ldg(reinterpret_cast<const float4 *>(y + i * C) + q);
          g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f; g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
        }
        a.x += g.x; a.y += g.y; a.z += g.z; a.w += g.w;
        b.x = fmaf(g.x, (v.x - mu.x) * is.x, b.x); b.y = fmaf(g.y, (v.y - mu.y) * is.y, b.y);
        b.z = fmaf(g.z, (v.z - mu.z) * is.z, b.z); b.w = fmaf(g.w, (v.w - mu.w) * is.w, b.w);
      }
    }
    float *dst = red + (size_t)r * 2 * C;
    *reinterpret_cast<float4 *>(dst + 4 * q) = a;
    *reinterpret_cast<float4 *>(dst + C + 4 * q) = b;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += kBnThreads) {
    double t = 0.0;
    for (int rr = 0; rr < R; ++rr) t += (double)red[(size_t)rr * 2 * C + c];
    atomicAdd(sums + c, t);
  }
}

// mean / invstd from the sums (+ running statistics, momentum update as torch.nn.BatchNorm1d: unbiased running variance)
__global__ void bn_finalize_kernel(const double *__restrict__ sums, int n, int C, float eps, float momentum,
                                   float *__restrict__ mean, float *__restrict__ invstd, float *__restrict__ running_mean,
                                   float *__restrict__ running_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double m = sums[c] / (double)n;
  double var = sums[C + c] / (double)n - m * m;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)m;
  invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) {
    const double unbiased = n > 1 ? var * (double)n / (double)(n - 1) : var;
    running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * m);
    running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
  }
}

__device__ __forceinline__ uint2 pack_bf16x4(float a, float b, float c, float d) {
  const __nv_bfloat162 p0 = __floats2bfloat162_rn(a, b), p1 = __floats2bfloat162_rn(c, d);
  return make_uint2(*reinterpret_cast<const uint32_t *>(&p0), *reinterpret_cast<const uint32_t *>(&p1));
}

__global__ void __launch_bounds__(kBnThreads)
    bn_apply_fwd_kernel(const float *__restrict__ x, const float *__restrict__ residual, const float *__restrict__ mean,
                        const float *__restrict__ invstd, const float *__restrict__ gamma, const float *__restrict__ beta,
                        long long n4, int C, int relu, float *__restrict__ y, __nv_bfloat16 *__restrict__ y_bf16) {
  const int L = C >> 2;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(t % L);
    const float4 v = __ldg(reinterpret_cast<const float4 *>(x) + t);
    const float4 mu = *reinterpret_cast<const float4 *>(mean + 4 * q), is = *reinterpret_cast<const float4 *>(invstd + 4 * q);
    float4 g = make_float4(1.f, 1.f, 1.f, 1.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gamma) g = *reinterpret_cast<const float4 *>(gamma + 4 * q);
    if (beta) b = *reinterpret_cast<const float4 *>(beta + 4 * q);
    float4 o;
    o.x = fmaf((v.x - mu.x) * is.x, g.x, b.x); o.y = fmaf((v.y - mu.y) * is.y, g.y, b.y);
    o.z = fmaf((v.z - mu.z) * is.z, g.z, b.z); o.w = fmaf((v.w - mu.w) * is.w, g.w, b.w);
    if (residual) {
      const float4 rr = __ldg(reinterpret_cast<const float4 *>(residual) + t);
      o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
    }
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    reinterpret_cast<float4 *>(y)[t] = o;
    if (y_bf16) reinterpret_cast<uint2 *>(y_bf16)[t] = pack_bf16x4(o.x, o.y, o.z, o.w);
  }
}

__global__ void __launch_bounds__(kBnThreads)
    bn_apply_bwd_kernel(const float *__restrict__ x, const float *__restrict__ dy, const float *__restrict__ y,
                        const float *__restrict__ mean, const float *__restrict__ invstd, const float *__restrict__ gamma,
                        const double *__restrict__ sums, long long n4, int n, int C, int relu, float *__restrict__ dx,
                        __nv_bfloat16 *__restrict__ dx_bf16, float *__restrict__ d_residual) {
  const int L = C >> 2;
  const double inv_n = 1.0 / (double)n;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(t % L);
    const float4 v = __ldg(reinterpret_cast<const float4 *>(x) + t);
    float4 g = __ldg(reinterpret_cast<const float4 *>(dy) + t);
    if (relu) {
      const float4 o = __ldg(reinterpret_cast<const float4 *>(y) + t);
      g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f; g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
    }
    if (d_residual) reinterpret_cast<float4 *>(d_residual)[t] = g;
    const float4 mu = *reinterpret_cast<const float4 *>(mean + 4 * q), is = *reinterpret_cast<const float4 *>(invstd + 4 * q);
    float4 w = make_float4(1.f, 1.f, 1.f, 1.f);
    if (gamma) w = *reinterpret_cast<const float4 *>(gamma + 4 * q);
    const float m0[4] = {(float)(sums[4 * q] * inv_n), (float)(sums[4 * q + 1] * inv_n), (float)(sums[4 * q + 2] * inv_n),
                         (float)(sums[4 * q + 3] * inv_n)};
    const float m1[4] = {(float)(sums[C + 4 * q] * inv_n), (float)(sums[C + 4 * q + 1] * inv_n),
                         (float)(sums[C + 4 * q + 2] * inv_n), (float)(sums[C + 4 * q + 3] * inv_n)};
    float4 o;
    o.x = w.x * is.x * (g.x - m0[0] - (v.x - mu.x) * is.x * m1[0]);
    o.y = w.y * is.y * (g.y - m0[1] - (v.y - mu.y) * is.y * m1[1]);
    o.z = w.z * is.z * (g.z - m0[2] - (v.z - mu.z) * is.z * m1[2]);
    o.w = w.w * is.w * (g.w - m0[3] - (v.w - mu.w) * is.w * m1[3]);
    reinterpret_cast<float4 *>(dx)[t] = o;
    if (dx_bf16) reinterpret_cast<uint2 *>(dx_bf16)[t] = pack_bf16x4(o.x, o.y, o.z, o.w);
  }
}

__global__ void bn_param_grads_kernel(const double *__restrict__ sums, int C, float *__restrict__ dgamma,
                                      float *__restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (dbeta) dbeta[c] = (float)sums[c];
  if (dgamma) dgamma[c] = (float)sums[C + c];
}

int grid_for(long long items) {
  const long long want = (items + kBnThreads - 1) / kBnThreads;
  const long long cap = (long long)bevf::kNumSMs * 8;
  return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

}  // namespace

BEVF_API int bevf_bn_train_forward(const float *x, const float *residual, const float *gamma, const float *beta, int n, int c,
                                   float eps, float momentum, int relu, float *running_mean, float *running_var,
                                   float *save_mean, float *save_invstd, double *sums_ws, float *y, void *y_bf16,
                                   void *stream) {
  BEVF_CHECK_ARG(n > 0 && c > 0 && c % 4 == 0 && c <= 256, "bn_train: rows > 0 and channels a multiple of 4 up to 256 (got %d x %d)", n, c);
  BEVF_CHECK_ARG(x && y && save_mean && save_invstd && sums_ws, "NULL tensor");
  BEVF_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(residual) |
                   reinterpret_cast<uintptr_t>(y_bf16)) & 15u) == 0, "bn_train: tensors must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  BEVF_CHECK_CUDA(cudaMemsetAsync(sums_ws, 0, sizeof(double) * 2 * c, st));
  const int L = c / 4, R = kBnThreads / L;
  const size_t smem = (size_t)R * 2 * c * sizeof(float);
  int blocks = bevf::ceil_div(n, R);
  if (blocks > bevf::kNumSMs * 4) blocks = bevf::kNumSMs * 4;
  bn_reduce_kernel<false><<<blocks, kBnThreads, smem, st>>>(x, nullptr, nullptr, nullptr, nullptr, n, c, 0, sums_ws);
  BEVF_CHECK_LAUNCH();
  bn_finalize_kernel<<<bevf::ceil_div(c, 128), 128, 0, st>>>(sums_ws, n, c, eps, momentum, save_mean, save_invstd, running_mean,
                                                         running_var);
  BEVF_CHECK_LAUNCH();
  const long long n4 = (long long)n * L;
  bn_apply_fwd_kernel<<<grid_for(n4), kBnThreads, 0, st>>>(x, residual, save_mean, save_invstd, gamma, beta, n4, c, relu, y,
                                                         (__nv_bfloat16 *)y_bf16);
  BEVF_CHECK_LAUNCH();
  return BEVF_OK;
}

BEVF_API int bevf_bn_train_backward(const float *x, const float *dy, const float *y, const float *gamma, const float *save_mean,
                                    const float *save_invstd, int n, int c, int relu, double *sums_ws, float *dx, void *dx_bf16,
                                    float *d_residual, float *dgamma, float *dbeta, void *stream) {
  BEVF_CHECK_ARG(n > 0 && c > 0 && c % 4 == 0 && c <= 256, "bn_train: rows > 0 and channels a multiple of 4 up to 256 (got %d x %d)", n, c);
  BEVF_CHECK_ARG(x && dy && dx && save_mean && save_invstd && sums_ws && (!relu || y), "NULL tensor");
  cudaStream_t st = (cudaStream_t)stream;
  BEVF_CHECK_CUDA(cudaMemsetAsync(sums_ws, 0, sizeof(double) * 2 * c, st));
  const int L = c / 4, R = kBnThreads / L;
  const size_t smem = (size_t)R * 2 * c * sizeof(float);
  int blocks = bevf::ceil_div(n, R);
  if (blocks > bevf::kNumSMs * 4) blocks = bevf::kNumSMs * 4;
  bn_reduce_kernel<true><<<blocks, kBnThreads, smem, st>>>(x, dy, y, save_mean, save_invstd, n, c, relu, sums_ws);
  BEVF_CHECK_LAUNCH();
  const long long n4 = (long long)n * L;
  bn_apply_bwd_kernel<<<grid_for(n4), kBnThreads, 0, st>>>(x, dy, y, save_mean, save_invstd, gamma, sums_ws, n4, n, c, relu, dx,
                                                         (__nv_bfloat16 *)dx_bf16, d_residual);
  BEVF_CHECK_LAUNCH();
  if (dgamma || dbeta) {
    bn_param_grads_kernel<<<bevf::ceil_div(c, 128), 128, 0, st>>>(sums_ws, c, dgamma, dbeta);
    BEVF_CHECK_LAUNCH();
  }
  return BEVF_OK;
}
