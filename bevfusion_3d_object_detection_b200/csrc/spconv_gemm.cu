// spconv_gemm.cu -- fp32 gather-GEMM-scatter for sparse convolution (the parity path, FFMA).
//
//   out[j, :] = epilogue( sum_k  feats[pair_fwd[k, j], :] @ W[k] )          W[k] is [Cin, Cout]
//
// Replaces spconv's ConvGemmOps.implicit_gemm in fp32 (reference call site
// projects/SparseConvolution/sparse_functional.py:287-314; spconv's default fp32 path is SIMT FFMA too:
// SPCONV_ALLOW_TF32 is off by default).  TF32/bf16 tensor-core math cannot meet the 1e-5 parity bar, so this
// kernel stays on the fp32 pipe; the tensor-core path is spconv_tc.cu.
// The epilogue optionally fuses what follows every conv in the encoder (mmdet3d/models/layers/
// sparse_block.py:137-154): bias, eval-mode BatchNorm1d folded to scale/shift, residual add, ReLU.
#include "common.cuh"

namespace {

constexpr int TM = 64;   // output rows per CTA
constexpr int KC = 16;   // input channels per smem step
constexpr int kThreads = 256;

struct Epilogue {
  const float *bias;      // [cout] or null
  const float *scale;     // [cout] or null  (BN: gamma / sqrt(var + eps))
  const float *shift;     // [cout] or null  (BN: beta - mean * scale)
  const float *residual;  // [n_out, cout] or null
  int relu;
};

template <int BN>
__global__ void __launch_bounds__(kThreads)
    spconv_gemm_f32_kernel(const float *__restrict__ feats, const float *__restrict__ w_kio,
                           const int *__restrict__ pair_fwd, int ld, int n_out_host,
                           const int *__restrict__ n_out_dev, int kv, int cin, int cout, Epilogue ep,
                           float *__restrict__ out) {
  constexpr int CPT = BN / 16;  // columns per thread
  __shared__ __align__(16) float As[KC][TM + 4];
  __shared__ __align__(16) float Bs[KC][BN];
  __shared__ int rows_idx[TM];

  const int n_out = n_out_dev ? min(*n_out_dev, ld) : n_out_host;
  const int j0 = blockIdx.x * TM;
  if (j0 >= n_out) return;
  const int n0 = blockIdx.y * BN;
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int ar = tid & 63, aq = tid >> 6;  // gather mapping: row ar, channel quad aq

  float acc[4][CPT];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[i][c] = 0.f;

  const bool vec_a = (cin & 3) == 0;
  const bool vec_b = (cout & 3) == 0;

  for (int k = 0; k < kv; ++k) {
    int idx = -1;
    if (tid < TM && j0 + tid < n_out) idx = __ldg(pair_fwd + (size_t)k * ld + j0 + tid);
    if (tid < TM) rows_idx[tid] = idx;
    if (!__syncthreads_or(idx >= 0)) continue;  // no row of this tile uses tap k
    const int my_row = rows_idx[ar];
    const float *wk = w_kio + (size_t)k * cin * cout;
    for (int c0 = 0; c0 < cin; c0 += KC) {
      // A: 64 rows x 16 channels, stored channel-major so the micro-kernel reads 4 rows with one LDS.128
      {
        const int kk0 = aq * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (my_row >= 0) {
          const float *src = feats + (size_t)my_row * cin + c0 + kk0;
          if (vec_a && c0 + kk0 + 3 < cin) {
            v = __ldg(reinterpret_cast<const float4 *>(src));
          } else {
            if (c0 + kk0 + 0 < cin) v.x = __ldg(src + 0);
            if (c0 + kk0 + 1 < cin) v.y = __ldg(src + 1);
            if (c0 + kk0 + 2 < cin) v.z = __ldg(src + 2);
            if (c0 + kk0 + 3 < cin) v.w = __ldg(src + 3);
          }
        }
        As[kk0 + 0][ar] = v.x;
        As[kk0 + 1][ar] = v.y;
        As[kk0 + 2][ar] = v.z;
        As[kk0 + 3][ar] = v.w;
      }
      // B: 16 x BN slice of W[k]
      for (int e = tid * 4; e < KC * BN; e += kThreads * 4) {
        const int kk = e / BN, col = e % BN;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c0 + kk < cin) {
          const float *src = wk + (size_t)(c0 + kk) * cout + n0 + col;
          if (vec_b && n0 + col + 3 < cout) {
            v = __ldg(reinterpret_cast<const float4 *>(src));
          } else {
            if (n0 + col + 0 < cout) v.x = __ldg(src + 0);
            if (n0 + col + 1 < cout) v.y = __ldg(src + 1);
            if (n0 + col + 2 < cout) v.z = __ldg(src + 2);
            if (n0 + col + 3 < cout) v.w = __ldg(src + 3);
          }
        }
        *reinterpret_cast<float4 *>(&Bs[kk][col]) = v;
      }
      __syncthreads();
      const int kmax = min(KC, cin - c0);
#pragma unroll 4
      for (int kk = 0; kk < kmax; ++kk) {
        const float4 a = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
        float b[CPT];
#pragma unroll
        for (int c = 0; c < CPT; ++c) b[c] = Bs[kk][tx * CPT + c];
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          acc[0][c] = fmaf(a.x, b[c], acc[0][c]);
          acc[1][c] = fmaf(a.y, b[c], acc[1][c]);
          acc[2][c] = fmaf(a.z, b[c], acc[2][c]);
          acc[3][c] = fmaf(a.w, b[c], acc[3][c]);
        }
      }
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = j0 + ty * 4 + i;
    if (j >= n_out) continue;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int col = n0 + tx * CPT + c;
      if (col >= cout) continue;
      float v = acc[i][c];
      if (ep.bias) v += __ldg(ep.bias + col);
      if (ep.scale) v = fmaf(v, __ldg(ep.scale + col), __ldg(ep.shift + col));
      if (ep.residual) v += __ldg(ep.residual + (size_t)j * cout + col);
      if (ep.relu) v = fmaxf(v, 0.f);
      out[(size_t)j * cout + col] = v;
    }
  }
}

// weight [Cout, kv, Cin] (spconv-2.x layout flattened) -> [kv, Cin, Cout]
__global__ void pack_weight_f32_kernel(const float *__restrict__ w, float *__restrict__ out, int kv, int cin,
                                       int cout) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)kv * cin * cout;
  if (t >= total) return;
  int co = (int)(t % cout);
  long long r = t / cout;
  int ci = (int)(r % cin);
  int k = (int)(r / cin);
  out[t] = w[((size_t)co * kv + k) * cin + ci];
}

}  // namespace

BEVF_API int bevf_spconv_pack_weight_f32(const float *w_okc, float *w_kio, int kv, int cin, int cout, void *stream) {
  BEVF_CHECK_ARG(kv > 0 && cin > 0 && cout > 0, "bad weight shape");
  long long total = (long long)kv * cin * cout;
  pack_weight_f32_kernel<<<bevf::ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(w_okc, w_kio, kv, cin, cout);
  BEVF_CHECK_LAUNCH();
  return BEVF_OK;
}

BEVF_API int bevf_spconv_gemm_f32(const float *feats, const float *w_kio, const int *pair_fwd, int ld, int n_out,
                                  const int *n_out_dev, int kv, int cin, int cout, const float *bias,
                                  const float *bn_scale, const float *bn_shift, const float *residual, int relu,
                                  float *out, void *stream) {
  BEVF_CHECK_ARG(kv > 0 && cin > 0 && cout > 0, "bad conv shape kv=%d cin=%d cout=%d", kv, cin, cout);
  BEVF_CHECK_ARG((bn_scale == nullptr) == (bn_shift == nullptr), "bn_scale and bn_shift go together");
  BEVF_CHECK_ARG(ld >= n_out && n_out >= 0, "bad n_out / ld");
  const int rows = n_out_dev ? ld : n_out;
  if (rows == 0) return BEVF_OK;
  BEVF_CHECK_ARG(feats && w_kio && pair_fwd && out, "NULL tensor");
  Epilogue ep{bias, bn_scale, bn_shift, residual, relu};
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(bevf::ceil_div(rows, TM), 1);
  if (cout <= 16) {
    spconv_gemm_f32_kernel<16><<<grid, kThreads, 0, st>>>(feats, w_kio, pair_fwd, ld, n_out, n_out_dev, kv, cin, cout,
                                                         ep, out);
  } else if (cout <= 32) {
    spconv_gemm_f32_kernel<32><<<grid, kThreads, 0, st>>>(feats, w_kio, pair_fwd, ld, n_out, n_out_dev, kv, cin, cout,
                                                         ep, out);
  } else if (cout <= 64) {
    spconv_gemm_f32_kernel<64><<<grid, kThreads, 0, st>>>(feats, w_kio, pair_fwd, ld, n_out, n_out_dev, kv, cin, cout,
                                                         ep, out);
  } else {
    grid.y = bevf::ceil_div(cout, 128);
    spconv_gemm_f32_kernel<128><<<grid, kThreads, 0, st>>>(feats, w_kio, pair_fwd, ld, n_out, n_out_dev, kv, cin,
                                                          cout, ep, out);
  }
  BEVF_CHECK_LAUNCH();
  return BEVF_OK;
}
