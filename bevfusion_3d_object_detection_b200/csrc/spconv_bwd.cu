// spconv_bwd.cu -- backward of the sparse convolution (training: BASELINE configs[2]).
//
//   forward   out[j, :]      = sum_k feats[pair_fwd[k, j], :] @ W[k]                      (W[k] is [Cin, Cout])
//   dgrad     d_feats[i, :]  = sum_k d_out[pair_bwd[k, i], :] @ W[k]^T    pair_bwd[k, pair_fwd[k, j]] = j
//   wgrad     d_W[k]         = sum_j feats[pair_fwd[k, j], :]^T (x) d_out[j, :]
//
// spconv does the same through its implicit-GEMM backward kernels (reference call site
// projects/SparseConvolution/sparse_functional.py: ConvGemmOps.implicit_gemm with the bwd pair tables); here the
// data gradient IS a forward gather-GEMM over the inverse rulebook (bevf_spconv_pair_bwd, this file) with the weights
// re-packed transposed, so it runs on the same fp32 / tcgen05 kernels as the forward, and the weight gradient is a
// tiled fp32 outer-product reduction (64 x 64 output tile per CTA, rows streamed through shared memory, split over
// row ranges, fp32 atomics into d_W).
//
// An input row feeds at most one output row per tap in every convolution geometry (out = (in + pad - k*dil) / stride is
// a function of (in, k)), so the inverse rulebook is a plain scatter without conflicts.
#include "common.cuh"

namespace {

__global__ void pair_bwd_fill_kernel(int *__restrict__ pair_bwd, long long total) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < total) pair_bwd[t] = -1;
}

__global__ void pair_bwd_scatter_kernel(const int *__restrict__ pair_fwd, int ld, int n_out_host,
                                        const int *__restrict__ n_out_dev, int kv, int *__restrict__ pair_bwd,
                                        int ld_in, int n_in) {
  const int n_out = n_out_dev ? min(*n_out_dev, ld) : n_out_host;
  const long long total = (long long)kv * n_out;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(t / n_out), j = (int)(t - (long long)k * n_out);
    const int i = __ldg(pair_fwd + (size_t)k * ld + j);
    if (i >= 0 && i < n_in) pair_bwd[(size_t)k * ld_in + i] = j;
  }
}

constexpr int WR = 16;        // rows per shared-memory step
constexpr int kWThreads = 256;

// d_W tile of (16 * MA) output channels x (16 * MB) input channels per CTA, MA x MB per thread: 64 x 64 for the wide
// layers, 16 x 16 / 32 x 32 for the narrow ones (a fixed 64 x 64 tile would do 16x the work of a 16 -> 16 layer).
// grid: (row splits, kv, cout tiles * cin tiles).  d_w layout [Cout, kv, Cin] (the spconv-2.x parameter layout).
template <int MA, int MB>
__global__ void __launch_bounds__(kWThreads)
    spconv_wgrad_f32_kernel(const float *__restrict__ feats, const float *__restrict__ d_out,
                            const int *__restrict__ pair_fwd, int ld, int n_out_host, const int *__restrict__ n_out_dev,
                            int kv, int cin, int cout, int rows_per_split, float *__restrict__ d_w) {
  constexpr int TCO = 16 * MA, TCI = 16 * MB;
  __shared__ __align__(16) float Gs[WR][TCO + 4];   // d_out rows   [r][co]
  __shared__ __align__(16) float Xs[WR][TCI + 4];   // gathered in  [r][ci]
  const int n_out = n_out_dev ? min(*n_out_dev, ld) : n_out_host;
  const int k = blockIdx.y;
  const int cin_tiles = (cin + TCI - 1) / TCI;
  const int co0 = ((int)blockIdx.z / cin_tiles) * TCO, ci0 = ((int)blockIdx.z % cin_tiles) * TCI;
  const int j_begin = blockIdx.x * rows_per_split;
  const int j_end = min(j_begin + rows_per_split, n_out);
  if (j_begin >= j_end) return;
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;          // thread tile: co = co0 + ty*MA + a, ci = ci0 + tx*MB + b
  const int lr = tid >> 4, lq = tid & 15;          // loader: row lr, MA (MB) channels from lq*MA (lq*MB)

  float acc[MA][MB];
#pragma unroll
  for (int a = 0; a < MA; ++a)
#pragma unroll
    for (int b = 0; b < MB; ++b) acc[a][b] = 0.f;
  bool touched = false;

  // software pipeline: the rulebook entry of step s + 2 and the two row segments of step s + 1 are in flight while
  // step s is multiplied (the entry -> row dependency would otherwise expose two global-memory latencies per 16 rows)
  const int *pk = pair_fwd + (size_t)k * ld;
  const bool vec_g = (cout % MA) == 0, vec_x = (cin % MB) == 0;   // aligned MA- / MB-wide segments
  auto load_rows = [&](int idx, int j, float (&g)[MA], float (&x)[MB]) {
#pragma unroll
    for (int c = 0; c < MA; ++c) g[c] = 0.f;
#pragma unroll
    for (int c = 0; c < MB; ++c) x[c] = 0.f;
    if (idx >= 0) {
      const float *gp = d_out + (size_t)j * cout + co0 + lq * MA;
      const float *xp = feats + (size_t)idx * cin + ci0 + lq * MB;
      bool done_g = false;
      if constexpr (MA == 4) {
        if (vec_g && co0 + lq * 4 + 4 <= cout) {
          const float4 t = __ldg(reinterpret_cast<const float4 *>(gp));
          g[0] = t.x; g[1] = t.y; g[2] = t.z; g[3] = t.w;
          done_g = true;
        }
      } else if constexpr (MA == 2) {
        if (vec_g && co0 + lq * 2 + 2 <= cout) {
          const float2 t = __ldg(reinterpret_cast<const float2 *>(gp));
          g[0] = t.x; g[1] = t.y;
          done_g = true;
        }
      }
      if (!done_g) {
#pragma unroll
        for (int c = 0; c < MA; ++c) if (co0 + lq * MA + c < cout) g[c] = __ldg(gp + c);
      }
      bool done_x = false;
      if constexpr (MB == 4) {
        if (vec_x && ci0 + lq * 4 + 4 <= cin) {
          const float4 t = __ldg(reinterpret_cast<const float4 *>(xp));
          x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
          done_x = true;
        }
      } else if constexpr (MB == 2) {
        if (vec_x && ci0 + lq * 2 + 2 <= cin) {
          const float2 t = __ldg(reinterpret_cast<const float2 *>(xp));
          x[0] = t.x; x[1] = t.y;
          done_x = true;
        }
      }
      if (!done_x) {
#pragma unroll
        for (int c = 0; c < MB; ++c) if (ci0 + lq * MB + c < cin) x[c] = __ldg(xp + c);
      }
    }
  };
  int idx_n = (j_begin + lr < j_end) ? __ldg(pk + j_begin + lr) : -1;
  float g_n[MA], x_n[MB];
  load_rows(idx_n, j_begin + lr, g_n, x_n);
  int idx_nn = (j_begin + WR + lr < j_end) ? __ldg(pk + j_begin + WR + lr) : -1;

  for (int j0 = j_begin; j0 < j_end; j0 += WR) {
    const int idx_c = idx_n;
    float g[MA], x[MB];
#pragma unroll
    for (int c = 0; c < MA; ++c) g[c] = g_n[c];
#pragma unroll
    for (int c = 0; c < MB; ++c) x[c] = x_n[c];
    idx_n = idx_nn;
    load_rows(idx_n, j0 + WR + lr, g_n, x_n);
    idx_nn = (j0 + 2 * WR + lr < j_end) ? __ldg(pk + j0 + 2 * WR + lr) : -1;
    // all threads are past the previous step's multiply here; skip steps whose 16 rows have no pair under this tap
    if (!__syncthreads_or(idx_c >= 0)) continue;
    touched = true;
    if constexpr (MA == 4) *reinterpret_cast<float4 *>(&Gs[lr][lq * 4]) = make_float4(g[0], g[1], g[2], g[3]);
    else {
#pragma unroll
      for (int c = 0; c < MA; ++c) Gs[lr][lq * MA + c] = g[c];
    }
    if constexpr (MB == 4) *reinterpret_cast<float4 *>(&Xs[lr][lq * 4]) = make_float4(x[0], x[1], x[2], x[3]);
    else {
#pragma unroll
      for (int c = 0; c < MB; ++c) Xs[lr][lq * MB + c] = x[c];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < WR; ++r) {
      float ga[MA], xa[MB];
      if constexpr (MA == 4) {
        const float4 t = *reinterpret_cast<const float4 *>(&Gs[r][ty * 4]);
        ga[0] = t.x; ga[1] = t.y; ga[2] = t.z; ga[3] = t.w;
      } else {
#pragma unroll
        for (int a = 0; a < MA; ++a) ga[a] = Gs[r][ty * MA + a];
      }
      if constexpr (MB == 4) {
        const float4 t = *reinterpret_cast<const float4 *>(&Xs[r][tx * 4]);
        xa[0] = t.x; xa[1] = t.y; xa[2] = t.z; xa[3] = t.w;
      } else {
#pragma unroll
        for (int b = 0; b < MB; ++b) xa[b] = Xs[r][tx * MB + b];
      }
#pragma unroll
      for (int a = 0; a < MA; ++a)
#pragma unroll
        for (int b = 0; b < MB; ++b) acc[a][b] = fmaf(ga[a], xa[b], acc[a][b]);
    }
  }
  if (!touched) return;
#pragma unroll
  for (int a = 0; a < MA; ++a) {
    const int co = co0 + ty * MA + a;
    if (co >= cout) continue;
#pragma unroll
    for (int b = 0; b < MB; ++b) {
      const int ci = ci0 + tx * MB + b;
      if (ci < cin && acc[a][b] != 0.f) atomicAdd(d_w + ((size_t)co * kv + k) * cin + ci, acc[a][b]);
    }
  }
}

template <int MA, int MB>
int launch_wgrad(const float *feats, const float *d_out, const int *pair_fwd, int ld, int n_out, const int *n_out_dev,
                 int kv, int cin, int cout, int rows, float *d_w, cudaStream_t st) {
  constexpr int TCO = 16 * MA, TCI = 16 * MB;
  // enough row splits to fill the machine a few times over, at least 512 rows each
  const int tiles = bevf::ceil_div(cout, TCO) * bevf::ceil_div(cin, TCI);
  int splits = bevf::ceil_div(4 * bevf::kNumSMs, kv * tiles);
  const int max_splits = bevf::ceil_div(rows, 512);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int rows_per_split = bevf::ceil_div(rows, splits);
  rows_per_split = bevf::ceil_div(rows_per_split, WR) * WR;
  splits = bevf::ceil_div(rows, rows_per_split);
  dim3 grid(splits, kv, tiles);
  spconv_wgrad_f32_kernel<MA, MB><<<grid, kWThreads, 0, st>>>(feats, d_out, pair_fwd, ld, n_out, n_out_dev, kv, cin,
                                                             cout, rows_per_split, d_w);
  BEVF_CHECK_LAUNCH();
  return BEVF_OK;
}

}  // namespace

BEVF_API int bevf_spconv_pair_bwd(const int *pair_fwd, int ld, int n_out, const int *n_out_dev, int kv, int *pair_bwd,
                                  int ld_in, int n_in, void *stream) {
  BEVF_CHECK_ARG(kv > 0 && ld >= n_out && n_out >= 0 && ld_in >= n_in && n_in >= 0, "bad rulebook shape");
  if (n_in == 0 || ld_in == 0) return BEVF_OK;
  BEVF_CHECK_ARG(pair_bwd && (pair_fwd || n_out == 0), "NULL rulebook");
  cudaStream_t st = (cudaStream_t)stream;
  const long long total_in = (long long)kv * ld_in;
  pair_bwd_fill_kernel<<<bevf::ceil_div(total_in, 256), 256, 0, st>>>(pair_bwd, total_in);
  BEVF_CHECK_LAUNCH();
  const int rows = n_out_dev ? ld : n_out;
  if (rows == 0) return BEVF_OK;
  const long long total = (long long)kv * rows;
  const int grid = (int)((total + 255) / 256 < 148LL * 16 ? (total + 255) / 256 : 148LL * 16);
  pair_bwd_scatter_kernel<<<grid, 256, 0, st>>>(pair_fwd, ld, n_out, n_out_dev, kv, pair_bwd, ld_in, n_in);
  BEVF_CHECK_LAUNCH();
  return BEVF_OK;
}

BEVF_API int bevf_spconv_wgrad_f32(const float *feats, const float *d_out, const int *pair_fwd, int ld, int n_out,
                                   const int *n_out_dev, int kv, int cin, int cout, float *d_weight_okc, void *stream) {
  BEVF_CHECK_ARG(kv > 0 && cin > 0 && cout > 0 && ld >= n_out && n_out >= 0, "bad wgrad shape");
  BEVF_CHECK_ARG(d_weight_okc, "NULL weight gradient");
  cudaStream_t st = (cudaStream_t)stream;
  BEVF_CHECK_CUDA(cudaMemsetAsync(d_weight_okc, 0, sizeof(float) * (size_t)cout * kv * cin, st));
  const int rows = n_out_dev ? ld : n_out;
  if (rows == 0) return BEVF_OK;
  BEVF_CHECK_ARG(feats && d_out && pair_fwd, "NULL tensor");
  const int ma = cout <= 16 ? 1 : cout <= 32 ? 2 : 4, mb = cin <= 16 ? 1 : cin <= 32 ? 2 : 4;
#define BEVF_WG_CASE(A, B) \
  if (ma == A && mb == B) return launch_wgrad<A, B>(feats, d_out, pair_fwd, ld, n_out, n_out_dev, kv, cin, cout, rows, d_weight_okc, st)
  BEVF_WG_CASE(1, 1); BEVF_WG_CASE(1, 2); BEVF_WG_CASE(1, 4);
  BEVF_WG_CASE(2, 1); BEVF_WG_CASE(2, 2); BEVF_WG_CASE(2, 4);
  BEVF_WG_CASE(4, 1); BEVF_WG_CASE(4, 2); BEVF_WG_CASE(4, 4);
#undef BEVF_WG_CASE
  return BEVF_OK;
}
