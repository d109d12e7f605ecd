// pack_out.cu -- lossless "occupied rows" form of the two BEV maps, written where the caller points: device memory or
// PINNED HOST memory (under unified addressing every cudaHostAlloc'd buffer is device-visible at its own address, so the
// kernels below store straight over the host link: no staging copy, and the number of rows -- which only the device
// knows -- never has to reach the host before the transfer starts).
//
// Why: the dense fp32 maps are 74.6 MB per frame, 79 % of what a frame moves over the host link, and that link is what
// bounds the end-to-end rate (profiles/README.md: 8 GPUs share ~190 GB/s).  But the LiDAR map is dense() of a sparse tensor
// with ~31 k active sites of 64.8 k cells (BEVFusionSparseEncoder tail, sparse_encoder.py:147-151) and the camera map is
// zero outside the cells the frustum reaches (35 % of 129.6 k; fixed per calibration: `interval_cell` of the pooling
// tables).  The rows + their coordinates are the same information in 30 MB; the host rebuilds the dense maps bit for bit.
//
//   bevf_pack_sparse_rows   header {n, c, cap, 0} | indices [cap, 4] int32 | rows [cap, c] fp32   (first n rows written)
//   bevf_pack_cells         out[ch, i] = dense[b, z*C + ch, x, y] of cell i (channel-major: coalesced on both sides)
//
// CTAs are 64 threads with few registers on purpose: they have to fit next to the persistent gather-GEMM CTAs (which
// leave ~4 K registers per SM) because they live as long as the transfer takes.
#include "common.cuh"

namespace {

constexpr int kPackThreads = 64;

__global__ void __launch_bounds__(kPackThreads)
    pack_rows_kernel(const float4 *__restrict__ feats, const int4 *__restrict__ indices, int cap, const int *__restrict__ n_dev,
                     int c4, int4 *__restrict__ hdr, int4 *__restrict__ dst_idx, float4 *__restrict__ dst_rows) {
  const int n = min(*n_dev, cap);
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nth = (long long)gridDim.x * blockDim.x;
  if (tid == 0) *hdr = make_int4(n, c4 * 4, cap, 0);
  for (long long i = tid; i < n; i += nth) dst_idx[i] = __ldg(indices + i);
  const long long total = (long long)n * c4;
  for (long long e = tid; e < total; e += nth) dst_rows[e] = __ldg(feats + e);
}

// thread -> 4 consecutive cells of one channel (16-byte store); cells are ascending, mostly consecutive along y
__global__ void __launch_bounds__(kPackThreads)
    pack_cells_kernel(const float *__restrict__ dense, const int *__restrict__ cells, int n_cells, int c, int nz, int plane,
                      float *__restrict__ dst, int pitch) {
  const int quads = (n_cells + 3) >> 2;
  const long long total = (long long)quads * c;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(t / quads), i0 = (int)(t - (long long)ch * quads) * 4;
    float v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      v[q] = 0.f;
      if (i0 + q < n_cells) {
        const int cell = __ldg(cells + i0 + q);
        const int bz = cell / plane, xy = cell - bz * plane;
        const int b = bz / nz, z = bz - b * nz;
        v[q] = __ldg(dense + ((size_t)(b * nz + z) * c + ch) * plane + xy);   // [B, nz*C, nx, ny] with channel z*C + ch
      }
    }
    *reinterpret_cast<float4 *>(dst + (size_t)ch * pitch + i0) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

}  // namespace

BEVF_API size_t bevf_pack_sparse_rows_bytes(int cap, int c) {
  if (cap < 0 || c < 0) return 0;
  return 16 + (size_t)cap * 16 + (size_t)cap * c * sizeof(float);
}

BEVF_API int bevf_pack_sparse_rows(const float *feats, const int *indices, int cap, const int *n_dev, int c, void *dst,
                                   size_t dst_bytes, int max_blocks, void *stream) {
  BEVF_CHECK_ARG(feats && indices && n_dev && dst && cap > 0 && c > 0 && c % 4 == 0, "bad arguments (c must be a multiple of 4)");
  BEVF_CHECK_ARG(dst_bytes >= bevf_pack_sparse_rows_bytes(cap, c), "destination too small: need %zu bytes",
                 bevf_pack_sparse_rows_bytes(cap, c));
  BEVF_CHECK_ARG(((reinterpret_cast<uintptr_t>(feats) | reinterpret_cast<uintptr_t>(indices) |
                   reinterpret_cast<uintptr_t>(dst)) & 15u) == 0, "feats / indices / dst must be 16-byte aligned");
  char *d = reinterpret_cast<char *>(dst);
  int grid = bevf::ceil_div((long long)cap * (c / 4), kPackThreads * 8);
  if (max_blocks > 0 && grid > max_blocks) grid = max_blocks;
  if (grid < 1) grid = 1;
  pack_rows_kernel<<<grid, kPackThreads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4 *>(feats), reinterpret_cast<const int4 *>(indices), cap, n_dev, c / 4,
      reinterpret_cast<int4 *>(d), reinterpret_cast<int4 *>(d + 16), reinterpret_cast<float4 *>(d + 16 + (size_t)cap * 16));
  BEVF_CHECK_LAUNCH();
  return BEVF_OK;
}

BEVF_API int bevf_pack_cells(const float *dense, const int *cells, int n_cells, int c, int nz, int plane, float *dst,
                             int pitch, int max_blocks, void *stream) {
  BEVF_CHECK_ARG(dense && dst && n_cells >= 0 && c > 0 && nz > 0 && plane > 0, "bad arguments");
  BEVF_CHECK_ARG(n_cells == 0 || cells, "cells is NULL");
  BEVF_CHECK_ARG(pitch >= ((n_cells + 3) & ~3) && pitch % 4 == 0 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0,
                 "pitch must be a multiple of 4 >= n_cells rounded up to 4, dst 16-byte aligned");
  if (n_cells == 0) return BEVF_OK;
  int grid = bevf::ceil_div((long long)((n_cells + 3) / 4) * c, kPackThreads * 8);
  if (max_blocks > 0 && grid > max_blocks) grid = max_blocks;
  if (grid < 1) grid = 1;
  pack_cells_kernel<<<grid, kPackThreads, 0, (cudaStream_t)stream>>>(dense, cells, n_cells, c, nz, plane, dst, pitch);
  BEVF_CHECK_LAUNCH();
  return BEVF_OK;
}
