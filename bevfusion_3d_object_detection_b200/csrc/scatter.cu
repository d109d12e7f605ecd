// scatter.cu -- dynamic point-to-voxel scatter (max / sum / mean) for sm_100a.
//
// Replaces projects/BEVFusion/bevfusion/ops/voxel/src/scatter_points_cuda.cu (reference), whose forward is
// at::unique_dim (a full radix sort of the N coordinate rows) followed by per-feature atomics.  Here the
// sorted-unique step is a rank query on an occupancy bitmap of the (per-call) coordinate bounding box:
//   K0 extents     per-column max over valid rows (host reads 4 ints: the reference syncs here too, :208)
//   K1 mark        key = lexicographic linear index; atomicOr into the bitmap
//   K2a/K2b        popcount scan of the bitmap words  ->  rank of every occupied key == its row in the
//                  ascending-lexicographic output (what unique_dim(sorted=true) returns)
//   K3 emit        out_coors[rank] = decode(key)
//   K4 reduce      coors_map, reduce_count, feature reduction with native float atomics
//                  (max uses the ordered-int trick instead of a CAS loop)
#include "bitmap_rank.cuh"
#include "common.cuh"

namespace {

struct Extents {
  int e[4];
  int ndim;
};

__device__ __forceinline__ bool row_key(const int *__restrict__ row, const Extents &E, unsigned long long &key) {
  unsigned long long k = 0;
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (j < E.ndim) {
      int v = row[j];
      ok = ok && (v >= 0);
      k = k * (unsigned long long)E.e[j] + (unsigned long long)(unsigned)v;
    }
  }
  key = k;
  return ok;
}

__global__ void extents_kernel(const int *__restrict__ coors, int n, int ndim, int *__restrict__ ext) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int v[4] = {0, 0, 0, 0};
  if (i < n) {
    bool ok = true;
    for (int j = 0; j < ndim; ++j) ok = ok && (coors[(size_t)i * ndim + j] >= 0);
    if (ok)
      for (int j = 0; j < ndim; ++j) v[j] = coors[(size_t)i * ndim + j] + 1;
  }
  for (int j = 0; j < ndim; ++j) {
    int m = v[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(ext + j, m);
  }
}

__global__ void mark_kernel(const int *__restrict__ coors, int n, Extents E, unsigned *__restrict__ bitmap) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long key;
  if (!row_key(coors + (size_t)i * E.ndim, E, key)) return;
  atomicOr(bitmap + (key >> 5), 1u << (unsigned)(key & 31));
}

struct EmitCoors {
  Extents E;
  int *out_coors;
  __device__ void operator()(int rank, unsigned long long key) const {
    int *o = out_coors + (size_t)rank * E.ndim;
    for (int j = E.ndim - 1; j >= 0; --j) {
      o[j] = (int)(key % (unsigned long long)E.e[j]);
      key /= (unsigned long long)E.e[j];
    }
  }
};

__device__ __forceinline__ void atomic_max_float(float *addr, float val) {
  if (val >= 0.f) atomicMax(reinterpret_cast<int *>(addr), __float_as_int(val));
  else atomicMin(reinterpret_cast<unsigned *>(addr), __float_as_uint(val));
}

__global__ void init_reduce_kernel(float *__restrict__ reduced, int *__restrict__ count, size_t rows, int c,
                                   float init) {
  size_t total = rows * c;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    reduced[i] = init;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (size_t)gridDim.x * blockDim.x)
    count[i] = 0;
}

// one thread per (point, feature)
__global__ void reduce_kernel(const float *__restrict__ feats, const int *__restrict__ coors, int n, int c,
                              Extents E, const unsigned *__restrict__ bitmap, const int *__restrict__ word_prefix,
                              int reduce_type, float *__restrict__ reduced, int *__restrict__ coors_map,
                              int *__restrict__ count) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * c) return;
  int i = (int)(t / c);
  int f = (int)(t - (long long)i * c);
  unsigned long long key;
  if (!row_key(coors + (size_t)i * E.ndim, E, key)) {
    if (f == 0) coors_map[i] = -1;
    return;
  }
  unsigned long long w = key >> 5;
  unsigned below = bitmap[w] & ((1u << (unsigned)(key & 31)) - 1u);
  int r = word_prefix[w] + __popc(below);
  if (f == 0) {
    coors_map[i] = r;
    atomicAdd(count + r, 1);
  }
  float v = feats[t];
  float *dst = reduced + (size_t)r * c + f;
  if (reduce_type == BEVF_REDUCE_MAX) atomic_max_float(dst, v);
  else atomicAdd(dst, v);
}

__global__ void mean_div_kernel(float *__restrict__ reduced, const int *__restrict__ count,
                                const int *__restrict__ m_dev, int c) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)(*m_dev) * c) return;
  reduced[t] = __fdiv_rn(reduced[t], (float)count[t / c]);
}

// ---- backward (scatter_points_cuda.cu:106-179) -------------------------------------------------------
__global__ void bwd_add_kernel(float *__restrict__ grad_feats, const float *__restrict__ grad_reduced,
                               const int *__restrict__ coors_map, const int *__restrict__ count, int n, int c,
                               int reduce_type) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * c) return;
  int i = (int)(t / c);
  int f = (int)(t - (long long)i * c);
  int r = coors_map[i];
  float g = 0.f;
  if (r >= 0) {
    g = grad_reduced[(size_t)r * c + f];
    if (reduce_type == BEVF_REDUCE_MEAN) g = __fdiv_rn(g, (float)count[r]);
  }
  grad_feats[t] = g;
}

__global__ void bwd_max_argmin_kernel(const float *__restrict__ feats, const float *__restrict__ reduced,
                                      const int *__restrict__ coors_map, int n, int c,
                                      int *__restrict__ reduce_from, float *__restrict__ grad_feats) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * c) return;
  grad_feats[t] = 0.f;
  int i = (int)(t / c);
  int f = (int)(t - (long long)i * c);
  int r = coors_map[i];
  if (r < 0) return;
  if (feats[t] == reduced[(size_t)r * c + f]) atomicMin(reduce_from + (size_t)r * c + f, i);
}

__global__ void bwd_max_scatter_kernel(float *__restrict__ grad_feats, const float *__restrict__ grad_reduced,
                                       const int *__restrict__ reduce_from, int m, int c, int n) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)m * c) return;
  int f = (int)(t % c);
  int i = reduce_from[t];
  if (i < n) grad_feats[(size_t)i * c + f] = grad_reduced[t];
}

__global__ void fill_int_kernel(int *__restrict__ p, size_t n, int v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] = v;
}

struct ScatterWs {
  unsigned *bitmap;
  int *word_prefix;
  int *block_counts;
  long long nwords;
  int nblk;
};

int key_space(const int *ext, int ndim, long long &nwords) {
  unsigned long long cells = 1;
  for (int j = 0; j < ndim; ++j) {
    if (ext[j] < 0) return BEVF_ERR_INVALID_ARGUMENT;
    cells *= (unsigned long long)(ext[j] > 0 ? ext[j] : 1);
    if (cells > (1ull << 34)) {
      bevf::set_error("dynamic_scatter: coordinate bounding box has more than 2^34 cells (unsupported)");
      return BEVF_ERR_UNSUPPORTED;
    }
  }
  nwords = (long long)((cells + 31) / 32);
  return BEVF_OK;
}

size_t carve_scatter(ScatterWs &w, void *ws, size_t bytes, long long nwords) {
  bevf::Workspace a(ws, bytes);
  w.nwords = nwords;
  w.nblk = bevf::rank_num_blocks(nwords);
  w.bitmap = a.take<unsigned>((size_t)nwords);
  w.word_prefix = a.take<int>((size_t)nwords);
  w.block_counts = a.take<int>((size_t)w.nblk);
  return a.off;
}

}  // namespace

BEVF_API int bevf_dynamic_scatter_extents(const int *coors, int n, int ndim, int *extents_dev, void *stream) {
  BEVF_CHECK_ARG(ndim >= 1 && ndim <= 4, "ndim must be in 1..4 (got %d)", ndim);
  cudaStream_t st = (cudaStream_t)stream;
  BEVF_CHECK_CUDA(cudaMemsetAsync(extents_dev, 0, 4 * sizeof(int), st));
  if (n == 0) return BEVF_OK;
  extents_kernel<<<bevf::ceil_div(n, 256), 256, 0, st>>>(coors, n, ndim, extents_dev);
  BEVF_CHECK_LAUNCH();
  return BEVF_OK;
}

BEVF_API size_t bevf_dynamic_scatter_workspace_bytes(int n, int ndim, const int *extents_host) {
  (void)n;
  long long nwords = 1;
  if (key_space(extents_host, ndim, nwords)) return 0;
  ScatterWs w;
  return carve_scatter(w, nullptr, 0, nwords) + 256;
}

BEVF_API int bevf_dynamic_scatter_forward(const float *feats, const int *coors, int n, int c, int ndim,
                                          const int *extents_host, int reduce_type, float *reduced,
                                          int *out_coors, int *coors_map, int *reduce_count, int *num_out_dev,
                                          void *workspace, size_t workspace_bytes, void *stream) {
  BEVF_CHECK_ARG(ndim >= 1 && ndim <= 4, "ndim must be in 1..4 (got %d)", ndim);
  BEVF_CHECK_ARG(reduce_type >= 0 && reduce_type <= 2, "unknown reduce type %d", reduce_type);
  BEVF_CHECK_ARG(n >= 0 && c > 0, "bad sizes n=%d c=%d", n, c);
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    BEVF_CHECK_CUDA(cudaMemsetAsync(num_out_dev, 0, sizeof(int), st));
    return BEVF_OK;
  }
  BEVF_CHECK_ARG(feats && coors && reduced && out_coors && coors_map && reduce_count && num_out_dev, "NULL tensor");
  long long nwords;
  int rc = key_space(extents_host, ndim, nwords);
  if (rc) return rc;
  ScatterWs w;
  size_t need = carve_scatter(w, workspace, workspace_bytes, nwords);
  if (!workspace || need > workspace_bytes) {
    bevf::set_error("dynamic_scatter workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
    return BEVF_ERR_WORKSPACE;
  }
  Extents E;
  E.ndim = ndim;
  for (int j = 0; j < 4; ++j) E.e[j] = (j < ndim && extents_host[j] > 0) ? extents_host[j] : 1;
  BEVF_CHECK_CUDA(cudaMemsetAsync(w.bitmap, 0, (size_t)nwords * sizeof(unsigned), st));
  mark_kernel<<<bevf::ceil_div(n, 256), 256, 0, st>>>(coors, n, E, w.bitmap);
  BEVF_CHECK_LAUNCH();
  rc = bevf::rank_build(w.bitmap, nwords, w.word_prefix, w.block_counts, num_out_dev, EmitCoors{E, out_coors}, st);
  if (rc) return rc;
  const float init = (reduce_type == BEVF_REDUCE_MAX) ? -__builtin_inff() : 0.f;
  init_reduce_kernel<<<bevf::kNumSMs * 4, 256, 0, st>>>(reduced, reduce_count, (size_t)n, c, init);
  BEVF_CHECK_LAUNCH();
  long long elems = (long long)n * c;
  reduce_kernel<<<bevf::ceil_div(elems, 256), 256, 0, st>>>(feats, coors, n, c, E, w.bitmap, w.word_prefix,
                                                            reduce_type, reduced, coors_map, reduce_count);
  BEVF_CHECK_LAUNCH();
  if (reduce_type == BEVF_REDUCE_MEAN) {
    mean_div_kernel<<<bevf::ceil_div(elems, 256), 256, 0, st>>>(reduced, reduce_count, num_out_dev, c);
    BEVF_CHECK_LAUNCH();
  }
  return BEVF_OK;
}

BEVF_API int bevf_dynamic_scatter_backward(float *grad_feats, const float *grad_reduced, const float *feats,
                                           const float *reduced, const int *coors_map, const int *reduce_count,
                                           int n, int m, int c, int reduce_type, void *workspace,
                                           size_t workspace_bytes, void *stream) {
  BEVF_CHECK_ARG(reduce_type >= 0 && reduce_type <= 2, "unknown reduce type %d", reduce_type);
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) return BEVF_OK;
  BEVF_CHECK_ARG(grad_feats != nullptr, "grad_feats is NULL");
  long long elems = (long long)n * c;
  if (m == 0) {
    BEVF_CHECK_CUDA(cudaMemsetAsync(grad_feats, 0, (size_t)elems * sizeof(float), st));
    return BEVF_OK;
  }
  if (reduce_type != BEVF_REDUCE_MAX) {
    bwd_add_kernel<<<bevf::ceil_div(elems, 256), 256, 0, st>>>(grad_feats, grad_reduced, coors_map, reduce_count, n,
                                                               c, reduce_type);
    BEVF_CHECK_LAUNCH();
    return BEVF_OK;
  }
  size_t need = (size_t)m * c * sizeof(int);
  if (!workspace || need > workspace_bytes) {
    bevf::set_error("dynamic_scatter backward workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
    return BEVF_ERR_WORKSPACE;
  }
  int *reduce_from = (int *)workspace;
  fill_int_kernel<<<bevf::kNumSMs * 2, 256, 0, st>>>(reduce_from, (size_t)m * c, n);
  BEVF_CHECK_LAUNCH();
  bwd_max_argmin_kernel<<<bevf::ceil_div(elems, 256), 256, 0, st>>>(feats, reduced, coors_map, n, c, reduce_from,
                                                                    grad_feats);
  BEVF_CHECK_LAUNCH();
  bwd_max_scatter_kernel<<<bevf::ceil_div((long long)m * c, 256), 256, 0, st>>>(grad_feats, grad_reduced,
                                                                               reduce_from, m, c, n);
  BEVF_CHECK_LAUNCH();
  return BEVF_OK;
}
