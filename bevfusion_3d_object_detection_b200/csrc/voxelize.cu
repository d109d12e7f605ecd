// voxelize.cu -- hard / dynamic voxelization for sm_100a.
//
// Replaces projects/BEVFusion/bevfusion/ops/voxel/src/voxelization_cuda.cu (reference): the deterministic
// path there is an O(N^2) predecessor scan (:106-147) plus a single-thread walk (:150-180) and four device
// synchronisations.  Here the same first-appearance semantics are produced in O(N) with:
//   K1  vox_insert      point -> cell key (fp32 sub, IEEE div, floor), 64-bit open-addressing hash slot
//                       {key:32 | first point index:32} claimed with atomicCAS, first index kept with a
//                       64-bit atomicMin (same key => same high word, so min acts on the index)
//   K2a vox_head_count  head flag (point is the first of its voxel) + per-block head counts
//   K2b vox_scan_assign block offset = sum of earlier block counts, warp-shuffle scan -> voxel id in
//                       first-appearance order; clamps at max_voxels; writes coors and voxel_num
//   K3  vox_rank        per voxel, the max_points smallest point indices are kept in an ascending list by a
//                       carry chain of atomicMin (a concurrent insertion sort whose result is race-free)
//   K4  vox_gather      coalesced write of voxels[v, k, :] = points[list[v][k], :] (+ zero padding) and
//                       num_points_per_voxel
// No host synchronisation; voxel_num stays on the device unless the *_sync wrapper is used.
#include <limits.h>

#include "common.cuh"

namespace {

constexpr unsigned long long kEmptySlot = 0xFFFFFFFFFFFFFFFFull;
constexpr int kSentinel = 0x7F7F7F7F;  // cudaMemset(0x7F): larger than any point index
constexpr int kInsertThreads = 256;
constexpr int kScanThreads = 256;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;

struct VoxParams {
  float vx, vy, vz;
  float x0, y0, z0;
  int gx, gy, gz;
};

// Reference arithmetic: floor((p - min) / voxel) in fp32 with a true IEEE divide
// (voxelization_cuda.cu:37-50; the reference build has no --use_fast_math).
__device__ __forceinline__ int cell_coord(float p, float lo, float vs) {
  return (int)floorf(__fdiv_rn(__fsub_rn(p, lo), vs));
}

__device__ __forceinline__ uint32_t hash_u32(uint32_t k) {
  k *= 0x9E3779B1u;
  k ^= k >> 15;
  k *= 0x85EBCA77u;
  k ^= k >> 13;
  return k;
}

// ---- dynamic voxelize ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dynamic_voxelize_kernel(const float *__restrict__ points, int n, int c,
                                                                 VoxParams P, int *__restrict__ coors) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float *p = points + (size_t)i * c;
  int *o = coors + (size_t)i * 3;
  int cx = cell_coord(p[0], P.x0, P.vx);
  if (cx < 0 || cx >= P.gx) {
    o[0] = -1;  // reference returns here: o[1], o[2] keep the caller's values
    return;
  }
  int cy = cell_coord(p[1], P.y0, P.vy);
  if (cy < 0 || cy >= P.gy) {
    o[0] = -1;
    o[1] = -1;
    return;
  }
  int cz = cell_coord(p[2], P.z0, P.vz);
  if (cz < 0 || cz >= P.gz) {
    o[0] = -1;
    o[1] = -1;
    o[2] = -1;
  } else {
    o[0] = cx;
    o[1] = cy;
    o[2] = cz;
  }
}

// ---- K1: hash insert -------------------------------------------------------------------------------
// Points of a block are staged through shared memory with 16-byte loads when the block's first float is
// 16-byte aligned (C=5: every 256-point block starts on a 5120-byte boundary).
__global__ void __launch_bounds__(kInsertThreads)
    vox_insert_kernel(const float *__restrict__ points, int n, int c, VoxParams P,
                      unsigned long long *__restrict__ table, uint32_t mask, int *__restrict__ slot_of_point) {
  extern __shared__ float4 smem4[];
  float *sm = reinterpret_cast<float *>(smem4);
  const int b0 = blockIdx.x * kInsertThreads;
  const int cnt = min(kInsertThreads, n - b0);
  const size_t f0 = (size_t)b0 * c;
  const int nf = cnt * c;
  const float *src = points + f0;
  if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
    const float4 *src4 = reinterpret_cast<const float4 *>(src);
    for (int t = threadIdx.x; t < nf / 4; t += kInsertThreads) smem4[t] = __ldg(src4 + t);
    for (int t = (nf / 4) * 4 + threadIdx.x; t < nf; t += kInsertThreads) sm[t] = __ldg(src + t);
  } else {
    for (int t = threadIdx.x; t < nf; t += kInsertThreads) sm[t] = __ldg(src + t);
  }
  __syncthreads();
  if ((int)threadIdx.x >= cnt) return;
  const int i = b0 + threadIdx.x;
  const float *p = sm + threadIdx.x * c;
  int cx = cell_coord(p[0], P.x0, P.vx);
  int cy = cell_coord(p[1], P.y0, P.vy);
  int cz = cell_coord(p[2], P.z0, P.vz);
  if (cx < 0 || cx >= P.gx || cy < 0 || cy >= P.gy || cz < 0 || cz >= P.gz) {
    slot_of_point[i] = -1;
    return;
  }
  const uint32_t key = ((uint32_t)cx * (uint32_t)P.gy + (uint32_t)cy) * (uint32_t)P.gz + (uint32_t)cz;
  const unsigned long long packed = ((unsigned long long)key << 32) | (unsigned long long)(uint32_t)i;
  uint32_t s = hash_u32(key) & mask;
  while (true) {
    unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(table + s);
    if (cur == kEmptySlot) {
      cur = atomicCAS(table + s, kEmptySlot, packed);
      if (cur == kEmptySlot) break;  // claimed
    }
    if ((uint32_t)(cur >> 32) == key) {
      if ((uint32_t)cur > (uint32_t)i) atomicMin(table + s, packed);
      break;
    }
    s = (s + 1) & mask;
  }
  slot_of_point[i] = (int)s;
}

// ---- K2a: head flags + per-block counts ------------------------------------------------------------
__global__ void __launch_bounds__(kScanThreads)
    vox_head_count_kernel(const unsigned long long *__restrict__ table, const int *__restrict__ slot_of_point,
                          int n, unsigned char *__restrict__ head, int *__restrict__ block_counts) {
  __shared__ int warp_sums[kScanThreads / 32];
  const int base = blockIdx.x * kScanTile;
  int cnt = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int i = base + k * kScanThreads + threadIdx.x;
    if (i < n) {
      int s = slot_of_point[i];
      unsigned char h = 0;
      if (s >= 0) h = ((uint32_t)table[s] == (uint32_t)i) ? 1 : 0;
      head[i] = h;
      cnt += h;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) t += warp_sums[w];
    block_counts[blockIdx.x] = t;
  }
}

// ---- K2b: scan + voxel id assignment ---------------------------------------------------------------
// Items are laid out blocked per thread (thread t owns points base + t*kScanItems .. +kScanItems-1) so the
// exclusive scan follows point order.
__global__ void __launch_bounds__(kScanThreads)
    vox_scan_assign_kernel(const unsigned long long *__restrict__ table, const int *__restrict__ slot_of_point,
                           const unsigned char *__restrict__ head, const int *__restrict__ block_counts, int n,
                           int max_voxels, VoxParams P, int *__restrict__ vid_of_slot, int *__restrict__ coors,
                           int coors_stride, int coors_off, int batch_idx, const int *__restrict__ row_offset,
                           int *__restrict__ voxel_num) {
  __shared__ int red[kScanThreads / 32];
  __shared__ int warp_excl[kScanThreads / 32];
  __shared__ int block_offset;
  // offset of this block = sum of the counts of all earlier blocks
  int part = 0;
  for (int b = threadIdx.x; b < (int)blockIdx.x; b += kScanThreads) part += block_counts[b];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) t += red[w];
    block_offset = t;
  }
  const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  unsigned char h[kScanItems];
  int local = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int i = base + k;
    h[k] = (i < n) ? head[i] : 0;
    local += h[k];
  }
  // warp inclusive scan of per-thread totals
  int incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, o);
    if ((int)(threadIdx.x & 31) >= o) incl += v;
  }
  __syncthreads();  // red[] reuse below is safe; block_offset visible
  if ((threadIdx.x & 31) == 31) red[threadIdx.x >> 5] = incl;
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = (threadIdx.x < kScanThreads / 32) ? red[threadIdx.x] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int u = __shfl_up_sync(0xffffffffu, inc, o);
      if ((int)threadIdx.x >= o) inc += u;
    }
    if (threadIdx.x < kScanThreads / 32) warp_excl[threadIdx.x] = inc - v;
  }
  __syncthreads();
  int run = block_offset + warp_excl[threadIdx.x >> 5] + (incl - local);
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (h[k]) {
      int i = base + k;
      int s = slot_of_point[i];
      int vid = run;
      run += 1;
      if (vid < max_voxels) {
        vid_of_slot[s] = vid;
        uint32_t key = (uint32_t)(table[s] >> 32);
        uint32_t cz = key % (uint32_t)P.gz;
        uint32_t r = key / (uint32_t)P.gz;
        uint32_t cy = r % (uint32_t)P.gy;
        uint32_t cx = r / (uint32_t)P.gy;
        int *o = coors + ((size_t)vid + (row_offset ? *row_offset : 0)) * coors_stride;
        if (coors_off) o[0] = batch_idx;
        o[coors_off + 0] = (int)cx;
        o[coors_off + 1] = (int)cy;
        o[coors_off + 2] = (int)cz;
      } else {
        vid_of_slot[s] = -1;
      }
    }
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == kScanThreads - 1) {
    // `run` of the last thread of the last block = total number of distinct voxels
    *voxel_num = min(run, max_voxels);
  }
}

// ---- K3: ordered slot assignment -------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    vox_rank_kernel(const int *__restrict__ slot_of_point, const int *__restrict__ vid_of_slot, int n,
                    int max_points, int *__restrict__ lists) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int s = slot_of_point[i];
  if (s < 0) return;
  int v = vid_of_slot[s];
  if (v < 0) return;
  int *list = lists + (size_t)v * max_points;
  // entries only ever decrease: if the last one is already below i, i cannot be among the smallest
  if (*reinterpret_cast<volatile int *>(list + max_points - 1) < i) return;
  int cur = i;
  for (int k = 0; k < max_points; ++k) {
    int old = atomicMin(list + k, cur);
    if (old == kSentinel) break;  // filled an empty entry, nothing displaced
    if (old > cur) cur = old;     // displaced a larger index: carry it to the next entry
  }
}

// ---- K4: gather ------------------------------------------------------------------------------------
// Row mode (caller pre-zeroed the outputs, reference contract): one thread per (voxel, slot) row, only
// occupied rows are written.
__global__ void __launch_bounds__(256)
    vox_gather_rows_kernel(const float *__restrict__ points, int c, const int *__restrict__ lists,
                           int max_points, const int *__restrict__ voxel_num, float *__restrict__ voxels,
                           int *__restrict__ npv) {
  const int m = *voxel_num;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)m * max_points) return;
  int k = (int)(t % max_points);
  int idx = lists[t];
  if (idx == kSentinel) return;
  const float *src = points + (size_t)idx * c;
  float *dst = voxels + (size_t)t * c;
  for (int f = 0; f < c; ++f) dst[f] = __ldg(src + f);
  if (k == max_points - 1 || lists[t + 1] == kSentinel) npv[t / max_points] = k + 1;
}

// Fill mode: every float of the first voxel_num voxels is written (occupied rows from the points, the rest
// zero) with 16-byte stores; the buffers may be uninitialised.
__global__ void __launch_bounds__(256)
    vox_gather_fill_kernel(const float *__restrict__ points, int c, const int *__restrict__ lists,
                           int max_points, const int *__restrict__ voxel_num, float *__restrict__ voxels,
                           int *__restrict__ npv) {
  const int m = *voxel_num;
  const long long total = (long long)m * max_points * c;
  long long e0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (e0 >= total) return;
  float v[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    long long e = e0 + q;
    v[q] = 0.f;
    if (e < total) {
      long long row = e / c;
      int f = (int)(e - row * c);
      int idx = lists[row];
      if (idx != kSentinel) {
        v[q] = __ldg(points + (size_t)idx * c + f);
        int k = (int)(row % max_points);
        if (f == 0 && (k == max_points - 1 || lists[row + 1] == kSentinel)) npv[row / max_points] = k + 1;
      }
    }
  }
  if (e0 + 3 < total && (reinterpret_cast<uintptr_t>(voxels) & 15u) == 0) {
    *reinterpret_cast<float4 *>(voxels + e0) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (e0 + q < total) voxels[e0 + q] = v[q];
  }
}

// Mean mode (bevfusion.py:251-253): feats[v,:] = sum_k points[list[v][k],:] / count, sizes[v] = count.
// One thread per (voxel, feature); the sum runs in slot order like feats.sum(dim=1) over the padded tensor.
__global__ void __launch_bounds__(256)
    vox_gather_mean_kernel(const float *__restrict__ points, int c, const int *__restrict__ lists,
                           int max_points, const int *__restrict__ voxel_num, float *__restrict__ feats,
                           int *__restrict__ sizes, const int *__restrict__ row_offset) {
  const int m = *voxel_num;
  const size_t row0 = row_offset ? (size_t)*row_offset : 0;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)m * c) return;
  int v = (int)(t / c);
  int f = (int)(t - (long long)v * c);
  const int *list = lists + (size_t)v * max_points;
  float acc = 0.f;
  int cnt = 0;
  for (int k = 0; k < max_points; ++k) {
    int idx = list[k];
    if (idx == kSentinel) break;
    acc += __ldg(points + (size_t)idx * c + f);
    cnt += 1;
  }
  feats[row0 * c + t] = __fdiv_rn(acc, (float)cnt);
  if (f == 0) sizes[row0 + v] = cnt;
}

__global__ void add_offset_kernel(const int *__restrict__ voxel_num, int *__restrict__ row_offset) {
  *row_offset += *voxel_num;
}

int make_params(const float *vs, const float *rng, VoxParams &P) {
  int grid[3];
  bevf_voxel_grid_size(vs, rng, grid);
  P.vx = vs[0]; P.vy = vs[1]; P.vz = vs[2];
  P.x0 = rng[0]; P.y0 = rng[1]; P.z0 = rng[2];
  P.gx = grid[0]; P.gy = grid[1]; P.gz = grid[2];
  if (P.gx <= 0 || P.gy <= 0 || P.gz <= 0) {
    bevf::set_error("voxel grid is empty: %d x %d x %d", P.gx, P.gy, P.gz);
    return BEVF_ERR_INVALID_ARGUMENT;
  }
  if ((long long)P.gx * P.gy * P.gz >= 0xFFFFFFFFll) {
    bevf::set_error("voxel grid %d x %d x %d has >= 2^32-1 cells (unsupported)", P.gx, P.gy, P.gz);
    return BEVF_ERR_UNSUPPORTED;
  }
  return BEVF_OK;
}

uint32_t table_capacity(int n) {
  uint32_t cap = 1024;
  while (cap < 2u * (uint32_t)n) cap <<= 1;
  return cap;
}

struct VoxWorkspace {
  unsigned long long *table;
  int *vid_of_slot;
  int *slot_of_point;
  int *lists;
  int *block_counts;
  unsigned char *head;
  uint32_t cap;
  int list_rows;
  int nblk;
};

size_t carve(VoxWorkspace &w, void *ws, size_t ws_bytes, int n, int max_points, int max_voxels) {
  bevf::Workspace a(ws, ws_bytes);
  w.cap = table_capacity(n);
  w.list_rows = max(1, min(max_voxels, n));
  w.nblk = max(1, bevf::ceil_div(n, kScanTile));
  w.table = a.take<unsigned long long>(w.cap);
  w.lists = a.take<int>((size_t)w.list_rows * max_points);
  w.vid_of_slot = a.take<int>(w.cap);
  w.slot_of_point = a.take<int>(n);
  w.block_counts = a.take<int>(w.nblk);
  w.head = a.take<unsigned char>(n);
  return a.off;
}

// Common front half (K1..K3).  On return lists / voxel_num_dev / coors are final.
int voxelize_core(const float *points, int n, int c, const VoxParams &P, int max_points, int max_voxels,
                  const VoxWorkspace &w, int *coors, int coors_stride, int coors_off, int batch_idx,
                  const int *row_offset_dev, int *voxel_num_dev, cudaStream_t st) {
  BEVF_CHECK_CUDA(cudaMemsetAsync(w.table, 0xFF, (size_t)w.cap * sizeof(unsigned long long), st));
  BEVF_CHECK_CUDA(cudaMemsetAsync(w.lists, 0x7F, (size_t)w.list_rows * max_points * sizeof(int), st));
  const int nb = bevf::ceil_div(n, kInsertThreads);
  vox_insert_kernel<<<nb, kInsertThreads, (size_t)kInsertThreads * c * sizeof(float), st>>>(
      points, n, c, P, w.table, w.cap - 1, w.slot_of_point);
  BEVF_CHECK_LAUNCH();
  vox_head_count_kernel<<<w.nblk, kScanThreads, 0, st>>>(w.table, w.slot_of_point, n, w.head, w.block_counts);
  BEVF_CHECK_LAUNCH();
  vox_scan_assign_kernel<<<w.nblk, kScanThreads, 0, st>>>(w.table, w.slot_of_point, w.head, w.block_counts, n,
                                                         max_voxels, P, w.vid_of_slot, coors, coors_stride,
                                                         coors_off, batch_idx, row_offset_dev, voxel_num_dev);
  BEVF_CHECK_LAUNCH();
  vox_rank_kernel<<<bevf::ceil_div(n, 256), 256, 0, st>>>(w.slot_of_point, w.vid_of_slot, n, max_points, w.lists);
  BEVF_CHECK_LAUNCH();
  return BEVF_OK;
}

int check_common(const float *points, int n, int c, int max_points, int max_voxels, int ndim) {
  BEVF_CHECK_ARG(ndim == 3, "NDim must be 3 (got %d)", ndim);
  BEVF_CHECK_ARG(n >= 0 && c >= 3, "points must be [N>=0, C>=3] (got N=%d C=%d)", n, c);
  BEVF_CHECK_ARG(c <= 64, "num_features %d > 64 is not supported", c);
  BEVF_CHECK_ARG(n == 0 || points != nullptr, "points is NULL");
  BEVF_CHECK_ARG(max_points > 0 && max_voxels > 0, "max_points / max_voxels must be positive (got %d, %d)",
                 max_points, max_voxels);
  return BEVF_OK;
}

}  // namespace

BEVF_API int bevf_voxel_grid_size(const float *vs, const float *rng, int *grid) {
  for (int j = 0; j < 3; ++j) grid[j] = (int)roundf((rng[3 + j] - rng[j]) / vs[j]);
  return BEVF_OK;
}

BEVF_API int bevf_dynamic_voxelize(const float *points, int n, int c, int *coors, const float *vs,
                                   const float *rng, int ndim, void *stream) {
  BEVF_CHECK_ARG(ndim == 3, "NDim must be 3 (got %d)", ndim);
  BEVF_CHECK_ARG(n >= 0 && c >= 3, "points must be [N>=0, C>=3] (got N=%d C=%d)", n, c);
  if (n == 0) return BEVF_OK;
  BEVF_CHECK_ARG(points && coors, "NULL tensor");
  VoxParams P;
  int rc = make_params(vs, rng, P);
  if (rc) return rc;
  dynamic_voxelize_kernel<<<bevf::ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(points, n, c, P, coors);
  BEVF_CHECK_LAUNCH();
  return BEVF_OK;
}

BEVF_API size_t bevf_hard_voxelize_workspace_bytes(int n, int max_points, int max_voxels) {
  VoxWorkspace w;
  if (n < 1) n = 1;
  if (max_points < 1) max_points = 1;
  if (max_voxels < 1) max_voxels = 1;
  return carve(w, nullptr, 0, n, max_points, max_voxels) + 256;
}

BEVF_API int bevf_hard_voxelize(const float *points, int n, int c, float *voxels, int *coors, int *npv,
                                const float *vs, const float *rng, int max_points, int max_voxels, int ndim,
                                int deterministic, int zero_fill, void *workspace, size_t workspace_bytes,
                                int *voxel_num_dev, void *stream) {
  (void)deterministic;  // both modes return the deterministic result (see header)
  int rc = check_common(points, n, c, max_points, max_voxels, ndim);
  if (rc) return rc;
  BEVF_CHECK_ARG(voxels && coors && npv && voxel_num_dev, "NULL output tensor");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    BEVF_CHECK_CUDA(cudaMemsetAsync(voxel_num_dev, 0, sizeof(int), st));
    return BEVF_OK;
  }
  VoxParams P;
  rc = make_params(vs, rng, P);
  if (rc) return rc;
  VoxWorkspace w;
  size_t need = carve(w, workspace, workspace_bytes, n, max_points, max_voxels);
  if (!workspace || need > workspace_bytes) {
    bevf::set_error("hard_voxelize workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
    return BEVF_ERR_WORKSPACE;
  }
  rc = voxelize_core(points, n, c, P, max_points, max_voxels, w, coors, 3, 0, 0, nullptr, voxel_num_dev, st);
  if (rc) return rc;
  if (zero_fill) {
    long long quads = ((long long)w.list_rows * max_points * c + 3) / 4;
    vox_gather_fill_kernel<<<bevf::ceil_div(quads, 256), 256, 0, st>>>(points, c, w.lists, max_points,
                                                                       voxel_num_dev, voxels, npv);
  } else {
    long long rows = (long long)w.list_rows * max_points;
    vox_gather_rows_kernel<<<bevf::ceil_div(rows, 256), 256, 0, st>>>(points, c, w.lists, max_points,
                                                                      voxel_num_dev, voxels, npv);
  }
  BEVF_CHECK_LAUNCH();
  return BEVF_OK;
}

BEVF_API int bevf_hard_voxelize_sync(const float *points, int n, int c, float *voxels, int *coors, int *npv,
                                     const float *vs, const float *rng, int max_points, int max_voxels, int ndim,
                                     int deterministic, int zero_fill, void *workspace, size_t workspace_bytes,
                                     int *voxel_num_dev, void *stream, int *voxel_num_host) {
  int rc = bevf_hard_voxelize(points, n, c, voxels, coors, npv, vs, rng, max_points, max_voxels, ndim,
                              deterministic, zero_fill, workspace, workspace_bytes, voxel_num_dev, stream);
  if (rc) return rc;
  BEVF_CHECK_CUDA(cudaMemcpyAsync(voxel_num_host, voxel_num_dev, sizeof(int), cudaMemcpyDeviceToHost,
                                  (cudaStream_t)stream));
  BEVF_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return BEVF_OK;
}

BEVF_API int bevf_voxelize_mean(const float *points, int n, int c, float *feats, int *coords4, int *sizes,
                                const float *vs, const float *rng, int max_points, int max_voxels, int batch_idx,
                                void *workspace, size_t workspace_bytes, int *voxel_num_dev, int *row_offset_dev,
                                void *stream) {
  int rc = check_common(points, n, c, max_points, max_voxels, 3);
  if (rc) return rc;
  BEVF_CHECK_ARG(feats && coords4 && sizes && voxel_num_dev, "NULL output tensor");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    BEVF_CHECK_CUDA(cudaMemsetAsync(voxel_num_dev, 0, sizeof(int), st));
    return BEVF_OK;
  }
  VoxParams P;
  rc = make_params(vs, rng, P);
  if (rc) return rc;
  VoxWorkspace w;
  size_t need = carve(w, workspace, workspace_bytes, n, max_points, max_voxels);
  if (!workspace || need > workspace_bytes) {
    bevf::set_error("voxelize_mean workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
    return BEVF_ERR_WORKSPACE;
  }
  rc = voxelize_core(points, n, c, P, max_points, max_voxels, w, coords4, 4, 1, batch_idx, row_offset_dev, voxel_num_dev,
                     st);
  if (rc) return rc;
  long long elems = (long long)w.list_rows * c;
  vox_gather_mean_kernel<<<bevf::ceil_div(elems, 256), 256, 0, st>>>(points, c, w.lists, max_points,
                                                                     voxel_num_dev, feats, sizes, row_offset_dev);
  BEVF_CHECK_LAUNCH();
  if (row_offset_dev) {
    add_offset_kernel<<<1, 1, 0, st>>>(voxel_num_dev, row_offset_dev);
    BEVF_CHECK_LAUNCH();
  }
  return BEVF_OK;
}
