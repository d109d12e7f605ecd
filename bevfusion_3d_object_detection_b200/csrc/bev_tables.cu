// bev_tables.cu -- device-side construction of every index table the pooling kernels read, from the frustum geometry
// (SURVEY 8f-1).
//
// Reference: BaseViewTransform.bev_pool_aux, projects/BEVFusion/bevfusion/depth_lss.py:118-176 -- quantise ~2 M
// frustum points, append the batch id, four boolean-mask gathers, an int64 rank, an (unstable) argsort of ~1.8 M keys
// -- followed by the interval construction of bev_pool() (ops/bev_pool/bev_pool.py:158-166).  With training-time
// augmentation the geometry changes every sample, so this runs per frame in front of bev_pool.
//
// Here: one pass quantises and histograms, the order is a stable LSD radix sort keyed by the OUTPUT CELL (ties stay
// in frustum order, the deterministic choice among the orders the reference's unstable argsort may produce), and the
// interval / run / tile tables fall out of prefix sums over the cell histogram and over the ray-major run flags.
// Nothing synchronises with the host; the three table sizes land in counts[] for the caller to read when it wants.
// The sort and scans are cub's device-wide primitives (library code, like cuBLAS for a plain GEMM); the sort key is
// the cell number, 18 bits at 360 x 360.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace {

constexpr int kTileYTables = 32;  // == kTileY of bev_pool.cu (the tile the second phase of the forward works on)

struct Grid3 {
  float lo[3], dx[3];
  int nx[3];  // (x, y, z) cell counts
};

__global__ void __launch_bounds__(256)
    tables_quantize_kernel(const float *__restrict__ geom, int nprime, int per_batch, Grid3 g, int cells,
                           int *__restrict__ cell_of_point, int *__restrict__ key, int *__restrict__ iota,
                           int *__restrict__ cell_count) {
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < nprime; p += gridDim.x * blockDim.x) {
    int v[3];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      // depth_lss.py:129: ((geom - (bx - dx / 2)) / dx).long() -- truncation toward zero, so (-1, 0) lands in cell 0
      const float u = __fdiv_rn(__fsub_rn(geom[(size_t)p * 3 + j], g.lo[j]), g.dx[j]);
      ok = ok && (u > -1.0f) && (u < (float)g.nx[j]);   // also false for NaN
      v[j] = ok ? (int)u : 0;
    }
    const int b = p / per_batch;
    // (b, z, x, y) memory order of the pooled map: depth_lss.py:158-169 ranks by (x, y, z, b); the fused kernels walk
    // cells in output-memory order instead (identical when B == 1 and nz == 1)
    const int cell = ok ? ((b * g.nx[2] + v[2]) * g.nx[0] + v[0]) * g.nx[1] + v[1] : -1;
    cell_of_point[p] = cell;
    key[p] = ok ? cell : cells;
    iota[p] = p;
    if (ok) atomicAdd(cell_count + cell, 1);
  }
}

__global__ void __launch_bounds__(256) tables_flag_cells_kernel(const int *__restrict__ count, int cells,
                                                                int *__restrict__ flag) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c <= cells) flag[c] = (c < cells && count[c] > 0) ? 1 : 0;
}

// interval tables + tile_starts from the two cell-level prefix sums
__global__ void __launch_bounds__(256)
    tables_intervals_kernel(const int *__restrict__ cell_count, const int *__restrict__ cell_offset,
                            const int *__restrict__ cell_interval, int cells, int ny, int tiles_y, int n_tiles,
                            int *__restrict__ interval_cell, int *__restrict__ interval_starts,
                            int *__restrict__ tile_starts, int *__restrict__ counts) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < cells && cell_count[c] > 0) {
    const int i = cell_interval[c];
    interval_cell[i] = c;
    interval_starts[i] = cell_offset[c];
  }
  if (c == cells) {
    const int n_int = cell_interval[cells], nk = cell_offset[cells];
    interval_starts[n_int] = nk;
    counts[0] = nk;
    counts[1] = n_int;
  }
  if (c <= n_tiles) {
    const int cell0 = (c == n_tiles) ? cells : (c / tiles_y) * ny + (c % tiles_y) * kTileYTables;
    tile_starts[c] = cell_interval[min(cell0, cells)];
  }
}

// ray-major position q = ((bn * fW + w) * D + d) * fH + h  <->  frustum index p = ((bn * D + d) * fH + h) * fW + w
__device__ __forceinline__ int ray_to_frustum(int q, int D, int fH, int fW, int &h) {
  h = q % fH;
  const int d = (q / fH) % D;
  const int w = (q / (fH * D)) % fW;
  const int bn = q / (fH * D * fW);
  return ((bn * D + d) * fH + h) * fW + w;
}

// a run starts where a kept point is the first row of its (camera, column, depth bin) or its upper neighbour is in
// another cell (or dropped)
__global__ void __launch_bounds__(256)
    tables_run_flags_kernel(const int *__restrict__ cell_of_point, int nprime, int D, int fH, int fW,
                            int *__restrict__ flag) {
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q <= nprime; q += gridDim.x * blockDim.x) {
    int f = 0;
    if (q < nprime) {
      int h;
      const int p = ray_to_frustum(q, D, fH, fW, h);
      const int c = cell_of_point[p];
      f = (c >= 0 && (h == 0 || cell_of_point[p - fW] != c)) ? 1 : 0;
    }
    flag[q] = f;
  }
}

__global__ void __launch_bounds__(256)
    tables_run_emit_kernel(const int *__restrict__ cell_of_point, const int *__restrict__ flag,
                           const int *__restrict__ run_scan, int nprime, int D, int fH, int fW, int cells,
                           int *__restrict__ run_p0, int *__restrict__ run_len, int *__restrict__ run_key,
                           int *__restrict__ run_count, int *__restrict__ col_run_starts, int *__restrict__ counts) {
  const int col_len = D * fH;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q <= nprime; q += gridDim.x * blockDim.x) {
    const int id = run_scan[q];
    if (q % col_len == 0) col_run_starts[q / col_len] = id;  // q == nprime writes the closing entry (= n_runs)
    if (q == nprime) {
      counts[2] = id;
      continue;
    }
    if (!flag[q]) continue;
    int h;
    const int p = ray_to_frustum(q, D, fH, fW, h);
    const int c = cell_of_point[p];
    int len = 1;
    while (h + len < fH && cell_of_point[p + len * fW] == c) ++len;
    run_p0[id] = p;
    run_len[id] = len;
    run_key[id] = c;
    atomicAdd(run_count + c, 1);
  }
}

__global__ void __launch_bounds__(256) tables_fill_kernel(int *__restrict__ a, int n, int v) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) a[i] = v;
}

__global__ void __launch_bounds__(256)
    tables_cell_runs_kernel(const int *__restrict__ cell_count, const int *__restrict__ cell_interval,
                            const int *__restrict__ run_offset, int cells, int *__restrict__ cell_run_starts) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < cells && cell_count[c] > 0) cell_run_starts[cell_interval[c]] = run_offset[c];
  if (c == cells) cell_run_starts[cell_interval[cells]] = run_offset[cells];
}

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

struct TablesWs {
  size_t key, key_sorted, iota, flag, scan, cell_count, cell_offset, cell_flag, cell_interval, run_count, run_offset,
      cub, total, cub_bytes;
};

TablesWs plan_tables(long long nprime, long long cells) {
  TablesWs w{};
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += align256(bytes);
    return o;
  };
  const size_t np = (size_t)(nprime + 1) * sizeof(int), nc = (size_t)(cells + 1) * sizeof(int);
  w.key = take(np);
  w.key_sorted = take(np);
  w.iota = take(np);
  w.flag = take(np);
  w.scan = take(np);
  w.cell_count = take(nc);   // cell_count and run_count are adjacent: one memset clears both
  w.run_count = take(nc);
  w.cell_offset = take(nc);
  w.cell_flag = take(nc);
  w.cell_interval = take(nc);
  w.run_offset = take(nc);
  size_t s1 = 0, s2 = 0, s3 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, s1, (const int *)nullptr, (int *)nullptr, (const int *)nullptr,
                                  (int *)nullptr, (int)nprime, 0, 32);
  cub::DeviceScan::ExclusiveSum(nullptr, s2, (const int *)nullptr, (int *)nullptr, (int)(nprime + 1));
  cub::DeviceScan::ExclusiveSum(nullptr, s3, (const int *)nullptr, (int *)nullptr, (int)(cells + 1));
  w.cub_bytes = std::max(s1, std::max(s2, s3));
  w.cub = take(w.cub_bytes);
  w.total = off;
  return w;
}

}  // namespace

BEVF_API size_t bevf_bev_pool_tables_workspace_bytes(long long nprime, int b, int nz, int nx, int ny) {
  if (nprime <= 0 || b <= 0 || nz <= 0 || nx <= 0 || ny <= 0) return 0;
  return plan_tables(nprime, (long long)b * nz * nx * ny).total;
}

BEVF_API int bevf_bev_pool_build_tables(const float *geom, int bn, int d, int fh, int fw, int b, const float *bx,
                                        const float *dx, int nx, int ny, int nz, int *cell_of_point, int *src,
                                        int *interval_starts, int *interval_cell, int *tile_starts, int *run_p0,
                                        int *run_len, int *col_run_starts, int *cell_run_ids, int *cell_run_starts,
                                        int *counts, void *workspace, size_t workspace_bytes, void *stream) {
  BEVF_CHECK_ARG(bn > 0 && d > 0 && fh > 0 && fw > 0 && b > 0 && bn % b == 0, "bad frustum shape");
  BEVF_CHECK_ARG(nx > 0 && ny > 0 && nz > 0 && bx && dx, "bad grid");
  const long long nprime = (long long)bn * d * fh * fw, cells = (long long)b * nz * nx * ny;
  BEVF_CHECK_ARG(nprime < (1ll << 31) - 1 && cells < (1ll << 30), "frustum or grid too large for 32-bit tables");
  BEVF_CHECK_ARG(geom && cell_of_point && src && interval_starts && interval_cell && tile_starts && run_p0 &&
                     run_len && col_run_starts && cell_run_ids && cell_run_starts && counts && workspace,
                 "NULL tensor");
  const TablesWs w = plan_tables(nprime, cells);
  BEVF_CHECK_ARG(workspace_bytes >= w.total, "workspace too small: %zu < %zu", workspace_bytes, w.total);
  cudaStream_t st = (cudaStream_t)stream;
  char *ws = (char *)workspace;
  auto at = [&](size_t o) { return reinterpret_cast<int *>(ws + o); };
  int *key = at(w.key), *key_sorted = at(w.key_sorted), *iota = at(w.iota), *flag = at(w.flag), *scan = at(w.scan);
  int *cell_count = at(w.cell_count), *run_count = at(w.run_count), *cell_offset = at(w.cell_offset);
  int *cell_flag = at(w.cell_flag), *cell_interval = at(w.cell_interval), *run_offset = at(w.run_offset);
  void *cub_ws = ws + w.cub;
  size_t cub_bytes = w.cub_bytes;

  Grid3 g;
  for (int j = 0; j < 3; ++j) {
    volatile float half = dx[j] / 2.0f;   // (bx - dx / 2.0) in fp32, as the tensor expression rounds it
    volatile float lo = bx[j] - half;
    g.lo[j] = lo;
    g.dx[j] = dx[j];
  }
  g.nx[0] = nx, g.nx[1] = ny, g.nx[2] = nz;
  const int np = (int)nprime, nc = (int)cells;
  const int grid_p = (int)std::min<long long>(bevf::ceil_div(nprime + 1, 256), (long long)bevf::kNumSMs * 16);
  const int grid_c = bevf::ceil_div(nc + 1, 256);
  int end_bit = 1;
  while ((1ll << end_bit) <= cells) ++end_bit;  // keys are 0..cells inclusive

  BEVF_CHECK_CUDA(cudaMemsetAsync(cell_count, 0, w.cell_offset - w.cell_count, st));  // cell_count + run_count
  tables_quantize_kernel<<<grid_p, 256, 0, st>>>(geom, np, np / b, g, nc, cell_of_point, key, iota, cell_count);
  BEVF_CHECK_LAUNCH();
  tables_flag_cells_kernel<<<grid_c, 256, 0, st>>>(cell_count, nc, cell_flag);
  BEVF_CHECK_LAUNCH();
  BEVF_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(cub_ws, cub_bytes, cell_count, cell_offset, nc + 1, st));
  BEVF_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(cub_ws, cub_bytes, cell_flag, cell_interval, nc + 1, st));
  const int n_tiles = bevf_bev_pool_num_tiles(b, nz, nx, ny);
  const int tiles_y = bevf::ceil_div(ny, kTileYTables);
  tables_intervals_kernel<<<bevf::ceil_div(std::max(nc, n_tiles) + 1, 256), 256, 0, st>>>(
      cell_count, cell_offset, cell_interval, nc, ny, tiles_y, n_tiles, interval_cell, interval_starts, tile_starts,
      counts);
  BEVF_CHECK_LAUNCH();
  // rank order: stable sort of the frustum indices by cell; the dropped points (key == cells) end up behind nk
  BEVF_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(cub_ws, cub_bytes, key, key_sorted, iota, src, np, 0, end_bit, st));

  // ray-major run tables
  tables_run_flags_kernel<<<grid_p, 256, 0, st>>>(cell_of_point, np, d, fh, fw, flag);
  BEVF_CHECK_LAUNCH();
  BEVF_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(cub_ws, cub_bytes, flag, scan, np + 1, st));
  tables_fill_kernel<<<grid_p, 256, 0, st>>>(key, np, nc);  // run keys past n_runs sort to the end
  BEVF_CHECK_LAUNCH();
  tables_run_emit_kernel<<<grid_p, 256, 0, st>>>(cell_of_point, flag, scan, np, d, fh, fw, nc, run_p0, run_len, key,
                                                 run_count, col_run_starts, counts);
  BEVF_CHECK_LAUNCH();
  BEVF_CHECK_CUDA(
      cub::DeviceRadixSort::SortPairs(cub_ws, cub_bytes, key, key_sorted, iota, cell_run_ids, np, 0, end_bit, st));
  BEVF_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(cub_ws, cub_bytes, run_count, run_offset, nc + 1, st));
  tables_cell_runs_kernel<<<grid_c, 256, 0, st>>>(cell_count, cell_interval, run_offset, nc, cell_run_starts);
  BEVF_CHECK_LAUNCH();
  return BEVF_OK;
}
