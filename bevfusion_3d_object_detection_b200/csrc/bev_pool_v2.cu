// bev_pool_v2.cu -- fused camera-to-BEV pooling, second generation: two launches, no layout pre-pass.
//
//   out[b, z*C + c, x, y] = sum over the frustum points (cam, d, h, w) that fall into cell (b, z, x, y) of
//                           depth[cam, d, h, w] * ctx[cam, c, h, w]
//
// Reference data path: depth_lss.py:699-725 (outer product), :179-204 (reshape / kept / sort / bev_pool / collapse Z),
// bev_pool_cuda.cu:20-42 (K1).  The frustum tensor (638 MB at config A) never exists; order of summation is fixed
// (deterministic), no atomics.
//
// A *run* is a maximal set of frustum points of one (camera, depth bin, column w) with consecutive rows h and the same
// BEV cell (per-calibration tables, bev_tables.cu); a level camera gives ~one run per (camera, d, w): 62 k runs for 1.8 M
// points.  partial[r, :] = sum_h depth * ctx of run r is a tiny dense contraction per pixel column
// ([D x fH] x [fH x C]); a cell then sums its 1-3 runs.
//
// phase 1 (ray-major).  A CTA owns FOUR adjacent pixel columns of one camera and a range of depth bins.  It reads depth
//   [cam, d, h, w0..w0+3] and context [cam, c, h, w0..w0+3] as 16-byte vectors STRAIGHT from the NCHW tensors the
//   depthnet emits (half of every 32-byte sector is used; the former NCHW->NHWC pre-pass and its 6.8 MB round trip are
//   gone) and transposes them into shared memory.  The contraction runs from shared memory with register tiles: a group
//   of C/8 lanes owns 4 runs x 8 channels (32 accumulators per lane; 30 of 32 lanes busy at C = 80), one LDS.128 pair of
//   context feeds 4 runs with packed fp32 FMAs (fma.rn.f32x2), depth weights are broadcast loads, masked only on the rows
//   some run of the tile does not cover.  Rows of `partial` are written at the run's position in CELL order (run_pos), so
//   that
// phase 2 (cell-major) finds every cell's 1-3 rows adjacent (no run-id indirection; a tile's rows are one contiguous,
//   L1-resident range): a lane owns a cell, a warp takes 4 channels per step, rows are added in ascending order and leave
//   as 128-byte channel rows of the [B, nz*C, nx, ny] map; tiles without points write zeros.
// Measured (config A, CUDA graph, inputs > L2): 65 us at batch 1 (first generation: 74 us in three launches), 50 us per
// frame at batch 4 (67 us); phase 1 ~31 us, phase 2 ~32 us at batch 1.
//
// Algorithmic bytes (config A): depth 13.4 MB + context 6.8 MB read once, 41.5 MB written, tables 0.9 MB; `partial`
// (20 MB) lives in L2 between the launches.
#include <stdlib.h>

#include "common.cuh"

namespace {

// experiments: BEVFRONT_POOL_SKIP=1 / 2 leaves phase 1 / phase 2 out (per-phase timing inside a graph)
const int g_dsplit = [] { const char *e = getenv("BEVFRONT_POOL_DSPLIT"); return e ? atoi(e) : 0; }();
const int g_skip = [] { const char *e = getenv("BEVFRONT_POOL_SKIP"); return e ? atoi(e) : 0; }();


constexpr int kWG = 4;          // pixel columns per CTA (one 16-byte vector of depth / context)
constexpr int kRB = 4;          // runs per lane group (register tile height)
constexpr int kP1Threads = 256;
constexpr int kTileY2 = 32;     // output cells per phase-2 tile
constexpr int kP2Threads = 128;

// packed fp32 FMA (sm_100: fma.rn.f32x2, two IEEE fp32 FMAs per instruction -- same rounding as two fmaf)
__device__ __forceinline__ void ffma2(float2 &d, const float2 a, const float2 b) {
  unsigned long long dd, aa, bb;
  asm("mov.b64 %0, {%1, %2};" : "=l"(dd) : "f"(d.x), "f"(d.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(bb) : "f"(b.x), "f"(b.y));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(dd));
}
struct Acc8 {   // 8 channels of one run
  float2 v[4];
};
__device__ __forceinline__ void fma8(Acc8 &a, float w, const float4 &v0, const float4 &v1) {
  const float2 ww = make_float2(w, w);
  ffma2(a.v[0], ww, make_float2(v0.x, v0.y));
  ffma2(a.v[1], ww, make_float2(v0.z, v0.w));
  ffma2(a.v[2], ww, make_float2(v1.x, v1.y));
  ffma2(a.v[3], ww, make_float2(v1.z, v1.w));
}

// first run in [lo, hi) whose depth bin is >= d_target (runs of a column are sorted by depth bin)
__device__ int lower_bound_run(const int *__restrict__ run_p0, int lo, int hi, int plane, int D, int d_target) {
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const int d = (__ldg(run_p0 + mid) / plane) % D;
    if (d < d_target) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(kP1Threads)
    bev_pool_p1_kernel(const float *__restrict__ depth, const float *__restrict__ ctx, const int *__restrict__ run_p0,
                       const int *__restrict__ run_len, const int *__restrict__ run_pos,
                       const int *__restrict__ col_run_starts, int D, int fH, int fW, int C, int dc, int csp, int dsp,
                       float *__restrict__ partial) {
  extern __shared__ __align__(16) float sm1[];
  __shared__ int s_a[kWG], s_b[kWG], s_toff[kWG + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int groups_w = fW / kWG;
  const int bn = blockIdx.x / groups_w, w0 = (blockIdx.x % groups_w) * kWG;
  const int d_lo = blockIdx.y * dc, d_hi = min(D, d_lo + dc);
  const int plane = fH * fW;
  float *ctx_s = sm1;                          // [kWG][fH][csp]   (csp = C + 4: 16-byte aligned rows, skewed banks)
  float *dm = sm1 + kWG * fH * csp;            // [kWG][dc][dsp]   (dsp = fH + 1)
  // ---- run ranges of the four columns inside this CTA's depth range: every thread looks at one run (one coalesced
  //      load, one latency -- two dependent binary searches cost 14 L2 round trips before the first task could start)
  if (tid < kWG) {
    const int col = bn * fW + w0 + tid;
    s_a[tid] = __ldg(col_run_starts + col + 1);   // min over the runs in range (none: a = re >= b = rs -> zero tasks)
    s_b[tid] = __ldg(col_run_starts + col);
  }
  __syncthreads();
  {
    const int rs0 = __ldg(col_run_starts + bn * fW + w0), re3 = __ldg(col_run_starts + bn * fW + w0 + kWG);
    for (int r = rs0 + tid; r < re3; r += kP1Threads) {
      const int p0 = __ldg(run_p0 + r);
      const int d = (p0 / plane) % D;
      if (d >= d_lo && d < d_hi) {
        const int wl = p0 % fW - w0;
        atomicMin(&s_a[wl], r);
        atomicMax(&s_b[wl], r + 1);
      }
    }
  }
  // ---- ... everybody stages: context [c][h][w0..3] -> ctx_s[w][h][c], depth [d][h][w0..3] -> dm[w][d - d_lo][h] ----
  {
    // kSU independent 16-byte loads in flight per thread before the first transposed store.  (Measured alternatives:
    // one load at a time, same time; 4-byte cp.async doing the transposition in flight, 25 % slower -- the cost is the
    // number of scattered 32-byte sector requests, 1.7 M per frame, not their latency; more depth splits re-stage the
    // context and cost 10 % each.)
    constexpr int kSU = 6;
    const float *cbase = ctx + ((size_t)bn * C * fH) * fW + w0;
    const int n_items = C * fH;                // (h, c): one 16-byte vector = 4 columns
    {
      int h = tid / C, c = tid - h * C;        // consecutive threads -> consecutive c: conflict-free transposed stores
      const int sh = kP1Threads / C, sc = kP1Threads - sh * C;
      for (int i0 = tid; i0 < n_items; i0 += kP1Threads * kSU) {
        float4 v[kSU];
        int hh[kSU], cc[kSU];
#pragma unroll
        for (int u = 0; u < kSU; ++u) {
          hh[u] = h; cc[u] = c;
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i0 + u * kP1Threads < n_items)
            v[u] = __ldg(reinterpret_cast<const float4 *>(cbase + ((size_t)c * fH + h) * fW));
          h += sh; c += sc;
          if (c >= C) { c -= C; ++h; }
        }
#pragma unroll
        for (int u = 0; u < kSU; ++u) {
          if (i0 + u * kP1Threads < n_items) {
            ctx_s[(0 * fH + hh[u]) * csp + cc[u]] = v[u].x;
            ctx_s[(1 * fH + hh[u]) * csp + cc[u]] = v[u].y;
            ctx_s[(2 * fH + hh[u]) * csp + cc[u]] = v[u].z;
            ctx_s[(3 * fH + hh[u]) * csp + cc[u]] = v[u].w;
          }
        }
      }
    }
    const float *dbase = depth + (((size_t)bn * D + d_lo) * fH) * fW + w0;
    const int n_d = (d_hi - d_lo) * fH;
    {
      int dd = tid / fH, h = tid - dd * fH;
      const int sd = kP1Threads / fH, shh = kP1Threads - sd * fH;
      for (int i0 = tid; i0 < n_d; i0 += kP1Threads * kSU) {
        float4 v[kSU];
        int ofs[kSU];
#pragma unroll
        for (int u = 0; u < kSU; ++u) {
          ofs[u] = dd * dsp + h;
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i0 + u * kP1Threads < n_d) v[u] = __ldg(reinterpret_cast<const float4 *>(dbase + (size_t)(i0 + u * kP1Threads) * fW));
          dd += sd; h += shh;
          if (h >= fH) { h -= fH; ++dd; }
        }
#pragma unroll
        for (int u = 0; u < kSU; ++u) {
          if (i0 + u * kP1Threads < n_d) {
            dm[0 * dc * dsp + ofs[u]] = v[u].x;
            dm[1 * dc * dsp + ofs[u]] = v[u].y;
            dm[2 * dc * dsp + ofs[u]] = v[u].z;
            dm[3 * dc * dsp + ofs[u]] = v[u].w;
          }
        }
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    int t = 0;
    for (int w = 0; w < kWG; ++w) {
      s_toff[w] = t;
      t += max(0, s_b[w] - s_a[w] + kRB - 1) / kRB;
    }
    s_toff[kWG] = t;
  }
  __syncthreads();
  // ---- contraction: lane groups of C/8 lanes, each owning kRB runs x 8 channels (two float4: c, c + C/2) ----------
  const int g8 = C >> 3;                       // lanes per group
  const int nsub = 32 / g8;                    // groups per warp
  const int sub = lane / g8, cl = lane - sub * g8;
  const bool lane_on = sub < nsub;
  const int n_tasks = s_toff[kWG];
  const int half = C >> 1;
  for (int t0 = warp * nsub; t0 < n_tasks; t0 += (kP1Threads >> 5) * nsub) {
    const int task = t0 + sub;
    const bool on = lane_on && task < n_tasks;
    int wl = 0;
    if (on) {
#pragma unroll
      for (int w = 1; w < kWG; ++w) wl += (task >= s_toff[w]) ? 1 : 0;
    }
    const int r0 = on ? s_a[wl] + (task - s_toff[wl]) * kRB : 0;
    const int cnt = on ? min(kRB, s_b[wl] - r0) : 0;
    int wofs[kRB], h0[kRB], hl[kRB];
    int hmin = fH, hmax = 0;
#pragma unroll
    for (int j = 0; j < kRB; ++j) {
      wofs[j] = 0; h0[j] = 0; hl[j] = 0;
      if (j < cnt) {
        const int p0 = __ldg(run_p0 + r0 + j);
        const int d = (p0 / plane) % D;
        h0[j] = (p0 % plane) / fW;
        hl[j] = __ldg(run_len + r0 + j);
        wofs[j] = (wl * dc + (d - d_lo)) * dsp;
        hmin = min(hmin, h0[j]);
        hmax = max(hmax, h0[j] + hl[j]);
      }
    }
    Acc8 acc[kRB];
#pragma unroll
    for (int j = 0; j < kRB; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[j].v[q] = make_float2(0.f, 0.f);
    const float *crow = ctx_s + (size_t)wl * fH * csp + cl * 4;
    const float *dptr[kRB];
    int hin_lo = 0, hin_hi = fH;            // rows every run of the tile covers: no masks needed there
#pragma unroll
    for (int j = 0; j < kRB; ++j) {
      dptr[j] = dm + wofs[j];
      if (j < cnt) {
        hin_lo = max(hin_lo, h0[j]);
        hin_hi = min(hin_hi, h0[j] + hl[j]);
      }
    }
    if (cnt < kRB) { hin_lo = hmax; hin_hi = hmax; }   // a partial tile takes the masked path throughout
    hin_lo = min(max(hin_lo, hmin), hmax);
    hin_hi = max(min(hin_hi, hmax), hin_lo);
    auto masked = [&](int ha, int hb) {
      for (int h = ha; h < hb; ++h) {
        const float4 v0 = *reinterpret_cast<const float4 *>(crow + h * csp);
        const float4 v1 = *reinterpret_cast<const float4 *>(crow + h * csp + half);
#pragma unroll
        for (int j = 0; j < kRB; ++j) {
          const float wv = ((unsigned)(h - h0[j]) < (unsigned)hl[j]) ? dptr[j][h] : 0.f;
          fma8(acc[j], wv, v0, v1);
        }
      }
    };
    masked(hmin, hin_lo);
#pragma unroll 4
    for (int h = hin_lo; h < hin_hi; ++h) {
      const float4 v0 = *reinterpret_cast<const float4 *>(crow + h * csp);
      const float4 v1 = *reinterpret_cast<const float4 *>(crow + h * csp + half);
#pragma unroll
      for (int j = 0; j < kRB; ++j) fma8(acc[j], dptr[j][h], v0, v1);
    }
    masked(hin_hi, hmax);
#pragma unroll
    for (int j = 0; j < kRB; ++j) {
      if (j < cnt) {
        float *dst = partial + (size_t)__ldg(run_pos + r0 + j) * C + cl * 4;
        *reinterpret_cast<float4 *>(dst) = make_float4(acc[j].v[0].x, acc[j].v[0].y, acc[j].v[1].x, acc[j].v[1].y);
        *reinterpret_cast<float4 *>(dst + half) = make_float4(acc[j].v[2].x, acc[j].v[2].y, acc[j].v[3].x, acc[j].v[3].y);
      }
    }
  }
}

// phase 2: one CTA per output tile (line = (b*nz + z)*nx + x, 32 cells along y).  partial rows are in cell order: cell t's
// rows are [cell_run_starts[t], cell_run_starts[t + 1]) and the tile's rows one contiguous range (14 KB: L1-resident).
// A lane owns one cell; a warp takes 4 channels per step: <= 3 predicated 16-byte loads of the cell's rows, sums in
// ascending row order, four 128-byte channel-row stores.  No staging, no barriers after the cell table.
__global__ void __launch_bounds__(kP2Threads)
    bev_pool_p2_kernel(const float *__restrict__ partial, const int *__restrict__ cell_run_starts,
                       const int *__restrict__ icell, const int *__restrict__ tile_starts, int C, int nz, int nx, int ny,
                       int tiles_y, float *__restrict__ out) {
  __shared__ int s_rs[kTileY2], s_re[kTileY2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ty = blockIdx.x % tiles_y;
  const int line = blockIdx.x / tiles_y;
  const int y0 = ty * kTileY2;
  const int ycnt = min(kTileY2, ny - y0);
  const int cell0 = line * ny + y0;
  const int lo = __ldg(tile_starts + blockIdx.x), t_end = __ldg(tile_starts + blockIdx.x + 1);
  const int x_ = line % nx;
  const int bz = line / nx;   // b*nz + z  ->  channel block z*C of sample b: [B, nz*C, nx, ny]
  float *obase = out + (((size_t)bz * C) * nx + x_) * (size_t)ny + y0;
  const size_t cstride = (size_t)nx * ny;
  if (lo == t_end) {   // no point falls into this tile: zeros
    if (lane < ycnt)
      for (int ch = warp; ch < C; ch += kP2Threads / 32) obase[ch * cstride + lane] = 0.f;
    return;
  }
  if (tid < kTileY2) { s_rs[tid] = 0; s_re[tid] = 0; }
  __syncthreads();
  for (int t = lo + tid; t < t_end; t += kP2Threads) {
    const int yy = __ldg(icell + t) - cell0;
    s_rs[yy] = __ldg(cell_run_starts + t);
    s_re[yy] = __ldg(cell_run_starts + t + 1);
  }
  __syncthreads();
  const int rs = lane < ycnt ? s_rs[lane] : 0;
  const int nr = lane < ycnt ? s_re[lane] - rs : 0;
  const float *p0 = partial + (size_t)rs * C;
  for (int ch = warp * 4; ch < C; ch += (kP2Threads / 32) * 4) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (nr > 0) a = __ldg(reinterpret_cast<const float4 *>(p0 + ch));
    if (nr > 1) { const float4 v = __ldg(reinterpret_cast<const float4 *>(p0 + C + ch)); a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
    if (nr > 2) { const float4 v = __ldg(reinterpret_cast<const float4 *>(p0 + 2 * C + ch)); a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
    for (int r = 3; r < nr; ++r) {
      const float4 v = __ldg(reinterpret_cast<const float4 *>(p0 + (size_t)r * C + ch));
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    if (lane < ycnt) {
      float *o = obase + ch * cstride + lane;
      o[0] = a.x; o[cstride] = a.y; o[2 * cstride] = a.z; o[3 * cstride] = a.w;
    }
  }
}

}  // namespace

BEVF_API int bevf_bev_pool_fused_forward_v2(const float *depth, const float *ctx_nchw, const int *run_p0, const int *run_len,
                                            const int *run_pos, int n_runs, const int *col_run_starts,
                                            const int *cell_run_starts, const int *interval_cell, const int *tile_starts,
                                            int n_int, int bn, int d, int fh, int fw, int c, int b, int nz, int nx, int ny,
                                            float *partial, float *out, void *stream) {
  BEVF_CHECK_ARG(bn > 0 && d > 0 && fh > 0 && fw > 0 && b > 0 && nz > 0 && nx > 0 && ny > 0 && c > 0, "bad dims");
  // (C % 16 == 0 keeps the 4-channel steps of phase 2 inside the channel range; 80 = 5 x 16)
  BEVF_CHECK_ARG(out && depth && ctx_nchw && tile_starts && (n_runs == 0 || (partial && col_run_starts && run_pos)),
                 "NULL tensor");
  if (fw % kWG != 0 || c % 16 != 0 || c > 256) {
    bevf::set_error("fused bev_pool v2 needs fW %% 4 == 0 and C %% 8 == 0, C <= 256 (got fW %d, C %d)", fw, c);
    return BEVF_ERR_UNSUPPORTED;
  }
  BEVF_CHECK_ARG(((reinterpret_cast<uintptr_t>(depth) | reinterpret_cast<uintptr_t>(ctx_nchw) |
                   reinterpret_cast<uintptr_t>(partial)) & 15u) == 0, "depth / ctx / partial must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles_y = bevf::ceil_div(ny, kTileY2);
  const long long blocks2 = (long long)b * nz * nx * tiles_y;
  BEVF_CHECK_ARG(blocks2 < (1ll << 31), "BEV grid too large");
  if (n_runs > 0 && g_skip != 1) {
    // depth bins per CTA: enough CTAs for ~2 per SM, shared memory <= ~100 KB
    const int groups = bn * (fw / kWG);
    const int csp = c + 4, dsp = fh + 1;
    // as few depth splits as fill the SMs once (every split stages the context again)
    int dsplit = bevf::ceil_div(bevf::kNumSMs, groups);
    if (g_dsplit > 0) dsplit = g_dsplit;
    if (dsplit < 1) dsplit = 1;
    int dc = bevf::ceil_div(d, dsplit);
    const size_t ctx_bytes = (size_t)kWG * fh * csp * sizeof(float);
    while ((size_t)kWG * dc * dsp * sizeof(float) + ctx_bytes > 100 * 1024 && dc > 8) dc = (dc + 1) / 2;
    if (dc < 1) dc = 1;
    dsplit = bevf::ceil_div(d, dc);
    const size_t smem1 = ctx_bytes + (size_t)kWG * dc * dsp * sizeof(float);
    BEVF_CHECK_ARG(smem1 <= 200 * 1024, "pixel columns do not fit in shared memory (%zu bytes)", smem1);
    static bevf::DeviceOnce conf1;
    if (conf1.first())
      BEVF_CHECK_CUDA(cudaFuncSetAttribute(bev_pool_p1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    bev_pool_p1_kernel<<<dim3(groups, dsplit), kP1Threads, smem1, st>>>(depth, ctx_nchw, run_p0, run_len, run_pos,
                                                                       col_run_starts, d, fh, fw, c, dc, csp, dsp, partial);
    BEVF_CHECK_LAUNCH();
  }
  if (g_skip != 2) {
    bev_pool_p2_kernel<<<(unsigned)blocks2, kP2Threads, 0, st>>>(partial, cell_run_starts, interval_cell, tile_starts, c, nz,
                                                                nx, ny, tiles_y, out);
    BEVF_CHECK_LAUNCH();
  }
  return BEVF_OK;
}
