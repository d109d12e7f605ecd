// spconv_rg.cu -- implicit gather-GEMM for the 16-CHANNEL sparse-conv layers: operand rows gathered straight from L2
// into warp-level MMA fragments ("register gather").
//
//   out[j, :] = epilogue( sum_k  feats_bf16[pair_fwd[k, j], :] @ W_bf16[k] )      bf16 operands, fp32 accumulation
//
// Why a second kernel.  At the first level of the BEVFusion encoder (conv_input, four 16 -> 16 SubM layers, the 16 -> 32
// strided conv) only 7-13 % of the (row, tap) pairs exist and a tap is K = 16: the tensor pipe is idle whatever the
// kernel, and the tcgen05 kernel (spconv_tc.cu) pays its per-item latency chain (gather -> tcgen05.st -> wait::st ->
// mbarrier -> tcgen05.mma -> commit -> mbarrier) for every tap of every tile, 1.9 cycles per (row, tap) per SM.  Here
// there is no ring: a warp owns 16 output rows, requests the rulebook entries of all 27 taps at once, its lanes load the
// neighbours' rows with 8-byte gathers directly in the mma.sync.m16n8k16 A-fragment layout (the K order inside a tap is
// free, so the weights are packed in the order the lanes load the channels: a 32-byte row is ONE sector read by the 4
// lanes of a quad), missing pairs load nothing, accumulators stay in registers for all taps, 24 warps per SM hide the
// L2 latency.  Weights live in shared memory in fragment order (LDS.64, conflict-free).
// Measured (config A, per launch): 16 -> 16 23.4 us (tcgen05 kernel 28.5), conv_input 30.5 (40), 16 -> 32 48 (53); ncu:
// issue slots 59 % busy -- 41 instructions per (16 rows, tap), most of them spent on pairs that do not exist.
// NOT used at Cin = 32 (BEVF_RG_CIN32): with 11 valid pairs per row the gathers re-read 225 MB per layer through L2
// where the tcgen05 kernel's TMA halo moves 77 MB (measured 80 vs 56 us).  Legacy mma.sync peaks at 593 TFLOP/s on B200
// (scripts/hmma_rate.cu: 2 cycles per m16n8k16 per SM), 36 % of the measured tcgen05 rate: irrelevant at K = 16, and
// the reason the 64- and 128-channel layers (92 % of the flops) stay on tcgen05 / TMEM.
//
// Reference boundary: spconv's implicit GEMM (projects/SparseConvolution/sparse_functional.py:287-314); same rulebook
// layout (pair_fwd[kv, n_out], -1 = no input), same fused epilogue as the other kernels.
#include <cuda_bf16.h>

#include "common.cuh"
#include "spconv_internal.cuh"

namespace {

constexpr int kRgWarps = 8;
constexpr int kMaxKv = 27;      // kernel volumes up to 3 x 3 x 3 (a multiple of every TG below)
#ifndef BEVF_RG_TG16
#define BEVF_RG_TG16 3   /* taps gathered together at Cin = 16 */
#endif
#ifndef BEVF_RG_TG32
#define BEVF_RG_TG32 3   /* ... at Cin = 32 */
#endif
#ifndef BEVF_RG_MT16
#define BEVF_RG_MT16 1   /* m16 tiles per warp iteration at Cin = 16 (Cout <= 32); measured: 1 -> 23.4 us, 2 -> 28 us per 16->16 layer */
#endif
#ifndef BEVF_RG_MT32
#define BEVF_RG_MT32 2   /* ... at Cin = 32 (Cout <= 32) */
#endif
#ifndef BEVF_RG_CIN32
#define BEVF_RG_CIN32 0  /* 1: also the 32-channel layers (measured slower than the tcgen05 kernel there: at 11 valid pairs per
                            row the gathers re-read 225 MB through L2 per layer where the TMA halo moves 77 MB) */
#endif
#ifndef BEVF_RG_MINB
#define BEVF_RG_MINB 1   /* __launch_bounds__ min blocks per SM (register cap) */
#endif

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// one lane's share of a feature row: CIN / 4 channels = CIN / 2 bytes (8 B at Cin 16, 16 B at Cin 32)
template <int CIN>
struct RowPart;
template <>
struct RowPart<16> {
  uint2 v;
  __device__ __forceinline__ void zero() { v = make_uint2(0u, 0u); }
  __device__ __forceinline__ void load(const uint8_t *row, int t) { v = __ldg(reinterpret_cast<const uint2 *>(row) + t); }
  __device__ __forceinline__ uint32_t lo(int) const { return v.x; }   // logical k 2t, 2t+1   of k-step s
  __device__ __forceinline__ uint32_t hi(int) const { return v.y; }   // logical k 2t+8, 2t+9
};
template <>
struct RowPart<32> {
  uint4 v;
  __device__ __forceinline__ void zero() { v = make_uint4(0u, 0u, 0u, 0u); }
  __device__ __forceinline__ void load(const uint8_t *row, int t) { v = __ldg(reinterpret_cast<const uint4 *>(row) + t); }
  __device__ __forceinline__ uint32_t lo(int s) const { return s == 0 ? v.x : v.z; }
  __device__ __forceinline__ uint32_t hi(int s) const { return s == 0 ? v.y : v.w; }
};

// MT = m16 tiles per warp iteration (2: 32 rows, 1: 16 rows when the accumulators of 64 output channels need the room)
template <int CIN, int COUT, int MT>
__global__ void __launch_bounds__(kRgWarps * 32, BEVF_RG_MINB)
    spconv_rg_kernel(const __nv_bfloat16 *__restrict__ feats, const uint2 *__restrict__ wfrag,
                     const int *__restrict__ pair_fwd, int ld, int n_out_host, const int *__restrict__ n_out_dev, int kv,
                     bevf::SpconvEpilogue ep, float *__restrict__ out_f32, __nv_bfloat16 *__restrict__ out_bf16) {
  constexpr int KS = CIN / 16;          // k-steps per tap
  constexpr int NB = COUT / 8;          // n8 blocks
  constexpr int ROWS = 16 * MT;         // rows per warp iteration
  constexpr int TG = CIN == 16 ? BEVF_RG_TG16 : BEVF_RG_TG32;   // taps gathered together (loads in flight per lane: TG * 2 * MT)
  extern __shared__ __align__(16) uint8_t smem_rg[];
  uint2 *wsm = reinterpret_cast<uint2 *>(smem_rg);                       // [kv][KS][NB][32]
  float *eps = reinterpret_cast<float *>(smem_rg + (size_t)kv * KS * NB * 32 * sizeof(uint2));   // bias | scale | shift
  const int n_out = n_out_dev ? min(*n_out_dev, ld) : n_out_host;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  {  // weights (constants) -> shared memory, 16 bytes per thread per step
    const uint4 *src = reinterpret_cast<const uint4 *>(wfrag);
    uint4 *dst = reinterpret_cast<uint4 *>(wsm);
    const int n16 = kv * KS * NB * 32 / 2;
    for (int i = tid; i < n16; i += kRgWarps * 32) dst[i] = __ldg(src + i);
    for (int c = tid; c < COUT; c += kRgWarps * 32) {
      eps[c] = ep.bias ? ep.bias[c] : 0.f;
      eps[COUT + c] = ep.scale ? ep.scale[c] : 1.f;
      eps[2 * COUT + c] = ep.scale ? ep.shift[c] : 0.f;
    }
  }
  __syncthreads();
  const int g = lane >> 2, t = lane & 3;
  const uint8_t *feats_b = reinterpret_cast<const uint8_t *>(feats);
  const int n_tiles = (n_out + ROWS - 1) / ROWS;
  const int n_warps = gridDim.x * kRgWarps;
  for (int tile = blockIdx.x * kRgWarps + warp; tile < n_tiles; tile += n_warps) {
    const int j0 = tile * ROWS;
    // the rulebook is streamed once from DRAM: pull the lines of this warp's next tile into L2 now (one instruction), so
    // that the register prefetch below (one chunk ahead) sees an L2 latency, not a DRAM latency, per chunk
    {
      const int jn = (tile + n_warps) * ROWS;
      if (lane < kv && jn < n_out) asm volatile("prefetch.global.L2 [%0];" ::"l"(pair_fwd + (size_t)lane * ld + jn));
    }
    float acc[MT][NB][4];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int j = 0; j < NB; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[m][j][q] = 0.f;
    // Rulebook entries of ALL taps first: lane L holds the entry of row j0 + L, one coalesced 128-byte line per tap and
    // warp, kv independent loads in flight per warp.  The rulebook is streamed once from DRAM, and it is this stream's
    // latency, not the gathers, that bounds the layer when only a chunk of taps is requested ahead (measured: 23 us per
    // 16 -> 16 layer with 3 lines per warp in flight).
    const int jr = j0 + lane;
    const bool row_here = lane < ROWS && jr < n_out;
    int idx[kMaxKv];
#pragma unroll
    for (int u = 0; u < kMaxKv; ++u) idx[u] = (row_here && u < kv) ? __ldg(pair_fwd + (size_t)u * ld + jr) : -1;
#pragma unroll
    for (int k0 = 0; k0 < kMaxKv; k0 += TG) {
      if (k0 >= kv) break;
      // gathers of the chunk: rows g, g + 8 (, g + 16, g + 24) of TG taps, all independent; missing pairs load nothing
      RowPart<CIN> a[TG][2 * MT];
      unsigned any = 0u;
#pragma unroll
      for (int u = 0; u < TG; ++u) {
#pragma unroll
        for (int h = 0; h < 2 * MT; ++h) {
          const int i = __shfl_sync(0xffffffffu, idx[k0 + u], g + 8 * h);
          a[u][h].zero();
          if (i >= 0) a[u][h].load(feats_b + (size_t)i * (CIN * 2), t);
        }
        any |= (__ballot_sync(0xffffffffu, idx[k0 + u] >= 0) != 0u) ? (1u << u) : 0u;
      }
#pragma unroll
      for (int u = 0; u < TG; ++u) {
        if (!((any >> u) & 1u)) continue;          // no row of the tile has this tap (warp-uniform)
        const uint2 *wk = wsm + (size_t)(k0 + u) * (KS * NB * 32) + lane;
#pragma unroll
        for (int s = 0; s < KS; ++s) {
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            const uint2 b = wk[(s * NB + j) * 32];
#pragma unroll
            for (int m = 0; m < MT; ++m)
              mma_bf16_16816(acc[m][j], a[u][2 * m].lo(s), a[u][2 * m + 1].lo(s), a[u][2 * m].hi(s), a[u][2 * m + 1].hi(s),
                             b.x, b.y);
          }
        }
      }
    }
    // epilogue: lane holds (row g [+8] of m-tile m, channels 8 j + 2 t, + 1)
#pragma unroll
    for (int m = 0; m < MT; ++m) {
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int j_row = j0 + 16 * m + g + 8 * hh;
        if (j_row >= n_out) continue;
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const int c = 8 * j + 2 * t;
          float x0 = acc[m][j][2 * hh], x1 = acc[m][j][2 * hh + 1];
          const float2 bb = *reinterpret_cast<const float2 *>(eps + c);
          const float2 ss = *reinterpret_cast<const float2 *>(eps + COUT + c);
          const float2 sh = *reinterpret_cast<const float2 *>(eps + 2 * COUT + c);
          x0 = fmaf(x0 + bb.x, ss.x, sh.x);
          x1 = fmaf(x1 + bb.y, ss.y, sh.y);
          if (ep.residual) {
            const float2 r = __ldg(reinterpret_cast<const float2 *>(ep.residual + (size_t)j_row * COUT + c));
            x0 += r.x; x1 += r.y;
          } else if (ep.residual_bf16) {
            const uint32_t r = __ldg(reinterpret_cast<const uint32_t *>(ep.residual_bf16 + (size_t)j_row * COUT + c));
            x0 += __uint_as_float(r << 16);
            x1 += __uint_as_float(r & 0xffff0000u);
          }
          if (ep.relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
          if (out_f32) *reinterpret_cast<float2 *>(out_f32 + (size_t)j_row * COUT + c) = make_float2(x0, x1);
          if (out_bf16) {
            const __nv_bfloat162 p = __floats2bfloat162_rn(x0, x1);
            *reinterpret_cast<uint32_t *>(out_bf16 + (size_t)j_row * COUT + c) = *reinterpret_cast<const uint32_t *>(&p);
          }
        }
      }
    }
  }
}

// weight fp32 [Cout, kv, Cin] -> fragment order [kv][KS][NB][32 lanes] x (b0, b1): lane (g, t) of n8 block j holds column
// n = 8 j + g and the four channels base .. base + 3, base = t * (cin_pad / 4) + 4 s, i.e. the channels the lane's own
// gather of k-step s delivers (logical k 2t, 2t+1 | 2t+8, 2t+9 of mma.m16n8k16)
__global__ void pack_weight_rg_kernel(const float *__restrict__ w, uint2 *__restrict__ out, int kv, int cin, int cin_pad,
                                      int cout) {
  const int KS = cin_pad / 16, NB = cout / 8;
  const long long total = (long long)kv * KS * NB * 32;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int lane = (int)(e % 32);
  const int j = (int)((e / 32) % NB);
  const int s = (int)((e / (32 * NB)) % KS);
  const int k = (int)(e / ((long long)32 * NB * KS));
  const int g = lane >> 2, t = lane & 3;
  const int n = 8 * j + g;
  const int base = t * (cin_pad / 4) + 4 * s;
  float v[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) v[q] = (base + q < cin) ? w[((size_t)n * kv + k) * cin + base + q] : 0.f;
  const __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
  out[e] = make_uint2(*reinterpret_cast<const uint32_t *>(&p0), *reinterpret_cast<const uint32_t *>(&p1));
}

template <int CIN, int COUT, int MT>
int launch_rg(const __nv_bfloat16 *feats, const uint8_t *w_rg, const int *pair_fwd, int ld, int n_out, const int *n_out_dev,
              int kv, const bevf::SpconvEpilogue &ep, float *out_f32, __nv_bfloat16 *out_bf16, cudaStream_t st) {
  const size_t smem = (size_t)kv * (CIN / 16) * (COUT / 8) * 32 * sizeof(uint2) + 3 * COUT * sizeof(float);
  if (smem > 200 * 1024) {
    bevf::set_error("sparse conv (register gather): %zu bytes of weights do not fit in shared memory", smem);
    return BEVF_ERR_UNSUPPORTED;
  }
  static bevf::DeviceOnce configured;
  if (configured.first())
    BEVF_CHECK_CUDA(cudaFuncSetAttribute(spconv_rg_kernel<CIN, COUT, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         200 * 1024));
  static thread_local int occ = 0, occ_dev = -1;
  static thread_local size_t occ_smem = 0;
  int dev_now = 0;
  cudaGetDevice(&dev_now);
  if (occ_dev != dev_now || occ_smem != smem) {
    BEVF_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spconv_rg_kernel<CIN, COUT, MT>, kRgWarps * 32, smem));
    if (occ < 1) occ = 1;
    occ_dev = dev_now;
    occ_smem = smem;
  }
  const int rows = n_out_dev ? ld : n_out;
  const int tiles = bevf::ceil_div(rows, 16 * MT);
  int grid = bevf::kNumSMs * occ;
  const int need = bevf::ceil_div(tiles, kRgWarps);
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  spconv_rg_kernel<CIN, COUT, MT><<<grid, kRgWarps * 32, smem, st>>>(feats, reinterpret_cast<const uint2 *>(w_rg), pair_fwd,
                                                                    ld, n_out, n_out_dev, kv, ep, out_f32, out_bf16);
  BEVF_CHECK_LAUNCH();
  return BEVF_OK;
}

}  // namespace

namespace bevf {

bool spconv_rg_supported(int cin_pad, int cout) {
  return (cin_pad == 16 || (BEVF_RG_CIN32 && cin_pad == 32)) && (cout == 16 || cout == 32 || cout == 64);
}

int spconv_rg_pack_weight(const float *w_okc, void *w_rg, int kv, int cin, int cin_pad, int cout, cudaStream_t st) {
  const long long total = (long long)kv * (cin_pad / 16) * (cout / 8) * 32;
  pack_weight_rg_kernel<<<ceil_div(total, 256), 256, 0, st>>>(w_okc, reinterpret_cast<uint2 *>(w_rg), kv, cin, cin_pad, cout);
  BEVF_CHECK_LAUNCH();
  return BEVF_OK;
}

int spconv_rg_launch(int cin_pad, int cout, const void *feats_bf16, const void *w_rg, const int *pair_fwd, int ld, int n_out,
                     const int *n_out_dev, int kv, const SpconvEpilogue &ep, float *out_f32, void *out_bf16,
                     cudaStream_t st) {
  const __nv_bfloat16 *f = (const __nv_bfloat16 *)feats_bf16;
  const uint8_t *w = (const uint8_t *)w_rg;
  __nv_bfloat16 *ob = (__nv_bfloat16 *)out_bf16;
#define BEVF_RG_CASE(CI, CO, MT) \
  if (cin_pad == CI && cout == CO) return launch_rg<CI, CO, MT>(f, w, pair_fwd, ld, n_out, n_out_dev, kv, ep, out_f32, ob, st)
  BEVF_RG_CASE(16, 16, BEVF_RG_MT16);
  BEVF_RG_CASE(16, 32, BEVF_RG_MT16);
  BEVF_RG_CASE(32, 16, BEVF_RG_MT32);
  BEVF_RG_CASE(32, 32, BEVF_RG_MT32);
  BEVF_RG_CASE(16, 64, 1);
  BEVF_RG_CASE(32, 64, 1);
#undef BEVF_RG_CASE
  set_error("sparse conv (register gather): unsupported channel pair %d -> %d", cin_pad, cout);
  return BEVF_ERR_UNSUPPORTED;
}

}  // namespace bevf
