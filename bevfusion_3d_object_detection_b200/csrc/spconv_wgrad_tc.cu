// spconv_wgrad_tc.cu -- weight gradient of the sparse convolution on the 5th-gen tensor cores (training, BASELINE
// configs[2]).
//
//   d_W[k][ci, co] = sum_j  feats[pair_fwd[k, j], ci] * d_out[j, co]            bf16 operands, fp32 accumulation in TMEM
//
// spconv computes it with the implicit-GEMM backward-weight kernels of cumm (mma.sync; reference call site
// projects/SparseConvolution/sparse_functional.py:287-314 and its backward).  Here the contraction runs over ROWS, so
// both tcgen05 operands are MN-major: a tile of gathered input rows [R, Cin] and a tile of output-gradient rows
// [R, Cout], each stored row after row at a 32 / 64 / 128-byte pitch in the matching swizzle, ARE the canonical
// MN-major shared-memory operands (K = row index; cute::UMMA canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte
// units) -- no transposition anywhere.  One tcgen05.mma has M = 128:
//   Cin = 16 / 32 / 64 : M = (8 / 4 / 2 taps) x Cin -- the tap tiles lie LBO bytes apart and share the B operand (d_out);
//   Cin = 128          : M = one tap's two 64-channel slabs;
// N = Cout, K = 16 rows per instruction.  A CTA owns 4 such "tap sets" (4 accumulators of Cout columns, at most 512 TMEM
// columns) and a range of output rows; it streams the rows in chunks of R = 64 through a 2-3 stage ring:
//   warps 0-7  producers: rulebook entries -> cp.async 16-byte gathers (zero-fill for missing pairs) straight into the
//              swizzled tiles, completion signalled on the stage's mbarrier (cp.async.mbarrier.arrive.noinc);
//   warp 8     tcgen05.mma issuer (one elected lane), tcgen05.commit frees the stage;
//   warps 0-3  epilogue at the end: tcgen05.ld, fp32 atomics into d_W[Cout, kv, Cin] (several row ranges per tap group).
// The fp32 FFMA kernel of spconv_bwd.cu stays as the fp32-parity path.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

constexpr int kR = 64;          // rows (K extent) of one chunk
constexpr int kSets = 4;        // tap sets (accumulators) per CTA
#ifndef BEVF_WG_PROD
#define BEVF_WG_PROD 256
#endif
constexpr int kProd = BEVF_WG_PROD;   // producer threads (the first 128 also run the epilogue: one TMEM lane each)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WG_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra WG_DONE;\n\t"
      "bra WG_WAIT;\n\t"
      "WG_DONE:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void *src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred px;\n\t"
      "elect.sync _|px, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// MN-major shared-memory matrix descriptor: start address, LBO = byte distance of the next MN block (one swizzle row
// width of elements), SBO = byte distance of the next group of 8 K rows, version 1, swizzle mode in [61, 64)
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

template <int CIN, int COUT>
struct WgCfg {
  static constexpr int PITCH_A = CIN * 2 < 128 ? CIN * 2 : 128;     // bytes of one gathered row inside a slab
  static constexpr int SLABS_A = CIN * 2 / PITCH_A;                 // 2 for Cin = 128
  static constexpr int T = CIN >= 128 ? 1 : 128 / CIN;              // taps per set (M = 128 = T * Cin)
  static constexpr int NBLK_A = T * SLABS_A;                        // MN blocks of one A operand (= 256 / PITCH_A)
  static constexpr int A_TILE = kR * PITCH_A;                       // one tap slab: kR rows
  static constexpr int A_SET = NBLK_A * A_TILE;                     // == 16 KB
  static constexpr int A_STAGE = kSets * A_SET;                     // == 64 KB
  static constexpr int PITCH_B = COUT * 2 < 128 ? COUT * 2 : 128;
  static constexpr int SLABS_B = COUT * 2 / PITCH_B;
  static constexpr int B_TILE = kR * PITCH_B;
  static constexpr int B_STAGE = SLABS_B * B_TILE;
  static constexpr int STAGE = A_STAGE + B_STAGE;                   // multiple of 1024
  static constexpr int NCH_A = CIN / 8, NCH_B = COUT / 8;           // 16-byte chunks per row
  static constexpr int CPS_A = PITCH_A / 16, CPS_B = PITCH_B / 16;  // chunks per slab row
  static constexpr uint32_t LAYOUT_A = PITCH_A == 128 ? 2u : PITCH_A == 64 ? 4u : 6u;
  static constexpr uint32_t LAYOUT_B = PITCH_B == 128 ? 2u : PITCH_B == 64 ? 4u : 6u;
  static constexpr int TMEM_COLS = kSets * COUT < 32 ? 32 : kSets * COUT;   // 64 / 128 / 256 / 512
  // D = f32, A = B = bf16, both MN-major (bits 15, 16), N >> 3 at [17, 23), M >> 4 at [24, 29)
  static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                    ((uint32_t)(COUT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  static_assert(A_SET == 16384 && STAGE % 1024 == 0, "tile arithmetic");
};

// byte offset of 16-byte chunk c of row r inside a tile of `pitch`-byte rows in the 32 / 64 / 128-byte swizzle
// (address bits [4, 7) ^= bits [7, 10), restricted to the chunks of one row)
template <int PITCH>
__device__ __forceinline__ uint32_t swz_off(int r, int c) {
  const uint32_t a0 = (uint32_t)r * PITCH;
  constexpr uint32_t mask = PITCH == 128 ? 7u : PITCH == 64 ? 3u : 1u;
  return a0 + ((((uint32_t)c) ^ ((a0 >> 7) & mask)) << 4);
}

template <int CIN, int COUT>
__global__ void __launch_bounds__(kProd + 32)
    spconv_wgrad_tc_kernel(const __nv_bfloat16 *__restrict__ feats, const __nv_bfloat16 *__restrict__ d_out,
                           const int *__restrict__ pair_fwd, int ld, int n_out, int kv, int cin_real, int stages,
                           int chunks_per_split, float *__restrict__ d_w) {
  using Cfg = WgCfg<CIN, COUT>;
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_off = (uint32_t)stages * Cfg::STAGE;
  const uint32_t bar_full = smem_base + bar_off;            // [stages] kProd async arrivals
  const uint32_t bar_empty = bar_full + 8 * 4;              // [stages] one tcgen05.commit
  const uint32_t bar_done = bar_empty + 8 * 4;              // accumulators complete
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + bar_off + 8 * 4 + 8 * 4 + 8);

  const int group = blockIdx.y;                             // tap sets [group * kSets, ...)
  const int set0 = group * kSets;
  const int n_sets = (kv + Cfg::T - 1) / Cfg::T;
  const int my_sets = min(kSets, n_sets - set0);
  const int n_chunks = (n_out + kR - 1) / kR;
  const int c_begin = blockIdx.x * chunks_per_split;
  const int c_end = min(c_begin + chunks_per_split, n_chunks);
  if (c_begin >= c_end || my_sets <= 0) return;             // whole CTA, before any barrier / TMEM state exists

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(bar_full + 8 * s, kProd);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kProd / 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(tmem_slot);

  if (warp < kProd / 32) {
    // ------------------------------------------ producers ------------------------------------------
    // work items of a chunk: (tap tile t, row r, 16-byte chunk c) of the gathered operand, then (row, chunk) of d_out;
    // consecutive threads take consecutive chunks of a row (coalesced 32..256-byte row reads)
    const uint8_t *feats_b = reinterpret_cast<const uint8_t *>(feats);
    const uint8_t *dout_b = reinterpret_cast<const uint8_t *>(d_out);
    const int taps_here = min(my_sets * Cfg::T, kv - set0 * Cfg::T);      // taps this CTA gathers
    constexpr int ITEMS_TAP = kR * Cfg::NCH_A;                            // items of one tap tile
    constexpr int IT_A = kSets * Cfg::T * ITEMS_TAP / kProd;              // items per thread (all 4 sets)
    constexpr int ITEMS_B = kR * Cfg::NCH_B;
    constexpr int IT_B = (ITEMS_B + kProd - 1) / kProd;
    int s = 0;
    uint32_t ph = 0;
    for (int ch = c_begin; ch < c_end; ++ch) {
      const int j0 = ch * kR;
      mbar_wait(bar_empty + 8 * s, ph ^ 1u);
      const uint32_t st = smem_base + (uint32_t)s * Cfg::STAGE;
      // rulebook entries first (independent loads in flight together), then the row gathers
      int idx[IT_A];
#pragma unroll
      for (int it = 0; it < IT_A; ++it) {
        const int i = it * kProd + tid;
        const int t = i / ITEMS_TAP, r = (i % ITEMS_TAP) / Cfg::NCH_A;
        const int j = j0 + r;
        idx[it] = (t < taps_here && j < n_out) ? __ldg(pair_fwd + (size_t)(set0 * Cfg::T + t) * ld + j) : -1;
      }
#pragma unroll
      for (int it = 0; it < IT_A; ++it) {
        const int i = it * kProd + tid;
        const int t = i / ITEMS_TAP, rem = i % ITEMS_TAP;
        const int r = rem / Cfg::NCH_A, c = rem % Cfg::NCH_A;
        if (t < taps_here) {   // tiles of taps past the kernel volume are never read by a valid accumulator row
          const uint32_t dst = st + (uint32_t)((t * Cfg::SLABS_A + c / Cfg::CPS_A) * Cfg::A_TILE) +
                               swz_off<Cfg::PITCH_A>(r, c % Cfg::CPS_A);
          const bool ok = idx[it] >= 0;
          cp_async16_zfill(dst, feats_b + ((size_t)(ok ? idx[it] : 0) * CIN + (size_t)c * 8) * 2, ok ? 16u : 0u);
        }
      }
#pragma unroll
      for (int it = 0; it < IT_B; ++it) {
        const int i = it * kProd + tid;
        if (i < ITEMS_B) {
          const int r = i / Cfg::NCH_B, c = i % Cfg::NCH_B;
          const int j = j0 + r;
          const bool ok = j < n_out;
          const uint32_t dst = st + (uint32_t)Cfg::A_STAGE + (uint32_t)((c / Cfg::CPS_B) * Cfg::B_TILE) +
                               swz_off<Cfg::PITCH_B>(r, c % Cfg::CPS_B);
          cp_async16_zfill(dst, dout_b + ((size_t)(ok ? j : 0) * COUT + (size_t)c * 8) * 2, ok ? 16u : 0u);
        }
      }
      cp_async_mbar_arrive_noinc(bar_full + 8 * s);   // arrives when this thread's copies above have landed
      if (++s == stages) { s = 0; ph ^= 1u; }
    }
    // ------------------------------------------ epilogue -------------------------------------------
    if (warp < 4) {                            // warps 0-3 own the four TMEM lane quadrants
    mbar_wait(bar_done, 0u);
    tc_fence_after();
    const int m = warp * 32 + lane;                        // accumulator row (TMEM lane): (tap in set, ci) or ci
    const int t_loc = CIN >= 128 ? 0 : m / CIN, ci = CIN >= 128 ? m : m % CIN;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
#pragma unroll 1
    for (int sidx = 0; sidx < my_sets; ++sidx) {
      const int k = (set0 + sidx) * Cfg::T + t_loc;
      const bool ok = k < kv && ci < cin_real;
#pragma unroll 1
      for (int c0 = 0; c0 < COUT; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tmem_base + lane_base + (uint32_t)(sidx * COUT + c0), r);
        if (ok) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float v = __uint_as_float(r[i]);
            if (v != 0.f) atomicAdd(d_w + ((size_t)(c0 + i) * kv + k) * cin_real + ci, v);
          }
        }
      }
    }
    }
    tc_fence_before();
  } else {
    // ------------------------------------------ MMA issuer -----------------------------------------
    int s = 0;
    uint32_t ph = 0;
    for (int ch = c_begin; ch < c_end; ++ch) {
      mbar_wait(bar_full + 8 * s, ph);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // cp.async (generic proxy) writes -> tcgen05 reads
      tc_fence_after();
      if (elect_one()) {
        const uint32_t st = smem_base + (uint32_t)s * Cfg::STAGE;
        const uint32_t b_s = st + (uint32_t)Cfg::A_STAGE;
        for (int sidx = 0; sidx < my_sets; ++sidx) {
          const uint32_t a_s = st + (uint32_t)(sidx * Cfg::A_SET);
#pragma unroll
          for (int kk = 0; kk < kR / 16; ++kk) {
            const uint64_t adesc = make_desc_mn(a_s + (uint32_t)(kk * 16 * Cfg::PITCH_A), (uint32_t)Cfg::A_TILE,
                                                8u * Cfg::PITCH_A, Cfg::LAYOUT_A);
            const uint64_t bdesc = make_desc_mn(b_s + (uint32_t)(kk * 16 * Cfg::PITCH_B), (uint32_t)Cfg::B_TILE,
                                                8u * Cfg::PITCH_B, Cfg::LAYOUT_B);
            umma_bf16(tmem_base + (uint32_t)(sidx * COUT), adesc, bdesc, Cfg::IDESC, (ch > c_begin || kk > 0) ? 1u : 0u);
          }
        }
        umma_commit(bar_empty + 8 * s);
        if (ch + 1 == c_end) umma_commit(bar_done);
      }
      __syncwarp();
      if (++s == stages) { s = 0; ph ^= 1u; }
    }
  }
  __syncthreads();
  if (warp == kProd / 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
  }
}

template <int CIN, int COUT>
int launch_wgrad_tc(const __nv_bfloat16 *feats, const __nv_bfloat16 *d_out, const int *pair_fwd, int ld, int n_out, int kv,
                    int cin_real, float *d_w, cudaStream_t st) {
  using Cfg = WgCfg<CIN, COUT>;
  static bevf::DeviceOnce configured;
  if (configured.first()) {
    BEVF_CHECK_CUDA(cudaFuncSetAttribute(spconv_wgrad_tc_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024));
  }
  const int fixed = 1024 + 128;   // alignment slack + barriers + TMEM slot
  int stages = (227 * 1024 - fixed) / Cfg::STAGE;
  if (stages > 3) stages = 3;
  if (stages < 2) {
    bevf::set_error("wgrad (tensor cores): a stage of %d bytes does not fit twice", Cfg::STAGE);
    return BEVF_ERR_UNSUPPORTED;
  }
  const int n_sets = bevf::ceil_div(kv, Cfg::T);
  const int groups = bevf::ceil_div(n_sets, kSets);
  const int n_chunks = bevf::ceil_div(n_out, kR);
  // two waves of CTAs (one CTA per SM: the ring takes the shared memory), at least 4 chunks per CTA
  int splits = bevf::ceil_div(2 * bevf::kNumSMs, groups);
  const int max_splits = bevf::ceil_div(n_chunks, 4);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  const int chunks_per_split = bevf::ceil_div(n_chunks, splits);
  splits = bevf::ceil_div(n_chunks, chunks_per_split);
  dim3 grid(splits, groups);
  spconv_wgrad_tc_kernel<CIN, COUT><<<grid, kProd + 32, stages * Cfg::STAGE + fixed, st>>>(
      feats, d_out, pair_fwd, ld, n_out, kv, cin_real, stages, chunks_per_split, d_w);
  BEVF_CHECK_LAUNCH();
  return BEVF_OK;
}

}  // namespace

BEVF_API int bevf_spconv_wgrad_bf16(const void *feats_bf16, const void *d_out_bf16, const int *pair_fwd, int ld, int n_out,
                                    int kv, int cin, int cin_pad, int cout, float *d_weight_okc, void *stream) {
  BEVF_CHECK_ARG(kv > 0 && cin > 0 && cin_pad >= cin && cout > 0 && ld >= n_out && n_out >= 0, "bad wgrad shape");
  BEVF_CHECK_ARG(d_weight_okc, "NULL weight gradient");
  cudaStream_t st = (cudaStream_t)stream;
  BEVF_CHECK_CUDA(cudaMemsetAsync(d_weight_okc, 0, sizeof(float) * (size_t)cout * kv * cin, st));
  if (n_out == 0) return BEVF_OK;
  BEVF_CHECK_ARG(feats_bf16 && d_out_bf16 && pair_fwd, "NULL tensor");
  BEVF_CHECK_ARG(((reinterpret_cast<uintptr_t>(feats_bf16) | reinterpret_cast<uintptr_t>(d_out_bf16)) & 15u) == 0,
                 "bf16 operands must be 16-byte aligned");
  const __nv_bfloat16 *f = (const __nv_bfloat16 *)feats_bf16, *g = (const __nv_bfloat16 *)d_out_bf16;
#define BEVF_WGT_CASE(CI, CO) \
  if (cin_pad == CI && cout == CO) return launch_wgrad_tc<CI, CO>(f, g, pair_fwd, ld, n_out, kv, cin, d_weight_okc, st)
  BEVF_WGT_CASE(16, 16); BEVF_WGT_CASE(16, 32); BEVF_WGT_CASE(16, 64); BEVF_WGT_CASE(16, 128);
  BEVF_WGT_CASE(32, 16); BEVF_WGT_CASE(32, 32); BEVF_WGT_CASE(32, 64); BEVF_WGT_CASE(32, 128);
  BEVF_WGT_CASE(64, 16); BEVF_WGT_CASE(64, 32); BEVF_WGT_CASE(64, 64); BEVF_WGT_CASE(64, 128);
  BEVF_WGT_CASE(128, 16); BEVF_WGT_CASE(128, 32); BEVF_WGT_CASE(128, 64); BEVF_WGT_CASE(128, 128);
#undef BEVF_WGT_CASE
  bevf::set_error("wgrad (tensor cores): unsupported channel pair %d -> %d", cin_pad, cout);
  return BEVF_ERR_UNSUPPORTED;
}
