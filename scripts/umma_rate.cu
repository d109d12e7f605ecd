// umma_rate.cu -- cycles per tcgen05.mma (M = 128, K = 16, bf16 -> fp32) issued back to back by one thread into one
// accumulator, A operand from tensor memory ("TS") or from shared memory ("SS"), for N = 16 .. 256.  One CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/umma_rate scripts/umma_rate.cu && build/umma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xffffffff;\n\tselp.u32 %0, 1, 0, px;\n\t}" : "=r"(pred));
  return pred != 0;
}

template <int N, bool TS, int NACC, int AROT = 1, int BSBO = 0>
__global__ void __launch_bounds__(128) rate_kernel(long long *out, int reps) {
  __shared__ __align__(1024) uint8_t a_s[128 * 32];        // A: 128 rows x K = 16 bf16, 8 x 16 B core matrices
  __shared__ __align__(1024) uint8_t b_s[32 * 1024];       // B: up to 256 rows x K = 16 (or the gather-GEMM's weight-tile layout, BSBO > 0)
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 128 * 32 / 4; i += 128) reinterpret_cast<uint32_t *>(a_s)[i] = 0;
  for (int i = threadIdx.x; i < 32 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(b_s)[i] = 0;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  if (warp == 0) {
    const uint64_t adesc = make_desc(smem_u32(a_s), 2048u, 128u);
    // BSBO == 0: the 8-row groups of B are contiguous (SBO = 128 B) and the second K chunk follows them (LBO = N/8 * 128 B);
    // BSBO > 0: the gather-GEMM's weight tile [N/8][Cin/8][8 x 16 B]: K chunks adjacent (LBO = 128 B), row groups BSBO bytes apart
    const uint64_t bdesc = BSBO ? make_desc(smem_u32(b_s), 128u, (uint32_t)BSBO) : make_desc(smem_u32(b_s), (uint32_t)(N / 8) * 128u, 128u);
    const uint32_t a_t = tmem + 480u;     // 8 columns of A behind the accumulators (NACC * N <= 480)
    long long t0 = 0, t1 = 0, t2 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        const uint32_t d = tmem + (uint32_t)((r % NACC) * N);
        if (TS)   // AROT > 1: every MMA reads a different 8-column A slice (as the gather-GEMM does: one slice per tap and K step)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                       ::"r"(d), "r"(a_t - (uint32_t)((r % AROT) * 8)), "l"(bdesc), "r"(IDESC), "r"(1u) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(d), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(1u) : "memory");
      }
      t1 = clock64();
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    __syncwarp();
    asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}"
                 ::"r"(smem_u32(&bar)) : "memory");
    t2 = clock64();
    const long long issue = __shfl_sync(0xffffffffu, t1 - t0, 0), tot = t2 - __shfl_sync(0xffffffffu, t0, 0);
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = issue; out[1] = tot; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int N, bool TS, int NACC, int AROT = 1, int BSBO = 0>
void run(const char *name, long long *d_out) {
  const int reps = 512;
  long long h[2];
  for (int it = 0; it < 2; ++it) {
    rate_kernel<N, TS, NACC, AROT, BSBO><<<148, 128>>>(d_out, reps);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s N=%d: %s\n", name, N, cudaGetErrorString(e)); return; }
  }
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%s N=%3d accumulators=%d A slices=%d B row-group stride=%d: issue %.1f clk/MMA, issue+drain %.1f clk/MMA (ideal math %d)\n", name, N, NACC, AROT, BSBO ? BSBO : 128,
         (double)h[0] / reps, (double)h[1] / reps, 128 * N / 256);
}

int main() {
  long long *d_out;
  cudaMalloc(&d_out, 16);
  run<16, true, 1>("TS", d_out);  run<32, true, 1>("TS", d_out);  run<64, true, 1>("TS", d_out);
  run<128, true, 1>("TS", d_out); run<256, true, 1>("TS", d_out);
  run<16, false, 1>("SS", d_out); run<32, false, 1>("SS", d_out); run<64, false, 1>("SS", d_out);
  run<128, false, 1>("SS", d_out); run<256, false, 1>("SS", d_out);
  run<64, true, 2>("TS", d_out);  run<64, false, 2>("SS", d_out); run<128, true, 2>("TS", d_out);
  // rotating A slices (16 distinct 8-column slices below column 480; accumulators stay below column 352)
  run<16, true, 1, 16>("TS", d_out); run<32, true, 1, 16>("TS", d_out); run<64, true, 1, 16>("TS", d_out);
  run<128, true, 1, 16>("TS", d_out); run<64, true, 2, 16>("TS", d_out);
  // the gather-GEMM's weight-tile layout: row groups Cin/8 * 128 bytes apart (Cin = 32 / 64 / 128)
  run<32, true, 1, 16, 512>("TS", d_out); run<64, true, 1, 16, 1024>("TS", d_out); run<128, true, 1, 16, 2048>("TS", d_out);
  run<64, true, 1, 16, 512>("TS", d_out); run<128, true, 1, 16, 1024>("TS", d_out);
  return 0;
}
