// hmma_rate.cu -- throughput of the legacy warp-level mma.sync.m16n8k16 (bf16 -> fp32) on sm_100a: cycles per
// instruction per SM for 4 / 8 / 16 / 32 resident warps, 8 independent accumulators per warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/hmma_rate scripts/hmma_rate.cu && build/hmma_rate
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int iters, float *out, long long *cyc) {
  float d[8][4];
  for (int i = 0; i < 8; ++i) for (int q = 0; q < 4; ++q) d[i][q] = 0.f;
  unsigned a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 5, a3 = 7, b0 = 11, b1 = 13;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) for (int q = 0; q < 4; ++q) s += d[i][q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float *out; long long *cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  for (int warps : {1, 4, 8, 16, 32}) {
    k<<<148, warps * 32>>>(iters, out, cyc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = (double)h[0];
    double mmas = (double)iters * 8 * warps;
    printf("warps/SM %2d: %.2f cycles per mma.m16n8k16 per SM  -> %.0f flops/clk/SM  (%.0f TFLOP/s at 148 SMs x 1.965 GHz)\n",
           warps, c / mmas, 4096.0 * mmas / c, 4096.0 * mmas / c * 148 * 1.965e9 / 1e12);
  }
  return 0;
}
