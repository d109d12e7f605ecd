// bitmap_rank.cuh -- occupancy bitmap + popcount prefix = "sorted unique with rank" in O(cells/32).
//
// Used by the dynamic scatter (rows of unique_dim(sorted=true)) and by the sparse-conv coordinate index
// (coordinate -> row lookup, sorted output sites of strided convolutions).  A key k is present iff bit k of
// the bitmap is set; rank(k) = word_prefix[k >> 5] + popc(bitmap[k >> 5] & ((1 << (k & 31)) - 1)) is its
// position among the present keys in ascending order.
#pragma once
#include "common.cuh"

namespace bevf {

constexpr int kRankThreads = 256;
constexpr int kRankItems = 4;
constexpr int kRankTile = kRankThreads * kRankItems;

inline int rank_num_blocks(long long nwords) {
  long long b = (nwords + kRankTile - 1) / kRankTile;
  return b < 1 ? 1 : (int)b;
}

static __global__ void __launch_bounds__(kRankThreads)
    rank_popc_count_kernel(const unsigned *__restrict__ bitmap, long long nwords, int *__restrict__ block_counts) {
  __shared__ int ws[kRankThreads / 32];
  long long base = (long long)blockIdx.x * kRankTile;
  int cnt = 0;
#pragma unroll
  for (int k = 0; k < kRankItems; ++k) {
    long long w = base + (long long)k * kRankThreads + threadIdx.x;
    if (w < nwords) cnt += __popc(bitmap[w]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < kRankThreads / 32; ++w) t += ws[w];
    block_counts[blockIdx.x] = t;
  }
}

// exclusive prefix of block_counts in place (single 1024-thread block); *total = number of set bits
static __global__ void __launch_bounds__(1024)
    rank_block_scan_kernel(int *__restrict__ block_counts, int nblk, int *__restrict__ total) {
  __shared__ int ws[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblk; base += 1024) {
    int i = base + threadIdx.x;
    int v = (i < nblk) ? block_counts[i] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int u = __shfl_up_sync(0xffffffffu, inc, o);
      if ((int)(threadIdx.x & 31) >= o) inc += u;
    }
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = ws[threadIdx.x];
      int winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, winc, o);
        if ((int)threadIdx.x >= o) winc += u;
      }
      ws[threadIdx.x] = winc - w;
    }
    __syncthreads();
    int excl = carry + ws[threadIdx.x >> 5] + inc - v;
    if (i < nblk) block_counts[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

// word_prefix[w] = set bits in words < w; emit(rank, key) is called for every set bit in ascending order
template <typename Emit>
__global__ void __launch_bounds__(kRankThreads)
    rank_scan_emit_kernel(const unsigned *__restrict__ bitmap, long long nwords,
                          const int *__restrict__ block_offsets, int *__restrict__ word_prefix, Emit emit) {
  __shared__ int ws[kRankThreads / 32];
  const long long base = (long long)blockIdx.x * kRankTile + (long long)threadIdx.x * kRankItems;
  unsigned wv[kRankItems];
  int local = 0;
#pragma unroll
  for (int k = 0; k < kRankItems; ++k) {
    long long w = base + k;
    wv[k] = (w < nwords) ? bitmap[w] : 0u;
    local += __popc(wv[k]);
  }
  int inc = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int u = __shfl_up_sync(0xffffffffu, inc, o);
    if ((int)(threadIdx.x & 31) >= o) inc += u;
  }
  if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = inc;
  __syncthreads();
  if (threadIdx.x < 32) {
    int w = (threadIdx.x < kRankThreads / 32) ? ws[threadIdx.x] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int u = __shfl_up_sync(0xffffffffu, winc, o);
      if ((int)threadIdx.x >= o) winc += u;
    }
    if (threadIdx.x < kRankThreads / 32) ws[threadIdx.x] = winc - w;
  }
  __syncthreads();
  int run = block_offsets[blockIdx.x] + ws[threadIdx.x >> 5] + inc - local;
#pragma unroll
  for (int k = 0; k < kRankItems; ++k) {
    long long w = base + k;
    if (w < nwords) {
      word_prefix[w] = run;
      unsigned bits = wv[k];
      while (bits) {
        int b = __ffs(bits) - 1;
        bits &= bits - 1;
        emit(run, ((unsigned long long)w << 5) | (unsigned)b);
        run += 1;
      }
    }
  }
}

// bitmap must already hold the marks; block_counts has rank_num_blocks(nwords) ints
template <typename Emit>
int rank_build(const unsigned *bitmap, long long nwords, int *word_prefix, int *block_counts, int *total_dev,
               Emit emit, cudaStream_t st) {
  const int nblk = rank_num_blocks(nwords);
  rank_popc_count_kernel<<<nblk, kRankThreads, 0, st>>>(bitmap, nwords, block_counts);
  BEVF_CHECK_LAUNCH();
  rank_block_scan_kernel<<<1, 1024, 0, st>>>(block_counts, nblk, total_dev);
  BEVF_CHECK_LAUNCH();
  rank_scan_emit_kernel<<<nblk, kRankThreads, 0, st>>>(bitmap, nwords, block_counts, word_prefix, emit);
  BEVF_CHECK_LAUNCH();
  return BEVF_OK;
}

}  // namespace bevf
