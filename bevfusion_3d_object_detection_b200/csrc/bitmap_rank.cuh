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

// ---- one-launch form -------------------------------------------------------------------------------------------------
// The bitmap is cut into <= kRankMaxChunks contiguous chunks (about two per SM).  A CTA takes the next chunk from an atomic
// ticket (so every predecessor of a running chunk is itself running or done: the waits below cannot deadlock), counts
// its set bits, publishes the count, sums its predecessors' counts (<= 591 flags, read 256 at a time), and scans its
// chunk a second time -- from L2 -- writing word_prefix and emitting.  One launch instead of three and no per-tile
// partials; `sync` = 1 + kRankMaxChunks ints, zeroed by the caller before the launch.
constexpr int kRankMaxChunks = 592;

template <typename Emit>
__device__ __forceinline__ void rank_chunk_scan_body(const unsigned *__restrict__ bitmap, long long nwords,
                                                     long long chunk_words, int nchunks, int *__restrict__ sync,
                                                     int *__restrict__ word_prefix, int *__restrict__ total, Emit emit) {
  __shared__ int ws[kRankThreads / 32];
  __shared__ int s_tile, s_val;
  int *counter = sync, *status = sync + 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(counter, 1);
  __syncthreads();
  const int tile = s_tile;
  if (tile >= nchunks) return;
  const long long w0 = (long long)tile * chunk_words;
  const long long w1 = w0 + chunk_words < nwords ? w0 + chunk_words : nwords;
  auto block_sum = [&](int v) {   // every thread gets the CTA-wide sum
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) ws[warp] = v;
    __syncthreads();
    int t = 0;
#pragma unroll
    for (int w = 0; w < kRankThreads / 32; ++w) t += ws[w];
    return t;
  };
  int cnt = 0;
  for (long long w = w0 + tid; w < w1; w += kRankThreads) cnt += __popc(__ldg(bitmap + w));
  const int agg = block_sum(cnt);
  if (tid == 0) atomicExch(status + tile, agg + 1);   // 0 = not yet published
  int before = 0;
  for (int t = tid; t < tile; t += kRankThreads) {
    int v;
    do {
      v = *reinterpret_cast<volatile int *>(status + t);
    } while (v == 0);
    before += v - 1;
  }
  int run_base = block_sum(before);
  if (tile == nchunks - 1 && tid == 0) *total = run_base + agg;
  for (long long tb = w0; tb < w1; tb += kRankTile) {
    const long long base = tb + (long long)tid * kRankItems;
    unsigned wv[kRankItems];
    int local = 0;
#pragma unroll
    for (int k = 0; k < kRankItems; ++k) {
      const long long w = base + k;
      wv[k] = (w < w1) ? __ldg(bitmap + w) : 0u;
      local += __popc(wv[k]);
    }
    int inc = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    __syncthreads();
    if (lane == 31) ws[warp] = inc;
    __syncthreads();
    int wpre = 0, ttot = 0;
#pragma unroll
    for (int w = 0; w < kRankThreads / 32; ++w) {
      if (w < warp) wpre += ws[w];
      ttot += ws[w];
    }
    int run = run_base + wpre + inc - local;
#pragma unroll
    for (int k = 0; k < kRankItems; ++k) {
      const long long w = base + k;
      if (w < w1) {
        word_prefix[w] = run;
        unsigned bits = wv[k];
        while (bits) {
          const int b = __ffs(bits) - 1;
          bits &= bits - 1;
          emit(run, ((unsigned long long)w << 5) | (unsigned)b);
          run += 1;
        }
      }
    }
    run_base += ttot;
  }
}

template <typename Emit>
__global__ void __launch_bounds__(kRankThreads)
    rank_chunk_scan_kernel(const unsigned *__restrict__ bitmap, long long nwords, long long chunk_words, int nchunks,
                           int *__restrict__ sync, int *__restrict__ word_prefix, int *__restrict__ total, Emit emit) {
  rank_chunk_scan_body(bitmap, nwords, chunk_words, nchunks, sync, word_prefix, total, emit);
}

// several independent bitmaps in ONE launch: blockIdx.y selects the segment (each with its own ticket / status words)
constexpr int kRankMaxSegments = 6;
struct RankSegment {
  const unsigned *bitmap;
  long long nwords, chunk_words;
  int nchunks;
  int *sync, *word_prefix, *total;
};
struct RankSegments {
  int n;
  RankSegment seg[kRankMaxSegments];
};
inline void rank_plan_chunks(long long nwords, long long &chunk_words, int &nchunks) {
  long long nc = (nwords + kRankTile - 1) / kRankTile;
  if (nc > kRankMaxChunks) nc = kRankMaxChunks;
  if (nc < 1) nc = 1;
  chunk_words = (nwords + nc - 1) / nc;
  chunk_words = (chunk_words + kRankTile - 1) / kRankTile * kRankTile;
  nc = (nwords + chunk_words - 1) / chunk_words;
  nchunks = nc < 1 ? 1 : (int)nc;
}
struct RankEmitNothing {
  __device__ void operator()(int, unsigned long long) const {}
};
static __global__ void __launch_bounds__(kRankThreads) rank_chunk_scan_multi_kernel(const __grid_constant__ RankSegments S) {
  const RankSegment &g = S.seg[blockIdx.y];
  if ((int)blockIdx.x >= g.nchunks) return;      // whole CTA (tickets are taken by the first nchunks CTAs of the segment only)
  rank_chunk_scan_body(g.bitmap, g.nwords, g.chunk_words, g.nchunks, g.sync, g.word_prefix, g.total, RankEmitNothing{});
}

// `sync` must hold 1 + kRankMaxChunks zeroed ints
template <typename Emit>
int rank_build_fused(const unsigned *bitmap, long long nwords, int *word_prefix, int *sync, int *total_dev, Emit emit,
                     cudaStream_t st) {
  long long nchunks = (nwords + kRankTile - 1) / kRankTile;
  if (nchunks > kRankMaxChunks) nchunks = kRankMaxChunks;
  if (nchunks < 1) nchunks = 1;
  long long chunk_words = (nwords + nchunks - 1) / nchunks;
  chunk_words = (chunk_words + kRankTile - 1) / kRankTile * kRankTile;
  nchunks = (nwords + chunk_words - 1) / chunk_words;
  if (nchunks < 1) nchunks = 1;
  rank_chunk_scan_kernel<<<(int)nchunks, kRankThreads, 0, st>>>(bitmap, nwords, chunk_words, (int)nchunks, sync, word_prefix,
                                                              total_dev, emit);
  BEVF_CHECK_LAUNCH();
  return BEVF_OK;
}

}  // namespace bevf
