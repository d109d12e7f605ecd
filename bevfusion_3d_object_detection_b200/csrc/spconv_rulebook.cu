// spconv_rulebook.cu -- coordinate index + rulebook (indice-pair) generation for sparse 3-D convolution.
//
// Restates what spconv>=2.3's SpconvOps.get_indice_pairs_implicit_gemm produces (call site in the reference:
// projects/SparseConvolution/sparse_functional.py:118-137): pair_fwd[kv, n_out] (input row feeding output j
// under kernel tap k, -1 = none).  spconv builds it with a hash table (+ thrust sort for the output order);
// here the lookup structure is an occupancy bitmap over the (batch, X, Y, Z) grid with a popcount prefix
// (bitmap_rank.cuh):
//   * coordinate -> row lookup is one 4-byte bitmap read for a miss (most of the 27 taps miss) and two more
//     reads for a hit; no probing loops, no atomicCAS;
//   * the output sites of a strided convolution fall out of the same scan already sorted by linear index,
//     which is this framework's canonical output order (SURVEY 7 "Rulebook bit-exact");
//   * a level produced by a strided conv needs no row permutation (row == rank), only the raw voxelizer
//     output (first-appearance order) does.
#include <cuda_bf16.h>

#include "bitmap_rank.cuh"
#include "common.cuh"

namespace {

struct Grid {
  int b, x, y, z;
};

struct IndexView {
  const unsigned *bitmap;
  const int *word_prefix;
  const int *perm;  // rank -> row, or nullptr when rows are already in ascending linear order
  Grid g;
};

struct ConvGeom {
  int k[3], s[3], p[3], d[3];
  int sh[3];   // log2(s) when s is a power of two, else -1
};

__device__ __forceinline__ long long lin_cell(const Grid &g, int b, int x, int y, int z) {
  return (((long long)b * g.x + x) * g.y + y) * g.z + z;
}
// IT = unsigned (grids below 2^31 cells: the hot kernels then run on 32-bit integer math) or long long
template <typename IT>
__device__ __forceinline__ IT lin_cell_t(const Grid &g, int b, int x, int y, int z) {
  return (((IT)b * (IT)g.x + (IT)x) * (IT)g.y + (IT)y) * (IT)g.z + (IT)z;
}

template <typename IT>
__device__ __forceinline__ int index_lookup(const IndexView &ix, int b, int x, int y, int z) {
  if ((unsigned)x >= (unsigned)ix.g.x || (unsigned)y >= (unsigned)ix.g.y || (unsigned)z >= (unsigned)ix.g.z) return -1;
  const IT cell = lin_cell_t<IT>(ix.g, b, x, y, z);
  const unsigned bits = __ldg(ix.bitmap + (cell >> 5));
  const unsigned bit = (unsigned)(cell & 31);
  if (!((bits >> bit) & 1u)) return -1;
  const int rank = __ldg(ix.word_prefix + (cell >> 5)) + __popc(bits & ((1u << bit) - 1u));
  return ix.perm ? __ldg(ix.perm + rank) : rank;
}

// The k[2] taps of one (kx, ky) column are z-neighbours, i.e. (z is the fastest axis) cells of ONE bitmap word most of
// the time: the word and its prefix are loaded once per column, not once per tap.  out[kz] = row or -1.
template <typename IT>
__device__ __forceinline__ void index_lookup_zcol(const IndexView &ix, int b, int x, int y, int z0, int dz, int kz_n,
                                                  int *__restrict__ dst, size_t dst_stride) {
  const bool xy_ok = (unsigned)x < (unsigned)ix.g.x && (unsigned)y < (unsigned)ix.g.y;
  const IT base = xy_ok ? lin_cell_t<IT>(ix.g, b, x, y, 0) : (IT)0;
  IT cur_w = ~(IT)0;
  unsigned bits = 0u;
  int pref = -1;
  for (int kz = 0; kz < kz_n; ++kz) {
    const int z = z0 + kz * dz;
    int row = -1;
    if (xy_ok && (unsigned)z < (unsigned)ix.g.z) {
      const IT cell = base + (IT)z;
      const IT w = cell >> 5;
      if (w != cur_w) {
        cur_w = w;
        bits = __ldg(ix.bitmap + w);
        pref = -1;
      }
      const unsigned bit = (unsigned)(cell & 31);
      if ((bits >> bit) & 1u) {
        if (pref < 0) pref = __ldg(ix.word_prefix + w);
        const int rank = pref + __popc(bits & ((1u << bit) - 1u));
        row = ix.perm ? __ldg(ix.perm + rank) : rank;
      }
    }
    dst[(size_t)kz * dst_stride] = row;
  }
}

constexpr int kRbBlocks = bevf::kNumSMs * 16;  // grid-stride kernels: the grid never grows with the buffer capacity

// every kernel below takes the row count either from the host (n) or, when n_dev != NULL, from device memory
// (min(*n_dev, n): n is then the capacity the grid was sized for) so a whole frame can run without a host sync
__global__ void mark_sites_kernel(const int *__restrict__ indices, int n, const int *__restrict__ n_dev, Grid g,
                                  unsigned *__restrict__ bitmap, int *__restrict__ error_flag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) n = min(n, *n_dev);
  if (i >= n) return;
  int4 c = __ldg(reinterpret_cast<const int4 *>(indices) + i);  // (b, x, y, z)
  if ((unsigned)c.x >= (unsigned)g.b || (unsigned)c.y >= (unsigned)g.x || (unsigned)c.z >= (unsigned)g.y ||
      (unsigned)c.w >= (unsigned)g.z) {
    *error_flag = 1;
    return;
  }
  long long cell = lin_cell(g, c.x, c.y, c.z, c.w);
  unsigned old = atomicOr(bitmap + (cell >> 5), 1u << (unsigned)(cell & 31));
  if ((old >> (unsigned)(cell & 31)) & 1u) *error_flag = 2;  // duplicate coordinate
}

struct EmitNothing {
  __device__ void operator()(int, unsigned long long) const {}
};

__global__ void fill_perm_kernel(const int *__restrict__ indices, int n, const int *__restrict__ n_dev, Grid g,
                                 const unsigned *__restrict__ bitmap, const int *__restrict__ word_prefix,
                                 int *__restrict__ perm) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) n = min(n, *n_dev);
  if (i >= n) return;
  int4 c = __ldg(reinterpret_cast<const int4 *>(indices) + i);
  if ((unsigned)c.x >= (unsigned)g.b || (unsigned)c.y >= (unsigned)g.x || (unsigned)c.z >= (unsigned)g.y ||
      (unsigned)c.w >= (unsigned)g.z)
    return;
  long long cell = lin_cell(g, c.x, c.y, c.z, c.w);
  unsigned bits = bitmap[cell >> 5];
  int rank = word_prefix[cell >> 5] + __popc(bits & ((1u << (unsigned)(cell & 31)) - 1u));
  perm[rank] = i;
}

// SubM: out sites == in sites; kernel centred; one work item per (site, tap-x/y column), the k[2] z-taps of a
// column share the bitmap word most of the time.  Grid-stride over n * kx * ky items (n possibly from the device).
template <typename IT>
__global__ void __launch_bounds__(256)
    subm_rulebook_kernel(const int *__restrict__ indices, int n, const int *__restrict__ n_dev, IndexView ix,
                         ConvGeom cg, int ld, int *__restrict__ pair_fwd) {
  if (n_dev) n = min(n, *n_dev);
  const unsigned kxy = (unsigned)(cg.k[0] * cg.k[1]);
  const unsigned total = (unsigned)n * kxy;
  for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const unsigned col = t / (unsigned)n;
    const int j = (int)(t - col * (unsigned)n);  // consecutive threads -> consecutive sites: coalesced stores
    const int kx = (int)(col / (unsigned)cg.k[1]), ky = (int)(col % (unsigned)cg.k[1]);
    const int4 c = __ldg(reinterpret_cast<const int4 *>(indices) + j);
    const int x = c.y + (kx - cg.k[0] / 2) * cg.d[0];
    const int y = c.z + (ky - cg.k[1] / 2) * cg.d[1];
    index_lookup_zcol<IT>(ix, c.x, x, y, c.w - (cg.k[2] / 2) * cg.d[2], cg.d[2], cg.k[2],
                          pair_fwd + (size_t)((int)col * cg.k[2]) * ld + j, (size_t)ld);
  }
}

// The hot geometry (every 3 x 3 x 3, dilation-1 layer of the encoder), SubM (S = 1, centred) and strided alike: blockIdx.y is
// kx (no division to split the work item), a thread takes one site and its three ky columns; the z triple of a column
// lies in one or two bitmap words that are loaded once; all address arithmetic is 32-bit adds from the site's own cell.
// ix.perm must be null (rows in ascending cell order).  x0/y0/z0 = input coordinate of tap (0, 0, 0) for the site.
template <bool SUBM>
__global__ void __launch_bounds__(256)
    rulebook_k3_kernel(const int *__restrict__ sites, int n_host, const int *__restrict__ n_dev, int n_cap, IndexView ix,
                       ConvGeom cg, int ld, int *__restrict__ pair_fwd) {
  int n = n_host;
  if (n_dev) n = min(*n_dev, n_cap);
  const int kx = blockIdx.y;
  const unsigned gz = (unsigned)ix.g.z, gyz = (unsigned)ix.g.y * gz;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const int4 c = __ldg(reinterpret_cast<const int4 *>(sites) + j);
    const int x = SUBM ? c.y + kx - 1 : c.y * cg.s[0] - cg.p[0] + kx;
    const int y0 = SUBM ? c.z - 1 : c.z * cg.s[1] - cg.p[1];
    const int z0 = SUBM ? c.w - 1 : c.w * cg.s[2] - cg.p[2];
    const bool x_ok = (unsigned)x < (unsigned)ix.g.x;
    const unsigned base_x = ((unsigned)c.x * (unsigned)ix.g.x + (unsigned)x) * gyz;
    int *dst = pair_fwd + (size_t)(kx * 9) * ld + j;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int y = y0 + ky;
      int r0 = -1, r1 = -1, r2 = -1;
      if (x_ok && (unsigned)y < (unsigned)ix.g.y) {
        const unsigned cell0 = base_x + (unsigned)y * gz + (unsigned)z0;   // may wrap when z0 < 0: only used if z valid
        const bool v0 = (unsigned)z0 < gz, v1 = (unsigned)(z0 + 1) < gz, v2 = (unsigned)(z0 + 2) < gz;
        const unsigned ca = v0 ? cell0 : (v1 ? cell0 + 1u : cell0 + 2u);   // first valid cell of the triple
        const unsigned cb = v2 ? cell0 + 2u : (v1 ? cell0 + 1u : cell0);   // last valid cell
        if (v0 | v1 | v2) {
          const unsigned wa = ca >> 5, wb = cb >> 5;
          const unsigned ba = __ldg(ix.bitmap + wa);
          const unsigned bb = wb != wa ? __ldg(ix.bitmap + wb) : ba;
          int pa = -1, pb = -1;
          auto look = [&](unsigned cell) {
            const unsigned w = cell >> 5, bit = cell & 31u;
            const unsigned bits = w == wa ? ba : bb;
            if (!((bits >> bit) & 1u)) return -1;
            int &pref = w == wa ? pa : pb;
            if (pref < 0) pref = __ldg(ix.word_prefix + w);
            return pref + __popc(bits & ((1u << bit) - 1u));
          };
          if (v0) r0 = look(cell0);
          if (v1) r1 = look(cell0 + 1u);
          if (v2) r2 = look(cell0 + 2u);
        }
      }
      dst[(size_t)(ky * 3 + 0) * ld] = r0;
      dst[(size_t)(ky * 3 + 1) * ld] = r1;
      dst[(size_t)(ky * 3 + 2) * ld] = r2;
    }
  }
}

// strided conv, pass 1: every input marks the output sites it reaches.  One thread per input site.  Per AXIS the taps
// that land on an output ((c + p - k d) a non-negative multiple of s, inside the grid) are found first -- <= K candidates
// per axis, with shifts instead of divisions for power-of-two strides (the 27 per-tap divisions were what this kernel
// spent its time on) -- then the candidate products are walked; the z outputs of one (ox, oy) are neighbouring cells of
// one bitmap word (two when the run crosses a word) and go out as ONE atomicOr, skipped when the bits are already set.
template <typename IT, int KMAX>
__global__ void __launch_bounds__(256)
    strided_mark_kernel(const int *__restrict__ indices, int n, const int *__restrict__ n_dev, ConvGeom cg, Grid og,
                        unsigned *__restrict__ out_bitmap) {
  if (n_dev) n = min(n, *n_dev);
  const int lim[3] = {og.x, og.y, og.z};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int4 c = __ldg(reinterpret_cast<const int4 *>(indices) + i);
    const int cc[3] = {c.y, c.z, c.w};
    int cand[3][KMAX];   // output coordinate per tap, -1 = the tap lands on no output (constant indices: registers)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        int q = -1;
        if (k < cg.k[a]) {
          const int t = cc[a] + cg.p[a] - k * cg.d[a];
          if (t >= 0) {
            if (cg.sh[a] >= 0) {
              if ((t & (cg.s[a] - 1)) == 0) q = t >> cg.sh[a];
            } else if (t % cg.s[a] == 0) {
              q = t / cg.s[a];
            }
            if (q >= lim[a]) q = -1;
          }
        }
        cand[a][k] = q;
      }
    }
#pragma unroll
    for (int kx = 0; kx < KMAX; ++kx) {
      if (cand[0][kx] < 0) continue;
#pragma unroll
      for (int ky = 0; ky < KMAX; ++ky) {
        if (cand[1][ky] < 0) continue;
        const IT base = lin_cell_t<IT>(og, c.x, cand[0][kx], cand[1][ky], 0);
        IT cur_w = ~(IT)0;
        unsigned mask = 0u;
#pragma unroll
        for (int kz = 0; kz < KMAX; ++kz) {
          if (cand[2][kz] < 0) continue;
          const IT cell = base + (IT)cand[2][kz];
          const IT w = cell >> 5;
          if (w != cur_w) {
            if (mask && (out_bitmap[cur_w] & mask) != mask) atomicOr(out_bitmap + cur_w, mask);
            cur_w = w;
            mask = 0u;
          }
          mask |= 1u << (unsigned)(cell & 31);
        }
        if (mask && (out_bitmap[cur_w] & mask) != mask) atomicOr(out_bitmap + cur_w, mask);
      }
    }
  }
}

// strided conv, pass 1b: output sites in ascending cell order from the scanned bitmap (one work item per word)
template <typename IT>
__global__ void __launch_bounds__(256)
    emit_sites_kernel(const unsigned *__restrict__ bitmap, const int *__restrict__ word_prefix, long long nwords,
                      Grid og, int *__restrict__ out_indices, int cap) {
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < nwords;
       w += (long long)gridDim.x * blockDim.x) {
    unsigned bits = __ldg(bitmap + w);
    if (!bits) continue;
    int rank = __ldg(word_prefix + w);
    // coordinates of the word's first cell (three divisions per word, not per site), then carried forward bit by bit
    IT key = (IT)w << 5;
    int z0 = (int)(key % (IT)og.z); key /= (IT)og.z;
    int y0 = (int)(key % (IT)og.y); key /= (IT)og.y;
    int x0 = (int)(key % (IT)og.x); key /= (IT)og.x;
    int b0 = (int)key;
    int last = 0;
    while (bits) {
      const int b = __ffs(bits) - 1;
      bits &= bits - 1;
      z0 += b - last;
      last = b;
      while (z0 >= og.z) {
        z0 -= og.z;
        if (++y0 >= og.y) {
          y0 = 0;
          if (++x0 >= og.x) { x0 = 0; ++b0; }
        }
      }
      if (rank < cap) reinterpret_cast<int4 *>(out_indices)[rank] = make_int4(b0, x0, y0, z0);
      rank += 1;
    }
  }
}

// ---- all strided levels of a chain from the level-0 coordinates ---------------------------------------------------------
// The sites of a strided level are the cells some site of the level below reaches; for ONE site the reachable outputs are,
// per axis, the interval ceil((c + p - (k-1)) / s) .. floor((c + p) / s) clipped to the grid, for a full interval of
// inputs [i0, i1] the interval ceil((i0 + p - (k-1)) / s) .. floor((i1 + p) / s) -- so the box a level-0 site reaches on
// level l follows from its box on level l-1 (every cell of which is active, being reached by that very site), and level l's
// site set is the union of the boxes of all level-0 sites.  One pass over the level-0 coordinates therefore fills the
// bitmaps of ALL levels; their rank scans and site lists are then independent and go out as one multi-segment launch each
// (per level before: mark + scan + emit + three memory operations, each waiting for the level below).  Dilation 1.
constexpr int kMaxChain = bevf::kRankMaxSegments;
struct ChainLevel {
  int k[3], s[3], p[3];
  Grid og;
  unsigned *bitmap;
  const int *word_prefix;
  const int *total;
  long long nwords;
  int *out_indices;
  int cap;
  int *n_out_dev;
};
struct ChainArgs {
  int nlev;
  Grid g0;
  ChainLevel lv[kMaxChain];
};

__device__ __forceinline__ int ceil_div_i(int a, int s) { return a >= 0 ? (a + s - 1) / s : -((-a) / s); }
__device__ __forceinline__ int floor_div_i(int a, int s) { return a >= 0 ? a / s : -((-a + s - 1) / s); }

__global__ void __launch_bounds__(256)
    mark_levels_kernel(const int *__restrict__ coords0, int n, const int *__restrict__ n_dev, const __grid_constant__ ChainArgs A) {
  if (n_dev) n = min(n, *n_dev);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int4 c = __ldg(reinterpret_cast<const int4 *>(coords0) + i);   // (b, x, y, z)
    if ((unsigned)c.x >= (unsigned)A.g0.b || (unsigned)c.y >= (unsigned)A.g0.x || (unsigned)c.z >= (unsigned)A.g0.y ||
        (unsigned)c.w >= (unsigned)A.g0.z)
      continue;                                                          // reported by the level-0 index build
    int lo[3] = {c.y, c.z, c.w}, hi[3] = {c.y, c.z, c.w};
    for (int l = 0; l < A.nlev; ++l) {
      const ChainLevel &L = A.lv[l];
      const int lim[3] = {L.og.x, L.og.y, L.og.z};
      bool empty = false;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const int nlo = max(ceil_div_i(lo[a] + L.p[a] - (L.k[a] - 1), L.s[a]), 0);
        const int nhi = min(floor_div_i(hi[a] + L.p[a], L.s[a]), lim[a] - 1);
        lo[a] = nlo; hi[a] = nhi;
        empty |= nlo > nhi;
      }
      if (empty) break;
      for (int x = lo[0]; x <= hi[0]; ++x) {
        for (int y = lo[1]; y <= hi[1]; ++y) {
          const unsigned long long base = (((unsigned long long)c.x * L.og.x + x) * L.og.y + y) * L.og.z;
          unsigned long long cur_w = ~0ull;
          unsigned mask = 0u;
          for (int z = lo[2]; z <= hi[2]; ++z) {                         // the z run of one (x, y): one or two bitmap words
            const unsigned long long cell = base + (unsigned)z, w = cell >> 5;
            if (w != cur_w) {
              if (mask && (L.bitmap[cur_w] & mask) != mask) atomicOr(L.bitmap + cur_w, mask);
              cur_w = w;
              mask = 0u;
            }
            mask |= 1u << (unsigned)(cell & 31);
          }
          if (mask && (L.bitmap[cur_w] & mask) != mask) atomicOr(L.bitmap + cur_w, mask);
        }
      }
    }
  }
}

// site lists of all levels of a chain: blockIdx.y = level (same per-word walk as emit_sites_kernel)
__global__ void __launch_bounds__(256) emit_sites_multi_kernel(const __grid_constant__ ChainArgs A) {
  const ChainLevel &L = A.lv[blockIdx.y];
  const Grid og = L.og;
  if (blockIdx.x == 0 && threadIdx.x == 0 && L.n_out_dev) *L.n_out_dev = *L.total;
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < L.nwords; w += (long long)gridDim.x * blockDim.x) {
    unsigned bits = __ldg(L.bitmap + w);
    if (!bits) continue;
    int rank = __ldg(L.word_prefix + w);
    unsigned long long key = (unsigned long long)w << 5;
    int z0 = (int)(key % (unsigned long long)og.z); key /= (unsigned long long)og.z;
    int y0 = (int)(key % (unsigned long long)og.y); key /= (unsigned long long)og.y;
    int x0 = (int)(key % (unsigned long long)og.x); key /= (unsigned long long)og.x;
    int b0 = (int)key;
    int last = 0;
    while (bits) {
      const int b = __ffs(bits) - 1;
      bits &= bits - 1;
      z0 += b - last;
      last = b;
      while (z0 >= og.z) {
        z0 -= og.z;
        if (++y0 >= og.y) {
          y0 = 0;
          if (++x0 >= og.x) { x0 = 0; ++b0; }
        }
      }
      if (rank < L.cap) reinterpret_cast<int4 *>(L.out_indices)[rank] = make_int4(b0, x0, y0, z0);
      rank += 1;
    }
  }
}

// strided conv, pass 2: for every output site and tap, look the input up
template <typename IT>
__global__ void __launch_bounds__(256)
    strided_rulebook_kernel(const int *__restrict__ out_indices, const int *__restrict__ n_out_dev, int n_out_host,
                            IndexView ix, ConvGeom cg, int ld, int *__restrict__ pair_fwd) {
  const int n_out = n_out_dev ? min(*n_out_dev, ld) : n_out_host;
  const unsigned kxy = (unsigned)(cg.k[0] * cg.k[1]);
  const unsigned total = (unsigned)n_out * kxy;
  for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const unsigned col = t / (unsigned)n_out;
    const int j = (int)(t - col * (unsigned)n_out);
    const int kx = (int)(col / (unsigned)cg.k[1]), ky = (int)(col % (unsigned)cg.k[1]);
    const int4 o = __ldg(reinterpret_cast<const int4 *>(out_indices) + j);
    const int x = o.y * cg.s[0] - cg.p[0] + kx * cg.d[0];
    const int y = o.z * cg.s[1] - cg.p[1] + ky * cg.d[1];
    index_lookup_zcol<IT>(ix, o.x, x, y, o.w * cg.s[2] - cg.p[2], cg.d[2], cg.k[2],
                          pair_fwd + (size_t)((int)col * cg.k[2]) * ld + j, (size_t)ld);
  }
}

// SparseConvTensor.dense(): [B, C, X, Y, Z]; one thread per (site, channel), channel fastest in the read
__global__ void __launch_bounds__(256)
    to_dense_kernel(const float *__restrict__ feats, const int *__restrict__ indices, int n,
                    const int *__restrict__ n_dev, int c, Grid g, float *__restrict__ dense) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) n = min(n, *n_dev);
  if (t >= (long long)n * c) return;
  const int i = (int)(t / c), ch = (int)(t % c);
  const int4 q = __ldg(reinterpret_cast<const int4 *>(indices) + i);
  const size_t vol = (size_t)g.x * g.y * g.z;
  dense[((size_t)q.x * c + ch) * vol + ((size_t)q.y * g.y + q.z) * g.z + q.w] = feats[t];
}

// backward of dense() / of the fused BEV tail: rows gathered back from the dense gradient at the active sites
__global__ void __launch_bounds__(256)
    from_dense_kernel(const float *__restrict__ dense, const int *__restrict__ indices, int n,
                      const int *__restrict__ n_dev, int c, Grid g, int bev_layout, float *__restrict__ feats) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) n = min(n, *n_dev);
  if (t >= (long long)n * c) return;
  const int i = (int)(t / c), ch = (int)(t % c);
  const int4 q = __ldg(reinterpret_cast<const int4 *>(indices) + i);
  size_t src;
  if (bev_layout) src = (((size_t)q.x * c * g.z + (size_t)ch * g.z + q.w) * g.x + q.y) * g.y + q.z;
  else src = ((size_t)q.x * c + ch) * ((size_t)g.x * g.y * g.z) + ((size_t)q.y * g.y + q.z) * g.z + q.w;
  feats[t] = __ldg(dense + src);
}

// BEVFusionSparseEncoder tail (sparse_encoder.py:147-151): dense() [N,C,X,Y,Z] -> permute(0,1,4,2,3) ->
// view(N, C*Z, X, Y), fused: bev[b, ch*Z + z, x, y].  A CTA takes 32 consecutive sites: their feature rows are read
// coalesced into shared memory, then each warp writes one channel at a time for the 32 sites -- with sites in
// ascending (x, y, z) order the 32 destinations of a channel are a few runs of consecutive y.
__global__ void __launch_bounds__(256)
    to_bev_kernel(const float *__restrict__ feats, const int *__restrict__ indices, int n,
                  const int *__restrict__ n_dev, int c, Grid g, float *__restrict__ bev) {
  extern __shared__ float tile[];  // [32][c + 1]
  __shared__ int4 site[32];
  if (n_dev) n = min(n, *n_dev);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i0 = blockIdx.x * 32; i0 < n; i0 += gridDim.x * 32) {
    const int cnt = min(32, n - i0);
    if (threadIdx.x < cnt) site[threadIdx.x] = __ldg(reinterpret_cast<const int4 *>(indices) + i0 + threadIdx.x);
    for (int e = threadIdx.x; e < cnt * c; e += blockDim.x) tile[(e / c) * (c + 1) + e % c] = feats[(size_t)i0 * c + e];
    __syncthreads();
    if (lane < cnt) {
      const int4 q = site[lane];
      const size_t base = ((size_t)q.x * c * g.z + q.w) * g.x * g.y + (size_t)q.y * g.y + q.z;
      const size_t cstride = (size_t)g.z * g.x * g.y;
      for (int ch = warp; ch < c; ch += 8) bev[base + ch * cstride] = tile[lane * (c + 1) + ch];
    }
    __syncthreads();
  }
}

// The same BEV tail driven from the OUTPUT side (the level's coordinate index gives cell -> row): a CTA owns one x line of
// one sample and 32 channels; it looks up the row of every (y, z) cell of the line (bitmap + popcount prefix: rows are in
// ascending cell order, so the line's rows are one contiguous range), stages that range's 128-byte channel chunks in
// shared memory with coalesced loads, and writes every (channel, z) plane line -- ny contiguous floats -- complete, zeros
// included.  No memset of the 33 MB map, no 4-byte scattered stores (to_bev_kernel: 27 us + the memset; this: one pass).
constexpr int kBevCh = 32;
__global__ void __launch_bounds__(256)
    to_bev_indexed_kernel(const float *__restrict__ feats, const unsigned *__restrict__ bitmap,
                          const int *__restrict__ word_prefix, const int *__restrict__ n_dev, int n_cap, int c, Grid g,
                          int tile_floats, float *__restrict__ bev) {
  extern __shared__ int sm_bev[];
  const int cells = g.y * g.z;
  int *rowidx = sm_bev;                                            // [y * Z + z] -> row or -1
  float *tile = reinterpret_cast<float *>(sm_bev + cells);         // [rows of the line][channels of the pass + 1]
  __shared__ int s_r0, s_r1;
  const int x = blockIdx.x % g.x, chunk = blockIdx.x / g.x, b = blockIdx.y;
  const int ch0 = chunk * kBevCh, nch = min(kBevCh, c - ch0);
  const int n = n_dev ? min(*n_dev, n_cap) : n_cap;
  const unsigned long long cell0 = ((unsigned long long)b * g.x + x) * (unsigned long long)cells;
  if (threadIdx.x == 0) { s_r0 = 0x7fffffff; s_r1 = -1; }
  __syncthreads();
  for (int t = threadIdx.x; t < cells; t += blockDim.x) {
    const unsigned long long cell = cell0 + (unsigned)t;
    const unsigned bits = __ldg(bitmap + (cell >> 5)), bit = (unsigned)(cell & 31);
    int row = -1;
    if ((bits >> bit) & 1u) {
      row = __ldg(word_prefix + (cell >> 5)) + __popc(bits & ((1u << bit) - 1u));
      if (row >= n) row = -1;                                      // beyond the level's capacity (flagged elsewhere)
    }
    rowidx[t] = row;
    if (row >= 0) { atomicMin(&s_r0, row); atomicMax(&s_r1, row); }
  }
  __syncthreads();
  const int r0 = s_r0, nrows = s_r1 - r0 + 1;                      // <= cells: the line's rows are consecutive
  // channels per pass: all of the chunk when the line's rows fit the tile (sized for ~90 % of a full line so that five
  // CTAs share an SM and the grid is one wave), fewer -- a multiple of 4 -- for the rare fuller line
  int gch = nch;
  if (nrows > 0 && nrows * (nch + 1) > tile_floats) gch = max(4, ((tile_floats / nrows - 1) / 4) * 4);
  const size_t plane = (size_t)g.x * g.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c0 = 0; c0 < nch; c0 += gch) {
    const int cn = min(gch, nch - c0), pitch = cn + 1;
    if (nrows > 0) {
      const int vec = cn / 4;                                      // 16-byte pieces per row chunk (c % 4 == 0)
      for (int e = threadIdx.x; e < nrows * vec; e += blockDim.x) {
        const int r = e / vec, q = e - r * vec;
        const float4 v = __ldg(reinterpret_cast<const float4 *>(feats + (size_t)(r0 + r) * c + ch0 + c0) + q);
        float *d = tile + r * pitch + 4 * q;
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
      }
    }
    __syncthreads();
    for (int p = warp; p < cn * g.z; p += blockDim.x >> 5) {       // plane line (channel, z): ny contiguous floats
      const int chl = p / g.z, z = p - chl * g.z;
      float *o = bev + (((size_t)b * c + ch0 + c0 + chl) * g.z + z) * plane + (size_t)x * g.y;
      for (int y = lane; y < g.y; y += 32) {
        const int row = rowidx[y * g.z + z];
        o[y] = row >= 0 ? tile[(row - r0) * pitch + chl] : 0.f;
      }
    }
    __syncthreads();
  }
}

// rows re-ordered by perm (rank -> row): out_indices[r] = indices[perm[r]], features copied as fp32 and/or as bf16
// zero-padded to cin_pad columns (the tensor-core operand of the first conv)
__global__ void __launch_bounds__(256)
    permute_rows_kernel(const float *__restrict__ feats, const int *__restrict__ indices, const int *__restrict__ perm,
                        int n, const int *__restrict__ n_dev, int c, int cin_pad, int width, float *__restrict__ out_f32,
                        unsigned short *__restrict__ out_bf16, int *__restrict__ out_indices) {
  if (n_dev) n = min(n, *n_dev);
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * width) return;
  const int r = (int)(t / width), col = (int)(t % width);
  const int src = __ldg(perm + r);
  if (col < 4) out_indices[(size_t)r * 4 + col] = __ldg(indices + (size_t)src * 4 + col);
  const float v = col < c ? __ldg(feats + (size_t)src * c + col) : 0.f;
  if (out_f32 && col < c) out_f32[(size_t)r * c + col] = v;
  if (out_bf16 && col < cin_pad) {
    __nv_bfloat16 b = __float2bfloat16_rn(v);
    out_bf16[(size_t)r * cin_pad + col] = *reinterpret_cast<unsigned short *>(&b);
  }
}

constexpr int kScalarInts = 64 + 1 + bevf::kRankMaxChunks + 7;

struct IndexMem {
  unsigned *bitmap;
  int *word_prefix;
  int *block_counts;
  int *scalars;  // [0] total, [1] error flag
  long long nwords;
};

int grid_words(int batch, const int *shape, long long &nwords) {
  BEVF_CHECK_ARG(batch > 0 && shape[0] > 0 && shape[1] > 0 && shape[2] > 0, "bad grid %d x (%d,%d,%d)", batch,
                 shape[0], shape[1], shape[2]);
  long long cells = (long long)batch * shape[0] * shape[1] * shape[2];
  if (cells >= (1ll << 36)) {
    bevf::set_error("sparse grid with %lld cells is not supported (limit 2^36)", cells);
    return BEVF_ERR_UNSUPPORTED;
  }
  nwords = (cells + 31) / 32;
  return BEVF_OK;
}

size_t carve_index(IndexMem &m, void *mem, size_t bytes, long long nwords) {
  bevf::Workspace a(mem, bytes);
  m.nwords = nwords;
  m.scalars = a.take<int>(kScalarInts);   // [0] total, [1] error flag, [64 ..] ticket + chunk flags of the rank scan
  m.bitmap = a.take<unsigned>((size_t)nwords);
  m.word_prefix = a.take<int>((size_t)nwords);
  m.block_counts = a.take<int>((size_t)bevf::rank_num_blocks(nwords));
  return a.off;
}

void fill_geom(ConvGeom &cg, const int *k, const int *s, const int *p, const int *d) {
  for (int j = 0; j < 3; ++j) {
    cg.k[j] = k[j];
    cg.s[j] = s ? s[j] : 1;
    cg.p[j] = p ? p[j] : 0;
    cg.d[j] = d ? d[j] : 1;
    cg.sh[j] = -1;
    for (int b = 0; b < 16; ++b)
      if (cg.s[j] == (1 << b)) cg.sh[j] = b;
  }
}

int check_geom(const int *k, const int *s, const int *p, const int *d) {
  for (int j = 0; j < 3; ++j) {
    BEVF_CHECK_ARG(k[j] >= 1 && k[j] <= 7, "kernel size %d out of range 1..7", k[j]);
    BEVF_CHECK_ARG(!s || s[j] >= 1, "stride must be >= 1");
    BEVF_CHECK_ARG(!p || p[j] >= 0, "padding must be >= 0");
    BEVF_CHECK_ARG(!d || d[j] >= 1, "dilation must be >= 1");
  }
  return BEVF_OK;
}

}  // namespace

BEVF_API int bevf_spconv_out_shape(const int *shape, const int *k, const int *s, const int *p, const int *d,
                                   int *out_shape) {
  for (int j = 0; j < 3; ++j) out_shape[j] = (shape[j] + 2 * p[j] - d[j] * (k[j] - 1) - 1) / s[j] + 1;
  return BEVF_OK;
}

BEVF_API size_t bevf_spconv_index_bytes(int batch, const int *shape) {
  long long nwords;
  if (grid_words(batch, shape, nwords)) return 0;
  IndexMem m;
  return carve_index(m, nullptr, 0, nwords) + 256;
}

BEVF_API int bevf_spconv_index_build(const int *indices, int n, const int *n_dev, int batch, const int *shape,
                                     void *index_mem, size_t index_bytes, int *perm, void *stream) {
  long long nwords;
  int rc = grid_words(batch, shape, nwords);
  if (rc) return rc;
  IndexMem m;
  size_t need = carve_index(m, index_mem, index_bytes, nwords);
  if (!index_mem || need > index_bytes) {
    bevf::set_error("spconv index memory too small: need %zu bytes, got %zu", need, index_bytes);
    return BEVF_ERR_WORKSPACE;
  }
  BEVF_CHECK_ARG(n >= 0 && (n == 0 || indices), "bad indices");
  BEVF_CHECK_ARG((reinterpret_cast<uintptr_t>(indices) & 15u) == 0, "indices must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  Grid g{batch, shape[0], shape[1], shape[2]};
  BEVF_CHECK_CUDA(cudaMemsetAsync(m.scalars, 0, kScalarInts * sizeof(int), st));
  BEVF_CHECK_CUDA(cudaMemsetAsync(m.bitmap, 0, (size_t)nwords * sizeof(unsigned), st));
  if (n > 0) {
    mark_sites_kernel<<<bevf::ceil_div(n, 256), 256, 0, st>>>(indices, n, n_dev, g, m.bitmap, m.scalars + 1);
    BEVF_CHECK_LAUNCH();
  }
  rc = bevf::rank_build_fused(m.bitmap, nwords, m.word_prefix, m.scalars + 64, m.scalars, EmitNothing{}, st);
  if (rc) return rc;
  if (perm && n > 0) {
    fill_perm_kernel<<<bevf::ceil_div(n, 256), 256, 0, st>>>(indices, n, n_dev, g, m.bitmap, m.word_prefix, perm);
    BEVF_CHECK_LAUNCH();
  }
  return BEVF_OK;
}

// error flag of the last index build: 0 ok, 1 coordinate out of the grid, 2 duplicate coordinate (device int)
BEVF_API const int *bevf_spconv_index_error_flag(void *index_mem, size_t index_bytes, int batch, const int *shape) {
  long long nwords;
  if (grid_words(batch, shape, nwords)) return nullptr;
  IndexMem m;
  carve_index(m, index_mem, index_bytes, nwords);
  return m.scalars + 1;
}

BEVF_API int bevf_spconv_subm_rulebook(const int *indices, int n, const int *n_dev, int batch, const int *shape,
                                       const int *ksize,
                                       const int *dilation, const void *index_mem, size_t index_bytes,
                                       const int *perm, int *pair_fwd, int ld, void *stream) {
  int rc = check_geom(ksize, nullptr, nullptr, dilation);
  if (rc) return rc;
  long long nwords;
  rc = grid_words(batch, shape, nwords);
  if (rc) return rc;
  BEVF_CHECK_ARG(ld >= n, "pair_fwd leading dimension %d < n %d", ld, n);
  if (n == 0) return BEVF_OK;
  IndexMem m;
  carve_index(m, const_cast<void *>(index_mem), index_bytes, nwords);
  IndexView ix{m.bitmap, m.word_prefix, perm, Grid{batch, shape[0], shape[1], shape[2]}};
  ConvGeom cg;
  fill_geom(cg, ksize, nullptr, nullptr, dilation);
  const long long threads = (long long)n * ksize[0] * ksize[1];
  BEVF_CHECK_ARG(threads < (1ll << 32), "rulebook with %lld (site, tap column) items is not supported", threads);
  const bool k3 = ksize[0] == 3 && ksize[1] == 3 && ksize[2] == 3 && (!dilation || (dilation[0] == 1 && dilation[1] == 1 &&
                                                                                  dilation[2] == 1));
  if (k3 && !perm && nwords < (1ll << 26)) {   // the encoder's SubM layers: specialised kernel, grid.y = kx
    const int bx = bevf::ceil_div(n, 256) < kRbBlocks / 3 ? bevf::ceil_div(n, 256) : kRbBlocks / 3;
    rulebook_k3_kernel<true><<<dim3(bx, 3), 256, 0, (cudaStream_t)stream>>>(indices, n, n_dev, n, ix, cg, ld, pair_fwd);
    BEVF_CHECK_LAUNCH();
    return BEVF_OK;
  }
  const int blocks = (int)(bevf::ceil_div(threads, 256) < kRbBlocks ? bevf::ceil_div(threads, 256) : kRbBlocks);
  if (nwords < (1ll << 26))
    subm_rulebook_kernel<unsigned><<<blocks, 256, 0, (cudaStream_t)stream>>>(indices, n, n_dev, ix, cg, ld, pair_fwd);
  else
    subm_rulebook_kernel<long long><<<blocks, 256, 0, (cudaStream_t)stream>>>(indices, n, n_dev, ix, cg, ld, pair_fwd);
  BEVF_CHECK_LAUNCH();
  return BEVF_OK;
}

BEVF_API int bevf_spconv_strided_sites(const int *in_indices, int n_in, const int *n_in_dev, int batch,
                                       const int *in_shape,
                                       const int *ksize, const int *stride, const int *padding, const int *dilation,
                                       void *out_index_mem, size_t out_index_bytes, int *out_indices, int cap,
                                       int *n_out_dev, void *stream) {
  int rc = check_geom(ksize, stride, padding, dilation);
  if (rc) return rc;
  int os[3];
  bevf_spconv_out_shape(in_shape, ksize, stride, padding, dilation, os);
  long long nwords;
  rc = grid_words(batch, os, nwords);
  if (rc) return rc;
  IndexMem m;
  size_t need = carve_index(m, out_index_mem, out_index_bytes, nwords);
  if (!out_index_mem || need > out_index_bytes) {
    bevf::set_error("spconv output index memory too small: need %zu bytes, got %zu", need, out_index_bytes);
    return BEVF_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  Grid og{batch, os[0], os[1], os[2]};
  ConvGeom cg;
  fill_geom(cg, ksize, stride, padding, dilation);
  BEVF_CHECK_CUDA(cudaMemsetAsync(m.scalars, 0, kScalarInts * sizeof(int), st));
  BEVF_CHECK_CUDA(cudaMemsetAsync(m.bitmap, 0, (size_t)nwords * sizeof(unsigned), st));
  if (n_in > 0) {
    const long long items = (long long)n_in;
    const int blocks = (int)(bevf::ceil_div(items, 256) < kRbBlocks ? bevf::ceil_div(items, 256) : kRbBlocks);
    const bool small_k = ksize[0] <= 3 && ksize[1] <= 3 && ksize[2] <= 3;
    if (nwords < (1ll << 26)) {
      if (small_k) strided_mark_kernel<unsigned, 3><<<blocks, 256, 0, st>>>(in_indices, n_in, n_in_dev, cg, og, m.bitmap);
      else strided_mark_kernel<unsigned, 7><<<blocks, 256, 0, st>>>(in_indices, n_in, n_in_dev, cg, og, m.bitmap);
    } else {
      if (small_k) strided_mark_kernel<long long, 3><<<blocks, 256, 0, st>>>(in_indices, n_in, n_in_dev, cg, og, m.bitmap);
      else strided_mark_kernel<long long, 7><<<blocks, 256, 0, st>>>(in_indices, n_in, n_in_dev, cg, og, m.bitmap);
    }
    BEVF_CHECK_LAUNCH();
  }
  // (emitting the sites from inside the scan was measured 2x slower than this separate pass: a scan thread owns four
  // consecutive words and would emit up to 128 sites serially)
  rc = bevf::rank_build_fused(m.bitmap, nwords, m.word_prefix, m.scalars + 64, m.scalars, EmitNothing{}, st);
  if (rc) return rc;
  {
    const int blocks = (int)(bevf::ceil_div(nwords, 256) < kRbBlocks ? bevf::ceil_div(nwords, 256) : kRbBlocks);
    if (nwords < (1ll << 26)) emit_sites_kernel<unsigned><<<blocks, 256, 0, st>>>(m.bitmap, m.word_prefix, nwords, og, out_indices, cap);
    else emit_sites_kernel<long long><<<blocks, 256, 0, st>>>(m.bitmap, m.word_prefix, nwords, og, out_indices, cap);
    BEVF_CHECK_LAUNCH();
  }
  if (n_out_dev)
    BEVF_CHECK_CUDA(cudaMemcpyAsync(n_out_dev, m.scalars, sizeof(int), cudaMemcpyDeviceToDevice, st));
  return BEVF_OK;
}

BEVF_API int bevf_spconv_strided_sites_chain(const int *coords0, int n0, const int *n0_dev, int batch, const int *shape0,
                                             int nlev, const int *ksizes, const int *strides, const int *paddings,
                                             void *const *index_mems, const size_t *index_bytes, int *const *out_indices,
                                             const int *caps, int *const *n_out_devs, void *stream) {
  BEVF_CHECK_ARG(nlev >= 1 && nlev <= kMaxChain, "a chain has 1..%d strided levels (got %d)", kMaxChain, nlev);
  BEVF_CHECK_ARG(n0 >= 0 && (n0 == 0 || coords0) && ksizes && strides && paddings && index_mems && index_bytes &&
                     out_indices && caps && n_out_devs, "NULL argument");
  BEVF_CHECK_ARG((reinterpret_cast<uintptr_t>(coords0) & 15u) == 0, "coordinates must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  ChainArgs A;
  bevf::RankSegments S;
  A.nlev = S.n = nlev;
  A.g0 = Grid{batch, shape0[0], shape0[1], shape0[2]};
  int shape[3] = {shape0[0], shape0[1], shape0[2]};
  const int ones[3] = {1, 1, 1};
  int max_chunks = 1;
  long long max_words = 1;
  for (int l = 0; l < nlev; ++l) {
    const int *k = ksizes + 3 * l, *sd = strides + 3 * l, *pd = paddings + 3 * l;
    int rc = check_geom(k, sd, pd, ones);
    if (rc) return rc;
    int os[3];
    bevf_spconv_out_shape(shape, k, sd, pd, ones, os);
    long long nwords;
    rc = grid_words(batch, os, nwords);
    if (rc) return rc;
    IndexMem m;
    const size_t need = carve_index(m, index_mems[l], index_bytes[l], nwords);
    if (!index_mems[l] || need > index_bytes[l]) {
      bevf::set_error("spconv index memory of chain level %d too small: need %zu bytes, got %zu", l + 1, need, index_bytes[l]);
      return BEVF_ERR_WORKSPACE;
    }
    BEVF_CHECK_ARG(out_indices[l] && (reinterpret_cast<uintptr_t>(out_indices[l]) & 15u) == 0, "bad out_indices of level %d", l + 1);
    // scalars (total, error flag, scan ticket + flags) and the bitmap are adjacent: one clear
    const size_t clear = (size_t)(reinterpret_cast<char *>(m.bitmap + nwords) - reinterpret_cast<char *>(m.scalars));
    BEVF_CHECK_CUDA(cudaMemsetAsync(m.scalars, 0, clear, st));
    ChainLevel &L = A.lv[l];
    for (int j = 0; j < 3; ++j) { L.k[j] = k[j]; L.s[j] = sd[j]; L.p[j] = pd[j]; shape[j] = os[j]; }
    L.og = Grid{batch, os[0], os[1], os[2]};
    L.bitmap = m.bitmap;
    L.word_prefix = m.word_prefix;
    L.total = m.scalars;
    L.nwords = nwords;
    L.out_indices = out_indices[l];
    L.cap = caps[l];
    L.n_out_dev = n_out_devs[l];
    bevf::RankSegment &R = S.seg[l];
    R.bitmap = m.bitmap;
    R.nwords = nwords;
    bevf::rank_plan_chunks(nwords, R.chunk_words, R.nchunks);
    R.sync = m.scalars + 64;
    R.word_prefix = m.word_prefix;
    R.total = m.scalars;
    if (R.nchunks > max_chunks) max_chunks = R.nchunks;
    if (nwords > max_words) max_words = nwords;
  }
  if (n0 > 0) {
    const int blocks = bevf::ceil_div(n0, 256) < kRbBlocks ? bevf::ceil_div(n0, 256) : kRbBlocks;
    mark_levels_kernel<<<blocks, 256, 0, st>>>(coords0, n0, n0_dev, A);
    BEVF_CHECK_LAUNCH();
  }
  bevf::rank_chunk_scan_multi_kernel<<<dim3(max_chunks, nlev), bevf::kRankThreads, 0, st>>>(S);
  BEVF_CHECK_LAUNCH();
  {
    const int blocks = (int)(bevf::ceil_div(max_words, 256) < kRbBlocks ? bevf::ceil_div(max_words, 256) : kRbBlocks);
    emit_sites_multi_kernel<<<dim3(blocks, nlev), 256, 0, st>>>(A);
    BEVF_CHECK_LAUNCH();
  }
  return BEVF_OK;
}

BEVF_API int bevf_spconv_strided_rulebook(const int *out_indices, int n_out, const int *n_out_dev, int batch,
                                          const int *in_shape, const int *ksize, const int *stride,
                                          const int *padding, const int *dilation, const void *in_index_mem,
                                          size_t in_index_bytes, const int *in_perm, int *pair_fwd, int ld,
                                          void *stream) {
  int rc = check_geom(ksize, stride, padding, dilation);
  if (rc) return rc;
  long long nwords;
  rc = grid_words(batch, in_shape, nwords);
  if (rc) return rc;
  BEVF_CHECK_ARG(ld >= n_out, "pair_fwd leading dimension %d < n_out %d", ld, n_out);
  if (ld == 0) return BEVF_OK;
  IndexMem m;
  carve_index(m, const_cast<void *>(in_index_mem), in_index_bytes, nwords);
  IndexView ix{m.bitmap, m.word_prefix, in_perm, Grid{batch, in_shape[0], in_shape[1], in_shape[2]}};
  ConvGeom cg;
  fill_geom(cg, ksize, stride, padding, dilation);
  const long long threads = (long long)(n_out_dev ? ld : n_out) * ksize[0] * ksize[1];
  BEVF_CHECK_ARG(threads < (1ll << 32), "rulebook with %lld (site, tap column) items is not supported", threads);
  if (threads == 0) return BEVF_OK;
  const bool k3 = ksize[0] == 3 && ksize[1] == 3 && ksize[2] == 3 && dilation[0] == 1 && dilation[1] == 1 && dilation[2] == 1;
  if (k3 && !in_perm && nwords < (1ll << 26)) {
    const int rows = n_out_dev ? ld : n_out;
    const int bx = bevf::ceil_div(rows, 256) < kRbBlocks / 3 ? bevf::ceil_div(rows, 256) : kRbBlocks / 3;
    rulebook_k3_kernel<false><<<dim3(bx, 3), 256, 0, (cudaStream_t)stream>>>(out_indices, n_out, n_out_dev, ld, ix, cg, ld,
                                                                            pair_fwd);
    BEVF_CHECK_LAUNCH();
    return BEVF_OK;
  }
  const int blocks = (int)(bevf::ceil_div(threads, 256) < kRbBlocks ? bevf::ceil_div(threads, 256) : kRbBlocks);
  if (nwords < (1ll << 26))
    strided_rulebook_kernel<unsigned><<<blocks, 256, 0, (cudaStream_t)stream>>>(out_indices, n_out_dev, n_out, ix, cg, ld,
                                                                                pair_fwd);
  else
    strided_rulebook_kernel<long long><<<blocks, 256, 0, (cudaStream_t)stream>>>(out_indices, n_out_dev, n_out, ix, cg, ld,
                                                                                 pair_fwd);
  BEVF_CHECK_LAUNCH();
  return BEVF_OK;
}

BEVF_API int bevf_sparse_to_dense(const float *feats, const int *indices, int n, const int *n_dev, int c, int batch,
                                  const int *shape, float *dense, int bev_layout, void *stream) {
  BEVF_CHECK_ARG(batch > 0 && c > 0 && n >= 0, "bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  size_t total = (size_t)batch * c * shape[0] * shape[1] * shape[2];
  BEVF_CHECK_CUDA(cudaMemsetAsync(dense, 0, total * sizeof(float), st));
  if (n == 0) return BEVF_OK;
  Grid g{batch, shape[0], shape[1], shape[2]};
  long long threads = (long long)n * c;
  if (bev_layout) {
    const int blocks = bevf::ceil_div(n, 32) < kRbBlocks ? bevf::ceil_div(n, 32) : kRbBlocks;
    BEVF_CHECK_ARG((size_t)32 * (c + 1) * sizeof(float) <= 48 * 1024, "too many channels for the BEV scatter (%d)", c);
    to_bev_kernel<<<blocks, 256, (size_t)32 * (c + 1) * sizeof(float), st>>>(feats, indices, n, n_dev, c, g, dense);
  }
  else to_dense_kernel<<<bevf::ceil_div(threads, 256), 256, 0, st>>>(feats, indices, n, n_dev, c, g, dense);
  BEVF_CHECK_LAUNCH();
  return BEVF_OK;
}

BEVF_API int bevf_sparse_to_bev_indexed(const float *feats, int n_cap, const int *n_dev, int c, int batch, const int *shape,
                                        const void *index_mem, size_t index_bytes, float *bev, void *stream) {
  BEVF_CHECK_ARG(batch > 0 && c > 0 && c % 4 == 0 && n_cap >= 0 && feats && bev && index_mem, "bad arguments (c % 4 == 0)");
  BEVF_CHECK_ARG((reinterpret_cast<uintptr_t>(feats) & 15u) == 0, "feats must be 16-byte aligned");
  long long nwords;
  int rc = grid_words(batch, shape, nwords);
  if (rc) return rc;
  IndexMem m;
  carve_index(m, const_cast<void *>(index_mem), index_bytes, nwords);
  Grid g{batch, shape[0], shape[1], shape[2]};
  const size_t cells = (size_t)shape[1] * shape[2];
  // tile: a full line x 32 channels when that keeps five CTAs on an SM (<= 44 KB), else what fits (at least 5 floats per
  // row of a full line: the kernel then takes the channels in several passes)
  size_t tile_floats = cells * (kBevCh + 1);
  const size_t budget = 44 * 1024;
  if (cells * sizeof(int) + tile_floats * sizeof(float) > budget && budget > cells * sizeof(int))
    tile_floats = (budget - cells * sizeof(int)) / sizeof(float);
  if (tile_floats < cells * 5) tile_floats = cells * 5;
  const size_t smem = cells * sizeof(int) + tile_floats * sizeof(float);
  if (smem > 200 * 1024) {
    bevf::set_error("BEV tail: a line of %zu cells does not fit in shared memory", cells);
    return BEVF_ERR_UNSUPPORTED;
  }
  static bevf::DeviceOnce conf;
  if (conf.first())
    BEVF_CHECK_CUDA(cudaFuncSetAttribute(to_bev_indexed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int chunks = bevf::ceil_div(c, kBevCh);
  to_bev_indexed_kernel<<<dim3(shape[0] * chunks, batch), 256, smem, (cudaStream_t)stream>>>(
      feats, m.bitmap, m.word_prefix, n_dev, n_cap, c, g, (int)tile_floats, bev);
  BEVF_CHECK_LAUNCH();
  return BEVF_OK;
}

BEVF_API int bevf_dense_to_sparse(const float *dense, const int *indices, int n, const int *n_dev, int c, int batch,
                                  const int *shape, int bev_layout, float *feats, void *stream) {
  BEVF_CHECK_ARG(batch > 0 && c > 0 && n >= 0, "bad sizes");
  if (n == 0) return BEVF_OK;
  BEVF_CHECK_ARG(dense && indices && feats, "NULL tensor");
  Grid g{batch, shape[0], shape[1], shape[2]};
  const long long threads = (long long)n * c;
  from_dense_kernel<<<bevf::ceil_div(threads, 256), 256, 0, (cudaStream_t)stream>>>(dense, indices, n, n_dev, c, g,
                                                                                   bev_layout, feats);
  BEVF_CHECK_LAUNCH();
  return BEVF_OK;
}

BEVF_API int bevf_spconv_permute_rows(const float *feats, const int *indices, const int *perm, int n, const int *n_dev,
                                      int c, int cin_pad, float *out_f32, void *out_bf16, int *out_indices,
                                      void *stream) {
  BEVF_CHECK_ARG(n >= 0 && c > 0, "bad shapes");
  BEVF_CHECK_ARG(out_bf16 == nullptr || cin_pad >= c, "cin_pad %d does not cover the %d feature columns", cin_pad, c);
  if (n == 0) return BEVF_OK;
  BEVF_CHECK_ARG(feats && indices && perm && out_indices, "NULL tensor");
  int width = c > 4 ? c : 4;
  if (out_bf16 && cin_pad > width) width = cin_pad;
  const long long threads = (long long)n * width;
  permute_rows_kernel<<<bevf::ceil_div(threads, 256), 256, 0, (cudaStream_t)stream>>>(
      feats, indices, perm, n, n_dev, c, out_bf16 ? cin_pad : 0, width, out_f32, (unsigned short *)out_bf16, out_indices);
  BEVF_CHECK_LAUNCH();
  return BEVF_OK;
}
