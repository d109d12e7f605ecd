// common.cuh -- shared helpers for libbevfront_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "bevfront_b200.h"

#define BEVF_API extern "C" __attribute__((visibility("default")))

namespace bevf {

// thread-local last-error text (bevf_last_error)
void set_error(const char *fmt, ...);
// process-wide count of kernel launches issued by this library (bevf_launch_count)
void count_launch();

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Simple bump allocator over a caller-owned workspace.
struct Workspace {
  char *base;
  size_t size;
  size_t off;
  __host__ Workspace(void *p, size_t n) : base((char *)p), size(n), off(0) {}
  template <typename T>
  __host__ T *take(size_t count) {
    off = align_up(off, 256);
    T *r = (T *)(base + off);
    off += count * sizeof(T);
    return r;
  }
  __host__ bool ok() const { return base != nullptr && off <= size; }
};

constexpr int kNumSMs = 148;  // B200

// Per-device "already configured" flag for function attributes (cudaFuncSetAttribute is per device / context, so a
// process that drives several GPUs must set it on each).  One instance per call site: `static DeviceOnce once;`
// then `if (once.first()) { ...set attributes... }`.  Thread-safe (atomic flags), at most 64 devices.
struct DeviceOnce {
  unsigned char done[64] = {0};
  bool first() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;   // unknown device: always configure
    return __atomic_exchange_n(&done[dev], (unsigned char)1, __ATOMIC_ACQ_REL) == 0;
  }
};

}  // namespace bevf

#define BEVF_CHECK_ARG(cond, ...)              \
  do {                                         \
    if (!(cond)) {                             \
      bevf::set_error(__VA_ARGS__);            \
      return BEVF_ERR_INVALID_ARGUMENT;        \
    }                                          \
  } while (0)

#define BEVF_CHECK_CUDA(expr)                                                                  \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      bevf::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return BEVF_ERR_CUDA;                                                                    \
    }                                                                                          \
  } while (0)

#define BEVF_CHECK_LAUNCH()             \
  do {                                  \
    bevf::count_launch();               \
    BEVF_CHECK_CUDA(cudaGetLastError()); \
  } while (0)
