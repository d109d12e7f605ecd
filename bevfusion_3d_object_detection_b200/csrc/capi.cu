// capi.cu -- ABI bookkeeping for libbevfront_b200 (version, last-error text).
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace bevf {
static thread_local char g_last_error[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace bevf

BEVF_API int bevf_abi_version(void) { return 1; }
BEVF_API const char *bevf_last_error(void) { return bevf::g_last_error; }
BEVF_API int bevf_compiled_arch(void) { return 100; }
BEVF_API long long bevf_launch_count(void) { return bevf::g_launches.load(std::memory_order_relaxed); }
