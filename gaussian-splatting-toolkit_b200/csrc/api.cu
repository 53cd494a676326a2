// api.cu — library identification + thread-local error reporting of the C ABI (include/gsr_b200.h).
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"

namespace gsr {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static std::atomic<unsigned long long> g_launches{0};
// GSR_NVTX=1: NVTX ranges over the entry points, one marker per kernel launch (header-only NVTX 3: no-ops unless a tool
// is attached to the process)
static bool nvtx_on() {
  static const bool v = [] {
    const char *e = getenv("GSR_NVTX");
    return e && e[0] == '1';
  }();
  return v;
}
void count_launch(const char *kernel_name) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (kernel_name && nvtx_on()) nvtxMarkA(kernel_name);
}
TraceScope::TraceScope(const char *name) : on(nvtx_on()) {
  if (on) nvtxRangePushA(name);
}
TraceScope::~TraceScope() {
  if (on) nvtxRangePop();
}
}  // namespace gsr

extern "C" {
GSR_API const char *gsr_version(void) { return "0.1.2+b200.2"; }
GSR_API const char *gsr_last_error(void) { return gsr::g_err; }
GSR_API int gsr_built_for_sm(void) { return 100; }
GSR_API unsigned long long gsr_launch_count(void) { return gsr::g_launches.load(std::memory_order_relaxed); }
}
