// api.cu — library identification + thread-local error reporting of the C ABI (include/gsr_b200.h).
#include <stdarg.h>

#include <atomic>

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
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace gsr

extern "C" {
GSR_API const char *gsr_version(void) { return "0.1.2+b200.2"; }
GSR_API const char *gsr_last_error(void) { return gsr::g_err; }
GSR_API int gsr_built_for_sm(void) { return 100; }
GSR_API unsigned long long gsr_launch_count(void) { return gsr::g_launches.load(std::memory_order_relaxed); }
}
