// Error plumbing and version of the C ABI (include/s4g_b200.h).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace s4g {

char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

static unsigned long long g_launches = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }

}  // namespace s4g

extern "C" int s4g_version(void) { return 1; }
extern "C" unsigned long long s4g_launch_count(void) { return __atomic_load_n(&s4g::g_launches, __ATOMIC_RELAXED); }
extern "C" const char* s4g_last_error(void) { return s4g::error_buffer(); }
