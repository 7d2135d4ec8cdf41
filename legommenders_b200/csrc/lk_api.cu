// Library-level entry points: version, per-thread error string, device check.
#include <stdarg.h>
#include <string.h>

#include "lk_common.cuh"
#include "../../include/legommenders_b200.h"

namespace lk {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static unsigned long long g_launches = 0;
void count_launches(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }
}  // namespace lk

extern "C" {

unsigned long long lk_launch_count(void) { return __atomic_load_n(&lk::g_launches, __ATOMIC_RELAXED); }

const char* lk_version(void) { return "legommenders_b200 0.1 (sm_100a)"; }

const char* lk_last_error(void) { return lk::g_err; }

int lk_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 0;
  return prop.major == 10 ? 1 : 0;
}

}  // extern "C"
