// Library-level entry points: version, per-thread error string, device check.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

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
bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("LK_PDL"); return !(e && e[0] == '0'); }();
  return on;
}
static int32_t* g_id_viol = nullptr;
int32_t* id_violations() { return g_id_viol; }
static unsigned long long g_launches = 0;
void count_launches(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }

// ---- per-call device timing inside native drivers (bench.py's kernel shares and the live roofline numerator) --------
struct ProfRec { char name[40]; double flops; cudaEvent_t e0, e1; };
static bool g_prof = false;
static std::vector<ProfRec> g_recs;
bool prof_enabled() { return g_prof; }
void prof_begin(const char* expr, double flops, cudaStream_t st) {
  ProfRec r;
  size_t n = 0;
  while (expr[n] && expr[n] != '(' && n + 1 < sizeof(r.name)) { r.name[n] = expr[n]; n++; }   // entry-point name or caller-made label
  r.name[n] = 0;
  r.flops = flops;
  cudaEventCreate(&r.e0);
  cudaEventCreate(&r.e1);
  cudaEventRecord(r.e0, st);
  g_recs.push_back(r);
}
void prof_end(cudaStream_t st) { cudaEventRecord(g_recs.back().e1, st); }
}  // namespace lk

extern "C" {

unsigned long long lk_launch_count(void) { return __atomic_load_n(&lk::g_launches, __ATOMIC_RELAXED); }

void lk_set_id_violation_counter(int32_t* device_counter) { lk::g_id_viol = device_counter; }

const char* lk_version(void) { return "legommenders_b200 0.1 (sm_100a)"; }

const char* lk_last_error(void) { return lk::g_err; }

void lk_profile_enable(int on) {
  for (auto& r : lk::g_recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  lk::g_recs.clear();
  lk::g_prof = on != 0;
}

int lk_profile_collect(char* names, int name_stride, float* ms, double* flops, int* calls, int cap) {
  int n = 0;
  for (auto& r : lk::g_recs) {
    if (cudaEventSynchronize(r.e1) != cudaSuccess) { lk::set_error("lk_profile_collect: event sync failed"); return LK_ERR_CUDA; }
    float t = 0.f;
    cudaEventElapsedTime(&t, r.e0, r.e1);
    int i = 0;
    for (; i < n; i++)
      if (strncmp(names + (size_t)i * name_stride, r.name, name_stride) == 0) break;
    if (i == n) {
      if (n == cap) continue;
      strncpy(names + (size_t)i * name_stride, r.name, name_stride - 1);
      names[(size_t)i * name_stride + name_stride - 1] = 0;
      ms[i] = 0.f; flops[i] = 0.0; calls[i] = 0;
      n++;
    }
    ms[i] += t; flops[i] += r.flops; calls[i] += 1;
  }
  return n;
}

int lk_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 0;
  return prop.major == 10 ? 1 : 0;
}

}  // extern "C"
