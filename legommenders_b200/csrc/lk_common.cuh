// Shared device/host helpers for the legommenders_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define LK_OK 0
#define LK_ERR_CUDA (-1)
#define LK_ERR_SHAPE (-2)
#define LK_ERR_ARG (-3)

namespace lk {

void set_error(const char* fmt, ...);
void count_launches(int n);   // kernels launched by this library (reported by lk_launch_count)
bool prof_enabled();          // lk_profile_enable(1): native drivers bracket every sub-call with CUDA events
void prof_begin(const char* expr, double flops, cudaStream_t st);
void prof_end(cudaStream_t st);

inline int check_launch(const char* what, int n = 1) {
  count_launches(n);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return LK_ERR_CUDA;
  }
  return LK_OK;
}

#define LK_REQUIRE(cond, code, ...)   \
  do {                                \
    if (!(cond)) {                    \
      lk::set_error(__VA_ARGS__);     \
      return (code);                  \
    }                                 \
  } while (0)

constexpr int kNumSMs = 148;  // B200

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------------
// A training step is ~80 short kernels in one stream; without PDL every boundary pays launch latency + grid ramp-up after
// the predecessor has fully drained.  Every kernel of this library starts with pdl_prologue(): it lets the NEXT kernel in
// the stream be scheduled as soon as this grid's last CTA is resident (griddepcontrol.launch_dependents), and then blocks
// until the PREVIOUS grid has completed and its memory is visible (griddepcontrol.wait) before touching global memory.
// Every kernel waits on its immediate predecessor before doing anything, so stream order is preserved transitively
// (reads after writes and buffer reuse alike); only launch latency and CTA scheduling overlap.  LK_PDL=0 turns it off.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
  pdl_trigger();
  pdl_wait();
}
bool pdl_enabled();

// Id bounds (the reference raises IndexError on an out-of-range id; here the kernels never dereference one): an id outside
// [0, rows) reads as an INVALID position (zero row / zero score) and is counted in the device counter registered with
// lk_set_id_violation_counter (null: not counted).  Host code reads the counter at its next natural synchronisation point.
int32_t* id_violations();
__device__ __forceinline__ bool id_in_range(int64_t id, int64_t rows, int32_t* viol) {
  if ((uint64_t)id < (uint64_t)rows) return true;
  if (viol) atomicAdd(viol, 1);
  return false;
}
// warp-cooperative form: every lane holds the same id and evaluates the (register-only) test itself; lane 0 alone reports a violation.
// No shuffle sits between the id load and the row loads that depend on it.
__device__ __forceinline__ bool id_in_range_warp(int64_t id, int64_t rows, int32_t* viol, int lane) {
  const bool ok = (uint64_t)id < (uint64_t)rows;
  if (!ok && lane == 0 && viol) atomicAdd(viol, 1);
  return ok;
}

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);   // errors surface through cudaGetLastError in check_launch
}
#define LK_LAUNCH(kernel, grid, block, smem, st, ...) \
  lk::launch_pdl(kernel, dim3(grid), dim3(block), (size_t)(smem), st, ##__VA_ARGS__)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// streaming (read-once) 16-byte load that does not pollute L1
__device__ __forceinline__ float4 ldg4_stream(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void f4_fma(float4& a, float s, const float4& x) {
  a.x = fmaf(s, x.x, a.x);
  a.y = fmaf(s, x.y, a.y);
  a.z = fmaf(s, x.z, a.z);
  a.w = fmaf(s, x.w, a.w);
}
__device__ __forceinline__ void f4_add(float4& a, const float4& x) {
  a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
}
__device__ __forceinline__ float f4_dot(const float4& a, const float4& b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}

// Counter-based RNG for dropout (statistics, not bits, are what the reference defines).
__device__ __forceinline__ uint32_t mix32(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return (uint32_t)x;
}
// keep-scale for element `idx` of dropout stream `seed`: 0 with prob p, 1/(1-p) otherwise.
__device__ __forceinline__ float dropout_scale(uint64_t seed, uint64_t idx, float p, float inv_keep) {
  uint32_t r = mix32(seed * 0x9E3779B97F4A7C15ULL + idx);
  return ((r >> 8) * (1.0f / 16777216.0f)) < p ? 0.f : inv_keep;
}

}  // namespace lk
