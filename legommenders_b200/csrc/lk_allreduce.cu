// Gradient all-reduce over NVLink peer memory (data-parallel training): the flat gradient bucket of every rank lives in symmetric
// memory (peer-mapped over NVSwitch); one kernel per rank reduces ITS 1/W slice by reading that slice from all W buffers (fixed order
// 0..W-1: every element is summed once, by one rank, so all ranks end up with bit-identical gradients) and writes the sum back into all W
// buffers.  3.5 MB on 8 GPUs: 3 MB in and 3 MB out per rank over NVLink instead of a ring / tree of NCCL steps (measured 63-71 us per step for
// ncclAllReduce at this size, profiles/r2_05).  The caller brackets the kernel with two cross-rank barriers (gradients ready / sums visible);
// they are the symmetric-memory handle's device-side barriers, enqueued on the same stream.
#include "lk_common.cuh"
#include "../../include/legommenders_b200.h"

namespace lk {
namespace ar {

constexpr int MAXW = 16;
struct Peers { float* p[MAXW]; };

__device__ __forceinline__ float4 ld_sys(const float* a) {      // peer memory: system-scope, never from a stale L1 line
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys(float* a, const float4& v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(512) allreduce_slice_kernel(const Peers peers, int rank, int W, int64_t lo4, int64_t hi4, float scale) {
  pdl_prologue();
  for (int64_t i = lo4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 s = ld_sys(peers.p[0] + i * 4);
    for (int r = 1; r < W; r++) f4_add(s, ld_sys(peers.p[r] + i * 4));
    s.x *= scale; s.y *= scale; s.z *= scale; s.w *= scale;
    for (int r = 0; r < W; r++) st_sys(peers.p[(rank + r) % W] + i * 4, s);      // start with the own buffer: spreads the W writers over the links
  }
}

}  // namespace ar
}  // namespace lk

using namespace lk;

extern "C" {

// peer_ptrs: HOST array of W device pointers (this rank's view of every rank's bucket, index = rank); n floats, n % 4 == 0
int lk_allreduce_p2p(void* const* peer_ptrs, int rank, int W, int64_t n, float scale, cudaStream_t st) {
  LK_REQUIRE(W >= 1 && W <= ar::MAXW && rank >= 0 && rank < W && n % 4 == 0, LK_ERR_ARG, "lk_allreduce_p2p: W=%d rank=%d n=%ld", W, rank, (long)n);
  if (n == 0) return LK_OK;
  ar::Peers peers;
  for (int r = 0; r < W; r++) {
    LK_REQUIRE(peer_ptrs[r] != nullptr && (uintptr_t)peer_ptrs[r] % 16 == 0, LK_ERR_ARG, "lk_allreduce_p2p: peer %d pointer", r);
    peers.p[r] = (float*)peer_ptrs[r];
  }
  const int64_t n4 = n / 4, per = (n4 + W - 1) / W;
  const int64_t lo4 = (int64_t)rank * per, hi4 = lo4 + per < n4 ? lo4 + per : n4;
  if (lo4 >= hi4) return LK_OK;
  int64_t blocks = (hi4 - lo4 + 511) / 512;
  if (blocks > 2 * kNumSMs) blocks = 2 * kNumSMs;
  LK_LAUNCH((ar::allreduce_slice_kernel), (unsigned)blocks, 512, 0, st, peers, rank, W, lo4, hi4, scale);
  return check_launch("allreduce_p2p");
}

}  // extern "C"
