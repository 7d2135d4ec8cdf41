// Gradient all-reduce over NVLink peer memory (data-parallel training): the flat gradient bucket of every rank lives in symmetric
// memory (peer-mapped over NVSwitch, plus the switch's multicast mapping).  One kernel per rank reduces ITS 1/W slice — in the switch
// (multimem.ld_reduce over the W replicas) or by reading the slice from all W buffers in rank order — and writes the sum into all W buffers
// (multimem.st, or W peer stores).  Every element is summed once, by one rank, so all ranks end up with bit-identical gradients.
// The two cross-rank rendezvous (gradients ready / sums visible) are flag words in symmetric memory polled inside the same launch
// (allreduce_fused_kernel); with flag_ptrs == NULL the caller brackets the plain slice kernel with its own barriers instead.
// 3.49 MB on 8 GPUs: 23.4 us (multicast) / 29.5 us (peer loads) against 50.6 us for ncclAllReduce, profiles/r2_05_allreduce.md.
#include "lk_common.cuh"
#include "../../include/legommenders_b200.h"

namespace lk {
namespace ar {

constexpr int MAXW = 16;
constexpr long long kSpinLimitClk = 120ll * 1900000000ll;      // ~2 minutes of SM clocks
struct Peers { float* p[MAXW]; };

__device__ __forceinline__ float4 ld_sys(const float* a) {      // peer memory: system-scope, never from a stale L1 line
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys(float* a, const float4& v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// The same two accesses through the NVSwitch multicast mapping of the bucket: the load is reduced IN THE SWITCH over all W replicas (one
// 16-byte response instead of W), the store is replicated by the switch to all W buckets (one 16-byte request instead of W).
__device__ __forceinline__ float4 ld_reduce_mc(const float* a) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a) : "memory");
  return v;
}
__device__ __forceinline__ void st_mc(float* a, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Cross-rank rendezvous of block b with block b of every peer, through flag words in symmetric memory: flags[b * MAXW + r] of rank q is written
// only by rank r's block b, with a monotonically increasing epoch (never reset, so there is no reset race).  Every rank launches the same grid.
// RELEASE = true publishes what this BLOCK wrote before the preceding __syncthreads (the fence of the signalling thread is cumulative over
// the block barrier: the same pattern as a grid-wide sync), so the other threads need no system fence of their own.
template <bool RELEASE>
__device__ __forceinline__ void signal_peers(const Peers& flags, int rank, int W, int epoch) {
  if ((int)threadIdx.x < W) {
    int* f = (int*)flags.p[threadIdx.x] + (size_t)blockIdx.x * MAXW + rank;
    if (RELEASE) asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
    else asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
  }
}
__device__ __forceinline__ void wait_peers(const Peers& flags, int rank, int W, int epoch) {
  if ((int)threadIdx.x < W) {
    const int* f = (const int*)flags.p[rank] + (size_t)blockIdx.x * MAXW + threadIdx.x;
    int v;
    const long long t0 = clock64();
    for (;;) {
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if (v - epoch >= 0) break;
      if (clock64() - t0 > kSpinLimitClk) {                // a peer died or never launched: fail loudly instead of hanging the GPU
        printf("lk_allreduce_p2p: rank %d block %d waited %llds for rank %d (flag %d, expected %d)\n", rank, (int)blockIdx.x,
               kSpinLimitClk / 1900000000ll, (int)threadIdx.x, v, epoch);
        __trap();
      }
    }
  }
  __syncthreads();
}

// FUSED: gradients-ready rendezvous -> slice reduction -> sums-visible rendezvous, one launch.  Block b only touches elements that block b of
// the other ranks does not, and leaves only after block b of every peer has finished (its reads of this rank's bucket and its writes into it),
// so when the grid completes every block of every peer is done with this rank's bucket: the next kernel on the stream may overwrite it.
template <int WT, bool FENCE_ALL = false>
__global__ void __launch_bounds__(1024) allreduce_fused_kernel(const __grid_constant__ Peers peers, const __grid_constant__ Peers flags, int rank, int W_rt, int64_t per4, int64_t n4,
                                                              float scale, int epoch, long long* trace, float* mc) {
  const int W = WT > 0 ? WT : W_rt;
  long long t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
  if (trace) t0 = clock64();
  pdl_prologue();                                          // the producer of this rank's gradients has completed and flushed
  if (trace) t1 = clock64();
  signal_peers<false>(flags, rank, W, 2 * epoch);          // nothing of THIS grid to publish; the producer grid's writes are complete (line above)
  wait_peers(flags, rank, W, 2 * epoch);
  if (trace) t2 = clock64();
  const int64_t lo4 = (int64_t)rank * per4, hi4 = lo4 + per4 < n4 ? lo4 + per4 : n4;
  for (int64_t i = lo4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 s;
    if (mc != nullptr) {                                   // switch-side reduction and broadcast (the summation order is the switch's)
      s = ld_reduce_mc(mc + i * 4);
      s.x *= scale; s.y *= scale; s.z *= scale; s.w *= scale;
      st_mc(mc + i * 4, s);
      continue;
    }
    if (WT > 0) {
      float4 v[WT > 0 ? WT : 1];
#pragma unroll
      for (int r = 0; r < WT; r++) v[r] = ld_sys(peers.p[r] + i * 4);       // all W loads in flight before the first add
      s = v[0];
#pragma unroll
      for (int r = 1; r < WT; r++) f4_add(s, v[r]);
    } else {
      s = ld_sys(peers.p[0] + i * 4);
      for (int r = 1; r < W; r++) f4_add(s, ld_sys(peers.p[r] + i * 4));
    }
    s.x *= scale; s.y *= scale; s.z *= scale; s.w *= scale;
#pragma unroll
    for (int r = 0; r < (WT > 0 ? WT : MAXW); r++)
      if (r < W) st_sys(peers.p[(rank + r) % W] + i * 4, s);                 // start with the own buffer: spreads the W writers over the links
  }
  if (trace) t3 = clock64();
  if (FENCE_ALL) __threadfence_system();
  __syncthreads();                                         // the block's peer stores happen-before the signalling threads' release
  if (trace) t4 = clock64();
  signal_peers<true>(flags, rank, W, 2 * epoch + 1);
  wait_peers(flags, rank, W, 2 * epoch + 1);
  if (trace && threadIdx.x == 0) {                        // LK_AR_TRACE: SM clocks of the phases of this block (scratch/bench_allreduce.py)
    long long* t = trace + (size_t)blockIdx.x * 6;
    t[0] = t1 - t0; t[1] = t2 - t1; t[2] = t3 - t2; t[3] = t4 - t3; t[4] = clock64() - t4; t[5] = clock64() - t0;
  }
}

__global__ void __launch_bounds__(512) allreduce_slice_kernel(const __grid_constant__ Peers peers, int rank, int W, int64_t lo4, int64_t hi4, float scale) {
  pdl_prologue();
  for (int64_t i = lo4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 s = ld_sys(peers.p[0] + i * 4);
    for (int r = 1; r < W; r++) f4_add(s, ld_sys(peers.p[r] + i * 4));
    s.x *= scale; s.y *= scale; s.z *= scale; s.w *= scale;
    for (int r = 0; r < W; r++) st_sys(peers.p[(rank + r) % W] + i * 4, s);
  }
}

}  // namespace ar
}  // namespace lk

using namespace lk;

extern "C" {

static long long* g_ar_trace = nullptr;
// debugging aid: 6 int64 per block (SM clocks: launch->dependency, rendezvous, reduction, fence, rendezvous, total) of every later fused launch
int lk_allreduce_set_trace(void* buf) { g_ar_trace = (long long*)buf; return LK_OK; }

// peer_ptrs: HOST array of W device pointers (this rank's view of every rank's bucket, index = rank); n floats, n % 4 == 0.
// flag_ptrs: HOST array of W device pointers to each rank's flag words (LK_ALLREDUCE_FLAG_WORDS int32, zeroed once, symmetric memory), or NULL:
//   with flags the rendezvous before and after the reduction are inside the launch (epoch = 1, 2, 3 ... per call, the same on every rank);
//   without, the caller brackets the launch with its own cross-rank barriers.
int lk_allreduce_p2p(void* const* peer_ptrs, void* const* flag_ptrs, void* multicast_ptr, int epoch, int rank, int W, int64_t n, float scale,
                     cudaStream_t st) {
  LK_REQUIRE(W >= 1 && W <= ar::MAXW && rank >= 0 && rank < W && n % 4 == 0, LK_ERR_ARG, "lk_allreduce_p2p: W=%d rank=%d n=%ld", W, rank, (long)n);
  if (n == 0) return LK_OK;
  ar::Peers peers, flags;
  for (int r = 0; r < W; r++) {
    LK_REQUIRE(peer_ptrs[r] != nullptr && (uintptr_t)peer_ptrs[r] % 16 == 0, LK_ERR_ARG, "lk_allreduce_p2p: peer %d pointer", r);
    peers.p[r] = (float*)peer_ptrs[r];
    if (flag_ptrs) {
      LK_REQUIRE(flag_ptrs[r] != nullptr, LK_ERR_ARG, "lk_allreduce_p2p: peer %d flag pointer", r);
      flags.p[r] = (float*)flag_ptrs[r];
    }
  }
  const int64_t n4 = n / 4, per = (n4 + W - 1) / W;
  if (flag_ptrs) {
    LK_REQUIRE(epoch > 0, LK_ERR_ARG, "lk_allreduce_p2p: epoch starts at 1");
    float* mc = (float*)multicast_ptr;
    LK_REQUIRE((uintptr_t)mc % 16 == 0, LK_ERR_ARG, "lk_allreduce_p2p: multicast pointer alignment");
    // the SAME grid on every rank (blocks pair up across ranks), one block per SM at most; one float4 per thread when the slice allows it:
    // the reduction is latency-bound (a round trip over NVSwitch per load), so parallelism rather than a grid-stride loop
    static const int thr_env = getenv("LK_AR_THREADS") ? atoi(getenv("LK_AR_THREADS")) : 0;
    int threads = thr_env > 0 ? thr_env : (per <= (int64_t)kNumSMs * 256 ? 256 : 512)    /* measured at W=2: 512 beats 1024 (19.2 vs 21.2 us) */;
    int64_t blocks = (per + threads - 1) / threads;
    if (blocks > kNumSMs) blocks = kNumSMs;
    static_assert(kNumSMs * ar::MAXW <= LK_ALLREDUCE_FLAG_WORDS, "flag words");
    switch (W) {
      case 2: LK_LAUNCH((ar::allreduce_fused_kernel<2>), (unsigned)blocks, threads, 0, st, peers, flags, rank, W, per, n4, scale, epoch, g_ar_trace, mc); break;
      case 4: LK_LAUNCH((ar::allreduce_fused_kernel<4>), (unsigned)blocks, threads, 0, st, peers, flags, rank, W, per, n4, scale, epoch, g_ar_trace, mc); break;
      case 8: LK_LAUNCH((ar::allreduce_fused_kernel<8>), (unsigned)blocks, threads, 0, st, peers, flags, rank, W, per, n4, scale, epoch, g_ar_trace, mc); break;
      default: LK_LAUNCH((ar::allreduce_fused_kernel<0>), (unsigned)blocks, threads, 0, st, peers, flags, rank, W, per, n4, scale, epoch, g_ar_trace, mc);
    }
    return check_launch("allreduce_p2p");
  }
  LK_REQUIRE(multicast_ptr == nullptr, LK_ERR_ARG, "lk_allreduce_p2p: the multicast path needs the flag words");
  const int64_t lo4 = (int64_t)rank * per, hi4 = lo4 + per < n4 ? lo4 + per : n4;
  if (lo4 >= hi4) return LK_OK;
  int64_t blocks = (hi4 - lo4 + 511) / 512;
  if (blocks > 2 * kNumSMs) blocks = 2 * kNumSMs;
  LK_LAUNCH((ar::allreduce_slice_kernel), (unsigned)blocks, 512, 0, st, peers, rank, W, lo4, hi4, scale);
  return check_launch("allreduce_p2p");
}

}  // extern "C"
