// Row-sharded word table (BASELINE config 4): the integer side of the all-to-all lookup, on the device.
//
// A rank needs the rows of the DISTINCT token ids of its batch; row i lives on rank i % W at local index i / W.  Per step:
//   lk_shard_plan     dedup (open-addressing hash, atomicCAS) + bucket by owner in ONE pass over the ids: the first thread to claim an id
//                     takes the next slot of its owner's FIXED-CAPACITY bucket, send_ids[owner*cap + slot] = id
//   lk_shard_inverse  every position -> index of its id's slot (= row of the receive buffer), -1 for unset positions
//   (NCCL all-to-all of the id buckets: equal splits, no sizes to exchange, no host round trip)
//   lk_shard_gather   owner side: out[j] = local[id / W] for the valid ids of the received buckets, zero rows for the padding
//   (NCCL all-to-all of the row buckets back)
// Slot order inside a bucket depends on thread scheduling; the VALUE every position finally gathers is a bit-exact copy of its table row,
// so results do not depend on it.  Bucket overflow (more distinct ids for one owner than `cap`) is counted in `overflow` and must be
// checked by the host before the results are trusted (sharding.ShardedTable re-plans with a larger capacity).
#include "lk_common.cuh"
#include "../../include/legommenders_b200.h"

namespace lk {
namespace shard {

__device__ __forceinline__ uint32_t hash64(int64_t x) {
  uint64_t z = (uint64_t)x * 0x9E3779B97F4A7C15ULL;
  z ^= z >> 32;
  return (uint32_t)z;
}

__global__ void __launch_bounds__(256) plan_kernel(const int64_t* __restrict__ ids, int64_t P, int W, int cap, long long* __restrict__ keys,
                                                   int32_t* __restrict__ vals, uint32_t hmask, int32_t* __restrict__ counts,
                                                   int64_t* __restrict__ send_ids, int32_t* __restrict__ overflow) {
  pdl_prologue();
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t id = ids[p];
    if (id < 0) continue;
    uint32_t h = hash64(id) & hmask;
    while (true) {
      const long long prev = atomicCAS((unsigned long long*)&keys[h], (unsigned long long)-1LL, (unsigned long long)id);
      if (prev == -1LL) {                       // this thread claimed the id
        const int owner = (int)(id % W);
        const int slot = atomicAdd(&counts[owner], 1);
        if (slot < cap) {
          send_ids[(int64_t)owner * cap + slot] = id;
          vals[h] = owner * cap + slot;
        } else {
          vals[h] = -1;
          atomicAdd(overflow, 1);
        }
        break;
      }
      if (prev == id) break;
      h = (h + 1) & hmask;
    }
  }
}

__global__ void __launch_bounds__(256) inverse_kernel(const int64_t* __restrict__ ids, int64_t P, const long long* __restrict__ keys,
                                                      const int32_t* __restrict__ vals, uint32_t hmask, int64_t* __restrict__ inverse) {
  pdl_prologue();
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t id = ids[p];
    int64_t r = -1;
    if (id >= 0) {
      uint32_t h = hash64(id) & hmask;
      while (keys[h] != id) h = (h + 1) & hmask;      // present by construction
      r = vals[h];
    }
    inverse[p] = r;
  }
}

// out[j,:] = local[ids[j] / W, :] for ids[j] >= 0, zeros otherwise; one warp per row, 16-byte loads
__global__ void __launch_bounds__(256) gather_kernel(const int64_t* __restrict__ ids, int64_t M, int W, const float* __restrict__ local,
                                                     int64_t local_rows, int E, float* __restrict__ out, int32_t* viol) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * 8;
  const int E4 = E >> 2;
  for (int64_t j = warp; j < M; j += nwarps) {
    const int64_t id = ids[j];
    bool ok = id >= 0;
    if (ok) ok = id_in_range_warp(id / W, local_rows, viol, lane);
    const float* src = local + (ok ? id / W : 0) * (int64_t)E;
    float* o = out + j * (int64_t)E;
    for (int c = lane; c < E4; c += 32) st4(o + c * 4, ok ? ldg4(src + c * 4) : f4_zero());
  }
}

}  // namespace shard
}  // namespace lk

using namespace lk;

extern "C" {

// hash capacity for P positions: next power of two >= 2P (load factor <= 0.5)
int64_t lk_shard_hash_slots(int64_t P) {
  int64_t s = 1024;
  while (s < 2 * P) s <<= 1;
  return s;
}

// keys [slots] int64 must be filled with -1, counts [W] and overflow [1] with 0, send_ids [W*cap] with -1 (cudaMemsetAsync 0xFF / 0) before the call
int lk_shard_plan(const int64_t* ids, int64_t P, int W, int64_t cap, int64_t* keys, int32_t* vals, int64_t slots, int32_t* counts,
                  int64_t* send_ids, int32_t* overflow, cudaStream_t st) {
  LK_REQUIRE(W >= 1 && cap >= 1 && slots >= 2 * P && (slots & (slots - 1)) == 0, LK_ERR_ARG, "lk_shard_plan: W=%d cap=%ld slots=%ld (power of two >= 2P)",
             W, (long)cap, (long)slots);
  LK_REQUIRE((int64_t)W * cap < ((int64_t)1 << 31), LK_ERR_ARG, "lk_shard_plan: W*cap overflows int32");
  if (P == 0) return LK_OK;
  int64_t blocks = (P + 255) / 256;
  if (blocks > (int64_t)kNumSMs * 8) blocks = (int64_t)kNumSMs * 8;
  LK_LAUNCH((shard::plan_kernel), (unsigned)blocks, 256, 0, st, ids, P, W, (int)cap, (long long*)keys, vals, (uint32_t)(slots - 1), counts, send_ids, overflow);
  return check_launch("shard_plan");
}

int lk_shard_inverse(const int64_t* ids, int64_t P, const int64_t* keys, const int32_t* vals, int64_t slots, int64_t* inverse, cudaStream_t st) {
  if (P == 0) return LK_OK;
  int64_t blocks = (P + 255) / 256;
  if (blocks > (int64_t)kNumSMs * 8) blocks = (int64_t)kNumSMs * 8;
  LK_LAUNCH((shard::inverse_kernel), (unsigned)blocks, 256, 0, st, ids, P, (const long long*)keys, vals, (uint32_t)(slots - 1), inverse);
  return check_launch("shard_inverse");
}

int lk_shard_gather(const int64_t* ids, int64_t M, int W, const float* local, int64_t local_rows, int64_t E, float* out, cudaStream_t st) {
  LK_REQUIRE(E % 4 == 0 && W >= 1, LK_ERR_SHAPE, "lk_shard_gather: E=%ld must be a multiple of 4", (long)E);
  if (M == 0) return LK_OK;
  int64_t blocks = (M + 7) / 8;
  if (blocks > (int64_t)kNumSMs * 32) blocks = (int64_t)kNumSMs * 32;
  LK_LAUNCH((shard::gather_kernel), (unsigned)blocks, 256, 0, st, ids, M, W, local, local_rows, (int)E, out, id_violations());
  return check_launch("shard_gather");
}

}  // extern "C"
