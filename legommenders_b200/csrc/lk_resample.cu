// Device-side Resampler: from B impression indices to the id lists and offsets of a packed training batch, one launch.
//
// Reference semantics (loader/resampler.py:139-259, restated in oracle/lego_oracle.py `candidates` / `pad_history`):
//   candidates of an impression = [positive, negatives...]  (label 0 = the positive, legommender.py:114-118)
//   negatives = random.sample(true_negs, min(K, len(true_negs)))            distinct POSITIONS of the user's negative list, random order
//             + (K - that) x random.randint(0, item_size - 1)               uniform item ids
//   history   = the user's clicks, right-padded with item 0 to H, `__clicks_mask__ = 1^len 0^pad`
// Re-design: the padding is never materialised (packed rows, see packing.py); the batch is described by
//   items[n]      B*C candidate ids ([pos, negs...] per impression) followed by the valid history items user by user
//   cu_items[n+1] token offsets (item_len = number of valid tokens of the item's ConcatInputer layout)
//   cu_users[B+1] history offsets
//   meta[4]       T = total token rows, n, longest item, longest history   (the host needs these four integers to size the step;
//                 they are read back one step AHEAD of use, batching.DeviceResampler)
// and `lk_pack_item_tokens` expands it to token ids.  Draws come from Philox4x32-10 keyed by (seed, impression row), counter = draw
// index: the batch is a pure function of (rows, seed), independent of launch geometry — tests replay it on the host.
#include "lk_common.cuh"
#include "../../include/legommenders_b200.h"

namespace lk {
namespace rs {

constexpr int RT = 1024;       // one CTA: a batch is a few thousand items

__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
  for (int r = 0; r < 10; r++) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// draw `i` of impression `row` under `seed`: uniform integer in [0, bound) (multiply-shift, bound < 2^32)
__host__ __device__ inline uint32_t draw(uint64_t seed, int64_t row, uint32_t i, uint32_t bound) {
  uint32_t o[4];
  philox4x32_10(i, 0u, (uint32_t)row, (uint32_t)((uint64_t)row >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), o);
  return (uint32_t)(((uint64_t)o[0] * bound) >> 32);
}

// candidates of one impression into cand[0..K]: the shared restatement used by the kernel and by lk_resample_reference (host)
__host__ __device__ inline void sample_candidates(uint64_t seed, int64_t row, int64_t pos, const int64_t* negs, int64_t n_negs, int K,
                                                  int64_t n_items, int64_t* cand) {
  cand[0] = pos;
  const int k = n_negs < K ? (int)n_negs : K;
  uint32_t d = 0;
  // Floyd: k distinct positions of [0, n_negs), then a Fisher-Yates pass for a uniformly random ORDER (random.sample returns one)
  int64_t chosen[LK_MAX_NEG];
  for (int j = 0; j < k; j++) {
    const int64_t top = n_negs - k + j;                         // candidate range [0, top]
    int64_t t = (int64_t)draw(seed, row, d++, (uint32_t)(top + 1));
    for (int q = 0; q < j; q++)
      if (chosen[q] == t) { t = top; break; }
    chosen[j] = t;
  }
  for (int j = k - 1; j > 0; j--) {
    const int r = (int)draw(seed, row, d++, (uint32_t)(j + 1));
    const int64_t tmp = chosen[j]; chosen[j] = chosen[r]; chosen[r] = tmp;
  }
  for (int j = 0; j < k; j++) cand[1 + j] = negs[chosen[j]];
  for (int j = k; j < K; j++) cand[1 + j] = (int64_t)draw(seed, row, d++, (uint32_t)n_items);
}

// block-wide exclusive scan of per-thread sums (RT threads); returns the total in every thread
__device__ __forceinline__ int block_exclusive_scan(int v, int& excl, int* warp_tot /* [32] */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  if (w == 0) {
    int t = warp_tot[lane], s = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += u;
    }
    warp_tot[lane] = s - t;                                     // exclusive warp offsets
    if (lane == 31) warp_tot[32] = s;                            // grand total
  }
  __syncthreads();
  excl = warp_tot[w] + inc - v;
  const int total = warp_tot[32];
  __syncthreads();
  return total;
}

__global__ void __launch_bounds__(RT) resample_kernel(const int64_t* __restrict__ rows, int B, int K, uint64_t seed,
                                                      const int64_t* __restrict__ imp_user, const int64_t* __restrict__ imp_pos,
                                                      const int64_t* __restrict__ neg_off, const int64_t* __restrict__ neg_items,
                                                      const int64_t* __restrict__ hist_off, const int64_t* __restrict__ hist_items,
                                                      const int32_t* __restrict__ item_len, int64_t n_items, int64_t n_imps, int64_t n_users,
                                                      int64_t* __restrict__ items, int32_t* __restrict__ cu_items, int32_t* __restrict__ cu_users,
                                                      int64_t* __restrict__ user_ids, int32_t* __restrict__ meta, int64_t items_cap,
                                                      int32_t* viol) {
  pdl_prologue();
  __shared__ int warp_tot[33];
  __shared__ int s_max[2];
  const int C = K + 1;
  const int tid = threadIdx.x;
  if (tid < 2) s_max[tid] = 0;
  // ---- phase 1: candidates + history lengths (thread per impression, strided) ------------------------------------------------------
  for (int b0 = 0; b0 < B; b0 += RT) {
    const int b = b0 + tid;
    int hl = 0;
    if (b < B) {
      int64_t row = rows[b];
      if (!id_in_range(row, n_imps, viol)) row = 0;
      int64_t u = imp_user[row];
      if (!id_in_range(u, n_users, viol)) u = 0;
      user_ids[b] = u;
      const int64_t o = neg_off[u];
      int64_t cand[LK_MAX_NEG + 1];
      sample_candidates(seed, row, imp_pos[row], neg_items + o, neg_off[u + 1] - o, K, n_items, cand);
      for (int j = 0; j < C; j++) items[(int64_t)b * C + j] = cand[j];
      hl = (int)(hist_off[u + 1] - hist_off[u]);
    }
    int excl;
    // (B <= RT in every configuration of BASELINE.json; larger batches chain the scan through a running base)
    const int total = block_exclusive_scan(hl, excl, warp_tot);
    const int base = b0 == 0 ? 0 : cu_users[b0];
    if (b < B) cu_users[b] = base + excl;
    if (tid == 0) cu_users[min(b0 + RT, B)] = base + total;
    atomicMax(&s_max[1], hl);
    __syncthreads();
  }
  const int n_hist = cu_users[B];
  const int64_t n = (int64_t)B * C + n_hist;
  if (n > items_cap) {               // caller's buffers are too small: report through meta, write nothing further
    if (tid == 0) { meta[0] = -1; meta[1] = (int32_t)n; meta[2] = meta[3] = 0; }
    return;
  }
  // ---- phase 2: valid history items user by user ------------------------------------------------------------------------------------
  for (int b = tid >> 5; b < B; b += RT / 32) {                   // warp per impression
    const int64_t u = user_ids[b];
    const int64_t ho = hist_off[u];
    const int o = cu_users[b], L = cu_users[b + 1] - o;
    for (int j = tid & 31; j < L; j += 32) items[(int64_t)B * C + o + j] = hist_items[ho + j];
  }
  __syncthreads();
  // ---- phase 3: token offsets -------------------------------------------------------------------------------------------------------------
  int run = 0;
  for (int64_t i0 = 0; i0 < n; i0 += RT) {
    const int64_t i = i0 + tid;
    int len = 0;
    if (i < n) {
      int64_t it = items[i];
      if (!id_in_range(it, n_items, viol)) { it = 0; items[i] = 0; }
      len = item_len[it];
      atomicMax(&s_max[0], len);
    }
    int excl;
    const int total = block_exclusive_scan(len, excl, warp_tot);
    if (i < n) cu_items[i] = run + excl;
    run += total;
  }
  __syncthreads();
  if (tid == 0) {
    cu_items[n] = run;
    meta[0] = run; meta[1] = (int32_t)n; meta[2] = s_max[0]; meta[3] = s_max[1];
  }
}

}  // namespace rs
}  // namespace lk

using namespace lk;

extern "C" {

int lk_resample_batch(const int64_t* rows, int64_t B, int K, uint64_t seed, const int64_t* imp_user, const int64_t* imp_pos,
                      const int64_t* neg_off, const int64_t* neg_items, const int64_t* hist_off, const int64_t* hist_items,
                      const int32_t* item_len, int64_t n_items, int64_t n_imps, int64_t n_users, int64_t* items, int32_t* cu_items,
                      int32_t* cu_users, int64_t* user_ids, int32_t* meta, int64_t items_cap, cudaStream_t st) {
  LK_REQUIRE(B > 0 && K >= 0 && K <= LK_MAX_NEG, LK_ERR_ARG, "lk_resample_batch: B=%ld, K=%d (at most %d negatives)", (long)B, K, LK_MAX_NEG);
  LK_REQUIRE(n_items > 0 && n_items < ((int64_t)1 << 32), LK_ERR_ARG, "lk_resample_batch: item vocabulary size %ld", (long)n_items);
  LK_REQUIRE(items_cap >= B * (K + 1), LK_ERR_ARG, "lk_resample_batch: item buffer smaller than the candidates alone");
  LK_LAUNCH((rs::resample_kernel), 1, rs::RT, 0, st, rows, (int)B, K, (unsigned long long)seed, imp_user, imp_pos, neg_off, neg_items, hist_off,
            hist_items, item_len, n_items, n_imps, n_users, items, cu_items, cu_users, user_ids, meta, items_cap, id_violations());
  return check_launch("resample_batch");
}

// Host restatement of the candidate draw for ONE impression (the same inline function the kernel runs): tests replay device batches with it.
int lk_resample_reference(uint64_t seed, int64_t row, int64_t pos, const int64_t* negs, int64_t n_negs, int K, int64_t n_items, int64_t* cand_out) {
  LK_REQUIRE(K >= 0 && K <= LK_MAX_NEG && n_items > 0, LK_ERR_ARG, "lk_resample_reference: bad arguments");
  rs::sample_candidates(seed, row, pos, negs, n_negs, K, n_items, cand_out);
  return LK_OK;
}

}  // extern "C"
