// Group metrics of the evaluation phase on the device (reference utils/metrics.py:88-160, 223-235, 313-369).
//
// MetricPool.calculate builds a pandas DataFrame, groups by `groups` (sorted keys, rows of a group in original order) and, for
// every group metric, forks a process pool that calls sklearn per group — ~150 s for the 2.66 M-row MIND-small test set.
// Here: one stable radix sort by group key, one run-length encode, and ONE kernel with a warp per group that derives
//   GAUC   = Mann-Whitney U with tie-averaged ranks  (= sklearn.roc_auc_score for binary labels)
//   MRR    = sum_i y_i / rank_i / sum_i y_i, rank from the STABLE descending sort (python sorted(reverse=True))
//   nDCG@k = sklearn.ndcg_score: linear gain, 1/log2(p+2) discount truncated at k, DCG averaged over score ties
// from exact integer pair counts (for row i: #scores greater, #equal, #equal-and-earlier, #negatives lower / equal), so no
// in-group sort is needed; group sizes are tens to hundreds of rows and the O(n^2) compares run out of L1.
// Floating point is fp64 from the integer counts on; per-group values are rounded to fp32 exactly like the reference's
// `torch.tensor(values, dtype=torch.float)` before the mean (accumulated in fp64, fixed order).
#include <cub/cub.cuh>

#include "lk_common.cuh"
#include "../../include/legommenders_b200.h"

namespace lk {

constexpr int MAX_K = 8;

struct MetricWs {
  long long *keys, *skeys, *run_key;
  int *vals, *svals, *run_len, *run_off, *num_runs;
  float* s;        // scores in group-sorted order
  float* y;        // labels (as float) in group-sorted order
  void* cub;
  size_t cub_bytes;
};

static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t metric_cub_bytes(int64_t R) {
  size_t a = 0, b = 0, c = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (long long*)nullptr, (long long*)nullptr, (int*)nullptr, (int*)nullptr, (int)R);
  cub::DeviceRunLengthEncode::Encode(nullptr, b, (long long*)nullptr, (long long*)nullptr, (int*)nullptr, (int*)nullptr, (int)R);
  cub::DeviceScan::ExclusiveSum(nullptr, c, (int*)nullptr, (int*)nullptr, (int)R);
  size_t m = a > b ? a : b;
  return al256(m > c ? m : c);
}

static size_t metric_carve(MetricWs& w, void* base, int64_t R) {
  char* p = (char*)base;
  size_t off = 0;
  auto take = [&](size_t bytes) { void* r = p ? p + off : nullptr; off += al256(bytes); return r; };
  w.keys = (long long*)take(R * 8); w.skeys = (long long*)take(R * 8); w.run_key = (long long*)take(R * 8);
  w.vals = (int*)take(R * 4); w.svals = (int*)take(R * 4); w.run_len = (int*)take(R * 4); w.run_off = (int*)take(R * 4);
  w.num_runs = (int*)take(4);
  w.s = (float*)take(R * 4); w.y = (float*)take(R * 4);
  w.cub_bytes = metric_cub_bytes(R);
  w.cub = take(w.cub_bytes);
  return off;
}

__global__ void metric_keys_kernel(const int64_t* __restrict__ groups, long long* __restrict__ keys, int* __restrict__ vals, int64_t R) {
  pdl_prologue();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R) return;
  keys[i] = (long long)groups[i];   // cub's radix sort orders signed keys correctly (ascending, like pandas groupby)
  vals[i] = (int)i;
}

__global__ void metric_gather_kernel(const int* __restrict__ svals, const float* __restrict__ scores, const int64_t* __restrict__ labels,
                                     float* __restrict__ s, float* __restrict__ y, int64_t R) {
  pdl_prologue();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R) return;
  const int r = svals[i];
  s[i] = scores[r];
  y[i] = (float)labels[r];
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct KList { int k[MAX_K]; int nk; };

// per_group layout: [2 + nk][cap] (row 0 GAUC, 1 MRR, 2.. nDCG@k)
__global__ void __launch_bounds__(256) group_metrics_kernel(const float* __restrict__ s, const float* __restrict__ y,
                                                            const int* __restrict__ run_off, const int* __restrict__ run_len,
                                                            const int* __restrict__ num_runs, const double* __restrict__ dcs,
                                                            int n_disc, KList ks, float* __restrict__ per_group, int64_t cap) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= *num_runs) return;
  const int off = run_off[g], n = run_len[g];
  const float* sg = s + off;
  const float* yg = y + off;
  long long u2 = 0;            // 2 * Mann-Whitney U (integers: ties count 1, wins count 2)
  double rr = 0.0, ysum = 0.0, npos = 0.0;
  double dcg[MAX_K];
#pragma unroll
  for (int q = 0; q < MAX_K; q++) dcg[q] = 0.0;
  for (int i = lane; i < n; i += 32) {
    const float si = sg[i], yi = yg[i];
    int gt = 0, eq = 0, eqb = 0, nlt = 0, neq = 0;
    for (int j = 0; j < n; j++) {
      const float sj = sg[j];
      const bool neg = yg[j] <= 0.f;
      gt += sj > si;
      const bool e = sj == si;
      eq += e;
      eqb += e && (j < i);
      nlt += neg && (sj < si);
      neq += neg && e;
    }
    if (yi > 0.f) {
      u2 += 2LL * nlt + neq;
      npos += 1.0;
    }
    if (yi != 0.f) {
      rr += (double)yi / (double)(1 + gt + eqb);
      ysum += (double)yi;
#pragma unroll
      for (int q = 0; q < MAX_K; q++) {
        if (q < ks.nk) {
          const int k = ks.k[q] < n_disc - 1 ? ks.k[q] : n_disc - 1;
          const int hi = gt + eq < k ? gt + eq : k, lo = gt < k ? gt : k;
          dcg[q] += (double)yi / (double)eq * (dcs[hi] - dcs[lo]);
        }
      }
    }
  }
  u2 = warp_sum_ll(u2);
  rr = warp_sum_d(rr);
  ysum = warp_sum_d(ysum);
  npos = warp_sum_d(npos);
#pragma unroll
  for (int q = 0; q < MAX_K; q++)
    if (q < ks.nk) dcg[q] = warp_sum_d(dcg[q]);
  if (lane != 0) return;
  const double nneg = (double)n - npos;
  // single-class group: roc_auc_score is undefined (sklearn 1.5 raises, >=1.8 returns nan) -> nan
  per_group[0 * cap + g] = (npos > 0.0 && nneg > 0.0) ? (float)(0.5 * (double)u2 / (npos * nneg)) : __int_as_float(0x7fc00000);
  per_group[1 * cap + g] = ysum != 0.0 ? (float)(rr / ysum) : __int_as_float(0x7fc00000);
  const int ip = (int)npos;    // ideal DCG: the positives first (binary relevance)
#pragma unroll
  for (int q = 0; q < MAX_K; q++) {
    if (q < ks.nk) {
      const int k = ks.k[q] < n_disc - 1 ? ks.k[q] : n_disc - 1;
      const double ideal = dcs[ip < k ? ip : k];
      per_group[(2 + q) * cap + g] = ideal > 0.0 ? (float)(dcg[q] / ideal) : 0.f;
    }
  }
}

// out[m] = mean over groups of per_group[m][:] (fp64 accumulation in a fixed order); out[nm] = number of groups
__global__ void __launch_bounds__(256) metric_mean_kernel(const float* __restrict__ per_group, const int* __restrict__ num_runs,
                                                          int64_t cap, int nm, double* __restrict__ out) {
  pdl_prologue();
  __shared__ double sh[256];
  const int m = blockIdx.x;
  const int G = *num_runs;
  double acc = 0.0;
  for (int i = threadIdx.x; i < G; i += 256) acc += (double)per_group[(int64_t)m * cap + i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[m] = G > 0 ? sh[0] / (double)G : 0.0;
    if (m == 0) out[nm] = (double)G;
  }
}

}  // namespace lk

using namespace lk;

extern "C" {

size_t lk_group_metrics_workspace_bytes(int64_t R, int nk) {
  MetricWs w;
  return metric_carve(w, nullptr, R) + al256((size_t)(2 + nk) * R * 4) + 256;
}

int lk_group_metrics(const float* scores, const int64_t* labels, const int64_t* groups, int64_t R, const int32_t* ks, int nk,
                     const double* disc_prefix, int64_t n_disc, double* out, float* per_group, void* workspace,
                     size_t workspace_bytes, cudaStream_t st) {
  LK_REQUIRE(nk >= 0 && nk <= MAX_K, LK_ERR_ARG, "lk_group_metrics: at most %d nDCG cut-offs", MAX_K);
  LK_REQUIRE(R > 0 && R < (1LL << 31), LK_ERR_SHAPE, "lk_group_metrics: row count %ld out of range", (long)R);
  LK_REQUIRE(n_disc >= 2, LK_ERR_ARG, "lk_group_metrics: discount prefix table too short");
  LK_REQUIRE(workspace && workspace_bytes >= lk_group_metrics_workspace_bytes(R, nk), LK_ERR_ARG, "lk_group_metrics: workspace too small");
  MetricWs w;
  size_t used = metric_carve(w, workspace, R);
  float* pg = per_group ? per_group : (float*)((char*)workspace + used);
  KList kl;
  kl.nk = nk;
  for (int i = 0; i < MAX_K; i++) kl.k[i] = i < nk ? ks[i] : 0;
  const unsigned nb = (unsigned)((R + 255) / 256);
  LK_LAUNCH((metric_keys_kernel), nb, 256, 0, st, groups, w.keys, w.vals, R);
  size_t cb = w.cub_bytes;
  cub::DeviceRadixSort::SortPairs(w.cub, cb, w.keys, w.skeys, w.vals, w.svals, (int)R, 0, 64, st);   // LSD radix sort: stable
  LK_LAUNCH((metric_gather_kernel), nb, 256, 0, st, w.svals, scores, labels, w.s, w.y, R);
  cb = w.cub_bytes;
  cub::DeviceRunLengthEncode::Encode(w.cub, cb, w.skeys, w.run_key, w.run_len, w.num_runs, (int)R, st);
  cb = w.cub_bytes;
  cub::DeviceScan::ExclusiveSum(w.cub, cb, w.run_len, w.run_off, (int)R, st);
  LK_LAUNCH((group_metrics_kernel), (unsigned)((R + 7) / 8), 256, 0, st, w.s, w.y, w.run_off, w.run_len, w.num_runs, disc_prefix, (int)n_disc, kl, pg, R);
  LK_LAUNCH((metric_mean_kernel), 2 + nk, 256, 0, st, pg, w.num_runs, R, 2 + nk, out);
  return check_launch("group_metrics", 5);
}

}  // extern "C"
