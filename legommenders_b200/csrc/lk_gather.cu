// EmbeddingHub token gather (north_star piece 1).
//
// Reference semantics (model/inputer/concat_inputer.py:105-113, simple_inputer.py:51-64,
// loader/embedding_hub.py:378-385, model/operators/pooling_operator.py:46-61):
//   valid(m) = mask ? mask[m] > 0 : ids[m] > -1 ;  row(m) = valid ? table[ids[m]] : 0
//   gather      : out[m,:]  (+)= row(m)
//   gather+pool : out[n,:]  = sum_t row(n,t) / (sum_t valid + 1e-8)   (mean) | max_t row(n,t) | sum_t row(n,t)
// Backward for trainable tables is a sorted-index segmented scatter-add (no atomics, deterministic):
//   radix sort (id, position) -> run-length encode -> fixed-size partial sums -> per-id ordered reduction.
#include <cub/cub.cuh>
#include <cuda_bf16.h>

#include "lk_common.cuh"
#include "../../include/legommenders_b200.h"

namespace lk {

constexpr int GW = 8;  // warps per block in the gather kernels

// ------------------------------------------------------------------------------------------------
// plain gather: one warp per row, 16-byte vector loads, fused validity mask
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GW * 32, 8) gather_rows_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ mask,
                                                              const float* __restrict__ table, int64_t V, int32_t* viol,
                                                              float* __restrict__ out, int64_t M, int E, int accumulate) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * GW + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * GW;
  const int E4 = E >> 2;
  for (int64_t m = warp; m < M; m += nwarps) {
    int64_t id = ids[m];
    bool valid = mask ? (mask[m] > 0) : (id > -1);
    if (valid) valid = id_in_range_warp(id, V, viol, lane);
    float* o = out + m * E;
    if (valid) {
      const float* src = table + id * (int64_t)E;
      for (int c = lane; c < E4; c += 32) {
        float4 v = ldg4(src + c * 4);
        if (accumulate) f4_add(v, *reinterpret_cast<const float4*>(o + c * 4));
        st4(o + c * 4, v);
      }
    } else if (!accumulate) {
      for (int c = lane; c < E4; c += 32) st4(o + c * 4, f4_zero());
    }
  }
}

// gather straight into split-bf16 planes (the A operand of the projection GEMM): row m of hi/lo = split(table[ids[m]]) or 0
__global__ void __launch_bounds__(GW * 32) gather_split_kernel(const int64_t* __restrict__ ids, const float* __restrict__ table,
                                                               int64_t V, int32_t* viol,
                                                               __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                               int64_t M, int E, int ld) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * GW + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * GW;
  const int L4 = ld >> 2;
  for (int64_t m = warp; m < M; m += nwarps) {
    int64_t id = ids[m];
    if (id > -1 && !id_in_range_warp(id, V, viol, lane)) id = -1;
    const float* src = table + id * (int64_t)E;
    for (int c = lane; c < L4; c += 32) {
      float4 v = f4_zero();
      if (id > -1 && c * 4 < E) v = ldg4(src + c * 4);
      __align__(8) __nv_bfloat16 h[4], l[4];
      const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; e++) {
        h[e] = __float2bfloat16_rn(x[e]);
        l[e] = __float2bfloat16_rn(x[e] - __bfloat162float(h[e]));
      }
      *reinterpret_cast<uint2*>(hi + m * ld + c * 4) = *reinterpret_cast<uint2*>(h);
      *reinterpret_cast<uint2*>(lo + m * ld + c * 4) = *reinterpret_cast<uint2*>(l);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// gather + masked pooling: one warp per (item, 128-float column block); 4 tokens in flight
// ------------------------------------------------------------------------------------------------
template <int MODE>  // 0 mean, 1 max, 2 sum
// 8 blocks per SM = 32 registers = every warp slot of the SM: the id bounds check had pushed the kernel to 36 registers (6 blocks, 48 of 64
// warps) and cost 7 % (4 M-row table) to 25 % (L2-resident table) of its throughput
__global__ void __launch_bounds__(GW * 32, 8) gather_pool_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ mask,
                                                              const float* __restrict__ table, int64_t V, int32_t* viol,
                                                              float* __restrict__ out, int64_t N, int S, int E) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t n = (int64_t)blockIdx.x * GW + (threadIdx.x >> 5);
  const int c = (blockIdx.y * 32 + lane) * 4;
  if (n >= N) return;
  const bool col_ok = c < E;
  const int64_t* idr = ids + n * S;
  const int64_t* mr = mask ? mask + n * S : nullptr;
  float4 acc = MODE == 1 ? make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY) : f4_zero();
  int cnt = 0, nbad = 0;
  for (int t0 = 0; t0 < S; t0 += 4) {
    float4 v[4];
    bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      int t = t0 + u;
      ok[u] = false;
      v[u] = f4_zero();
      if (t < S) {
        int64_t id = idr[t];
        ok[u] = mr ? (mr[t] > 0) : (id > -1);
        const bool oob = ok[u] && (uint64_t)id >= (uint64_t)V;   // out of range: an invalid position, counted once after the loop
        nbad += oob ? 1 : 0;                                     // (no branch or atomic between the four loads of a step)
        ok[u] = ok[u] && !oob;
        if (ok[u] && col_ok) v[u] = ldg4(table + id * (int64_t)E + c);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (t0 + u >= S) continue;
      cnt += ok[u] ? 1 : 0;
      if (MODE == 1) {
        acc.x = fmaxf(acc.x, v[u].x); acc.y = fmaxf(acc.y, v[u].y);
        acc.z = fmaxf(acc.z, v[u].z); acc.w = fmaxf(acc.w, v[u].w);
      } else {
        f4_add(acc, v[u]);
      }
    }
  }
  if (nbad && lane == 0 && blockIdx.y == 0 && viol) atomicAdd(viol, nbad);   // every lane saw the same ids: lane 0 of column block 0 reports
  if (!col_ok) return;
  if (MODE == 0) {
    float inv = 1.0f / ((float)cnt + 1e-8f);
    acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
  }
  st4(out + n * (int64_t)E + c, acc);
}

// ------------------------------------------------------------------------------------------------
// sorted-index segmented scatter-add
// ------------------------------------------------------------------------------------------------
constexpr int SEG = 32;  // rows per partial sum

__global__ void make_keys_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ mask, int* __restrict__ keys,
                                 int* __restrict__ vals, int64_t P, int V) {
  pdl_prologue();
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  int64_t id = ids[p];
  bool valid = mask ? (mask[p] > 0) : (id > -1);
  keys[p] = valid ? (int)id : V;   // invalid positions sort to the end under sentinel V
  vals[p] = (int)p;
}

__global__ void partial_counts_kernel(const int* __restrict__ run_len, const int* __restrict__ num_runs, int* __restrict__ npart,
                                      int cap) {
  pdl_prologue();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap) return;
  npart[i] = (i < *num_runs) ? (run_len[i] + SEG - 1) / SEG : 0;
}

// partial p of run r sums rows [run_off[r] + k*SEG, min(run_off[r+1], ...+SEG)) of the sorted order
__global__ void __launch_bounds__(GW * 32) partial_sums_kernel(const int* __restrict__ sorted_pos, const int* __restrict__ run_off,
                                                               const int* __restrict__ run_len, const int* __restrict__ part_off,
                                                               const int* __restrict__ num_runs, const float* __restrict__ src,
                                                               const float* __restrict__ scale, int row_div, int E,
                                                               float* __restrict__ partial, int max_partials) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * GW + (threadIdx.x >> 5);
  const int R = *num_runs;
  if (p >= max_partials) return;
  const int total = part_off[R - 1] + (run_len[R - 1] + SEG - 1) / SEG;
  if (p >= total) return;
  // binary search: last run r with part_off[r] <= p
  int lo = 0, hi = R - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (part_off[mid] <= p) lo = mid; else hi = mid - 1;
  }
  const int r = lo;
  const int beg = run_off[r] + (p - part_off[r]) * SEG;
  const int end = min(run_off[r] + run_len[r], beg + SEG);
  const int c = (blockIdx.y * 32 + lane) * 4;
  if (c >= E) return;
  float4 acc = f4_zero();
  for (int i = beg; i < end; i++) {
    int pos = sorted_pos[i];
    float4 v = ldg4(src + (int64_t)(pos / row_div) * E + c);
    if (scale) { float s = scale[pos]; v.x *= s; v.y *= s; v.z *= s; v.w *= s; }
    f4_add(acc, v);
  }
  st4(partial + (int64_t)p * E + c, acc);
}

__global__ void __launch_bounds__(GW * 32) run_reduce_kernel(const int* __restrict__ run_key, const int* __restrict__ run_len,
                                                             const int* __restrict__ part_off, const int* __restrict__ num_runs,
                                                             const float* __restrict__ partial, float* __restrict__ dtable, int V,
                                                             int E, int accumulate, int max_runs) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * GW + (threadIdx.x >> 5);
  if (r >= max_runs || r >= *num_runs) return;
  const int key = run_key[r];
  if (key >= V) return;  // sentinel run of invalid positions
  const int c = (blockIdx.y * 32 + lane) * 4;
  if (c >= E) return;
  const int np = (run_len[r] + SEG - 1) / SEG;
  const float* src = partial + (int64_t)part_off[r] * E + c;
  float4 acc = f4_zero();
  for (int i = 0; i < np; i++) f4_add(acc, ldg4_stream(src + (int64_t)i * E));
  float* o = dtable + (int64_t)key * E + c;
  if (accumulate) f4_add(acc, *reinterpret_cast<const float4*>(o));
  st4(o, acc);
}

// ------------------------------------------------------------------------------------------------
// small tables (V <= 64 rows: category / special-token tables): no sort.  A block owns SM_CHUNK consecutive positions;
// SM_LANES row-lanes walk them in order, each into its own shared [V][E] accumulator; the lanes are then summed in fixed
// order into partial[block][V][E] and a second kernel reduces the blocks in fixed order -> deterministic, no atomics.
// ------------------------------------------------------------------------------------------------
constexpr int SM_MAX_V = 64, SM_CHUNK = 512, SM_LANES = 4;

__global__ void __launch_bounds__(256) scatter_small_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ mask,
                                                            const float* __restrict__ src, const float* __restrict__ scale,
                                                            int row_div, int64_t P, int V, int E, float* __restrict__ partial) {
  pdl_prologue();
  extern __shared__ __align__(16) float acc_s[];   // [SM_LANES][V][E]
  const int E4 = E >> 2;
  const int cols = blockDim.x / SM_LANES;          // column threads per row-lane
  const int rl = threadIdx.x / cols, ct = threadIdx.x - rl * cols;
  float* mine = acc_s + (size_t)rl * V * E;
  for (int i = ct; i < V * E4; i += cols) reinterpret_cast<float4*>(mine)[i] = f4_zero();
  __syncthreads();
  const int64_t p0 = (int64_t)blockIdx.x * SM_CHUNK;
  const int64_t p1 = p0 + SM_CHUNK < P ? p0 + SM_CHUNK : P;
  for (int64_t p = p0 + rl; p < p1; p += SM_LANES) {
    const int64_t id = ids[p];
    const bool valid = mask ? (mask[p] > 0) : (id > -1);
    if (!valid) continue;
    const float sc = scale ? scale[p] : 1.f;
    const float* s = src + (p / row_div) * (int64_t)E;
    float* a = mine + (size_t)id * E;
    for (int c = ct; c < E4; c += cols) {
      float4 v = ldg4(s + c * 4);
      float4 o = *reinterpret_cast<float4*>(a + c * 4);
      f4_fma(o, sc, v);
      *reinterpret_cast<float4*>(a + c * 4) = o;
    }
  }
  __syncthreads();
  float* out = partial + (size_t)blockIdx.x * V * E;
  for (int i = threadIdx.x; i < V * E4; i += blockDim.x) {
    float4 t = reinterpret_cast<const float4*>(acc_s)[i];
#pragma unroll
    for (int l = 1; l < SM_LANES; l++) f4_add(t, reinterpret_cast<const float4*>(acc_s + (size_t)l * V * E)[i]);
    reinterpret_cast<float4*>(out)[i] = t;
  }
}

__global__ void scatter_small_finish_kernel(const float* __restrict__ partial, float* __restrict__ dtable, int nblk, int VE4,
                                            int accumulate) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= VE4) return;
  float4 t = accumulate ? reinterpret_cast<const float4*>(dtable)[i] : f4_zero();
  for (int b = 0; b < nblk; b++) f4_add(t, ldg4_stream(partial + ((size_t)b * VE4 + i) * 4));
  reinterpret_cast<float4*>(dtable)[i] = t;
}

static bool small_table(int64_t V, int64_t E) { return V <= SM_MAX_V && (size_t)SM_LANES * V * E * 4 <= 96 * 1024; }
static size_t small_ws_bytes(int64_t P, int64_t V, int64_t E) { return (size_t)((P + SM_CHUNK - 1) / SM_CHUNK) * V * E * 4 + 256; }

// ------------------------------------------------------------------------------------------------
// backward of the ConcatInputer embedding stage in ONE pass over dx (concat_inputer.py:105-113 + embedding_hub.py:95-96):
//   x[t] = valid(title[t]) · dropout(W·glove[title[t]] + b) + cat_table[cat[t]] + special_table[sp[t]]
//   => dP[t]   = dx[t] · dropout · valid     -> split-bf16 planes (operand of the projection's weight gradient) + column sums (db)
//      dcat[c] = Σ_{t: cat[t]=c} dx[t],   dspecial likewise              (deterministic: block partials, fixed-order finish)
// A block owns eb_rows(T) consecutive rows; EB_LANES row-lanes walk them in order with private shared accumulators.  The rows per block
// are chosen so that the grid is ONE balanced wave of the 2 blocks per SM the shared accumulators allow (128-row blocks at T = 42.6 k
// were 333 blocks on 296 slots: a full second wave for 37 blocks).
// ------------------------------------------------------------------------------------------------
constexpr int EB_MAX_ROWS = 512, EB_MIN_ROWS = 32, EB_LANES = 4;
static int eb_rows(int64_t T) {
  int64_t r = (T + 2 * kNumSMs - 1) / (2 * kNumSMs);
  r = (r + EB_LANES - 1) / EB_LANES * EB_LANES;
  return (int)(r < EB_MIN_ROWS ? EB_MIN_ROWS : r > EB_MAX_ROWS ? EB_MAX_ROWS : r);
}

__global__ void __launch_bounds__(256) concat_embed_bwd_kernel(const float* __restrict__ dx, const int64_t* __restrict__ title,
                                                               const int64_t* __restrict__ cat, const int64_t* __restrict__ special,
                                                               int64_t T, int D, int Vc, int Vs, float drop_p, unsigned long long seed,
                                                               __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int ld,
                                                               float* __restrict__ part, int EB_ROWS) {
  pdl_prologue();
  extern __shared__ __align__(16) float sm[];           // [EB_LANES][1 + Vc + Vs][D]
  __shared__ int s_title[EB_MAX_ROWS], s_cat[EB_MAX_ROWS], s_sp[EB_MAX_ROWS];
  const int D4 = D >> 2, NT = 1 + Vc + Vs;
  const int cols = blockDim.x / EB_LANES;
  const int rl = threadIdx.x / cols, ct = threadIdx.x - rl * cols;
  const int64_t r0 = (int64_t)blockIdx.x * EB_ROWS;
  const int nrows = (int)((T - r0) < EB_ROWS ? (T - r0) : EB_ROWS);
  for (int i = threadIdx.x; i < nrows; i += blockDim.x) {
    s_title[i] = title[r0 + i] > -1 ? 1 : 0;
    s_cat[i] = (int)cat[r0 + i];
    s_sp[i] = (int)special[r0 + i];
  }
  float* mine = sm + (size_t)rl * NT * D;
  for (int i = ct; i < NT * D4; i += cols) reinterpret_cast<float4*>(mine)[i] = f4_zero();
  __syncthreads();
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  // one column quad per thread (D <= 4 * cols): the next row's dx is requested before this row is processed, so the walk over the
  // rows is not a chain of exposed global-load latencies
  float4 nxt = f4_zero();
  if (rl < nrows && ct < D4) nxt = ldg4_stream(dx + (r0 + rl) * D + ct * 4);
  for (int i = rl; i < nrows; i += EB_LANES) {
    const int64_t row = r0 + i;
    const bool tv = s_title[i] != 0;
    const int ci = s_cat[i], si = s_sp[i];
    const float4 cur = nxt;
    if (i + EB_LANES < nrows && ct < D4) nxt = ldg4_stream(dx + (row + EB_LANES) * D + ct * 4);
    for (int c = ct; c < D4; c += cols) {
      const float4 g = cur;
      float4 m = f4_zero();
      if (tv) {
        m = g;
        if (drop_p > 0.f) {
          const uint64_t e = (uint64_t)row * D + c * 4;
          m.x *= dropout_scale(seed, e, drop_p, inv_keep); m.y *= dropout_scale(seed, e + 1, drop_p, inv_keep);
          m.z *= dropout_scale(seed, e + 2, drop_p, inv_keep); m.w *= dropout_scale(seed, e + 3, drop_p, inv_keep);
        }
        f4_add(*reinterpret_cast<float4*>(mine + c * 4), m);
      }
      __align__(8) __nv_bfloat16 h[4], l[4];
      const float xv[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
      for (int e = 0; e < 4; e++) {
        h[e] = __float2bfloat16_rn(xv[e]);
        l[e] = __float2bfloat16_rn(xv[e] - __bfloat162float(h[e]));
      }
      *reinterpret_cast<uint2*>(hi + row * ld + c * 4) = *reinterpret_cast<uint2*>(h);
      *reinterpret_cast<uint2*>(lo + row * ld + c * 4) = *reinterpret_cast<uint2*>(l);
      if (ci > -1) f4_add(*reinterpret_cast<float4*>(mine + (size_t)(1 + ci) * D + c * 4), g);
      if (si > -1) f4_add(*reinterpret_cast<float4*>(mine + (size_t)(1 + Vc + si) * D + c * 4), g);
    }
  }
  __syncthreads();
  float* out = part + (size_t)blockIdx.x * NT * D;
  for (int i = threadIdx.x; i < NT * D4; i += blockDim.x) {
    float4 t = reinterpret_cast<const float4*>(sm)[i];
#pragma unroll
    for (int l2 = 1; l2 < EB_LANES; l2++) f4_add(t, reinterpret_cast<const float4*>(sm + (size_t)l2 * NT * D)[i]);
    reinterpret_cast<float4*>(out)[i] = t;
  }
}

// out[c] = Σ_b part[b, c] in a fixed order: 32 part-lanes x 32 columns per block, then an ordered shared-memory reduction
__global__ void __launch_bounds__(1024) partial_finish_kernel(const float* __restrict__ part, int nparts, int64_t stride, int cols,
                                                              float* __restrict__ out) {
  pdl_prologue();
  __shared__ float red[32][33];
  const int cx = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float s = 0.f;
  if (c < cols)
    for (int i = pl; i < nparts; i += 32) s += part[(int64_t)i * stride + c];
  red[pl][cx] = s;
  __syncthreads();
  if (pl == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 32; i++) t += red[i][cx];
    out[c] = t;
  }
}

// ------------------------------------------------------------------------------------------------
// packed token ids of a list of items from device-resident per-item token tables (the Resampler's per-item cache,
// loader/resampler.py:113-126, kept on the device so that a training batch crosses PCIe as item ids, not as [B, 55, S] trees)
//   out_c[cu[n] + t] = table_c[item[n], t]   for t < cu[n+1] - cu[n]      (valid tokens are left-packed, concat_inputer.py:58-87)
// ------------------------------------------------------------------------------------------------
struct PackCols { const int64_t* table[4]; int64_t* out[4]; int ncols; };

__global__ void __launch_bounds__(256) pack_item_tokens_kernel(PackCols pc, const int64_t* __restrict__ items, const int* __restrict__ cu,
                                                               int64_t n, int S) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= n) return;
  const int64_t id = items[w];
  const int r0 = cu[w], L = cu[w + 1] - r0;
  for (int c = 0; c < pc.ncols; c++) {
    const int64_t* src = pc.table[c] + id * S;
    int64_t* dst = pc.out[c] + r0;
    for (int t = lane; t < L; t += 32) dst[t] = src[t];
  }
}

struct ScatterWs {
  int *keys, *vals, *skeys, *svals, *run_key, *run_len, *run_off, *npart, *part_off, *num_runs;
  float* partial;
  void* cub;
  size_t cub_bytes;
};

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t cub_bytes_for(int64_t P, int cap) {
  size_t a = 0, b = 0, c = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (int*)nullptr, (int*)nullptr, (int*)nullptr, (int*)nullptr, (int)P);
  cub::DeviceRunLengthEncode::Encode(nullptr, b, (int*)nullptr, (int*)nullptr, (int*)nullptr, (int*)nullptr, (int)P);
  cub::DeviceScan::ExclusiveSum(nullptr, c, (int*)nullptr, (int*)nullptr, cap);
  size_t m = a > b ? a : b;
  return align_up(m > c ? m : c);
}

static int run_cap(int64_t P, int64_t V) { return (int)((P < V + 1) ? P : V + 1); }
static int64_t partial_cap(int64_t P, int64_t V) { return (P + SEG - 1) / SEG + run_cap(P, V); }

static size_t carve(ScatterWs& w, void* base, int64_t P, int64_t V, int E) {
  int cap = run_cap(P, V);
  char* p = (char*)base;
  size_t off = 0;
  auto take = [&](size_t bytes) { void* r = p ? p + off : nullptr; off += align_up(bytes); return r; };
  w.keys = (int*)take(P * 4); w.vals = (int*)take(P * 4); w.skeys = (int*)take(P * 4); w.svals = (int*)take(P * 4);
  w.run_key = (int*)take((size_t)cap * 4); w.run_len = (int*)take((size_t)cap * 4); w.run_off = (int*)take((size_t)cap * 4);
  w.npart = (int*)take((size_t)cap * 4); w.part_off = (int*)take((size_t)cap * 4); w.num_runs = (int*)take(4);
  w.partial = (float*)take((size_t)partial_cap(P, V) * E * 4);
  w.cub_bytes = cub_bytes_for(P, cap);
  w.cub = take(w.cub_bytes);
  return off;
}

}  // namespace lk

using namespace lk;

extern "C" {

int lk_gather_rows(const int64_t* ids, const int64_t* mask, const float* table, int64_t V, float* out, int64_t M, int64_t E,
                   int accumulate, cudaStream_t st) {
  LK_REQUIRE(E % 4 == 0, LK_ERR_SHAPE, "lk_gather_rows: row width %ld must be a multiple of 4 floats (16-byte loads)", (long)E);
  if (M == 0) return LK_OK;
  int64_t blocks = (M + GW - 1) / GW;
  if (blocks > (int64_t)kNumSMs * 64) blocks = (int64_t)kNumSMs * 64;
  LK_LAUNCH((gather_rows_kernel), (unsigned)blocks, GW * 32, 0, st, ids, mask, table, V, id_violations(), out, M, (int)E, accumulate);
  return check_launch("gather_rows");
}

int lk_gather_split_bf16(const int64_t* ids, const float* table, int64_t V, void* hi, void* lo, int64_t M, int64_t E, int64_t ld,
                         cudaStream_t st) {
  LK_REQUIRE(E % 4 == 0 && ld % 8 == 0 && ld >= E, LK_ERR_SHAPE, "lk_gather_split_bf16: E=%ld must be a multiple of 4, ld=%ld of 8", (long)E, (long)ld);
  if (M == 0) return LK_OK;
  int64_t blocks = (M + GW - 1) / GW;
  if (blocks > (int64_t)kNumSMs * 64) blocks = (int64_t)kNumSMs * 64;
  LK_LAUNCH((gather_split_kernel), (unsigned)blocks, GW * 32, 0, st, ids, table, V, id_violations(), (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, M, (int)E,
            (int)ld);
  return check_launch("gather_split");
}

int lk_gather_pool(const int64_t* ids, const int64_t* mask, const float* table, int64_t V, float* out, int64_t N, int64_t S, int64_t E,
                   int mode, cudaStream_t st) {
  LK_REQUIRE(E % 4 == 0, LK_ERR_SHAPE, "lk_gather_pool: row width %ld must be a multiple of 4 floats", (long)E);
  LK_REQUIRE(mode >= 0 && mode <= 2, LK_ERR_ARG, "lk_gather_pool: mode must be 0 (mean), 1 (max) or 2 (sum)");
  if (N == 0) return LK_OK;
  dim3 grid((unsigned)((N + GW - 1) / GW), (unsigned)((E + 127) / 128));
  if (mode == 0) LK_LAUNCH((gather_pool_kernel<0>), grid, GW * 32, 0, st, ids, mask, table, V, id_violations(), out, N, (int)S, (int)E);
  else if (mode == 1) LK_LAUNCH((gather_pool_kernel<1>), grid, GW * 32, 0, st, ids, mask, table, V, id_violations(), out, N, (int)S, (int)E);
  else LK_LAUNCH((gather_pool_kernel<2>), grid, GW * 32, 0, st, ids, mask, table, V, id_violations(), out, N, (int)S, (int)E);
  return check_launch("gather_pool");
}

int lk_pack_item_tokens(const int64_t* const* tables, int64_t* const* outs, int ncols, const int64_t* items, const int32_t* cu, int64_t n,
                        int64_t S, cudaStream_t st) {
  LK_REQUIRE(ncols >= 1 && ncols <= 4, LK_ERR_ARG, "lk_pack_item_tokens: 1..4 token columns");
  if (n == 0) return LK_OK;
  PackCols pc;
  pc.ncols = ncols;
  for (int c = 0; c < 4; c++) { pc.table[c] = c < ncols ? tables[c] : nullptr; pc.out[c] = c < ncols ? outs[c] : nullptr; }
  LK_LAUNCH((pack_item_tokens_kernel), (unsigned)((n + 7) / 8), 256, 0, st, pc, items, cu, n, (int)S);
  return check_launch("pack_item_tokens");
}

int64_t lk_concat_embed_bwd_blocks(int64_t T) { return T > 0 ? (T + eb_rows(T) - 1) / eb_rows(T) : 0; }

size_t lk_concat_embed_bwd_workspace_bytes(int64_t T, int64_t D, int64_t n_cats, int64_t n_special) {
  return (size_t)lk_concat_embed_bwd_blocks(T) * (1 + n_cats + n_special) * D * sizeof(float) + 256;
}

int lk_concat_embed_bwd(const float* dx, const int64_t* title_ids, const int64_t* cat_ids, const int64_t* special_ids, int64_t T,
                        int64_t D, int64_t n_cats, int64_t n_special, float drop_p, uint64_t seed, void* dp_hi, void* dp_lo, int64_t ld,
                        float* g_bias, float* g_cat, float* g_special, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const int NT = (int)(1 + n_cats + n_special);
  const size_t smem = (size_t)EB_LANES * NT * D * sizeof(float);
  LK_REQUIRE(D % 4 == 0 && ld % 8 == 0 && ld >= D && D / 4 <= 64, LK_ERR_SHAPE, "lk_concat_embed_bwd: D=%ld must be a multiple of 4 and <= 256, ld a multiple of 8", (long)D);
  LK_REQUIRE(smem <= 200 * 1024, LK_ERR_SHAPE, "lk_concat_embed_bwd: tables with %d rows do not fit the shared accumulators", NT - 1);
  LK_REQUIRE(workspace && workspace_bytes >= lk_concat_embed_bwd_workspace_bytes(T, D, n_cats, n_special), LK_ERR_ARG,
             "lk_concat_embed_bwd: workspace too small");
  const int nblk = (int)lk_concat_embed_bwd_blocks(T);
  if (T > 0) {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(concat_embed_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr = true; }
    LK_LAUNCH((concat_embed_bwd_kernel), nblk, 256, smem, st, dx, title_ids, cat_ids, special_ids, T, (int)D, (int)n_cats, (int)n_special, drop_p,
                                                     (unsigned long long)seed, (__nv_bfloat16*)dp_hi, (__nv_bfloat16*)dp_lo, (int)ld,
                                                     (float*)workspace, eb_rows(T));
  }
  if (!g_bias && !g_cat && !g_special) return check_launch("concat_embed_bwd");   // deferred: the caller reduces workspace[nblk, NT*D] itself
  const float* part = (const float*)workspace;
  const int64_t stride = (int64_t)NT * D;
  LK_LAUNCH((partial_finish_kernel), (unsigned)((D + 31) / 32), 1024, 0, st, part, nblk, stride, (int)D, g_bias);
  LK_LAUNCH((partial_finish_kernel), (unsigned)((n_cats * D + 31) / 32), 1024, 0, st, part + D, nblk, stride, (int)(n_cats * D), g_cat);
  LK_LAUNCH((partial_finish_kernel), (unsigned)((n_special * D + 31) / 32), 1024, 0, st, part + (1 + n_cats) * D, nblk, stride, (int)(n_special * D),
                                                                                g_special);
  return check_launch("concat_embed_bwd", 4);
}

size_t lk_scatter_add_workspace_bytes(int64_t P, int64_t V, int64_t E) {
  if (small_table(V, E)) return small_ws_bytes(P, V, E);
  ScatterWs w;
  return carve(w, nullptr, P, V, (int)E) + 256;
}

int lk_scatter_add_sorted(const int64_t* ids, const int64_t* mask, const float* src, const float* scale, int64_t row_div,
                          float* dtable, int64_t P, int64_t V, int64_t E, int accumulate, void* workspace,
                          size_t workspace_bytes, cudaStream_t st) {
  LK_REQUIRE(E % 4 == 0, LK_ERR_SHAPE, "lk_scatter_add_sorted: row width %ld must be a multiple of 4 floats", (long)E);
  LK_REQUIRE(P < (1LL << 31) && V < (1LL << 30), LK_ERR_SHAPE, "lk_scatter_add_sorted: P or V too large for 32-bit keys");
  LK_REQUIRE(workspace_bytes >= lk_scatter_add_workspace_bytes(P, V, E), LK_ERR_ARG, "lk_scatter_add_sorted: workspace too small");
  if (small_table(V, E) && P > 0) {
    const int nblk = (int)((P + SM_CHUNK - 1) / SM_CHUNK);
    const size_t smem = (size_t)SM_LANES * V * E * 4;
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(scatter_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); attr = true; }
    LK_LAUNCH((scatter_small_kernel), nblk, 256, smem, st, ids, mask, src, scale, (int)row_div, P, (int)V, (int)E, (float*)workspace);
    const int VE4 = (int)(V * E / 4);
    LK_LAUNCH((scatter_small_finish_kernel), (VE4 + 127) / 128, 128, 0, st, (const float*)workspace, dtable, nblk, VE4, accumulate);
    return check_launch("scatter_add_small", 2);
  }
  if (!accumulate) cudaMemsetAsync(dtable, 0, (size_t)V * E * sizeof(float), st);
  if (P == 0) return LK_OK;
  ScatterWs w;
  carve(w, workspace, P, V, (int)E);
  const int cap = run_cap(P, V);
  LK_LAUNCH((make_keys_kernel), (unsigned)((P + 255) / 256), 256, 0, st, ids, mask, w.keys, w.vals, P, (int)V);
  int end_bit = 1;
  while ((1LL << end_bit) <= V) end_bit++;
  size_t cb = w.cub_bytes;
  cub::DeviceRadixSort::SortPairs(w.cub, cb, w.keys, w.skeys, w.vals, w.svals, (int)P, 0, end_bit, st);
  cb = w.cub_bytes;
  cub::DeviceRunLengthEncode::Encode(w.cub, cb, w.skeys, w.run_key, w.run_len, w.num_runs, (int)P, st);
  cb = w.cub_bytes;
  cub::DeviceScan::ExclusiveSum(w.cub, cb, w.run_len, w.run_off, cap, st);
  LK_LAUNCH((partial_counts_kernel), (cap + 255) / 256, 256, 0, st, w.run_len, w.num_runs, w.npart, cap);
  cb = w.cub_bytes;
  cub::DeviceScan::ExclusiveSum(w.cub, cb, w.npart, w.part_off, cap, st);
  const int maxp = (int)partial_cap(P, V);
  dim3 g1((maxp + GW - 1) / GW, (unsigned)((E + 127) / 128));
  LK_LAUNCH((partial_sums_kernel), g1, GW * 32, 0, st, w.svals, w.run_off, w.run_len, w.part_off, w.num_runs, src, scale, (int)row_div, (int)E,
                                            w.partial, maxp);
  dim3 g2((cap + GW - 1) / GW, (unsigned)((E + 127) / 128));
  LK_LAUNCH((run_reduce_kernel), g2, GW * 32, 0, st, w.run_key, w.run_len, w.part_off, w.num_runs, w.partial, dtable, (int)V, (int)E, 1, cap);
  return check_launch("scatter_add_sorted", 4);
}

}  // extern "C"
