// DotPredictor scoring fused with the loss (north_star piece 4), cached-evaluation scoring (piece 5),
// and the fused Adam update that closes a training step.
//
// Reference: model/legommender.py:268-290 + model/predictors/dot_predictor.py:7-10 (scores),
// model/legommender.py:114-118,254,263 (CrossEntropyLoss with label 0 / BCEWithLogitsLoss, mean),
// model/legommender.py:153-157,202-203 (cache indexing), base_lego.py:198-204 (torch.optim.Adam defaults).
#include "lk_common.cuh"
#include "../../include/legommenders_b200.h"

namespace lk {

constexpr int SW = 8;  // warps per block

__device__ __forceinline__ float warp_dot(const float* __restrict__ a, const float* __restrict__ b, int D, int lane) {
  float s = 0.f;
  for (int c = lane * 4; c < D; c += 128) s += f4_dot(ldg4(a + c), ldg4(b + c));
  return warp_sum(s);
}

// scores[b,c] = <u[b], v[b,c]>; optional softmax-CE with label 0: probs, rowloss
__global__ void __launch_bounds__(SW * 32) dot_ce_fwd_kernel(const float* __restrict__ U, const float* __restrict__ V,
                                                             float* __restrict__ scores, float* __restrict__ probs,
                                                             float* __restrict__ rowloss, int64_t B, int C, int D) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * SW + (threadIdx.x >> 5);
  if (b >= B) return;
  const float* u = U + b * (int64_t)D;
  float mx = -INFINITY;
  for (int c = 0; c < C; c++) {
    float s = warp_dot(u, V + (b * C + c) * (int64_t)D, D, lane);
    if (lane == 0) scores[b * C + c] = s;
    mx = fmaxf(mx, s);
  }
  if (!probs) return;
  __syncwarp();
  float sum = 0.f;
  for (int c = 0; c < C; c++) sum += expf(scores[b * C + c] - mx);
  const float lse = mx + logf(sum);
  if (lane == 0) {
    for (int c = 0; c < C; c++) probs[b * C + c] = expf(scores[b * C + c] - lse);
    rowloss[b] = lse - scores[b * C];
  }
}

// ranking mode: z[b] = <u[b], v[b]>, rowloss = softplus(z) - y z  (BCEWithLogitsLoss, numerically stable form)
__global__ void __launch_bounds__(SW * 32) dot_bce_fwd_kernel(const float* __restrict__ U, const float* __restrict__ V,
                                                              const float* __restrict__ y, float* __restrict__ scores,
                                                              float* __restrict__ rowloss, int64_t B, int D) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * SW + (threadIdx.x >> 5);
  if (b >= B) return;
  float z = warp_dot(U + b * (int64_t)D, V + b * (int64_t)D, D, lane);
  if (lane == 0) {
    scores[b] = z;
    if (rowloss) rowloss[b] = fmaxf(z, 0.f) - z * y[b] + log1pf(expf(-fabsf(z)));
  }
}

// deterministic mean of a vector by one CTA
__global__ void mean_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n) {
  pdl_prologue();
  __shared__ float sh[256];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 256) s += x[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0] / (float)n;
}

// g[b,c] = (probs[b,c] - [c==0]) * dloss / B ; dU[b] = sum_c g v[b,c] ; dV[b,c] = g u[b]
__global__ void __launch_bounds__(SW * 32) dot_ce_bwd_kernel(const float* __restrict__ U, const float* __restrict__ V,
                                                             const float* __restrict__ probs, const float* __restrict__ dloss,
                                                             float* __restrict__ dU, float* __restrict__ dV, int64_t B, int C,
                                                             int D) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * SW + (threadIdx.x >> 5);
  if (b >= B) return;
  const float gs = (dloss ? dloss[0] : 1.f) / (float)B;
  const float* u = U + b * (int64_t)D;
  for (int d = lane * 4; d < D; d += 128) {
    const float4 uv = ldg4(u + d);
    float4 acc = f4_zero();
    for (int c = 0; c < C; c++) {
      float g = (probs[b * C + c] - (c == 0 ? 1.f : 0.f)) * gs;
      f4_fma(acc, g, ldg4(V + (b * C + c) * (int64_t)D + d));
      st4(dV + (b * C + c) * (int64_t)D + d, make_float4(g * uv.x, g * uv.y, g * uv.z, g * uv.w));
    }
    st4(dU + b * (int64_t)D + d, acc);
  }
}

// generic: given dscores[b,c], dU[b] = sum_c ds v[b,c]; dV[b,c] = ds u[b]   (used for BCE and raw-score backward)
__global__ void __launch_bounds__(SW * 32) dot_bwd_kernel(const float* __restrict__ U, const float* __restrict__ V,
                                                          const float* __restrict__ dS, float* __restrict__ dU,
                                                          float* __restrict__ dV, int64_t B, int C, int D) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * SW + (threadIdx.x >> 5);
  if (b >= B) return;
  const float* u = U + b * (int64_t)D;
  for (int d = lane * 4; d < D; d += 128) {
    const float4 uv = ldg4(u + d);
    float4 acc = f4_zero();
    for (int c = 0; c < C; c++) {
      float g = dS[b * C + c];
      f4_fma(acc, g, ldg4(V + (b * C + c) * (int64_t)D + d));
      st4(dV + (b * C + c) * (int64_t)D + d, make_float4(g * uv.x, g * uv.y, g * uv.z, g * uv.w));
    }
    st4(dU + b * (int64_t)D + d, acc);
  }
}

__global__ void bce_dscore_kernel(const float* __restrict__ z, const float* __restrict__ y, const float* __restrict__ dloss,
                                  float* __restrict__ dz, int64_t B) {
  pdl_prologue();
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float sig = 1.f / (1.f + expf(-z[b]));
  dz[b] = (sig - y[b]) * dloss[0] / (float)B;
}

// cached evaluation: out[r] = <U[uid[r]], I[iid[r]]>; one warp per row, two rows in flight per warp
__global__ void __launch_bounds__(SW * 32) cached_scores_kernel(const float* __restrict__ U, const float* __restrict__ I,
                                                                const int64_t* __restrict__ uid, const int64_t* __restrict__ iid,
                                                                float* __restrict__ out, int64_t R, int D, int64_t n_users,
                                                                int64_t n_items, int32_t* viol) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * SW + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * SW;
  for (int64_t r = warp * 2; r < R; r += nwarps * 2) {
    const bool two = r + 1 < R;
    int64_t ua = uid[r], ia = iid[r], ub = two ? uid[r + 1] : ua, ib = two ? iid[r + 1] : ia;
    // out-of-range ids (the reference raises IndexError): counted, scored as 0, never dereferenced
    const bool ok0 = id_in_range_warp(ua, n_users, viol, lane) & id_in_range_warp(ia, n_items, viol, lane);
    const bool ok1 = !two || (id_in_range_warp(ub, n_users, viol, lane) & id_in_range_warp(ib, n_items, viol, lane));
    if (!ok0) ua = ia = 0;
    if (!ok1) ub = ib = 0;
    const float* u0 = U + ua * (int64_t)D;
    const float* v0 = I + ia * (int64_t)D;
    const float* u1 = U + ub * (int64_t)D;
    const float* v1 = I + ib * (int64_t)D;
    float s0 = 0.f, s1 = 0.f;
    for (int c = lane * 4; c < D; c += 128) {
      float4 a0 = ldg4(u0 + c), b0 = ldg4(v0 + c), a1 = ldg4(u1 + c), b1 = ldg4(v1 + c);
      s0 += f4_dot(a0, b0);
      s1 += f4_dot(a1, b1);
    }
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
    if (lane == 0) {
      out[r] = ok0 ? s0 : 0.f;
      if (two) out[r + 1] = ok1 ? s1 : 0.f;
    }
  }
}

// out[r,:] = table[ids[r],:]  (cache indexing model/legommender.py:153-157; ids are always valid)
__global__ void __launch_bounds__(SW * 32) index_rows_kernel(const float* __restrict__ table, const int64_t* __restrict__ ids,
                                                             float* __restrict__ out, int64_t R, int D, int64_t V, int32_t* viol) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * SW + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * SW;
  for (int64_t r = warp; r < R; r += nwarps) {
    const int64_t id = ids[r];
    const bool ok = id_in_range_warp(id, V, viol, lane);
    const float* src = table + (ok ? id : 0) * (int64_t)D;
    for (int c = lane * 4; c < D; c += 128) st4(out + r * (int64_t)D + c, ok ? ldg4(src + c) : f4_zero());
  }
}

__global__ void fill_kernel(float* __restrict__ p, float v, int64_t n) {
  pdl_prologue();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            int64_t n, float lr_over_bc1, float b1, float b2, float eps, float inv_sqrt_bc2, float grad_scale) {
  pdl_prologue();
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  if (i + 3 < n) {
    float4 gv = *reinterpret_cast<const float4*>(g + i);
    float4 mv = *reinterpret_cast<const float4*>(m + i);
    float4 vv = *reinterpret_cast<const float4*>(v + i);
    float4 pv = *reinterpret_cast<const float4*>(p + i);
    float ge[4] = {gv.x * grad_scale, gv.y * grad_scale, gv.z * grad_scale, gv.w * grad_scale};
    float me[4] = {mv.x, mv.y, mv.z, mv.w}, ve[4] = {vv.x, vv.y, vv.z, vv.w}, pe[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
    for (int q = 0; q < 4; q++) {
      me[q] = b1 * me[q] + (1.f - b1) * ge[q];
      ve[q] = b2 * ve[q] + (1.f - b2) * ge[q] * ge[q];
      pe[q] -= lr_over_bc1 * me[q] / (sqrtf(ve[q]) * inv_sqrt_bc2 + eps);
    }
    st4(m + i, make_float4(me[0], me[1], me[2], me[3]));
    st4(v + i, make_float4(ve[0], ve[1], ve[2], ve[3]));
    st4(p + i, make_float4(pe[0], pe[1], pe[2], pe[3]));
  } else {
    for (int64_t k = i; k < n; k++) {
      float ge = g[k] * grad_scale;
      float me = b1 * m[k] + (1.f - b1) * ge;
      float ve = b2 * v[k] + (1.f - b2) * ge * ge;
      m[k] = me; v[k] = ve;
      p[k] -= lr_over_bc1 * me / (sqrtf(ve) * inv_sqrt_bc2 + eps);
    }
  }
}

}  // namespace lk

using namespace lk;

extern "C" {

int lk_dot_scores(const float* U, const float* V, float* scores, int64_t B, int64_t C, int64_t D, cudaStream_t st) {
  LK_REQUIRE(D % 4 == 0, LK_ERR_SHAPE, "lk_dot_scores: D=%ld must be a multiple of 4", (long)D);
  if (B == 0) return LK_OK;
  LK_LAUNCH((dot_ce_fwd_kernel), (unsigned)((B + SW - 1) / SW), SW * 32, 0, st, U, V, scores, nullptr, nullptr, B, (int)C, (int)D);
  return check_launch("dot_scores");
}

int lk_dot_ce_fwd(const float* U, const float* V, float* scores, float* probs, float* rowloss, float* loss, int64_t B, int64_t C,
                  int64_t D, cudaStream_t st) {
  LK_REQUIRE(D % 4 == 0, LK_ERR_SHAPE, "lk_dot_ce_fwd: D=%ld must be a multiple of 4", (long)D);
  LK_REQUIRE(B > 0 && C > 0, LK_ERR_SHAPE, "lk_dot_ce_fwd: empty batch");
  LK_LAUNCH((dot_ce_fwd_kernel), (unsigned)((B + SW - 1) / SW), SW * 32, 0, st, U, V, scores, probs, rowloss, B, (int)C, (int)D);
  LK_LAUNCH((mean_kernel), 1, 256, 0, st, rowloss, loss, B);
  return check_launch("dot_ce_fwd", 2);
}

int lk_dot_ce_bwd(const float* U, const float* V, const float* probs, const float* dloss, float* dU, float* dV, int64_t B,
                  int64_t C, int64_t D, cudaStream_t st) {
  LK_REQUIRE(D % 4 == 0, LK_ERR_SHAPE, "lk_dot_ce_bwd: D=%ld must be a multiple of 4", (long)D);
  if (B == 0) return LK_OK;
  LK_LAUNCH((dot_ce_bwd_kernel), (unsigned)((B + SW - 1) / SW), SW * 32, 0, st, U, V, probs, dloss, dU, dV, B, (int)C, (int)D);
  return check_launch("dot_ce_bwd");
}

int lk_dot_bwd(const float* U, const float* V, const float* dS, float* dU, float* dV, int64_t B, int64_t C, int64_t D,
               cudaStream_t st) {
  LK_REQUIRE(D % 4 == 0, LK_ERR_SHAPE, "lk_dot_bwd: D=%ld must be a multiple of 4", (long)D);
  if (B == 0) return LK_OK;
  LK_LAUNCH((dot_bwd_kernel), (unsigned)((B + SW - 1) / SW), SW * 32, 0, st, U, V, dS, dU, dV, B, (int)C, (int)D);
  return check_launch("dot_bwd");
}

int lk_dot_bce_fwd(const float* U, const float* V, const float* y, float* scores, float* rowloss, float* loss, int64_t B, int64_t D,
                   cudaStream_t st) {
  LK_REQUIRE(D % 4 == 0, LK_ERR_SHAPE, "lk_dot_bce_fwd: D=%ld must be a multiple of 4", (long)D);
  LK_REQUIRE(B > 0, LK_ERR_SHAPE, "lk_dot_bce_fwd: empty batch");
  LK_LAUNCH((dot_bce_fwd_kernel), (unsigned)((B + SW - 1) / SW), SW * 32, 0, st, U, V, y, scores, rowloss, B, (int)D);
  LK_LAUNCH((mean_kernel), 1, 256, 0, st, rowloss, loss, B);
  return check_launch("dot_bce_fwd", 2);
}

int lk_dot_bce_bwd(const float* U, const float* V, const float* y, const float* scores, const float* dloss, float* dz, float* dU,
                   float* dV, int64_t B, int64_t D, cudaStream_t st) {
  LK_REQUIRE(D % 4 == 0, LK_ERR_SHAPE, "lk_dot_bce_bwd: D=%ld must be a multiple of 4", (long)D);
  if (B == 0) return LK_OK;
  LK_LAUNCH((bce_dscore_kernel), (unsigned)((B + 255) / 256), 256, 0, st, scores, y, dloss, dz, B);
  LK_LAUNCH((dot_bwd_kernel), (unsigned)((B + SW - 1) / SW), SW * 32, 0, st, U, V, dz, dU, dV, B, 1, (int)D);
  return check_launch("dot_bce_bwd", 2);
}

int lk_cached_scores(const float* U, int64_t n_users, const float* I, int64_t n_items, const int64_t* uid, const int64_t* iid, float* out,
                     int64_t R, int64_t D, cudaStream_t st) {
  LK_REQUIRE(D % 4 == 0, LK_ERR_SHAPE, "lk_cached_scores: D=%ld must be a multiple of 4", (long)D);
  if (R == 0) return LK_OK;
  int64_t blocks = (R + 2 * SW - 1) / (2 * SW);
  if (blocks > (int64_t)kNumSMs * 32) blocks = (int64_t)kNumSMs * 32;
  LK_LAUNCH((cached_scores_kernel), (unsigned)blocks, SW * 32, 0, st, U, I, uid, iid, out, R, (int)D, n_users, n_items, id_violations());
  return check_launch("cached_scores");
}

int lk_index_rows(const float* table, int64_t V, const int64_t* ids, float* out, int64_t R, int64_t D, cudaStream_t st) {
  LK_REQUIRE(D % 4 == 0, LK_ERR_SHAPE, "lk_index_rows: D=%ld must be a multiple of 4", (long)D);
  if (R == 0) return LK_OK;
  int64_t blocks = (R + SW - 1) / SW;
  if (blocks > (int64_t)kNumSMs * 32) blocks = (int64_t)kNumSMs * 32;
  LK_LAUNCH((index_rows_kernel), (unsigned)blocks, SW * 32, 0, st, table, ids, out, R, (int)D, V, id_violations());
  return check_launch("index_rows");
}

int lk_fill_f32(float* p, float value, int64_t n, cudaStream_t st) {
  if (n == 0) return LK_OK;
  LK_LAUNCH((fill_kernel), (unsigned)((n + 255) / 256), 256, 0, st, p, value, n);
  return check_launch("fill");
}

int lk_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                 int64_t step, float grad_scale, cudaStream_t st) {
  LK_REQUIRE(step >= 1, LK_ERR_ARG, "lk_adam_step: step must be >= 1");
  if (n == 0) return LK_OK;
  double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  int64_t n4 = (n + 3) / 4;
  LK_LAUNCH((adam_kernel), (unsigned)((n4 + 255) / 256), 256, 0, st, p, g, m, v, n, (float)(lr / bc1), beta1, beta2, eps,
                                                          (float)(1.0 / sqrt(bc2)), grad_scale);
  return check_launch("adam_step");
}

}  // extern "C"
