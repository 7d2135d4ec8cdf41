// Row-wise pieces of the BERT-style blocks behind TransformerOperator / FastformerOperator (model/operators/transformer_operator.py:29-38 ->
// transformers BertModel; model/common/fastformer.py:146-226): LayerNorm over the feature axis with an optional fused residual add, exact
// (erf) GELU, and stand-alone counter-based dropout.  One warp per row, 16-byte accesses; the affine-parameter gradients leave as per-block
// partial sums that are finished in a fixed order (deterministic, no atomics).
#include "lk_common.cuh"
#include "../../include/legommenders_b200.h"

namespace lk {
namespace norm {

constexpr int NW = 8;            // warps (rows) per CTA
constexpr int MAXV = 8;          // float4 per lane: D <= 1024

// y = LN(x + res) * w + b ; xs (optional) receives x + res (what the backward normalises again); mean / rstd per row
__global__ void __launch_bounds__(NW * 32) layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ res, const float* __restrict__ w,
                                                                const float* __restrict__ b, float* __restrict__ y, float* __restrict__ xs,
                                                                float* __restrict__ mean, float* __restrict__ rstd, int64_t rows, int D, float eps) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * NW + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int D4 = D >> 2;
  float4 v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; i++) {
    const int c = lane + 32 * i;
    if (c < D4) {
      v[i] = ldg4(x + r * D + c * 4);
      if (res) f4_add(v[i], ldg4(res + r * D + c * 4));
      if (xs) st4(xs + r * D + c * 4, v[i]);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float mu = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; i++) {
    const int c = lane + 32 * i;
    if (c < D4) {
      const float a = v[i].x - mu, bb = v[i].y - mu, cc = v[i].z - mu, d = v[i].w - mu;
      q += (a * a + bb * bb) + (cc * cc + d * d);
    }
  }
  const float rs = rsqrtf(warp_sum(q) / (float)D + eps);
  if (lane == 0) { mean[r] = mu; rstd[r] = rs; }
#pragma unroll
  for (int i = 0; i < MAXV; i++) {
    const int c = lane + 32 * i;
    if (c < D4) {
      const float4 ww = ldg4(w + c * 4), bv = ldg4(b + c * 4);
      st4(y + r * D + c * 4, make_float4((v[i].x - mu) * rs * ww.x + bv.x, (v[i].y - mu) * rs * ww.y + bv.y, (v[i].z - mu) * rs * ww.z + bv.z,
                                         (v[i].w - mu) * rs * ww.w + bv.w));
    }
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * w ; per-block partials of dw = sum dy * xhat and db = sum dy
__global__ void __launch_bounds__(NW * 32) layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ xs, const float* __restrict__ w,
                                                                const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ dx,
                                                                float* __restrict__ dwb_part, int64_t rows, int D, int rows_per_block) {
  pdl_prologue();
  extern __shared__ float red[];            // [NW][2 * D]
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int D4 = D >> 2;
  float4 aw[MAXV], ab[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; i++) aw[i] = ab[i] = f4_zero();
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  for (int64_t r = r0 + wp; r < min(rows, r0 + rows_per_block); r += NW) {
    const float mu = mean[r], rs = rstd[r];
    float4 xh[MAXV], g[MAXV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; i++) {
      const int c = lane + 32 * i;
      if (c < D4) {
        const float4 xv = ldg4(xs + r * D + c * 4), d = ldg4(dy + r * D + c * 4), ww = ldg4(w + c * 4);
        xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        g[i] = make_float4(d.x * ww.x, d.y * ww.y, d.z * ww.z, d.w * ww.w);
        s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
        s2 += f4_dot(g[i], xh[i]);
        aw[i].x = fmaf(d.x, xh[i].x, aw[i].x); aw[i].y = fmaf(d.y, xh[i].y, aw[i].y); aw[i].z = fmaf(d.z, xh[i].z, aw[i].z); aw[i].w = fmaf(d.w, xh[i].w, aw[i].w);
        f4_add(ab[i], d);
      }
    }
    const float m1 = warp_sum(s1) / (float)D, m2 = warp_sum(s2) / (float)D;
#pragma unroll
    for (int i = 0; i < MAXV; i++) {
      const int c = lane + 32 * i;
      if (c < D4)
        st4(dx + r * D + c * 4, make_float4(rs * (g[i].x - m1 - xh[i].x * m2), rs * (g[i].y - m1 - xh[i].y * m2), rs * (g[i].z - m1 - xh[i].z * m2),
                                            rs * (g[i].w - m1 - xh[i].w * m2)));
    }
  }
  // block partial: the NW warps' accumulators summed in warp order
#pragma unroll
  for (int i = 0; i < MAXV; i++) {
    const int c = lane + 32 * i;
    if (c < D4) {
      *reinterpret_cast<float4*>(red + wp * 2 * D + c * 4) = aw[i];
      *reinterpret_cast<float4*>(red + wp * 2 * D + D + c * 4) = ab[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * D; c += NW * 32) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < NW; k++) t += red[k * 2 * D + c];
    dwb_part[(size_t)blockIdx.x * 2 * D + c] = t;
  }
}

__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad(float x) {
  return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * expf(-0.5f * x * x);
}
// mode 0: y = gelu(x) ; mode 1: y = dy * gelu'(x)
__global__ void __launch_bounds__(256) gelu_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ y, int64_t n4, int mode) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = ldg4(x + i * 4);
    float4 o;
    if (mode == 0) o = make_float4(gelu_f(v.x), gelu_f(v.y), gelu_f(v.z), gelu_f(v.w));
    else {
      const float4 d = ldg4(dy + i * 4);
      o = make_float4(d.x * gelu_grad(v.x), d.y * gelu_grad(v.y), d.z * gelu_grad(v.z), d.w * gelu_grad(v.w));
    }
    st4(y + i * 4, o);
  }
}
// y = x * keep_scale(seed, index): the same call with dy in place of x is the backward
__global__ void __launch_bounds__(256) dropout_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, float p, float inv_keep,
                                                      unsigned long long seed) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = x[i] * dropout_scale(seed, (uint64_t)i, p, inv_keep);
}

}  // namespace norm
}  // namespace lk

using namespace lk;
using namespace lk::norm;

extern "C" {

int lk_layernorm_fwd(const float* x, const float* res, const float* w, const float* b, float* y, float* xs, float* mean, float* rstd, int64_t rows,
                     int64_t D, float eps, cudaStream_t st) {
  LK_REQUIRE(D % 4 == 0 && D <= 128 * MAXV, LK_ERR_SHAPE, "lk_layernorm_fwd: D=%ld (multiple of 4, <= %d)", (long)D, 128 * MAXV);
  if (rows == 0) return LK_OK;
  LK_LAUNCH((layernorm_fwd_kernel), (unsigned)((rows + NW - 1) / NW), NW * 32, 0, st, x, res, w, b, y, xs, mean, rstd, rows, (int)D, eps);
  return check_launch("layernorm_fwd");
}

int64_t lk_layernorm_bwd_parts(int64_t rows) {
  int64_t blocks = (rows + 255) / 256;
  return blocks < 1 ? 1 : blocks;
}

// dwb_part: [lk_layernorm_bwd_parts(rows), 2 * D] (dw partials, then db partials, per block)
int lk_layernorm_bwd(const float* dy, const float* xs, const float* w, const float* mean, const float* rstd, float* dx, float* dwb_part, int64_t rows,
                     int64_t D, cudaStream_t st) {
  LK_REQUIRE(D % 4 == 0 && D <= 128 * MAXV, LK_ERR_SHAPE, "lk_layernorm_bwd: D=%ld (multiple of 4, <= %d)", (long)D, 128 * MAXV);
  if (rows == 0) return LK_OK;
  const int64_t blocks = lk_layernorm_bwd_parts(rows);
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(layernorm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NW * 2 * 128 * MAXV * 4); attr = true; }
  LK_LAUNCH((layernorm_bwd_kernel), (unsigned)blocks, NW * 32, NW * 2 * D * sizeof(float), st, dy, xs, w, mean, rstd, dx, dwb_part, rows, (int)D, 256);
  return check_launch("layernorm_bwd");
}

int lk_gelu(const float* x, const float* dy, float* y, int64_t n, int mode, cudaStream_t st) {
  LK_REQUIRE(n % 4 == 0 && (mode == 0 || (mode == 1 && dy)), LK_ERR_ARG, "lk_gelu: n=%ld must be a multiple of 4; mode 1 needs dy", (long)n);
  if (n == 0) return LK_OK;
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
  LK_LAUNCH((gelu_kernel), (unsigned)blocks, 256, 0, st, x, dy, y, n / 4, mode);
  return check_launch("gelu");
}

int lk_dropout(const float* x, float* y, int64_t n, float p, uint64_t seed, cudaStream_t st) {
  LK_REQUIRE(p >= 0.f && p < 1.f, LK_ERR_ARG, "lk_dropout: p=%f", p);
  if (n == 0) return LK_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
  LK_LAUNCH((dropout_kernel), (unsigned)blocks, 256, 0, st, x, y, n, p, 1.f / (1.f - p), (unsigned long long)seed);
  return check_launch("dropout");
}

}  // extern "C"
