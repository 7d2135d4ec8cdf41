// Fastformer additive attention (model/common/fastformer.py:62-143, FastSelfAttention): the two per-head softmax poolings over the sequence and
// the two broadcast products around them.  Dense pieces (query / key / *_att / transform, BertSelfOutput, BertIntermediate, BertOutput) are
// contractions + LayerNorm / GELU kernels of the library.
//   head_pool   w[h, s] = softmax_s(score[s, h] * scale + (1 - mask[s]) * (-10000));  out[h*dh + j] = sum_s w[h, s] v[s, h*dh + j]
//               (the reference's additive -10000 mask, restated literally — not -inf)
//   bcast_mul   y[s, :] = a[s, :] * v[:]   (pooled vector broadcast over the sequence);  backward: da = dy * v,  dv = sum_s dy * a
//   add         y = a + b
#include "lk_common.cuh"
#include "../../include/legommenders_b200.h"

namespace lk {
namespace fast {

constexpr int FT = 256;
constexpr int MAXH = 32, MAXS = 128;

__global__ void __launch_bounds__(FT) head_pool_fwd_kernel(const float* __restrict__ score, const int64_t* __restrict__ mask, const float* __restrict__ v,
                                                           float* __restrict__ out, float* __restrict__ wsave, int S, int H, int D, float scale) {
  pdl_prologue();
  extern __shared__ float w_s[];                      // [H][S]
  const int b = blockIdx.x, lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const float* sc = score + (size_t)b * S * H;
  const int64_t* mk = mask + (size_t)b * S;
  for (int h = wp; h < H; h += FT / 32) {
    float mx = -INFINITY;
    for (int s = lane; s < S; s += 32) mx = fmaxf(mx, sc[s * H + h] * scale + (mk[s] > 0 ? 0.f : -10000.f));
    mx = warp_max(mx);
    float z = 0.f;
    for (int s = lane; s < S; s += 32) {
      const float e = expf(sc[s * H + h] * scale + (mk[s] > 0 ? 0.f : -10000.f) - mx);
      w_s[h * S + s] = e;
      z += e;
    }
    const float inv = 1.f / warp_sum(z);
    for (int s = lane; s < S; s += 32) {
      const float w = w_s[h * S + s] * inv;
      w_s[h * S + s] = w;
      wsave[((size_t)b * H + h) * S + s] = w;
    }
  }
  __syncthreads();
  const int dh = D / H;
  const float* vb = v + (size_t)b * S * D;
  for (int d = threadIdx.x; d < D; d += FT) {
    const float* w = w_s + (d / dh) * S;
    float acc = 0.f;
    for (int s = 0; s < S; s++) acc = fmaf(w[s], vb[(size_t)s * D + d], acc);
    out[(size_t)b * D + d] = acc;
  }
}

// dout [B, D] -> dv [B, S, D], dscore [B, S, H]
__global__ void __launch_bounds__(FT) head_pool_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ wsave, const float* __restrict__ v,
                                                           float* __restrict__ dv, float* __restrict__ dscore, int S, int H, int D, float scale) {
  pdl_prologue();
  extern __shared__ float sm[];                       // w [H][S], dw [H][S]
  float* w_s = sm;
  float* dw_s = sm + H * S;
  const int b = blockIdx.x, lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int dh = D / H;
  const float* vb = v + (size_t)b * S * D;
  const float* g = dout + (size_t)b * D;
  for (int i = threadIdx.x; i < H * S; i += FT) w_s[i] = wsave[(size_t)b * H * S + i];
  __syncthreads();
  for (int i = wp; i < H * S; i += FT / 32) {          // dw[h, s] = <dout[h-block], v[s, h-block]>
    const int h = i / S, s = i - h * S;
    float t = 0.f;
    for (int j = lane; j < dh; j += 32) t = fmaf(g[h * dh + j], vb[(size_t)s * D + h * dh + j], t);
    t = warp_sum(t);
    if (lane == 0) dw_s[i] = t;
  }
  __syncthreads();
  for (int h = wp; h < H; h += FT / 32) {
    float t = 0.f;
    for (int s = lane; s < S; s += 32) t = fmaf(w_s[h * S + s], dw_s[h * S + s], t);
    t = warp_sum(t);
    for (int s = lane; s < S; s += 32) dscore[((size_t)b * S + s) * H + h] = scale * w_s[h * S + s] * (dw_s[h * S + s] - t);
  }
  for (int d = threadIdx.x; d < D; d += FT) {
    const float* w = w_s + (d / dh) * S;
    const float gd = g[d];
    for (int s = 0; s < S; s++) dv[((size_t)b * S + s) * D + d] = w[s] * gd;
  }
}

// y[b, s, :] = a[b, s, :] * v[b, :]
__global__ void __launch_bounds__(FT) bcast_mul_kernel(const float* __restrict__ a, const float* __restrict__ v, float* __restrict__ y, int64_t rows, int S,
                                                       int D4) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows * D4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / D4;
    const int c = (int)(i - r * D4);
    const float4 x = ldg4(a + i * 4), w = ldg4(v + ((r / S) * D4 + c) * 4);
    st4(y + i * 4, make_float4(x.x * w.x, x.y * w.y, x.z * w.z, x.w * w.w));
  }
}
// dv[b, :] = sum_s dy[b, s, :] * a[b, s, :]   (fixed order over s)
__global__ void __launch_bounds__(FT) bcast_mul_dv_kernel(const float* __restrict__ dy, const float* __restrict__ a, float* __restrict__ dv, int S, int D) {
  pdl_prologue();
  const int b = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += FT) {
    float acc = 0.f;
    for (int s = 0; s < S; s++) acc = fmaf(dy[((size_t)b * S + s) * D + d], a[((size_t)b * S + s) * D + d], acc);
    dv[(size_t)b * D + d] = acc;
  }
}
__global__ void __launch_bounds__(FT) add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, int64_t n4) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 x = ldg4(a + i * 4);
    f4_add(x, ldg4(b + i * 4));
    st4(y + i * 4, x);
  }
}

static unsigned grid_for(int64_t work) {
  int64_t g = (work + FT - 1) / FT;
  if (g > (int64_t)kNumSMs * 16) g = (int64_t)kNumSMs * 16;
  return (unsigned)(g < 1 ? 1 : g);
}

}  // namespace fast
}  // namespace lk

using namespace lk;
using namespace lk::fast;

extern "C" {

int lk_head_pool_fwd(const float* score, const int64_t* mask, const float* v, float* out, float* w, int64_t B, int64_t S, int64_t H, int64_t D, float scale,
                     cudaStream_t st) {
  LK_REQUIRE(H >= 1 && H <= MAXH && S >= 1 && S <= MAXS && D % H == 0, LK_ERR_SHAPE, "lk_head_pool_fwd: H=%ld (<=%d) S=%ld (<=%d) D=%ld", (long)H, MAXH, (long)S,
             MAXS, (long)D);
  if (B == 0) return LK_OK;
  LK_LAUNCH((head_pool_fwd_kernel), (unsigned)B, FT, H * S * sizeof(float), st, score, mask, v, out, w, (int)S, (int)H, (int)D, scale);
  return check_launch("head_pool_fwd");
}

int lk_head_pool_bwd(const float* dout, const float* w, const float* v, float* dv, float* dscore, int64_t B, int64_t S, int64_t H, int64_t D, float scale,
                     cudaStream_t st) {
  LK_REQUIRE(H >= 1 && H <= MAXH && S >= 1 && S <= MAXS && D % H == 0, LK_ERR_SHAPE, "lk_head_pool_bwd: bad shape");
  if (B == 0) return LK_OK;
  LK_LAUNCH((head_pool_bwd_kernel), (unsigned)B, FT, 2 * H * S * sizeof(float), st, dout, w, v, dv, dscore, (int)S, (int)H, (int)D, scale);
  return check_launch("head_pool_bwd");
}

int lk_bcast_mul(const float* a, const float* v, float* y, int64_t B, int64_t S, int64_t D, cudaStream_t st) {
  LK_REQUIRE(D % 4 == 0, LK_ERR_SHAPE, "lk_bcast_mul: D=%ld must be a multiple of 4", (long)D);
  if (B * S == 0) return LK_OK;
  LK_LAUNCH((bcast_mul_kernel), grid_for(B * S * (D / 4)), FT, 0, st, a, v, y, B * S, (int)S, (int)(D / 4));
  return check_launch("bcast_mul");
}

int lk_bcast_mul_dv(const float* dy, const float* a, float* dv, int64_t B, int64_t S, int64_t D, cudaStream_t st) {
  if (B == 0) return LK_OK;
  LK_LAUNCH((bcast_mul_dv_kernel), (unsigned)B, FT, 0, st, dy, a, dv, (int)S, (int)D);
  return check_launch("bcast_mul_dv");
}

int lk_add(const float* a, const float* b, float* y, int64_t n, cudaStream_t st) {
  LK_REQUIRE(n % 4 == 0, LK_ERR_SHAPE, "lk_add: n=%ld must be a multiple of 4", (long)n);
  if (n == 0) return LK_OK;
  LK_LAUNCH((add_kernel), grid_for(n / 4), FT, 0, st, a, b, y, n / 4);
  return check_launch("add");
}

}  // extern "C"
