// MINER: poly-attention user encoder and target-aware predictor (model/operators/poly_attention_operator.py:45-58, model/predictors/
// miner_predictor.py:30-64).  The dense parts (tanh(Linear), the code logits, gelu(Linear)) are contractions of the library; what is left are
// per-user / per-impression softmax-weighted sums over a handful of rows, one CTA each:
//   poly_pool   w[c, s] = softmax_s(mask[s] ? logit[s, c] : 1e-30);  out[c, :] = sum_s w[c, s] x[s, :]
//               (the reference fills masked positions with 1e-30 — a logit of ~0, NOT -inf: padded clicks take part with weight exp(0);
//                restated literally)
//   miner_score s[i, c] = <v_i, u_c>;  a[i, c] = <v_i, p_c>;  out[i] = sum_c softmax_c(a[i, :])[c] s[i, c]   (weighted) | max_c s | mean_c s
#include "lk_common.cuh"
#include "../../include/legommenders_b200.h"

namespace lk {
namespace poly {

constexpr int PT = 256;
constexpr int MAXC = 64, MAXS = 128;

__device__ __forceinline__ float dot_row(const float* __restrict__ a, const float* __restrict__ b, int D, int lane) {
  float s = 0.f;
  for (int c = lane * 4; c < D; c += 128) s += f4_dot(ldg4(a + c), ldg4(b + c));
  return warp_sum(s);
}

__global__ void __launch_bounds__(PT) poly_pool_fwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ mask, const float* __restrict__ x,
                                                           float* __restrict__ out, float* __restrict__ wsave, int S, int C, int D) {
  pdl_prologue();
  extern __shared__ float w_s[];                      // [C][S]
  const int b = blockIdx.x, lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const float* lg = logits + (size_t)b * S * C;
  const int64_t* mk = mask + (size_t)b * S;
  for (int c = wp; c < C; c += PT / 32) {             // a warp per code: softmax over the sequence
    float mx = -INFINITY;
    for (int s = lane; s < S; s += 32) mx = fmaxf(mx, mk[s] > 0 ? lg[s * C + c] : 1e-30f);
    mx = warp_max(mx);
    float z = 0.f;
    for (int s = lane; s < S; s += 32) {
      const float e = expf((mk[s] > 0 ? lg[s * C + c] : 1e-30f) - mx);
      w_s[c * S + s] = e;
      z += e;
    }
    const float inv = 1.f / warp_sum(z);
    for (int s = lane; s < S; s += 32) {
      const float w = w_s[c * S + s] * inv;
      w_s[c * S + s] = w;
      wsave[((size_t)b * C + c) * S + s] = w;
    }
  }
  __syncthreads();
  const float* xb = x + (size_t)b * S * D;
  for (int d = threadIdx.x; d < D; d += PT) {
    for (int c0 = 0; c0 < C; c0 += 16) {              // 16 codes at a time in registers
      float acc[16];
#pragma unroll
      for (int i = 0; i < 16; i++) acc[i] = 0.f;
      for (int s = 0; s < S; s++) {
        const float xv = xb[(size_t)s * D + d];
#pragma unroll
        for (int i = 0; i < 16; i++)
          if (c0 + i < C) acc[i] = fmaf(w_s[(c0 + i) * S + s], xv, acc[i]);
      }
#pragma unroll
      for (int i = 0; i < 16; i++)
        if (c0 + i < C) out[((size_t)b * C + c0 + i) * D + d] = acc[i];
    }
  }
}

__global__ void __launch_bounds__(PT) poly_pool_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ wsave, const int64_t* __restrict__ mask,
                                                           const float* __restrict__ x, float* __restrict__ dx, float* __restrict__ dlogits, int S, int C,
                                                           int D) {
  pdl_prologue();
  extern __shared__ float sm[];                       // w [C][S], dw [C][S]
  float* w_s = sm;
  float* dw_s = sm + C * S;
  const int b = blockIdx.x, lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const float* xb = x + (size_t)b * S * D;
  const float* db = dout + (size_t)b * C * D;
  for (int i = threadIdx.x; i < C * S; i += PT) w_s[i] = wsave[(size_t)b * C * S + i];
  __syncthreads();
  for (int i = wp; i < C * S; i += PT / 32) {          // dw[c, s] = <dout[c, :], x[s, :]>
    const int c = i / S, s = i - c * S;
    const float v = dot_row(db + (size_t)c * D, xb + (size_t)s * D, D, lane);
    if (lane == 0) dw_s[i] = v;
  }
  __syncthreads();
  for (int c = wp; c < C; c += PT / 32) {              // softmax backward per code; masked positions carry a constant logit: no gradient
    float t = 0.f;
    for (int s = lane; s < S; s += 32) t = fmaf(w_s[c * S + s], dw_s[c * S + s], t);
    t = warp_sum(t);
    for (int s = lane; s < S; s += 32)
      dlogits[((size_t)b * S + s) * C + c] = mask[(size_t)b * S + s] > 0 ? w_s[c * S + s] * (dw_s[c * S + s] - t) : 0.f;
  }
  for (int d = threadIdx.x; d < D; d += PT) {          // dx[s, :] = sum_c w[c, s] dout[c, :]
    for (int s = 0; s < S; s++) {
      float acc = 0.f;
      for (int c = 0; c < C; c++) acc = fmaf(w_s[c * S + s], db[(size_t)c * D + d], acc);
      dx[((size_t)b * S + s) * D + d] = acc;
    }
  }
}

// user [B, C, D], proj [B, C, D] (gelu(Linear(user)); unused for mode != 0), items [B, K1, D] -> out [B, K1]; saved: sc, wt [B, K1, C]
__global__ void __launch_bounds__(PT) miner_fwd_kernel(const float* __restrict__ user, const float* __restrict__ proj, const float* __restrict__ items,
                                                       float* __restrict__ out, float* __restrict__ sc, float* __restrict__ wt, int K1, int C, int D,
                                                       int mode) {
  pdl_prologue();
  extern __shared__ float sm[];                        // s [K1][C], a [K1][C]
  float* s_s = sm;
  float* a_s = sm + K1 * C;
  const int b = blockIdx.x, lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const float* u = user + (size_t)b * C * D;
  const float* p = proj + (size_t)b * C * D;
  const float* v = items + (size_t)b * K1 * D;
  for (int i = wp; i < K1 * C; i += PT / 32) {
    const int k = i / C, c = i - k * C;
    const float s = dot_row(v + (size_t)k * D, u + (size_t)c * D, D, lane);
    const float a = mode == 0 ? dot_row(v + (size_t)k * D, p + (size_t)c * D, D, lane) : 0.f;
    if (lane == 0) { s_s[i] = s; a_s[i] = a; }
  }
  __syncthreads();
  for (int k = wp; k < K1; k += PT / 32) {             // a warp per candidate
    float res;
    if (mode == 0) {
      float mx = -INFINITY;
      for (int c = lane; c < C; c += 32) mx = fmaxf(mx, a_s[k * C + c]);
      mx = warp_max(mx);
      float z = 0.f;
      for (int c = lane; c < C; c += 32) { const float e = expf(a_s[k * C + c] - mx); a_s[k * C + c] = e; z += e; }
      const float inv = 1.f / warp_sum(z);
      float t = 0.f;
      for (int c = lane; c < C; c += 32) {
        const float w = a_s[k * C + c] * inv;
        wt[((size_t)b * K1 + k) * C + c] = w;
        t = fmaf(w, s_s[k * C + c], t);
      }
      res = warp_sum(t);
    } else if (mode == 1) {
      float mx = -INFINITY;
      for (int c = lane; c < C; c += 32) mx = fmaxf(mx, s_s[k * C + c]);
      res = warp_max(mx);
      bool taken = false;                              // one-hot "weights" of the FIRST maximum (torch.max's gradient convention)
      for (int c0 = 0; c0 < C; c0 += 32) {
        const int c = c0 + lane;
        const bool hit = c < C && s_s[k * C + c] == res;
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        const int first = (m && !taken) ? __ffs(m) - 1 : -1;
        if (c < C) wt[((size_t)b * K1 + k) * C + c] = (lane == first) ? 1.f : 0.f;
        taken = taken || m != 0;
      }
    } else {
      float t = 0.f;
      for (int c = lane; c < C; c += 32) { t += s_s[k * C + c]; wt[((size_t)b * K1 + k) * C + c] = 1.f / (float)C; }
      res = warp_sum(t) / (float)C;
    }
    for (int c = lane; c < C; c += 32) sc[((size_t)b * K1 + k) * C + c] = s_s[k * C + c];
    if (lane == 0) out[(size_t)b * K1 + k] = res;
  }
}

// dout [B, K1] -> duser, dproj [B, C, D], ditems [B, K1, D]
__global__ void __launch_bounds__(PT) miner_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ user, const float* __restrict__ proj,
                                                       const float* __restrict__ items, const float* __restrict__ sc, const float* __restrict__ wt,
                                                       float* __restrict__ duser, float* __restrict__ dproj, float* __restrict__ ditems, int K1, int C,
                                                       int D, int mode) {
  pdl_prologue();
  extern __shared__ float sm[];                        // ds [K1][C], da [K1][C]
  float* ds_s = sm;
  float* da_s = sm + K1 * C;
  const int b = blockIdx.x, lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  for (int k = wp; k < K1; k += PT / 32) {
    const float g = dout[(size_t)b * K1 + k];
    const float* w = wt + ((size_t)b * K1 + k) * C;
    const float* s = sc + ((size_t)b * K1 + k) * C;
    float t = 0.f;
    if (mode == 0)
      for (int c = lane; c < C; c += 32) t = fmaf(w[c], s[c], t);
    t = warp_sum(t);
    for (int c = lane; c < C; c += 32) {
      ds_s[k * C + c] = g * w[c];                                        // d out / d s = w (weighted, one-hot max, 1/C mean)
      da_s[k * C + c] = mode == 0 ? g * w[c] * (s[c] - t) : 0.f;        // through the softmax of a
    }
  }
  __syncthreads();
  const float* u = user + (size_t)b * C * D;
  const float* p = proj + (size_t)b * C * D;
  const float* v = items + (size_t)b * K1 * D;
  for (int d = threadIdx.x; d < D; d += PT) {
    for (int k = 0; k < K1; k++) {
      float acc = 0.f;
      for (int c = 0; c < C; c++) {
        acc = fmaf(ds_s[k * C + c], u[(size_t)c * D + d], acc);
        if (mode == 0) acc = fmaf(da_s[k * C + c], p[(size_t)c * D + d], acc);
      }
      ditems[((size_t)b * K1 + k) * D + d] = acc;
    }
    for (int c = 0; c < C; c++) {
      float au = 0.f, ap = 0.f;
      for (int k = 0; k < K1; k++) {
        const float vv = v[(size_t)k * D + d];
        au = fmaf(ds_s[k * C + c], vv, au);
        ap = fmaf(da_s[k * C + c], vv, ap);
      }
      duser[((size_t)b * C + c) * D + d] = au;
      dproj[((size_t)b * C + c) * D + d] = ap;
    }
  }
}

}  // namespace poly
}  // namespace lk

using namespace lk;
using namespace lk::poly;

extern "C" {

int lk_poly_pool_fwd(const float* logits, const int64_t* mask, const float* x, float* out, float* w, int64_t B, int64_t S, int64_t C, int64_t D,
                     cudaStream_t st) {
  LK_REQUIRE(C >= 1 && C <= MAXC && S >= 1 && S <= MAXS && D % 4 == 0, LK_ERR_SHAPE, "lk_poly_pool_fwd: C=%ld (<=%d) S=%ld (<=%d) D=%ld", (long)C, MAXC,
             (long)S, MAXS, (long)D);
  if (B == 0) return LK_OK;
  LK_LAUNCH((poly_pool_fwd_kernel), (unsigned)B, PT, C * S * sizeof(float), st, logits, mask, x, out, w, (int)S, (int)C, (int)D);
  return check_launch("poly_pool_fwd");
}

int lk_poly_pool_bwd(const float* dout, const float* w, const int64_t* mask, const float* x, float* dx, float* dlogits, int64_t B, int64_t S, int64_t C,
                     int64_t D, cudaStream_t st) {
  LK_REQUIRE(C >= 1 && C <= MAXC && S >= 1 && S <= MAXS && D % 4 == 0, LK_ERR_SHAPE, "lk_poly_pool_bwd: bad shape");
  if (B == 0) return LK_OK;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(poly_pool_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * MAXC * MAXS * 4); attr = true; }
  LK_LAUNCH((poly_pool_bwd_kernel), (unsigned)B, PT, 2 * C * S * sizeof(float), st, dout, w, mask, x, dx, dlogits, (int)S, (int)C, (int)D);
  return check_launch("poly_pool_bwd");
}

int lk_miner_fwd(const float* user, const float* proj, const float* items, float* out, float* sc, float* wt, int64_t B, int64_t K1, int64_t C, int64_t D,
                 int mode, cudaStream_t st) {
  LK_REQUIRE(C >= 1 && C <= MAXC && K1 >= 1 && K1 <= 64 && D % 4 == 0 && mode >= 0 && mode <= 2, LK_ERR_SHAPE, "lk_miner_fwd: bad shape / mode");
  LK_REQUIRE(mode != 0 || proj, LK_ERR_ARG, "lk_miner_fwd: the weighted score needs the projected user codes");
  if (B == 0) return LK_OK;
  LK_LAUNCH((miner_fwd_kernel), (unsigned)B, PT, 2 * K1 * C * sizeof(float), st, user, proj ? proj : user, items, out, sc, wt, (int)K1, (int)C, (int)D, mode);
  return check_launch("miner_fwd");
}

int lk_miner_bwd(const float* dout, const float* user, const float* proj, const float* items, const float* sc, const float* wt, float* duser, float* dproj,
                 float* ditems, int64_t B, int64_t K1, int64_t C, int64_t D, int mode, cudaStream_t st) {
  LK_REQUIRE(C >= 1 && C <= MAXC && K1 >= 1 && K1 <= 64 && D % 4 == 0 && mode >= 0 && mode <= 2, LK_ERR_SHAPE, "lk_miner_bwd: bad shape / mode");
  if (B == 0) return LK_OK;
  LK_LAUNCH((miner_bwd_kernel), (unsigned)B, PT, 2 * K1 * C * sizeof(float), st, dout, user, proj ? proj : user, items, sc, wt, duser, dproj, ditems, (int)K1,
            (int)C, (int)D, mode);
  return check_launch("miner_bwd");
}

}  // extern "C"
