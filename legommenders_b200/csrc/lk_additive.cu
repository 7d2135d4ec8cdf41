// AdditiveAttention pooling (north_star piece 3; also the tail of every item encoder).
//
// Reference model/common/attention.py:31-38, restated literally (no max-subtraction):
//   s[t] = w2 · tanh(W1 x[t] + b1);  a[t] = exp(s[t]) * mask[t];  alpha = a / (sum_t a + eps32);  out = sum_t alpha[t] x[t]
// The W1 GEMM (+bias+tanh) is done by the linear kernels; this file fuses everything after it.
// One CTA owns one sequence; masked positions are never loaded.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "lk_common.cuh"
#include "../../include/legommenders_b200.h"

namespace lk {

constexpr int AT = 256;        // threads per CTA
constexpr int MAXS = 128;      // max sequence length
constexpr float EPS32 = 1.1920928955078125e-07f;

// Work split inside a CTA: the row-wise dot products go warp-per-row; the column-wise sums over rows split the rows over
// G = AT / (cols/4) groups of column-quad threads (4 groups at 256 columns) whose partial sums are combined through shared
// memory in a fixed order.  A 50-item click history then costs 13 dependent loads per thread instead of 50.
struct ColSplit { int cq, groups, rg, q; bool active; };
__device__ __forceinline__ ColSplit col_split(int cols) {
  ColSplit c;
  c.cq = cols >> 2;
  c.groups = c.cq >= AT ? 1 : AT / c.cq;
  c.rg = c.groups == 1 ? 0 : threadIdx.x / c.cq;
  c.q = c.groups == 1 ? threadIdx.x : threadIdx.x - c.rg * c.cq;
  c.active = c.rg < c.groups;
  return c;
}
// sum of `v` over the row groups for column quad `q` (groups > 1: q < cq <= AT/2); result valid in group 0
__device__ __forceinline__ float4 group_sum(float4 v, const ColSplit& c, float4* red) {
  if (c.groups == 1) return v;
  __syncthreads();
  if (c.active && c.rg > 0) red[(c.rg - 1) * c.cq + c.q] = v;
  __syncthreads();
  if (c.rg == 0)
    for (int g = 1; g < c.groups; g++) f4_add(v, red[(g - 1) * c.cq + c.q]);
  return v;
}

// Rows of X either as fp32 or as the split-bf16 image (hi + lo, 16 mantissa bits) the fused GEMM chain leaves behind (lk_tc_chain):
// same bytes per element, and the fp32 copy of `lin` then never has to be written at all.
struct RowSrc {
  const float* f32;
  const __nv_bfloat16 *hi, *lo;
  int64_t ld;                    // pitch of the planes
};
__device__ __forceinline__ float4 load4(const RowSrc& X, int64_t row, int D, int c) {
  if (X.f32) return ldg4(X.f32 + row * D + c);
  const uint2 h = __ldg(reinterpret_cast<const uint2*>(X.hi + row * X.ld + c));
  const uint2 l = __ldg(reinterpret_cast<const uint2*>(X.lo + row * X.ld + c));
  return make_float4(__uint_as_float(h.x << 16) + __uint_as_float(l.x << 16), __uint_as_float(h.x & 0xffff0000u) + __uint_as_float(l.x & 0xffff0000u),
                     __uint_as_float(h.y << 16) + __uint_as_float(l.y << 16), __uint_as_float(h.y & 0xffff0000u) + __uint_as_float(l.y & 0xffff0000u));
}

template <int OCC>
__global__ void __launch_bounds__(AT, OCC) additive_pool_fwd_kernel(const RowSrc X, const float* __restrict__ Hd, const float* __restrict__ s_part,
                                                               const float* __restrict__ w2, const int64_t* __restrict__ mask,
                                                               const int* __restrict__ cu, float* __restrict__ out,
                                                               float* __restrict__ alpha, int Smax, int D, int A) {
  pdl_prologue();
  __shared__ float a_s[MAXS];
  __shared__ float4 red[AT];
  const int64_t n = blockIdx.x;
  const int64_t r0 = cu ? cu[n] : n * Smax;            // first row of this sequence
  const int S = cu ? cu[n + 1] - cu[n] : Smax;
  if (cu) mask = nullptr;                               // packed rows are all valid
  const int64_t mrow = n * Smax;
  alpha += r0;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (s_part) {      // scores already reduced by the producer of Hd (four partial sums per row, fixed order)
    for (int t = threadIdx.x; t < S; t += AT) {
      const bool valid = mask ? mask[mrow + t] > 0 : true;
      const float4 sp = ldg4(s_part + (r0 + t) * 4);
      a_s[t] = valid ? expf((sp.x + sp.y) + (sp.z + sp.w)) : 0.f;
    }
  } else {
    const float* Hr = Hd + r0 * A;
    for (int t = w; t < S; t += AT / 32) {
      bool valid = mask ? mask[mrow + t] > 0 : true;
      float s = 0.f;
      if (valid) {
        const float* h = Hr + t * (int64_t)A;
        for (int c = lane * 4; c < A; c += 128) s += f4_dot(ldg4(h + c), ldg4(w2 + c));
        s = warp_sum(s);
      }
      if (lane == 0) a_s[t] = valid ? expf(s) : 0.f;
    }
  }
  __syncthreads();
  float Z = 0.f;
  for (int t = 0; t < S; t++) Z += a_s[t];
  const float inv = 1.f / (Z + EPS32);
  for (int t = threadIdx.x; t < S; t += AT) alpha[t] = a_s[t] * inv;
  const ColSplit cs = col_split(D);
  for (int q0 = 0; q0 < cs.cq; q0 += AT) {               // one round unless D > 4*AT
    const int q = q0 + cs.q;
    float4 acc = f4_zero();
    if (cs.active && q < cs.cq) {
#pragma unroll 4
      for (int t = cs.rg; t < S; t += cs.groups) {
        const float al = a_s[t] * inv;
        if (al != 0.f) f4_fma(acc, al, load4(X, r0 + t, D, q * 4));
      }
    }
    acc = group_sum(acc, cs, red);
    if (cs.rg == 0 && cs.active && q < cs.cq) st4(out + n * (int64_t)D + q * 4, acc);
  }
}

// (hi, lo) bf16 planes of four floats (the operand format of the tcgen05 contractions, lk_tc.cuh:split4)
__device__ __forceinline__ void split4_planes(const float4& v, uint2& hi, uint2& lo) {
  uint32_t h01, h23, l01, l23;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h01) : "f"(v.y), "f"(v.x));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h23) : "f"(v.w), "f"(v.z));
  const float r0 = v.x - __uint_as_float(h01 << 16), r1 = v.y - __uint_as_float(h01 & 0xffff0000u);
  const float r2 = v.z - __uint_as_float(h23 << 16), r3 = v.w - __uint_as_float(h23 & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l01) : "f"(r1), "f"(r0));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l23) : "f"(r3), "f"(r2));
  hi = make_uint2(h01, h23);
  lo = make_uint2(l01, l23);
}

// Where dpre goes: fp32 rows, or directly the split-bf16 planes the next contractions read (dpre·W1 and dpreᵀ·lin) together with the
// per-sequence column sums of dpre (the b1 gradient) — the separate split pass over dpre and the fp32 copy then never exist.
struct DpreOut {
  float* f32;
  __nv_bfloat16 *hi, *lo;
  int64_t ld;
  float* colsum_part;            // [N, A] with planes
};

// dX[t,:] (+)= alpha[t] * dOut ;  dpre[t,:] = ds[t] * w2 * (1 - h^2) ;  dw2_part[n,:] = sum_t ds[t] * h[t,:]
template <int OCC>
__global__ void __launch_bounds__(AT, OCC) additive_pool_bwd_kernel(const RowSrc X, const float* __restrict__ Hd,
                                                               const float* __restrict__ w2, const float* __restrict__ alpha,
                                                               const int* __restrict__ cu, const float* __restrict__ dOut,
                                                               float* __restrict__ dX, const DpreOut dp,
                                                               float* __restrict__ dw2_part, int Smax, int D, int A,
                                                               int accumulate_dx) {
  pdl_prologue();
  __shared__ float al_s[MAXS], da_s[MAXS], ds_s[MAXS];
  __shared__ float4 red[AT];
  const int64_t n = blockIdx.x;
  const int64_t r0 = cu ? cu[n] : n * Smax;
  const int S = cu ? cu[n + 1] - cu[n] : Smax;
  Hd += r0 * A; alpha += r0; dX += r0 * D;
  float* dpre = dp.f32 ? dp.f32 + r0 * A : nullptr;
  __nv_bfloat16* dhi = dp.hi ? dp.hi + r0 * dp.ld : nullptr;
  __nv_bfloat16* dlo = dp.hi ? dp.lo + r0 * dp.ld : nullptr;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int t = threadIdx.x; t < S; t += AT) al_s[t] = alpha[t];
  __syncthreads();
  const float* g = dOut + n * (int64_t)D;
  for (int t = w; t < S; t += AT / 32) {
    float s = 0.f;
    if (al_s[t] != 0.f) {
      for (int c = lane * 4; c < D; c += 128) s += f4_dot(load4(X, r0 + t, D, c), ldg4(g + c));
      s = warp_sum(s);
    }
    if (lane == 0) da_s[t] = s;
  }
  __syncthreads();
  float cs = 0.f;
  for (int t = 0; t < S; t++) cs = fmaf(al_s[t], da_s[t], cs);
  for (int t = threadIdx.x; t < S; t += AT) ds_s[t] = al_s[t] * (da_s[t] - cs);
  __syncthreads();
  // dX: every (row, column quad) pair is independent
  const int DQ = D >> 2;
  for (int idx = threadIdx.x; idx < S * DQ; idx += AT) {
    const int t = idx / DQ, c = (idx - t * DQ) * 4;
    const float al = al_s[t];
    const float4 gv = ldg4(g + c);
    float4 v = make_float4(al * gv.x, al * gv.y, al * gv.z, al * gv.w);
    float* o = dX + t * (int64_t)D + c;
    if (accumulate_dx) f4_add(v, *reinterpret_cast<const float4*>(o));
    st4(o, v);
  }
  // dpre rows and the per-sequence partial of dw2
  const ColSplit sp = col_split(A);
  for (int q0 = 0; q0 < sp.cq; q0 += AT) {
    const int q = q0 + sp.q;
    float4 acc = f4_zero(), bsum = f4_zero();
    if (sp.active && q < sp.cq) {
      const float4 wv = ldg4(w2 + q * 4);
#pragma unroll 4
      for (int t = sp.rg; t < S; t += sp.groups) {
        const float ds = ds_s[t];
        float4 o = f4_zero();
        if (ds != 0.f) {
          const float4 h = ldg4(Hd + t * (int64_t)A + q * 4);
          f4_fma(acc, ds, h);
          o.x = ds * wv.x * (1.f - h.x * h.x); o.y = ds * wv.y * (1.f - h.y * h.y);
          o.z = ds * wv.z * (1.f - h.z * h.z); o.w = ds * wv.w * (1.f - h.w * h.w);
        }
        if (dpre) st4(dpre + t * (int64_t)A + q * 4, o);
        if (dhi) {
          uint2 h2, l2;
          split4_planes(o, h2, l2);
          *reinterpret_cast<uint2*>(dhi + t * dp.ld + q * 4) = h2;
          *reinterpret_cast<uint2*>(dlo + t * dp.ld + q * 4) = l2;
          f4_add(bsum, o);
        }
      }
    }
    acc = group_sum(acc, sp, red);
    if (sp.rg == 0 && sp.active && q < sp.cq) st4(dw2_part + n * (int64_t)A + q * 4, acc);
    if (dp.colsum_part) {
      bsum = group_sum(bsum, sp, red);
      if (sp.rg == 0 && sp.active && q < sp.cq) st4(dp.colsum_part + n * (int64_t)A + q * 4, bsum);
    }
  }
}

// masked mean / max pooling of already-gathered embeddings (model/operators/pooling_operator.py:46-56)
template <int MODE>
__global__ void masked_pool_kernel(const float* __restrict__ X, const int64_t* __restrict__ mask, float* __restrict__ out,
                                   int64_t N, int S, int D) {
  pdl_prologue();
  int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int D4 = D >> 2;
  if (i4 >= N * D4) return;
  int64_t n = i4 / D4;
  int c = (int)(i4 % D4) * 4;
  float4 acc = MODE == 1 ? make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY) : f4_zero();
  int cnt = 0;
  for (int t = 0; t < S; t++) {
    bool valid = mask[n * S + t] > 0;
    float4 v = valid ? ldg4(X + (n * S + t) * (int64_t)D + c) : f4_zero();
    cnt += valid;
    if (MODE == 1) {
      acc.x = fmaxf(acc.x, v.x); acc.y = fmaxf(acc.y, v.y); acc.z = fmaxf(acc.z, v.z); acc.w = fmaxf(acc.w, v.w);
    } else {
      f4_add(acc, v);
    }
  }
  if (MODE == 0) {
    float inv = 1.f / ((float)cnt + 1e-8f);
    acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
  }
  st4(out + n * (int64_t)D + c, acc);
}

// backward of masked mean pooling: dX[n,t,:] = mask ? dOut[n,:] / (cnt + 1e-8) : 0
__global__ void masked_mean_pool_bwd_kernel(const float* __restrict__ dOut, const int64_t* __restrict__ mask,
                                            float* __restrict__ dX, int64_t N, int S, int D) {
  pdl_prologue();
  int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int D4 = D >> 2;
  if (i4 >= N * D4) return;
  int64_t n = i4 / D4;
  int c = (int)(i4 % D4) * 4;
  int cnt = 0;
  for (int t = 0; t < S; t++) cnt += mask[n * S + t] > 0;
  float inv = 1.f / ((float)cnt + 1e-8f);
  float4 g = ldg4(dOut + n * (int64_t)D + c);
  g.x *= inv; g.y *= inv; g.z *= inv; g.w *= inv;
  for (int t = 0; t < S; t++) st4(dX + (n * S + t) * (int64_t)D + c, mask[n * S + t] > 0 ? g : f4_zero());
}

}  // namespace lk

using namespace lk;

extern "C" {

// 8 CTAs per SM (32 registers) instead of 4: these kernels wait on memory between block barriers, more resident sequences hide it
static bool pool_occ8() {
  static const bool v = !(getenv("LK_POOL_OCC") && atoi(getenv("LK_POOL_OCC")) == 4);
  return v;
}

static int pool_fwd(const RowSrc& X, const float* Hd, const float* s_part, const float* w2, const int64_t* mask, const int32_t* cu, float* out,
                    float* alpha, int64_t N, int64_t S, int64_t D, int64_t A, cudaStream_t st) {
  LK_REQUIRE(D % 4 == 0 && A % 4 == 0, LK_ERR_SHAPE, "lk_additive_pool_fwd: D=%ld, A=%ld must be multiples of 4", (long)D, (long)A);
  LK_REQUIRE(S <= MAXS, LK_ERR_SHAPE, "lk_additive_pool_fwd: S=%ld exceeds %d", (long)S, MAXS);
  LK_REQUIRE(Hd || s_part, LK_ERR_ARG, "lk_additive_pool_fwd: needs the hidden rows or their w2 row dots");
  if (N == 0) return LK_OK;
  if (pool_occ8()) LK_LAUNCH((additive_pool_fwd_kernel<8>), (unsigned)N, AT, 0, st, X, Hd, s_part, w2, mask, cu, out, alpha, (int)S, (int)D, (int)A);
  else LK_LAUNCH((additive_pool_fwd_kernel<4>), (unsigned)N, AT, 0, st, X, Hd, s_part, w2, mask, cu, out, alpha, (int)S, (int)D, (int)A);
  return check_launch("additive_pool_fwd");
}
static int pool_bwd(const RowSrc& X, const float* Hd, const float* w2, const float* alpha, const int32_t* cu, const float* dOut, float* dX,
                    const DpreOut& dpre, float* dw2_part, int64_t N, int64_t S, int64_t D, int64_t A, int accumulate_dx, cudaStream_t st) {
  LK_REQUIRE(dpre.f32 || dpre.hi, LK_ERR_ARG, "lk_additive_pool_bwd: no destination for dpre");
  LK_REQUIRE(!dpre.hi || (dpre.lo && dpre.ld % 4 == 0 && dpre.ld >= A), LK_ERR_ARG, "lk_additive_pool_bwd: dpre planes (both, pitch %% 4 == 0, >= A)");
  LK_REQUIRE(D % 4 == 0 && A % 4 == 0, LK_ERR_SHAPE, "lk_additive_pool_bwd: D=%ld, A=%ld must be multiples of 4", (long)D, (long)A);
  LK_REQUIRE(S <= MAXS, LK_ERR_SHAPE, "lk_additive_pool_bwd: S=%ld exceeds %d", (long)S, MAXS);
  if (N == 0) return LK_OK;
  if (pool_occ8())
    LK_LAUNCH((additive_pool_bwd_kernel<8>), (unsigned)N, AT, 0, st, X, Hd, w2, alpha, cu, dOut, dX, dpre, dw2_part, (int)S, (int)D, (int)A, accumulate_dx);
  else
    LK_LAUNCH((additive_pool_bwd_kernel<4>), (unsigned)N, AT, 0, st, X, Hd, w2, alpha, cu, dOut, dX, dpre, dw2_part, (int)S, (int)D, (int)A, accumulate_dx);
  return check_launch("additive_pool_bwd");
}

int lk_additive_pool_fwd(const float* X, const float* Hd, const float* w2, const int64_t* mask, const int32_t* cu, float* out,
                         float* alpha, int64_t N, int64_t S, int64_t D, int64_t A, cudaStream_t st) {
  return pool_fwd(RowSrc{X, nullptr, nullptr, 0}, Hd, nullptr, w2, mask, cu, out, alpha, N, S, D, A, st);
}

int lk_additive_pool_bwd(const float* X, const float* Hd, const float* w2, const float* alpha, const int32_t* cu, const float* dOut,
                         float* dX, float* dpre, float* dw2_part, int64_t N, int64_t S, int64_t D, int64_t A, int accumulate_dx,
                         cudaStream_t st) {
  return pool_bwd(RowSrc{X, nullptr, nullptr, 0}, Hd, w2, alpha, cu, dOut, dX, DpreOut{dpre, nullptr, nullptr, 0, nullptr}, dw2_part, N, S, D, A,
                  accumulate_dx, st);
}

int lk_additive_pool_fwd_planes(const void* X_hi, const void* X_lo, int64_t ldx, const float* s_part, const int32_t* cu, float* out, float* alpha,
                                int64_t N, int64_t S, int64_t D, cudaStream_t st) {
  LK_REQUIRE(X_hi && X_lo && ldx % 4 == 0 && ldx >= D && s_part && cu, LK_ERR_ARG, "lk_additive_pool_fwd_planes: packed rows as planes + row-dot partials");
  return pool_fwd(RowSrc{nullptr, (const __nv_bfloat16*)X_hi, (const __nv_bfloat16*)X_lo, ldx}, nullptr, s_part, nullptr, nullptr, cu, out, alpha, N, S, D, 4,
                  st);
}

int lk_additive_pool_bwd_planes(const void* X_hi, const void* X_lo, int64_t ldx, const float* Hd, const float* w2, const float* alpha, const int32_t* cu,
                                const float* dOut, float* dX, float* dpre, void* dpre_hi, void* dpre_lo, int64_t ld_dpre, float* dpre_colsum_part,
                                float* dw2_part, int64_t N, int64_t S, int64_t D, int64_t A, cudaStream_t st) {
  LK_REQUIRE(X_hi && X_lo && ldx % 4 == 0 && ldx >= D && cu, LK_ERR_ARG, "lk_additive_pool_bwd_planes: packed rows as planes");
  LK_REQUIRE(!dpre_colsum_part || dpre_hi, LK_ERR_ARG, "lk_additive_pool_bwd_planes: the column-sum partials ride on the plane output");
  return pool_bwd(RowSrc{nullptr, (const __nv_bfloat16*)X_hi, (const __nv_bfloat16*)X_lo, ldx}, Hd, w2, alpha, cu, dOut, dX,
                  DpreOut{dpre, (__nv_bfloat16*)dpre_hi, (__nv_bfloat16*)dpre_lo, ld_dpre, dpre_colsum_part}, dw2_part, N, S, D, A, 0, st);
}

int lk_masked_pool(const float* X, const int64_t* mask, float* out, int64_t N, int64_t S, int64_t D, int mode, cudaStream_t st) {
  LK_REQUIRE(D % 4 == 0, LK_ERR_SHAPE, "lk_masked_pool: D=%ld must be a multiple of 4", (long)D);
  LK_REQUIRE(mode == 0 || mode == 1, LK_ERR_ARG, "lk_masked_pool: mode must be 0 (mean) or 1 (max)");
  if (N == 0) return LK_OK;
  int64_t total = N * (D / 4);
  if (mode == 0) LK_LAUNCH((masked_pool_kernel<0>), (unsigned)((total + 255) / 256), 256, 0, st, X, mask, out, N, (int)S, (int)D);
  else LK_LAUNCH((masked_pool_kernel<1>), (unsigned)((total + 255) / 256), 256, 0, st, X, mask, out, N, (int)S, (int)D);
  return check_launch("masked_pool");
}

int lk_masked_mean_pool_bwd(const float* dOut, const int64_t* mask, float* dX, int64_t N, int64_t S, int64_t D, cudaStream_t st) {
  LK_REQUIRE(D % 4 == 0, LK_ERR_SHAPE, "lk_masked_mean_pool_bwd: D=%ld must be a multiple of 4", (long)D);
  if (N == 0) return LK_OK;
  int64_t total = N * (D / 4);
  LK_LAUNCH((masked_mean_pool_bwd_kernel), (unsigned)((total + 255) / 256), 256, 0, st, dOut, mask, dX, N, (int)S, (int)D);
  return check_launch("masked_mean_pool_bwd");
}

}  // extern "C"
