// Multi-head self-attention core for the NRMS item / user encoders (north_star piece 2).
//
// Reference: nn.MultiheadAttention(batch_first) as called in model/operators/attention_operator.py:49-55
// (SURVEY Appendix C): q,k,v are the three D-wide column blocks of qkv = x·in_projᵀ + b; per head
// logits = (q·dh^-0.5)·kᵀ, -inf on padded keys (mask<=0), softmax over keys, dropout(p) on the
// probabilities in training, ctx = probs·v.  Sequences are short (S = 33 / 50), so one warp owns one
// (sequence, head): K/V/Q tiles live in shared memory, keys are spread over lanes for the softmax and
// head-dim over lanes for the PV product.  The backward recomputes the probabilities from the saved
// row log-sum-exp instead of storing [N,H,S,S].
#include "lk_common.cuh"
#include "../../include/legommenders_b200.h"

namespace lk {

constexpr int MAX_KPL = 4;  // keys per lane -> S <= 128

struct MhaParams {
  const float* qkv;      // [N,S,3D]
  const int64_t* mask;   // [N,S] key validity
  float* ctx;            // [N,S,D]
  float* lse;            // [N,H,S]
  int64_t N;
  int S, D, H, dh;
  float scale;
  float drop_p;
  unsigned long long seed;
};

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) mha_fwd_kernel(MhaParams p) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t task = (int64_t)blockIdx.x * WARPS + w;
  if (task >= p.N * p.H) return;
  const int64_t n = task / p.H;
  const int h = (int)(task % p.H);
  const int S = p.S, dh = p.dh, ld = dh + 1;
  float* Qs = smem + (size_t)w * (3 * S * ld + S);
  float* Ks = Qs + S * ld;
  float* Vs = Ks + S * ld;
  float* Ps = Vs + S * ld;

  const float* base = p.qkv + n * S * 3 * (int64_t)p.D + h * dh;
  for (int idx = lane; idx < S * dh; idx += 32) {
    int t = idx / dh, d = idx - t * dh;
    const float* row = base + (int64_t)t * 3 * p.D + d;
    Qs[t * ld + d] = __ldg(row) * p.scale;
    Ks[t * ld + d] = __ldg(row + p.D);
    Vs[t * ld + d] = __ldg(row + 2 * p.D);
  }
  bool kvalid[MAX_KPL];
#pragma unroll
  for (int u = 0; u < MAX_KPL; u++) {
    int j = lane + 32 * u;
    kvalid[u] = (j < S) && (p.mask[n * S + j] > 0);
  }
  __syncwarp();
  const float inv_keep = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;

  for (int i = 0; i < S; i++) {
    float l[MAX_KPL];
    float mx = -INFINITY;
#pragma unroll
    for (int u = 0; u < MAX_KPL; u++) {
      int j = lane + 32 * u;
      l[u] = -INFINITY;
      if (j < S && kvalid[u]) {
        float a = 0.f;
        for (int d = 0; d < dh; d++) a = fmaf(Qs[i * ld + d], Ks[j * ld + d], a);
        l[u] = a;
      }
      mx = fmaxf(mx, l[u]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int u = 0; u < MAX_KPL; u++) {
      l[u] = expf(l[u] - mx);   // all-masked row: (-inf) - (-inf) = NaN, as torch does
      if (lane + 32 * u < S) sum += l[u];
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
#pragma unroll
    for (int u = 0; u < MAX_KPL; u++) {
      int j = lane + 32 * u;
      if (j < S) {
        float pr = l[u] * inv;
        if (p.drop_p > 0.f) pr *= dropout_scale(p.seed, ((uint64_t)task * S + i) * S + j, p.drop_p, inv_keep);
        Ps[j] = pr;
      }
    }
    if (lane == 0) p.lse[task * S + i] = mx + logf(sum);
    __syncwarp();
    for (int d = lane; d < dh; d += 32) {
      float a = 0.f;
      for (int j = 0; j < S; j++) a = fmaf(Ps[j], Vs[j * ld + d], a);
      p.ctx[(n * S + i) * (int64_t)p.D + h * dh + d] = a;
    }
    __syncwarp();
  }
}

struct MhaBwdParams {
  const float* qkv;      // [N,S,3D]
  const int64_t* mask;   // [N,S]
  const float* lse;      // [N,H,S]
  const float* dctx;     // [N,S,D]
  float* dqkv;           // [N,S,3D]
  int64_t N;
  int S, D, H, dh;
  float scale;
  float drop_p;
  unsigned long long seed;
};

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) mha_bwd_kernel(MhaBwdParams p) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t task = (int64_t)blockIdx.x * WARPS + w;
  if (task >= p.N * p.H) return;
  const int64_t n = task / p.H;
  const int h = (int)(task % p.H);
  const int S = p.S, dh = p.dh, ld = dh + 1;
  float* Qs = smem + (size_t)w * (6 * S * ld + S);
  float* Ks = Qs + S * ld;
  float* Vs = Ks + S * ld;
  float* Gs = Vs + S * ld;    // dctx tile
  float* dKs = Gs + S * ld;
  float* dVs = dKs + S * ld;
  float* dSs = dVs + S * ld;

  const float* base = p.qkv + n * S * 3 * (int64_t)p.D + h * dh;
  const float* gbase = p.dctx + n * S * (int64_t)p.D + h * dh;
  for (int idx = lane; idx < S * dh; idx += 32) {
    int t = idx / dh, d = idx - t * dh;
    const float* row = base + (int64_t)t * 3 * p.D + d;
    Qs[t * ld + d] = __ldg(row) * p.scale;
    Ks[t * ld + d] = __ldg(row + p.D);
    Vs[t * ld + d] = __ldg(row + 2 * p.D);
    Gs[t * ld + d] = __ldg(gbase + (int64_t)t * p.D + d);
    dKs[t * ld + d] = 0.f;
    dVs[t * ld + d] = 0.f;
  }
  bool kvalid[MAX_KPL];
#pragma unroll
  for (int u = 0; u < MAX_KPL; u++) {
    int j = lane + 32 * u;
    kvalid[u] = (j < S) && (p.mask[n * S + j] > 0);
  }
  __syncwarp();
  const float inv_keep = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
  float* dq_out = p.dqkv + n * S * 3 * (int64_t)p.D + h * dh;

  for (int i = 0; i < S; i++) {
    const float lse_i = p.lse[task * S + i];
    float pr[MAX_KPL], dP[MAX_KPL], dsc[MAX_KPL];
    float Di = 0.f;
#pragma unroll
    for (int u = 0; u < MAX_KPL; u++) {
      int j = lane + 32 * u;
      pr[u] = 0.f; dP[u] = 0.f; dsc[u] = 1.f;
      if (j < S && kvalid[u]) {
        float a = 0.f, g = 0.f;
        for (int d = 0; d < dh; d++) {
          a = fmaf(Qs[i * ld + d], Ks[j * ld + d], a);
          g = fmaf(Gs[i * ld + d], Vs[j * ld + d], g);
        }
        pr[u] = expf(a - lse_i);
        if (p.drop_p > 0.f) dsc[u] = dropout_scale(p.seed, ((uint64_t)task * S + i) * S + j, p.drop_p, inv_keep);
        dP[u] = g * dsc[u];
        Di = fmaf(pr[u], dP[u], Di);
      }
    }
    Di = warp_sum(Di);
#pragma unroll
    for (int u = 0; u < MAX_KPL; u++) {
      int j = lane + 32 * u;
      if (j < S) {
        float dS = 0.f;
        if (kvalid[u]) {
          dS = pr[u] * (dP[u] - Di);
          const float pd = pr[u] * dsc[u];
          for (int d = 0; d < dh; d++) {
            dVs[j * ld + d] = fmaf(pd, Gs[i * ld + d], dVs[j * ld + d]);
            dKs[j * ld + d] = fmaf(dS, Qs[i * ld + d], dKs[j * ld + d]);   // Qs already carries the dh^-0.5 scale
          }
        }
        dSs[j] = dS;
      }
    }
    __syncwarp();
    for (int d = lane; d < dh; d += 32) {
      float a = 0.f;
      for (int j = 0; j < S; j++) a = fmaf(dSs[j], Ks[j * ld + d], a);
      dq_out[(int64_t)i * 3 * p.D + d] = a * p.scale;
    }
    __syncwarp();
  }
  for (int idx = lane; idx < S * dh; idx += 32) {
    int t = idx / dh, d = idx - t * dh;
    float* row = dq_out + (int64_t)t * 3 * p.D + d;
    row[p.D] = dKs[t * ld + d];
    row[2 * p.D] = dVs[t * ld + d];
  }
}

}  // namespace lk

using namespace lk;

extern "C" {

int lk_mha_fwd(const float* qkv, const int64_t* mask, float* ctx, float* lse, int64_t N, int64_t S, int64_t D, int64_t H,
               float drop_p, uint64_t seed, cudaStream_t st) {
  LK_REQUIRE(H > 0 && D % H == 0, LK_ERR_SHAPE, "lk_mha_fwd: D=%ld not divisible by heads=%ld", (long)D, (long)H);
  LK_REQUIRE(S <= 32 * MAX_KPL, LK_ERR_SHAPE, "lk_mha_fwd: S=%ld exceeds %d", (long)S, 32 * MAX_KPL);
  if (N == 0) return LK_OK;
  const int dh = (int)(D / H);
  constexpr int WARPS = 4;
  size_t smem = (size_t)WARPS * (3 * S * (dh + 1) + S) * sizeof(float);
  LK_REQUIRE(smem <= 227 * 1024, LK_ERR_SHAPE, "lk_mha_fwd: tile does not fit shared memory");
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(mha_fwd_kernel<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  MhaParams p{qkv, mask, ctx, lse, N, (int)S, (int)D, (int)H, dh, 1.0f / sqrtf((float)dh), drop_p, (unsigned long long)seed};
  int64_t tasks = N * H;
  mha_fwd_kernel<WARPS><<<(unsigned)((tasks + WARPS - 1) / WARPS), WARPS * 32, smem, st>>>(p);
  return check_launch("mha_fwd");
}

int lk_mha_bwd(const float* qkv, const int64_t* mask, const float* lse, const float* dctx, float* dqkv, int64_t N, int64_t S,
               int64_t D, int64_t H, float drop_p, uint64_t seed, cudaStream_t st) {
  LK_REQUIRE(H > 0 && D % H == 0, LK_ERR_SHAPE, "lk_mha_bwd: D=%ld not divisible by heads=%ld", (long)D, (long)H);
  LK_REQUIRE(S <= 32 * MAX_KPL, LK_ERR_SHAPE, "lk_mha_bwd: S=%ld exceeds %d", (long)S, 32 * MAX_KPL);
  if (N == 0) return LK_OK;
  const int dh = (int)(D / H);
  constexpr int WARPS = 2;
  size_t smem = (size_t)WARPS * (6 * S * (dh + 1) + S) * sizeof(float);
  LK_REQUIRE(smem <= 227 * 1024, LK_ERR_SHAPE, "lk_mha_bwd: tile does not fit shared memory");
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(mha_bwd_kernel<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  MhaBwdParams p{qkv, mask, lse, dctx, dqkv, N, (int)S, (int)D, (int)H, dh, 1.0f / sqrtf((float)dh), drop_p, (unsigned long long)seed};
  int64_t tasks = N * H;
  mha_bwd_kernel<WARPS><<<(unsigned)((tasks + WARPS - 1) / WARPS), WARPS * 32, smem, st>>>(p);
  return check_launch("mha_bwd");
}

}  // extern "C"
