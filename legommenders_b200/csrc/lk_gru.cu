// GRU user encoder (LSTUR; model/operators/gru_operator.py:25-54 = nn.GRU(batch_first, 1 layer) over pack_padded_sequence -> last hidden).
//
//   r = sigmoid(gi_r + W_hr h + b_hr)   z = sigmoid(gi_z + W_hz h + b_hz)   n = tanh(gi_n + r * (W_hn h + b_hn))   h' = (1 - z) * n + z * h
// with gi = x W_ih^T + b_ih for ALL time steps computed up front by one tensor-core contraction (lk_tc_gemm); what is sequential — the
// H x 3H recurrent product and the gates — is one kernel: a CTA per sequence, a thread per hidden unit, h in shared memory, the recurrent
// weights streamed from L2 in the transposed layout [H, 3H] so that the threads of a warp read consecutive addresses.  Only the first
// `len[b]` steps of a sequence run (pack_padded_sequence semantics); the last hidden state of each sequence is the output.
// The backward kernel walks the same steps in reverse and leaves d(gi) and d(gh) per step; the four weight / bias gradients and dx are then
// ordinary contractions over [B*S, 3H] (host side: ops.gru_last_hidden).
#include "lk_common.cuh"
#include "../../include/legommenders_b200.h"

namespace lk {
namespace gru {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// gi [B, S, 3H]; whhT [H, 3H]; bhh [3H]; len [B]; outputs: last [B, H]; saved: hs [B, S, H], gates [B, S, 3H] (r, z, n), hnp [B, S, H]
__global__ void gru_fwd_kernel(const float* __restrict__ gi, const float* __restrict__ whhT, const float* __restrict__ bhh,
                               const int32_t* __restrict__ len, float* __restrict__ last, float* __restrict__ hs, float* __restrict__ gates,
                               float* __restrict__ hnp, int S, int H) {
  pdl_prologue();
  extern __shared__ float h_s[];          // [H]
  const int b = blockIdx.x, j = threadIdx.x;
  const int L = min(len[b], S);
  h_s[j] = 0.f;
  __syncthreads();
  const float br = bhh[j], bz = bhh[H + j], bn = bhh[2 * H + j];
  float h = 0.f;
  for (int t = 0; t < L; t++) {
    float ar = br, az = bz, an = bn;
#pragma unroll 4
    for (int k = 0; k < H; k++) {
      const float hk = h_s[k];
      const float* w = whhT + (size_t)k * 3 * H;
      ar = fmaf(w[j], hk, ar);
      az = fmaf(w[H + j], hk, az);
      an = fmaf(w[2 * H + j], hk, an);
    }
    const float* g = gi + ((size_t)b * S + t) * 3 * H;
    const float r = sigmoidf_(g[j] + ar), z = sigmoidf_(g[H + j] + az);
    const float n = tanhf(g[2 * H + j] + r * an);
    h = (1.f - z) * n + z * h;
    const size_t o = (size_t)b * S + t;
    gates[o * 3 * H + j] = r; gates[o * 3 * H + H + j] = z; gates[o * 3 * H + 2 * H + j] = n;
    hnp[o * H + j] = an;
    hs[o * H + j] = h;
    __syncthreads();                      // every thread has read the old h
    h_s[j] = h;
    __syncthreads();
  }
  last[(size_t)b * H + j] = h;
}

// dlast [B, H] -> dgi, dgh [B, S, 3H] (zero beyond len); whh [3H, H] in its own layout (thread k reads column k: coalesced)
__global__ void gru_bwd_kernel(const float* __restrict__ dlast, const float* __restrict__ whh, const int32_t* __restrict__ len,
                               const float* __restrict__ hs, const float* __restrict__ gates, const float* __restrict__ hnp,
                               float* __restrict__ dgi, float* __restrict__ dgh, int S, int H) {
  pdl_prologue();
  extern __shared__ float d_s[];          // [3H] d(gh) of the current step
  const int b = blockIdx.x, j = threadIdx.x;
  const int L = min(len[b], S);
  float dh = dlast[(size_t)b * H + j];
  for (int t = S - 1; t >= L; t--) {      // steps that never ran
    const size_t o = ((size_t)b * S + t) * 3 * H;
    dgi[o + j] = dgi[o + H + j] = dgi[o + 2 * H + j] = 0.f;
    dgh[o + j] = dgh[o + H + j] = dgh[o + 2 * H + j] = 0.f;
  }
  for (int t = L - 1; t >= 0; t--) {
    const size_t o = (size_t)b * S + t;
    const float r = gates[o * 3 * H + j], z = gates[o * 3 * H + H + j], n = gates[o * 3 * H + 2 * H + j];
    const float hprev = t > 0 ? hs[(o - 1) * H + j] : 0.f;
    const float dn = dh * (1.f - z) * (1.f - n * n);
    const float dz = dh * (hprev - n) * z * (1.f - z);
    const float dr = dn * hnp[o * H + j] * r * (1.f - r);
    dgi[o * 3 * H + j] = dr; dgi[o * 3 * H + H + j] = dz; dgi[o * 3 * H + 2 * H + j] = dn;
    const float dnh = dn * r;
    dgh[o * 3 * H + j] = dr; dgh[o * 3 * H + H + j] = dz; dgh[o * 3 * H + 2 * H + j] = dnh;
    __syncthreads();                      // the previous step's d_s has been consumed
    d_s[j] = dr; d_s[H + j] = dz; d_s[2 * H + j] = dnh;
    __syncthreads();
    float acc = dh * z;
#pragma unroll 4
    for (int i = 0; i < 3 * H; i++) acc = fmaf(d_s[i], whh[(size_t)i * H + j], acc);
    dh = acc;
  }
}

}  // namespace gru
}  // namespace lk

using namespace lk;

extern "C" {

int lk_gru_fwd(const float* gi, const float* whhT, const float* bhh, const int32_t* len, float* last, float* hs, float* gates, float* hnp,
               int64_t B, int64_t S, int64_t H, cudaStream_t st) {
  LK_REQUIRE(H >= 32 && H <= 1024 && H % 32 == 0, LK_ERR_SHAPE, "lk_gru_fwd: hidden size %ld (multiple of 32, <= 1024)", (long)H);
  if (B == 0) return LK_OK;
  LK_LAUNCH((gru::gru_fwd_kernel), (unsigned)B, (unsigned)H, H * sizeof(float), st, gi, whhT, bhh, len, last, hs, gates, hnp, (int)S, (int)H);
  return check_launch("gru_fwd");
}

int lk_gru_bwd(const float* dlast, const float* whh, const int32_t* len, const float* hs, const float* gates, const float* hnp, float* dgi,
               float* dgh, int64_t B, int64_t S, int64_t H, cudaStream_t st) {
  LK_REQUIRE(H >= 32 && H <= 1024 && H % 32 == 0, LK_ERR_SHAPE, "lk_gru_bwd: hidden size %ld (multiple of 32, <= 1024)", (long)H);
  if (B == 0) return LK_OK;
  LK_LAUNCH((gru::gru_bwd_kernel), (unsigned)B, (unsigned)H, 3 * H * sizeof(float), st, dlast, whh, len, hs, gates, hnp, dgi, dgh, (int)S, (int)H);
  return check_launch("gru_bwd");
}

}  // extern "C"
