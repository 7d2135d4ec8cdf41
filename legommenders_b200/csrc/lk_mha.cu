// Multi-head self-attention core for the NRMS item / user encoders (north_star piece 2).
//
// Reference: nn.MultiheadAttention(batch_first) as called in model/operators/attention_operator.py:49-55
// (SURVEY Appendix C): q,k,v are the three D-wide column blocks of qkv = x·in_projᵀ + b; per head
// logits = (q·dh^-0.5)·kᵀ, -inf on padded keys, softmax over keys, dropout(p) on the probabilities in
// training, ctx = probs·v.
//
// Sequences are short (<= 33 item tokens, <= 50 history items) and there are thousands of them, so ONE CTA owns
// ONE sequence with all its heads: the K and V tiles [L, D] are staged in shared memory with coalesced loads and
// every thread owns one (head, query) pair — its q row and its output accumulator live in registers, keys are
// walked with warp-broadcast 16-byte shared loads, and the softmax needs no cross-lane traffic at all.
// The backward recomputes probabilities from the saved row log-sum-exp (no [N,H,S,S] tensor is ever stored):
// phase A (thread = head,query) produces D_i and dQ, phase B (thread = head,key) produces dK and dV.
//
// Two sequence layouts are served: dense [N, S, *] with a key-validity mask (the reference's padded layout) and
// packed rows with cumulative offsets `cu` (padding-free execution: pad tokens and pad history slots, which the
// reference computes and then masks away, are never touched).
#include "lk_common.cuh"
#include "../../include/legommenders_b200.h"

namespace lk {

struct SeqView {
  const int* cu;         // [N+1] packed row offsets, or null for the dense layout
  const int64_t* mask;   // dense layout: [N,S] key validity (null = all valid)
  int S;                 // dense sequence length / upper bound on packed lengths
};

__device__ __forceinline__ void seq_range(const SeqView& v, int64_t n, int64_t& row0, int& L) {
  if (v.cu) { row0 = v.cu[n]; L = v.cu[n + 1] - v.cu[n]; }
  else { row0 = n * v.S; L = v.S; }
}

struct MhaParams {
  const float* qkv;      // [rows, 3D]
  float* ctx;            // [rows, D]
  float* lse;            // [rows, H]
  const float* dctx;     // bwd: [rows, D]
  float* dqkv;           // bwd: [rows, 3D]
  SeqView seq;
  int D, H;
  float scale, drop_p;
  unsigned long long seed;
};

template <int DH>
__device__ __forceinline__ float dot_smem(const float (&q)[DH], const float* __restrict__ k) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;   // four independent chains: the dot is latency-, not issue-bound
#pragma unroll
  for (int d = 0; d < DH; d += 4) {
    const float4 kv = *reinterpret_cast<const float4*>(k + d);
    s0 = fmaf(q[d], kv.x, s0); s1 = fmaf(q[d + 1], kv.y, s1); s2 = fmaf(q[d + 2], kv.z, s2); s3 = fmaf(q[d + 3], kv.w, s3);
  }
  return (s0 + s1) + (s2 + s3);
}

// ------------------------------------------------------------------------------------------------ forward
template <int DH>
__global__ void __launch_bounds__(512) mha_fwd_seq_kernel(MhaParams p) {
  extern __shared__ __align__(16) float smem[];
  const int64_t n = blockIdx.x;
  int64_t row0; int L;
  seq_range(p.seq, n, row0, L);
  if (L == 0) return;
  const int D = p.D, H = p.H;
  float* Ks = smem;                       // [L][D]
  float* Vs = Ks + (size_t)p.seq.S * D;   // [L][D]
  float* valid = Vs + (size_t)p.seq.S * D;   // [L] 1/0

  const float* base = p.qkv + row0 * 3 * (int64_t)D;
  for (int idx = threadIdx.x; idx < L * (D / 4); idx += blockDim.x) {
    const int t = idx / (D / 4), c = (idx - t * (D / 4)) * 4;
    const float* r = base + (int64_t)t * 3 * D + c;
    *reinterpret_cast<float4*>(Ks + t * D + c) = ldg4(r + D);
    *reinterpret_cast<float4*>(Vs + t * D + c) = ldg4(r + 2 * D);
  }
  for (int t = threadIdx.x; t < L; t += blockDim.x)
    valid[t] = (p.seq.cu || !p.seq.mask || p.seq.mask[n * p.seq.S + t] > 0) ? 1.f : 0.f;
  __syncthreads();

  const float inv_keep = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
  for (int w = threadIdx.x; w < L * H; w += blockDim.x) {
    const int h = w / L, i = w - h * L;
    float q[DH], acc[DH];
    const float* qr = base + (int64_t)i * 3 * D + h * DH;
#pragma unroll
    for (int d = 0; d < DH; d += 4) {
      const float4 v = ldg4(qr + d);
      q[d] = v.x * p.scale; q[d + 1] = v.y * p.scale; q[d + 2] = v.z * p.scale; q[d + 3] = v.w * p.scale;
      acc[d] = acc[d + 1] = acc[d + 2] = acc[d + 3] = 0.f;
    }
    const float* kh = Ks + h * DH;
    const float* vh = Vs + h * DH;
    float mx = -INFINITY;
    for (int j = 0; j < L; j++)
      if (valid[j] != 0.f) mx = fmaxf(mx, dot_smem<DH>(q, kh + j * D));
    float l = 0.f;
    const uint64_t didx = (((uint64_t)(row0 + i)) * H + h) * (uint64_t)p.seq.S;
    for (int j = 0; j < L; j++) {
      if (valid[j] == 0.f) continue;
      float pr = expf(dot_smem<DH>(q, kh + j * D) - mx);
      l += pr;
      if (p.drop_p > 0.f) pr *= dropout_scale(p.seed, didx + j, p.drop_p, inv_keep);
      const float* vr = vh + j * D;
#pragma unroll
      for (int d = 0; d < DH; d += 4) {
        const float4 vv = *reinterpret_cast<const float4*>(vr + d);
        acc[d] = fmaf(pr, vv.x, acc[d]); acc[d + 1] = fmaf(pr, vv.y, acc[d + 1]);
        acc[d + 2] = fmaf(pr, vv.z, acc[d + 2]); acc[d + 3] = fmaf(pr, vv.w, acc[d + 3]);
      }
    }
    // all keys masked: mx = -inf, l = 0 -> 0 * inf = NaN, as torch's softmax over an all -inf row
    const float inv = 1.f / l;
    float* o = p.ctx + (row0 + i) * (int64_t)D + h * DH;
#pragma unroll
    for (int d = 0; d < DH; d += 4) st4(o + d, make_float4(acc[d] * inv, acc[d + 1] * inv, acc[d + 2] * inv, acc[d + 3] * inv));
    p.lse[(row0 + i) * H + h] = mx + logf(l);
  }
}

// ------------------------------------------------------------------------------------------------ backward
template <int DH>
__global__ void __launch_bounds__(416) mha_bwd_seq_kernel(MhaParams p) {
  extern __shared__ __align__(16) float smem[];
  const int64_t n = blockIdx.x;
  int64_t row0; int L;
  seq_range(p.seq, n, row0, L);
  if (L == 0) return;
  const int D = p.D, H = p.H, S = p.seq.S;
  float* Qs = smem;                        // [L][D] raw q
  float* Ks = Qs + (size_t)S * D;
  float* Vs = Ks + (size_t)S * D;
  float* Gs = Vs + (size_t)S * D;          // dctx
  float* lse_s = Gs + (size_t)S * D;       // [H][S]
  float* Di_s = lse_s + (size_t)H * S;     // [H][S]
  float* valid = Di_s + (size_t)H * S;     // [S]

  const float* base = p.qkv + row0 * 3 * (int64_t)D;
  const float* gbase = p.dctx + row0 * (int64_t)D;
  for (int idx = threadIdx.x; idx < L * (D / 4); idx += blockDim.x) {
    const int t = idx / (D / 4), c = (idx - t * (D / 4)) * 4;
    const float* r = base + (int64_t)t * 3 * D + c;
    *reinterpret_cast<float4*>(Qs + t * D + c) = ldg4(r);
    *reinterpret_cast<float4*>(Ks + t * D + c) = ldg4(r + D);
    *reinterpret_cast<float4*>(Vs + t * D + c) = ldg4(r + 2 * D);
    *reinterpret_cast<float4*>(Gs + t * D + c) = ldg4(gbase + (int64_t)t * D + c);
  }
  for (int w = threadIdx.x; w < L * H; w += blockDim.x) {
    const int h = w / L, i = w - h * L;
    lse_s[h * S + i] = p.lse[(row0 + i) * H + h];
  }
  for (int t = threadIdx.x; t < L; t += blockDim.x)
    valid[t] = (p.seq.cu || !p.seq.mask || p.seq.mask[n * S + t] > 0) ? 1.f : 0.f;
  __syncthreads();

  const float inv_keep = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
  float* dbase = p.dqkv + row0 * 3 * (int64_t)D;

  // ---- phase A: thread = (head, query): D_i = sum_j p_ij dP_ij ; dQ_i = scale * sum_j p_ij (dP_ij - D_i) K_j
  for (int w = threadIdx.x; w < L * H; w += blockDim.x) {
    const int h = w / L, i = w - h * L;
    float q[DH], g[DH], dq[DH];
#pragma unroll
    for (int d = 0; d < DH; d++) {
      q[d] = Qs[i * D + h * DH + d] * p.scale;
      g[d] = Gs[i * D + h * DH + d];
      dq[d] = 0.f;
    }
    const float lse_i = lse_s[h * S + i];
    const float* kh = Ks + h * DH;
    const float* vh = Vs + h * DH;
    const uint64_t didx = (((uint64_t)(row0 + i)) * H + h) * (uint64_t)S;
    float Di = 0.f;
    for (int j = 0; j < L; j++) {
      if (valid[j] == 0.f) continue;
      const float pr = expf(dot_smem<DH>(q, kh + j * D) - lse_i);
      float dP = dot_smem<DH>(g, vh + j * D);
      if (p.drop_p > 0.f) dP *= dropout_scale(p.seed, didx + j, p.drop_p, inv_keep);
      Di = fmaf(pr, dP, Di);
    }
    Di_s[h * S + i] = Di;
    for (int j = 0; j < L; j++) {
      if (valid[j] == 0.f) continue;
      const float* kr = kh + j * D;
      const float pr = expf(dot_smem<DH>(q, kr) - lse_i);
      float dP = dot_smem<DH>(g, vh + j * D);
      if (p.drop_p > 0.f) dP *= dropout_scale(p.seed, didx + j, p.drop_p, inv_keep);
      const float dS = pr * (dP - Di);
#pragma unroll
      for (int d = 0; d < DH; d += 4) {
        const float4 kv = *reinterpret_cast<const float4*>(kr + d);
        dq[d] = fmaf(dS, kv.x, dq[d]); dq[d + 1] = fmaf(dS, kv.y, dq[d + 1]);
        dq[d + 2] = fmaf(dS, kv.z, dq[d + 2]); dq[d + 3] = fmaf(dS, kv.w, dq[d + 3]);
      }
    }
    float* o = dbase + (int64_t)i * 3 * D + h * DH;
#pragma unroll
    for (int d = 0; d < DH; d += 4)
      st4(o + d, make_float4(dq[d] * p.scale, dq[d + 1] * p.scale, dq[d + 2] * p.scale, dq[d + 3] * p.scale));
  }
  __syncthreads();

  // ---- phase B: thread = (head, key), two register-light passes:
  //      B1: dV_j = sum_i (p_ij * drop_ij) dO_i        B2: dK_j = scale * sum_i p_ij (dP_ij - D_i) q_i
  for (int w = threadIdx.x; w < L * H; w += blockDim.x) {
    const int h = w / L, j = w - h * L;
    float k[DH], acc[DH];
#pragma unroll
    for (int d = 0; d < DH; d++) { k[d] = Ks[j * D + h * DH + d] * p.scale; acc[d] = 0.f; }
    const float* qh = Qs + h * DH;
    const float* gh = Gs + h * DH;
    const bool ok = valid[j] != 0.f;
    if (ok) {
      for (int i = 0; i < L; i++) {
        const float* gr = gh + i * D;
        float pd = expf(dot_smem<DH>(k, qh + i * D) - lse_s[h * S + i]);
        if (p.drop_p > 0.f) pd *= dropout_scale(p.seed, (((uint64_t)(row0 + i)) * H + h) * (uint64_t)S + j, p.drop_p, inv_keep);
#pragma unroll
        for (int d = 0; d < DH; d += 4) {
          const float4 gv = *reinterpret_cast<const float4*>(gr + d);
          acc[d] = fmaf(pd, gv.x, acc[d]); acc[d + 1] = fmaf(pd, gv.y, acc[d + 1]);
          acc[d + 2] = fmaf(pd, gv.z, acc[d + 2]); acc[d + 3] = fmaf(pd, gv.w, acc[d + 3]);
        }
      }
    }
    float* o = dbase + (int64_t)j * 3 * D + h * DH;
#pragma unroll
    for (int d = 0; d < DH; d += 4) st4(o + 2 * D + d, make_float4(acc[d], acc[d + 1], acc[d + 2], acc[d + 3]));

    float v[DH];
#pragma unroll
    for (int d = 0; d < DH; d++) { v[d] = Vs[j * D + h * DH + d]; acc[d] = 0.f; }
    if (ok) {
      for (int i = 0; i < L; i++) {
        const float* qr = qh + i * D;
        const float pr = expf(dot_smem<DH>(k, qr) - lse_s[h * S + i]);
        float dP = dot_smem<DH>(v, gh + i * D);
        if (p.drop_p > 0.f) dP *= dropout_scale(p.seed, (((uint64_t)(row0 + i)) * H + h) * (uint64_t)S + j, p.drop_p, inv_keep);
        const float dS = pr * (dP - Di_s[h * S + i]) * p.scale;
#pragma unroll
        for (int d = 0; d < DH; d += 4) {
          const float4 qv = *reinterpret_cast<const float4*>(qr + d);
          acc[d] = fmaf(dS, qv.x, acc[d]); acc[d + 1] = fmaf(dS, qv.y, acc[d + 1]);
          acc[d + 2] = fmaf(dS, qv.z, acc[d + 2]); acc[d + 3] = fmaf(dS, qv.w, acc[d + 3]);
        }
      }
    }
#pragma unroll
    for (int d = 0; d < DH; d += 4) st4(o + D + d, make_float4(acc[d], acc[d + 1], acc[d + 2], acc[d + 3]));
  }
}

template <int DH>
static int launch_fwd(const MhaParams& p, int64_t N, int threads, cudaStream_t st) {
  size_t smem = ((size_t)2 * p.seq.S * p.D + p.seq.S) * sizeof(float);
  LK_REQUIRE(smem <= 227 * 1024, LK_ERR_SHAPE, "lk_mha_fwd: K/V tiles (%zu B) do not fit shared memory", smem);
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(mha_fwd_seq_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); attr = true; }
  mha_fwd_seq_kernel<DH><<<(unsigned)N, threads, smem, st>>>(p);
  return check_launch("mha_fwd");
}
template <int DH>
static int launch_bwd(const MhaParams& p, int64_t N, int threads, cudaStream_t st) {
  size_t smem = ((size_t)4 * p.seq.S * p.D + 2 * (size_t)p.H * p.seq.S + p.seq.S) * sizeof(float);
  LK_REQUIRE(smem <= 227 * 1024, LK_ERR_SHAPE, "lk_mha_bwd: tiles (%zu B) do not fit shared memory", smem);
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(mha_bwd_seq_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); attr = true; }
  mha_bwd_seq_kernel<DH><<<(unsigned)N, threads, smem, st>>>(p);
  return check_launch("mha_bwd");
}

static int pick_threads(int64_t S, int64_t H) {
  int64_t t = (S * H + 31) / 32 * 32;
  return (int)(t < 64 ? 64 : (t > 416 ? 416 : t));
}

}  // namespace lk

using namespace lk;

extern "C" {

int lk_mha_fwd(const float* qkv, const int64_t* mask, const int32_t* cu, float* ctx, float* lse, int64_t N, int64_t S, int64_t D,
               int64_t H, float drop_p, uint64_t seed, cudaStream_t st) {
  LK_REQUIRE(H > 0 && D % H == 0 && D % 4 == 0, LK_ERR_SHAPE, "lk_mha_fwd: D=%ld not divisible by heads=%ld", (long)D, (long)H);
  if (N == 0 || S == 0) return LK_OK;
  const int dh = (int)(D / H);
  MhaParams p{qkv, ctx, lse, nullptr, nullptr, {cu, mask, (int)S}, (int)D, (int)H, 1.0f / sqrtf((float)dh), drop_p, (unsigned long long)seed};
  const int threads = pick_threads(S, H);
  switch (dh) {
    case 8: return launch_fwd<8>(p, N, threads, st);
    case 16: return launch_fwd<16>(p, N, threads, st);
    case 32: return launch_fwd<32>(p, N, threads, st);
    case 64: return launch_fwd<64>(p, N, threads, st);
  }
  LK_REQUIRE(false, LK_ERR_SHAPE, "lk_mha_fwd: head dim %d not in {8,16,32,64}", dh);
}

int lk_mha_bwd(const float* qkv, const int64_t* mask, const int32_t* cu, const float* lse, const float* dctx, float* dqkv, int64_t N,
               int64_t S, int64_t D, int64_t H, float drop_p, uint64_t seed, cudaStream_t st) {
  LK_REQUIRE(H > 0 && D % H == 0 && D % 4 == 0, LK_ERR_SHAPE, "lk_mha_bwd: D=%ld not divisible by heads=%ld", (long)D, (long)H);
  if (N == 0 || S == 0) return LK_OK;
  const int dh = (int)(D / H);
  MhaParams p{qkv, nullptr, const_cast<float*>(lse), dctx, dqkv, {cu, mask, (int)S}, (int)D, (int)H, 1.0f / sqrtf((float)dh), drop_p,
              (unsigned long long)seed};
  const int threads = pick_threads(S, H);
  switch (dh) {
    case 8: return launch_bwd<8>(p, N, threads, st);
    case 16: return launch_bwd<16>(p, N, threads, st);
    case 32: return launch_bwd<32>(p, N, threads, st);
    case 64: return launch_bwd<64>(p, N, threads, st);
  }
  LK_REQUIRE(false, LK_ERR_SHAPE, "lk_mha_bwd: head dim %d not in {8,16,32,64}", dh);
}

}  // extern "C"
