// Multi-head self-attention core for the NRMS item / user encoders (north_star piece 2).
//
// Reference: nn.MultiheadAttention(batch_first) as called in model/operators/attention_operator.py:49-55
// (SURVEY Appendix C): q,k,v are the three D-wide column blocks of qkv = x·in_projᵀ + b; per head
// logits = (q·dh^-0.5)·kᵀ, -inf on padded keys, softmax over keys, dropout(p) on the probabilities in
// training, ctx = probs·v.
//
// Shape of the problem: thousands of short sequences (<= 33 item tokens, <= 50/100 history items), head dim 32.  The tiles
// are far too small for tcgen05 and fp32 parity rules out single-pass tensor-core math, so this is an fp32 CUDA-core
// kernel whose limiter is the shared-memory return path (128 B/clk/SM: one operand word per lane per FMA if nothing is
// reused).  The work split is therefore chosen to cut shared-memory words per FMA by R (2 for head dim 32: with one shuffle
// step per dot product the instruction issue rate and the shared-memory return path are then about equally loaded):
//   * one CTA owns one sequence and a group of HG heads (HG*DH = 128 columns); two [L, 128] tiles sit in shared memory;
//   * R adjacent lanes (a "quad") share a group of R rows of one head; each lane owns a W = DH/R wide slice of the head
//     dimension for ALL R rows, so one W-float shared load feeds R*W FMAs; partial dot products are completed with
//     log2(R) xor-shuffles inside the quad.
//   forward     quad = R queries:  one pass over the keys with an online softmax (K/V tiles)
//   backward A  quad = R queries:  p_ij from the saved log-sum-exp, dS_ij = p_ij (dP_ij - D_i) with D_i = dO_i·O_i, dQ_i
//   backward B  quad = R keys:     dV_j, dK_j in one pass over the queries (Q/dO tiles re-staged in the same memory)
// Probabilities are never stored; exp is ex2.approx on log2(e)-prescaled logits (rel. error 2^-22).  Results leave through a
// shared tile: coalesced stores, optional split-bf16 plane output (ctx and dqkv only feed tensor-core contractions) and
// deterministic per-sequence column sums of dqkv (the in_proj bias gradient).
//
// Two sequence layouts are served: dense [N, S, *] with a key-validity mask (the reference's padded layout) and
// packed rows with cumulative offsets `cu` (padding-free execution: pad tokens and pad history slots, which the
// reference computes and then masks away, are never touched).
#include <cuda_bf16.h>
#include <stdlib.h>

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
  const float* qkv;                 // [rows, 3D]
  float* ctx;                       // [rows, D] fp32 (fwd: written; bwd: read for D_i = dO·O)
  __nv_bfloat16 *ctx_hi, *ctx_lo;   // fwd: optional split-bf16 image of ctx, pitch D
  float* lse;                       // [rows, H]  natural-log log-sum-exp of the scaled logits
  const float* dctx;                // bwd: [rows, D]
  float* dqkv;                      // bwd: [rows, 3D] fp32 (nullable when planes are requested)
  __nv_bfloat16 *dq_hi, *dq_lo;     // bwd: optional split-bf16 image of dqkv, pitch 3D
  float* colsum_part;               // bwd: optional [N, 3D] per-sequence column sums of dqkv
  SeqView seq;
  int D, H, HG;                     // HG heads per CTA
  int ov_items;                     // work items beyond one CTA-full (their results are parked, see unpark)
  float scale, drop_p;
  uint32_t drop_thr;                // keep iff (hash >> 8) >= drop_thr, drop_thr = p * 2^24
  unsigned long long seed;
};

constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
constexpr int MHA_THREADS = 256;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// attention dropout: keep bit of probability (row, head, key j) = one 64-bit hash per (row, head), a 32-bit finaliser per key
__device__ __forceinline__ uint32_t attn_hash_base(unsigned long long seed, uint64_t row_head) {
  return mix32(seed * 0x9E3779B97F4A7C15ULL + row_head);
}
__device__ __forceinline__ bool attn_keep(uint32_t hb, int j, uint32_t thr) {
  uint32_t x = hb ^ ((uint32_t)(j + 1) * 0x9E3779B9u);
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  return (x >> 8) >= thr;
}

template <int R>
__device__ __forceinline__ float quad_sum(float v) {
  if (R >= 2) v += __shfl_xor_sync(0xffffffffu, v, 1);
  if (R >= 4) v += __shfl_xor_sync(0xffffffffu, v, 2);
  if (R >= 8) v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

template <int W>
__device__ __forceinline__ void lds_slice(float (&v)[W], const float* __restrict__ src) {
#pragma unroll
  for (int w = 0; w < W; w += 4) {
    const float4 x = *reinterpret_cast<const float4*>(src + w);
    v[w] = x.x; v[w + 1] = x.y; v[w + 2] = x.z; v[w + 3] = x.w;
  }
}
template <int W>
__device__ __forceinline__ void ldg_slice(float (&v)[W], const float* __restrict__ src, float s) {
#pragma unroll
  for (int w = 0; w < W; w += 4) {
    const float4 x = ldg4(src + w);
    v[w] = x.x * s; v[w + 1] = x.y * s; v[w + 2] = x.z * s; v[w + 3] = x.w * s;
  }
}
template <int W>
__device__ __forceinline__ float dot_slice(const float (&a)[W], const float (&b)[W]) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int w = 0; w < W; w += 2) { s0 = fmaf(a[w], b[w], s0); s1 = fmaf(a[w + 1], b[w + 1], s1); }
  return s0 + s1;
}
template <int W>
__device__ __forceinline__ void sts_slice(float* __restrict__ dst, const float (&v)[W], float s) {
#pragma unroll
  for (int w = 0; w < W; w += 4) *reinterpret_cast<float4*>(dst + w) = make_float4(v[w] * s, v[w + 1] * s, v[w + 2] * s, v[w + 3] * s);
}

// Shared tiles are [S][TP] with TP = TW + 4 (TW = HG*DH columns of this CTA): rows stay 16-byte aligned and the row-per-lane
// writes of the result tiles spread over the banks.
__device__ __forceinline__ void stage_tile(float* T, const float* __restrict__ src, int64_t ld, int L, int TW) {
  const int C4 = TW >> 2, TP = TW + 4;
  const int sh = (C4 & (C4 - 1)) == 0 ? 31 - __clz(C4) : -1;       // power-of-two tile widths: shifts instead of divisions
  for (int idx = threadIdx.x; idx < L * C4; idx += blockDim.x) {
    const int t = sh >= 0 ? idx >> sh : idx / C4, c = (idx - t * C4) * 4;
    cp_async16(T + t * TP + c, src + (int64_t)t * ld + c);
  }
}

// write a result tile to global rows (pitch ld, in elements) as fp32 and/or split-bf16 planes, 16 bytes per thread, coalesced;
// optionally its column sums (fixed row order -> deterministic)
__device__ __forceinline__ void flush_tile(const float* __restrict__ T, int L, int TW, float* __restrict__ f32, __nv_bfloat16* __restrict__ hi,
                                           __nv_bfloat16* __restrict__ lo, int64_t ld, float* __restrict__ colsum) {
  const int C4 = TW >> 2, TP = TW + 4;
  const int sh = (C4 & (C4 - 1)) == 0 ? 31 - __clz(C4) : -1;
  for (int idx = threadIdx.x; idx < L * C4; idx += blockDim.x) {
    const int t = sh >= 0 ? idx >> sh : idx / C4, c = (idx - t * C4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(T + t * TP + c);
    if (f32) st4(f32 + (int64_t)t * ld + c, v);
    if (hi) {
      __align__(8) __nv_bfloat16 h[4], l[4];
      const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; e++) {
        h[e] = __float2bfloat16_rn(x[e]);
        l[e] = __float2bfloat16_rn(x[e] - __bfloat162float(h[e]));
      }
      *reinterpret_cast<uint2*>(hi + (int64_t)t * ld + c) = *reinterpret_cast<uint2*>(h);
      *reinterpret_cast<uint2*>(lo + (int64_t)t * ld + c) = *reinterpret_cast<uint2*>(l);
    }
  }
  if (colsum) {
    for (int c = threadIdx.x; c < TW; c += blockDim.x) {
      float s = 0.f;
      for (int t = 0; t < L; t++) s += T[t * TP + c];
      colsum[c] = s;
    }
  }
}

// decomposition of a work item: (local head, row group, slice)
struct Item { int hl, g, ds; bool active; };
template <int R>
__device__ __forceinline__ Item item_of(int w, int G, int items) {
  Item it;
  it.active = w < items;
  const int wc = it.active ? w : 0;
  it.ds = wc % R;
  const int t = wc / R;
  it.g = t % G;
  it.hl = t / G;
  return it;
}
// keep bits of this quad's R (row, key) pairs: lane ds computed `mine`; returns the R bits of the quad
template <int R>
__device__ __forceinline__ uint32_t quad_bits(bool mine) {
  const uint32_t b = __ballot_sync(0xffffffffu, mine);
  return (b >> ((threadIdx.x & 31) & ~(R - 1))) & ((1u << R) - 1u);
}

// Result tiles alias operand tiles that are dead by the time the results exist (one __syncthreads in between), which is what lets a
// third backward / fifth forward CTA fit an SM.  A sequence needing more work items than the CTA has threads is processed in
// rounds, LAST round first: the later rounds park their results in a small overflow area (slot = item - blockDim), only round 0 —
// whose results are still in registers when every thread has finished reading the operands — writes the aliased tile directly.
template <int R, int W>
__device__ __forceinline__ void park_or_store(float* tile, float* ov, bool first_round, int slot, int i, int TP, int col, int r,
                                              const float (&v)[W], float scale, bool row_ok) {
  if (first_round) { if (row_ok) sts_slice<W>(tile + i * TP + col, v, scale); }
  else sts_slice<W>(ov + (slot * R + r) * W, v, scale);
}
// overflow area -> tile (call between two __syncthreads)
template <int DH, int R>
__device__ __forceinline__ void unpark(float* tile, const float* ov, int items, int G, int L, int TP) {
  constexpr int W = DH / R, W4 = W / 4;
  const int n_ov = items - (int)blockDim.x;
  for (int idx = threadIdx.x; idx < n_ov * R * W4; idx += blockDim.x) {
    const int slot = idx / (R * W4), rem = idx - slot * (R * W4), r = rem / W4, w4 = rem - r * W4;
    const Item it = item_of<R>((int)blockDim.x + slot, G, items);
    const int i = it.g * R + r;
    if (i < L)
      *reinterpret_cast<float4*>(tile + i * TP + it.hl * DH + it.ds * W + w4 * 4) = *reinterpret_cast<const float4*>(ov + (slot * R + r) * W + w4 * 4);
  }
}

// ------------------------------------------------------------------------------------------------ forward
template <int DH, int R>
__global__ void __maxnreg__(DH <= 32 ? 96 : 168) mha_fwd_seq_kernel(MhaParams p) {   // 4 CTAs of 160 threads per SM at head dim 32
  pdl_prologue();
  constexpr int W = DH / R;
  extern __shared__ __align__(16) float smem[];
  const int64_t n = blockIdx.x;
  int64_t row0; int L;
  seq_range(p.seq, n, row0, L);
  if (L == 0) return;
  const int D = p.D, H = p.H, S = p.seq.S, HG = p.HG, TW = HG * DH, TP = TW + 4;
  const int h0 = blockIdx.y * HG, c0 = h0 * DH;          // first head / first column of this CTA
  float* Ks = smem;                        // [S][TP]; becomes the result tile once every thread is done with the keys
  float* Vs = Ks + (size_t)S * TP;
  float* Os = Ks;
  float* valid = Vs + (size_t)S * TP;      // [S]
  float* Ov = valid + ((S + 3) & ~3);      // results of the rounds after the first: [items - blockDim][R][W]

  const float* base = p.qkv + row0 * 3 * (int64_t)D;
  stage_tile(Ks, base + D + c0, 3 * D, L, TW);
  stage_tile(Vs, base + 2 * D + c0, 3 * D, L, TW);
  for (int t = threadIdx.x; t < L; t += blockDim.x)
    valid[t] = (p.seq.cu || !p.seq.mask || p.seq.mask[n * S + t] > 0) ? 1.f : 0.f;
  cp_async_wait_all();
  __syncthreads();

  const float inv_keep = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
  const int G = (L + R - 1) / R, items = HG * G * R;
  const int rounds = (items + (int)blockDim.x - 1) / (int)blockDim.x;
  for (int rd = rounds - 1; rd >= 0; rd--) {
    const int w0 = rd * (int)blockDim.x;
    const bool warp_on = w0 + (int)(threadIdx.x & ~31u) < items;     // warp-uniform: idle warps skip the key loop altogether
    const Item it = item_of<R>(w0 + threadIdx.x, G, items);
    const int h = h0 + it.hl, col = it.hl * DH + it.ds * W;
    float acc[R][W], m[R], l[R];
    if (warp_on) {
      float q[R][W];
#pragma unroll
      for (int r = 0; r < R; r++) {
        const int i = min(it.g * R + r, L - 1);
        ldg_slice<W>(q[r], base + (int64_t)i * 3 * D + c0 + col, p.scale * kLog2e);
        m[r] = -INFINITY; l[r] = 0.f;
#pragma unroll
        for (int w = 0; w < W; w++) acc[r][w] = 0.f;
      }
      const int my_i = min(it.g * R + it.ds, L - 1);
      const uint32_t hb = attn_hash_base(p.seed, (uint64_t)(row0 + my_i) * H + h);
      for (int j = 0; j < L; j++) {
        if (valid[j] == 0.f) continue;                      // CTA-uniform
        float kv[W], sc[R];
        lds_slice<W>(kv, Ks + j * TP + col);
#pragma unroll
        for (int r = 0; r < R; r++) sc[r] = quad_sum<R>(dot_slice<W>(q[r], kv));
        uint32_t bits = (1u << R) - 1u;
        if (p.drop_p > 0.f) bits = quad_bits<R>(attn_keep(hb, j, p.drop_thr));
        lds_slice<W>(kv, Vs + j * TP + col);
#pragma unroll
        for (int r = 0; r < R; r++) {
          if (sc[r] > m[r]) {                               // rescale the running sums (rare after the first keys)
            const float c = ex2(m[r] - sc[r]);              // 2^-inf = 0 on the first valid key
            l[r] *= c;
#pragma unroll
            for (int w = 0; w < W; w++) acc[r][w] *= c;
            m[r] = sc[r];
          }
          float pr = ex2(sc[r] - m[r]);
          l[r] += pr;
          pr = ((bits >> r) & 1u) ? pr * inv_keep : 0.f;
#pragma unroll
          for (int w = 0; w < W; w++) acc[r][w] = fmaf(pr, kv[w], acc[r][w]);
        }
      }
    }
    if (rd == 0) __syncthreads();                           // nobody reads the key tile any more: it takes the results
    // all keys masked: l = 0 -> 0 * inf = NaN, as torch's softmax over an all -inf row
    if (warp_on && it.active) {
#pragma unroll
      for (int r = 0; r < R; r++) {
        const int i = it.g * R + r;
        park_or_store<R, W>(Os, Ov, rd == 0, w0 + (int)threadIdx.x - (int)blockDim.x, i, TP, col, r, acc[r], 1.f / l[r], i < L);
        if (i < L && r == it.ds) p.lse[(row0 + i) * H + h] = (m[r] + log2f(l[r])) * kLn2;
      }
    }
  }
  __syncthreads();
  if (rounds > 1) {
    unpark<DH, R>(Os, Ov, items, G, L, TP);
    __syncthreads();
  }
  const int64_t off = row0 * (int64_t)D + c0;
  flush_tile(Os, L, TW, p.ctx ? p.ctx + off : nullptr, p.ctx_hi ? p.ctx_hi + off : nullptr, p.ctx_lo ? p.ctx_lo + off : nullptr, D, nullptr);
}

// ------------------------------------------------------------------------------------------------ backward
// A: quad = R queries   p_ij = 2^(s2_ij - lse2_i);  dP_ij = (dO_i·v_j) keep_ij;  D_i = dO_i·O_i;  dS_ij = p_ij (dP_ij - D_i)
//                       dQ_i = scale Σ_j dS_ij k_j
// B: quad = R keys      dV_j = Σ_i p_ij keep_ij dO_i;   dK_j = scale Σ_i dS_ij q_i
// SP = true: phase A leaves Pd = p∘keep and dS in shared memory ([HG][L][LP] each) and phase B is two plain axpys per pair
// (160 instead of 224 FMAs per pair, one exp and one hash per pair); SP = false recomputes them in phase B (long sequences
// whose L x L matrices do not fit shared memory).
template <int DH, int R, bool SP>
__global__ void __maxnreg__(DH <= 32 ? (SP ? 168 : 200) : 255) mha_bwd_seq_kernel(MhaParams p) {
  pdl_prologue();
  constexpr int W = DH / R;
  extern __shared__ __align__(16) float smem[];
  const int64_t n = blockIdx.x;
  int64_t row0; int L;
  seq_range(p.seq, n, row0, L);
  if (L == 0) return;
  const int D = p.D, H = p.H, S = p.seq.S, HG = p.HG, TW = HG * DH, TP = TW + 4;
  const int h0 = blockIdx.y * HG, c0 = h0 * DH;
  float* T0 = smem;                        // A: K              B: Q, then the dK results
  float* T1 = T0 + (size_t)S * TP;         // A: V, then dQ     B: dO, then the dV results
  float* lse_s = T1 + (size_t)S * TP;      // [HG][S]  log2-domain log-sum-exp
  float* Di_s = lse_s + (size_t)HG * S;    // [HG][S]
  uint32_t* hb_s = reinterpret_cast<uint32_t*>(Di_s + (size_t)HG * S);   // [HG][S] dropout hash of (row, head)
  float* valid = reinterpret_cast<float*>(hb_s + (size_t)HG * S);        // [S]
  const int LP = (S + 1) & ~1;                                           // even row pitch: phase B reads key pairs as float2
  float* Pd_s = valid + ((S + 3) & ~3);                                  // SP: [HG][S][LP]  p * keep
  float* dS_s = Pd_s + (SP ? (size_t)HG * S * LP : 0);                   // SP: [HG][S][LP]  dS * scale
  float* Ov0 = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(dS_s + (SP ? (size_t)HG * S * LP : 0)) + 15) & ~(uintptr_t)15);
                                                                         // results of the rounds after the first (see unpark)
  float* Ov1 = Ov0 + (size_t)p.ov_items * R * W;

  const float* base = p.qkv + row0 * 3 * (int64_t)D;
  const float* gbase = p.dctx + row0 * (int64_t)D;
  const float* obase = p.ctx + row0 * (int64_t)D;
  stage_tile(T0, base + D + c0, 3 * D, L, TW);
  stage_tile(T1, base + 2 * D + c0, 3 * D, L, TW);
  for (int w = threadIdx.x; w < HG * L; w += blockDim.x) {
    const int hl = w / L, i = w - hl * L;
    lse_s[hl * S + i] = p.lse[(row0 + i) * H + h0 + hl] * kLog2e;
    hb_s[hl * S + i] = attn_hash_base(p.seed, (uint64_t)(row0 + i) * H + h0 + hl);
  }
  for (int t = threadIdx.x; t < L; t += blockDim.x)
    valid[t] = (p.seq.cu || !p.seq.mask || p.seq.mask[n * S + t] > 0) ? 1.f : 0.f;
  cp_async_wait_all();
  __syncthreads();

  const float inv_keep = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
  const int G = (L + R - 1) / R, items = HG * G * R;
  const int rounds = (items + (int)blockDim.x - 1) / (int)blockDim.x;
  float* dbase = p.dqkv ? p.dqkv + row0 * 3 * (int64_t)D + c0 : nullptr;
  __nv_bfloat16* hbase = p.dq_hi ? p.dq_hi + row0 * 3 * (int64_t)D + c0 : nullptr;
  __nv_bfloat16* lbase = p.dq_lo ? p.dq_lo + row0 * 3 * (int64_t)D + c0 : nullptr;
  float* csum = p.colsum_part ? p.colsum_part + n * 3 * (int64_t)D + c0 : nullptr;

  // ---- phase A ---------------------------------------------------------------------------------------------------
  for (int rd = rounds - 1; rd >= 0; rd--) {
    const int w0 = rd * (int)blockDim.x;
    const bool warp_on = w0 + (int)(threadIdx.x & ~31u) < items;     // warp-uniform: idle warps skip the key loop
    const Item it = item_of<R>(w0 + threadIdx.x, G, items);
    const int col = it.hl * DH + it.ds * W;
    float dq[R][W];
    if (warp_on) {
      float q[R][W], g[R][W], Di[R], lse2[R];
#pragma unroll
      for (int r = 0; r < R; r++) {
        const int i = min(it.g * R + r, L - 1);
        float o[W];
        ldg_slice<W>(q[r], base + (int64_t)i * 3 * D + c0 + col, p.scale * kLog2e);
        ldg_slice<W>(g[r], gbase + (int64_t)i * D + c0 + col, 1.f);
        ldg_slice<W>(o, obase + (int64_t)i * D + c0 + col, 1.f);
        Di[r] = quad_sum<R>(dot_slice<W>(g[r], o));
        lse2[r] = lse_s[it.hl * S + i];
#pragma unroll
        for (int w = 0; w < W; w++) dq[r][w] = 0.f;
      }
#pragma unroll
      for (int r = 0; r < R; r++) {                       // lane ds publishes row ds (static register indexing)
        const int i = it.g * R + r;
        if (r == it.ds && it.active && i < L) Di_s[it.hl * S + i] = Di[r];
      }
      const uint32_t hb = hb_s[it.hl * S + min(it.g * R + it.ds, L - 1)];
      for (int j = 0; j < L; j++) {
        if (valid[j] == 0.f) continue;
        float kv[W], vv[W], ps[R], pd[R];
        lds_slice<W>(kv, T0 + j * TP + col);
        lds_slice<W>(vv, T1 + j * TP + col);
#pragma unroll
        for (int r = 0; r < R; r++) {
          ps[r] = quad_sum<R>(dot_slice<W>(q[r], kv));
          pd[r] = quad_sum<R>(dot_slice<W>(g[r], vv));
        }
        uint32_t bits = (1u << R) - 1u;
        if (p.drop_p > 0.f) bits = quad_bits<R>(attn_keep(hb, j, p.drop_thr));
#pragma unroll
        for (int r = 0; r < R; r++) {
          const float pr = ex2(ps[r] - lse2[r]);
          const bool kept = (bits >> r) & 1u;
          const float dP = kept ? pd[r] * inv_keep : 0.f;
          const float dS = pr * (dP - Di[r]);
#pragma unroll
          for (int w = 0; w < W; w++) dq[r][w] = fmaf(dS, kv[w], dq[r][w]);
          if (SP && r == it.ds && it.active) {              // lane ds publishes row ds of the quad
            const int i = min(it.g * R + r, L - 1);
            Pd_s[(it.hl * S + i) * LP + j] = kept ? pr * inv_keep : 0.f;
            dS_s[(it.hl * S + i) * LP + j] = dS * p.scale;
          }
        }
      }
    }
    if (rd == 0) __syncthreads();                           // K and V are dead: the V tile takes dQ
    if (warp_on && it.active) {
#pragma unroll
      for (int r = 0; r < R; r++) {
        const int i = it.g * R + r;
        park_or_store<R, W>(T1, Ov0, rd == 0, w0 + (int)threadIdx.x - (int)blockDim.x, i, TP, col, r, dq[r], p.scale, i < L);
      }
    }
  }
  __syncthreads();
  stage_tile(T0, base + c0, 3 * D, L, TW);                  // phase B's Q tile streams in while dQ is flushed
  if (rounds > 1) {
    unpark<DH, R>(T1, Ov0, items, G, L, TP);
    __syncthreads();
  }
  flush_tile(T1, L, TW, dbase, hbase, lbase, 3 * D, csum);
  __syncthreads();
  // ---- phase B: Q and dO tiles over K and V --------------------------------------------------------------------------
  stage_tile(T1, gbase + c0, D, L, TW);
  cp_async_wait_all();
  __syncthreads();
  for (int rd = rounds - 1; rd >= 0; rd--) {
    const int w0 = rd * (int)blockDim.x;
    const bool warp_on = w0 + (int)(threadIdx.x & ~31u) < items;
    const Item it = item_of<R>(w0 + threadIdx.x, G, items);
    const int col = it.hl * DH + it.ds * W;
    float dk[R][W], dv[R][W];
    float dk_scale = 1.f;
    if (warp_on) {
      if (SP) {
        // phase A skipped masked keys: their columns of Pd / dS were never written, and rows of a quad beyond L were clamped
        bool kok[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
          const int j = it.g * R + r;
          kok[r] = j < L && valid[min(j, L - 1)] != 0.f;
#pragma unroll
          for (int w = 0; w < W; w++) { dk[r][w] = 0.f; dv[r][w] = 0.f; }
        }
        const float* pd_h = Pd_s + (it.hl * S * LP + it.g * R);
        const float* ds_h = dS_s + (it.hl * S * LP + it.g * R);
        const float* qp = T0 + col;
        const float* gp = T1 + col;
#pragma unroll 2
        for (int i = 0; i < L; i++, pd_h += LP, ds_h += LP, qp += TP, gp += TP) {
          float qv[W], gv[W], pk[R], dsv[R];
          lds_slice<W>(qv, qp);
          lds_slice<W>(gv, gp);
          if (R == 2) {
            const float2 a = *reinterpret_cast<const float2*>(pd_h), b = *reinterpret_cast<const float2*>(ds_h);
            pk[0] = a.x; pk[R - 1] = a.y; dsv[0] = b.x; dsv[R - 1] = b.y;
          } else {
#pragma unroll
            for (int r = 0; r < R; r++) { pk[r] = pd_h[r]; dsv[r] = ds_h[r]; }
          }
#pragma unroll
          for (int r = 0; r < R; r++) {
            const float a = kok[r] ? pk[r] : 0.f, b = kok[r] ? dsv[r] : 0.f;
#pragma unroll
            for (int w = 0; w < W; w++) {
              dv[r][w] = fmaf(a, gv[w], dv[r][w]);
              dk[r][w] = fmaf(b, qv[w], dk[r][w]);
            }
          }
        }
        // dS_s already carries the scale
      } else {
        float k[R][W], v[R][W];
        uint32_t kvalid = 0;
#pragma unroll
        for (int r = 0; r < R; r++) {
          const int j = it.g * R + r, jc = min(j, L - 1);
          ldg_slice<W>(k[r], base + (int64_t)jc * 3 * D + D + c0 + col, p.scale * kLog2e);
          ldg_slice<W>(v[r], base + (int64_t)jc * 3 * D + 2 * D + c0 + col, 1.f);
          if (j < L && valid[jc] != 0.f) kvalid |= 1u << r;
#pragma unroll
          for (int w = 0; w < W; w++) { dk[r][w] = 0.f; dv[r][w] = 0.f; }
        }
        const int my_j = it.g * R + it.ds;
        const float* lse_h = lse_s + it.hl * S;
        const float* Di_h = Di_s + it.hl * S;
        const uint32_t* hb_h = hb_s + it.hl * S;
        for (int i = 0; i < L; i++) {
          float qv[W], gv[W], ps[R], pd[R];
          lds_slice<W>(qv, T0 + i * TP + col);
          lds_slice<W>(gv, T1 + i * TP + col);
#pragma unroll
          for (int r = 0; r < R; r++) {
            ps[r] = quad_sum<R>(dot_slice<W>(k[r], qv));
            pd[r] = quad_sum<R>(dot_slice<W>(v[r], gv));
          }
          uint32_t bits = (1u << R) - 1u;
          if (p.drop_p > 0.f) bits = quad_bits<R>(attn_keep(hb_h[i], my_j, p.drop_thr));
          bits &= kvalid;
          const float lse_i = lse_h[i], Di = Di_h[i];
#pragma unroll
          for (int r = 0; r < R; r++) {
            const float pr = ((kvalid >> r) & 1u) ? ex2(ps[r] - lse_i) : 0.f;
            const bool kept = (bits >> r) & 1u;
            const float pk = kept ? pr * inv_keep : 0.f;               // p * keep
            const float dS = pr * ((kept ? pd[r] * inv_keep : 0.f) - Di);
#pragma unroll
            for (int w = 0; w < W; w++) {
              dv[r][w] = fmaf(pk, gv[w], dv[r][w]);
              dk[r][w] = fmaf(dS, qv[w], dk[r][w]);
            }
          }
        }
        dk_scale = p.scale;   // scale*log2e sits on k here, so dk accumulated raw q: dK = scale * sum dS q
      }
    }
    if (rd == 0) __syncthreads();                           // Q and dO are dead: their tiles take dK and dV
    if (warp_on && it.active) {
#pragma unroll
      for (int r = 0; r < R; r++) {
        const int j = it.g * R + r;
        const int slot = w0 + (int)threadIdx.x - (int)blockDim.x;
        park_or_store<R, W>(T0, Ov0, rd == 0, slot, j, TP, col, r, dk[r], dk_scale, j < L);
        park_or_store<R, W>(T1, Ov1, rd == 0, slot, j, TP, col, r, dv[r], 1.f, j < L);
      }
    }
  }
  __syncthreads();
  if (rounds > 1) {
    unpark<DH, R>(T0, Ov0, items, G, L, TP);
    unpark<DH, R>(T1, Ov1, items, G, L, TP);
    __syncthreads();
  }
  flush_tile(T0, L, TW, dbase ? dbase + D : nullptr, hbase ? hbase + D : nullptr, lbase ? lbase + D : nullptr, 3 * D, csum ? csum + D : nullptr);
  flush_tile(T1, L, TW, dbase ? dbase + 2 * D : nullptr, hbase ? hbase + 2 * D : nullptr, lbase ? lbase + 2 * D : nullptr, 3 * D,
             csum ? csum + 2 * D : nullptr);
}

// ================================================================================================ tensor-core path (head dim 32)
// Sequences of at most 64 tokens (every item, every 50-click history) run on mma.sync.m16n8k16 bf16 tiles, one WARP per
// (sequence, head) and no shared operand staging at all: fragments are read straight from the fp32 rows (8-byte loads, L1/L2
// resident), split into (hi, lo) bf16 in registers, and every product issues three MMAs (lo·hi, hi·lo, hi·hi) into fp32
// accumulators — the same error-compensated scheme as the tcgen05 GEMM, so fp32 parity holds.  Transposed operands
// (V for P·V, K for dS·K, Q / dO for the key-side gradients, dS^T / P^T) are produced from the row-major fragments with
// movmatrix (8x8 b16 transpose across lanes).  Softmax, masking and dropout act on accumulator fragments: every (query, key)
// pair is owned by exactly one lane, so nothing is computed twice.  Dropout bits are the same function of (row, head, key) as
// in the SIMT kernels, which remain the path for longer sequences and other head dims.
namespace tcm {

__device__ __forceinline__ void split2(float x0, float x1, uint32_t& h, uint32_t& l) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x1), "f"(x0));
  const float r0 = x0 - __uint_as_float(h << 16), r1 = x1 - __uint_as_float(h & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(r1), "f"(r0));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// c += (ah + al) · (bh + bl) without the lo·lo term; small terms first
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0, uint32_t bh1,
                                     uint32_t bl0, uint32_t bl1) {
  mma16816(c, al, bh0, bh1);
  mma16816(c, ah, bl0, bl1);
  mma16816(c, ah, bh0, bh1);
}
__device__ __forceinline__ uint32_t movm(uint32_t x) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(x));
  return d;
}
__device__ __forceinline__ float2 ld2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_add(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}
// A fragments (16 rows x 32 columns, two k-steps) of rows r0 / r1 of a row-major fp32 matrix, scaled
__device__ __forceinline__ void load_a(const float* __restrict__ row0p, const float* __restrict__ row1p, int t, float s, uint32_t (&h)[2][4],
                                       uint32_t (&l)[2][4]) {
#pragma unroll
  for (int kk = 0; kk < 2; kk++) {
    const float2 a = ld2(row0p + kk * 16 + 2 * t), b = ld2(row1p + kk * 16 + 2 * t);
    const float2 c = ld2(row0p + kk * 16 + 8 + 2 * t), d = ld2(row1p + kk * 16 + 8 + 2 * t);
    split2(a.x * s, a.y * s, h[kk][0], l[kk][0]);
    split2(b.x * s, b.y * s, h[kk][1], l[kk][1]);
    split2(c.x * s, c.y * s, h[kk][2], l[kk][2]);
    split2(d.x * s, d.y * s, h[kk][3], l[kk][3]);
  }
}
// A fragments from a shared tile (plain loads: the tile pointer is a shared-memory address)
__device__ __forceinline__ void load_a_s(const float* row0p, const float* row1p, int t, float s, uint32_t (&h)[2][4], uint32_t (&l)[2][4]) {
#pragma unroll
  for (int kk = 0; kk < 2; kk++) {
    const float2 a = *reinterpret_cast<const float2*>(row0p + kk * 16 + 2 * t), b = *reinterpret_cast<const float2*>(row1p + kk * 16 + 2 * t);
    const float2 c = *reinterpret_cast<const float2*>(row0p + kk * 16 + 8 + 2 * t), d = *reinterpret_cast<const float2*>(row1p + kk * 16 + 8 + 2 * t);
    split2(a.x * s, a.y * s, h[kk][0], l[kk][0]);
    split2(b.x * s, b.y * s, h[kk][1], l[kk][1]);
    split2(c.x * s, c.y * s, h[kk][2], l[kk][2]);
    split2(d.x * s, d.y * s, h[kk][3], l[kk][3]);
  }
}
// validity bits of up to 64 keys of this sequence (dense layout: the key mask; packed: every key below L)
__device__ __forceinline__ unsigned long long key_bits(const MhaParams& p, int64_t n, int L, int lane) {
  const bool dense = !p.seq.cu && p.seq.mask;
  const int S = p.seq.S;
  const bool a = lane < L && (!dense || p.seq.mask[n * S + lane] > 0);
  const bool b = lane + 32 < L && (!dense || p.seq.mask[n * S + lane + 32] > 0);
  const uint32_t lo = __ballot_sync(0xffffffffu, a), hi = __ballot_sync(0xffffffffu, b);
  return ((unsigned long long)hi << 32) | lo;
}

constexpr int TC_MAX_L = 64;
constexpr int TCP = 136;   // fp32 pitch of a 128-column (4 heads x 32) tile: 8-byte fragment loads of 8 rows x 4 lanes hit 32 distinct banks

// [L][128] columns of a row-major fp32 matrix -> shared tile with pitch TCP, 16 bytes per cp.async
__device__ __forceinline__ void stage128(float* T, const float* __restrict__ src, int64_t ld, int L) {
  for (int idx = threadIdx.x; idx < L * 32; idx += blockDim.x) {
    const int r = idx >> 5, c = (idx & 31) * 4;
    cp_async16(T + r * TCP + c, src + (int64_t)r * ld + c);
  }
}

// MW warps share a head: warp (head, mw) takes the 16-query tiles mw, mw + MW, ...  MW = 1 when there are thousands of sequences
// (items); MW = 4 when a launch has few (the 64 click histories of a batch), so that a 50-token history is four warps deep, not one.
template <int MW>
__global__ void __launch_bounds__(128 * MW, MW == 1 ? 4 : 1) mha_fwd_tc_kernel(MhaParams p) {
  pdl_prologue();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int hw = warp & 3, mw = warp >> 2;
  const int64_t n = blockIdx.x;
  int64_t row0; int L;
  seq_range(p.seq, n, row0, L);
  if (L == 0) return;
  extern __shared__ __align__(16) float smem[];
  const int D = p.D, H = p.H, S = p.seq.S, h = blockIdx.y * 4 + hw, c0 = h * 32, cw = hw * 32;
  const float* base = p.qkv + row0 * 3 * (int64_t)D;
  float* Ks = smem;                               // [L][TCP]: the four heads of this CTA
  float* Vs = Ks + (size_t)S * TCP;
  float* Qs = Vs + (size_t)S * TCP;
  stage128(Qs, base + blockIdx.y * 128, 3 * D, L);
  stage128(Ks, base + D + blockIdx.y * 128, 3 * D, L);
  stage128(Vs, base + 2 * D + blockIdx.y * 128, 3 * D, L);
  const unsigned long long kv = key_bits(p, n, L, lane);
  const float qs = p.scale * kLog2e;
  const float inv_keep = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
  const int nt = (L + 7) >> 3, mt = (L + 15) >> 4;
  cp_async_wait_all();
  __syncthreads();
  for (int m = mw; m < mt; m += MW) {
    const int i0 = m * 16 + g, i1 = i0 + 8, i0c = min(i0, L - 1), i1c = min(i1, L - 1);
    uint32_t qh[2][4], ql[2][4];
    load_a_s(Qs + i0c * TCP + cw, Qs + i1c * TCP + cw, t, qs, qh, ql);
    float c[8][4];
#pragma unroll
    for (int nn = 0; nn < 8; nn++) {
      c[nn][0] = c[nn][1] = c[nn][2] = c[nn][3] = 0.f;
      if (nn < nt) {                                               // warp-uniform
        const float* kp = Ks + min(nn * 8 + g, L - 1) * TCP + cw + 2 * t;
#pragma unroll
        for (int kk = 0; kk < 2; kk++) {
          const float2 k0 = *reinterpret_cast<const float2*>(kp + kk * 16), k1 = *reinterpret_cast<const float2*>(kp + kk * 16 + 8);
          uint32_t bh0, bl0, bh1, bl1;
          split2(k0.x, k0.y, bh0, bl0);
          split2(k1.x, k1.y, bh1, bl1);
          mma3(c[nn], qh[kk], ql[kk], bh0, bh1, bl0, bl1);
        }
      }
    }
    // ---- softmax over the keys of rows i0 (c[.][0..1]) and i1 (c[.][2..3]); scores are in the log2 domain
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nn = 0; nn < 8; nn++) {
      if (nn < nt) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int j = nn * 8 + 2 * t + e;
          if (!((kv >> j) & 1ull)) { c[nn][e] = -INFINITY; c[nn][2 + e] = -INFINITY; }
          mx0 = fmaxf(mx0, c[nn][e]);
          mx1 = fmaxf(mx1, c[nn][2 + e]);
        }
      }
    }
    mx0 = quad_max(mx0); mx1 = quad_max(mx1);
    const uint32_t hb0 = attn_hash_base(p.seed, (uint64_t)(row0 + i0c) * H + h), hb1 = attn_hash_base(p.seed, (uint64_t)(row0 + i1c) * H + h);
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int nn = 0; nn < 8; nn++) {
      if (nn < nt) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int j = nn * 8 + 2 * t + e;
          const float p0 = ex2(c[nn][e] - mx0), p1 = ex2(c[nn][2 + e] - mx1);   // masked keys: 2^-inf = 0; all masked: NaN as torch
          l0 += p0; l1 += p1;
          bool k0 = true, k1 = true;
          if (p.drop_p > 0.f) { k0 = attn_keep(hb0, j, p.drop_thr); k1 = attn_keep(hb1, j, p.drop_thr); }
          c[nn][e] = k0 ? p0 * inv_keep : 0.f;
          c[nn][2 + e] = k1 ? p1 * inv_keep : 0.f;
        }
      }
    }
    l0 = quad_add(l0); l1 = quad_add(l1);
    if (t == 0) {
      if (i0 < L) p.lse[(row0 + i0) * H + h] = (mx0 + log2f(l0)) * kLn2;
      if (i1 < L) p.lse[(row0 + i1) * H + h] = (mx1 + log2f(l1)) * kLn2;
    }
    // ---- O = P·V: the probability fragments of two key tiles are the A operand of one 16-key step
    float o[4][4];
#pragma unroll
    for (int dn = 0; dn < 4; dn++) o[dn][0] = o[dn][1] = o[dn][2] = o[dn][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
      if (kk * 16 < L) {
        uint32_t ph[4], pl[4];
        split2(c[2 * kk][0], c[2 * kk][1], ph[0], pl[0]);
        split2(c[2 * kk][2], c[2 * kk][3], ph[1], pl[1]);
        split2(c[2 * kk + 1][0], c[2 * kk + 1][1], ph[2], pl[2]);
        split2(c[2 * kk + 1][2], c[2 * kk + 1][3], ph[3], pl[3]);
        const float* va = Vs + min(kk * 16 + g, L - 1) * TCP + cw + 2 * t;
        const float* vb = Vs + min(kk * 16 + 8 + g, L - 1) * TCP + cw + 2 * t;
#pragma unroll
        for (int dn = 0; dn < 4; dn++) {
          const float2 xa = *reinterpret_cast<const float2*>(va + dn * 8), xb = *reinterpret_cast<const float2*>(vb + dn * 8);
          uint32_t ah, al, bh, bl;
          split2(xa.x, xa.y, ah, al);
          split2(xb.x, xb.y, bh, bl);
          mma3(o[dn], ph, pl, movm(ah), movm(bh), movm(al), movm(bl));
        }
      }
    }
    const float r0 = 1.f / l0, r1 = 1.f / l1;
#pragma unroll
    for (int dn = 0; dn < 4; dn++) {
      const int col = c0 + dn * 8 + 2 * t;
#pragma unroll
      for (int rr = 0; rr < 2; rr++) {
        const int i = rr ? i1 : i0;
        if (i < L) {
          const float x0 = o[dn][2 * rr] * (rr ? r1 : r0), x1 = o[dn][2 * rr + 1] * (rr ? r1 : r0);
          const int64_t off = (row0 + i) * (int64_t)D + col;
          if (p.ctx) *reinterpret_cast<float2*>(p.ctx + off) = make_float2(x0, x1);
          if (p.ctx_hi) {
            uint32_t hh, ll;
            split2(x0, x1, hh, ll);
            *reinterpret_cast<uint32_t*>(p.ctx_hi + off) = hh;
            *reinterpret_cast<uint32_t*>(p.ctx_lo + off) = ll;
          }
        }
      }
    }
  }
}

__device__ __forceinline__ float col_add(float v) {     // sum over the 8 row groups of a fragment column (lanes with equal t)
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  return v + __shfl_xor_sync(0xffffffffu, v, 16);
}

// Backward, one warp per (sequence, head).  Outer loop over 16-key tiles j (their K / V fragments and the dK / dV accumulators stay
// in registers), inner loop over 16-query tiles i: S and dP blocks are recomputed once per (i, j) pair, turned into dS and P∘keep
// in place, and feed three more contractions: dQ(i) += dS·K(j) (accumulated in a shared tile whose columns this warp owns),
// dK(j) += dS^T·Q(i), dV(j) += (P∘keep)^T·dO(i).  Q and dO of the CTA's four heads are staged once in shared memory.
// JW warps share a head: warp (head, jw) takes the key tiles jw, jw + JW, ... and accumulates its dQ contribution in its OWN shared
// tile; the JW tiles are added in a fixed order at the end (deterministic, no atomics).  JW = 1 for the items, 2 for small launches.
template <int JW>
__global__ void __launch_bounds__(128 * JW, JW == 1 ? 3 : 1) mha_bwd_tc_kernel(MhaParams p) {
  pdl_prologue();
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int hw = warp & 3, jw = warp >> 2;
  const int64_t n = blockIdx.x;
  int64_t row0; int L;
  seq_range(p.seq, n, row0, L);
  if (L == 0) return;
  const int D = p.D, H = p.H, S = p.seq.S, h = blockIdx.y * 4 + hw, c0 = h * 32, cw = hw * 32;
  const float* base = p.qkv + row0 * 3 * (int64_t)D;
  const float* gbase = p.dctx + row0 * (int64_t)D;
  const float* obase = p.ctx + row0 * (int64_t)D;
  float* Qs = smem;                                // [S][TCP] q (unscaled)
  float* Gs = Qs + (size_t)S * TCP;                // dO
  float* dQall = Gs + (size_t)S * TCP;             // [JW][S][TCP] dQ accumulators, one tile per key-tile warp
  float* Ds = dQall + (size_t)JW * S * TCP;        // [4][TC_MAX_L]  D_i = dO_i · O_i
  float* Cs = Ds + 4 * TC_MAX_L;                   // [JW][4][64] column sums of the dK / dV rows of each warp (JW > 1)
  float* dQs = dQall + (size_t)jw * S * TCP;
  stage128(Qs, base + blockIdx.y * 128, 3 * D, L);
  stage128(Gs, gbase + blockIdx.y * 128, D, L);
  for (int idx = threadIdx.x; idx < JW * L * 32; idx += blockDim.x) {
    const int w = idx / (L * 32), r = idx - w * (L * 32);
    *reinterpret_cast<float4*>(dQall + ((size_t)w * S + (r >> 5)) * TCP + (r & 31) * 4) = f4_zero();
  }
  const unsigned long long kv = key_bits(p, n, L, lane);
  const float qs = p.scale * kLog2e;
  const float inv_keep = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
  const int mt = (L + 15) >> 4;
  cp_async_wait_all();
  __syncthreads();
  for (int m = jw; m < mt; m += JW) {
    const int i0 = m * 16 + g, i1 = i0 + 8, i0c = min(i0, L - 1), i1c = min(i1, L - 1);
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int dn = 0; dn < 4; dn++) {
      const float2 a = *reinterpret_cast<const float2*>(Gs + i0c * TCP + cw + dn * 8 + 2 * t), oa = ld2(obase + (int64_t)i0c * D + c0 + dn * 8 + 2 * t);
      const float2 b = *reinterpret_cast<const float2*>(Gs + i1c * TCP + cw + dn * 8 + 2 * t), ob = ld2(obase + (int64_t)i1c * D + c0 + dn * 8 + 2 * t);
      d0 = fmaf(a.x, oa.x, fmaf(a.y, oa.y, d0));
      d1 = fmaf(b.x, ob.x, fmaf(b.y, ob.y, d1));
    }
    d0 = quad_add(d0); d1 = quad_add(d1);
    if (t == 0) {
      if (i0 < L) Ds[hw * TC_MAX_L + i0] = d0;
      if (i1 < L) Ds[hw * TC_MAX_L + i1] = d1;
    }
  }
  if (JW == 1) __syncwarp(); else __syncthreads();

  float ksum[4][2], vsum[4][2];
#pragma unroll
  for (int dn = 0; dn < 4; dn++) ksum[dn][0] = ksum[dn][1] = vsum[dn][0] = vsum[dn][1] = 0.f;
  float* dbase = p.dqkv ? p.dqkv + row0 * 3 * (int64_t)D + c0 : nullptr;
  __nv_bfloat16* hbase = p.dq_hi ? p.dq_hi + row0 * 3 * (int64_t)D + c0 : nullptr;
  __nv_bfloat16* lbase = p.dq_lo ? p.dq_lo + row0 * 3 * (int64_t)D + c0 : nullptr;
  auto emit = [&](int64_t off, float x0, float x1) {       // two adjacent result columns: fp32 and / or planes
    if (dbase) *reinterpret_cast<float2*>(dbase + off) = make_float2(x0, x1);
    if (hbase) {
      uint32_t hh, ll;
      split2(x0, x1, hh, ll);
      *reinterpret_cast<uint32_t*>(hbase + off) = hh;
      *reinterpret_cast<uint32_t*>(lbase + off) = ll;
    }
  };

  for (int jt = jw; jt < mt; jt += JW) {
    uint32_t kh[2][2][2], kl[2][2][2], vh[2][2][2], vl[2][2][2];
#pragma unroll
    for (int nn = 0; nn < 2; nn++) {
      const int64_t jr = min(jt * 16 + nn * 8 + g, L - 1);
      const float* kp = base + jr * 3 * D + D + c0 + 2 * t;
      const float* vp = kp + D;
#pragma unroll
      for (int kk = 0; kk < 2; kk++) {
        const float2 k0 = ld2(kp + kk * 16), k1 = ld2(kp + kk * 16 + 8), v0 = ld2(vp + kk * 16), v1 = ld2(vp + kk * 16 + 8);
        split2(k0.x, k0.y, kh[nn][kk][0], kl[nn][kk][0]);
        split2(k1.x, k1.y, kh[nn][kk][1], kl[nn][kk][1]);
        split2(v0.x, v0.y, vh[nn][kk][0], vl[nn][kk][0]);
        split2(v1.x, v1.y, vh[nn][kk][1], vl[nn][kk][1]);
      }
    }
    float dk[4][4], dv[4][4];
#pragma unroll
    for (int dn = 0; dn < 4; dn++) dk[dn][0] = dk[dn][1] = dk[dn][2] = dk[dn][3] = dv[dn][0] = dv[dn][1] = dv[dn][2] = dv[dn][3] = 0.f;
    for (int m = 0; m < mt; m++) {
      const int i0 = m * 16 + g, i1 = i0 + 8, i0c = min(i0, L - 1), i1c = min(i1, L - 1);
      const bool r0ok = i0 < L, r1ok = i1 < L;
      uint32_t qh[2][4], ql[2][4], gh[2][4], gl[2][4];
      load_a_s(Qs + i0c * TCP + cw, Qs + i1c * TCP + cw, t, qs, qh, ql);
      load_a_s(Gs + i0c * TCP + cw, Gs + i1c * TCP + cw, t, 1.f, gh, gl);
      float cs[2][4], cp[2][4];
#pragma unroll
      for (int nn = 0; nn < 2; nn++) {
        cs[nn][0] = cs[nn][1] = cs[nn][2] = cs[nn][3] = 0.f;
        cp[nn][0] = cp[nn][1] = cp[nn][2] = cp[nn][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < 2; kk++) {
          mma3(cs[nn], qh[kk], ql[kk], kh[nn][kk][0], kh[nn][kk][1], kl[nn][kk][0], kl[nn][kk][1]);
          mma3(cp[nn], gh[kk], gl[kk], vh[nn][kk][0], vh[nn][kk][1], vl[nn][kk][0], vl[nn][kk][1]);
        }
      }
      const float l20 = p.lse[(row0 + i0c) * H + h] * kLog2e, l21 = p.lse[(row0 + i1c) * H + h] * kLog2e;
      const float D0 = Ds[hw * TC_MAX_L + i0c], D1 = Ds[hw * TC_MAX_L + i1c];
      const uint32_t hb0 = attn_hash_base(p.seed, (uint64_t)(row0 + i0c) * H + h), hb1 = attn_hash_base(p.seed, (uint64_t)(row0 + i1c) * H + h);
#pragma unroll
      for (int nn = 0; nn < 2; nn++) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int j = jt * 16 + nn * 8 + 2 * t + e;
          const bool ok = (kv >> j) & 1ull;
          const float p0 = (ok && r0ok) ? ex2(cs[nn][e] - l20) : 0.f, p1 = (ok && r1ok) ? ex2(cs[nn][2 + e] - l21) : 0.f;
          bool k0 = true, k1 = true;
          if (p.drop_p > 0.f) { k0 = attn_keep(hb0, j, p.drop_thr); k1 = attn_keep(hb1, j, p.drop_thr); }
          const float dP0 = k0 ? cp[nn][e] * inv_keep : 0.f, dP1 = k1 ? cp[nn][2 + e] * inv_keep : 0.f;
          cs[nn][e] = p0 * (dP0 - D0);                         // dS (the softmax scale is applied on the way out)
          cs[nn][2 + e] = p1 * (dP1 - D1);
          cp[nn][e] = k0 ? p0 * inv_keep : 0.f;                // P ∘ keep
          cp[nn][2 + e] = k1 ? p1 * inv_keep : 0.f;
        }
      }
      uint32_t sh[4], sl[4], ph[4], pl[4];
      split2(cs[0][0], cs[0][1], sh[0], sl[0]); split2(cs[0][2], cs[0][3], sh[1], sl[1]);
      split2(cs[1][0], cs[1][1], sh[2], sl[2]); split2(cs[1][2], cs[1][3], sh[3], sl[3]);
      split2(cp[0][0], cp[0][1], ph[0], pl[0]); split2(cp[0][2], cp[0][3], ph[1], pl[1]);
      split2(cp[1][0], cp[1][1], ph[2], pl[2]); split2(cp[1][2], cp[1][3], ph[3], pl[3]);
      // dQ(i) += scale * dS · K(j)
#pragma unroll
      for (int dn = 0; dn < 4; dn++) {
        const int kk = dn >> 1, r = dn & 1;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        mma3(acc, sh, sl, movm(kh[0][kk][r]), movm(kh[1][kk][r]), movm(kl[0][kk][r]), movm(kl[1][kk][r]));
        if (r0ok) {
          float2* q = reinterpret_cast<float2*>(dQs + i0 * TCP + cw + dn * 8 + 2 * t);
          float2 v = *q; v.x = fmaf(acc[0], p.scale, v.x); v.y = fmaf(acc[1], p.scale, v.y); *q = v;
        }
        if (r1ok) {
          float2* q = reinterpret_cast<float2*>(dQs + i1 * TCP + cw + dn * 8 + 2 * t);
          float2 v = *q; v.x = fmaf(acc[2], p.scale, v.x); v.y = fmaf(acc[3], p.scale, v.y); *q = v;
        }
      }
      // key side: dK(j) += dS^T · Q(i),  dV(j) += (P∘keep)^T · dO(i)
      uint32_t th[4] = {movm(sh[0]), movm(sh[2]), movm(sh[1]), movm(sh[3])}, tl[4] = {movm(sl[0]), movm(sl[2]), movm(sl[1]), movm(sl[3])};
      uint32_t uh[4] = {movm(ph[0]), movm(ph[2]), movm(ph[1]), movm(ph[3])}, ul[4] = {movm(pl[0]), movm(pl[2]), movm(pl[1]), movm(pl[3])};
#pragma unroll
      for (int dn = 0; dn < 4; dn++) {
        const int kk = dn >> 1, r = dn & 1;
        mma3(dk[dn], th, tl, movm(qh[kk][2 * r]), movm(qh[kk][2 * r + 1]), movm(ql[kk][2 * r]), movm(ql[kk][2 * r + 1]));
        mma3(dv[dn], uh, ul, movm(gh[kk][2 * r]), movm(gh[kk][2 * r + 1]), movm(gl[kk][2 * r]), movm(gl[kk][2 * r + 1]));
      }
    }
    // rows j of dK (q carried scale*log2e: undo the log2e) and dV
    const int j0 = jt * 16 + g, j1 = j0 + 8;
#pragma unroll
    for (int dn = 0; dn < 4; dn++) {
#pragma unroll
      for (int rr = 0; rr < 2; rr++) {
        const int j = rr ? j1 : j0;
        if (j < L) {
          const float x0 = dk[dn][2 * rr] * kLn2, x1 = dk[dn][2 * rr + 1] * kLn2, y0 = dv[dn][2 * rr], y1 = dv[dn][2 * rr + 1];
          const int64_t off = (int64_t)j * 3 * D + dn * 8 + 2 * t;
          emit(off + D, x0, x1);
          emit(off + 2 * D, y0, y1);
          ksum[dn][0] += x0; ksum[dn][1] += x1; vsum[dn][0] += y0; vsum[dn][1] += y1;
        }
      }
    }
  }
  float* csum = p.colsum_part ? p.colsum_part + n * 3 * (int64_t)D + c0 : nullptr;
  if (JW > 1) {
    // the key-tile warps of a head hand their dK / dV column sums to warp (head, 0), which also adds up the JW dQ tiles
    if (csum) {
#pragma unroll
      for (int dn = 0; dn < 4; dn++) {
        const float k0 = col_add(ksum[dn][0]), k1 = col_add(ksum[dn][1]), v0 = col_add(vsum[dn][0]), v1 = col_add(vsum[dn][1]);
        if (g == 0) {
          float* c = Cs + ((size_t)jw * 4 + hw) * 64 + dn * 8 + 2 * t;
          c[0] = k0; c[1] = k1; c[32] = v0; c[33] = v1;
        }
      }
    }
    __syncthreads();
    if (jw != 0) return;
  } else {
    __syncwarp();
  }
  // dQ rows out of this head's columns of the shared tile(s), and the per-sequence column sums (in_proj bias gradient)
#pragma unroll
  for (int dn = 0; dn < 4; dn++) {
    float q0 = 0.f, q1 = 0.f;
    for (int i = g; i < L; i += 8) {
      float2 v = *reinterpret_cast<const float2*>(dQall + i * TCP + cw + dn * 8 + 2 * t);
#pragma unroll
      for (int w = 1; w < JW; w++) {
        const float2 u = *reinterpret_cast<const float2*>(dQall + ((size_t)w * S + i) * TCP + cw + dn * 8 + 2 * t);
        v.x += u.x; v.y += u.y;
      }
      emit((int64_t)i * 3 * D + dn * 8 + 2 * t, v.x, v.y);
      q0 += v.x; q1 += v.y;
    }
    if (csum) {
      q0 = col_add(q0); q1 = col_add(q1);
      float k0, k1, v0, v1;
      if (JW == 1) {
        k0 = col_add(ksum[dn][0]); k1 = col_add(ksum[dn][1]); v0 = col_add(vsum[dn][0]); v1 = col_add(vsum[dn][1]);
      } else {
        k0 = k1 = v0 = v1 = 0.f;
#pragma unroll
        for (int w = 0; w < JW; w++) {
          const float* c = Cs + ((size_t)w * 4 + hw) * 64 + dn * 8 + 2 * t;
          k0 += c[0]; k1 += c[1]; v0 += c[32]; v1 += c[33];
        }
      }
      if (g == 0) {
        const int col = dn * 8 + 2 * t;
        csum[col] = q0; csum[col + 1] = q1;
        csum[D + col] = k0; csum[D + col + 1] = k1;
        csum[2 * D + col] = v0; csum[2 * D + col + 1] = v1;
      }
    }
  }
}

}  // namespace tcm

// heads per CTA: the largest divisor of H with HG*DH <= 128 columns
static int pick_hg(int H, int dh) {
  int hg = 128 / dh;
  if (hg < 1) hg = 1;
  if (hg > H) hg = H;
  while (H % hg) hg--;
  return hg;
}
static int pick_threads(int64_t S, int hg, int R) {
  int64_t t = (int64_t)hg * ((S + R - 1) / R) * R;
  t = (t + 31) / 32 * 32;
  // registers are granted to a CTA in units of 4 warps: a 160-thread CTA pays for 256.  When the longest sequences overshoot
  // 128 threads by one warp, run 128 threads and let those (rare) sequences take a second round.
  if (t > 128 && t <= 160) t = 128;
  return (int)(t < 64 ? 64 : (t > MHA_THREADS ? MHA_THREADS : t));
}

static int overflow_items(int64_t S, int hg, int R, int threads) {
  const int64_t items = (int64_t)hg * ((S + R - 1) / R) * R;
  return items > threads ? (int)(items - threads) : 0;
}

template <int DH, int R>
static int launch_fwd(MhaParams p, int64_t N, cudaStream_t st) {
  p.HG = pick_hg(p.H, DH);
  const int TP = p.HG * DH + 4, S = p.seq.S, threads = pick_threads(S, p.HG, R);
  p.ov_items = overflow_items(S, p.HG, R, threads);
  size_t smem = ((size_t)2 * S * TP + ((S + 3) & ~3) + (size_t)p.ov_items * DH) * sizeof(float);
  LK_REQUIRE(smem <= 227 * 1024, LK_ERR_SHAPE, "lk_mha_fwd: tiles (%zu B) do not fit shared memory", smem);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(mha_fwd_seq_kernel<DH, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(mha_fwd_seq_kernel<DH, R>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    attr = true;
  }
  dim3 grid((unsigned)N, (unsigned)(p.H / p.HG));
  LK_LAUNCH((mha_fwd_seq_kernel<DH, R>), grid, threads, smem, st, p);
  return check_launch("mha_fwd");
}
template <int DH, int R, bool SP>
static int launch_bwd_sp(const MhaParams& p, int64_t N, size_t smem, int threads, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(mha_bwd_seq_kernel<DH, R, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(mha_bwd_seq_kernel<DH, R, SP>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    attr = true;
  }
  dim3 grid((unsigned)N, (unsigned)(p.H / p.HG));
  LK_LAUNCH((mha_bwd_seq_kernel<DH, R, SP>), grid, threads, smem, st, p);
  return check_launch("mha_bwd");
}
template <int DH, int R>
static int launch_bwd(MhaParams p, int64_t N, cudaStream_t st) {
  p.HG = pick_hg(p.H, DH);
  const int TP = p.HG * DH + 4, S = p.seq.S, LP = (S + 1) & ~1, threads = pick_threads(S, p.HG, R);
  p.ov_items = overflow_items(S, p.HG, R, threads);
  const size_t base = ((size_t)2 * S * TP + 3 * (size_t)p.HG * S + ((S + 3) & ~3) + 2 * (size_t)p.ov_items * DH + 4) * sizeof(float);
  const size_t with_p = base + (size_t)2 * p.HG * S * LP * sizeof(float);
  if (with_p <= 227 * 1024) return launch_bwd_sp<DH, R, true>(p, N, with_p, threads, st);
  LK_REQUIRE(base <= 227 * 1024, LK_ERR_SHAPE, "lk_mha_bwd: tiles (%zu B) do not fit shared memory", base);
  return launch_bwd_sp<DH, R, false>(p, N, base, threads, st);
}

// tensor-core kernels: head dim 32, heads in groups of four, at most 64 tokens per sequence (LK_MHA_TC=0 forces the SIMT kernels)
static bool tc_path_ok(const MhaParams& p, int64_t S) {
  static const bool on = [] { const char* e = getenv("LK_MHA_TC"); return !(e && e[0] == '0'); }();
  return on && p.D / p.H == 32 && p.H % 4 == 0 && S <= tcm::TC_MAX_L;
}

// few sequences with more than one 16-token tile: spread each head over four warps
static bool tc_small_launch(int64_t N, int64_t H, int64_t S) { return S > 16 && N * (H / 4) <= 2 * kNumSMs; }

static void set_dropout(MhaParams& p, float drop_p, uint64_t seed) {
  p.drop_p = drop_p;
  p.seed = (unsigned long long)seed;
  double t = (double)drop_p * 16777216.0;
  p.drop_thr = t <= 0.0 ? 0u : (t >= 16777216.0 ? 16777216u : (uint32_t)t);
}

}  // namespace lk

using namespace lk;

extern "C" {

int lk_mha_fwd(const float* qkv, const int64_t* mask, const int32_t* cu, float* ctx, void* ctx_hi, void* ctx_lo, float* lse, int64_t N,
               int64_t S, int64_t D, int64_t H, float drop_p, uint64_t seed, cudaStream_t st) {
  LK_REQUIRE(H > 0 && D % H == 0 && D % 4 == 0, LK_ERR_SHAPE, "lk_mha_fwd: D=%ld not divisible by heads=%ld", (long)D, (long)H);
  LK_REQUIRE(ctx || ctx_hi, LK_ERR_ARG, "lk_mha_fwd: no output requested");
  LK_REQUIRE(!ctx_hi || ctx_lo, LK_ERR_ARG, "lk_mha_fwd: both output planes are needed");
  if (N == 0 || S == 0) return LK_OK;
  const int dh = (int)(D / H);
  MhaParams p{};
  p.qkv = qkv; p.ctx = ctx; p.ctx_hi = (__nv_bfloat16*)ctx_hi; p.ctx_lo = (__nv_bfloat16*)ctx_lo; p.lse = lse;
  p.seq = SeqView{cu, mask, (int)S};
  p.D = (int)D; p.H = (int)H; p.scale = 1.0f / sqrtf((float)dh);
  set_dropout(p, drop_p, seed);
  switch (dh) {
    case 8: return launch_fwd<8, 2>(p, N, st);
    case 16: return launch_fwd<16, 2>(p, N, st);
    case 32:
      if (tc_path_ok(p, S)) {
        const size_t smem = (size_t)3 * S * tcm::TCP * sizeof(float);
        static bool attr = false;
        if (!attr) {
          const int mx = 3 * tcm::TC_MAX_L * tcm::TCP * (int)sizeof(float);
          cudaFuncSetAttribute(tcm::mha_fwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
          cudaFuncSetAttribute(tcm::mha_fwd_tc_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
          cudaFuncSetAttribute(tcm::mha_fwd_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
          attr = true;
        }
        const dim3 grid((unsigned)N, (unsigned)(H / 4));
        if (tc_small_launch(N, H, S)) LK_LAUNCH((tcm::mha_fwd_tc_kernel<4>), grid, 512, smem, st, p);
        else LK_LAUNCH((tcm::mha_fwd_tc_kernel<1>), grid, 128, smem, st, p);
        return check_launch("mha_fwd_tc");
      }
      return launch_fwd<32, 2>(p, N, st);
    case 64: return launch_fwd<64, 4>(p, N, st);
  }
  LK_REQUIRE(false, LK_ERR_SHAPE, "lk_mha_fwd: head dim %d not in {8,16,32,64}", dh);
}

int lk_mha_bwd(const float* qkv, const int64_t* mask, const int32_t* cu, const float* ctx, const float* lse, const float* dctx,
               float* dqkv, void* dq_hi, void* dq_lo, float* colsum_part, int64_t N, int64_t S, int64_t D, int64_t H, float drop_p,
               uint64_t seed, cudaStream_t st) {
  LK_REQUIRE(H > 0 && D % H == 0 && D % 4 == 0, LK_ERR_SHAPE, "lk_mha_bwd: D=%ld not divisible by heads=%ld", (long)D, (long)H);
  LK_REQUIRE(dqkv || dq_hi, LK_ERR_ARG, "lk_mha_bwd: no output requested");
  LK_REQUIRE(!dq_hi || dq_lo, LK_ERR_ARG, "lk_mha_bwd: both output planes are needed");
  LK_REQUIRE(ctx && lse && dctx, LK_ERR_ARG, "lk_mha_bwd: the forward's ctx and lse are needed (D_i = dO·O)");
  if (N == 0 || S == 0) return LK_OK;
  const int dh = (int)(D / H);
  MhaParams p{};
  p.qkv = qkv; p.ctx = const_cast<float*>(ctx); p.lse = const_cast<float*>(lse); p.dctx = dctx; p.dqkv = dqkv;
  p.dq_hi = (__nv_bfloat16*)dq_hi; p.dq_lo = (__nv_bfloat16*)dq_lo; p.colsum_part = colsum_part;
  p.seq = SeqView{cu, mask, (int)S};
  p.D = (int)D; p.H = (int)H; p.scale = 1.0f / sqrtf((float)dh);
  set_dropout(p, drop_p, seed);
  switch (dh) {
    case 8: return launch_bwd<8, 2>(p, N, st);
    case 16: return launch_bwd<16, 2>(p, N, st);
    case 32:
      if (tc_path_ok(p, S)) {
        const bool small = tc_small_launch(N, H, S);
        const int JW = small ? 2 : 1;   // 2 x 128 threads keep the 168 registers the kernel needs (4 would cap it at 128 and spill)
        // (measured, r2: forcing 4 CTAs/SM on the JW = 1 kernel = 128 registers, 260 B of spills: 187 -> 234 us per launch — rejected)
        const size_t smem = ((size_t)(2 + JW) * S * tcm::TCP + 4 * tcm::TC_MAX_L + (size_t)JW * 4 * 64) * sizeof(float);
        LK_REQUIRE(smem <= 227 * 1024, LK_ERR_SHAPE, "lk_mha_bwd: tiles (%zu B) do not fit shared memory", smem);
        static bool attr = false;
        if (!attr) {
          cudaFuncSetAttribute(tcm::mha_bwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (3 * tcm::TC_MAX_L * tcm::TCP + 4 * tcm::TC_MAX_L + 4 * 64) * (int)sizeof(float));
          cudaFuncSetAttribute(tcm::mha_bwd_tc_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
          cudaFuncSetAttribute(tcm::mha_bwd_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
          attr = true;
        }
        const dim3 grid((unsigned)N, (unsigned)(H / 4));
        if (small) LK_LAUNCH((tcm::mha_bwd_tc_kernel<2>), grid, 256, smem, st, p);
        else LK_LAUNCH((tcm::mha_bwd_tc_kernel<1>), grid, 128, smem, st, p);
        return check_launch("mha_bwd_tc");
      }
      return launch_bwd<32, 2>(p, N, st);
    case 64: return launch_bwd<64, 4>(p, N, st);
  }
  LK_REQUIRE(false, LK_ERR_SHAPE, "lk_mha_bwd: head dim %d not in {8,16,32,64}", dh);
}

}  // extern "C"
