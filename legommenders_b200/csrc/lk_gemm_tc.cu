// tcgen05 / TMEM / TMA GEMM for the dense contractions of the hot path (QKV / out-proj / linear / additive-W1 /
// GloVe + LLM projections and their gradients; reference call sites: nn.Linear + nn.MultiheadAttention projections in
// model/operators/attention_operator.py:32-57, model/common/attention.py:17-19, loader/embedding_hub.py:95-96).
//
// Numerics: the reference computes in fp32 and parity is 1e-4 (BASELINE.json), which single-pass bf16/tf32 tensor-core
// math misses.  Operands are therefore carried as TWO bf16 planes (hi = bf16(x), lo = bf16(x - hi); 16
// mantissa bits, the same 4 bytes/element as fp32) and every k-step issues three MMAs into one fp32 TMEM accumulator:
//     D += A_hi·B_hi + A_hi·B_lo + A_lo·B_hi          (the dropped lo·lo term is ~2^-18 relative)
//
// Structure (one persistent CTA per SM, 192 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor 128B-swizzled tiles of the four planes into a 3-stage smem ring
//   warp 1      MMA issuer:   one elected lane issues tcgen05.mma (M=128,N=128,K=16, kind::f16, bf16 in / fp32 out)
//   warps 2-5   epilogue:     tcgen05.ld of the fp32 accumulator (double-buffered in TMEM, 2 x 128 columns) ->
//                             bias / tanh / relu / dropout / row mask -> 16-byte stores
// Operand layouts: K-major tiles ([rows, k] with k contiguous) and MN-major tiles ([k, rows] with rows contiguous, used by
// the weight-gradient contraction whose reduction runs over token rows) — both through 128-byte-swizzle UMMA descriptors.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "lk_tc.cuh"
#include "../../include/legommenders_b200.h"

namespace lk {
namespace tc {

// Two tile shapes.  128x128 (3 stages of 64 KB) is the general one.  128x256 (2 stages of 96 KB, the whole TMEM as the
// double-buffered accumulator) reads every A tile once instead of twice; see pick_bn for where it pays.
constexpr int ACC_STAGES = 2;
// The operand ring is SIX slots of 32 KiB.  A k-block occupies [A_hi|A_lo] + [B_hi] + [B_lo] (three slots, BN = 256) or [A_hi|A_lo] + [B_hi|B_lo]
// (two slots, BN = 128): finer slots keep ~190 KiB of loads in flight whatever the tile width.  (Round 1 used 2 x 96 KiB stages for BN = 256: the
// producer ran a single k-block ahead and every k-block waited out an L2 round trip — 11 us per 128x256x256 tile against 5 us of MMA issue.)
constexpr int SLOT_BYTES = 2 * TILE_BYTES;       // 32 KiB
constexpr int SLOTS = 6;
template <int BN> struct Cfg {
  static constexpr int TILE_B = BN * BK * 2;                    // one plane of the B tile (16 or 32 KiB)
  static constexpr int SLOTS_PER_KB = BN == 256 ? 3 : 2;
  static constexpr int TMEM_COLS = ACC_STAGES * BN;             // fp32 accumulator columns (256 or 512)
  static constexpr int EPI_CHUNKS = BN / 4 / 16;                // 16-column chunks per epilogue warp
};
constexpr int EPI_WARPS = 16;              // warp (q, cg): TMEM lanes 32q..32q+31, column group cg (a quarter of the tile's columns)
constexpr int NUM_THREADS = 64 + EPI_WARPS * 32;   // TMA warp, MMA warp, epilogue warps
constexpr int EPI_LD = 16;                 // staged epilogue rows are unpadded; an XOR swizzle of the 16-byte column keeps both
                                           // the row-per-lane writes and the 4-lanes-per-row reads free of bank conflicts
constexpr int SMEM_BYTES = 3 * 4 * TILE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + EPI_WARPS * 32 * EPI_LD * 4 /*epilogue staging*/;
static_assert(SLOTS * SLOT_BYTES == 3 * 4 * TILE_BYTES, "192 KiB operand ring");
static_assert(SMEM_BYTES <= 227 * 1024, "tc_gemm shared memory budget");

struct Params {
  float* C;                 // [GM, ldc] (or split partials [splits, GM, GN] when partial != null)
  float* partial;
  int GM, GN, GK, ldc;
  int m_tiles, n_tiles, k_blocks, splits, kb_per_split;
  const float* bias;        // [GN] or null
  const int64_t* rowmask;   // [GM] or null
  int rowmask_is_ids;       // 0: keep row when rowmask > 0;  1: rowmask holds ids, keep row when id > -1
  int act, accumulate;
  int store_c;              // 0: C is only read (accumulate) — the result leaves as planes / column sums
  float drop_p;
  unsigned long long seed;
  // fused producers of the NEXT contraction's operands (all optional)
  const int64_t* add_ids[2];   // [GM] row addends: result[row,:] += add_tab[i][add_ids[i][row], :] where the id is > -1
  const float* add_tab[2];     // [*, GN]
  __nv_bfloat16 *out_hi, *out_lo;   // split-bf16 image of the result, pitch ld_planes
  int ld_planes;
  float* colsum_part;          // [m_tiles*4, GN] per-(tile, lane-quarter) column sums of the result rows < GM
};

// Epilogue of one 128x128 accumulator tile by 16 warps: warp (q, cg) owns TMEM lanes 32q..32q+31 and columns
// 32*cg..32*cg+31.  Per 16-column chunk: tcgen05.ld (row = lane) -> bias / activation / dropout / row mask in registers
// -> transpose through a warp-private shared tile -> 64-byte-contiguous row segments to global (a warp store covers 8 rows).

template <int BN>
__device__ __forceinline__ void store_tile(const Params& p, uint32_t tmem_base, int acc, int q, int cg, int lane, int m0, int n0,
                                           int split, bool has_k, float inv_keep, float* stage) {
  constexpr int EPI_CHUNKS = Cfg<BN>::EPI_CHUNKS;
  const int row = m0 + q * 32 + lane;
  const bool row_ok = row < p.GM;
  float* out;
  int ldo;
  if (p.partial) { out = p.partial + (size_t)split * p.GM * p.GN; ldo = p.GN; } else { out = p.C; ldo = p.ldc; }   // C may be null (planes only)
  const bool fused = p.partial == nullptr;
  float rm = 1.f;
  if (fused && p.rowmask && row_ok) rm = p.rowmask[row] > (p.rowmask_is_ids ? -1 : 0) ? 1.f : 0.f;
  const float* add0 = nullptr;
  const float* add1 = nullptr;
  if (fused && row_ok) {
    if (p.add_ids[0]) { const int64_t id = p.add_ids[0][row]; if (id > -1) add0 = p.add_tab[0] + id * (int64_t)p.GN; }
    if (p.add_ids[1]) { const int64_t id = p.add_ids[1][row]; if (id > -1) add1 = p.add_tab[1] + id * (int64_t)p.GN; }
  }
  // the TMEM load of chunk c+1 is in flight while chunk c is processed and stored
  uint32_t v[16];
  const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + cg * (16 * EPI_CHUNKS));
  auto tmem_ld16 = [&](int c) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr0 + (uint32_t)(c * 16)));
  };
  if (n0 + cg * (16 * EPI_CHUNKS) < p.GN) tmem_ld16(0);
#pragma unroll 1
  for (int c = 0; c < EPI_CHUNKS; c++) {
    const int nc0 = n0 + cg * (16 * EPI_CHUNKS) + c * 16;
    if (nc0 >= p.GN) break;                                  // warp-uniform
    float4 bb[4];
    if (fused && p.bias) {
#pragma unroll
      for (int j = 0; j < 4; j++) bb[j] = (nc0 + 4 * j < p.GN) ? ldg4(p.bias + nc0 + 4 * j) : f4_zero();
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    float x[16];
#pragma unroll
    for (int j = 0; j < 16; j++) x[j] = has_k ? __uint_as_float(v[j]) : 0.f;
    if (c + 1 < EPI_CHUNKS && nc0 + 16 < p.GN) tmem_ld16(c + 1);
    if (fused) {    // every branch below is warp-uniform and sits OUTSIDE the per-element loops
      if (p.bias) {
#pragma unroll
        for (int j = 0; j < 4; j++) { x[4 * j] += bb[j].x; x[4 * j + 1] += bb[j].y; x[4 * j + 2] += bb[j].z; x[4 * j + 3] += bb[j].w; }
      }
      if (p.act == 1) {
#pragma unroll
        for (int j = 0; j < 16; j++) x[j] = tanhf(x[j]);
      } else if (p.act == 2) {
#pragma unroll
        for (int j = 0; j < 16; j++) x[j] = fmaxf(x[j], 0.f);
      }
      if (p.drop_p > 0.f) {
#pragma unroll
        for (int j = 0; j < 16; j++) x[j] *= dropout_scale(p.seed, (uint64_t)row * p.GN + nc0 + j, p.drop_p, inv_keep);
      }
      if (p.rowmask) {
#pragma unroll
        for (int j = 0; j < 16; j++) x[j] *= rm;
      }
    }
    if (add0) {
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        if (nc0 + j < p.GN) { const float4 a = ldg4(add0 + nc0 + j); x[j] += a.x; x[j + 1] += a.y; x[j + 2] += a.z; x[j + 3] += a.w; }
    }
    if (add1) {
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        if (nc0 + j < p.GN) { const float4 a = ldg4(add1 + nc0 + j); x[j] += a.x; x[j + 1] += a.y; x[j + 2] += a.z; x[j + 3] += a.w; }
    }
#pragma unroll
    for (int j = 0; j < 16; j += 4)
      *reinterpret_cast<float4*>(stage + lane * EPI_LD + (((j >> 2) ^ ((lane >> 1) & 3)) << 2)) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
    __syncwarp();
    float4 cs = f4_zero();
    const bool want_cs = fused && p.colsum_part != nullptr;
    const int c4 = (lane & 3) * 4, n = nc0 + c4;
#pragma unroll
    for (int it = 0; it < 4; it++) {
      const int r = it * 8 + (lane >> 2);
      const int grow = m0 + q * 32 + r;
      if (grow < p.GM && n < p.GN) {
        float4 val = *reinterpret_cast<const float4*>(stage + r * EPI_LD + ((((lane & 3)) ^ ((r >> 1) & 3)) << 2));
        if (out) {
          float* o = out + (size_t)grow * ldo + n;
          if (fused && p.accumulate) f4_add(val, *reinterpret_cast<const float4*>(o));
          if (!fused || p.store_c) st4(o, val);
        }
        if (fused && p.out_hi) {
          uint2 h, l;
          split4(val, h, l);
          *reinterpret_cast<uint2*>(p.out_hi + (size_t)grow * p.ld_planes + n) = h;
          *reinterpret_cast<uint2*>(p.out_lo + (size_t)grow * p.ld_planes + n) = l;
        }
        if (want_cs) f4_add(cs, val);
      }
    }
    if (want_cs) {     // warp-uniform branch
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
        cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
      }
      if (lane < 4 && n < p.GN) st4(p.colsum_part + ((size_t)(m0 / BM) * 4 + q) * p.GN + n, cs);
    }
    __syncwarp();
  }
}

template <bool A_MN, bool B_MN, int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
               const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl, const Params p) {
  pdl_trigger();   // the wait comes after the CTA set-up below (barriers, TMEM allocation, descriptor prefetch touch no global data)
  extern __shared__ uint8_t smem_raw[];
  constexpr int TILE_B = Cfg<BN>::TILE_B, TMEM_COLS = Cfg<BN>::TMEM_COLS;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B needs 1024-byte alignment
  uint64_t* bars = (uint64_t*)(smem + SLOTS * SLOT_BYTES);
  uint64_t* full_bar = bars;                     // [SLOTS]
  uint64_t* empty_bar = bars + SLOTS;            // [SLOTS]
  uint64_t* tfull_bar = bars + 2 * SLOTS;        // [ACC_STAGES]
  uint64_t* tempty_bar = tfull_bar + ACC_STAGES; // [ACC_STAGES]
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + ACC_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles * p.splits;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&mapAh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&mapAl) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&mapBh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&mapBl) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < SLOTS; i++) {
      mbar_init(smem_u32(&full_bar[i]), 1);
      mbar_init(smem_u32(&empty_bar[i]), 1);
    }
    for (int i = 0; i < ACC_STAGES; i++) {
      mbar_init(smem_u32(&tfull_bar[i]), 1);
      mbar_init(smem_u32(&tempty_bar[i]), EPI_WARPS);   // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // the previous kernel in the stream has completed: operands may be read, outputs overwritten

  if (warp == 0) {
    // ------------------------------------------------ TMA producer ------------------------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int split = tile / (p.m_tiles * p.n_tiles);
        const int rem = tile - split * (p.m_tiles * p.n_tiles);
        const int m0 = (rem / p.n_tiles) * BM, n0 = (rem % p.n_tiles) * BN;
        const int kb0 = split * p.kb_per_split, kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; kb++) {
          const int k0 = kb * BK;
          // slot 1: both planes of the A tile
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          uint32_t fb = smem_u32(&full_bar[stage]);
          mbar_expect_tx(fb, 2 * TILE_BYTES);
          uint32_t sa = smem_u32(smem + stage * SLOT_BYTES);
          if (A_MN) {
            tma_load_2d(sa, &mapAh, fb, m0, k0);
            tma_load_2d(sa + TILE_BYTES / 2, &mapAh, fb, m0 + 64, k0);
            tma_load_2d(sa + TILE_BYTES, &mapAl, fb, m0, k0);
            tma_load_2d(sa + TILE_BYTES + TILE_BYTES / 2, &mapAl, fb, m0 + 64, k0);
          } else {
            tma_load_2d(sa, &mapAh, fb, k0, m0);
            tma_load_2d(sa + TILE_BYTES, &mapAl, fb, k0, m0);
          }
          if (++stage == SLOTS) { stage = 0; phase ^= 1; }
          // slot(s) 2 (3): the B planes — one slot per plane for BN = 256, one slot for both for BN = 128
#pragma unroll
          for (int pl = 0; pl < 2; pl++) {
            if (BN == 256 || pl == 0) {
              mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
              fb = smem_u32(&full_bar[stage]);
              mbar_expect_tx(fb, BN == 256 ? TILE_B : 2 * TILE_B);
              sa = smem_u32(smem + stage * SLOT_BYTES);
            }
            const uint32_t dst = sa + (BN == 256 ? 0 : pl * TILE_B);
            const CUtensorMap* mb = pl ? &mapBl : &mapBh;
            if (B_MN) {
#pragma unroll
              for (int c = 0; c < BN / 64; c++) tma_load_2d(dst + c * (TILE_BYTES / 2), mb, fb, n0 + 64 * c, k0);   // 64 MN-elements x 64 k per box
            } else {
              tma_load_2d(dst, mb, fb, k0, n0);
            }
            if (BN == 256 || pl == 1) {
              if (++stage == SLOTS) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer ---------------------------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(A_MN, B_MN, BN);
      constexpr uint32_t a_kstep = A_MN ? (UMMA_K * 128) : (UMMA_K * 2);   // bytes per UMMA_K step inside a tile
      constexpr uint32_t b_kstep = B_MN ? (UMMA_K * 128) : (UMMA_K * 2);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int split = tile / (p.m_tiles * p.n_tiles);
        const int kb0 = split * p.kb_per_split, kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
        mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; kb++) {
          const int a_slot = stage;
          mbar_wait(smem_u32(&full_bar[stage]), phase);            // A planes
          const uint32_t sa = smem_u32(smem + stage * SLOT_BYTES);
          if (++stage == SLOTS) { stage = 0; phase ^= 1; }
          mbar_wait(smem_u32(&full_bar[stage]), phase);            // B_hi (BN = 256) or both B planes (BN = 128)
          tc_fence_after();
          const uint32_t sb = smem_u32(smem + stage * SLOT_BYTES);
          const int bh_slot = stage;
          if (++stage == SLOTS) { stage = 0; phase ^= 1; }
          if (BN == 256) {
            // hi plane first: A_lo·B_hi and A_hi·B_hi, then its slot goes back to the producer while the lo plane may still be in flight
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; k++) {
              const uint64_t ah = make_desc(sa + k * a_kstep, A_MN), al = make_desc(sa + TILE_BYTES + k * a_kstep, A_MN);
              const uint64_t bh = make_desc(sb + k * b_kstep, B_MN);
              umma_bf16(d_tmem, al, bh, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
              umma_bf16(d_tmem, ah, bh, idesc, 1u);
            }
            umma_commit(smem_u32(&empty_bar[bh_slot]));
            mbar_wait(smem_u32(&full_bar[stage]), phase);          // B_lo
            tc_fence_after();
            const uint32_t sl = smem_u32(smem + stage * SLOT_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; k++) {
              const uint64_t ah = make_desc(sa + k * a_kstep, A_MN);
              const uint64_t bl = make_desc(sl + k * b_kstep, B_MN);
              umma_bf16(d_tmem, ah, bl, idesc, 1u);
            }
            umma_commit(smem_u32(&empty_bar[stage]));
            umma_commit(smem_u32(&empty_bar[a_slot]));
            if (++stage == SLOTS) { stage = 0; phase ^= 1; }
          } else {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; k++) {
              const uint64_t ah = make_desc(sa + k * a_kstep, A_MN), al = make_desc(sa + TILE_BYTES + k * a_kstep, A_MN);
              const uint64_t bh = make_desc(sb + k * b_kstep, B_MN), bl = make_desc(sb + TILE_B + k * b_kstep, B_MN);
              umma_bf16(d_tmem, al, bh, idesc, (kb > kb0 || k > 0) ? 1u : 0u);   // small terms first
              umma_bf16(d_tmem, ah, bl, idesc, 1u);
              umma_bf16(d_tmem, ah, bh, idesc, 1u);
            }
            umma_commit(smem_u32(&empty_bar[bh_slot]));
            umma_commit(smem_u32(&empty_bar[a_slot]));
          }
        }
        umma_commit(smem_u32(&tfull_bar[acc]));          // accumulator complete -> epilogue
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------ epilogue (warps 2..9) ----------------------------------------
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int cg = (warp - 2) >> 2;         // which column group of the tile
    float* stage = reinterpret_cast<float*>(smem + SLOTS * SLOT_BYTES + 256) + (warp - 2) * 32 * EPI_LD;
    int acc = 0;
    uint32_t acc_phase = 0;
    const float inv_keep = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int split = tile / (p.m_tiles * p.n_tiles);
      const int rem = tile - split * (p.m_tiles * p.n_tiles);
      const int m0 = (rem / p.n_tiles) * BM, n0 = (rem % p.n_tiles) * BN;
      const int kb0 = split * p.kb_per_split, kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
      const bool has_k = kb1 > kb0;
      mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
      tc_fence_after();
      store_tile<BN>(p, tmem_base, acc, q, cg, lane, m0, n0, split, has_k, inv_keep, stage);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

// im2col image of [rows, C] token rows (sequences of S rows) as split-bf16 planes: four columns per thread, consecutive threads on
// consecutive columns of one output row (8-byte stores per plane, 16-byte loads; every input row is read `taps` times, from L1/L2)
__global__ void __launch_bounds__(256) im2col_split_kernel(const float* __restrict__ X, int64_t rows, int S, int C, int taps,
                                                           __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int64_t ld_out) {
  pdl_prologue();
  const int ld4 = (int)(ld_out >> 2), half = taps >> 1, width = taps * C;
  const int64_t total = rows * ld4;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / ld4;
    const int col = (int)(idx - r * ld4) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col < width) {
      const int j = col / C, c = col - j * C;
      const int s = (int)(r % S) + j - half;
      if (s >= 0 && s < S) v = __ldg(reinterpret_cast<const float4*>(X + (r + j - half) * C + c));
    }
    uint2 h, l;
    split4(v, h, l);
    *reinterpret_cast<uint2*>(hi + r * ld_out + col) = h;
    *reinterpret_cast<uint2*>(lo + r * ld_out + col) = l;
  }
}

// fp32 -> (hi, lo) bf16 planes, output pitch ld_out (>= cols, multiple of 8, pad columns zeroed).  A block owns SPLIT_ROWS
// rows x 256 columns (32 column-threads x 8 elements, 8 row-threads); optionally it also emits per-block column sums of the
// fp32 input (the bias gradient db = sum_m dY[m,:] rides on the pass that has to read dY anyway).
constexpr int SPLIT_ROWS = 64;
__global__ void __launch_bounds__(256) split_bf16_kernel(const float* __restrict__ X, int64_t rows, int cols, int64_t ld_in,
                                                         __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int64_t ld_out,
                                                         float* __restrict__ colsum_part) {
  pdl_prologue();
  __shared__ float red[8][256 + 8];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 256 + tx * 8;
  const int64_t r0 = (int64_t)blockIdx.y * SPLIT_ROWS;
  float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c < ld_out) {
#pragma unroll 2
    for (int i = ty; i < SPLIT_ROWS; i += 8) {
      const int64_t r = r0 + i;
      if (r >= rows) break;
      float x[8];
      if (c + 7 < cols) {
        const float4 a = ldg4_stream(X + r * ld_in + c), b2 = ldg4_stream(X + r * ld_in + c + 4);
        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b2.x; x[5] = b2.y; x[6] = b2.z; x[7] = b2.w;
      } else {
#pragma unroll
        for (int e = 0; e < 8; e++) x[e] = (c + e < cols) ? X[r * ld_in + c + e] : 0.f;
      }
      __align__(16) __nv_bfloat16 h[8], l[8];
#pragma unroll
      for (int e = 0; e < 8; e++) {
        h[e] = __float2bfloat16_rn(x[e]);
        l[e] = __float2bfloat16_rn(x[e] - __bfloat162float(h[e]));
        cs[e] += x[e];
      }
      *reinterpret_cast<uint4*>(hi + r * ld_out + c) = *reinterpret_cast<uint4*>(h);
      *reinterpret_cast<uint4*>(lo + r * ld_out + c) = *reinterpret_cast<uint4*>(l);
    }
  }
  if (colsum_part) {
#pragma unroll
    for (int e = 0; e < 8; e++) red[ty][tx * 8 + e] = cs[e];
    __syncthreads();
    const int cc = threadIdx.x;   // 256 threads <-> 256 columns of this block
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) t += red[i][cc];
    if (blockIdx.x * 256 + cc < cols) colsum_part[(int64_t)blockIdx.y * cols + blockIdx.x * 256 + cc] = t;
  }
}

// several matrices in one launch (blockIdx.y = segment): each thread converts 8 contiguous columns of one row
struct SplitSegs { lk_split_seg seg[LK_SPLIT_MAX_SEGS]; };
__global__ void __launch_bounds__(256) split_bf16_multi_kernel(const SplitSegs segs) {
  pdl_prologue();
  const lk_split_seg& g = segs.seg[blockIdx.y];
  const int64_t chunks = g.ld_out >> 3, total = g.rows * chunks;
  __nv_bfloat16* hi = (__nv_bfloat16*)g.hi;
  __nv_bfloat16* lo = (__nv_bfloat16*)g.lo;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / chunks;
    const int c = (int)(i - r * chunks) * 8;
    float x[8];
    if (c + 7 < g.cols) {
      const float4 a = ldg4(g.X + r * g.ld_in + c), b2 = ldg4(g.X + r * g.ld_in + c + 4);
      x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b2.x; x[5] = b2.y; x[6] = b2.z; x[7] = b2.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; e++) x[e] = (c + e < g.cols) ? g.X[r * g.ld_in + c + e] : 0.f;
    }
    __align__(16) __nv_bfloat16 h[8], l[8];
#pragma unroll
    for (int e = 0; e < 8; e++) {
      h[e] = __float2bfloat16_rn(x[e]);
      l[e] = __float2bfloat16_rn(x[e] - __bfloat162float(h[e]));
    }
    *reinterpret_cast<uint4*>(hi + r * g.ld_out + c) = *reinterpret_cast<uint4*>(h);
    *reinterpret_cast<uint4*>(lo + r * g.ld_out + c) = *reinterpret_cast<uint4*>(l);
  }
}

// out[c] = sum_i part[i, c] in a fixed order: 32 part-lanes x 32 columns per block, then an ordered shared-memory reduction
__global__ void __launch_bounds__(1024) colsum_finish2_kernel(const float* __restrict__ part, float* __restrict__ out, int nparts, int cols) {
  pdl_prologue();
  __shared__ float red[32][33];
  const int cx = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float s = 0.f;
  if (c < cols)
    for (int i = pl; i < nparts; i += 32) s += part[(int64_t)i * cols + c];
  red[pl][cx] = s;
  __syncthreads();
  if (pl == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 32; i++) t += red[i][cx];
    out[c] = t;
  }
}

// several column reductions in one launch (blockIdx.y = job): 32 part-lanes x 32 columns per block, ordered shared-memory combine
struct ColsumJobs { lk_colsum_job job[LK_COLSUM_MAX_JOBS]; };
__global__ void __launch_bounds__(1024) colsum_finish_multi_kernel(const ColsumJobs jobs) {
  pdl_prologue();
  __shared__ float red[32][33];
  const lk_colsum_job& j = jobs.job[blockIdx.y];
  const int cx = threadIdx.x & 31, pl = threadIdx.x >> 5;
  for (int64_t cb = blockIdx.x; cb * 32 < j.cols; cb += gridDim.x) {
    const int64_t c = cb * 32 + cx;
    float s0 = 0.f, s1 = 0.f;
    if (c < j.cols) {
      int64_t i = pl;
      // eight independent loads in flight per lane: the per-sequence partials (thousands of rows) made this loop a chain of exposed
      // memory latencies with two
      float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (; i + 224 < j.nparts; i += 256) {
#pragma unroll
        for (int u = 0; u < 8; u++) a[u] += __ldg(j.part + (i + 32 * u) * j.stride + c);
      }
      s0 = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
      for (; i + 32 < j.nparts; i += 64) { s0 += j.part[i * j.stride + c]; s1 += j.part[(i + 32) * j.stride + c]; }
      if (i < j.nparts) s0 += j.part[i * j.stride + c];
    }
    red[pl][cx] = s0 + s1;
    __syncthreads();
    if (pl == 0 && c < j.cols) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 32; i++) t += red[i][cx];
      j.out[c] = j.accumulate ? j.out[c] + t : t;
    }
    __syncthreads();
  }
}

// transposed split through a 32x32 smem tile: out[c, r]
__global__ void split_bf16_t_kernel(const float* __restrict__ X, int rows, int cols, int64_t ld_in, __nv_bfloat16* __restrict__ hi,
                                    __nv_bfloat16* __restrict__ lo, int64_t ld_out) {
  pdl_prologue();
  __shared__ float t[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int r = r0 + i, c = c0 + threadIdx.x;
    t[i][threadIdx.x] = (r < rows && c < cols) ? X[(int64_t)r * ld_in + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, r = r0 + threadIdx.x;   // output row = c, output col = r
    if (c < cols && r < ld_out) {
      float x = (r < rows) ? t[threadIdx.x][i] : 0.f;
      __nv_bfloat16 h = __float2bfloat16_rn(x);
      hi[(int64_t)c * ld_out + r] = h;
      lo[(int64_t)c * ld_out + r] = __float2bfloat16_rn(x - __bfloat162float(h));
    }
  }
}

// Tile width.  Measured on the NRMS step (profiles/r1_04): the wide tile wins a few percent at N = 768 (QKV, its weight gradient)
// and loses 10-15 % at N = 256, where 333 wide tiles quantise badly over 148 SMs and the two-stage ring covers less latency —
// so it is used for 512 <= N <= 4096 only.  LK_TC_BN=128 / 256 forces a shape (256 still needs N % 256 == 0).
static int pick_bn(int64_t GN) {
  static const int forced = [] { const char* e = getenv("LK_TC_BN"); return e ? atoi(e) : 0; }();
  if (forced == 128 || GN % 256 != 0) return 128;
  if (forced == 256) return 256;
  return (GN >= 512 && GN <= 4096) ? 256 : 128;     // (the wide tile was also measured for the 256x256xT weight gradients: 108 -> 123 us, r2_31)
}

static int pick_splits(int64_t GM, int64_t GN, int64_t GK) {
  const int BN = pick_bn(GN);
  int64_t tiles = ((GM + BM - 1) / BM) * ((GN + BN - 1) / BN);
  int64_t kb = (GK + BK - 1) / BK;
  if (tiles >= kNumSMs || kb < 8) return 1;
  int64_t s = kNumSMs / tiles;       // one wave of (tile, split) work items: twice as many only doubles the partials to write and re-read
  if (s > kb / 4) s = kb / 4;
  return (int)(s < 1 ? 1 : s);
}

}  // namespace tc
}  // namespace lk

using namespace lk;
using namespace lk::tc;

extern "C" {

size_t lk_split_bf16_workspace_bytes(int64_t rows, int64_t cols) {
  return (size_t)((rows + SPLIT_ROWS - 1) / SPLIT_ROWS) * cols * sizeof(float) + 256;
}

int lk_split_bf16(const float* X, int64_t rows, int64_t cols, int64_t ld_in, void* hi, void* lo, int64_t ld_out, int transpose,
                  float* colsum, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  LK_REQUIRE(ld_out % 8 == 0, LK_ERR_SHAPE, "lk_split_bf16: output pitch %ld must be a multiple of 8 (16-byte TMA rows)", (long)ld_out);
  if (rows == 0 || cols == 0) return LK_OK;
  if (!transpose) {
    LK_REQUIRE(ld_out >= cols && ld_in % 4 == 0, LK_ERR_SHAPE, "lk_split_bf16: bad pitches");
    LK_REQUIRE(!colsum || (workspace && workspace_bytes >= lk_split_bf16_workspace_bytes(rows, cols)), LK_ERR_ARG,
               "lk_split_bf16: workspace too small for the fused column sums");
    const int nparts = (int)((rows + SPLIT_ROWS - 1) / SPLIT_ROWS);
    dim3 grid((unsigned)((ld_out + 255) / 256), (unsigned)nparts);
    LK_LAUNCH((split_bf16_kernel), grid, 256, 0, st, X, rows, (int)cols, ld_in, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, ld_out,
                                            colsum ? (float*)workspace : nullptr);
    if (colsum) {
      LK_LAUNCH((colsum_finish2_kernel), (unsigned)((cols + 31) / 32), 1024, 0, st, (const float*)workspace, colsum, nparts, (int)cols);
      return check_launch("split_bf16", 2);
    }
  } else {
    LK_REQUIRE(ld_out >= rows && !colsum, LK_ERR_SHAPE, "lk_split_bf16: transposed pitch too small / no fused sums when transposing");
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((ld_out + 31) / 32));
    LK_LAUNCH((split_bf16_t_kernel), grid, dim3(32, 8), 0, st, X, (int)rows, (int)cols, ld_in, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, ld_out);
  }
  return check_launch("split_bf16");
}

int lk_colsum_finish_multi(const lk_colsum_job* jobs, int n_jobs, cudaStream_t st) {
  LK_REQUIRE(n_jobs >= 0 && n_jobs <= LK_COLSUM_MAX_JOBS, LK_ERR_ARG, "lk_colsum_finish_multi: %d jobs (max %d)", n_jobs, LK_COLSUM_MAX_JOBS);
  ColsumJobs cj = {};
  int n = 0;
  int64_t most = 0;
  for (int i = 0; i < n_jobs; i++) {
    const lk_colsum_job& j = jobs[i];
    LK_REQUIRE(j.cols >= 0 && j.nparts >= 0 && j.stride >= j.cols, LK_ERR_SHAPE, "lk_colsum_finish_multi: job %d has a bad shape", i);
    LK_REQUIRE(j.cols == 0 || (j.out && (j.part || j.nparts == 0)), LK_ERR_ARG, "lk_colsum_finish_multi: job %d has null pointers", i);
    if (j.cols == 0) continue;
    cj.job[n++] = j;
    if ((j.cols + 31) / 32 > most) most = (j.cols + 31) / 32;
  }
  if (n == 0) return LK_OK;
  if (most > 64) most = 64;       // wider jobs loop over their column blocks
  LK_LAUNCH((colsum_finish_multi_kernel), dim3((unsigned)most, (unsigned)n), 1024, 0, st, cj);
  return check_launch("colsum_finish_multi");
}

int lk_split_bf16_partial(const float* X, int64_t rows, int64_t cols, int64_t ld_in, void* hi, void* lo, int64_t ld_out, float* part,
                          cudaStream_t st) {
  LK_REQUIRE(ld_out % 8 == 0 && ld_out >= cols && ld_in % 4 == 0, LK_ERR_SHAPE, "lk_split_bf16_partial: bad pitches");
  LK_REQUIRE(part != nullptr, LK_ERR_ARG, "lk_split_bf16_partial: no buffer for the partial column sums");
  if (rows == 0 || cols == 0) return LK_OK;
  const int nparts = (int)((rows + SPLIT_ROWS - 1) / SPLIT_ROWS);
  dim3 grid((unsigned)((ld_out + 255) / 256), (unsigned)nparts);
  LK_LAUNCH((split_bf16_kernel), grid, 256, 0, st, X, rows, (int)cols, ld_in, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, ld_out, part);
  return check_launch("split_bf16_partial");
}

// Conv1d(k, 'same') as ONE contraction (model/operators/cnn_operator.py:54-58): the operand planes of the im2col image of X.
//   Xcol[r, j*C + c] = X[r + j - taps/2, c] when that row is inside r's sequence (fixed length S), else 0
// so that Y = Xcol · Wrᵀ with Wr[o, j*C + c] = W[o, c, j] and, with X := dY and the flipped kernel, dX = dYcol · Wdᵀ.
int lk_im2col_split_bf16(const float* X, int64_t rows, int64_t S, int64_t C, int taps, void* hi, void* lo, int64_t ld_out, cudaStream_t st) {
  LK_REQUIRE(S > 0 && rows % S == 0 && C % 4 == 0 && taps > 0 && (taps & 1), LK_ERR_SHAPE,
             "lk_im2col_split_bf16: rows=%ld S=%ld C=%ld taps=%d (C %% 4 == 0, odd taps, whole sequences)", (long)rows, (long)S, (long)C, taps);
  LK_REQUIRE(ld_out % 8 == 0 && ld_out >= taps * C, LK_ERR_SHAPE, "lk_im2col_split_bf16: output pitch %ld", (long)ld_out);
  if (rows == 0) return LK_OK;
  const int64_t total = rows * (ld_out / 4);
  int64_t blocks = (total + 255) / 256;
  if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
  LK_LAUNCH((im2col_split_kernel), (unsigned)blocks, 256, 0, st, X, rows, (int)S, (int)C, taps, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, ld_out);
  return check_launch("im2col_split_bf16");
}

int lk_split_bf16_multi(const lk_split_seg* segs, int n_segs, cudaStream_t st) {
  LK_REQUIRE(n_segs >= 0 && n_segs <= LK_SPLIT_MAX_SEGS, LK_ERR_ARG, "lk_split_bf16_multi: %d segments (max %d)", n_segs, LK_SPLIT_MAX_SEGS);
  if (n_segs == 0) return LK_OK;
  SplitSegs ss = {};
  int64_t most = 0;
  for (int i = 0; i < n_segs; i++) {
    const lk_split_seg& g = segs[i];
    LK_REQUIRE(g.ld_out % 8 == 0 && g.ld_out >= g.cols && g.ld_in % 4 == 0 && g.rows >= 0, LK_ERR_SHAPE,
               "lk_split_bf16_multi: segment %d has bad pitches (cols=%ld ld_in=%ld ld_out=%ld)", i, (long)g.cols, (long)g.ld_in, (long)g.ld_out);
    LK_REQUIRE(((uintptr_t)g.X | (uintptr_t)g.hi | (uintptr_t)g.lo) % 16 == 0, LK_ERR_ARG, "lk_split_bf16_multi: segment %d is not 16-byte aligned", i);
    ss.seg[i] = g;
    const int64_t t = g.rows * (g.ld_out >> 3);
    if (t > most) most = t;
  }
  if (most == 0) return LK_OK;
  int64_t bx = (most + 255) / 256;
  if (bx > 2 * kNumSMs) bx = 2 * kNumSMs;
  LK_LAUNCH((split_bf16_multi_kernel), dim3((unsigned)bx, (unsigned)n_segs), 256, 0, st, ss);
  return check_launch("split_bf16_multi");
}

size_t lk_tc_gemm_workspace_bytes(int64_t GM, int64_t GN, int64_t GK) {
  int s = pick_splits(GM, GN, GK);
  const size_t colsum = (size_t)((GM + BM - 1) / BM) * 4 * GN * sizeof(float);   // per-(tile, quarter) column-sum partials
  return (s > 1 ? (size_t)s * GM * GN * sizeof(float) : colsum) + 256;
}

int lk_tc_gemm_ex(const void* A_hi, const void* A_lo, int64_t lda, int a_mn, const void* B_hi, const void* B_lo, int64_t ldb, int b_mn,
                  float* C, int64_t ldc, int64_t GM, int64_t GN, int64_t GK, const lk_gemm_epilogue* ep, void* workspace,
                  size_t workspace_bytes, cudaStream_t st) {
  static const lk_gemm_epilogue none = {};
  if (!ep) ep = &none;
  LK_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && ldc % 4 == 0 && GN % 4 == 0, LK_ERR_SHAPE, "lk_tc_gemm: pitches must be 16-byte multiples");
  LK_REQUIRE(!(a_mn && !b_mn), LK_ERR_ARG, "lk_tc_gemm: MN-major A with K-major B is not instantiated");
  LK_REQUIRE(((uintptr_t)A_hi | (uintptr_t)A_lo | (uintptr_t)B_hi | (uintptr_t)B_lo | (uintptr_t)C | (uintptr_t)ep->out_hi |
              (uintptr_t)ep->out_lo) % 16 == 0, LK_ERR_ARG, "lk_tc_gemm: operands must be 16-byte aligned");
  LK_REQUIRE(C || ep->out_hi || ep->colsum || ep->colsum_part, LK_ERR_ARG, "lk_tc_gemm: no output requested");
  LK_REQUIRE(!ep->out_hi || (ep->out_lo && ep->ld_planes % 8 == 0 && ep->ld_planes >= GN), LK_ERR_ARG, "lk_tc_gemm: bad output planes");
  LK_REQUIRE(!ep->accumulate || C, LK_ERR_ARG, "lk_tc_gemm: accumulate needs C");
  if (GM == 0 || GN == 0) return LK_OK;
  CUtensorMap mAh, mAl, mBh, mBl;
  int rc;
  const int BN = pick_bn(GN);
  if (a_mn) {   // A stored [GK, GM] (GM contiguous)
    if ((rc = make_map(&mAh, A_hi, GM, GK, lda, 64))) return rc;
    if ((rc = make_map(&mAl, A_lo, GM, GK, lda, 64))) return rc;
  } else {      // A stored [GM, GK] (GK contiguous)
    if ((rc = make_map(&mAh, A_hi, GK, GM, lda, BM))) return rc;
    if ((rc = make_map(&mAl, A_lo, GK, GM, lda, BM))) return rc;
  }
  if (b_mn) {   // B stored [GK, GN] (GN contiguous)
    if ((rc = make_map(&mBh, B_hi, GN, GK, ldb, 64))) return rc;
    if ((rc = make_map(&mBl, B_lo, GN, GK, ldb, 64))) return rc;
  } else {      // B stored [GN, GK] (GK contiguous)
    if ((rc = make_map(&mBh, B_hi, GK, GN, ldb, BN))) return rc;
    if ((rc = make_map(&mBl, B_lo, GK, GN, ldb, BN))) return rc;
  }
  Params p;
  p.C = C; p.GM = (int)GM; p.GN = (int)GN; p.GK = (int)GK; p.ldc = (int)ldc;
  p.m_tiles = (int)((GM + BM - 1) / BM);
  p.n_tiles = (int)((GN + BN - 1) / BN);
  p.k_blocks = (int)((GK + BK - 1) / BK);
  const bool plain = !ep->bias && !ep->rowmask && ep->act == 0 && ep->drop_p == 0.f && !ep->add_ids0 && !ep->add_ids1 && !ep->out_hi &&
                     !ep->colsum && !ep->colsum_part;
  p.splits = (plain && C) ? pick_splits(GM, GN, GK) : 1;       // a fused epilogue needs the whole reduction in one accumulator
  p.kb_per_split = (p.k_blocks + p.splits - 1) / p.splits;
  p.splits = (p.k_blocks + p.kb_per_split - 1) / p.kb_per_split;   // no empty splits
  p.partial = nullptr;
  if (p.splits > 1) {
    LK_REQUIRE(workspace && workspace_bytes >= (size_t)p.splits * GM * GN * sizeof(float), LK_ERR_ARG, "lk_tc_gemm: workspace too small");
    LK_REQUIRE(plain && C, LK_ERR_ARG, "lk_tc_gemm: split reduction has no fused epilogue");
    p.partial = (float*)workspace;
  }
  p.bias = ep->bias; p.rowmask = ep->rowmask; p.rowmask_is_ids = ep->rowmask_is_ids; p.act = ep->act; p.accumulate = ep->accumulate;
  p.store_c = (C && !ep->store_c_off) ? 1 : 0;
  p.drop_p = ep->drop_p; p.seed = (unsigned long long)ep->seed;
  p.add_ids[0] = ep->add_ids0; p.add_tab[0] = ep->add_tab0; p.add_ids[1] = ep->add_ids1; p.add_tab[1] = ep->add_tab1;
  p.out_hi = (__nv_bfloat16*)ep->out_hi; p.out_lo = (__nv_bfloat16*)ep->out_lo; p.ld_planes = (int)ep->ld_planes;
  p.colsum_part = nullptr;
  LK_REQUIRE(!(ep->colsum && ep->colsum_part), LK_ERR_ARG, "lk_tc_gemm: colsum and colsum_part are mutually exclusive");
  if (ep->colsum_part) p.colsum_part = ep->colsum_part;      // deferred: the caller finishes the partials
  if (ep->colsum) {
    LK_REQUIRE(workspace && workspace_bytes >= (size_t)p.m_tiles * 4 * GN * sizeof(float), LK_ERR_ARG, "lk_tc_gemm: workspace too small for column sums");
    p.colsum_part = (float*)workspace;
  }

  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(tc_gemm_kernel<false, false, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaFuncSetAttribute(tc_gemm_kernel<false, true, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaFuncSetAttribute(tc_gemm_kernel<true, true, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaFuncSetAttribute(tc_gemm_kernel<false, false, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaFuncSetAttribute(tc_gemm_kernel<false, true, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaFuncSetAttribute(tc_gemm_kernel<true, true, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    attr_set = true;
  }
  int total = p.m_tiles * p.n_tiles * p.splits;
  int grid = total < kNumSMs ? total : kNumSMs;
  if (BN == 256) {
    if (a_mn) LK_LAUNCH((tc_gemm_kernel<true, true, 256>), grid, NUM_THREADS, SMEM_BYTES, st, mAh, mAl, mBh, mBl, p);
    else if (b_mn) LK_LAUNCH((tc_gemm_kernel<false, true, 256>), grid, NUM_THREADS, SMEM_BYTES, st, mAh, mAl, mBh, mBl, p);
    else LK_LAUNCH((tc_gemm_kernel<false, false, 256>), grid, NUM_THREADS, SMEM_BYTES, st, mAh, mAl, mBh, mBl, p);
  } else {
    if (a_mn) LK_LAUNCH((tc_gemm_kernel<true, true, 128>), grid, NUM_THREADS, SMEM_BYTES, st, mAh, mAl, mBh, mBl, p);
    else if (b_mn) LK_LAUNCH((tc_gemm_kernel<false, true, 128>), grid, NUM_THREADS, SMEM_BYTES, st, mAh, mAl, mBh, mBl, p);
    else LK_LAUNCH((tc_gemm_kernel<false, false, 128>), grid, NUM_THREADS, SMEM_BYTES, st, mAh, mAl, mBh, mBl, p);
  }
  rc = check_launch("tc_gemm");
  if (rc) return rc;
  if (p.splits > 1) return lk_splitk_reduce(p.partial, C, GM, GN, ldc, p.splits, ep->accumulate, st);
  if (ep->colsum) {
    LK_LAUNCH((colsum_finish2_kernel), (unsigned)((GN + 31) / 32), 1024, 0, st, p.colsum_part, ep->colsum, p.m_tiles * 4, (int)GN);
    return check_launch("tc_gemm_colsum");
  }
  return LK_OK;
}

int lk_tc_gemm(const void* A_hi, const void* A_lo, int64_t lda, int a_mn, const void* B_hi, const void* B_lo, int64_t ldb, int b_mn,
               float* C, int64_t ldc, int64_t GM, int64_t GN, int64_t GK, const float* bias, const int64_t* rowmask, int act,
               float drop_p, uint64_t seed, int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  lk_gemm_epilogue ep = {};
  ep.bias = bias; ep.rowmask = rowmask; ep.act = act; ep.drop_p = drop_p; ep.seed = seed; ep.accumulate = accumulate;
  return lk_tc_gemm_ex(A_hi, A_lo, lda, a_mn, B_hi, B_lo, ldb, b_mn, C, ldc, GM, GN, GK, &ep, workspace, workspace_bytes, st);
}

}  // extern "C"
