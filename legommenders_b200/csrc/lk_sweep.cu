// Catalog scoring sweep with the per-user top-k fused into the epilogue (BASELINE config 5: U users x 1M LLM-embedded items).
//
// Reference path: scores = DotPredictor(user, item) for every (user, item) pair (model/predictors/dot_predictor.py:7-10 over the projected item
// table, model/legommender.py:239-246), followed by a ranking.  Materialised that is U x N fp32 (16.8 GB for 4096 x 1M) written once and read
// once; here the scores never leave the SM:
//
//   work item = (128-user tile, contiguous item range).  The user tile's split-bf16 planes stay resident in shared memory (A operand,
//   2 x 64 KiB); the range's item planes stream through a 6-stage TMA ring of single-plane [128 items x 64 k] tiles (B operand, K-major);
//   each 128 x 128 score tile accumulates in one of FOUR TMEM buffers (3 MMAs per product: lo*hi, hi*lo, hi*hi), so the MMA issuer runs up
//   to three tiles ahead of the four epilogue warps.  An epilogue thread owns one user row: it reads the row's 128 scores
//   (tcgen05.ld 32x32b) and keeps that user's k best (value, item) pairs of the whole range in registers — a score enters only if it
//   beats the current k-th best, which after the first few tiles is rare.  Per work item [128, k] pairs are written; the ranges (and, with
//   several GPUs, the ranks' item shards) are merged by a k-way selection over R*k candidates per user on the host side of the ABI.
//   Ties are broken towards the smaller item index (torch.topk / a stable descending sort of the full score row).
#include <math_constants.h>
#include <stdlib.h>

#include "lk_tc.cuh"
#include "../../include/legommenders_b200.h"

namespace lk {
namespace sweep {

using namespace lk::tc;

constexpr int KD = 256;                          // representation width (hidden_size)
constexpr int KB = KD / BK;                      // 4 k-chunks
constexpr int TN = 128;                          // items per score tile (MMA N)
constexpr int B_STAGES = 6;
constexpr int B_STAGE_BYTES = TN * BK * 2;       // one plane of a [128 items x 64 k] tile: 16 KiB
constexpr int A_PLANE_BYTES = KB * TILE_BYTES;   // 64 KiB
constexpr int ACC = 4;                           // TMEM accumulator buffers (4 x 128 columns)
constexpr int EPI_WARPS = 4;
constexpr int FIRST_EPI_WARP = 4;                // warps 4..7: warp & 3 = TMEM lane quarter
constexpr int NUM_THREADS = (FIRST_EPI_WARP + EPI_WARPS) * 32;
constexpr int MAXK = LK_SWEEP_MAX_K;
constexpr int SMEM_BYTES = 2 * A_PLANE_BYTES + B_STAGES * B_STAGE_BYTES + 1024 + 256;
static_assert(SMEM_BYTES <= 227 * 1024, "sweep kernel shared memory budget");

struct Params {
  int U, N;                 // users, items (of this shard)
  int u_tiles, ranges, tiles_per_range;   // work items = u_tiles * ranges; a range = tiles_per_range item tiles of 128
  int k, dbg;
  float* out_val;           // [ranges, U, k]
  int32_t* out_idx;         // [ranges, U, k]   item index within this shard
};
struct Maps { CUtensorMap u_hi, u_lo, i_hi, i_lo; };

__global__ void __launch_bounds__(NUM_THREADS, 1) sweep_topk_kernel(const __grid_constant__ Maps maps, const Params p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* a_smem = smem;
  uint8_t* b_smem = smem + 2 * A_PLANE_BYTES;
  uint64_t* bars = (uint64_t*)(b_smem + B_STAGES * B_STAGE_BYTES);
  uint64_t* full_bar = bars;                     // [B_STAGES]
  uint64_t* empty_bar = full_bar + B_STAGES;     // [B_STAGES]
  uint64_t* a_full_bar = empty_bar + B_STAGES;   // [1] user tile landed
  uint64_t* a_free_bar = a_full_bar + 1;         // [1] every MMA of the work item has read the user tile
  uint64_t* tfull_bar = a_free_bar + 1;          // [ACC]
  uint64_t* tempty_bar = tfull_bar + ACC;        // [ACC]
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + ACC);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_work = p.u_tiles * p.ranges;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&maps.u_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&maps.u_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&maps.i_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&maps.i_lo) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < B_STAGES; i++) { mbar_init(smem_u32(&full_bar[i]), 1); mbar_init(smem_u32(&empty_bar[i]), 1); }
    mbar_init(smem_u32(a_full_bar), 1);
    mbar_init(smem_u32(a_free_bar), 1);
    for (int i = 0; i < ACC; i++) { mbar_init(smem_u32(&tfull_bar[i]), 1); mbar_init(smem_u32(&tempty_bar[i]), EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  // work items are ordered range-major: the CTAs running at any moment share a few item ranges (L2 reuse of the item planes)
  if (warp == 0) {
    // ------------------------------------------- item-tile producer ---------------------------------------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int range = w / p.u_tiles;
        for (int t = 0; t < p.tiles_per_range; t++) {
          const int n0 = (range * p.tiles_per_range + t) * TN;
          if (n0 >= p.N) break;
          for (int kb = 0; kb < KB; kb++) {
#pragma unroll
            for (int pl = 0; pl < 2; pl++) {
              mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
              const uint32_t fb = smem_u32(&full_bar[stage]);
              mbar_expect_tx(fb, B_STAGE_BYTES);
              tma_load_2d(smem_u32(b_smem + stage * B_STAGE_BYTES), pl ? &maps.i_lo : &maps.i_hi, fb, kb * BK, n0);
              if (++stage == B_STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------- user-tile producer ---------------------------------------------------------------
    if (lane == 0) {
      uint32_t wphase = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int m0 = (w % p.u_tiles) * BM;
        mbar_wait(smem_u32(a_free_bar), wphase ^ 1);
        const uint32_t fb = smem_u32(a_full_bar);
        mbar_expect_tx(fb, 2 * A_PLANE_BYTES);
        for (int kb = 0; kb < KB; kb++) {
          tma_load_2d(smem_u32(a_smem + kb * TILE_BYTES), &maps.u_hi, fb, kb * BK, m0);
          tma_load_2d(smem_u32(a_smem + A_PLANE_BYTES + kb * TILE_BYTES), &maps.u_lo, fb, kb * BK, m0);
        }
        wphase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------- MMA issuer ------------------------------------------------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(false, false, TN);
      int stage = 0;
      uint32_t phase = 0, wphase = 0, n = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int range = w / p.u_tiles;
        mbar_wait(smem_u32(a_full_bar), wphase);
        tc_fence_after();
        for (int t = 0; t < p.tiles_per_range; t++, n++) {
          if ((range * p.tiles_per_range + t) * TN >= p.N) break;
          const uint32_t acc = n % ACC, acc_phase = (n / ACC) & 1;
          mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * TN;
          for (int kb = 0; kb < KB; kb++) {
            const uint32_t sa = smem_u32(a_smem + kb * TILE_BYTES);
            mbar_wait(smem_u32(&full_bar[stage]), phase);            // item hi plane
            tc_fence_after();
            uint32_t sb = smem_u32(b_smem + stage * B_STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; k++) {
              const uint64_t ah = make_desc(sa + k * (UMMA_K * 2), false), al = make_desc(sa + A_PLANE_BYTES + k * (UMMA_K * 2), false);
              const uint64_t bh = make_desc(sb + k * (UMMA_K * 2), false);
              umma_bf16(d_tmem, al, bh, idesc, (kb > 0 || k > 0) ? 1u : 0u);
              umma_bf16(d_tmem, ah, bh, idesc, 1u);
            }
            umma_commit(smem_u32(&empty_bar[stage]));
            if (++stage == B_STAGES) { stage = 0; phase ^= 1; }
            mbar_wait(smem_u32(&full_bar[stage]), phase);            // item lo plane
            tc_fence_after();
            sb = smem_u32(b_smem + stage * B_STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; k++) {
              const uint64_t ah = make_desc(sa + k * (UMMA_K * 2), false);
              const uint64_t bl = make_desc(sb + k * (UMMA_K * 2), false);
              umma_bf16(d_tmem, ah, bl, idesc, 1u);
            }
            umma_commit(smem_u32(&empty_bar[stage]));
            if (++stage == B_STAGES) { stage = 0; phase ^= 1; }
          }
          umma_commit(smem_u32(&tfull_bar[acc]));
        }
        umma_commit(smem_u32(a_free_bar));                            // the user tile may be replaced
        wphase ^= 1;
      }
    }
  } else if (warp >= FIRST_EPI_WARP) {
    // ------------------------------------------- epilogue: one thread per user row ---------------------------------------------------
    const int q = warp & 3;
    uint32_t n = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      const int range = w / p.u_tiles, ut = w % p.u_tiles;
      const int row = ut * BM + q * 32 + lane;
      float bv[MAXK];
      int bi[MAXK];
#pragma unroll
      for (int j = 0; j < MAXK; j++) { bv[j] = -CUDART_INF_F; bi[j] = 0x7fffffff; }
      float thr = -CUDART_INF_F;      // current k-th best (bv[k-1])
      for (int t = 0; t < p.tiles_per_range; t++, n++) {
        const int n0 = (range * p.tiles_per_range + t) * TN;
        if (n0 >= p.N) break;
        const uint32_t acc = n % ACC, acc_phase = (n / ACC) & 1;
        mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * TN;
#pragma unroll 1
        for (int c = 0; c < TN; c += 32) {
          uint32_t v[32];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, "
              "%28, %29, %30, %31}, [%32];"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
                "=r"(v[31])
              : "r"(taddr + (uint32_t)c));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const int lim = p.N - (n0 + c);                       // columns past the last item are TMA zero fill: not candidates
          // hot path: a 32-bit mask of this row's scores above its current k-th best — register compares only, almost always zero after the
          // first few tiles.  (Written as a plain per-element `if`, the compiler if-converts the insertion and every element pays for it: 36 ms
          // for the 1M sweep; a warp-wide scan of all 32 elements whenever ANY lane has a candidate: 14 ms; this: the lanes that have
          // candidates walk only their own.)
          uint32_t m = 0;
#pragma unroll
          for (int e = 0; e < 32; e++) m |= (__uint_as_float(v[e]) > thr ? 1u : 0u) << e;
          if (lim < 32) m &= lim <= 0 ? 0u : (0xffffffffu >> (32 - lim));
          if (p.dbg != 1 && m != 0 && p.dbg != 2) {                 // divergent: only lanes with candidates enter
            float loc[32];                                          // dynamic indexing of the chunk lives in local memory, on this path only
#pragma unroll
            for (int e = 0; e < 32; e++) loc[e] = __uint_as_float(v[e]);
            while (m) {
              const int e = __ffs(m) - 1;
              m &= m - 1;
              const float s = loc[e];
              if (s > thr) {                                        // thr may have risen since the mask was taken
                // insert (s, item) keeping bv descending; equal scores keep the earlier (smaller) item first
                float cv = s;
                int ci = n0 + c + e;
                bool ins = false;
#pragma unroll
                for (int j = 0; j < MAXK; j++) {
                  if (j < p.k) {
                    ins = ins || cv > bv[j];          // strict: a new score equal to a kept one goes AFTER it (items arrive in index order)
                    if (ins) { const float tv = bv[j]; const int ti = bi[j]; bv[j] = cv; bi[j] = ci; cv = tv; ci = ti; }
                  }
                }
#pragma unroll
                for (int j = 0; j < MAXK; j++)
                  if (j == p.k - 1) thr = bv[j];
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));
      }
      if (row < p.U) {
        float* ov = p.out_val + ((size_t)range * p.U + row) * p.k;
        int32_t* oi = p.out_idx + ((size_t)range * p.U + row) * p.k;
#pragma unroll
        for (int j = 0; j < MAXK; j++)
          if (j < p.k) { ov[j] = bv[j]; oi[j] = bi[j]; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

}  // namespace sweep
}  // namespace lk

using namespace lk;
using namespace lk::sweep;

extern "C" {

// number of item ranges the sweep is cut into for U users and N items (so that u_tiles * ranges work items fill the SMs a few times over)
int64_t lk_sweep_ranges(int64_t U, int64_t N) {
  const int64_t u_tiles = (U + BM - 1) / BM, tiles = (N + TN - 1) / TN;
  int64_t r = (4 * kNumSMs + u_tiles - 1) / u_tiles;          // ~4 work items per SM ...
  for (int64_t c = r; c <= 2 * r + 1; c++)                    // ... and, when a nearby count makes u_tiles * ranges a multiple of the SM count, that one
    if ((u_tiles * c) % kNumSMs == 0) { r = c; break; }
  if (r > tiles) r = tiles;
  if (r < 1) r = 1;
  const int64_t per = (tiles + r - 1) / r;                     // tiles per range -> the number of NON-EMPTY ranges of that size
  return (tiles + per - 1) / per;
}

int lk_sweep_topk(const void* U_hi, const void* U_lo, int64_t ldu, int64_t U, const void* I_hi, const void* I_lo, int64_t ldi, int64_t N, int64_t D,
                  int k, int64_t ranges, float* out_val, int32_t* out_idx, cudaStream_t st) {
  LK_REQUIRE(D == KD, LK_ERR_SHAPE, "lk_sweep_topk: representation width %ld (this kernel is specialised to %d)", (long)D, KD);
  LK_REQUIRE(k >= 1 && k <= MAXK, LK_ERR_ARG, "lk_sweep_topk: k=%d (1..%d)", k, MAXK);
  LK_REQUIRE(ldu % 8 == 0 && ldi % 8 == 0 && ldu >= D && ldi >= D, LK_ERR_SHAPE, "lk_sweep_topk: plane pitches");
  LK_REQUIRE(U > 0 && N > 0 && N < ((int64_t)1 << 31) && ranges >= 1, LK_ERR_ARG, "lk_sweep_topk: U=%ld N=%ld ranges=%ld", (long)U, (long)N, (long)ranges);
  Maps maps;
  int rc;
  if ((rc = make_map(&maps.u_hi, U_hi, D, U, ldu, BM))) return rc;
  if ((rc = make_map(&maps.u_lo, U_lo, D, U, ldu, BM))) return rc;
  if ((rc = make_map(&maps.i_hi, I_hi, D, N, ldi, TN))) return rc;
  if ((rc = make_map(&maps.i_lo, I_lo, D, N, ldi, TN))) return rc;
  Params p;
  p.U = (int)U; p.N = (int)N; p.k = k;
  p.u_tiles = (int)((U + BM - 1) / BM);
  const int64_t tiles = (N + TN - 1) / TN;
  p.ranges = (int)ranges;
  p.tiles_per_range = (int)((tiles + ranges - 1) / ranges);
  LK_REQUIRE((int64_t)p.tiles_per_range * (ranges - 1) < tiles, LK_ERR_ARG, "lk_sweep_topk: %ld ranges leave some empty (%ld item tiles)", (long)ranges, (long)tiles);
  p.out_val = out_val; p.out_idx = out_idx;
  { const char* e = getenv("LK_SWEEP_DBG"); p.dbg = e ? atoi(e) : 0; }
  static bool attr_set = false;
  if (!attr_set) { cudaFuncSetAttribute(sweep_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES); attr_set = true; }
  const int n_work = p.u_tiles * p.ranges;
  LK_LAUNCH((sweep_topk_kernel), n_work < kNumSMs ? n_work : kNumSMs, NUM_THREADS, SMEM_BYTES, st, maps, p);
  return check_launch("sweep_topk");
}

}  // extern "C"
