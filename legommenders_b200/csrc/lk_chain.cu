// Fused chain of up to three 256x256 contractions over the same 128-row tile — the K = 256 GEMM chains of an NRMS encoder:
//   forward   ctx -> out_proj -> linear -> W1 (+tanh, + w2 row dot)      model/operators/attention_operator.py:55-58, model/common/attention.py:31-33
//   backward  dpre -> dlin (+ alpha*drep) -> dout -> dctx                  (autograd of the same three nn.Linear)
// Run as three separate GEMMs every intermediate tile makes a round trip through HBM/L2 and every launch pays the fixed
// per-tile cost of a K = 256 problem (4 k-blocks).  Here ONE persistent CTA per SM keeps the 128 x 256 intermediate on chip:
//
//   TMEM   : two fp32 accumulators of 128 x 256 (2 x 256 columns); GEMM n accumulates into buffer n & 1 while the epilogue
//            warps drain buffer (n-1) & 1.
//   smem   : the A operand of the running GEMM as split-bf16 planes, 4 k-chunks of [128 rows x 64 k] (K-major, 128-byte
//            swizzle; 2 x 64 KiB) + a 6-stage ring of single-plane weight tiles [128 n x 64 k] (16 KiB per stage).
//   roles  : warp 0 weight-tile TMA producer | warp 1 tcgen05.mma issuer (M128 N128 K16, 3 MMAs per product: lo*hi, hi*lo,
//            hi*hi) | warp 2 A-chunk TMA producer (first GEMM of a tile) | warps 3-18 epilogue.
//   epilogue of GEMM g, k-chunk by k-chunk (all 16 warps on the same 64 columns): tcgen05.ld -> bias / + fp32 addend / tanh ->
//            column-sum partials, w2 row-dot partials -> the (hi, lo) planes of the tile are written straight into the A buffer in
//            the swizzled UMMA layout (st.shared + fence.proxy.async) and the chunk's mbarrier is signalled: GEMM g+1 starts on
//            k-chunk 0 while chunks 1..3 are still being drained.  The same shared-memory chunk is the source of the planes' GLOBAL
//            image (TMA store by one elected thread: no LSU traffic at all).  The accumulator is read with tcgen05.ld.16x256b (MMA-fragment
//            layout: four lanes hold 32 contiguous bytes of a row), so fp32 results, addends and the swizzled shared-memory writes are
//            sector-exact and bank-conflict free without any transpose.  (Row-per-lane 32x32b loads made the first version LSU-bound —
//            32 lines per store instruction — and a shuffle transpose made the second one issue-bound: 133 / 125 us against 3 x 39 us.)
//   A-chunk recycling: the first GEMM of the NEXT tile streams its A chunks by TMA into the same buffer, chunk kb as soon
//            as the last GEMM of the current tile has consumed it (tcgen05.commit per k-chunk).
#include <stdlib.h>

#include "lk_tc.cuh"
#include "../../include/legommenders_b200.h"

namespace lk {
namespace chain {

using namespace lk::tc;

constexpr int ND = 256;                          // every contraction of the chain is [128 x 256] x [256 x 256]
constexpr int KB = ND / BK;                      // 4 k-chunks
constexpr int HALF = 128;                        // MMA N (one half of the accumulator per weight tile)
constexpr int B_STAGES = 6;
constexpr int B_STAGE_BYTES = HALF * BK * 2;     // ONE plane of a [128 n x 64 k] weight tile: 16 KiB (hi and lo are separate stages:
                                                 // 80 KiB in flight per SM hide the L2 latency that a 3 x 32 KiB ring did not)
constexpr int A_PLANE_BYTES = KB * TILE_BYTES;   // 64 KiB per plane
constexpr int EPI_WARPS = 16;
constexpr int FIRST_EPI_WARP = 3;
constexpr int NUM_THREADS = (FIRST_EPI_WARP + EPI_WARPS) * 32;   // 608
constexpr int MAX_G = 3;
constexpr int N_BARS = 2 * B_STAGES + 3 * KB + 4;
constexpr int SMEM_BYTES = 2 * A_PLANE_BYTES + B_STAGES * B_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers + tmem slot*/;
static_assert(N_BARS * 8 + 8 <= 256, "barrier block");
static_assert(SMEM_BYTES <= 227 * 1024, "chain kernel shared memory budget");

struct Stage {
  const float* bias;         // [256] or null
  const float* addsrc;       // fp32 [M, 256] added to the accumulator, or null
  int act;                   // 0 none, 1 tanh
  float* out_f32;            // [M, 256] or null
  int out_planes;            // 1: split-bf16 image of the result to global (TMA store from the A buffer; stages that feed a next one only)
  float* colsum_part;        // [m_tiles * 4, 256] or null
  const float* dotvec;       // [256] or null: rowdot_part[row, 0..3] = partial sums of result[row, :] . dotvec
  float* rowdot_part;        // [M, 4]
};
struct Params {
  int M, m_tiles, ng;
  long long* trace;          // debug (LK_CHAIN_TRACE): clock64 stamps of CTA 0's roles, 256 slots per role

  Stage st[MAX_G];
};
struct Maps {
  CUtensorMap a_hi, a_lo, b_hi[MAX_G], b_lo[MAX_G], o_hi[MAX_G - 1], o_lo[MAX_G - 1];   // o_*: plane outputs of the stages that feed a next one
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)map), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}

#define TRACE(role, idx) do { if (p.trace && blockIdx.x == 0 && (idx) < 256) p.trace[(role) * 256 + (idx)] = clock64(); } while (0)

template <bool B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1) chain_kernel(const __grid_constant__ Maps maps, const Params p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* a_smem = smem;                                   // [hi: 4 chunks x 16 KiB][lo: 4 chunks x 16 KiB]
  uint8_t* b_smem = smem + 2 * A_PLANE_BYTES;               // B_STAGES x 32 KiB
  uint64_t* bars = (uint64_t*)(b_smem + B_STAGES * B_STAGE_BYTES);
  uint64_t* full_bar = bars;                                // [B_STAGES]   weight tile landed (tx)
  uint64_t* empty_bar = full_bar + B_STAGES;                // [B_STAGES]   weight tile consumed (commit)
  uint64_t* a_tma_bar = empty_bar + B_STAGES;               // [KB]         A chunk landed by TMA (first GEMM of a tile)
  uint64_t* a_epi_bar = a_tma_bar + KB;                     // [KB]         A chunk written by the epilogue warps (later GEMMs)
  uint64_t* a_free_bar = a_epi_bar + KB;                    // [KB]         A chunk consumed by the last GEMM of the tile (commit)
  uint64_t* tfull_bar = a_free_bar + KB;                    // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                     // [2]
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ng = p.ng;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&maps.a_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&maps.a_lo) : "memory");
    for (int g = 0; g < ng; g++) {
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&maps.b_hi[g]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&maps.b_lo[g]) : "memory");
      if (g + 1 < ng && p.st[g].out_planes) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&maps.o_hi[g]) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&maps.o_lo[g]) : "memory");
      }
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < B_STAGES; i++) { mbar_init(smem_u32(&full_bar[i]), 1); mbar_init(smem_u32(&empty_bar[i]), 1); }
    for (int i = 0; i < KB; i++) {
      mbar_init(smem_u32(&a_tma_bar[i]), 1);
      mbar_init(smem_u32(&a_epi_bar[i]), EPI_WARPS);
      mbar_init(smem_u32(&a_free_bar[i]), 2);      // tcgen05.commit of the last GEMM + the plane-store thread (its TMA store has read the chunk)
    }
    for (int i = 0; i < 2; i++) { mbar_init(smem_u32(&tfull_bar[i]), 1); mbar_init(smem_u32(&tempty_bar[i]), EPI_WARPS); }
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
  if (threadIdx.x == 0) TRACE(3, 0);

  if (warp == 0) {
    // ------------------------------------------- weight-tile producer: (tile, g, kb, half) in MMA order -----------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
        for (int g = 0; g < ng; g++) {
          const CUtensorMap* mh = &maps.b_hi[g];
          const CUtensorMap* ml = &maps.b_lo[g];
          for (int kb = 0; kb < KB; kb++) {
#pragma unroll
            for (int hp = 0; hp < 4; hp++) {          // (half 0, hi) (half 0, lo) (half 1, hi) (half 1, lo)
              const CUtensorMap* m = (hp & 1) ? ml : mh;
              mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
              const uint32_t fb = smem_u32(&full_bar[stage]);
              mbar_expect_tx(fb, B_STAGE_BYTES);
              TRACE(2, ((tile / gridDim.x * ng + g) * KB + kb) * 4 + hp);
              const uint32_t sb = smem_u32(b_smem + stage * B_STAGE_BYTES);
              const int k0 = kb * BK, n0 = (hp >> 1) * HALF;
              if (B_MN) {        // W stored [K, N] (N contiguous): boxes of 64 n x 64 k
                tma_load_2d(sb, m, fb, n0, k0);
                tma_load_2d(sb + TILE_BYTES / 2, m, fb, n0 + 64, k0);
              } else {           // W stored [N, K] (K contiguous): one box of 64 k x 128 n
                tma_load_2d(sb, m, fb, k0, n0);
              }
              if (++stage == B_STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------- A-chunk producer (first GEMM of every tile) --------------------------------
    if (lane == 0) {
      uint32_t tphase = 0;        // a_free completes once per tile, a_tma once per tile
      for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
        const int m0 = tile * BM;
        for (int kb = 0; kb < KB; kb++) {
          mbar_wait(smem_u32(&a_free_bar[kb]), tphase ^ 1);       // the previous tile's last GEMM has consumed this chunk
          const uint32_t fb = smem_u32(&a_tma_bar[kb]);
          mbar_expect_tx(fb, 2 * TILE_BYTES);
          tma_load_2d(smem_u32(a_smem + kb * TILE_BYTES), &maps.a_hi, fb, kb * BK, m0);
          tma_load_2d(smem_u32(a_smem + A_PLANE_BYTES + kb * TILE_BYTES), &maps.a_lo, fb, kb * BK, m0);
        }
        tphase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------- MMA issuer ---------------------------------------------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(false, B_MN, HALF);
      constexpr uint32_t b_kstep = B_MN ? (UMMA_K * 128) : (UMMA_K * 2);
      int stage = 0;
      uint32_t phase = 0, tphase = 0, ephase = 0;
      uint32_t n = 0;             // running GEMM count of this CTA -> accumulator buffer n & 1
      for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
        for (int g = 0; g < ng; g++, n++) {
          const uint32_t acc = n & 1, acc_phase = (n >> 1) & 1;
          mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1);
          tc_fence_after();
          for (int kb = 0; kb < KB; kb++) {
            if (g == 0) mbar_wait(smem_u32(&a_tma_bar[kb]), tphase);
            else mbar_wait(smem_u32(&a_epi_bar[kb]), ephase);
            tc_fence_after();
            TRACE(0, (n * KB + kb) * 3);
            const uint32_t sa = smem_u32(a_smem + kb * TILE_BYTES);
#pragma unroll
            for (int h = 0; h < 2; h++) {
              const uint32_t d_tmem = tmem_base + acc * ND + h * HALF;
              // hi plane of the weight tile: A_lo·B_hi and A_hi·B_hi
              mbar_wait(smem_u32(&full_bar[stage]), phase);
              tc_fence_after();
              if (h == 0) TRACE(0, (n * KB + kb) * 3 + 1);
              uint32_t sb = smem_u32(b_smem + stage * B_STAGE_BYTES);
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; k++) {
                const uint64_t ah = make_desc(sa + k * (UMMA_K * 2), false);
                const uint64_t al = make_desc(sa + A_PLANE_BYTES + k * (UMMA_K * 2), false);
                const uint64_t bh = make_desc(sb + k * b_kstep, B_MN);
                umma_bf16(d_tmem, al, bh, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                umma_bf16(d_tmem, ah, bh, idesc, 1u);
              }
              umma_commit(smem_u32(&empty_bar[stage]));
              if (++stage == B_STAGES) { stage = 0; phase ^= 1; }
              // lo plane: A_hi·B_lo
              mbar_wait(smem_u32(&full_bar[stage]), phase);
              tc_fence_after();
              sb = smem_u32(b_smem + stage * B_STAGE_BYTES);
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; k++) {
                const uint64_t ah = make_desc(sa + k * (UMMA_K * 2), false);
                const uint64_t bl = make_desc(sb + k * b_kstep, B_MN);
                umma_bf16(d_tmem, ah, bl, idesc, 1u);
              }
              umma_commit(smem_u32(&empty_bar[stage]));
              if (++stage == B_STAGES) { stage = 0; phase ^= 1; }
            }
            TRACE(0, (n * KB + kb) * 3 + 2);
            if (g == ng - 1) umma_commit(smem_u32(&a_free_bar[kb]));          // chunk kb may be refilled for the next tile
          }
          umma_commit(smem_u32(&tfull_bar[acc]));
          if (g > 0) ephase ^= 1;
        }
        tphase ^= 1;
      }
    }
  } else {
    // ------------------------------------------- epilogue warps ------------------------------------------------------------------
    const int q = warp & 3;                          // TMEM lane quarter this warp may read
    const int cg = (warp - FIRST_EPI_WARP) >> 2;     // which 16 of a chunk's 64 columns
    const bool storer = warp == FIRST_EPI_WARP + 1 && lane == 0;   // issues the planes' TMA stores and releases A chunks (warp 4: q = 0, cg = 0)
    const int free_stage = ng >= 2 ? ng - 2 : 0;     // the epilogue in which the storer releases the A chunks for the next tile
    uint32_t n = 0, ephase = 0;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
      for (int g = 0; g < ng; g++, n++) {
        const Stage& s = p.st[g];
        const uint32_t acc = n & 1, acc_phase = (n >> 1) & 1;
        const bool feeds_next = g + 1 < ng;
        // The two instantiations are the two flavours the encoders need: K-major weights <-> forward epilogues (bias, tanh, w2 row dot),
        // MN-major weights <-> backward epilogues (fp32 addend, column sums).  Gating at compile time keeps the register budget (96).
        const float* bias = B_MN ? nullptr : s.bias;
        const float* dotvec = B_MN ? nullptr : s.dotvec;
        const float* addsrc = B_MN ? s.addsrc : nullptr;
        float* colsum_part = B_MN ? s.colsum_part : nullptr;
        const int act = B_MN ? 0 : s.act;
        const int lr = lane >> 2, lc = (lane & 3) * 2;
        bool rok[4];
#pragma unroll
        for (int i = 0; i < 4; i++) rok[i] = tile * BM + q * 32 + 8 * i + lr < p.M;       // row 8i + lr: i = 2h + r
        // epilogue operands of chunk j+1 are fetched while chunk j is processed (chunk 0: before the accumulator is even complete):
        // a global-load latency inside every chunk would throttle the NEXT contraction, whose k-block j waits for chunk j
        float2 nb[2], nw[2], nad[4][2];
        auto prefetch = [&](int j) {
          const int c0 = j * BK + cg * 16 + lc;
          if (bias) { nb[0] = __ldg(reinterpret_cast<const float2*>(bias + c0)); nb[1] = __ldg(reinterpret_cast<const float2*>(bias + c0 + 8)); }
          if (dotvec) { nw[0] = __ldg(reinterpret_cast<const float2*>(dotvec + c0)); nw[1] = __ldg(reinterpret_cast<const float2*>(dotvec + c0 + 8)); }
          if (addsrc) {
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
              for (int c = 0; c < 2; c++)
                nad[i][c] = rok[i] ? __ldcs(reinterpret_cast<const float2*>(addsrc + (size_t)(tile * BM + q * 32 + 8 * i + lr) * ND + c0 + 8 * c))
                                   : make_float2(0.f, 0.f);
          }
        };
        prefetch(0);
        if (storer) TRACE(1, n * 12);
        mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
        tc_fence_after();
        if (storer) TRACE(1, n * 12 + 1);
        // every epilogue warp is past the previous epilogue (whose storer has waited for its TMA stores to finish READING the A buffer)
        if (feeds_next) asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
        // Fragment layout of tcgen05.ld.16x256b.x2 (probed, scratch/probe/tmem_frag.cu): lane L, register 8h + 4c + 2r + e holds
        // row 16h + 8r + (L>>2), column 8c + 2*(L&3) + e of the warp's 32 x 16 block — four lanes cover 32 contiguous bytes of a row,
        // so global / shared accesses are sector-exact without any transpose.
        float rowdot[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t v[16];
        const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + acc * ND + cg * 16;
        auto tmem_ld = [&](int j) {
#pragma unroll
          for (int h = 0; h < 2; h++)
            asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(v[8 * h]), "=r"(v[8 * h + 1]), "=r"(v[8 * h + 2]), "=r"(v[8 * h + 3]), "=r"(v[8 * h + 4]), "=r"(v[8 * h + 5]),
                           "=r"(v[8 * h + 6]), "=r"(v[8 * h + 7])
                         : "r"(taddr0 + ((uint32_t)(16 * h) << 16) + (uint32_t)(j * BK)));
        };
        // A stage that feeds the next contraction AND stores fp32 makes two passes over the accumulator: the first only produces the next
        // operand (the critical path: k-block j of the next contraction waits for chunk j), the second re-reads TMEM and stores fp32
        // while the next contraction's MMAs already run (fp32 stores drain at ~32 B/clk per SM: 1.9 us per chunk when done in line).
        const int npass = (feeds_next && s.out_f32) ? 2 : 1;
#pragma unroll 1
        for (int pass = 0; pass < npass; pass++) {
        const bool first = pass == 0, last = pass == npass - 1;
        if (!first) prefetch(0);
        tmem_ld(0);
#pragma unroll 1
        for (int j = 0; j < KB; j++) {
          const int c0 = j * BK + cg * 16 + lc;          // this lane's first column (its others: +1, +8, +9)
          float2 ad[4][2], bv[2], w2v[2];
          bv[0] = bv[1] = w2v[0] = w2v[1] = make_float2(0.f, 0.f);
          if (bias) { bv[0] = nb[0]; bv[1] = nb[1]; }
          if (dotvec) { w2v[0] = nw[0]; w2v[1] = nw[1]; }
          if (addsrc) {
#pragma unroll
            for (int i = 0; i < 4; i++) { ad[i][0] = nad[i][0]; ad[i][1] = nad[i][1]; }
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (storer) TRACE(1, n * 12 + 2 + 2 * j);
          float x[4][2][2];                              // [row i = 2h + r][column group c][e]
#pragma unroll
          for (int h = 0; h < 2; h++)
#pragma unroll
            for (int c = 0; c < 2; c++)
#pragma unroll
              for (int r = 0; r < 2; r++)
#pragma unroll
                for (int e = 0; e < 2; e++) x[2 * h + r][c][e] = __uint_as_float(v[8 * h + 4 * c + 2 * r + e]);
          if (j + 1 < KB) { tmem_ld(j + 1); prefetch(j + 1); }
#pragma unroll
          for (int i = 0; i < 4; i++)
#pragma unroll
            for (int c = 0; c < 2; c++) {
              x[i][c][0] += bv[c].x; x[i][c][1] += bv[c].y;
              if (addsrc) { x[i][c][0] += ad[i][c].x; x[i][c][1] += ad[i][c].y; }
              if (act == 1) { x[i][c][0] = tanhf(x[i][c][0]); x[i][c][1] = tanhf(x[i][c][1]); }
              if (!rok[i]) x[i][c][0] = x[i][c][1] = 0.f;      // rows past M: out of the sums, finite in the next GEMM
            }
          if (feeds_next && first) {
            // (hi, lo) planes into the A buffer, swizzled K-major UMMA layout: the 16-byte chunk ch of row r lives at r*128 + ((ch ^ (r & 7)) << 4);
            // a store instruction covers 8 rows x 16 bytes: 32 distinct banks
#pragma unroll
            for (int i = 0; i < 4; i++) {
              const int r = q * 32 + 8 * i + lr;
              uint8_t* rowp = a_smem + j * TILE_BYTES + r * 128 + (lane & 3) * 4;
#pragma unroll
              for (int c = 0; c < 2; c++) {
                uint32_t hp, lp;
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hp) : "f"(x[i][c][1]), "f"(x[i][c][0]));
                const float r0 = x[i][c][0] - __uint_as_float(hp << 16), r1 = x[i][c][1] - __uint_as_float(hp & 0xffff0000u);
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lp) : "f"(r1), "f"(r0));
                const int off = ((cg * 2 + c) ^ (r & 7)) << 4;
                *reinterpret_cast<uint32_t*>(rowp + off) = hp;
                *reinterpret_cast<uint32_t*>(rowp + A_PLANE_BYTES + off) = lp;
              }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to tcgen05.mma and to the TMA store
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&a_epi_bar[j]));
            if (storer) TRACE(1, n * 12 + 3 + 2 * j);
            if (storer) {
              if (s.out_planes) {
                mbar_wait(smem_u32(&a_epi_bar[j]), ephase);                 // all 16 warps have written chunk j
                tma_store_2d(&maps.o_hi[g], smem_u32(a_smem + j * TILE_BYTES), j * BK, tile * BM);
                tma_store_2d(&maps.o_lo[g], smem_u32(a_smem + A_PLANE_BYTES + j * TILE_BYTES), j * BK, tile * BM);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (j > 0) {
                  asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");     // chunk j-1 has been read by its store
                  if (g == free_stage) mbar_arrive(smem_u32(&a_free_bar[j - 1]));
                }
              } else if (g == free_stage) {
                mbar_arrive(smem_u32(&a_free_bar[j]));
              }
            }
          } else if (storer && g == free_stage && first) {
            mbar_arrive(smem_u32(&a_free_bar[j]));                           // single-contraction chain: nothing of this thread's is pending
          }
          if (s.out_f32 && last) {
#pragma unroll
            for (int i = 0; i < 4; i++)
              if (rok[i]) {
                float* o = s.out_f32 + (size_t)(tile * BM + q * 32 + 8 * i + lr) * ND + c0;
                *reinterpret_cast<float2*>(o) = make_float2(x[i][0][0], x[i][0][1]);
                *reinterpret_cast<float2*>(o + 8) = make_float2(x[i][1][0], x[i][1][1]);
              }
          }
          if (dotvec && first) {
            const float2 w0 = w2v[0], w1 = w2v[1];
#pragma unroll
            for (int i = 0; i < 4; i++)
              rowdot[i] = fmaf(x[i][0][0], w0.x, fmaf(x[i][0][1], w0.y, fmaf(x[i][1][0], w1.x, fmaf(x[i][1][1], w1.y, rowdot[i]))));
          }
          if (colsum_part && first) {     // warp-uniform: per-lane sums over its 4 rows, then over the 8 row groups (lane bits 2..4)
            float cs[4];
#pragma unroll
            for (int c = 0; c < 2; c++)
#pragma unroll
              for (int e = 0; e < 2; e++) {
                float t = (x[0][c][e] + x[1][c][e]) + (x[2][c][e] + x[3][c][e]);
                t += __shfl_xor_sync(0xffffffffu, t, 4);
                t += __shfl_xor_sync(0xffffffffu, t, 8);
                t += __shfl_xor_sync(0xffffffffu, t, 16);
                cs[2 * c + e] = t;
              }
            if (lane < 4) {
              float* o = colsum_part + ((size_t)tile * 4 + q) * ND + c0;
              *reinterpret_cast<float2*>(o) = make_float2(cs[0], cs[1]);
              *reinterpret_cast<float2*>(o + 8) = make_float2(cs[2], cs[3]);
            }
          }
        }
        if (first && storer && feeds_next && s.out_planes) {
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          if (g == free_stage) mbar_arrive(smem_u32(&a_free_bar[KB - 1]));
        }
        }   // pass
        if (dotvec) {            // the four lanes of a row group hold the partial dots of the same four rows
#pragma unroll
          for (int i = 0; i < 4; i++) {
            float t = rowdot[i];
            t += __shfl_xor_sync(0xffffffffu, t, 1);
            t += __shfl_xor_sync(0xffffffffu, t, 2);
            if ((lane & 3) == 0 && rok[i]) s.rowdot_part[(size_t)(tile * BM + q * 32 + 8 * i + lr) * 4 + cg] = t;
          }
        }
        if (feeds_next) ephase ^= 1;
        tc_fence_before();
        __syncwarp();
        if (storer) TRACE(1, n * 12 + 10);
        if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));
      }
    }
    if (storer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // global writes of the last stores complete before exit
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

}  // namespace chain
}  // namespace lk

using namespace lk;
using namespace lk::chain;

extern "C" {

// debug: copy the clock64 stamps of the last traced launch (LK_CHAIN_TRACE=1) to the host; returns the number of slots
static long long* g_trace_dev = nullptr;
int lk_tc_chain_trace(long long* host_out, int cap) {
  if (!g_trace_dev) return 0;
  const int n = cap < 1024 ? cap : 1024;
  cudaMemcpy(host_out, g_trace_dev, n * sizeof(long long), cudaMemcpyDeviceToHost);
  return n;
}

int lk_tc_chain(const void* A_hi, const void* A_lo, int64_t lda, int64_t M, const lk_chain_stage* stages, int n_stages, int b_mn,
                cudaStream_t st) {
  LK_REQUIRE(n_stages >= 1 && n_stages <= MAX_G, LK_ERR_ARG, "lk_tc_chain: 1..%d contractions", MAX_G);
  LK_REQUIRE(lda % 8 == 0 && lda >= ND, LK_ERR_SHAPE, "lk_tc_chain: A pitch %ld", (long)lda);
  LK_REQUIRE(((uintptr_t)A_hi | (uintptr_t)A_lo) % 16 == 0, LK_ERR_ARG, "lk_tc_chain: A planes must be 16-byte aligned");
  if (M == 0) return LK_OK;
  Maps maps;
  Params p = {};
  p.M = (int)M;
  p.m_tiles = (int)((M + BM - 1) / BM);
  p.ng = n_stages;
  {
    static long long* trace = [] {
      long long* t = nullptr;
      const char* e = getenv("LK_CHAIN_TRACE");
      if (e && e[0] == '1') { cudaMalloc(&t, 4 * 256 * sizeof(long long)); cudaMemset(t, 0, 4 * 256 * sizeof(long long)); }
      g_trace_dev = t;
      return t;
    }();
    p.trace = trace;
  }
  int rc;
  if ((rc = make_map(&maps.a_hi, A_hi, ND, M, lda, BM))) return rc;
  if ((rc = make_map(&maps.a_lo, A_lo, ND, M, lda, BM))) return rc;
  for (int g = 0; g < MAX_G; g++) {
    const lk_chain_stage& s = stages[g < n_stages ? g : 0];
    LK_REQUIRE(s.w_hi && s.w_lo && s.ldw % 8 == 0 && s.ldw >= ND, LK_ERR_ARG, "lk_tc_chain: stage %d has no weight planes / bad pitch", g);
    LK_REQUIRE(!(s.out_hi && g + 1 >= n_stages), LK_ERR_ARG, "lk_tc_chain: plane output of the last contraction is not supported (its tile never enters shared memory)");
    LK_REQUIRE(((uintptr_t)s.w_hi | (uintptr_t)s.w_lo | (uintptr_t)s.out_f32 | (uintptr_t)s.out_hi | (uintptr_t)s.out_lo | (uintptr_t)s.addsrc |
                (uintptr_t)s.bias | (uintptr_t)s.dotvec | (uintptr_t)s.colsum_part) % 16 == 0, LK_ERR_ARG, "lk_tc_chain: stage %d operands must be 16-byte aligned", g);
    LK_REQUIRE(!s.out_hi || (s.out_lo && s.ld_planes % 8 == 0 && s.ld_planes >= ND), LK_ERR_ARG, "lk_tc_chain: stage %d bad output planes", g);
    LK_REQUIRE(!s.dotvec || s.rowdot_part, LK_ERR_ARG, "lk_tc_chain: stage %d row dot without an output", g);
    LK_REQUIRE(b_mn ? (!s.bias && !s.dotvec && s.act == 0) : (!s.addsrc && !s.colsum_part), LK_ERR_ARG,
               "lk_tc_chain: stage %d: bias / tanh / row dot belong to the K-major (forward) flavour, addend / column sums to the MN-major (backward) one", g);
    // weights are square [256, 256]: stored [N, K] (b_mn = 0) or [K, N] (b_mn = 1); box 64 wide, 128 (K-major) or 64 (MN-major) rows
    if ((rc = make_map(&maps.b_hi[g], s.w_hi, ND, ND, s.ldw, b_mn ? 64 : HALF))) return rc;
    if ((rc = make_map(&maps.b_lo[g], s.w_lo, ND, ND, s.ldw, b_mn ? 64 : HALF))) return rc;
    if (g < n_stages) {
      Stage& d = p.st[g];
      d.bias = s.bias; d.addsrc = s.addsrc; d.act = s.act; d.out_f32 = s.out_f32;
      d.out_planes = s.out_hi ? 1 : 0;
      if (s.out_hi) {
        if ((rc = make_map(&maps.o_hi[g], s.out_hi, ND, M, s.ld_planes, BM))) return rc;
        if ((rc = make_map(&maps.o_lo[g], s.out_lo, ND, M, s.ld_planes, BM))) return rc;
      }
      d.colsum_part = s.colsum_part; d.dotvec = s.dotvec; d.rowdot_part = s.rowdot_part;
    }
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(chain_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaFuncSetAttribute(chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    attr_set = true;
  }
  const int grid = p.m_tiles < kNumSMs ? p.m_tiles : kNumSMs;
  if (b_mn) LK_LAUNCH((chain_kernel<true>), grid, NUM_THREADS, SMEM_BYTES, st, maps, p);
  else LK_LAUNCH((chain_kernel<false>), grid, NUM_THREADS, SMEM_BYTES, st, maps, p);
  return check_launch("tc_chain");
}

}  // extern "C"
