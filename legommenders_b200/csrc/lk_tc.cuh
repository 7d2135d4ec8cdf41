// tcgen05 / TMEM / TMA building blocks shared by the GEMM kernels of this library (lk_gemm_tc.cu, lk_chain.cu): mbarrier and TMA
// wrappers, UMMA shared-memory / instruction descriptors (128-byte swizzle), the fp32 -> (hi, lo) bf16 split, tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "lk_common.cuh"

namespace lk {
namespace tc {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int TILE_BYTES = BM * BK * 2;        // one plane of a 128-row operand tile: 16 KiB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 128-byte-swizzle UMMA shared-memory descriptor (cute/arch/mma_sm100_desc.hpp: SmemDescriptor).
//   K-major : rows of 64 bf16 (128 B); 8-row groups 1024 B apart (SBO); LBO unused.
//   MN-major: k-rows of 64 MN-elements (128 B); 8-k groups 1024 B apart (SBO); next 64 MN-elements BK*128 B away (LBO).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, bool mn_major) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(mn_major ? ((BK * 128) >> 4) : 0) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
// instruction descriptor (cute/arch/mma_sm100_desc.hpp: InstrDescriptor): fp32 accumulate, bf16 A/B
__host__ __device__ constexpr uint32_t make_idesc(bool a_mn, bool b_mn, int bn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// (hi, lo) bf16 planes of four floats: two packed cvt.rn.bf16x2 per plane; hi's fp32 image is its bits shifted up
__device__ __forceinline__ void split4(const float4& v, uint2& hi, uint2& lo) {
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

// 2-D bf16 tensor map: `inner` contiguous elements per row, `outer` rows, row pitch `pitch_elems`; 128B-swizzled box.
static inline int make_map(CUtensorMap* m, const void* base, int64_t inner, int64_t outer, int64_t pitch_elems, int box_outer) {
  EncodeTiledFn fn = encode_fn();
  LK_REQUIRE(fn != nullptr, LK_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)pitch_elems * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LK_REQUIRE(r == CUDA_SUCCESS, LK_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): inner=%ld outer=%ld pitch=%ld", (int)r, (long)inner,
             (long)outer, (long)pitch_elems);
  return LK_OK;
}

}  // namespace tc
}  // namespace lk
