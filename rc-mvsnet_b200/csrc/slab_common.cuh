// TMA + kind::f16 (bf16) helpers shared by the split-bf16 tensor-core convolution kernels (conv3d_slab_bf16.cu: 3-D slab
// kernel; conv2d_slab_bf16.cu: its 2-D, persistent-tile variant for FeatureNet).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "umma.cuh"

namespace rcmvs {
namespace slab {

using namespace rcmvs::umma;

constexpr int kThreads = 32 * 6;

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], "
      "[%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// Instruction descriptor for kind::f16 with bf16 operands (cute::UMMA::InstrDescriptor): c=F32 [4,6)=1, a=BF16 [7,10)=1,
// b=BF16 [10,13)=1, K-major A and B (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(1u), "r"(kDescHi)
      : "memory");
}

// same with an explicit accumulate flag (0: the MMA overwrites the accumulator — first MMA of a tile, no zeroing pass)
__device__ __forceinline__ void umma_bf16_lohi_acc(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                                   uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi)
      : "memory");
}

// ---- host side ----------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 5-D map over split-bf16 activations [2·B][D][H][W][C] with box {8, 32, rows, 1, 1}
static inline int make_tmap(CUtensorMap* map, const void* x, int B, int D, int H, int W, int C, int rows) {
  EncodeTiledFn fn = encode_fn();
  RCMVS_REQUIRE(fn != nullptr, RCMVS_ERR_CUDA, "rcmvs_conv3d_slab_fwd: cuTensorMapEncodeTiled not available");
  const cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)2 * B};
  const cuuint64_t gstr[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2,
                              (cuuint64_t)D * H * W * C * 2};
  const cuuint32_t box[5] = {8, (cuuint32_t)PW, (cuuint32_t)rows, 1, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RCMVS_REQUIRE(r == CUDA_SUCCESS, RCMVS_ERR_CUDA, "rcmvs_conv3d_slab_fwd: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return RCMVS_OK;
}


// pre-split store of 4 consecutive channels: hi = bf16(v), lo = bf16(v - hi) into the two planes of a split tensor
__device__ __forceinline__ void store_split4(__nv_bfloat16* hi_plane, int64_t part_stride, int64_t o, float4 r) {
  const __nv_bfloat16 h0 = __float2bfloat16_rn(r.x), h1 = __float2bfloat16_rn(r.y);
  const __nv_bfloat16 h2 = __float2bfloat16_rn(r.z), h3 = __float2bfloat16_rn(r.w);
  const __nv_bfloat162 hi01 = __halves2bfloat162(h0, h1), hi23 = __halves2bfloat162(h2, h3);
  const __nv_bfloat162 lo01 = __floats2bfloat162_rn(r.x - __bfloat162float(h0), r.y - __bfloat162float(h1));
  const __nv_bfloat162 lo23 = __floats2bfloat162_rn(r.z - __bfloat162float(h2), r.w - __bfloat162float(h3));
  uint2 hv, lv;
  hv.x = *reinterpret_cast<const uint32_t*>(&hi01); hv.y = *reinterpret_cast<const uint32_t*>(&hi23);
  lv.x = *reinterpret_cast<const uint32_t*>(&lo01); lv.y = *reinterpret_cast<const uint32_t*>(&lo23);
  *reinterpret_cast<uint2*>(hi_plane + o) = hv;
  *reinterpret_cast<uint2*>(hi_plane + part_stride + o) = lv;
}

}  // namespace slab
}  // namespace rcmvs
