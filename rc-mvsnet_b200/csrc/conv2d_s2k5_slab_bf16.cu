// f1 FeatureNet's two down-sampling convolutions (reference models/modules.py:394-395 conv1.0 = Conv2d(8, 16, 5, stride 2,
// pad 2) and :399-400 conv2.0 = Conv2d(16, 32, 5, stride 2, pad 2), each conv -> BN -> ReLU, :28-62) on tcgen05 with
// TMA-fed split-bf16 operands ("bf16x3", see conv3d_slab_bf16.cu).  The 2-D, 5x5 sibling of conv3d_s2_slab_bf16.cu:
//
//   * input (2o + k - 2) per axis, k = 0..4: tap k reads parity k & 1 at index o + (k >> 1) - 1.  An input tile is fetched
//     as four parity sub-planes P[pr][pc][r][c] = x[2·(oy0 - 1 + r) + pr][2·(ox0 - 1 + c) + pc][:] by TMA boxes with
//     elementStrides {1, 2, 2, 1, 1}; out-of-range coordinates are zero-filled = the convolution's padding;
//   * in sub-plane coordinates every tap is an offset (dr, dc) in {-1, 0, +1}² = a start-address offset of the UMMA A
//     descriptor over a flat M = 128 run (4 output rows x 32 columns, columns 0 and 31 = halo); all 25 taps accumulate
//     into one TMEM accumulator, the epilogue adds the BN shift, applies ReLU and stores fp32 and / or split-bf16;
//   * 16 input channels: one K = 16 step per tap (25 steps).  8 input channels: one step covers TWO taps, K-chunk 1 being
//     addressed through the descriptor's leading-dimension byte offset — (kh, kw) + (kh + 2, kw): next row of the same
//     sub-plane (LBO = 512 B) for kh = 0, 1; (4, kw) + (4, kw + 2): next position (LBO = 16 B) for kw = 0, 1; the last tap
//     (4, 4) pairs with a block of zeros (13 steps).
//
// Every (image, row-tile, column-tile) item is independent: the kernel is persistent (one CTA per SM walks a contiguous
// range of items, weight image staged once, TMEM accumulator groups double-buffer the MMAs against the epilogue).
// Warp roles: warps 0-3 epilogue, warp 4 TMA producer, warp 5 MMA issuer.
#include "slab_common.cuh"

namespace rcmvs {
namespace slab2d5 {

using namespace rcmvs::slab;

constexpr int TWO = 30;                                            // real output columns per tile (columns 0, 31 = halo)

template <int CIN, int COUT, int KM, int NS>
struct Cfg {
  static constexpr int NCH = CIN / 8;                              // 16-byte channel chunks per position
  static constexpr int NSTEP = CIN == 8 ? 13 : 25;                 // K = 16 steps
  static constexpr int ROWS_S = 4 * KM + 2;                        // sub-plane rows (1 halo row above, 1 below)
  static constexpr int OROWS = 4 * KM;
  static constexpr int PLANE_B = ROWS_S * PW * 16;                 // one chunk plane of one parity sub-plane
  static constexpr int SUB_B = NCH * PLANE_B;
  static constexpr int PART_B = 4 * SUB_B;
  static constexpr int STAGE_B = 2 * PART_B;                       // hi, lo
  static constexpr int FRONT = 128;                                // the (-1, -1) taps of the first sub-plane start 16 B early
  static constexpr int ZERO_OFF = FRONT + NS * STAGE_B + 128;      // (CIN = 8) 2 KB of zeros: K-chunk 1 of the unpaired tap
  static constexpr int A_BYTES = FRONT + NS * STAGE_B + 128 + 2048 + 128;
  static constexpr int BBLK_B = 2 * COUT * 16;                     // one step: 2 chunks x COUT rows x 16 B
  static constexpr int BTERM_B = NSTEP * BBLK_B;
  static constexpr int B_BYTES = 2 * BTERM_B;                      // hi, lo
  static constexpr int NG = 512 / (KM * COUT) >= 4 ? 4 : 2;        // accumulator groups in flight
  static constexpr int TCOLS = NG * KM * COUT <= 128 ? 128 : (NG * KM * COUT <= 256 ? 256 : 512);
  static constexpr size_t SMEM = (size_t)A_BYTES + B_BYTES + 512;
  static_assert((CIN == 8 || CIN == 16) && COUT % 16 == 0 && PLANE_B % 128 == 0 && NG * KM * COUT <= 512, "configuration");
};

struct Params {
  const __nv_bfloat16* wimg;
  const float* bias;           // [Cout] or null (folded BatchNorm shift)
  float* y;                    // [N][Ho][Wo][Cout] fp32 or null
  __nv_bfloat16* ys;           // [2][N][Ho][Wo][Cout] split-bf16 or null
  int N, H, W, Ho, Wo;
  int tiles_x, tiles_y, items, items_per_cta;
  int relu;
};

// step -> (tap of K-chunk 0, LBO kind of K-chunk 1)   [CIN = 8: 13 steps; CIN = 16: step = tap, chunk 1 = channels 8..15]
//   steps 0..9 : (kh = s / 5, kw = s % 5) + (kh + 2, kw)      LBO = one row (512 B)
//   steps 10,11: (4, kw = s - 10) + (4, kw + 2)               LBO = one position (16 B)
//   step  12   : (4, 4) + zero block
__device__ __forceinline__ constexpr int step_kh(int s, int cin) { return cin == 16 ? s / 5 : (s < 10 ? s / 5 : 4); }
__device__ __forceinline__ constexpr int step_kw(int s, int cin) { return cin == 16 ? s % 5 : (s < 10 ? s % 5 : (s < 12 ? s - 10 : 4)); }

template <int CIN, int COUT, int KM, int NS>
__global__ void __launch_bounds__(kThreads, 1)
conv2d_s2k5_slab_bf16_kernel(const __grid_constant__ CUtensorMap tmap, const Params p) {
  using C = Cfg<CIN, COUT, KM, NS>;
  constexpr int NG = C::NG;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* a_smem = smem + C::FRONT;                          // stage ring (the 128 bytes in front are zeroed)
  unsigned char* b_smem = smem + C::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::A_BYTES + C::B_BYTES);
  uint64_t* full = bars;                 // [NS]  TMA -> MMA
  uint64_t* empty = bars + NS;           // [NS]  MMA -> TMA
  uint64_t* grp_full = bars + 2 * NS;    // [NG]  MMA -> epilogue
  uint64_t* grp_free = grp_full + NG;    // [NG]  epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(grp_free + NG);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item0 = blockIdx.x * p.items_per_cta;
  const int nitems = max(0, min(p.items, item0 + p.items_per_cta) - item0);
  auto item_coords = [&](int it, int& n, int& oy0, int& ox0) {
    const int tx = it % p.tiles_x, ty = (it / p.tiles_x) % p.tiles_y;
    n = it / (p.tiles_x * p.tiles_y);
    oy0 = ty * C::OROWS;
    ox0 = tx * TWO;
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    for (int i = 0; i < NG; ++i) { mbar_init(grp_full + i, 1); mbar_init(grp_free + i, 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4 && lane == 0) tma_prefetch_desc(&tmap);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)C::TCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {
    const float4* src = reinterpret_cast<const float4*>(p.wimg);
    float4* dst = reinterpret_cast<float4*>(b_smem);
    for (int i = threadIdx.x; i < C::B_BYTES / 16; i += kThreads) dst[i] = __ldg(src + i);
    for (int i = threadIdx.x; i < (C::A_BYTES - C::FRONT - NS * C::STAGE_B) / 16; i += kThreads)      // overrun pad + zero block
      reinterpret_cast<float4*>(a_smem + NS * C::STAGE_B)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x < C::FRONT / 16) reinterpret_cast<float4*>(smem)[threadIdx.x] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // =============================== TMA producer: one stage per item ==============================================
    if (lane == 0) {
      for (int i = 0; i < nitems; ++i) {
        const int buf = i % NS;
        int n, oy0, ox0;
        item_coords(item0 + i, n, oy0, ox0);
        mbar_wait(empty + buf, ((i / NS) & 1) ^ 1);
        mbar_arrive_expect_tx(full + buf, (uint32_t)C::STAGE_B);
        const uint32_t dst0 = smem_u32(a_smem + (size_t)buf * C::STAGE_B);
#pragma unroll
        for (int part = 0; part < 2; ++part)
#pragma unroll
          for (int sp = 0; sp < 4; ++sp)
#pragma unroll
            for (int ch = 0; ch < C::NCH; ++ch)          // map dims {C, W, H, N, 2}, traversal stride 2 along W and H
              tma_load_5d(dst0 + part * C::PART_B + sp * C::SUB_B + ch * C::PLANE_B, &tmap, full + buf, ch * 8,
                          2 * (ox0 - 1) + (sp & 1), 2 * (oy0 - 1) + (sp >> 1), n, part);
      }
    }
  } else if (warp == 5) {
    // =============================== MMA issuer ====================================================================
    const bool leader = elect_one();
    const uint32_t b_lo0 = desc_lo(smem_u32(b_smem), COUT * 16);
    constexpr uint32_t idesc = make_idesc_bf16(128, COUT);
    const uint32_t zb = smem_u32(smem) + C::ZERO_OFF;
    for (int i = 0; i < nitems; ++i) {
      const int g = i % NG, buf = i % NS;
      if (i >= NG) mbar_wait(grp_free + g, ((i / NG) - 1) & 1);
      mbar_wait(full + buf, (i / NS) & 1);
      tc_fence_after();
      const uint32_t a_base = smem_u32(a_smem + (size_t)buf * C::STAGE_B);
#pragma unroll
      for (int st = 0; st < C::NSTEP; ++st) {
        const int kh = step_kh(st, CIN), kw = step_kw(st, CIN);
        const int sp = (kh & 1) * 2 + (kw & 1);
        const int off = ((kh >> 1) - 1) * PW + ((kw >> 1) - 1);              // positions relative to the output's own
#pragma unroll
        for (int mt = 0; mt < KM; ++mt) {
          constexpr uint32_t kBLoOff = C::BTERM_B >> 4;
          const uint32_t a_addr = a_base + sp * C::SUB_B + (uint32_t)(((1 + 4 * mt) * PW + off) * 16);
          uint32_t lbo_hi, lbo_lo;
          if constexpr (CIN == 16) {
            lbo_hi = lbo_lo = (uint32_t)C::PLANE_B;                          // K-chunk 1 = channels 8..15 (next chunk plane)
          } else {
            lbo_hi = st < 10 ? 512u : (st < 12 ? 16u : zb - a_addr);
            lbo_lo = st < 12 ? lbo_hi : zb - (a_addr + C::PART_B);
          }
          const uint32_t a_hi_d = desc_lo(a_addr, lbo_hi), a_lo_d = desc_lo(a_addr + C::PART_B, lbo_lo);
          const uint32_t b_d = b_lo0 + (uint32_t)((st * C::BBLK_B) >> 4);
          const uint32_t d = tmem_base + (uint32_t)(mt * (NG * COUT) + g * COUT);
          if (leader) {
            umma_bf16_lohi_acc(d, a_lo_d, b_d, idesc, st == 0 ? 0u : 1u);         // x_lo · w_hi (first MMA of the item overwrites)
            umma_bf16_lohi_acc(d, a_hi_d, b_d + kBLoOff, idesc, 1u);              // x_hi · w_lo
            umma_bf16_lohi_acc(d, a_hi_d, b_d, idesc, 1u);                        // x_hi · w_hi
          }
        }
      }
      if (leader) {
        umma_commit(empty + buf);
        umma_commit(grp_full + g);
      }
      __syncwarp();
    }
  } else {
    // =============================== epilogue (128 threads) ========================================================
    // thread (warp w, lane l) of M-tile mt holds output row oy0 + 4·mt + w, column ox0 + l - 1 (l = 0, 31: halo)
    const int64_t part_stride = (int64_t)p.N * p.Ho * p.Wo * COUT;
    for (int q = 0; q < nitems; ++q) {
      const int g = q % NG;
      int n, oy0, ox0;
      item_coords(item0 + q, n, oy0, ox0);
      const int gx = ox0 + lane - 1;
      const bool col_ok = lane >= 1 && lane <= TWO && gx < p.Wo;
      mbar_wait(grp_full + g, (q / NG) & 1);
      tc_fence_after();
#pragma unroll
      for (int mt = 0; mt < KM; ++mt) {
        uint32_t vr[COUT];
        const uint32_t ta = tmem_base + ((uint32_t)(warp * 32) << 16) + mt * (NG * COUT) + g * COUT;
#pragma unroll
        for (int c0 = 0; c0 < COUT; c0 += 16) tmem_ld16_nowait(ta + c0, &vr[c0]);
        tmem_ld_wait();
        if (mt == KM - 1) {
          tc_fence_before();
          mbar_arrive(grp_free + g);
        }
        const int gy = oy0 + 4 * mt + warp;
        if (!col_ok || gy >= p.Ho) continue;
        const int64_t o = (((int64_t)n * p.Ho + gy) * p.Wo + gx) * COUT;
#pragma unroll
        for (int c = 0; c < COUT; c += 4) {
          float4 r = make_float4(__uint_as_float(vr[c]), __uint_as_float(vr[c + 1]), __uint_as_float(vr[c + 2]),
                                 __uint_as_float(vr[c + 3]));
          if (p.bias) {
            const float4 t4 = ldg4(p.bias + c);
            r.x += t4.x; r.y += t4.y; r.z += t4.z; r.w += t4.w;
          }
          if (p.relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
          if (p.y) st4(p.y + o + c, r);
          if (p.ys) store_split4(p.ys, part_stride, o + c, r);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TCOLS)
                 : "memory");
  }
}

// Weight image: torch Conv2d weight [Cout][Cin][5][5] x BN scale -> img[term][step][chunk][row = cout][e]
//   Cin = 16: step = kh*5 + kw, channel = chunk*8 + e
//   Cin = 8 : chunk 0 = tap (step_kh, step_kw); chunk 1 = (kh + 2, kw) for steps 0..9, (4, kw + 2) for steps 10, 11, zero for 12
__global__ void pack_s2k5_weight_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                        __nv_bfloat16* __restrict__ img, int Cout, int Cin) {
  const int nstep = Cin == 8 ? 13 : 25;
  const int per_term = nstep * 2 * Cout * 8;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per_term) return;
  const int e = i & 7;
  int r = i >> 3;
  const int n = r % Cout; r /= Cout;
  const int chunk = r & 1;
  const int st = r >> 1;
  int kh = step_kh(st, Cin), kw = step_kw(st, Cin), ci = e;
  bool used = true;
  if (Cin == 16) {
    ci = chunk * 8 + e;
  } else if (chunk == 1) {
    if (st < 10) kh += 2;
    else if (st < 12) kw += 2;
    else used = false;
  }
  float v = 0.f;
  if (used) {
    v = w[((int64_t)n * Cin + ci) * 25 + kh * 5 + kw];
    if (scale) v *= scale[n];
  }
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  img[i] = hi;
  img[per_term + i] = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// 5-D map over split-bf16 activations, dims {C, W, H, N, 2}, traversal stride 2 along W and H: box {8, 2·32, 2·rows, 1, 1}
static int make_tmap_s2(CUtensorMap* map, const void* x, int N, int H, int W, int C, int rows) {
  EncodeTiledFn fn = encode_fn();
  RCMVS_REQUIRE(fn != nullptr, RCMVS_ERR_CUDA, "rcmvs_conv2d_s2k5_slab_fwd: cuTensorMapEncodeTiled not available");
  const cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N, 2};
  const cuuint64_t gstr[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)N * H * W * C * 2};
  const cuuint32_t box[5] = {8, (cuuint32_t)(2 * PW), (cuuint32_t)(2 * rows), 1, 1};
  const cuuint32_t estr[5] = {1, 2, 2, 1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RCMVS_REQUIRE(r == CUDA_SUCCESS, RCMVS_ERR_CUDA, "rcmvs_conv2d_s2k5_slab_fwd: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return RCMVS_OK;
}

template <int CIN, int COUT, int KM, int NS>
static int launch(const void* x, Params p, cudaStream_t st) {
  using C = Cfg<CIN, COUT, KM, NS>;
  static_assert(C::SMEM <= 232448, "shared memory budget exceeded");
  auto kern = conv2d_s2k5_slab_bf16_kernel<CIN, COUT, KM, NS>;
  static SmemAttrOnce smem_attr;
  {
    cudaError_t e = smem_attr.ensure(kern, (int)C::SMEM);
    RCMVS_REQUIRE(e == cudaSuccess, RCMVS_ERR_CUDA, "rcmvs_conv2d_s2k5_slab_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  CUtensorMap map;
  const int rc = make_tmap_s2(&map, x, p.N, p.H, p.W, CIN, C::ROWS_S);
  if (rc != RCMVS_OK) return rc;
  p.tiles_x = ceil_div(p.Wo, TWO);
  p.tiles_y = ceil_div(p.Ho, C::OROWS);
  p.items = p.tiles_x * p.tiles_y * p.N;
  const int ctas = p.items < 148 ? p.items : 148;                 // persistent: one CTA per SM
  p.items_per_cta = ceil_div(p.items, ctas);
  kern<<<ceil_div(p.items, p.items_per_cta), kThreads, C::SMEM, st>>>(map, p);
  RCMVS_CHECK_LAUNCH("rcmvs_conv2d_s2k5_slab_fwd");
  return RCMVS_OK;
}

static bool supported(int Cin, int Cout) { return (Cin == 8 && Cout == 16) || (Cin == 16 && Cout == 32); }

}  // namespace slab2d5
}  // namespace rcmvs

extern "C" int64_t rcmvs_conv2d_s2k5_slab_weight_elems(int Cin, int Cout) {
  return (int64_t)2 * (Cin == 8 ? 13 : 25) * 2 * Cout * 8;
}

extern "C" int rcmvs_conv2d_s2k5_slab_pack(const float* w, const float* scale, void* img, int Cout, int Cin,
                                           rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(w && img, RCMVS_ERR_BAD_ARG, "rcmvs_conv2d_s2k5_slab_pack: null pointer");
  RCMVS_REQUIRE(slab2d5::supported(Cin, Cout), RCMVS_ERR_UNSUPPORTED, "rcmvs_conv2d_s2k5_slab_pack: unsupported Cin=%d Cout=%d",
                Cin, Cout);
  const int per_term = (int)(rcmvs_conv2d_s2k5_slab_weight_elems(Cin, Cout) / 2);
  slab2d5::pack_s2k5_weight_kernel<<<ceil_div(per_term, 256), 256, 0, as_stream(stream)>>>(
      w, scale, reinterpret_cast<__nv_bfloat16*>(img), Cout, Cin);
  RCMVS_CHECK_LAUNCH("rcmvs_conv2d_s2k5_slab_pack");
  return RCMVS_OK;
}

extern "C" int rcmvs_conv2d_s2k5_slab_fwd(const void* x_split, const void* wimg, const float* bias, float* y, void* y_split,
                                          int N, int H, int W, int Cin, int Cout, int relu, rcmvs_stream_t stream) {
  using namespace rcmvs;
  using namespace rcmvs::slab2d5;
  RCMVS_REQUIRE(x_split && wimg && (y || y_split), RCMVS_ERR_BAD_ARG, "rcmvs_conv2d_s2k5_slab_fwd: null pointer");
  RCMVS_REQUIRE(slab2d5::supported(Cin, Cout), RCMVS_ERR_UNSUPPORTED, "rcmvs_conv2d_s2k5_slab_fwd: unsupported Cin=%d Cout=%d",
                Cin, Cout);
  RCMVS_REQUIRE(N > 0 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0, RCMVS_ERR_UNSUPPORTED,
                "rcmvs_conv2d_s2k5_slab_fwd: H, W must be even (got %d, %d)", H, W);
  RCMVS_REQUIRE((reinterpret_cast<uintptr_t>(x_split) & 15) == 0, RCMVS_ERR_BAD_ARG,
                "rcmvs_conv2d_s2k5_slab_fwd: x_split must be 16-byte aligned (TMA)");
  Params p;
  p.wimg = reinterpret_cast<const __nv_bfloat16*>(wimg);
  p.bias = bias; p.y = y; p.ys = reinterpret_cast<__nv_bfloat16*>(y_split);
  p.N = N; p.H = H; p.W = W; p.Ho = H / 2; p.Wo = W / 2; p.relu = relu;
  p.tiles_x = p.tiles_y = p.items = p.items_per_cta = 0;
  cudaStream_t st = as_stream(stream);
  if (Cin == 8) return launch<8, 16, 3, 3>(x_split, p, st);
  return launch<16, 32, 2, 2>(x_split, p, st);
}
