// f1 FeatureNet's 3x3 stride-1 convolutions on the tensor cores (reference models/modules.py:28-62 Conv2d = conv -> BN ->
// ReLU; :363-464 FeatureNet: conv1.1/1.2 16->16, conv2.1/2.2 32->32, out2 32->16, out3 32->8 = 17 of the FPN's 27 GFLOP
// for 5 views at 512x640).  The 2-D sibling of conv3d_slab_bf16.cu, same ingredients:
//   * operands = split-bf16 planes [2][N][H][W][C] written by the producing layer, staged by TMA box loads whose
//     out-of-bounds zero fill is the convolution's padding, landing directly in the UMMA canonical K-major layout;
//   * tcgen05.mma kind::f16, three products per operand pair (bf16x3), fp32 accumulators in TMEM;
//   * M = 128 input positions = a flat run of the padded (ROWS x 32) tile, kw = +16 bytes of descriptor start address,
//     N folds (kh, cout), the sum over kh is done by the epilogue through shared memory.
// There is no depth dimension, hence no TMEM plane ring: every (image, row-tile, column-tile) item is independent, so the
// kernel is PERSISTENT — one CTA per SM walks a contiguous range of items, the weight image is staged once per CTA and
// the TMEM accumulator groups double-buffer MMA against the epilogue.
// Warp roles: warps 0-3 epilogue, warp 4 TMA producer, warp 5 MMA issuer.
#include "slab_common.cuh"

namespace rcmvs {
namespace slab2d {

using namespace rcmvs::slab;

template <int CIN, int COUT, int KM, int NS>
struct Cfg {
  static constexpr int CP = COUT < 8 ? 8 : COUT;
  static constexpr int G = ((3 * CP + 15) / 16) * 16;             // TMEM columns / B rows: 3 kh blocks of CP channels
  static constexpr int KS = CIN / 16;
  // CIN = 8 (conv0.0 / conv0.1): one K = 16 step covers TWO kw taps — K-chunk 1 is the next position (LBO = 16 B); the
  // second step is (kw = 2, zero block)
  static constexpr int NSTEP = CIN == 8 ? 2 : 3 * KS;
  static constexpr int NCH = CIN / 8;
  static constexpr int ROWS = 4 * KM;
  static constexpr int OROWS = ROWS - 2;
  static constexpr int PLANE_B = ROWS * PW * 16;
  static constexpr int SLAB_B = NCH * PLANE_B;
  static constexpr int STAGE_B = 2 * SLAB_B;
  static constexpr int ZERO_OFF = NS * STAGE_B + 128;             // (CIN = 8) 2 KB of zeros
  static constexpr int A_BYTES = NS * STAGE_B + 128 + (CIN == 8 ? 2048 + 128 : 0);
  static constexpr int BBLK_B = 2 * G * 16;                       // one step: 2 chunks x G rows x 16 B
  static constexpr int BTERM_B = NSTEP * BBLK_B;
  static constexpr int B_BYTES = 2 * BTERM_B;
  static constexpr int NG = 512 / (KM * G) >= 4 ? 4 : 512 / (KM * G);   // accumulator groups in flight (2..4)
  static constexpr int EX_B = 3 * OROWS * PW * CP * 4;
  static constexpr size_t SMEM = (size_t)A_BYTES + B_BYTES + EX_B + 512;
  static_assert((CIN == 8 || CIN % 16 == 0) && NG >= 2 && PLANE_B % 128 == 0, "configuration");
};

struct Params {
  const __nv_bfloat16* wimg;
  const float* bias;           // [Cout] or null (folded BatchNorm shift)
  float* y;                    // [N][H][W][Cout] fp32 or null
  __nv_bfloat16* ys;           // [2][N][H][W][Cout] split-bf16 or null
  int N, H, W;
  int tiles_x, tiles_y, items, items_per_cta;
  int relu;
};

template <int CIN, int COUT, int KM, int NS>
__global__ void __launch_bounds__(kThreads, 1)
conv2d_slab_bf16_kernel(const __grid_constant__ CUtensorMap tmap, const Params p) {
  using C = Cfg<CIN, COUT, KM, NS>;
  constexpr int G = C::G, NG = C::NG, CP = C::CP;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* a_smem = smem;
  unsigned char* b_smem = smem + C::A_BYTES;
  float* ex = reinterpret_cast<float*>(smem + C::A_BYTES + C::B_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::A_BYTES + C::B_BYTES + C::EX_B);
  uint64_t* full = bars;                 // [NS]  TMA -> MMA
  uint64_t* empty = bars + NS;           // [NS]  MMA -> TMA
  uint64_t* grp_full = bars + 2 * NS;    // [NG]  MMA -> epilogue
  uint64_t* grp_free = grp_full + NG;    // [NG]  epilogue -> MMA (drained and zeroed)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(grp_free + NG);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item0 = blockIdx.x * p.items_per_cta;
  const int nitems = max(0, min(p.items, item0 + p.items_per_cta) - item0);
  auto item_coords = [&](int it, int& n, int& y0, int& x0) {
    const int tx = it % p.tiles_x, ty = (it / p.tiles_x) % p.tiles_y;
    n = it / (p.tiles_x * p.tiles_y);
    y0 = ty * C::OROWS;
    x0 = tx * TW;
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    for (int i = 0; i < NG; ++i) { mbar_init(grp_full + i, 1); mbar_init(grp_free + i, 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4 && lane == 0) tma_prefetch_desc(&tmap);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {
    const float4* src = reinterpret_cast<const float4*>(p.wimg);
    float4* dst = reinterpret_cast<float4*>(b_smem);
    for (int i = threadIdx.x; i < C::B_BYTES / 16; i += kThreads) dst[i] = __ldg(src + i);
    for (int i = threadIdx.x; i < (C::A_BYTES - NS * C::STAGE_B) / 16; i += kThreads)       // overrun pad (+ zero block)
      reinterpret_cast<float4*>(a_smem + NS * C::STAGE_B)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp < 4) {
    for (int c = 0; c < NG * KM * G; c += 16) tmem_st16_zero(tmem_base + ((uint32_t)(warp * 32) << 16) + c);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 4) {
    // =============================== TMA producer: one stage per item ==============================================
    if (lane == 0) {
      for (int i = 0; i < nitems; ++i) {
        const int buf = i % NS;
        int n, y0, x0;
        item_coords(item0 + i, n, y0, x0);
        mbar_wait(empty + buf, ((i / NS) & 1) ^ 1);
        mbar_arrive_expect_tx(full + buf, (uint32_t)C::STAGE_B);
        const uint32_t dst0 = smem_u32(a_smem + (size_t)buf * C::STAGE_B);
#pragma unroll
        for (int part = 0; part < 2; ++part)
#pragma unroll
          for (int ch = 0; ch < C::NCH; ++ch)          // map dims {C, W, H, N, 2}: images = the map's 4th dimension
            tma_load_5d(dst0 + part * C::SLAB_B + ch * C::PLANE_B, &tmap, full + buf, ch * 8, x0 - 1, y0 - 1, n, part);
      }
    }
  } else if (warp == 5) {
    // =============================== MMA issuer ====================================================================
    const bool leader = elect_one();
    const uint32_t b_lo0 = desc_lo(smem_u32(b_smem), G * 16);
    constexpr uint32_t idesc = make_idesc_bf16(128, G);
    for (int i = 0; i < nitems; ++i) {
      const int g = i % NG, buf = i % NS;
      if (i >= NG) mbar_wait(grp_free + g, ((i / NG) - 1) & 1);
      mbar_wait(full + buf, (i / NS) & 1);
      tc_fence_after();
      const uint32_t a_base = smem_u32(a_smem + (size_t)buf * C::STAGE_B);
#pragma unroll
      for (int st = 0; st < C::NSTEP; ++st) {
        constexpr int KSD = C::KS > 0 ? C::KS : 1;
        const int kw = CIN == 8 ? 2 * st : st / KSD, ks = CIN == 8 ? 0 : st % KSD;
#pragma unroll
        for (int mt = 0; mt < KM; ++mt) {
          constexpr uint32_t kLoOff = C::SLAB_B >> 4, kBLoOff = C::BTERM_B >> 4;
          uint32_t a_d, a_lo_d;
          if constexpr (CIN == 8) {
            const uint32_t a_addr = a_base + (uint32_t)((mt * 128 + kw) * 16);
            if (st == 0) {
              a_d = desc_lo(a_addr, 16);
              a_lo_d = a_d + kLoOff;
            } else {
              const uint32_t zb = smem_u32(a_smem) + C::ZERO_OFF;
              a_d = desc_lo(a_addr, zb - a_addr);
              a_lo_d = desc_lo(a_addr + C::SLAB_B, zb - (a_addr + C::SLAB_B));
            }
          } else {
            a_d = desc_lo(a_base, C::PLANE_B) + (uint32_t)((ks * 2 * C::PLANE_B) >> 4) + (uint32_t)(mt * 128 + kw);
            a_lo_d = a_d + kLoOff;
          }
          const uint32_t b_d = b_lo0 + (uint32_t)((st * C::BBLK_B) >> 4);
          const uint32_t d = tmem_base + (uint32_t)(mt * (NG * G) + g * G);
          if (leader) {
            umma_bf16_lohi(d, a_lo_d, b_d, idesc);                  // x_lo · w_hi
            umma_bf16_lohi(d, a_d, b_d + kBLoOff, idesc);           // x_hi · w_lo
            umma_bf16_lohi(d, a_d, b_d, idesc);                     // x_hi · w_hi
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
    const int64_t part_stride = (int64_t)p.N * p.H * p.W * COUT;
    for (int q = 0; q < nitems; ++q) {
      const int g = q % NG;
      int n, y0, x0;
      item_coords(item0 + q, n, y0, x0);
      const int gx = x0 + lane;
      const bool col_ok = lane < TW && gx < p.W;
      mbar_wait(grp_full + g, (q / NG) & 1);
      tc_fence_after();
      // one M-tile at a time (G up to 96 accumulator columns per thread): TMEM -> registers -> its three kh exchange planes
#pragma unroll
      for (int mt = 0; mt < KM; ++mt) {
        const uint32_t ta = tmem_base + ((uint32_t)(warp * 32) << 16) + mt * (NG * G) + g * G;
        uint32_t vr[G];
#pragma unroll
        for (int c0 = 0; c0 < G; c0 += 16) tmem_ld16_nowait(ta + c0, &vr[c0]);
        tmem_ld_wait();
#pragma unroll
        for (int c0 = 0; c0 < G; c0 += 16) tmem_st16_zero(ta + c0);
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const int j = 4 * mt + warp - kh;              // input row 4*mt + warp feeds output row j through tap kh
          if (j >= 0 && j < C::OROWS) {
            float* e = ex + ((kh * C::OROWS + j) * (CP / 4) * PW + lane) * 4;
#pragma unroll
            for (int c = 0; c < CP; c += 4)
              *reinterpret_cast<float4*>(e + (c / 4) * PW * 4) =
                  make_float4(__uint_as_float(vr[kh * CP + c]), __uint_as_float(vr[kh * CP + c + 1]),
                              __uint_as_float(vr[kh * CP + c + 2]), __uint_as_float(vr[kh * CP + c + 3]));
          }
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(grp_free + g);
      epi_bar_sync();
      for (int j = warp; j < C::OROWS; j += 4) {
        const int gy = y0 + j;
        if (!col_ok || gy >= p.H) continue;
        const float* e = ex + (j * (CP / 4) * PW + lane) * 4;
        constexpr int EXP = C::OROWS * PW * CP;
        const int64_t o = (((int64_t)n * p.H + gy) * p.W + gx) * COUT;
#pragma unroll
        for (int c = 0; c < COUT; c += 4) {
          float4 r = *reinterpret_cast<const float4*>(e + (c / 4) * PW * 4);
          const float4 r1 = *reinterpret_cast<const float4*>(e + EXP + (c / 4) * PW * 4);
          const float4 r2 = *reinterpret_cast<const float4*>(e + 2 * EXP + (c / 4) * PW * 4);
          r.x += r1.x + r2.x; r.y += r1.y + r2.y; r.z += r1.z + r2.z; r.w += r1.w + r2.w;
          if (p.bias) {
            const float4 t4 = ldg4(p.bias + c);
            r.x += t4.x; r.y += t4.y; r.z += t4.z; r.w += t4.w;
          }
          if (p.relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
          if (p.y) st4(p.y + o + c, r);
          if (p.ys) store_split4(p.ys, part_stride, o + c, r);
        }
      }
      epi_bar_sync();                              // exchange buffer is reused by the next item
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// Weight image: torch Conv2d weight [Cout][Cin][3][3] x BN scale -> img[term][step][chunk][row = kh*CP + n][e]
//   Cin >= 16: step = kw*KS + ks, channel = ks*16 + chunk*8 + e;  Cin == 8: step s covers taps kw = 2s + chunk (kw = 3: zero)
//   term 0 = bf16(w·s), term 1 = bf16(rest)
__global__ void pack_slab2d_weight_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                          __nv_bfloat16* __restrict__ img, int Cout, int Cin, int CP, int G) {
  const int KS = Cin / 16;
  const int nstep = Cin == 8 ? 2 : 3 * KS;
  const int per_term = nstep * 2 * G * 8;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per_term) return;
  const int e = i & 7;
  int r = i >> 3;
  const int row = r % G; r /= G;
  const int chunk = r & 1; r >>= 1;
  const int st = r;
  const int kw = Cin == 8 ? 2 * st + chunk : st / KS;
  const int ci = Cin == 8 ? e : (st % KS) * 16 + chunk * 8 + e;
  const int kh = row / CP, n = row % CP;
  float v = 0.f;
  if (kh < 3 && n < Cout && kw < 3) {
    v = w[((int64_t)n * Cin + ci) * 9 + kh * 3 + kw];
    if (scale) v *= scale[n];
  }
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  img[i] = hi;
  img[per_term + i] = __float2bfloat16_rn(v - __bfloat162float(hi));
}

template <int CIN, int COUT, int KM, int NS>
static int launch(const void* x, Params p, cudaStream_t st) {
  using C = Cfg<CIN, COUT, KM, NS>;
  static_assert(C::SMEM <= 232448, "shared memory budget exceeded");
  auto kern = conv2d_slab_bf16_kernel<CIN, COUT, KM, NS>;
  static SmemAttrOnce smem_attr;
  {
    cudaError_t e = smem_attr.ensure(kern, (int)C::SMEM);
    RCMVS_REQUIRE(e == cudaSuccess, RCMVS_ERR_CUDA, "rcmvs_conv2d_slab_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  CUtensorMap map;
  // the 5-D map of the 3-D kernel with the images as its depth dimension and batch 1: dims {C, W, H, N, 2}
  const int rc = make_tmap(&map, x, 1, p.N, p.H, p.W, CIN, C::ROWS);
  if (rc != RCMVS_OK) return rc;
  p.tiles_x = ceil_div(p.W, TW);
  p.tiles_y = ceil_div(p.H, C::OROWS);
  p.items = p.tiles_x * p.tiles_y * p.N;
  const int ctas = p.items < 148 ? p.items : 148;                 // persistent: one CTA per SM
  p.items_per_cta = ceil_div(p.items, ctas);
  kern<<<ceil_div(p.items, p.items_per_cta), kThreads, C::SMEM, st>>>(map, p);
  RCMVS_CHECK_LAUNCH("rcmvs_conv2d_slab_fwd");
  return RCMVS_OK;
}

static bool supported(int Cin, int Cout) {
  return (Cin == 8 && Cout == 8) || (Cin == 16 && Cout == 16) || (Cin == 32 && (Cout == 32 || Cout == 16 || Cout == 8));
}

}  // namespace slab2d
}  // namespace rcmvs

extern "C" int64_t rcmvs_conv2d_slab_weight_elems(int Cin, int Cout) {
  const int cp = Cout < 8 ? 8 : Cout;
  const int g = ((3 * cp + 15) / 16) * 16;
  return (int64_t)2 * (Cin == 8 ? 2 : 3 * (Cin / 16)) * 2 * g * 8;
}

extern "C" int rcmvs_conv2d_slab_pack(const float* w, const float* scale, void* img, int Cout, int Cin,
                                      rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(w && img, RCMVS_ERR_BAD_ARG, "rcmvs_conv2d_slab_pack: null pointer");
  RCMVS_REQUIRE(slab2d::supported(Cin, Cout), RCMVS_ERR_UNSUPPORTED, "rcmvs_conv2d_slab_pack: unsupported Cin=%d Cout=%d", Cin,
                Cout);
  const int cp = Cout < 8 ? 8 : Cout;
  const int g = ((3 * cp + 15) / 16) * 16;
  const int per_term = (Cin == 8 ? 2 : 3 * (Cin / 16)) * 2 * g * 8;
  slab2d::pack_slab2d_weight_kernel<<<ceil_div(per_term, 256), 256, 0, as_stream(stream)>>>(
      w, scale, reinterpret_cast<__nv_bfloat16*>(img), Cout, Cin, cp, g);
  RCMVS_CHECK_LAUNCH("rcmvs_conv2d_slab_pack");
  return RCMVS_OK;
}

extern "C" int rcmvs_conv2d_slab_fwd(const void* x_split, const void* wimg, const float* bias, float* y, void* y_split, int N,
                                     int H, int W, int Cin, int Cout, int relu, rcmvs_stream_t stream) {
  using namespace rcmvs;
  using namespace rcmvs::slab2d;
  RCMVS_REQUIRE(x_split && wimg && (y || y_split), RCMVS_ERR_BAD_ARG, "rcmvs_conv2d_slab_fwd: null pointer");
  RCMVS_REQUIRE(N > 0 && H > 0 && W > 0, RCMVS_ERR_BAD_ARG, "rcmvs_conv2d_slab_fwd: bad dims");
  RCMVS_REQUIRE(slab2d::supported(Cin, Cout), RCMVS_ERR_UNSUPPORTED, "rcmvs_conv2d_slab_fwd: unsupported Cin=%d Cout=%d", Cin,
                Cout);
  RCMVS_REQUIRE((reinterpret_cast<uintptr_t>(x_split) & 15) == 0, RCMVS_ERR_BAD_ARG,
                "rcmvs_conv2d_slab_fwd: x_split must be 16-byte aligned (TMA)");
  Params p;
  p.wimg = reinterpret_cast<const __nv_bfloat16*>(wimg);
  p.bias = bias; p.y = y; p.ys = reinterpret_cast<__nv_bfloat16*>(y_split);
  p.N = N; p.H = H; p.W = W; p.relu = relu;
  p.tiles_x = p.tiles_y = p.items = p.items_per_cta = 0;
  cudaStream_t st = as_stream(stream);
  if (Cin == 8 && Cout == 8) return launch<8, 8, 3, 8>(x_split, p, st);         // conv0.0 (3 -> 8, input padded to 8), conv0.1
  if (Cin == 16 && Cout == 16) return launch<16, 16, 3, 6>(x_split, p, st);
  if (Cin == 32 && Cout == 32) return launch<32, 32, 2, 3>(x_split, p, st);
  if (Cin == 32 && Cout == 16) return launch<32, 16, 3, 3>(x_split, p, st);
  if (Cin == 32 && Cout == 8) return launch<32, 8, 3, 3>(x_split, p, st);
  return RCMVS_ERR_UNSUPPORTED;
}
