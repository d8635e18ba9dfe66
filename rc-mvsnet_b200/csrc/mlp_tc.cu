// a10 Renderer_ours MLP on the tensor cores (reference models/render_models.py:45-49 embedding, :143-220 Renderer_ours;
// called per ray chunk from models/renderer.py:42-63).  Same inputs, outputs and fp32 parameters as the fused FFMA kernel
// `render_mlp_kernel` (csrc/render.cu); this path runs every 128 / 64-wide Linear layer as a TMA-fed tcgen05 GEMM on
// error-compensated split-bf16 operands ("bf16x3", see conv3d_slab_bf16.cu):
//
//   prep     embedding gamma(ndc) (63 -> 64 columns), view directions, mod = pts_bias(feat) (20 -> 128, FFMA, fp32)
//   L0..L4   h = relu((W h + b) * mod)                      128-wide, K = 64 / 128          (render_models.py:196-203)
//   L5       h = relu((W [gamma, h] + b) * mod), K = 192;   alpha = relu(alpha_linear(h)) in the epilogue (:205)
//   F        feature = feature_linear(h)                    (no activation, :209)
//   V        h2 = relu(views_linears.0([feature, dir])), 64-wide, K = 192 (131 used);  rgb = sigmoid(rgb_linear(h2))
//            in the epilogue; out = [rgb, alpha]            (:210-216)
//
// Activations live in global memory as split-bf16 row-major matrices [2][M][ld] (hi plane, lo plane) written by the
// previous layer's epilogue; a [8 columns x 128 rows] TMA box of such a matrix IS a K-major SWIZZLE_NONE UMMA operand
// chunk.  The concatenations of the reference are column ranges of one buffer: layer 4 writes its output next to the
// embedding, the feature layer next to the view direction.
//
// GEMM kernel: persistent CTAs over 128-row tiles; the layer's weight image stays in shared memory; A arrives in
// K blocks of 64 through a 3-stage TMA ring; two TMEM accumulators double-buffer the MMAs against the epilogue.
// Warp roles: warps 0-3 epilogue (thread = one row), warp 4 TMA producer, warp 5 MMA issuer.
#include "slab_common.cuh"

namespace rcmvs {
namespace mlptc {

using namespace rcmvs::slab;

constexpr int TM = 128;                         // rows per tile
constexpr int KB = 64;                          // K block (columns) per pipeline stage
constexpr int NS = 3;
constexpr int STAGE_B = 2 * (KB / 8) * TM * 16; // hi + lo, 8 chunks of [128 rows][16 B]
constexpr int LD0 = 192, LDH = 128, LD2 = 192;  // row lengths of X0 (gamma | h4), HA / HB, X2 (feature | dir | 0)

struct GemmParams {
  const __nv_bfloat16* wimg;   // [2][K/8][N][8]
  const float* bias;           // [N]
  const float* mod;            // [Mp][128] fp32 or null
  __nv_bfloat16* y;            // split output [2][Mp][ldy] (columns yo .. yo+N) or null
  int ldy, yo;
  int64_t Mp, P;               // padded / real row count
  int K, k0;                   // K (multiple of 64) columns starting at k0 of the A matrix
  int tiles;
  // heads
  const float* walpha; const float* balpha; float* alpha;     // MODE 1 writes alpha[Mp]
  const float* wrgb; const float* brgb; float* out;           // MODE 3 reads alpha, writes out[P][4]
};

// MODE 0: y = relu((acc + b) * mod)   1: same + alpha head   2: y = acc + b   3: N = 64, relu + rgb head
template <int N, int MODE>
__global__ void __launch_bounds__(kThreads, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap tmap, const GemmParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* a_smem = smem;
  unsigned char* b_smem = smem + NS * STAGE_B;
  const int b_bytes = 2 * (p.K / 8) * N * 16;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_smem + 2 * (192 / 8) * N * 16);
  uint64_t* full = bars;                 // [NS]
  uint64_t* empty = bars + NS;           // [NS]
  uint64_t* acc_full = bars + 2 * NS;    // [2]
  uint64_t* acc_empty = acc_full + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  constexpr uint32_t TCOLS = 2 * N;      // 256 or 128

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = p.K / KB;
  const int my_tiles = (p.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4 && lane == 0) tma_prefetch_desc(&tmap);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {
    const float4* src = reinterpret_cast<const float4*>(p.wimg);
    float4* dst = reinterpret_cast<float4*>(b_smem);
    for (int i = threadIdx.x; i < b_bytes / 16; i += kThreads) dst[i] = __ldg(src + i);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // =============================== TMA producer =================================================================
    if (lane == 0) {
      int it = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const int m0 = ((int)blockIdx.x + t * (int)gridDim.x) * TM;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int buf = it % NS;
          mbar_wait(empty + buf, ((it / NS) & 1) ^ 1);
          mbar_arrive_expect_tx(full + buf, (uint32_t)STAGE_B);
          const uint32_t dst0 = smem_u32(a_smem + (size_t)buf * STAGE_B);
#pragma unroll
          for (int part = 0; part < 2; ++part)
#pragma unroll
            for (int ch = 0; ch < KB / 8; ++ch)      // map dims {ld, Mp, 1, 1, 2}: box {8 columns, 128 rows}
              tma_load_5d(dst0 + (part * (KB / 8) + ch) * (TM * 16), &tmap, full + buf, p.k0 + kb * KB + ch * 8, m0, 0, 0, part);
        }
      }
    }
  } else if (warp == 5) {
    // =============================== MMA issuer ====================================================================
    const bool leader = elect_one();
    const uint32_t b_base = smem_u32(b_smem);
    constexpr uint32_t idesc = make_idesc_bf16(128, N);
    const uint32_t bterm = (uint32_t)((p.K / 8) * N * 16) >> 4;           // lo weights follow the hi weights
    int it = 0;
    for (int t = 0; t < my_tiles; ++t) {
      const int acc = t & 1;
      mbar_wait(acc_empty + acc, ((t >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d = tmem_base + (uint32_t)(acc * N);
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int buf = it % NS;
        mbar_wait(full + buf, (it / NS) & 1);
        tc_fence_after();
        const uint32_t a_hi0 = desc_lo(smem_u32(a_smem + (size_t)buf * STAGE_B), TM * 16);   // LBO = next 8-column chunk
        constexpr uint32_t kLoOff = ((KB / 8) * TM * 16) >> 4;
#pragma unroll
        for (int ks = 0; ks < KB / 16; ++ks) {
          const uint32_t a_d = a_hi0 + (uint32_t)((2 * ks * TM * 16) >> 4);
          const uint32_t b_d = desc_lo(b_base + (uint32_t)((kb * (KB / 8) + 2 * ks) * N * 16), N * 16);
          if (leader) {
            umma_bf16_lohi_acc(d, a_d + kLoOff, b_d, idesc, (kb == 0 && ks == 0) ? 0u : 1u);   // x_lo · w_hi
            umma_bf16_lohi_acc(d, a_d, b_d + bterm, idesc, 1u);                                 // x_hi · w_lo
            umma_bf16_lohi_acc(d, a_d, b_d, idesc, 1u);                                         // x_hi · w_hi
          }
        }
        if (leader) umma_commit(empty + buf);
      }
      if (leader) umma_commit(acc_full + acc);
      __syncwarp();
    }
  } else {
    // =============================== epilogue (128 threads, thread = one row) =====================================
    for (int t = 0; t < my_tiles; ++t) {
      const int acc = t & 1;
      const int64_t row = (int64_t)((int)blockIdx.x + t * (int)gridDim.x) * TM + warp * 32 + lane;
      mbar_wait(acc_full + acc, (t >> 1) & 1);
      tc_fence_after();
      float dot = 0.f;                                   // MODE 1: alpha head
      float r3[3] = {0.f, 0.f, 0.f};                     // MODE 3: rgb head
#pragma unroll 1
      for (int g = 0; g < N / 32; ++g) {
        uint32_t vr[32];
        const uint32_t ta = tmem_base + ((uint32_t)(warp * 32) << 16) + acc * N + g * 32;
        tmem_ld16_nowait(ta, &vr[0]);
        tmem_ld16_nowait(ta + 16, &vr[16]);
        tmem_ld_wait();
        if (g == N / 32 - 1) {
          tc_fence_before();
          mbar_arrive(acc_empty + acc);
        }
#pragma unroll
        for (int c8 = 0; c8 < 32; c8 += 8) {
          const int col = g * 32 + c8;
          float v[8];
          const float4 b0 = ldg4(p.bias + col), b1 = ldg4(p.bias + col + 4);
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = __uint_as_float(vr[c8 + k]) + bb[k];
          if constexpr (MODE == 0 || MODE == 1) {
            const float4 m0 = ldg4(p.mod + row * 128 + col), m1 = ldg4(p.mod + row * 128 + col + 4);
            const float mm[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = fmaxf(v[k] * mm[k], 0.f);
          }
          if constexpr (MODE == 1) {
            const float4 w0 = ldg4(p.walpha + col), w1 = ldg4(p.walpha + col + 4);
            const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int k = 0; k < 8; ++k) dot = fmaf(v[k], ww[k], dot);
          }
          if constexpr (MODE == 3) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              v[k] = fmaxf(v[k], 0.f);
              r3[0] = fmaf(v[k], __ldg(p.wrgb + (col + k) * 3 + 0), r3[0]);
              r3[1] = fmaf(v[k], __ldg(p.wrgb + (col + k) * 3 + 1), r3[1]);
              r3[2] = fmaf(v[k], __ldg(p.wrgb + (col + k) * 3 + 2), r3[2]);
            }
          } else {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * k]), h1 = __float2bfloat16_rn(v[2 * k + 1]);
              const __nv_bfloat162 hh = __halves2bfloat162(h0, h1);
              const __nv_bfloat162 ll = __floats2bfloat162_rn(v[2 * k] - __bfloat162float(h0), v[2 * k + 1] - __bfloat162float(h1));
              hi[k] = *reinterpret_cast<const uint32_t*>(&hh);
              lo[k] = *reinterpret_cast<const uint32_t*>(&ll);
            }
            __nv_bfloat16* yo = p.y + row * p.ldy + p.yo + col;
            *reinterpret_cast<uint4*>(yo) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(yo + p.Mp * p.ldy) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
      }
      if constexpr (MODE == 1) p.alpha[row] = fmaxf(dot + __ldg(p.balpha), 0.f);
      if constexpr (MODE == 3) {
        if (row < p.P)
          st4(p.out + row * 4, make_float4(1.f / (1.f + expf(-(r3[0] + __ldg(p.brgb)))), 1.f / (1.f + expf(-(r3[1] + __ldg(p.brgb + 1)))),
                                           1.f / (1.f + expf(-(r3[2] + __ldg(p.brgb + 2)))), p.alpha[row]));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCOLS) : "memory");
  }
}

// ---- prep: embedding, view directions, multiplicative modulation --------------------------------------------------------------
// thread = one row (point).  X0[:, 0:64] = [x, sin(x·2^0..2^9) (30), cos(...) (30), 0]; X2[:, 128:192] = [dir (3), 0...];
// mod = pts_bias(feat).  Rows P..Mp-1 (tile padding) are written as zeros.
__global__ void __launch_bounds__(128)
mlp_prep_kernel(const float* __restrict__ ndc, const float* __restrict__ feat, const float* __restrict__ dirs,
                const float* __restrict__ wbias_t /* [20][128] */, const float* __restrict__ bbias, __nv_bfloat16* __restrict__ x0,
                __nv_bfloat16* __restrict__ x2, float* __restrict__ mod, int64_t P, int64_t Mp, int S) {
  __shared__ __align__(16) float ws[20 * 128 + 128];
  for (int i = threadIdx.x; i < 20 * 128; i += 128) ws[i] = __ldg(wbias_t + i);
  ws[20 * 128 + threadIdx.x] = __ldg(bbias + threadIdx.x);
  __syncthreads();
  const int64_t row = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (row >= Mp) return;
  const bool real = row < P;
  auto store8 = [&](__nv_bfloat16* base, int64_t ld, int col, const float (&v)[8]) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * k]), h1 = __float2bfloat16_rn(v[2 * k + 1]);
      const __nv_bfloat162 hh = __halves2bfloat162(h0, h1);
      const __nv_bfloat162 ll = __floats2bfloat162_rn(v[2 * k] - __bfloat162float(h0), v[2 * k + 1] - __bfloat162float(h1));
      hi[k] = *reinterpret_cast<const uint32_t*>(&hh);
      lo[k] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    *reinterpret_cast<uint4*>(base + row * ld + col) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(base + Mp * ld + row * ld + col) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  };
  float x[3] = {0.f, 0.f, 0.f};
  if (real) { x[0] = __ldg(ndc + row * 3); x[1] = __ldg(ndc + row * 3 + 1); x[2] = __ldg(ndc + row * 3 + 2); }
  // embedding column c: c < 3: x[c]; 3 <= c < 63: j = (c-3) % 30, f = j / 3, a = j % 3, c < 33 ? sin : cos  (render_models.py:45-49)
#pragma unroll 1
  for (int c0 = 0; c0 < 64; c0 += 8) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = c0 + k;
      float e = 0.f;
      if (real) {
        if (c < 3) e = x[c];
        else if (c < 63) {
          const int j = (c - 3) % 30, f = j / 3, a = j % 3;
          const float xx = x[a] * (float)(1 << f);
          e = (c < 33) ? sinf(xx) : cosf(xx);
        }
      }
      v[k] = e;
    }
    store8(x0, LD0, c0, v);
  }
  {
    float d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (real) { const int64_t ray = row / S; d[0] = __ldg(dirs + ray * 3); d[1] = __ldg(dirs + ray * 3 + 1); d[2] = __ldg(dirs + ray * 3 + 2); }
    store8(x2, LD2, 128, d);
    const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int c = 136; c < 192; c += 8) store8(x2, LD2, c, z);
  }
  float f[20];
#pragma unroll
  for (int k = 0; k < 20; ++k) f[k] = real ? __ldg(feat + row * 20 + k) : 0.f;
#pragma unroll 1
  for (int n0 = 0; n0 < 128; n0 += 4) {
    float4 a = *reinterpret_cast<const float4*>(ws + 20 * 128 + n0);
#pragma unroll
    for (int k = 0; k < 20; ++k) {
      const float4 w = *reinterpret_cast<const float4*>(ws + k * 128 + n0);
      a.x = fmaf(f[k], w.x, a.x); a.y = fmaf(f[k], w.y, a.y); a.z = fmaf(f[k], w.z, a.z); a.w = fmaf(f[k], w.w, a.w);
    }
    st4(mod + row * 128 + n0, a);
  }
}

// Weight image of one layer from the transposed fp32 weight Wt [K_in][N]:  img[term][k'/8][n][k' % 8], k' in [0, Kpad):
// source row = k' (k' < gap_at), zero (gap_at <= k' < gap_at + gap_len or beyond the input), k' - gap_len otherwise
__global__ void pack_linear_kernel(const float* __restrict__ wt, __nv_bfloat16* __restrict__ img, int N, int K_in, int Kpad,
                                   int gap_at, int gap_len) {
  const int per_term = Kpad * N;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per_term) return;
  const int e = i & 7;
  const int n = (i >> 3) % N;
  const int ch = i / (8 * N);
  const int kp = ch * 8 + e;
  int ks = kp < gap_at ? kp : (kp < gap_at + gap_len ? -1 : kp - gap_len);
  if (ks >= K_in) ks = -1;
  const float v = ks >= 0 ? wt[(int64_t)ks * N + n] : 0.f;
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  img[i] = hi;
  img[per_term + i] = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// layers: {N, K_in, Kpad, gap_at, gap_len}: L0, L1..L4, L5, feature, views
struct LayerDef { int N, K_in, Kpad, gap_at, gap_len; };
static const LayerDef kLayers[8] = {{128, 63, 64, 63, 1},    {128, 128, 128, 128, 0}, {128, 128, 128, 128, 0}, {128, 128, 128, 128, 0},
                                    {128, 128, 128, 128, 0}, {128, 191, 192, 63, 1},  {128, 128, 128, 128, 0}, {64, 131, 192, 131, 61}};
static int64_t layer_offset(int l) {           // element offset of layer l's image
  int64_t o = 0;
  for (int i = 0; i < l; ++i) o += 2ll * kLayers[i].Kpad * kLayers[i].N;
  return o;
}

// 5-D map over a split matrix [2][Mp][ld]: dims {ld, Mp, 1, 1, 2}, box {8 columns, 128 rows, 1, 1, 1}
static int make_tmap_rows(CUtensorMap* map, const void* x, int64_t Mp, int ld) {
  EncodeTiledFn fn = encode_fn();
  RCMVS_REQUIRE(fn != nullptr, RCMVS_ERR_CUDA, "rcmvs_render_mlp_tc: cuTensorMapEncodeTiled not available");
  const cuuint64_t gdim[5] = {(cuuint64_t)ld, (cuuint64_t)Mp, 1, 1, 2};
  const cuuint64_t gstr[4] = {(cuuint64_t)ld * 2, (cuuint64_t)Mp * ld * 2, (cuuint64_t)Mp * ld * 2, (cuuint64_t)Mp * ld * 2};
  const cuuint32_t box[5] = {8, (cuuint32_t)TM, 1, 1, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RCMVS_REQUIRE(r == CUDA_SUCCESS, RCMVS_ERR_CUDA, "rcmvs_render_mlp_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return RCMVS_OK;
}

template <int N, int MODE>
static int launch_gemm(const void* a, int lda, GemmParams p, cudaStream_t st) {
  constexpr size_t SMEM = (size_t)NS * STAGE_B + 2 * (192 / 8) * N * 16 + 256;
  static_assert(SMEM <= 232448, "shared memory budget exceeded");
  auto kern = linear_tc_kernel<N, MODE>;
  static SmemAttrOnce smem_attr;
  {
    cudaError_t e = smem_attr.ensure(kern, (int)SMEM);
    RCMVS_REQUIRE(e == cudaSuccess, RCMVS_ERR_CUDA, "rcmvs_render_mlp_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  CUtensorMap map;
  const int rc = make_tmap_rows(&map, a, p.Mp, lda);
  if (rc != RCMVS_OK) return rc;
  p.tiles = (int)(p.Mp / TM);
  const int grid = p.tiles < 148 ? p.tiles : 148;
  kern<<<grid, kThreads, SMEM, st>>>(map, p);
  RCMVS_CHECK_LAUNCH("rcmvs_render_mlp_tc");
  return RCMVS_OK;
}

}  // namespace mlptc
}  // namespace rcmvs

extern "C" int64_t rcmvs_render_mlp_tc_weight_elems(void) { return rcmvs::mlptc::layer_offset(8); }

extern "C" int64_t rcmvs_render_mlp_tc_workspace_bytes(int64_t P) {
  const int64_t Mp = (P + 127) / 128 * 128;
  // X0 [2][Mp][192] + HA, HB [2][Mp][128] + X2 [2][Mp][192] bf16, mod [Mp][128] + alpha [Mp] fp32
  return Mp * (2 * 192 * 2 + 2 * 2 * 128 * 2 + 2 * 192 * 2 + 128 * 4 + 4);
}

extern "C" int rcmvs_render_mlp_tc_pack(const float* const* weights, void* wimg, rcmvs_stream_t stream) {
  using namespace rcmvs;
  using namespace rcmvs::mlptc;
  RCMVS_REQUIRE(weights && wimg, RCMVS_ERR_BAD_ARG, "rcmvs_render_mlp_tc_pack: null pointer");
  // the 22 pointers of rcmvs_render_mlp: pts_linears.i at 2i, feature_linear at 16, views_linears.0 at 18
  const int src[8] = {0, 2, 4, 6, 8, 10, 16, 18};
  __nv_bfloat16* img = reinterpret_cast<__nv_bfloat16*>(wimg);
  for (int l = 0; l < 8; ++l) {
    RCMVS_REQUIRE(weights[src[l]], RCMVS_ERR_BAD_ARG, "rcmvs_render_mlp_tc_pack: null weight %d", src[l]);
    const LayerDef& L = kLayers[l];
    const int per_term = L.Kpad * L.N;
    pack_linear_kernel<<<ceil_div(per_term, 256), 256, 0, as_stream(stream)>>>(weights[src[l]], img + layer_offset(l), L.N, L.K_in,
                                                                             L.Kpad, L.gap_at, L.gap_len);
  }
  RCMVS_CHECK_LAUNCH("rcmvs_render_mlp_tc_pack");
  return RCMVS_OK;
}

extern "C" int rcmvs_render_mlp_tc(const float* ndc, const float* feat, const float* dirs, const float* const* weights,
                                   const void* wimg, void* workspace, float* out, int64_t P, int S, rcmvs_stream_t stream) {
  using namespace rcmvs;
  using namespace rcmvs::mlptc;
  RCMVS_REQUIRE(ndc && feat && dirs && weights && wimg && workspace && out, RCMVS_ERR_BAD_ARG, "rcmvs_render_mlp_tc: null pointer");
  RCMVS_REQUIRE(P > 0 && S > 0, RCMVS_ERR_BAD_ARG, "rcmvs_render_mlp_tc: bad dims");
  for (int i = 0; i < 22; ++i) RCMVS_REQUIRE(weights[i], RCMVS_ERR_BAD_ARG, "rcmvs_render_mlp_tc: null weight %d", i);
  RCMVS_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, RCMVS_ERR_BAD_ARG, "rcmvs_render_mlp_tc: workspace alignment");
  const int64_t Mp = (P + 127) / 128 * 128;
  cudaStream_t st = as_stream(stream);
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
  __nv_bfloat16* X0 = reinterpret_cast<__nv_bfloat16*>(ws); ws += Mp * 2 * LD0 * 2;
  __nv_bfloat16* HA = reinterpret_cast<__nv_bfloat16*>(ws); ws += Mp * 2 * LDH * 2;
  __nv_bfloat16* HB = reinterpret_cast<__nv_bfloat16*>(ws); ws += Mp * 2 * LDH * 2;
  __nv_bfloat16* X2 = reinterpret_cast<__nv_bfloat16*>(ws); ws += Mp * 2 * LD2 * 2;
  float* MOD = reinterpret_cast<float*>(ws); ws += Mp * 128 * 4;
  float* ALPHA = reinterpret_cast<float*>(ws);
  const __nv_bfloat16* img = reinterpret_cast<const __nv_bfloat16*>(wimg);

  mlp_prep_kernel<<<(unsigned)(Mp / 128), 128, 0, st>>>(ndc, feat, dirs, weights[12], weights[13], X0, X2, MOD, P, Mp, S);
  RCMVS_CHECK_LAUNCH("rcmvs_render_mlp_tc");

  GemmParams g;
  g.Mp = Mp; g.P = P; g.mod = MOD; g.tiles = 0;
  g.walpha = weights[14]; g.balpha = weights[15]; g.alpha = ALPHA; g.wrgb = weights[20]; g.brgb = weights[21]; g.out = out;
  int rc;
  // L0: gamma (X0 cols 0..63) -> HA
  g.wimg = img + layer_offset(0); g.bias = weights[1]; g.y = HA; g.ldy = LDH; g.yo = 0; g.K = 64; g.k0 = 0;
  if ((rc = launch_gemm<128, 0>(X0, LD0, g, st)) != RCMVS_OK) return rc;
  // L1: HA -> HB, L2: HB -> HA, L3: HA -> HB
  g.K = 128;
  g.wimg = img + layer_offset(1); g.bias = weights[3]; g.y = HB;
  if ((rc = launch_gemm<128, 0>(HA, LDH, g, st)) != RCMVS_OK) return rc;
  g.wimg = img + layer_offset(2); g.bias = weights[5]; g.y = HA;
  if ((rc = launch_gemm<128, 0>(HB, LDH, g, st)) != RCMVS_OK) return rc;
  g.wimg = img + layer_offset(3); g.bias = weights[7]; g.y = HB;
  if ((rc = launch_gemm<128, 0>(HA, LDH, g, st)) != RCMVS_OK) return rc;
  // L4: HB -> X0 cols 64..191 (the skip concatenation [gamma, h], render_models.py:202-203)
  g.wimg = img + layer_offset(4); g.bias = weights[9]; g.y = X0; g.ldy = LD0; g.yo = 64;
  if ((rc = launch_gemm<128, 0>(HB, LDH, g, st)) != RCMVS_OK) return rc;
  // L5: X0 (K = 192) -> HA, alpha head
  g.wimg = img + layer_offset(5); g.bias = weights[11]; g.y = HA; g.ldy = LDH; g.yo = 0; g.K = 192;
  if ((rc = launch_gemm<128, 1>(X0, LD0, g, st)) != RCMVS_OK) return rc;
  // feature: HA -> X2 cols 0..127 (next to the view direction)
  g.wimg = img + layer_offset(6); g.bias = weights[17]; g.y = X2; g.ldy = LD2; g.yo = 0; g.K = 128; g.mod = nullptr;
  if ((rc = launch_gemm<128, 2>(HA, LDH, g, st)) != RCMVS_OK) return rc;
  // views + rgb head: X2 (K = 192, 131 used) -> out
  g.wimg = img + layer_offset(7); g.bias = weights[19]; g.y = nullptr; g.K = 192;
  return launch_gemm<64, 3>(X2, LD2, g, st);
}
