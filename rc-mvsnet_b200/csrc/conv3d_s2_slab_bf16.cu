// a3 CostRegNet conv1 (8 -> 16 channels, 3x3x3, stride 2, pad 1; reference models/modules.py:118-163 Conv3d, :472 conv1;
// also conv1 of the renderer's CostReg, models/render_models.py:696) on tcgen05 with split-bf16 operands ("bf16x3", see
// conv3d_slab_bf16.cu) fed by STRIDED TMA boxes.
//
// A stride-2 convolution reads input (2o + k - 1) per axis: tap k = 1 reads the EVEN samples at index o, taps k = 0 / 2 read
// the ODD samples at index o - 1 / o.  One input plane z is therefore fetched as four parity sub-planes
//      P[pr][pc][r][c] = x[z][2·(oy0 - 1 + r) + pr][2·(ox0 - 1 + c) + pc][0..7]
// by four TMA box loads with elementStrides = {1, 2, 2, 1, 1} (cuTensorMapEncodeTiled: a box of 2N elements with traversal
// stride 2 lands N elements, densely), out-of-range coordinates zero-filled = the convolution's padding.  In sub-plane
// coordinates every tap is a plain offset (dr, dc) in {-1, 0}², i.e. a start-address offset of the UMMA A descriptor over a
// flat M = 128 run (4 output rows x 32 columns, column 0 = halo).  With 8 input channels one K = 16 MMA step covers TWO
// taps: K-chunk 1 is addressed through the descriptor's leading-dimension byte offset (LBO), which is the distance
// between the two taps' sub-plane positions:
//      pair 0..2: (kh, kw=1) + (kh, kw=2)   even-column -> odd-column sub-plane            LBO = SUB_B
//      pair 3   : (kh=0, kw=0) + (kh=2, kw=0)  same sub-plane, next row                    LBO = 512 B
//      pair 4   : (kh=1, kw=0) + zero weights (chunk 1 points at finite slab data)         LBO = 2·SUB_B
// kd: an even input plane z only meets kd = 1 (output plane z/2, N = 16); an odd one meets kd = 2 (plane (z-1)/2) and
// kd = 0 (plane (z+1)/2): N = 32 over two neighbouring column groups of a TMEM ring of per-output-plane accumulators.
// All 27 taps are summed inside TMEM; the epilogue only adds the BN shift, applies ReLU and stores fp32 and / or the
// pre-split form for the next tensor-core layer.
//
// Warp roles as in the stride-1 kernel: warps 0-3 epilogue, warp 4 TMA producer, warp 5 MMA issuer.
#include "slab_common.cuh"

namespace rcmvs {
namespace slab2 {

using namespace rcmvs::slab;

constexpr int TWO = 31;                                            // real output columns per tile (column 0 = halo)
constexpr int NG = 4;                                              // TMEM ring: output planes in flight

// (CIN, COUT) = (8, 16): conv1 — 5 steps of two taps each (see above).  (16, 32): conv3 (half -> quarter resolution,
// modules.py:476) — 16 channels fill a K = 16 step, so the 9 (kh, kw) taps are 9 steps and K-chunk 1 is the second chunk
// plane of the same sub-plane.
template <int CIN, int COUT, int KM, int NS>
struct Cfg {
  static constexpr int NCH = CIN / 8;
  static constexpr int NSTEP = CIN == 8 ? 5 : 9;
  static constexpr int ROWS_S = 4 * KM + 1;                        // sub-plane rows (1 halo row above)
  static constexpr int OROWS = 4 * KM;
  static constexpr int PLANE_B = ROWS_S * PW * 16;                 // one 8-channel chunk plane of a sub-plane
  static constexpr int SUB_B = NCH * PLANE_B;                      // one parity sub-plane of one part
  static constexpr int PART_B = 4 * SUB_B;
  static constexpr int STAGE_B = 2 * PART_B;                       // hi, lo
  static constexpr int A_BYTES = NS * STAGE_B + 128;
  static constexpr int ODD_B = NSTEP * 2 * (2 * COUT) * 16;        // steps x 2 K-chunks x 2·COUT rows (kd=2 | kd=0) x 16 B
  static constexpr int EVEN_B = NSTEP * 2 * COUT * 16;             // steps x 2 K-chunks x COUT rows (kd=1)
  static constexpr int BTERM_B = ODD_B + EVEN_B;
  static constexpr int B_BYTES = 2 * BTERM_B;                      // hi, lo
  static constexpr int TCOLS = NG * KM * COUT;                     // accumulator columns in use
  static constexpr int TALLOC = TCOLS <= 128 ? 128 : 256;
  static constexpr size_t SMEM = (size_t)A_BYTES + B_BYTES + 512;
  static_assert(PLANE_B % 128 == 0, "TMA destination alignment");
  static_assert(TCOLS <= 256 && (CIN == 8 || CIN == 16), "configuration");
};

struct Params {
  const __nv_bfloat16* wimg;
  const float* bias;           // [COUT] or null
  float* y;                    // [B][Do][Ho][Wo][COUT] fp32 or null
  __nv_bfloat16* ys;           // [2][B][Do][Ho][Wo][COUT] split-bf16 or null
  int B, D, H, W;              // input dims (even)
  int Do, Ho, Wo;
  int d_chunk, chunks;         // output planes per CTA
  int relu;
};

// tap pair p: sub-plane of K-chunk 0 (sp = 2·pr + pc), offset (dr, dc) in positions, LBO in units of SUB_B (0 = one row)
//            p0 (kh0: kw1,kw2)  p1 (kh1: kw1,kw2)  p2 (kh2: kw1,kw2)  p3 (kw0: kh0,kh2)  p4 (kh1,kw0 + zero)
__device__ __forceinline__ constexpr int pair_sp(int p) { return p == 0 ? 2 : p == 1 ? 0 : p == 2 ? 2 : p == 3 ? 3 : 1; }
__device__ __forceinline__ constexpr int pair_off(int p) { return p == 0 ? -PW : p == 1 ? 0 : p == 2 ? 0 : p == 3 ? -PW - 1 : -1; }
__device__ __forceinline__ constexpr int pair_lbo_sub(int p) { return p == 3 ? 0 : p == 4 ? 2 : 1; }
// 16 input channels: step t = tap (kh, kw) = (t / 3, t % 3): parity sub-plane and offset of that tap
__device__ __forceinline__ constexpr int tap_sp(int t) { return ((t / 3) == 1 ? 0 : 2) + ((t % 3) == 1 ? 0 : 1); }
__device__ __forceinline__ constexpr int tap_off(int t) { return ((t / 3) == 0 ? -PW : 0) + ((t % 3) == 0 ? -1 : 0); }

template <int CIN, int COUT, int KM, int NS>
__global__ void __launch_bounds__(kThreads, 1)
conv3d_s2_slab_bf16_kernel(const __grid_constant__ CUtensorMap tmap, const Params p) {
  using C = Cfg<CIN, COUT, KM, NS>;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* a_smem = smem;
  unsigned char* b_smem = smem + C::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::A_BYTES + C::B_BYTES);
  uint64_t* full = bars;                 // [NS]  TMA -> MMA
  uint64_t* empty = bars + NS;           // [NS]  MMA -> TMA
  uint64_t* grp_full = bars + 2 * NS;    // [NG]  MMA -> epilogue   (output plane complete)
  uint64_t* grp_free = grp_full + NG;    // [NG]  epilogue -> MMA   (group drained and zeroed)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(grp_free + NG);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ox0 = blockIdx.x * TWO, oy0 = blockIdx.y * C::OROWS;
  const int b = blockIdx.z / p.chunks;
  const int od0 = (blockIdx.z % p.chunks) * p.d_chunk;
  const int od1 = min(p.Do, od0 + p.d_chunk);
  const int nplanes = od1 - od0;          // output planes od0..od1-1 ; input planes z = 2·od0-1 .. 2·od1-1
  const int nslabs = 2 * nplanes + 1;
  const int z0 = 2 * od0 - 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    for (int i = 0; i < NG; ++i) { mbar_init(grp_full + i, 1); mbar_init(grp_free + i, 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4 && lane == 0) tma_prefetch_desc(&tmap);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)C::TALLOC)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {
    const float4* src = reinterpret_cast<const float4*>(p.wimg);
    float4* dst = reinterpret_cast<float4*>(b_smem);
    for (int i = threadIdx.x; i < C::B_BYTES / 16; i += kThreads) dst[i] = __ldg(src + i);
    if (threadIdx.x < 8) reinterpret_cast<float4*>(a_smem + NS * C::STAGE_B)[threadIdx.x] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp < 4) {
    for (int c = 0; c < C::TCOLS; c += 16) tmem_st16_zero(tmem_base + ((uint32_t)(warp * 32) << 16) + c);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 4) {
    // =============================== TMA producer =================================================================
    if (lane == 0) {
      for (int ii = 0; ii < nslabs; ++ii) {
        const int buf = ii % NS;
        mbar_wait(empty + buf, ((ii / NS) & 1) ^ 1);
        mbar_arrive_expect_tx(full + buf, (uint32_t)C::STAGE_B);
        const int z = z0 + ii;
        const uint32_t dst0 = smem_u32(a_smem + (size_t)buf * C::STAGE_B);
#pragma unroll
        for (int part = 0; part < 2; ++part)
#pragma unroll
          for (int sp = 0; sp < 4; ++sp)
#pragma unroll
            for (int ch = 0; ch < C::NCH; ++ch)
              tma_load_5d(dst0 + part * C::PART_B + sp * C::SUB_B + ch * C::PLANE_B, &tmap, full + buf, ch * 8,
                          2 * (ox0 - 1) + (sp & 1), 2 * (oy0 - 1) + (sp >> 1), z, part * p.B + b);
      }
    }
  } else if (warp == 5) {
    // =============================== MMA issuer ==================================================================
    const bool leader = elect_one();
    const uint32_t b_base = smem_u32(b_smem);
    constexpr uint32_t idesc16 = make_idesc_bf16(128, COUT), idesc32 = make_idesc_bf16(128, 2 * COUT);   // one / two kd blocks
    for (int i = 0; i < nslabs; ++i) {
      const bool odd = (i & 1) == 0;                         // z = z0 + i with z0 odd
      const int q_lo = odd ? i / 2 - 1 : (i - 1) / 2;        // odd z: kd = 2 -> plane (z-1)/2 ; even z: kd = 1 -> plane z/2
      const int q_hi = i / 2;                                // odd z: kd = 0 -> plane (z+1)/2 (first touched here)
      if (odd && q_hi < nplanes && q_hi >= NG) mbar_wait(grp_free + (q_hi % NG), ((q_hi / NG) - 1) & 1);
      const int buf = i % NS;
      mbar_wait(full + buf, (i / NS) & 1);
      tc_fence_after();
      const uint32_t a_base = smem_u32(a_smem + (size_t)buf * C::STAGE_B);
      // runs: (TMEM column group, B row offset in 16-byte rows, idesc)
      const bool lo_ok = q_lo >= 0 && q_lo < nplanes, hi_ok = odd && q_hi < nplanes;
      const int g_lo = (q_lo + NG) % NG, g_hi = q_hi % NG;
      const bool fused = odd && lo_ok && hi_ok && g_hi == g_lo + 1;       // two neighbouring groups: one N = 32 MMA
#pragma unroll
      for (int pr = 0; pr < C::NSTEP; ++pr) {
        const int sp = CIN == 8 ? pair_sp(pr) : tap_sp(pr), off = CIN == 8 ? pair_off(pr) : tap_off(pr);
        const int lbs = pair_lbo_sub(pr % 5);
        const uint32_t lbo = CIN == 8 ? (lbs == 0 ? 512u : (uint32_t)(lbs * C::SUB_B)) : (uint32_t)C::PLANE_B;
        const uint32_t a0 = desc_lo(a_base + sp * C::SUB_B, lbo);
        const uint32_t b_odd = desc_lo(b_base + pr * (2 * 2 * COUT * 16), 2 * COUT * 16);
        const uint32_t b_even = desc_lo(b_base + C::ODD_B + pr * (2 * COUT * 16), COUT * 16);
        constexpr uint32_t kLoOff = C::PART_B >> 4, kBLoOff = C::BTERM_B >> 4;
#pragma unroll
        for (int mt = 0; mt < KM; ++mt) {
          const uint32_t a_hi_d = a0 + (uint32_t)((1 + 4 * mt) * PW + off);
          const uint32_t dm = tmem_base + (uint32_t)(mt * (NG * COUT));
          if (leader) {
            if (!odd) {
              const uint32_t d = dm + g_lo * COUT;
              umma_bf16_lohi(d, a_hi_d + kLoOff, b_even, idesc16);                 // x_lo · w_hi
              umma_bf16_lohi(d, a_hi_d, b_even + kBLoOff, idesc16);                // x_hi · w_lo
              umma_bf16_lohi(d, a_hi_d, b_even, idesc16);                          // x_hi · w_hi
            } else if (fused) {
              const uint32_t d = dm + g_lo * COUT;
              umma_bf16_lohi(d, a_hi_d + kLoOff, b_odd, idesc32);
              umma_bf16_lohi(d, a_hi_d, b_odd + kBLoOff, idesc32);
              umma_bf16_lohi(d, a_hi_d, b_odd, idesc32);
            } else {
              if (lo_ok) {                                                          // kd = 2 rows 0..COUT-1
                const uint32_t d = dm + g_lo * COUT;
                umma_bf16_lohi(d, a_hi_d + kLoOff, b_odd, idesc16);
                umma_bf16_lohi(d, a_hi_d, b_odd + kBLoOff, idesc16);
                umma_bf16_lohi(d, a_hi_d, b_odd, idesc16);
              }
              if (hi_ok) {                                                          // kd = 0 rows COUT..2·COUT-1
                const uint32_t d = dm + g_hi * COUT;
                umma_bf16_lohi(d, a_hi_d + kLoOff, b_odd + COUT, idesc16);
                umma_bf16_lohi(d, a_hi_d, b_odd + COUT + kBLoOff, idesc16);
                umma_bf16_lohi(d, a_hi_d, b_odd + COUT, idesc16);
              }
            }
          }
        }
      }
      if (leader) {
        umma_commit(empty + buf);
        if (odd && lo_ok) umma_commit(grp_full + g_lo);       // output plane q_lo has received its last input plane
      }
      __syncwarp();
    }
  } else {
    // =============================== epilogue (128 threads) ======================================================
    // thread (warp w, lane l) of M-tile mt holds output row oy0 + 4·mt + w, column ox0 + l - 1 (l = 0: halo)
    const int gx = ox0 + lane - 1;
    const bool col_ok = lane >= 1 && gx < p.Wo;
    const int64_t part_stride = (int64_t)p.B * p.Do * p.Ho * p.Wo * COUT;
    float bs[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) bs[c] = p.bias ? __ldg(p.bias + c) : 0.f;
    for (int q = 0; q < nplanes; ++q) {
      const int g = q % NG;
      mbar_wait(grp_full + g, (q / NG) & 1);
      tc_fence_after();
      uint32_t vr[KM][COUT];
#pragma unroll
      for (int mt = 0; mt < KM; ++mt)
#pragma unroll
        for (int c0 = 0; c0 < COUT; c0 += 16)
          tmem_ld16_nowait(tmem_base + ((uint32_t)(warp * 32) << 16) + mt * (NG * COUT) + g * COUT + c0, &vr[mt][c0]);
      tmem_ld_wait();
#pragma unroll
      for (int mt = 0; mt < KM; ++mt)
#pragma unroll
        for (int c0 = 0; c0 < COUT; c0 += 16)
          tmem_st16_zero(tmem_base + ((uint32_t)(warp * 32) << 16) + mt * (NG * COUT) + g * COUT + c0);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(grp_free + g);
      const int od = od0 + q;
#pragma unroll
      for (int mt = 0; mt < KM; ++mt) {
        const int gy = oy0 + 4 * mt + warp;
        if (!col_ok || gy >= p.Ho) continue;
        const int64_t o = ((((int64_t)b * p.Do + od) * p.Ho + gy) * p.Wo + gx) * COUT;
#pragma unroll
        for (int c = 0; c < COUT; c += 4) {
          float4 r = make_float4(__uint_as_float(vr[mt][c]) + bs[c], __uint_as_float(vr[mt][c + 1]) + bs[c + 1],
                                 __uint_as_float(vr[mt][c + 2]) + bs[c + 2], __uint_as_float(vr[mt][c + 3]) + bs[c + 3]);
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TALLOC)
                 : "memory");
  }
}

// Weight image: torch Conv3d weight [Cout][Cin][3][3][3] x BN scale ->
//   img[term][ odd: step x chunk(2) x row(2·Cout: kd=2 couts | kd=0 couts) x e(8) | even: step x chunk(2) x row(Cout: kd=1) x e(8) ]
//   Cin = 8: step = tap pair, chunk = tap within the pair, channel = e;  Cin = 16: step = tap, channel = chunk*8 + e
//   term 0 = bf16(w·scale), term 1 = bf16(w·scale - term 0)
__global__ void pack_s2_weight_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                      __nv_bfloat16* __restrict__ img, int Cout, int Cin) {
  const int nstep = Cin == 8 ? 5 : 9;
  const int odd_e = nstep * 2 * 2 * Cout * 8, even_e = nstep * 2 * Cout * 8, per_term = odd_e + even_e;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per_term) return;
  // Cin = 8: taps (kh, kw) of the two K-chunks of each pair; kh = -1: zero weights
  const int KH[5][2] = {{0, 0}, {1, 1}, {2, 2}, {0, 2}, {1, -1}};
  const int KW[5][2] = {{1, 2}, {1, 2}, {1, 2}, {0, 0}, {0, 0}};
  const bool odd = i < odd_e;
  const int rows = odd ? 2 * Cout : Cout;
  int r = odd ? i : i - odd_e;
  const int e = r & 7; r >>= 3;
  const int row = r % rows; r /= rows;
  const int chunk = r & 1;
  const int st = r >> 1;
  const int kd = odd ? (row < Cout ? 2 : 0) : 1, n = row % Cout;
  int kh, kw, ci;
  if (Cin == 8) { kh = KH[st][chunk]; kw = KW[st][chunk]; ci = e; }
  else { kh = st / 3; kw = st % 3; ci = chunk * 8 + e; }
  float v = 0.f;
  if (kh >= 0) {
    v = w[((int64_t)n * Cin + ci) * 27 + (kd * 9 + kh * 3 + kw)];
    if (scale) v *= scale[n];
  }
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  img[i] = hi;
  img[per_term + i] = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// 5-D map over split-bf16 activations [2·B][D][H][W][C], traversal stride 2 along W and H: box {8, 2·32, 2·rows, 1, 1}
static int make_tmap_s2(CUtensorMap* map, const void* x, int B, int D, int H, int W, int CIN, int rows) {
  EncodeTiledFn fn = encode_fn();
  RCMVS_REQUIRE(fn != nullptr, RCMVS_ERR_CUDA, "rcmvs_conv3d_s2_slab_fwd: cuTensorMapEncodeTiled not available");
  const cuuint64_t gdim[5] = {(cuuint64_t)CIN, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)2 * B};
  const cuuint64_t gstr[4] = {(cuuint64_t)CIN * 2, (cuuint64_t)W * CIN * 2, (cuuint64_t)H * W * CIN * 2,
                              (cuuint64_t)D * H * W * CIN * 2};
  const cuuint32_t box[5] = {8, (cuuint32_t)(2 * PW), (cuuint32_t)(2 * rows), 1, 1};
  const cuuint32_t estr[5] = {1, 2, 2, 1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RCMVS_REQUIRE(r == CUDA_SUCCESS, RCMVS_ERR_CUDA, "rcmvs_conv3d_s2_slab_fwd: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return RCMVS_OK;
}

template <int CIN, int COUT, int KM, int NS>
static int launch(const void* x, const Params& p, cudaStream_t st) {
  using C = Cfg<CIN, COUT, KM, NS>;
  static_assert(C::SMEM <= 232448, "shared memory budget exceeded");
  auto kern = conv3d_s2_slab_bf16_kernel<CIN, COUT, KM, NS>;
  static SmemAttrOnce smem_attr;
  {
    cudaError_t e = smem_attr.ensure(kern, (int)C::SMEM);
    RCMVS_REQUIRE(e == cudaSuccess, RCMVS_ERR_CUDA, "rcmvs_conv3d_s2_slab_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  CUtensorMap map;
  const int rc = make_tmap_s2(&map, x, p.B, p.D, p.H, p.W, CIN, C::ROWS_S);
  if (rc != RCMVS_OK) return rc;
  dim3 grid(ceil_div(p.Wo, TWO), ceil_div(p.Ho, C::OROWS), p.B * p.chunks);
  RCMVS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, RCMVS_ERR_UNSUPPORTED, "rcmvs_conv3d_s2_slab_fwd: grid too large");
  kern<<<grid, kThreads, C::SMEM, st>>>(map, p);
  RCMVS_CHECK_LAUNCH("rcmvs_conv3d_s2_slab_fwd");
  return RCMVS_OK;
}

// M-tiles per CTA and output-plane chunks: one CTA per SM, minimise waves x (input planes per CTA) x KM
static void plan(int B, int Do, int Ho, int Wo, int km_lo, int km_hi, int& km, int& chunks) {
  long best = -1;
  for (int k = km_lo; k <= km_hi; ++k) {
    const int tiles = ceil_div(Wo, TWO) * ceil_div(Ho, 4 * k) * B;
    for (int c = 1; c <= Do; ++c) {
      const int dc = ceil_div(Do, c);
      const int cc = ceil_div(Do, dc);
      const long cost = (long)ceil_div(tiles * cc, 148) * (2 * dc + 1) * k;
      if (best < 0 || cost < best) { best = cost; km = k; chunks = cc; }
    }
  }
}

static bool supported(int Cin, int Cout) { return (Cin == 8 && Cout == 16) || (Cin == 16 && Cout == 32); }

}  // namespace slab2
}  // namespace rcmvs

extern "C" int64_t rcmvs_conv3d_s2_slab_weight_elems(int Cin, int Cout) {
  const int nstep = Cin == 8 ? 5 : 9;
  return (int64_t)2 * (nstep * 2 * 2 * Cout * 8 + nstep * 2 * Cout * 8);
}

extern "C" int rcmvs_conv3d_s2_slab_pack(const float* w, const float* scale, void* img, int Cout, int Cin,
                                         rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(w && img, RCMVS_ERR_BAD_ARG, "rcmvs_conv3d_s2_slab_pack: null pointer");
  RCMVS_REQUIRE(slab2::supported(Cin, Cout), RCMVS_ERR_UNSUPPORTED, "rcmvs_conv3d_s2_slab_pack: unsupported Cin=%d Cout=%d", Cin,
                Cout);
  const int per_term = (int)(rcmvs_conv3d_s2_slab_weight_elems(Cin, Cout) / 2);
  slab2::pack_s2_weight_kernel<<<ceil_div(per_term, 256), 256, 0, as_stream(stream)>>>(
      w, scale, reinterpret_cast<__nv_bfloat16*>(img), Cout, Cin);
  RCMVS_CHECK_LAUNCH("rcmvs_conv3d_s2_slab_pack");
  return RCMVS_OK;
}

extern "C" int rcmvs_conv3d_s2_slab_fwd(const void* x_split, const void* wimg, const float* bias, float* y, void* y_split,
                                        int B, int D, int H, int W, int Cin, int Cout, int relu, rcmvs_stream_t stream) {
  using namespace rcmvs;
  using namespace rcmvs::slab2;
  RCMVS_REQUIRE(x_split && wimg && (y || y_split), RCMVS_ERR_BAD_ARG, "rcmvs_conv3d_s2_slab_fwd: null pointer");
  RCMVS_REQUIRE(slab2::supported(Cin, Cout), RCMVS_ERR_UNSUPPORTED, "rcmvs_conv3d_s2_slab_fwd: unsupported Cin=%d Cout=%d", Cin,
                Cout);
  RCMVS_REQUIRE(B > 0 && D >= 2 && H >= 2 && W >= 2 && D % 2 == 0 && H % 2 == 0 && W % 2 == 0, RCMVS_ERR_UNSUPPORTED,
                "rcmvs_conv3d_s2_slab_fwd: D, H, W must be even (got %d, %d, %d)", D, H, W);
  RCMVS_REQUIRE((reinterpret_cast<uintptr_t>(x_split) & 15) == 0, RCMVS_ERR_BAD_ARG,
                "rcmvs_conv3d_s2_slab_fwd: x_split must be 16-byte aligned (TMA)");
  Params p;
  p.wimg = reinterpret_cast<const __nv_bfloat16*>(wimg);
  p.bias = bias; p.y = y; p.ys = reinterpret_cast<__nv_bfloat16*>(y_split);
  p.B = B; p.D = D; p.H = H; p.W = W; p.Do = D / 2; p.Ho = H / 2; p.Wo = W / 2; p.relu = relu;
  int km = 3, chunks = 1;
  cudaStream_t st = as_stream(stream);
  if (Cin == 8) {
    plan(B, p.Do, p.Ho, p.Wo, 3, 4, km, chunks);
    p.d_chunk = ceil_div(p.Do, chunks);
    p.chunks = ceil_div(p.Do, p.d_chunk);
    return km == 4 ? launch<8, 16, 4, 3>(x_split, p, st) : launch<8, 16, 3, 4>(x_split, p, st);
  }
  plan(B, p.Do, p.Ho, p.Wo, 2, 2, km, chunks);
  p.d_chunk = ceil_div(p.Do, chunks);
  p.chunks = ceil_div(p.Do, p.d_chunk);
  return launch<16, 32, 2, 2>(x_split, p, st);
}
