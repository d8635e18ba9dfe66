// a3 CostRegNet conv4 (32 -> 32 channels, 3x3x3, stride 1, quarter resolution; reference models/modules.py:118-163 Conv3d,
// :478 conv4) on tcgen05 with TMA-fed split-bf16 operands ("bf16x3", see conv3d_slab_bf16.cu).
//
// With 32 output channels the (kd, kh, cout) fold of conv3d_slab_bf16.cu does not fit one MMA (N = 288), so this member of
// the family keeps only kd in N (three blocks of 32 columns = the accumulators of three output planes in a TMEM ring)
// and walks all nine (kh, kw) taps as start-address offsets of the A descriptor over a flat M = 128 run of OUTPUT
// positions (4 rows x 32 columns, columns 0 and 31 = halo): tap (kh, kw) reads position + (kh-1)·32 + (kw-1) of the
// zero-padded input slab.  Everything is summed inside TMEM; the epilogue adds the BN shift, applies ReLU, adds the skip
// tensor and stores fp32 and / or split-bf16 — no shared-memory exchange.
//
// Warp roles: warps 0-3 epilogue, warp 4 TMA producer, warp 5 MMA issuer.
#include "slab_common.cuh"

namespace rcmvs {
namespace slabo {

using namespace rcmvs::slab;

constexpr int TWO = 30;                                            // real output columns per tile (columns 0, 31 = halo)
constexpr int NG = 4;                                              // TMEM ring: output planes in flight

template <int CIN, int COUT, int KM, int NS>
struct Cfg {
  static constexpr int KS = CIN / 16, NCH = CIN / 8;
  static constexpr int ROWS_S = 4 * KM + 2;                        // slab rows (1 halo row above, 1 below)
  static constexpr int OROWS = 4 * KM;
  static constexpr int PLANE_B = ROWS_S * PW * 16;                 // one 8-channel chunk plane
  static constexpr int SLAB_B = NCH * PLANE_B;                     // one part (hi or lo)
  static constexpr int STAGE_B = 2 * SLAB_B;
  static constexpr int FRONT = 128;                                // the (-1, -1) tap of the first chunk plane starts 16 B early
  static constexpr int A_BYTES = FRONT + NS * STAGE_B + 128;       // + the (+1, +1) overrun of the last M-tile
  static constexpr int BBLK_B = 2 * 3 * COUT * 16;                 // one (tap, K step): 2 chunks x 3·COUT rows x 16 B
  static constexpr int BTERM_B = 9 * KS * BBLK_B;
  static constexpr int B_BYTES = 2 * BTERM_B;                      // hi, lo
  static constexpr int TCOLS = NG * KM * COUT;
  static constexpr int TALLOC = TCOLS <= 128 ? 128 : (TCOLS <= 256 ? 256 : 512);
  static constexpr size_t SMEM = (size_t)A_BYTES + B_BYTES + 512;
  static_assert(CIN % 16 == 0 && COUT % 16 == 0 && 3 * COUT <= 256 && TCOLS <= 512 && PLANE_B % 128 == 0, "configuration");
};

struct Params {
  const __nv_bfloat16* wimg;
  const float* bias;           // [Cout] or null
  const float* skip;           // [B][D][H][W][Cout] fp32 or null
  float* y;                    // [B][D][H][W][Cout] fp32 or null
  __nv_bfloat16* ys;           // [2][B][D][H][W][Cout] split-bf16 or null
  int B, D, H, W;
  int d_chunk, chunks;
  int relu;
};

template <int CIN, int COUT, int KM, int NS>
__global__ void __launch_bounds__(kThreads, 1)
conv3d_o_slab_bf16_kernel(const __grid_constant__ CUtensorMap tmap, const Params p) {
  using C = Cfg<CIN, COUT, KM, NS>;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* a_smem = smem + C::FRONT;
  unsigned char* b_smem = smem + C::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::A_BYTES + C::B_BYTES);
  uint64_t* full = bars;                 // [NS]  TMA -> MMA
  uint64_t* empty = bars + NS;           // [NS]  MMA -> TMA
  uint64_t* grp_full = bars + 2 * NS;    // [NG]  MMA -> epilogue   (plane complete)
  uint64_t* grp_free = grp_full + NG;    // [NG]  epilogue -> MMA   (group drained and zeroed)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(grp_free + NG);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x0 = blockIdx.x * TWO, y0 = blockIdx.y * C::OROWS;
  const int b = blockIdx.z / p.chunks;
  const int d0 = (blockIdx.z % p.chunks) * p.d_chunk;
  const int d1 = min(p.D, d0 + p.d_chunk);
  const int nplanes = d1 - d0;           // output planes d0..d1-1 ; input slabs z = d0-1 .. d1  (nplanes + 2)
  const int nslabs = nplanes + 2;

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
    if (threadIdx.x < 8) {
      reinterpret_cast<float4*>(smem)[threadIdx.x] = make_float4(0.f, 0.f, 0.f, 0.f);
      reinterpret_cast<float4*>(a_smem + NS * C::STAGE_B)[threadIdx.x] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp < 4) {
    // all accumulators start at zero: every MMA accumulates, a group is re-zeroed by the epilogue after it is drained
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
        const int z = d0 - 1 + ii;
        const uint32_t dst0 = smem_u32(a_smem + (size_t)buf * C::STAGE_B);
#pragma unroll
        for (int part = 0; part < 2; ++part)
#pragma unroll
          for (int ch = 0; ch < C::NCH; ++ch)
            tma_load_5d(dst0 + part * C::SLAB_B + ch * C::PLANE_B, &tmap, full + buf, ch * 8, x0 - 1, y0 - 1, z,
                        part * p.B + b);
      }
    }
  } else if (warp == 5) {
    // =============================== MMA issuer ==================================================================
    const bool leader = elect_one();
    const uint32_t b_lo0 = desc_lo(smem_u32(b_smem), 3 * COUT * 16);
    for (int i = 0; i < nslabs; ++i) {
      // slab i = input plane z = d0-1+i contributes to output planes (relative) q = i-2+j, block j = 0,1,2 (kd = 2-j)
      const int jlo = max(0, 2 - i), jhi = min(2, nplanes + 1 - i);
      const int qnew = i;                                                  // plane first touched by this slab (j = 2)
      if (qnew < nplanes && qnew >= NG) mbar_wait(grp_free + (qnew % NG), ((qnew / NG) - 1) & 1);
      const int buf = i % NS;
      mbar_wait(full + buf, (i / NS) & 1);
      tc_fence_after();
      const uint32_t a_lo0 = desc_lo(smem_u32(a_smem + (size_t)buf * C::STAGE_B), C::PLANE_B);
      const int g0 = (i - 2 + jlo + NG) % NG;
      const int n_all = jhi - jlo + 1;
      const int n0 = min(n_all, NG - g0), n1 = n_all - n0;                 // second run wraps to group 0
      const uint32_t d0c = tmem_base + g0 * COUT, d1c = tmem_base;
      const uint32_t idesc0 = make_idesc_bf16(128, n0 * COUT), idesc1 = make_idesc_bf16(128, (n1 > 0 ? n1 : 1) * COUT);
      const uint32_t boff0 = (uint32_t)(jlo * COUT), boff1 = (uint32_t)((jlo + n0) * COUT);   // in 16-byte rows
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int off = (tap / 3 - 1) * PW + (tap % 3 - 1);
#pragma unroll
        for (int ks = 0; ks < C::KS; ++ks) {
#pragma unroll
          for (int mt = 0; mt < KM; ++mt) {
            constexpr uint32_t kLoOff = C::SLAB_B >> 4, kBLoOff = C::BTERM_B >> 4;
            const uint32_t a_hi_d = a_lo0 + (uint32_t)((ks * 2 * C::PLANE_B) >> 4) + (uint32_t)((1 + 4 * mt) * PW + off);
            const uint32_t b_blk_d = b_lo0 + (uint32_t)(((tap * C::KS + ks) * C::BBLK_B) >> 4);
            const uint32_t dm = (uint32_t)(mt * (NG * COUT));
            if (leader) {
              umma_bf16_lohi(d0c + dm, a_hi_d + kLoOff, b_blk_d + boff0, idesc0);            // x_lo · w_hi
              umma_bf16_lohi(d0c + dm, a_hi_d, b_blk_d + boff0 + kBLoOff, idesc0);           // x_hi · w_lo
              umma_bf16_lohi(d0c + dm, a_hi_d, b_blk_d + boff0, idesc0);                     // x_hi · w_hi
              if (n1 > 0) {
                umma_bf16_lohi(d1c + dm, a_hi_d + kLoOff, b_blk_d + boff1, idesc1);
                umma_bf16_lohi(d1c + dm, a_hi_d, b_blk_d + boff1 + kBLoOff, idesc1);
                umma_bf16_lohi(d1c + dm, a_hi_d, b_blk_d + boff1, idesc1);
              }
            }
          }
        }
      }
      if (leader) {
        umma_commit(empty + buf);                                   // slab consumed once these MMAs retire
        if (i >= 2) umma_commit(grp_full + ((i - 2) % NG));         // output plane q = i-2 is complete
      }
      __syncwarp();
    }
  } else {
    // =============================== epilogue (128 threads) ======================================================
    // thread (warp w, lane l) of M-tile mt holds output row y0 + 4·mt + w, column x0 + l - 1 (l = 0, 31: halo)
    const int gx = x0 + lane - 1;
    const bool col_ok = lane >= 1 && lane <= TWO && gx < p.W;
    const int64_t part_stride = (int64_t)p.B * p.D * p.H * p.W * COUT;
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
      const int d = d0 + q;
#pragma unroll
      for (int mt = 0; mt < KM; ++mt) {
        const int gy = y0 + 4 * mt + warp;
        if (!col_ok || gy >= p.H) continue;
        const int64_t o = ((((int64_t)b * p.D + d) * p.H + gy) * p.W + gx) * COUT;
#pragma unroll
        for (int c = 0; c < COUT; c += 4) {
          float4 r = make_float4(__uint_as_float(vr[mt][c]), __uint_as_float(vr[mt][c + 1]), __uint_as_float(vr[mt][c + 2]),
                                 __uint_as_float(vr[mt][c + 3]));
          if (p.bias) {
            const float4 t4 = ldg4(p.bias + c);
            r.x += t4.x; r.y += t4.y; r.z += t4.z; r.w += t4.w;
          }
          if (p.relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
          if (p.skip) {
            const float4 t4 = ldg4(p.skip + o + c);
            r.x += t4.x; r.y += t4.y; r.z += t4.z; r.w += t4.w;
          }
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

// Weight image: torch Conv3d weight [Cout][Cin][27] x BN scale ->
//   img[term][tap = kh*3+kw][ks][chunk][row = jblk*COUT + n][e],  kd = 2 - jblk, channel = ks*16 + chunk*8 + e
__global__ void pack_o_slab_weight_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                          __nv_bfloat16* __restrict__ img, int Cout, int Cin) {
  const int KS = Cin / 16;
  const int per_term = 9 * KS * 2 * 3 * Cout * 8;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per_term) return;
  const int e = i & 7;
  int r = i >> 3;
  const int row = r % (3 * Cout); r /= (3 * Cout);
  const int chunk = r & 1; r >>= 1;
  const int ks = r % KS;
  const int tap = r / KS;
  const int jblk = row / Cout, n = row % Cout;
  const int kd = 2 - jblk, ci = ks * 16 + chunk * 8 + e;
  float v = w[((int64_t)n * Cin + ci) * 27 + (kd * 9 + tap)];
  if (scale) v *= scale[n];
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  img[i] = hi;
  img[per_term + i] = __float2bfloat16_rn(v - __bfloat162float(hi));
}

template <int CIN, int COUT, int KM, int NS>
static int launch(const void* x, const Params& p, cudaStream_t st) {
  using C = Cfg<CIN, COUT, KM, NS>;
  static_assert(C::SMEM <= 232448, "shared memory budget exceeded");
  auto kern = conv3d_o_slab_bf16_kernel<CIN, COUT, KM, NS>;
  static SmemAttrOnce smem_attr;
  {
    cudaError_t e = smem_attr.ensure(kern, (int)C::SMEM);
    RCMVS_REQUIRE(e == cudaSuccess, RCMVS_ERR_CUDA, "rcmvs_conv3d_o_slab_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  CUtensorMap map;
  const int rc = make_tmap(&map, x, p.B, p.D, p.H, p.W, CIN, C::ROWS_S);
  if (rc != RCMVS_OK) return rc;
  dim3 grid(ceil_div(p.W, TWO), ceil_div(p.H, C::OROWS), p.B * p.chunks);
  RCMVS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, RCMVS_ERR_UNSUPPORTED, "rcmvs_conv3d_o_slab_fwd: grid too large");
  kern<<<grid, kThreads, C::SMEM, st>>>(map, p);
  RCMVS_CHECK_LAUNCH("rcmvs_conv3d_o_slab_fwd");
  return RCMVS_OK;
}

static bool supported(int Cin, int Cout) { return Cin == 32 && Cout == 32; }

}  // namespace slabo
}  // namespace rcmvs

extern "C" int64_t rcmvs_conv3d_o_slab_weight_elems(int Cin, int Cout) { return (int64_t)2 * 9 * (Cin / 16) * 2 * 3 * Cout * 8; }

extern "C" int rcmvs_conv3d_o_slab_pack(const float* w, const float* scale, void* img, int Cout, int Cin,
                                        rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(w && img, RCMVS_ERR_BAD_ARG, "rcmvs_conv3d_o_slab_pack: null pointer");
  RCMVS_REQUIRE(slabo::supported(Cin, Cout), RCMVS_ERR_UNSUPPORTED, "rcmvs_conv3d_o_slab_pack: unsupported Cin=%d Cout=%d", Cin,
                Cout);
  const int per_term = (int)(rcmvs_conv3d_o_slab_weight_elems(Cin, Cout) / 2);
  slabo::pack_o_slab_weight_kernel<<<ceil_div(per_term, 256), 256, 0, as_stream(stream)>>>(
      w, scale, reinterpret_cast<__nv_bfloat16*>(img), Cout, Cin);
  RCMVS_CHECK_LAUNCH("rcmvs_conv3d_o_slab_pack");
  return RCMVS_OK;
}

extern "C" int rcmvs_conv3d_o_slab_fwd(const void* x_split, const void* wimg, const float* bias, const float* skip, float* y,
                                       void* y_split, int B, int D, int H, int W, int Cin, int Cout, int relu,
                                       rcmvs_stream_t stream) {
  using namespace rcmvs;
  using namespace rcmvs::slabo;
  RCMVS_REQUIRE(x_split && wimg && (y || y_split), RCMVS_ERR_BAD_ARG, "rcmvs_conv3d_o_slab_fwd: null pointer");
  RCMVS_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, RCMVS_ERR_BAD_ARG, "rcmvs_conv3d_o_slab_fwd: bad dims");
  RCMVS_REQUIRE(slabo::supported(Cin, Cout), RCMVS_ERR_UNSUPPORTED, "rcmvs_conv3d_o_slab_fwd: unsupported Cin=%d Cout=%d", Cin,
                Cout);
  RCMVS_REQUIRE((reinterpret_cast<uintptr_t>(x_split) & 15) == 0, RCMVS_ERR_BAD_ARG,
                "rcmvs_conv3d_o_slab_fwd: x_split must be 16-byte aligned (TMA)");
  Params p;
  p.wimg = reinterpret_cast<const __nv_bfloat16*>(wimg);
  p.bias = bias; p.skip = skip; p.y = y; p.ys = reinterpret_cast<__nv_bfloat16*>(y_split);
  p.B = B; p.D = D; p.H = H; p.W = W; p.relu = relu;
  // depth chunks: one CTA per SM; minimise waves(tiles x chunks / 148) x slabs per CTA (planes + 2 halo slabs)
  constexpr int KM = 2;
  const int tiles = ceil_div(W, TWO) * ceil_div(H, 4 * KM) * B;
  int chunks = 1;
  {
    long best = -1;
    for (int c = 1; c <= D; ++c) {
      const int dc = ceil_div(D, c);
      const int cc = ceil_div(D, dc);
      const long cost = (long)ceil_div(tiles * cc, 148) * (dc + 2);
      if (best < 0 || cost < best) { best = cost; chunks = cc; }
    }
  }
  p.d_chunk = ceil_div(D, chunks);
  p.chunks = ceil_div(D, p.d_chunk);
  return launch<32, 32, KM, 2>(x_split, p, as_stream(stream));
}
