// a3 CostRegNet conv11 (ConvTranspose3d 16 -> 8, k3, stride 2, pad 1, output_padding 1 + BN + ReLU + skip; reference
// models/modules.py:166-210 Deconv3d, :488 conv11, :499 the skip add) on tcgen05 with TMA-fed split-bf16 operands
// ("bf16x3", see conv3d_slab_bf16.cu).
//
// out[2i + a] per dimension:  a = 0 <- (k = 1, i);  a = 1 <- (k = 2, i) + (k = 0, i + 1).  Every output parity class
// (a, b, c) is a small stride-1 convolution over the INPUT grid, so the flat-run trick applies in input space: M = 128
// input positions (4 rows x 32 columns, column 31 = halo for dx = 1), the 8 input shifts (dz, dy, dx) in {0,1}^3 are
// descriptor start offsets (+dy·32 + dx positions, slab iz + dz) and N = 8 classes x 8 couts holds all parity classes
// (weights of classes that do not use a shift are zero rows of B).  K = 16 = the 16 input channels in ONE step, so an
// input M-tile costs 8 shifts x 3 terms = 24 MMAs (the round-1 3xTF32 kernel: 48) and the slabs arrive by TMA (zero
// fill = the i + 1 neighbours beyond the volume) instead of through loader warps.  The kernel is bound by its epilogue
// (each input voxel scatters 8 x 32 B of output and reads 8 x 32 B of skip tensor), so it is sized for TWO CTAs per SM
// (107 KB of shared memory, 256 TMEM columns each): eight epilogue warps per SM keep twice the loads in flight.
//
// conv9 (32 -> 16, quarter -> half resolution) runs on the same kernel: two K steps, N = 128, the shifts with dz = 1 only
// address the 64 columns of the classes that use them, the epilogue walks the classes in two passes, and the result is
// written pre-split for conv11 (one CTA per SM: its weight image alone is 96 KB).
//
// Warp roles: warps 0-3 epilogue, warp 4 TMA producer, warp 5 MMA issuer.
#include "slab_common.cuh"

namespace rcmvs {
namespace slabt {

using namespace rcmvs::slab;

constexpr int TWI = 31;                                            // input columns per tile (column 31 = halo for dx = 1)

template <int CIN, int COUT, int KM, int NS>
struct Cfg {
  static constexpr int N = 8 * COUT;                               // 8 output parity classes x COUT
  static constexpr int KS = CIN / 16, NCH = CIN / 8;
  static constexpr int ROWS = 4 * KM;                              // input rows per CTA tile (all produce outputs)
  static constexpr int PLANE_B = (ROWS + 1) * PW * 16;             // one 8-channel chunk plane (+1 halo row for dy = 1)
  static constexpr int SLAB_B = NCH * PLANE_B;                     // one part (hi or lo)
  static constexpr int STAGE_B = 2 * SLAB_B;
  static constexpr int A_BYTES = NS * STAGE_B + 128;               // + overrun of the (dy, dx) = (1, 1) shift
  // B blocks: the four dz = 0 shifts carry all 8 classes (N rows), the four dz = 1 shifts only the classes with a = 1
  // (the upper N/2 rows / TMEM columns)
  static constexpr int BBLK0_B = 2 * N * 16, BBLK1_B = N * 16;     // per (shift, K step): 2 chunks x rows x 16 B
  static constexpr int BTERM_B = 4 * KS * (BBLK0_B + BBLK1_B);
  static constexpr int B_BYTES = 2 * BTERM_B;                      // hi, lo
  static constexpr int TCOLS = 2 * KM * N;                         // double-buffered accumulators
  static constexpr int CPP = 64 / COUT;                            // classes per epilogue pass (64 accumulator columns)
  static constexpr size_t SMEM = (size_t)A_BYTES + B_BYTES + 256;
  static_assert(PLANE_B % 128 == 0, "TMA destination alignment");
  static_assert(TCOLS == 256 || TCOLS == 512, "TMEM allocation is a power of two");
  static_assert(CIN % 16 == 0 && (COUT == 8 || COUT == 16), "configuration");
};

struct Params {
  const __nv_bfloat16* wimg;
  const float* bias;           // [COUT] or null
  const float* skip;           // [B][2D][2H][2W][COUT] or null
  float* y;                    // [B][2D][2H][2W][COUT] fp32 or null
  __nv_bfloat16* ys;           // [2][B][2D][2H][2W][COUT] split-bf16 or null
  int B, D, H, W;              // input dims
  int d_chunk, chunks;
  int relu;
};

template <int CIN, int COUT, int KM, int NS, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
convT3d_slab_bf16_kernel(const __grid_constant__ CUtensorMap tmap, const Params p) {
  using C = Cfg<CIN, COUT, KM, NS>;
  constexpr int N = C::N;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* a_smem = smem;
  unsigned char* b_smem = smem + C::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::A_BYTES + C::B_BYTES);
  uint64_t* full = bars;                 // [NS]  TMA -> MMA
  uint64_t* empty = bars + NS;           // [NS]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * NS;    // [2]   MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;    // [2]   epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x0 = blockIdx.x * TWI, y0 = blockIdx.y * C::ROWS;
  const int b = blockIdx.z / p.chunks;
  const int z0 = (blockIdx.z % p.chunks) * p.d_chunk;
  const int z1 = min(p.D, z0 + p.d_chunk);
  const int nsteps = z1 - z0;            // input planes z0..z1-1 ; slabs z0 .. z1  (nsteps + 1)
  const int nslabs = nsteps + 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, 128); }
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
    if (threadIdx.x < 8) reinterpret_cast<float4*>(a_smem + NS * C::STAGE_B)[threadIdx.x] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // =============================== TMA producer =================================================================
    // slab ii = input plane z0 + ii: 2 parts x 2 chunk boxes {8 ch, 32 cols, ROWS+1 rows, 1 plane} at (x0, y0, z); columns /
    // rows / planes beyond the volume are zero-filled (the i + 1 neighbours of the last voxels contribute nothing)
    if (lane == 0) {
      for (int ii = 0; ii < nslabs; ++ii) {
        const int buf = ii % NS;
        mbar_wait(empty + buf, ((ii / NS) & 1) ^ 1);
        mbar_arrive_expect_tx(full + buf, (uint32_t)C::STAGE_B);
        const uint32_t dst0 = smem_u32(a_smem + (size_t)buf * C::STAGE_B);
#pragma unroll
        for (int part = 0; part < 2; ++part)
#pragma unroll
          for (int ch = 0; ch < C::NCH; ++ch)
            tma_load_5d(dst0 + part * C::SLAB_B + ch * C::PLANE_B, &tmap, full + buf, ch * 8, x0, y0, z0 + ii, part * p.B + b);
      }
    }
  } else if (warp == 5) {
    // =============================== MMA issuer ==================================================================
    const bool leader = elect_one();
    const uint32_t b_base = smem_u32(b_smem);
    constexpr uint32_t idesc_full = make_idesc_bf16(128, N), idesc_half = make_idesc_bf16(128, N / 2);
    for (int s = 0; s < nsteps; ++s) {
      const int stage = s & 1;
      mbar_wait(acc_empty + stage, ((s >> 1) & 1) ^ 1);
      tc_fence_after();
#pragma unroll
      for (int dz = 0; dz < 2; ++dz) {
        const int i = s + dz;
        const int buf = i % NS;
        mbar_wait(full + buf, (i / NS) & 1);       // (slab s + 1 is waited for again as slab (s+1) + 0: already complete)
        tc_fence_after();
        const uint32_t a_lo0 = desc_lo(smem_u32(a_smem + (size_t)buf * C::STAGE_B), C::PLANE_B);
        // dz = 0: all 8 classes (N columns); dz = 1: only the classes with a = 1 = the upper N/2 columns / rows
        const uint32_t idesc = dz == 0 ? idesc_full : idesc_half;
        const uint32_t col_off = dz == 0 ? 0u : (uint32_t)(N / 2);
        const uint32_t blk_b = dz == 0 ? (uint32_t)C::BBLK0_B : (uint32_t)C::BBLK1_B;
        const uint32_t b_lo0 = desc_lo(b_base + (dz == 0 ? 0 : 4 * C::KS * C::BBLK0_B), blk_b / 2);
#pragma unroll
        for (int dyx = 0; dyx < 4; ++dyx) {
#pragma unroll
          for (int ks = 0; ks < C::KS; ++ks) {
#pragma unroll
            for (int mt = 0; mt < KM; ++mt) {
              constexpr uint32_t kLoOff = C::SLAB_B >> 4, kBLoOff = C::BTERM_B >> 4;
              const uint32_t a_d = a_lo0 + (uint32_t)((ks * 2 * C::PLANE_B) >> 4) + (uint32_t)(mt * 128 + (dyx >> 1) * PW + (dyx & 1));
              const uint32_t b_d = b_lo0 + (uint32_t)(((dyx * C::KS + ks) * blk_b) >> 4);
              const uint32_t d = tmem_base + (uint32_t)(stage * KM * N + mt * N) + col_off;
              const bool first = dz == 0 && dyx == 0 && ks == 0;
              if (leader) {
                umma_bf16_lohi_acc(d, a_d + kLoOff, b_d, idesc, first ? 0u : 1u);     // x_lo · w_hi (first MMA overwrites)
                umma_bf16_lohi_acc(d, a_d, b_d + kBLoOff, idesc, 1u);                 // x_hi · w_lo
                umma_bf16_lohi_acc(d, a_d, b_d, idesc, 1u);                           // x_hi · w_hi
              }
            }
          }
        }
        if (dz == 0 && leader) umma_commit(empty + buf);                 // slab iz is not needed by later planes
      }
      if (leader) umma_commit(acc_full + stage);
      __syncwarp();
    }
    // the last slab (z1) is only ever used as a dz = 1 neighbour: nobody waits on its `empty`, nothing to release
  } else {
    // =============================== epilogue (128 threads) ======================================================
    const int gx = x0 + lane;
    const bool col_ok = lane < TWI && gx < p.W;
    const int Do = 2 * p.D, Ho = 2 * p.H, Wo = 2 * p.W;
    // every thread needs 8 x 32 B of the skip tensor per M-tile: requested BEFORE the wait for the accumulators / right
    // after the previous M-tile's stores, not one load per class in front of its add (latency-bound otherwise)
    float4 sk[16];
    const bool use_skip = p.skip != nullptr;
    const int64_t part_stride = (int64_t)p.B * Do * Ho * Wo * COUT;
    constexpr int CPP = C::CPP, NPASS = 8 / CPP, Q = COUT / 4;      // float4 groups per class
    // one pass = CPP classes = 64 accumulator columns = 16 float4 of skip tensor
    auto load_skip = [&](int iz, int mt, int pass) {
      const int gy = y0 + mt * 4 + warp;
      const bool ok = use_skip && col_ok && gy < p.H;
#pragma unroll
      for (int cc = 0; cc < CPP; ++cc) {
        const int cls = pass * CPP + cc;
        const int oz = 2 * iz + (cls >> 2), oy = 2 * gy + ((cls >> 1) & 1), ox = 2 * gx + (cls & 1);
        const int64_t o = ((((int64_t)b * Do + oz) * Ho + oy) * Wo + ox) * COUT;
#pragma unroll
        for (int h = 0; h < Q; ++h) sk[cc * Q + h] = ok ? ldg4(p.skip + o + 4 * h) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    for (int s = 0; s < nsteps; ++s) {
      const int stage = s & 1;
      const int iz = z0 + s;
      load_skip(iz, 0, 0);
      mbar_wait(acc_full + stage, (s >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int it = 0; it < KM * NPASS; ++it) {
        const int mt = it / NPASS, pass = it % NPASS;
        uint32_t vr[64];
        const uint32_t ta = tmem_base + ((uint32_t)(warp * 32) << 16) + stage * KM * N + mt * N + pass * 64;
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 16) tmem_ld16_nowait(ta + c0, &vr[c0]);
        tmem_ld_wait();
        if (it == KM * NPASS - 1) {
          tc_fence_before();
          mbar_arrive(acc_empty + stage);
        }
        const int gy = y0 + mt * 4 + warp;
        if (col_ok && gy < p.H) {
#pragma unroll
          for (int cc = 0; cc < CPP; ++cc) {
            const int cls = pass * CPP + cc;
            const int oz = 2 * iz + (cls >> 2), oy = 2 * gy + ((cls >> 1) & 1), ox = 2 * gx + (cls & 1);
            const int64_t o = ((((int64_t)b * Do + oz) * Ho + oy) * Wo + ox) * COUT;
#pragma unroll
            for (int h = 0; h < Q; ++h) {
              float4 r = make_float4(__uint_as_float(vr[cc * COUT + 4 * h]), __uint_as_float(vr[cc * COUT + 4 * h + 1]),
                                     __uint_as_float(vr[cc * COUT + 4 * h + 2]), __uint_as_float(vr[cc * COUT + 4 * h + 3]));
              if (p.bias) {
                const float4 bb = ldg4(p.bias + 4 * h);
                r.x += bb.x; r.y += bb.y; r.z += bb.z; r.w += bb.w;
              }
              if (p.relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
              const float4 q = sk[cc * Q + h];
              r.x += q.x; r.y += q.y; r.z += q.z; r.w += q.w;
              if (p.y) st4(p.y + o + 4 * h, r);
              if (p.ys) store_split4(p.ys, part_stride, o + 4 * h, r);
            }
          }
        }
        if (it + 1 < KM * NPASS) load_skip(iz, (it + 1) / NPASS, (it + 1) % NPASS);
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

// Weight image: torch ConvTranspose3d weight [Cin][Cout][27] x BN scale ->
//   img[term][ dz=0: shift(4: dy*2+dx) x ks x chunk x row(N: cls*COUT + n) x e | dz=1: shift(4) x ks x chunk x row(N/2: classes 4..7) x e ]
//   channel = ks*16 + chunk*8 + e; per dimension the tap used by output parity a with input shift d is
//   k = (a == 0 ? (d == 0 ? 1 : none) : (d == 0 ? 2 : 0)); unused (class, shift) rows are zero
__global__ void pack_convT_slab_weight_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                              __nv_bfloat16* __restrict__ img, int Cin, int Cout) {
  const int KS = Cin / 16, N = 8 * Cout;
  const int e0 = 4 * KS * 2 * N * 8, e1 = 4 * KS * 2 * (N / 2) * 8, per_term = e0 + e1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per_term) return;
  const int dz = i >= e0;
  const int rows = dz ? N / 2 : N;
  int r = dz ? i - e0 : i;
  const int e = r & 7; r >>= 3;
  const int row = r % rows; r /= rows;
  const int chunk = r & 1; r >>= 1;
  const int ks = r % KS;
  const int dyx = r / KS;
  const int cls = (dz ? 4 : 0) + row / Cout, n = row % Cout;
  const int par[3] = {cls >> 2, (cls >> 1) & 1, cls & 1};
  const int dsh[3] = {dz, dyx >> 1, dyx & 1};
  int k[3] = {0, 0, 0};
  bool used = true;
  for (int a = 0; a < 3; ++a) {
    if (par[a] == 0) { if (dsh[a] == 0) k[a] = 1; else used = false; }
    else k[a] = dsh[a] == 0 ? 2 : 0;
  }
  float v = 0.f;
  if (used) {
    const int ci = ks * 16 + chunk * 8 + e;
    v = w[((int64_t)ci * Cout + n) * 27 + (k[0] * 9 + k[1] * 3 + k[2])];
    if (scale) v *= scale[n];
  }
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  img[i] = hi;
  img[per_term + i] = __float2bfloat16_rn(v - __bfloat162float(hi));
}

template <int CIN, int COUT, int KM, int NS, int MINB>
static int launch(const void* x, const Params& p, cudaStream_t st) {
  using C = Cfg<CIN, COUT, KM, NS>;
  static_assert(C::SMEM <= (MINB == 2 ? 113 * 1024 : 232448), "shared memory budget");
  auto kern = convT3d_slab_bf16_kernel<CIN, COUT, KM, NS, MINB>;
  static SmemAttrOnce smem_attr;
  {
    cudaError_t e = smem_attr.ensure(kern, (int)C::SMEM);
    RCMVS_REQUIRE(e == cudaSuccess, RCMVS_ERR_CUDA, "rcmvs_conv_transpose3d_slab_fwd: cudaFuncSetAttribute: %s",
                  cudaGetErrorString(e));
  }
  CUtensorMap map;
  const int rc = make_tmap(&map, x, p.B, p.D, p.H, p.W, CIN, C::ROWS + 1);
  if (rc != RCMVS_OK) return rc;
  dim3 grid(ceil_div(p.W, TWI), ceil_div(p.H, C::ROWS), p.B * p.chunks);
  RCMVS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, RCMVS_ERR_UNSUPPORTED, "rcmvs_conv_transpose3d_slab_fwd: grid too large");
  kern<<<grid, kThreads, C::SMEM, st>>>(map, p);
  RCMVS_CHECK_LAUNCH("rcmvs_conv_transpose3d_slab_fwd");
  return RCMVS_OK;
}

static bool supported(int Cin, int Cout) { return (Cin == 16 && Cout == 8) || (Cin == 32 && Cout == 16); }

}  // namespace slabt
}  // namespace rcmvs

extern "C" int64_t rcmvs_conv_transpose3d_slab_weight_elems(int Cin, int Cout) {
  const int KS = Cin / 16, N = 8 * Cout;
  return (int64_t)2 * (4 * KS * 2 * N * 8 + 4 * KS * 2 * (N / 2) * 8);
}

extern "C" int rcmvs_conv_transpose3d_slab_pack(const float* w, const float* scale, void* img, int Cout, int Cin,
                                                rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(w && img, RCMVS_ERR_BAD_ARG, "rcmvs_conv_transpose3d_slab_pack: null pointer");
  RCMVS_REQUIRE(slabt::supported(Cin, Cout), RCMVS_ERR_UNSUPPORTED,
                "rcmvs_conv_transpose3d_slab_pack: unsupported Cin=%d Cout=%d", Cin, Cout);
  const int per_term = (int)(rcmvs_conv_transpose3d_slab_weight_elems(Cin, Cout) / 2);
  slabt::pack_convT_slab_weight_kernel<<<ceil_div(per_term, 256), 256, 0, as_stream(stream)>>>(
      w, scale, reinterpret_cast<__nv_bfloat16*>(img), Cin, Cout);
  RCMVS_CHECK_LAUNCH("rcmvs_conv_transpose3d_slab_pack");
  return RCMVS_OK;
}

extern "C" int rcmvs_conv_transpose3d_slab_fwd(const void* x_split, const void* wimg, const float* bias, const float* skip,
                                               float* y, void* y_split, int B, int D, int H, int W, int Cin, int Cout,
                                               int relu, rcmvs_stream_t stream) {
  using namespace rcmvs;
  using namespace rcmvs::slabt;
  RCMVS_REQUIRE(x_split && wimg && (y || y_split), RCMVS_ERR_BAD_ARG, "rcmvs_conv_transpose3d_slab_fwd: null pointer");
  RCMVS_REQUIRE(supported(Cin, Cout), RCMVS_ERR_UNSUPPORTED, "rcmvs_conv_transpose3d_slab_fwd: unsupported Cin=%d Cout=%d", Cin,
                Cout);
  RCMVS_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, RCMVS_ERR_BAD_ARG, "rcmvs_conv_transpose3d_slab_fwd: bad dims");
  RCMVS_REQUIRE((reinterpret_cast<uintptr_t>(x_split) & 15) == 0, RCMVS_ERR_BAD_ARG,
                "rcmvs_conv_transpose3d_slab_fwd: x_split must be 16-byte aligned (TMA)");
  Params p;
  p.wimg = reinterpret_cast<const __nv_bfloat16*>(wimg);
  p.bias = bias; p.skip = skip; p.y = y; p.ys = reinterpret_cast<__nv_bfloat16*>(y_split);
  p.B = B; p.D = D; p.H = H; p.W = W; p.relu = relu;
  // depth chunks: minimise waves(tiles x chunks / resident CTAs) x slabs per CTA (planes + 1 halo slab)
  constexpr int KM = 2;
  const int resident = Cout == 8 ? 296 : 148;
  const int tiles = ceil_div(W, TWI) * ceil_div(H, 4 * KM) * B;
  int chunks = 1;
  {
    long best = -1;
    for (int c = 1; c <= D; ++c) {
      const int dc = ceil_div(D, c);
      const int cc = ceil_div(D, dc);
      const long cost = (long)ceil_div(tiles * cc, resident) * (dc + 1);
      if (best < 0 || cost < best) { best = cost; chunks = cc; }
    }
  }
  p.d_chunk = ceil_div(D, chunks);
  p.chunks = ceil_div(D, p.d_chunk);
  cudaStream_t st = as_stream(stream);
  if (Cout == 8) return launch<16, 8, KM, 4, 2>(x_split, p, st);       // conv11: two CTAs per SM
  return launch<32, 16, KM, 2, 1>(x_split, p, st);                       // conv9: 96 KB weight image, one CTA per SM
}
