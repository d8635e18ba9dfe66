// a3 CostRegNet building blocks, exact-fp32 CUDA-core path (reference models/modules.py:118-210: Conv3d/Deconv3d =
// conv(k3,p1,bias=False) -> BatchNorm3d -> ReLU, then the skip add in CostRegNet.forward :497-499).
//
// Eval-mode BatchNorm is folded at pack time (scale into the weights, shift into `bias`), so one kernel does
// conv + bias + ReLU + skip and each activation is written exactly once, channels-last [D][H][W][C].
//
// Mapping (direct convolution, register-blocked):
//   CTA   = output tile 1(d) x TH(h) x 32(w), all Cout
//   warp  = 32 consecutive w positions; thread = 4 output rows (h) x CO_T(<=8) output channels -> 32 accumulators
//   smem  = input halo tile for CB=8 input channels and ONE depth tap (kd) at a time, stored [row][c4][col] as float4
//           so that a warp's LDS.128 for one tap is 32 consecutive 16-byte words (conflict-free at stride 1), plus
//           the 9xCBxCOUT weight slice of that kd (warp-uniform broadcasts).  Small tiles (<= 36 KB) keep several
//           CTAs resident per SM; layers with few output voxels additionally split Cout over CTAs (COUT = channels
//           per CTA, grid.z carries the split) so that the 1/4- and 1/8-resolution levels still fill 148 SMs.
//   Per (kd,kw,c4) a thread loads 6 input float4 and 24 weight float4 for 384 FMAs.
// This is the fp32 reference-precision path; the tcgen05 implicit-GEMM path lives in conv3d_umma.cu.
#include "common.cuh"

namespace rcmvs {

constexpr int CB = 8;  // input channels staged per pass

template <int COUT>
struct ConvTile {
  static constexpr int CO_T = COUT >= 8 ? 8 : COUT;
  static constexpr int COG = COUT / CO_T;
  static constexpr int NWH = (COG >= 4) ? 1 : (COG == 2 ? 2 : 4);
  static constexpr int TH = 4 * NWH;
  static constexpr int THREADS = 32 * NWH * COG;
};

template <int COUT, int STRIDE>
struct ConvSmem {
  using T = ConvTile<COUT>;
  static constexpr int IH = (T::TH - 1) * STRIDE + 3;
  static constexpr int IW = 31 * STRIDE + 3;
  static constexpr int X_F4 = IH * 2 * IW;                // float4 count: ONE depth slab (kd) at a time
  static constexpr int W_F = 9 * CB * COUT;               // float count: the 9 in-plane taps of that kd
  static constexpr size_t BYTES = (size_t)X_F4 * 16 + (size_t)W_F * 4;
};

template <int COUT, int STRIDE>
__global__ void __launch_bounds__(ConvTile<COUT>::THREADS)
conv3d_direct_kernel(const float* __restrict__ x, const float* __restrict__ wp, const float* __restrict__ bias,
                     const float* __restrict__ skip, float* __restrict__ y, int D, int H, int W, int Cin, int Do,
                     int Ho, int Wo, int relu, int Ctot, int nsplit, int ksplit, float* __restrict__ wksp,
                     int64_t ws_stride) {
  using T = ConvTile<COUT>;
  using S = ConvSmem<COUT, STRIDE>;
  constexpr int CO_T = T::CO_T, TH = T::TH, IH = S::IH, IW = S::IW;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* xs = reinterpret_cast<float4*>(smem_raw);
  float* ws = reinterpret_cast<float*>(smem_raw + (size_t)S::X_F4 * 16);

  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int hg = wid % T::NWH;
  const int cog = wid / T::NWH;
  const int ox0 = blockIdx.x * 32, oy0 = blockIdx.y * TH;
  const int kidx = blockIdx.z % ksplit;                      // split-K slice (input-channel chunks) of this CTA
  const int zz = blockIdx.z / ksplit;
  const int co0 = (zz % nsplit) * COUT;                      // first output channel of this CTA
  const int bz = zz / nsplit;
  const int b = bz / Do, od = bz % Do;
  const int nchunks = Cin / CB, cper = (nchunks + ksplit - 1) / ksplit;
  const int cb_begin = kidx * cper * CB, cb_end = min(Cin, (kidx + 1) * cper * CB);
  const int iz0 = od * STRIDE - 1, iy0 = oy0 * STRIDE - 1, ix0 = ox0 * STRIDE - 1;
  x += (int64_t)b * D * H * W * Cin;

  float acc[4][CO_T];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int c = 0; c < CO_T; ++c) acc[j][c] = 0.f;

  for (int cb0 = cb_begin; cb0 < cb_end; cb0 += CB) {
#pragma unroll 1
    for (int kd = 0; kd < 3; ++kd) {
      const int gz = iz0 + kd;
      if (gz < 0 || gz >= D) continue;                       // zero padding in depth: nothing to accumulate (CTA-uniform)
      __syncthreads();
      // ---- stage the input halo slab for channels [cb0, cb0+8): element e -> (iy, ix, c4), c4 fastest ----
      for (int e = threadIdx.x; e < S::X_F4; e += T::THREADS) {
        const int c4 = e & 1;
        const int r = e >> 1;
        const int ix = r % IW, iy = r / IW;
        const int gy = iy0 + iy, gx = ix0 + ix;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gy >= 0 && gy < H && gx >= 0 && gx < W)
          v = ldg4(x + (((int64_t)gz * H + gy) * W + gx) * Cin + cb0 + c4 * 4);
        // stride 2: columns are stored de-interleaved by parity ([even 0,2,..][odd 1,3,..]) so that the lanes of a warp,
        // which read column 2*lane + kw, touch consecutive float4 (no 2-way bank conflict on every activation load)
        const int slot = STRIDE == 2 ? ((ix & 1) ? (IW + 1) / 2 + (ix >> 1) : (ix >> 1)) : ix;
        xs[(iy * 2 + c4) * IW + slot] = v;
      }
      // ---- stage the weight slice [9][CB][COUT] of this kd ----
      for (int e = threadIdx.x; e < S::W_F; e += T::THREADS) {
        const int co = e % COUT;
        const int r = e / COUT;
        const int ci = r % CB, tap = r / CB;
        ws[e] = __ldg(wp + ((int64_t)(kd * 9 + tap) * Cin + cb0 + ci) * Ctot + co0 + co);
      }
      __syncthreads();

#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
        for (int c4 = 0; c4 < 2; ++c4) {
          constexpr int NR = 3 * STRIDE + 3;  // distinct input rows touched by 4 output rows
          float4 xr[NR];
#pragma unroll
          for (int r = 0; r < NR; ++r)
            xr[r] = xs[((hg * 4 * STRIDE + r) * 2 + c4) * IW +
                       (STRIDE == 2 ? ((kw & 1) ? (IW + 1) / 2 + lane : lane + (kw >> 1)) : lane + kw)];
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const int tap = kh * 3 + kw;
            if constexpr (COUT == 1) {
              // single output channel (`prob`): the 4 input-channel weights of this tap are one float4
              const float4 w4 = *reinterpret_cast<const float4*>(ws + (tap * CB) + c4 * 4);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 xv4 = xr[j * STRIDE + kh];
                acc[j][0] = fmaf(xv4.x, w4.x, fmaf(xv4.y, w4.y, fmaf(xv4.z, w4.z, fmaf(xv4.w, w4.w, acc[j][0]))));
              }
              continue;
            }
#pragma unroll
            for (int ci = 0; ci < 4; ++ci) {
              float wv[CO_T];
              const float* wrow = ws + ((tap * CB) + c4 * 4 + ci) * COUT + cog * CO_T;
              if constexpr (CO_T == 8) {
                const float4 w0 = *reinterpret_cast<const float4*>(wrow);
                const float4 w1 = *reinterpret_cast<const float4*>(wrow + 4);
                wv[0] = w0.x; wv[1] = w0.y; wv[2] = w0.z; wv[3] = w0.w;
                wv[4] = w1.x; wv[5] = w1.y; wv[6] = w1.z; wv[7] = w1.w;
              } else {
#pragma unroll
                for (int c = 0; c < CO_T; ++c) wv[c] = wrow[c];
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 xv4 = xr[j * STRIDE + kh];
                const float xv = ci == 0 ? xv4.x : (ci == 1 ? xv4.y : (ci == 2 ? xv4.z : xv4.w));
#pragma unroll
                for (int c = 0; c < CO_T; ++c) acc[j][c] = fmaf(xv, wv[c], acc[j][c]);
              }
            }
          }
        }
      }
    }
  }

  // ---- epilogue: + bias, ReLU, + skip (after the ReLU), streaming store ----
  const int ox = ox0 + lane;
  if (ox >= Wo) return;
  float bv[CO_T];
#pragma unroll
  for (int c = 0; c < CO_T; ++c) bv[c] = bias ? __ldg(bias + co0 + cog * CO_T + c) : 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int oy = oy0 + hg * 4 + j;
    if (oy >= Ho) continue;
    const int64_t o = ((((int64_t)b * Do + od) * Ho + oy) * Wo + ox) * Ctot + co0 + cog * CO_T;
    if (ksplit > 1) {
      // deterministic split-K: raw partial sums go to this slice's workspace; splitk_reduce_kernel finishes the layer
      float* wsp = wksp + (int64_t)kidx * ws_stride + o;
#pragma unroll
      for (int c = 0; c < CO_T; ++c) wsp[c] = acc[j][c];
      continue;
    }
    float v[CO_T];
#pragma unroll
    for (int c = 0; c < CO_T; ++c) {
      v[c] = acc[j][c] + bv[c];
      if (relu) v[c] = fmaxf(v[c], 0.f);
    }
    if constexpr (CO_T == 8) {
      if (skip) {
        const float4 s0 = ldg4(skip + o), s1 = ldg4(skip + o + 4);
        v[0] += s0.x; v[1] += s0.y; v[2] += s0.z; v[3] += s0.w;
        v[4] += s1.x; v[5] += s1.y; v[6] += s1.z; v[7] += s1.w;
      }
      st4(y + o, make_float4(v[0], v[1], v[2], v[3]));
      st4(y + o + 4, make_float4(v[4], v[5], v[6], v[7]));
    } else {
#pragma unroll
      for (int c = 0; c < CO_T; ++c) y[o + c] = v[c] + (skip ? __ldg(skip + o + c) : 0.f);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// `prob` layer (8 -> 1 channels, stride 1; reference models/modules.py:488): memory-bound, 27 x 8 MACs per output.
// Input-stationary plane march: a CTA owns a 16 x 32 (h x w) column of the volume and walks a chunk of depth planes.
// Every input plane is staged ONCE (cp.async, double-buffered) and scattered into the three output planes it touches
// (kd = 0/1/2 -> out z+1 / z / z-1), whose partial sums stay in registers; plane z-1 is complete after plane z.
// Per staged plane a thread does 36 LDS.128 of activations for 864 FMAs (the per-kd slab kernel above needs 3x that).
// ------------------------------------------------------------------------------------------------------------------
namespace march {
constexpr int TH = 16, TW = 32, IH = TH + 2, IW = TW + 2;
constexpr int PLANE_F4 = IH * 2 * IW;                        // float4 per staged plane: [iy][c4][ix]
constexpr int THREADS = 128;
constexpr size_t SMEM = (size_t)2 * PLANE_F4 * 16 + 27 * 8 * 4;

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
}  // namespace march

__global__ void __launch_bounds__(march::THREADS)
conv3d_c8_to_c1_march_kernel(const float* __restrict__ x, const float* __restrict__ wp, const float* __restrict__ bias,
                             const float* __restrict__ skip, float* __restrict__ y, int D, int H, int W, int relu,
                             int planes_per_chunk, int nchunks) {
  using namespace march;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* xs = reinterpret_cast<float4*>(smem_raw);
  float* ws = reinterpret_cast<float*>(smem_raw + (size_t)2 * PLANE_F4 * 16);
  const int lane = threadIdx.x & 31, hg = threadIdx.x >> 5;
  const int ox0 = blockIdx.x * TW, oy0 = blockIdx.y * TH;
  const int b = blockIdx.z / nchunks, chunk = blockIdx.z % nchunks;
  const int z0 = chunk * planes_per_chunk, z1 = min(D, z0 + planes_per_chunk);
  x += (int64_t)b * D * H * W * 8;
  const int64_t ybase = (int64_t)b * D * H * W;

  for (int e = threadIdx.x; e < 27 * 8; e += THREADS) ws[e] = __ldg(wp + e);      // packed [27][8][1]

  auto stage = [&](int z, int buf) {
    // plane z of the halo tile -> xs[buf]; out-of-volume voxels (incl. whole planes z = -1, D) are zero-filled
    const bool zin = z >= 0 && z < D;
    const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(xs + (size_t)buf * PLANE_F4);
    for (int e = threadIdx.x; e < PLANE_F4; e += THREADS) {
      const int c4 = e & 1, r = e >> 1;
      const int ix = r % IW, iy = r / IW;
      const int gy = oy0 - 1 + iy, gx = ox0 - 1 + ix;
      const bool ok = zin && gy >= 0 && gy < H && gx >= 0 && gx < W;
      const float* src = ok ? x + (((int64_t)z * H + gy) * W + gx) * 8 + c4 * 4 : x;
      cp_async16_zfill(dst0 + ((iy * 2 + c4) * IW + ix) * 16, src, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  float acc[3][4];                                   // [0]: out plane z-1, [1]: z, [2]: z+1 (z = plane being consumed)
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[a][j] = 0.f;

  const float bv = bias ? __ldg(bias) : 0.f;
  const int ox = ox0 + lane;
  stage(z0 - 1, 0);
  int buf = 0;
  for (int z = z0 - 1; z <= z1; ++z, buf ^= 1) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                 // plane z landed; everyone is done reading the other buffer
    if (z < z1) stage(z + 1, buf ^ 1);
    if (z >= 0 && z < D) {                           // CTA-uniform: padding planes contribute nothing
      const float4* xp = xs + (size_t)buf * PLANE_F4;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
        for (int c4 = 0; c4 < 2; ++c4) {
          float4 xr[6];
#pragma unroll
          for (int r = 0; r < 6; ++r) xr[r] = xp[((hg * 4 + r) * 2 + c4) * IW + lane + kw];
#pragma unroll
          for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
              const float4 w4 = *reinterpret_cast<const float4*>(ws + ((kd * 9 + kh * 3 + kw) * 8) + c4 * 4);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 v = xr[j + kh];
                float& a = acc[2 - kd][j];           // input plane z feeds out plane z + 1 - kd
                a = fmaf(v.x, w4.x, fmaf(v.y, w4.y, fmaf(v.z, w4.z, fmaf(v.w, w4.w, a))));
              }
            }
          }
        }
      }
    }
    // out plane z-1 has now seen input planes z-2, z-1, z
    const int oz = z - 1;
    if (oz >= z0 && oz < z1 && ox < W) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int oy = oy0 + hg * 4 + j;
        if (oy >= H) continue;
        const int64_t o = ybase + ((int64_t)oz * H + oy) * W + ox;
        float v = acc[0][j] + bv;
        if (relu) v = fmaxf(v, 0.f);
        if (skip) v += __ldg(skip + o);
        y[o] = v;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[0][j] = acc[1][j]; acc[1][j] = acc[2][j]; acc[2][j] = 0.f; }
  }
}

static int launch_march_c8_c1(const float* x, const float* wp, const float* bias, const float* skip, float* y, int B,
                              int D, int H, int W, int relu, cudaStream_t st) {
  using namespace march;
  static SmemAttrOnce smem_attr;
  {
    cudaError_t e = smem_attr.ensure(conv3d_c8_to_c1_march_kernel, (int)SMEM);
    RCMVS_REQUIRE(e == cudaSuccess, RCMVS_ERR_CUDA, "rcmvs_conv3d_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  // depth chunks: as long as possible (every chunk re-reads 2 halo planes) while the grid still covers ~3 waves
  const int64_t tiles = (int64_t)ceil_div(W, TW) * ceil_div(H, TH) * B;
  int ppc = D;
  while (ppc > 8 && tiles * ceil_div(D, ppc) < 3 * 148) ppc = (ppc + 1) / 2;
  const int nchunks = ceil_div(D, ppc);
  RCMVS_REQUIRE((int64_t)B * nchunks <= 65535 && ceil_div(H, TH) <= 65535, RCMVS_ERR_UNSUPPORTED,
                "rcmvs_conv3d_fwd: grid too large");
  dim3 grid(ceil_div(W, TW), ceil_div(H, TH), B * nchunks);
  conv3d_c8_to_c1_march_kernel<<<grid, THREADS, SMEM, st>>>(x, wp, bias, skip, y, D, H, W, relu, ppc, nchunks);
  RCMVS_CHECK_LAUNCH("rcmvs_conv3d_fwd(march)");
  return RCMVS_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// ConvTranspose3d(k=3, stride=2, padding=1, output_padding=1): out[z] += x[i]·w[k] with z = 2i + k - 1.
// Per dimension: even output (a=0) <- (k=1, i); odd output (a=1) <- (k=2, i) and (k=0, i+1).
// Thread = one input-space voxel -> its 2x2x2 output block x 8 output channels (64 accumulators).
// CTA = input tile 1(d) x TH(h) x 32(w); smem = (2 x (TH+1) x 33) input voxels x 8 channels + 27x8xCout weights.
// ------------------------------------------------------------------------------------------------------------------
template <int COUT>
struct TConvTile {
  static constexpr int COG = COUT / 8;
  static constexpr int NWH = (COG >= 4) ? 1 : (COG == 2 ? 2 : 4);
  static constexpr int TH = NWH;  // one input row per warp-row
  static constexpr int THREADS = 32 * NWH * COG;
  static constexpr int IH = TH + 1, IW = 33;
  static constexpr int X_F4 = 2 * IH * 2 * IW;
  static constexpr int W_F = 27 * CB * COUT;
  static constexpr size_t BYTES = (size_t)X_F4 * 16 + (size_t)W_F * 4;
};

template <int COUT>
__global__ void __launch_bounds__(TConvTile<COUT>::THREADS)
conv_transpose3d_direct_kernel(const float* __restrict__ x, const float* __restrict__ wp,
                               const float* __restrict__ bias, const float* __restrict__ skip, float* __restrict__ y,
                               int D, int H, int W, int Cin, int relu, int Ctot, int nsplit) {
  using T = TConvTile<COUT>;
  constexpr int IH = T::IH, IW = T::IW;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* xs = reinterpret_cast<float4*>(smem_raw);
  float* ws = reinterpret_cast<float*>(smem_raw + (size_t)T::X_F4 * 16);

  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int hg = wid % T::NWH;
  const int cog = wid / T::NWH;
  const int ix0 = blockIdx.x * 32, iy0 = blockIdx.y * T::TH;
  const int co0 = (blockIdx.z % nsplit) * COUT;
  const int bz = blockIdx.z / nsplit;
  const int b = bz / D, iz = bz % D;
  x += (int64_t)b * D * H * W * Cin;

  float acc[8][8];
#pragma unroll
  for (int o = 0; o < 8; ++o)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[o][c] = 0.f;

  for (int cb0 = 0; cb0 < Cin; cb0 += CB) {
    __syncthreads();
    for (int e = threadIdx.x; e < T::X_F4; e += T::THREADS) {
      const int c4 = e & 1;
      int r = e >> 1;
      const int ix = r % IW; r /= IW;
      const int iy = r % IH;
      const int dz = r / IH;
      const int gz = iz + dz, gy = iy0 + iy, gx = ix0 + ix;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gz < D && gy < H && gx < W) v = ldg4(x + (((int64_t)gz * H + gy) * W + gx) * Cin + cb0 + c4 * 4);
      xs[((dz * IH + iy) * 2 + c4) * IW + ix] = v;
    }
    for (int e = threadIdx.x; e < T::W_F; e += T::THREADS) {
      const int co = e % COUT;
      const int r = e / COUT;
      const int ci = r % CB, tap = r / CB;
      ws[e] = __ldg(wp + ((int64_t)tap * Cin + cb0 + ci) * Ctot + co0 + co);
    }
    __syncthreads();

#pragma unroll
    for (int c4 = 0; c4 < 2; ++c4) {
      float4 xn[8];  // neighbourhood (dz,dy,dx) in {0,1}^3
#pragma unroll
      for (int n = 0; n < 8; ++n)
        xn[n] = xs[(((n >> 2) * IH + hg + ((n >> 1) & 1)) * 2 + c4) * IW + lane + (n & 1)];
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
#pragma unroll
        for (int o = 0; o < 8; ++o) {       // output parity (a,b,c) = (o>>2, (o>>1)&1, o&1)
          const int a = o >> 2, bq = (o >> 1) & 1, cq = o & 1;
#pragma unroll
          for (int ta = 0; ta <= a; ++ta)   // a=0: one term (k=1,di=0); a=1: (k=2,di=0),(k=0,di=1)
#pragma unroll
            for (int tb = 0; tb <= bq; ++tb)
#pragma unroll
              for (int tc = 0; tc <= cq; ++tc) {
                const int kd = a ? (ta ? 0 : 2) : 1, dz = a ? ta : 0;
                const int kh = bq ? (tb ? 0 : 2) : 1, dy = bq ? tb : 0;
                const int kw = cq ? (tc ? 0 : 2) : 1, dx = cq ? tc : 0;
                const int tap = (kd * 3 + kh) * 3 + kw;
                const float4 xv4 = xn[dz * 4 + dy * 2 + dx];
                const float xv = ci == 0 ? xv4.x : (ci == 1 ? xv4.y : (ci == 2 ? xv4.z : xv4.w));
                const float* wrow = ws + ((tap * CB) + c4 * 4 + ci) * COUT + cog * 8;
                const float4 w0 = *reinterpret_cast<const float4*>(wrow);
                const float4 w1 = *reinterpret_cast<const float4*>(wrow + 4);
                acc[o][0] = fmaf(xv, w0.x, acc[o][0]); acc[o][1] = fmaf(xv, w0.y, acc[o][1]);
                acc[o][2] = fmaf(xv, w0.z, acc[o][2]); acc[o][3] = fmaf(xv, w0.w, acc[o][3]);
                acc[o][4] = fmaf(xv, w1.x, acc[o][4]); acc[o][5] = fmaf(xv, w1.y, acc[o][5]);
                acc[o][6] = fmaf(xv, w1.z, acc[o][6]); acc[o][7] = fmaf(xv, w1.w, acc[o][7]);
              }
        }
      }
    }
  }

  const int ix = ix0 + lane, iy = iy0 + hg;
  if (ix >= W || iy >= H) return;
  const int Do = 2 * D, Ho = 2 * H, Wo = 2 * W;
  float bv[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) bv[c] = bias ? __ldg(bias + co0 + cog * 8 + c) : 0.f;
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    const int oz = 2 * iz + (o >> 2), oy = 2 * iy + ((o >> 1) & 1), ox = 2 * ix + (o & 1);
    const int64_t off = ((((int64_t)b * Do + oz) * Ho + oy) * Wo + ox) * Ctot + co0 + cog * 8;
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      v[c] = acc[o][c] + bv[c];
      if (relu) v[c] = fmaxf(v[c], 0.f);
    }
    if (skip) {
      const float4 s0 = ldg4(skip + off), s1 = ldg4(skip + off + 4);
      v[0] += s0.x; v[1] += s0.y; v[2] += s0.z; v[3] += s0.w;
      v[4] += s1.x; v[5] += s1.y; v[6] += s1.z; v[7] += s1.w;
    }
    st4(y + off, make_float4(v[0], v[1], v[2], v[3]));
    st4(y + off + 4, make_float4(v[4], v[5], v[6], v[7]));
  }
}

// ------------------------------------------------------------------------------------------------------------------
// weight packing: torch OIDHW (conv) / IODHW (transposed conv) x BN scale -> [27][Cin][Cout]
// ------------------------------------------------------------------------------------------------------------------
__global__ void pack_weight_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                   float* __restrict__ packed, int Cout, int Cin, int transposed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 27 * Cin * Cout) return;
  const int co = i % Cout;
  const int ci = (i / Cout) % Cin;
  const int tap = i / (Cout * Cin);
  const float v = transposed ? w[((int64_t)ci * Cout + co) * 27 + tap] : w[((int64_t)co * Cin + ci) * 27 + tap];
  packed[i] = scale ? v * scale[co] : v;
}

// y = act(Σ_k ws[k] + bias) (+ skip): fixed summation order -> deterministic
__global__ void splitk_reduce_kernel(const float* __restrict__ ws, int ksplit, int64_t ws_stride,
                                     const float* __restrict__ bias, const float* __restrict__ skip,
                                     float* __restrict__ y, int64_t n4, int Ctot, int relu) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 a = ldg4(ws + i * 4);
  for (int k = 1; k < ksplit; ++k) {
    const float4 q = ldg4(ws + (int64_t)k * ws_stride + i * 4);
    a.x += q.x; a.y += q.y; a.z += q.z; a.w += q.w;
  }
  if (bias) {
    const float4 q = ldg4(bias + (int)((i * 4) % Ctot));
    a.x += q.x; a.y += q.y; a.z += q.z; a.w += q.w;
  }
  if (relu) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
  if (skip) {
    const float4 q = ldg4(skip + i * 4);
    a.x += q.x; a.y += q.y; a.z += q.z; a.w += q.w;
  }
  st4(y + i * 4, a);
}

template <int COUT, int STRIDE>
static int launch_conv(const float* x, const float* wp, const float* bias, const float* skip, float* y, int B, int D,
                       int H, int W, int Cin, int relu, int Ctot, int ksplit, float* ws, cudaStream_t st) {
  using T = ConvTile<COUT>;
  using S = ConvSmem<COUT, STRIDE>;
  const int Do = (D - 1) / STRIDE + 1, Ho = (H - 1) / STRIDE + 1, Wo = (W - 1) / STRIDE + 1;
  const int nsplit = Ctot / COUT;
  const int64_t out_elems = (int64_t)B * Do * Ho * Wo * Ctot;
  auto kern = conv3d_direct_kernel<COUT, STRIDE>;
  static SmemAttrOnce smem_attr;
  {
    cudaError_t e = smem_attr.ensure(kern, (int)S::BYTES);
    RCMVS_REQUIRE(e == cudaSuccess, RCMVS_ERR_CUDA, "rcmvs_conv3d_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  RCMVS_REQUIRE((int64_t)B * Do * nsplit * ksplit <= 65535 && ceil_div(Ho, T::TH) <= 65535, RCMVS_ERR_UNSUPPORTED,
                "rcmvs_conv3d_fwd: grid too large");
  dim3 grid(ceil_div(Wo, 32), ceil_div(Ho, T::TH), B * Do * nsplit * ksplit);
  kern<<<grid, T::THREADS, S::BYTES, st>>>(x, wp, bias, skip, y, D, H, W, Cin, Do, Ho, Wo, relu, Ctot, nsplit, ksplit,
                                           ws, out_elems);
  RCMVS_CHECK_LAUNCH("rcmvs_conv3d_fwd");
  if (ksplit > 1) {
    const int64_t n4 = out_elems / 4;
    splitk_reduce_kernel<<<(unsigned)ceil_div<int64_t>(n4, 256), 256, 0, st>>>(ws, ksplit, out_elems, bias, skip, y, n4,
                                                                             Ctot, relu);
    RCMVS_CHECK_LAUNCH("rcmvs_conv3d_fwd(split-K reduce)");
  }
  return RCMVS_OK;
}

// split-K factor: only for layers whose grid cannot fill ~2 waves even after the Cout split (tiny deep levels)
static int pick_ksplit(int Cin, int Cout, int cpc, int B, int D, int H, int W, int stride) {
  const int nchunks = Cin / CB;
  if (nchunks < 2 || Cout % 4 != 0) return 1;
  const int Do = (D - 1) / stride + 1, Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  const int th = cpc >= 32 ? 4 : (cpc == 16 ? 8 : 16);
  const int64_t ctas = (int64_t)ceil_div(Wo, 32) * ceil_div(Ho, th) * B * Do * (Cout / cpc);
  if (ctas >= 2 * 148) return 1;
  int k = (int)(ceil_div<int64_t>(2 * 148, ctas) < nchunks ? ceil_div<int64_t>(2 * 148, ctas) : nchunks);
  return max(1, min(k, 8));
}

// channels per CTA: all of Cout for the large layers; halved while the grid would not fill ~2 waves of 148 SMs
static int pick_cout_per_cta(int Cout, int B, int D, int H, int W, int stride) {
  if (Cout < 16) return Cout;
  const int Do = (D - 1) / stride + 1, Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  int c = Cout;
  auto ctas = [&](int cc) {
    const int th = cc >= 32 ? 4 : (cc == 16 ? 8 : 16);
    return (int64_t)ceil_div(Wo, 32) * ceil_div(Ho, th) * B * Do * (Cout / cc);
  };
  while (c > 8 && ctas(c) < 2 * 148) c /= 2;
  return c;
}

template <int COUT>
static int launch_tconv(const float* x, const float* wp, const float* bias, const float* skip, float* y, int B, int D,
                        int H, int W, int Cin, int relu, int Ctot, cudaStream_t st) {
  using T = TConvTile<COUT>;
  const int nsplit = Ctot / COUT;
  auto kern = conv_transpose3d_direct_kernel<COUT>;
  static SmemAttrOnce smem_attr;
  {
    cudaError_t e = smem_attr.ensure(kern, (int)T::BYTES);
    RCMVS_REQUIRE(e == cudaSuccess, RCMVS_ERR_CUDA, "rcmvs_conv_transpose3d_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  RCMVS_REQUIRE((int64_t)B * D * nsplit <= 65535 && ceil_div(H, T::TH) <= 65535, RCMVS_ERR_UNSUPPORTED,
                "rcmvs_conv_transpose3d_fwd: grid too large");
  dim3 grid(ceil_div(W, 32), ceil_div(H, T::TH), B * D * nsplit);
  kern<<<grid, T::THREADS, T::BYTES, st>>>(x, wp, bias, skip, y, D, H, W, Cin, relu, Ctot, nsplit);
  RCMVS_CHECK_LAUNCH("rcmvs_conv_transpose3d_fwd");
  return RCMVS_OK;
}

}  // namespace rcmvs

extern "C" int rcmvs_pack_conv3d_weight(const float* w, const float* scale, float* packed, int Cout, int Cin,
                                        int transposed, rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(w && packed, RCMVS_ERR_BAD_ARG, "rcmvs_pack_conv3d_weight: null pointer");
  RCMVS_REQUIRE(Cout > 0 && Cin > 0, RCMVS_ERR_BAD_ARG, "rcmvs_pack_conv3d_weight: bad dims");
  const int n = 27 * Cin * Cout;
  pack_weight_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(w, scale, packed, Cout, Cin, transposed);
  RCMVS_CHECK_LAUNCH("rcmvs_pack_conv3d_weight");
  return RCMVS_OK;
}

extern "C" int64_t rcmvs_conv3d_workspace_floats(int B, int D, int H, int W, int Cin, int Cout, int stride) {
  using namespace rcmvs;
  if (!(Cout == 1 || Cout == 8 || Cout == 16 || Cout == 32 || Cout == 64) || Cin % 8 != 0 || (stride != 1 && stride != 2))
    return 0;
  const int cpc = pick_cout_per_cta(Cout, B, D, H, W, stride);
  const int ks = pick_ksplit(Cin, Cout, cpc, B, D, H, W, stride);
  if (ks <= 1) return 0;
  const int Do = (D - 1) / stride + 1, Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  return (int64_t)ks * B * Do * Ho * Wo * Cout;
}

extern "C" int rcmvs_conv3d_fwd_ws(const float* x, const float* packed_w, const float* bias, const float* skip,
                                   float* y, float* workspace, int B, int D, int H, int W, int Cin, int Cout,
                                   int stride, int relu, rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(x && packed_w && y, RCMVS_ERR_BAD_ARG, "rcmvs_conv3d_fwd: null pointer");
  RCMVS_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, RCMVS_ERR_BAD_ARG, "rcmvs_conv3d_fwd: bad dims");
  RCMVS_REQUIRE(Cin % 8 == 0 && Cin > 0, RCMVS_ERR_UNSUPPORTED, "rcmvs_conv3d_fwd: Cin=%d must be a multiple of 8", Cin);
  RCMVS_REQUIRE(stride == 1 || stride == 2, RCMVS_ERR_UNSUPPORTED, "rcmvs_conv3d_fwd: stride must be 1 or 2");
  cudaStream_t st = as_stream(stream);
  RCMVS_REQUIRE(Cout == 1 || Cout == 8 || Cout == 16 || Cout == 32 || Cout == 64, RCMVS_ERR_UNSUPPORTED,
                "rcmvs_conv3d_fwd: Cout=%d not in {1,8,16,32,64}", Cout);
  if (Cin == 8 && Cout == 1 && stride == 1) return launch_march_c8_c1(x, packed_w, bias, skip, y, B, D, H, W, relu, st);
  const int cpc = pick_cout_per_cta(Cout, B, D, H, W, stride);
  const int ksplit = workspace ? pick_ksplit(Cin, Cout, cpc, B, D, H, W, stride) : 1;
#define RCMVS_CONV_CASE(CO)                                                                                           \
  case CO:                                                                                                            \
    return stride == 1 ? launch_conv<CO, 1>(x, packed_w, bias, skip, y, B, D, H, W, Cin, relu, Cout, ksplit, workspace, st) \
                       : launch_conv<CO, 2>(x, packed_w, bias, skip, y, B, D, H, W, Cin, relu, Cout, ksplit, workspace, st)
  switch (cpc) {
    RCMVS_CONV_CASE(1);
    RCMVS_CONV_CASE(8);
    RCMVS_CONV_CASE(16);
    RCMVS_CONV_CASE(32);
    RCMVS_CONV_CASE(64);
    default: break;
  }
#undef RCMVS_CONV_CASE
  set_error("rcmvs_conv3d_fwd: Cout=%d not in {1,8,16,32,64}", Cout);
  return RCMVS_ERR_UNSUPPORTED;
}

extern "C" int rcmvs_conv3d_fwd(const float* x, const float* packed_w, const float* bias, const float* skip, float* y,
                                int B, int D, int H, int W, int Cin, int Cout, int stride, int relu,
                                rcmvs_stream_t stream) {
  return rcmvs_conv3d_fwd_ws(x, packed_w, bias, skip, y, nullptr, B, D, H, W, Cin, Cout, stride, relu, stream);
}

extern "C" int rcmvs_conv_transpose3d_fwd(const float* x, const float* packed_w, const float* bias, const float* skip,
                                          float* y, int B, int D, int H, int W, int Cin, int Cout, int relu,
                                          rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(x && packed_w && y, RCMVS_ERR_BAD_ARG, "rcmvs_conv_transpose3d_fwd: null pointer");
  RCMVS_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, RCMVS_ERR_BAD_ARG, "rcmvs_conv_transpose3d_fwd: bad dims");
  RCMVS_REQUIRE(Cin % 8 == 0 && Cin > 0, RCMVS_ERR_UNSUPPORTED, "rcmvs_conv_transpose3d_fwd: Cin=%d must be a multiple of 8", Cin);
  cudaStream_t st = as_stream(stream);
  RCMVS_REQUIRE(Cout == 8 || Cout == 16 || Cout == 32, RCMVS_ERR_UNSUPPORTED,
                "rcmvs_conv_transpose3d_fwd: Cout=%d not in {8,16,32}", Cout);
  // channels per CTA: halve while the grid does not fill ~2 waves (input-space tiles of 32 x TH voxels)
  int cpc = Cout;
  {
    auto ctas = [&](int cc) {
      const int th = cc >= 32 ? 1 : (cc == 16 ? 2 : 4);
      return (int64_t)ceil_div(W, 32) * ceil_div(H, th) * B * D * (Cout / cc);
    };
    while (cpc > 8 && ctas(cpc) < 2 * 148) cpc /= 2;
  }
  switch (cpc) {
    case 8: return launch_tconv<8>(x, packed_w, bias, skip, y, B, D, H, W, Cin, relu, Cout, st);
    case 16: return launch_tconv<16>(x, packed_w, bias, skip, y, B, D, H, W, Cin, relu, Cout, st);
    case 32: return launch_tconv<32>(x, packed_w, bias, skip, y, B, D, H, W, Cin, relu, Cout, st);
    default: break;
  }
  set_error("rcmvs_conv_transpose3d_fwd: Cout=%d not in {8,16,32}", Cout);
  return RCMVS_ERR_UNSUPPORTED;
}
