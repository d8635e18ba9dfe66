// f1 FeatureNet (reference models/modules.py:363-464, fpn): 2-D convolutions in exact fp32, channels-last, with the
// eval-mode BatchNorm folded, bias + ReLU fused and — for the FPN lateral connections — the nearest-neighbour 2x
// upsample + add (modules.py:451,455) fused into the 1x1 convolution's epilogue.
//
// Same register-blocked direct-convolution scheme as conv3d_direct.cu: CTA = 32 (w) x TH (h) output tile, thread =
// 4 output rows x 8 output channels, shared-memory halo tile of 8 input channels at a time stored [row][c4][col] as
// float4 (a warp's LDS.128 for one tap is 32 consecutive 16-byte words), weight slice [K*K][8][Cout] in shared
// memory read as warp-uniform broadcasts.  Kernel sizes 1, 3, 5; strides 1, 2; Cout 8, 16, 32.
#include <cuda_bf16.h>

#include "common.cuh"

namespace rcmvs {

constexpr int CB2 = 8;

template <int COUT>
struct Conv2Tile {
  static constexpr int COG = COUT / 8;
  static constexpr int NWH = (COG >= 4) ? 1 : (COG == 2 ? 2 : 4);
  static constexpr int TH = 4 * NWH;
  static constexpr int THREADS = 32 * NWH * COG;
};

template <int COUT, int STRIDE, int KS>
struct Conv2Smem {
  using T = Conv2Tile<COUT>;
  static constexpr int IH = (T::TH - 1) * STRIDE + KS;
  static constexpr int IW = 31 * STRIDE + KS;
  static constexpr int X_F4 = IH * 2 * IW;
  static constexpr int W_F = KS * KS * CB2 * COUT;
  static constexpr size_t BYTES = (size_t)X_F4 * 16 + (size_t)W_F * 4;
};

template <int COUT, int STRIDE, int KS>
__global__ void __launch_bounds__(Conv2Tile<COUT>::THREADS)
conv2d_direct_kernel(const float* __restrict__ x, const float* __restrict__ wp, const float* __restrict__ bias,
                     const float* __restrict__ up, float* __restrict__ y, __nv_bfloat16* __restrict__ ys,
                     int64_t ys_part_stride, int H, int W, int Cin, int Ho, int Wo, int relu) {
  using T = Conv2Tile<COUT>;
  using S = Conv2Smem<COUT, STRIDE, KS>;
  constexpr int TH = T::TH, IH = S::IH, IW = S::IW, PAD = KS / 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* xs = reinterpret_cast<float4*>(smem_raw);
  float* ws = reinterpret_cast<float*>(smem_raw + (size_t)S::X_F4 * 16);

  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int hg = wid % T::NWH, cog = wid / T::NWH;
  const int ox0 = blockIdx.x * 32, oy0 = blockIdx.y * TH, n = blockIdx.z;
  const int iy0 = oy0 * STRIDE - PAD, ix0 = ox0 * STRIDE - PAD;
  x += (int64_t)n * H * W * Cin;

  float acc[4][8];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[j][c] = 0.f;

  for (int cb0 = 0; cb0 < Cin; cb0 += CB2) {
    __syncthreads();
    for (int e = threadIdx.x; e < S::X_F4; e += T::THREADS) {
      const int c4 = e & 1;
      int r = e >> 1;
      const int ix = r % IW;
      const int iy = r / IW;
      const int gy = iy0 + iy, gx = ix0 + ix;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = ldg4(x + ((int64_t)gy * W + gx) * Cin + cb0 + c4 * 4);
      // stride 2: columns de-interleaved by parity so a warp's reads of column 2*lane + kw are consecutive float4
      const int slot = STRIDE == 2 ? ((ix & 1) ? (IW + 1) / 2 + (ix >> 1) : (ix >> 1)) : ix;
      xs[(iy * 2 + c4) * IW + slot] = v;
    }
    for (int e = threadIdx.x; e < S::W_F; e += T::THREADS) {
      const int co = e % COUT;
      const int r = e / COUT;
      const int ci = r % CB2, tap = r / CB2;
      ws[e] = __ldg(wp + ((int64_t)tap * Cin + cb0 + ci) * COUT + co);
    }
    __syncthreads();

#pragma unroll(KS == 5 ? 1 : KS)
    for (int kw = 0; kw < KS; ++kw) {
#pragma unroll
      for (int c4 = 0; c4 < 2; ++c4) {
        constexpr int NR = 3 * STRIDE + KS;
        float4 xr[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r)
          xr[r] = xs[((hg * 4 * STRIDE + r) * 2 + c4) * IW +
                     (STRIDE == 2 ? ((kw & 1) ? (IW + 1) / 2 + lane + (kw >> 1) : lane + (kw >> 1)) : lane + kw)];
#pragma unroll
        for (int kh = 0; kh < KS; ++kh) {
          const int tap = kh * KS + kw;
#pragma unroll
          for (int ci = 0; ci < 4; ++ci) {
            const float* wrow = ws + ((tap * CB2) + c4 * 4 + ci) * COUT + cog * 8;
            const float4 w0 = *reinterpret_cast<const float4*>(wrow);
            const float4 w1 = *reinterpret_cast<const float4*>(wrow + 4);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 xv4 = xr[j * STRIDE + kh];
              const float xv = ci == 0 ? xv4.x : (ci == 1 ? xv4.y : (ci == 2 ? xv4.z : xv4.w));
              acc[j][0] = fmaf(xv, w0.x, acc[j][0]); acc[j][1] = fmaf(xv, w0.y, acc[j][1]);
              acc[j][2] = fmaf(xv, w0.z, acc[j][2]); acc[j][3] = fmaf(xv, w0.w, acc[j][3]);
              acc[j][4] = fmaf(xv, w1.x, acc[j][4]); acc[j][5] = fmaf(xv, w1.y, acc[j][5]);
              acc[j][6] = fmaf(xv, w1.z, acc[j][6]); acc[j][7] = fmaf(xv, w1.w, acc[j][7]);
            }
          }
        }
      }
    }
  }

  const int ox = ox0 + lane;
  if (ox >= Wo) return;
  float bv[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) bv[c] = bias ? __ldg(bias + cog * 8 + c) : 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int oy = oy0 + hg * 4 + j;
    if (oy >= Ho) continue;
    const int64_t o = (((int64_t)n * Ho + oy) * Wo + ox) * COUT + cog * 8;
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      v[c] = acc[j][c] + bv[c];
      if (relu) v[c] = fmaxf(v[c], 0.f);
    }
    if (up) {   // + nearest-neighbour 2x upsample of the coarser pyramid level (modules.py:451,455)
      const int64_t u = (((int64_t)n * (Ho >> 1) + (oy >> 1)) * (Wo >> 1) + (ox >> 1)) * COUT + cog * 8;
      const float4 s0 = ldg4(up + u), s1 = ldg4(up + u + 4);
      v[0] += s0.x; v[1] += s0.y; v[2] += s0.z; v[3] += s0.w;
      v[4] += s1.x; v[5] += s1.y; v[6] += s1.z; v[7] += s1.w;
    }
    if (y) {
      st4(y + o, make_float4(v[0], v[1], v[2], v[3]));
      st4(y + o + 4, make_float4(v[4], v[5], v[6], v[7]));
    }
    if (ys) {   // pre-split copy (hi = bf16(v), lo = bf16(v - hi)) for a tensor-core consumer (conv2d_slab_bf16.cu)
      __nv_bfloat16 h[8], l[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        h[c] = __float2bfloat16_rn(v[c]);
        l[c] = __float2bfloat16_rn(v[c] - __bfloat162float(h[c]));
      }
      *reinterpret_cast<uint4*>(ys + o) = *reinterpret_cast<const uint4*>(h);
      *reinterpret_cast<uint4*>(ys + ys_part_stride + o) = *reinterpret_cast<const uint4*>(l);
    }
  }
}

// [N][3][H][W] -> [N][H][W][8] (channels 3..7 zero) so that every FeatureNet layer sees Cin % 8 == 0
__global__ void images_to_nhwc8_kernel(const float* __restrict__ img, float* __restrict__ out, int64_t HW, int C) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (i >= HW) return;
  const float* p = img + (int64_t)n * C * HW + i;
  float v[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) v[c] = c < C ? __ldg(p + (int64_t)c * HW) : 0.f;
  float* o = out + ((int64_t)n * HW + i) * 8;
  st4(o, make_float4(v[0], v[1], v[2], v[3]));
  st4(o + 4, make_float4(v[4], v[5], v[6], v[7]));
}

// same, written as split-bf16 planes [2][N][H][W][8] for the tensor-core first layer
__global__ void images_to_nhwc8_split_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int64_t HW,
                                             int C, int64_t part_stride) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (i >= HW) return;
  const float* p = img + (int64_t)n * C * HW + i;
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float v0 = 2 * k < C ? __ldg(p + (int64_t)(2 * k) * HW) : 0.f, v1 = 2 * k + 1 < C ? __ldg(p + (int64_t)(2 * k + 1) * HW) : 0.f;
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
    const __nv_bfloat162 hh = __halves2bfloat162(h0, h1);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(v0 - __bfloat162float(h0), v1 - __bfloat162float(h1));
    hi[k] = *reinterpret_cast<const uint32_t*>(&hh);
    lo[k] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  const int64_t o = ((int64_t)n * HW + i) * 8;
  *reinterpret_cast<uint4*>(out + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(out + part_stride + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// torch [Cout][Cin][K][K] x scale[Cout] -> [K*K][Cin_pad][Cout] (rows Cin..Cin_pad-1 zero)
__global__ void pack_conv2d_weight_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                          float* __restrict__ packed, int Cout, int Cin, int Cin_pad, int taps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= taps * Cin_pad * Cout) return;
  const int co = i % Cout;
  const int ci = (i / Cout) % Cin_pad;
  const int tap = i / (Cout * Cin_pad);
  float v = 0.f;
  if (ci < Cin) {
    v = w[((int64_t)co * Cin + ci) * taps + tap];
    if (scale) v *= scale[co];
  }
  packed[i] = v;
}

// FPN laterals (modules.py:451, 455: intra = up2(coarser) + inner(c)): 1x1 convolution to 32 channels + bias + the
// nearest-2x-upsampled coarser map, as a streaming kernel — the generic tile kernel above spends its time staging halo
// tiles a 1x1 tap does not need.  Thread = one pixel x 8 output channels (4 threads per pixel: a warp writes 8 pixels =
// 1 KB contiguous), weights [CIN][32] in shared memory (4 distinct 32-byte rows per warp: broadcast reads).
template <int CIN>
__global__ void __launch_bounds__(256)
conv1x1_c32_kernel(const float* __restrict__ x, const float* __restrict__ wp, const float* __restrict__ bias,
                   const float* __restrict__ up, float* __restrict__ y, __nv_bfloat16* __restrict__ ys,
                   int64_t ys_part_stride, int64_t npix, int H, int W, int relu) {
  __shared__ __align__(16) float ws[CIN * 32];
  for (int i = threadIdx.x; i < CIN * 32; i += 256) ws[i] = __ldg(wp + i);
  __syncthreads();
  const int cg = threadIdx.x & 3;
  float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
  if (bias) { b0 = ldg4(bias + cg * 8); b1 = ldg4(bias + cg * 8 + 4); }
  for (int64_t pix = (int64_t)blockIdx.x * 64 + (threadIdx.x >> 2); pix < npix; pix += (int64_t)gridDim.x * 64) {
    float4 xv[CIN / 4];
#pragma unroll
    for (int i = 0; i < CIN / 4; ++i) xv[i] = ldg4(x + pix * CIN + 4 * i);
    float4 u0 = make_float4(0.f, 0.f, 0.f, 0.f), u1 = u0;
    if (up) {
      const int px = (int)(pix % W), py = (int)((pix / W) % H);
      const int64_t n = pix / ((int64_t)W * H);
      const int64_t uo = ((n * (H / 2) + (py >> 1)) * (W / 2) + (px >> 1)) * 32 + cg * 8;
      u0 = ldg4(up + uo); u1 = ldg4(up + uo + 4);
    }
    float4 a0 = b0, a1 = b1;
#pragma unroll
    for (int i = 0; i < CIN / 4; ++i) {
      const float xs4[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 w0 = *reinterpret_cast<const float4*>(ws + (4 * i + k) * 32 + cg * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(ws + (4 * i + k) * 32 + cg * 8 + 4);
        a0.x = fmaf(xs4[k], w0.x, a0.x); a0.y = fmaf(xs4[k], w0.y, a0.y); a0.z = fmaf(xs4[k], w0.z, a0.z); a0.w = fmaf(xs4[k], w0.w, a0.w);
        a1.x = fmaf(xs4[k], w1.x, a1.x); a1.y = fmaf(xs4[k], w1.y, a1.y); a1.z = fmaf(xs4[k], w1.z, a1.z); a1.w = fmaf(xs4[k], w1.w, a1.w);
      }
    }
    if (relu) {
      a0.x = fmaxf(a0.x, 0.f); a0.y = fmaxf(a0.y, 0.f); a0.z = fmaxf(a0.z, 0.f); a0.w = fmaxf(a0.w, 0.f);
      a1.x = fmaxf(a1.x, 0.f); a1.y = fmaxf(a1.y, 0.f); a1.z = fmaxf(a1.z, 0.f); a1.w = fmaxf(a1.w, 0.f);
    }
    a0.x += u0.x; a0.y += u0.y; a0.z += u0.z; a0.w += u0.w;
    a1.x += u1.x; a1.y += u1.y; a1.z += u1.z; a1.w += u1.w;
    const int64_t o = pix * 32 + cg * 8;
    if (y) { st4(y + o, a0); st4(y + o + 4, a1); }
    if (ys) {
      const float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * k]), h1 = __float2bfloat16_rn(v[2 * k + 1]);
        const __nv_bfloat162 hh = __halves2bfloat162(h0, h1);
        const __nv_bfloat162 ll = __floats2bfloat162_rn(v[2 * k] - __bfloat162float(h0), v[2 * k + 1] - __bfloat162float(h1));
        hi[k] = *reinterpret_cast<const uint32_t*>(&hh);
        lo[k] = *reinterpret_cast<const uint32_t*>(&ll);
      }
      *reinterpret_cast<uint4*>(ys + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(ys + ys_part_stride + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

template <int CIN>
static int launch_conv1x1_c32(const float* x, const float* wp, const float* bias, const float* up, float* y, void* ys, int N,
                              int H, int W, int relu, cudaStream_t st) {
  const int64_t npix = (int64_t)N * H * W;
  const int64_t want = ceil_div<int64_t>(npix, 64);
  const int grid = (int)(want < 148 * 16 ? want : 148 * 16);
  conv1x1_c32_kernel<CIN><<<grid, 256, 0, st>>>(x, wp, bias, up, y, reinterpret_cast<__nv_bfloat16*>(ys), npix * 32, npix, H, W,
                                                relu);
  RCMVS_CHECK_LAUNCH("rcmvs_conv2d_fwd");
  return RCMVS_OK;
}

template <int COUT, int STRIDE, int KS>
static int launch_conv2d(const float* x, const float* wp, const float* bias, const float* up, float* y, void* ys, int N,
                         int H, int W, int Cin, int relu, cudaStream_t st) {
  using T = Conv2Tile<COUT>;
  using S = Conv2Smem<COUT, STRIDE, KS>;
  const int Ho = (H - 1) / STRIDE + 1, Wo = (W - 1) / STRIDE + 1;
  auto kern = conv2d_direct_kernel<COUT, STRIDE, KS>;
  static SmemAttrOnce smem_attr;
  {
    cudaError_t e = smem_attr.ensure(kern, (int)S::BYTES);
    RCMVS_REQUIRE(e == cudaSuccess, RCMVS_ERR_CUDA, "rcmvs_conv2d_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  RCMVS_REQUIRE(N <= 65535 && ceil_div(Ho, T::TH) <= 65535, RCMVS_ERR_UNSUPPORTED, "rcmvs_conv2d_fwd: grid too large");
  dim3 grid(ceil_div(Wo, 32), ceil_div(Ho, T::TH), N);
  kern<<<grid, T::THREADS, S::BYTES, st>>>(x, wp, bias, up, y, reinterpret_cast<__nv_bfloat16*>(ys),
                                           (int64_t)N * Ho * Wo * COUT, H, W, Cin, Ho, Wo, relu);
  RCMVS_CHECK_LAUNCH("rcmvs_conv2d_fwd");
  return RCMVS_OK;
}

}  // namespace rcmvs

extern "C" int rcmvs_images_to_nhwc8(const float* img, float* out, int N, int C, int H, int W, rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(img && out, RCMVS_ERR_BAD_ARG, "rcmvs_images_to_nhwc8: null pointer");
  RCMVS_REQUIRE(N > 0 && C > 0 && C <= 8 && H > 0 && W > 0 && N <= 65535, RCMVS_ERR_BAD_ARG, "rcmvs_images_to_nhwc8: bad dims");
  const int64_t HW = (int64_t)H * W;
  dim3 grid((unsigned)ceil_div<int64_t>(HW, 256), N);
  images_to_nhwc8_kernel<<<grid, 256, 0, as_stream(stream)>>>(img, out, HW, C);
  RCMVS_CHECK_LAUNCH("rcmvs_images_to_nhwc8");
  return RCMVS_OK;
}

extern "C" int rcmvs_images_to_nhwc8_split(const float* img, void* out_split, int N, int C, int H, int W,
                                           rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(img && out_split, RCMVS_ERR_BAD_ARG, "rcmvs_images_to_nhwc8_split: null pointer");
  RCMVS_REQUIRE(N > 0 && C > 0 && C <= 8 && H > 0 && W > 0 && N <= 65535, RCMVS_ERR_BAD_ARG,
                "rcmvs_images_to_nhwc8_split: bad dims");
  const int64_t HW = (int64_t)H * W;
  dim3 grid((unsigned)ceil_div<int64_t>(HW, 256), N);
  images_to_nhwc8_split_kernel<<<grid, 256, 0, as_stream(stream)>>>(img, reinterpret_cast<__nv_bfloat16*>(out_split), HW, C,
                                                                    (int64_t)N * HW * 8);
  RCMVS_CHECK_LAUNCH("rcmvs_images_to_nhwc8_split");
  return RCMVS_OK;
}

extern "C" int rcmvs_pack_conv2d_weight(const float* w, const float* scale, float* packed, int Cout, int Cin,
                                        int Cin_pad, int K, rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(w && packed, RCMVS_ERR_BAD_ARG, "rcmvs_pack_conv2d_weight: null pointer");
  RCMVS_REQUIRE(Cout > 0 && Cin > 0 && Cin_pad >= Cin && K > 0, RCMVS_ERR_BAD_ARG, "rcmvs_pack_conv2d_weight: bad dims");
  const int n = K * K * Cin_pad * Cout;
  pack_conv2d_weight_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(w, scale, packed, Cout, Cin, Cin_pad, K * K);
  RCMVS_CHECK_LAUNCH("rcmvs_pack_conv2d_weight");
  return RCMVS_OK;
}

extern "C" int rcmvs_conv2d_fwd_split(const float* x, const float* packed_w, const float* bias, const float* up, float* y,
                                      void* y_split, int N, int H, int W, int Cin, int Cout, int K, int stride, int relu,
                                      rcmvs_stream_t stream);

extern "C" int rcmvs_conv2d_fwd(const float* x, const float* packed_w, const float* bias, const float* up, float* y,
                                int N, int H, int W, int Cin, int Cout, int K, int stride, int relu,
                                rcmvs_stream_t stream) {
  return rcmvs_conv2d_fwd_split(x, packed_w, bias, up, y, nullptr, N, H, W, Cin, Cout, K, stride, relu, stream);
}

extern "C" int rcmvs_conv2d_fwd_split(const float* x, const float* packed_w, const float* bias, const float* up, float* y,
                                      void* y_split, int N, int H, int W, int Cin, int Cout, int K, int stride, int relu,
                                      rcmvs_stream_t stream) {
  using namespace rcmvs;
  void* ys = y_split;
  RCMVS_REQUIRE(x && packed_w && (y || ys), RCMVS_ERR_BAD_ARG, "rcmvs_conv2d_fwd: null pointer");
  RCMVS_REQUIRE(N > 0 && H > 0 && W > 0, RCMVS_ERR_BAD_ARG, "rcmvs_conv2d_fwd: bad dims");
  RCMVS_REQUIRE(Cin % 8 == 0 && Cin > 0, RCMVS_ERR_UNSUPPORTED, "rcmvs_conv2d_fwd: Cin=%d must be a multiple of 8", Cin);
  RCMVS_REQUIRE(!up || (stride == 1 && H % 2 == 0 && W % 2 == 0), RCMVS_ERR_UNSUPPORTED,
                "rcmvs_conv2d_fwd: the fused upsample-add needs stride 1 and even H, W");
  cudaStream_t st = as_stream(stream);
  if (K == 1 && stride == 1 && Cout == 32) {            // FPN laterals / out1: streaming 1x1 kernel
    if (Cin == 8) return launch_conv1x1_c32<8>(x, packed_w, bias, up, y, ys, N, H, W, relu, st);
    if (Cin == 16) return launch_conv1x1_c32<16>(x, packed_w, bias, up, y, ys, N, H, W, relu, st);
    if (Cin == 32) return launch_conv1x1_c32<32>(x, packed_w, bias, up, y, ys, N, H, W, relu, st);
  }
#define RCMVS_C2(CO, S, KK) \
  if (Cout == CO && stride == S && K == KK) return launch_conv2d<CO, S, KK>(x, packed_w, bias, up, y, ys, N, H, W, Cin, relu, st)
  RCMVS_C2(8, 1, 3); RCMVS_C2(16, 1, 3); RCMVS_C2(32, 1, 3);
  RCMVS_C2(16, 2, 5); RCMVS_C2(32, 2, 5);
  RCMVS_C2(32, 1, 1);
#undef RCMVS_C2
  set_error("rcmvs_conv2d_fwd: no kernel for Cout=%d K=%d stride=%d", Cout, K, stride);
  return RCMVS_ERR_UNSUPPORTED;
}
