// a5 depth hypotheses (reference models/modules.py:549-588 + the resampling in models/casmvsnet.py:191-217).
// The reference materialises a full-resolution [B,D,H,W] tensor (63-252 MB at stage 1) only to resize it; here the
// planes are produced directly at the stage resolution in one pass: 4·D·h·w bytes written, prev depth read from L2.
#include "common.cuh"

namespace rcmvs {

__global__ void hyp_first_kernel(const float* __restrict__ depth_values, float* __restrict__ planes, int B, int Dv,
                                 int D) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  int b = i / D, d = i % D;
  float dmin = depth_values[(int64_t)b * Dv];
  float dmax = depth_values[(int64_t)b * Dv + Dv - 1];
  float interval = __fdiv_rn(__fsub_rn(dmax, dmin), (float)(D - 1));      // modules.py:577
  planes[i] = __fadd_rn(dmin, __fmul_rn((float)d, interval));             // modules.py:579-580 (mul, then add)
}

// ATen area_pixel_compute_source_index, align_corners=False
__device__ __forceinline__ void src_index(int dst, float scale, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float src = fmaxf(__fsub_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), 0.5f), 0.f);
  i0 = min((int)src, in_size - 1);
  i1 = min(i0 + 1, in_size - 1);
  l1 = __fsub_rn(src, (float)i0);
  l0 = __fsub_rn(1.f, l1);
}

__device__ __forceinline__ float bilerp(float a, float b, float c, float d, float ly0, float ly1, float lx0,
                                        float lx1) {
  // h0λ·(w0λ·p00 + w1λ·p01) + h1λ·(w0λ·p10 + w1λ·p11), unfused like ATen's scalar path
  float top = __fadd_rn(__fmul_rn(lx0, a), __fmul_rn(lx1, b));
  float bot = __fadd_rn(__fmul_rn(lx0, c), __fmul_rn(lx1, d));
  return __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
}

__device__ __forceinline__ float upsampled_prev(const float* __restrict__ prev, int hp, int wp, float sy, float sx,
                                                int Y, int X) {
  int y0, y1, x0, x1;
  float ly0, ly1, lx0, lx1;
  src_index(Y, sy, hp, y0, y1, ly0, ly1);
  src_index(X, sx, wp, x0, x1, lx0, lx1);
  return bilerp(__ldg(prev + (int64_t)y0 * wp + x0), __ldg(prev + (int64_t)y0 * wp + x1),
                __ldg(prev + (int64_t)y1 * wp + x0), __ldg(prev + (int64_t)y1 * wp + x1), ly0, ly1, lx0, lx1);
}

__global__ void hyp_next_kernel(const float* __restrict__ prev, float* __restrict__ hyp, int hp, int wp, int H,
                                int W, int h, int w, int D, float half_range, const float* __restrict__ dv, int Dv,
                                double ratio) {
  if (dv) {
    // depth_interval = (max - min) / Dv of batch item 0 in Python floats (casmvsnet.py:172-174), times the stage
    // ratio (:203-204), times ndepth / 2 (modules.py:553-554), rounded to fp32 where it meets the tensor
    const double interval = ((double)__ldg(dv + Dv - 1) - (double)__ldg(dv)) / (double)Dv;
    half_range = (float)((double)D / 2.0 * (ratio * interval));
  }
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int b = blockIdx.z;
  if (x >= w) return;
  prev += (int64_t)b * hp * wp;
  hyp += (int64_t)b * D * h * w;
  const float up_sy = (float)hp / (float)H, up_sx = (float)wp / (float)W;     // scale = in/out
  const float dn_sy = (float)H / (float)h, dn_sx = (float)W / (float)w;
  int Y0, Y1, X0, X1;
  float ly0, ly1, lx0, lx1;
  src_index(y, dn_sy, H, Y0, Y1, ly0, ly1);
  src_index(x, dn_sx, W, X0, X1, lx0, lx1);
  float cmin[4], cint[4];
  const int Ys[4] = {Y0, Y0, Y1, Y1};
  const int Xs[4] = {X0, X1, X0, X1};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float cur = upsampled_prev(prev, hp, wp, up_sy, up_sx, Ys[k], Xs[k]);      // casmvsnet.py:196-198
    float lo = __fsub_rn(cur, half_range);                                      // modules.py:553
    float hi = __fadd_rn(cur, half_range);                                      // modules.py:554
    cmin[k] = lo;
    cint[k] = __fdiv_rn(__fsub_rn(hi, lo), (float)(D - 1));                     // modules.py:559
  }
  for (int d = 0; d < D; ++d) {
    float s[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) s[k] = __fadd_rn(cmin[k], __fmul_rn((float)d, cint[k]));   // modules.py:561-564
    hyp[((int64_t)d * h + y) * w + x] = bilerp(s[0], s[1], s[2], s[3], ly0, ly1, lx0, lx1);  // casmvsnet.py:212-217
  }
}

}  // namespace rcmvs

extern "C" int rcmvs_depth_hypotheses_first(const float* depth_values, float* planes, int B, int Dv, int D,
                                            rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(depth_values && planes, RCMVS_ERR_BAD_ARG, "rcmvs_depth_hypotheses_first: null pointer");
  RCMVS_REQUIRE(B > 0 && Dv > 1 && D > 1, RCMVS_ERR_BAD_ARG, "rcmvs_depth_hypotheses_first: need Dv>1, D>1");
  hyp_first_kernel<<<ceil_div(B * D, 128), 128, 0, as_stream(stream)>>>(depth_values, planes, B, Dv, D);
  RCMVS_CHECK_LAUNCH("rcmvs_depth_hypotheses_first");
  return RCMVS_OK;
}

extern "C" int rcmvs_depth_hypotheses_next(const float* prev_depth, float* hyp, int B, int hp, int wp, int H, int W,
                                           int h, int w, int D, double interval_pixel, rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(prev_depth && hyp, RCMVS_ERR_BAD_ARG, "rcmvs_depth_hypotheses_next: null pointer");
  RCMVS_REQUIRE(B > 0 && hp > 0 && wp > 0 && H > 0 && W > 0 && h > 0 && w > 0 && D > 1, RCMVS_ERR_BAD_ARG,
                "rcmvs_depth_hypotheses_next: bad dims");
  RCMVS_REQUIRE(h <= 65535 && B <= 65535, RCMVS_ERR_UNSUPPORTED, "rcmvs_depth_hypotheses_next: h or B too large");
  // `ndepth / 2 * depth_inteval_pixel` is a Python float in the reference, rounded to fp32 when it meets the tensor
  float half_range = (float)((double)D / 2.0 * interval_pixel);
  dim3 grid(ceil_div(w, 128), h, B);
  hyp_next_kernel<<<grid, 128, 0, as_stream(stream)>>>(prev_depth, hyp, hp, wp, H, W, h, w, D, half_range, nullptr, 0,
                                                       0.0);
  RCMVS_CHECK_LAUNCH("rcmvs_depth_hypotheses_next");
  return RCMVS_OK;
}

extern "C" int rcmvs_depth_hypotheses_next_dev(const float* prev_depth, const float* depth_values, int Dv,
                                               double interval_ratio, float* hyp, int B, int hp, int wp, int H, int W,
                                               int h, int w, int D, rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(prev_depth && hyp && depth_values, RCMVS_ERR_BAD_ARG, "rcmvs_depth_hypotheses_next_dev: null pointer");
  RCMVS_REQUIRE(B > 0 && hp > 0 && wp > 0 && H > 0 && W > 0 && h > 0 && w > 0 && D > 1 && Dv > 1, RCMVS_ERR_BAD_ARG,
                "rcmvs_depth_hypotheses_next_dev: bad dims");
  RCMVS_REQUIRE(h <= 65535 && B <= 65535, RCMVS_ERR_UNSUPPORTED, "rcmvs_depth_hypotheses_next_dev: h or B too large");
  dim3 grid(ceil_div(w, 128), h, B);
  hyp_next_kernel<<<grid, 128, 0, as_stream(stream)>>>(prev_depth, hyp, hp, wp, H, W, h, w, D, 0.f, depth_values, Dv,
                                                       interval_ratio);
  RCMVS_CHECK_LAUNCH("rcmvs_depth_hypotheses_next_dev");
  return RCMVS_OK;
}
