// f3: geometric-consistency filter of the reference's fusion step (eval_rcmvsnet_dtu.py:281-336 reproject_with_depth /
// check_geometric_consistency and :366-389 of filter_depth), all NS source views of one reference view in ONE launch.
//
// The reference (numpy + cv2, one process-pool task per scan) does, per (ref, src) pair: back-project every reference
// pixel with its depth, project into the source view, sample the source depth map with cv2.remap(INTER_LINEAR),
// back-project with the sampled depth, project into the reference view again and keep the pixel iff the round trip
// moved it by < img_dist_thresh px and changed its depth by < depth_thresh relative.  Then, over the sources:
// geo_mask_sum, averaged depth (Σ reprojected depths + ref depth) / (geo_mask_sum + 1), final = photo & geo masks.
//
// Arithmetic follows the reference's numpy promotion rules so that the masks are reproducible: the four 3x3 / 4x4
// matrices per pair (K_ref^-1, E_src·E_ref^-1, K_src^-1, E_ref·E_src^-1) are float32 (host, numpy float32 inverse as in
// the reference) and every per-pixel product is float64 (int64 pixel grids * float32 depth promote to float64 in
// numpy); x_src / y_src are rounded to float32 only for the remap, the unrounded float64 values are re-projected
// (:300, :306 of the reference).  cv2.remap for float maps is restated from OpenCV's algorithm: coordinates are
// quantised to 1/32 px (cvRound(x*32), round-half-even), the 2x2 weights come from the 32x32 float table
// (1-f/32)*(1-g/32)..., taps outside the image contribute the constant border value 0, the products are summed left to
// right in float.
//
// One thread per reference pixel, looping over the sources: the reference depth is read once, every source costs 4
// depth taps.  Bytes: read 4·H·W·(1 + NS) (+ confidence), write ≤ 4·H·W·(2 + 2·NS): HBM-trivial next to the inference
// that produced the maps; the point is removing the seconds of single-threaded numpy per reference view.
#include "common.cuh"

namespace rcmvs {

struct GeoPair {
  float Kri[9];    // inv(K_ref)            float32
  float T_rs[16];  // E_src @ inv(E_ref)    float32
  float Ks[9];     // K_src
  float Ksi[9];    // inv(K_src)            float32
  float T_sr[16];  // E_ref @ inv(E_src)    float32
  float Kr[9];     // K_ref
};

// cvRound(float): round to nearest even; out-of-range / NaN -> INT_MIN like the SSE conversion OpenCV uses
__device__ __forceinline__ int cv_round(float v) {
  return (fabsf(v) < 2147483648.f) ? __float2int_rn(v) : (int)0x80000000;
}
__device__ __forceinline__ int sat_short(int v) { return min(max(v, -32768), 32767); }

// cv2.remap(src, x, y, INTER_LINEAR) for one sample (BORDER_CONSTANT, value 0), src [Hs][Ws] float32
__device__ __forceinline__ float cv_remap_linear(const float* __restrict__ src, int Hs, int Ws, float x, float y) {
  const int sx = cv_round(x * 32.f), sy = cv_round(y * 32.f);
  const int ix = sat_short(sx >> 5), iy = sat_short(sy >> 5);
  const float fx = (float)(sx & 31) * (1.f / 32.f), fy = (float)(sy & 31) * (1.f / 32.f);
  const float wx0 = 1.f - fx, wy0 = 1.f - fy;
  const float w0 = __fmul_rn(wy0, wx0), w1 = __fmul_rn(wy0, fx), w2 = __fmul_rn(fy, wx0), w3 = __fmul_rn(fy, fx);
  const bool x0ok = (unsigned)ix < (unsigned)Ws, x1ok = (unsigned)(ix + 1) < (unsigned)Ws;
  const bool y0ok = (unsigned)iy < (unsigned)Hs, y1ok = (unsigned)(iy + 1) < (unsigned)Hs;
  const float v0 = (x0ok && y0ok) ? __ldg(src + (int64_t)iy * Ws + ix) : 0.f;
  const float v1 = (x1ok && y0ok) ? __ldg(src + (int64_t)iy * Ws + ix + 1) : 0.f;
  const float v2 = (x0ok && y1ok) ? __ldg(src + (int64_t)(iy + 1) * Ws + ix) : 0.f;
  const float v3 = (x1ok && y1ok) ? __ldg(src + (int64_t)(iy + 1) * Ws + ix + 1) : 0.f;
  // left-to-right float sum of float products, no contraction (OpenCV's scalar remapBilinear)
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v0, w0), __fmul_rn(v1, w1)), __fmul_rn(v2, w2)), __fmul_rn(v3, w3));
}

__global__ void __launch_bounds__(128)
geo_consistency_kernel(const float* __restrict__ depth_ref, const float* __restrict__ conf,
                       const float* __restrict__ depth_src, const GeoPair* __restrict__ pairs, int NS, int H, int W,
                       int Hs, int Ws, float img_dist_thresh, float depth_thresh, float prob_thresh,
                       int num_consistent, uint8_t* __restrict__ masks, float* __restrict__ reproj,
                       float* __restrict__ xy_src, int32_t* __restrict__ geo_sum, float* __restrict__ depth_avg,
                       uint8_t* __restrict__ final_mask) {
  extern __shared__ GeoPair sp[];
  for (int i = threadIdx.x; i < NS * (int)(sizeof(GeoPair) / 4); i += blockDim.x)
    reinterpret_cast<float*>(sp)[i] = __ldg(reinterpret_cast<const float*>(pairs) + i);
  __syncthreads();
  const int x = blockIdx.x * 128 + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const int64_t pix = (int64_t)y * W + x, HW = (int64_t)H * W;
  const float dref = __ldg(depth_ref + pix);
  const double d = (double)dref, xd = (double)x * d, yd = (double)y * d;       // (x, y, 1) * depth, float64
  int nsum = 0;
  float dsum = 0.f;                       // the reference sums the float32 reprojected maps, then adds the ref depth
  for (int s = 0; s < NS; ++s) {
    const GeoPair& p = sp[s];
    // xyz_ref = inv(K_ref) @ (x·d, y·d, d)
    double r[3], q[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) r[i] = (double)p.Kri[i * 3] * xd + (double)p.Kri[i * 3 + 1] * yd + (double)p.Kri[i * 3 + 2] * d;
    // xyz_src = (E_src E_ref^-1) @ (xyz_ref, 1)   — numpy sums the 4 products of a row in order
#pragma unroll
    for (int i = 0; i < 3; ++i)
      q[i] = (double)p.T_rs[i * 4] * r[0] + (double)p.T_rs[i * 4 + 1] * r[1] + (double)p.T_rs[i * 4 + 2] * r[2] +
             (double)p.T_rs[i * 4 + 3];
#pragma unroll
    for (int i = 0; i < 3; ++i) r[i] = (double)p.Ks[i * 3] * q[0] + (double)p.Ks[i * 3 + 1] * q[1] + (double)p.Ks[i * 3 + 2] * q[2];
    const double xs = r[0] / r[2], ys = r[1] / r[2];
    const float xs_f = (float)xs, ys_f = (float)ys;
    const float sd = cv_remap_linear(depth_src + (int64_t)s * Hs * Ws, Hs, Ws, xs_f, ys_f);
    // xyz_src = inv(K_src) @ ((xs, ys, 1) * sampled depth)
    const double sdd = (double)sd, a0 = xs * sdd, a1 = ys * sdd;
#pragma unroll
    for (int i = 0; i < 3; ++i) q[i] = (double)p.Ksi[i * 3] * a0 + (double)p.Ksi[i * 3 + 1] * a1 + (double)p.Ksi[i * 3 + 2] * sdd;
#pragma unroll
    for (int i = 0; i < 3; ++i)
      r[i] = (double)p.T_sr[i * 4] * q[0] + (double)p.T_sr[i * 4 + 1] * q[1] + (double)p.T_sr[i * 4 + 2] * q[2] +
             (double)p.T_sr[i * 4 + 3];
    const float drep = (float)r[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) q[i] = (double)p.Kr[i * 3] * r[0] + (double)p.Kr[i * 3 + 1] * r[1] + (double)p.Kr[i * 3 + 2] * r[2];
    const float xr = (float)(q[0] / q[2]), yr = (float)(q[1] / q[2]);
    // dist: float32 minus int64 grid -> float64 in numpy; relative depth difference stays float32
    const double ex = (double)xr - (double)x, ey = (double)yr - (double)y;
    const double dist = sqrt(ex * ex + ey * ey);
    const float rel = __fdiv_rn(fabsf(__fsub_rn(drep, dref)), dref);
    const bool ok = (dist < (double)img_dist_thresh) && (rel < depth_thresh);     // NaN compares false, as in numpy
    const float dkeep = ok ? drep : 0.f;
    nsum += ok ? 1 : 0;
    dsum = (s == 0) ? dkeep : __fadd_rn(dsum, dkeep);
    if (masks) masks[(int64_t)s * HW + pix] = ok ? 1 : 0;
    if (reproj) reproj[(int64_t)s * HW + pix] = dkeep;
    if (xy_src) {
      xy_src[((int64_t)s * 2) * HW + pix] = xs_f;
      xy_src[((int64_t)s * 2 + 1) * HW + pix] = ys_f;
    }
  }
  if (geo_sum) geo_sum[pix] = nsum;
  // (sum(all_srcview_depth_ests) + ref_depth_est) / (geo_mask_sum + 1): float32 / int32 -> float64 in numpy
  if (depth_avg) depth_avg[pix] = (float)((double)__fadd_rn(dsum, dref) / (double)(nsum + 1));
  if (final_mask) {
    const bool photo = conf ? (__ldg(conf + pix) > prob_thresh) : true;
    final_mask[pix] = (photo && nsum >= num_consistent) ? 1 : 0;
  }
}

}  // namespace rcmvs

extern "C" int rcmvs_geo_consistency(const float* depth_ref, const float* conf, const float* depth_src,
                                     const float* pair_mats, int NS, int H, int W, int Hs, int Ws,
                                     float img_dist_thresh, float depth_thresh, float prob_thresh, int num_consistent,
                                     uint8_t* masks, float* depth_reproj, float* xy_src, int32_t* geo_mask_sum,
                                     float* depth_avg, uint8_t* final_mask, rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(depth_ref && depth_src && pair_mats, RCMVS_ERR_BAD_ARG, "rcmvs_geo_consistency: null pointer");
  RCMVS_REQUIRE(NS > 0 && H > 0 && W > 0 && Hs > 0 && Ws > 0 && H <= 65535, RCMVS_ERR_BAD_ARG,
                "rcmvs_geo_consistency: bad dims");
  RCMVS_REQUIRE(NS * sizeof(GeoPair) <= 48 * 1024, RCMVS_ERR_UNSUPPORTED, "rcmvs_geo_consistency: at most %d sources",
                (int)(48 * 1024 / sizeof(GeoPair)));
  static_assert(sizeof(GeoPair) == 68 * 4, "GeoPair layout = 68 floats per pair (include/rcmvs.h)");
  dim3 grid(ceil_div(W, 128), H);
  geo_consistency_kernel<<<grid, 128, NS * sizeof(GeoPair), as_stream(stream)>>>(
      depth_ref, conf, depth_src, reinterpret_cast<const GeoPair*>(pair_mats), NS, H, W, Hs, Ws, img_dist_thresh,
      depth_thresh, prob_thresh, num_consistent, masks, depth_reproj, xy_src, geo_mask_sum, depth_avg, final_mask);
  RCMVS_CHECK_LAUNCH("rcmvs_geo_consistency");
  return RCMVS_OK;
}
