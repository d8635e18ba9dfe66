// a13: losses.homography.inverse_warping (reference losses/homography.py:6-200), forward and the gradient w.r.t.
// depth (the unsupervised photometric training signal).  One thread per reference pixel; the ~60 tiny torch kernels
// (+ torch.inverse) of the reference collapse into one launch.  Bug-compatible on purpose: the source pixel is
// projected with K_left (:55-57), bilinear weights use the CLAMPED corners (:191-194) and the validity mask is
// x0>=0 & x1<=W-1 & y0>=0 & y0<=H-1 (:155).
#include "inverse_warp.cuh"

extern "C" int rcmvs_inverse_warp_fwd(const float* img, const float* left_cam, const float* right_cam,
                                      const float* depth, float* warped, float* mask, int B, int h, int w, int C,
                                      rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(img && left_cam && right_cam && depth && warped && mask, RCMVS_ERR_BAD_ARG,
                "rcmvs_inverse_warp_fwd: null pointer");
  RCMVS_REQUIRE(B > 0 && h > 0 && w > 0 && C > 0 && h <= 65535 && B <= 65535, RCMVS_ERR_BAD_ARG,
                "rcmvs_inverse_warp_fwd: bad dims");
  dim3 grid(ceil_div(w, 128), h, B);
  inverse_warp_kernel<false><<<grid, 128, 0, as_stream(stream)>>>(img, left_cam, right_cam, depth, warped, mask,
                                                                  nullptr, nullptr, h, w, C, 32, (int64_t)h * w * C);
  RCMVS_CHECK_LAUNCH("rcmvs_inverse_warp_fwd");
  return RCMVS_OK;
}

extern "C" int rcmvs_inverse_warp_bwd(const float* img, const float* left_cam, const float* right_cam,
                                      const float* depth, const float* grad_warped, float* grad_depth, int B, int h,
                                      int w, int C, rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(img && left_cam && right_cam && depth && grad_warped && grad_depth, RCMVS_ERR_BAD_ARG,
                "rcmvs_inverse_warp_bwd: null pointer");
  RCMVS_REQUIRE(B > 0 && h > 0 && w > 0 && C > 0 && h <= 65535 && B <= 65535, RCMVS_ERR_BAD_ARG,
                "rcmvs_inverse_warp_bwd: bad dims");
  dim3 grid(ceil_div(w, 128), h, B);
  inverse_warp_kernel<true><<<grid, 128, 0, as_stream(stream)>>>(img, left_cam, right_cam, depth, nullptr, nullptr,
                                                                 grad_warped, grad_depth, h, w, C, 32, (int64_t)h * w * C);
  RCMVS_CHECK_LAUNCH("rcmvs_inverse_warp_bwd");
  return RCMVS_OK;
}
