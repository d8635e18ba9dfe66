// a13 device code shared by inverse_warp.cu (the C entry points) and unsup_loss.cu (the fused f2 loss): camera set-up and
// the per-pixel warp kernel of losses.homography.inverse_warping (reference losses/homography.py:6-200).
#pragma once
#include "common.cuh"

namespace rcmvs {

struct IwCam {
  float A[9];   // K_l · R_rel · K_l^-1   (applied to (x, y, 1)·depth)
  float b[3];   // K_l · t_rel
};

__device__ inline void iw_setup(const float* __restrict__ lc, const float* __restrict__ rc, IwCam& c) {
  // lc, rc: [2][4][4]  ([0] = extrinsic, [1][:3,:3] = intrinsic)
  float Kl[9], Rl[9], Rr[9], tl[3], tr[3];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) { Kl[i * 3 + j] = lc[16 + i * 4 + j]; Rl[i * 3 + j] = lc[i * 4 + j]; Rr[i * 3 + j] = rc[i * 4 + j]; }
    tl[i] = lc[i * 4 + 3];
    tr[i] = rc[i * 4 + 3];
  }
  // K_l^-1 by cofactors (fp64 internally)
  double k[9];
  for (int i = 0; i < 9; ++i) k[i] = Kl[i];
  const double det = k[0] * (k[4] * k[8] - k[5] * k[7]) - k[1] * (k[3] * k[8] - k[5] * k[6]) + k[2] * (k[3] * k[7] - k[4] * k[6]);
  double ki[9] = {(k[4] * k[8] - k[5] * k[7]) / det, (k[2] * k[7] - k[1] * k[8]) / det, (k[1] * k[5] - k[2] * k[4]) / det,
                  (k[5] * k[6] - k[3] * k[8]) / det, (k[0] * k[8] - k[2] * k[6]) / det, (k[2] * k[3] - k[0] * k[5]) / det,
                  (k[3] * k[7] - k[4] * k[6]) / det, (k[1] * k[6] - k[0] * k[7]) / det, (k[0] * k[4] - k[1] * k[3]) / det};
  double Rrel[9], trel[3], KR[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int m = 0; m < 3; ++m) s += (double)Rr[i * 3 + m] * Rl[j * 3 + m];       // R_r · R_l^T
      Rrel[i * 3 + j] = s;
    }
  for (int i = 0; i < 3; ++i) {
    double s = 0;
    for (int m = 0; m < 3; ++m) s += Rrel[i * 3 + m] * tl[m];
    trel[i] = tr[i] - s;
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int m = 0; m < 3; ++m) s += k[i * 3 + m] * Rrel[m * 3 + j];
      KR[i * 3 + j] = s;
    }
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int m = 0; m < 3; ++m) s += KR[i * 3 + m] * ki[m * 3 + j];
      c.A[i * 3 + j] = (float)s;
    }
    double s = 0;
    for (int m = 0; m < 3; ++m) s += k[i * 3 + m] * trel[m];
    c.b[i] = (float)s;
  }
}

// cam_stride / img_stride: floats between consecutive batch items of lcam, rcam / img (the fused loss passes views of a
// [B][V][...] tensor); ACCUM: grad_depth += (the loss sums the contributions of all source views).
template <bool BWD, bool ACCUM = false>
__global__ void __launch_bounds__(128) inverse_warp_kernel(const float* __restrict__ img, const float* __restrict__ lcam,
                                                           const float* __restrict__ rcam, const float* __restrict__ depth,
                                                           float* __restrict__ warped, float* __restrict__ mask,
                                                           const float* __restrict__ gout, float* __restrict__ gdepth,
                                                           int h, int w, int C, int64_t cam_stride, int64_t img_stride) {
  __shared__ IwCam cam;
  const int b = blockIdx.z;
  if (threadIdx.x == 0) iw_setup(lcam + (int64_t)b * cam_stride, rcam + (int64_t)b * cam_stride, cam);
  __syncthreads();
  const int x = blockIdx.x * 128 + threadIdx.x, y = blockIdx.y;
  if (x >= w) return;
  const int64_t pix = ((int64_t)b * h + y) * w + x;
  const float d = __ldg(depth + pix);
  const float xf = (float)x, yf = (float)y;
  // p = (A·(x,y,1))·d + b
  const float ax = cam.A[0] * xf + cam.A[1] * yf + cam.A[2], ay = cam.A[3] * xf + cam.A[4] * yf + cam.A[5],
              az = cam.A[6] * xf + cam.A[7] * yf + cam.A[8];
  const float X = fmaf(ax, d, cam.b[0]), Y = fmaf(ay, d, cam.b[1]), Z = fmaf(az, d, cam.b[2]) + 1e-10f;
  const float px = X / Z, py = Y / Z;                                      // :91-100
  const float fx = floorf(px), fy = floorf(py);
  const bool fin = fabsf(px) < 1e8f && fabsf(py) < 1e8f;
  const int x0 = fin ? (int)fx : -4, y0 = fin ? (int)fy : -4;
  const int x1 = x0 + 1, y1 = y0 + 1;
  const float m = (x0 >= 0 && x1 <= w - 1 && y0 >= 0 && y0 <= h - 1) ? 1.f : 0.f;      // :155
  const int x0c = min(max(x0, 0), w - 1), x1c = min(max(x1, 0), w - 1);
  const int y0c = min(max(y0, 0), h - 1), y1c = min(max(y1, 0), h - 1);
  const float dx = fin ? (float)x1c - px : 0.f, dy = fin ? (float)y1c - py : 0.f;       // :191-194 (clamped corners)
  const float* base = img + (int64_t)b * img_stride;
  const float* pa = base + ((int64_t)y0c * w + x0c) * C;
  const float* pb = base + ((int64_t)y1c * w + x0c) * C;
  const float* pc = base + ((int64_t)y0c * w + x1c) * C;
  const float* pd = base + ((int64_t)y1c * w + x1c) * C;
  if (!BWD) {
    const float wa = dx * dy, wb = dx * (1.f - dy), wc = (1.f - dx) * dy, wd = (1.f - dx) * (1.f - dy);
    for (int c = 0; c < C; ++c)
      warped[pix * C + c] = wa * __ldg(pa + c) + wb * __ldg(pb + c) + wc * __ldg(pc + c) + wd * __ldg(pd + c);
    mask[pix] = m;
  } else {
    // d(out)/d(px) = -d(out)/d(dx),  d(out)/d(py) = -d(out)/d(dy)
    float gx = 0.f, gy = 0.f;
    for (int c = 0; c < C; ++c) {
      const float a = __ldg(pa + c), bb = __ldg(pb + c), cc = __ldg(pc + c), dd = __ldg(pd + c);
      const float g = __ldg(gout + pix * C + c);
      gx += g * -(dy * a + (1.f - dy) * bb - dy * cc - (1.f - dy) * dd);
      gy += g * -(dx * a - dx * bb + (1.f - dx) * cc - (1.f - dx) * dd);
    }
    // px = X/Z: dpx/dd = (ax·Z − X·az)/Z²
    const float iz = 1.f / Z;
    const float dpx = (ax - px * az) * iz, dpy = (ay - py * az) * iz;
    const float gd = fin ? gx * dpx + gy * dpy : 0.f;
    if (ACCUM) gdepth[pix] += gd; else gdepth[pix] = gd;
  }
}

}  // namespace rcmvs
