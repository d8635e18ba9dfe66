// f2: the unsupervised photometric loss of one cascade stage, value AND gradient w.r.t. depth in one C call
// (reference losses/unsup_loss.py:9-94 UnSupLoss.forward; losses/modules.py SSIM :6-41, depth_smoothness :56-69,
// compute_reconstr_loss :72-82; losses/homography.py inverse_warping).
//
// The reference evaluates, per stage and per source view, inverse_warping (~60 tiny kernels), two masked smooth-L1
// terms on image gradients, an SSIM with six 3x3 pools, a top-1 selection over views and a smoothness term, then
// autograd walks all of it backwards: several hundred launches of kernels that each touch a few MB.  Because every
// term reduces to a SCALAR (the per-view reconstruction loss is mean-reduced BEFORE the per-pixel top-1, :66-68), the
// whole thing is: warp, three sets of per-block partial sums, two tiny finalisation kernels, and a backward that is a
// 3x3 gather per pixel followed by the warp's depth gradient.  11 + 4(V-1) launches per stage, fixed-order (two-level)
// reductions -> deterministic.
//
//   total = 12·reconstr + 6·ssim + smooth_weight·smooth
//   L_v      = 0.5·mean sl1(e) + 0.5·(mean sl1(∂x e) + mean sl1(∂y e)),  e = (warped_v − ref)·mask_v
//   reconstr = mean_p [ min_v (L_v + 1e4·(1−mask_v(p))) if < 1e4 else 0 ]
//   ssim     = Σ_{v<3} mean( pool(mask_v) · clamp((1 − SSIM(ref, warped_v))/2, 0, 1) )     (3x3 valid pools)
//   smooth   = mean |∂x depth · exp(−mean_c|∂x ref|)| + mean |∂y depth · exp(−mean_c|∂y ref|)|
#include "inverse_warp.cuh"

namespace rcmvs {
namespace loss {

constexpr int kMaxSrc = 15;
constexpr int TX = 128;

__device__ __forceinline__ float sl1(float d) { const float a = fabsf(d); return a < 1.f ? 0.5f * d * d : a - 0.5f; }
__device__ __forceinline__ float sl1_grad(float d) { return fminf(fmaxf(d, -1.f), 1.f); }

// block-wide sum of N values per thread -> thread 0 holds the totals (fixed order: deterministic)
template <int N, int WARPS = 4>
__device__ __forceinline__ void block_sum(float (&v)[N], float* sm /* [N][WARPS] */) {
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < N; ++i) sm[i * WARPS + warp] = v[i];
  __syncthreads();
  if (threadIdx.x == 0)
#pragma unroll
    for (int i = 0; i < N; ++i) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < WARPS; ++k) t += sm[i * WARPS + k];
      v[i] = t;
    }
}

// ---- per view: partial sums of the photo / gradient / SSIM terms + the SSIM backward coefficients -----------------
// imgs [B][V][h][w][3]; warped [B][h][w][3], mask [B][h][w] of THIS view; coef [B][h-2][w-2][3][3] or null (v >= 3)
__global__ void __launch_bounds__(TX)
terms_kernel(const float* __restrict__ imgs, const float* __restrict__ warped, const float* __restrict__ mask,
             float* __restrict__ coef, float* __restrict__ partial /* [NB][4] */, int V, int h, int w) {
  __shared__ float sm[16];
  const int x = blockIdx.x * TX + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
  const int64_t hw = (int64_t)h * w;
  const float* ref = imgs + (int64_t)b * V * hw * 3;
  const float* wp = warped + (int64_t)b * hw * 3;
  const float* mk = mask + (int64_t)b * hw;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};       // photo, grad-x, grad-y, ssim
  if (x < w) {
    const int64_t p = (int64_t)y * w + x;
    const float m0 = mk[p];
    float e0[3], ex[3], ey[3];
    const bool hx = x + 1 < w, hy = y + 1 < h;
    const float mx = hx ? mk[p + 1] : 0.f, my = hy ? mk[p + w] : 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      e0[c] = (wp[p * 3 + c] - ref[p * 3 + c]) * m0;
      ex[c] = hx ? (wp[(p + 1) * 3 + c] - ref[(p + 1) * 3 + c]) * mx : 0.f;
      ey[c] = hy ? (wp[(p + w) * 3 + c] - ref[(p + w) * 3 + c]) * my : 0.f;
      acc[0] += sl1(e0[c]);
      if (hx) acc[1] += sl1(ex[c] - e0[c]);
      if (hy) acc[2] += sl1(ey[c] - e0[c]);
    }
    if (coef && x + 2 < w && y + 2 < h) {
      // 3x3 window with top-left corner (x, y)
      float pm = 0.f;
#pragma unroll
      for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i) pm += mk[p + j * w + i];
      pm *= (1.f / 9.f);
      float* cq = coef + (((int64_t)b * (h - 2) + y) * (w - 2) + x) * 9;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const float xv = ref[(p + j * w + i) * 3 + c], yv = wp[(p + j * w + i) * 3 + c];
            sx += xv; sy += yv; sxx = fmaf(xv, xv, sxx); syy = fmaf(yv, yv, syy); sxy = fmaf(xv, yv, sxy);
          }
        const float mu_x = sx * (1.f / 9.f), mu_y = sy * (1.f / 9.f);
        const float vx = sxx * (1.f / 9.f) - mu_x * mu_x, vy = syy * (1.f / 9.f) - mu_y * mu_y;
        const float cxy = sxy * (1.f / 9.f) - mu_x * mu_y;
        constexpr float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
        const float A1 = 2.f * mu_x * mu_y + C1, A2 = 2.f * cxy + C2;
        const float B1 = mu_x * mu_x + mu_y * mu_y + C1, B2 = vx + vy + C2;
        const float inv = 1.f / (B1 * B2);
        const float S = A1 * A2 * inv;
        const float val = 0.5f * (1.f - S);
        acc[3] += pm * fminf(fmaxf(val, 0.f), 1.f);
        // d(pm·clamp(val))/dS = -0.5·pm inside the clamp range (torch.clamp passes the gradient on the closed interval)
        const float dS = (val >= 0.f && val <= 1.f) ? -0.5f * pm : 0.f;
        const float dS_dmu = (2.f * mu_x * A2 - 2.f * mu_x * A1) * inv - S * (2.f * mu_y / B1 - 2.f * mu_y / B2);
        const float dS_dey2 = -S / B2;
        const float dS_dexy = 2.f * A1 * inv;
        cq[c * 3 + 0] = dS * dS_dmu * (1.f / 9.f);
        cq[c * 3 + 1] = dS * dS_dey2 * (2.f / 9.f);
        cq[c * 3 + 2] = dS * dS_dexy * (1.f / 9.f);
      }
    }
  }
  block_sum<4>(acc, sm);
  if (threadIdx.x == 0) {
    const int64_t blk = ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) partial[blk * 4 + i] = acc[i];
  }
}

// scal layout (floats): [0..15) L_v, [16..31) ssim_v, [32..47) g_photo_v, [48..63) g_gx_v, [64..79) g_gy_v, [80] g_ssim
__global__ void __launch_bounds__(256)
finalize_terms_kernel(const float* __restrict__ partial /* [NV][NB][4] */, float* __restrict__ scal, int NV, int64_t NB,
                      int B, int h, int w) {
  __shared__ double sm[4][256];
  for (int v = 0; v < NV; ++v) {
    double a[4] = {0, 0, 0, 0};
    for (int64_t i = threadIdx.x; i < NB; i += 256)
#pragma unroll
      for (int k = 0; k < 4; ++k) a[k] += (double)partial[((int64_t)v * NB + i) * 4 + k];
#pragma unroll
    for (int k = 0; k < 4; ++k) sm[k][threadIdx.x] = a[k];
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if (threadIdx.x < s)
#pragma unroll
        for (int k = 0; k < 4; ++k) sm[k][threadIdx.x] += sm[k][threadIdx.x + s];
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      const double n0 = (double)B * h * w * 3, nx = (double)B * h * (w - 1) * 3, ny = (double)B * (h - 1) * w * 3;
      const double ns = (double)B * (h - 2) * (w - 2) * 3;
      scal[v] = (float)(0.5 * (sm[0][0] / n0) + 0.5 * (sm[1][0] / nx + sm[2][0] / ny));
      scal[16 + v] = (float)(sm[3][0] / ns);
    }
    __syncthreads();
  }
}

// ---- top-1 over views (per-pixel argmin of the scalar L_v among valid views) + smoothness partial sums --------------
__global__ void __launch_bounds__(TX)
select_smooth_kernel(const float* __restrict__ imgs, const float* __restrict__ masks /* [NV][B][h][w] */,
                     const float* __restrict__ depth, const float* __restrict__ scal,
                     float* __restrict__ partial /* [NB][NV + 2] */, int V, int B, int h, int w) {
  __shared__ float sm[(kMaxSrc + 2) * 4];
  const int NV = V - 1;
  const int x = blockIdx.x * TX + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
  const int64_t hw = (int64_t)h * w;
  float cnt[kMaxSrc + 2];
#pragma unroll
  for (int i = 0; i < kMaxSrc + 2; ++i) cnt[i] = 0.f;
  if (x < w) {
    const int64_t p = (int64_t)y * w + x;
    int best = -1;
    float bl = 1e4f;
    for (int v = 0; v < NV; ++v) {
      // value = L_v + 1e4·(1 − mask): a masked-out view can never be < 1e4 (L_v >= 0)
      const float val = __ldg(scal + v) + 1e4f * (1.f - masks[((int64_t)v * B + b) * hw + p]);
      if (val < bl) { bl = val; best = v; }
    }
#pragma unroll
    for (int v = 0; v < kMaxSrc; ++v) cnt[v] = (v == best) ? 1.f : 0.f;
    const float* ref = imgs + (int64_t)b * V * hw * 3;
    const float* dp = depth + (int64_t)b * hw;
    const float d0 = dp[p];
    if (x + 1 < w) {
      const float a = (fabsf(ref[p * 3] - ref[(p + 1) * 3]) + fabsf(ref[p * 3 + 1] - ref[(p + 1) * 3 + 1]) +
                       fabsf(ref[p * 3 + 2] - ref[(p + 1) * 3 + 2])) * (1.f / 3.f);
      cnt[kMaxSrc] = fabsf((d0 - dp[p + 1]) * expf(-a));
    }
    if (y + 1 < h) {
      const float a = (fabsf(ref[p * 3] - ref[(p + w) * 3]) + fabsf(ref[p * 3 + 1] - ref[(p + w) * 3 + 1]) +
                       fabsf(ref[p * 3 + 2] - ref[(p + w) * 3 + 2])) * (1.f / 3.f);
      cnt[kMaxSrc + 1] = fabsf((d0 - dp[p + w]) * expf(-a));
    }
  }
  block_sum<kMaxSrc + 2>(cnt, sm);
  if (threadIdx.x == 0) {
    const int64_t blk = ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    for (int v = 0; v < NV; ++v) partial[blk * (NV + 2) + v] = cnt[v];
    partial[blk * (NV + 2) + NV] = cnt[kMaxSrc];
    partial[blk * (NV + 2) + NV + 1] = cnt[kMaxSrc + 1];
  }
}

__global__ void __launch_bounds__(256)
finalize_total_kernel(const float* __restrict__ partial, float* __restrict__ scal, float* __restrict__ terms, int NV,
                      int64_t NB, int B, int h, int w, float smooth_weight) {
  __shared__ double sm[256];
  __shared__ double tot[kMaxSrc + 2];
  for (int k = 0; k < NV + 2; ++k) {
    double a = 0;
    for (int64_t i = threadIdx.x; i < NB; i += 256) a += (double)partial[i * (NV + 2) + k];
    sm[threadIdx.x] = a;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
      __syncthreads();
    }
    if (threadIdx.x == 0) tot[k] = sm[0];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double npix = (double)B * h * w;
    double rec = 0, ss = 0;
    for (int v = 0; v < NV; ++v) rec += tot[v] * (double)scal[v];
    rec /= npix;
    for (int v = 0; v < NV && v < 2; ++v) ss += (double)scal[16 + v];
    const double smooth = tot[NV] / ((double)B * h * (w - 1)) + tot[NV + 1] / ((double)B * (h - 1) * w);
    terms[0] = (float)(12.0 * rec + 6.0 * ss + (double)smooth_weight * smooth);
    terms[1] = (float)rec;
    terms[2] = (float)ss;
    terms[3] = (float)smooth;
    for (int v = 0; v < NV; ++v) {
      const double gl = 12.0 * tot[v] / npix;                       // d total / d L_v
      scal[32 + v] = (float)(gl * 0.5 / ((double)B * h * w * 3));
      scal[48 + v] = (float)(gl * 0.5 / ((double)B * h * (w - 1) * 3));
      scal[64 + v] = (float)(gl * 0.5 / ((double)B * (h - 1) * w * 3));
    }
    scal[80] = (float)(6.0 / ((double)B * (h - 2) * (w - 2) * 3));
  }
}

// ---- backward: d total / d warped_v -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(TX)
grad_warped_kernel(const float* __restrict__ imgs, const float* __restrict__ warped, const float* __restrict__ mask,
                   const float* __restrict__ coef, const float* __restrict__ scal, int v, float* __restrict__ gw, int V,
                   int h, int w) {
  const int x = blockIdx.x * TX + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
  if (x >= w) return;
  const int64_t hw = (int64_t)h * w, p = (int64_t)y * w + x;
  const float* ref = imgs + (int64_t)b * V * hw * 3;
  const float* wp = warped + (int64_t)b * hw * 3;
  const float* mk = mask + (int64_t)b * hw;
  const float gp = __ldg(scal + 32 + v), ggx = __ldg(scal + 48 + v), ggy = __ldg(scal + 64 + v);
  const float gs = coef ? __ldg(scal + 80) : 0.f;
  const float m0 = mk[p];
  const bool l = x > 0, r = x + 1 < w, u = y > 0, d = y + 1 < h;
  const float ml = l ? mk[p - 1] : 0.f, mr = r ? mk[p + 1] : 0.f, mu = u ? mk[p - w] : 0.f, md = d ? mk[p + w] : 0.f;
  float g[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float e0 = (wp[p * 3 + c] - ref[p * 3 + c]) * m0;
    float acc = gp * sl1_grad(e0);
    if (r) acc -= ggx * sl1_grad((wp[(p + 1) * 3 + c] - ref[(p + 1) * 3 + c]) * mr - e0);
    if (l) acc += ggx * sl1_grad(e0 - (wp[(p - 1) * 3 + c] - ref[(p - 1) * 3 + c]) * ml);
    if (d) acc -= ggy * sl1_grad((wp[(p + w) * 3 + c] - ref[(p + w) * 3 + c]) * md - e0);
    if (u) acc += ggy * sl1_grad(e0 - (wp[(p - w) * 3 + c] - ref[(p - w) * 3 + c]) * mu);
    g[c] = acc * m0;                                     // e = (warped − ref)·mask
  }
  if (coef) {
    // every 3x3 window containing p: top-left corners (x-2..x, y-2..y) clipped to the valid range
    const float* cb = coef + (int64_t)b * (h - 2) * (w - 2) * 9;
    float k0[3] = {0.f, 0.f, 0.f}, k1[3] = {0.f, 0.f, 0.f}, k2[3] = {0.f, 0.f, 0.f};
    for (int qy = max(y - 2, 0); qy <= min(y, h - 3); ++qy)
      for (int qx = max(x - 2, 0); qx <= min(x, w - 3); ++qx) {
        const float* cq = cb + ((int64_t)qy * (w - 2) + qx) * 9;
#pragma unroll
        for (int c = 0; c < 3; ++c) { k0[c] += cq[c * 3]; k1[c] += cq[c * 3 + 1]; k2[c] += cq[c * 3 + 2]; }
      }
#pragma unroll
    for (int c = 0; c < 3; ++c) g[c] += gs * (k0[c] + wp[p * 3 + c] * k1[c] + ref[p * 3 + c] * k2[c]);
  }
  float* o = gw + ((int64_t)b * hw + p) * 3;
  o[0] = g[0]; o[1] = g[1]; o[2] = g[2];
}

// d (smooth_weight·smooth) / d depth, written (not accumulated): the warp gradients of the views are added afterwards
__global__ void __launch_bounds__(TX)
smooth_grad_kernel(const float* __restrict__ imgs, const float* __restrict__ depth, float* __restrict__ gd, int V, int B,
                   int h, int w, float smooth_weight) {
  const int x = blockIdx.x * TX + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
  if (x >= w) return;
  const int64_t hw = (int64_t)h * w, p = (int64_t)y * w + x;
  const float* ref = imgs + (int64_t)b * V * hw * 3;
  const float* dp = depth + (int64_t)b * hw;
  const float cx = smooth_weight / ((float)B * h * (w - 1)), cy = smooth_weight / ((float)B * (h - 1) * w);
  auto edge = [&](int64_t a, int64_t bb) {      // d |(d_a − d_b)·w_ab| / d d_a
    const float wgt = expf(-(fabsf(ref[a * 3] - ref[bb * 3]) + fabsf(ref[a * 3 + 1] - ref[bb * 3 + 1]) +
                             fabsf(ref[a * 3 + 2] - ref[bb * 3 + 2])) * (1.f / 3.f));
    const float s = (dp[a] - dp[bb]) * wgt;
    return (s > 0.f ? 1.f : (s < 0.f ? -1.f : 0.f)) * wgt;
  };
  float g = 0.f;
  if (x + 1 < w) g += cx * edge(p, p + 1);
  if (x > 0) g -= cx * edge(p - 1, p);
  if (y + 1 < h) g += cy * edge(p, p + w);
  if (y > 0) g -= cy * edge(p - w, p);
  gd[(int64_t)b * hw + p] = g;
}

// [N][3][H][W] -> [N][H/step][W/step][3]: F.interpolate(scale 1/step, nearest, recompute_scale_factor) = strided pick
__global__ void subsample_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int H, int W, int h, int w,
                                      int step, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % 3);
  const int64_t p = i / 3;
  const int x = (int)(p % w), y = (int)((p / w) % h);
  const int64_t n = p / ((int64_t)w * h);
  dst[i] = __ldg(src + ((n * 3 + c) * H + (int64_t)y * step) * W + (int64_t)x * step);
}

// ---- masked mean loss + gradient (AugLossMultiStage, SL1Loss, img2mse) ------------------------------------------------
// element i = (b, y, x) of pred [B][h][w]; target / mask are read at (y·step, x·step) of [B][h·step][w·step] planes
// (F.interpolate(nearest, 1/step) of the pseudo depth and of the filter mask, losses/aug_loss.py:44-52).
// mask_kind 0: all elements, 1: mask plane > 0.5, 2: target > 0 (SL1Loss default, losses/sl1loss.py:10-11)
__device__ __forceinline__ bool ml_active(int kind, float tgt, const float* mask, int64_t mi) {
  return kind == 0 ? true : (kind == 1 ? __ldg(mask + mi) > 0.5f : tgt > 0.f);
}
__global__ void __launch_bounds__(256)
masked_loss_partial_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                           const float* __restrict__ mask, int64_t mask_bstride, int kind, int mode, int h, int w,
                           int step, int64_t n, float* __restrict__ partial /* [gridDim.x][2] */) {
  __shared__ float sm[16];
  float acc[2] = {0.f, 0.f};
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const int x = (int)(i % w), y = (int)((i / w) % h);
    const int64_t b = i / ((int64_t)w * h);
    const int64_t ti = (b * h * step + (int64_t)y * step) * ((int64_t)w * step) + (int64_t)x * step;
    const float t = __ldg(target + ti);
    if (ml_active(kind, t, mask, b * mask_bstride + ((int64_t)y * step) * ((int64_t)w * step) + (int64_t)x * step)) {
      const float d = pred[i] - t;
      acc[0] += mode == 0 ? sl1(d) : d * d;
      acc[1] += 1.f;
    }
  }
  block_sum<2, 8>(acc, sm);
  if (threadIdx.x == 0) { partial[blockIdx.x * 2] = acc[0]; partial[blockIdx.x * 2 + 1] = acc[1]; }
}
__global__ void masked_loss_finalize_kernel(const float* __restrict__ partial, int nblocks, float scale,
                                            float* __restrict__ loss, float* __restrict__ coef) {
  if (threadIdx.x != 0) return;
  double s = 0, c = 0;
  for (int i = 0; i < nblocks; ++i) { s += (double)partial[i * 2]; c += (double)partial[i * 2 + 1]; }
  loss[0] = (float)((double)scale * s / c);            // empty selection -> NaN, like torch's mean of nothing
  coef[0] = (float)((double)scale / c);
}
__global__ void __launch_bounds__(256)
masked_loss_grad_kernel(const float* __restrict__ pred, const float* __restrict__ target, const float* __restrict__ mask,
                        int64_t mask_bstride, int kind, int mode, int h, int w, int step, int64_t n,
                        const float* __restrict__ coef, float* __restrict__ grad) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const int x = (int)(i % w), y = (int)((i / w) % h);
  const int64_t b = i / ((int64_t)w * h);
  const int64_t ti = (b * h * step + (int64_t)y * step) * ((int64_t)w * step) + (int64_t)x * step;
  const float t = __ldg(target + ti);
  float g = 0.f;
  if (ml_active(kind, t, mask, b * mask_bstride + ((int64_t)y * step) * ((int64_t)w * step) + (int64_t)x * step)) {
    const float d = pred[i] - t;
    g = __ldg(coef) * (mode == 0 ? sl1_grad(d) : 2.f * d);
  }
  grad[i] = g;
}

struct Layout {
  int64_t warped, mask, gw, coef, part, scal, total;
};
static Layout layout(int B, int V, int h, int w) {
  const int NV = V - 1;
  const int64_t hw = (int64_t)h * w, NB = (int64_t)ceil_div(w, TX) * h * B;
  Layout L;
  int64_t o = 0;
  L.warped = o; o += (int64_t)NV * B * hw * 3;
  L.mask = o; o += (int64_t)NV * B * hw;
  L.gw = o; o += (int64_t)B * hw * 3;
  L.coef = o; o += (int64_t)2 * B * (h - 2) * (w - 2) * 9;
  L.part = o; o += NB * 4 * NV > NB * (NV + 2) ? NB * 4 * NV : NB * (NV + 2);
  L.scal = o; o += 96;
  L.total = o;
  return L;
}

}  // namespace loss
}  // namespace rcmvs

extern "C" int64_t rcmvs_unsup_loss_workspace_floats(int B, int V, int h, int w) {
  if (B <= 0 || V < 2 || h < 3 || w < 3) return 0;
  return rcmvs::loss::layout(B, V, h, w).total;
}

extern "C" int rcmvs_images_subsample_nhwc(const float* imgs, float* out, int N, int H, int W, int step,
                                           rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(imgs && out, RCMVS_ERR_BAD_ARG, "rcmvs_images_subsample_nhwc: null pointer");
  RCMVS_REQUIRE(N > 0 && H > 0 && W > 0 && step > 0, RCMVS_ERR_BAD_ARG, "rcmvs_images_subsample_nhwc: bad dims");
  const int h = (H + step - 1) / step, w = (W + step - 1) / step;        // floor(H * (1/step)) for H % step == 0
  RCMVS_REQUIRE(H % step == 0 && W % step == 0, RCMVS_ERR_UNSUPPORTED,
                "rcmvs_images_subsample_nhwc: H, W must be multiples of the step");
  const int64_t total = (int64_t)N * h * w * 3;
  loss::subsample_nhwc_kernel<<<(unsigned)ceil_div<int64_t>(total, 256), 256, 0, as_stream(stream)>>>(imgs, out, H, W, h,
                                                                                                    w, step, total);
  RCMVS_CHECK_LAUNCH("rcmvs_images_subsample_nhwc");
  return RCMVS_OK;
}

extern "C" int rcmvs_unsup_loss_fwd_bwd(const float* imgs, const float* cams, const float* depth, float smooth_weight,
                                        float* terms, float* grad_depth, float* workspace, int B, int V, int h, int w,
                                        rcmvs_stream_t stream) {
  using namespace rcmvs;
  using namespace rcmvs::loss;
  RCMVS_REQUIRE(imgs && cams && depth && terms && workspace, RCMVS_ERR_BAD_ARG, "rcmvs_unsup_loss_fwd_bwd: null pointer");
  RCMVS_REQUIRE(B > 0 && V >= 2 && V - 1 <= kMaxSrc && h >= 3 && w >= 3 && h <= 65535 && B <= 65535, RCMVS_ERR_BAD_ARG,
                "rcmvs_unsup_loss_fwd_bwd: bad dims B=%d V=%d h=%d w=%d", B, V, h, w);
  cudaStream_t st = as_stream(stream);
  const Layout L = layout(B, V, h, w);
  const int NV = V - 1;
  const int64_t hw = (int64_t)h * w, NB = (int64_t)ceil_div(w, TX) * h * B;
  float* ws = workspace;
  const dim3 grid(ceil_div(w, TX), h, B);
  const int64_t cam_stride = (int64_t)V * 32, img_stride = (int64_t)V * hw * 3;
  for (int v = 0; v < NV; ++v) {
    float* wv = ws + L.warped + (int64_t)v * B * hw * 3;
    float* mv = ws + L.mask + (int64_t)v * B * hw;
    // source image = view v+1 of every batch item, cameras: left = view 0, right = view v+1
    inverse_warp_kernel<false><<<grid, 128, 0, st>>>(imgs + (int64_t)(v + 1) * hw * 3, cams, cams + (int64_t)(v + 1) * 32,
                                                     depth, wv, mv, nullptr, nullptr, h, w, 3, cam_stride, img_stride);
    float* cf = v < 2 ? ws + L.coef + (int64_t)v * B * (h - 2) * (w - 2) * 9 : nullptr;
    terms_kernel<<<grid, TX, 0, st>>>(imgs, wv, mv, cf, ws + L.part + (int64_t)v * NB * 4, V, h, w);
  }
  finalize_terms_kernel<<<1, 256, 0, st>>>(ws + L.part, ws + L.scal, NV, NB, B, h, w);
  select_smooth_kernel<<<grid, TX, 0, st>>>(imgs, ws + L.mask, depth, ws + L.scal, ws + L.part, V, B, h, w);
  finalize_total_kernel<<<1, 256, 0, st>>>(ws + L.part, ws + L.scal, terms, NV, NB, B, h, w, smooth_weight);
  if (grad_depth) {
    smooth_grad_kernel<<<grid, TX, 0, st>>>(imgs, depth, grad_depth, V, B, h, w, smooth_weight);
    for (int v = 0; v < NV; ++v) {
      float* wv = ws + L.warped + (int64_t)v * B * hw * 3;
      float* mv = ws + L.mask + (int64_t)v * B * hw;
      float* cf = v < 2 ? ws + L.coef + (int64_t)v * B * (h - 2) * (w - 2) * 9 : nullptr;
      grad_warped_kernel<<<grid, TX, 0, st>>>(imgs, wv, mv, cf, ws + L.scal, v, ws + L.gw, V, h, w);
      inverse_warp_kernel<true, true><<<grid, 128, 0, st>>>(imgs + (int64_t)(v + 1) * hw * 3, cams,
                                                            cams + (int64_t)(v + 1) * 32, depth, nullptr, nullptr,
                                                            ws + L.gw, grad_depth, h, w, 3, cam_stride, img_stride);
    }
  }
  RCMVS_CHECK_LAUNCH("rcmvs_unsup_loss_fwd_bwd");
  return RCMVS_OK;
}

extern "C" int64_t rcmvs_masked_loss_workspace_floats(void) { return 2 * 1184 + 8; }

extern "C" int rcmvs_masked_loss_fwd_bwd(const float* pred, const float* target, const float* mask,
                                         int64_t mask_batch_stride, int mask_kind, int mode, float scale, float* loss,
                                         float* grad, float* workspace, int B, int h, int w, int step,
                                         rcmvs_stream_t stream) {
  using namespace rcmvs;
  using namespace rcmvs::loss;
  RCMVS_REQUIRE(pred && target && loss && workspace, RCMVS_ERR_BAD_ARG, "rcmvs_masked_loss_fwd_bwd: null pointer");
  RCMVS_REQUIRE(B > 0 && h > 0 && w > 0 && step > 0 && mask_kind >= 0 && mask_kind <= 2 && (mode == 0 || mode == 1),
                RCMVS_ERR_BAD_ARG, "rcmvs_masked_loss_fwd_bwd: bad arguments");
  RCMVS_REQUIRE(mask_kind != 1 || mask, RCMVS_ERR_BAD_ARG, "rcmvs_masked_loss_fwd_bwd: mask_kind 1 needs a mask");
  const int64_t n = (int64_t)B * h * w;
  const int64_t want_blocks = ceil_div<int64_t>(n, 256);
  const int nblocks = (int)(want_blocks < 1184 ? want_blocks : 1184);      // <= 8 CTAs on each of 148 SMs
  cudaStream_t st = as_stream(stream);
  masked_loss_partial_kernel<<<nblocks, 256, 0, st>>>(pred, target, mask, mask_batch_stride, mask_kind, mode, h, w, step,
                                                      n, workspace);
  float* coef = workspace + 2 * 1184;
  masked_loss_finalize_kernel<<<1, 32, 0, st>>>(workspace, nblocks, scale, loss, coef);
  if (grad)
    masked_loss_grad_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, st>>>(pred, target, mask, mask_batch_stride,
                                                                               mask_kind, mode, h, w, step, n, coef, grad);
  RCMVS_CHECK_LAUNCH("rcmvs_masked_loss_fwd_bwd");
  return RCMVS_OK;
}
