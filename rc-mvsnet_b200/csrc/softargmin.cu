// a4: softmax over depth + soft-argmin depth + photometric confidence in ONE kernel
// (reference models/casmvsnet.py:111-122 + models/modules.py:519-525: softmax, mul+sum, pad, avg_pool3d, gather =
// 5-6 passes over the probability volume plus a padded copy).  Two kernels: the wide one (below) for inference, and the
// one-thread-per-pixel kernel that can also write the probability volume (training) and takes any D: it reads the cost
// column three times from L2 (max, Σexp, Σp·d) and never writes the probability volume unless asked.
// HBM bytes: read D·h·w (+ D·h·w hypotheses), write 2·h·w.
#include "common.cuh"

namespace rcmvs {

template <bool PER_PIXEL>
__global__ void __launch_bounds__(128)
softargmin_conf_kernel(const float* __restrict__ cost, const float* __restrict__ planes, float* __restrict__ depth,
                       float* __restrict__ conf, int32_t* __restrict__ idx_out, float* __restrict__ prob, int D,
                       int64_t hw) {
  const int64_t p = (int64_t)blockIdx.x * 128 + threadIdx.x;
  const int b = blockIdx.y;
  if (p >= hw) return;
  cost += (int64_t)b * D * hw + p;
  planes += PER_PIXEL ? (int64_t)b * D * hw + p : (int64_t)b * D;
  if (prob) prob += (int64_t)b * D * hw + p;

  float m = -INFINITY;
#pragma unroll 4
  for (int d = 0; d < D; ++d) m = fmaxf(m, __ldg(cost + (int64_t)d * hw));
  float sum = 0.f;
#pragma unroll 4
  for (int d = 0; d < D; ++d) sum += expf(__ldg(cost + (int64_t)d * hw) - m);
  // p_d = exp(c_d - max) / sum ; depth = Σ p_d·depth_d ; fidx = Σ p_d·d   (ascending d, fp32)
  float dsum = 0.f, isum = 0.f;
#pragma unroll 4
  for (int d = 0; d < D; ++d) {
    const float pd = __fdiv_rn(expf(__ldg(cost + (int64_t)d * hw) - m), sum);
    const float dv = PER_PIXEL ? __ldg(planes + (int64_t)d * hw) : __ldg(planes + d);
    dsum = __fadd_rn(dsum, __fmul_rn(pd, dv));
    isum = __fadd_rn(isum, __fmul_rn(pd, (float)d));
    if (prob) prob[(int64_t)d * hw] = pd;
  }
  int idx = (int)isum;                       // .long(): truncation toward zero (casmvsnet.py:119-120)
  idx = min(max(idx, 0), D - 1);             // casmvsnet.py:121
  // conf = p[idx-1] + p[idx] + p[idx+1] + p[idx+2], zero outside [0,D)  (pad (1,2) + 4*avg_pool(4))
  float c4 = 0.f;
#pragma unroll
  for (int k = -1; k <= 2; ++k) {
    const int d = idx + k;
    if (d >= 0 && d < D) c4 += __fdiv_rn(expf(__ldg(cost + (int64_t)d * hw) - m), sum);
  }
  depth[(int64_t)b * hw + p] = dsum;
  conf[(int64_t)b * hw + p] = c4;
  if (idx_out) idx_out[(int64_t)b * hw + p] = idx;
}

// Wide variant: CTA = 32 pixels x NW warps; warp w owns planes d = w, w + NW, ... of the CTA's 32 pixels, so every load
// is a coalesced 128-byte row and the cost column is read from memory ONCE (kept in registers across the three
// reductions: max, Σ exp, Σ p·d / Σ p·i, which go through shared memory in ascending-warp order — deterministic).
// 8x the threads of the one-thread-per-pixel kernel: the op is latency-bound (160 CTAs at stage 1).
template <bool PER_PIXEL, int NW, int KMAX>
__global__ void __launch_bounds__(32 * NW)
softargmin_conf_wide_kernel(const float* __restrict__ cost, const float* __restrict__ planes, float* __restrict__ depth,
                            float* __restrict__ conf, int32_t* __restrict__ idx_out, int D, int64_t hw) {
  __shared__ float red[4][NW][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t p = min((int64_t)blockIdx.x * 32 + lane, hw - 1);          // clamped lanes compute, only in-range ones store
  const bool store = (int64_t)blockIdx.x * 32 + lane < hw;
  const int b = blockIdx.y;
  cost += (int64_t)b * D * hw + p;
  planes += PER_PIXEL ? (int64_t)b * D * hw + p : (int64_t)b * D;
  const int kd = D / NW;                                                     // planes per thread (D % NW == 0, kd <= KMAX)
  float c[KMAX];
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    c[k] = k < kd ? __ldg(cost + (int64_t)(w + k * NW) * hw) : -INFINITY;
    m = fmaxf(m, c[k]);
  }
  red[0][w][lane] = m;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < NW; ++j) m = fmaxf(m, red[0][j][lane]);
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    c[k] = k < kd ? expf(c[k] - m) : 0.f;
    s += c[k];
  }
  red[1][w][lane] = s;
  __syncthreads();
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < NW; ++j) sum += red[1][j][lane];
  float dsum = 0.f, isum = 0.f;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    if (k < kd) {
      const int d = w + k * NW;
      c[k] = __fdiv_rn(c[k], sum);                                           // p_d
      const float dv = PER_PIXEL ? __ldg(planes + (int64_t)d * hw) : __ldg(planes + d);
      dsum = __fadd_rn(dsum, __fmul_rn(c[k], dv));
      isum = __fadd_rn(isum, __fmul_rn(c[k], (float)d));
    }
  }
  red[2][w][lane] = dsum;
  red[3][w][lane] = isum;
  __syncthreads();
  float dtot = 0.f, itot = 0.f;
#pragma unroll
  for (int j = 0; j < NW; ++j) { dtot += red[2][j][lane]; itot += red[3][j][lane]; }
  int idx = (int)itot;                       // .long(): truncation toward zero (casmvsnet.py:119-120)
  idx = min(max(idx, 0), D - 1);             // casmvsnet.py:121
  // conf = p[idx-1] + p[idx] + p[idx+1] + p[idx+2], zero outside [0,D): every thread adds the planes it holds
  float c4 = 0.f;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int d = w + k * NW;
    if (k < kd && d >= idx - 1 && d <= idx + 2) c4 += c[k];
  }
  __syncthreads();                           // red[0] is reused
  red[0][w][lane] = c4;
  __syncthreads();
  if (w == 0 && store) {
    float ct = 0.f;
#pragma unroll
    for (int j = 0; j < NW; ++j) ct += red[0][j][lane];
    depth[(int64_t)b * hw + p] = dtot;
    conf[(int64_t)b * hw + p] = ct;
    if (idx_out) idx_out[(int64_t)b * hw + p] = idx;
  }
}

}  // namespace rcmvs

extern "C" int rcmvs_softargmin_conf_fwd(const float* cost, const float* depth_planes, int per_pixel, float* depth,
                                         float* conf, int32_t* idx, float* prob, int B, int D, int h, int w,
                                         rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(cost && depth_planes && depth && conf, RCMVS_ERR_BAD_ARG, "rcmvs_softargmin_conf_fwd: null pointer");
  RCMVS_REQUIRE(B > 0 && D > 0 && h > 0 && w > 0 && B <= 65535, RCMVS_ERR_BAD_ARG,
                "rcmvs_softargmin_conf_fwd: bad dims");
  const int64_t hw = (int64_t)h * w;
  // the wide kernel pays four block barriers per 32 pixels: measured 17 vs 53 us at D = 192 (128x160), but 51 vs 29 us at
  // D = 64 (256x320) and 89 vs 16 us at D = 8 (512x640) — only deep columns take it
  if (!prob && D % 8 == 0 && D >= 128 && D / 8 <= 24) {
    dim3 gridw((unsigned)ceil_div<int64_t>(hw, 32), B);
    if (per_pixel)
      softargmin_conf_wide_kernel<true, 8, 24><<<gridw, 256, 0, as_stream(stream)>>>(cost, depth_planes, depth, conf, idx, D, hw);
    else
      softargmin_conf_wide_kernel<false, 8, 24><<<gridw, 256, 0, as_stream(stream)>>>(cost, depth_planes, depth, conf, idx, D, hw);
    RCMVS_CHECK_LAUNCH("rcmvs_softargmin_conf_fwd");
    return RCMVS_OK;
  }
  dim3 grid((unsigned)ceil_div<int64_t>(hw, 128), B);
  if (per_pixel)
    softargmin_conf_kernel<true><<<grid, 128, 0, as_stream(stream)>>>(cost, depth_planes, depth, conf, idx, prob, D, hw);
  else
    softargmin_conf_kernel<false><<<grid, 128, 0, as_stream(stream)>>>(cost, depth_planes, depth, conf, idx, prob, D, hw);
  RCMVS_CHECK_LAUNCH("rcmvs_softargmin_conf_fwd");
  return RCMVS_OK;
}
