// a4: softmax over depth + soft-argmin depth + photometric confidence in ONE kernel
// (reference models/casmvsnet.py:111-122 + models/modules.py:519-525: softmax, mul+sum, pad, avg_pool3d, gather =
// 5-6 passes over the probability volume plus a padded copy).  Here each thread owns one pixel, reads the cost
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

}  // namespace rcmvs

extern "C" int rcmvs_softargmin_conf_fwd(const float* cost, const float* depth_planes, int per_pixel, float* depth,
                                         float* conf, int32_t* idx, float* prob, int B, int D, int h, int w,
                                         rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(cost && depth_planes && depth && conf, RCMVS_ERR_BAD_ARG, "rcmvs_softargmin_conf_fwd: null pointer");
  RCMVS_REQUIRE(B > 0 && D > 0 && h > 0 && w > 0 && B <= 65535, RCMVS_ERR_BAD_ARG,
                "rcmvs_softargmin_conf_fwd: bad dims");
  const int64_t hw = (int64_t)h * w;
  dim3 grid((unsigned)ceil_div<int64_t>(hw, 128), B);
  if (per_pixel)
    softargmin_conf_kernel<true><<<grid, 128, 0, as_stream(stream)>>>(cost, depth_planes, depth, conf, idx, prob, D, hw);
  else
    softargmin_conf_kernel<false><<<grid, 128, 0, as_stream(stream)>>>(cost, depth_planes, depth, conf, idx, prob, D, hw);
  RCMVS_CHECK_LAUNCH("rcmvs_softargmin_conf_fwd");
  return RCMVS_OK;
}
