// Relative projections src·inverse(ref) (reference models/modules.py:315-317) on device, fp64, one thread each.
// Replaces torch.inverse (batched LU + host sync on `info`) + matmul + slicing: no host round trip, capturable.
#include "common.cuh"

namespace rcmvs {

__device__ void kext_mul(const float* P /*[2][4][4]*/, double out[16]) {
  // out = E with top 3x4 replaced by K[:3,:3] @ E[:3,:4]   (models/casmvsnet.py:78-81)
  const float* E = P;
  const float* K = P + 16;
  for (int i = 0; i < 16; ++i) out[i] = (double)E[i];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) {
      // the reference does this product in fp32; reproduce its rounding so both paths see the same matrix
      float acc = 0.f;
      for (int k = 0; k < 3; ++k) acc = fmaf(K[r * 4 + k], E[k * 4 + c], acc);
      out[r * 4 + c] = (double)acc;
    }
}

__device__ bool invert4(const double* A, double* inv) {
  double m[4][8];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      m[r][c] = A[r * 4 + c];
      m[r][4 + c] = (r == c) ? 1.0 : 0.0;
    }
  for (int col = 0; col < 4; ++col) {
    int piv = col;
    double best = fabs(m[col][col]);
    for (int r = col + 1; r < 4; ++r)
      if (fabs(m[r][col]) > best) { best = fabs(m[r][col]); piv = r; }
    if (best == 0.0) return false;
    if (piv != col)
      for (int c = 0; c < 8; ++c) { double t = m[col][c]; m[col][c] = m[piv][c]; m[piv][c] = t; }
    double d = 1.0 / m[col][col];
    for (int c = 0; c < 8; ++c) m[col][c] *= d;
    for (int r = 0; r < 4; ++r)
      if (r != col) {
        double f = m[r][col];
        for (int c = 0; c < 8; ++c) m[r][c] -= f * m[col][c];
      }
  }
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) inv[r * 4 + c] = m[r][4 + c];
  return true;
}

__device__ void rel_to_rt(const double* S, const double* R, float* rt) {
  double inv[16];
  if (!invert4(R, inv)) {
    for (int i = 0; i < 12; ++i) rt[i] = __int_as_float(0x7fc00000);  // NaN: singular reference projection
    return;
  }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) {
      double acc = 0.0;
      for (int k = 0; k < 4; ++k) acc += S[r * 4 + k] * inv[k * 4 + c];
      rt[r * 4 + c] = (float)acc;
    }
}

__global__ void relative_projections_kernel(const float* __restrict__ proj, float* __restrict__ rt, int B, int V) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * (V - 1)) return;
  int b = i / (V - 1), v = i % (V - 1) + 1;
  double R[16], S[16];
  kext_mul(proj + ((int64_t)b * V + 0) * 32, R);
  kext_mul(proj + ((int64_t)b * V + v) * 32, S);
  rel_to_rt(S, R, rt + (int64_t)i * 12);
}

__global__ void relative_pair_kernel(const float* __restrict__ src, const float* __restrict__ ref,
                                     float* __restrict__ rt, int B) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double R[16], S[16];
  for (int i = 0; i < 16; ++i) { R[i] = ref[b * 16 + i]; S[i] = src[b * 16 + i]; }
  rel_to_rt(S, R, rt + (int64_t)b * 12);
}

}  // namespace rcmvs

extern "C" int rcmvs_relative_projections(const float* proj, float* rt, int B, int V, rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(proj && rt, RCMVS_ERR_BAD_ARG, "rcmvs_relative_projections: null pointer");
  RCMVS_REQUIRE(B > 0 && V > 1, RCMVS_ERR_BAD_ARG, "rcmvs_relative_projections: need B>0, V>1 (B=%d V=%d)", B, V);
  int n = B * (V - 1);
  relative_projections_kernel<<<ceil_div(n, 64), 64, 0, as_stream(stream)>>>(proj, rt, B, V);
  RCMVS_CHECK_LAUNCH("rcmvs_relative_projections");
  return RCMVS_OK;
}

extern "C" int rcmvs_relative_projection_pair(const float* src_proj, const float* ref_proj, float* rt, int B,
                                              rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(src_proj && ref_proj && rt, RCMVS_ERR_BAD_ARG, "rcmvs_relative_projection_pair: null pointer");
  RCMVS_REQUIRE(B > 0, RCMVS_ERR_BAD_ARG, "rcmvs_relative_projection_pair: B must be positive");
  relative_pair_kernel<<<ceil_div(B, 64), 64, 0, as_stream(stream)>>>(src_proj, ref_proj, rt, B);
  RCMVS_CHECK_LAUNCH("rcmvs_relative_projection_pair");
  return RCMVS_OK;
}
