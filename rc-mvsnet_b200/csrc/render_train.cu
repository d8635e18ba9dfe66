// Training path of the rendering-consistency branch (reference train_rcmvsnet.py:279-325 calls Rendering_Consistency_Net in
// .train() mode under autograd): the differentiable pieces around the convolution blocks (csrc/train.cu) —
//   sgemm                 C (+)= op(A)·op(B) (+ bias): the MLP's Linear layers forward, data gradient and (split-K, atomic)
//                         weight gradient (reference models/render_models.py:192-220 Renderer_ours.forward: 10 nn.Linear)
//   mul_relu / act        h = relu(Linear(h) ⊙ bias) (the multiplicative modulation, :200), relu, sigmoid + backward
//   embed                 the 63-d positional embedding of the NDC sample position (:45-49), no gradient needed
//   point_features_bwd    gradient of the 8-channel trilinear volume lookup w.r.t. the encoded volume (render_utils.py:304-330)
//   composite_bwd         gradient of alpha compositing w.r.t. raw = (rgb, sigma) (renderer.py:18-26, 65-93)
//   resize_depth_bwd      adjoint of Neural_Volume_Net's linear resize along depth (render_models.py:755-756)
// All fp32 FFMA; the forward inference kernels (csrc/render.cu) are unchanged.
#include "common.cuh"

namespace rcmvs {
namespace rtrain {

// ---------------------------------------------------------------------------------------------------------------------
// SGEMM: C[M][N] (+)= op(A)[M][K] · op(B)[K][N] (+ bias[N]);  TA: A stored [K][M];  TB: B stored [N][K].
// 128 x 64 tile, 256 threads, 8 x 4 micro-tile, K step 8; grid.z = split-K slices (then C must be zero-filled: atomics).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int GM = 128, GN = 64, GK = 8;

template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, const float* __restrict__ B, const float* __restrict__ bias, float* __restrict__ C,
             int M, int N, int K, int lda, int ldb, int ldc, int k_per_split) {
  __shared__ float As[GK][GM + 4];
  __shared__ float Bs[GK][GN + 4];
  const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
  const int k_begin = blockIdx.z * k_per_split, k_end = min(K, k_begin + k_per_split);
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;          // 16 x 16 threads: rows ty*8.., cols tx*4..
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = k_begin; k0 < k_end; k0 += GK) {
    // A tile: GM x GK = 1024 elements, 4 per thread
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = threadIdx.x + e * 256;
      int m, k;
      if (TA) { m = idx % GM; k = idx / GM; } else { k = idx % GK; m = idx / GK; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < k_end) v = TA ? __ldg(A + (int64_t)gk * lda + gm) : __ldg(A + (int64_t)gm * lda + gk);
      As[k][m] = v;
    }
    // B tile: GK x GN = 512 elements, 2 per thread
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int idx = threadIdx.x + e * 256;
      int n, k;
      if (TB) { k = idx % GK; n = idx / GK; } else { n = idx % GN; k = idx / GN; }
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < N && gk < k_end) v = TB ? __ldg(B + (int64_t)gn * ldb + gk) : __ldg(B + (int64_t)gk * ldb + gn);
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      float a[8], b[4];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]), a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const bool split = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = m0 + ty * 8 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (bias && blockIdx.z == 0) v += __ldg(bias + gn);
      if (split) atomicAdd(C + (int64_t)gm * ldc + gn, v);
      else C[(int64_t)gm * ldc + gn] = v;
    }
  }
}

// ---- elementwise --------------------------------------------------------------------------------------------------------
// mode 0: h = relu(a·b); mode 1: relu(a); mode 2: sigmoid(a)
__global__ void act_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ h, int64_t n,
                               int mode) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = a[i];
  h[i] = mode == 0 ? fmaxf(x * b[i], 0.f) : (mode == 1 ? fmaxf(x, 0.f) : 1.f / (1.f + expf(-x)));
}
// mode 0: da = g·b·[h>0], db = g·a·[h>0]; mode 1: da = g·[h>0]; mode 2: da = g·h·(1−h)   (h = the forward output)
__global__ void act_bwd_kernel(const float* __restrict__ g, const float* __restrict__ a, const float* __restrict__ b,
                               const float* __restrict__ h, float* __restrict__ da, float* __restrict__ db, int64_t n,
                               int mode) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i], hi = h[i];
  if (mode == 0) {
    const float m = hi > 0.f ? gi : 0.f;
    da[i] = m * b[i];
    db[i] = m * a[i];
  } else if (mode == 1) {
    da[i] = hi > 0.f ? gi : 0.f;
  } else {
    da[i] = gi * hi * (1.f - hi);
  }
}

// embed: x [P][3] -> out [P][ld] at column 0: [x (3), sin(2^k x) for k = 0..9 (k-major, then axis), cos likewise] = 63
__global__ void embed_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t P, int ld) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const float v[3] = {__ldg(x + i * 3), __ldg(x + i * 3 + 1), __ldg(x + i * 3 + 2)};
  float* o = out + i * ld;
  o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
  float f = 1.f;
#pragma unroll
  for (int k = 0; k < 10; ++k) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float s, c;
      sincosf(v[a] * f, &s, &c);
      o[3 + k * 3 + a] = s;
      o[33 + k * 3 + a] = c;
    }
    f *= 2.f;
  }
}

// ---- gradient of the 8-channel trilinear lookup w.r.t. the volume [D][h][w][8] (align_corners=True, zeros padding) -----------
__global__ void __launch_bounds__(128)
point_features_bwd_kernel(const float* __restrict__ ndc, const float* __restrict__ gfeat /*[P][ld], first 8 columns*/,
                          float* __restrict__ gvol, int64_t P, int ld, int D, int h, int w) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const float gx = __ldg(ndc + i * 3 + 0) * 2.f - 1.f, gy = __ldg(ndc + i * 3 + 1) * 2.f - 1.f, gz = __ldg(ndc + i * 3 + 2) * 2.f - 1.f;
  const float ix = (gx + 1.f) * 0.5f * (float)(w - 1), iy = (gy + 1.f) * 0.5f * (float)(h - 1), iz = (gz + 1.f) * 0.5f * (float)(D - 1);
  const bool fin = fabsf(ix) < 1e8f && fabsf(iy) < 1e8f && fabsf(iz) < 1e8f;
  const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
  const int x0 = fin ? (int)fx : -4, y0 = fin ? (int)fy : -4, z0 = fin ? (int)fz : -4;
  const float ax = ix - fx, ay = iy - fy, az = iz - fz;
  const float4 g0 = ldg4(gfeat + i * ld), g1 = ldg4(gfeat + i * ld + 4);
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int dx = c & 1, dy = (c >> 1) & 1, dz = c >> 2;
    const int xx = x0 + dx, yy = y0 + dy, zz = z0 + dz;
    if (xx >= 0 && xx < w && yy >= 0 && yy < h && zz >= 0 && zz < D) {
      const float wgt = (dx ? ax : 1.f - ax) * (dy ? ay : 1.f - ay) * (dz ? az : 1.f - az);
      float* p = gvol + (((int64_t)zz * h + yy) * w + xx) * 8;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(wgt * g0.x), "f"(wgt * g0.y),
                   "f"(wgt * g0.z), "f"(wgt * g0.w) : "memory");
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p + 4), "f"(wgt * g1.x), "f"(wgt * g1.y),
                   "f"(wgt * g1.z), "f"(wgt * g1.w) : "memory");
    }
  }
}

// ---- alpha compositing backward: one thread per ray (1024 rays x 128 samples: latency-trivial) ----------------------------
// forward (renderer.py:18-26, 65-93): a_i = 1 − exp(−s_i), T_i = Π_{j<i} (1 − a_j + 1e-10), w_i = a_i·T_i,
// rgb = Σ w_i c_i, depth = Σ w_i z_i.  grad inputs: g_rgb [N][3], g_depth [N], g_w [N][S] (or null) -> g_raw [N][S][4].
__global__ void composite_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ g_rgb,
                                     const float* __restrict__ g_depth, const float* __restrict__ g_w,
                                     float* __restrict__ g_raw, int N, int S) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= N) return;
  const float* rw = raw + (int64_t)r * S * 4;
  const float* zz = z + (int64_t)r * S;
  float* go = g_raw + (int64_t)r * S * 4;
  const float gr = g_rgb ? g_rgb[r * 3] : 0.f, gg = g_rgb ? g_rgb[r * 3 + 1] : 0.f, gb = g_rgb ? g_rgb[r * 3 + 2] : 0.f;
  const float gd = g_depth ? g_depth[r] : 0.f;
  // forward transmittance, kept in the gradient buffer's 4th slot temporarily
  float T = 1.f;
  for (int i = 0; i < S; ++i) {
    go[i * 4 + 3] = T;
    const float a = 1.f - expf(-rw[i * 4 + 3]);
    T *= (1.f - a + 1e-10f);
  }
  float suffix = 0.f;                                     // Σ_{k>i} gw_k·w_k
  for (int i = S - 1; i >= 0; --i) {
    const float e = expf(-rw[i * 4 + 3]);
    const float a = 1.f - e, Ti = go[i * 4 + 3], wi = a * Ti;
    float gw = gr * rw[i * 4] + gg * rw[i * 4 + 1] + gb * rw[i * 4 + 2] + gd * zz[i];
    if (g_w) gw += g_w[(int64_t)r * S + i];
    const float da = gw * Ti - suffix / (1.f - a + 1e-10f);
    go[i * 4] = gr * wi; go[i * 4 + 1] = gg * wi; go[i * 4 + 2] = gb * wi;
    go[i * 4 + 3] = da * e;
    suffix += gw * wi;
  }
}

// ---- adjoint of resize_depth_cl (linear along D, align_corners=True, channel padding dropped) ------------------------------
__global__ void resize_depth_bwd_kernel(const float* __restrict__ gy /*[B][D1][S][Cpad]*/, float* __restrict__ gx /*[B][D0][S][C]*/,
                                        int D0, int D1, int64_t S, int C, int Cpad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)D0 * S * C;
  if (i >= total) return;
  const int c = (int)(i % C);
  const int64_t s = (i / C) % S;
  const int d0 = (int)(i / ((int64_t)C * S));
  const int b = blockIdx.y;
  const float scale = D1 > 1 ? (float)(D0 - 1) / (float)(D1 - 1) : 0.f;
  const float* gb = gy + (int64_t)b * D1 * S * Cpad;
  // outputs d with i0(d) in {d0-1, d0}: d in [ (d0-1)/scale, (d0+1)/scale ] — scan that window with the forward's own arithmetic
  int lo = 0, hi = D1 - 1;
  if (scale > 0.f) {
    lo = max(0, (int)floorf((float)(d0 - 1) / scale) - 1);
    hi = min(D1 - 1, (int)ceilf((float)(d0 + 1) / scale) + 1);
  }
  float acc = 0.f;
  for (int d = lo; d <= hi; ++d) {
    const float src = scale * (float)d;
    const int i0 = min((int)src, D0 - 1), i1 = min(i0 + 1, D0 - 1);
    const float l1 = src - (float)i0, l0 = 1.f - l1;
    const float g = __ldg(gb + ((int64_t)d * S + s) * Cpad + c);
    if (i0 == d0) acc = fmaf(l0, g, acc);
    if (i1 == d0) acc = fmaf(l1, g, acc);
  }
  gx[(int64_t)b * total + i] = acc;
}

}  // namespace rtrain
}  // namespace rcmvs

using namespace rcmvs;
using namespace rcmvs::rtrain;

extern "C" int rcmvs_sgemm(const float* A, const float* B, const float* bias, float* C, int M, int N, int K, int lda, int ldb,
                           int ldc, int trans_a, int trans_b, int splits, rcmvs_stream_t stream) {
  RCMVS_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0 && splits >= 1, RCMVS_ERR_BAD_ARG, "rcmvs_sgemm: bad arguments");
  int kps = ceil_div(ceil_div(K, splits), GK) * GK;
  splits = ceil_div(K, kps);
  dim3 grid(ceil_div(N, GN), ceil_div(M, GM), splits);
  RCMVS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, RCMVS_ERR_UNSUPPORTED, "rcmvs_sgemm: grid too large");
  cudaStream_t st = as_stream(stream);
  if (!trans_a && !trans_b) sgemm_kernel<false, false><<<grid, 256, 0, st>>>(A, B, bias, C, M, N, K, lda, ldb, ldc, kps);
  else if (!trans_a && trans_b) sgemm_kernel<false, true><<<grid, 256, 0, st>>>(A, B, bias, C, M, N, K, lda, ldb, ldc, kps);
  else if (trans_a && !trans_b) sgemm_kernel<true, false><<<grid, 256, 0, st>>>(A, B, bias, C, M, N, K, lda, ldb, ldc, kps);
  else sgemm_kernel<true, true><<<grid, 256, 0, st>>>(A, B, bias, C, M, N, K, lda, ldb, ldc, kps);
  RCMVS_CHECK_LAUNCH("rcmvs_sgemm");
  return RCMVS_OK;
}

extern "C" int rcmvs_act_fwd(const float* a, const float* b, float* h, int64_t n, int mode, rcmvs_stream_t stream) {
  RCMVS_REQUIRE(a && h && n > 0 && mode >= 0 && mode <= 2 && (mode != 0 || b), RCMVS_ERR_BAD_ARG, "rcmvs_act_fwd: bad arguments");
  act_fwd_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, as_stream(stream)>>>(a, b, h, n, mode);
  RCMVS_CHECK_LAUNCH("rcmvs_act_fwd");
  return RCMVS_OK;
}

extern "C" int rcmvs_act_bwd(const float* g, const float* a, const float* b, const float* h, float* da, float* db, int64_t n,
                             int mode, rcmvs_stream_t stream) {
  RCMVS_REQUIRE(g && h && da && n > 0 && mode >= 0 && mode <= 2 && (mode != 0 || (a && b && db)), RCMVS_ERR_BAD_ARG,
                "rcmvs_act_bwd: bad arguments");
  act_bwd_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, as_stream(stream)>>>(g, a, b, h, da, db, n, mode);
  RCMVS_CHECK_LAUNCH("rcmvs_act_bwd");
  return RCMVS_OK;
}

extern "C" int rcmvs_embed(const float* x, float* out, int64_t P, int ld, rcmvs_stream_t stream) {
  RCMVS_REQUIRE(x && out && P > 0 && ld >= 63, RCMVS_ERR_BAD_ARG, "rcmvs_embed: bad arguments");
  embed_kernel<<<(unsigned)ceil_div<int64_t>(P, 128), 128, 0, as_stream(stream)>>>(x, out, P, ld);
  RCMVS_CHECK_LAUNCH("rcmvs_embed");
  return RCMVS_OK;
}

extern "C" int rcmvs_point_features_bwd(const float* ndc, const float* grad_feat, float* grad_vol, int64_t P, int ld, int D,
                                        int h, int w, rcmvs_stream_t stream) {
  RCMVS_REQUIRE(ndc && grad_feat && grad_vol && P > 0 && ld >= 8 && ld % 4 == 0, RCMVS_ERR_BAD_ARG,
                "rcmvs_point_features_bwd: bad arguments");
  point_features_bwd_kernel<<<(unsigned)ceil_div<int64_t>(P, 128), 128, 0, as_stream(stream)>>>(ndc, grad_feat, grad_vol, P, ld,
                                                                                            D, h, w);
  RCMVS_CHECK_LAUNCH("rcmvs_point_features_bwd");
  return RCMVS_OK;
}

extern "C" int rcmvs_composite_bwd(const float* raw, const float* z, const float* grad_rgb, const float* grad_depth,
                                   const float* grad_weights, float* grad_raw, int N, int S, rcmvs_stream_t stream) {
  RCMVS_REQUIRE(raw && z && grad_raw && N > 0 && S > 0, RCMVS_ERR_BAD_ARG, "rcmvs_composite_bwd: bad arguments");
  composite_bwd_kernel<<<ceil_div(N, 64), 64, 0, as_stream(stream)>>>(raw, z, grad_rgb, grad_depth, grad_weights, grad_raw, N, S);
  RCMVS_CHECK_LAUNCH("rcmvs_composite_bwd");
  return RCMVS_OK;
}

extern "C" int rcmvs_resize_depth_cl_bwd(const float* grad_y, float* grad_x, int B, int D0, int D1, int64_t S, int C, int Cpad,
                                         rcmvs_stream_t stream) {
  RCMVS_REQUIRE(grad_y && grad_x && B > 0 && D0 > 0 && D1 > 0 && S > 0 && C > 0 && Cpad >= C && B <= 65535, RCMVS_ERR_BAD_ARG,
                "rcmvs_resize_depth_cl_bwd: bad arguments");
  const int64_t total = (int64_t)D0 * S * C;
  dim3 grid((unsigned)ceil_div<int64_t>(total, 256), B);
  resize_depth_bwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(grad_y, grad_x, D0, D1, S, C, Cpad);
  RCMVS_CHECK_LAUNCH("rcmvs_resize_depth_cl_bwd");
  return RCMVS_OK;
}
