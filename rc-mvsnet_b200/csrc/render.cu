// a7-a12: the rendering-consistency branch (reference models/render_consist_net.py:54-76).
//   resize_depth_cl      Neural_Volume_Net's trilinear D 48->128 resize (render_models.py:755-756), channels-last,
//                        zero-padding the 41 input channels to a multiple of 8 for the conv kernels
//   gu_sample            Gaussian-Uniform depth candidates (render_utils.py:203-224): one warp per ray, Philox
//                        normals, in-register bitonic sort; replaces a 1024-iteration Python loop (~7k launches)
//   ray_points           ray directions, sample points, NDC coordinates (render_utils.py:86-146, :225)
//   point_features       8-ch trilinear volume lookup + 3 x (RGB bilinear border + in-bounds mask)
//                        (render_utils.py:247-279, 304-330; renderer.py:154-166)
//   render_mlp           positional embedding + Renderer_ours MLP fused per 128-point tile, activations in shared
//                        memory, fp32 FFMA (render_models.py:45-49, 192-220)
//   composite            alpha compositing per ray with a warp scan (renderer.py:18-26, 65-93)
#include <cuda_bf16.h>
#include <curand_kernel.h>

#include "common.cuh"

namespace rcmvs {

// ------------------------------------------------------------------------------------------------------------------
__global__ void resize_depth_cl_kernel(const float* __restrict__ x, float* __restrict__ y, int D0, int D1, int64_t S,
                                       int C, int Cpad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)D1 * S * Cpad;
  if (i >= total) return;
  const int c = (int)(i % Cpad);
  const int64_t s = (i / Cpad) % S;
  const int d = (int)(i / ((int64_t)Cpad * S));
  const int b = blockIdx.y;
  float v = 0.f;
  if (c < C) {
    // align_corners=True: src = d * (D0-1)/(D1-1)   (ATen area_pixel_compute_scale / source_index)
    const float scale = D1 > 1 ? (float)(D0 - 1) / (float)(D1 - 1) : 0.f;
    const float src = scale * (float)d;
    const int i0 = min((int)src, D0 - 1);
    const int i1 = min(i0 + 1, D0 - 1);
    const float l1 = src - (float)i0, l0 = 1.f - l1;
    const float* xb = x + (int64_t)b * D0 * S * C;
    v = l0 * __ldg(xb + ((int64_t)i0 * S + s) * C + c) + l1 * __ldg(xb + ((int64_t)i1 * S + s) * C + c);
  }
  y[(int64_t)b * total + i] = v;
}

// Same resize, written as split-bf16 planes [2][B][D1][S][Cpad] (hi, lo): the operand format of the TMA-fed tensor-core
// convolution (csrc/conv3d_slab_bf16.cu) that runs Neural_Volume_Net's conv0, the largest layer of the renderer.
// Thread = one (plane, position, 8-channel chunk): 16 scalar loads (C = 41 is not a multiple of 4), two 16-byte stores.
__global__ void __launch_bounds__(256)
resize_depth_cl_split_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int B, int D0, int D1, int64_t S,
                             int C, int Cpad) {
  const int nch = Cpad / 8;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total8 = (int64_t)D1 * S * nch;
  if (i >= total8) return;
  const int ch = (int)(i % nch);
  const int64_t s = (i / nch) % S;
  const int d = (int)(i / ((int64_t)nch * S));
  const int b = blockIdx.y;
  const float scale = D1 > 1 ? (float)(D0 - 1) / (float)(D1 - 1) : 0.f;
  const float src = scale * (float)d;
  const int i0 = min((int)src, D0 - 1);
  const int i1 = min(i0 + 1, D0 - 1);
  const float l1 = src - (float)i0, l0 = 1.f - l1;
  const float* x0 = x + (((int64_t)b * D0 + i0) * S + s) * C + ch * 8;
  const float* x1 = x + (((int64_t)b * D0 + i1) * S + s) * C + ch * 8;
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float v[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c = ch * 8 + 2 * k + e;
      v[e] = c < C ? l0 * __ldg(x0 + 2 * k + e) + l1 * __ldg(x1 + 2 * k + e) : 0.f;
    }
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v[0]), h1 = __float2bfloat16_rn(v[1]);
    const __nv_bfloat162 hh = __halves2bfloat162(h0, h1);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(v[0] - __bfloat162float(h0), v[1] - __bfloat162float(h1));
    hi[k] = *reinterpret_cast<const uint32_t*>(&hh);
    lo[k] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  const int64_t total = (int64_t)D1 * S * Cpad;
  const int64_t o = (int64_t)b * total + i * 8;
  *reinterpret_cast<uint4*>(y + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(y + (int64_t)B * total + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// ------------------------------------------------------------------------------------------------------------------
// Gaussian-Uniform sampler.  128 samples per ray = 4 per lane, element index i = lane*4 + k.
__device__ __forceinline__ void cmpswap(float& a, float& b, bool asc) {
  const float lo = fminf(a, b), hi = fmaxf(a, b);
  a = asc ? lo : hi;
  b = asc ? hi : lo;
}

__global__ void __launch_bounds__(128) gu_sample_kernel(const float* __restrict__ ray_depth, float near, float far,
                                                        unsigned long long seed, float* __restrict__ out, int N,
                                                        int n_gauss) {
  const int ray = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (ray >= N) return;
  curandStatePhilox4_32_10_t st;
  curand_init(seed, (unsigned long long)ray * 32 + lane, 0, &st);
  float v[4];
  if (ray < n_gauss) {
    const float d = __ldg(ray_depth + ray);
    const float sd = fminf(fabsf(far - d), fabsf(d - near)) / 3.f;       // render_utils.py:206
    const float4 g = curand_normal4(&st);
    v[0] = fmaf(sd, g.x, d); v[1] = fmaf(sd, g.y, d); v[2] = fmaf(sd, g.z, d); v[3] = fmaf(sd, g.w, d);
    // bitonic sort of 128 elements, ascending (torch.sort, render_utils.py:207)
    for (int size = 2; size <= 128; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        if (stride >= 4) {
          const int lstride = stride >> 2;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int i = lane * 4 + k;
            const float other = __shfl_xor_sync(0xffffffffu, v[k], lstride);
            const bool asc = (i & size) == 0;
            const bool lower = (lane & lstride) == 0;
            v[k] = (asc == lower) ? fminf(v[k], other) : fmaxf(v[k], other);
          }
        } else if (stride == 2) {
          const bool asc0 = ((lane * 4) & size) == 0;      // size >= 4 here, so all 4 in-lane elements agree
          cmpswap(v[0], v[2], asc0);
          cmpswap(v[1], v[3], asc0);
        } else {
          const bool asc01 = ((lane * 4) & size) == 0, asc23 = ((lane * 4 + 2) & size) == 0;
          cmpswap(v[0], v[1], asc01);
          cmpswap(v[2], v[3], asc23);
        }
      }
    }
  } else {
    // stratified uniform in [near, far]  (render_utils.py:213-224)
    const float4 u = curand_uniform4(&st);
    const float uu[4] = {1.f - u.x, 1.f - u.y, 1.f - u.z, 1.f - u.w};     // curand is (0,1], torch.rand [0,1)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = lane * 4 + k;
      auto zlin = [&](int j) {
        const float t = (float)j / 127.f;
        return near * (1.f - t) + far * t;
      };
      const float z = zlin(i);
      const float lo = i == 0 ? z : 0.5f * (zlin(i - 1) + z);
      const float hi = i == 127 ? z : 0.5f * (z + zlin(i + 1));
      v[k] = lo + (hi - lo) * uu[k];
    }
  }
  st4(out + (int64_t)ray * 128 + lane * 4, make_float4(v[0], v[1], v[2], v[3]));
}

// ------------------------------------------------------------------------------------------------------------------
struct RayCam {
  float K[9];      // intrinsics of the ray camera (= reference view)
  float c2w[12];   // rows of c2w[:3,:4]
  float w2c[12];   // rows of w2c_ref[:3,:4]
};

__global__ void ray_points_kernel(const float* __restrict__ ys, const float* __restrict__ xs,
                                  const float* __restrict__ K, const float* __restrict__ c2w,
                                  const float* __restrict__ w2c, const float* __restrict__ z, float near, float far,
                                  int W, int H, float* __restrict__ pts, float* __restrict__ ndc,
                                  float* __restrict__ dirs, int N, int S) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * S) return;
  const int r = i / S;
  const float x = __ldg(xs + r), y = __ldg(ys + r);
  const float dx = (x - __ldg(K + 2)) / __ldg(K + 0), dy = (y - __ldg(K + 5)) / __ldg(K + 4);
  // rays_d = dirs @ c2w[:3,:3]^T  (render_utils.py:103)
  float d[3], o[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    d[a] = dx * __ldg(c2w + a * 4 + 0) + dy * __ldg(c2w + a * 4 + 1) + __ldg(c2w + a * 4 + 2);
    o[a] = __ldg(c2w + a * 4 + 3);
  }
  if (i % S == 0) { dirs[r * 3 + 0] = d[0]; dirs[r * 3 + 1] = d[1]; dirs[r * 3 + 2] = d[2]; }
  const float zz = __ldg(z + i);
  float p[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) p[a] = o[a] + zz * d[a];                 // render_utils.py:225
  pts[(int64_t)i * 3 + 0] = p[0]; pts[(int64_t)i * 3 + 1] = p[1]; pts[(int64_t)i * 3 + 2] = p[2];
  // get_ndc_coordinate (render_utils.py:121-131)
  float c[3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
    c[a] = p[0] * __ldg(w2c + a * 4 + 0) + p[1] * __ldg(w2c + a * 4 + 1) + p[2] * __ldg(w2c + a * 4 + 2) +
           __ldg(w2c + a * 4 + 3);
  float q[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) q[a] = c[0] * __ldg(K + a * 3 + 0) + c[1] * __ldg(K + a * 3 + 1) + c[2] * __ldg(K + a * 3 + 2);
  ndc[(int64_t)i * 3 + 0] = q[0] / q[2] / (float)(W - 1);
  ndc[(int64_t)i * 3 + 1] = q[1] / q[2] / (float)(H - 1);
  ndc[(int64_t)i * 3 + 2] = (q[2] - near) / (far - near);
}

// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) point_features_kernel(
    const float* __restrict__ vol /*[D][h][w][8]*/, const float* __restrict__ ndc, const float* __restrict__ pts,
    const float* __restrict__ imgs /*[3][3][H][W]*/, const float* __restrict__ w2cs /*[3][4][4]*/,
    const float* __restrict__ Ks /*[3][3][3]*/, float* __restrict__ out /*[P][20]*/, int64_t P, int D, int h, int w,
    int H, int W) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  float f[20];
  // ---- 8-channel trilinear, zeros padding, align_corners=True (grid = ndc*2-1 -> index = ndc*(size-1)) ----
  {
    const float gx = __ldg(ndc + i * 3 + 0) * 2.f - 1.f, gy = __ldg(ndc + i * 3 + 1) * 2.f - 1.f,
                gz = __ldg(ndc + i * 3 + 2) * 2.f - 1.f;
    const float ix = (gx + 1.f) * 0.5f * (float)(w - 1), iy = (gy + 1.f) * 0.5f * (float)(h - 1),
                iz = (gz + 1.f) * 0.5f * (float)(D - 1);
    const bool fin = fabsf(ix) < 1e8f && fabsf(iy) < 1e8f && fabsf(iz) < 1e8f;
    const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
    const int x0 = fin ? (int)fx : -4, y0 = fin ? (int)fy : -4, z0 = fin ? (int)fz : -4;
    const float ax = ix - fx, ay = iy - fy, az = iz - fz;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int dx = c & 1, dy = (c >> 1) & 1, dz = c >> 2;
      const int xx = x0 + dx, yy = y0 + dy, zz = z0 + dz;
      if (xx >= 0 && xx < w && yy >= 0 && yy < h && zz >= 0 && zz < D) {
        const float wgt = (dx ? ax : 1.f - ax) * (dy ? ay : 1.f - ay) * (dz ? az : 1.f - az);
        const float* p = vol + (((int64_t)zz * h + yy) * w + xx) * 8;
        const float4 v0 = ldg4(p), v1 = ldg4(p + 4);
        a0.x = fmaf(wgt, v0.x, a0.x); a0.y = fmaf(wgt, v0.y, a0.y); a0.z = fmaf(wgt, v0.z, a0.z); a0.w = fmaf(wgt, v0.w, a0.w);
        a1.x = fmaf(wgt, v1.x, a1.x); a1.y = fmaf(wgt, v1.y, a1.y); a1.z = fmaf(wgt, v1.z, a1.z); a1.w = fmaf(wgt, v1.w, a1.w);
      }
    }
    f[0] = a0.x; f[1] = a0.y; f[2] = a0.z; f[3] = a0.w; f[4] = a1.x; f[5] = a1.y; f[6] = a1.z; f[7] = a1.w;
  }
  // ---- 3 views x (RGB bilinear with border padding + in-bounds mask) ----
  const float px = __ldg(pts + i * 3 + 0), py = __ldg(pts + i * 3 + 1), pz = __ldg(pts + i * 3 + 2);
#pragma unroll
  for (int v = 0; v < 3; ++v) {
    const float* M = w2cs + v * 16;
    const float* K = Ks + v * 9;
    float c[3], q[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
      c[a] = px * __ldg(M + a * 4 + 0) + py * __ldg(M + a * 4 + 1) + pz * __ldg(M + a * 4 + 2) + __ldg(M + a * 4 + 3);
#pragma unroll
    for (int a = 0; a < 3; ++a) q[a] = c[0] * __ldg(K + a * 3 + 0) + c[1] * __ldg(K + a * 3 + 1) + c[2] * __ldg(K + a * 3 + 2);
    const float gx = q[0] / q[2] / (float)(W - 1) * 2.f - 1.f, gy = q[1] / q[2] / (float)(H - 1) * 2.f - 1.f;
    const float mask = (gx > -1.f && gx < 1.f && gy > -1.f && gy < 1.f) ? 1.f : 0.f;
    // border padding: clip the un-normalised coordinate to [0, size-1] (ATen clip_coordinates); NaN -> 0
    float ix = (gx + 1.f) * 0.5f * (float)(W - 1), iy = (gy + 1.f) * 0.5f * (float)(H - 1);
    ix = fminf(fmaxf(ix, 0.f), (float)(W - 1));
    iy = fminf(fmaxf(iy, 0.f), (float)(H - 1));
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy;
    const float ax = ix - fx, ay = iy - fy;
    const int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
    const float w00 = (1.f - ax) * (1.f - ay), w01 = (x0 + 1 <= W - 1) ? ax * (1.f - ay) : 0.f,
                w10 = (y0 + 1 <= H - 1) ? (1.f - ax) * ay : 0.f,
                w11 = (x0 + 1 <= W - 1 && y0 + 1 <= H - 1) ? ax * ay : 0.f;
    const float* img = imgs + (int64_t)v * 3 * H * W;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const float* pl = img + (int64_t)ch * H * W;
      f[8 + v * 4 + ch] = w00 * __ldg(pl + (int64_t)y0 * W + x0) + w01 * __ldg(pl + (int64_t)y0 * W + x1) +
                          w10 * __ldg(pl + (int64_t)y1 * W + x0) + w11 * __ldg(pl + (int64_t)y1 * W + x1);
    }
    f[8 + v * 4 + 3] = mask;
  }
#pragma unroll
  for (int c = 0; c < 20; c += 4) st4(out + i * 20 + c, make_float4(f[c], f[c + 1], f[c + 2], f[c + 3]));
}

// ------------------------------------------------------------------------------------------------------------------
// Fused embedding + MLP.  Tile = 128 points, 256 threads; each thread owns 8 points x 8 outputs of a 128x128 layer.
// Weights are pre-transposed [K][N] in global memory (read through L1, 32-byte runs per thread, shared by the tile);
// activations live in shared memory and are updated in place between __syncthreads().
struct MlpW {
  const float* w[6];     // pts_linears.i  transposed [K_i][128]  (K = 63,128,128,128,128,191)
  const float* b[6];
  const float* wbias;    // pts_bias       [20][128]
  const float* bbias;
  const float* walpha;   // alpha_linear   [128]
  const float* balpha;
  const float* wfeat;    // feature_linear [128][128]
  const float* bfeat;
  const float* wview;    // views_linears.0 [131][64]
  const float* bview;
  const float* wrgb;     // rgb_linear     [64][3]
  const float* brgb;
};

constexpr int MLP_TP = 128;      // points per tile
constexpr int MLP_LD = 132;      // activation row pitch (floats)
constexpr int MLP_PE = 64;       // embedding row pitch

template <int N_T>   // outputs per thread along n: 8 (N=128) or 4 (N=64)
__device__ __forceinline__ void mlp_gemm(float (&acc)[8][N_T], const float* __restrict__ A, int lda, int K,
                                         const float* __restrict__ Wt, int N, int p0, int n0) {
  // unrolled: the weight rows of four k steps are requested before the first FMA block needs them (ncu: the loop was
  // latency-bound on these L1 / L2 loads at 8 warps per SM, profiles/r02_render_branch_ncu.md)
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = A[(p0 + j) * lda + k];
    float wv[N_T];
    const float4 w0 = ldg4(Wt + (int64_t)k * N + n0);
    wv[0] = w0.x; wv[1] = w0.y; wv[2] = w0.z; wv[3] = w0.w;
    if constexpr (N_T == 8) {
      const float4 w1 = ldg4(Wt + (int64_t)k * N + n0 + 4);
      wv[4] = w1.x; wv[5] = w1.y; wv[6] = w1.z; wv[7] = w1.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int n = 0; n < N_T; ++n) acc[j][n] = fmaf(a[j], wv[n], acc[j][n]);
  }
}

__global__ void __launch_bounds__(256) render_mlp_kernel(const float* __restrict__ ndc, const float* __restrict__ feat,
                                                         const float* __restrict__ dirs, const MlpW W,
                                                         float* __restrict__ out, int64_t P, int S) {
  extern __shared__ __align__(16) float sm[];
  float* hA = sm;                                   // [128][132]
  float* pe = hA + MLP_TP * MLP_LD;                 // [128][64]   embedded ndc (63 used)
  float* bias = pe + MLP_TP * MLP_PE;               // [128][128]  pts_bias(feat)
  float* dirs_s = bias + MLP_TP * 128;              // [128][4]
  float* alpha_s = dirs_s + MLP_TP * 4;             // [128]
  const int tid = threadIdx.x;
  const int64_t base = (int64_t)blockIdx.x * MLP_TP;

  // ---- embedding: [x, sin(x*2^0..2^9) (30), cos(...) (30)]  (render_models.py:45-49) ----
  for (int e = tid; e < MLP_TP * 64; e += 256) {
    const int p = e >> 6, c = e & 63;
    const int64_t gp = min(base + p, P - 1);
    float v = 0.f;
    if (c < 3) v = __ldg(ndc + gp * 3 + c);
    else if (c < 63) {
      const int j = (c - 3) % 30, f = j / 3, a = j % 3;
      const float x = __ldg(ndc + gp * 3 + a) * (float)(1 << f);
      v = (c < 33) ? sinf(x) : cosf(x);
    }
    pe[p * MLP_PE + c] = v;
  }
  for (int e = tid; e < MLP_TP * 20; e += 256) {    // stage feat in hA[:, 0:20] for the bias GEMM
    const int p = e / 20, c = e % 20;
    hA[p * MLP_LD + c] = __ldg(feat + min(base + p, P - 1) * 20 + c);
  }
  for (int e = tid; e < MLP_TP * 3; e += 256) {
    const int p = e / 3, c = e % 3;
    dirs_s[p * 4 + c] = __ldg(dirs + (min(base + p, P - 1) / S) * 3 + c);
  }
  __syncthreads();

  const int p0 = (tid >> 4) * 8, n0 = (tid & 15) * 8;
  float acc[8][8];
  auto zero = [&]() {
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int n = 0; n < 8; ++n) acc[j][n] = 0.f;
  };
  // bias = pts_bias(feat)
  zero();
  mlp_gemm<8>(acc, hA, MLP_LD, 20, W.wbias, 128, p0, n0);
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int n = 0; n < 8; ++n) bias[(p0 + j) * 128 + n0 + n] = acc[j][n] + __ldg(W.bbias + n0 + n);
  __syncthreads();

  // 6 modulated layers: h = relu((W h + b) * bias); layer 5 consumes [pe, h]  (render_models.py:196-203)
  for (int l = 0; l < 6; ++l) {
    zero();
    if (l == 0) {
      mlp_gemm<8>(acc, pe, MLP_PE, 63, W.w[0], 128, p0, n0);
    } else if (l == 5) {
      mlp_gemm<8>(acc, pe, MLP_PE, 63, W.w[5], 128, p0, n0);
      mlp_gemm<8>(acc, hA, MLP_LD, 128, W.w[5] + 63 * 128, 128, p0, n0);
    } else {
      mlp_gemm<8>(acc, hA, MLP_LD, 128, W.w[l], 128, p0, n0);
    }
    __syncthreads();                               // everyone finished reading hA
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int n = 0; n < 8; ++n)
        hA[(p0 + j) * MLP_LD + n0 + n] =
            fmaxf((acc[j][n] + __ldg(W.b[l] + n0 + n)) * bias[(p0 + j) * 128 + n0 + n], 0.f);
    __syncthreads();
  }
  // alpha = relu(alpha_linear(h))
  if (tid < MLP_TP) {
    float s = __ldg(W.balpha);
    for (int k = 0; k < 128; ++k) s = fmaf(hA[tid * MLP_LD + k], __ldg(W.walpha + k), s);
    alpha_s[tid] = fmaxf(s, 0.f);
  }
  // feature = feature_linear(h) (no activation), in place
  zero();
  mlp_gemm<8>(acc, hA, MLP_LD, 128, W.wfeat, 128, p0, n0);
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int n = 0; n < 8; ++n) hA[(p0 + j) * MLP_LD + n0 + n] = acc[j][n] + __ldg(W.bfeat + n0 + n);
  __syncthreads();
  // h2 = relu(views_linears.0([feature, dir]))  (131 -> 64): 8 points x 4 outputs per thread
  {
    float a2[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int n = 0; n < 4; ++n) a2[j][n] = 0.f;
    const int m0 = (tid & 15) * 4;
    mlp_gemm<4>(a2, hA, MLP_LD, 128, W.wview, 64, p0, m0);
    mlp_gemm<4>(a2, dirs_s, 4, 3, W.wview + 128 * 64, 64, p0, m0);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int n = 0; n < 4; ++n) hA[(p0 + j) * MLP_LD + m0 + n] = fmaxf(a2[j][n] + __ldg(W.bview + m0 + n), 0.f);
    __syncthreads();
  }
  // rgb = sigmoid(rgb_linear(h2)); out = [rgb, alpha]
  if (tid < MLP_TP && base + tid < P) {
    float r[3] = {__ldg(W.brgb), __ldg(W.brgb + 1), __ldg(W.brgb + 2)};
    for (int k = 0; k < 64; ++k) {
      const float hv = hA[tid * MLP_LD + k];
      r[0] = fmaf(hv, __ldg(W.wrgb + k * 3 + 0), r[0]);
      r[1] = fmaf(hv, __ldg(W.wrgb + k * 3 + 1), r[1]);
      r[2] = fmaf(hv, __ldg(W.wrgb + k * 3 + 2), r[2]);
    }
    st4(out + (base + tid) * 4, make_float4(1.f / (1.f + expf(-r[0])), 1.f / (1.f + expf(-r[1])),
                                            1.f / (1.f + expf(-r[2])), alpha_s[tid]));
  }
}

// ------------------------------------------------------------------------------------------------------------------
// alpha = 1-exp(-sigma); T = exclusive cumprod(1-alpha+1e-10); w = alpha*T; rgb = Σ w c; depth = Σ w z
__global__ void __launch_bounds__(128) composite_kernel(const float* __restrict__ raw, const float* __restrict__ z,
                                                        float* __restrict__ rgb, float* __restrict__ depth,
                                                        float* __restrict__ weights, float* __restrict__ alpha_out,
                                                        int N, int S) {
  const int ray = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (ray >= N) return;
  const int per = S / 32;                      // consecutive samples per lane (4 for S=128)
  float run = 1.f;
  float a_[8], t_[8];
  for (int k = 0; k < per; ++k) {
    const int s = lane * per + k;
    const float sigma = __ldg(raw + ((int64_t)ray * S + s) * 4 + 3);
    a_[k] = 1.f - expf(-sigma);
    t_[k] = run;                               // product of the lane's earlier factors
    run *= (1.f - a_[k] + 1e-10f);
  }
  // exclusive scan of the per-lane products across the warp
  float incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl *= up;
  }
  float excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) excl = 1.f;
  float r = 0.f, g = 0.f, b = 0.f, d = 0.f;
  for (int k = 0; k < per; ++k) {
    const int s = lane * per + k;
    const int64_t o = (int64_t)ray * S + s;
    const float wgt = a_[k] * (excl * t_[k]);
    r = fmaf(wgt, __ldg(raw + o * 4 + 0), r);
    g = fmaf(wgt, __ldg(raw + o * 4 + 1), g);
    b = fmaf(wgt, __ldg(raw + o * 4 + 2), b);
    d = fmaf(wgt, __ldg(z + o), d);
    weights[o] = wgt;
    alpha_out[o] = a_[k];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    r += __shfl_xor_sync(0xffffffffu, r, o);
    g += __shfl_xor_sync(0xffffffffu, g, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    d += __shfl_xor_sync(0xffffffffu, d, o);
  }
  if (lane == 0) {
    rgb[ray * 3 + 0] = r; rgb[ray * 3 + 1] = g; rgb[ray * 3 + 2] = b;
    depth[ray] = d;
  }
}

}  // namespace rcmvs

extern "C" int rcmvs_resize_depth_cl_split(const float* x, void* y_split, int B, int D0, int D1, int64_t S, int C, int Cpad,
                                           rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(x && y_split && B > 0 && D0 > 0 && D1 > 0 && S > 0 && C > 0 && Cpad >= C && B <= 65535, RCMVS_ERR_BAD_ARG,
                "rcmvs_resize_depth_cl_split: bad arguments");
  RCMVS_REQUIRE(Cpad % 8 == 0, RCMVS_ERR_BAD_ARG, "rcmvs_resize_depth_cl_split: Cpad must be a multiple of 8");
  const int64_t total8 = (int64_t)D1 * S * (Cpad / 8);
  dim3 grid((unsigned)ceil_div<int64_t>(total8, 256), B);
  resize_depth_cl_split_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, reinterpret_cast<__nv_bfloat16*>(y_split), B, D0, D1, S,
                                                                    C, Cpad);
  RCMVS_CHECK_LAUNCH("rcmvs_resize_depth_cl_split");
  return RCMVS_OK;
}

extern "C" int rcmvs_resize_depth_cl(const float* x, float* y, int B, int D0, int D1, int64_t S, int C, int Cpad,
                                     rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(x && y, RCMVS_ERR_BAD_ARG, "rcmvs_resize_depth_cl: null pointer");
  RCMVS_REQUIRE(B > 0 && D0 > 0 && D1 > 0 && S > 0 && C > 0 && Cpad >= C && B <= 65535, RCMVS_ERR_BAD_ARG,
                "rcmvs_resize_depth_cl: bad dims");
  const int64_t total = (int64_t)D1 * S * Cpad;
  dim3 grid((unsigned)ceil_div<int64_t>(total, 256), B);
  resize_depth_cl_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, y, D0, D1, S, C, Cpad);
  RCMVS_CHECK_LAUNCH("rcmvs_resize_depth_cl");
  return RCMVS_OK;
}

extern "C" int rcmvs_gu_sample(const float* ray_depth, float near, float far, uint64_t seed, float* out, int N,
                               int n_gauss, int S, rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(ray_depth && out, RCMVS_ERR_BAD_ARG, "rcmvs_gu_sample: null pointer");
  RCMVS_REQUIRE(N > 0 && n_gauss >= 0 && n_gauss <= N, RCMVS_ERR_BAD_ARG, "rcmvs_gu_sample: bad ray counts");
  RCMVS_REQUIRE(S == 128, RCMVS_ERR_UNSUPPORTED, "rcmvs_gu_sample: N_samples must be 128 (got %d)", S);
  gu_sample_kernel<<<ceil_div(N, 4), 128, 0, as_stream(stream)>>>(ray_depth, near, far, (unsigned long long)seed, out,
                                                                  N, n_gauss);
  RCMVS_CHECK_LAUNCH("rcmvs_gu_sample");
  return RCMVS_OK;
}

extern "C" int rcmvs_ray_points(const float* ys, const float* xs, const float* K, const float* c2w,
                                const float* w2c_ref, const float* z, float near, float far, int W, int H,
                                float* pts, float* ndc, float* dirs, int N, int S, rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(ys && xs && K && c2w && w2c_ref && z && pts && ndc && dirs, RCMVS_ERR_BAD_ARG,
                "rcmvs_ray_points: null pointer");
  RCMVS_REQUIRE(N > 0 && S > 0 && W > 1 && H > 1, RCMVS_ERR_BAD_ARG, "rcmvs_ray_points: bad dims");
  ray_points_kernel<<<ceil_div(N * S, 256), 256, 0, as_stream(stream)>>>(ys, xs, K, c2w, w2c_ref, z, near, far, W, H,
                                                                         pts, ndc, dirs, N, S);
  RCMVS_CHECK_LAUNCH("rcmvs_ray_points");
  return RCMVS_OK;
}

extern "C" int rcmvs_point_features(const float* vol_cl, const float* ndc, const float* pts, const float* imgs,
                                    const float* w2cs, const float* Ks, float* out, int64_t P, int D, int h, int w,
                                    int H, int W, rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(vol_cl && ndc && pts && imgs && w2cs && Ks && out, RCMVS_ERR_BAD_ARG,
                "rcmvs_point_features: null pointer");
  RCMVS_REQUIRE(P > 0 && D > 0 && h > 0 && w > 0 && H > 1 && W > 1, RCMVS_ERR_BAD_ARG, "rcmvs_point_features: bad dims");
  point_features_kernel<<<(unsigned)ceil_div<int64_t>(P, 128), 128, 0, as_stream(stream)>>>(vol_cl, ndc, pts, imgs,
                                                                                            w2cs, Ks, out, P, D, h, w,
                                                                                            H, W);
  RCMVS_CHECK_LAUNCH("rcmvs_point_features");
  return RCMVS_OK;
}

extern "C" int rcmvs_render_mlp(const float* ndc, const float* feat, const float* dirs, const float* const* weights,
                                float* out, int64_t P, int S, rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(ndc && feat && dirs && weights && out, RCMVS_ERR_BAD_ARG, "rcmvs_render_mlp: null pointer");
  RCMVS_REQUIRE(P > 0 && S > 0, RCMVS_ERR_BAD_ARG, "rcmvs_render_mlp: bad dims");
  MlpW W;
  for (int i = 0; i < 6; ++i) { W.w[i] = weights[2 * i]; W.b[i] = weights[2 * i + 1]; }
  W.wbias = weights[12]; W.bbias = weights[13]; W.walpha = weights[14]; W.balpha = weights[15];
  W.wfeat = weights[16]; W.bfeat = weights[17]; W.wview = weights[18]; W.bview = weights[19];
  W.wrgb = weights[20]; W.brgb = weights[21];
  for (int i = 0; i < 22; ++i) RCMVS_REQUIRE(weights[i], RCMVS_ERR_BAD_ARG, "rcmvs_render_mlp: null weight %d", i);
  const size_t smem = (size_t)(MLP_TP * MLP_LD + MLP_TP * MLP_PE + MLP_TP * 128 + MLP_TP * 4 + MLP_TP) * 4;
  static SmemAttrOnce smem_attr;
  {
    cudaError_t e = smem_attr.ensure(render_mlp_kernel, (int)smem);
    RCMVS_REQUIRE(e == cudaSuccess, RCMVS_ERR_CUDA, "rcmvs_render_mlp: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  render_mlp_kernel<<<(unsigned)ceil_div<int64_t>(P, MLP_TP), 256, smem, as_stream(stream)>>>(ndc, feat, dirs, W, out,
                                                                                              P, S);
  RCMVS_CHECK_LAUNCH("rcmvs_render_mlp");
  return RCMVS_OK;
}

extern "C" int rcmvs_composite(const float* raw, const float* z, float* rgb, float* depth, float* weights,
                               float* alpha, int N, int S, rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(raw && z && rgb && depth && weights && alpha, RCMVS_ERR_BAD_ARG, "rcmvs_composite: null pointer");
  RCMVS_REQUIRE(N > 0 && S > 0 && S % 32 == 0 && S <= 256, RCMVS_ERR_UNSUPPORTED,
                "rcmvs_composite: N_samples must be a multiple of 32, at most 256");
  composite_kernel<<<ceil_div(N, 4), 128, 0, as_stream(stream)>>>(raw, z, rgb, depth, weights, alpha, N, S);
  RCMVS_CHECK_LAUNCH("rcmvs_composite");
  return RCMVS_OK;
}
