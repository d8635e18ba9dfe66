// a1 homo_warping (reference models/modules.py:304-339) and the fused plane sweep + variance reduction that replaces
// DepthNet's inner loop (reference models/casmvsnet.py:65-100 / 254-288).
//
// Fused kernel data flow (per reference view, per stage):
//   read   V·C·h·w   fp32 source/ref features, channels-last [V][h][w][C]  (<= 52 MB at DTU sizes: L2-resident)
//   read   D·h·w     depth hypotheses (or D scalars at stage 1)
//   write  C·D·h·w   variance volume, channels-last [D][h][w][C], exactly once, streaming stores
// The reference materialises (V-1) warped volumes and streams the volume 3(V-1)+4 times; here Σf and Σf² live in
// registers and nothing but the final variance touches HBM.
//
// Thread mapping: one thread = one reference pixel × one group of 4 channels (float4).  Consecutive lanes hold the
// channel groups of one pixel, then the next pixel, so a warp's tap load is (32/CQ) contiguous C·4-byte runs and a
// warp's store is one contiguous (32/CQ)·C·4 = 512-byte run.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"

namespace rcmvs {

constexpr int kMaxSrc = 15;  // V-1 upper bound (rt staged in shared memory)

struct Tap {
  int off00, off01, off10, off11;  // element offsets of the 4 taps (pixel index, not yet × C), clamped in-bounds
  float w00, w01, w10, w11;        // bilinear weights, zeroed for out-of-bounds taps
  bool any;
};

// Pixel-space sample position -> taps.  grid_sample(bilinear, zeros, align_corners=True): tap (xi,yi) contributes
// iff 0<=xi<w and 0<=yi<h; NaN/inf/huge coordinates contribute nothing.
__device__ __forceinline__ Tap make_tap(float px, float py, int h, int w) {
  Tap t;
  const bool finite = (fabsf(px) < 1e8f) && (fabsf(py) < 1e8f);  // false for NaN
  const float fx = floorf(px), fy = floorf(py);
  const float ax = px - fx, ay = py - fy;  // exact
  const int x0 = finite ? (int)fx : -4, y0 = finite ? (int)fy : -4;
  const bool vx0 = (x0 >= 0) && (x0 < w), vx1 = (x0 + 1 >= 0) && (x0 + 1 < w);
  const bool vy0 = (y0 >= 0) && (y0 < h), vy1 = (y0 + 1 >= 0) && (y0 + 1 < h);
  const float bx = 1.f - ax, by = 1.f - ay;
  t.w00 = (vx0 && vy0) ? bx * by : 0.f;
  t.w01 = (vx1 && vy0) ? ax * by : 0.f;
  t.w10 = (vx0 && vy1) ? bx * ay : 0.f;
  t.w11 = (vx1 && vy1) ? ax * ay : 0.f;
  t.any = (vx0 || vx1) && (vy0 || vy1);
  const int cx0 = min(max(x0, 0), w - 1), cx1 = min(max(x0 + 1, 0), w - 1);
  const int cy0 = min(max(y0, 0), h - 1), cy1 = min(max(y0 + 1, 0), h - 1);
  t.off00 = cy0 * w + cx0;
  t.off01 = cy0 * w + cx1;
  t.off10 = cy1 * w + cx0;
  t.off11 = cy1 * w + cx1;
  return t;
}

__device__ __forceinline__ void project(const float* __restrict__ rt, float rx, float ry, float rz, float d,
                                        float& px, float& py) {
  // p = rot·(x,y,1)·depth + trans ; (px,py) = p.xy / p.z   (modules.py:325-329)
  const float X = fmaf(rx, d, rt[3]);
  const float Y = fmaf(ry, d, rt[7]);
  const float Z = fmaf(rz, d, rt[11]);
  px = __fdiv_rn(X, Z);
  py = __fdiv_rn(Y, Z);
}

// ------------------------------------------------------------------------------------------------------------------
// Fused plane sweep + variance
// ------------------------------------------------------------------------------------------------------------------
template <int C, bool PER_PIXEL, bool NOREF>
__global__ void __launch_bounds__(256)
costvolume_var_kernel(const float* __restrict__ feats, const float* __restrict__ rt_g,
                      const float* __restrict__ depth,
                      float* __restrict__ var, float* __restrict__ var_noref, int V, int D, int h, int w,
                      int d_chunk, int chunks_per_batch) {
  constexpr int CQ = C / 4;        // float4 groups per pixel
  constexpr int PXB = 256 / CQ;    // pixels per block
  const int q = threadIdx.x % CQ;
  const int x = blockIdx.x * PXB + threadIdx.x / CQ;
  const int y = blockIdx.y;
  const int b = blockIdx.z / chunks_per_batch;
  const int d_begin = (blockIdx.z % chunks_per_batch) * d_chunk;
  const int d_end = min(D, d_begin + d_chunk);
  __shared__ float rt_s[kMaxSrc * 12];
  for (int i = threadIdx.x; i < (V - 1) * 12; i += 256) rt_s[i] = __ldg(rt_g + (int64_t)b * (V - 1) * 12 + i);
  __syncthreads();
  if (x >= w) return;

  const int64_t hw = (int64_t)h * w;
  feats += (int64_t)b * V * hw * C + q * 4;
  var += (int64_t)b * D * hw * C + q * 4;
  if (NOREF) var_noref += (int64_t)b * D * hw * C + q * 4;
  depth += PER_PIXEL ? (int64_t)b * D * hw + (int64_t)y * w + x : (int64_t)b * D;

  const float4 ref = ldg4(feats + ((int64_t)y * w + x) * C);
  const float inv_v = 1.f / (float)V;
  const float xf = (float)x, yf = (float)y;

  for (int d = d_begin; d < d_end; ++d) {
    const float dep = PER_PIXEL ? __ldg(depth + (int64_t)d * hw) : __ldg(depth + d);
    float4 s = ref;
    float4 s2 = make_float4(ref.x * ref.x, ref.y * ref.y, ref.z * ref.z, ref.w * ref.w);
    float4 n = make_float4(0.f, 0.f, 0.f, 0.f), n2 = n;
    for (int v = 0; v < V - 1; ++v) {
      const float* m = rt_s + v * 12;
      const float rx = fmaf(m[0], xf, fmaf(m[1], yf, m[2]));
      const float ry = fmaf(m[4], xf, fmaf(m[5], yf, m[6]));
      const float rz = fmaf(m[8], xf, fmaf(m[9], yf, m[10]));
      float px, py;
      project(m, rx, ry, rz, dep, px, py);
      const Tap t = make_tap(px, py, h, w);
      if (t.any) {
        const float* src = feats + (int64_t)(v + 1) * hw * C;
        const float4 a = ldg4(src + (int64_t)t.off00 * C);
        const float4 bb = ldg4(src + (int64_t)t.off01 * C);
        const float4 c = ldg4(src + (int64_t)t.off10 * C);
        const float4 e = ldg4(src + (int64_t)t.off11 * C);
        float4 f;
        f.x = fmaf(e.x, t.w11, fmaf(c.x, t.w10, fmaf(bb.x, t.w01, a.x * t.w00)));
        f.y = fmaf(e.y, t.w11, fmaf(c.y, t.w10, fmaf(bb.y, t.w01, a.y * t.w00)));
        f.z = fmaf(e.z, t.w11, fmaf(c.z, t.w10, fmaf(bb.z, t.w01, a.z * t.w00)));
        f.w = fmaf(e.w, t.w11, fmaf(c.w, t.w10, fmaf(bb.w, t.w01, a.w * t.w00)));
        s.x += f.x; s.y += f.y; s.z += f.z; s.w += f.w;
        s2.x = fmaf(f.x, f.x, s2.x); s2.y = fmaf(f.y, f.y, s2.y);
        s2.z = fmaf(f.z, f.z, s2.z); s2.w = fmaf(f.w, f.w, s2.w);
        if (NOREF) {
          n.x += f.x; n.y += f.y; n.z += f.z; n.w += f.w;
          n2.x = fmaf(f.x, f.x, n2.x); n2.y = fmaf(f.y, f.y, n2.y);
          n2.z = fmaf(f.z, f.z, n2.z); n2.w = fmaf(f.w, f.w, n2.w);
        }
      }
    }
    // var = Σf²/V − (Σf/V)²   (casmvsnet.py:99)
    float4 o;
    float mx = s.x * inv_v, my = s.y * inv_v, mz = s.z * inv_v, mw = s.w * inv_v;
    o.x = fmaf(-mx, mx, s2.x * inv_v);
    o.y = fmaf(-my, my, s2.y * inv_v);
    o.z = fmaf(-mz, mz, s2.z * inv_v);
    o.w = fmaf(-mw, mw, s2.w * inv_v);
    const int64_t oidx = (((int64_t)d * h + y) * w + x) * C;
    st4_stream(var + oidx, o);
    if (NOREF) {
      mx = n.x * inv_v; my = n.y * inv_v; mz = n.z * inv_v; mw = n.w * inv_v;
      o.x = fmaf(-mx, mx, n2.x * inv_v);
      o.y = fmaf(-my, my, n2.y * inv_v);
      o.z = fmaf(-mz, mz, n2.z * inv_v);
      o.w = fmaf(-mw, mw, n2.w * inv_v);
      st4_stream(var_noref + oidx, o);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// v2 of the fused sweep (eval path, no var_noref): ncu on v1 (profiles/r01_costvolume_v1_ncu.md) showed it is
// instruction-issue bound — every one of the C/4 lanes of a pixel redid the projection, the IEEE divisions and the tap
// setup for every (plane, view).  Here the LPP = C/4 lanes of a pixel each take ONE plane of a group of LPP planes,
// compute its sample position once, and the group then walks the LPP planes broadcasting (tap base | flags, ax, ay)
// with three width-LPP shuffles.  Σf, Σf² for the LPP planes stay in registers.
// ------------------------------------------------------------------------------------------------------------------
template <int C, bool PER_PIXEL>
__global__ void __launch_bounds__(256)
costvolume_var_v2_kernel(const float* __restrict__ feats, const float* __restrict__ rt_g,
                         const float* __restrict__ depth, float* __restrict__ var, int V, int D, int h, int w,
                         int d_chunk, int chunks_per_batch) {
  constexpr int LPP = C / 4;
  constexpr int PXB = 256 / LPP;
  const int q = threadIdx.x % LPP;
  const int xr = blockIdx.x * PXB + threadIdx.x / LPP;
  const int x = min(xr, w - 1);                 // out-of-range lanes still take part in the shuffles
  const int y = blockIdx.y;
  const int b = blockIdx.z / chunks_per_batch;
  const int d_begin = (blockIdx.z % chunks_per_batch) * d_chunk;
  const int d_end = min(D, d_begin + d_chunk);
  __shared__ float rt_s[kMaxSrc * 12];
  for (int i = threadIdx.x; i < (V - 1) * 12; i += 256) rt_s[i] = __ldg(rt_g + (int64_t)b * (V - 1) * 12 + i);
  __syncthreads();

  const int64_t hw = (int64_t)h * w;
  feats += (int64_t)b * V * hw * C + q * 4;
  var += (int64_t)b * D * hw * C + q * 4;
  depth += PER_PIXEL ? (int64_t)b * D * hw + (int64_t)y * w + x : (int64_t)b * D;
  const float4 ref = ldg4(feats + ((int64_t)y * w + x) * C);
  const float4 ref2 = make_float4(ref.x * ref.x, ref.y * ref.y, ref.z * ref.z, ref.w * ref.w);
  const float inv_v = 1.f / (float)V;
  const float xf = (float)x, yf = (float)y;

  for (int dg = d_begin; dg < d_end; dg += LPP) {
    float4 s[LPP], s2[LPP];
#pragma unroll
    for (int j = 0; j < LPP; ++j) { s[j] = ref; s2[j] = ref2; }
    const int my_d = min(dg + q, D - 1);
    const float dep = PER_PIXEL ? __ldg(depth + (int64_t)my_d * hw) : __ldg(depth + my_d);
    for (int v = 0; v < V - 1; ++v) {
      const float* m = rt_s + v * 12;
      // this lane's plane: p = rot·(x,y,1)·depth + trans ; (px,py) = p.xy / p.z   (modules.py:325-329)
      const float rx = fmaf(m[0], xf, fmaf(m[1], yf, m[2]));
      const float ry = fmaf(m[4], xf, fmaf(m[5], yf, m[6]));
      const float rz = fmaf(m[8], xf, fmaf(m[9], yf, m[10]));
      const float iz = __frcp_rn(fmaf(rz, dep, m[11]));
      const float px = fmaf(rx, dep, m[3]) * iz, py = fmaf(ry, dep, m[7]) * iz;
      const bool finite = (fabsf(px) < 1e8f) && (fabsf(py) < 1e8f);
      const float fx = floorf(px), fy = floorf(py);
      const float my_ax = px - fx, my_ay = py - fy;
      const int x0 = finite ? (int)fx : -4, y0 = finite ? (int)fy : -4;
      const bool vx0 = (x0 >= 0) && (x0 < w), vx1 = (x0 + 1 >= 0) && (x0 + 1 < w);
      const bool vy0 = (y0 >= 0) && (y0 < h), vy1 = (y0 + 1 >= 0) && (y0 + 1 < h);
      const int cx0 = min(max(x0, 0), w - 1), cx1 = min(max(x0 + 1, 0), w - 1);
      const int cy0 = min(max(y0, 0), h - 1), cy1 = min(max(y0 + 1, 0), h - 1);
      // tap base (< 2^24) | dx<<24 | dy<<25 | valid00<<26 | valid01<<27 | valid10<<28 | valid11<<29
      const unsigned my_pk = (unsigned)(cy0 * w + cx0) | ((unsigned)(cx1 != cx0) << 24) | ((unsigned)(cy1 != cy0) << 25) |
                             ((unsigned)(vx0 && vy0) << 26) | ((unsigned)(vx1 && vy0) << 27) |
                             ((unsigned)(vx0 && vy1) << 28) | ((unsigned)(vx1 && vy1) << 29);
      const float* src = feats + (int64_t)(v + 1) * hw * C;
#pragma unroll
      for (int j = 0; j < LPP; ++j) {
        const unsigned pk = __shfl_sync(0xffffffffu, my_pk, j, LPP);
        const float ax = __shfl_sync(0xffffffffu, my_ax, j, LPP);
        const float ay = __shfl_sync(0xffffffffu, my_ay, j, LPP);
        if (pk >> 26) {
          const int base = (int)(pk & 0xFFFFFFu);
          const int o01 = base + (int)((pk >> 24) & 1u);
          const int o10 = base + (((pk >> 25) & 1u) ? w : 0);
          const int o11 = o10 + (int)((pk >> 24) & 1u);
          const float bx = 1.f - ax, by = 1.f - ay;
          const float w00 = (pk & (1u << 26)) ? bx * by : 0.f, w01 = (pk & (1u << 27)) ? ax * by : 0.f;
          const float w10 = (pk & (1u << 28)) ? bx * ay : 0.f, w11 = (pk & (1u << 29)) ? ax * ay : 0.f;
          const float4 a = ldg4(src + (int64_t)base * C), bb = ldg4(src + (int64_t)o01 * C);
          const float4 c = ldg4(src + (int64_t)o10 * C), e = ldg4(src + (int64_t)o11 * C);
          float4 f;
          f.x = fmaf(e.x, w11, fmaf(c.x, w10, fmaf(bb.x, w01, a.x * w00)));
          f.y = fmaf(e.y, w11, fmaf(c.y, w10, fmaf(bb.y, w01, a.y * w00)));
          f.z = fmaf(e.z, w11, fmaf(c.z, w10, fmaf(bb.z, w01, a.z * w00)));
          f.w = fmaf(e.w, w11, fmaf(c.w, w10, fmaf(bb.w, w01, a.w * w00)));
          s[j].x += f.x; s[j].y += f.y; s[j].z += f.z; s[j].w += f.w;
          s2[j].x = fmaf(f.x, f.x, s2[j].x); s2[j].y = fmaf(f.y, f.y, s2[j].y);
          s2[j].z = fmaf(f.z, f.z, s2[j].z); s2[j].w = fmaf(f.w, f.w, s2[j].w);
        }
      }
    }
    if (xr < w) {
#pragma unroll
      for (int j = 0; j < LPP; ++j) {
        const int d = dg + j;
        if (d < d_end) {
          float4 o;
          const float mx = s[j].x * inv_v, my = s[j].y * inv_v, mz = s[j].z * inv_v, mw = s[j].w * inv_v;
          o.x = fmaf(-mx, mx, s2[j].x * inv_v);
          o.y = fmaf(-my, my, s2[j].y * inv_v);
          o.z = fmaf(-mz, mz, s2[j].z * inv_v);
          o.w = fmaf(-mw, mw, s2[j].w * inv_v);
          st4_stream(var + (((int64_t)d * h + y) * w + x) * C, o);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// v3 of the fused sweep (round 2).  ncu on v2 (profiles/r01_costvolume_final_ncu.md): 329 M warp instructions for 110 M
// essential ones (4 LDG.128 + 24 FP per float4, view and plane) — instruction-issue bound at 2.6 IPC, 62 % L1.  v3 keeps
// the plane-distributed mapping (lane q of a pixel owns plane dg+q and broadcasts its sample) and removes work per tap:
//   * the owning lane folds validity, clamping and the bilinear weights into FOUR final tap weights and ONE in-bounds
//     base texel (the 2x2 quad is shifted inside the image at the borders and the weights move with it), so the LPP-1
//     other lanes do no unpacking / predicate / weight arithmetic: 5 shuffles, 1 address, 4 loads at constant offsets
//     {0, C, w·C, w·C + C} from the base, 24 FP;
//   * consecutive planes of a group that fall into the same texel quad (disparity between neighbouring planes is
//     0.1-0.5 px on DTU geometry) reuse the four texels from registers — fewer L1 wavefronts (the 128 B/clk/SM L1 path
//     bounds v2-style designs at ~0.22 ms for stage 1);
//   * the perspective division is an IEEE division again (v2 multiplied by a reciprocal: up to 2 ulp of the coordinate
//     away from the reference's arithmetic).
// OUT_SPLIT: write the variance as two bf16 planes [2][B][D][h][w][C] (hi, lo) — the operand format of the TMA-fed
// tensor-core convolution that consumes it (csrc/conv3d_slab_bf16.cu) — instead of fp32; same bytes.
// ------------------------------------------------------------------------------------------------------------------
template <int C, bool PER_PIXEL, bool OUT_SPLIT>
__global__ void __launch_bounds__(256)
costvolume_var_v3_kernel(const float* __restrict__ feats, const float* __restrict__ rt_g,
                         const float* __restrict__ depth, float* __restrict__ var, __nv_bfloat16* __restrict__ var_split,
                         int B, int V, int D, int h, int w, int d_chunk, int chunks_per_batch) {
  constexpr int LPP = C / 4;
  constexpr int PXB = 256 / LPP;
  const int q = threadIdx.x % LPP;
  const int xr = blockIdx.x * PXB + threadIdx.x / LPP;
  const int x = min(xr, w - 1);                 // out-of-range lanes still take part in the shuffles
  const int y = blockIdx.y;
  const int b = blockIdx.z / chunks_per_batch;
  const int d_begin = (blockIdx.z % chunks_per_batch) * d_chunk;
  const int d_end = min(D, d_begin + d_chunk);
  __shared__ float rt_s[kMaxSrc * 12];
  for (int i = threadIdx.x; i < (V - 1) * 12; i += 256) rt_s[i] = __ldg(rt_g + (int64_t)b * (V - 1) * 12 + i);
  __syncthreads();

  const int64_t hw = (int64_t)h * w;
  feats += (int64_t)b * V * hw * C + q * 4;
  depth += PER_PIXEL ? (int64_t)b * D * hw + (int64_t)y * w + x : (int64_t)b * D;
  const float4 ref = ldg4(feats + ((int64_t)y * w + x) * C);
  const float4 ref2 = make_float4(ref.x * ref.x, ref.y * ref.y, ref.z * ref.z, ref.w * ref.w);
  const float inv_v = 1.f / (float)V;
  const float xf = (float)x, yf = (float)y;
  const float* feats1 = feats + w * C;          // the quad's second row

  for (int dg = d_begin; dg < d_end; dg += LPP) {
    float2 s[LPP][2], s2[LPP][2];
#pragma unroll
    for (int j = 0; j < LPP; ++j) {
      s[j][0] = make_float2(ref.x, ref.y); s[j][1] = make_float2(ref.z, ref.w);
      s2[j][0] = make_float2(ref2.x, ref2.y); s2[j][1] = make_float2(ref2.z, ref2.w);
    }
    const int my_d = min(dg + q, D - 1);
    const float dep = PER_PIXEL ? __ldg(depth + (int64_t)my_d * hw) : __ldg(depth + my_d);
    for (int v = 0; v < V - 1; ++v) {
      const float* m = rt_s + v * 12;
      // this lane's plane: p = rot·(x,y,1)·depth + trans ; (px,py) = p.xy / p.z   (modules.py:325-329)
      const float rx = fmaf(m[0], xf, fmaf(m[1], yf, m[2]));
      const float ry = fmaf(m[4], xf, fmaf(m[5], yf, m[6]));
      const float rz = fmaf(m[8], xf, fmaf(m[9], yf, m[10]));
      const float Z = fmaf(rz, dep, m[11]);
      const float px = __fdiv_rn(fmaf(rx, dep, m[3]), Z), py = __fdiv_rn(fmaf(ry, dep, m[7]), Z);
      const bool finite = (fabsf(px) < 1e8f) && (fabsf(py) < 1e8f);     // false for NaN
      const float fx = floorf(px), fy = floorf(py);
      const float ax = px - fx, ay = py - fy;
      const int x0 = finite ? (int)fx : -4, y0 = finite ? (int)fy : -4;
      // shift the 2x2 quad inside the image; the per-axis weights of its two columns / rows move with it:
      //   x0 in [0, w-2]: (1-ax, ax);  x0 = -1: only tap x1 = 0 is valid -> (ax, 0);  x0 = w-1: only tap x0 -> (0, 1-ax)
      const int bx = min(max(x0, 0), w - 2), by = min(max(y0, 0), h - 2);
      const float wx0 = (x0 == bx) ? 1.f - ax : ((x0 == -1) ? ax : 0.f);
      const float wx1 = (x0 == bx) ? ax : ((x0 == w - 1) ? 1.f - ax : 0.f);
      const float wy0 = (y0 == by) ? 1.f - ay : ((y0 == -1) ? ay : 0.f);
      const float wy1 = (y0 == by) ? ay : ((y0 == h - 1) ? 1.f - ay : 0.f);
      const float my_w00 = wy0 * wx0, my_w01 = wy0 * wx1, my_w10 = wy1 * wx0, my_w11 = wy1 * wx1;
      // (all four taps outside -> all weights are 0 and the base stays inside the image: no branch needed)
      // element offset of the quad's first texel from `feats` (the view offset is folded in: < 2^31, checked by the launcher)
      const unsigned my_base = (unsigned)(v + 1) * (unsigned)hw * C + (unsigned)(by * w + bx) * C;
      unsigned prev = 0xffffffffu;
      float2 a0 = make_float2(0.f, 0.f), a1 = a0, b0 = a0, b1 = a0, c0 = a0, c1 = a0, e0 = a0, e1 = a0;
#pragma unroll
      for (int j = 0; j < LPP; ++j) {
        const unsigned base = __shfl_sync(0xffffffffu, my_base, j, LPP);
        const float w00 = __shfl_sync(0xffffffffu, my_w00, j, LPP);
        const float w01 = __shfl_sync(0xffffffffu, my_w01, j, LPP);
        const float w10 = __shfl_sync(0xffffffffu, my_w10, j, LPP);
        const float w11 = __shfl_sync(0xffffffffu, my_w11, j, LPP);
        if (base != prev) {                      // neighbouring planes mostly sample the same texel quad
          const float* t = feats + base;         // two 64-bit address formations per reload (rows y0 and y0+1)
          const float* t1 = feats1 + base;
          const float4 a = ldg4(t), bb = ldg4(t + C), c = ldg4(t1), e = ldg4(t1 + C);
          a0 = make_float2(a.x, a.y); a1 = make_float2(a.z, a.w);
          b0 = make_float2(bb.x, bb.y); b1 = make_float2(bb.z, bb.w);
          c0 = make_float2(c.x, c.y); c1 = make_float2(c.z, c.w);
          e0 = make_float2(e.x, e.y); e1 = make_float2(e.z, e.w);
          prev = base;
        }
        // packed fp32x2 arithmetic (FFMA2 / FADD2 / FMUL2 on sm_100: two IEEE fp32 results per instruction)
        const float2 q00 = make_float2(w00, w00), q01 = make_float2(w01, w01), q10 = make_float2(w10, w10),
                     q11 = make_float2(w11, w11);
        const float2 f0 = __ffma2_rn(e0, q11, __ffma2_rn(c0, q10, __ffma2_rn(b0, q01, __fmul2_rn(a0, q00))));
        const float2 f1 = __ffma2_rn(e1, q11, __ffma2_rn(c1, q10, __ffma2_rn(b1, q01, __fmul2_rn(a1, q00))));
        s[j][0] = __fadd2_rn(s[j][0], f0);
        s[j][1] = __fadd2_rn(s[j][1], f1);
        s2[j][0] = __ffma2_rn(f0, f0, s2[j][0]);
        s2[j][1] = __ffma2_rn(f1, f1, s2[j][1]);
      }
    }
    if (xr < w) {
#pragma unroll
      for (int j = 0; j < LPP; ++j) {
        const int d = dg + j;
        if (d < d_end) {
          float4 o;
          const float mx = s[j][0].x * inv_v, my = s[j][0].y * inv_v, mz = s[j][1].x * inv_v, mw = s[j][1].y * inv_v;
          o.x = fmaf(-mx, mx, s2[j][0].x * inv_v);
          o.y = fmaf(-my, my, s2[j][0].y * inv_v);
          o.z = fmaf(-mz, mz, s2[j][1].x * inv_v);
          o.w = fmaf(-mw, mw, s2[j][1].y * inv_v);
          const int64_t oidx = ((((int64_t)b * D + d) * h + y) * w + x) * C + q * 4;
          if (OUT_SPLIT) {
            const __nv_bfloat16 h0 = __float2bfloat16_rn(o.x), h1 = __float2bfloat16_rn(o.y);
            const __nv_bfloat16 h2 = __float2bfloat16_rn(o.z), h3 = __float2bfloat16_rn(o.w);
            const __nv_bfloat162 hi01 = __halves2bfloat162(h0, h1), hi23 = __halves2bfloat162(h2, h3);
            const __nv_bfloat162 lo01 = __floats2bfloat162_rn(o.x - __bfloat162float(h0), o.y - __bfloat162float(h1));
            const __nv_bfloat162 lo23 = __floats2bfloat162_rn(o.z - __bfloat162float(h2), o.w - __bfloat162float(h3));
            uint2 hv, lv;
            hv.x = *reinterpret_cast<const uint32_t*>(&hi01); hv.y = *reinterpret_cast<const uint32_t*>(&hi23);
            lv.x = *reinterpret_cast<const uint32_t*>(&lo01); lv.y = *reinterpret_cast<const uint32_t*>(&lo23);
            __stcs(reinterpret_cast<uint2*>(var_split + oidx), hv);
            __stcs(reinterpret_cast<uint2*>(var_split + (int64_t)B * D * hw * C + oidx), lv);
          } else {
            st4_stream(var + oidx, o);
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// v4: the v3 kernel with TWO channel quads per lane (quads q and q + LPP, LPP = C/8 lanes per pixel): the per-plane
// bookkeeping of a lane group (5 shuffles, the quad-reuse test, address formation) is shared by twice the channels, and
// with 8 channels (stage 3) a pixel is a single lane — no shuffles at all.  Each of a lane's two 16-byte loads per tap
// still falls into a fully used 64-byte run of the lane group (quads q and q + LPP are LPP·16 bytes apart), so the
// sector efficiency of v3 is kept.  Per channel the arithmetic is that of v3 / v1 (bitwise, tested).
// ------------------------------------------------------------------------------------------------------------------
template <int C, bool PER_PIXEL, bool OUT_SPLIT>
__global__ void __launch_bounds__(128)
costvolume_var_v4_kernel(const float* __restrict__ feats, const float* __restrict__ rt_g,
                         const float* __restrict__ depth, float* __restrict__ var, __nv_bfloat16* __restrict__ var_split,
                         int B, int V, int D, int h, int w, int d_chunk, int chunks_per_batch) {
  constexpr int LPP = C / 8;                    // lanes per pixel
  constexpr int PXB = 128 / LPP;
  constexpr int QB = LPP * 4;                   // channel offset of a lane's second quad
  const int q = threadIdx.x % LPP;
  const int xr = blockIdx.x * PXB + threadIdx.x / LPP;
  const int x = min(xr, w - 1);                 // out-of-range lanes still take part in the shuffles
  const int y = blockIdx.y;
  const int b = blockIdx.z / chunks_per_batch;
  const int d_begin = (blockIdx.z % chunks_per_batch) * d_chunk;
  const int d_end = min(D, d_begin + d_chunk);
  __shared__ float rt_s[kMaxSrc * 12];
  for (int i = threadIdx.x; i < (V - 1) * 12; i += 128) rt_s[i] = __ldg(rt_g + (int64_t)b * (V - 1) * 12 + i);
  __syncthreads();

  const int64_t hw = (int64_t)h * w;
  feats += (int64_t)b * V * hw * C + q * 4;
  depth += PER_PIXEL ? (int64_t)b * D * hw + (int64_t)y * w + x : (int64_t)b * D;
  const float4 refa = ldg4(feats + ((int64_t)y * w + x) * C), refb = ldg4(feats + ((int64_t)y * w + x) * C + QB);
  const float inv_v = 1.f / (float)V;
  const float xf = (float)x, yf = (float)y;
  const float* feats1 = feats + w * C;          // the quad's second row

  for (int dg = d_begin; dg < d_end; dg += LPP) {
    // [plane][0,1: quad a (xy, zw)  2,3: quad b]
    float2 s[LPP][4], s2[LPP][4];
#pragma unroll
    for (int j = 0; j < LPP; ++j) {
      s[j][0] = make_float2(refa.x, refa.y); s[j][1] = make_float2(refa.z, refa.w);
      s[j][2] = make_float2(refb.x, refb.y); s[j][3] = make_float2(refb.z, refb.w);
      s2[j][0] = make_float2(refa.x * refa.x, refa.y * refa.y); s2[j][1] = make_float2(refa.z * refa.z, refa.w * refa.w);
      s2[j][2] = make_float2(refb.x * refb.x, refb.y * refb.y); s2[j][3] = make_float2(refb.z * refb.z, refb.w * refb.w);
    }
    const int my_d = min(dg + q, D - 1);
    const float dep = PER_PIXEL ? __ldg(depth + (int64_t)my_d * hw) : __ldg(depth + my_d);
    for (int v = 0; v < V - 1; ++v) {
      const float* m = rt_s + v * 12;
      // this lane's plane: p = rot·(x,y,1)·depth + trans ; (px,py) = p.xy / p.z   (modules.py:325-329)
      const float rx = fmaf(m[0], xf, fmaf(m[1], yf, m[2]));
      const float ry = fmaf(m[4], xf, fmaf(m[5], yf, m[6]));
      const float rz = fmaf(m[8], xf, fmaf(m[9], yf, m[10]));
      const float Z = fmaf(rz, dep, m[11]);
      const float px = __fdiv_rn(fmaf(rx, dep, m[3]), Z), py = __fdiv_rn(fmaf(ry, dep, m[7]), Z);
      const bool finite = (fabsf(px) < 1e8f) && (fabsf(py) < 1e8f);     // false for NaN
      const float fx = floorf(px), fy = floorf(py);
      const float ax = px - fx, ay = py - fy;
      const int x0 = finite ? (int)fx : -4, y0 = finite ? (int)fy : -4;
      const int bx = min(max(x0, 0), w - 2), by = min(max(y0, 0), h - 2);
      const float wx0 = (x0 == bx) ? 1.f - ax : ((x0 == -1) ? ax : 0.f);
      const float wx1 = (x0 == bx) ? ax : ((x0 == w - 1) ? 1.f - ax : 0.f);
      const float wy0 = (y0 == by) ? 1.f - ay : ((y0 == -1) ? ay : 0.f);
      const float wy1 = (y0 == by) ? ay : ((y0 == h - 1) ? 1.f - ay : 0.f);
      const float my_w00 = wy0 * wx0, my_w01 = wy0 * wx1, my_w10 = wy1 * wx0, my_w11 = wy1 * wx1;
      const unsigned my_base = (unsigned)(v + 1) * (unsigned)hw * C + (unsigned)(by * w + bx) * C;
      unsigned prev = 0xffffffffu;
      float2 t00[4], t01[4], t10[4], t11[4];    // the four texels: [0,1] quad a, [2,3] quad b
#pragma unroll
      for (int k = 0; k < 4; ++k) t00[k] = t01[k] = t10[k] = t11[k] = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < LPP; ++j) {
        const unsigned base = LPP == 1 ? my_base : __shfl_sync(0xffffffffu, my_base, j, LPP);
        const float w00 = LPP == 1 ? my_w00 : __shfl_sync(0xffffffffu, my_w00, j, LPP);
        const float w01 = LPP == 1 ? my_w01 : __shfl_sync(0xffffffffu, my_w01, j, LPP);
        const float w10 = LPP == 1 ? my_w10 : __shfl_sync(0xffffffffu, my_w10, j, LPP);
        const float w11 = LPP == 1 ? my_w11 : __shfl_sync(0xffffffffu, my_w11, j, LPP);
        if (base != prev) {                      // neighbouring planes mostly sample the same texel quad
          const float* t = feats + base;
          const float* t1 = feats1 + base;
          const float4 a = ldg4(t), bb = ldg4(t + C), c = ldg4(t1), e = ldg4(t1 + C);
          const float4 a2 = ldg4(t + QB), bb2 = ldg4(t + C + QB), c2 = ldg4(t1 + QB), e2 = ldg4(t1 + C + QB);
          t00[0] = make_float2(a.x, a.y); t00[1] = make_float2(a.z, a.w); t00[2] = make_float2(a2.x, a2.y); t00[3] = make_float2(a2.z, a2.w);
          t01[0] = make_float2(bb.x, bb.y); t01[1] = make_float2(bb.z, bb.w); t01[2] = make_float2(bb2.x, bb2.y); t01[3] = make_float2(bb2.z, bb2.w);
          t10[0] = make_float2(c.x, c.y); t10[1] = make_float2(c.z, c.w); t10[2] = make_float2(c2.x, c2.y); t10[3] = make_float2(c2.z, c2.w);
          t11[0] = make_float2(e.x, e.y); t11[1] = make_float2(e.z, e.w); t11[2] = make_float2(e2.x, e2.y); t11[3] = make_float2(e2.z, e2.w);
          prev = base;
        }
        const float2 q00 = make_float2(w00, w00), q01 = make_float2(w01, w01), q10 = make_float2(w10, w10),
                     q11 = make_float2(w11, w11);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __ffma2_rn(t11[k], q11, __ffma2_rn(t10[k], q10, __ffma2_rn(t01[k], q01, __fmul2_rn(t00[k], q00))));
          s[j][k] = __fadd2_rn(s[j][k], f);
          s2[j][k] = __ffma2_rn(f, f, s2[j][k]);
        }
      }
    }
    if (xr < w) {
#pragma unroll
      for (int j = 0; j < LPP; ++j) {
        const int d = dg + j;
        if (d < d_end) {
#pragma unroll
          for (int hq = 0; hq < 2; ++hq) {
            float4 o;
            const float mx = s[j][2 * hq].x * inv_v, my = s[j][2 * hq].y * inv_v, mz = s[j][2 * hq + 1].x * inv_v,
                        mw = s[j][2 * hq + 1].y * inv_v;
            o.x = fmaf(-mx, mx, s2[j][2 * hq].x * inv_v);
            o.y = fmaf(-my, my, s2[j][2 * hq].y * inv_v);
            o.z = fmaf(-mz, mz, s2[j][2 * hq + 1].x * inv_v);
            o.w = fmaf(-mw, mw, s2[j][2 * hq + 1].y * inv_v);
            const int64_t oidx = ((((int64_t)b * D + d) * h + y) * w + x) * C + q * 4 + hq * QB;
            if (OUT_SPLIT) {
              const __nv_bfloat16 h0 = __float2bfloat16_rn(o.x), h1 = __float2bfloat16_rn(o.y);
              const __nv_bfloat16 h2 = __float2bfloat16_rn(o.z), h3 = __float2bfloat16_rn(o.w);
              const __nv_bfloat162 hi01 = __halves2bfloat162(h0, h1), hi23 = __halves2bfloat162(h2, h3);
              const __nv_bfloat162 lo01 = __floats2bfloat162_rn(o.x - __bfloat162float(h0), o.y - __bfloat162float(h1));
              const __nv_bfloat162 lo23 = __floats2bfloat162_rn(o.z - __bfloat162float(h2), o.w - __bfloat162float(h3));
              uint2 hv, lv;
              hv.x = *reinterpret_cast<const uint32_t*>(&hi01); hv.y = *reinterpret_cast<const uint32_t*>(&hi23);
              lv.x = *reinterpret_cast<const uint32_t*>(&lo01); lv.y = *reinterpret_cast<const uint32_t*>(&lo23);
              __stcs(reinterpret_cast<uint2*>(var_split + oidx), hv);
              __stcs(reinterpret_cast<uint2*>(var_split + (int64_t)B * D * hw * C + oidx), lv);
            } else {
              st4_stream(var + oidx, o);
            }
          }
        }
      }
    }
  }
}

// A/B switch for profiling sessions: RCMVS_SWEEP=v2 runs the round-1 kernel on the fp32-output eval path
// default kernel per channel count (measured): see kSweepDefault
static int kSweepDefault(int C) { return 3; }
// 0 = policy below; 2 / 3 / 4 force a kernel (tests and tools set it through ctypes in_dll; not part of the ABI)
extern "C" int rcmvs_sweep_variant = 0;
static int sweep_variant_for(int C) {        // RCMVS_SWEEP=v2|v3|v4 forces one kernel for every stage
  if (rcmvs_sweep_variant) return rcmvs_sweep_variant;
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RCMVS_SWEEP");
    v = (e && e[0] == 'v' && e[1] >= '2' && e[1] <= '4') ? e[1] - '0' : 0;
  }
  if (v) return v;
  return kSweepDefault(C);
}


template <int C>
static int launch_costvolume(const float* feats, const float* rt, const float* depth, int per_pixel, float* var,
                             float* var_noref, void* var_split, int B, int V, int D, int h, int w, cudaStream_t st) {
  constexpr int PXB = 256 / (C / 4);
  // enough CTAs for >= ~4 waves of 148 SMs x 8 resident CTAs, but >= 8 planes per CTA to amortise the setup
  const int xy_blocks = ceil_div(w, PXB) * h * B;
  int chunks = max(1, min(ceil_div(D, 8), ceil_div(148 * 8 * 4, xy_blocks)));
  const int d_chunk = ceil_div(D, chunks);
  chunks = ceil_div(D, d_chunk);
  dim3 grid(ceil_div(w, PXB), h, B * chunks);
  const bool hw_fits_24bit = (int64_t)h * w < (1 << 24);
#define RCMVS_CV_LAUNCH(PP, NR)                                                                              \
  costvolume_var_kernel<C, PP, NR><<<grid, 256, 0, st>>>(feats, rt, depth, var, var_noref, V, D, h, w, d_chunk, \
                                                         chunks)
  if (!var_noref && h >= 2 && w >= 2 && (int64_t)h * w * C < (1ll << 31) && (var_split || sweep_variant_for(C) != 2)) {
    // eval path: plane-distributed v3 kernel; planes per CTA rounded to the group size
    constexpr int LPP = C / 4;
    const int dch = ceil_div(d_chunk, LPP) * LPP;
    const int nch = ceil_div(D, dch);
    dim3 grid2(ceil_div(w, PXB), h, B * nch);
    __nv_bfloat16* vs = reinterpret_cast<__nv_bfloat16*>(var_split);
#define RCMVS_V3(PP, SP) costvolume_var_v3_kernel<C, PP, SP><<<grid2, 256, 0, st>>>(feats, rt, depth, var, vs, B, V, D, h, w, dch, nch)
    if (sweep_variant_for(C) == 4) {
      // two channel quads per lane: C/8 lanes per pixel, C/8 planes per iteration, 128 threads = the same pixels per CTA
      constexpr int LPP4 = C / 8;
      const int dch4 = ceil_div(d_chunk, LPP4) * LPP4;
      const int nch4 = ceil_div(D, dch4);
      dim3 grid4(ceil_div(w, 128 / LPP4), h, B * nch4);
#define RCMVS_V4(PP, SP) costvolume_var_v4_kernel<C, PP, SP><<<grid4, 128, 0, st>>>(feats, rt, depth, var, vs, B, V, D, h, w, dch4, nch4)
      if (per_pixel) { if (vs) RCMVS_V4(true, true); else RCMVS_V4(true, false); }
      else { if (vs) RCMVS_V4(false, true); else RCMVS_V4(false, false); }
#undef RCMVS_V4
    } else if (per_pixel) { if (vs) RCMVS_V3(true, true); else RCMVS_V3(true, false); }
    else { if (vs) RCMVS_V3(false, true); else RCMVS_V3(false, false); }
#undef RCMVS_V3
  } else if (var_split) {
    set_error("rcmvs_costvolume_var_fwd_split: needs h, w >= 2 and no var_noref");
    return RCMVS_ERR_UNSUPPORTED;
  } else if (!var_noref && hw_fits_24bit) {
    // round-1 v2 kernel (degenerate 1-pixel-wide shapes, RCMVS_SWEEP=v2)
    constexpr int LPP = C / 4;
    const int dch = ceil_div(d_chunk, LPP) * LPP;
    const int nch = ceil_div(D, dch);
    dim3 grid2(ceil_div(w, PXB), h, B * nch);
    if (per_pixel) costvolume_var_v2_kernel<C, true><<<grid2, 256, 0, st>>>(feats, rt, depth, var, V, D, h, w, dch, nch);
    else costvolume_var_v2_kernel<C, false><<<grid2, 256, 0, st>>>(feats, rt, depth, var, V, D, h, w, dch, nch);
  } else if (per_pixel) {
    if (var_noref) RCMVS_CV_LAUNCH(true, true); else RCMVS_CV_LAUNCH(true, false);
  } else {
    if (var_noref) RCMVS_CV_LAUNCH(false, true); else RCMVS_CV_LAUNCH(false, false);
  }
#undef RCMVS_CV_LAUNCH
  RCMVS_CHECK_LAUNCH("rcmvs_costvolume_var_fwd");
  return RCMVS_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Standalone homo_warping (API parity with the reference function): NCHW in, NCDHW out, any C.
// One thread per (x, y, d); coordinates computed once and reused for all channels; stores are 128 B coalesced rows.
// ------------------------------------------------------------------------------------------------------------------
template <bool PER_PIXEL>
__global__ void __launch_bounds__(128)
homo_warp_fwd_kernel(const float* __restrict__ src, const float* __restrict__ rt_g, const float* __restrict__ depth,
                     float* __restrict__ out, int C, int D, int h, int w) {
  const int x = blockIdx.x * 128 + threadIdx.x;
  const int y = blockIdx.y;
  const int b = blockIdx.z / D, d = blockIdx.z % D;
  if (x >= w) return;
  const int64_t hw = (int64_t)h * w;
  const float* m = rt_g + (int64_t)b * 12;
  const float xf = (float)x, yf = (float)y;
  const float rx = fmaf(__ldg(m + 0), xf, fmaf(__ldg(m + 1), yf, __ldg(m + 2)));
  const float ry = fmaf(__ldg(m + 4), xf, fmaf(__ldg(m + 5), yf, __ldg(m + 6)));
  const float rz = fmaf(__ldg(m + 8), xf, fmaf(__ldg(m + 9), yf, __ldg(m + 10)));
  const float dep = PER_PIXEL ? __ldg(depth + ((int64_t)b * D + d) * hw + (int64_t)y * w + x)
                              : __ldg(depth + (int64_t)b * D + d);
  const float px = __fdiv_rn(fmaf(rx, dep, __ldg(m + 3)), fmaf(rz, dep, __ldg(m + 11)));
  const float py = __fdiv_rn(fmaf(ry, dep, __ldg(m + 7)), fmaf(rz, dep, __ldg(m + 11)));
  const Tap t = make_tap(px, py, h, w);
  src += (int64_t)b * C * hw;
  out += ((int64_t)b * C * D + d) * hw + (int64_t)y * w + x;
  if (!t.any) {
    for (int c = 0; c < C; ++c) __stcs(out + (int64_t)c * D * hw, 0.f);
    return;
  }
#pragma unroll 4
  for (int c = 0; c < C; ++c) {
    const float* p = src + (int64_t)c * hw;
    const float v = fmaf(__ldg(p + t.off11), t.w11,
                         fmaf(__ldg(p + t.off10), t.w10, fmaf(__ldg(p + t.off01), t.w01, __ldg(p + t.off00) * t.w00)));
    __stcs(out + (int64_t)c * D * hw, v);
  }
}

// Backward w.r.t. src_fea only (the grid is built under torch.no_grad in the reference, modules.py:314):
// grad_src[c, tap] += w_tap * grad_out[c, d, y, x]  — bilinear scatter-add with red.global.add.f32.
template <bool PER_PIXEL>
__global__ void __launch_bounds__(128)
homo_warp_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ rt_g, const float* __restrict__ depth,
                     float* __restrict__ gsrc, int C, int D, int h, int w) {
  const int x = blockIdx.x * 128 + threadIdx.x;
  const int y = blockIdx.y;
  const int b = blockIdx.z / D, d = blockIdx.z % D;
  if (x >= w) return;
  const int64_t hw = (int64_t)h * w;
  const float* m = rt_g + (int64_t)b * 12;
  const float xf = (float)x, yf = (float)y;
  const float rx = fmaf(__ldg(m + 0), xf, fmaf(__ldg(m + 1), yf, __ldg(m + 2)));
  const float ry = fmaf(__ldg(m + 4), xf, fmaf(__ldg(m + 5), yf, __ldg(m + 6)));
  const float rz = fmaf(__ldg(m + 8), xf, fmaf(__ldg(m + 9), yf, __ldg(m + 10)));
  const float dep = PER_PIXEL ? __ldg(depth + ((int64_t)b * D + d) * hw + (int64_t)y * w + x)
                              : __ldg(depth + (int64_t)b * D + d);
  const float px = __fdiv_rn(fmaf(rx, dep, __ldg(m + 3)), fmaf(rz, dep, __ldg(m + 11)));
  const float py = __fdiv_rn(fmaf(ry, dep, __ldg(m + 7)), fmaf(rz, dep, __ldg(m + 11)));
  const Tap t = make_tap(px, py, h, w);
  if (!t.any) return;
  gsrc += (int64_t)b * C * hw;
  gout += ((int64_t)b * C * D + d) * hw + (int64_t)y * w + x;
  for (int c = 0; c < C; ++c) {
    const float g = __ldg(gout + (int64_t)c * D * hw);
    float* p = gsrc + (int64_t)c * hw;
    if (t.w00 != 0.f) atomicAdd(p + t.off00, g * t.w00);
    if (t.w01 != 0.f) atomicAdd(p + t.off01, g * t.w01);
    if (t.w10 != 0.f) atomicAdd(p + t.off10, g * t.w10);
    if (t.w11 != 0.f) atomicAdd(p + t.off11, g * t.w11);
  }
}

}  // namespace rcmvs

extern "C" int rcmvs_costvolume_var_fwd(const float* feats, const float* rt, const float* depth, int per_pixel,
                                        float* var, float* var_noref, int B, int V, int C, int D, int h, int w,
                                        rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(feats && rt && depth && var, RCMVS_ERR_BAD_ARG, "rcmvs_costvolume_var_fwd: null pointer");
  RCMVS_REQUIRE(B > 0 && V > 1 && D > 0 && h > 0 && w > 0, RCMVS_ERR_BAD_ARG,
                "rcmvs_costvolume_var_fwd: bad dims B=%d V=%d D=%d h=%d w=%d", B, V, D, h, w);
  RCMVS_REQUIRE(V - 1 <= kMaxSrc, RCMVS_ERR_UNSUPPORTED, "rcmvs_costvolume_var_fwd: at most %d source views", kMaxSrc);
  RCMVS_REQUIRE(h <= 65535, RCMVS_ERR_UNSUPPORTED, "rcmvs_costvolume_var_fwd: h > 65535");
  cudaStream_t st = as_stream(stream);
  switch (C) {
    case 8: return launch_costvolume<8>(feats, rt, depth, per_pixel, var, var_noref, nullptr, B, V, D, h, w, st);
    case 16: return launch_costvolume<16>(feats, rt, depth, per_pixel, var, var_noref, nullptr, B, V, D, h, w, st);
    case 32: return launch_costvolume<32>(feats, rt, depth, per_pixel, var, var_noref, nullptr, B, V, D, h, w, st);
    default: break;
  }
  set_error("rcmvs_costvolume_var_fwd: C=%d not in {8,16,32}", C);
  return RCMVS_ERR_UNSUPPORTED;
}

extern "C" int rcmvs_costvolume_var_fwd_split(const float* feats, const float* rt, const float* depth, int per_pixel,
                                              void* var_split, int B, int V, int C, int D, int h, int w,
                                              rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(feats && rt && depth && var_split, RCMVS_ERR_BAD_ARG, "rcmvs_costvolume_var_fwd_split: null pointer");
  RCMVS_REQUIRE(B > 0 && V > 1 && D > 0 && h > 0 && w > 0, RCMVS_ERR_BAD_ARG,
                "rcmvs_costvolume_var_fwd_split: bad dims B=%d V=%d D=%d h=%d w=%d", B, V, D, h, w);
  RCMVS_REQUIRE(V - 1 <= kMaxSrc && h <= 65535, RCMVS_ERR_UNSUPPORTED, "rcmvs_costvolume_var_fwd_split: V or h too large");
  cudaStream_t st = as_stream(stream);
  switch (C) {
    case 8: return launch_costvolume<8>(feats, rt, depth, per_pixel, nullptr, nullptr, var_split, B, V, D, h, w, st);
    case 16: return launch_costvolume<16>(feats, rt, depth, per_pixel, nullptr, nullptr, var_split, B, V, D, h, w, st);
    case 32: return launch_costvolume<32>(feats, rt, depth, per_pixel, nullptr, nullptr, var_split, B, V, D, h, w, st);
    default: break;
  }
  set_error("rcmvs_costvolume_var_fwd_split: C=%d not in {8,16,32}", C);
  return RCMVS_ERR_UNSUPPORTED;
}

extern "C" int rcmvs_homo_warp_fwd(const float* src_fea, const float* rt, const float* depth, int per_pixel,
                                   float* out, int B, int C, int D, int h, int w, rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(src_fea && rt && depth && out, RCMVS_ERR_BAD_ARG, "rcmvs_homo_warp_fwd: null pointer");
  RCMVS_REQUIRE(B > 0 && C > 0 && D > 0 && h > 0 && w > 0, RCMVS_ERR_BAD_ARG, "rcmvs_homo_warp_fwd: bad dims");
  RCMVS_REQUIRE(h <= 65535 && (int64_t)B * D <= 65535, RCMVS_ERR_UNSUPPORTED, "rcmvs_homo_warp_fwd: h or B*D > 65535");
  dim3 grid(ceil_div(w, 128), h, B * D);
  if (per_pixel) homo_warp_fwd_kernel<true><<<grid, 128, 0, as_stream(stream)>>>(src_fea, rt, depth, out, C, D, h, w);
  else homo_warp_fwd_kernel<false><<<grid, 128, 0, as_stream(stream)>>>(src_fea, rt, depth, out, C, D, h, w);
  RCMVS_CHECK_LAUNCH("rcmvs_homo_warp_fwd");
  return RCMVS_OK;
}

extern "C" int rcmvs_homo_warp_bwd(const float* grad_out, const float* rt, const float* depth, int per_pixel,
                                   float* grad_src, int B, int C, int D, int h, int w, rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(grad_out && rt && depth && grad_src, RCMVS_ERR_BAD_ARG, "rcmvs_homo_warp_bwd: null pointer");
  RCMVS_REQUIRE(B > 0 && C > 0 && D > 0 && h > 0 && w > 0, RCMVS_ERR_BAD_ARG, "rcmvs_homo_warp_bwd: bad dims");
  RCMVS_REQUIRE(h <= 65535 && (int64_t)B * D <= 65535, RCMVS_ERR_UNSUPPORTED, "rcmvs_homo_warp_bwd: h or B*D > 65535");
  dim3 grid(ceil_div(w, 128), h, B * D);
  if (per_pixel) homo_warp_bwd_kernel<true><<<grid, 128, 0, as_stream(stream)>>>(grad_out, rt, depth, grad_src, C, D, h, w);
  else homo_warp_bwd_kernel<false><<<grid, 128, 0, as_stream(stream)>>>(grad_out, rt, depth, grad_src, C, D, h, w);
  RCMVS_CHECK_LAUNCH("rcmvs_homo_warp_bwd");
  return RCMVS_OK;
}
