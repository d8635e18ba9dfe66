// Training-path kernels (BASELINE configs[3]: train_rcmvsnet.py unsupervised step).  The reference's training forward and
// backward are autograd over cuDNN / cuBLAS / ATen kernels: Conv3d/Deconv3d/Conv2d = conv -> BatchNorm (batch statistics,
// running-stat update) -> ReLU (models/modules.py:28-62, 118-210), the variance cost volume (models/casmvsnet.py:65-100),
// softmax + depth regression (:111-112).  Here every one of those ops has a hand-written forward (conv kernels of
// conv3d_direct.cu / conv2d_direct.cu + the batch-norm kernels below) and a hand-written backward:
//
//   bn_stats / bn_finalize / bn_apply           train-mode BatchNorm over channels-last [M][C] (+ ReLU + skip add)
//   bn_bwd_reduce / bn_bwd_apply                its backward (dgamma, dbeta, gradient w.r.t. the convolution output)
//   conv wgrad (2-D and 3-D, stride 1/2, and    G[cs][cb][tap] = Σ_o S[o][cs] · Bg[s·o + tap - pad][cb]: one kernel serves
//     transposed)                               Conv (S = dy, Bg = x) and ConvTranspose (S = x, Bg = dy)
//   conv dgrad                                  3-D: the forward kernels with re-packed weights (stride-1 conv <-> flipped
//                                               conv, stride-2 conv <-> transposed conv); 2-D + the 1-channel `prob`: below
//   costvolume_var_bwd                          gradient of the fused plane sweep + variance w.r.t. all V feature maps
//   softargmin_bwd                              gradient of softmax + depth regression w.r.t. the cost volume
//   sum2x2                                      backward of the FPN's nearest 2x upsample (modules.py:451,455)
//   adam_step                                   fused Adam over the flat parameter / gradient buffers (train_rcmvsnet.py:537)
//
// Precision: fp32 FFMA with fp64 block-level accumulation for the statistics (the reference's autograd is fp32/TF32).
#include "common.cuh"

namespace rcmvs {
namespace train {

// ---------------------------------------------------------------------------------------------------------------------
// BatchNorm (training): x [M][C] channels-last
// ---------------------------------------------------------------------------------------------------------------------
// Block-level reduction of per-thread channel partials (thread = channel quad g = threadIdx.x % cq, C = 4·cq): lanes that hold
// the same quad are summed by shuffles first (cq a power of two <= 32), so that 8 warps, not 256 threads, meet at each
// shared-memory double — a contended shared fp64 atomicAdd is a compare-and-swap loop.
__device__ __forceinline__ void reduce_channel_partials(double (&a)[4], double (&b)[4], double* sm, int C, int cq, int g) {
  const bool by_warp = (cq & (cq - 1)) == 0 && cq <= 32;
  if (by_warp) {
    for (int off = 16; off >= cq; off >>= 1) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        a[k] += __shfl_xor_sync(0xffffffffu, a[k], off);
        b[k] += __shfl_xor_sync(0xffffffffu, b[k], off);
      }
    }
  }
  if (by_warp ? (int)(threadIdx.x & 31) < cq : (int)threadIdx.x < (256 / cq) * cq) {
#pragma unroll
    for (int k = 0; k < 4; ++k) { atomicAdd(&sm[g * 4 + k], a[k]); atomicAdd(&sm[C + g * 4 + k], b[k]); }
  }
}

// partial sums: stats[0..C) += Σ x, stats[C..2C) += Σ x²  (double atomics; caller zero-fills)
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ x, double* __restrict__ stats, int64_t M,
                                                        int C, int rows_per_cta) {
  extern __shared__ double sm[];                 // [2][C]
  for (int i = threadIdx.x; i < 2 * C; i += 256) sm[i] = 0.0;
  __syncthreads();
  const int cq = C / 4;                          // float4 groups per row
  const int g = threadIdx.x % cq, r0 = threadIdx.x / cq, rstep = 256 / cq;
  const int64_t row_begin = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t row_end = min(M, row_begin + rows_per_cta);
  double ds[4] = {0, 0, 0, 0}, dq[4] = {0, 0, 0, 0};
  for (int64_t rb = row_begin + r0; rb < row_end; rb += (int64_t)rstep * 64) {   // fp32 partials over <= 64 rows, then doubles
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
    const int cnt = (int)min((int64_t)64, (row_end - rb + rstep - 1) / rstep);
    const float* xp = x + rb * C + g * 4;
#pragma unroll 8
    for (int j = 0; j < cnt; ++j) {
      const float4 v = ldg4(xp + (int64_t)j * rstep * C);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
    }
    ds[0] += s.x; ds[1] += s.y; ds[2] += s.z; ds[3] += s.w; dq[0] += q.x; dq[1] += q.y; dq[2] += q.z; dq[3] += q.w;
  }
  reduce_channel_partials(ds, dq, sm, C, cq, g);
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += 256) atomicAdd(stats + i, sm[i]);
}

// mean / invstd + running-stat update (momentum m: running = (1-m)·running + m·batch, unbiased variance for running_var)
__global__ void bn_finalize_kernel(const double* __restrict__ stats, float* __restrict__ mean, float* __restrict__ invstd,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, double M, int C,
                                   float eps, float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mu = stats[c] / M;
  double var = stats[C + c] / M - mu * mu;
  if (var < 0) var = 0;
  mean[c] = (float)mu;
  invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mu);
  if (running_var) running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * var * (M > 1 ? M / (M - 1) : 1.0));
}

// z = act((x - mean)·invstd·gamma + beta) (+ skip) (+ nearest-2x-upsampled `up`: 2-D FPN lateral, rows = [N][H][W])
__global__ void __launch_bounds__(256)
bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ invstd,
                const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ skip,
                float* __restrict__ z, int64_t n4, int C, int relu) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n4) return;
  const int c = (int)((i * 4) % C);
  const float4 v = ldg4(x + i * 4), mu = ldg4(mean + c), is = ldg4(invstd + c), ga = ldg4(gamma + c), be = ldg4(beta + c);
  float4 r;
  r.x = fmaf((v.x - mu.x) * is.x, ga.x, be.x); r.y = fmaf((v.y - mu.y) * is.y, ga.y, be.y);
  r.z = fmaf((v.z - mu.z) * is.z, ga.z, be.z); r.w = fmaf((v.w - mu.w) * is.w, ga.w, be.w);
  if (relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
  if (skip) {
    const float4 s = ldg4(skip + i * 4);
    r.x += s.x; r.y += s.y; r.z += s.z; r.w += s.w;
  }
  st4(z + i * 4, r);
}

// backward, pass 1: sums[0..C) += Σ dz', sums[C..2C) += Σ dz'·xhat, dz' = dz·[bn(x) > 0] when relu
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const float* __restrict__ dz, const float* __restrict__ x, const float* __restrict__ mean,
                     const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                     double* __restrict__ sums, int64_t M, int C, int relu, int rows_per_cta) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 2 * C; i += 256) sm[i] = 0.0;
  __syncthreads();
  const int cq = C / 4;
  const int g = threadIdx.x % cq, r0 = threadIdx.x / cq, rstep = 256 / cq;
  const int64_t row_begin = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t row_end = min(M, row_begin + rows_per_cta);
  const float4 mu = ldg4(mean + g * 4), is = ldg4(invstd + g * 4), ga = ldg4(gamma + g * 4), be = ldg4(beta + g * 4);
  const float gg[4] = {ga.x, ga.y, ga.z, ga.w}, bb[4] = {be.x, be.y, be.z, be.w};
  double da[4] = {0, 0, 0, 0}, db[4] = {0, 0, 0, 0};
  for (int64_t rb = row_begin + r0; rb < row_end; rb += (int64_t)rstep * 64) {
    float a[4] = {0, 0, 0, 0}, bq[4] = {0, 0, 0, 0};
    const int cnt = (int)min((int64_t)64, (row_end - rb + rstep - 1) / rstep);
    const int64_t o0 = rb * C + g * 4;
#pragma unroll 4
    for (int j = 0; j < cnt; ++j) {
      const int64_t o = o0 + (int64_t)j * rstep * C;
      const float4 v = ldg4(x + o), d = ldg4(dz + o);
      const float xh[4] = {(v.x - mu.x) * is.x, (v.y - mu.y) * is.y, (v.z - mu.z) * is.z, (v.w - mu.w) * is.w};
      const float dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float dv = (relu && !(fmaf(xh[k], gg[k], bb[k]) > 0.f)) ? 0.f : dd[k];
        a[k] += dv;
        bq[k] = fmaf(dv, xh[k], bq[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { da[k] += a[k]; db[k] += bq[k]; }
  }
  reduce_channel_partials(da, db, sm, C, cq, g);
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += 256) atomicAdd(sums + i, sm[i]);
}

// backward, pass 2: dx = gamma·invstd·(dz' − Σdz'/M − xhat·Σ(dz'·xhat)/M);  dgamma = Σ dz'·xhat, dbeta = Σ dz' (block 0 writes)
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ dz, const float* __restrict__ x, const float* __restrict__ mean,
                    const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                    const double* __restrict__ sums, float* __restrict__ dx, float* __restrict__ dgamma,
                    float* __restrict__ dbeta, int64_t n4, double M, int C, int relu, int frozen) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const bool accumulate = (frozen & 2) != 0;       // flags: bit 0 = frozen statistics, bit 1 = dgamma / dbeta are += targets
  frozen &= 1;
  if (blockIdx.x == 0 && threadIdx.x < C) {
    if (dbeta) dbeta[threadIdx.x] = (accumulate ? dbeta[threadIdx.x] : 0.f) + (float)sums[threadIdx.x];
    if (dgamma) dgamma[threadIdx.x] = (accumulate ? dgamma[threadIdx.x] : 0.f) + (float)sums[C + threadIdx.x];
  }
  if (i >= n4) return;
  const int c = (int)((i * 4) % C);
  const float4 v = ldg4(x + i * 4), d = ldg4(dz + i * 4);
  const float4 mu = ldg4(mean + c), is = ldg4(invstd + c), ga = ldg4(gamma + c), be = ldg4(beta + c);
  const float vv[4] = {v.x, v.y, v.z, v.w}, dd[4] = {d.x, d.y, d.z, d.w}, mm[4] = {mu.x, mu.y, mu.z, mu.w};
  const float ii[4] = {is.x, is.y, is.z, is.w}, gg[4] = {ga.x, ga.y, ga.z, ga.w}, bb[4] = {be.x, be.y, be.z, be.w};
  float r[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float xh = (vv[k] - mm[k]) * ii[k];
    const float dv = (relu && !(fmaf(xh, gg[k], bb[k]) > 0.f)) ? 0.f : dd[k];
    // frozen statistics (eval-mode BatchNorm with autograd): mean / invstd are constants, the two correction terms vanish
    const float m1 = frozen ? 0.f : (float)(sums[c + k] / M), m2 = frozen ? 0.f : (float)(sums[C + c + k] / M);
    r[k] = gg[k] * ii[k] * (dv - m1 - xh * m2);
  }
  st4(dx + i * 4, make_float4(r[0], r[1], r[2], r[3]));
}

// dz' = dz·[z > 0]  (ReLU backward without BatchNorm; not used by the reference nets' conv blocks, kept for completeness)

// ---------------------------------------------------------------------------------------------------------------------
// Weight gradient, 2-D (KD = 1) and 3-D (KD = 3):  G[cs][cb][tap] += Σ_rows Σ_ox S[o][cs] · Bg[s·o + tap − pad][cb]
//   S  [B][Ds][Hs][Ws][Cs]  the tensor on the OUTPUT side of the strided convolution (dy of a Conv, x of a ConvTranspose)
//   Bg [B][Db][Hb][Wb][Cb]  the tensor on its INPUT side (x of a Conv, dy of a ConvTranspose)
// Thread = one (cs, cb) pair holding all KD·K·K taps in registers; it walks output rows with a sliding window of K
// input values per (kd, kh) so a position costs STRIDE·KD·K loads (coalesced over cb, L1-resident: the CS warps of a CTA
// read the same addresses) + 1 broadcast load for KD·K·K FMAs.  Partial sums are added to G with fp32 atomics.
// ---------------------------------------------------------------------------------------------------------------------
// MODE 0: window slides by register moves; 1: rotating window; 2: moves + software prefetch (the S values and the entering
// window columns of position ox + 1 are requested before the FMA block of position ox)
template <int KD, int K, int STRIDE, int MODE, int CSR>
__global__ void __launch_bounds__(256)
wgrad_kernel(const float* __restrict__ S, const float* __restrict__ Bg, float* __restrict__ G, int B, int Ds, int Hs,
             int Ws, int Cs, int Db, int Hb, int Wb, int Cb, int rows_per_cta, int cbb, int csb) {
  constexpr int TAPS = KD * K * K, PAD = K / 2, PADD = KD / 2;
  // 256 threads = `groups` x csb x cbb: few-channel layers (the full-resolution ones) do not fill a CTA with (cs, cb)
  // pairs, so the remaining threads form further groups that take every groups-th row of the CTA's row range.
  // A thread owns CSR consecutive S channels (one 16-byte load when CSR = 4) x one Bg channel: the sliding window is
  // loaded once for CSR·TAPS FMAs.
  const int cb_l = threadIdx.x % cbb, cs_l = (threadIdx.x / cbb) % csb, grp = threadIdx.x / (cbb * csb);
  const int groups = 256 / (cbb * csb);
  const int cs = (blockIdx.y * csb + cs_l) * CSR, cb = blockIdx.z * cbb + cb_l;
  const bool active = cs < Cs && cb < Cb;
  float acc[CSR][TAPS];
#pragma unroll
  for (int c = 0; c < CSR; ++c)
#pragma unroll
    for (int t = 0; t < TAPS; ++t) acc[c][t] = 0.f;
  const int64_t nrows = (int64_t)B * Ds * Hs;
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_cta, r_end = min(nrows, r_begin + rows_per_cta);
  if (active) {
    for (int64_t r = r_begin + grp; r < r_end; r += groups) {
      const int oy = (int)(r % Hs), od = (int)((r / Hs) % Ds), b = (int)(r / ((int64_t)Hs * Ds));
      const float* srow = S + r * Ws * Cs + cs;
      const float* brow[KD * K];
#pragma unroll
      for (int kd = 0; kd < KD; ++kd)
#pragma unroll
        for (int kh = 0; kh < K; ++kh) {
          const int iz = STRIDE * od + kd - PADD, iy = STRIDE * oy + kh - PAD;
          brow[kd * K + kh] = (iz >= 0 && iz < Db && iy >= 0 && iy < Hb)
                                  ? Bg + (((int64_t)b * Db + iz) * Hb + iy) * Wb * Cb + cb : nullptr;
        }
      float win[KD * K][K];
#pragma unroll
      for (int q = 0; q < KD * K; ++q)
#pragma unroll
        for (int kw = 0; kw < K; ++kw) {                              // window of ox = 0: ix = kw - PAD -> slot kw
          const int ix = kw - PAD;
          win[q][kw] = (brow[q] && ix >= 0 && ix < Wb) ? __ldg(brow[q] + (int64_t)ix * Cb) : 0.f;
        }
      auto load_s = [&](int ox, float (&sv)[CSR]) {
        if constexpr (CSR == 4) {
          const float4 t = ldg4(srow + (int64_t)ox * Cs);
          sv[0] = t.x; sv[1] = t.y; sv[2] = t.z; sv[3] = t.w;
        } else {
#pragma unroll
          for (int c = 0; c < CSR; ++c) sv[c] = __ldg(srow + (int64_t)ox * Cs + c);
        }
      };
      if constexpr (MODE == 1) {
        // rotating register window: input column ix lives in slot (ix + PAD) mod K, so sliding by STRIDE only
        // overwrites the STRIDE oldest slots (no register moves); unrolled over one period U of the slot pattern
        constexpr int U = (K % STRIDE == 0) ? K / STRIDE : K;          // STRIDE·U ≡ 0 (mod K)
        for (int ox0 = 0; ox0 < Ws; ox0 += U) {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int ox = ox0 + u;
            if (ox < Ws) {
              float sv[CSR];
              load_s(ox, sv);
#pragma unroll
              for (int q = 0; q < KD * K; ++q)
#pragma unroll
                for (int kw = 0; kw < K; ++kw) {
                  const float wv = win[q][(STRIDE * u + kw) % K];
#pragma unroll
                  for (int c = 0; c < CSR; ++c) acc[c][q * K + kw] = fmaf(sv[c], wv, acc[c][q * K + kw]);
                }
#pragma unroll
              for (int q = 0; q < KD * K; ++q)
#pragma unroll
                for (int j = 0; j < STRIDE && j < K; ++j) {
                  const int ix = STRIDE * ox - PAD + K + j;
                  win[q][(STRIDE * u + j) % K] = (brow[q] && ix < Wb) ? __ldg(brow[q] + (int64_t)ix * Cb) : 0.f;
                }
            }
          }
        }
      } else if constexpr (MODE == 2) {
        constexpr int NEWC = STRIDE < K ? STRIDE : K;                  // columns entering the window per step
        float sv[CSR];
        load_s(0, sv);
        for (int ox = 0; ox < Ws; ++ox) {
          float svn[CSR], wn[KD * K][NEWC];
          const int oxn = min(ox + 1, Ws - 1);                          // (the last prefetch re-reads valid data, unused)
          load_s(oxn, svn);
#pragma unroll
          for (int q = 0; q < KD * K; ++q)
#pragma unroll
            for (int j = 0; j < NEWC; ++j) {
              const int ix = STRIDE * (ox + 1) + (K - NEWC + j) - PAD;
              wn[q][j] = (brow[q] && ix >= 0 && ix < Wb) ? __ldg(brow[q] + (int64_t)ix * Cb) : 0.f;
            }
#pragma unroll
          for (int q = 0; q < KD * K; ++q)
#pragma unroll
            for (int kw = 0; kw < K; ++kw) {
              const float wv = win[q][kw];
#pragma unroll
              for (int c = 0; c < CSR; ++c) acc[c][q * K + kw] = fmaf(sv[c], wv, acc[c][q * K + kw]);
            }
#pragma unroll
          for (int q = 0; q < KD * K; ++q) {
#pragma unroll
            for (int kw = 0; kw + STRIDE < K; ++kw) win[q][kw] = win[q][kw + STRIDE];
#pragma unroll
            for (int j = 0; j < NEWC; ++j) win[q][K - NEWC + j] = wn[q][j];
          }
#pragma unroll
          for (int c = 0; c < CSR; ++c) sv[c] = svn[c];
        }
      } else {
        for (int ox = 0; ox < Ws; ++ox) {
          float sv[CSR];
          load_s(ox, sv);
#pragma unroll
          for (int q = 0; q < KD * K; ++q)
#pragma unroll
            for (int kw = 0; kw < K; ++kw) {
              const float wv = win[q][kw];
#pragma unroll
              for (int c = 0; c < CSR; ++c) acc[c][q * K + kw] = fmaf(sv[c], wv, acc[c][q * K + kw]);
            }
#pragma unroll
          for (int q = 0; q < KD * K; ++q) {                           // slide by STRIDE (register moves)
#pragma unroll
            for (int kw = 0; kw + STRIDE < K; ++kw) win[q][kw] = win[q][kw + STRIDE];
#pragma unroll
            for (int kw = (K > STRIDE ? K - STRIDE : 0); kw < K; ++kw) {
              const int ix = STRIDE * (ox + 1) + kw - PAD;
              win[q][kw] = (brow[q] && ix >= 0 && ix < Wb) ? __ldg(brow[q] + (int64_t)ix * Cb) : 0.f;
            }
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < CSR; ++c)
#pragma unroll
      for (int t = 0; t < TAPS; ++t) atomicAdd(G + ((int64_t)(cs + c) * Cb + cb) * TAPS + t, acc[c][t]);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Weight gradient, shared-memory variant (the default where it applies: Cs % 4 == 0, Cb % 4 == 0).  The register-window
// kernel above streams its operands from L1 / L2 with one or two loads in flight per warp: ncu showed it latency-bound
// (FMA pipe 12-24 %, 2-4 warps stalled on long-scoreboard per issue, L1 hit rate 13-43 %).  Here a CTA walks a range of
// output rows; for each row the S row [Ws][CSB] and the KD·K input rows [Wb + 2·PAD][CBB] (zero halo columns, zero rows
// outside the volume: no bounds checks in the inner loop) are staged in shared memory with coalesced 16-byte loads, and
// all 256 threads work on that row: thread = (4 S channels, 1 Bg channel, x slice), x positions ox = slice, slice + NSL, …
// so that the lanes of a warp (Bg channel fastest, then slice) read consecutive shared-memory words.  Every position costs
// one LDS.128 + KD·K·K LDS.32 for 4·KD·K·K FMAs.  Partial sums are combined through shared-memory atomics at the end
// of the row range, then added to G with one global atomic per weight and CTA.
// ---------------------------------------------------------------------------------------------------------------------
// 16-byte asynchronous global -> shared copy; `valid` false writes zeros (src-size 0) and reads nothing
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}

// CSR = S channels per thread: 4, or 1 when Cs is not a multiple of 4 (`prob`: Cs = 1).
// RING: the Bg rows of one (batch, od) run live in K + STRIDE slots per kd plane, slot = (STRIDE·oy + kh) mod (K + STRIDE):
// going from output row oy to oy + 1 only STRIDE new rows per plane are fetched (3 instead of 9 for the 3x3x3 stride-1
// layers), into slots the current row does not read, WHILE the current row is being accumulated (the S row is double
// buffered) — staging costs a third and hides behind the FMAs.  A new (batch, od) run restarts with a full, synchronous fill.
template <int KD, int K, int STRIDE, int CSR, bool RING>
__global__ void __launch_bounds__(256, (CSR * KD * K * K > 36 ? (RING ? 1 : 2) : 3))   // 25/27 taps x 4 channels: 100-108 accumulators
wgrad_smem_kernel(const float* __restrict__ S, const float* __restrict__ Bg, float* __restrict__ G, int B, int Ds, int Hs,
                  int Ws, int Cs, int Db, int Hb, int Wb, int Cb, int rows_per_cta, int CSB, int CBB) {
  constexpr int TAPS = KD * K * K, PAD = K / 2, PADD = KD / 2;
  constexpr int NSLOT = RING ? K + STRIDE : K, NSB = RING ? 2 : 1;
  extern __shared__ __align__(16) unsigned char wg_smem_raw[];
  float* sm = reinterpret_cast<float*>(wg_smem_raw);
  const int Wbp = Wb + 2 * PAD;
  float* s_rows = sm;                                 // [NSB][Ws][CSB]
  float* b_rows = sm + (size_t)NSB * Ws * CSB;        // [KD][NSLOT][Wbp][CBB]
  const int csg_n = CSB / CSR;                        // S channel groups in this CTA
  const int NSL = 256 / (csg_n * CBB);                // x slices
  const int cb_l = threadIdx.x % CBB, slice = (threadIdx.x / CBB) % NSL, cs_g = threadIdx.x / (CBB * NSL);
  const bool active = cs_g < csg_n;
  const int cs0 = blockIdx.y * CSB, cb0 = blockIdx.z * CBB;
  float acc[CSR][TAPS];
#pragma unroll
  for (int c = 0; c < CSR; ++c)
#pragma unroll
    for (int t = 0; t < TAPS; ++t) acc[c][t] = 0.f;
  const int64_t nrows = (int64_t)B * Ds * Hs;
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_cta, r_end = min(nrows, r_begin + rows_per_cta);
  const int cs4 = CSB / 4, cb4 = CBB / 4;
  const int row_elems = Wbp * CBB, plane_elems = NSLOT * row_elems;

  // global -> shared with 16-byte cp.async (no register staging, every thread keeps all its copies in flight; src-size 0
  // zero-fills the halo).  `full`: all K rows of every plane; otherwise only the STRIDE rows that are new against row r - 1.
  auto stage = [&](int64_t r, int buf, bool full) {
    const int oy = (int)(r % Hs), od = (int)((r / Hs) % Ds), b = (int)(r / ((int64_t)Hs * Ds));
    float* s_row = s_rows + (size_t)(RING ? buf : 0) * Ws * CSB;
    const float* src = S + r * Ws * Cs + cs0;
    if constexpr (CSR == 4) {
      for (int e = threadIdx.x; e < Ws * cs4; e += 256) {
        const int ox = e / cs4, c4 = e % cs4;
        cp_async16(s_row + ox * CSB + 4 * c4, src + (int64_t)ox * Cs + 4 * c4, true);
      }
    } else {
      for (int e = threadIdx.x; e < Ws * CSB; e += 256) s_row[e] = __ldg(src + (int64_t)(e / CSB) * Cs + e % CSB);
    }
#pragma unroll
    for (int q = 0; q < KD * K; ++q) {
      const int kdi = q / K, kh = q % K;
      if (!full && kh < K - STRIDE) continue;
      const int iz = STRIDE * od + kdi - PADD, iy = STRIDE * oy + kh - PAD;
      const int slot = RING ? (STRIDE * oy + kh) % NSLOT : kh;
      const bool ok = iz >= 0 && iz < Db && iy >= 0 && iy < Hb;
      const float* brow = ok ? Bg + (((int64_t)b * Db + iz) * Hb + iy) * Wb * Cb + cb0 : Bg;
      float* dst = b_rows + (size_t)kdi * plane_elems + (size_t)slot * row_elems;
      for (int e = threadIdx.x; e < Wbp * cb4; e += 256) {
        const int ixp = e / cb4, c4 = e % cb4, ix = ixp - PAD;
        const bool in = ok && ix >= 0 && ix < Wb;
        cp_async16(dst + ixp * CBB + 4 * c4, in ? brow + (int64_t)ix * Cb + 4 * c4 : Bg, in);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int buf = 0;
  if (r_begin < r_end) {
    stage(r_begin, 0, true);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  for (int64_t r = r_begin; r < r_end; ++r) {
    const bool next = r + 1 < r_end;
    const bool same = RING && next && (r + 1) % Hs != 0;       // the next row continues this (batch, od) run
    if (same) stage(r + 1, buf ^ 1, false);                    // lands in the other S buffer and in slots row r does not read
    if (active) {
      const int sbase = STRIDE * (int)(r % Hs);
      int roff[KD * K];                                        // (plane, slot) offsets of this row's KD·K Bg rows
#pragma unroll
      for (int q = 0; q < KD * K; ++q) roff[q] = (q / K) * plane_elems + (RING ? (sbase + q % K) % NSLOT : q % K) * row_elems;
      const float* s_row = s_rows + (size_t)(RING ? buf : 0) * Ws * CSB;
      for (int ox = slice; ox < Ws; ox += NSL) {
        float sv[CSR];
        if constexpr (CSR == 4) {
          const float4 sv4 = *reinterpret_cast<const float4*>(s_row + ox * CSB + 4 * cs_g);
          sv[0] = sv4.x; sv[1] = sv4.y; sv[2] = sv4.z; sv[3] = sv4.w;
        } else {
          sv[0] = s_row[ox * CSB + cs_g];
        }
        const float* bp = b_rows + (size_t)(ox * STRIDE) * CBB + cb_l;
#pragma unroll
        for (int kw = 0; kw < K; ++kw) {
          const float* bpk = bp + kw * CBB;
#pragma unroll
          for (int q = 0; q < KD * K; ++q) {
            const float wv = bpk[roff[q]];
#pragma unroll
            for (int c = 0; c < CSR; ++c) acc[c][q * K + kw] = fmaf(sv[c], wv, acc[c][q * K + kw]);
          }
        }
      }
    }
    if (same) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();                                         // row r consumed by all, row r + 1 complete for all
    } else if (next) {
      __syncthreads();
      stage(r + 1, buf ^ 1, true);
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
    }
    buf ^= 1;
  }
  // ---- combine the x slices: shared-memory atomics on red[cs][cb][tap], then one global atomic per weight ------------
  __syncthreads();
  float* red = sm;
  const int nred = CSB * CBB * TAPS;
  for (int e = threadIdx.x; e < nred; e += 256) red[e] = 0.f;
  __syncthreads();
  // lanes of a warp = CBB channels x 32 / CBB slices of ONE S channel group (CBB, csg_n, NSL are powers of two on this
  // path): the slices meet by shuffles, so 32 / CBB times fewer threads contend for each shared-memory sum
  const bool by_warp = CBB < 32 && (CBB & (CBB - 1)) == 0 && NSL * CBB >= 32 && (NSL & (NSL - 1)) == 0;
  if (by_warp) {
    for (int off = 16; off >= CBB; off >>= 1) {
#pragma unroll
      for (int c = 0; c < CSR; ++c)
#pragma unroll
        for (int t = 0; t < TAPS; ++t) acc[c][t] += __shfl_xor_sync(0xffffffffu, acc[c][t], off);
    }
  }
  if (active && (!by_warp || (int)(threadIdx.x & 31) < CBB)) {
#pragma unroll
    for (int c = 0; c < CSR; ++c)
#pragma unroll
      for (int t = 0; t < TAPS; ++t) atomicAdd(red + ((CSR * cs_g + c) * CBB + cb_l) * TAPS + t, acc[c][t]);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < nred; e += 256) {
    const int t = e % TAPS, cb = (e / TAPS) % CBB, cs = e / (TAPS * CBB);
    if (cs0 + cs < Cs && cb0 + cb < Cb) atomicAdd(G + ((int64_t)(cs0 + cs) * Cb + cb0 + cb) * TAPS + t, red[e]);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// 2-D data gradient (FeatureNet): dx[n,y,x,ci] = Σ_{kh,kw,co} dy[n,(y+P−kh)/s,(x+P−kw)/s,co]·w[co][ci][kh][kw]
// (terms with non-integer or out-of-range output coordinates vanish).  Thread = one input pixel x 4 input channels;
// the re-laid-out weights [tap][co][ci] sit in shared memory.  Optional `up2_add`: dx += the gradient arriving through a
// second consumer (FPN: a map feeds both a 3x3 conv and the next lateral) is handled by the caller (autograd sums).
// ---------------------------------------------------------------------------------------------------------------------
template <int K, int STRIDE>
__global__ void __launch_bounds__(256)
conv2d_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ wd /* [K*K][Cout][Cin] */,
                    float* __restrict__ dx, int N, int H, int W, int Cin, int Ho, int Wo, int Cout) {
  extern __shared__ __align__(16) float wsm[];
  const int nw = K * K * Cout * Cin;
  for (int i = threadIdx.x; i < nw; i += 256) wsm[i] = __ldg(wd + i);
  __syncthreads();
  constexpr int P = K / 2;
  const int cq = Cin / 4;
  const int64_t gid = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int g = (int)(gid % cq);
  const int64_t pix = gid / cq;
  if (pix >= (int64_t)N * H * W) return;
  const int x = (int)(pix % W), y = (int)((pix / W) % H), n = (int)(pix / ((int64_t)W * H));
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int kh = 0; kh < K; ++kh) {
    const int ty = y + P - kh;
    if (ty < 0 || (ty % STRIDE) != 0) continue;
    const int oy = ty / STRIDE;
    if (oy >= Ho) continue;
#pragma unroll
    for (int kw = 0; kw < K; ++kw) {
      const int tx = x + P - kw;
      if (tx < 0 || (tx % STRIDE) != 0) continue;
      const int ox = tx / STRIDE;
      if (ox >= Wo) continue;
      const float* dp = dy + (((int64_t)n * Ho + oy) * Wo + ox) * Cout;
      const float* wp = wsm + (kh * K + kw) * Cout * Cin + g * 4;
      for (int co = 0; co < Cout; co += 4) {
        const float4 d = ldg4(dp + co);
        const float4 w0 = *reinterpret_cast<const float4*>(wp + (co + 0) * Cin);
        const float4 w1 = *reinterpret_cast<const float4*>(wp + (co + 1) * Cin);
        const float4 w2 = *reinterpret_cast<const float4*>(wp + (co + 2) * Cin);
        const float4 w3 = *reinterpret_cast<const float4*>(wp + (co + 3) * Cin);
        acc.x = fmaf(d.x, w0.x, fmaf(d.y, w1.x, fmaf(d.z, w2.x, fmaf(d.w, w3.x, acc.x))));
        acc.y = fmaf(d.x, w0.y, fmaf(d.y, w1.y, fmaf(d.z, w2.y, fmaf(d.w, w3.y, acc.y))));
        acc.z = fmaf(d.x, w0.z, fmaf(d.y, w1.z, fmaf(d.z, w2.z, fmaf(d.w, w3.z, acc.z))));
        acc.w = fmaf(d.x, w0.w, fmaf(d.y, w1.w, fmaf(d.z, w2.w, fmaf(d.w, w3.w, acc.w))));
      }
    }
  }
  st4(dx + pix * Cin + g * 4, acc);
}

// Stride-1 variant, register-blocked: thread = 4 consecutive input pixels x 4 input channels.  Per (kh, 4 output channels)
// it loads the K + 3 dy columns its four pixels touch once (float4 each) and the K x 4 weight rows once, for K·4·16 FMAs —
// 10.7 FMAs per load at K = 3 against 3.2 for one pixel per thread — and a CTA walks over many tiles, so the weights are
// staged in shared memory once per CTA instead of once per 32-64 pixels.
template <int K>
__global__ void __launch_bounds__(256)
conv2d_dgrad_s1_kernel(const float* __restrict__ dy, const float* __restrict__ wd /* [K*K][Cout][Cin] */,
                       float* __restrict__ dx, int N, int H, int W, int Cin, int Cout) {
  extern __shared__ __align__(16) float wsm[];
  const int nw = K * K * Cout * Cin;
  for (int i = threadIdx.x; i < nw / 4; i += 256) reinterpret_cast<float4*>(wsm)[i] = ldg4(wd + 4 * i);
  __syncthreads();
  constexpr int P = K / 2, NC = K + 3;
  const int cq = Cin / 4, Wt = (W + 3) / 4;
  const int64_t total = (int64_t)N * H * Wt * cq;
  for (int64_t gid = (int64_t)blockIdx.x * 256 + threadIdx.x; gid < total; gid += (int64_t)gridDim.x * 256) {
    const int g = (int)(gid % cq);
    const int64_t tile = gid / cq;
    const int x0 = (int)(tile % Wt) * 4, y = (int)((tile / Wt) % H), n = (int)(tile / ((int64_t)Wt * H));
    float4 acc[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) acc[p] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int kh = 0; kh < K; ++kh) {
      const int oy = y + P - kh;
      if (oy < 0 || oy >= H) continue;
      const float* drow = dy + ((int64_t)n * H + oy) * W * Cout;
      for (int co = 0; co < Cout; co += 4) {
        float4 d[NC];                                  // dy columns x0 - P .. x0 + 3 + P
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const int ox = x0 - P + c;
          d[c] = (ox >= 0 && ox < W) ? ldg4(drow + (int64_t)ox * Cout + co) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int kw = 0; kw < K; ++kw) {
          const float* wp = wsm + ((kh * K + kw) * Cout + co) * Cin + g * 4;
          const float4 w0 = *reinterpret_cast<const float4*>(wp);
          const float4 w1 = *reinterpret_cast<const float4*>(wp + Cin);
          const float4 w2 = *reinterpret_cast<const float4*>(wp + 2 * Cin);
          const float4 w3 = *reinterpret_cast<const float4*>(wp + 3 * Cin);
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const float4 dv = d[p + 2 * P - kw];       // pixel x0 + p reads dy column x0 + p + P - kw
            acc[p].x = fmaf(dv.x, w0.x, fmaf(dv.y, w1.x, fmaf(dv.z, w2.x, fmaf(dv.w, w3.x, acc[p].x))));
            acc[p].y = fmaf(dv.x, w0.y, fmaf(dv.y, w1.y, fmaf(dv.z, w2.y, fmaf(dv.w, w3.y, acc[p].y))));
            acc[p].z = fmaf(dv.x, w0.z, fmaf(dv.y, w1.z, fmaf(dv.z, w2.z, fmaf(dv.w, w3.z, acc[p].z))));
            acc[p].w = fmaf(dv.x, w0.w, fmaf(dv.y, w1.w, fmaf(dv.z, w2.w, fmaf(dv.w, w3.w, acc[p].w))));
          }
        }
      }
    }
    float* out = dx + (((int64_t)n * H + y) * W + x0) * Cin + g * 4;
#pragma unroll
    for (int p = 0; p < 4; ++p)
      if (x0 + p < W) st4(out + (int64_t)p * Cin, acc[p]);
  }
}

// torch Conv2d weight [Cout][Cin][K][K] -> [K*K][Cout][Cin_pad] (input channels contiguous, zero-padded)
__global__ void pack_dgrad2d_kernel(const float* __restrict__ w, float* __restrict__ wd, int Cout, int Cin, int Cin_pad,
                                    int KK) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= KK * Cout * Cin_pad) return;
  const int ci = i % Cin_pad, co = (i / Cin_pad) % Cout, tap = i / (Cin_pad * Cout);
  wd[i] = ci < Cin ? w[((int64_t)co * Cin + ci) * KK + tap] : 0.f;
}

// ---------------------------------------------------------------------------------------------------------------------
// data gradient of `prob` (Conv3d 8 -> 1, k3 p1, no BN): dy8[v][ci] = Σ_k dcost[v − k + 1]·w[0][ci][k]
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
conv3d_c1_dgrad_kernel(const float* __restrict__ dcost, const float* __restrict__ w /* [1][8][27] */,
                       float* __restrict__ dx /* [B][D][H][W][8] */, int B, int D, int H, int W) {
  __shared__ float ws[27 * 8];
  for (int i = threadIdx.x; i < 27 * 8; i += 256) {
    const int ci = i % 8, k = i / 8;
    ws[i] = __ldg(w + ci * 27 + k);
  }
  __syncthreads();
  const int64_t v = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (v >= (int64_t)B * D * H * W) return;
  const int x = (int)(v % W), y = (int)((v / W) % H), z = (int)((v / ((int64_t)W * H)) % D);
  const int64_t b = v / ((int64_t)W * H * D);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int kd = 0; kd < 3; ++kd) {
    const int oz = z - kd + 1;
    if (oz < 0 || oz >= D) continue;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int oy = y - kh + 1;
      if (oy < 0 || oy >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int ox = x - kw + 1;
        if (ox < 0 || ox >= W) continue;
        const float g = __ldg(dcost + ((b * D + oz) * H + oy) * W + ox);
        const float* wk = ws + ((kd * 3 + kh) * 3 + kw) * 8;
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = fmaf(g, wk[c], acc[c]);
      }
    }
  }
  st4(dx + v * 8, make_float4(acc[0], acc[1], acc[2], acc[3]));
  st4(dx + v * 8 + 4, make_float4(acc[4], acc[5], acc[6], acc[7]));
}

// dgrad weights for the forward kernels' packed format [27][Cin'][Cout'] of a STRIDE-1 conv: Cin' = Cout, Cout' = Cin,
// packed[t][co][ci] = w[co][ci][26 − t]  (flipped taps; the stride-2 / transposed cases reuse rcmvs_pack_conv3d_weight)
__global__ void pack_dgrad3d_s1_kernel(const float* __restrict__ w, float* __restrict__ packed, int Cout, int Cin) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 27 * Cout * Cin) return;
  const int ci = i % Cin, co = (i / Cin) % Cout, t = i / (Cin * Cout);
  packed[i] = w[((int64_t)co * Cin + ci) * 27 + (26 - t)];
}

// ---------------------------------------------------------------------------------------------------------------------
// Backward of the fused plane sweep + variance (models/casmvsnet.py:65-100):
//   var = Σ_v f_v²/V − (Σ_v f_v/V)²  ->  d var / d f_v = (2/V)(f_v − mean);   var_noref likewise over the sources only.
// Thread = one reference pixel x 4 channels: per plane it re-samples the V−1 sources (keeping them in registers), forms the
// means, accumulates the reference view's gradient in registers and scatters the sources' gradients to their 4 taps with
// vector reductions (red.global.add.v4.f32).  The sample grid carries no gradient (modules.py:314).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kMaxSrcB = 7;
__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int C, bool PER_PIXEL>
__global__ void __launch_bounds__(256)
costvolume_var_bwd_kernel(const float* __restrict__ feats, const float* __restrict__ rt_g, const float* __restrict__ depth,
                          const float* __restrict__ gvar, const float* __restrict__ gnoref, float* __restrict__ gfeats,
                          int V, int D, int h, int w) {
  constexpr int CQ = C / 4, PXB = 256 / CQ;
  const int q = threadIdx.x % CQ;
  const int x = blockIdx.x * PXB + threadIdx.x / CQ, y = blockIdx.y, b = blockIdx.z;
  __shared__ float rt_s[kMaxSrcB * 12];
  for (int i = threadIdx.x; i < (V - 1) * 12; i += 256) rt_s[i] = __ldg(rt_g + (int64_t)b * (V - 1) * 12 + i);
  __syncthreads();
  if (x >= w) return;
  const int64_t hw = (int64_t)h * w;
  feats += (int64_t)b * V * hw * C + q * 4;
  gfeats += (int64_t)b * V * hw * C + q * 4;
  gvar += (int64_t)b * D * hw * C + q * 4;
  if (gnoref) gnoref += (int64_t)b * D * hw * C + q * 4;
  depth += PER_PIXEL ? (int64_t)b * D * hw + (int64_t)y * w + x : (int64_t)b * D;
  const float4 ref = ldg4(feats + ((int64_t)y * w + x) * C);
  const float inv_v = 1.f / (float)V, two_v = 2.f / (float)V;
  const float xf = (float)x, yf = (float)y;
  float4 gref = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int d = 0; d < D; ++d) {
    const float dep = PER_PIXEL ? __ldg(depth + (int64_t)d * hw) : __ldg(depth + d);
    const int64_t o = (((int64_t)d * h + y) * w + x) * C;
    const float4 g = ldg4(gvar + o);
    const float4 gn = gnoref ? ldg4(gnoref + o) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 f[kMaxSrcB];
    int off[kMaxSrcB][4];
    float wt[kMaxSrcB][4];
    float4 s = ref, n = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int v = 0; v < kMaxSrcB; ++v) {
      if (v >= V - 1) break;
      const float* m = rt_s + v * 12;
      const float rx = fmaf(m[0], xf, fmaf(m[1], yf, m[2]));
      const float ry = fmaf(m[4], xf, fmaf(m[5], yf, m[6]));
      const float rz = fmaf(m[8], xf, fmaf(m[9], yf, m[10]));
      const float Z = fmaf(rz, dep, m[11]);
      const float px = __fdiv_rn(fmaf(rx, dep, m[3]), Z), py = __fdiv_rn(fmaf(ry, dep, m[7]), Z);
      const bool finite = (fabsf(px) < 1e8f) && (fabsf(py) < 1e8f);
      const float fx = floorf(px), fy = floorf(py);
      const float ax = px - fx, ay = py - fy;
      const int x0 = finite ? (int)fx : -4, y0 = finite ? (int)fy : -4;
      const bool vx0 = (x0 >= 0) && (x0 < w), vx1 = (x0 + 1 >= 0) && (x0 + 1 < w);
      const bool vy0 = (y0 >= 0) && (y0 < h), vy1 = (y0 + 1 >= 0) && (y0 + 1 < h);
      const int cx0 = min(max(x0, 0), w - 1), cx1 = min(max(x0 + 1, 0), w - 1);
      const int cy0 = min(max(y0, 0), h - 1), cy1 = min(max(y0 + 1, 0), h - 1);
      wt[v][0] = (vx0 && vy0) ? (1.f - ax) * (1.f - ay) : 0.f;
      wt[v][1] = (vx1 && vy0) ? ax * (1.f - ay) : 0.f;
      wt[v][2] = (vx0 && vy1) ? (1.f - ax) * ay : 0.f;
      wt[v][3] = (vx1 && vy1) ? ax * ay : 0.f;
      off[v][0] = (cy0 * w + cx0) * C; off[v][1] = (cy0 * w + cx1) * C;
      off[v][2] = (cy1 * w + cx0) * C; off[v][3] = (cy1 * w + cx1) * C;
      const float* src = feats + (int64_t)(v + 1) * hw * C;
      const float4 a = ldg4(src + off[v][0]), bb = ldg4(src + off[v][1]), c = ldg4(src + off[v][2]), e = ldg4(src + off[v][3]);
      f[v].x = fmaf(e.x, wt[v][3], fmaf(c.x, wt[v][2], fmaf(bb.x, wt[v][1], a.x * wt[v][0])));
      f[v].y = fmaf(e.y, wt[v][3], fmaf(c.y, wt[v][2], fmaf(bb.y, wt[v][1], a.y * wt[v][0])));
      f[v].z = fmaf(e.z, wt[v][3], fmaf(c.z, wt[v][2], fmaf(bb.z, wt[v][1], a.z * wt[v][0])));
      f[v].w = fmaf(e.w, wt[v][3], fmaf(c.w, wt[v][2], fmaf(bb.w, wt[v][1], a.w * wt[v][0])));
      s.x += f[v].x; s.y += f[v].y; s.z += f[v].z; s.w += f[v].w;
      n.x += f[v].x; n.y += f[v].y; n.z += f[v].z; n.w += f[v].w;
    }
    const float4 mean = make_float4(s.x * inv_v, s.y * inv_v, s.z * inv_v, s.w * inv_v);
    const float4 nmean = make_float4(n.x * inv_v, n.y * inv_v, n.z * inv_v, n.w * inv_v);
    gref.x = fmaf(g.x * two_v, ref.x - mean.x, gref.x); gref.y = fmaf(g.y * two_v, ref.y - mean.y, gref.y);
    gref.z = fmaf(g.z * two_v, ref.z - mean.z, gref.z); gref.w = fmaf(g.w * two_v, ref.w - mean.w, gref.w);
#pragma unroll
    for (int v = 0; v < kMaxSrcB; ++v) {
      if (v >= V - 1) break;
      float4 gf;
      gf.x = two_v * (g.x * (f[v].x - mean.x) + gn.x * (f[v].x - nmean.x));
      gf.y = two_v * (g.y * (f[v].y - mean.y) + gn.y * (f[v].y - nmean.y));
      gf.z = two_v * (g.z * (f[v].z - mean.z) + gn.z * (f[v].z - nmean.z));
      gf.w = two_v * (g.w * (f[v].w - mean.w) + gn.w * (f[v].w - nmean.w));
      float* dst = gfeats + (int64_t)(v + 1) * hw * C;
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (wt[v][t] != 0.f)
          red_add_v4(dst + off[v][t], make_float4(gf.x * wt[v][t], gf.y * wt[v][t], gf.z * wt[v][t], gf.w * wt[v][t]));
    }
  }
  red_add_v4(gfeats + ((int64_t)y * w + x) * C, gref);
}

// ---------------------------------------------------------------------------------------------------------------------
// Backward of softmax over depth + depth regression (casmvsnet.py:111-112): dcost[d] = g·p_d·(hyp_d − depth)
// ---------------------------------------------------------------------------------------------------------------------
template <bool PER_PIXEL>
__global__ void __launch_bounds__(128)
softargmin_bwd_kernel(const float* __restrict__ cost, const float* __restrict__ planes, const float* __restrict__ gdepth,
                      float* __restrict__ gcost, int D, int64_t hw) {
  const int64_t p = (int64_t)blockIdx.x * 128 + threadIdx.x;
  const int b = blockIdx.y;
  if (p >= hw) return;
  cost += (int64_t)b * D * hw + p;
  gcost += (int64_t)b * D * hw + p;
  planes += PER_PIXEL ? (int64_t)b * D * hw + p : (int64_t)b * D;
  const float g = __ldg(gdepth + (int64_t)b * hw + p);
  float m = -INFINITY;
  for (int d = 0; d < D; ++d) m = fmaxf(m, __ldg(cost + (int64_t)d * hw));
  float sum = 0.f;
  for (int d = 0; d < D; ++d) sum += expf(__ldg(cost + (int64_t)d * hw) - m);
  float dsum = 0.f;
  for (int d = 0; d < D; ++d) {
    const float pd = __fdiv_rn(expf(__ldg(cost + (int64_t)d * hw) - m), sum);
    dsum = fmaf(pd, PER_PIXEL ? __ldg(planes + (int64_t)d * hw) : __ldg(planes + d), dsum);
  }
  for (int d = 0; d < D; ++d) {
    const float pd = __fdiv_rn(expf(__ldg(cost + (int64_t)d * hw) - m), sum);
    const float dv = PER_PIXEL ? __ldg(planes + (int64_t)d * hw) : __ldg(planes + d);
    gcost[(int64_t)d * hw] = g * pd * (dv - dsum);
  }
}

// out[n][y][x][c] = Σ_{dy,dx in {0,1}} in[n][2y+dy][2x+dx][c]   (backward of the nearest-neighbour 2x upsample)
__global__ void sum2x2_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int Ho, int Wo, int C) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int cq = C / 4;
  if (i >= (int64_t)N * Ho * Wo * cq) return;
  const int g = (int)(i % cq);
  const int64_t pix = i / cq;
  const int x = (int)(pix % Wo), y = (int)((pix / Wo) % Ho);
  const int64_t n = pix / ((int64_t)Wo * Ho);
  const int W2 = 2 * Wo;
  const float* p = in + (((n * 2 * Ho + 2 * y) * W2) + 2 * x) * C + g * 4;
  const float4 a = ldg4(p), b = ldg4(p + C), c = ldg4(p + (int64_t)W2 * C), d = ldg4(p + (int64_t)W2 * C + C);
  st4(out + pix * C + g * 4, make_float4(a.x + b.x + c.x + d.x, a.y + b.y + c.y + d.y, a.z + b.z + c.z + d.z, a.w + b.w + c.w + d.w));
}

// out[n][y][x][c] = a[n][y][x][c] + up[n][y/2][x/2][c]     (FPN lateral: nearest 2x upsample + add, modules.py:451,455)
__global__ void up2_add_kernel(const float* __restrict__ a, const float* __restrict__ up, float* __restrict__ out, int N,
                               int H, int W, int C) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int cq = C / 4;
  if (i >= (int64_t)N * H * W * cq) return;
  const int g = (int)(i % cq);
  const int64_t pix = i / cq;
  const int x = (int)(pix % W), y = (int)((pix / W) % H);
  const int64_t n = pix / ((int64_t)W * H);
  const float4 u = ldg4(up + ((n * (H / 2) + y / 2) * (W / 2) + x / 2) * C + g * 4), v = ldg4(a + pix * C + g * 4);
  st4(out + pix * C + g * 4, make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w));
}

// fused Adam (torch.optim.Adam semantics, no amsgrad): one launch over the flat parameter / gradient / state buffers
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay, float bc1,
                            float bc2_sqrt, float grad_scale) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float gi = g[i] * grad_scale;
  if (weight_decay != 0.f) gi = fmaf(weight_decay, p[i], gi);
  const float mi = fmaf(beta1, m[i], (1.f - beta1) * gi);
  const float vi = fmaf(beta2, v[i], (1.f - beta2) * gi * gi);
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] -= (lr / bc1) * (mi / denom);
}

static int rows_per_cta(int64_t M, int target_ctas) {
  int64_t r = ceil_div<int64_t>(M, target_ctas);
  if (r < 64) r = 64;
  return (int)r;
}

}  // namespace train
}  // namespace rcmvs

using namespace rcmvs;
using namespace rcmvs::train;

extern "C" int rcmvs_bn_stats(const float* x, double* stats, int64_t M, int C, rcmvs_stream_t stream) {
  RCMVS_REQUIRE(x && stats && M > 0 && C >= 4 && C % 4 == 0 && C <= 256, RCMVS_ERR_BAD_ARG, "rcmvs_bn_stats: bad arguments");
  const int rpc = rows_per_cta(M, 148 * 8);
  bn_stats_kernel<<<(unsigned)ceil_div<int64_t>(M, rpc), 256, 2 * C * sizeof(double), as_stream(stream)>>>(x, stats, M, C, rpc);
  RCMVS_CHECK_LAUNCH("rcmvs_bn_stats");
  return RCMVS_OK;
}

extern "C" int rcmvs_bn_finalize(const double* stats, float* mean, float* invstd, float* running_mean, float* running_var,
                                 int64_t M, int C, float eps, float momentum, rcmvs_stream_t stream) {
  RCMVS_REQUIRE(stats && mean && invstd && M > 0 && C > 0, RCMVS_ERR_BAD_ARG, "rcmvs_bn_finalize: bad arguments");
  bn_finalize_kernel<<<ceil_div(C, 64), 64, 0, as_stream(stream)>>>(stats, mean, invstd, running_mean, running_var,
                                                                     (double)M, C, eps, momentum);
  RCMVS_CHECK_LAUNCH("rcmvs_bn_finalize");
  return RCMVS_OK;
}

extern "C" int rcmvs_bn_apply(const float* x, const float* mean, const float* invstd, const float* gamma,
                              const float* beta, const float* skip, float* z, int64_t M, int C, int relu,
                              rcmvs_stream_t stream) {
  RCMVS_REQUIRE(x && mean && invstd && gamma && beta && z && M > 0 && C % 4 == 0, RCMVS_ERR_BAD_ARG,
                "rcmvs_bn_apply: bad arguments");
  const int64_t n4 = M * C / 4;
  bn_apply_kernel<<<(unsigned)ceil_div<int64_t>(n4, 256), 256, 0, as_stream(stream)>>>(x, mean, invstd, gamma, beta, skip, z,
                                                                                     n4, C, relu);
  RCMVS_CHECK_LAUNCH("rcmvs_bn_apply");
  return RCMVS_OK;
}

// the two passes separately, for SyncBatchNorm: the caller all-reduces `sums` (and M) over the ranks in between
extern "C" int rcmvs_bn_bwd_reduce(const float* dz, const float* x, const float* mean, const float* invstd,
                                   const float* gamma, const float* beta, double* sums, int64_t M, int C, int relu,
                                   rcmvs_stream_t stream) {
  RCMVS_REQUIRE(dz && x && mean && invstd && gamma && beta && sums && M > 0 && C % 4 == 0 && C <= 256, RCMVS_ERR_BAD_ARG,
                "rcmvs_bn_bwd_reduce: bad arguments");
  const int rpc = rows_per_cta(M, 148 * 8);
  bn_bwd_reduce_kernel<<<(unsigned)ceil_div<int64_t>(M, rpc), 256, 2 * C * sizeof(double), as_stream(stream)>>>(
      dz, x, mean, invstd, gamma, beta, sums, M, C, relu, rpc);
  RCMVS_CHECK_LAUNCH("rcmvs_bn_bwd_reduce");
  return RCMVS_OK;
}

extern "C" int rcmvs_bn_bwd_apply(const float* dz, const float* x, const float* mean, const float* invstd,
                                  const float* gamma, const float* beta, const double* sums, float* dx, float* dgamma,
                                  float* dbeta, int64_t M, int64_t M_total, int C, int relu, int frozen,
                                  rcmvs_stream_t stream) {
  RCMVS_REQUIRE(dz && x && mean && invstd && gamma && beta && sums && dx && M > 0 && C % 4 == 0 && C <= 256, RCMVS_ERR_BAD_ARG,
                "rcmvs_bn_bwd_apply: bad arguments");
  const int64_t n4 = M * C / 4;
  bn_bwd_apply_kernel<<<(unsigned)ceil_div<int64_t>(n4, 256), 256, 0, as_stream(stream)>>>(
      dz, x, mean, invstd, gamma, beta, sums, dx, dgamma, dbeta, n4, (double)M_total, C, relu, frozen);
  RCMVS_CHECK_LAUNCH("rcmvs_bn_bwd_apply");
  return RCMVS_OK;
}

extern "C" int rcmvs_bn_bwd(const float* dz, const float* x, const float* mean, const float* invstd, const float* gamma,
                            const float* beta, double* sums, float* dx, float* dgamma, float* dbeta, int64_t M, int C,
                            int relu, int frozen, rcmvs_stream_t stream) {
  RCMVS_REQUIRE(dz && x && mean && invstd && gamma && beta && sums && dx && M > 0 && C % 4 == 0 && C <= 256, RCMVS_ERR_BAD_ARG,
                "rcmvs_bn_bwd: bad arguments");
  cudaStream_t st = as_stream(stream);
  const int rpc = rows_per_cta(M, 148 * 8);
  bn_bwd_reduce_kernel<<<(unsigned)ceil_div<int64_t>(M, rpc), 256, 2 * C * sizeof(double), st>>>(dz, x, mean, invstd, gamma,
                                                                                             beta, sums, M, C, relu, rpc);
  const int64_t n4 = M * C / 4;
  bn_bwd_apply_kernel<<<(unsigned)ceil_div<int64_t>(n4, 256), 256, 0, st>>>(dz, x, mean, invstd, gamma, beta, sums, dx,
                                                                          dgamma, dbeta, n4, (double)M, C, relu, frozen);
  RCMVS_CHECK_LAUNCH("rcmvs_bn_bwd");
  return RCMVS_OK;
}

// A/B switches for tools/train_experiments.py (ctypes in_dll; not part of the ABI): bit 0 = rotating register window
// instead of sliding by register moves, bit 1 = one S channel per thread instead of four, bit 2 = software prefetch of the next position's operands,
// bit 3 = register-window kernel instead of the shared-memory kernel, bit 4 = shared-memory kernel without the row ring,
// bit 5 = one-pixel-per-thread 2-D data-gradient kernel for the stride-1 layers.  Measured per training step
// (B200, 4 views 512x640): 0 -> 94.7 ms, 1 -> 106.1 ms, 2 -> 96.9 ms, 3 -> 109.5 ms.
extern "C" int rcmvs_wgrad_variant = 0;

// dims = {B, Ds, Hs, Ws, Cs, Db, Hb, Wb, Cb}; dim3 = 1: 3-D (27 taps, k = 3), 0: 2-D (Ds = Db = 1, k in {1,3,5})
extern "C" int rcmvs_conv_wgrad(const float* S, const float* Bg, float* G, int B, int Ds, int Hs, int Ws, int Cs, int Db,
                                int Hb, int Wb, int Cb, int k, int stride, int dim3d, rcmvs_stream_t stream) {
  RCMVS_REQUIRE(S && Bg && G, RCMVS_ERR_BAD_ARG, "rcmvs_conv_wgrad: null pointer");
  RCMVS_REQUIRE(B > 0 && Ds > 0 && Hs > 0 && Ws > 0 && Cs > 0 && Cb > 0 && (stride == 1 || stride == 2), RCMVS_ERR_BAD_ARG,
                "rcmvs_conv_wgrad: bad dims");
  cudaStream_t st0 = as_stream(stream);
  if (!(rcmvs_wgrad_variant & 8) && Cb % 4 == 0 && (dim3d ? k == 3 : (Ds == 1 && Db == 1))) {
    // shared-memory variant: channel blocks so that pairs (Cs-block / CSR x Cb-block) <= 256 and the rows fit 100 KB (two
    // CTAs per SM) or, failing that, 200 KB (one)
    const int KD = dim3d ? 3 : 1, pad = k / 2, taps = KD * k * k;
    const int csr = Cs % 4 == 0 ? 4 : 1;
    int CSB = csr == 4 ? (Cs < 32 ? Cs : 32) : 1;
    int CBB = Cb < 32 ? Cb : 32;
    while ((CSB / csr) * CBB > 256) CBB /= 2;
    const bool ring = !(rcmvs_wgrad_variant & 16);
    const int nslot = ring ? k + stride : k, nsb = ring ? 2 : 1;
    auto smem_of = [&](int csb_, int cbb_) {
      const size_t rows = ((size_t)nsb * Ws * csb_ + (size_t)KD * nslot * (Wb + 2 * pad) * cbb_) * 4;
      const size_t red = (size_t)csb_ * cbb_ * taps * 4;
      return rows > red ? rows : red;
    };
    const size_t kTwo = 100 * 1024, kOne = 200 * 1024;
    const bool one_cta = ring && csr * taps > 36;               // (those variants hold > 128 registers: one CTA per SM anyway)
    int csb2 = CSB, cbb2 = CBB;
    while (smem_of(csb2, cbb2) > kTwo && cbb2 > 4) cbb2 /= 2;
    while (smem_of(csb2, cbb2) > kTwo && csb2 > csr && csr == 4 && csb2 > 4) csb2 /= 2;
    if (!one_cta && smem_of(csb2, cbb2) <= kTwo) { CSB = csb2; CBB = cbb2; }
    else { while (smem_of(CSB, CBB) > kOne && CBB > 4) CBB /= 2; }
    const bool pow2 = (CBB & (CBB - 1)) == 0 && ((CSB / csr) & (CSB / csr - 1)) == 0;
    if (smem_of(CSB, CBB) <= kOne && Cs % CSB == 0 && Cb % CBB == 0 && pow2 && CBB % 4 == 0) {
      const size_t smem = smem_of(CSB, CBB);
      const int64_t nrows = (int64_t)B * Ds * Hs;
      const int gy = Cs / CSB, gz = Cb / CBB;
      int64_t target = one_cta ? 148 * 2 : 148 * 4;                     // ~2 waves of 1 or 2 CTAs per SM ...
      if (nrows * gy * gz <= 2 * target) target = 148;                  // ... or, for the small 2-D layers, one wave of long CTAs:
      int64_t rpc64 = ceil_div<int64_t>(nrows * gy * gz, target);       // the ring and the epilogue amortise over the rows
      if (rpc64 < 1) rpc64 = 1;
      if (rpc64 > nrows) rpc64 = nrows;
      dim3 grid((unsigned)ceil_div<int64_t>(nrows, rpc64), gy, gz);
#define RCMVS_WS_(KD_, KK, SS, CC, RR)                                                                                     \
  do {                                                                                                                     \
    static SmemAttrOnce attr;                                                                                              \
    cudaError_t e = attr.ensure(wgrad_smem_kernel<KD_, KK, SS, CC, RR>, (int)kOne);                                        \
    RCMVS_REQUIRE(e == cudaSuccess, RCMVS_ERR_CUDA, "rcmvs_conv_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); \
    wgrad_smem_kernel<KD_, KK, SS, CC, RR><<<grid, 256, smem, st0>>>(S, Bg, G, B, Ds, Hs, Ws, Cs, Db, Hb, Wb, Cb, (int)rpc64, \
                                                                     CSB, CBB);                                            \
  } while (0)
#define RCMVS_WS(KD_, KK, SS)                                                                                              \
  do {                                                                                                                     \
    if (csr == 4) { if (ring) RCMVS_WS_(KD_, KK, SS, 4, true); else RCMVS_WS_(KD_, KK, SS, 4, false); }                    \
    else { if (ring) RCMVS_WS_(KD_, KK, SS, 1, true); else RCMVS_WS_(KD_, KK, SS, 1, false); }                             \
  } while (0)
      bool launched = true;
      if (dim3d) { if (stride == 1) RCMVS_WS(3, 3, 1); else RCMVS_WS(3, 3, 2); }
      else if (k == 1 && stride == 1) RCMVS_WS(1, 1, 1);
      else if (k == 3 && stride == 1) RCMVS_WS(1, 3, 1);
      else if (k == 3 && stride == 2) RCMVS_WS(1, 3, 2);
      else if (k == 5 && stride == 1) RCMVS_WS(1, 5, 1);
      else if (k == 5 && stride == 2) RCMVS_WS(1, 5, 2);
      else launched = false;
#undef RCMVS_WS
#undef RCMVS_WS_
      if (launched) {
        RCMVS_CHECK_LAUNCH("rcmvs_conv_wgrad");
        return RCMVS_OK;
      }
    }
  }
  const int cbb = Cb >= 32 ? 32 : (Cb >= 16 ? 16 : (Cb >= 8 ? 8 : (Cb >= 4 ? 4 : (Cb >= 2 ? 2 : 1))));
  const int csr = (Cs % 4 == 0 && !(rcmvs_wgrad_variant & 2)) ? 4 : 1;   // S channels per thread
  const int csg = Cs / csr;                                             // channel groups along S
  int csb = 1;
  while (csb < csg && csb * cbb < 256) csb *= 2;         // power of two >= min(csg, 256 / cbb)
  const int groups = 256 / (cbb * csb);
  const int64_t nrows = (int64_t)B * Ds * Hs;
  const int gy = ceil_div(csg, csb), gz = ceil_div(Cb, cbb);
  // ~6 waves of CTAs over the row dimension, at least 2 rows per thread group so the atomics stay a small share
  int64_t rpc64 = ceil_div<int64_t>(nrows * gy * gz, 148 * 6);
  if (rpc64 < 2 * groups) rpc64 = 2 * groups;
  if (rpc64 > nrows) rpc64 = nrows;
  const int rpc = (int)rpc64;
  dim3 grid((unsigned)ceil_div<int64_t>(nrows, rpc), gy, gz);
  cudaStream_t st = as_stream(stream);
  // measured per training step (B200, 4 views 512x640): moves 98.0 ms, moves + prefetch 101-103 ms (197 registers), rotating 106 ms
  const int mode = (rcmvs_wgrad_variant & 1) ? 1 : ((rcmvs_wgrad_variant & 4) ? 2 : 0);
#define RCMVS_WG_(KD, KK, SS, MM, CC) \
  wgrad_kernel<KD, KK, SS, MM, CC><<<grid, 256, 0, st>>>(S, Bg, G, B, Ds, Hs, Ws, Cs, Db, Hb, Wb, Cb, rpc, cbb, csb)
#define RCMVS_WG(KD, KK, SS)                                                          \
  do {                                                                                \
    if (csr == 4) {                                                                   \
      if (mode == 2) RCMVS_WG_(KD, KK, SS, 2, 4);                                     \
      else if (mode == 1) RCMVS_WG_(KD, KK, SS, 1, 4);                                \
      else RCMVS_WG_(KD, KK, SS, 0, 4);                                               \
    } else {                                                                          \
      if (mode == 2) RCMVS_WG_(KD, KK, SS, 2, 1);                                     \
      else if (mode == 1) RCMVS_WG_(KD, KK, SS, 1, 1);                                \
      else RCMVS_WG_(KD, KK, SS, 0, 1);                                               \
    }                                                                                 \
  } while (0)
  if (dim3d) {
    RCMVS_REQUIRE(k == 3, RCMVS_ERR_UNSUPPORTED, "rcmvs_conv_wgrad: 3-D kernels are 3x3x3");
    if (stride == 1) RCMVS_WG(3, 3, 1); else RCMVS_WG(3, 3, 2);
  } else {
    RCMVS_REQUIRE(Ds == 1 && Db == 1, RCMVS_ERR_BAD_ARG, "rcmvs_conv_wgrad: 2-D needs Ds = Db = 1");
    if (k == 1 && stride == 1) RCMVS_WG(1, 1, 1);
    else if (k == 3 && stride == 1) RCMVS_WG(1, 3, 1);
    else if (k == 3 && stride == 2) RCMVS_WG(1, 3, 2);
    else if (k == 5 && stride == 1) RCMVS_WG(1, 5, 1);
    else if (k == 5 && stride == 2) RCMVS_WG(1, 5, 2);
    else { set_error("rcmvs_conv_wgrad: unsupported 2-D k=%d stride=%d", k, stride); return RCMVS_ERR_UNSUPPORTED; }
  }
#undef RCMVS_WG
#undef RCMVS_WG_
  RCMVS_CHECK_LAUNCH("rcmvs_conv_wgrad");
  return RCMVS_OK;
}

extern "C" int rcmvs_conv2d_dgrad(const float* dy, const float* w, float* wd_ws, float* dx, int N, int H, int W, int Cin,
                                  int Cout, int k, int stride, rcmvs_stream_t stream) {
  RCMVS_REQUIRE(dy && w && wd_ws && dx, RCMVS_ERR_BAD_ARG, "rcmvs_conv2d_dgrad: null pointer");
  RCMVS_REQUIRE(N > 0 && H > 0 && W > 0 && Cout % 4 == 0 && (stride == 1 || stride == 2), RCMVS_ERR_BAD_ARG,
                "rcmvs_conv2d_dgrad: bad dims");
  const int cin_pad = (Cin + 3) / 4 * 4;
  RCMVS_REQUIRE(cin_pad == Cin, RCMVS_ERR_UNSUPPORTED, "rcmvs_conv2d_dgrad: Cin must be a multiple of 4");
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  const int kk = k * k;
  cudaStream_t st = as_stream(stream);
  pack_dgrad2d_kernel<<<ceil_div(kk * Cout * cin_pad, 256), 256, 0, st>>>(w, wd_ws, Cout, Cin, cin_pad, kk);
  const size_t smem = (size_t)kk * Cout * cin_pad * 4;
  RCMVS_REQUIRE(smem <= 200 * 1024, RCMVS_ERR_UNSUPPORTED, "rcmvs_conv2d_dgrad: weights do not fit shared memory");
  const int64_t threads = (int64_t)N * H * W * (cin_pad / 4);
  const unsigned grid = (unsigned)ceil_div<int64_t>(threads, 256);
#define RCMVS_DG(KK, SS)                                                                                        \
  {                                                                                                             \
    static SmemAttrOnce attr;                                                                                   \
    cudaError_t e = attr.ensure(conv2d_dgrad_kernel<KK, SS>, 200 * 1024);                                       \
    RCMVS_REQUIRE(e == cudaSuccess, RCMVS_ERR_CUDA, "rcmvs_conv2d_dgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); \
    conv2d_dgrad_kernel<KK, SS><<<grid, 256, smem, st>>>(dy, wd_ws, dx, N, H, W, cin_pad, Ho, Wo, Cout);        \
  }
#define RCMVS_DG1(KK)                                                                                           \
  {                                                                                                             \
    static SmemAttrOnce attr;                                                                                   \
    cudaError_t e = attr.ensure(conv2d_dgrad_s1_kernel<KK>, 200 * 1024);                                        \
    RCMVS_REQUIRE(e == cudaSuccess, RCMVS_ERR_CUDA, "rcmvs_conv2d_dgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); \
    const int64_t thr4 = (int64_t)N * H * ((W + 3) / 4) * (cin_pad / 4);                                        \
    const int per_sm = smem > 64 * 1024 ? 2 : (smem > 32 * 1024 ? 4 : 6);                                       \
    const unsigned g4 = (unsigned)std::min<int64_t>(ceil_div<int64_t>(thr4, 256), (int64_t)148 * per_sm);       \
    conv2d_dgrad_s1_kernel<KK><<<g4, 256, smem, st>>>(dy, wd_ws, dx, N, H, W, cin_pad, Cout);                   \
  }
  if (k == 1 && stride == 1) RCMVS_DG(1, 1)
  else if (k == 3 && stride == 1 && !(rcmvs_wgrad_variant & 32)) RCMVS_DG1(3)
  else if (k == 5 && stride == 1 && !(rcmvs_wgrad_variant & 32)) RCMVS_DG1(5)
  else if (k == 3 && stride == 1) RCMVS_DG(3, 1)
  else if (k == 3 && stride == 2) RCMVS_DG(3, 2)
  else if (k == 5 && stride == 1) RCMVS_DG(5, 1)
  else if (k == 5 && stride == 2) RCMVS_DG(5, 2)
  else { set_error("rcmvs_conv2d_dgrad: unsupported k=%d stride=%d", k, stride); return RCMVS_ERR_UNSUPPORTED; }
#undef RCMVS_DG
#undef RCMVS_DG1
  RCMVS_CHECK_LAUNCH("rcmvs_conv2d_dgrad");
  return RCMVS_OK;
}

extern "C" int rcmvs_conv3d_c1_dgrad(const float* dcost, const float* w, float* dx, int B, int D, int H, int W,
                                     rcmvs_stream_t stream) {
  RCMVS_REQUIRE(dcost && w && dx && B > 0 && D > 0 && H > 0 && W > 0, RCMVS_ERR_BAD_ARG, "rcmvs_conv3d_c1_dgrad: bad arguments");
  const int64_t n = (int64_t)B * D * H * W;
  conv3d_c1_dgrad_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, as_stream(stream)>>>(dcost, w, dx, B, D, H, W);
  RCMVS_CHECK_LAUNCH("rcmvs_conv3d_c1_dgrad");
  return RCMVS_OK;
}

extern "C" int rcmvs_pack_conv3d_dgrad_s1(const float* w, float* packed, int Cout, int Cin, rcmvs_stream_t stream) {
  RCMVS_REQUIRE(w && packed && Cout > 0 && Cin > 0, RCMVS_ERR_BAD_ARG, "rcmvs_pack_conv3d_dgrad_s1: bad arguments");
  pack_dgrad3d_s1_kernel<<<ceil_div(27 * Cout * Cin, 256), 256, 0, as_stream(stream)>>>(w, packed, Cout, Cin);
  RCMVS_CHECK_LAUNCH("rcmvs_pack_conv3d_dgrad_s1");
  return RCMVS_OK;
}

extern "C" int rcmvs_costvolume_var_bwd(const float* feats, const float* rt, const float* depth, int per_pixel,
                                        const float* grad_var, const float* grad_noref, float* grad_feats, int B, int V,
                                        int C, int D, int h, int w, rcmvs_stream_t stream) {
  RCMVS_REQUIRE(feats && rt && depth && grad_var && grad_feats, RCMVS_ERR_BAD_ARG, "rcmvs_costvolume_var_bwd: null pointer");
  RCMVS_REQUIRE(B > 0 && V > 1 && V - 1 <= kMaxSrcB && D > 0 && h > 0 && w > 0 && h <= 65535 && B <= 65535,
                RCMVS_ERR_BAD_ARG, "rcmvs_costvolume_var_bwd: bad dims (at most %d source views)", kMaxSrcB);
  RCMVS_REQUIRE((int64_t)h * w * C < (1ll << 31), RCMVS_ERR_UNSUPPORTED, "rcmvs_costvolume_var_bwd: feature map too large");
  cudaStream_t st = as_stream(stream);
#define RCMVS_VB(CC)                                                                                                   \
  {                                                                                                                    \
    dim3 grid(ceil_div(w, 256 / (CC / 4)), h, B);                                                                      \
    if (per_pixel) costvolume_var_bwd_kernel<CC, true><<<grid, 256, 0, st>>>(feats, rt, depth, grad_var, grad_noref, grad_feats, V, D, h, w); \
    else costvolume_var_bwd_kernel<CC, false><<<grid, 256, 0, st>>>(feats, rt, depth, grad_var, grad_noref, grad_feats, V, D, h, w);          \
  }
  if (C == 8) RCMVS_VB(8) else if (C == 16) RCMVS_VB(16) else if (C == 32) RCMVS_VB(32)
  else { set_error("rcmvs_costvolume_var_bwd: C=%d not in {8,16,32}", C); return RCMVS_ERR_UNSUPPORTED; }
#undef RCMVS_VB
  RCMVS_CHECK_LAUNCH("rcmvs_costvolume_var_bwd");
  return RCMVS_OK;
}

extern "C" int rcmvs_softargmin_bwd(const float* cost, const float* depth_planes, int per_pixel, const float* grad_depth,
                                    float* grad_cost, int B, int D, int h, int w, rcmvs_stream_t stream) {
  RCMVS_REQUIRE(cost && depth_planes && grad_depth && grad_cost && B > 0 && D > 0 && h > 0 && w > 0 && B <= 65535,
                RCMVS_ERR_BAD_ARG, "rcmvs_softargmin_bwd: bad arguments");
  const int64_t hw = (int64_t)h * w;
  dim3 grid((unsigned)ceil_div<int64_t>(hw, 128), B);
  if (per_pixel) softargmin_bwd_kernel<true><<<grid, 128, 0, as_stream(stream)>>>(cost, depth_planes, grad_depth, grad_cost, D, hw);
  else softargmin_bwd_kernel<false><<<grid, 128, 0, as_stream(stream)>>>(cost, depth_planes, grad_depth, grad_cost, D, hw);
  RCMVS_CHECK_LAUNCH("rcmvs_softargmin_bwd");
  return RCMVS_OK;
}

extern "C" int rcmvs_sum2x2(const float* in, float* out, int N, int Ho, int Wo, int C, rcmvs_stream_t stream) {
  RCMVS_REQUIRE(in && out && N > 0 && Ho > 0 && Wo > 0 && C % 4 == 0, RCMVS_ERR_BAD_ARG, "rcmvs_sum2x2: bad arguments");
  const int64_t n = (int64_t)N * Ho * Wo * (C / 4);
  sum2x2_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, as_stream(stream)>>>(in, out, N, Ho, Wo, C);
  RCMVS_CHECK_LAUNCH("rcmvs_sum2x2");
  return RCMVS_OK;
}

extern "C" int rcmvs_up2_add(const float* a, const float* up, float* out, int N, int H, int W, int C, rcmvs_stream_t stream) {
  RCMVS_REQUIRE(a && up && out && N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C % 4 == 0, RCMVS_ERR_BAD_ARG,
                "rcmvs_up2_add: bad arguments");
  const int64_t n = (int64_t)N * H * W * (C / 4);
  up2_add_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, as_stream(stream)>>>(a, up, out, N, H, W, C);
  RCMVS_CHECK_LAUNCH("rcmvs_up2_add");
  return RCMVS_OK;
}

extern "C" int rcmvs_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                               float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                               rcmvs_stream_t stream) {
  RCMVS_REQUIRE(params && grads && exp_avg && exp_avg_sq && n > 0 && step > 0, RCMVS_ERR_BAD_ARG, "rcmvs_adam_step: bad arguments");
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
  adam_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, as_stream(stream)>>>(params, grads, exp_avg, exp_avg_sq, n, lr,
                                                                                beta1, beta2, eps, weight_decay, bc1,
                                                                                sqrtf(bc2), grad_scale);
  RCMVS_CHECK_LAUNCH("rcmvs_adam_step");
  return RCMVS_OK;
}
