// a3 CostRegNet stride-1 3x3x3 convolutions, round-2 engine: TMA-fed tcgen05 (kind::f16, bf16 operands, fp32 accumulators
// in TMEM) with error-compensated split-bf16 operands ("bf16x3").  Reference models/modules.py:118-163 (Conv3d = conv ->
// BN -> ReLU), :473-475 (conv0 / conv2 of CostRegNet).
//
// Why bf16x3: the parity bar (depth within 1e-3 relative, BASELINE north_star) needs better than TF32 operands on
// ill-conditioned input.  Round 1 used 3xTF32 (x = hi + lo in TF32, 3 MMAs of K=8 per 8 channels).  The
// same three products with a bf16 split, x = hi + lo, hi = bf16(x), lo = bf16(x - hi) (|x - hi - lo| <= 2^-17 |x|):
//      x·w ~= x_lo·w_hi + x_hi·w_lo + x_hi·w_hi          (dropped terms <= 3·2^-17 relative)
// run on kind::f16 MMAs of K=16: half the tensor-core instructions per channel and half the shared-memory bytes per
// operand.  On the full-size golden (5 views 512x640 D=192/64/8) the emulated scheme is 1.7e-5 relative L-inf from
// the reference's fp32 depth (plain TF32: 5e-3), see DESIGN §5.
//
// Activations arrive PRE-SPLIT from the producing kernel as two bf16 planes [2][B][D][H][W][C] (hi, lo: the same 4
// bytes per element as fp32), so operand staging is pure data movement: one elected thread issues 5-D TMA box loads
// (cp.async.bulk.tensor) of {8 channels, 32 columns, ROWS rows, 1 plane} per 8-channel chunk, out-of-bounds
// coordinates are zero-filled by the TMA unit = the convolution's zero padding in x, y AND z, and the landed box IS the
// UMMA canonical K-major SWIZZLE_NONE operand layout (core matrix = 8 positions x 16 bytes).  No loader warps, no
// bounds checks, no register staging.
//
// MMA structure (unchanged from the TF32 slab kernel, conv3d_umma.cu): M = 128 input positions = a flat run of the
// zero-padded (ROWS x 32) slab plane, kw = +1 position = +16 bytes of descriptor start address; N folds (kd, kh, cout):
// the three kd blocks of N are the accumulators of three different OUTPUT planes held in a TMEM ring of per-plane column
// groups (the sum over kd happens inside TMEM), the sum over kh is done by the epilogue through shared memory.
//
// Warp roles: warps 0-3 epilogue (TMEM -> registers -> smem exchange -> bias/ReLU/skip -> global, fp32 and/or split-bf16
// for the next tensor-core layer), warp 4 TMA producer, warp 5 MMA issuer.
#include "slab_common.cuh"

namespace rcmvs {
namespace slab {

template <int CIN, int COUT, int KM, int NS>
struct Cfg {
  static constexpr int CP = COUT < 8 ? 8 : COUT;                  // channel slots per kh block
  static constexpr int G = ((3 * CP + 15) / 16) * 16;             // TMEM columns / B rows per plane block
  static constexpr int KS = CIN / 16;                             // K steps (16 bf16 channels each) per kw tap, CIN >= 16
  // CIN = 8: one K = 16 step covers TWO kw taps of the 8 channels — K-chunk 1 is the next position (LBO = 16 B); the
  // second step is (kw = 2, zero): its chunk 1 reads a zero block (and has zero weights)
  static constexpr int NSTEP = CIN == 8 ? 2 : 3 * KS;             // MMA steps per (slab, M-tile, term)
  static constexpr int NCH = CIN / 8;                             // 16-byte channel chunks
  static constexpr int ROWS = 4 * KM;
  static constexpr int OROWS = ROWS - 2;
  static constexpr int PLANE_B = ROWS * PW * 16;                  // one chunk plane = one TMA box
  static constexpr int SLAB_B = NCH * PLANE_B;                    // one part (hi or lo)
  static constexpr int STAGE_B = 2 * SLAB_B;
  static constexpr int ZERO_OFF = NS * STAGE_B + 128;             // (CIN = 8) 2 KB of zeros: K-chunk 1 of the (kw = 2, zero) step
  static constexpr int A_BYTES = NS * STAGE_B + 128 + (CIN == 8 ? 2048 + 128 : 0);   // + the kw overrun of the last M-tile
  static constexpr int BBLK_B = 2 * 3 * G * 16;                   // one step: 2 chunks x 3G rows x 16 B
  static constexpr int BTERM_B = NSTEP * BBLK_B;
  static constexpr int B_BYTES = 2 * BTERM_B;                     // hi, lo
  static constexpr int NG = 512 / (KM * G);
  static constexpr int EX_B = 3 * OROWS * PW * CP * 4;
  static constexpr size_t SMEM = (size_t)A_BYTES + B_BYTES + EX_B + 512;
  static_assert(CIN == 8 || CIN % 16 == 0, "K step = 16 bf16 channels");
  static_assert(NG >= 4, "TMEM ring needs at least 4 plane groups");
  static_assert(PLANE_B % 128 == 0, "TMA destination alignment");
};

struct Params {
  const __nv_bfloat16* wimg;   // packed B image
  const float* bias;           // [Cout] or null
  const float* skip;           // [B][D][H][W][Cout] fp32 or null
  float* y;                    // [B][D][H][W][Cout] fp32 or null
  __nv_bfloat16* ys;           // [2][B][D][H][W][Cout] split-bf16 or null
  int B, D, H, W;
  int d_chunk, chunks;
  int relu;
  int dbg;                     // experiment switches (tools/slab_experiments.py); 0 in the product path
  long long* trace;            // experiment: clock64 timeline of two CTAs (null in the product path)
};

template <int CIN, int COUT, int KM, int NS>
__global__ void __launch_bounds__(kThreads, 1)
conv3d_slab_bf16_kernel(const __grid_constant__ CUtensorMap tmap, const Params p) {
  using C = Cfg<CIN, COUT, KM, NS>;
  constexpr int G = C::G, NG = C::NG, CP = C::CP;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* a_smem = smem;
  unsigned char* b_smem = smem + C::A_BYTES;
  float* ex = reinterpret_cast<float*>(smem + C::A_BYTES + C::B_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::A_BYTES + C::B_BYTES + C::EX_B);
  uint64_t* full = bars;                 // [NS]  TMA -> MMA
  uint64_t* empty = bars + NS;           // [NS]  MMA -> TMA
  uint64_t* grp_full = bars + 2 * NS;    // [NG]  MMA -> epilogue   (plane complete)
  uint64_t* grp_free = grp_full + NG;    // [NG]  epilogue -> MMA   (group drained and zeroed)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(grp_free + NG);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * C::OROWS;
  const int b = blockIdx.z / p.chunks;
  const int d0 = (blockIdx.z % p.chunks) * p.d_chunk;
  const int d1 = min(p.D, d0 + p.d_chunk);
  const int nplanes = d1 - d0;           // output planes d0..d1-1 ; input slabs z = d0-1 .. d1  (nplanes + 2)
  const int nslabs = nplanes + 2;
  // experiment: timeline of the first CTA and of the middle one: [0] start, [1] prologue done, [2] end, [3] nslabs,
  // [8+i] producer got the stage of slab i, [72+i] issuer saw slab i landed, [136+i] issuer finished issuing slab i,
  // [200+q] epilogue saw plane q complete, [264+q] epilogue stored plane q
  long long* tr = nullptr;
  if (p.trace) {
    const unsigned lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    const unsigned total = gridDim.x * gridDim.y * gridDim.z;
    if (lin == 0) tr = p.trace;
    else if (lin == total / 2) tr = p.trace + 512;
  }
  if (tr && threadIdx.x == 0) { tr[0] = clock64(); tr[3] = nslabs; }

  // ---- one-time setup -------------------------------------------------------------------------------------------
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    for (int i = 0; i < NG; ++i) { mbar_init(grp_full + i, 1); mbar_init(grp_free + i, 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4 && lane == 0) tma_prefetch_desc(&tmap);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {
    const float4* src = reinterpret_cast<const float4*>(p.wimg);
    float4* dst = reinterpret_cast<float4*>(b_smem);
    for (int i = threadIdx.x; i < C::B_BYTES / 16; i += kThreads) dst[i] = __ldg(src + i);
    // the overrun pad behind the last stage is read (into discarded accumulator rows) by the kw = 1, 2 taps
    for (int i = threadIdx.x; i < (C::A_BYTES - NS * C::STAGE_B) / 16; i += kThreads)
      reinterpret_cast<float4*>(a_smem + NS * C::STAGE_B)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp < 4) {
    // all accumulators start at zero: every MMA accumulates, a group is re-zeroed by the epilogue after it is drained
    for (int c = 0; c < 512; c += 16) tmem_st16_zero(tmem_base + ((uint32_t)(warp * 32) << 16) + c);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tr && threadIdx.x == 0) tr[1] = clock64();

  if (warp == 4) {
    // =============================== TMA producer =================================================================
    // slab ii = input plane z = d0-1+ii: 2 parts x NCH boxes {8 ch, 32 cols, ROWS rows, 1 plane} at (x0-1, y0-1, z).
    // Out-of-range x / y / z are zero-filled by the TMA unit (the convolution's padding), the full box always lands.
    if (lane == 0) {
      for (int ii = 0; ii < nslabs; ++ii) {
        const int buf = ii % NS;
        mbar_wait(empty + buf, ((ii / NS) & 1) ^ 1);
        if (tr && ii < 64) tr[8 + ii] = clock64();
        if (p.dbg & 8) { mbar_arrive(full + buf); continue; }   // experiment: no TMA traffic
        mbar_arrive_expect_tx(full + buf, (uint32_t)C::STAGE_B);
        const int z = d0 - 1 + ii;
        const uint32_t dst0 = smem_u32(a_smem + (size_t)buf * C::STAGE_B);
#pragma unroll
        for (int part = 0; part < 2; ++part)
#pragma unroll
          for (int ch = 0; ch < C::NCH; ++ch)
            tma_load_5d(dst0 + part * C::SLAB_B + ch * C::PLANE_B, &tmap, full + buf, ch * 8, x0 - 1, y0 - 1, z,
                        part * p.B + b);
      }
    }
  } else if (warp == 5) {
    // =============================== MMA issuer ==================================================================
    const bool leader = elect_one();
    const uint32_t b_base = smem_u32(b_smem);
    for (int i = 0; i < nslabs; ++i) {
      // slab i = input plane z = d0-1+i contributes to output planes (relative) q = i-2+j, block j = 0,1,2 (kd = 2-j)
      const int jlo = max(0, 2 - i), jhi = min(2, nplanes + 1 - i);       // valid blocks: 0 <= q < nplanes
      const int qnew = i;                                                  // plane first touched by this slab (j=2)
      if (qnew < nplanes && qnew >= NG) mbar_wait(grp_free + (qnew % NG), ((qnew / NG) - 1) & 1);
      const int buf = i % NS;
      mbar_wait(full + buf, (i / NS) & 1);
      tc_fence_after();
      if (tr && leader && i < 64) tr[72 + i] = clock64();
      const uint32_t a_base = smem_u32(a_smem + (size_t)buf * C::STAGE_B);
      // split the valid block range into (at most two) ring-contiguous runs; per run: TMEM column, B row offset, N
      const int g0 = (i - 2 + jlo + NG) % NG;
      const int n_all = jhi - jlo + 1;
      const int n0 = min(n_all, NG - g0), n1 = n_all - n0;                 // second run wraps to group 0
      const uint32_t d0c = tmem_base + g0 * G, d1c = tmem_base;
      const uint32_t idesc0 = make_idesc_bf16(128, (p.dbg & 32) ? G : n0 * G), idesc1 = make_idesc_bf16(128, (n1 > 0 ? n1 : 1) * G);
      const uint32_t boff0 = (uint32_t)(jlo * G), boff1 = (uint32_t)((jlo + n0) * G);   // in 16-byte rows
      const uint32_t b_lo0 = desc_lo(b_base, 3 * G * 16);
#pragma unroll
      for (int st = 0; st < C::NSTEP; ++st) {
        constexpr int KSD = C::KS > 0 ? C::KS : 1;
        const int kw = CIN == 8 ? 2 * st : st / KSD, ks = CIN == 8 ? 0 : st % KSD;
#pragma unroll
        for (int mt = 0; mt < KM; ++mt) {
          constexpr uint32_t kLoOff = C::SLAB_B >> 4, kBLoOff = C::BTERM_B >> 4;
          uint32_t a_hi_d, a_lo_d;
          if constexpr (CIN == 8) {
            const uint32_t a_off = (uint32_t)((mt * 128 + kw) * 16);               // bytes from the slab start
            if (st == 0) {                                                        // (kw0, kw1): chunk 1 = next position
              a_hi_d = desc_lo(a_base + a_off, 16);
              a_lo_d = a_hi_d + kLoOff;
            } else {                                                              // (kw2, zero): chunk 1 = zero block
              const uint32_t zb = smem_u32(a_smem) + C::ZERO_OFF;
              a_hi_d = desc_lo(a_base + a_off, zb - (a_base + a_off));
              a_lo_d = desc_lo(a_base + C::SLAB_B + a_off, zb - (a_base + C::SLAB_B + a_off));
            }
          } else {
            a_hi_d = desc_lo(a_base, C::PLANE_B) + (uint32_t)((ks * 2 * C::PLANE_B) >> 4) + (uint32_t)(mt * 128 + kw);
            a_lo_d = a_hi_d + kLoOff;
          }
          const uint32_t b_blk_d = b_lo0 + (uint32_t)((st * C::BBLK_B) >> 4);
          const uint32_t dm = (uint32_t)(mt * (NG * G));
          if (leader && !((p.dbg & 16) && kw > 0)) {
            if (!(p.dbg & 1)) {
              umma_bf16_lohi(d0c + dm, a_lo_d, b_blk_d + boff0, idesc0);                   // x_lo · w_hi
              umma_bf16_lohi(d0c + dm, a_hi_d, b_blk_d + boff0 + kBLoOff, idesc0);         // x_hi · w_lo
            }
            umma_bf16_lohi(d0c + dm, a_hi_d, b_blk_d + boff0, idesc0);                     // x_hi · w_hi
            if (n1 > 0) {
              if (!(p.dbg & 1)) {
                umma_bf16_lohi(d1c + dm, a_lo_d, b_blk_d + boff1, idesc1);
                umma_bf16_lohi(d1c + dm, a_hi_d, b_blk_d + boff1 + kBLoOff, idesc1);
              }
              umma_bf16_lohi(d1c + dm, a_hi_d, b_blk_d + boff1, idesc1);
            }
          }
        }
      }
      if (leader) {
        umma_commit(empty + buf);                                   // slab consumed once these MMAs retire
        if (i >= 2) umma_commit(grp_full + ((i - 2) % NG));         // output plane q = i-2 is complete
        if (tr && i < 64) tr[136 + i] = clock64();
      }
      __syncwarp();
    }
  } else {
    // =============================== epilogue (128 threads) ======================================================
    const int gx = x0 + lane;
    const bool col_ok = lane < TW && gx < p.W;
    const int64_t part_stride = (int64_t)p.B * p.D * p.H * p.W * COUT;       // elements between the hi and lo planes of ys
    for (int q = 0; q < nplanes; ++q) {
      const int g = q % NG;
      mbar_wait(grp_full + g, (q / NG) & 1);
      tc_fence_after();
      if (tr && threadIdx.x == 0 && q < 64) tr[200 + q] = clock64();
      uint32_t vr[KM][G];
#pragma unroll
      for (int mt = 0; mt < KM; ++mt) {
        const uint32_t ta = tmem_base + ((uint32_t)(warp * 32) << 16) + mt * (NG * G) + g * G;
#pragma unroll
        for (int c0 = 0; c0 < G; c0 += 16) tmem_ld16_nowait(ta + c0, &vr[mt][c0]);
      }
      tmem_ld_wait();                                            // one wait for all loads of this plane
#pragma unroll
      for (int mt = 0; mt < KM; ++mt) {
        const uint32_t ta = tmem_base + ((uint32_t)(warp * 32) << 16) + mt * (NG * G) + g * G;
#pragma unroll
        for (int c0 = 0; c0 < G; c0 += 16) tmem_st16_zero(ta + c0);
      }
      float v[KM][G];
#pragma unroll
      for (int mt = 0; mt < KM; ++mt)
#pragma unroll
        for (int c = 0; c < G; ++c) v[mt][c] = __uint_as_float(vr[mt][c]);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(grp_free + g);
      if (p.dbg & 4) continue;                                   // experiment: drain only
      // ---- sum over kh: input row ri = 4*mt + warp contributes its kh block to output row j = ri - kh.  Each kh
      //      writes its own exchange plane (no read-modify-write), one barrier, then the store phase adds the three.
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
        for (int mt = 0; mt < KM; ++mt) {
          const int j = 4 * mt + warp - kh;
          if (j >= 0 && j < C::OROWS) {
            // exchange layout [kh][row][channel quad][lane][4]: a warp's 128-bit stores are dense (bank-conflict free)
            float* e = ex + ((kh * C::OROWS + j) * (CP / 4) * PW + lane) * 4;
#pragma unroll
            for (int c = 0; c < CP; c += 4)
              *reinterpret_cast<float4*>(e + (c / 4) * PW * 4) =
                  make_float4(v[mt][kh * CP + c], v[mt][kh * CP + c + 1], v[mt][kh * CP + c + 2], v[mt][kh * CP + c + 3]);
          }
        }
      }
      epi_bar_sync();
      // ---- bias / ReLU / skip, coalesced store: thread handles rows j = warp, warp+4, ... ------------------------
      const int d = d0 + q;
      for (int j = warp; j < C::OROWS; j += 4) {
        const int gy = y0 + j;
        if (!col_ok || gy >= p.H) continue;
        const float* e = ex + (j * (CP / 4) * PW + lane) * 4;
        constexpr int EXP = C::OROWS * PW * CP;                  // floats per kh plane
        const int64_t o = ((((int64_t)b * p.D + d) * p.H + gy) * p.W + gx) * COUT;
        if constexpr (COUT == 1) {
          float r = e[0] + e[EXP] + e[2 * EXP];
          if (p.bias) r += __ldg(p.bias);
          if (p.relu) r = fmaxf(r, 0.f);
          if (p.skip) r += __ldg(p.skip + o);
          if (p.y) p.y[o] = r;
        } else {
#pragma unroll
          for (int c = 0; c < COUT; c += 4) {
            float4 r = *reinterpret_cast<const float4*>(e + (c / 4) * PW * 4);
            {
              const float4 r1 = *reinterpret_cast<const float4*>(e + EXP + (c / 4) * PW * 4);
              const float4 r2 = *reinterpret_cast<const float4*>(e + 2 * EXP + (c / 4) * PW * 4);
              r.x += r1.x + r2.x; r.y += r1.y + r2.y; r.z += r1.z + r2.z; r.w += r1.w + r2.w;
            }
            if (p.bias) {
              const float4 t4 = ldg4(p.bias + c);
              r.x += t4.x; r.y += t4.y; r.z += t4.z; r.w += t4.w;
            }
            if (p.relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
            if (p.skip) {
              const float4 t4 = ldg4(p.skip + o + c);
              r.x += t4.x; r.y += t4.y; r.z += t4.z; r.w += t4.w;
            }
            if (p.y) st4(p.y + o + c, r);
            if (p.ys) {
              // pre-split copy for the next tensor-core layer: hi = bf16(v), lo = bf16(v - hi)
              const __nv_bfloat16 h0 = __float2bfloat16_rn(r.x), h1 = __float2bfloat16_rn(r.y);
              const __nv_bfloat16 h2 = __float2bfloat16_rn(r.z), h3 = __float2bfloat16_rn(r.w);
              const __nv_bfloat162 hi01 = __halves2bfloat162(h0, h1), hi23 = __halves2bfloat162(h2, h3);
              const __nv_bfloat162 lo01 = __floats2bfloat162_rn(r.x - __bfloat162float(h0), r.y - __bfloat162float(h1));
              const __nv_bfloat162 lo23 = __floats2bfloat162_rn(r.z - __bfloat162float(h2), r.w - __bfloat162float(h3));
              uint2 hv, lv;
              hv.x = *reinterpret_cast<const uint32_t*>(&hi01); hv.y = *reinterpret_cast<const uint32_t*>(&hi23);
              lv.x = *reinterpret_cast<const uint32_t*>(&lo01); lv.y = *reinterpret_cast<const uint32_t*>(&lo23);
              *reinterpret_cast<uint2*>(p.ys + o + c) = hv;
              *reinterpret_cast<uint2*>(p.ys + part_stride + o + c) = lv;
            }
          }
        }
      }
      epi_bar_sync();                              // exchange buffer is reused by the next plane
      if (tr && threadIdx.x == 0 && q < 64) tr[264 + q] = clock64();
    }
  }

  // ---- teardown -------------------------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (tr && threadIdx.x == 0) tr[2] = clock64();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// Weight image: torch Conv3d weight [Cout][Cin][27] x BN scale ->
//   img[term][step][chunk][row = jblk*G + kh*CP + n][e],  kd = 2 - jblk
//     Cin >= 16: step = kw*KS + ks, channel = ks*16 + chunk*8 + e
//     Cin == 8 : step s covers taps kw = 2s + chunk (kw = 3: zero), channel = e
//   term 0 = bf16(w·scale), term 1 = bf16(w·scale - term 0); zero for padded rows
__global__ void pack_slab_weight_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                        __nv_bfloat16* __restrict__ img, int Cout, int Cin, int CP, int G) {
  const int KS = Cin / 16;
  const int nstep = Cin == 8 ? 2 : 3 * KS;
  const int per_term = nstep * 2 * 3 * G * 8;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per_term) return;
  const int e = i & 7;
  int r = i >> 3;
  const int row = r % (3 * G); r /= (3 * G);
  const int chunk = r & 1; r >>= 1;
  const int st = r;
  const int kw = Cin == 8 ? 2 * st + chunk : st / KS;
  const int ci = Cin == 8 ? e : (st % KS) * 16 + chunk * 8 + e;
  const int jblk = row / G, rr = row % G;
  const int kh = rr / CP, n = rr % CP;
  float v = 0.f;
  if (kh < 3 && n < Cout && kw < 3) {
    const int kd = 2 - jblk;
    v = w[((int64_t)n * Cin + ci) * 27 + (kd * 9 + kh * 3 + kw)];
    if (scale) v *= scale[n];
  }
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  img[i] = hi;
  img[per_term + i] = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// fp32 channels-last -> split-bf16 planes (API boundary / tests; producers on the hot path write the split form directly)
__global__ void split_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                  int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = x[i];
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi[i] = h;
  lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}

template <int CIN, int COUT, int KM, int NS>
static int launch(const void* x, const Params& p, cudaStream_t st) {
  using C = Cfg<CIN, COUT, KM, NS>;
  static_assert(C::SMEM <= 232448, "shared memory budget exceeded");
  auto kern = conv3d_slab_bf16_kernel<CIN, COUT, KM, NS>;
  static SmemAttrOnce smem_attr;
  {
    cudaError_t e = smem_attr.ensure(kern, (int)C::SMEM);
    RCMVS_REQUIRE(e == cudaSuccess, RCMVS_ERR_CUDA, "rcmvs_conv3d_slab_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  CUtensorMap map;
  const int rc = make_tmap(&map, x, p.B, p.D, p.H, p.W, CIN, C::ROWS);
  if (rc != RCMVS_OK) return rc;
  dim3 grid(ceil_div(p.W, TW), ceil_div(p.H, C::OROWS), p.B * p.chunks);
  RCMVS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, RCMVS_ERR_UNSUPPORTED, "rcmvs_conv3d_slab_fwd: grid too large");
  kern<<<grid, kThreads, C::SMEM, st>>>(map, p);
  RCMVS_CHECK_LAUNCH("rcmvs_conv3d_slab_fwd");
  return RCMVS_OK;
}

static bool supported(int Cin, int Cout) {
  return (Cout == 8 && (Cin == 8 || Cin == 16 || Cin == 32 || Cin == 48)) || (Cout == 16 && Cin == 16) || (Cout == 1 && Cin == 8);
}

// depth chunks + M-tiles per CTA: one CTA per SM (shared memory), so minimise
//   waves(tiles x chunks / 148) x slabs per CTA (planes + 2 halo slabs) x KM
static void plan(int B, int D, int H, int W, int Cin, int Cout, int& km, int& chunks) {
  long best = -1;
  const int km_lo = 2, km_hi = (Cout == 16 || Cin == 48) ? 2 : 3;        // shared-memory budget
  for (int k = km_lo; k <= km_hi; ++k) {
    const int tiles = ceil_div(W, TW) * ceil_div(H, 4 * k - 2) * B;
    for (int c = 1; c <= max(1, D / 4); ++c) {
      const int dc = ceil_div(D, c);
      const int cc = ceil_div(D, dc);
      const long cost = (long)ceil_div(tiles * cc, 148) * (dc + 2) * k;
      if (best < 0 || cost < best) { best = cost; km = k; chunks = cc; }
    }
  }
}

}  // namespace slab
}  // namespace rcmvs

extern "C" int64_t rcmvs_conv3d_slab_weight_elems(int Cin, int Cout) {
  const int cp = Cout < 8 ? 8 : Cout;
  const int g = ((3 * cp + 15) / 16) * 16;
  const int nstep = Cin == 8 ? 2 : 3 * (Cin / 16);
  return (int64_t)2 * nstep * 2 * 3 * g * 8;
}

extern "C" int rcmvs_conv3d_slab_pack(const float* w, const float* scale, void* img, int Cout, int Cin,
                                      rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(w && img, RCMVS_ERR_BAD_ARG, "rcmvs_conv3d_slab_pack: null pointer");
  RCMVS_REQUIRE(slab::supported(Cin, Cout), RCMVS_ERR_UNSUPPORTED, "rcmvs_conv3d_slab_pack: unsupported Cin=%d Cout=%d", Cin,
                Cout);
  const int cp = Cout < 8 ? 8 : Cout;
  const int g = ((3 * cp + 15) / 16) * 16;
  const int per_term = (Cin == 8 ? 2 : 3 * (Cin / 16)) * 2 * 3 * g * 8;
  slab::pack_slab_weight_kernel<<<ceil_div(per_term, 256), 256, 0, as_stream(stream)>>>(
      w, scale, reinterpret_cast<__nv_bfloat16*>(img), Cout, Cin, cp, g);
  RCMVS_CHECK_LAUNCH("rcmvs_conv3d_slab_pack");
  return RCMVS_OK;
}

extern "C" int rcmvs_split_bf16(const float* x, void* xs, int64_t n, rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(x && xs && n > 0, RCMVS_ERR_BAD_ARG, "rcmvs_split_bf16: bad arguments");
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(xs);
  slab::split_bf16_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, as_stream(stream)>>>(x, hi, hi + n, n);
  RCMVS_CHECK_LAUNCH("rcmvs_split_bf16");
  return RCMVS_OK;
}

// experiment switches for tools/slab_experiments.py (set through ctypes in_dll; not part of the ABI, 0 = product behaviour)
extern "C" int rcmvs_slab_debug_flags = 0;
extern "C" long long* rcmvs_slab_trace = nullptr;      // device buffer of 1024 int64 (tools/slab_experiments.py)

extern "C" int rcmvs_conv3d_slab_fwd(const void* x_split, const void* wimg, const float* bias, const float* skip,
                                     float* y, void* y_split, int B, int D, int H, int W, int Cin, int Cout, int relu,
                                     rcmvs_stream_t stream) {
  using namespace rcmvs;
  using namespace rcmvs::slab;
  RCMVS_REQUIRE(x_split && wimg && (y || y_split), RCMVS_ERR_BAD_ARG, "rcmvs_conv3d_slab_fwd: null pointer");
  RCMVS_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, RCMVS_ERR_BAD_ARG, "rcmvs_conv3d_slab_fwd: bad dims");
  RCMVS_REQUIRE(slab::supported(Cin, Cout), RCMVS_ERR_UNSUPPORTED, "rcmvs_conv3d_slab_fwd: unsupported Cin=%d Cout=%d", Cin,
                Cout);
  RCMVS_REQUIRE((reinterpret_cast<uintptr_t>(x_split) & 15) == 0, RCMVS_ERR_BAD_ARG,
                "rcmvs_conv3d_slab_fwd: x_split must be 16-byte aligned (TMA)");
  Params p;
  p.wimg = reinterpret_cast<const __nv_bfloat16*>(wimg);
  p.bias = bias; p.skip = skip; p.y = y; p.ys = reinterpret_cast<__nv_bfloat16*>(y_split);
  p.B = B; p.D = D; p.H = H; p.W = W; p.relu = relu; p.dbg = rcmvs_slab_debug_flags; p.trace = rcmvs_slab_trace;
  int km = 2, chunks = 1;
  plan(B, D, H, W, Cin, Cout, km, chunks);
  p.d_chunk = ceil_div(D, chunks);
  p.chunks = ceil_div(D, p.d_chunk);
  cudaStream_t st = as_stream(stream);
  if (Cin == 32 && Cout == 8) return km == 3 ? launch<32, 8, 3, 3>(x_split, p, st) : launch<32, 8, 2, 4>(x_split, p, st);
  if (Cin == 8 && Cout == 1) return km == 3 ? launch<8, 1, 3, 8>(x_split, p, st) : launch<8, 1, 2, 8>(x_split, p, st);   // `prob`
  if (Cin == 8 && Cout == 8) return km == 3 ? launch<8, 8, 3, 8>(x_split, p, st) : launch<8, 8, 2, 8>(x_split, p, st);   // conv0, stage 3
  if (Cin == 16 && Cout == 8) return km == 3 ? launch<16, 8, 3, 5>(x_split, p, st) : launch<16, 8, 2, 6>(x_split, p, st);
  if (Cin == 16 && Cout == 16) return launch<16, 16, 2, 5>(x_split, p, st);
  if (Cin == 48 && Cout == 8) return launch<48, 8, 2, 3>(x_split, p, st);      // Neural_Volume_Net conv0 (41 -> 48 padded)
  return RCMVS_ERR_UNSUPPORTED;
}
