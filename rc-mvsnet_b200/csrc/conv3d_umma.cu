// a3 CostRegNet stride-1 3x3x3 convolutions on the 5th-gen tensor cores (tcgen05.mma kind::tf32, accumulators in
// TMEM), reference models/modules.py:118-163 (Conv3d = conv -> BN -> ReLU) and :490 (`prob`).
//
// v1 of this kernel (profiles/r01_conv3d_umma_v1_ncu.md) issued one M128xN16xK8 MMA per (tap, K-step): with N=16 every
// MMA re-fetches its 4 KB A tile from shared memory, 27 x Cin/8 times per output tile, and the 128 B/clk shared-memory
// operand path (not the tensor pipe, not HBM) bounded the kernel.  v2 folds the kd and kh taps into the N dimension:
//
//   for one INPUT depth slab z and one kw:   E_z[m, (kd, kh, cout)] += Σ_cin x_z[m + kw, cin] · w[kd, kh, kw, cin, cout]
//
//   M = 128 input positions = a flat run of a zero-padded (rows x 32) plane (kw = +1 row of the run = +16 bytes of
//       descriptor start address);  N = 3 (kd) x G, G = 3 (kh) x Cout rounded up to 16;  K = 8 channels per MMA.
//   The three kd blocks of N are the accumulators of three different OUTPUT planes (z+1, z, z-1): TMEM holds a ring of
//   per-plane column groups, so the sum over kd happens inside TMEM for free.  The sum over kh (contributions of input
//   row r to output rows r, r-1, r-2) is done by the epilogue through a small shared-memory exchange.
//   A is fetched 3 (kw) x Cin/8 times per slab instead of 27 x Cin/8 -> 9x less operand traffic; N grows to 96..144.
//
// Shared-memory operand layout = UMMA canonical K-major, SWIZZLE_NONE ("core matrix" = 8 rows x 16 bytes):
//   A slab : [Cin/4 chunks][slab voxels][4 floats]   row pitch 16 B, SBO = 128 B, LBO = plane size
//   B image: [term][kw][K-step][2 chunks][3 x G rows][4 floats], pre-packed in global memory, copied once per CTA
// cp.async (16 B, zero-fill = the conv's zero padding) fills the slabs, double/triple buffered along d.
//
// Precision: NSPLIT=1 plain TF32 (what cuDNN gives the reference on GPU under torch defaults);
//            NSPLIT=3 error-compensated 3xTF32: x = x_hi + x_lo, w = w_hi + w_lo, acc += x_lo·w_hi + x_hi·w_lo + x_hi·w_hi.
//
// Warp roles: warps 0-3 epilogue, 2 loader warps per slab buffer (cp.async + hi/lo split), last warp MMA issuer.
#include <stdlib.h>

#include "umma.cuh"

namespace rcmvs {
namespace umma {

// Stage-ablation mask for profiling (skip MMA issue / loads / stores ...): only in builds made with
// RCMVS_BUILD_DEBUG_HOOKS=1 (-DRCMVS_TC_DEBUG_HOOKS); the shipped library ignores the environment.
static inline int tc_debug_mask() {
#ifdef RCMVS_TC_DEBUG_HOOKS
  const char* s = getenv("RCMVS_TC_DBG");
  return s ? atoi(s) : 0;
#else
  return 0;
#endif
}

// COUT real output channels; G columns per (plane, M-tile) group = 3 kh blocks of COUT, rounded up to 16.
template <int CIN, int COUT, int NSPLIT, int KM, int NS, int NGRP = 1>
struct Cfg {
  static constexpr int CP = COUT < 8 ? 8 : COUT;                  // channel slots per kh block (Cout=1 -> 8)
  static constexpr int G = ((3 * CP + 15) / 16) * 16;             // TMEM columns / B rows per plane block
  static constexpr int KS = CIN / 8;
  static constexpr int NCH = CIN / 4;
  static constexpr int ROWS = 4 * KM;                             // input rows covered by the M-tiles
  static constexpr int OROWS = ROWS - 2;                          // output rows per CTA tile
  // voxels per slab plane (+ kw overrun) + 8/NCH voxels of padding: the channel-chunk planes then start 128/NCH bytes
  // apart modulo 128, so the NCH chunks x (8/NCH) voxels a quarter-warp of the loader stores with one 128-bit store
  // instruction fall into distinct banks (a plane stride of 0 mod 128 made them collide NCH-fold)
  static constexpr int SLABV = ROWS * PW + 8 + 8 / NCH;
  static constexpr int PLANE_B = SLABV * 16;
  static constexpr int SLAB_B = NCH * PLANE_B;
  static constexpr int NPART = NSPLIT == 3 ? 2 : 1;
  static constexpr int A_BYTES = NS * NPART * SLAB_B;
  static constexpr int BBLK_B = 2 * 3 * G * 16;                   // one (kw, K-step): 2 chunks x 3G rows x 16 B
  static constexpr int BTERM_B = 3 * KS * BBLK_B;
  static constexpr int BGRP_B = NPART * BTERM_B;                  // weight image of one input-channel group
  static constexpr int B_BYTES = NGRP * BGRP_B;
  static constexpr int NG = 512 / (KM * G);                       // plane groups in the TMEM ring
  static constexpr int EX_B = 3 * OROWS * PW * CP * 4;            // epilogue exchange buffer: one plane per kh
  static constexpr size_t SMEM = (size_t)A_BYTES + B_BYTES + EX_B + 512;
  static_assert(NG >= 4, "TMEM ring needs at least 4 plane groups");
};

struct Params {
  const float* x;       // [B][D][H][W][cin_total]
  const float* wimg;    // packed B image for this channel group
  const float* bias;    // [Cout] or null
  const float* skip;    // [B][D][H][W][Cout] or null
  float* y;             // [B][D][H][W][Cout]
  int D, H, W, cin_total, cin_off;
  int d_chunk, chunks;  // depth planes per CTA, chunks per batch item
  int accumulate_in;    // add the existing contents of y (partial sums of an earlier channel group)
  int finalize;         // apply bias / ReLU / skip (last channel group)
  int relu;
  int dbg;              // development only (RCMVS_TC_DBG): 1 = skip MMA issue, 2 = skip epilogue math/stores, 4 = skip loads
};

// NGRP > 1: the CTA walks NGRP input-channel groups of CIN channels per depth slab (each group is its own slab in
// the ring and has its own weight image); all groups accumulate into the same TMEM plane groups, so a 32-channel
// layer in 3xTF32 mode needs one launch and one epilogue instead of two launches chained through global memory.
template <int CIN, int COUT, int NSPLIT, int KM, int NS, int NGRP>
__global__ void __launch_bounds__(threads_for(NS), 1) conv3d_umma_kernel(const Params p) {
  using C = Cfg<CIN, COUT, NSPLIT, KM, NS, NGRP>;
  constexpr int G = C::G, NG = C::NG, CP = C::CP;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* a_smem = smem;
  unsigned char* b_smem = smem + C::A_BYTES;
  float* ex = reinterpret_cast<float*>(smem + C::A_BYTES + C::B_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::A_BYTES + C::B_BYTES + C::EX_B);
  uint64_t* full = bars;                 // [NS]  loaders -> MMA
  uint64_t* empty = bars + NS;           // [NS]  MMA -> loaders
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
  const int nstream = nslabs * NGRP;       // (slab, channel group) items flowing through the ring
  const int64_t vox_stride = p.cin_total;
  const float* xb = p.x + (int64_t)b * p.D * p.H * p.W * vox_stride + p.cin_off;

  // ---- one-time setup -------------------------------------------------------------------------------------------
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(full + i, kLoaderThreads); mbar_init(empty + i, 1); }
    for (int i = 0; i < NG; ++i) { mbar_init(grp_full + i, 1); mbar_init(grp_free + i, 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {
    const float4* src = reinterpret_cast<const float4*>(p.wimg);
    float4* dst = reinterpret_cast<float4*>(b_smem);
    for (int i = threadIdx.x; i < C::B_BYTES / 16; i += threads_for(NS)) dst[i] = __ldg(src + i);
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

  if (warp >= 4 && warp < 4 + kWPB * NS) {
    // =============================== slab loaders ================================================================
    // kWPB warps own one ring buffer each for the whole kernel: wait until the MMA warp released it, fill it with
    // cp.async, wait for the data, (split hi/lo), publish.  The NS buffers run independent load->MMA->free cycles,
    // so global-memory latency of one slab overlaps the MMAs of the others.
    const int mybuf = (warp - 4) / kWPB;
    const int t = (warp - 4) % kWPB * 32 + lane;
    constexpr int PIECES = C::ROWS * PW * C::NCH;
    const uint32_t slab = smem_u32(a_smem + (size_t)mybuf * C::NPART * C::SLAB_B);
    for (int ii = mybuf; ii < nstream; ii += NS) {
      mbar_wait(empty + mybuf, ((ii / NS) & 1) ^ 1);
      const int i = ii / NGRP, grp = ii % NGRP;
      const int z = d0 - 1 + i;
      const bool zok = (z >= 0) && (z < p.D);
      const float* xg = xb + grp * CIN;                          // this group's first channel
      if constexpr (NSPLIT == 3) {
        // 3xTF32: plain loads into registers, split x = hi + lo (hi = x & 0xFFFFE000, exactly representable in TF32)
        // on the way to shared memory — one pass instead of cp.async + re-read + two stores.  8 loads in flight per thread.
        constexpr int VSTEP = kLoaderThreads / C::NCH;
        static_assert(PW % VSTEP == 0, "loader mapping");
        constexpr int CPR = PW / VSTEP;                           // pieces per row per thread
        constexpr int NP = C::ROWS * CPR;                         // pieces per thread
        constexpr int BATCH = (NP % 8 == 0) ? 8 : (NP % 6 == 0 ? 6 : 4);
        static_assert(NP % BATCH == 0, "loader batch");
        const int ch = t % C::NCH, v0 = t / C::NCH;
        const float* zbase = xg + (int64_t)z * p.H * p.W * vox_stride + ch * 4;
        unsigned char* hi_p = a_smem + (size_t)mybuf * C::NPART * C::SLAB_B + ch * C::PLANE_B;
        unsigned char* lo_p = hi_p + C::SLAB_B;
        if (!(p.dbg & 4)) {
          for (int p0 = 0; p0 < NP; p0 += BATCH) {
            float4 v[BATCH];
#pragma unroll
            for (int u = 0; u < BATCH; ++u) {
              const int pi = p0 + u;
              const int row = pi / CPR, col = v0 + (pi % CPR) * VSTEP;
              const int gy = y0 - 1 + row, gx = x0 - 1 + col;
              const bool ok = pi < NP && zok && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
              v[u] = ok ? ldg4(zbase + ((int64_t)gy * p.W + gx) * vox_stride) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < BATCH; ++u) {
              const int pi = p0 + u;
              if (pi >= NP) break;
              const int row = pi / CPR, col = v0 + (pi % CPR) * VSTEP;
              const int off = (row * PW + col) * 16;
              float4 h, l;
              h.x = __uint_as_float(__float_as_uint(v[u].x) & 0xFFFFE000u); l.x = v[u].x - h.x;
              h.y = __uint_as_float(__float_as_uint(v[u].y) & 0xFFFFE000u); l.y = v[u].y - h.y;
              h.z = __uint_as_float(__float_as_uint(v[u].z) & 0xFFFFE000u); l.z = v[u].z - h.z;
              h.w = __uint_as_float(__float_as_uint(v[u].w) & 0xFFFFE000u); l.w = v[u].w - h.w;
              *reinterpret_cast<float4*>(hi_p + off) = h;
              *reinterpret_cast<float4*>(lo_p + off) = l;
            }
          }
        }
      } else {
        if (!(p.dbg & 4)) {
          constexpr int VSTEP = kLoaderThreads / C::NCH;          // voxels covered by the buffer's threads per pass
          static_assert(PW % VSTEP == 0, "loader mapping");
          const int ch = t % C::NCH, v0 = t / C::NCH;
          const float* zbase = xg + (int64_t)z * p.H * p.W * vox_stride + ch * 4;
          const uint32_t dst0 = slab + ch * C::PLANE_B;
          for (int row = 0; row < C::ROWS; ++row) {
            const int gy = y0 - 1 + row;
            const bool yok = zok && gy >= 0 && gy < p.H;
            const float* rbase = zbase + (int64_t)gy * p.W * vox_stride;
#pragma unroll
            for (int col = v0; col < PW; col += VSTEP) {
              const int gx = x0 - 1 + col;
              const bool ok = yok && gx >= 0 && gx < p.W;
              const float* src = ok ? rbase + (int64_t)gx * vox_stride : xb;
              cp_async16(dst0 + (row * PW + col) * 16, src, ok);
            }
          }
        }
        cp_async_commit();
        cp_async_wait<0>();
      }
      fence_proxy_async();
      mbar_arrive(full + mybuf);
    }
  } else if (warp == 4 + kWPB * NS) {
    // =============================== MMA issuer ==================================================================
    // The whole warp runs the (warp-uniform) control flow so descriptor arithmetic stays on the uniform datapath;
    // one elected lane issues tcgen05.mma / tcgen05.commit.
    const bool leader = elect_one();
    const uint32_t b_base = smem_u32(b_smem);
    for (int ii = 0; ii < nstream; ++ii) {
      const int i = ii / NGRP, grp = ii % NGRP;
      // slab i = input plane z = d0-1+i contributes to output planes (relative) q = i-2+j, block j = 0,1,2 (kd = 2-j)
      const int jlo = max(0, 2 - i), jhi = min(2, nplanes + 1 - i);       // valid blocks: 0 <= q < nplanes
      const int qnew = i;                                                  // plane first touched by this slab (j=2)
      if (grp == 0 && qnew < nplanes && qnew >= NG) {                      // its group must have been drained
        mbar_wait(grp_free + (qnew % NG), ((qnew / NG) - 1) & 1);
      }
      const int buf = ii % NS;
      mbar_wait(full + buf, (ii / NS) & 1);
      tc_fence_after();
      const uint32_t a_base = smem_u32(a_smem + (size_t)buf * C::NPART * C::SLAB_B);
      // split the valid block range into (at most two) ring-contiguous runs; per run: TMEM column, B row offset, N
      const int g0 = (i - 2 + jlo + NG) % NG;
      const int n_all = jhi - jlo + 1;
      const int n0 = min(n_all, NG - g0), n1 = n_all - n0;                 // second run wraps to group 0
      const uint32_t d0c = tmem_base + g0 * G, d1c = tmem_base;
      const uint32_t idesc0 = make_idesc(128, n0 * G), idesc1 = make_idesc(128, (n1 > 0 ? n1 : 1) * G);
      const uint32_t boff0 = (uint32_t)(jlo * G), boff1 = (uint32_t)((jlo + n0) * G);   // in 16-byte rows
      const uint32_t a_lo0 = desc_lo(a_base, C::PLANE_B);                  // hi-part slab
      const uint32_t b_lo0 = desc_lo(b_base + grp * C::BGRP_B, 3 * G * 16);
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
        for (int ks = 0; ks < C::KS; ++ks) {
#pragma unroll
          for (int mt = 0; mt < KM; ++mt) {
            constexpr uint32_t kLoOff = C::SLAB_B >> 4, kBLoOff = C::BTERM_B >> 4;
            const uint32_t a_hi_d = a_lo0 + (uint32_t)((ks * 2 * C::PLANE_B) >> 4) + (uint32_t)(mt * 128 + kw);
            const uint32_t b_blk_d = b_lo0 + (uint32_t)(((kw * C::KS + ks) * C::BBLK_B) >> 4);
            const uint32_t dm = (uint32_t)(mt * (NG * G));
            if (leader && !(p.dbg & 1)) {
              if constexpr (NSPLIT == 3) {
                umma_tf32_lohi(d0c + dm, a_hi_d + kLoOff, b_blk_d + boff0, idesc0);            // x_lo · w_hi
                umma_tf32_lohi(d0c + dm, a_hi_d, b_blk_d + boff0 + kBLoOff, idesc0);           // x_hi · w_lo
                umma_tf32_lohi(d0c + dm, a_hi_d, b_blk_d + boff0, idesc0);                     // x_hi · w_hi
                if (n1 > 0) {
                  umma_tf32_lohi(d1c + dm, a_hi_d + kLoOff, b_blk_d + boff1, idesc1);
                  umma_tf32_lohi(d1c + dm, a_hi_d, b_blk_d + boff1 + kBLoOff, idesc1);
                  umma_tf32_lohi(d1c + dm, a_hi_d, b_blk_d + boff1, idesc1);
                }
              } else {
                umma_tf32_lohi(d0c + dm, a_hi_d, b_blk_d + boff0, idesc0);
                if (n1 > 0) umma_tf32_lohi(d1c + dm, a_hi_d, b_blk_d + boff1, idesc1);
              }
            }
          }
        }
      }
      if (leader) {
        umma_commit(empty + buf);                                   // slab consumed once these MMAs retire
        if (i >= 2 && grp == NGRP - 1) umma_commit(grp_full + ((i - 2) % NG));   // output plane q = i-2 is complete
      }
      __syncwarp();
    }
  } else if (warp < 4) {
    // =============================== epilogue (128 threads) ======================================================
    const int gx = x0 + lane;
    const bool col_ok = lane < TW && gx < p.W;
    for (int q = 0; q < nplanes; ++q) {
      const int g = q % NG;
      mbar_wait(grp_full + g, (q / NG) & 1);
      tc_fence_after();
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
      if (p.dbg & 2) continue;
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
          if (p.accumulate_in) r += p.y[o];
          if (p.finalize) {
            if (p.bias) r += __ldg(p.bias);
            if (p.relu) r = fmaxf(r, 0.f);
            if (p.skip) r += __ldg(p.skip + o);
          }
          p.y[o] = r;
        } else {
#pragma unroll
          for (int c = 0; c < COUT; c += 4) {
            float4 r = *reinterpret_cast<const float4*>(e + (c / 4) * PW * 4);
            {
              const float4 r1 = *reinterpret_cast<const float4*>(e + EXP + (c / 4) * PW * 4);
              const float4 r2 = *reinterpret_cast<const float4*>(e + 2 * EXP + (c / 4) * PW * 4);
              r.x += r1.x + r2.x; r.y += r1.y + r2.y; r.z += r1.z + r2.z; r.w += r1.w + r2.w;
            }
            if (p.accumulate_in) {
              const float4 t4 = *reinterpret_cast<const float4*>(p.y + o + c);
              r.x += t4.x; r.y += t4.y; r.z += t4.z; r.w += t4.w;
            }
            if (p.finalize) {
              if (p.bias) {
                const float4 t4 = ldg4(p.bias + c);
                r.x += t4.x; r.y += t4.y; r.z += t4.z; r.w += t4.w;
              }
              if (p.relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
              if (p.skip) {
                const float4 t4 = ldg4(p.skip + o + c);
                r.x += t4.x; r.y += t4.y; r.z += t4.z; r.w += t4.w;
              }
            }
            st4(p.y + o + c, r);
          }
        }
      }
      epi_bar_sync();                              // exchange buffer is reused by the next plane
    }
  }

  // ---- teardown -------------------------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// Weight image: torch Conv3d weight [Cout][Cin_total][27] x BN scale, for input channels [cin_off, cin_off+CIN):
//   img[term][kw][ks][chunk][row = jblk*G + kh*CP + n][e],  kd = 2 - jblk,
//   value = split_term( w[n][cin_off + ks*8 + chunk*4 + e][kd][kh][kw] * scale[n] ), zero for padded rows
__global__ void pack_umma_weight_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                        float* __restrict__ img, int Cout, int cin_total, int cin_off, int CIN,
                                        int CP, int G, int nsplit) {
  const int KS = CIN / 8;
  const int per_term = 3 * KS * 2 * 3 * G * 4;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per_term) return;
  const int e = i & 3;
  int r = i >> 2;
  const int row = r % (3 * G); r /= (3 * G);
  const int chunk = r & 1; r >>= 1;
  const int ks = r % KS;
  const int kw = r / KS;
  const int jblk = row / G, rr = row % G;
  const int kh = rr / CP, n = rr % CP;
  float v = 0.f;
  if (kh < 3 && n < Cout) {
    const int kd = 2 - jblk;
    const int ci = cin_off + ks * 8 + chunk * 4 + e;
    v = w[((int64_t)n * cin_total + ci) * 27 + (kd * 9 + kh * 3 + kw)];
    if (scale) v *= scale[n];
  }
  if (nsplit == 3) {
    const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    img[i] = hi;
    img[per_term + i] = v - hi;
  } else {
    img[i] = v;
  }
}

template <int CIN, int COUT, int NSPLIT, int KM, int NS, int NGRP = 1>
static int launch(const Params& p, int B, cudaStream_t st) {
  using C = Cfg<CIN, COUT, NSPLIT, KM, NS, NGRP>;
  static_assert(C::SMEM <= 232448, "shared memory budget exceeded");
  auto kern = conv3d_umma_kernel<CIN, COUT, NSPLIT, KM, NS, NGRP>;
  static SmemAttrOnce smem_attr;
  {
    cudaError_t e = smem_attr.ensure(kern, (int)C::SMEM);
    RCMVS_REQUIRE(e == cudaSuccess, RCMVS_ERR_CUDA, "rcmvs_conv3d_tc_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  dim3 grid(ceil_div(p.W, TW), ceil_div(p.H, C::OROWS), B * p.chunks);
  RCMVS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, RCMVS_ERR_UNSUPPORTED, "rcmvs_conv3d_tc_fwd: grid too large");
  kern<<<grid, threads_for(NS), C::SMEM, st>>>(p);
  RCMVS_CHECK_LAUNCH("rcmvs_conv3d_tc_fwd");
  return RCMVS_OK;
}

// (Cin, Cout) pairs with an instantiated kernel: conv0 of every stage, conv2, prob
static bool supported(int Cin, int Cout) {
  return (Cout == 8 && (Cin == 8 || Cin == 16 || Cin == 32)) || (Cout == 1 && Cin == 8) || (Cout == 16 && Cin == 16);
}

struct Plan {
  int cin_k, km, cp, g;
};
static Plan plan(int Cin, int Cout, int nsplit) {
  Plan pl;
  pl.cp = Cout < 8 ? 8 : Cout;
  pl.g = ((3 * pl.cp + 15) / 16) * 16;
  pl.cin_k = (nsplit == 3 && Cin == 32) ? 16 : Cin;     // hi+lo slabs of 32 channels do not fit: two channel groups,
  pl.km = (Cout == 16 || (nsplit == 3 && Cin == 32)) ? 2 : 3;   // ... walked inside one launch (NGRP = 2)
  return pl;
}


// ==================================================================================================================
// ConvTranspose3d(k=3, stride=2, padding=1, output_padding=1) on tcgen05 (reference models/modules.py:166-210 Deconv3d;
// conv7/9/11 of CostRegNet :484-488).  out[2i+a] per dimension: a=0 <- (k=1, i);  a=1 <- (k=2, i) + (k=0, i+1).
// Every output parity class (a,b,c) is a small stride-1 convolution over the INPUT grid, so the same flat-run trick
// applies in input space: M = 128 input positions, the 8 input shifts (dz,dy,dx) in {0,1}^3 are descriptor start
// offsets (+dy*32+dx voxels, slab iz+dz), and N = 8 classes x Cout holds all parity classes at once (weights of
// classes that do not use a shift are zero rows of B).  Accumulators: double-buffered TMEM stages; the epilogue
// scatters the 8 classes of every input voxel to its 2x2x2 output block with bias / ReLU / skip fused.
// ==================================================================================================================
template <int CIN, int COUT, int NSPLIT, int KM, int NS>
struct TCfg {
  static constexpr int N = 8 * COUT;
  static constexpr int KS = CIN / 8;
  static constexpr int NCH = CIN / 4;
  static constexpr int ROWS = 4 * KM;                              // input rows per CTA tile (all produce outputs)
  static constexpr int TWI = 31;                                   // input columns per tile (col 31 = halo for dx=1)
  static constexpr int SLABV = (ROWS + 1) * PW + 8 + 8 / NCH;     // + bank-spreading pad, see Cfg
  static constexpr int PLANE_B = SLABV * 16;
  static constexpr int SLAB_B = NCH * PLANE_B;
  static constexpr int NPART = NSPLIT == 3 ? 2 : 1;
  static constexpr int A_BYTES = NS * NPART * SLAB_B;
  static constexpr int BBLK_B = 2 * N * 16;                        // one (shift, K-step)
  static constexpr int BTERM_B = 8 * KS * BBLK_B;
  static constexpr int B_BYTES = NPART * BTERM_B;
  static constexpr int TMEM_COLS = 2 * KM * N <= 256 ? 256 : 512;
  static constexpr size_t SMEM = (size_t)A_BYTES + B_BYTES + 512;
  static_assert(2 * KM * N <= 512, "TMEM budget exceeded");
};

template <int CIN, int COUT, int NSPLIT, int KM, int NS>
__global__ void __launch_bounds__(threads_for(NS), 1) convT3d_umma_kernel(const Params p) {
  using C = TCfg<CIN, COUT, NSPLIT, KM, NS>;
  constexpr int N = C::N;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* a_smem = smem;
  unsigned char* b_smem = smem + C::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::A_BYTES + C::B_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + NS;
  uint64_t* acc_full = bars + 2 * NS;
  uint64_t* acc_empty = bars + 2 * NS + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x0 = blockIdx.x * C::TWI, y0 = blockIdx.y * C::ROWS;
  const int b = blockIdx.z / p.chunks;
  const int z0 = (blockIdx.z % p.chunks) * p.d_chunk;
  const int z1 = min(p.D, z0 + p.d_chunk);
  const int nsteps = z1 - z0;            // input planes z0..z1-1 ; slabs z0 .. z1  (nsteps + 1)
  const int nslabs = nsteps + 1;
  const int64_t vox_stride = p.cin_total;
  const float* xb = p.x + (int64_t)b * p.D * p.H * p.W * vox_stride + p.cin_off;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(full + i, kLoaderThreads); mbar_init(empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {
    const float4* src = reinterpret_cast<const float4*>(p.wimg);
    float4* dst = reinterpret_cast<float4*>(b_smem);
    for (int i = threadIdx.x; i < C::B_BYTES / 16; i += threads_for(NS)) dst[i] = __ldg(src + i);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 4 && warp < 4 + kWPB * NS) {
    // =============================== slab loaders (kWPB warps own one ring buffer each) ============================
    const int mybuf = (warp - 4) / kWPB;
    const int t = (warp - 4) % kWPB * 32 + lane;
    constexpr int PIECES = (C::ROWS + 1) * PW * C::NCH;
    const uint32_t slab = smem_u32(a_smem + (size_t)mybuf * C::NPART * C::SLAB_B);
    for (int i = mybuf; i < nslabs; i += NS) {
      mbar_wait(empty + mybuf, ((i / NS) & 1) ^ 1);
      const int z = z0 + i;
      const bool zok = z < p.D;
      if constexpr (NSPLIT == 3) {
        // 3xTF32: plain loads into registers, split x = hi + lo on the way to shared memory (one pass instead of
        // cp.async + re-read + two stores); BATCH loads in flight per thread
        static_assert(PIECES % kLoaderThreads == 0, "loader mapping");
        constexpr int NP = PIECES / kLoaderThreads;
        constexpr int BATCH = NP % 6 == 0 ? 6 : (NP % 4 == 0 ? 4 : (NP % 3 == 0 ? 3 : 1));
        unsigned char* hi_p = a_smem + (size_t)mybuf * C::NPART * C::SLAB_B;
        unsigned char* lo_p = hi_p + C::SLAB_B;
        for (int p0 = 0; p0 < NP; p0 += BATCH) {
          float4 v[BATCH];
#pragma unroll
          for (int u = 0; u < BATCH; ++u) {
            const int e = t + (p0 + u) * kLoaderThreads;
            const int ch = e % C::NCH, vv = e / C::NCH;
            const int gy = y0 + vv / PW, gx = x0 + (vv % PW);
            const bool ok = zok && gy < p.H && gx < p.W;
            v[u] = ok ? ldg4(xb + (((int64_t)z * p.H + gy) * p.W + gx) * vox_stride + ch * 4)
                      : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < BATCH; ++u) {
            const int e = t + (p0 + u) * kLoaderThreads;
            const int off = (e % C::NCH) * C::PLANE_B + (e / C::NCH) * 16;
            float4 h, l;
            h.x = __uint_as_float(__float_as_uint(v[u].x) & 0xFFFFE000u); l.x = v[u].x - h.x;
            h.y = __uint_as_float(__float_as_uint(v[u].y) & 0xFFFFE000u); l.y = v[u].y - h.y;
            h.z = __uint_as_float(__float_as_uint(v[u].z) & 0xFFFFE000u); l.z = v[u].z - h.z;
            h.w = __uint_as_float(__float_as_uint(v[u].w) & 0xFFFFE000u); l.w = v[u].w - h.w;
            *reinterpret_cast<float4*>(hi_p + off) = h;
            *reinterpret_cast<float4*>(lo_p + off) = l;
          }
        }
      } else {
        for (int e = t; e < PIECES; e += kLoaderThreads) {
          const int ch = e % C::NCH;
          const int v = e / C::NCH;
          const int gy = y0 + v / PW, gx = x0 + (v % PW);
          const bool ok = zok && gy < p.H && gx < p.W;
          const float* src = ok ? xb + (((int64_t)z * p.H + gy) * p.W + gx) * vox_stride + ch * 4 : xb;
          cp_async16(slab + ch * C::PLANE_B + v * 16, src, ok);
        }
        cp_async_commit();
        cp_async_wait<0>();
      }
      fence_proxy_async();
      mbar_arrive(full + mybuf);
    }
  } else if (warp == 4 + kWPB * NS) {
    // =============================== MMA issuer (warp-uniform, one elected lane issues) ============================
    const bool leader = elect_one();
    const uint32_t b_lo0 = desc_lo(smem_u32(b_smem), N * 16);
    constexpr uint32_t idesc = make_idesc(128, N);
    for (int s = 0; s < nsteps; ++s) {
      const int stage = s & 1;
      mbar_wait(acc_empty + stage, ((s >> 1) & 1) ^ 1);
      tc_fence_after();
#pragma unroll
      for (int dz = 0; dz < 2; ++dz) {
        const int i = s + dz;
        const int buf = i % NS;
        mbar_wait(full + buf, (i / NS) & 1);
        tc_fence_after();
        const uint32_t a_lo0 = desc_lo(smem_u32(a_smem + (size_t)buf * C::NPART * C::SLAB_B), C::PLANE_B);
#pragma unroll
        for (int dyx = 0; dyx < 4; ++dyx) {
          const int sh = dz * 4 + dyx;                                   // shift index (dz,dy,dx)
#pragma unroll
          for (int ks = 0; ks < C::KS; ++ks) {
#pragma unroll
            for (int mt = 0; mt < KM; ++mt) {
              constexpr uint32_t kLoOff = C::SLAB_B >> 4, kBLoOff = C::BTERM_B >> 4;
              const uint32_t a_d = a_lo0 + (uint32_t)((ks * 2 * C::PLANE_B) >> 4) +
                                   (uint32_t)(mt * 128 + (dyx >> 1) * PW + (dyx & 1));
              const uint32_t b_d = b_lo0 + (uint32_t)(((sh * C::KS + ks) * C::BBLK_B) >> 4);
              const uint32_t d = tmem_base + (uint32_t)(stage * KM * N + mt * N);
              const bool first = (dz == 0 && dyx == 0 && ks == 0);
              if (leader) {
                if constexpr (NSPLIT == 3) {
                  umma_tf32_lohi_acc(d, a_d + kLoOff, b_d, idesc, first ? 0u : 1u);
                  umma_tf32_lohi_acc(d, a_d, b_d + kBLoOff, idesc, 1u);
                  umma_tf32_lohi_acc(d, a_d, b_d, idesc, 1u);
                } else {
                  umma_tf32_lohi_acc(d, a_d, b_d, idesc, first ? 0u : 1u);
                }
              }
            }
          }
        }
        if (dz == 0 && leader) umma_commit(empty + buf);                 // slab iz is not needed by later steps
      }
      if (leader) umma_commit(acc_full + stage);
      __syncwarp();
    }
  } else if (warp < 4) {
    // =============================== epilogue ======================================================================
    const int gx = x0 + lane;
    const bool col_ok = lane < C::TWI && gx < p.W;
    const int Do = 2 * p.D, Ho = 2 * p.H, Wo = 2 * p.W;
    // The skip tensor is as large as the output and every thread needs 8 x COUT of it per M-tile: with one CTA per SM
    // the 4 epilogue warps cannot hide that latency, so (COUT = 8) the values are requested BEFORE the wait for the
    // accumulators / right after the previous M-tile's stores instead of one load per class in front of its add.
    constexpr bool kPrefetchSkip = COUT == 8;
    float4 sk[kPrefetchSkip ? 16 : 1];
    const bool use_skip = p.skip != nullptr && p.finalize;
    auto load_skip = [&](int iz, int mt) {
      if constexpr (kPrefetchSkip) {
        const int gy = y0 + mt * 4 + warp;
        const bool ok = use_skip && col_ok && gy < p.H && iz < p.D;
#pragma unroll
        for (int cls = 0; cls < 8; ++cls) {
          const int oz = 2 * iz + (cls >> 2), oy = 2 * gy + ((cls >> 1) & 1), ox = 2 * gx + (cls & 1);
          const int64_t o = ((((int64_t)b * Do + oz) * Ho + oy) * Wo + ox) * COUT;
          sk[cls * 2] = ok ? ldg4(p.skip + o) : make_float4(0.f, 0.f, 0.f, 0.f);
          sk[cls * 2 + 1] = ok ? ldg4(p.skip + o + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
    for (int s = 0; s < nsteps; ++s) {
      const int stage = s & 1;
      const int iz = z0 + s;
      load_skip(iz, 0);
      mbar_wait(acc_full + stage, (s >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int mt = 0; mt < KM; ++mt) {
        float v[N];
#pragma unroll
        for (int c0 = 0; c0 < N; c0 += 16)
          tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + stage * KM * N + mt * N + c0, &v[c0]);
        if (mt == KM - 1) {
          tc_fence_before();
          mbar_arrive(acc_empty + stage);
        }
        const int gy = y0 + mt * 4 + warp;
        if (!col_ok || gy >= p.H) {
          if (mt + 1 < KM) load_skip(iz, mt + 1);
          continue;
        }
#pragma unroll
        for (int cls = 0; cls < 8; ++cls) {
          const int oz = 2 * iz + (cls >> 2), oy = 2 * gy + ((cls >> 1) & 1), ox = 2 * gx + (cls & 1);
          const int64_t o = ((((int64_t)b * Do + oz) * Ho + oy) * Wo + ox) * COUT;
#pragma unroll
          for (int c = 0; c < COUT; c += 4) {
            float4 r = make_float4(v[cls * COUT + c], v[cls * COUT + c + 1], v[cls * COUT + c + 2], v[cls * COUT + c + 3]);
            if (p.accumulate_in) {
              const float4 q = *reinterpret_cast<const float4*>(p.y + o + c);
              r.x += q.x; r.y += q.y; r.z += q.z; r.w += q.w;
            }
            if (p.finalize) {
              if (p.bias) {
                const float4 q = ldg4(p.bias + c);
                r.x += q.x; r.y += q.y; r.z += q.z; r.w += q.w;
              }
              if (p.relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
              if (p.skip) {
                const float4 q = kPrefetchSkip ? sk[kPrefetchSkip ? cls * 2 + c / 4 : 0] : ldg4(p.skip + o + c);
                r.x += q.x; r.y += q.y; r.z += q.z; r.w += q.w;
              }
            }
            st4(p.y + o + c, r);
          }
        }
        if (mt + 1 < KM) load_skip(iz, mt + 1);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
  }
}

// Weight image of the transposed convolution: torch ConvTranspose3d weight [Cin_total][Cout][27] x BN scale ->
//   img[term][shift = dz*4+dy*2+dx][ks][chunk][row = cls*COUT + n][e]; per dimension the tap used by output parity a
//   with input shift d is k = (a==0 ? (d==0 ? 1 : none) : (d==0 ? 2 : 0)).
__global__ void pack_umma_tweight_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                         float* __restrict__ img, int Cout, int cin_off, int CIN, int nsplit) {
  const int KS = CIN / 8, N = 8 * Cout;
  const int per_term = 8 * KS * 2 * N * 4;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per_term) return;
  const int e = i & 3;
  int r = i >> 2;
  const int row = r % N; r /= N;
  const int chunk = r & 1; r >>= 1;
  const int ks = r % KS;
  const int sh = r / KS;
  const int cls = row / Cout, n = row % Cout;
  const int par[3] = {cls >> 2, (cls >> 1) & 1, cls & 1};
  const int dsh[3] = {sh >> 2, (sh >> 1) & 1, sh & 1};
  int k[3];
  bool used = true;
  for (int a = 0; a < 3; ++a) {
    if (par[a] == 0) { if (dsh[a] == 0) k[a] = 1; else used = false; }
    else k[a] = dsh[a] == 0 ? 2 : 0;
  }
  float v = 0.f;
  if (used) {
    const int ci = cin_off + ks * 8 + chunk * 4 + e;
    v = w[((int64_t)ci * Cout + n) * 27 + (k[0] * 9 + k[1] * 3 + k[2])];
    if (scale) v *= scale[n];
  }
  if (nsplit == 3) {
    const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    img[i] = hi;
    img[per_term + i] = v - hi;
  } else {
    img[i] = v;
  }
}

template <int CIN, int COUT, int NSPLIT, int KM, int NS>
static int launch_t(const Params& p, int B, cudaStream_t st) {
  using C = TCfg<CIN, COUT, NSPLIT, KM, NS>;
  static_assert(C::SMEM <= 232448, "shared memory budget exceeded");
  auto kern = convT3d_umma_kernel<CIN, COUT, NSPLIT, KM, NS>;
  static SmemAttrOnce smem_attr;
  {
    cudaError_t e = smem_attr.ensure(kern, (int)C::SMEM);
    RCMVS_REQUIRE(e == cudaSuccess, RCMVS_ERR_CUDA, "rcmvs_conv_transpose3d_tc_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  dim3 grid(ceil_div(p.W, C::TWI), ceil_div(p.H, C::ROWS), B * p.chunks);
  RCMVS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, RCMVS_ERR_UNSUPPORTED, "rcmvs_conv_transpose3d_tc_fwd: grid too large");
  kern<<<grid, threads_for(NS), C::SMEM, st>>>(p);
  RCMVS_CHECK_LAUNCH("rcmvs_conv_transpose3d_tc_fwd");
  return RCMVS_OK;
}

static bool t_supported(int Cin, int Cout) { return (Cin == 16 && Cout == 8) || (Cin == 32 && Cout == 16); }
// channels per launch (channel groups accumulate through y): B image + slabs must fit 227 KB
static int t_cin_k(int Cin, int Cout, int nsplit) {
  if (Cout == 8) return 16;
  return nsplit == 3 ? 8 : 16;
}

}  // namespace umma
}  // namespace rcmvs

extern "C" int64_t rcmvs_conv3d_tc_weight_floats(int Cin, int Cout, int nsplit) {
  const rcmvs::umma::Plan pl = rcmvs::umma::plan(Cin, Cout, nsplit);
  return (int64_t)(nsplit == 3 ? 2 : 1) * 3 * (Cin / 8) * 2 * 3 * pl.g * 4;
}

extern "C" int rcmvs_conv3d_tc_pack(const float* w, const float* scale, float* img, int Cout, int Cin, int nsplit,
                                    rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(w && img, RCMVS_ERR_BAD_ARG, "rcmvs_conv3d_tc_pack: null pointer");
  RCMVS_REQUIRE(umma::supported(Cin, Cout) && (nsplit == 1 || nsplit == 3), RCMVS_ERR_UNSUPPORTED,
                "rcmvs_conv3d_tc_pack: unsupported Cin=%d Cout=%d nsplit=%d", Cin, Cout, nsplit);
  const umma::Plan pl = umma::plan(Cin, Cout, nsplit);
  const int terms = nsplit == 3 ? 2 : 1;
  const int per_term = 3 * (pl.cin_k / 8) * 2 * 3 * pl.g * 4;
  for (int g = 0; g < Cin / pl.cin_k; ++g) {
    umma::pack_umma_weight_kernel<<<ceil_div(per_term, 256), 256, 0, as_stream(stream)>>>(
        w, scale, img + (int64_t)g * terms * per_term, Cout, Cin, g * pl.cin_k, pl.cin_k, pl.cp, pl.g, nsplit);
  }
  RCMVS_CHECK_LAUNCH("rcmvs_conv3d_tc_pack");
  return RCMVS_OK;
}

extern "C" int rcmvs_conv3d_tc_fwd(const float* x, const float* wimg, const float* bias, const float* skip, float* y,
                                   int B, int D, int H, int W, int Cin, int Cout, int nsplit, int relu,
                                   rcmvs_stream_t stream) {
  using namespace rcmvs;
  using namespace rcmvs::umma;
  RCMVS_REQUIRE(x && wimg && y, RCMVS_ERR_BAD_ARG, "rcmvs_conv3d_tc_fwd: null pointer");
  RCMVS_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, RCMVS_ERR_BAD_ARG, "rcmvs_conv3d_tc_fwd: bad dims");
  RCMVS_REQUIRE(umma::supported(Cin, Cout) && (nsplit == 1 || nsplit == 3), RCMVS_ERR_UNSUPPORTED,
                "rcmvs_conv3d_tc_fwd: unsupported Cin=%d Cout=%d nsplit=%d", Cin, Cout, nsplit);
  const Plan pl = plan(Cin, Cout, nsplit);
  const int orows = 4 * pl.km - 2;
  // depth chunks: one CTA per SM (shared memory), so pick the chunk count that minimises
  //   waves(cols*chunks / 148) x slabs per CTA (planes + 2 halo slabs)
  const int cols = ceil_div(W, TW) * ceil_div(H, orows) * B;
  int chunks = 1;
  {
    long best = -1;
    for (int c = 1; c <= max(1, D / 4); ++c) {
      const int dc = ceil_div(D, c);
      const int cc = ceil_div(D, dc);
      const long cost = (long)ceil_div(cols * cc, 148) * (dc + 2);
      if (best < 0 || cost < best) { best = cost; chunks = cc; }
    }
  }
  const int d_chunk = ceil_div(D, chunks);
  chunks = ceil_div(D, d_chunk);
  cudaStream_t st = as_stream(stream);
  {
    Params p;
    p.x = x; p.wimg = wimg; p.bias = bias; p.skip = skip; p.y = y;
    p.D = D; p.H = H; p.W = W; p.cin_total = Cin; p.cin_off = 0;
    p.d_chunk = d_chunk; p.chunks = chunks;
    p.accumulate_in = 0; p.finalize = 1; p.relu = relu;
    p.dbg = tc_debug_mask();
    int rc = RCMVS_ERR_UNSUPPORTED;
#define RCMVS_TC_CASE(CI, CO, NSP, KMV, NSV, NGR) \
  if (Cin == CI * NGR && pl.cin_k == CI && Cout == CO && nsplit == NSP) rc = launch<CI, CO, NSP, KMV, NSV, NGR>(p, B, st)
    RCMVS_TC_CASE(8, 8, 1, 3, 6, 1); else RCMVS_TC_CASE(8, 8, 3, 3, 5, 1);
    else RCMVS_TC_CASE(16, 8, 1, 3, 5, 1); else RCMVS_TC_CASE(16, 8, 3, 3, 3, 1);
    else RCMVS_TC_CASE(32, 8, 1, 3, 3, 1); else RCMVS_TC_CASE(16, 8, 3, 2, 3, 2);
    else RCMVS_TC_CASE(8, 1, 1, 3, 6, 1); else RCMVS_TC_CASE(8, 1, 3, 3, 5, 1);
    else RCMVS_TC_CASE(16, 16, 1, 2, 5, 1); else RCMVS_TC_CASE(16, 16, 3, 2, 3, 1);
    else set_error("rcmvs_conv3d_tc_fwd: no kernel for cin_k=%d Cout=%d nsplit=%d", pl.cin_k, Cout, nsplit);
#undef RCMVS_TC_CASE
    if (rc != RCMVS_OK) return rc;
  }
  return RCMVS_OK;
}

extern "C" int64_t rcmvs_conv_transpose3d_tc_weight_floats(int Cin, int Cout, int nsplit) {
  return (int64_t)(nsplit == 3 ? 2 : 1) * 8 * (Cin / 8) * 2 * (8 * Cout) * 4;
}

extern "C" int rcmvs_conv_transpose3d_tc_pack(const float* w, const float* scale, float* img, int Cout, int Cin,
                                              int nsplit, rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(w && img, RCMVS_ERR_BAD_ARG, "rcmvs_conv_transpose3d_tc_pack: null pointer");
  RCMVS_REQUIRE(umma::t_supported(Cin, Cout) && (nsplit == 1 || nsplit == 3), RCMVS_ERR_UNSUPPORTED,
                "rcmvs_conv_transpose3d_tc_pack: unsupported Cin=%d Cout=%d nsplit=%d", Cin, Cout, nsplit);
  const int cin_k = umma::t_cin_k(Cin, Cout, nsplit);
  const int terms = nsplit == 3 ? 2 : 1;
  const int per_term = 8 * (cin_k / 8) * 2 * (8 * Cout) * 4;
  for (int g = 0; g < Cin / cin_k; ++g)
    umma::pack_umma_tweight_kernel<<<ceil_div(per_term, 256), 256, 0, as_stream(stream)>>>(
        w, scale, img + (int64_t)g * terms * per_term, Cout, g * cin_k, cin_k, nsplit);
  RCMVS_CHECK_LAUNCH("rcmvs_conv_transpose3d_tc_pack");
  return RCMVS_OK;
}

extern "C" int rcmvs_conv_transpose3d_tc_fwd(const float* x, const float* wimg, const float* bias, const float* skip,
                                             float* y, int B, int D, int H, int W, int Cin, int Cout, int nsplit,
                                             int relu, rcmvs_stream_t stream) {
  using namespace rcmvs;
  using namespace rcmvs::umma;
  RCMVS_REQUIRE(x && wimg && y, RCMVS_ERR_BAD_ARG, "rcmvs_conv_transpose3d_tc_fwd: null pointer");
  RCMVS_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, RCMVS_ERR_BAD_ARG, "rcmvs_conv_transpose3d_tc_fwd: bad dims");
  RCMVS_REQUIRE(t_supported(Cin, Cout) && (nsplit == 1 || nsplit == 3), RCMVS_ERR_UNSUPPORTED,
                "rcmvs_conv_transpose3d_tc_fwd: unsupported Cin=%d Cout=%d nsplit=%d", Cin, Cout, nsplit);
  const int cin_k = t_cin_k(Cin, Cout, nsplit);
  const int km = 2, rows = 4 * km;
  const int cols = ceil_div(W, 31) * ceil_div(H, rows) * B;
  int chunks = 1;
  {
    long best = -1;
    for (int c = 1; c <= max(1, D / 2); ++c) {
      const int dc = ceil_div(D, c);
      const int cc = ceil_div(D, dc);
      const long cost = (long)ceil_div(cols * cc, 148) * (dc + 1);
      if (best < 0 || cost < best) { best = cost; chunks = cc; }
    }
  }
  const int d_chunk = ceil_div(D, chunks);
  chunks = ceil_div(D, d_chunk);
  const int groups = Cin / cin_k;
  const int terms = nsplit == 3 ? 2 : 1;
  const int per_term = 8 * (cin_k / 8) * 2 * (8 * Cout) * 4;
  cudaStream_t st = as_stream(stream);
  for (int g = 0; g < groups; ++g) {
    Params p;
    p.x = x; p.wimg = wimg + (int64_t)g * terms * per_term; p.bias = bias; p.skip = skip; p.y = y;
    p.D = D; p.H = H; p.W = W; p.cin_total = Cin; p.cin_off = g * cin_k;
    p.d_chunk = d_chunk; p.chunks = chunks;
    p.accumulate_in = g > 0; p.finalize = g == groups - 1; p.relu = relu;
    p.dbg = 0;
    int rc = RCMVS_ERR_UNSUPPORTED;
    if (Cout == 8 && nsplit == 1) rc = launch_t<16, 8, 1, 2, 5>(p, B, st);
    else if (Cout == 8 && nsplit == 3) rc = launch_t<16, 8, 3, 2, 3>(p, B, st);
    else if (Cout == 16 && nsplit == 1) rc = launch_t<16, 16, 1, 2, 4>(p, B, st);
    else if (Cout == 16 && nsplit == 3) rc = launch_t<8, 16, 3, 2, 4>(p, B, st);
    if (rc != RCMVS_OK) return rc;
  }
  return RCMVS_OK;
}

// ==================================================================================================================
// Gather implicit-GEMM on tcgen05 for the SMALL levels of the 3-D U-Net (1/4 and 1/8 resolution: conv3..conv6, conv7,
// conv9; reference models/modules.py:478-486).  These layers have few voxels (7.7 k .. 61 k at stage 1) but many
// channels (16..64): the slab kernels above cannot fill the machine there and the FFMA kernels are latency-bound
// (0.1-0.2 ms per layer for 1-3 GFLOP).  Here M = 128 OUTPUT voxels taken from a flat index space (no tile shape, no
// width waste), N = Cout, and the K loop walks (tap, 16-channel group) items.  For every item a loader warp GATHERS
// the 128 x 16-channel A tile straight from global/L2 and copies the matching 16 x Cout weight block; one loader warp
// per pipeline stage keeps NS items in flight, one thread issues the MMAs, accumulators stay in TMEM for the tile.
//   * stride-1 / stride-2 convolutions: rows = consecutive output voxels, 27 taps, input = s*o + k - 1.
//   * transposed convolution (k3 s2 p1 op1): rows = 128 consecutive voxels of ONE output parity class (oz&1,oy&1,ox&1);
//     a class only receives the taps whose parity matches (1, 2, 4 or 8 of the 27), so the K loop is 27/8 taps long on
//     average instead of 27: input = j + d with o = 2j + parity.
//   * 3xTF32: the hi/lo halves of the ACTIVATION are either produced on the fly by the loader (registers, `x_lo` null)
//     or - cheaper, the split is done once instead of once per tap - read pre-split from the producing layer's epilogue
//     (`y_hi`/`y_lo` outputs), in which case A and B both move with cp.async and the loader does no arithmetic.
// ==================================================================================================================
namespace rcmvs {
namespace umma {

struct GParams {
  const float* x;      // [B][D][H][W][Cin]  (the hi half when x_lo is given)
  const float* x_lo;   // lo half of the pre-split input, or null
  const float* wimg;   // [27][Cin/16][terms][2 ks][2 chunks][COUT][4]
  const float* bias;
  const float* skip;   // [B][Do][Ho][Wo][COUT] or null
  float* y;            // [B][Do][Ho][Wo][COUT] or null
  float* y_hi;         // optional pre-split copy of the result for a following gather layer
  float* y_lo;
  int D, H, W, Cin, lg, Do, Ho, Wo;   // lg = log2(Cin/16)
  int mode;            // 0: conv stride 1, 1: conv stride 2, 2: transposed conv (stride 2, output_padding 1)
  int relu;
  int tiles_per_class; // transposed: M tiles per parity class
  int dbg;             // development only (RCMVS_TC_DBG): 1 = skip MMA issue, 2 = skip A gather, 4 = skip weight copy
};

template <int COUT, int NSPLIT, bool PRESPLIT, int NS, bool CIN8 = false>
struct GCfg {
  static constexpr int NPART = NSPLIT == 3 ? 2 : 1;
  static constexpr bool ASYNC_A = NSPLIT == 1 || PRESPLIT;          // A tile moves with cp.async (no arithmetic)
  static constexpr int CHUNK_B = 128 * 16 + 32;                     // one 4-channel chunk plane: 128 rows x 16 B, +32 B
                                                                    // so the 4 chunks a quarter-warp stores hit distinct banks
  static constexpr int A_PART = 4 * CHUNK_B;                        // 16 channels
  static constexpr int B_PART = 2 * 2 * COUT * 16;                  // 2 K-steps x 2 chunks x COUT rows x 16 B
  static constexpr int STAGE_B = NPART * (A_PART + B_PART);
  static constexpr int TMEM_COLS = COUT <= 32 ? 32 : 64;
  static constexpr size_t SMEM = (size_t)NS * STAGE_B + 256 + 128 * 8 + 32 * 16;   // stages | barriers | rows | taps
  // cp.async loaders never block on their own copies (the copies arrive on the mbarrier themselves), so 4 warps feed
  // a ring of NS stages; the register path puts TWO warps on every stage (64 rows each): the per-item instruction chain
  // of a warp (gather, hi/lo split, 128-bit stores) is what bounds the ring's round trip
  static constexpr int NLW = ASYNC_A ? 4 : 2 * NS;
  static constexpr int THREADS = 32 * (4 + NLW + 1);
  // short rings fit twice on an SM: two CTAs = two MMA issue threads and two rings whose stalls interleave
  static constexpr int MIN_CTAS = (SMEM <= 113 * 1024 && THREADS <= 512) ? 2 : 1;
};

// tap t of this tile -> weight tap index (kd*9+kh*3+kw) and input offset (dz,dy,dx) relative to the row's base voxel
__device__ __forceinline__ void gather_tap(int mode, int t, int pz, int py, int px, int& wtap, int& dz, int& dy, int& dx) {
  if (mode == 2) {
    // per dimension: parity 0 -> k = 1, d = 0; parity 1 -> (k = 0, d = +1), (k = 2, d = 0)
    const int nx = 1 + px, ny = 1 + py;
    const int tx = t % nx, ty = (t / nx) % ny, tz = t / (nx * ny);
    const int kw = px ? 2 * tx : 1, kh = py ? 2 * ty : 1, kd = pz ? 2 * tz : 1;
    dx = px ? 1 - tx : 0; dy = py ? 1 - ty : 0; dz = pz ? 1 - tz : 0;
    wtap = kd * 9 + kh * 3 + kw;
  } else {
    const int kd = t / 9, kh = (t / 3) % 3, kw = t % 3;
    dz = kd - 1; dy = kh - 1; dx = kw - 1;
    wtap = t;
  }
}

// CIN8: 8-channel input (conv1, 8 -> 16 stride 2): one tap is a single K-step, so an item carries TWO live taps
// (chunks 0-1 = first tap, chunks 2-3 = second tap; the weight image is per tap: [27][terms][2 chunks][COUT][4]).
template <int COUT, int NSPLIT, bool PRESPLIT, int NS, bool CIN8>
__global__ void __launch_bounds__(GCfg<COUT, NSPLIT, PRESPLIT, NS>::THREADS, GCfg<COUT, NSPLIT, PRESPLIT, NS>::MIN_CTAS)
    conv3d_gather_umma_kernel(const GParams p) {
  using C = GCfg<COUT, NSPLIT, PRESPLIT, NS>;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)NS * C::STAGE_B);
  uint64_t* full = bars;
  uint64_t* empty = bars + NS;
  uint64_t* acc_full = bars + 2 * NS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;

  // ---- which rows this tile owns ---------------------------------------------------------------------------------
  int pz = 0, py = 0, px = 0, tile = blockIdx.x, ntaps = 27;
  int64_t Mspace;                       // size of the flat index space the rows are taken from
  int Rh, Rw;                           // its (h, w) extents
  if (p.mode == 2) {
    const int cls = blockIdx.x / p.tiles_per_class;
    tile = blockIdx.x - cls * p.tiles_per_class;
    pz = (cls >> 2) & 1; py = (cls >> 1) & 1; px = cls & 1;
    ntaps = (1 + pz) * (1 + py) * (1 + px);
    Mspace = (int64_t)p.D * p.H * p.W; Rh = p.H; Rw = p.W;
  } else {
    Mspace = (int64_t)p.Do * p.Ho * p.Wo; Rh = p.Ho; Rw = p.Wo;
  }
  const int64_t m0 = (int64_t)tile * 128;
  const int ngroups = 1 << p.lg;

  // ---- row table: element offset of each row's base input voxel + the set of taps that stay inside the volume --------
  int2* geo = reinterpret_cast<int2*>(smem + (size_t)NS * C::STAGE_B + 256);
  unsigned* tapmask = reinterpret_cast<unsigned*>(tmem_slot + 1);      // taps that reach at least one row of this tile
  if (threadIdx.x == 128) *tapmask = 0u;
  __syncthreads();
  if (threadIdx.x < 128) {
    const int s = p.mode == 1 ? 2 : 1;
    const int64_t m = m0 + threadIdx.x;
    int2 g = make_int2(0, 0);
    if (m < Mspace) {
      const int bx = (int)(m % Rw) * s, by = (int)((m / Rw) % Rh) * s, bz = (int)(m / ((int64_t)Rw * Rh)) * s;
      g.x = ((bz * p.H + by) * p.W + bx) * p.Cin;
      for (int t = 0; t < ntaps; ++t) {
        int wtap, dz, dy, dx;
        gather_tap(p.mode, t, pz, py, px, wtap, dz, dy, dx);
        const int iz = bz + dz, iy = by + dy, ix = bx + dx;
        if (iz >= 0 && iy >= 0 && ix >= 0 && iz < p.D && iy < p.H && ix < p.W) g.y |= 1 << t;
      }
    }
    geo[threadIdx.x] = g;
    if (g.y) atomicOr(tapmask, (unsigned)g.y);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(full + i, C::ASYNC_A ? 32 : 2); mbar_init(empty + i, 1); }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // taps that fall outside the volume for every row (shallow volumes: kd = 0 / 2 of a 1-plane input) are skipped;
  // the live ones are listed once: {bit in the row masks, weight tap kd*9+kh*3+kw, element offset of the tap}
  const unsigned tmask = *tapmask;
  const int nlive = __popc(tmask);
  int4* taps = reinterpret_cast<int4*>(smem + (size_t)NS * C::STAGE_B + 256 + 128 * 8);
  if (threadIdx.x < nlive) {
    const int tt = __fns(tmask, 0, threadIdx.x + 1);
    int wtap, dz, dy, dx;
    gather_tap(p.mode, tt, pz, py, px, wtap, dz, dy, dx);
    taps[threadIdx.x] = make_int4(tt, wtap, ((dz * p.H + dy) * p.W + dx) * p.Cin, 0);
  }
  __syncthreads();
  const int nitems = CIN8 ? (nlive + 1) >> 1 : nlive << p.lg;

  if (warp >= 4 && warp < 4 + C::NLW) {
    // =============================== gather loaders ==================================================================
    // 4 consecutive lanes fetch the 64 contiguous bytes (16 channels) of one row, so a warp instruction touches 8 cache
    // lines instead of 32; a lane serves chunk c = lane&3 of rows (lane>>2) + 8j, j = 0..15.
    const int lw = warp - 4;
    const int c = lane & 3, row0 = lane >> 2;
    const int64_t xbo = (int64_t)b * p.D * p.H * p.W * p.Cin;
    const int coff = CIN8 ? (c & 1) * 4 : c * 4;                    // this lane's channel offset inside the tap
    const float* xb = p.x + xbo + coff;
    const float* xlb = C::ASYNC_A && NSPLIT == 3 ? p.x_lo + xbo + coff : nullptr;
    constexpr int BV = C::NPART * C::B_PART / 16 / 32;              // 16-byte weight pieces per lane (2..16)
    const uint32_t lane_dst = c * C::CHUNK_B + row0 * 16;           // + j * 128 per row step

    // item -> (tap bit `t` of this lane in the row masks, weight block(s), element offset of the tap + channel group)
    auto item_geom = [&](int it, int& t, int2& wblk, int& toff) {
      if constexpr (CIN8) {
        const int te = 2 * it + (c >> 1);                  // chunks 0-1: first tap of the pair, chunks 2-3: second
        wblk = make_int2(taps[2 * it].y, 2 * it + 1 < nlive ? taps[2 * it + 1].y : -1);
        if (te < nlive) {
          const int4 q = taps[te];
          t = q.x; toff = q.z;
        } else {
          t = 31; toff = 0;                                // bit 31 of a row mask is never set -> zero rows
        }
      } else {
        const int4 q = taps[it >> p.lg];                   // the (it / ngroups)-th live tap of this tile
        const int grp = it & (ngroups - 1);
        t = q.x;
        wblk = make_int2(q.y * ngroups + grp, 0);
        toff = q.z + grp * 16;
      }
    };
    constexpr int BV0 = C::ASYNC_A ? 0 : 1;                 // register path: warp `half` copies its half of the pieces
    auto copy_weights = [&](int2 wblk, uint32_t stage_u, int half) {
      const uint32_t wdst = stage_u + C::NPART * C::A_PART;
      if (p.dbg & 4) return;
      if constexpr (CIN8) {
        // smem [term][ks][chunk][COUT] <- per-tap image blocks [term][chunk][COUT]; ks 0/1 = first/second tap
        const float4* wimg4 = reinterpret_cast<const float4*>(p.wimg);
#pragma unroll
        for (int e = 0; e < BV; ++e) {
          if (BV0 && (e & 1) != half) continue;
          const int piece = lane + 32 * e;
          const int term = piece / (4 * COUT), ks = (piece / (2 * COUT)) & 1, within = piece % (2 * COUT);
          const int wt = ks ? wblk.y : wblk.x;
          if (wt >= 0) cp_async16(wdst + piece * 16, wimg4 + ((int64_t)wt * C::NPART + term) * (2 * COUT) + within, true);
        }
      } else {
        const unsigned char* wsrc =
            reinterpret_cast<const unsigned char*>(p.wimg) + (int64_t)wblk.x * (C::NPART * C::B_PART);
#pragma unroll
        for (int e = 0; e < BV; ++e)
          if (!BV0 || (e & 1) == half) cp_async16(wdst + (lane + 32 * e) * 16, wsrc + (lane + 32 * e) * 16, true);
      }
    };

    if constexpr (C::ASYNC_A) {
      for (int it = lw; it < nitems; it += C::NLW) {
        int t, toff;
        int2 wblk;
        item_geom(it, t, wblk, toff);
        const int st = it % NS;
        const uint32_t stage_u = smem_u32(smem + (size_t)st * C::STAGE_B);
        mbar_wait(empty + st, ((it / NS) & 1) ^ 1);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int2 g = geo[row0 + 8 * j];
          const bool ok = ((g.y >> t) & 1) && !(p.dbg & 2);
          const int off = ok ? g.x + toff : 0;
          const uint32_t dst = stage_u + lane_dst + j * 128;
          cp_async16(dst, xb + off, ok);
          if constexpr (NSPLIT == 3) cp_async16(dst + C::A_PART, xlb + off, ok);
        }
        copy_weights(wblk, stage_u, 0);
        cp_async_mbar_arrive_noinc(full + st);
      }
      cp_async_commit();
      cp_async_wait<0>();                 // nothing of this warp is in flight when the CTA tears down
    } else {
      // split-in-loader path: two warps per stage (64 rows each); the A rows of the NEXT item are gathered into
      // registers while this one is consumed
      const int st = lw >> 1, half = lw & 1;
      const int rbase = half * 64 + row0;
      float4 v[8];
      auto gather = [&](int it, int2& wblk) {
        int t, toff;
        item_geom(it, t, wblk, toff);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int2 g = geo[rbase + 8 * j];
          const bool ok = ((g.y >> t) & 1) && !(p.dbg & 2);
          v[j] = ok ? ldg4(xb + g.x + toff) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      unsigned char* stage = smem + (size_t)st * C::STAGE_B;
      int2 wblk = make_int2(0, 0);
      if (st < nitems) gather(st, wblk);
      for (int it = st; it < nitems; it += NS) {
        mbar_wait(empty + st, ((it / NS) & 1) ^ 1);
        copy_weights(wblk, smem_u32(stage), half);
        cp_async_commit();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (p.dbg & 16) break;
          const uint32_t off = c * C::CHUNK_B + (rbase + 8 * j) * 16;
          const float4 x4 = v[j];
          float4 h, l;
          h.x = __uint_as_float(__float_as_uint(x4.x) & 0xFFFFE000u); l.x = x4.x - h.x;
          h.y = __uint_as_float(__float_as_uint(x4.y) & 0xFFFFE000u); l.y = x4.y - h.y;
          h.z = __uint_as_float(__float_as_uint(x4.z) & 0xFFFFE000u); l.z = x4.z - h.z;
          h.w = __uint_as_float(__float_as_uint(x4.w) & 0xFFFFE000u); l.w = x4.w - h.w;
          *reinterpret_cast<float4*>(stage + off) = h;
          *reinterpret_cast<float4*>(stage + C::A_PART + off) = l;
        }
        if (it + NS < nitems) gather(it + NS, wblk);
        cp_async_wait<0>();
        if (!(p.dbg & 8)) fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(full + st);
      }
    }
  } else if (warp == 4 + C::NLW) {
    // =============================== MMA issuer ====================================================================
    const bool leader = elect_one();
    constexpr uint32_t idesc = make_idesc(128, COUT);
    for (int it = 0; it < nitems; ++it) {
      const int st = it % NS;
      mbar_wait(full + st, (it / NS) & 1);
      if constexpr (C::ASYNC_A) fence_proxy_async();       // cp.async data observed through the mbarrier -> async proxy
      tc_fence_after();
      const uint32_t sbase = smem_u32(smem + (size_t)st * C::STAGE_B);
      const uint32_t a_lo0 = desc_lo(sbase, C::CHUNK_B);
      const uint32_t b_lo0 = desc_lo(sbase + C::NPART * C::A_PART, COUT * 16);
      const int nk = (CIN8 && it == nitems - 1 && (nlive & 1)) ? 1 : 2;     // odd tap count: last item is half full
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const uint32_t a_d = a_lo0 + (uint32_t)((ks * 2 * C::CHUNK_B) >> 4);
        const uint32_t b_d = b_lo0 + (uint32_t)((ks * 2 * COUT * 16) >> 4);
        const uint32_t acc = (it == 0 && ks == 0) ? 0u : 1u;
        if (leader && ks < nk && !(p.dbg & 1)) {
          if constexpr (NSPLIT == 3) {
            umma_tf32_lohi_acc(tmem_base, a_d + (C::A_PART >> 4), b_d, idesc, acc);                 // x_lo · w_hi
            umma_tf32_lohi_acc(tmem_base, a_d, b_d + (C::B_PART >> 4), idesc, 1u);                  // x_hi · w_lo
            umma_tf32_lohi_acc(tmem_base, a_d, b_d, idesc, 1u);                                     // x_hi · w_hi
          } else {
            umma_tf32_lohi_acc(tmem_base, a_d, b_d, idesc, acc);
          }
        }
      }
      if (leader) {
        umma_commit(empty + st);
        if (it == nitems - 1) umma_commit(acc_full);
      }
      __syncwarp();
    }
  } else if (warp < 4) {
    // =============================== epilogue: one output voxel per thread ==========================================
    const int64_t m = m0 + warp * 32 + lane;
    int64_t ovox = m;
    if (p.mode == 2 && m < Mspace) {
      const int jx = (int)(m % p.W), jy = (int)((m / p.W) % p.H), jz = (int)(m / ((int64_t)p.W * p.H));
      ovox = ((int64_t)(2 * jz + pz) * p.Ho + (2 * jy + py)) * p.Wo + (2 * jx + px);
    }
    // the skip row is requested before the (long) wait for the accumulators, so its latency never shows
    constexpr bool kPrefetchSkip = COUT <= 32;
    const int64_t o = ((int64_t)b * p.Do * p.Ho * p.Wo + ovox) * COUT;
    float4 sk[kPrefetchSkip ? COUT / 4 : 1];
    if constexpr (kPrefetchSkip) {
#pragma unroll
      for (int c = 0; c < COUT; c += 4)
        sk[c / 4] = (p.skip && m < Mspace) ? ldg4(p.skip + o + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    while (!mbar_try_wait(acc_full, 0)) __nanosleep(512);      // idle for the whole main loop: stay off the issue ports
    tc_fence_after();
    float v[COUT];
#pragma unroll
    for (int c0 = 0; c0 < COUT; c0 += 16) tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, &v[c0]);
    if (m < Mspace) {
#pragma unroll
      for (int c = 0; c < COUT; c += 4) {
        float4 r = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
        if (p.bias) {
          const float4 q = ldg4(p.bias + c);
          r.x += q.x; r.y += q.y; r.z += q.z; r.w += q.w;
        }
        if (p.relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
        if (p.skip) {
          const float4 q = kPrefetchSkip ? sk[kPrefetchSkip ? c / 4 : 0] : ldg4(p.skip + o + c);
          r.x += q.x; r.y += q.y; r.z += q.z; r.w += q.w;
        }
        if (p.y) st4(p.y + o + c, r);
        if (p.y_hi) {
          float4 h;
          h.x = __uint_as_float(__float_as_uint(r.x) & 0xFFFFE000u);
          h.y = __uint_as_float(__float_as_uint(r.y) & 0xFFFFE000u);
          h.z = __uint_as_float(__float_as_uint(r.z) & 0xFFFFE000u);
          h.w = __uint_as_float(__float_as_uint(r.w) & 0xFFFFE000u);
          st4(p.y_hi + o + c, h);
          st4(p.y_lo + o + c, make_float4(r.x - h.x, r.y - h.y, r.z - h.z, r.w - h.w));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
  }
}

// img[tap][grp][term][ks][chunk][n][e] = split_term( w(n, ci = grp*16 + ks*8 + chunk*4 + e, tap) * scale[n] )
__global__ void pack_gather_weight_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                          float* __restrict__ img, int Cout, int Cin, int transposed, int nsplit) {
  const int terms = nsplit == 3 ? 2 : 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (Cin == 8) {                                   // per-tap image: [tap][term][chunk][n][e], ci = chunk*4 + e
    const int blk8 = 2 * Cout * 4;
    if (i >= 27 * blk8) return;
    const int e = i & 3, n = (i >> 2) % Cout, chunk = ((i >> 2) / Cout) & 1, tap = i / blk8;
    const int ci = chunk * 4 + e;
    float v = transposed ? w[((int64_t)ci * Cout + n) * 27 + tap] : w[((int64_t)n * Cin + ci) * 27 + tap];
    if (scale) v *= scale[n];
    const int64_t base = (int64_t)tap * terms * blk8 + ((chunk * Cout + n) * 4 + e);
    const float hi = nsplit == 3 ? __uint_as_float(__float_as_uint(v) & 0xFFFFE000u) : v;
    img[base] = hi;
    if (nsplit == 3) img[base + blk8] = v - hi;
    return;
  }
  const int blk = 2 * 2 * Cout * 4;                 // floats per (tap, grp, term)
  const int ngroups = Cin / 16;
  const int total = 27 * ngroups * blk;
  if (i >= total) return;
  const int e = i & 3;
  int r = i >> 2;
  const int n = r % Cout; r /= Cout;
  const int chunk = r & 1; r >>= 1;
  const int ks = r & 1; r >>= 1;
  const int grp = r % ngroups;
  const int tap = r / ngroups;
  const int ci = grp * 16 + ks * 8 + chunk * 4 + e;
  float v = transposed ? w[((int64_t)ci * Cout + n) * 27 + tap] : w[((int64_t)n * Cin + ci) * 27 + tap];
  if (scale) v *= scale[n];
  const int64_t base = ((int64_t)(tap * ngroups + grp) * terms) * blk + (((ks * 2 + chunk) * Cout + n) * 4 + e);
  if (nsplit == 3) {
    const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    img[base] = hi;
    img[base + blk] = v - hi;
  } else {
    img[base] = v;
  }
}

template <int COUT, int NSPLIT, bool PRESPLIT, int NS, bool CIN8 = false>
static int launch_g(const GParams& p, int B, cudaStream_t st) {
  using C = GCfg<COUT, NSPLIT, PRESPLIT, NS>;
  static_assert(C::SMEM <= 232448, "shared memory budget exceeded");
  auto kern = conv3d_gather_umma_kernel<COUT, NSPLIT, PRESPLIT, NS, CIN8>;
  static SmemAttrOnce smem_attr;
  {
    cudaError_t e = smem_attr.ensure(kern, (int)C::SMEM);
    RCMVS_REQUIRE(e == cudaSuccess, RCMVS_ERR_CUDA, "rcmvs_conv3d_gemm_tc_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  const int64_t Mspace = p.mode == 2 ? (int64_t)p.D * p.H * p.W : (int64_t)p.Do * p.Ho * p.Wo;
  const int64_t tiles = ceil_div<int64_t>(Mspace, 128);
  dim3 grid((unsigned)(p.mode == 2 ? 8 * tiles : tiles), B);
  kern<<<grid, C::THREADS, C::SMEM, st>>>(p);
  RCMVS_CHECK_LAUNCH("rcmvs_conv3d_gemm_tc_fwd");
  return RCMVS_OK;
}

}  // namespace umma
}  // namespace rcmvs

extern "C" int64_t rcmvs_conv3d_gemm_tc_weight_floats(int Cin, int Cout, int nsplit) {
  if (Cin == 8) return (int64_t)27 * (nsplit == 3 ? 2 : 1) * 2 * Cout * 4;
  return (int64_t)27 * (Cin / 16) * (nsplit == 3 ? 2 : 1) * 2 * 2 * Cout * 4;
}

extern "C" int rcmvs_conv3d_gemm_tc_pack(const float* w, const float* scale, float* img, int Cout, int Cin,
                                         int transposed, int nsplit, rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(w && img, RCMVS_ERR_BAD_ARG, "rcmvs_conv3d_gemm_tc_pack: null pointer");
  RCMVS_REQUIRE(((Cin == 8 && Cout == 16) ||
                 ((Cin == 16 || Cin == 32 || Cin == 64) && (Cout == 16 || Cout == 32 || Cout == 64))) &&
                    (nsplit == 1 || nsplit == 3),
                RCMVS_ERR_UNSUPPORTED, "rcmvs_conv3d_gemm_tc_pack: unsupported Cin=%d Cout=%d nsplit=%d", Cin, Cout, nsplit);
  const int total = Cin == 8 ? 27 * 2 * Cout * 4 : 27 * (Cin / 16) * 2 * 2 * Cout * 4;
  umma::pack_gather_weight_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(w, scale, img, Cout, Cin,
                                                                                       transposed, nsplit);
  RCMVS_CHECK_LAUNCH("rcmvs_conv3d_gemm_tc_pack");
  return RCMVS_OK;
}

extern "C" int rcmvs_conv3d_gemm_tc_fwd(const float* x, const float* x_lo, const float* wimg, const float* bias,
                                        const float* skip, float* y, float* y_hi, float* y_lo, int B, int D, int H,
                                        int W, int Cin, int Cout, int mode, int nsplit, int relu,
                                        rcmvs_stream_t stream) {
  using namespace rcmvs;
  using namespace rcmvs::umma;
  RCMVS_REQUIRE(x && wimg && (y || y_hi), RCMVS_ERR_BAD_ARG, "rcmvs_conv3d_gemm_tc_fwd: null pointer");
  RCMVS_REQUIRE((y_hi == nullptr) == (y_lo == nullptr), RCMVS_ERR_BAD_ARG,
                "rcmvs_conv3d_gemm_tc_fwd: y_hi and y_lo go together");
  RCMVS_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && B <= 65535, RCMVS_ERR_BAD_ARG, "rcmvs_conv3d_gemm_tc_fwd: bad dims");
  RCMVS_REQUIRE(((Cin == 8 && Cout == 16 && !x_lo) ||
                 ((Cin == 16 || Cin == 32 || Cin == 64) && (Cout == 16 || Cout == 32 || Cout == 64))) &&
                    (nsplit == 1 || nsplit == 3) && mode >= 0 && mode <= 2,
                RCMVS_ERR_UNSUPPORTED, "rcmvs_conv3d_gemm_tc_fwd: unsupported Cin=%d Cout=%d mode=%d nsplit=%d", Cin,
                Cout, mode, nsplit);
  RCMVS_REQUIRE(!(x_lo && nsplit == 1), RCMVS_ERR_BAD_ARG, "rcmvs_conv3d_gemm_tc_fwd: x_lo only with nsplit=3");
  RCMVS_REQUIRE((int64_t)D * H * W * Cin < (int64_t)1 << 31, RCMVS_ERR_UNSUPPORTED,
                "rcmvs_conv3d_gemm_tc_fwd: one sample of the input must stay below 2^31 elements");
  GParams p;
  p.x = x; p.x_lo = x_lo; p.wimg = wimg; p.bias = bias; p.skip = skip; p.y = y; p.y_hi = y_hi; p.y_lo = y_lo;
  p.D = D; p.H = H; p.W = W; p.Cin = Cin; p.lg = Cin <= 16 ? 0 : (Cin == 32 ? 1 : 2); p.mode = mode; p.relu = relu;
  if (mode == 2) { p.Do = 2 * D; p.Ho = 2 * H; p.Wo = 2 * W; }
  else if (mode == 1) { p.Do = (D - 1) / 2 + 1; p.Ho = (H - 1) / 2 + 1; p.Wo = (W - 1) / 2 + 1; }
  else { p.Do = D; p.Ho = H; p.Wo = W; }
  p.tiles_per_class = (int)ceil_div<int64_t>((int64_t)D * H * W, 128);
  p.dbg = tc_debug_mask();
  cudaStream_t st = as_stream(stream);
  const bool pre = x_lo != nullptr;
  if (Cin == 8) return nsplit == 3 ? launch_g<16, 3, false, 8, true>(p, B, st) : launch_g<16, 1, false, 8, true>(p, B, st);
#define RCMVS_G_CASE(CO, NSP, PRE, NSV) \
  if (Cout == CO && nsplit == NSP && pre == PRE) return launch_g<CO, NSP, PRE, NSV>(p, B, st)
  RCMVS_G_CASE(16, 1, false, 8); RCMVS_G_CASE(16, 3, false, 4); RCMVS_G_CASE(16, 3, true, 8);
  RCMVS_G_CASE(32, 1, false, 8); RCMVS_G_CASE(32, 3, false, 4); RCMVS_G_CASE(32, 3, true, 8);
  RCMVS_G_CASE(64, 1, false, 8); RCMVS_G_CASE(64, 3, false, 6); RCMVS_G_CASE(64, 3, true, 8);
#undef RCMVS_G_CASE
  return RCMVS_ERR_UNSUPPORTED;
}
