/* rcmvs.h — C ABI of librcmvs.so: the B200 (sm_100a) implementation of RC-MVSNet's per-view depth-inference
 * hot path.  The reference (Boese0601/RC-MVSNet) has no FFI of its own — its boundary is the Python API of
 * models/modules.py, models/casmvsnet.py, models/renderer.py, losses/homography.py.  Each entry
 * point below replaces the body of one of those Python functions and cites it; the Python packages under
 * rc-mvsnet_b200/{models,losses} keep the reference signatures and call these through ctypes (rcmvs/_ffi.py).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 unless noted; the caller (torch) owns all buffers incl. workspace;
 *   - the library never allocates device memory, never synchronises, and only enqueues on `stream`;
 *   - return 0 on success, a negative rcmvs_status otherwise; rcmvs_last_error() gives a thread-local message;
 *   - "NHWC"/"NDHWC" = channels-last internal layout; "NCHW"/"NCDHW" = the reference's layout;
 *   - B = batch, V = views (ref first), C = channels, D = depth planes, (h,w) = stage resolution.
 */
#ifndef RCMVS_H_
#define RCMVS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* rcmvs_stream_t; /* cudaStream_t */

enum rcmvs_status {
  RCMVS_OK = 0,
  RCMVS_ERR_BAD_ARG = -1,     /* null pointer, non-positive dim */
  RCMVS_ERR_UNSUPPORTED = -2, /* shape/channel count outside what the kernels are instantiated for */
  RCMVS_ERR_CUDA = -3,        /* launch failure; message holds cudaGetErrorString */
  RCMVS_ERR_ARCH = -4         /* device is not sm_100 */
};

int rcmvs_version(void);
const char* rcmvs_last_error(void);
/* 0 if the current device can run the library (compute capability 10.x), RCMVS_ERR_ARCH otherwise. */
int rcmvs_check_device(void);

/* ---- layout ------------------------------------------------------------------------------------------------ */
/* [N][C][S] -> [N][S][C] and back (S = H*W or D*H*W).  Used at the API boundary only. */
int rcmvs_planar_to_cl(const float* src, float* dst, int N, int C, int64_t S, rcmvs_stream_t stream);
int rcmvs_cl_to_planar(const float* src, float* dst, int N, int C, int64_t S, rcmvs_stream_t stream);

/* ---- projections: models/casmvsnet.py:78-81 + models/modules.py:315-317 ------------------------------------- */
/* proj [B][V][2][4][4] (extrinsic, intrinsic) -> rt [B][V-1][12]: row-major 3x4 [rot|trans] of
 * (K_v·E_v) · inverse(K_0·E_0), computed on device in fp64 and rounded to fp32 once. */
int rcmvs_relative_projections(const float* proj, float* rt, int B, int V, rcmvs_stream_t stream);
/* Same for already-multiplied 4x4 projections as homo_warping receives them: src_proj, ref_proj [B][4][4]. */
int rcmvs_relative_projection_pair(const float* src_proj, const float* ref_proj, float* rt, int B,
                                   rcmvs_stream_t stream);

/* ---- a5 depth hypotheses: models/modules.py:549-588, models/casmvsnet.py:191-217 ---------------------------- */
/* Stage 1: depth_values [B][Dv] -> planes [B][D]: dmin + i*(dmax-dmin)/(D-1). */
int rcmvs_depth_hypotheses_first(const float* depth_values, float* planes, int B, int Dv, int D,
                                 rcmvs_stream_t stream);
/* Stage >=2: prev_depth [B][hp][wp] -> hyp [B][D][h][w]: bilinear upsample to (H,W) (align_corners=False),
 * cur -/+ D/2*interval, D planes, then bilinear resize to (h,w). */
int rcmvs_depth_hypotheses_next(const float* prev_depth, float* hyp, int B, int hp, int wp, int H, int W, int h,
                                int w, int D, double interval_pixel, rcmvs_stream_t stream);
/* Same, with the pixel interval derived on the device: interval_ratio * (depth_values[0][Dv-1] - depth_values[0][0])
 * / Dv in double precision (the reference pulls min/max to the host, casmvsnet.py:172-174; this keeps the forward
 * free of host synchronisation). */
int rcmvs_depth_hypotheses_next_dev(const float* prev_depth, const float* depth_values, int Dv, double interval_ratio,
                                    float* hyp, int B, int hp, int wp, int H, int W, int h, int w, int D,
                                    rcmvs_stream_t stream);

/* ---- a1 homo_warping: models/modules.py:304-339 ------------------------------------------------------------- */
/* src_fea [B][C][h][w] (NCHW), rt [B][12], depth [B][D] (per_pixel=0) or [B][D][h][w] (per_pixel=1)
 * -> out [B][C][D][h][w].  Bilinear, zero padding per tap, align_corners=True.  C % 4 == 0, or C == 3. */
int rcmvs_homo_warp_fwd(const float* src_fea, const float* rt, const float* depth, int per_pixel, float* out,
                        int B, int C, int D, int h, int w, rcmvs_stream_t stream);
/* grad_out [B][C][D][h][w] -> grad_src [B][C][h][w] (must be zero-filled by the caller); the sample grid carries
 * no gradient in the reference (modules.py:314 torch.no_grad). */
int rcmvs_homo_warp_bwd(const float* grad_out, const float* rt, const float* depth, int per_pixel,
                        float* grad_src, int B, int C, int D, int h, int w, rcmvs_stream_t stream);

/* ---- a1+a2 fused plane sweep + variance: models/casmvsnet.py:65-100 / 254-288 -------------------------------- */
/* feats [B][V][h][w][C] (channels-last, ref view first), rt [B][V-1][12], depth as above
 * -> var [B][D][h][w][C] = Σ_v f_v²/V − (Σ_v f_v/V)²  (written once, channels-last so conv0 reads it directly).
 * var_noref (optional, may be NULL): same sums over the V-1 warped sources only, still divided by V
 * (training-mode DepthNet, casmvsnet.py:100).  C in {8,16,32}. */
int rcmvs_costvolume_var_fwd(const float* feats, const float* rt, const float* depth, int per_pixel, float* var,
                             float* var_noref, int B, int V, int C, int D, int h, int w, rcmvs_stream_t stream);

/* Same sweep, variance written as two bf16 planes var_split [2][B][D][h][w][C] (hi = bf16(v), lo = bf16(v - hi)): the
 * pre-split operand format of rcmvs_conv3d_slab_fwd (same bytes as fp32, no conversion pass in between).  h, w >= 2. */
int rcmvs_costvolume_var_fwd_split(const float* feats, const float* rt, const float* depth, int per_pixel,
                                   void* var_split, int B, int V, int C, int D, int h, int w, rcmvs_stream_t stream);

/* ---- a3 CostRegNet building blocks: models/modules.py:118-210, 470-501 --------------------------------------- */
/* Weight packing (once per checkpoint): w [Cout][Cin][3][3][3] (Conv3d) or [Cin][Cout][3][3][3]
 * (ConvTranspose3d) times the folded BatchNorm scale[Cout] (NULL = 1) -> packed [27][Cin][Cout]. */
int rcmvs_pack_conv3d_weight(const float* w, const float* scale, float* packed, int Cout, int Cin,
                             int transposed, rcmvs_stream_t stream);
/* y = act(conv3d(x, k=3, pad=1, stride) + bias) (+ skip);  x [B][D][H][W][Cin], y [B][Do][Ho][Wo][Cout].
 * bias/skip may be NULL; relu in {0,1}; stride in {1,2}.  The skip is added AFTER the ReLU (modules.py:497-499). */
int rcmvs_conv3d_fwd(const float* x, const float* packed_w, const float* bias, const float* skip, float* y, int B,
                     int D, int H, int W, int Cin, int Cout, int stride, int relu, rcmvs_stream_t stream);
/* Same layer with a caller-provided workspace of rcmvs_conv3d_workspace_floats() floats (0 = not needed): lets the
 * small 1/4- and 1/8-resolution layers split the input-channel reduction over more CTAs (deterministic split-K:
 * per-slice partial sums + a fixed-order reduce that applies bias / ReLU / skip). */
int64_t rcmvs_conv3d_workspace_floats(int B, int D, int H, int W, int Cin, int Cout, int stride);
int rcmvs_conv3d_fwd_ws(const float* x, const float* packed_w, const float* bias, const float* skip, float* y,
                        float* workspace, int B, int D, int H, int W, int Cin, int Cout, int stride, int relu,
                        rcmvs_stream_t stream);
/* ConvTranspose3d(k=3, stride=2, padding=1, output_padding=1): x [B][D][H][W][Cin] -> y [B][2D][2H][2W][Cout]. */
int rcmvs_conv_transpose3d_fwd(const float* x, const float* packed_w, const float* bias, const float* skip,
                               float* y, int B, int D, int H, int W, int Cin, int Cout, int relu,
                               rcmvs_stream_t stream);

/* Tensor-core (tcgen05.mma kind::tf32, TMEM accumulators) path for the stride-1 convolutions, same semantics as
 * rcmvs_conv3d_fwd with stride 1.  nsplit = 1: plain TF32 operands (the precision cuDNN uses for the reference on
 * GPU by torch default); nsplit = 3: error-compensated 3xTF32, fp32-equivalent.  Cin in {8,16,32}, Cout in {1,8,16}.
 * The weight image (operand layout of the kernel) is built once per checkpoint by rcmvs_conv3d_tc_pack from the
 * torch [Cout][Cin][3][3][3] weight and the folded BatchNorm scale; its size in floats is
 * rcmvs_conv3d_tc_weight_floats(). */
int64_t rcmvs_conv3d_tc_weight_floats(int Cin, int Cout, int nsplit);
int rcmvs_conv3d_tc_pack(const float* w, const float* scale, float* img, int Cout, int Cin, int nsplit,
                         rcmvs_stream_t stream);
int rcmvs_conv3d_tc_fwd(const float* x, const float* wimg, const float* bias, const float* skip, float* y, int B,
                        int D, int H, int W, int Cin, int Cout, int nsplit, int relu, rcmvs_stream_t stream);

/* Tensor-core transposed convolution (k3, s2, p1, output_padding 1), same semantics as rcmvs_conv_transpose3d_fwd.
 * (Cin, Cout) in {(16,8), (32,16)} = conv11 / conv9 of CostRegNet; the weight image is built from the torch
 * ConvTranspose3d weight [Cin][Cout][3][3][3]. */
int64_t rcmvs_conv_transpose3d_tc_weight_floats(int Cin, int Cout, int nsplit);
int rcmvs_conv_transpose3d_tc_pack(const float* w, const float* scale, float* img, int Cout, int Cin, int nsplit,
                                   rcmvs_stream_t stream);
int rcmvs_conv_transpose3d_tc_fwd(const float* x, const float* wimg, const float* bias, const float* skip, float* y,
                                  int B, int D, int H, int W, int Cin, int Cout, int nsplit, int relu,
                                  rcmvs_stream_t stream);

/* Round-2 engine for the stride-1 convolutions (conv0 / conv2 of CostRegNet, reference models/modules.py:473-475):
 * TMA-fed tcgen05.mma kind::f16 with error-compensated split-bf16 operands ("bf16x3": x = hi + lo, hi = bf16(x),
 * lo = bf16(x - hi); x·w ~= x_lo·w_hi + x_hi·w_lo + x_hi·w_hi, relative operand error <= 3·2^-17).  The activation
 * arrives PRE-SPLIT as two bf16 planes x_split [2][B][D][H][W][Cin] (hi plane, lo plane; 16-byte aligned), written by the
 * producing kernel (rcmvs_costvolume_var_fwd_split, y_split of this function, rcmvs_conv3d_fwd_split) or, at an API
 * boundary, by rcmvs_split_bf16.  Output: fp32 y and/or split-bf16 y_split (either may be NULL).
 * (Cin, Cout) in {(32,8), (16,8), (16,16), (48,8), (8,8), (8,1)} ((8,1) = `prob`, fp32 output only); with 8 input channels
 * one K = 16 step covers two kw taps.
 * Weight image: rcmvs_conv3d_slab_weight_elems() bf16 elements. */
int64_t rcmvs_conv3d_slab_weight_elems(int Cin, int Cout);
int rcmvs_conv3d_slab_pack(const float* w, const float* scale, void* img, int Cout, int Cin, rcmvs_stream_t stream);
int rcmvs_conv3d_slab_fwd(const void* x_split, const void* wimg, const float* bias, const float* skip, float* y,
                          void* y_split, int B, int D, int H, int W, int Cin, int Cout, int relu, rcmvs_stream_t stream);
/* 32-output-channel member: conv4 of CostRegNet (32 -> 32, quarter resolution, reference models/modules.py:478).  Only kd
 * is folded into N (TMEM ring of output planes); the nine (kh, kw) taps are descriptor start offsets, no epilogue
 * exchange.  Same operand formats as rcmvs_conv3d_slab_fwd.  (Cin, Cout) = (32, 32). */
int64_t rcmvs_conv3d_o_slab_weight_elems(int Cin, int Cout);
int rcmvs_conv3d_o_slab_pack(const float* w, const float* scale, void* img, int Cout, int Cin, rcmvs_stream_t stream);
int rcmvs_conv3d_o_slab_fwd(const void* x_split, const void* wimg, const float* bias, const float* skip, float* y,
                            void* y_split, int B, int D, int H, int W, int Cin, int Cout, int relu, rcmvs_stream_t stream);

/* Stride-2 members of the same family: conv1 (8 -> 16 channels) and conv3 (16 -> 32) of CostRegNet, 3x3x3, stride 2, pad 1
 * (reference models/modules.py:472, :476; conv1 of the renderer's CostReg, models/render_models.py:696).  The input plane
 * is fetched as four parity sub-planes by TMA boxes with traversal stride 2 (elementStrides); all 27 taps are summed in
 * TMEM (8 channels: one K = 16 MMA step covers two taps).  x_split [2][B][D][H][W][Cin] (D, H, W even) ->
 * y [B][D/2][H/2][W/2][Cout] fp32 and / or y_split (either may be NULL).  (Cin, Cout) in {(8,16), (16,32)}. */
int64_t rcmvs_conv3d_s2_slab_weight_elems(int Cin, int Cout);
int rcmvs_conv3d_s2_slab_pack(const float* w, const float* scale, void* img, int Cout, int Cin, rcmvs_stream_t stream);
int rcmvs_conv3d_s2_slab_fwd(const void* x_split, const void* wimg, const float* bias, float* y, void* y_split, int B,
                             int D, int H, int W, int Cin, int Cout, int relu, rcmvs_stream_t stream);
/* Transposed members of the family: conv11 (ConvTranspose3d 16 -> 8) and conv9 (32 -> 16) of CostRegNet, k3, stride 2,
 * pad 1, output_padding 1 + folded BN + ReLU + skip add (reference models/modules.py:166-210, :486-488, :498-499).
 * x_split [2][B][D][H][W][Cin] -> y [B][2D][2H][2W][Cout] fp32 and / or y_split [2][B][2D][2H][2W][Cout] (either may be
 * NULL); skip (fp32, same shape as y) may be NULL.  (Cin, Cout) in {(16,8), (32,16)}. */
int64_t rcmvs_conv_transpose3d_slab_weight_elems(int Cin, int Cout);
int rcmvs_conv_transpose3d_slab_pack(const float* w, const float* scale, void* img, int Cout, int Cin,
                                     rcmvs_stream_t stream);
int rcmvs_conv_transpose3d_slab_fwd(const void* x_split, const void* wimg, const float* bias, const float* skip, float* y,
                                    void* y_split, int B, int D, int H, int W, int Cin, int Cout, int relu,
                                    rcmvs_stream_t stream);
/* x [n] fp32 -> xs [2][n] bf16 (hi plane, lo plane). */
int rcmvs_split_bf16(const float* x, void* xs, int64_t n, rcmvs_stream_t stream);

/* Gather implicit-GEMM on tcgen05 for the small 1/4- and 1/8-resolution layers (conv3..conv7, conv9; reference
 * models/modules.py:478-486): any D/H/W, Cin and Cout in {16,32,64} (plus 8 -> 16, conv1); mode 0 = conv stride 1, 1 = conv stride 2,
 * 2 = transposed conv (k3, s2, p1, output_padding 1).  The weight image is built from the torch weight (Conv3d
 * [Cout][Cin][27] or, with transposed = 1, ConvTranspose3d [Cin][Cout][27]) and the folded BatchNorm scale.
 * With nsplit = 3 the input may come pre-split: x = hi half, x_lo = lo half (x_hi + x_lo == activation exactly), as
 * written by the producing layer through y_hi / y_lo; x_lo = NULL makes the kernel split on the fly.  y may be NULL
 * when only the split pair is wanted. */
int64_t rcmvs_conv3d_gemm_tc_weight_floats(int Cin, int Cout, int nsplit);
int rcmvs_conv3d_gemm_tc_pack(const float* w, const float* scale, float* img, int Cout, int Cin, int transposed,
                              int nsplit, rcmvs_stream_t stream);
int rcmvs_conv3d_gemm_tc_fwd(const float* x, const float* x_lo, const float* wimg, const float* bias,
                             const float* skip, float* y, float* y_hi, float* y_lo, int B, int D, int H, int W,
                             int Cin, int Cout, int mode, int nsplit, int relu, rcmvs_stream_t stream);

/* ---- a4 softmax + soft-argmin + photometric confidence: models/casmvsnet.py:111-122, modules.py:519-525 ------ */
/* cost [B][D][h][w]; depth planes [B][D] (per_pixel=0) or [B][D][h][w]; -> depth [B][h][w], conf [B][h][w],
 * idx [B][h][w] int32 (optional, may be NULL: clamp(trunc(Σ p·i))), prob [B][D][h][w] (optional, may be NULL). */
int rcmvs_softargmin_conf_fwd(const float* cost, const float* depth_planes, int per_pixel, float* depth,
                              float* conf, int32_t* idx, float* prob, int B, int D, int h, int w,
                              rcmvs_stream_t stream);

/* ---- f1 FeatureNet building blocks: models/modules.py:28-62, 363-464 ------------------------------------------------ */
/* img [N][C<=8][H][W] -> out [N][H][W][8] (zero-padded channels-last input of the first convolution). */
int rcmvs_images_to_nhwc8(const float* img, float* out, int N, int C, int H, int W, rcmvs_stream_t stream);
/* same, written as split-bf16 planes out_split [2][N][H][W][8] for the tensor-core first layer (rcmvs_conv2d_slab_fwd, Cin = 8) */
int rcmvs_images_to_nhwc8_split(const float* img, void* out_split, int N, int C, int H, int W, rcmvs_stream_t stream);
/* torch Conv2d weight [Cout][Cin][K][K] x folded BatchNorm scale[Cout] (NULL = 1) -> packed [K*K][Cin_pad][Cout]. */
int rcmvs_pack_conv2d_weight(const float* w, const float* scale, float* packed, int Cout, int Cin, int Cin_pad, int K,
                             rcmvs_stream_t stream);
/* y = act(conv2d(x, K, pad K/2, stride) + bias) (+ nearest-2x-upsampled `up`);  x [N][H][W][Cin],
 * y [N][Ho][Wo][Cout], up [N][Ho/2][Wo/2][Cout] or NULL.  (K,stride,Cout) in {(3,1,8|16|32),(5,2,16|32),(1,1,32)}. */
int rcmvs_conv2d_fwd(const float* x, const float* packed_w, const float* bias, const float* up, float* y, int N, int H,
                     int W, int Cin, int Cout, int K, int stride, int relu, rcmvs_stream_t stream);

/* Same layer, additionally (or only: y may be NULL) writing the result as split-bf16 planes y_split [2][N][Ho][Wo][Cout] for a
 * tensor-core consumer (rcmvs_conv2d_slab_fwd). */
int rcmvs_conv2d_fwd_split(const float* x, const float* packed_w, const float* bias, const float* up, float* y, void* y_split,
                           int N, int H, int W, int Cin, int Cout, int K, int stride, int relu, rcmvs_stream_t stream);
/* FeatureNet's 3x3 stride-1 convolutions on tcgen05 (kind::f16, bf16x3) fed by TMA, persistent over (image, tile) items:
 * x_split [2][N][H][W][Cin] bf16 (hi, lo) -> y fp32 [N][H][W][Cout] and/or y_split; bias = folded BatchNorm shift or NULL.
 * (Cin, Cout) in {(8,8), (16,16), (32,32), (32,16), (32,8)}: conv0.0/0.1 (the 3 image channels padded to 8; two kw taps per
 * K = 16 step), conv1.1/1.2, conv2.1/2.2, out2, out3 (models/modules.py:391-417). */
int64_t rcmvs_conv2d_slab_weight_elems(int Cin, int Cout);
int rcmvs_conv2d_slab_pack(const float* w, const float* scale, void* img, int Cout, int Cin, rcmvs_stream_t stream);
int rcmvs_conv2d_slab_fwd(const void* x_split, const void* wimg, const float* bias, float* y, void* y_split, int N, int H, int W,
                          int Cin, int Cout, int relu, rcmvs_stream_t stream);

/* FeatureNet's two down-sampling convolutions on the same tensor-core family: conv1.0 = Conv2d(8, 16, 5, stride 2, pad 2)
 * and conv2.0 = Conv2d(16, 32, 5, stride 2, pad 2) + folded BN + ReLU (reference models/modules.py:394-395, 399-400).
 * Parity sub-planes by strided TMA boxes, all 25 taps as descriptor offsets (two taps per K = 16 step for 8 channels).
 * x_split [2][N][H][W][Cin] (H, W even) -> y [N][H/2][W/2][Cout] fp32 and / or y_split.  (Cin, Cout) in {(8,16), (16,32)}. */
int64_t rcmvs_conv2d_s2k5_slab_weight_elems(int Cin, int Cout);
int rcmvs_conv2d_s2k5_slab_pack(const float* w, const float* scale, void* img, int Cout, int Cin, rcmvs_stream_t stream);
int rcmvs_conv2d_s2k5_slab_fwd(const void* x_split, const void* wimg, const float* bias, float* y, void* y_split, int N,
                               int H, int W, int Cin, int Cout, int relu, rcmvs_stream_t stream);

/* ---- a7 Neural_Volume_Net resize: models/render_models.py:755-756 ------------------------------------------------ */
/* x [B][D0][S][C] channels-last (S = h*w) -> y [B][D1][S][Cpad]: linear along D with align_corners=True, channels
 * C..Cpad-1 zero-filled (the 41-channel volume is padded to 48 for the convolution kernels). */
int rcmvs_resize_depth_cl(const float* x, float* y, int B, int D0, int D1, int64_t S, int C, int Cpad,
                          rcmvs_stream_t stream);

/* Same, written as split-bf16 planes y_split [2][B][D1][S][Cpad] for rcmvs_conv3d_slab_fwd (Cin = 48, Cout = 8). */
int rcmvs_resize_depth_cl_split(const float* x, void* y_split, int B, int D0, int D1, int64_t S, int C, int Cpad,
                                rcmvs_stream_t stream);

/* ---- a8 Gaussian-Uniform sampler + rays: models/render_utils.py:86-146, 149-243 ------------------------------------ */
/* ray_depth [N] -> out [N][S] (S = 128): rays [0,n_gauss) ~ N(d, (min(|far-d|,|d-near|)/3)^2) sorted ascending,
 * rays [n_gauss,N) stratified-uniform in [near,far].  Philox4x32-10 keyed by (seed, ray, lane). */
int rcmvs_gu_sample(const float* ray_depth, float near, float far, uint64_t seed, float* out, int N, int n_gauss,
                    int S, rcmvs_stream_t stream);
/* ys, xs [N] pixel coordinates (float), K [3][3], c2w [4][4], w2c_ref [4][4], z [N][S] ->
 * pts [N][S][3] = o + z*dir, ndc [N][S][3] = (pixel/(W-1,H-1), (z_cam-near)/(far-near)), dirs [N][3] (unnormalised). */
int rcmvs_ray_points(const float* ys, const float* xs, const float* K, const float* c2w, const float* w2c_ref,
                     const float* z, float near, float far, int W, int H, float* pts, float* ndc, float* dirs, int N,
                     int S, rcmvs_stream_t stream);

/* ---- a9 per-point features: models/render_utils.py:247-279, 304-330; models/renderer.py:154-166 --------------------- */
/* vol_cl [D][h][w][8] channels-last, ndc/pts [P][3], imgs [3][3][H][W] (un-normalised source images),
 * w2cs [3][4][4], Ks [3][3][3] (poses 0..2) -> out [P][20] = [8 trilinear (zeros padding), 3 x (RGB border-padded
 * bilinear, in-bounds mask)]. */
int rcmvs_point_features(const float* vol_cl, const float* ndc, const float* pts, const float* imgs,
                         const float* w2cs, const float* Ks, float* out, int64_t P, int D, int h, int w, int H, int W,
                         rcmvs_stream_t stream);

/* ---- a10 embedding + Renderer_ours MLP: models/render_models.py:45-49, 192-220 --------------------------------------- */
/* ndc [P][3], feat [P][20], dirs [P/S][3] (view direction per ray, reference-camera frame) -> out [P][4] = rgb, sigma.
 * weights: 22 device pointers {pts_linears.0..5 (W^T [K][128], b), pts_bias (W^T [20][128], b), alpha_linear
 * (w [128], b), feature_linear (W^T [128][128], b), views_linears.0 (W^T [131][64], b), rgb_linear (W^T [64][3], b)}. */
int rcmvs_render_mlp(const float* ndc, const float* feat, const float* dirs, const float* const* weights, float* out,
                     int64_t P, int S, rcmvs_stream_t stream);

/* Tensor-core variant of the same MLP (same inputs, outputs and the same 22 fp32 weight pointers): every 128 / 64-wide Linear
 * layer runs as a TMA-fed tcgen05 GEMM on split-bf16 operands (3 products per operand pair, fp32 accumulation), the
 * embedding / modulation / alpha / rgb heads are fused into a prep kernel and the GEMM epilogues.
 * wimg: rcmvs_render_mlp_tc_weight_elems() bf16 elements filled by rcmvs_render_mlp_tc_pack (once per checkpoint);
 * workspace: rcmvs_render_mlp_tc_workspace_bytes(P) bytes of device memory, 256-byte aligned (activations between layers). */
int64_t rcmvs_render_mlp_tc_weight_elems(void);
int64_t rcmvs_render_mlp_tc_workspace_bytes(int64_t P);
int rcmvs_render_mlp_tc_pack(const float* const* weights, void* wimg, rcmvs_stream_t stream);
int rcmvs_render_mlp_tc(const float* ndc, const float* feat, const float* dirs, const float* const* weights, const void* wimg,
                        void* workspace, float* out, int64_t P, int S, rcmvs_stream_t stream);

/* ---- a11 compositing: models/renderer.py:18-26, 65-93 ------------------------------------------------------------------ */
/* raw [N][S][4], z [N][S] -> rgb [N][3], depth [N], weights [N][S], alpha [N][S]. */
int rcmvs_composite(const float* raw, const float* z, float* rgb, float* depth, float* weights, float* alpha, int N,
                    int S, rcmvs_stream_t stream);

/* ---- a13 inverse_warping: losses/homography.py:6-200 ------------------------------------------------------------------- */
/* img [B][h][w][C] (NHWC), cams [B][2][4][4], depth [B][h][w] -> warped [B][h][w][C], mask [B][h][w]. */
int rcmvs_inverse_warp_fwd(const float* img, const float* left_cam, const float* right_cam, const float* depth,
                           float* warped, float* mask, int B, int h, int w, int C, rcmvs_stream_t stream);
/* gradient w.r.t. depth: grad_warped [B][h][w][C] -> grad_depth [B][h][w]. */
int rcmvs_inverse_warp_bwd(const float* img, const float* left_cam, const float* right_cam, const float* depth,
                           const float* grad_warped, float* grad_depth, int B, int h, int w, int C,
                           rcmvs_stream_t stream);

/* ---- f2 unsupervised loss stack: losses/unsup_loss.py:9-94, losses/modules.py:6-82, losses/aug_loss.py, sl1loss.py ---- */
/* imgs [N][3][H][W] -> out [N][H/step][W/step][3]: F.interpolate(scale_factor=1/step, nearest) + NCHW->NHWC
 * (unsup_loss.py:27-33, 55-61). */
int rcmvs_images_subsample_nhwc(const float* imgs, float* out, int N, int H, int W, int step, rcmvs_stream_t stream);
/* UnSupLoss.forward of one stage, value AND gradient w.r.t. depth in one call: imgs [B][V][h][w][3] (stage-resolution
 * NHWC images, view 0 = reference), cams [B][V][2][4][4], depth [B][h][w] -> terms[4] = {12·reconstr + 6·ssim +
 * smooth_weight·smooth, reconstr, ssim, smooth} (device), grad_depth [B][h][w] = d terms[0] / d depth (may be NULL).
 * workspace: rcmvs_unsup_loss_workspace_floats(B, V, h, w) floats. */
int64_t rcmvs_unsup_loss_workspace_floats(int B, int V, int h, int w);
int rcmvs_unsup_loss_fwd_bwd(const float* imgs, const float* cams, const float* depth, float smooth_weight,
                             float* terms, float* grad_depth, float* workspace, int B, int V, int h, int w,
                             rcmvs_stream_t stream);
/* scale · mean over the selected elements of smooth-L1 (mode 0) or squared error (mode 1) of pred [B][h][w] against
 * target [B][h·step][w·step] read at (y·step, x·step), + gradient w.r.t. pred (may be NULL).  mask_kind 0: all elements;
 * 1: mask[b·mask_batch_stride + (y·step)·(w·step) + x·step] > 0.5; 2: target > 0.  AugLossMultiStage (aug_loss.py:41-58),
 * SL1Loss (sl1loss.py:9-13), img2mse.  workspace: rcmvs_masked_loss_workspace_floats() floats. */
int64_t rcmvs_masked_loss_workspace_floats(void);
int rcmvs_masked_loss_fwd_bwd(const float* pred, const float* target, const float* mask, int64_t mask_batch_stride,
                              int mask_kind, int mode, float scale, float* loss, float* grad, float* workspace, int B,
                              int h, int w, int step, rcmvs_stream_t stream);

/* ---- training path (BASELINE configs[3]; train_rcmvsnet.py:330-376, 397-445, 311): forward + backward of every
 * conv -> BatchNorm(batch statistics) -> ReLU block (models/modules.py:28-62, 118-210), of the variance cost volume and of
 * the soft-argmin, on own kernels.  All activations channels-last [M][C] / [B][D][H][W][C] / [N][H][W][C]. ---------------- */
/* stats [2][C] double (zero-filled by the caller) += {Σ x, Σ x²} over the M rows of x [M][C]. */
int rcmvs_bn_stats(const float* x, double* stats, int64_t M, int C, rcmvs_stream_t stream);
/* mean, invstd [C] from stats; running_mean / running_var (may be NULL) updated with `momentum` (unbiased variance). */
int rcmvs_bn_finalize(const double* stats, float* mean, float* invstd, float* running_mean, float* running_var, int64_t M,
                      int C, float eps, float momentum, rcmvs_stream_t stream);
/* z = act((x - mean)·invstd·gamma + beta) (+ skip, added after the ReLU: modules.py:497-499). */
int rcmvs_bn_apply(const float* x, const float* mean, const float* invstd, const float* gamma, const float* beta,
                   const float* skip, float* z, int64_t M, int C, int relu, rcmvs_stream_t stream);
/* BatchNorm(+ReLU) backward: dz, x (the convolution output the forward normalised) -> dx (gradient w.r.t. x), dgamma,
 * dbeta [C]; sums = 2·C doubles of zero-filled workspace.  frozen = 1: mean / invstd were running statistics (eval-mode
 * BatchNorm under autograd), so the batch-statistic correction terms are dropped; frozen bit 1 (value 2): dgamma / dbeta
 * are accumulated into (+=) instead of overwritten (gradient buffers of an optimiser, cf. rcmvs_conv_wgrad's G). */
int rcmvs_bn_bwd(const float* dz, const float* x, const float* mean, const float* invstd, const float* gamma,
                 const float* beta, double* sums, float* dx, float* dgamma, float* dbeta, int64_t M, int C, int relu,
                 int frozen, rcmvs_stream_t stream);
/* The same in two calls (SyncBatchNorm: all-reduce `sums` over the ranks in between; M_total = rows over all ranks). */
int rcmvs_bn_bwd_reduce(const float* dz, const float* x, const float* mean, const float* invstd, const float* gamma,
                        const float* beta, double* sums, int64_t M, int C, int relu, rcmvs_stream_t stream);
int rcmvs_bn_bwd_apply(const float* dz, const float* x, const float* mean, const float* invstd, const float* gamma,
                       const float* beta, const double* sums, float* dx, float* dgamma, float* dbeta, int64_t M,
                       int64_t M_total, int C, int relu, int frozen, rcmvs_stream_t stream);
/* Weight gradient G[cs][cb][taps] += Σ_o S[o][cs]·Bg[stride·o + tap - pad][cb] (fp32 atomics: G holds zeros for a plain
 * gradient, or an optimiser's gradient buffer to accumulate into; torch weight layout): Conv: S = dy [B][Ds][Hs][Ws][Cs=Cout], Bg = x [B][Db][Hb][Wb][Cb=Cin] -> dW [Cout][Cin][taps];
 * ConvTranspose(k3,s2,p1,op1): S = x (Cs = Cin), Bg = dy (Cb = Cout), stride 2 -> dW [Cin][Cout][27].
 * dim3d = 1: 3x3x3; dim3d = 0: 2-D (Ds = Db = 1), k in {1,3,5}. */
int rcmvs_conv_wgrad(const float* S, const float* Bg, float* G, int B, int Ds, int Hs, int Ws, int Cs, int Db, int Hb,
                     int Wb, int Cb, int k, int stride, int dim3d, rcmvs_stream_t stream);
/* 2-D data gradient: dy [N][Ho][Wo][Cout], torch weight w [Cout][Cin][k][k] -> dx [N][H][W][Cin] (Cin % 4 == 0: the
 * 3-channel image layer needs none); wd_ws = k·k·Cout·Cin floats of workspace. */
int rcmvs_conv2d_dgrad(const float* dy, const float* w, float* wd_ws, float* dx, int N, int H, int W, int Cin, int Cout,
                       int k, int stride, rcmvs_stream_t stream);
/* 3-D data gradients reuse the forward kernels: stride-1 conv -> rcmvs_conv3d_fwd with weights packed by
 * rcmvs_pack_conv3d_dgrad_s1 ([27][Cout][Cin], taps flipped); stride-2 conv -> rcmvs_conv_transpose3d_fwd with
 * rcmvs_pack_conv3d_weight(w, transposed = 1); transposed conv -> rcmvs_conv3d_fwd stride 2 with
 * rcmvs_pack_conv3d_weight(w, transposed = 0).  The 1-output-channel `prob` layer: */
int rcmvs_pack_conv3d_dgrad_s1(const float* w, float* packed, int Cout, int Cin, rcmvs_stream_t stream);
int rcmvs_conv3d_c1_dgrad(const float* dcost, const float* w, float* dx, int B, int D, int H, int W, rcmvs_stream_t stream);
/* Backward of rcmvs_costvolume_var_fwd: grad_var (and grad_noref, may be NULL) [B][D][h][w][C] -> grad_feats
 * [B][V][h][w][C] (zero-filled by the caller; vector reductions).  At most 7 source views. */
int rcmvs_costvolume_var_bwd(const float* feats, const float* rt, const float* depth, int per_pixel, const float* grad_var,
                             const float* grad_noref, float* grad_feats, int B, int V, int C, int D, int h, int w,
                             rcmvs_stream_t stream);
/* Backward of softmax + depth regression: grad_cost[d] = grad_depth · p_d · (plane_d - depth). */
int rcmvs_softargmin_bwd(const float* cost, const float* depth_planes, int per_pixel, const float* grad_depth,
                         float* grad_cost, int B, int D, int h, int w, rcmvs_stream_t stream);
/* FPN lateral (modules.py:451,455): out = a + nearest-2x-upsample(up); its backward w.r.t. up = 2x2 block sums. */
int rcmvs_up2_add(const float* a, const float* up, float* out, int N, int H, int W, int C, rcmvs_stream_t stream);
int rcmvs_sum2x2(const float* in, float* out, int N, int Ho, int Wo, int C, rcmvs_stream_t stream);
/* Fused Adam (torch.optim.Adam semantics, train_rcmvsnet.py:537) over flat buffers; grads are multiplied by grad_scale
 * first (1/world_size after the NCCL sum all-reduce). */
int rcmvs_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                    float beta2, float eps, float weight_decay, int step, float grad_scale, rcmvs_stream_t stream);

/* ---- training path of the rendering-consistency branch (train_rcmvsnet.py:279-325; models/render_models.py:192-220,
 * models/renderer.py:65-93, models/render_utils.py:304-330) ---------------------------------------------------------------- */
/* C[M][N] (+)= op(A)[M][K]·op(B)[K][N] (+ bias[N]): trans_a: A stored [K][M] (lda = M-stride); trans_b: B stored [N][K].
 * splits > 1: split-K with atomic accumulation, C zero-filled by the caller.  The MLP's Linear layers: forward
 * (X, W^T: trans_b = 1), data gradient (dY, W), weight gradient (dY^T, X: trans_a = 1, split-K). */
int rcmvs_sgemm(const float* A, const float* B, const float* bias, float* C, int M, int N, int K, int lda, int ldb, int ldc,
                int trans_a, int trans_b, int splits, rcmvs_stream_t stream);
/* mode 0: h = relu(a·b) (multiplicative bias modulation, render_models.py:200); 1: relu(a); 2: sigmoid(a) — and backward
 * (h = the forward output; db only for mode 0). */
int rcmvs_act_fwd(const float* a, const float* b, float* h, int64_t n, int mode, rcmvs_stream_t stream);
int rcmvs_act_bwd(const float* g, const float* a, const float* b, const float* h, float* da, float* db, int64_t n, int mode,
                  rcmvs_stream_t stream);
/* x [P][3] -> out[p][0..63) (row pitch ld): [x, sin(2^k x) k = 0..9, cos(2^k x)] (render_models.py:45-49). */
int rcmvs_embed(const float* x, float* out, int64_t P, int ld, rcmvs_stream_t stream);
/* grad_feat [P][ld] (first 8 columns = the trilinear volume features) -> grad_vol [D][h][w][8] (zero-filled by the caller). */
int rcmvs_point_features_bwd(const float* ndc, const float* grad_feat, float* grad_vol, int64_t P, int ld, int D, int h, int w,
                             rcmvs_stream_t stream);
/* raw [N][S][4], z [N][S], grad_rgb [N][3], grad_depth [N], grad_weights [N][S] (each may be NULL) -> grad_raw [N][S][4]. */
int rcmvs_composite_bwd(const float* raw, const float* z, const float* grad_rgb, const float* grad_depth,
                        const float* grad_weights, float* grad_raw, int N, int S, rcmvs_stream_t stream);
/* adjoint of rcmvs_resize_depth_cl: grad_y [B][D1][S][Cpad] -> grad_x [B][D0][S][C]. */
int rcmvs_resize_depth_cl_bwd(const float* grad_y, float* grad_x, int B, int D0, int D1, int64_t S, int C, int Cpad,
                              rcmvs_stream_t stream);

/* ---- f3 geometric-consistency filter: eval_rcmvsnet_dtu.py:281-336 (reproject_with_depth,
 * check_geometric_consistency) + :366-389 of filter_depth, all NS source views of one reference view per launch ---- */
/* depth_ref [H][W], conf [H][W] or NULL, depth_src [NS][Hs][Ws]; pair_mats [NS][68] float32 per source:
 * {inv(K_ref) 9, E_src·inv(E_ref) 16, K_src 9, inv(K_src) 9, E_ref·inv(E_src) 16, K_ref 9} (float32 on the host, as the
 * reference computes them).  Outputs (each may be NULL): masks uint8 [NS][H][W], depth_reproj [NS][H][W] (0 where the
 * mask is 0), xy_src [NS][2][H][W] (float32 source coordinates), geo_mask_sum int32 [H][W], depth_avg [H][W] =
 * (Σ depth_reproj + depth_ref) / (geo_mask_sum + 1), final_mask uint8 [H][W] = conf > prob_thresh & sum >= num_consistent. */
int rcmvs_geo_consistency(const float* depth_ref, const float* conf, const float* depth_src, const float* pair_mats,
                          int NS, int H, int W, int Hs, int Ws, float img_dist_thresh, float depth_thresh,
                          float prob_thresh, int num_consistent, uint8_t* masks, float* depth_reproj, float* xy_src,
                          int32_t* geo_mask_sum, float* depth_avg, uint8_t* final_mask, rcmvs_stream_t stream);

/* ---- f4 wire format: PFM depth / confidence maps (datasets/data_io.py:8-73; eval_rcmvsnet_dtu.py:199-253) — HOST functions,
 * host pointers.  Rows in memory are top-down (what read_pfm returns / save_pfm takes); the file stores them bottom-up. ------ */
int rcmvs_pfm_read_header(const char* path, int* width, int* height, int* channels, float* scale);
int rcmvs_pfm_read(const char* path, float* out, int height, int width, int channels);
int rcmvs_pfm_write(const char* path, const float* img, int height, int width, int channels, float scale);

#ifdef __cplusplus
}
#endif
#endif /* RCMVS_H_ */
