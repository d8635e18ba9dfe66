"""Drop-in for the reference's models/modules.py on the depth-inference hot path.

Same public names, constructor arguments, state_dict keys and call signatures as the reference
(models/modules.py: homo_warping :304, CostRegNet :470, FeatureNet :363, depth_regression :519,
get_depth_range_samples :569, get_cur_depth_range_samples :549), but the bodies enqueue the sm_100a kernels of
librcmvs.so through rcmvs.ops.  CUDA tensors only — there is no CPU fallback.
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from rcmvs import ops
from rcmvs.cache import PackedCache
from rcmvs.strict import library_fallback

__all__ = ["homo_warping", "CostRegNet", "FeatureNet", "Conv2d", "Conv3d", "Deconv3d", "depth_regression",
           "get_depth_range_samples", "get_cur_depth_range_samples", "cas_mvsnet_loss", "fold_bn"]


def conv_precision():
    """RCMVS_CONV selects how CostRegNet's convolutions run:
         'bf16x3' (default) conv0 / conv2 on the TMA-fed tcgen05 kind::f16 kernel with error-compensated split-bf16
                  operands (3 products, operand error <= 3·2^-17: 1.7e-5 relative on the full-size golden depth), the other
                  tensor-core layers as in 'tf32x3';
         'tf32x3' error-compensated 3xTF32 on tcgen05 tensor cores — fp32-equivalent accuracy (round-1 default);
         'tf32'   single TF32 pass on tcgen05 (the precision cuDNN gives the reference on GPU by torch default);
         'fp32'   exact fp32 FFMA direct-convolution kernels (no tensor cores)."""
    import os
    mode = os.environ.get("RCMVS_CONV", "bf16x3")
    if mode not in ("bf16x3", "tf32x3", "tf32", "fp32"):
        raise ValueError("RCMVS_CONV must be bf16x3, tf32x3, tf32 or fp32")
    return mode


def fold_bn(bn):
    """Eval-mode BatchNorm as y = x*scale + shift."""
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    return scale, bn.bias - bn.running_mean * scale


def _versions(mod):
    return tuple((p.data_ptr(), p._version) for p in list(mod.parameters()) + list(mod.buffers()))


# ---------------------------------------------------------------------------------------------------------------
# a1  homo_warping
# ---------------------------------------------------------------------------------------------------------------
class _HomoWarp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src_fea, rt, depth_values):
        ctx.save_for_backward(rt, depth_values)
        ctx.hw = src_fea.shape[2:]
        return ops.homo_warp_fwd(src_fea, rt, depth_values)

    @staticmethod
    def backward(ctx, grad_out):
        rt, depth_values = ctx.saved_tensors
        # the sampling grid is built under torch.no_grad() in the reference: gradient reaches src_fea only
        return ops.homo_warp_bwd(grad_out.contiguous(), rt, depth_values, *ctx.hw), None, None


def homo_warping(src_fea, src_proj, ref_proj, depth_values):
    """src_fea [B,C,H,W], src_proj/ref_proj [B,4,4], depth_values [B,D] or [B,D,H,W] -> [B,C,D,H,W]."""
    with torch.no_grad():
        rt = ops.relative_projection_pair(src_proj.float().contiguous(), ref_proj.float().contiguous())
        depth_values = depth_values.detach().float().contiguous()
    return _HomoWarp.apply(src_fea.float().contiguous(), rt, depth_values)


# ---------------------------------------------------------------------------------------------------------------
# 2-D / 3-D conv blocks (state_dict layout: <name>.conv.weight, <name>.bn.{weight,bias,running_mean,running_var,
# num_batches_tracked})
# ---------------------------------------------------------------------------------------------------------------
class _ConvNormAct(nn.Module):
    conv_cls = None
    bn_cls = None

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, relu=True, bn=True, bn_momentum=0.1,
                 **kwargs):
        super().__init__()
        kwargs.pop("init_method", None)
        kwargs.pop("norm", None)
        kwargs.pop("num_groups", None)
        assert stride in (1, 2)
        self.out_channels, self.kernel_size, self.stride, self.relu = out_channels, kernel_size, stride, relu
        self.conv = self.conv_cls(in_channels, out_channels, kernel_size, stride=stride, bias=not bn, **kwargs)
        self.bn = self.bn_cls(out_channels, momentum=bn_momentum) if bn else None

    def forward(self, x):
        y = self.conv(x)
        if self.bn is not None:
            y = self.bn(y)
        return F.relu(y, inplace=True) if self.relu else y


class Conv2d(_ConvNormAct):
    conv_cls, bn_cls = nn.Conv2d, nn.BatchNorm2d


class Conv3d(_ConvNormAct):
    conv_cls, bn_cls = nn.Conv3d, nn.BatchNorm3d


class Deconv3d(_ConvNormAct):
    conv_cls, bn_cls = nn.ConvTranspose3d, nn.BatchNorm3d


# ---------------------------------------------------------------------------------------------------------------
# FeatureNet (fpn) — SURVEY §8 row f1: hand-written fp32 conv2d kernels (csrc/conv2d_direct.cu), run ONCE on all V
# views as a batch with BatchNorm folded, emitting channels-last maps the sweep kernel consumes directly.
# ---------------------------------------------------------------------------------------------------------------
class FeatureNet(nn.Module):
    def __init__(self, base_channels, num_stage=3, stride=4, arch_mode="fpn"):
        super().__init__()
        if arch_mode != "fpn":
            raise NotImplementedError("rcmvs FeatureNet implements arch_mode='fpn' (the reference default)")
        assert num_stage in (2, 3)
        c = base_channels
        self.arch_mode, self.stride, self.base_channels, self.num_stage = arch_mode, stride, c, num_stage

        def chain(*specs):
            return nn.Sequential(*[Conv2d(i, o, k, s, padding=k // 2) for (i, o, k, s) in specs])

        self.conv0 = chain((3, c, 3, 1), (c, c, 3, 1))
        self.conv1 = chain((c, 2 * c, 5, 2), (2 * c, 2 * c, 3, 1), (2 * c, 2 * c, 3, 1))
        self.conv2 = chain((2 * c, 4 * c, 5, 2), (4 * c, 4 * c, 3, 1), (4 * c, 4 * c, 3, 1))
        self.out1 = nn.Conv2d(4 * c, 4 * c, 1, bias=False)
        self.inner1 = nn.Conv2d(2 * c, 4 * c, 1, bias=True)
        self.out_channels = [4 * c]
        if num_stage == 3:
            self.inner2 = nn.Conv2d(c, 4 * c, 1, bias=True)
            self.out2 = nn.Conv2d(4 * c, 2 * c, 3, padding=1, bias=False)
            self.out3 = nn.Conv2d(4 * c, c, 3, padding=1, bias=False)
            self.out_channels += [2 * c, c]
        else:
            self.out2 = nn.Conv2d(4 * c, c, 3, padding=1, bias=False)
            self.out_channels += [c]
        self._folded = PackedCache()

    def _packed(self):
        """BN-folded, re-laid-out weights for the librcmvs conv2d kernels, rebuilt when parameters change (stream-safe:
        a hit on another CUDA stream waits for the packing kernels, rcmvs/cache.py)."""
        def build():
            with torch.no_grad():
                P = {}
                for name in ("conv0", "conv1", "conv2"):
                    for i, blk in enumerate(getattr(self, name)):
                        scale, shift = (t.contiguous() for t in fold_bn(blk.bn))
                        w = blk.conv.weight.detach().contiguous()
                        cin_pad = (w.shape[1] + 7) // 8 * 8
                        P["%s.%d" % (name, i)] = (ops.pack_conv2d_weight(w, scale, cin_pad), shift, blk.stride)
                        wp = F.pad(w, (0, 0, 0, 0, 0, cin_pad - w.shape[1])) if cin_pad != w.shape[1] else w   # 3 -> 8 input channels
                        if tc and blk.stride == 1 and w.shape[2] == 3 and name != "conv0" and \
                                (cin_pad, w.shape[0]) in ops.SLAB2D_SUPPORTED:
                            P["%s.%d.slab" % (name, i)] = (ops.conv2d_slab_pack(wp.contiguous(), scale), shift)
                        if tc and blk.stride == 2 and w.shape[2] == 5 and (w.shape[1], w.shape[0]) in ops.S2K5_SUPPORTED:
                            P["%s.%d.s2slab" % (name, i)] = (ops.conv2d_s2k5_slab_pack(w, scale), shift)
                for name in ("out1", "inner1", "inner2", "out2", "out3"):
                    conv = getattr(self, name, None)
                    if conv is not None:
                        b = conv.bias.detach().contiguous() if conv.bias is not None else None
                        w = conv.weight.detach().contiguous()
                        P[name] = (ops.pack_conv2d_weight(w, None), b, 1)
                        if tc and w.shape[2] == 3 and (w.shape[1], w.shape[0]) in ops.SLAB2D_SUPPORTED:
                            P[name + ".slab"] = (ops.conv2d_slab_pack(w, None), b)
            return P
        tc = conv_precision() == "bf16x3" and self.base_channels == 8
        return self._folded.get((_versions(self), tc), build)

    def forward_cl(self, x, side_stream=None):
        """x [N,3,H,W] -> {"stageK": [N,h,w,C] channels-last contiguous}; every layer is one librcmvs launch
        (conv + folded BN + ReLU; the FPN's upsample+add is fused into the 1x1 lateral convolution).

        With `side_stream` the FPN tail (the stage-2/3 feature maps, ~0.5 ms of FFMA work nobody needs before stage 2)
        is enqueued on that stream and the result carries `"_ready"`, a CUDA event the consumer's stream must wait for
        before it touches `stage2` / `stage3`; the caller's stream goes straight on with stage 1."""
        P = self._packed()
        tc = "conv1.1.slab" in P          # bf16x3 mode: the 3x3 stride-1 layers with 16 / 32 input channels run on tcgen05

        def cbr(t, key, split_out=False):
            """conv + folded BN + ReLU.  `t` fp32 [N,H,W,C] or split-bf16 [2,N,H,W,C]; returns fp32, or the split form when
            the consumer is a tensor-core layer (`split_out`)."""
            if t.dtype == torch.bfloat16:
                img, shift = P[key + ".slab"]
                return ops.conv2d_slab(t, img, img_cout[key], shift, True, want_y=not split_out, want_split=split_out)
            w, b, stride = P[key]
            return ops.conv2d(t, w, b, None, stride=stride, relu=True, want_y=not split_out, want_split=split_out)

        img_cout = {"conv1.1": 16, "conv1.2": 16, "conv2.1": 32, "conv2.2": 32}

        def head(t, key, cout):
            """plain 3x3 convolution (out2 / out3)."""
            if t.dtype == torch.bfloat16:
                return ops.conv2d_slab(t, P[key + ".slab"][0], cout, None, False)
            return ops.conv2d(t, P[key][0], None, None, relu=False)

        def tail(out):
            # up2(c2) + inner1(c1): fp32 for the next lateral, split-bf16 for the tensor-core out2
            if tc:
                intra, intra_s = ops.conv2d(c1, P["inner1"][0], P["inner1"][1], c2, relu=False, want_split=True)
            else:
                intra = intra_s = ops.conv2d(c1, P["inner1"][0], P["inner1"][1], c2, relu=False)
            out["stage2"] = head(intra_s, "out2", 16)
            if self.num_stage == 3:
                intra2 = ops.conv2d(c0, P["inner2"][0], P["inner2"][1], intra, relu=False, want_y=not tc, want_split=tc)
                out["stage3"] = head(intra2, "out3", 8)

        # the 5x5 stride-2 layers run on the tensor cores too when their input has even height and width: their producers
        # then hand over the activation in both forms (fp32 for the FPN lateral, split-bf16 for the tensor-core consumer)
        s2 = tc and "conv1.0.s2slab" in P and "conv2.0.s2slab" in P and x.shape[-2] % 4 == 0 and x.shape[-1] % 4 == 0

        def down(ts, key, cout):
            img, shift = P[key + ".s2slab"]
            return ops.conv2d_s2k5_slab(ts, img, cout, shift, True, want_y=False, want_split=True)

        # conv0.0 / conv0.1 stay on the exact-fp32 FFMA kernel: their rounding is amplified by everything downstream, and the
        # split-bf16 scheme there (measured) pushed the ill-conditioned stage 3 of the 7 x 1056x1920 case from 6e-4 to 1.2e-3
        c0 = cbr(ops.images_to_nhwc8(x.float().contiguous()), "conv0.0")
        if s2:
            w, b, stride = P["conv0.1"]
            c0, c0s = ops.conv2d(c0, w, b, None, stride=stride, relu=True, want_y=True, want_split=True)
        if s2:
            img, shift = P["conv1.2.slab"]
            c1, c1s = ops.conv2d_slab(cbr(down(c0s, "conv1.0", 16), "conv1.1", True), img, 16, shift, True, want_y=True,
                                      want_split=True)
            c2 = cbr(cbr(down(c1s, "conv2.0", 32), "conv2.1", True), "conv2.2")
        else:
            c0 = cbr(c0, "conv0.1")
            c1 = cbr(cbr(cbr(c0, "conv1.0", tc), "conv1.1", tc), "conv1.2")
            c2 = cbr(cbr(cbr(c1, "conv2.0", tc), "conv2.1", tc), "conv2.2")
        out = {}
        if side_stream is None:
            out["stage1"] = ops.conv2d(c2, P["out1"][0], None, None, relu=False)
            tail(out)
            return out
        main = torch.cuda.current_stream(x.device)
        fork = torch.cuda.Event()
        fork.record(main)
        out["stage1"] = ops.conv2d(c2, P["out1"][0], None, None, relu=False)
        side_stream.wait_event(fork)
        with torch.cuda.stream(side_stream):
            tail(out)
            ready = torch.cuda.Event()
            ready.record(side_stream)
        for t_ in (c0, c1, c2):
            t_.record_stream(side_stream)          # read over there: do not recycle before the tail has run
        for k in ("stage2", "stage3"):
            if k in out:
                out[k].record_stream(main)         # allocated over there, consumed on the caller's stream
        out["_ready"] = ready
        return out

    def forward(self, x):
        """x [N,3,H,W] -> {"stage1": [N,4c,H/4,W/4], "stage2": [N,2c,H/2,W/2], "stage3": [N,c,H,W]}.
        Eval / no-grad: hand-written fp32 kernels, outputs are NCHW-shaped views of channels-last storage.
        Training (batch-statistics BatchNorm + autograd): torch library ops."""
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))
        if not self.training and not needs_grad and x.is_cuda and self.base_channels == 8:
            with torch.no_grad():
                return {k: v.permute(0, 3, 1, 2) for k, v in self.forward_cl(x).items()}
        if not x.is_cuda or self.base_channels != 8:
            library_fallback("FeatureNet.forward (CPU tensor / non-default width)", "torch conv2d + batch_norm")
            return self._forward_torch(x)
        return {k: v.permute(0, 3, 1, 2) for k, v in self.forward_cl_train(x).items()}

    def forward_cl_train(self, x):
        """Differentiable forward on own kernels (rcmvs.autograd): batch-statistics BatchNorm when self.training, frozen
        running statistics otherwise; every conv / BN / ReLU / FPN-add has a hand-written backward.  Channels-last maps."""
        from rcmvs import autograd as A
        tr = self.training

        def chain(seq, t):
            for blk in seq:
                t = A.conv_block(t, blk.conv, blk.bn, relu=True, training=tr)
            return t
        c0 = chain(self.conv0, ops.images_to_nhwc8(x.float().contiguous()))
        c1 = chain(self.conv1, c0)
        c2 = chain(self.conv2, c1)
        out = {"stage1": A.conv_block(c2, self.out1, None, relu=False)}
        intra = A.up2_add(A.conv_block(c1, self.inner1, None, relu=False), c2)
        out["stage2"] = A.conv_block(intra, self.out2, None, relu=False)
        if self.num_stage == 3:
            intra = A.up2_add(A.conv_block(c0, self.inner2, None, relu=False), intra)
            out["stage3"] = A.conv_block(intra, self.out3, None, relu=False)
        return out

    def _forward_torch(self, x):
        c0 = self.conv0(x)
        c1 = self.conv1(c0)
        c2 = self.conv2(c1)
        out = {"stage1": self.out1(c2)}
        intra = F.interpolate(c2, scale_factor=2, mode="nearest") + self.inner1(c1)
        out["stage2"] = self.out2(intra)
        if self.num_stage == 3:
            intra = F.interpolate(intra, scale_factor=2, mode="nearest") + self.inner2(c0)
            out["stage3"] = self.out3(intra)
        return out


def as_channels_last_views(feat):
    """[N,C,h,w] (any memory format) -> contiguous [N,h,w,C] tensor (zero-copy if already channels_last)."""
    p = feat.permute(0, 2, 3, 1)
    return p if p.is_contiguous() else ops.to_channels_last(feat.contiguous())


def pack_conv3d_layer(w, scale, shift, transposed, stride, relu, nsplit, slab=False):
    """One 3x3x3 (transposed) convolution + folded BatchNorm as a list of candidate engines (`CostRegNet._run` picks
    one per input shape): the FFMA kernel always; with nsplit = 1 (TF32) or 3 (3xTF32) also the tcgen05 kernels that
    support the layer's shape (slab kernel `tc` / `tct`, gather implicit-GEMM `gemm`)."""
    cout, cin = (w.shape[1], w.shape[0]) if transposed else (w.shape[0], w.shape[1])
    cands = [("direct", ops.pack_conv3d_weight(w, scale, transposed), shift)]
    if nsplit > 0 and transposed and (cin, cout) in ops.TCT_SUPPORTED:
        cands.append(("tct", ops.conv_transpose3d_tc_pack(w, scale, nsplit), shift, cout, nsplit))
    if nsplit > 0 and not transposed and stride == 1 and (cin, cout) in ops.TC_SUPPORTED:
        cands.append(("tc", ops.conv3d_tc_pack(w, scale, nsplit), shift, cout, nsplit))
    if nsplit > 0 and ops.gemm_tc_supported(cin, cout):
        gmode = 2 if transposed else (1 if stride == 2 else 0)
        cands.append(("gemm", ops.conv3d_gemm_tc_pack(w, scale, transposed, nsplit), shift, cout, nsplit, gmode))
    if slab and not transposed and stride == 1 and (cin, cout) in ops.SLAB_SUPPORTED:
        cands.append(("slab", ops.conv3d_slab_pack(w, scale), shift, cout))
    if slab and not transposed and stride == 1 and (cin, cout) in ops.SLABO_SUPPORTED:
        cands.append(("slabo", ops.conv3d_o_slab_pack(w, scale), shift, cout))
    if slab and transposed and (cin, cout) in ops.SLABT_SUPPORTED:
        cands.append(("slabt", ops.conv_transpose3d_slab_pack(w, scale), shift, cout))
    if slab and not transposed and stride == 2 and (cin, cout) in ops.SLAB2_SUPPORTED:
        cands.append(("slab2", ops.conv3d_s2_slab_pack(w, scale), shift, cout))
    return {"cands": cands, "tr": transposed, "stride": stride, "relu": relu, "choice": {}}


# ---------------------------------------------------------------------------------------------------------------
# a3  CostRegNet
# ---------------------------------------------------------------------------------------------------------------
class CostRegNet(nn.Module):
    """3-D U-Net regulariser.  Parameters/keys as the reference (modules.py:470-490); the eval-mode forward runs
    librcmvs conv kernels on channels-last volumes with BatchNorm folded and bias+ReLU+skip fused."""

    _LAYERS = ("conv0", "conv1", "conv2", "conv3", "conv4", "conv5", "conv6", "conv7", "conv9", "conv11")

    def __init__(self, in_channels, base_channels):
        super().__init__()
        c = base_channels
        self.conv0 = Conv3d(in_channels, c, padding=1)
        self.conv1 = Conv3d(c, 2 * c, stride=2, padding=1)
        self.conv2 = Conv3d(2 * c, 2 * c, padding=1)
        self.conv3 = Conv3d(2 * c, 4 * c, stride=2, padding=1)
        self.conv4 = Conv3d(4 * c, 4 * c, padding=1)
        self.conv5 = Conv3d(4 * c, 8 * c, stride=2, padding=1)
        self.conv6 = Conv3d(8 * c, 8 * c, padding=1)
        self.conv7 = Deconv3d(8 * c, 4 * c, stride=2, padding=1, output_padding=1)
        self.conv9 = Deconv3d(4 * c, 2 * c, stride=2, padding=1, output_padding=1)
        self.conv11 = Deconv3d(2 * c, c, stride=2, padding=1, output_padding=1)
        self.prob = nn.Conv3d(c, 1, 3, stride=1, padding=1, bias=False)
        self._packed = PackedCache()

    def packed(self):
        """{layer: [candidate engines]} rebuilt when any parameter/buffer or the precision mode changes.  Every
        candidate computes the same layer (conv + folded BN + bias + ReLU + skip); which one runs is decided per
        input shape by `_choose`: a static table (`_default_engine`) distilled from on-device timings, or a one-off
        timing of the candidates under RCMVS_AUTOTUNE=1."""
        mode = conv_precision()

        def build():
            nsplit = {"bf16x3": 3, "tf32x3": 3, "tf32": 1, "fp32": 0}[mode]
            with torch.no_grad():
                P = {}
                for name in self._LAYERS + ("prob",):
                    if name == "prob":
                        conv, scale, shift, tr, relu = self.prob, None, None, False, False
                    else:
                        blk = getattr(self, name)
                        conv, relu = blk.conv, True
                        scale, shift = (t.contiguous() for t in fold_bn(blk.bn))
                        tr = isinstance(conv, nn.ConvTranspose3d)
                    P[name] = pack_conv3d_layer(conv.weight.detach().contiguous(), scale, shift, tr, conv.stride[0],
                                                relu, nsplit, slab=(mode == "bf16x3"))
            return P
        return self._packed.get((_versions(self), mode), build)

    @staticmethod
    def _call(cand, layer, x, skip, out="y"):
        """`out`: "y" fp32 result; "split" / "both": also / only the pre-split bf16 form (slab engines only)."""
        kind = cand[0]
        if kind in ("slab", "slab2", "slabt", "slabo"):   # TMA-fed split-bf16 tcgen05 kernels; the input is pre-split by its producer when possible
            xs = x if x.dtype == torch.bfloat16 else ops.split_bf16(x)
            wy, ws = out in ("y", "both"), out in ("split", "both")
            if kind == "slabo":
                return ops.conv3d_o_slab(xs, cand[1], cand[3], cand[2], skip, layer["relu"], want_y=wy, want_split=ws)
            if kind == "slabt":
                return ops.conv_transpose3d_slab(xs, cand[1], cand[3], cand[2], skip, layer["relu"], want_y=wy, want_split=ws)
            if kind == "slab2":
                return ops.conv3d_s2_slab(xs, cand[1], cand[3], cand[2], layer["relu"], want_y=wy, want_split=ws)
            return ops.conv3d_slab(xs, cand[1], cand[3], cand[2], skip, layer["relu"], want_y=wy, want_split=ws)
        if out != "y":
            raise RuntimeError("rcmvs: only the slab engines write split-bf16 activations")
        if x.dtype == torch.bfloat16:
            raise RuntimeError("rcmvs: split-bf16 activations can only feed the slab tensor-core engine")
        if kind == "gemm":
            return ops.conv3d_gemm_tc(x, cand[1], cand[3], cand[5], cand[2], skip, cand[4], layer["relu"])
        if kind == "tc":
            return ops.conv3d_tc(x, cand[1], cand[3], cand[2], skip, cand[4], layer["relu"])
        if kind == "tct":
            return ops.conv_transpose3d_tc(x, cand[1], cand[3], cand[2], skip, cand[4], layer["relu"])
        if layer["tr"]:
            return ops.conv_transpose3d(x, cand[1], cand[2], skip=skip, relu=layer["relu"])
        return ops.conv3d(x, cand[1], cand[2], skip, stride=layer["stride"], relu=layer["relu"])

    @classmethod
    def _choose(cls, layer, shape, x=None, skip=None):
        """Index of the engine that runs `layer` for inputs of `shape` (cached per shape)."""
        cands = layer["cands"]
        if len(cands) == 1:
            return 0
        key = tuple(shape)
        if key not in layer["choice"]:
            forced = os.environ.get("RCMVS_ENGINE")            # 'direct' | 'tc' | 'gemm': pin every layer to one engine
            if forced:
                layer["choice"][key] = next((i for i, c in enumerate(cands) if c[0].startswith(forced)), 0)
            elif os.environ.get("RCMVS_AUTOTUNE", "0") == "1" and x is not None:
                best, best_ms = 0, float("inf")
                for i, c in enumerate(cands):
                    cls._call(c, layer, x, skip)               # warm-up (sets kernel attributes)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(3):
                        cls._call(c, layer, x, skip)
                    e1.record()
                    e1.synchronize()
                    ms = e0.elapsed_time(e1)
                    if ms < best_ms:
                        best, best_ms = i, ms
                layer["choice"][key] = best
            else:
                layer["choice"][key] = cls._default_engine(layer, cands, key)
        return layer["choice"][key]

    @classmethod
    def _run(cls, layer, x, skip=None, out="y"):
        """Run a layer with the engine chosen for this input shape (static table, or timed once with CUDA events under
        RCMVS_AUTOTUNE=1, then cached)."""
        shape = x.shape[1:] if x.dtype == torch.bfloat16 else x.shape
        return cls._call(layer["cands"][cls._choose(layer, shape, x, skip)], layer, x, skip, out)

    @classmethod
    def _kind(cls, layer, shape):
        """Name of the engine the static table picks for `layer` at this input shape."""
        return layer["cands"][cls._choose(layer, tuple(shape))][0]

    @staticmethod
    def _default_engine(layer, cands, shape):
        """Static table distilled from RCMVS_AUTOTUNE=1 / tools/time_layers.py runs on B200 at DTU / BASELINE shapes
        (profiles/): stride-1 conv0 / conv2 -> slab tcgen05 kernel; `prob` (8->1, memory-bound) -> FFMA kernel;
        the 1/4- and 1/8-resolution levels conv3..conv7 and the transposed conv9 -> gather-GEMM tcgen05 kernel;
        transposed conv11 -> slab tcgen05 kernel;
        everything else FFMA."""
        kinds = [c[0] for c in cands]
        cin = shape[-1]
        if "slab" in kinds:                                                # bf16x3 mode: conv0 (Cin 16/32), conv2
            return kinds.index("slab")
        if "slab2" in kinds:                                               # bf16x3 mode: conv1 (8 -> 16, stride 2)
            return kinds.index("slab2")
        if "slabo" in kinds:                                               # bf16x3 mode: conv4 (32 -> 32)
            return kinds.index("slabo")
        if "slabt" in kinds:                                               # bf16x3 mode: conv9 / conv11 (transposed)
            return kinds.index("slabt")
        if "tc" in kinds:
            cout = cands[kinds.index("tc")][3]
            return kinds.index("tc") if cout != 1 else 0
        if "gemm" in kinds and cin != 8:                                   # conv1 (8 -> 16, stride 2) stays on FFMA
            return kinds.index("gemm")
        if "tct" in kinds:                                                 # conv11 (16 -> 8, to full resolution)
            return kinds.index("tct")
        return 0

    def engines(self):
        """{layer: {input shape: engine name}} chosen so far (bench.py reports it)."""
        if self._packed.last is None:
            return {}
        return {n: {str(k): l["cands"][i][0] for k, i in l["choice"].items()} for n, l in self._packed.last.items()}

    @staticmethod
    def _static_engines():
        """bf16x3 mode with the built-in engine table (no RCMVS_ENGINE pin, no RCMVS_AUTOTUNE)."""
        return conv_precision() == "bf16x3" and not os.environ.get("RCMVS_ENGINE") and \
            os.environ.get("RCMVS_AUTOTUNE", "0") != "1"

    def wants_split_input(self):
        """True when conv0 runs on the split-bf16 slab engine: the sweep kernel then writes the variance pre-split
        ([2,B,D,h,w,C] bf16: hi, lo) and no conversion pass sits between the two kernels."""
        if not self._static_engines():
            return False
        return (self.conv0.conv.in_channels, self.conv0.conv.out_channels) in ops.SLAB_SUPPORTED

    def forward_cl(self, x_cl):
        """x_cl [B,D,h,w,Cin] channels-last fp32 (or its split-bf16 form [2,B,D,h,w,Cin]) -> cost [B,D,h,w]."""
        dims = x_cl.shape[2:5] if x_cl.dtype == torch.bfloat16 else x_cl.shape[1:4]
        if any(int(v) % 8 for v in dims):
            # three stride-2 levels: the reference fails at the first skip add (modules.py:497) for such shapes
            raise ValueError("CostRegNet needs D, H, W divisible by 8, got %s" % (tuple(dims),))
        P, run = self.packed(), self._run
        B = x_cl.shape[1] if x_cl.dtype == torch.bfloat16 else x_cl.shape[0]
        D, H, W = (int(v) for v in dims)
        c = self.conv0.conv.out_channels
        chain = self._static_engines()                          # activations may flow pre-split between slab engines
        k0 = self._kind(P["conv0"], (B, D, H, W, self.conv0.conv.in_channels)) if chain else ""
        k1 = self._kind(P["conv1"], (B, D, H, W, c)) if chain else ""
        k2 = self._kind(P["conv2"], (B, D // 2, H // 2, W // 2, 2 * c)) if chain else ""
        if k0 == "slab" and k1 == "slab2":
            c0, c0s = run(P["conv0"], x_cl, out="both")           # fp32 for the conv11 skip, split for conv1
            c1 = run(P["conv1"], c0s, out="split" if k2 == "slab" else "y")
        else:
            c0 = run(P["conv0"], x_cl)
            if k1 == "slab2":      # fp32 input (stage 3: conv0 has 8 input channels and runs on the TF32 slab engine): a separate
                c1 = self._call(P["conv1"]["cands"][0], P["conv1"], c0, None)   # split pass costs more than slab2 saves (measured)
            else:
                c1 = run(P["conv1"], c0)
        k3 = self._kind(P["conv3"], (B, D // 2, H // 2, W // 2, 2 * c)) if chain else ""
        k4 = self._kind(P["conv4"], (B, D // 4, H // 4, W // 4, 4 * c)) if chain else ""
        if k2 == "slab" and k3 == "slab2":
            c2, c2s = run(P["conv2"], c1, out="both")            # fp32 for the conv9 skip, split for conv3
            c3 = run(P["conv3"], c2s, out="split" if k4 == "slabo" else "y")
        else:
            c2 = run(P["conv2"], c1)
            c3 = run(P["conv3"], c2)
        c4 = run(P["conv4"], c3)
        y = run(P["conv6"], run(P["conv5"], c4))
        y = run(P["conv7"], y, c4)                               # skip + relu(bn(convT(y)))  (modules.py:497-499)
        if chain and self._kind(P["conv9"], (B, D // 4, H // 4, W // 4, 4 * c)) == "slabt" and \
                self._kind(P["conv11"], (B, D // 2, H // 2, W // 2, 2 * c)) == "slabt":
            y = run(P["conv9"], y, c2, out="split")             # only conv11 reads it: handed over pre-split
        else:
            y = run(P["conv9"], y, c2)
        kp = self._kind(P["prob"], (B, D, H, W, c)) if chain else ""
        y = run(P["conv11"], y, c0, out="split" if (kp == "slab" and self._kind(P["conv11"], (B, D // 2, H // 2, W // 2, 2 * c)) == "slabt") else "y")
        return run(P["prob"], y).squeeze(-1)

    def needs_autograd(self, x=None):
        """The fused inference kernels apply when nothing needs a gradient and BatchNorm runs on its running statistics."""
        if self.training:
            return True
        return torch.is_grad_enabled() and ((x is not None and x.requires_grad) or
                                            any(p.requires_grad for p in self.parameters()))

    def forward(self, x):
        """x [B,C,D,h,w] -> [B,1,D,h,w] (reference signature)."""
        if self.needs_autograd(x):
            if not x.is_cuda:
                library_fallback("CostRegNet.forward (CPU tensor)", "torch conv3d + batch_norm")
                return self._forward_torch(x)
            from rcmvs import autograd as A
            return self.forward_cl_train(A.to_channels_last(x)).unsqueeze(1)
        with torch.no_grad():
            return self.forward_cl(ops.to_channels_last(x.float().contiguous())).unsqueeze(1)

    def forward_cl_train(self, x_cl):
        """Differentiable forward on own kernels: x_cl [B,D,h,w,Cin] -> cost [B,D,h,w].  conv -> BatchNorm (batch statistics
        when self.training, running statistics otherwise) -> ReLU (+ skip after the ReLU, modules.py:497-499), `prob` plain."""
        from rcmvs import autograd as A
        if any(int(v) % 8 for v in x_cl.shape[1:4]):
            raise ValueError("CostRegNet needs D, H, W divisible by 8, got %s" % (tuple(x_cl.shape[1:4]),))
        tr = self.training
        blk = lambda t, m, skip=None: A.conv_block(t, m.conv, m.bn, skip=skip, relu=True, training=tr)
        c0 = blk(x_cl, self.conv0)
        c2 = blk(blk(c0, self.conv1), self.conv2)
        c4 = blk(blk(c2, self.conv3), self.conv4)
        y = blk(blk(c4, self.conv5), self.conv6)
        y = blk(y, self.conv7, c4)
        y = blk(y, self.conv9, c2)
        y = blk(y, self.conv11, c0)
        return A.conv_block(y, self.prob, None, relu=False).squeeze(-1)

    def _forward_torch(self, x):
        c0 = self.conv0(x)
        c2 = self.conv2(self.conv1(c0))
        c4 = self.conv4(self.conv3(c2))
        y = self.conv6(self.conv5(c4))
        y = c4 + self.conv7(y)
        y = c2 + self.conv9(y)
        y = c0 + self.conv11(y)
        return self.prob(y)


# ---------------------------------------------------------------------------------------------------------------
# a4 / a5 helpers with the reference signatures
# ---------------------------------------------------------------------------------------------------------------
def depth_regression(p, depth_values):
    """Σ_d p[:,d]·depth_values[:,d]; kept as a torch expression for API parity (the fused CUDA regression used by
    DepthNet is ops.softargmin_conf)."""
    if depth_values.dim() <= 2:
        depth_values = depth_values.view(*depth_values.shape, 1, 1)
    return torch.sum(p * depth_values, 1)


def get_cur_depth_range_samples(cur_depth, ndepth, depth_inteval_pixel, shape, max_depth=192.0, min_depth=0.0):
    """cur_depth [B,H,W] -> [B,ndepth,H,W] planes at the SAME resolution (no resize): reference :549-566."""
    assert cur_depth.shape == torch.Size(shape), "cur_depth:{}, input shape:{}".format(cur_depth.shape, shape)
    B, H, W = cur_depth.shape
    return ops.depth_hypotheses_next(cur_depth.float().contiguous(), ndepth, depth_inteval_pixel, H, W, H, W)


def get_depth_range_samples(cur_depth, ndepth, depth_inteval_pixel, device, dtype, shape, max_depth=192.0,
                            min_depth=0.0):
    """cur_depth [B,D0] or [B,H,W] -> [B,ndepth,H,W]: reference :569-588."""
    if cur_depth.dim() == 2:
        planes = ops.depth_hypotheses_first(cur_depth.float().contiguous(), ndepth)
        return planes.view(planes.shape[0], ndepth, 1, 1).expand(-1, -1, shape[1], shape[2]).contiguous()
    return get_cur_depth_range_samples(cur_depth, ndepth, depth_inteval_pixel, shape, max_depth, min_depth)


def cas_mvsnet_loss(inputs, depth_gt_ms, mask_ms, **kwargs):
    """Supervised multi-stage smooth-L1 (reference :527-546); host glue, unchanged semantics."""
    weights = kwargs.get("dlossw", None)
    total = torch.zeros((), dtype=torch.float32, device=mask_ms["stage1"].device)
    depth_loss = None
    for key in [k for k in inputs.keys() if "stage" in k]:
        mask = mask_ms[key] > 0.5
        depth_loss = F.smooth_l1_loss(inputs[key]["depth"][mask], depth_gt_ms[key][mask], reduction="mean")
        wgt = weights[int(key.replace("stage", "")) - 1] if weights is not None else 1.0
        total = total + wgt * depth_loss
    return total, depth_loss
