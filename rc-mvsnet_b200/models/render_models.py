"""Drop-in for the NeRF-side model definitions of the reference's models/render_models.py: Embedder / get_embedder
(:15-66), Renderer_ours (:143-220), RenderNet (:538-565), create_nerf_mvs (:567-655), CostReg (:690-734),
Neural_Volume_Net (:736-760).  Same parameter names (state_dict keys 'network_fn.nerf.*', 'MVSNet.cost_reg_2.*').
The eval/no-grad forward runs on librcmvs kernels; with autograd enabled the same math runs on torch ops so that
train_rcmvsnet.py keeps working until the fused backward kernels land.
Unlike the reference this module does NOT switch on torch.autograd.set_detect_anomaly at import (render_models.py:2)."""
import ctypes

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from rcmvs import ops
from rcmvs.cache import PackedCache
from rcmvs.strict import library_fallback
from .modules import fold_bn, _versions

__all__ = ["Embedder", "get_embedder", "Renderer_ours", "RenderNet", "create_nerf_mvs", "CostReg",
           "Neural_Volume_Net", "weights_init"]


def weights_init(m):
    if isinstance(m, nn.Linear):
        nn.init.kaiming_normal_(m.weight.data)
        if m.bias is not None:
            nn.init.zeros_(m.bias.data)


class Embedder:
    """[x, sin(2^k x) for all k, cos(2^k x) for all k] — frequency-major, then axis (reference :45-49)."""

    def __init__(self, **kwargs):
        self.kwargs = kwargs
        d, n = kwargs["input_dims"], kwargs["num_freqs"]
        if kwargs.get("log_sampling", True):
            self.freq_bands = 2.0 ** torch.linspace(0.0, kwargs["max_freq_log2"], steps=n)
        else:
            self.freq_bands = torch.linspace(1.0, 2.0 ** kwargs["max_freq_log2"], steps=n)
        self.out_dim = d * (1 + 2 * n) if kwargs.get("include_input", True) else d * 2 * n

    def embed(self, inputs):
        fb = self.freq_bands.to(inputs.device)
        scaled = (inputs.unsqueeze(-2) * fb.view(*[1] * (inputs.dim() - 1), -1, 1)).reshape(*inputs.shape[:-1], -1)
        return torch.cat((inputs, torch.sin(scaled), torch.cos(scaled)), dim=-1)


def get_embedder(multires, i=0, input_dims=3):
    if i == -1:
        return nn.Identity(), 3
    eo = Embedder(include_input=True, input_dims=input_dims, max_freq_log2=multires - 1, num_freqs=multires,
                  log_sampling=True, periodic_fns=[torch.sin, torch.cos])
    return (lambda x, eo=eo: eo.embed(x)), eo.out_dim


class Renderer_ours(nn.Module):
    def __init__(self, D=8, W=256, input_ch=3, input_ch_views=3, output_ch=4, input_ch_feat=8, skips=[4],
                 use_viewdirs=False):
        super().__init__()
        self.D, self.W, self.input_ch, self.input_ch_views, self.skips = D, W, input_ch, input_ch_views, skips
        self.use_viewdirs = use_viewdirs
        self.in_ch_pts, self.in_ch_views, self.in_ch_feat = input_ch, input_ch_views, input_ch_feat
        self.pts_linears = nn.ModuleList(
            [nn.Linear(input_ch, W)] + [nn.Linear(W + input_ch, W) if i in skips else nn.Linear(W, W)
                                        for i in range(D - 1)])
        self.pts_bias = nn.Linear(input_ch_feat, W)
        self.views_linears = nn.ModuleList([nn.Linear(input_ch_views + W, W // 2)])
        if use_viewdirs:
            self.feature_linear = nn.Linear(W, W)
            self.alpha_linear = nn.Linear(W, 1)
            self.rgb_linear = nn.Linear(W // 2, 3)
        else:
            self.output_linear = nn.Linear(W, output_ch)
        for m in (self.pts_linears, self.views_linears):
            m.apply(weights_init)
        if use_viewdirs:
            for m in (self.feature_linear, self.alpha_linear, self.rgb_linear):
                m.apply(weights_init)
        self._packed = PackedCache()

    def _trunk(self, input_pts, input_feats):
        h, bias = input_pts, self.pts_bias(input_feats)
        for i, lin in enumerate(self.pts_linears):
            h = F.relu(lin(h) * bias)                      # multiplicative modulation (reference :200)
            if i in self.skips:
                h = torch.cat([input_pts, h], -1)
        return h

    def forward_alpha(self, x):
        input_pts, input_feats = torch.split(x, [self.in_ch_pts, x.shape[-1] - self.in_ch_pts], dim=-1)
        return torch.relu(self.alpha_linear(self._trunk(input_pts, input_feats)))

    def forward(self, x):
        nf = x.shape[-1] - self.in_ch_pts - self.in_ch_views
        input_pts, input_feats, input_views = torch.split(x, [self.in_ch_pts, nf, self.in_ch_views], dim=-1)
        h = self._trunk(input_pts, input_feats)
        if not self.use_viewdirs:
            return self.output_linear(h)
        alpha = torch.relu(self.alpha_linear(h))
        h = torch.cat([self.feature_linear(h), input_views], -1)
        for lin in self.views_linears:
            h = F.relu(lin(h))
        return torch.cat([torch.sigmoid(self.rgb_linear(h)), alpha], -1)

    def forward_train(self, ndc, feat, dirs):
        """Differentiable forward on own kernels (rcmvs.autograd: SGEMM Linear layers, fused activation kernels, all with
        hand-written backward): ndc [N,S,3] (no gradient), feat [N,S,20], dirs [N,3] -> raw [N,S,4] = (rgb, sigma).
        Same math as forward(): multiplicative bias modulation (:200), skip concat after layer 4, relu'd density."""
        from rcmvs import autograd as A
        N, S = ndc.shape[:2]
        M = N * S
        pts = A.embed(ndc.reshape(M, 3))                                           # [M,63]
        feats = feat.reshape(M, feat.shape[-1])
        views = dirs[:, None].expand(N, S, dirs.shape[-1]).reshape(M, -1).contiguous()
        bias = A.linear(feats, self.pts_bias)
        h = pts
        for i, lin in enumerate(self.pts_linears):
            h = A.mul_relu(A.linear(h, lin), bias)
            if i in self.skips:
                h = torch.cat([pts, h], -1)
        alpha = A.relu(A.linear(h, self.alpha_linear))
        h2 = torch.cat([A.linear(h, self.feature_linear), views], -1)
        for lin in self.views_linears:
            h2 = A.relu(A.linear(h2, lin))
        rgb = A.sigmoid(A.linear(h2, self.rgb_linear))
        return torch.cat([rgb, alpha], -1).view(N, S, 4)

    # ---- CUDA path ------------------------------------------------------------------------------------------------
    def cuda_supported(self):
        return (self.use_viewdirs and self.D == 6 and self.W == 128 and self.in_ch_pts == 63 and self.in_ch_feat == 20
                and self.in_ch_views == 3 and list(self.skips) == [4])

    def packed(self):
        """Host array of 22 device pointers to transposed weights/biases (rcmvs_render_mlp), cached by version."""
        def build():
            with torch.no_grad():
                ts = []
                for lin in self.pts_linears:
                    ts += [lin.weight.t().contiguous(), lin.bias.contiguous()]
                ts += [self.pts_bias.weight.t().contiguous(), self.pts_bias.bias.contiguous(),
                       self.alpha_linear.weight.reshape(-1).contiguous(), self.alpha_linear.bias.contiguous(),
                       self.feature_linear.weight.t().contiguous(), self.feature_linear.bias.contiguous(),
                       self.views_linears[0].weight.t().contiguous(), self.views_linears[0].bias.contiguous(),
                       self.rgb_linear.weight.t().contiguous(), self.rgb_linear.bias.contiguous()]
                arr = (ctypes.c_void_p * 22)(*[t.data_ptr() for t in ts])
                # split-bf16 images of the eight wide Linear layers for the tensor-core path (rcmvs_render_mlp_tc)
                wimg = ops.render_mlp_tc_pack(arr, ts[0].device) if ts[0].is_cuda else None
            return arr, ts, wimg                   # keep the tensors alive
        return self._packed.get(_versions(self), build)

    def forward_cuda(self, ndc, feat, dirs):
        """ndc [N,S,3], feat [N,S,20], dirs [N,3] -> raw [N,S,4] (embedding fused in the kernel).  Default: the fused fp32
        FFMA kernel (csrc/render.cu).  RCMVS_MLP_TC=1: Linear layers on tcgen05 (csrc/mlp_tc.cu, split-bf16: 2.6x faster,
        equal to 2e-4 on well-conditioned weights; with the shipped checkpoint on synthetic input — sigma up to 1e7 — 4 of
        12 288 golden densities differ by up to 8e-3 relative instead of <= 2e-3, hence opt-in)."""
        arr, _, wimg = self.packed()
        if os.environ.get("RCMVS_MLP_TC", "0") == "1" and wimg is not None:
            return ops.render_mlp_tc(ndc.contiguous(), feat.contiguous(), dirs.contiguous(), arr, wimg)
        return ops.render_mlp(ndc.contiguous(), feat.contiguous(), dirs.contiguous(), arr)


class RenderNet(nn.Module):
    def __init__(self, D=8, W=256, input_ch_pts=3, input_ch_views=3, input_ch_feat=8, skips=[4], net_type="v2"):
        super().__init__()
        self.in_ch_pts, self.in_ch_views, self.in_ch_feat = input_ch_pts, input_ch_views, input_ch_feat
        if net_type != "v0":
            raise NotImplementedError("rcmvs implements net_type='v0' (Renderer_ours), the reference default")
        self.nerf = Renderer_ours(D=D, W=W, input_ch_feat=input_ch_feat, input_ch=input_ch_pts, output_ch=4,
                                  skips=skips, input_ch_views=input_ch_views, use_viewdirs=True)

    def forward_alpha(self, x):
        return self.nerf.forward_alpha(x)

    def forward(self, x):
        return self.nerf(x)


class _BatchNormAnyRank(nn.BatchNorm3d):
    """The reference applies nn.BatchNorm2d to 5-D tensors, which only works after SyncBatchNorm conversion
    (train_rcmvsnet.py:524-525); this variant accepts any rank so the module also runs un-converted."""

    def _check_input_dim(self, input):
        return


class ConvBnReLU3D(nn.Module):
    """conv -> norm, NO ReLU despite the name (reference :675-686)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, pad=1, norm_act=_BatchNormAnyRank):
        super().__init__()
        self.conv = nn.Conv3d(in_channels, out_channels, kernel_size, stride=stride, padding=pad, bias=False)
        self.bn = norm_act(out_channels)

    def forward(self, x):
        return self.bn(self.conv(x))


class CostReg(nn.Module):
    def __init__(self, in_channels, norm_act=_BatchNormAnyRank, base_channels=4):
        super().__init__()
        c = base_channels
        self.conv0 = ConvBnReLU3D(in_channels, c, norm_act=norm_act)
        self.conv1 = ConvBnReLU3D(c, 2 * c, stride=2, norm_act=norm_act)
        self.conv2 = ConvBnReLU3D(2 * c, 2 * c, norm_act=norm_act)
        self.conv3 = ConvBnReLU3D(2 * c, 4 * c, stride=2, norm_act=norm_act)
        self.conv4 = ConvBnReLU3D(4 * c, 4 * c, norm_act=norm_act)
        self.conv5 = ConvBnReLU3D(4 * c, 8 * c, stride=2, norm_act=norm_act)
        self.conv6 = ConvBnReLU3D(8 * c, 8 * c, norm_act=norm_act)

        def up(i, o, n):
            return nn.Sequential(nn.ConvTranspose3d(i, o, 3, padding=1, output_padding=1, stride=2, bias=False),
                                 norm_act(n))
        self.conv7, self.conv9, self.conv11 = up(8 * c, 4 * c, 32), up(4 * c, 2 * c, 16), up(2 * c, c, 8)
        self._packed = PackedCache()

    def forward(self, x):
        c0 = self.conv0(x)
        c2 = self.conv2(self.conv1(c0))
        c4 = self.conv4(self.conv3(c2))
        y = self.conv6(self.conv5(c4))
        y = c4 + self.conv7(y)
        y = c2 + self.conv9(y)
        return c0 + self.conv11(y)

    def packed(self, cin_pad):
        def build():
            with torch.no_grad():
                P = {}
                for name in ("conv0", "conv1", "conv2", "conv3", "conv4", "conv5", "conv6"):
                    blk = getattr(self, name)
                    scale, shift = (t.contiguous() for t in fold_bn(blk.bn))
                    w = blk.conv.weight.detach()
                    if name == "conv0" and w.shape[1] < cin_pad:      # 41 -> 48 input channels, zero weights
                        w = F.pad(w, (0, 0, 0, 0, 0, 0, 0, cin_pad - w.shape[1]))
                    P[name] = (ops.pack_conv3d_weight(w.contiguous(), scale, False), shift)
                    if name in ("conv0", "conv2") and (w.shape[1], w.shape[0]) in ops.SLAB_SUPPORTED:
                        P[name + ".slab"] = (ops.conv3d_slab_pack(w.contiguous(), scale), shift)   # tcgen05 + TMA, bf16x3
                    if name in ("conv1", "conv3") and (w.shape[1], w.shape[0]) in ops.SLAB2_SUPPORTED:
                        P[name + ".slab2"] = (ops.conv3d_s2_slab_pack(w.contiguous(), scale), shift)  # stride-2 TMA kernel
                    if name == "conv4" and (w.shape[1], w.shape[0]) in ops.SLABO_SUPPORTED:
                        P["conv4.slabo"] = (ops.conv3d_o_slab_pack(w.contiguous(), scale), shift)
                for name in ("conv7", "conv9", "conv11"):
                    seq = getattr(self, name)
                    scale, shift = (t.contiguous() for t in fold_bn(seq[1]))
                    P[name] = (ops.pack_conv3d_weight(seq[0].weight.detach().contiguous(), scale, True), shift)
                    wt = seq[0].weight.detach().contiguous()
                    if (wt.shape[0], wt.shape[1]) in ops.SLABT_SUPPORTED:
                        P[name + ".slabt"] = (ops.conv_transpose3d_slab_pack(wt, scale), shift)      # transposed TMA kernel
            return P
        return self._packed.get((_versions(self), cin_pad), build)

    def forward_cl_train(self, x_cl, training):
        """Differentiable forward on own kernels: conv -> BatchNorm (batch statistics when `training`), NO ReLU
        (reference :685-686), skip adds (:727-733)."""
        from rcmvs import autograd as A
        blk = lambda t, m: A.conv_block(t, m.conv, m.bn, relu=False, training=training)
        up = lambda t, seq, skip: A.conv_block(t, seq[0], seq[1], skip=skip, relu=False, training=training)
        c0 = blk(x_cl, self.conv0)
        c2 = blk(blk(c0, self.conv1), self.conv2)
        c4 = blk(blk(c2, self.conv3), self.conv4)
        y = blk(blk(c4, self.conv5), self.conv6)
        y = up(y, self.conv7, c4)
        y = up(y, self.conv9, c2)
        return up(y, self.conv11, c0)

    def forward_cl(self, x_cl):
        """x_cl [B,D,h,w,Cpad] -> [B,D,h,w,8]; eval-mode BN folded, no ReLU anywhere."""
        P = self.packed(x_cl.shape[-1])          # (the split form has its channel count last as well)
        tma = x_cl.dtype == torch.bfloat16 and all(k in P for k in ("conv1.slab2", "conv2.slab", "conv3.slab2", "conv4.slabo",
                                                                     "conv9.slabt", "conv11.slabt"))
        if tma:
            # pre-split input from resize_depth_cl_split: every layer down to quarter resolution and the two upper transposed
            # convolutions on the TMA-fed tensor-core kernels, activations handed over in split-bf16 form (no ReLU anywhere)
            c0, c0s = ops.conv3d_slab(x_cl, P["conv0.slab"][0], 8, P["conv0.slab"][1], None, relu=False, want_split=True)
            c1s = ops.conv3d_s2_slab(c0s, P["conv1.slab2"][0], 16, P["conv1.slab2"][1], relu=False, want_y=False, want_split=True)
            c2, c2s = ops.conv3d_slab(c1s, P["conv2.slab"][0], 16, P["conv2.slab"][1], None, relu=False, want_split=True)
            c3s = ops.conv3d_s2_slab(c2s, P["conv3.slab2"][0], 32, P["conv3.slab2"][1], relu=False, want_y=False, want_split=True)
            c4 = ops.conv3d_o_slab(c3s, P["conv4.slabo"][0], 32, P["conv4.slabo"][1], None, relu=False)
            y = ops.conv3d(ops.conv3d(c4, *P["conv5"], stride=2, relu=False), *P["conv6"], relu=False)
            y = ops.conv_transpose3d(y, *P["conv7"], skip=c4, relu=False)
            if os.environ.get("RCMVS_NVN_TCT", "0") == "1":     # measured: with the two upper transposed convs on bf16x3 as well,
                # 4 of 163 840 sampled point features of the full-size golden miss the 5e-3 absolute bar (6.8e-3): off by default
                y9s = ops.conv_transpose3d_slab(ops.split_bf16(y), P["conv9.slabt"][0], 16, P["conv9.slabt"][1], c2, relu=False,
                                                want_y=False, want_split=True)
                return ops.conv_transpose3d_slab(y9s, P["conv11.slabt"][0], 8, P["conv11.slabt"][1], c0, relu=False)
            y = ops.conv_transpose3d(y, *P["conv9"], skip=c2, relu=False)
            return ops.conv_transpose3d(y, *P["conv11"], skip=c0, relu=False)
        if x_cl.dtype == torch.bfloat16 and "conv1.slab2" in P and "conv2.slab" in P:
            c0, c0s = ops.conv3d_slab(x_cl, P["conv0.slab"][0], 8, P["conv0.slab"][1], None, relu=False, want_split=True)
            c1s = ops.conv3d_s2_slab(c0s, P["conv1.slab2"][0], 16, P["conv1.slab2"][1], relu=False, want_y=False, want_split=True)
            c2 = ops.conv3d_slab(c1s, P["conv2.slab"][0], 16, P["conv2.slab"][1], None, relu=False)
        else:
            if x_cl.dtype == torch.bfloat16:
                c0 = ops.conv3d_slab(x_cl, P["conv0.slab"][0], 8, P["conv0.slab"][1], None, relu=False)
            else:
                c0 = ops.conv3d(x_cl, *P["conv0"], relu=False)
            c2 = ops.conv3d(ops.conv3d(c0, *P["conv1"], stride=2, relu=False), *P["conv2"], relu=False)
        c4 = ops.conv3d(ops.conv3d(c2, *P["conv3"], stride=2, relu=False), *P["conv4"], relu=False)
        y = ops.conv3d(ops.conv3d(c4, *P["conv5"], stride=2, relu=False), *P["conv6"], relu=False)
        y = ops.conv_transpose3d(y, *P["conv7"], skip=c4, relu=False)
        y = ops.conv_transpose3d(y, *P["conv9"], skip=c2, relu=False)
        return ops.conv_transpose3d(y, *P["conv11"], skip=c0, relu=False)


class Neural_Volume_Net(nn.Module):
    def __init__(self, num_groups=1, norm_act=_BatchNormAnyRank, levels=1):
        super().__init__()
        self.levels, self.n_depths, self.G, self.N_importance, self.chunk = levels, [128, 32, 8], num_groups, 0, 1024
        self.cost_reg_2 = CostReg(32 + 9, norm_act, base_channels=8)

    def forward(self, volume_feature, pad=0):
        """volume_feature [B,41,D0,h,w] -> [1,B*8,128,h,w] (reference :753-760)."""
        B, C, _, H, W = volume_feature.shape
        if self.needs_autograd(volume_feature):
            if not volume_feature.is_cuda:
                library_fallback("Neural_Volume_Net.forward (CPU tensor)", "torch conv3d + batch_norm")
                v = F.interpolate(volume_feature, size=[128, H, W], mode="trilinear", align_corners=True)
                v = self.cost_reg_2(v)
                return v.reshape(1, -1, *v.shape[2:])
            from rcmvs import autograd as A
            return A.to_planar(self.forward_cl_train(volume_feature)).reshape(1, -1, 128, H, W)
        with torch.no_grad():
            return ops.to_planar(self.forward_cl(volume_feature)).reshape(1, -1, 128, H, W)

    def needs_autograd(self, x=None):
        if self.training:
            return True
        return torch.is_grad_enabled() and ((x is not None and x.requires_grad) or
                                            any(p.requires_grad for p in self.parameters()))

    def forward_cl_train(self, volume_feature):
        """Differentiable: NCDHW in -> channels-last encoded volume [B,128,h,w,8] (layout change, depth resize and the 3-D
        U-Net all with hand-written backward)."""
        from rcmvs import autograd as A
        x = A.to_channels_last(volume_feature)
        x = A.resize_depth_cl(x, 128, (x.shape[-1] + 7) // 8 * 8)
        return self.cost_reg_2.forward_cl_train(x, self.training)

    def forward_cl(self, volume_feature):
        """NCDHW in -> channels-last encoded volume [B,128,h,w,8] (what the point-feature kernel consumes)."""
        x = ops.to_channels_last(volume_feature.float().contiguous())
        cpad = (x.shape[-1] + 7) // 8 * 8
        from .modules import conv_precision
        if conv_precision() == "bf16x3" and (cpad, 8) in ops.SLAB_SUPPORTED and min(x.shape[2:4]) >= 2:
            return self.cost_reg_2.forward_cl(ops.resize_depth_cl_split(x, 128, cpad))
        return self.cost_reg_2.forward_cl(ops.resize_depth_cl(x, 128, cpad))


def create_nerf_mvs(args, pts_embedder=True, use_mvs=False, dir_embedder=True, share_warp=False):
    """Same return tuple as the reference (:567-655): (render_kwargs_train, render_kwargs_test, start, grad_vars)."""
    from .renderer import run_network_mvs
    embed_fn, input_ch = get_embedder(args.multires, args.i_embed, input_dims=args.pts_dim) if pts_embedder \
        else (None, args.pts_dim)
    embeddirs_fn, input_ch_views = get_embedder(args.multires_views, args.i_embed, input_dims=args.dir_dim) \
        if dir_embedder else (None, args.dir_dim)
    model = RenderNet(D=args.netdepth, W=args.netwidth, input_ch_pts=input_ch, skips=[4],
                      input_ch_views=input_ch_views, input_ch_feat=args.feat_dim, net_type=args.net_type)
    grad_vars = list(model.parameters())
    model_fine = None
    if args.N_importance > 0:
        raise NotImplementedError("N_importance > 0 is never used by the reference defaults")
    encoding = None
    if use_mvs:
        if not share_warp:
            raise NotImplementedError
        encoding = Neural_Volume_Net()
        grad_vars += list(encoding.parameters())
    if args.ckpt is not None and args.ckpt != "None":
        ckpt = torch.load(args.ckpt, map_location="cpu")
        if use_mvs:
            encoding.load_state_dict(ckpt["network_mvs_state_dict"])
        model.load_state_dict(ckpt["network_fn_state_dict"])

    def network_query_fn(pts, viewdirs, rays_feats, network_fn):
        return run_network_mvs(pts, viewdirs, rays_feats, network_fn, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn,
                               netchunk=args.netchunk)
    train = {"network_query_fn": network_query_fn, "perturb": args.perturb, "N_importance": args.N_importance,
             "network_fine": model_fine, "N_samples": args.N_samples, "network_fn": model, "network_mvs": encoding,
             "use_viewdirs": args.use_viewdirs, "white_bkgd": args.white_bkgd, "raw_noise_std": args.raw_noise_std}
    test = dict(train)
    test["perturb"] = False
    return train, test, 0, grad_vars
