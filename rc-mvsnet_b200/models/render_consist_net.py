"""Drop-in for the reference's models/render_consist_net.py.

`Rendering_Consistency_Net(args)(volume_feature_warp, pseudo_depth, batch)` keeps the reference call signature and
its 8-tuple result `(rgb, input_feat, weights, depth_pred, alpha, {}, rays_depth, target_s)` (reference :54-76 — the
reference binds the 2nd/3rd entries to names `disp, acc`, but what `rendering` hands back there is the per-point
feature tensor and the compositing weights).  State-dict keys are `MVSNet.cost_reg_2.*` and `network_fn.nerf.*`, so
`pretrain/model_000014_nerf.ckpt` loads with strict=True.

Under torch.no_grad() the whole branch runs on librcmvs kernels:
    Neural_Volume_Net (resize + 3-D U-Net, channels-last)  ->  pixel pick + Gaussian-Uniform depth candidates
    -> ray points / NDC -> per-point features -> fused embedding + MLP -> alpha compositing.
With autograd enabled (training) the same math runs on differentiable torch ops; the sampler stays on CUDA.
"""
import torch
import torch.nn as nn

from .render_models import create_nerf_mvs
from .render_utils import build_rays_norm, filter_keys, sub_selete_data
from .renderer import rendering

__all__ = ["Rendering_Consistency_Net"]

_IMAGENET_MEAN = (0.485, 0.456, 0.406)
_IMAGENET_STD = (0.229, 0.224, 0.225)
_N_RAYS = 1024            # hard-coded in the reference forward (:68)


class Rendering_Consistency_Net(nn.Module):
    def __init__(self, args):
        super().__init__()
        args.feat_dim = 8 + 3 * 4                      # 8 volume channels + 3 views x (RGB + in-bounds mask)
        self.args, self.idx = args, 0
        train_kw, self.render_kwargs_test, _, self.grad_vars = create_nerf_mvs(
            args, use_mvs=True, dir_embedder=False, pts_embedder=True, share_warp=True)
        filter_keys(train_kw)
        # sub-modules / callables the reference exposes as attributes (and whose names fix the state-dict keys)
        self.MVSNet = train_kw.pop("network_mvs")
        self.network_fn = train_kw.pop("network_fn")
        self.network_query_fn = train_kw.pop("network_query_fn")
        self.white_bkgd = train_kw.pop("white_bkgd")
        train_kw["NDC_local"] = False
        self.render_kwargs_train = train_kw

    # ---- helpers -----------------------------------------------------------------------------------------------------
    def decode_batch(self, batch, idx=(0, 1, 2, 3)):
        """NeRF-side tensors of a sample on this module's device + the squeezed pose dictionary."""
        device = next(self.parameters()).device        # the reference pins a module-global cuda:0 instead
        data = sub_selete_data(batch, device, list(idx), filtKey=[])
        poses = {name: data[name].squeeze() for name in ("w2cs", "intrinsics", "c2ws", "near_fars")}
        return data, poses

    @staticmethod
    def unpreprocess(data, shape=(1, 1, 3, 1, 1)):
        """Undo the ImageNet normalisation: x*std + mean, written like the reference as (x - (-mean/std)) / (1/std)."""
        shift = torch.tensor([-m / s for m, s in zip(_IMAGENET_MEAN, _IMAGENET_STD)], device=data.device).view(*shape)
        scale = torch.tensor([1.0 / s for s in _IMAGENET_STD], device=data.device).view(*shape)
        return (data - shift) / scale

    def _encode(self, volume_feature_warp, fused):
        """-> (NCDHW encoded volume or None, channels-last [128,h,w,8] volume or None).  Both branches run on librcmvs
        kernels: the fused inference kernels under no_grad with eval-mode BatchNorm, the differentiable ones
        (rcmvs.autograd) otherwise — the mode the reference actually calls this module in (train_rcmvsnet.py:280-312)."""
        if fused:
            return None, self.MVSNet.forward_cl(volume_feature_warp)[0]
        if volume_feature_warp.is_cuda:
            return None, self.MVSNet.forward_cl_train(volume_feature_warp)[0]
        return self.MVSNet(volume_feature_warp, pad=self.args.pad), None

    # ---- forward -----------------------------------------------------------------------------------------------------
    def forward(self, volume_feature_warp, pseudo_depth, batch, samples=None):
        """`samples` (optional, not in the reference) injects {'pixel_coordinates', 'depth_candidates'} for parity tests."""
        batch.pop("scan", None)
        data, poses = self.decode_batch(batch)
        n_views, height, width = data["depths_h"].shape[1:]
        depth_per_view = pseudo_depth.expand(n_views, height, width).unsqueeze(0)      # same pseudo depth for every view

        fused = not torch.is_grad_enabled() and not self.MVSNet.training
        volume, volume_cl = self._encode(volume_feature_warp, fused)
        images = self.unpreprocess(data["imgs"])

        pts, dirs, target_rgb, ndc, z_vals, origins, rays_depth, _ = build_rays_norm(
            images, depth_per_view, poses, poses["w2cs"], poses["c2ws"], poses["intrinsics"], data["near_fars"],
            _N_RAYS, self.args.N_samples, pad=self.args.pad, samples=samples)

        rgb, point_feat, weights, depth_pred, alpha, extras = rendering(
            self.args, poses, pts, ndc, z_vals, origins, dirs, volume, images[:, -3:], img_feat=None,
            network_fn=self.network_fn, network_query_fn=self.network_query_fn, white_bkgd=self.white_bkgd,
            volume_cl=volume_cl)
        return rgb, point_feat, weights, depth_pred, alpha, extras, rays_depth, target_rgb
