"""Drop-in for the reference's models/render_consist_net.py: Rendering_Consistency_Net(args).forward(
volume_feature_warp, pseudo_depth, batch) -> (rgb, input_feat, weights, depth_pred, alpha, {}, rays_depth, target_s)
(reference :54-76; the reference names the 2nd/3rd entries `disp, acc` but `rendering` returns input_feat, weights)."""
import torch
import torch.nn as nn

from .render_models import create_nerf_mvs
from .render_utils import build_rays_norm, filter_keys, sub_selete_data
from .renderer import rendering

__all__ = ["Rendering_Consistency_Net"]


class Rendering_Consistency_Net(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.args = args
        self.args.feat_dim = 8 + 3 * 4
        self.idx = 0
        self.render_kwargs_train, self.render_kwargs_test, start, self.grad_vars = create_nerf_mvs(
            args, use_mvs=True, dir_embedder=False, pts_embedder=True, share_warp=True)
        filter_keys(self.render_kwargs_train)
        kw = self.render_kwargs_train
        self.MVSNet, self.network_fn = kw.pop("network_mvs"), kw.pop("network_fn")
        self.network_query_fn, self.white_bkgd = kw.pop("network_query_fn"), kw.pop("white_bkgd")
        kw["NDC_local"] = False

    def decode_batch(self, batch, idx=list(range(4))):
        device = next(self.parameters()).device           # the reference hard-codes cuda:0 (:8,:38)
        data = sub_selete_data(batch, device, idx, filtKey=[])
        pose_ref = {k: data[k].squeeze() for k in ("w2cs", "intrinsics", "c2ws", "near_fars")}
        return data, pose_ref

    @staticmethod
    def unpreprocess(data, shape=(1, 1, 3, 1, 1)):
        mean = torch.tensor([-0.485 / 0.229, -0.456 / 0.224, -0.406 / 0.225], device=data.device).view(*shape)
        std = torch.tensor([1 / 0.229, 1 / 0.224, 1 / 0.225], device=data.device).view(*shape)
        return (data - mean) / std

    def forward(self, volume_feature_warp, pseudo_depth, batch, samples=None):
        batch.pop("scan", None)
        data, pose_ref = self.decode_batch(batch)
        imgs, near_fars, depths_h = data["imgs"], data["near_fars"], data["depths_h"]
        _, V, H, W = depths_h.shape
        pseudo_depth_h = pseudo_depth.expand(V, H, W).unsqueeze(0)
        fused = not torch.is_grad_enabled()
        volume_cl = None
        if fused:
            vcl = self.MVSNet.forward_cl(volume_feature_warp)            # [1,128,h,w,8] channels-last
            volume_cl, volume_feature = vcl[0], None
        else:
            volume_feature = self.MVSNet(volume_feature_warp, pad=self.args.pad)
        imgs = self.unpreprocess(imgs)
        N_rays, N_samples = 1024, self.args.N_samples
        c2ws, w2cs, intrinsics = pose_ref["c2ws"], pose_ref["w2cs"], pose_ref["intrinsics"]
        rays_pts, rays_dir, target_s, rays_NDC, depth_candidates, rays_o, rays_depth, _ = build_rays_norm(
            imgs, pseudo_depth_h, pose_ref, w2cs, c2ws, intrinsics, near_fars, N_rays, N_samples, pad=self.args.pad,
            samples=samples)
        rgb, input_feat, weights, depth_pred, alpha, ret = rendering(
            self.args, pose_ref, rays_pts, rays_NDC, depth_candidates, rays_o, rays_dir, volume_feature,
            imgs[:, -3:], img_feat=None, network_fn=self.network_fn, network_query_fn=self.network_query_fn,
            white_bkgd=self.white_bkgd, volume_cl=volume_cl)
        return rgb, input_feat, weights, depth_pred, alpha, ret, rays_depth, target_s
