"""Drop-in for the reference's models/renderer.py: depth2dist :5-11, raw2alpha :18-26, batchify :28-40,
run_network_mvs :42-63, raw2outputs :65-93, gen_dir_feature :141-152, gen_pts_feats :154-166, rendering :168-195.
No-grad calls run the fused CUDA kernels (point_features / render_mlp / composite); with autograd the same math
runs on torch ops."""
import torch
import torch.nn.functional as F

from rcmvs import ops
from rcmvs.strict import library_fallback
from .render_utils import build_color_volume, index_point_feature

__all__ = ["depth2dist", "raw2alpha", "batchify", "run_network_mvs", "raw2outputs", "gen_dir_feature",
           "gen_pts_feats", "rendering"]


def depth2dist(z_vals, cos_angle):
    d = z_vals[..., 1:] - z_vals[..., :-1]
    d = torch.cat([d, torch.full_like(d[..., :1], 1e10)], -1)
    return d * cos_angle.unsqueeze(-1)


def raw2alpha(sigma, dist, net_type):
    """alpha = 1-exp(-sigma): `dist` is ignored, as in the reference (:18-26)."""
    alpha = 1.0 - torch.exp(-sigma)
    T = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1.0 - alpha + 1e-10], -1), -1)[:, :-1]
    return alpha, alpha * T, F.softmax(sigma, 1)


def batchify(fn, chunk):
    if chunk is None:
        return fn

    def ret(inputs, alpha_only):
        f = fn.forward_alpha if alpha_only else fn
        return torch.cat([f(inputs[i:i + chunk]) for i in range(0, inputs.shape[0], chunk)], 0)
    return ret


def run_network_mvs(pts, viewdirs, alpha_feat, fn, embed_fn, embeddirs_fn, netchunk=1024):
    if (not torch.is_grad_enabled() and viewdirs is not None and embeddirs_fn is None and alpha_feat is not None
            and pts.is_cuda and hasattr(fn, "nerf") and fn.nerf.cuda_supported() and pts.dim() == 3):
        dirs = viewdirs if viewdirs.dim() == 2 else viewdirs[:, 0]
        return fn.nerf.forward_cuda(pts.float(), alpha_feat.float(), dirs.float())
    if (viewdirs is not None and embeddirs_fn is None and alpha_feat is not None and pts.is_cuda and hasattr(fn, "nerf")
            and fn.nerf.cuda_supported() and pts.dim() == 3):
        dirs = viewdirs if viewdirs.dim() == 2 else viewdirs[:, 0]
        return fn.nerf.forward_train(pts.float(), alpha_feat.float(), dirs.float())      # differentiable, own kernels
    library_fallback("run_network_mvs (non-default network configuration)", "cuBLAS linear layers")
    if embed_fn is not None:
        pts = embed_fn(pts)
    if alpha_feat is not None:
        pts = torch.cat((pts, alpha_feat), dim=-1)
    if viewdirs is not None:
        if viewdirs.dim() != 3:
            viewdirs = viewdirs[:, None].expand(-1, pts.shape[1], -1)
        if embeddirs_fn is not None:
            viewdirs = embeddirs_fn(viewdirs)
        pts = torch.cat([pts, viewdirs], -1)
    out = batchify(fn, netchunk)(pts, viewdirs is None)
    return out.reshape(list(pts.shape[:-1]) + [out.shape[-1]])


def raw2outputs(raw, z_vals, dists, white_bkgd=False, net_type="v2"):
    if not torch.is_grad_enabled() and raw.is_cuda and raw.shape[-1] == 4 and z_vals.shape[1] % 32 == 0:
        rgb_map, depth_map, weights, alpha = ops.composite(raw.float().contiguous(), z_vals.float().contiguous())
    elif raw.is_cuda and raw.shape[-1] == 4 and z_vals.shape[1] % 32 == 0:
        from rcmvs import autograd as A
        rgb_map, depth_map, weights, alpha = A.composite(raw, z_vals)                   # differentiable, own kernels
    else:
        library_fallback("raw2outputs (CPU tensor / unusual shape)", "ATen cumprod / reductions")
        alpha, weights, _ = raw2alpha(raw[..., 3], dists, net_type)
        rgb_map = torch.sum(weights[..., None] * raw[..., :3], -2)
        depth_map = torch.sum(weights * z_vals, -1)
    acc_map = torch.sum(weights, -1)
    disp_map = 1.0 / torch.max(1e-10 * torch.ones_like(depth_map), depth_map / acc_map)
    if white_bkgd:
        rgb_map = rgb_map + (1.0 - acc_map[..., None])
    return rgb_map, disp_map, acc_map, weights, depth_map, alpha


def gen_dir_feature(w2c_ref, rays_dir):
    return rays_dir @ w2c_ref[:3, :3].t()


def gen_pts_feats(imgs, volume_feature, rays_pts, pose_ref, rays_ndc, feat_dim, img_feat=None, img_downscale=1.0,
                  use_color_volume=False, net_type="v0", volume_cl=None):
    """[N,S,20] = [8-ch volume lookup, 3 x (RGB + in-bounds mask)].  `volume_cl` ([D,h,w,8] channels-last) selects
    the fused CUDA gather; otherwise the torch-ops (differentiable) path of the reference is used."""
    if volume_cl is not None and not torch.is_grad_enabled():
        return ops.point_features(volume_cl, rays_ndc.contiguous(), rays_pts.contiguous(), imgs[0].contiguous(),
                                  pose_ref["w2cs"][:3].float().contiguous(),
                                  pose_ref["intrinsics"][:3].float().contiguous())
    if volume_cl is not None:
        from rcmvs import autograd as A
        return A.point_features(volume_cl, rays_ndc.contiguous(), rays_pts.contiguous(), imgs[0].contiguous(),
                                pose_ref["w2cs"][:3].float().contiguous(), pose_ref["intrinsics"][:3].float().contiguous())
    library_fallback("gen_pts_feats (planar volume input)", "ATen grid_sample")
    ray_feats = index_point_feature(volume_feature, rays_ndc)
    return torch.cat([ray_feats, build_color_volume(rays_pts, pose_ref, imgs, img_feat, with_mask=True,
                                                    downscale=img_downscale)], -1)


def rendering(args, pose_ref, rays_pts, rays_ndc, depth_candidates, rays_o, rays_dir, volume_feature=None, imgs=None,
              network_fn=None, img_feat=None, network_query_fn=None, white_bkgd=False, volume_cl=None, **kwargs):
    cos_angle = torch.norm(rays_dir, dim=-1)
    angle = gen_dir_feature(pose_ref["w2cs"][0], rays_dir / cos_angle.unsqueeze(-1))
    input_feat = gen_pts_feats(imgs, volume_feature, rays_pts, pose_ref, rays_ndc, args.feat_dim, img_feat,
                               args.img_downscale, args.use_color_volume, args.net_type, volume_cl=volume_cl)
    raw = network_query_fn(rays_ndc, angle, input_feat, network_fn)
    dists = depth2dist(depth_candidates, cos_angle)
    rgb_map, disp_map, acc_map, weights, depth_map, alpha = raw2outputs(raw, depth_candidates, dists, white_bkgd,
                                                                        args.net_type)
    return rgb_map, input_feat, weights, depth_map, alpha, {}
