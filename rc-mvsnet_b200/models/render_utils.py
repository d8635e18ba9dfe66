"""Drop-in for the hot-path part of the reference's models/render_utils.py (ray generation :86-108, NDC :112-146,
Gaussian-Uniform sampling :149-243, colour gathering :247-279, volume lookup :304-330, batch helpers :365-381).
Visualisation / PFM / render-path helpers of that file are outside the path and stay in the reference tree."""
import torch
import torch.nn.functional as F

from rcmvs import ops

__all__ = ["get_rays_mvs", "get_ndc_coordinate", "build_rays_norm", "build_color_volume", "index_point_feature",
           "normal_vect", "filter_keys", "sub_selete_data", "img2mse", "mse2psnr"]

img2mse = lambda x, y: torch.mean((x - y) ** 2)  # noqa: E731
mse2psnr = lambda x: -10. * torch.log(x) / torch.log(torch.Tensor([10.]))  # noqa: E731


def normal_vect(vect, dim=-1):
    return vect / (torch.sqrt(torch.sum(vect ** 2, dim=dim, keepdim=True)) + 1e-7)


def sample_pixels(H, W, N, device):
    """N uniform-random integer pixels, drawn on the CPU generator in the reference's order (xs, then ys) so that a
    seeded run picks bit-identical pixels (render_utils.py:93).  Returns float tensors (ys, xs) on `device`."""
    xs = torch.randint(0, W, (N,)).float()
    ys = torch.randint(0, H, (N,)).float()
    return ys.to(device), xs.to(device)


def get_rays_mvs(H, W, intrinsic, c2w, N=1024, isRandom=True, is_precrop_iters=False, chunk=-1, idx=-1):
    device = c2w.device
    if isRandom:
        ys, xs = sample_pixels(H, W, N, device)
    else:
        ys, xs = torch.meshgrid(torch.linspace(0, H - 1, H), torch.linspace(0, W - 1, W), indexing="ij")
        ys, xs = ys.reshape(-1), xs.reshape(-1)
        if chunk > 0:
            ys, xs = ys[idx * chunk:(idx + 1) * chunk], xs[idx * chunk:(idx + 1) * chunk]
        ys, xs = ys.to(device), xs.to(device)
    dirs = torch.stack([(xs - intrinsic[0, 2]) / intrinsic[0, 0], (ys - intrinsic[1, 2]) / intrinsic[1, 1],
                        torch.ones_like(xs)], -1)
    return c2w[:3, -1].clone(), dirs @ c2w[:3, :3].t(), torch.stack((ys, xs))


def get_ndc_coordinate(w2c_ref, intrinsic_ref, point_samples, inv_scale, near=2, far=6, pad=0, lindisp=False):
    N_rays, N_samples = point_samples.shape[:2]
    p = point_samples.reshape(-1, 3)
    if w2c_ref is not None:
        p = torch.matmul(p, w2c_ref[:3, :3].t()) + w2c_ref[:3, 3:].reshape(1, 3)
    q = p @ intrinsic_ref.t()
    xy = q[:, :2] / q[:, -1:] / inv_scale.reshape(1, 2)
    z = (q[:, 2:] - near) / (far - near) if not lindisp else (1.0 / q[:, 2:] - 1.0 / near) / (1.0 / far - 1.0 / near)
    return torch.cat([xy, z], -1).view(N_rays, N_samples, 3)


def gaussian_uniform_candidates(rays_depth, near, far, N_samples):
    """[N] pseudo depths -> [N,N_samples]: first half Gaussian around the depth (sorted), second half stratified
    uniform (render_utils.py:203-224) — one CUDA kernel, seeded from torch's CPU generator."""
    seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    N = rays_depth.shape[0]
    return ops.gu_sample(rays_depth.detach().float().contiguous(), float(near), float(far), seed, N - N // 2, N_samples)


def build_rays_norm(imgs, depths, pose_ref, w2cs, c2ws, intrinsics, near_fars, N_rays, N_samples, pad=0,
                    is_precrop_iters=False, ref_idx=0, importanceSampling=False, with_depth=False, is_volume=False,
                    samples=None):
    """Same outputs as the reference (render_utils.py:149-243) for batch size 1.  `samples` optionally injects
    {'pixel_coordinates': [2,N] (ys,xs), 'depth_candidates': [N,S]} (used by the parity tests)."""
    device = imgs.device
    N, V, C, H, W = imgs.shape
    w2c_ref, K_ref = pose_ref["w2cs"][ref_idx], pose_ref["intrinsics"][ref_idx]
    near_ref, far_ref = pose_ref["near_fars"][ref_idx, 0], pose_ref["near_fars"][ref_idx, 1]
    K, c2w = intrinsics[0], c2ws[0]
    if samples is not None and "pixel_coordinates" in samples:
        ys, xs = (t.float().to(device).contiguous() for t in samples["pixel_coordinates"])
    else:
        ys, xs = sample_pixels(H, W, N_rays, device)
    yi, xi = ys.long(), xs.long()
    colors = imgs[0, 0][:, yi, xi].permute(1, 0)                         # [N_rays,3]
    rays_depth = depths[0, 0][yi, xi]
    near, far = near_fars[0, 0, 0], near_fars[0, 0, 1]
    if samples is not None and "depth_candidates" in samples:
        z = samples["depth_candidates"].float().to(device).contiguous()
    else:
        z = gaussian_uniform_candidates(rays_depth, near, far, N_samples)
    pts, ndc, dirs = ops.ray_points(ys, xs, K.float().contiguous(), c2w.float().contiguous(),
                                    w2c_ref.float().contiguous(), z, float(near_ref), float(far_ref), W, H)
    rays_o = c2w[:3, -1].reshape(3, 1).expand(3, ys.shape[0])
    inv_scale = torch.tensor([W - 1, H - 1], device=device)
    ndc_parameters = {"w2c_ref": w2c_ref, "intrinsic_ref": K_ref, "inv_scale": inv_scale, "near": near_ref,
                      "far": far_ref}
    return pts, dirs, colors, ndc, z, rays_o, rays_depth, ndc_parameters


def build_color_volume(point_samples, pose_ref, imgs, img_feat=None, downscale=1.0, with_mask=False):
    """torch-ops version (differentiable) of render_utils.py:247-279; the fused CUDA gather is ops.point_features."""
    N, V, C, H, W = imgs.shape
    inv_scale = torch.tensor([W - 1, H - 1], device=imgs.device)
    out = []
    for idx in range(V):
        ndc = get_ndc_coordinate(pose_ref["w2cs"][idx], pose_ref["intrinsics"][idx], point_samples, inv_scale)[None]
        grid = ndc[..., :2] * 2.0 - 1.0
        data = F.grid_sample(imgs[:, idx], grid, align_corners=True, mode="bilinear", padding_mode="border")
        if with_mask:
            m = (grid > -1.0) * (grid < 1.0)
            data = torch.cat((data, (m[..., 0] * m[..., 1]).float().unsqueeze(1)), dim=1)
        out.append(data[0].permute(1, 2, 0))
    return torch.cat(out, -1)


def index_point_feature(volume_feature, ray_coordinate_ref, chunk=-1):
    """torch-ops version of render_utils.py:304-330 ([B,C,D,h,w] volume, [N,S,3] NDC -> [N,S,C])."""
    H, W = ray_coordinate_ref.shape[-3:-1]
    grid = ray_coordinate_ref.view(-1, 1, H, W, 3) * 2 - 1.0
    return F.grid_sample(volume_feature, grid, align_corners=True, mode="bilinear")[:, :, 0].permute(2, 3, 0, 1).squeeze()


def filter_keys(d):
    d.pop("N_samples")
    d.pop("ndc", None)
    d.pop("lindisp", None)
    return d


def sub_selete_data(data_batch, device, idx, filtKey=[], filtIndex=None):
    out = {}
    for k, v in data_batch.items():
        if k in ("imgs_aug", "proj_matrices", "depth", "depth_values", "mask", "center_imgs"):
            continue
        v = v.float()
        out[k] = v if v.is_cuda else v.to(device)
    return out
