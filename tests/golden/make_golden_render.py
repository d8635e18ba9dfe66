"""Golden vectors for the rendering-consistency branch (a7-a13), produced by RUNNING THE UNMODIFIED REFERENCE on CPU.

    PYTHONPATH=/root/reference python tests/golden/make_golden_render.py

CPU shims: Tensor.cuda is a no-op and BatchNorm2d accepts 5-D input inside Neural_Volume_Net (what
SyncBatchNorm conversion does implicitly in train_rcmvsnet.py:524-525).  Random sampling is seeded; the sampled
pixels / depth candidates are stored so the CUDA path is checked on identical samples.
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "rc-mvsnet_b200"))
sys.path.insert(0, "/root/reference")
sys.argv = sys.argv[:1]
torch.Tensor.cuda = lambda self, *a, **k: self            # shim 1


class _BN2dAnyRank(nn.BatchNorm3d):                        # shim 2: BatchNorm2d applied to 5-D tensors
    def _check_input_dim(self, input):
        return


nn.BatchNorm2d = _BN2dAnyRank                              # before the reference binds it as a default argument
from rcmvs import synth  # noqa: E402
import models.render_models as rm  # noqa: E402  (reference)
from models.render_consist_net import Rendering_Consistency_Net  # noqa: E402
import models.render_consist_net as rcn  # noqa: E402
import models.renderer as rr  # noqa: E402
from models.casmvsnet import CascadeMVSNet  # noqa: E402
from losses.homography import inverse_warping  # noqa: E402

torch.set_num_threads(8)


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def save(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrs)
    print("wrote", path, "%.2f MB" % (os.path.getsize(path) / 1e6))


ck_nerf = torch.load("/root/reference/pretrain/model_000014_nerf.ckpt", map_location="cpu")["model"]
save("nerf_model_000014.npz", **{k: v.numpy() for k, v in ck_nerf.items()})
ck_cas = torch.load("/root/reference/pretrain/model_000014_cas.ckpt", map_location="cpu")["model"]

H = W = 64
batch = synth.render_batch(H, W, seed=0)
imgs_mvs, proj, dv = synth.mvs_inputs(4, H, W, kind="textured", seed=5)
cas = CascadeMVSNet(ndepths=[16, 8, 8])
cas.load_state_dict(ck_cas, strict=True)
cas.eval()
cas.DepthNet.train()
with torch.no_grad():
    out, vol = cas(t(imgs_mvs), {k: t(v) for k, v in proj.items()}, t(dv))
pseudo_depth = out["depth"].detach()                       # [1,H,W]

args = synth.RenderArgs()
net = Rendering_Consistency_Net(args)
net.load_state_dict(ck_nerf, strict=True)
net.eval()

cap = {}
orig_build = rcn.build_rays_norm
orig_render = rcn.rendering


def spy_build(*a, **k):
    r = orig_build(*a, **k)
    names = ["rays_pts", "rays_dir", "target_s", "rays_NDC", "depth_candidates", "rays_o", "rays_depth"]
    for n, v in zip(names, r[:7]):
        cap[n] = v.detach().numpy().copy()
    return r


orig_get_rays = rcn.get_rays_mvs if hasattr(rcn, "get_rays_mvs") else None
import models.render_utils as ru  # noqa: E402
_orig_get_rays = ru.get_rays_mvs


def spy_rays(*a, **k):
    r = _orig_get_rays(*a, **k)
    cap["pixel_coordinates"] = r[2].numpy().copy()        # [2,N] (ys, xs)
    return r


ru.get_rays_mvs = spy_rays
rcn.build_rays_norm = spy_build
vol_hook = {}
net.MVSNet.register_forward_hook(lambda m, i, o: vol_hook.__setitem__("volume", o.detach().numpy().copy()))
orig_query = net.network_query_fn


def spy_query(pts, viewdirs, feats, fn):
    raw = orig_query(pts, viewdirs, feats, fn)
    cap["raw"] = raw.detach().numpy().copy()
    cap["angle"] = viewdirs.detach().numpy().copy()
    return raw


net.network_query_fn = spy_query
torch.manual_seed(123)
b = {k: t(v) for k, v in batch.items()}
with torch.no_grad():
    rgb, input_feat, weights, depth_pred, alpha, ret, rays_depth, target_s = net(vol, pseudo_depth, b)

NR = 96     # rays stored for the per-point tensors (first 48 Gaussian + 48 uniform rays)
sel = np.r_[0:48, 512:560]
save("render_small.npz",
     imgs=batch["imgs"][0], w2cs=batch["w2cs"][0], c2ws=batch["c2ws"][0], near_fars=batch["near_fars"][0],
     intrinsics=batch["intrinsics"][0], depths_h=batch["depths_h"][0], proj_mats=batch["proj_mats"][0],
     volume_feature_in=vol[0].numpy(), pseudo_depth=pseudo_depth[0].numpy(), volume=vol_hook["volume"][0],
     pixel_coordinates=cap["pixel_coordinates"], depth_candidates=cap["depth_candidates"],
     rays_pts=cap["rays_pts"][sel], rays_NDC=cap["rays_NDC"][sel], rays_dir=cap["rays_dir"], target_s=cap["target_s"],
     rays_depth=cap["rays_depth"], angle=cap["angle"], sel=sel,
     input_feat=input_feat.numpy()[sel], raw=cap["raw"][sel], rgb=rgb.numpy(), depth_pred=depth_pred.numpy(),
     weights=weights.numpy()[sel], alpha=alpha.numpy()[sel])

# ---- a13 inverse_warping incl. gradient w.r.t. depth ---------------------------------------------------------------
rng = np.random.default_rng(4)
h, w = 24, 32
cams = synth.cameras(2, h * 4, w * 4, baseline=40.0)["stage1"]          # [1,2,2,4,4]
img = rng.standard_normal((1, h, w, 3), dtype=np.float32)
depth = (650 + 30 * rng.standard_normal((1, h, w))).astype(np.float32)
depth[0, :2, :3] = 5.0                                                    # force some out-of-image projections
dt = t(depth).requires_grad_(True)
warped, mask = inverse_warping(t(img), t(cams[:, 0]), t(cams[:, 1]), dt)
gw = t(rng.standard_normal(warped.shape, dtype=np.float32))
(warped * gw).sum().backward()
save("inverse_warp_small.npz", img=img, left_cam=cams[:, 0], right_cam=cams[:, 1], depth=depth,
     warped=warped.detach().numpy(), mask=mask.numpy(), grad_out=gw.numpy(), grad_depth=dt.grad.numpy())
