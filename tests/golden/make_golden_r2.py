"""Round-2 golden vectors, produced by RUNNING THE UNMODIFIED REFERENCE on CPU (authoring container only).

    PYTHONPATH=/root/reference python tests/golden/make_golden_r2.py [idx] [tanks] [render_full] [geo] [loss] [train_grad] [mask]

  idx          depth_index of the three stages of mvs_small.npz, computed with the reference's own op sequence
               (models/casmvsnet.py:111-121: softmax -> depth_regression with arange -> .long() -> clamp)
  tanks        BASELINE configs[4]: 7 views 1056x1920, cascade 48/32/8 (eval_rcmvsnet_tanks.py:48-56 defaults)
  render_full  BASELINE configs[2]: CascadeMVSNet (train class) + Rendering_Consistency_Net at 4 views 512x640,
               48/32/8, 1024 rays x 128 samples; the sampled pixels / depth candidates are stored for injection
  geo          eval_rcmvsnet_dtu.py:281-338 reproject_with_depth / check_geometric_consistency on synthetic depth maps
  loss         losses/unsup_loss.py UnSupLoss (+ gradient w.r.t. depth) on a small synthetic case
  train_grad   the reference's FeatureNet and CostRegNet (stage 1, 32 input channels) in train() mode on CPU: outputs, the
               gradients of every parameter and of the input for fixed random cotangents, BatchNorm running statistics
               after the step — what pins the torch checker of the GPU gradient tests (tests/test_gpu_r2.py) to the reference
  mask         losses/aug_loss.py random_image_mask under fixed np.random seeds (host RNG call order of the training step)
  tanks_sensitivity  (prints only) the reference re-run on ~1-ulp-perturbed images: its own conditioning at cfg 5

Inputs always come from rcmvs.synth with the seeds written next to each case, so the GPU tests regenerate them.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "rc-mvsnet_b200"))
sys.path.insert(0, "/root/reference")
what = set(sys.argv[1:]) or {"idx", "tanks", "render_full", "geo", "loss"}
sys.argv = sys.argv[:1]
torch.Tensor.cuda = lambda self, *a, **k: self            # CPU shim 1


class _BN2dAnyRank(nn.BatchNorm3d):                        # CPU shim 2: BatchNorm2d on 5-D input (what SyncBN does)
    def _check_input_dim(self, input):
        return


nn.BatchNorm2d = _BN2dAnyRank                              # before the reference binds it as a default argument
from rcmvs import synth  # noqa: E402

torch.set_num_threads(8)


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def save(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrs)
    print("wrote", path, "%.2f MB" % (os.path.getsize(path) / 1e6), flush=True)


ck_cas = torch.load("/root/reference/pretrain/model_000014_cas.ckpt", map_location="cpu")["model"]

# ---- idx ---------------------------------------------------------------------------------------------------------
if "idx" in what:
    import torch.nn.functional as F
    from models.modules import depth_regression  # reference
    g = dict(np.load(os.path.join(HERE, "mvs_small.npz")))
    out = {}
    for s in (1, 2, 3):
        cost = t(g["cost_stage%d" % s])[None]                              # [1,D,h,w]
        D = cost.shape[1]
        prob = F.softmax(cost, dim=1)
        idx = depth_regression(prob, depth_values=torch.arange(D, dtype=torch.float)).long().clamp(min=0, max=D - 1)
        out["idx_stage%d" % s] = idx[0].numpy().astype(np.int32)
    save("idx_small.npz", **out)

# ---- tanks -------------------------------------------------------------------------------------------------------
if "tanks" in what:
    from models.casmvsnet import CascadeMVSNet_eval  # reference
    V, H, W, nd = 7, 1056, 1920, (48, 32, 8)
    imgs, proj, dv = synth.mvs_inputs(V, H, W, kind="textured", seed=21)
    m = CascadeMVSNet_eval(ndepths=list(nd))
    m.load_state_dict(ck_cas, strict=True)
    m.eval()
    with torch.no_grad():
        out = m(t(imgs), {k: t(v) for k, v in proj.items()}, t(dv))
    save("tanks_7x1056x1920_48_32_8.npz", seed=np.int64(21),
         depth_stage1=out["stage1"]["depth"][0].numpy(), depth_stage2=out["stage2"]["depth"][0].numpy(),
         depth_stage3=out["stage3"]["depth"][0].numpy(),
         conf_stage3=out["stage3"]["photometric_confidence"][0].numpy().astype(np.float16))
    del m, out, imgs

# ---- render_full ---------------------------------------------------------------------------------------------------
if "render_full" in what:
    from models.casmvsnet import CascadeMVSNet  # reference
    from models.render_consist_net import Rendering_Consistency_Net
    import models.render_consist_net as rcn
    import models.render_utils as ru
    ck_nerf = torch.load("/root/reference/pretrain/model_000014_nerf.ckpt", map_location="cpu")["model"]
    H, W = 512, 640
    batch = synth.render_batch(H, W, seed=3)
    imgs_mvs, proj, dv = synth.mvs_inputs(4, H, W, kind="textured", seed=11)
    cas = CascadeMVSNet(ndepths=[48, 32, 8])
    cas.load_state_dict(ck_cas, strict=True)
    cas.eval()
    cas.DepthNet.train()          # training branch of DepthNet (the eval branch corrupts the no-ref sums, casmvsnet.py:94-96)
    with torch.no_grad():
        out, vol = cas(t(imgs_mvs), {k: t(v) for k, v in proj.items()}, t(dv))
    pseudo_depth = out["depth"].detach()
    net = Rendering_Consistency_Net(synth.RenderArgs())
    net.load_state_dict(ck_nerf, strict=True)
    net.eval()
    cap = {}
    orig_build = rcn.build_rays_norm
    orig_rays = ru.get_rays_mvs

    def spy_build(*a, **k):
        r = orig_build(*a, **k)
        cap["depth_candidates"] = r[4].detach().numpy().copy()
        cap["rays_depth"] = r[6].detach().numpy().copy()
        cap["target_s"] = r[2].detach().numpy().copy()
        return r

    def spy_rays(*a, **k):
        r = orig_rays(*a, **k)
        cap["pixel_coordinates"] = r[2].numpy().copy()
        return r
    rcn.build_rays_norm = spy_build
    ru.get_rays_mvs = spy_rays
    torch.manual_seed(321)
    with torch.no_grad():
        rgb, input_feat, weights, depth_pred, alpha, ret, rays_depth, target_s = net(
            vol, pseudo_depth, {k: t(v) for k, v in batch.items()})
    save("render_full_512x640.npz", mvs_seed=np.int64(11), batch_seed=np.int64(3),
         pseudo_depth=pseudo_depth[0].numpy(), depth_stage1=out["stage1"]["depth"][0].numpy(),
         pixel_coordinates=cap["pixel_coordinates"], depth_candidates=cap["depth_candidates"],
         rays_depth=cap["rays_depth"], target_s=cap["target_s"], rgb=rgb.numpy(), depth_pred=depth_pred.numpy(),
         weights=weights.numpy().astype(np.float16), alpha=alpha.numpy().astype(np.float16),
         input_feat_first64=input_feat.numpy()[:64])
    rcn.build_rays_norm, ru.get_rays_mvs = orig_build, orig_rays

# ---- geo (f3) --------------------------------------------------------------------------------------------------------
if "geo" in what:
    # eval_rcmvsnet_dtu.py imports plyfile / matplotlib / PIL at module level and parses argv: stub what is missing here
    for name in ("plyfile", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["plyfile"].PlyData = sys.modules["plyfile"].PlyElement = object
    src = open("/root/reference/eval_rcmvsnet_dtu.py").read()
    start = src.index("# project the reference point cloud into the source view")
    end = src.index("def filter_depth(")
    ns = {"np": np, "cv2": __import__("cv2")}
    exec(compile(src[start:end], "eval_rcmvsnet_dtu.py[281-338]", "exec"), ns)    # the two reference functions, unmodified
    rng = np.random.default_rng(17)
    H, W, NS = 96, 128, 4
    proj = synth.cameras(NS + 1, H, W)["stage3"][0]                               # [V,2,4,4] full-res cameras
    ext = proj[:, 0].astype(np.float32)
    K = proj[:, 1, :3, :3].astype(np.float32)
    # depth maps: the plane z=650 of the reference frame seen from every view + noise + a few gross outliers
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    depths = []
    for v in range(NS + 1):
        T = ext[0].astype(np.float64) @ np.linalg.inv(ext[v].astype(np.float64))   # cam v -> ref cam
        d = np.stack([xs, ys, np.ones_like(xs)], 0).reshape(3, -1).astype(np.float64)
        d = T[:3, :3] @ (np.linalg.inv(K[v].astype(np.float64)) @ d)
        tt = (650.0 - T[2, 3]) / d[2]
        # depth along the view-v camera z axis of the hit point = tt (ray dir has z=1 in cam v)
        dep = tt.reshape(H, W) + 0.4 * rng.standard_normal((H, W))
        out_mask = rng.uniform(size=(H, W)) < 0.05
        dep = np.where(out_mask, dep * rng.uniform(0.7, 1.4, (H, W)), dep)
        depths.append(dep.astype(np.float32))
    res = {"ext": ext, "K": K, "depths": np.stack(depths)}
    for v in range(1, NS + 1):
        m, dr, x2, y2 = ns["check_geometric_consistency"](depths[0], K[0], ext[0], depths[v], K[v], ext[v], 0.5, 0.01)
        res["mask_%d" % v] = m
        res["depth_reproj_%d" % v] = dr.astype(np.float32)
        res["x2d_src_%d" % v] = x2.astype(np.float32)
        res["y2d_src_%d" % v] = y2.astype(np.float32)
    # eval_rcmvsnet_dtu.py:385-389 with the script defaults (prob_thres 0.8, num_consistency 3)
    conf = rng.uniform(0.5, 1.0, (H, W)).astype(np.float32)
    geo_sum = sum(res["mask_%d" % v].astype(np.int32) for v in range(1, NS + 1))
    res["confidence"] = conf
    res["geo_mask_sum"] = geo_sum.astype(np.int32)
    res["depth_est_averaged"] = ((sum(res["depth_reproj_%d" % v] for v in range(1, NS + 1)) + depths[0]) /
                                 (geo_sum + 1)).astype(np.float32)
    res["final_mask"] = np.logical_and(conf > 0.8, geo_sum >= 3)
    save("geo_filter_small.npz", **res)

# ---- loss (f2) -------------------------------------------------------------------------------------------------------
if "loss" in what:
    from losses.unsup_loss import UnSupLoss, UnsupLossMultiStage  # reference
    rng = np.random.default_rng(23)
    V, H, W = 4, 64, 80
    imgs, proj, dv = synth.mvs_inputs(V, H, W, kind="textured", seed=31)
    imgs01 = (0.5 + 0.2 * imgs).clip(0, 1).astype(np.float32)                      # the loss sees un-normalised-ish images
    res = {"imgs": imgs01[0], "seed": np.int64(31)}
    for k in proj:
        res["proj_" + k] = proj[k][0]
    crit = UnSupLoss()
    depths = {}
    for s, key in enumerate(("stage1", "stage2", "stage3")):
        sc = (4, 2, 1)[s]
        h, w = H // sc, W // sc
        d = (650 + 15 * rng.standard_normal((1, h, w))).astype(np.float32)
        dt = t(d).requires_grad_(True)
        loss = crit(t(imgs01), t(proj[key]), dt, s)
        loss.backward()
        res["depth_" + key] = d[0]
        res["loss_" + key] = np.float32(loss.item())
        res["grad_" + key] = dt.grad[0].numpy()
        res["ssim_" + key] = np.float32(crit.ssim_loss.item())
        res["reconstr_" + key] = np.float32(crit.reconstr_loss.item())
        res["smooth_" + key] = np.float32(crit.smooth_loss.item())
        depths[key] = dt.detach()
    ms = UnsupLossMultiStage()
    inputs = {k: {"depth": depths[k].clone().requires_grad_(True)} for k in depths}
    total, scalars = ms(inputs, t(imgs01), {k: t(v) for k, v in proj.items()}, dlossw=[0.5, 1.0, 2.0])
    res["multistage_total"] = np.float32(total.item())
    save("unsup_loss_small.npz", **res)

# ---- tanks_sensitivity: how far does the REFERENCE move under ~1-ulp input noise? (justifies the cfg-5 tolerance) ------
if "tanks_sensitivity" in what:
    from models.casmvsnet import CascadeMVSNet_eval  # reference
    V, H, W, nd = 7, 1056, 1920, (48, 32, 8)
    imgs, proj, dv = synth.mvs_inputs(V, H, W, kind="textured", seed=21)
    rng = np.random.default_rng(0)
    imgs2 = (imgs.astype(np.float64) * (1 + 1e-7 * rng.standard_normal(imgs.shape))).astype(np.float32)
    m = CascadeMVSNet_eval(ndepths=list(nd))
    m.load_state_dict(ck_cas, strict=True)
    m.eval()
    with torch.no_grad():
        out = m(t(imgs2), {k: t(v) for k, v in proj.items()}, t(dv))
    g = dict(np.load(os.path.join(HERE, "tanks_7x1056x1920_48_32_8.npz")))
    for k in ("stage1", "stage2", "stage3"):
        d = out[k]["depth"][0].numpy()
        err = np.abs(g["depth_" + k] - d) / np.abs(d).max()
        print(k, "reference(x) vs reference(x·(1 + 1e-7·n)): rel L-inf %.3e, p99.9 %.3e" % (err.max(), np.percentile(err, 99.9)))
    # measured in the authoring container: stage1 1.111e-06 / 7.1e-07, stage2 1.497e-05 / 6.9e-06, stage3 1.641e-04 / 2.2e-05


# ---- train_grad ----------------------------------------------------------------------------------------------------
if "train_grad" in what:
    from models.casmvsnet import CascadeMVSNet  # reference (train class)
    torch.manual_seed(0)
    cas = CascadeMVSNet(ndepths=[16, 8, 8])
    cas.load_state_dict(ck_cas, strict=True)
    cas.train()
    rng = np.random.default_rng(11)
    out = {}
    # FeatureNet (reference models/modules.py:363-464), one "view" batch of 2 images
    x = t(rng.standard_normal((2, 3, 32, 40), dtype=np.float32)).requires_grad_(True)
    f = cas.feature(x)
    cot = {k: t(rng.standard_normal(tuple(v.shape), dtype=np.float32)) for k, v in f.items()}
    sum((f[k] * cot[k]).sum() for k in f).backward()
    out["feat_x"], out["feat_dx"] = x.detach().numpy(), x.grad.numpy()
    for k in f:
        out["feat_out_" + k], out["feat_cot_" + k] = f[k].detach().numpy(), cot[k].numpy()
    for n, p_ in cas.feature.named_parameters():
        out["feat_grad/" + n] = p_.grad.numpy().copy()
    for n, b in cas.feature.named_buffers():
        if n.endswith("running_mean") or n.endswith("running_var"):
            out["feat_buf/" + n] = b.numpy().copy()
    # CostRegNet of stage 1 (reference models/modules.py:470-501): 32 input channels
    reg = cas.cost_regularization[0]
    v = t(rng.standard_normal((1, 32, 8, 16, 24), dtype=np.float32)).requires_grad_(True)
    y = reg(v)
    cy = t(rng.standard_normal(tuple(y.shape), dtype=np.float32))
    (y * cy).sum().backward()
    out["reg_x"], out["reg_dx"], out["reg_out"], out["reg_cot"] = v.detach().numpy(), v.grad.numpy(), y.detach().numpy(), cy.numpy()
    for n, p_ in reg.named_parameters():
        out["reg_grad/" + n] = p_.grad.numpy().copy()
    for n, b in reg.named_buffers():
        if n.endswith("running_mean") or n.endswith("running_var"):
            out["reg_buf/" + n] = b.numpy().copy()
    save("train_grad_small.npz", **out)


# ---- mask ----------------------------------------------------------------------------------------------------------
if "mask" in what:
    from losses.aug_loss import random_image_mask  # reference
    rng = np.random.default_rng(3)
    img = t(rng.standard_normal((2, 3, 48, 60), dtype=np.float32))
    out = {"img": img.numpy(), "filter_size": np.array([16, 20])}
    for seed in (0, 5, 11):
        np.random.seed(seed)
        a, m = random_image_mask(img, (16, 20))
        b, m2 = random_image_mask(img, (16, 20))           # second draw from the same stream
        out["masked_%d" % seed], out["mask_%d" % seed] = a.numpy(), m.numpy().astype(np.uint8)
        out["mask2_%d" % seed] = m2.numpy().astype(np.uint8)
    np.random.seed(1)
    a, m = random_image_mask(img, (48, 60))                # filter = whole image: (img, None)
    out["full_is_none"] = np.array([m is None and a is img])
    save("aug_mask_small.npz", **out)
