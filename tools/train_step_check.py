"""Training-path check on GPU (BASELINE configs[3] in miniature): one unsupervised-style optimisation step through
the drop-in CascadeMVSNet (train class, autograd path with the CUDA warp + its scatter-add backward), the CUDA
inverse_warping (forward + depth gradient) and Rendering_Consistency_Net (torch-op autograd path + CUDA sampler).
Optionally under DDP (torchrun) to exercise the NCCL gradient all-reduce.  Prints loss before/after."""
import os, sys
sys.path[:0] = [os.path.join(os.path.dirname(__file__), ".."), os.path.join(os.path.dirname(__file__), "..", "rc-mvsnet_b200")]
import numpy as np, torch, torch.nn.functional as F
import models
from losses.homography import inverse_warping
from rcmvs import synth, weights

G = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(1 + rank)
H = W = 64
cas = models.CascadeMVSNet(ndepths=[16, 8, 8]); weights.load_model(cas, os.path.join(G, "cas_model_000014.npz")); cas.to(dev).train()
nerf = models.Rendering_Consistency_Net(synth.RenderArgs()); nerf.load_state_dict(weights.load_npz_state_dict(os.path.join(G, "nerf_model_000014.npz"))); nerf.to(dev).train()
if world > 1:
    cas = torch.nn.parallel.DistributedDataParallel(cas, device_ids=[local]); nerf = torch.nn.parallel.DistributedDataParallel(nerf, device_ids=[local])
opt = torch.optim.Adam(list(cas.parameters()) + list(nerf.parameters()), lr=1e-4)
imgs, proj, dv = synth.mvs_inputs(4, H, W, kind="textured", seed=5 + rank)
batch = synth.render_batch(H, W, seed=rank)
cu = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
imgs_t, proj_t, dv_t = cu(imgs), {k: cu(v) for k, v in proj.items()}, cu(dv)

def step():
    opt.zero_grad()
    out, vol = cas(imgs_t, proj_t, dv_t)
    loss = 0
    for s, k in enumerate(("stage1", "stage2", "stage3")):          # photometric term via the CUDA inverse_warping
        d = out[k]["depth"]
        h, w = d.shape[-2:]
        ref = F.interpolate(imgs_t[:, 0], (h, w), mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
        for v in range(1, 4):
            src = F.interpolate(imgs_t[:, v], (h, w), mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
            warped, mask = inverse_warping(src.contiguous(), proj_t[k][:, 0], proj_t[k][:, v], d)
            loss = loss + ((warped - ref).abs() * mask).mean()
    b = {k: cu(v) for k, v in batch.items()}
    rgb, feat, wts, depth_pred, alpha, ret, rays_depth, target = nerf(vol, out["depth"].detach(), b)
    loss = loss + F.mse_loss(rgb, target) + 1e-3 * F.smooth_l1_loss(depth_pred, rays_depth)
    loss.backward()
    gn = sum(float(p.grad.norm()) ** 2 for p in list(cas.parameters()) + list(nerf.parameters()) if p.grad is not None) ** 0.5
    opt.step()
    return float(loss), gn

l0, g0 = step(); l1, g1 = step(); l2, g2 = step()
torch.cuda.synchronize()
print("rank %d/%d loss %.5f -> %.5f -> %.5f  grad-norm %.3e  finite=%s" % (rank, world, l0, l1, l2, g0, np.isfinite([l0, l1, l2, g0]).all()))
if world > 1:
    p0 = next(cas.parameters()).detach().flatten()[:4].clone(); ref = p0.clone(); dist.broadcast(ref, 0)
    print("rank %d params in sync after DDP steps: %s" % (rank, torch.equal(p0, ref)))
    dist.destroy_process_group()
