"""Generate golden vectors by RUNNING THE UNMODIFIED REFERENCE (authoring container only).

    PYTHONPATH=/root/reference python tests/golden/make_golden.py

/root/reference does not exist on the GPU box, so the outputs (*.npz next to this file) are committed.
The inputs come from rcmvs.synth (seeded); the weights are the reference's shipped pretrain/model_000014_cas.ckpt
('model' entry only, converted to numpy -> cas_model_000014.npz, a data fixture, not source).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "rc-mvsnet_b200"))
sys.path.insert(0, "/root/reference")
sys.argv = sys.argv[:1]

from rcmvs import synth  # noqa: E402
from models.modules import homo_warping  # noqa: E402  (reference)
from models.casmvsnet import CascadeMVSNet_eval, CascadeMVSNet  # noqa: E402  (reference)

torch.manual_seed(0)
np.random.seed(0)
torch.set_num_threads(8)


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def save(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrs)
    print("wrote", path, "%.2f MB" % (os.path.getsize(path) / 1e6))


# ---- weights -----------------------------------------------------------------------------------------------
ck = torch.load("/root/reference/pretrain/model_000014_cas.ckpt", map_location="cpu")["model"]
save("cas_model_000014.npz", **{k: v.numpy() for k, v in ck.items()})

# ---- a1: homo_warping ---------------------------------------------------------------------------------------
rng = np.random.default_rng(1)
C, h, w, D, V = 8, 24, 32, 6, 3
proj = synth.cameras(V, h * 4, w * 4, baseline=60.0)["stage1"]          # [1,V,2,4,4] at this resolution
projs = synth.stage_projections(proj)[0]                                  # [V,4,4]
feats = rng.standard_normal((V, C, h, w), dtype=np.float32)
base = (425 + 2.65 * 192 * np.linspace(0.02, 0.98, D)).astype(np.float32)
depth_pp = (base[:, None, None] + 6.0 * rng.standard_normal((D, h, w))).astype(np.float32)
outs_pp, outs_1d = [], []
for v in range(1, V):
    outs_pp.append(homo_warping(t(feats[v:v + 1]), t(projs[v:v + 1]), t(projs[0:1]), t(depth_pp[None]))[0].numpy())
    outs_1d.append(homo_warping(t(feats[v:v + 1]), t(projs[v:v + 1]), t(projs[0:1]), t(base[None]))[0].numpy())
# a grazing/out-of-bounds heavy case: large baseline and a depth plane behind the source camera
proj_far = synth.cameras(2, h * 4, w * 4, baseline=400.0)["stage1"]
projs_far = synth.stage_projections(proj_far)[0]
depth_far = np.array([-50.0, 1.0, 120.0, 500.0, 900.0, 5000.0], np.float32)
out_far = homo_warping(t(feats[1:2]), t(projs_far[1:2]), t(projs_far[0:1]), t(depth_far[None]))[0].numpy()
save("warp_small.npz", feats=feats, projs=projs, depth_pp=depth_pp, depth_1d=base,
     out_pp=np.stack(outs_pp), out_1d=np.stack(outs_1d),
     projs_far=projs_far, depth_far=depth_far, out_far=out_far)

# ---- a2-a6: CascadeMVSNet_eval with intermediates ------------------------------------------------------------
def run_eval(H, W, V, ndepths, kind, seed):
    imgs, proj, dv = synth.mvs_inputs(V, H, W, kind=kind, seed=seed)
    model = CascadeMVSNet_eval(ndepths=list(ndepths))
    model.load_state_dict(ck, strict=True)
    model.eval()
    cap = {"features": [], "hyp": [], "variance": [], "cost": []}
    def fhook(m, i, o):
        cap["features"].append({k: v.numpy().copy() for k, v in o.items()})
    model.feature.register_forward_hook(fhook)
    for cr in model.cost_regularization:
        def crhook(m, i, o):
            cap["variance"].append(i[0][0].numpy().copy())
            cap["cost"].append(o[0, 0].numpy().copy())
        cr.register_forward_hook(crhook)
    orig = model.DepthNet.forward

    def spy(features, proj_matrices, depth_values, **kw):
        cap["hyp"].append(depth_values[0].numpy().copy())
        return orig(features, proj_matrices, depth_values=depth_values, **kw)
    model.DepthNet.forward = spy
    with torch.no_grad():
        out = model(t(imgs), {k: t(v) for k, v in proj.items()}, t(dv))
    arrs = {"imgs": imgs[0], "depth_values": dv[0]}
    for s in range(3):
        k = "stage%d" % (s + 1)
        arrs["proj_" + k] = proj[k][0]
        arrs["hyp_" + k] = cap["hyp"][s]
        arrs["variance_" + k] = cap["variance"][s]
        arrs["cost_" + k] = cap["cost"][s]
        arrs["depth_" + k] = out[k]["depth"][0].numpy()
        arrs["conf_" + k] = out[k]["photometric_confidence"][0].numpy()
        for v in range(V):
            arrs["feat_%s_v%d" % (k, v)] = cap["features"][v][k][0]
    return arrs


save("mvs_small.npz", **run_eval(64, 64, 3, (16, 8, 8), "textured", 0))
a = run_eval(64, 96, 4, (16, 8, 8), "noise", 3)
# the noise case only pins depth/conf + hypotheses (keeps the fixture small)
save("mvs_small_noise.npz", **{k: v for k, v in a.items() if not (k.startswith("variance") or k.startswith("cost_") or k.startswith("feat_"))})

# ---- train-mode DepthNet extras (volume_feature_no_ref) -------------------------------------------------------
imgs, proj, dv = synth.mvs_inputs(4, 64, 64, kind="textured", seed=5)
model = CascadeMVSNet(ndepths=[16, 8, 8])
model.load_state_dict(ck, strict=True)
model.eval()            # BatchNorm uses running statistics (deterministic fixture) ...
model.DepthNet.train()  # ... but DepthNet follows its training branch (casmvsnet.py:86-90): the eval branch squares
                        # warped_volume in place before the no-ref sums (:94-96) and its result is discarded upstream
with torch.no_grad():
    out, vol = model(t(imgs), {k: t(v) for k, v in proj.items()}, t(dv))
save("mvs_train_small.npz", imgs=imgs[0], depth_values=dv[0], **{"proj_" + k: v[0] for k, v in proj.items()},
     volume_feature_no_ref=vol[0].numpy(), depth_stage1=out["stage1"]["depth"][0].numpy(),
     depth_stage3=out["stage3"]["depth"][0].numpy())

# ---- BASELINE configs[1] at full size, run by the reference itself: 5 views 512x640, D=192/64/8 ------------------
a = run_eval(512, 640, 5, (192, 64, 8), "textured", 4)
save("mvs_full_192_64_8.npz", **{k: v for k, v in a.items() if k.startswith("depth_stage")},
     conf_stage3=a["conf_stage3"].astype(np.float16))
