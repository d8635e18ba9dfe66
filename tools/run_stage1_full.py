"""Run FeatureNet + the whole stage-1 regularisation (sweep, CostRegNet, soft-argmin) at BASELINE shapes a few times
(target for `ncu -k regex:...` captures of the individual TMA kernels)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rc-mvsnet_b200"))
from rcmvs import ops, synth, weights
import models
dev = torch.device("cuda:0")
V, H, W, D = 5, 512, 640, 192
imgs, proj, dv = synth.mvs_inputs(V, H, W, kind="textured", seed=0)
m = models.CascadeMVSNet_eval(ndepths=[192, 64, 8])
weights.load_model(m, os.path.join(ROOT, "tests", "golden", "cas_model_000014.npz"))
m.to(dev).eval()
args = (torch.from_numpy(imgs).to(dev), {k: torch.from_numpy(v).to(dev) for k, v in proj.items()}, torch.from_numpy(dv).to(dev))
with torch.no_grad():
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
        out = m(*args)
    torch.cuda.synchronize()
print("ok", tuple(out["depth"].shape), m.cost_regularization[0].engines())
