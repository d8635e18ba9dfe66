"""Run the two headline kernels at BASELINE stage-1 shapes a few times (target for `ncu -k regex:...`):
fused plane sweep + variance (5 views, C=32, 128x160, D=192) and conv0 (32->8) on its output."""
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
with torch.no_grad():
    f = m.feature.forward_cl(torch.from_numpy(imgs[0]).to(dev))
    feats = f["stage1"].view(1, V, *f["stage1"].shape[1:])
    rt = ops.relative_projections(torch.from_numpy(proj["stage1"]).to(dev))
    hyp = ops.depth_hypotheses_first(torch.from_numpy(dv).to(dev), D)
    cr = m.cost_regularization[0]
    P = cr.packed()
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
        var = ops.costvolume_variance_split(feats, rt, hyp) if cr.wants_split_input() else ops.costvolume_variance(feats, rt, hyp)
        c0 = cr._run(P["conv0"], var)
    torch.cuda.synchronize()
print("ok", tuple(var.shape), tuple(c0.shape), cr.engines().get("conv0"))
