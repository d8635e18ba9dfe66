"""Diagnostic (GPU): where does the Tanks-shaped cascade differ from the reference golden?"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rc-mvsnet_b200"))
import models
from rcmvs import synth, weights
G = os.path.join(ROOT, "tests", "golden")
g = dict(np.load(os.path.join(G, "tanks_7x1056x1920_48_32_8.npz")))
dev = torch.device("cuda:0")
imgs, proj, dv = synth.mvs_inputs(7, 1056, 1920, kind="textured", seed=int(g["seed"]))
for mode in ("bf16x3", "fp32"):
    os.environ["RCMVS_CONV"] = mode
    m = models.CascadeMVSNet_eval(ndepths=[48, 32, 8])
    weights.load_model(m, os.path.join(G, "cas_model_000014.npz"))
    m.to(dev).eval()
    with torch.no_grad():
        out = m(torch.from_numpy(imgs).to(dev), {k: torch.from_numpy(v).to(dev) for k, v in proj.items()}, torch.from_numpy(dv).to(dev))
    for k in ("stage1", "stage2", "stage3"):
        got = out[k]["depth"][0].cpu().numpy(); ref = g["depth_" + k]
        err = np.abs(got - ref); mx = np.abs(ref).max()
        y, x = np.unravel_index(err.argmax(), err.shape)
        print(mode, k, "rel max %.3e at (%d,%d) of %s; p99.99 %.3e p99.9 %.3e; n>1e-4: %d n>5e-4: %d; got %.4f ref %.4f"
              % (err.max() / mx, y, x, err.shape, np.percentile(err, 99.99) / mx, np.percentile(err, 99.9) / mx,
                 (err / mx > 1e-4).sum(), (err / mx > 5e-4).sum(), got[y, x], ref[y, x]))
