"""Depth error of the CUDA path against the reference-made full-size goldens, per stage and precision mode
(BASELINE configs[1]: 5 views 512x640 D=192/64/8; the test asserts < 1e-3, this prints the measured figures)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rc-mvsnet_b200"))
os.environ["RCMVS_STRICT"] = "1"
import models
from rcmvs import synth, weights
dev = torch.device("cuda:0")
G = os.path.join(ROOT, "tests", "golden")
g = np.load(os.path.join(G, "mvs_full_192_64_8.npz"))
imgs, proj, dv = synth.mvs_inputs(5, 512, 640, kind="textured", seed=4)
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
args = (cu(imgs), {k: cu(v) for k, v in proj.items()}, cu(dv))
for mode in ("bf16x3", "tf32x3", "fp32", "tf32"):
    os.environ["RCMVS_CONV"] = mode
    m = models.CascadeMVSNet_eval(ndepths=[192, 64, 8])
    weights.load_model(m, os.path.join(G, "cas_model_000014.npz"))
    m.to(dev).eval()
    with torch.no_grad():
        out = m(*args)
    rels = []
    for k in ("stage1", "stage2", "stage3"):
        got = out[k]["depth"][0].cpu().numpy()
        rels.append(np.abs(got - g["depth_" + k]).max() / np.abs(g["depth_" + k]).max())
    print("RCMVS_CONV=%-7s relative L-inf of depth vs the reference: stage1 %.2e  stage2 %.2e  stage3 %.2e" % (mode, *rels), flush=True)
