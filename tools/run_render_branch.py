"""Run BASELINE configs[2] (CascadeMVSNet train class + Rendering_Consistency_Net forward) once or twice: target for ncu."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rc-mvsnet_b200"))
os.environ.setdefault("RCMVS_STRICT", "1")
import bench
r = bench.render_branch_timing(torch.device("cuda:0"), bench.peaks())
print({k: v for k, v in r.items() if k != "workload"})
