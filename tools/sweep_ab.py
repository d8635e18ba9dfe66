"""A/B of the plane-sweep kernels v3 / v4 at the three BASELINE stage shapes (CUDA events, L2 flushed between launches)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rc-mvsnet_b200"))
from rcmvs import ops, synth, _ffi
dev = torch.device("cuda:0")
flag = ctypes.c_int.in_dll(_ffi.lib(), "rcmvs_sweep_variant")
flush = torch.empty(64 << 20, device=dev)
V, H, W = 5, 512, 640
_, proj, dv = synth.mvs_inputs(V, H, W, kind="textured", seed=0)
for (C, sc, D, per_pixel) in ((32, 4, 192, False), (16, 2, 64, True), (8, 1, 8, True)):
    h, w = H // sc, W // sc
    feats = torch.randn(1, V, h, w, C, device=dev)
    rt = ops.relative_projections(torch.from_numpy(proj["stage%d" % {4: 1, 2: 2, 1: 3}[sc]]).to(dev))
    if per_pixel:
        hyp = (torch.linspace(500, 800, D, device=dev)[None, :, None, None] + torch.rand(1, D, h, w, device=dev)).contiguous()
    else:
        hyp = ops.depth_hypotheses_first(torch.from_numpy(dv).to(dev), D)
    for var in (3, 4):
        flag.value = var
        for _ in range(3):
            ops.costvolume_variance_split(feats, rt, hyp)
        ts = []
        for _ in range(9):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.costvolume_variance_split(feats, rt, hyp); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        flag.value = 0
        print("C=%d %dx%d D=%d per_pixel=%d  v%d: %.4f ms" % (C, h, w, D, per_pixel, var, sorted(ts)[4]), flush=True)
