"""A/B of the weight-gradient kernel variants inside the real training step (csrc/train.cu rcmvs_wgrad_variant):
bit 0 = rotating register window (instead of sliding by register moves), bit 1 = one S channel per thread (instead of 4),
bit 2 = software prefetch, bit 3 = register-window kernel instead of the shared-memory kernel, bit 4 = shared-memory kernel without the row ring."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rc-mvsnet_b200"))
from rcmvs import _ffi, train
dev = torch.device("cuda:0")
flag = ctypes.c_int.in_dll(_ffi.lib(), "rcmvs_wgrad_variant")
for v, g in ((0, True), (32, True), (0, True)):
    flag.value = v
    r = train.bench_train_step(dev, steps=5, graph=g)
    print("wgrad variant %d %s: %.3f ms/step  loss %.6f" % (v, r["launch"][:10], r["ms_per_step"], r["loss"]), flush=True)
flag.value = 0
