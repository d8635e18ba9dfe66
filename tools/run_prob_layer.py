"""Launch the 8->1 plane-march kernel at the stage-1 BASELINE shape (development tool for ncu captures, GPU only)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "rc-mvsnet_b200"))
import torch
from rcmvs import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
D, H, W = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (192, 128, 160)
x = torch.randn(1, D, H, W, 8, device=dev)
w = ops.pack_conv3d_weight(torch.randn(1, 8, 3, 3, 3, device=dev) * 0.05, None, False)
for _ in range(3):
    ops.conv3d(x, w, None, None, stride=1, relu=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.conv3d(x, w, None, None, stride=1, relu=False)
e1.record(); e1.synchronize()
ms = e0.elapsed_time(e1) / 20
print("prob %s: %.4f ms, %.0f GB/s algorithmic" % ((D, H, W), ms, (x.numel() * 4 + D * H * W * 4) / ms / 1e6))
