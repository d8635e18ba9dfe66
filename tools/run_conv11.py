"""Run conv11 (transposed 16 -> 8, to full resolution) of stage 1 on the slab tcgen05 kernel: ncu target (GPU only)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "rc-mvsnet_b200"))
from rcmvs import ops
nsplit = int(sys.argv[1]) if len(sys.argv) > 1 else 3
D, H, W = (96, 64, 80)
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.randn(1, D, H, W, 16, device=dev)
skip = torch.randn(1, 2 * D, 2 * H, 2 * W, 8, device=dev)
w = torch.randn(16, 8, 3, 3, 3, device=dev) * 0.05
img = ops.conv_transpose3d_tc_pack(w, None, nsplit)
for _ in range(3):
    y = ops.conv_transpose3d_tc(x, img, 8, None, skip, nsplit, True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    y = ops.conv_transpose3d_tc(x, img, 8, None, skip, nsplit, True)
e1.record(); torch.cuda.synchronize()
print("conv11 nsplit=%d: %.3f ms" % (nsplit, e0.elapsed_time(e1) / 10))
