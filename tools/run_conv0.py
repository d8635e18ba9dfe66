"""Run conv0 of stage 1 (Cin=32 -> 8, 192x128x160) alone a few times: target for `ncu -k regex:conv3d_umma`."""
import os, sys, torch
sys.path[:0] = [os.path.join(os.path.dirname(__file__), ".."), os.path.join(os.path.dirname(__file__), "..", "rc-mvsnet_b200")]
from rcmvs import ops
nsplit = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cin = int(sys.argv[2]) if len(sys.argv) > 2 else 32
D, H, W = (192, 128, 160)
dev = torch.device("cuda:0")
x = torch.randn(1, D, H, W, cin, device=dev)
w = torch.randn(8, cin, 3, 3, 3, device=dev) * 0.05
img = ops.conv3d_tc_pack(w, None, nsplit)
for _ in range(3):
    y = ops.conv3d_tc(x, img, 8, None, None, nsplit, True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    y = ops.conv3d_tc(x, img, 8, None, None, nsplit, True)
e1.record(); torch.cuda.synchronize()
print("conv0 cin=%d nsplit=%d: %.3f ms" % (cin, nsplit, e0.elapsed_time(e1) / 10))
