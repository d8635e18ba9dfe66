"""Launch the gather-GEMM kernel on one layer shape (development tool for ncu captures, GPU only).
usage: python tools/run_gemm_layer.py Cin Cout mode D H W [presplit] [nsplit]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "rc-mvsnet_b200"))
import torch
from rcmvs import ops

cin, cout, mode, D, H, W = (int(a) for a in sys.argv[1:7])
pre = len(sys.argv) > 7 and sys.argv[7] == "1"
nsplit = int(sys.argv[8]) if len(sys.argv) > 8 else 3
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.randn(1, D, H, W, cin, device=dev)
w = torch.randn((cin, cout, 3, 3, 3) if mode == 2 else (cout, cin, 3, 3, 3), device=dev) * 0.05
img = ops.conv3d_gemm_tc_pack(w, None, mode == 2, nsplit)
hi = (x.view(torch.int32) & -8192).view(torch.float32)
lo = x - hi


def run():
    if pre:
        ops.conv3d_gemm_tc(hi, img, cout, mode, None, None, nsplit, True, x_lo=lo, want_split=True, want_y=False)
    else:
        ops.conv3d_gemm_tc(x, img, cout, mode, None, None, nsplit, True)


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    run()
e1.record(); e1.synchronize()
print("cin=%d cout=%d mode=%d %s pre=%d nsplit=%d: %.4f ms" % (cin, cout, mode, (D, H, W), pre, nsplit, e0.elapsed_time(e1) / 20))
