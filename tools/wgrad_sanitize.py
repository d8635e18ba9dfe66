"""Small weight-gradient / data-gradient / BatchNorm-reduction cases for compute-sanitizer (racecheck: the row ring of
wgrad_smem_kernel prefetches with cp.async into shared-memory slots while the current row is read; memcheck: halo and
tail handling).  Every case is also compared with the register-window kernel (rcmvs_wgrad_variant bit 3) or torch."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rc-mvsnet_b200"))
from rcmvs import _ffi, autograd as A
dev = torch.device("cuda:0")
flag = ctypes.c_int.in_dll(_ffi.lib(), "rcmvs_wgrad_variant")
torch.manual_seed(0)
cases = [  # (dim3d, k, stride, S shape, Bg shape): enough rows for several ring steps per CTA
    (True, 3, 1, (1, 8, 40, 24, 8), (1, 8, 40, 24, 8)),
    (True, 3, 1, (1, 6, 50, 20, 1), (1, 6, 50, 20, 8)),       # prob: Cs = 1
    (True, 3, 2, (1, 4, 40, 12, 16), (1, 8, 80, 24, 8)),
    (True, 3, 2, (2, 3, 30, 10, 32), (2, 6, 60, 20, 16)),
    (False, 3, 1, (2, 160, 36, 8), (2, 160, 36, 8)),
    (False, 5, 2, (1, 150, 20, 16), (1, 300, 40, 8)),
    (False, 1, 1, (1, 200, 24, 32), (1, 200, 24, 16)),
]
worst = 0.0
for dim3d, k, stride, ss, bs in cases:
    S, Bg = torch.randn(ss, device=dev), torch.randn(bs, device=dev)
    flag.value = 0
    G = A.conv_wgrad(S, Bg, k, stride, dim3d)
    flag.value = 8
    R = A.conv_wgrad(S, Bg, k, stride, dim3d)
    flag.value = 0
    err = float((G - R).abs().max() / R.abs().max())
    worst = max(worst, err)
    print("wgrad dim3d=%d k=%d s=%d S%s Bg%s  rel err vs register-window kernel %.2e" % (dim3d, k, stride, ss, bs, err), flush=True)
    assert err < 1e-4
# 2-D data gradient (register-blocked stride-1 kernel) against the one-pixel kernel
for cin, cout, k, hw in ((8, 8, 3, (19, 37)), (32, 16, 3, (12, 50)), (16, 16, 1, (9, 21))):
    dy = torch.randn(2, hw[0], hw[1], cout, device=dev)
    w = torch.randn(cout, cin, k, k, device=dev)
    a = A.conv2d_dgrad(dy, w, hw[0], hw[1], 1)
    flag.value = 32
    b = A.conv2d_dgrad(dy, w, hw[0], hw[1], 1)
    flag.value = 0
    err = float((a - b).abs().max() / b.abs().max())
    print("dgrad2d cin=%d cout=%d k=%d %s  rel err vs one-pixel kernel %.2e" % (cin, cout, k, hw, err), flush=True)
    assert err < 1e-5
# BatchNorm reductions (warp-shuffle + shared fp64 atomics)
for C, M in ((8, 5000), (32, 777), (64, 1300)):
    x = torch.randn(M, C, device=dev)
    mean, invstd, _ = A.bn_batch_stats(x, None, None, 0.1, 1e-5)
    err = float((mean - x.mean(0)).abs().max())
    print("bn_stats C=%d M=%d  mean err %.2e" % (C, M, err), flush=True)
    assert err < 1e-5
torch.cuda.synchronize()
print("OK worst wgrad rel err %.2e" % worst)
