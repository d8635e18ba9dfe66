"""What bounds the TMA-fed split-bf16 slab kernel?  Times conv0 of stage 1 (32 -> 8 channels, 192x128x160) under the
experiment switches of csrc/conv3d_slab_bf16.cu (rcmvs_slab_debug_flags; results are WRONG under any non-zero flag):
  1 only the x_hi·w_hi term (1/3 of the MMAs)
  4 epilogue drains TMEM only (no exchange / store) 8 no TMA loads at all
 16 only kw = 0 (1/3 of the MMAs, all three terms)  32 N = one kd block (1/3 of the columns, same A fetches)
Usage: python tools/slab_experiments.py [Cin Cout D H W]"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rc-mvsnet_b200"))
from rcmvs import ops, _ffi
dev = torch.device("cuda:0")
a = [int(v) for v in sys.argv[1:]]
Cin, Cout, D, H, W = a if len(a) == 5 else (32, 8, 192, 128, 160)
flags = ctypes.c_int.in_dll(_ffi.lib(), "rcmvs_slab_debug_flags")
torch.manual_seed(0)
x = torch.randn(1, D, H, W, Cin, device=dev)
xs = ops.split_bf16(x)
w = torch.randn(Cout, Cin, 3, 3, 3, device=dev) * 0.05
img = ops.conv3d_slab_pack(w, None)
bias = torch.zeros(Cout, device=dev)
flush = torch.empty(64 << 20, device=dev, dtype=torch.float32)      # 256 MB > L2


def run(f, reps=10):
    flags.value = f
    for _ in range(3):
        ops.conv3d_slab(xs, img, Cout, bias, None, True, want_y=False, want_split=True)
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.conv3d_slab(xs, img, Cout, bias, None, True, want_y=False, want_split=True)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    flags.value = 0
    ts.sort()
    return ts[len(ts) // 2]


names = {0: "baseline", 1: "hh term only", 16: "kw=0 only", 32: "N = 1 block", 4: "epilogue drains only", 8: "no TMA loads",
         1 | 4: "hh only + drain only", 8 | 4: "no TMA + drain only", 1 | 8 | 4: "hh + no TMA + drain",
         1 | 16 | 8 | 4: "1/9 MMAs, no TMA, drain only", 32 | 8 | 4: "N=1blk, no TMA, drain", 1 | 16: "1/9 of the MMAs"}
print("shape Cin=%d Cout=%d D=%d H=%d W=%d" % (Cin, Cout, D, H, W))
for f, n in names.items():
    print("flags %2d  %-32s %.4f ms" % (f, n, run(f)), flush=True)


# ---- clock64 timeline of the first and the middle CTA (baseline flags) ------------------------------------------------------
trace = torch.zeros(1024, dtype=torch.int64, device=dev)
tp = ctypes.c_void_p.in_dll(_ffi.lib(), "rcmvs_slab_trace")
for f in (0, 1 | 8 | 4):
    tp.value = trace.data_ptr()
    flags.value = f
    trace.zero_()
    ops.conv3d_slab(xs, img, Cout, bias, None, True, want_y=False, want_split=True)
    torch.cuda.synchronize()
    tp.value = None
    flags.value = 0
    t = trace.cpu().numpy()
    for which in (0, 1):
        r = t[which * 512:(which + 1) * 512]
        n = int(r[3])
        t0 = r[0]
        print("flags %d CTA %s: prologue %d cycles, total %d cycles, %d slabs" % (f, "first" if which == 0 else "middle", r[1] - t0, r[2] - t0, n))
        print("  producer got stage   :", " ".join(str(int(v - t0)) for v in r[8:8 + n]))
        print("  issuer saw slab      :", " ".join(str(int(v - t0)) for v in r[72:72 + n]))
        print("  issuer issued slab   :", " ".join(str(int(v - t0)) for v in r[136:136 + n]))
        print("  epilogue saw plane   :", " ".join(str(int(v - t0)) for v in r[200:200 + n - 2]))
        print("  epilogue stored plane:", " ".join(str(int(v - t0)) for v in r[264:264 + n - 2]), flush=True)
