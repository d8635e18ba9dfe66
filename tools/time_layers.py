"""Per-layer timing of CostRegNet engines at the BASELINE stage shapes (development tool, GPU only).
usage: python tools/time_layers.py [stage ...]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "rc-mvsnet_b200"))
import torch
from models.modules import CostRegNet

SHAPES = {1: (32, 192, 128, 160), 2: (16, 64, 256, 320), 3: (8, 8, 512, 640)}     # BASELINE configs[1]


def main():
    stages = [int(a) for a in sys.argv[1:]] or [1, 2, 3]
    dev = torch.device("cuda:0")
    for s in stages:
        cin, D, H, W = SHAPES[s]
        torch.manual_seed(0)
        net = CostRegNet(cin, 8).to(dev).eval()
        P = net.packed()
        x = torch.randn(1, D, H, W, cin, device=dev)
        acts = {}
        with torch.no_grad():
            run = lambda n, a, sk=None: CostRegNet._call(P[n]["cands"][0], P[n], a, sk)
            acts["conv0"] = (x, None); c0 = run("conv0", x)
            acts["conv1"] = (c0, None); c1 = run("conv1", c0)
            acts["conv2"] = (c1, None); c2 = run("conv2", c1)
            acts["conv3"] = (c2, None); c3 = run("conv3", c2)
            acts["conv4"] = (c3, None); c4 = run("conv4", c3)
            acts["conv5"] = (c4, None); c5 = run("conv5", c4)
            acts["conv6"] = (c5, None); c6 = run("conv6", c5)
            acts["conv7"] = (c6, c4); y = run("conv7", c6, c4)
            acts["conv9"] = (y, c2); y = run("conv9", y, c2)
            acts["conv11"] = (y, c0); y = run("conv11", y, c0)
            acts["prob"] = (y, None)
            for name, (a, sk) in acts.items():
                line = "S%d %-6s in=%-22s" % (s, name, tuple(a.shape[1:]))
                ref = None
                for c in P[name]["cands"]:
                    out = CostRegNet._call(c, P[name], a, sk)
                    torch.cuda.synchronize()
                    if ref is None:
                        ref = out
                    err = (out - ref).abs().max().item() / max(ref.abs().max().item(), 1e-20)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(10):
                        CostRegNet._call(c, P[name], a, sk)
                    e1.record(); e1.synchronize()
                    line += "  %s %.3f ms (rel %.1e)" % (c[0], e0.elapsed_time(e1) / 10, err)
                print(line, flush=True)


if __name__ == "__main__":
    main()
