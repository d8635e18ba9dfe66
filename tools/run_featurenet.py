"""Run FeatureNet (5 views, 512x640) a few times: target for `ncu -k regex:conv2d_direct` (development tool, GPU only)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "rc-mvsnet_b200"))
import torch
from models.modules import FeatureNet
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = FeatureNet(base_channels=8, num_stage=3, stride=4, arch_mode="fpn").to(dev).eval()
x = torch.randn(5, 3, 512, 640, device=dev)
with torch.no_grad():
    for _ in range(3):
        net.forward_cl(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        net.forward_cl(x)
    e1.record(); e1.synchronize()
print("featurenet 5x512x640: %.3f ms" % (e0.elapsed_time(e1) / 10))
