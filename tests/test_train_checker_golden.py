"""The torch checker of the GPU gradient tests, pinned to the unmodified reference (CPU).

tests/test_gpu_r2.py compares the CUDA training path (rcmvs.autograd: own convolution / BatchNorm / weight-gradient kernels)
with the same modules run on torch library ops + autograd (`FeatureNet._forward_torch`, `CostRegNet._forward_torch`).
This file closes the chain on the other side: those torch paths, in train() mode with the shipped checkpoint, reproduce
the outputs, every parameter gradient, the input gradient and the BatchNorm running statistics of the reference's own
FeatureNet / CostRegNet (models/modules.py:363-464, 470-501) as recorded by tests/golden/make_golden_r2.py train_grad.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rc-mvsnet_b200")]
G = os.path.join(ROOT, "tests", "golden")


def _model():
    import models
    from rcmvs import weights
    m = models.CascadeMVSNet(ndepths=[16, 8, 8])
    weights.load_model(m, os.path.join(G, "cas_model_000014.npz"))
    return m.train()


def _close(a, b, what, rel=2e-5):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = np.abs(a - b).max()
    assert err <= rel * (np.abs(b).max() + 1e-12), (what, err, np.abs(b).max())


def test_featurenet_torch_path_train_mode_matches_reference_gradients():
    g = np.load(os.path.join(G, "train_grad_small.npz"))
    torch.set_num_threads(4)
    net = _model().feature
    x = torch.from_numpy(g["feat_x"]).requires_grad_(True)
    f = net._forward_torch(x)
    assert sorted(f) == ["stage1", "stage2", "stage3"]
    for k in f:
        _close(f[k].detach().numpy(), g["feat_out_" + k], "feature " + k)
    sum((f[k] * torch.from_numpy(g["feat_cot_" + k])).sum() for k in f).backward()
    _close(x.grad.numpy(), g["feat_dx"], "d/d image")
    names = [n for n, _ in net.named_parameters()]
    assert sorted("feat_grad/" + n for n in names) == sorted(k for k in g.files if k.startswith("feat_grad/"))
    for n, p in net.named_parameters():
        _close(p.grad.numpy(), g["feat_grad/" + n], "grad " + n)
    bufs = dict(net.named_buffers())
    for k in (k for k in g.files if k.startswith("feat_buf/")):
        _close(bufs[k[len("feat_buf/"):]].numpy(), g[k], k, rel=1e-5)


def test_costregnet_torch_path_train_mode_matches_reference_gradients():
    g = np.load(os.path.join(G, "train_grad_small.npz"))
    torch.set_num_threads(4)
    reg = _model().cost_regularization[0]
    v = torch.from_numpy(g["reg_x"]).requires_grad_(True)
    y = reg._forward_torch(v)
    _close(y.detach().numpy(), g["reg_out"], "cost volume")
    (y * torch.from_numpy(g["reg_cot"])).sum().backward()
    _close(v.grad.numpy(), g["reg_dx"], "d/d variance volume")
    names = [n for n, _ in reg.named_parameters()]
    assert sorted("reg_grad/" + n for n in names) == sorted(k for k in g.files if k.startswith("reg_grad/"))
    for n, p in reg.named_parameters():
        _close(p.grad.numpy(), g["reg_grad/" + n], "grad " + n)
    bufs = dict(reg.named_buffers())
    for k in (k for k in g.files if k.startswith("reg_buf/")):
        _close(bufs[k[len("reg_buf/"):]].numpy(), g[k], k, rel=1e-5)
