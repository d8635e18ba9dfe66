"""CPU: the f2 oracle (oracle/ref_loss_torch.py) against value + gradient produced by the unmodified reference."""
import os

import numpy as np
import torch

from oracle import ref_loss_torch as L

G = os.path.join(os.path.dirname(__file__), "golden")


def test_loss_oracle_matches_reference_golden():
    g = dict(np.load(os.path.join(G, "unsup_loss_small.npz")))
    imgs = torch.from_numpy(g["imgs"][None])
    for s, key in enumerate(("stage1", "stage2", "stage3")):
        depth = torch.from_numpy(g["depth_" + key][None]).requires_grad_(True)
        total, rec, ss, sm = L.unsup_loss_stage(imgs, torch.from_numpy(g["proj_" + key][None]), depth, s)
        total.backward()
        np.testing.assert_allclose(total.item(), g["loss_" + key], rtol=2e-5)
        np.testing.assert_allclose(rec.item(), g["reconstr_" + key], rtol=2e-5)
        np.testing.assert_allclose(float(ss), g["ssim_" + key], rtol=2e-5)
        np.testing.assert_allclose(sm.item(), g["smooth_" + key], rtol=2e-5)
        want = g["grad_" + key]
        assert np.abs(depth.grad[0].numpy() - want).max() <= 1e-4 * np.abs(want).max() + 1e-9
