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


def test_random_image_mask_follows_the_reference_rng_order():
    """losses.aug_loss.random_image_mask (host side of the training step's second forward): under the same np.random seed the
    drop-in zeroes the very rectangles the reference's function does (x drawn before y, consecutive draws from one stream),
    and hands back (img, None) when the filter covers the image (reference losses/aug_loss.py:7-26)."""
    import torch
    from losses import random_image_mask
    g = np.load(os.path.join(G, "aug_mask_small.npz"))
    img = torch.from_numpy(g["img"])
    fs = tuple(int(v) for v in g["filter_size"])
    for seed in (0, 5, 11):
        np.random.seed(seed)
        a, m = random_image_mask(img, fs)
        _, m2 = random_image_mask(img, fs)
        assert np.array_equal(m.numpy().astype(np.uint8), g["mask_%d" % seed])
        assert np.array_equal(m2.numpy().astype(np.uint8), g["mask2_%d" % seed])
        assert np.array_equal(a.numpy(), g["masked_%d" % seed])
    np.random.seed(1)
    a, m = random_image_mask(img, (48, 60))
    assert bool(g["full_is_none"][0]) and m is None and a is img
