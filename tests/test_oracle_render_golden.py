"""Pin oracle/ref_render_np.py (a7-a13) against golden vectors produced by the unmodified reference."""
import os

import numpy as np
import pytest

from oracle import ref_np as O
from oracle import ref_render_np as R

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def g():
    return dict(np.load(os.path.join(G, "render_small.npz"))), dict(np.load(os.path.join(G, "nerf_model_000014.npz")))


def test_neural_volume_net(g):
    g, P = g
    vol = R.neural_volume_net(g["volume_feature_in"], O.sub_params(P, "MVSNet.cost_reg_2."))
    np.testing.assert_allclose(vol, g["volume"], atol=2e-3, rtol=1e-3)


def test_rays_points_ndc(g):
    g, _ = g
    H, W = g["imgs"].shape[-2:]
    ys, xs = g["pixel_coordinates"]
    o, d = R.rays_from_pixels(ys, xs, g["intrinsics"][0], g["c2ws"][0])
    np.testing.assert_allclose(d, g["rays_dir"], atol=1e-6)
    sel = g["sel"]
    pts = o[None, None] + g["depth_candidates"][sel][..., None] * d[sel][:, None]
    np.testing.assert_allclose(pts, g["rays_pts"], atol=2e-4)
    ndc = R.ndc_coordinates(pts, g["w2cs"][0], g["intrinsics"][0], W, H, g["near_fars"][0, 0], g["near_fars"][0, 1])
    np.testing.assert_allclose(ndc, g["rays_NDC"], atol=2e-5)
    np.testing.assert_array_equal(g["rays_depth"], g["pseudo_depth"][ys.astype(int), xs.astype(int)])
    # Gaussian half is sorted ascending, uniform half is stratified inside [near, far]
    dc = g["depth_candidates"]
    assert (np.diff(dc[:512], axis=1) >= 0).all()
    assert (dc[512:] >= g["near_fars"][0, 0]).all() and (dc[512:] <= g["near_fars"][0, 1]).all()


def test_point_features_mlp_composite(g):
    g, P = g
    imgs = g["imgs"]
    mean = np.array([-0.485 / 0.229, -0.456 / 0.224, -0.406 / 0.225], np.float32).reshape(1, 3, 1, 1)
    std = np.array([1 / 0.229, 1 / 0.224, 1 / 0.225], np.float32).reshape(1, 3, 1, 1)
    unp = (imgs - mean) / std                                            # render_consist_net.py:44-51
    feat8 = R.sample_volume_trilinear(g["volume"], g["rays_NDC"])
    col = R.color_features(g["rays_pts"], g["w2cs"], g["intrinsics"], unp[-3:])
    feat = np.concatenate([feat8, col], -1)
    np.testing.assert_allclose(feat, g["input_feat"], atol=3e-4)
    sel = g["sel"]
    dirs = np.repeat(g["angle"][sel][:, None], 128, 1)
    raw = R.render_mlp(g["rays_NDC"].reshape(-1, 3), g["input_feat"].reshape(-1, 20), dirs.reshape(-1, 3),
                       O.sub_params(P, "network_fn.nerf."))
    raw = raw.reshape(len(sel), 128, 4)
    # the shipped NeRF weights drive activations to ~1e6-1e7 on synthetic input (sigma up to 1.3e7): compare sigma
    # relatively, rgb (post-sigmoid of O(1e5) logits) with an absolute tolerance that reflects fp32 cancellation there
    np.testing.assert_allclose(raw[..., 3], g["raw"][..., 3], rtol=1e-3, atol=1e-2)
    np.testing.assert_allclose(raw[..., :3], g["raw"][..., :3], atol=5e-3)
    rgb, depth, w, alpha = R.raw2outputs(g["raw"], g["depth_candidates"][sel])
    np.testing.assert_allclose(rgb, g["rgb"][sel], atol=1e-5)
    np.testing.assert_allclose(depth, g["depth_pred"][sel], rtol=1e-5)
    np.testing.assert_allclose(w, g["weights"], atol=1e-6)
    np.testing.assert_allclose(alpha, g["alpha"], atol=1e-6)


def test_inverse_warping():
    g = dict(np.load(os.path.join(G, "inverse_warp_small.npz")))
    warped, mask = R.inverse_warping(g["img"][0], g["left_cam"][0], g["right_cam"][0], g["depth"][0])
    np.testing.assert_array_equal(mask, g["mask"][0, ..., 0])
    np.testing.assert_allclose(warped, g["warped"][0], atol=2e-4)
    assert 0.02 < (1 - mask).mean() < 0.9
