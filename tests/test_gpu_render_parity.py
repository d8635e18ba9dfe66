"""GPU parity of the rendering-consistency branch (a7-a12) and inverse_warping (a13) through the C ABI, against the
golden vectors made by the unmodified reference and the numpy oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_np as O
from oracle import ref_render_np as R

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return dict(np.load(os.path.join(G, name)))


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def cu(x, dev):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


@pytest.fixture(scope="module")
def net(dev):
    import models
    from rcmvs import synth, weights
    n = models.Rendering_Consistency_Net(synth.RenderArgs())
    n.load_state_dict(weights.load_npz_state_dict(os.path.join(G, "nerf_model_000014.npz")), strict=True)
    return n.to(dev).eval()


def test_neural_volume_net_golden(dev, net):
    g = load("render_small.npz")
    with torch.no_grad():
        vol = net.MVSNet(cu(g["volume_feature_in"][None], dev))                  # reference signature / layout
    assert vol.shape == (1, 8, 128, 16, 16)
    np.testing.assert_allclose(vol[0].cpu().numpy(), g["volume"], atol=3e-3, rtol=1e-3)


def test_resize_depth_align_corners(dev):
    from rcmvs import ops
    rng = np.random.default_rng(0)
    x = rng.standard_normal((5, 7, 3, 4), dtype=np.float32)                     # [C,D0,h,w]
    want = R.resize_depth_linear(x, 19)
    got = ops.resize_depth_cl(cu(x.transpose(1, 2, 3, 0)[None], dev), 19, 8)
    np.testing.assert_allclose(got[0, ..., :5].permute(3, 0, 1, 2).cpu().numpy(), want, atol=1e-6)
    assert (got[..., 5:] == 0).all()


def test_ray_points_and_features_golden(dev):
    from rcmvs import ops
    g = load("render_small.npz")
    H, W = g["imgs"].shape[-2:]
    ys, xs = g["pixel_coordinates"]
    sel = g["sel"]
    z = g["depth_candidates"][sel]
    pts, ndc, dirs = ops.ray_points(cu(ys[sel], dev), cu(xs[sel], dev), cu(g["intrinsics"][0], dev),
                                    cu(g["c2ws"][0], dev), cu(g["w2cs"][0], dev), cu(z, dev),
                                    float(g["near_fars"][0, 0]), float(g["near_fars"][0, 1]), W, H)
    np.testing.assert_allclose(dirs.cpu().numpy(), g["rays_dir"][sel], atol=1e-6)
    np.testing.assert_allclose(pts.cpu().numpy(), g["rays_pts"], atol=3e-4)
    np.testing.assert_allclose(ndc.cpu().numpy(), g["rays_NDC"], atol=3e-5)
    mean = np.array([-0.485 / 0.229, -0.456 / 0.224, -0.406 / 0.225], np.float32).reshape(1, 3, 1, 1)
    std = np.array([1 / 0.229, 1 / 0.224, 1 / 0.225], np.float32).reshape(1, 3, 1, 1)
    unp = (g["imgs"] - mean) / std
    vol_cl = cu(g["volume"].transpose(1, 2, 3, 0), dev)
    feat = ops.point_features(vol_cl, cu(g["rays_NDC"], dev), cu(g["rays_pts"], dev), cu(unp[-3:], dev),
                              cu(g["w2cs"][:3], dev), cu(g["intrinsics"][:3], dev))
    np.testing.assert_allclose(feat.cpu().numpy(), g["input_feat"], atol=5e-4)


def test_render_mlp_and_composite_golden(dev, net):
    from rcmvs import ops
    g = load("render_small.npz")
    sel = g["sel"]
    raw = net.network_fn.nerf.forward_cuda(cu(g["rays_NDC"], dev), cu(g["input_feat"], dev), cu(g["angle"][sel], dev))
    raw = raw.cpu().numpy()
    # sigma reaches ~1e7 with the shipped weights on synthetic input: relative check; rgb = sigmoid of huge logits
    np.testing.assert_allclose(raw[..., 3], g["raw"][..., 3], rtol=2e-3, atol=2e-2)
    np.testing.assert_allclose(raw[..., :3], g["raw"][..., :3], atol=1e-2)
    rgb, depth, w, alpha = ops.composite(cu(g["raw"], dev), cu(g["depth_candidates"][sel], dev))
    np.testing.assert_allclose(rgb.cpu().numpy(), g["rgb"][sel], atol=1e-5)
    np.testing.assert_allclose(depth.cpu().numpy(), g["depth_pred"][sel], rtol=1e-5)
    np.testing.assert_allclose(w.cpu().numpy(), g["weights"], atol=1e-6)
    np.testing.assert_allclose(alpha.cpu().numpy(), g["alpha"], atol=1e-6)


def test_rendering_consistency_net_end_to_end_golden(dev, net):
    """Whole forward with the reference's sampled pixels / depth candidates injected."""
    g = load("render_small.npz")
    batch = {k: cu(g[k][None], dev) for k in ("imgs", "depths_h", "w2cs", "c2ws", "near_fars", "intrinsics", "proj_mats")}
    samples = {"pixel_coordinates": cu(g["pixel_coordinates"], dev), "depth_candidates": cu(g["depth_candidates"], dev)}
    with torch.no_grad():
        rgb, feat, weights, depth_pred, alpha, ret, rays_depth, target = net(
            cu(g["volume_feature_in"][None], dev), cu(g["pseudo_depth"][None], dev), batch, samples=samples)
    sel = g["sel"]
    np.testing.assert_array_equal(rays_depth.cpu().numpy(), g["rays_depth"])
    np.testing.assert_allclose(target.cpu().numpy(), g["target_s"], atol=1e-6)
    np.testing.assert_allclose(feat.cpu().numpy()[sel], g["input_feat"], atol=3e-3)
    np.testing.assert_allclose(depth_pred.cpu().numpy(), g["depth_pred"], rtol=1e-3)      # north_star tolerance
    np.testing.assert_allclose(rgb.cpu().numpy(), g["rgb"], atol=2e-2)
    assert ret == {}


def test_gaussian_uniform_sampler_distribution(dev):
    """RNG parity with 1024 sequential torch.normal calls is impossible bit-for-bit: check the
    distributional contract instead — sorted Gaussian half with the reference's mean/std, stratified uniform half."""
    from rcmvs import ops
    N = 1024
    near, far = 425.0, 933.8
    d = torch.linspace(450, 900, N, device=dev)
    z = ops.gu_sample(d, near, far, 1234, N // 2).cpu().numpy()
    zg, zu = z[:N // 2], z[N // 2:]
    assert (np.diff(zg, axis=1) >= 0).all()
    dd = d.cpu().numpy()[:N // 2]
    sd = np.minimum(np.abs(far - dd), np.abs(dd - near)) / 3
    assert np.abs(((zg - dd[:, None]) / sd[:, None]).mean()) < 0.02
    assert abs(((zg - dd[:, None]) / sd[:, None]).std() - 1.0) < 0.02
    t = np.linspace(0, 1, 128, dtype=np.float32)
    zl = near * (1 - t) + far * t
    mids = 0.5 * (zl[1:] + zl[:-1])
    lo, hi = np.r_[zl[:1], mids], np.r_[mids, zl[-1:]]
    assert (zu >= lo[None] - 1e-3).all() and (zu <= hi[None] + 1e-3).all()
    u = (zu - lo[None])[:, 1:-1] / (hi - lo)[None, 1:-1]
    assert abs(u.mean() - 0.5) < 0.01
    z2 = ops.gu_sample(d, near, far, 1234, N // 2).cpu().numpy()
    assert np.array_equal(z, z2)                                      # deterministic for a fixed seed
    assert not np.array_equal(z, ops.gu_sample(d, near, far, 99, N // 2).cpu().numpy())


def test_inverse_warping_golden_and_depth_gradient(dev):
    from losses.homography import inverse_warping
    g = load("inverse_warp_small.npz")
    depth = cu(g["depth"], dev).requires_grad_(True)
    warped, mask = inverse_warping(cu(g["img"], dev), cu(g["left_cam"], dev), cu(g["right_cam"], dev), depth)
    np.testing.assert_array_equal(mask.cpu().numpy(), g["mask"])
    wn, m = warped.detach().cpu().numpy(), g["mask"][..., 0] > 0
    np.testing.assert_allclose(wn[m], g["warped"][m], atol=3e-4)
    # masked-out pixels project thousands of pixels outside the image: `x1 - x` cancels catastrophically in fp32 in
    # both implementations, the values are unused by the loss (mask = 0) — only a loose bound is meaningful there
    np.testing.assert_allclose(wn[~m], g["warped"][~m], atol=5e-3)
    (warped * cu(g["grad_out"], dev)).sum().backward()
    gd, want = depth.grad.cpu().numpy(), g["grad_depth"]
    assert np.abs(gd - want)[m].max() < 2e-3 * max(1.0, np.abs(want[m]).max())
    w2, m2 = R.inverse_warping(g["img"][0], g["left_cam"][0], g["right_cam"][0], g["depth"][0])
    np.testing.assert_allclose(wn[0][m[0]], w2[m[0]], atol=3e-4)
