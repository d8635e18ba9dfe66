"""GPU parity: every kernel is called through the C ABI (rcmvs.ops -> ctypes -> librcmvs.so) and compared with the
CPU oracle (oracle/ref_np.py) on seeded inputs, and with the golden vectors made by the unmodified reference.

Tolerances (floating point path; stated per test):
  * warp / variance: absolute, a few fp32 ulps of the feature magnitude (sample coordinates agree to ~1e-4 px);
  * CostRegNet cost volume: 1e-3 absolute on logits of magnitude O(10);
  * depth maps: 1e-3 relative L-inf (BASELINE.json north_star), integer depth index bit-exact away from fp ties.
"""
import os

import numpy as np
import pytest
import torch

from oracle import ref_np as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return dict(np.load(os.path.join(G, name)))


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    from rcmvs import _ffi
    assert _ffi.lib().rcmvs_check_device() == 0, _ffi.lib().rcmvs_last_error()
    return torch.device("cuda:0")


def cu(x, dev):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


# ---- layout / projections ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(2, 8, 5, 7), (1, 32, 3, 17, 33), (3, 3, 64, 64), (1, 1, 9, 9)])
def test_layout_roundtrip_bit_exact(dev, shape):
    from rcmvs import ops
    x = torch.randn(*shape, device=dev)
    cl = ops.to_channels_last(x)
    perm = (0,) + tuple(range(2, x.dim())) + (1,)
    assert torch.equal(cl, x.permute(*perm).contiguous())
    assert torch.equal(ops.to_planar(cl), x)


def test_relative_projections_match_oracle(dev):
    """rot|trans = top 3x4 of src_proj @ inverse(ref_proj) in the reference's own fp32 arithmetic (modules.py:315-317):
    bit-exact against torch-CPU (LAPACK LU + matmul) for the synthetic rigs, within a few ulp of the largest entry for
    general DTU-like cameras (MKL's LU orders a few operations differently)."""
    from rcmvs import ops, synth
    proj = synth.cameras(5, 512, 640, batch=2)["stage2"]
    rt = ops.relative_projections(cu(proj, dev)).cpu().numpy()
    for b in range(2):
        projs = torch.from_numpy(O.stage_projection(proj[b]))
        for v in range(1, 5):
            want = torch.matmul(projs[v], torch.inverse(projs[0]))[:3].numpy().ravel()
            np.testing.assert_array_equal(rt[b, v - 1], want)
            rot, trans = O.relative_projection(projs[v].numpy(), projs[0].numpy())
            np.testing.assert_allclose(rt[b, v - 1], np.concatenate([rot, trans[:, None]], 1).ravel(), rtol=3e-6,
                                       atol=3e-6 * np.abs(want).max())
    rng = np.random.default_rng(0)

    def rot3(a, bb, c):
        ca, sa, cb, sb, cc, sc = np.cos(a), np.sin(a), np.cos(bb), np.sin(bb), np.cos(c), np.sin(c)
        return (np.array([[1, 0, 0], [0, ca, -sa], [0, sa, ca]]) @ np.array([[cb, 0, sb], [0, 1, 0], [-sb, 0, cb]]) @
                np.array([[cc, -sc, 0], [sc, cc, 0], [0, 0, 1]]))
    P = np.zeros((64, 2, 2, 4, 4), np.float32)
    for i in range(64):
        for v in range(2):
            E = np.eye(4)
            E[:3, :3] = rot3(*rng.uniform(-0.6, 0.6, 3))
            E[:3, 3] = rng.uniform(-400, 400, 3) + np.array([0, 0, 600])
            P[i, v, 0] = E
            P[i, v, 1, :3, :3] = [[361.5 * rng.uniform(0.8, 4.5), 0, 80 * rng.uniform(0.8, 8)],
                                  [0, 361.5 * rng.uniform(0.8, 4.5), 64 * rng.uniform(0.8, 8)], [0, 0, 1]]
    rt = ops.relative_projections(cu(P, dev)).cpu().numpy()
    for i in range(64):
        projs = torch.from_numpy(O.stage_projection(P[i]))
        want = torch.matmul(projs[1], torch.inverse(projs[0]))[:3].numpy().ravel()
        # ulp-level differences of the inverse (MKL orders a few LU operations differently, and differently per host CPU)
        # are amplified by the cancellation in src @ inverse: bounded relative to the largest entry
        assert np.abs(rt[i, 0] - want).max() <= 1e-4 * np.abs(want).max()


# ---- a1 homo_warping ---------------------------------------------------------------------------------------------
def test_homo_warp_golden_reference_vectors(dev):
    from models.modules import homo_warping
    g = load("warp_small.npz")
    for i, v in enumerate(range(1, g["feats"].shape[0])):
        args = (cu(g["feats"][v:v + 1], dev), cu(g["projs"][v:v + 1], dev), cu(g["projs"][0:1], dev))
        out = homo_warping(*args, cu(g["depth_pp"][None], dev))[0].cpu().numpy()
        np.testing.assert_allclose(out, g["out_pp"][i], atol=2e-4, rtol=0)
        out = homo_warping(*args, cu(g["depth_1d"][None], dev))[0].cpu().numpy()
        np.testing.assert_allclose(out, g["out_1d"][i], atol=2e-4, rtol=0)
    out = homo_warping(cu(g["feats"][1:2], dev), cu(g["projs_far"][1:2], dev), cu(g["projs_far"][0:1], dev),
                       cu(g["depth_far"][None], dev))[0].cpu().numpy()
    assert np.isfinite(out).all()
    np.testing.assert_allclose(out, g["out_far"], atol=2e-4, rtol=0)


def test_homo_warp_baseline_config1_vs_oracle(dev):
    """BASELINE.json configs[0]: 3 src views, 128x160, D=48, C=32, random features/poses."""
    from models.modules import homo_warping
    from rcmvs import synth
    rng = np.random.default_rng(7)
    C, h, w, D, V = 32, 128, 160, 48, 4
    projs = synth.stage_projections(synth.cameras(V, 4 * h, 4 * w)["stage1"])[0]
    feats = rng.standard_normal((V, C, h, w), dtype=np.float32)
    base = np.linspace(425, 933.8, D).astype(np.float32)
    depth = (base[:, None, None] + rng.uniform(-3, 3, (D, h, w))).astype(np.float32)
    for v in range(1, V):
        got = homo_warping(cu(feats[v:v + 1], dev), cu(projs[v:v + 1], dev), cu(projs[0:1], dev),
                           cu(depth[None], dev))[0].cpu().numpy()
        want = O.homo_warp(feats[v], projs[v], projs[0], depth)
        err = np.abs(got - want)
        # coordinates differ by <~1e-4 px (fp64 vs fp32 inverse, fused vs unfused FMAs); N(0,1) features
        assert err.max() < 2e-3 and err.mean() < 2e-5, (err.max(), err.mean())


def test_homo_warp_backward_is_adjoint_of_forward(dev):
    """<warp(x), g> == <x, warp^T(g)> (linearity in src_fea) — size-independent property, fp64 accumulation."""
    from models.modules import homo_warping
    from rcmvs import synth
    rng = np.random.default_rng(3)
    C, h, w, D = 8, 64, 80, 12
    projs = synth.stage_projections(synth.cameras(2, 4 * h, 4 * w)["stage1"])[0]
    x = cu(rng.standard_normal((1, C, h, w), dtype=np.float32), dev).requires_grad_(True)
    depth = cu(np.linspace(450, 900, D).astype(np.float32)[None], dev)
    y = homo_warping(x, cu(projs[1:2], dev), cu(projs[0:1], dev), depth)
    g = torch.randn_like(y)
    y.backward(g)
    lhs = (y.detach().double() * g.double()).sum().item()
    rhs = (x.detach().double() * x.grad.double()).sum().item()
    assert abs(lhs - rhs) < 1e-4 * max(1.0, abs(lhs)), (lhs, rhs)
    # and against torch's own grid_sample backward on the same coordinates (library cross-check)
    rot, trans = O.relative_projection(projs[1], projs[0])
    ix, iy = O.warp_coords(rot, trans, np.broadcast_to(depth.cpu().numpy()[0][:, None, None], (D, h, w)), h, w)
    grid = torch.stack([cu(ix, dev) / ((w - 1) / 2) - 1, cu(iy, dev) / ((h - 1) / 2) - 1], -1).view(1, D * h, w, 2)
    x2 = x.detach().clone().requires_grad_(True)
    y2 = torch.nn.functional.grid_sample(x2, grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    y2.backward(g.view(1, C, D * h, w))
    assert (x.grad - x2.grad).abs().max().item() < 5e-3


# ---- a5 hypotheses -----------------------------------------------------------------------------------------------
def test_hypotheses_match_reference_golden(dev):
    from rcmvs import ops
    g = load("mvs_small.npz")
    H, W = g["imgs"].shape[-2:]
    interval = (float(g["depth_values"][-1]) - float(g["depth_values"][0])) / g["depth_values"].shape[0]
    p1 = ops.depth_hypotheses_first(cu(g["depth_values"][None], dev), 16).cpu().numpy()[0]
    np.testing.assert_array_equal(p1, g["hyp_stage1"][:, 0, 0])                       # bit-exact
    h2 = ops.depth_hypotheses_next(cu(g["depth_stage1"][None], dev), 8, 2 * interval, H, W, H // 2, W // 2)
    np.testing.assert_allclose(h2.cpu().numpy()[0], g["hyp_stage2"], atol=1.3e-4, rtol=0)   # <= 2 ulp @ 650
    h3 = ops.depth_hypotheses_next(cu(g["depth_stage2"][None], dev), 8, 1 * interval, H, W, H, W)
    np.testing.assert_allclose(h3.cpu().numpy()[0], g["hyp_stage3"], atol=1.3e-4, rtol=0)
    # interval derived on the device (what the cascade uses: no host read-back) == host-float path, bit for bit
    dv = cu(g["depth_values"][None], dev)
    h2d = ops.depth_hypotheses_next_dev(cu(g["depth_stage1"][None], dev), dv, 8, 2, H, W, H // 2, W // 2)
    h3d = ops.depth_hypotheses_next_dev(cu(g["depth_stage2"][None], dev), dv, 8, 1, H, W, H, W)
    assert torch.equal(h2d, h2) and torch.equal(h3d, h3)


def test_hypotheses_api_functions(dev):
    from models.modules import get_depth_range_samples
    g = load("mvs_small.npz")
    H, W = g["imgs"].shape[-2:]
    full = get_depth_range_samples(cu(g["depth_values"][None], dev), 16, 0.0, dev, torch.float32, [1, H, W])
    assert full.shape == (1, 16, H, W)
    np.testing.assert_array_equal(full[0, :, 3, 5].cpu().numpy(), g["hyp_stage1"][:, 0, 0])


# ---- a1+a2 fused sweep + variance --------------------------------------------------------------------------------
@pytest.mark.parametrize("stage", [1, 2, 3])
def test_costvolume_variance_golden(dev, stage):
    from rcmvs import ops
    g = load("mvs_small.npz")
    k = "stage%d" % stage
    V = g["imgs"].shape[0]
    feats = np.stack([g["feat_%s_v%d" % (k, v)] for v in range(V)])                   # [V,C,h,w]
    feats_cl = cu(feats.transpose(0, 2, 3, 1)[None], dev)
    rt = ops.relative_projections(cu(g["proj_" + k][None], dev))
    var = ops.costvolume_variance(feats_cl, rt, cu(g["hyp_" + k][None], dev))
    got = var[0].permute(3, 0, 1, 2).cpu().numpy()
    scale = max(1.0, float(np.abs(g["variance_" + k]).max()))
    np.testing.assert_allclose(got, g["variance_" + k], atol=1e-4 * scale, rtol=1e-4)
    if stage == 1:   # uniform planes passed as [B,D] must give the same volume
        var1 = ops.costvolume_variance(feats_cl, rt, cu(g["hyp_" + k][:, 0, 0][None], dev))
        assert torch.equal(var1, var)


def test_costvolume_variance_noref_and_oob_vs_oracle(dev):
    from rcmvs import ops, synth
    rng = np.random.default_rng(11)
    V, C, h, w, D = 4, 16, 40, 56, 10
    proj = synth.cameras(V, 4 * h, 4 * w, baseline=120.0)["stage1"]     # wide baseline: many out-of-bounds taps
    feats = rng.standard_normal((V, C, h, w), dtype=np.float32)
    depth = (np.linspace(300, 1200, D)[:, None, None] + rng.uniform(-5, 5, (D, h, w))).astype(np.float32)
    rt = ops.relative_projections(cu(proj, dev))
    var, noref = ops.costvolume_variance(cu(feats.transpose(0, 2, 3, 1)[None], dev), rt, cu(depth[None], dev), True)
    projs = O.stage_projection(proj[0])
    want, want_nr = O.costvolume_variance(list(feats), projs, depth, with_no_ref=True)
    np.testing.assert_allclose(var[0].permute(3, 0, 1, 2).cpu().numpy(), want, atol=3e-3, rtol=0)
    np.testing.assert_allclose(noref[0].permute(3, 0, 1, 2).cpu().numpy(), want_nr, atol=3e-3, rtol=0)


# ---- a3 conv3d / conv_transpose3d / CostRegNet ---------------------------------------------------------------------
@pytest.mark.parametrize("cin,cout,stride,dhw", [
    (8, 8, 1, (4, 9, 37)), (32, 8, 1, (3, 18, 34)), (16, 8, 1, (2, 5, 64)), (8, 16, 2, (6, 10, 44)),
    (16, 16, 1, (3, 8, 40)), (16, 32, 2, (4, 12, 20)), (32, 32, 1, (3, 6, 10)), (32, 64, 2, (6, 8, 12)),
    (64, 64, 1, (2, 4, 5)), (8, 1, 1, (5, 17, 33)), (8, 8, 2, (5, 7, 9)), (8, 1, 1, (1, 1, 1)), (8, 1, 1, (40, 33, 70)),
    (8, 16, 2, (1, 1, 1)), (8, 16, 2, (5, 7, 9)), (8, 16, 2, (9, 17, 70)), (8, 16, 2, (24, 33, 129)),
])
def test_conv3d_vs_oracle(dev, cin, cout, stride, dhw):
    from rcmvs import ops
    rng = np.random.default_rng(cin * 100 + cout + stride)
    D, H, W = dhw
    x = rng.standard_normal((cin, D, H, W), dtype=np.float32)
    w = (rng.standard_normal((cout, cin, 3, 3, 3)) / np.sqrt(27 * cin)).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    bias = rng.standard_normal(cout).astype(np.float32)
    want = O.conv3d(x, w * scale[:, None, None, None, None], stride=stride) + bias[:, None, None, None]
    want = np.maximum(want, 0)
    skip = rng.standard_normal(want.shape, dtype=np.float32)
    pw = ops.pack_conv3d_weight(cu(w, dev), cu(scale, dev))
    x_cl = cu(x.transpose(1, 2, 3, 0)[None], dev)
    got = ops.conv3d(x_cl, pw, cu(bias, dev), cu(skip.transpose(1, 2, 3, 0)[None], dev), stride=stride, relu=True)
    np.testing.assert_allclose(got[0].permute(3, 0, 1, 2).cpu().numpy(), want + skip, atol=2e-5, rtol=1e-5)
    got = ops.conv3d(x_cl, pw, None, None, stride=stride, relu=False)
    want_lin = O.conv3d(x, w * scale[:, None, None, None, None], stride=stride)
    np.testing.assert_allclose(got[0].permute(3, 0, 1, 2).cpu().numpy(), want_lin, atol=2e-5, rtol=1e-5)


@pytest.mark.parametrize("nsplit", [3, 1])
@pytest.mark.parametrize("cin,cout,dhw", [
    (8, 8, (3, 9, 37)), (16, 8, (2, 8, 30)), (32, 8, (4, 17, 61)), (16, 16, (9, 5, 64)), (8, 1, (5, 12, 33)),
    (8, 8, (1, 1, 1)), (32, 8, (20, 8, 31)), (8, 8, (23, 21, 95)), (16, 16, (13, 14, 31)),
])
def test_conv3d_tensor_core_vs_oracle(dev, cin, cout, dhw, nsplit):
    """tcgen05 implicit-GEMM path.  nsplit=3 (3xTF32) must be fp32-equivalent; nsplit=1 (plain TF32) carries
    ~2^-11 relative operand truncation, i.e. ~1e-3 of the accumulated magnitude."""
    from rcmvs import ops
    rng = np.random.default_rng(cin * 1000 + cout * 10 + nsplit)
    D, H, W = dhw
    x = rng.standard_normal((cin, D, H, W), dtype=np.float32)
    w = (rng.standard_normal((cout, cin, 3, 3, 3)) / np.sqrt(27 * cin)).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    bias = rng.standard_normal(cout).astype(np.float32)
    lin = O.conv3d(x, w * scale[:, None, None, None, None])
    want = np.maximum(lin + bias[:, None, None, None], 0)
    skip = rng.standard_normal(want.shape, dtype=np.float32)
    img = ops.conv3d_tc_pack(cu(w, dev), cu(scale, dev), nsplit)
    x_cl = cu(x.transpose(1, 2, 3, 0)[None], dev)
    got = ops.conv3d_tc(x_cl, img, cout, cu(bias, dev), cu(skip.transpose(1, 2, 3, 0)[None], dev), nsplit, True)
    tol = 3e-5 if nsplit == 3 else 6e-3
    np.testing.assert_allclose(got[0].permute(3, 0, 1, 2).cpu().numpy(), want + skip, atol=tol, rtol=tol)
    got = ops.conv3d_tc(x_cl, img, cout, None, None, nsplit, False)
    np.testing.assert_allclose(got[0].permute(3, 0, 1, 2).cpu().numpy(), lin, atol=tol, rtol=tol)


@pytest.mark.parametrize("cin,cout,dhw", [(64, 32, (2, 3, 5)), (32, 16, (3, 6, 33)), (16, 8, (4, 9, 35)),
                                          (16, 8, (1, 1, 1))])
def test_conv_transpose3d_vs_oracle(dev, cin, cout, dhw):
    from rcmvs import ops
    rng = np.random.default_rng(cin + cout)
    D, H, W = dhw
    x = rng.standard_normal((cin, D, H, W), dtype=np.float32)
    w = (rng.standard_normal((cin, cout, 3, 3, 3)) / np.sqrt(27 * cin)).astype(np.float32)
    bias = rng.standard_normal(cout).astype(np.float32)
    want = np.maximum(O.conv_transpose3d(x, w) + bias[:, None, None, None], 0)
    skip = rng.standard_normal(want.shape, dtype=np.float32)
    pw = ops.pack_conv3d_weight(cu(w, dev), None, transposed=True)
    got = ops.conv_transpose3d(cu(x.transpose(1, 2, 3, 0)[None], dev), pw, cu(bias, dev),
                               cu(skip.transpose(1, 2, 3, 0)[None], dev), relu=True)
    assert got.shape[1:4] == (2 * D, 2 * H, 2 * W)
    np.testing.assert_allclose(got[0].permute(3, 0, 1, 2).cpu().numpy(), want + skip, atol=2e-5, rtol=1e-5)


@pytest.mark.parametrize("nsplit", [3, 1])
@pytest.mark.parametrize("cin,cout,dhw", [(16, 8, (3, 9, 35)), (16, 8, (1, 1, 1)), (16, 8, (7, 17, 62)),
                                          (32, 16, (3, 6, 33)), (32, 16, (5, 12, 31))])
def test_conv_transpose3d_tensor_core_vs_oracle(dev, cin, cout, dhw, nsplit):
    from rcmvs import ops
    rng = np.random.default_rng(cin * 7 + cout + nsplit)
    D, H, W = dhw
    x = rng.standard_normal((cin, D, H, W), dtype=np.float32)
    w = (rng.standard_normal((cin, cout, 3, 3, 3)) / np.sqrt(27 * cin / 8)).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    bias = rng.standard_normal(cout).astype(np.float32)
    want = np.maximum(O.conv_transpose3d(x, w * scale[None, :, None, None, None]) + bias[:, None, None, None], 0)
    skip = rng.standard_normal(want.shape, dtype=np.float32)
    img = ops.conv_transpose3d_tc_pack(cu(w, dev), cu(scale, dev), nsplit)
    got = ops.conv_transpose3d_tc(cu(x.transpose(1, 2, 3, 0)[None], dev), img, cout, cu(bias, dev),
                                  cu(skip.transpose(1, 2, 3, 0)[None], dev), nsplit, True)
    assert got.shape[1:4] == (2 * D, 2 * H, 2 * W)
    tol = 3e-5 if nsplit == 3 else 6e-3
    np.testing.assert_allclose(got[0].permute(3, 0, 1, 2).cpu().numpy(), want + skip, atol=tol, rtol=tol)


@pytest.mark.parametrize("nsplit", [3, 1])
@pytest.mark.parametrize("cin,cout,mode,dhw", [
    (16, 16, 0, (5, 7, 33)), (16, 32, 1, (8, 10, 22)), (16, 32, 1, (7, 9, 21)), (32, 32, 0, (6, 8, 10)),
    (32, 64, 1, (6, 8, 10)), (64, 64, 0, (3, 4, 5)), (64, 64, 0, (1, 1, 1)), (64, 32, 2, (3, 4, 5)),
    (32, 16, 2, (3, 6, 11)), (32, 32, 0, (12, 16, 20)), (16, 16, 0, (1, 1, 129)), (8, 16, 1, (8, 10, 22)),
    (8, 16, 1, (1, 7, 9)), (8, 16, 0, (3, 5, 7)), (8, 16, 1, (16, 32, 40)),
])
def test_conv3d_gather_gemm_tensor_core_vs_oracle(dev, cin, cout, mode, dhw, nsplit):
    """Gather implicit-GEMM tcgen05 kernel of the 1/4- and 1/8-resolution levels: stride-1, stride-2 and transposed
    convolutions, ragged last M tile, batch 2."""
    from rcmvs import ops
    rng = np.random.default_rng(cin * 131 + cout * 7 + mode * 3 + nsplit)
    D, H, W = dhw
    xs = [rng.standard_normal((cin, D, H, W), dtype=np.float32) for _ in range(2)]
    scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    bias = rng.standard_normal(cout).astype(np.float32)
    if mode == 2:
        w = (rng.standard_normal((cin, cout, 3, 3, 3)) / np.sqrt(27 * cin / 8)).astype(np.float32)
        lin = [O.conv_transpose3d(x, w * scale[None, :, None, None, None]) for x in xs]
    else:
        w = (rng.standard_normal((cout, cin, 3, 3, 3)) / np.sqrt(27 * cin)).astype(np.float32)
        lin = [O.conv3d(x, w * scale[:, None, None, None, None], stride=1 + mode) for x in xs]
    lin = np.stack(lin)
    want = np.maximum(lin + bias[None, :, None, None, None], 0)
    skip = rng.standard_normal(want.shape, dtype=np.float32)
    img = ops.conv3d_gemm_tc_pack(cu(w, dev), cu(scale, dev), mode == 2, nsplit)
    x_cl = cu(np.stack(xs).transpose(0, 2, 3, 4, 1), dev)
    got = ops.conv3d_gemm_tc(x_cl, img, cout, mode, cu(bias, dev), cu(skip.transpose(0, 2, 3, 4, 1), dev), nsplit, True)
    assert tuple(got.shape) == (2,) + want.shape[2:] + (cout,)
    tol = 3e-5 if nsplit == 3 else 6e-3
    np.testing.assert_allclose(got.permute(0, 4, 1, 2, 3).cpu().numpy(), want + skip, atol=tol, rtol=tol)
    got = ops.conv3d_gemm_tc(x_cl, img, cout, mode, None, None, nsplit, False)
    np.testing.assert_allclose(got.permute(0, 4, 1, 2, 3).cpu().numpy(), lin, atol=tol, rtol=tol)
    if nsplit == 3 and cin != 8:
        # pre-split input (hi + lo == x exactly) and pre-split output: same numbers as the split-on-the-fly path
        import torch
        hi = (x_cl.view(torch.int32) & -8192).view(torch.float32)
        y, y_hi, y_lo = ops.conv3d_gemm_tc(hi, img, cout, mode, None, None, 3, False, x_lo=x_cl - hi, want_split=True)
        assert torch.equal(y, got)
        assert torch.equal(y_hi + y_lo, y) and int((y_hi.view(torch.int32) & 8191).abs().max()) == 0
        none_y, y_hi2, _ = ops.conv3d_gemm_tc(hi, img, cout, mode, None, None, 3, False, x_lo=x_cl - hi, want_y=False,
                                              want_split=True)
        assert none_y is None and torch.equal(y_hi2, y_hi)


def _model(dev, ndepths, cls_name="CascadeMVSNet_eval"):
    import models
    from rcmvs import weights
    m = getattr(models, cls_name)(ndepths=list(ndepths))
    weights.load_model(m, os.path.join(G, "cas_model_000014.npz"))
    return m.to(dev).eval()


@pytest.mark.parametrize("mode", ["bf16x3", "tf32x3", "fp32", "tf32"])
@pytest.mark.parametrize("stage", [1, 2, 3])
def test_costregnet_golden(dev, stage, mode, monkeypatch):
    monkeypatch.setenv("RCMVS_CONV", mode)
    g = load("mvs_small.npz")
    m = _model(dev, (16, 8, 8))
    k = "stage%d" % stage
    with torch.no_grad():
        cost = m.cost_regularization[stage - 1](cu(g["variance_" + k][None], dev))      # reference signature
    assert cost.shape[:2] == (1, 1)
    # logits of magnitude O(10); fp32 / 3xTF32: summation-order noise only; plain TF32: operand truncation
    tol = 5e-2 if mode == "tf32" else 1e-3
    np.testing.assert_allclose(cost[0, 0].cpu().numpy(), g["cost_" + k], atol=tol, rtol=1e-4 if mode != "tf32" else 1e-2)


# ---- f1 FeatureNet -----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cin,cout,k,stride,hw", [(8, 8, 3, 1, (9, 37)), (8, 16, 5, 2, (20, 70)), (16, 16, 3, 1, (7, 33)),
                                                  (16, 32, 5, 2, (18, 34)), (32, 32, 3, 1, (5, 40)), (32, 16, 3, 1, (6, 31)),
                                                  (32, 8, 3, 1, (12, 65)), (16, 32, 1, 1, (8, 36)), (8, 32, 1, 1, (4, 34))])
def test_conv2d_vs_oracle(dev, cin, cout, k, stride, hw):
    from rcmvs import ops
    rng = np.random.default_rng(cin * 100 + cout + k)
    H, W = hw
    x = rng.standard_normal((cin, H, W), dtype=np.float32)
    w = (rng.standard_normal((cout, cin, k, k)) / np.sqrt(k * k * cin)).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    bias = rng.standard_normal(cout).astype(np.float32)
    want = np.maximum(O.conv2d(x, w * scale[:, None, None, None], bias, stride=stride, pad=k // 2), 0)
    pw = ops.pack_conv2d_weight(cu(w, dev), cu(scale, dev))
    got = ops.conv2d(cu(x.transpose(1, 2, 0)[None], dev), pw, cu(bias, dev), None, stride=stride, relu=True)
    np.testing.assert_allclose(got[0].permute(2, 0, 1).cpu().numpy(), want, atol=2e-5, rtol=1e-5)
    if stride == 1 and H % 2 == 0 and W % 2 == 0:
        up = rng.standard_normal((cout, H // 2, W // 2), dtype=np.float32)
        want_up = O.conv2d(x, w, None, pad=k // 2) + np.repeat(np.repeat(up, 2, 1), 2, 2)
        got = ops.conv2d(cu(x.transpose(1, 2, 0)[None], dev), ops.pack_conv2d_weight(cu(w, dev)), None,
                         cu(up.transpose(1, 2, 0)[None], dev), relu=False)
        np.testing.assert_allclose(got[0].permute(2, 0, 1).cpu().numpy(), want_up, atol=2e-5, rtol=1e-5)


def test_featurenet_golden(dev):
    """Hand-written FeatureNet (conv2d_direct.cu) against the reference's per-view feature maps."""
    g = load("mvs_small.npz")
    m = _model(dev, (16, 8, 8))
    with torch.no_grad():
        f = m.feature(cu(g["imgs"], dev))                       # all views in one batch
    for k in ("stage1", "stage2", "stage3"):
        assert f[k].shape[1] == {"stage1": 32, "stage2": 16, "stage3": 8}[k]
        for v in range(g["imgs"].shape[0]):
            np.testing.assert_allclose(f[k][v].cpu().numpy(), g["feat_%s_v%d" % (k, v)], atol=3e-4, rtol=1e-4)


# ---- a4 softmax / soft-argmin / confidence -------------------------------------------------------------------------
@pytest.mark.parametrize("stage", [1, 2, 3])
def test_softargmin_conf_golden(dev, stage):
    from rcmvs import ops
    g = load("mvs_small.npz")
    k = "stage%d" % stage
    depth, conf, idx, prob = ops.softargmin_conf(cu(g["cost_" + k][None], dev), cu(g["hyp_" + k][None], dev),
                                                 want_idx=True, want_prob=True)
    np.testing.assert_allclose(depth[0].cpu().numpy(), g["depth_" + k], rtol=3e-6, atol=0)
    np.testing.assert_allclose(conf[0].cpu().numpy(), g["conf_" + k], atol=3e-6, rtol=0)
    _, _, want_idx, want_p = O.softargmin_conf(g["cost_" + k], g["hyp_" + k])
    np.testing.assert_allclose(prob[0].cpu().numpy(), want_p, atol=1e-6, rtol=0)
    # integer intermediate: bit-exact except where Σp·i lies within fp32 rounding of an integer (reported, must be rare)
    mism = (idx[0].cpu().numpy() != want_idx)
    assert mism.mean() < 1e-3, mism.mean()
    assert np.allclose(np.exp(np.log(prob[0].cpu().numpy().sum(0))), 1, atol=1e-5)


def test_softargmin_edge_planes(dev):
    """One-hot costs at the first/last plane: index clamps, zero padding of the 4-window (casmvsnet.py:117-122)."""
    from rcmvs import ops
    D, h, w = 8, 2, 3
    for hot, want_conf_taps in ((0, (0, 1, 2)), (D - 1, (D - 2, D - 1))):
        cost = np.zeros((1, D, h, w), np.float32)
        cost[:, hot] = 4.0
        planes = np.arange(D, dtype=np.float32)[None] * 2 + 10
        depth, conf, idx = ops.softargmin_conf(cu(cost, dev), cu(planes, dev), want_idx=True)
        d0, c0, i0, p0 = O.softargmin_conf(cost[0], np.broadcast_to(planes[0][:, None, None], (D, h, w)))
        np.testing.assert_allclose(depth[0].cpu().numpy(), d0, rtol=1e-6)
        np.testing.assert_allclose(conf[0].cpu().numpy(), c0, atol=1e-6)
        np.testing.assert_array_equal(idx[0].cpu().numpy(), i0)


# ---- a6 whole cascade ------------------------------------------------------------------------------------------------
def _run_model(m, g, dev):
    imgs = cu(g["imgs"][None], dev)
    proj = {k: cu(g["proj_" + k][None], dev) for k in ("stage1", "stage2", "stage3")}
    with torch.no_grad():
        return m(imgs, proj, cu(g["depth_values"][None], dev))


@pytest.mark.parametrize("mode", ["bf16x3", "tf32x3", "fp32"])
@pytest.mark.parametrize("fixture", ["mvs_small.npz", "mvs_small_noise.npz"])
def test_cascade_eval_matches_reference_golden(dev, fixture, mode, monkeypatch):
    monkeypatch.setenv("RCMVS_CONV", mode)
    g = load(fixture)
    out = _run_model(_model(dev, (16, 8, 8)), g, dev)
    assert set(out.keys()) >= {"depth", "photometric_confidence", "stage1", "stage2", "stage3"}
    for k in ("stage1", "stage2", "stage3"):
        d = out[k]["depth"][0].cpu().numpy()
        rel = np.abs(d - g["depth_" + k]).max() / np.abs(g["depth_" + k]).max()
        assert rel < 1e-3, (k, rel)                                   # north_star tolerance
        c = out[k]["photometric_confidence"][0].cpu().numpy()
        assert np.abs(c - g["conf_" + k]).mean() < 1e-3
    assert torch.equal(out["depth"], out["stage3"]["depth"])


def test_cascade_eval_plain_tf32_on_photo_consistent_input(dev, monkeypatch):
    """RCMVS_CONV=tf32 (single TF32 pass = the reference's own GPU precision under torch defaults) stays inside the
    1e-3 bar on photo-consistent input; on pure-noise input the soft-argmin is chaotic and only the fp32-equivalent
    modes are held to the bar (precision)."""
    monkeypatch.setenv("RCMVS_CONV", "tf32")
    g = load("mvs_small.npz")
    out = _run_model(_model(dev, (16, 8, 8)), g, dev)
    for k in ("stage1", "stage2", "stage3"):
        d = out[k]["depth"][0].cpu().numpy()
        assert np.abs(d - g["depth_" + k]).max() / np.abs(g["depth_" + k]).max() < 1e-3


def test_cascade_eval_midsize_vs_oracle(dev):
    """128x160 (BASELINE config-1 resolution), 4 views, cascade 16/8/8, photo-consistent input: CUDA path vs the
    numpy oracle end to end."""
    from rcmvs import synth
    imgs, proj, dv = synth.mvs_inputs(4, 128, 160, kind="textured", seed=2)
    P = load("cas_model_000014.npz")
    ref = O.cascade_forward_eval(imgs[0], {k: v[0] for k, v in proj.items()}, dv[0], P, ndepths=(16, 8, 8))
    m = _model(dev, (16, 8, 8))
    with torch.no_grad():
        out = m(cu(imgs, dev), {k: cu(v, dev) for k, v in proj.items()}, cu(dv, dev))
    for k in ("stage1", "stage2", "stage3"):
        d = out[k]["depth"][0].cpu().numpy()
        rel = np.abs(d - ref[k]["depth"]).max() / np.abs(ref[k]["depth"]).max()
        assert rel < 1e-3, (k, rel)


def test_cascade_full_size_vs_reference_golden(dev):
    """BASELINE configs[1]: 5 views 512x640, D=192/64/8, against depth maps produced by the unmodified reference at
    this exact size (tests/golden/make_golden.py), plus size-independent properties: finite outputs, confidence in
    [0,1], stage-1 depth inside the swept range, determinism, batch independence."""
    from rcmvs import synth
    g = load("mvs_full_192_64_8.npz")
    imgs, proj, dv = synth.mvs_inputs(5, 512, 640, kind="textured", seed=4)
    m = _model(dev, (192, 64, 8))
    args = (cu(imgs, dev), {k: cu(v, dev) for k, v in proj.items()}, cu(dv, dev))
    with torch.no_grad():
        out = m(*args)
        out2 = m(*args)
    d = out["depth"]
    assert d.shape == (1, 512, 640) and torch.isfinite(d).all()
    assert torch.equal(d, out2["depth"])
    for k in ("stage1", "stage2", "stage3"):
        got = out[k]["depth"][0].cpu().numpy()
        rel = np.abs(got - g["depth_" + k]).max() / np.abs(g["depth_" + k]).max()
        assert rel < 1e-3, (k, rel)                                   # north_star tolerance
    c = out["photometric_confidence"]
    assert (c >= 0).all() and (c <= 1 + 1e-5).all()
    assert np.abs(c[0].cpu().numpy() - g["conf_stage3"].astype(np.float32)).mean() < 2e-3
    s1 = out["stage1"]["depth"]
    assert (s1 >= dv.min() - 1e-3).all() and (s1 <= dv.max() + 1e-3).all()
    imgs2, proj2, dv2 = synth.mvs_inputs(5, 512, 640, kind="textured", seed=9)
    both = (cu(np.concatenate([imgs2, imgs]), dev),
            {k: cu(np.concatenate([proj2[k], proj[k]]), dev) for k in proj}, cu(np.concatenate([dv2, dv]), dev))
    with torch.no_grad():
        ob = m(*both)
    assert torch.allclose(ob["depth"][1], d[0], rtol=1e-5, atol=1e-3)


def test_train_class_volume_feature_no_ref(dev):
    g = load("mvs_train_small.npz")
    m = _model(dev, (16, 8, 8), "CascadeMVSNet")
    imgs = cu(g["imgs"][None], dev)
    proj = {k: cu(g["proj_" + k][None], dev) for k in ("stage1", "stage2", "stage3")}
    with torch.no_grad():
        out, vol = m(imgs, proj, cu(g["depth_values"][None], dev))
    assert vol.shape == (1, 3 * 3 + 32, 16, 16, 16)
    np.testing.assert_allclose(vol[0].cpu().numpy(), g["volume_feature_no_ref"], atol=2e-3, rtol=1e-3)
    rel = np.abs(out["stage3"]["depth"][0].cpu().numpy() - g["depth_stage3"]).max() / g["depth_stage3"].max()
    assert rel < 1e-3


def test_costregnet_rejects_shapes_not_divisible_by_8(dev):
    m = _model(dev, (16, 8, 8))
    with pytest.raises(ValueError, match="divisible by 8"):
        with torch.no_grad():
            m.cost_regularization[2](torch.zeros(1, 8, 12, 16, 16, device=dev))


def test_tanks_shaped_eval_properties(dev):
    """BASELINE configs[4] shape: 7 views, 1056x1920, cascade 48/32/8 — size-independent properties only (the CPU
    oracle needs minutes here): finite, confidence in [0,1], stage-1 depth inside the swept range, deterministic."""
    from rcmvs import synth
    imgs, proj, dv = synth.mvs_inputs(7, 1056, 1920, kind="noise", seed=1)
    m = _model(dev, (48, 32, 8))
    args = (cu(imgs, dev), {k: cu(v, dev) for k, v in proj.items()}, cu(dv, dev))
    with torch.no_grad():
        out = m(*args)
        out2 = m(*args)
    assert out["depth"].shape == (1, 1056, 1920) and torch.isfinite(out["depth"]).all()
    assert torch.equal(out["depth"], out2["depth"])
    c = out["photometric_confidence"]
    assert (c >= 0).all() and (c <= 1 + 1e-5).all()
    s1 = out["stage1"]["depth"]
    assert (s1 >= dv.min() - 1e-3).all() and (s1 <= dv.max() + 1e-3).all()


def test_host_pipeline_matches_direct_calls(dev):
    """rcmvs.pipeline.HostPipeline (copies on their own streams, two input slots) returns what a plain call returns,
    for every submit of a sequence that reuses the slots."""
    from rcmvs import synth
    from rcmvs.pipeline import HostPipeline
    m = _model(dev, (16, 8, 8))
    pipe = HostPipeline(m, dev)
    outs, want = [], []
    for seed in range(5):
        imgs, proj, dv = synth.mvs_inputs(3, 64, 64, seed=seed)
        ih = torch.from_numpy(imgs).pin_memory()
        ph = {k: torch.from_numpy(v).pin_memory() for k, v in proj.items()}
        dh = torch.from_numpy(dv).pin_memory()
        d_h = torch.empty((1, 64, 64)).pin_memory()
        c_h = torch.empty((1, 64, 64)).pin_memory()
        pipe.submit(ih, ph, dh, d_h, c_h)
        outs.append((d_h, c_h, ih, ph, dh))
    pipe.synchronize()
    with torch.no_grad():
        for d_h, c_h, ih, ph, dh in outs:
            o = m(ih.to(dev), {k: v.to(dev) for k, v in ph.items()}, dh.to(dev))
            assert torch.equal(o["depth"].cpu(), d_h) and torch.equal(o["photometric_confidence"].cpu(), c_h)
