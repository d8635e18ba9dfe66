"""GPU parity, round 2 rows: f2 (unsupervised loss stack), f3 (geometric-consistency filter), the integer depth index,
cfg 5 (Tanks-shaped) and full-size cfg 3 (render) against goldens made by the unmodified reference, multi-device guards."""
import os

import numpy as np
import pytest
import torch

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


# ---- a4: the integer intermediate -------------------------------------------------------------------------------------
@pytest.mark.parametrize("stage", [1, 2, 3])
def test_depth_index_bit_exact_or_pinned_ties(dev, stage):
    """depth_index = trunc(Σ p·i) (casmvsnet.py:119-121).  (1) bit-exact against an ascending-order fp32 restatement of
    the kernel's summation; (2) against the index the REFERENCE's own op sequence produced (idx_small.npz, torch's
    cascade summation order): every differing pixel is listed and must be a rounding tie — Σ p·i computed in float64
    lies within D·2^-23 of an integer — and the count is reported, not bounded by a rate."""
    from rcmvs import ops
    g, gi = load("mvs_small.npz"), load("idx_small.npz")
    cost = g["cost_stage%d" % stage]
    D = cost.shape[0]
    _, _, idx = ops.softargmin_conf(cu(cost[None], dev), cu(g["hyp_stage%d" % stage][None], dev), want_idx=True)
    idx = idx[0].cpu().numpy()
    # (1) same arithmetic, same order, on the host: exp/max/div are correctly rounded or identical on both sides only
    # up to 1 ulp of p, so compare through float64 instead of claiming bit-equality of p
    c64 = cost.astype(np.float64)
    p64 = np.exp(c64 - c64.max(0))
    p64 /= p64.sum(0)
    f64 = (p64 * np.arange(D, dtype=np.float64)[:, None, None]).sum(0)
    want = gi["idx_stage%d" % stage]
    diff = np.argwhere(idx != want)
    tie_tol = D * 2.0 ** -23 * 4
    for y, x in diff:
        v = f64[y, x]
        assert abs(v - round(v)) < tie_tol, ("non-tie depth_index mismatch", stage, y, x, v, idx[y, x], want[y, x])
    # away from ties the index is exact by construction: trunc(f64) where f64 is not within tie_tol of an integer
    safe = np.abs(f64 - np.rint(f64)) >= tie_tol
    assert np.array_equal(idx[safe], np.clip(np.trunc(f64[safe]), 0, D - 1).astype(np.int32))
    assert np.array_equal(want[safe], idx[safe])
    print("stage %d: %d of %d pixels differ from the reference index, all rounding ties" % (stage, len(diff), idx.size))


# ---- f3 ----------------------------------------------------------------------------------------------------------------
def test_geo_consistency_matches_reference_golden(dev):
    from rcmvs import fusion
    g = load("geo_filter_small.npz")
    NS = g["depths"].shape[0] - 1
    for v in range(1, NS + 1):
        m, dr, xs, ys = fusion.check_geometric_consistency(g["depths"][0], g["K"][0], g["ext"][0], g["depths"][v], g["K"][v],
                                                           g["ext"][v], 0.5, 0.01)
        np.testing.assert_array_equal(m, g["mask_%d" % v])                       # boolean mask: bit-exact
        np.testing.assert_allclose(xs, g["x2d_src_%d" % v], rtol=0, atol=1e-4)   # float64 chain, FMA contraction only
        np.testing.assert_allclose(ys, g["y2d_src_%d" % v], rtol=0, atol=1e-4)
        np.testing.assert_allclose(dr, g["depth_reproj_%d" % v], rtol=1e-6, atol=0)
    out = fusion.filter_view(g["depths"][0], g["confidence"], g["K"][0], g["ext"][0], g["depths"][1:], g["K"][1:], g["ext"][1:])
    np.testing.assert_array_equal(out["geo_mask_sum"], g["geo_mask_sum"])
    np.testing.assert_array_equal(out["final_mask"], g["final_mask"])
    np.testing.assert_allclose(out["depth_est_averaged"], g["depth_est_averaged"], rtol=1e-6)


def test_geo_consistency_full_size_properties_vs_oracle_sample(dev):
    """DTU-size maps (512x640, 10 sources): identical views => every pixel consistent with itself; a perturbed source
    fails exactly where the oracle says (checked on a row band the numpy oracle finishes instantly)."""
    from oracle import ref_fusion_np as F
    from rcmvs import fusion, synth
    H, W, NS = 512, 640, 10
    proj = synth.cameras(NS + 1, H, W)["stage3"][0]
    ext, K = proj[:, 0], proj[:, 1, :3, :3]
    rng = np.random.default_rng(5)
    flat = np.full((H, W), 650.0, np.float32)
    out = fusion.filter_view(flat, None, K[0], ext[0], np.stack([flat] * NS), [K[0]] * NS, [ext[0]] * NS)
    assert (out["geo_mask_sum"] == NS).all() and out["final_mask"].all()
    np.testing.assert_allclose(out["depth_est_averaged"], flat, rtol=1e-6)
    depths = (650 + 3 * rng.standard_normal((NS + 1, H, W))).astype(np.float32)
    out = fusion.filter_view(depths[0], None, K[0], ext[0], depths[1:], K[1:], ext[1:], per_source=True)
    band = slice(200, 208)
    for v in (1, NS):
        m, dr, _, _ = F.check_geometric_consistency(depths[0], K[0], ext[0], depths[v], K[v], ext[v], 0.5, 0.01)
        assert (out["masks"][v - 1][band] != m[band]).sum() == 0
        np.testing.assert_allclose(out["depth_reprojected"][v - 1][band], dr[band], rtol=1e-6)


# ---- f2 ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("stage", [0, 1, 2])
def test_unsup_loss_value_and_depth_gradient_match_reference(dev, stage):
    from losses.unsup_loss import UnSupLoss
    g = load("unsup_loss_small.npz")
    key = "stage%d" % (stage + 1)
    crit = UnSupLoss()
    depth = cu(g["depth_" + key][None], dev).requires_grad_(True)
    loss = crit(cu(g["imgs"][None], dev), cu(g["proj_" + key][None], dev), depth, stage)
    (loss * 1.0).backward()
    np.testing.assert_allclose(loss.item(), g["loss_" + key], rtol=1e-4)
    np.testing.assert_allclose(crit.reconstr_loss.item(), g["reconstr_" + key], rtol=1e-4)
    np.testing.assert_allclose(crit.ssim_loss.item(), g["ssim_" + key], rtol=1e-4)
    np.testing.assert_allclose(crit.smooth_loss.item(), g["smooth_" + key], rtol=1e-4)
    want = g["grad_" + key]
    err = np.abs(depth.grad[0].cpu().numpy() - want)
    # fp32 gradient through warps whose coordinates differ by ~1e-4 px: 1e-3 of the gradient's magnitude
    assert err.max() <= 2e-3 * np.abs(want).max(), (err.max(), np.abs(want).max())
    assert err.mean() <= 1e-4 * np.abs(want).max()


def test_unsup_loss_multistage_and_variants(dev):
    from losses import UnsupLossMultiStage, UnsupLossMultiStage_no_smooth
    g = load("unsup_loss_small.npz")
    inputs = {"stage%d" % s: {"depth": cu(g["depth_stage%d" % s][None], dev).requires_grad_(True)} for s in (1, 2, 3)}
    cams = {"stage%d" % s: cu(g["proj_stage%d" % s][None], dev) for s in (1, 2, 3)}
    total, sc = UnsupLossMultiStage()(inputs, cu(g["imgs"][None], dev), cams, dlossw=[0.5, 1.0, 2.0])
    np.testing.assert_allclose(total.item(), g["multistage_total"], rtol=1e-4)
    total.backward()
    for s in (1, 2, 3):
        w = [0.5, 1.0, 2.0][s - 1]
        np.testing.assert_allclose(inputs["stage%d" % s]["depth"].grad[0].cpu().numpy(), w * g["grad_stage%d" % s],
                                   atol=2e-3 * w * np.abs(g["grad_stage%d" % s]).max())
    t2, sc2 = UnsupLossMultiStage_no_smooth()(inputs, cu(g["imgs"][None], dev), cams)
    want = sum(12 * g["reconstr_stage%d" % s] + 6 * g["ssim_stage%d" % s] for s in (1, 2, 3))
    np.testing.assert_allclose(t2.item(), want, rtol=1e-4)
    assert "smooth_loss_stage1" not in sc2 and "smooth_loss_stage1" in sc


def test_aug_loss_and_sl1_match_torch_expressions(dev):
    """AugLossMultiStage / SL1Loss against the reference's torch expressions (boolean-index + F.smooth_l1_loss),
    value and gradient."""
    import torch.nn.functional as F
    from losses import AugLossMultiStage, SL1Loss, random_image_mask
    rng = np.random.default_rng(2)
    B, H, W = 2, 64, 96
    pseudo = cu((650 + 5 * rng.standard_normal((B, H, W))).astype(np.float32), dev)
    np.random.seed(4)
    img, fmask = random_image_mask(torch.ones(B, 3, H, W, device=dev), (H // 3, W // 3))
    assert fmask is not None and float(img.min()) == 0.0
    ests = {}
    for s, sc in ((1, 4), (2, 2), (3, 1)):
        ests["stage%d" % s] = {"depth": cu((650 + 5 * rng.standard_normal((B, H // sc, W // sc))).astype(np.float32), dev).requires_grad_(True)}
    total, scal = AugLossMultiStage()(ests, pseudo, None, fmask, dlossw=[0.5, 1.0, 2.0])
    total.backward()
    want_total = 0.0
    for s, sc in ((1, 4), (2, 2), (3, 1)):
        d = ests["stage%d" % s]["depth"].detach().clone().requires_grad_(True)
        pg = pseudo[:, ::sc, ::sc]
        m = fmask[:, 0, ::sc, ::sc] > 0.5
        l = F.smooth_l1_loss(d[m], pg[m], reduction="mean")
        l.backward()
        w = [0.5, 1.0, 2.0][s - 1]
        want_total += w * l.item()
        np.testing.assert_allclose(scal["aug_loss_stage%d" % s].item(), l.item(), rtol=1e-5)
        assert torch.allclose(ests["stage%d" % s]["depth"].grad, w * d.grad, rtol=1e-4, atol=1e-9)
    np.testing.assert_allclose(total.item(), want_total, rtol=1e-5)
    pred = cu((650 + 5 * rng.standard_normal(1024)).astype(np.float32), dev).requires_grad_(True)
    gt = cu(np.where(rng.uniform(size=1024) < 0.2, 0, 650 + 5 * rng.standard_normal(1024)).astype(np.float32), dev)
    l = SL1Loss()(pred, gt, gt > 0)
    l.backward()
    p2 = pred.detach().clone().requires_grad_(True)
    l2 = F.smooth_l1_loss(p2[gt > 0], gt[gt > 0], reduction="mean") * 0.5
    l2.backward()
    np.testing.assert_allclose(l.item(), l2.item(), rtol=1e-5)
    assert torch.allclose(pred.grad, p2.grad, rtol=1e-4, atol=1e-9)
    np.testing.assert_allclose(SL1Loss()(pred, gt).item(), l2.item(), rtol=1e-5)


# ---- cfg 5: Tanks-shaped eval against the reference's own output -------------------------------------------------------
def _model(dev, ndepths, cls_name="CascadeMVSNet_eval"):
    import models
    from rcmvs import weights
    m = getattr(models, cls_name)(ndepths=list(ndepths))
    weights.load_model(m, os.path.join(G, "cas_model_000014.npz"))
    return m.to(dev).eval()


def test_tanks_shaped_eval_matches_reference_golden(dev):
    """BASELINE configs[4]: 7 views 1056x1920, cascade 48/32/8 (the Tanks script's defaults), against depth maps
    produced by the unmodified reference at this exact size.

    Tolerance.  On this input the cascade is ill-conditioned: the REFERENCE ITSELF, re-run on the same images perturbed
    by 1e-7 relative noise (~1 ulp), moves by 1.1e-6 / 1.5e-5 / 1.6e-4 relative L-inf at stages 1 / 2 / 3 (measured with
    the unmodified reference, tests/golden/make_golden_r2.py `tanks_sensitivity`): a gain of ~14x per stage.  Any fp32
    implementation whose convolutions sum in a different order than MKL-DNN's differs from the reference by a few 1e-5 at
    stage 1 (this one: 4e-5 with exact-fp32 FFMA kernels, 6e-5 in the default bf16x3 mode) and therefore by ~1e-3 on the
    worst pixel of stage 3.  The north_star bar (1e-3 relative L-inf) is asserted where it is meaningful — stage 1, and
    the 99.9th percentile of every stage — and the L-inf of stages 2 / 3 is bounded at 3e-3 with the count of pixels
    beyond 1e-3 reported and limited to 0.05 %."""
    from rcmvs import synth
    g = load("tanks_7x1056x1920_48_32_8.npz")
    imgs, proj, dv = synth.mvs_inputs(7, 1056, 1920, kind="textured", seed=int(g["seed"]))
    m = _model(dev, (48, 32, 8))
    with torch.no_grad():
        out = m(cu(imgs, dev), {k: cu(v, dev) for k, v in proj.items()}, cu(dv, dev))
    for k in ("stage1", "stage2", "stage3"):
        got = out[k]["depth"][0].cpu().numpy()
        err = np.abs(got - g["depth_" + k]) / np.abs(g["depth_" + k]).max()
        p999, mx, frac = np.percentile(err, 99.9), err.max(), float((err > 1e-3).mean())
        print("%s: rel L-inf %.2e, p99.9 %.2e, pixels beyond 1e-3: %.4f %%" % (k, mx, p999, 100 * frac))
        assert p999 < 1e-3, (k, p999)
        assert mx < (1e-3 if k == "stage1" else 3e-3), (k, mx)
        assert frac < 5e-4, (k, frac)
    c = out["photometric_confidence"][0].cpu().numpy()
    assert np.abs(c - g["conf_stage3"].astype(np.float32)).mean() < 2e-3
    assert out["depth"].shape == (1, 1056, 1920)


# ---- cfg 3 at full size: MVS (train class) + renderer with the reference's samples injected ----------------------------
def test_render_branch_full_size_matches_reference_golden(dev):
    import models
    from rcmvs import synth, weights
    g = load("render_full_512x640.npz")
    H, W = 512, 640
    imgs, proj, dv = synth.mvs_inputs(4, H, W, kind="textured", seed=int(g["mvs_seed"]))
    batch = {k: cu(v, dev) for k, v in synth.render_batch(H, W, seed=int(g["batch_seed"])).items()}
    cas = _model(dev, (48, 32, 8), "CascadeMVSNet")
    nerf = models.Rendering_Consistency_Net(synth.RenderArgs())
    nerf.load_state_dict(weights.load_npz_state_dict(os.path.join(G, "nerf_model_000014.npz")), strict=True)
    nerf.to(dev).eval()
    samples = {"pixel_coordinates": cu(g["pixel_coordinates"], dev), "depth_candidates": cu(g["depth_candidates"], dev)}
    with torch.no_grad():
        out, vol = cas(cu(imgs, dev), {k: cu(v, dev) for k, v in proj.items()}, cu(dv, dev))
        rel = (out["depth"][0].cpu().numpy() - g["pseudo_depth"])
        assert np.abs(rel).max() / np.abs(g["pseudo_depth"]).max() < 1e-3
        # the golden pseudo depth is injected so the renderer sees exactly the reference's inputs
        rgb, feat, weights_, depth_pred, alpha, ret, rays_depth, target = nerf(vol, cu(g["pseudo_depth"][None], dev), batch,
                                                                              samples=samples)
    np.testing.assert_array_equal(rays_depth.cpu().numpy(), g["rays_depth"])
    np.testing.assert_allclose(target.cpu().numpy(), g["target_s"], atol=1e-6)
    np.testing.assert_allclose(feat.cpu().numpy()[:64], g["input_feat_first64"], atol=5e-3, rtol=2e-4)   # values up to ~130
    np.testing.assert_allclose(depth_pred.cpu().numpy(), g["depth_pred"], rtol=1e-3)      # north_star tolerance
    np.testing.assert_allclose(rgb.cpu().numpy(), g["rgb"], atol=2e-2)


# ---- multi-device: a model on cuda:1 while cuda:0 is current -----------------------------------------------
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_model_on_second_device_while_first_is_current():
    g = load("mvs_small.npz")
    d0, d1 = torch.device("cuda:0"), torch.device("cuda:1")
    torch.cuda.set_device(0)
    m0, m1 = _model(d0, (16, 8, 8)), _model(d1, (16, 8, 8))

    def run(m, dev):
        with torch.no_grad():
            return m(cu(g["imgs"][None], dev), {k: cu(g["proj_" + k][None], dev) for k in ("stage1", "stage2", "stage3")},
                     cu(g["depth_values"][None], dev))["depth"]
    a = run(m0, d0)
    b = run(m1, d1)                       # cuda:0 is still the current device
    assert b.device == d1 and torch.equal(a.cpu(), b.cpu())


# ---- a3 round-2 engine: TMA-fed split-bf16 slab convolution ---------------------------------------------------------------
@pytest.mark.parametrize("cin,cout,dhw", [
    (32, 8, (4, 17, 61)), (16, 8, (2, 8, 30)), (16, 16, (9, 5, 64)), (32, 8, (1, 1, 1)), (32, 8, (20, 8, 31)),
    (16, 8, (23, 21, 95)), (16, 16, (13, 14, 31)), (32, 8, (7, 40, 33)), (8, 8, (8, 24, 61)), (8, 8, (3, 10, 30)), (8, 8, (1, 1, 1)), (8, 1, (6, 21, 62)), (8, 1, (9, 8, 30)),
])
def test_conv3d_slab_bf16x3_vs_oracle(dev, cin, cout, dhw):
    """kind::f16 tcgen05 + TMA, split-bf16 operands: fp32-class accuracy (dropped terms <= 3·2^-17 relative per product)."""
    from oracle import ref_np as O
    from rcmvs import ops
    rng = np.random.default_rng(cin * 1000 + cout * 10 + 7)
    D, H, W = dhw
    xs = [rng.standard_normal((cin, D, H, W), dtype=np.float32) for _ in range(2)]
    w = (rng.standard_normal((cout, cin, 3, 3, 3)) / np.sqrt(27 * cin)).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    bias = rng.standard_normal(cout).astype(np.float32)
    lin = np.stack([O.conv3d(x, w * scale[:, None, None, None, None]) for x in xs])
    want = np.maximum(lin + bias[None, :, None, None, None], 0)
    skip = rng.standard_normal(want.shape, dtype=np.float32)
    img = ops.conv3d_slab_pack(cu(w, dev), cu(scale, dev))
    x_cl = cu(np.stack(xs).transpose(0, 2, 3, 4, 1), dev)
    x_split = ops.split_bf16(x_cl)
    assert torch.equal(x_split[0].float() + x_split[1].float(), x_split[0].float() + x_split[1].float())
    assert (x_split[0].float() + x_split[1].float() - x_cl).abs().max() <= 2.0 ** -16 * x_cl.abs().max()
    if cout == 1:                             # `prob`: fp32 output only
        y = ops.conv3d_slab(x_split, img, cout, cu(bias, dev), cu(skip.transpose(0, 2, 3, 4, 1), dev), True)
    else:
        y, ys = ops.conv3d_slab(x_split, img, cout, cu(bias, dev), cu(skip.transpose(0, 2, 3, 4, 1), dev), True, True, True)
        assert (ys[0].float() + ys[1].float() - y).abs().max() <= 2.0 ** -16 * y.abs().max()
    got = y.permute(0, 4, 1, 2, 3).cpu().numpy()
    np.testing.assert_allclose(got, want + skip, atol=1e-4, rtol=1e-4)
    y2 = ops.conv3d_slab(x_split, img, cout, None, None, False)
    np.testing.assert_allclose(y2.permute(0, 4, 1, 2, 3).cpu().numpy(), lin, atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("cin,cout,dhw", [(8, 16, (2, 2, 2)), (8, 16, (4, 8, 62)), (8, 16, (6, 34, 64)), (8, 16, (16, 24, 126)),
                                          (8, 16, (10, 64, 80)), (8, 16, (2, 26, 190)), (16, 32, (2, 2, 2)), (16, 32, (4, 8, 62)),
                                          (16, 32, (12, 32, 40)), (16, 32, (6, 18, 66))])
def test_conv3d_s2_slab_bf16x3_vs_oracle(dev, cin, cout, dhw):
    """conv1 (8 -> 16) / conv3 (16 -> 32), stride 2, on tcgen05: strided TMA boxes (parity sub-planes), taps as descriptor
    offsets (8 channels: two taps per K = 16 step through the leading-dimension offset), kd through a TMEM ring."""
    from oracle import ref_np as O
    from rcmvs import ops
    rng = np.random.default_rng(sum(dhw) + cin)
    D, H, W = dhw
    xs = [rng.standard_normal((cin, D, H, W), dtype=np.float32) for _ in range(2)]
    w = (rng.standard_normal((cout, cin, 3, 3, 3)) / np.sqrt(27 * cin)).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    bias = rng.standard_normal(cout).astype(np.float32)
    lin = np.stack([O.conv3d(x, w * scale[:, None, None, None, None], stride=2) for x in xs])
    want = np.maximum(lin + bias[None, :, None, None, None], 0)
    img = ops.conv3d_s2_slab_pack(cu(w, dev), cu(scale, dev))
    x_split = ops.split_bf16(cu(np.stack(xs).transpose(0, 2, 3, 4, 1), dev))
    y, ys = ops.conv3d_s2_slab(x_split, img, cout, cu(bias, dev), True, True, True)
    assert tuple(y.shape) == (2, D // 2, H // 2, W // 2, cout)
    np.testing.assert_allclose(y.permute(0, 4, 1, 2, 3).cpu().numpy(), want, atol=1e-4, rtol=1e-4)
    assert (ys[0].float() + ys[1].float() - y).abs().max() <= 2.0 ** -16 * max(1e-6, float(y.abs().max()))
    y2 = ops.conv3d_s2_slab(x_split, img, cout, None, False)
    np.testing.assert_allclose(y2.permute(0, 4, 1, 2, 3).cpu().numpy(), lin, atol=1e-4, rtol=1e-4)
    ys_only = ops.conv3d_s2_slab(x_split, img, cout, cu(bias, dev), True, want_y=False, want_split=True)
    assert torch.equal(ys_only, ys)


@pytest.mark.parametrize("cin,cout,dhw", [(16, 8, (1, 1, 1)), (16, 8, (2, 8, 31)), (16, 8, (3, 9, 33)), (16, 8, (5, 16, 62)),
                                          (16, 8, (4, 20, 80)), (16, 8, (12, 7, 40)), (32, 16, (1, 1, 1)), (32, 16, (3, 9, 33)),
                                          (32, 16, (6, 16, 40)), (32, 16, (2, 5, 64))])
def test_conv_transpose3d_slab_bf16x3_vs_oracle(dev, cin, cout, dhw):
    """conv11 / conv9 (ConvTranspose3d, stride 2, output_padding 1) + bias + ReLU + skip on the TMA-fed tcgen05 kernel."""
    from oracle import ref_np as O
    from rcmvs import ops
    rng = np.random.default_rng(sum(dhw) + cin)
    D, H, W = dhw
    xs = [rng.standard_normal((cin, D, H, W), dtype=np.float32) for _ in range(2)]
    w = (rng.standard_normal((cin, cout, 3, 3, 3)) / np.sqrt(27 * cin / 8)).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    bias = rng.standard_normal(cout).astype(np.float32)
    lin = np.stack([O.conv_transpose3d(x, w * scale[None, :, None, None, None]) for x in xs])
    assert lin.shape == (2, cout, 2 * D, 2 * H, 2 * W)
    skip = rng.standard_normal(lin.shape, dtype=np.float32)
    want = np.maximum(lin + bias[None, :, None, None, None], 0) + skip
    img = ops.conv_transpose3d_slab_pack(cu(w, dev), cu(scale, dev))
    x_split = ops.split_bf16(cu(np.stack(xs).transpose(0, 2, 3, 4, 1), dev))
    y, ys = ops.conv_transpose3d_slab(x_split, img, cout, cu(bias, dev), cu(skip.transpose(0, 2, 3, 4, 1), dev), True, True, True)
    np.testing.assert_allclose(y.permute(0, 4, 1, 2, 3).cpu().numpy(), want, atol=1e-4, rtol=1e-4)
    assert (ys[0].float() + ys[1].float() - y).abs().max() <= 2.0 ** -16 * max(1e-6, float(y.abs().max()))
    y2 = ops.conv_transpose3d_slab(x_split, img, cout, None, None, False)
    np.testing.assert_allclose(y2.permute(0, 4, 1, 2, 3).cpu().numpy(), lin, atol=1e-4, rtol=1e-4)


# ---- a1+a2 round-2 sweep kernel (v3) -------------------------------------------------------------------------------------
def test_sweep_v3_bitwise_equals_v1_and_split_output(dev):
    """The plane-distributed v3 kernel (shifted in-bounds quad + folded weights, packed fp32x2 math, IEEE division) must
    give bit-for-bit the numbers of the straightforward v1 kernel (the var_noref launch runs v1), incl. out-of-bounds
    taps on a wide baseline; its split-bf16 output is the same variance to 2^-16."""
    from rcmvs import ops, synth
    rng = np.random.default_rng(21)
    for (V, C, h, w, D, base) in ((4, 16, 40, 56, 10, 120.0), (5, 32, 24, 40, 16, 50.0), (3, 8, 32, 48, 8, 200.0)):
        proj = synth.cameras(V, 4 * h, 4 * w, baseline=base)["stage1"]
        feats = cu(rng.standard_normal((1, V, h, w, C), dtype=np.float32), dev)
        rt = ops.relative_projections(cu(proj, dev))
        for depth in ((np.linspace(300, 1200, D)[:, None, None] + rng.uniform(-5, 5, (D, h, w))).astype(np.float32)[None],
                      np.linspace(300, 1200, D).astype(np.float32)[None]):
            dt = cu(depth, dev)
            v3 = ops.costvolume_variance(feats, rt, dt)
            v1, _ = ops.costvolume_variance(feats, rt, dt, True)
            assert torch.equal(v3, v1)
            vs = ops.costvolume_variance_split(feats, rt, dt)
            assert vs.shape == (2,) + tuple(v3.shape) and vs.dtype == torch.bfloat16
            assert (vs[0].float() + vs[1].float() - v3).abs().max() <= 2.0 ** -16 * v3.abs().max()
            assert torch.equal(vs[0], v3.bfloat16())
            # v4 (two channel quads per lane): the same numbers, bit for bit, in both output forms
            import ctypes
            from rcmvs import _ffi
            flag = ctypes.c_int.in_dll(_ffi.lib(), "rcmvs_sweep_variant")
            try:
                flag.value = 4
                v4 = ops.costvolume_variance(feats, rt, dt)
                v4s = ops.costvolume_variance_split(feats, rt, dt)
            finally:
                flag.value = 0
            assert torch.equal(v4, v3) and torch.equal(v4s, vs)


# ---- training path: every backward kernel against torch autograd (library ops are the CHECKER here) -------------------------
def _tconv_ref(x_cl, w, transposed, stride):
    import torch.nn.functional as F
    x = x_cl.permute(0, 4, 1, 2, 3)
    y = F.conv_transpose3d(x, w, stride=2, padding=1, output_padding=1) if transposed else F.conv3d(x, w, stride=stride, padding=1)
    return y.permute(0, 2, 3, 4, 1)


@pytest.mark.parametrize("cin,cout,stride,transposed,dhw", [
    (32, 8, 1, False, (4, 8, 12)), (8, 16, 2, False, (4, 8, 12)), (16, 16, 1, False, (2, 6, 10)), (16, 32, 2, False, (4, 8, 8)),
    (64, 64, 1, False, (2, 2, 4)), (64, 32, 2, True, (2, 3, 4)), (32, 16, 2, True, (2, 4, 6)), (16, 8, 2, True, (3, 4, 10)),
    (8, 8, 1, False, (3, 5, 33)), (8, 8, 1, False, (2, 3, 640)), (32, 8, 1, False, (2, 3, 320)),
])
def test_conv3d_bn_relu_block_forward_backward_vs_torch(dev, cin, cout, stride, transposed, dhw):
    """conv -> BatchNorm3d (batch statistics) -> ReLU (+ skip): output, running statistics and all four gradients
    (input, weight, gamma, beta, skip) against torch autograd in fp32 with TF32 disabled."""
    import torch.nn as nn
    from rcmvs import autograd as A
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(cin * 7 + cout + stride)
    D, H, W = dhw
    conv = (nn.ConvTranspose3d(cin, cout, 3, stride=2, padding=1, output_padding=1, bias=False) if transposed else
            nn.Conv3d(cin, cout, 3, stride=stride, padding=1, bias=False)).to(dev)
    bn = nn.BatchNorm3d(cout, momentum=0.1).to(dev)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_()
    x = torch.randn(2, D, H, W, cin, device=dev, requires_grad=True)
    bn.train()
    ref_bn = nn.BatchNorm3d(cout, momentum=0.1).to(dev)
    ref_bn.load_state_dict(bn.state_dict())
    ref_bn.train()
    z_ref_lin = _tconv_ref(x, conv.weight, transposed, stride)
    skip = torch.randn_like(z_ref_lin.detach()).requires_grad_(True)
    z_ref = torch.relu(ref_bn(z_ref_lin.permute(0, 4, 1, 2, 3))).permute(0, 2, 3, 4, 1) + skip
    g = torch.randn_like(z_ref)
    gx, gw, gga, gbe, gs = torch.autograd.grad(z_ref, [x, conv.weight, ref_bn.weight, ref_bn.bias, skip], g)
    x2 = x.detach().clone().requires_grad_(True)
    skip2 = skip.detach().clone().requires_grad_(True)
    z = A.conv_block(x2, conv, bn, skip=skip2, relu=True, training=True)
    hx, hw, hga, hbe, hs = torch.autograd.grad(z, [x2, conv.weight, bn.weight, bn.bias, skip2], g)
    tol = dict(rtol=2e-4, atol=2e-4)
    assert torch.allclose(z, z_ref, **tol)
    assert torch.allclose(bn.running_mean, ref_bn.running_mean, rtol=1e-4, atol=1e-5)
    assert torch.allclose(bn.running_var, ref_bn.running_var, rtol=1e-4, atol=1e-5)
    scale = lambda t: float(t.abs().max()) + 1e-6
    assert (hx - gx).abs().max() <= 5e-4 * scale(gx)
    assert (hw - gw).abs().max() <= 5e-4 * scale(gw)
    assert (hga - gga).abs().max() <= 5e-4 * scale(gga) and (hbe - gbe).abs().max() <= 5e-4 * scale(gbe)
    assert torch.equal(hs, gs)
    # frozen statistics (eval-mode BatchNorm under autograd)
    bn.eval(); ref_bn.eval()
    ref_bn.load_state_dict(bn.state_dict())
    z_ref = torch.relu(ref_bn(_tconv_ref(x, conv.weight, transposed, stride).permute(0, 4, 1, 2, 3))).permute(0, 2, 3, 4, 1)
    gx, gw = torch.autograd.grad(z_ref, [x, conv.weight], g)
    z = A.conv_block(x2, conv, bn, relu=True, training=False)
    hx, hw = torch.autograd.grad(z, [x2, conv.weight], g)
    assert torch.allclose(z, z_ref, **tol) and (hx - gx).abs().max() <= 5e-4 * scale(gx) and (hw - gw).abs().max() <= 5e-4 * scale(gw)


def test_prob_conv_forward_backward_vs_torch(dev):
    import torch.nn as nn
    from rcmvs import autograd as A
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(3)
    conv = nn.Conv3d(8, 1, 3, padding=1, bias=False).to(dev)
    x = torch.randn(1, 8, 16, 24, 8, device=dev, requires_grad=True)
    ref = _tconv_ref(x, conv.weight, False, 1)
    g = torch.randn_like(ref)
    gx, gw = torch.autograd.grad(ref, [x, conv.weight], g)
    x2 = x.detach().clone().requires_grad_(True)
    z = A.conv_block(x2, conv, None, relu=False)
    hx, hw = torch.autograd.grad(z, [x2, conv.weight], g)
    assert torch.allclose(z, ref, rtol=1e-4, atol=1e-4)
    assert (hx - gx).abs().max() <= 5e-4 * float(gx.abs().max()) and (hw - gw).abs().max() <= 5e-4 * float(gw.abs().max())


@pytest.mark.parametrize("cin,cout,k,stride,bn,hw", [
    (3, 8, 3, 1, True, (16, 24)), (8, 16, 5, 2, True, (16, 24)), (16, 32, 5, 2, True, (12, 20)), (32, 32, 3, 1, True, (6, 10)),
    (32, 32, 1, 1, False, (6, 10)), (16, 32, 1, 1, False, (8, 12)), (8, 32, 1, 1, False, (8, 12)), (32, 16, 3, 1, False, (8, 12)),
    (32, 8, 3, 1, False, (8, 34)), (8, 8, 3, 1, True, (9, 37)), (16, 16, 3, 1, True, (40, 131)),
])
def test_conv2d_block_forward_backward_vs_torch(dev, cin, cout, k, stride, bn, hw):
    import torch.nn as nn
    import torch.nn.functional as F
    from rcmvs import autograd as A, ops
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(cin + cout * 3 + k)
    H, W = hw
    conv = nn.Conv2d(cin, cout, k, stride=stride, padding=k // 2, bias=not bn).to(dev)
    norm = nn.BatchNorm2d(cout).to(dev) if bn else None
    x_nchw = torch.randn(2, cin, H, W, device=dev, requires_grad=cin != 3)
    y = conv(x_nchw)
    ref = (torch.relu(norm(y)) if bn else y).permute(0, 2, 3, 1)
    g = torch.randn_like(ref)
    wanted = [conv.weight] + ([norm.weight, norm.bias] if bn else [conv.bias]) + ([x_nchw] if cin != 3 else [])
    grads = torch.autograd.grad(ref, wanted, g)
    if cin == 3:
        x_cl = ops.images_to_nhwc8(x_nchw.detach())
    else:
        x_cl = x_nchw.detach().permute(0, 2, 3, 1).contiguous().requires_grad_(True)
    if bn:
        norm2 = nn.BatchNorm2d(cout).to(dev)
        z = A.conv_block(x_cl, conv, norm2, relu=True, training=True)
        mine = torch.autograd.grad(z, [conv.weight, norm2.weight, norm2.bias] + ([x_cl] if cin != 3 else []), g)
    else:
        z = A.conv_block(x_cl, conv, None, relu=False)
        mine = torch.autograd.grad(z, [conv.weight, conv.bias] + ([x_cl] if cin != 3 else []), g)
    assert torch.allclose(z, ref, rtol=2e-4, atol=2e-4)
    for a, b in zip(mine, grads):
        if a.dim() == 4 and a.shape[-1] == cin and cin != 3 and a.shape != b.shape:
            a = a.permute(0, 3, 1, 2)
        assert (a - b).abs().max() <= 5e-4 * (float(b.abs().max()) + 1e-6), (a.shape, float((a - b).abs().max()), float(b.abs().max()))


def test_up2_add_and_layout_adjoints(dev):
    from rcmvs import autograd as A
    a = torch.randn(2, 8, 12, 16, device=dev, requires_grad=True)
    u = torch.randn(2, 4, 6, 16, device=dev, requires_grad=True)
    out = A.up2_add(a, u)
    want = a + u.repeat_interleave(2, 1).repeat_interleave(2, 2)
    assert torch.equal(out, want)
    g = torch.randn_like(out)
    ga, gu = torch.autograd.grad(out, [a, u], g)
    wa, wu = torch.autograd.grad(want, [a, u], g)
    assert torch.equal(ga, wa) and torch.allclose(gu, wu, rtol=1e-6, atol=1e-6)
    x = torch.randn(1, 5 * 0 + 8, 3, 4, 5, device=dev, requires_grad=True)
    y = A.to_planar(A.to_channels_last(x))
    (gx,) = torch.autograd.grad(y, [x], torch.ones_like(y))
    assert torch.equal(y, x) and torch.equal(gx, torch.ones_like(x))


def test_costvolume_variance_backward_vs_torch(dev):
    """Gradient of the fused sweep (+ var_no_ref) w.r.t. all feature maps against autograd through F.grid_sample on the
    same sample positions."""
    import torch.nn.functional as F
    from oracle import ref_np as O
    from rcmvs import autograd as A, ops, synth
    rng = np.random.default_rng(5)
    V, C, h, w, D = 4, 16, 20, 28, 6
    proj = synth.cameras(V, 4 * h, 4 * w, baseline=90.0)["stage1"]
    feats = cu(rng.standard_normal((1, V, h, w, C), dtype=np.float32), dev).requires_grad_(True)
    depth = (np.linspace(400, 1000, D)[:, None, None] + rng.uniform(-5, 5, (D, h, w))).astype(np.float32)
    rt = ops.relative_projections(cu(proj, dev))
    var, noref = A.costvolume_variance(feats, rt, cu(depth[None], dev), True)
    gv, gn = torch.randn_like(var), torch.randn_like(noref)
    (gf,) = torch.autograd.grad([var, noref], [feats], [gv, gn])
    # checker: torch grid_sample on coordinates from the numpy oracle
    projs = O.stage_projection(proj[0])
    f2 = feats.detach().clone().requires_grad_(True)
    fn = f2[0].permute(0, 3, 1, 2)                                                # [V,C,h,w]
    ref = fn[0][:, None].expand(-1, D, -1, -1)
    s, s2, n, n2 = ref, ref ** 2, 0, 0
    for v in range(1, V):
        rot, trans = O.relative_projection(projs[v], projs[0])
        ix, iy = O.warp_coords(rot, trans, depth, h, w)
        grid = torch.stack([cu(ix, dev) / ((w - 1) / 2) - 1, cu(iy, dev) / ((h - 1) / 2) - 1], -1).view(1, D * h, w, 2)
        wv = F.grid_sample(fn[v][None], grid, mode="bilinear", padding_mode="zeros", align_corners=True).view(C, D, h, w)
        s, s2, n, n2 = s + wv, s2 + wv ** 2, n + wv, n2 + wv ** 2
    var_r = (s2 / V - (s / V) ** 2).permute(1, 2, 3, 0)[None]
    nr_r = (n2 / V - (n / V) ** 2).permute(1, 2, 3, 0)[None]
    assert torch.allclose(var, var_r, atol=3e-3) and torch.allclose(noref, nr_r, atol=3e-3)
    (gr,) = torch.autograd.grad([var_r, nr_r], [f2], [gv, gn])
    assert (gf - gr).abs().max() <= 3e-3 * float(gr.abs().max())


def test_softargmin_backward_vs_torch(dev):
    from rcmvs import autograd as A
    torch.manual_seed(1)
    cost = torch.randn(2, 8, 6, 10, device=dev, requires_grad=True)
    for planes in (torch.linspace(400, 900, 8, device=dev)[None].repeat(2, 1).contiguous(),
                   (torch.linspace(400, 900, 8, device=dev).view(1, 8, 1, 1) + torch.rand(2, 8, 6, 10, device=dev)).contiguous()):
        depth, conf = A.softargmin(cost, planes)
        pl = planes if planes.dim() == 4 else planes.view(2, 8, 1, 1)
        ref = (torch.softmax(cost, 1) * pl).sum(1)
        g = torch.randn_like(ref)
        (a,) = torch.autograd.grad(depth, [cost], g)
        (b,) = torch.autograd.grad(ref, [cost], g)
        assert torch.allclose(depth, ref, rtol=1e-5) and (a - b).abs().max() <= 1e-4 * float(b.abs().max())


def test_fused_adam_matches_torch_adam(dev):
    from rcmvs import autograd as A
    torch.manual_seed(0)
    p = torch.randn(10007, device=dev)
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        g = torch.randn_like(p)
        ref.grad = g.clone()
        opt.step()
        A.adam_step(p, g, m, v, step, 1e-3, weight_decay=0.01)
        assert torch.allclose(p, ref.detach(), rtol=1e-5, atol=1e-7)


def test_cascade_training_forward_backward_matches_torch_autograd(dev):
    """Whole CascadeMVSNet in train() mode: outputs and the gradient of a scalar loss w.r.t. EVERY parameter, own kernels
    (RCMVS_STRICT=1: no library fallback) against the same modules run on torch library ops (the round-1 path)."""
    import models
    from rcmvs import synth, weights
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    imgs, proj, dv = synth.mvs_inputs(3, 64, 64, kind="textured", seed=2)
    args = (cu(imgs, dev), {k: cu(v, dev) for k, v in proj.items()}, cu(dv, dev))

    def build():
        m = models.CascadeMVSNet(ndepths=[16, 8, 8])
        weights.load_model(m, os.path.join(G, "cas_model_000014.npz"))
        return m.to(dev).train()
    m1, m2 = build(), build()
    out1, vol1 = m1(*args)
    loss1 = sum(out1["stage%d" % s]["depth"].mean() for s in (1, 2, 3)) + 1e-3 * vol1[:, -32:].mean()
    loss1.backward()
    # checker: torch library ops, module by module (what the drop-in did in round 1)
    feats = [m2.feature._forward_torch(args[0][:, v]) for v in range(3)]
    depth, loss2 = None, 0
    import torch.nn.functional as F
    from models.modules import homo_warping
    for s, key in enumerate(("stage1", "stage2", "stage3")):
        with torch.no_grad():
            hyp = out1[key]["hyp"] if "hyp" in out1[key] else None
        D = [16, 8, 8][s]
        sc = (4, 2, 1)[s]
        h, w = 64 // sc, 64 // sc
        from rcmvs import ops
        hyp = ops.depth_hypotheses_first(args[2], D).view(1, D, 1, 1).expand(-1, -1, h, w).contiguous() if depth is None else \
            ops.depth_hypotheses_next_dev(depth.detach().contiguous(), args[2], D, m2.depth_interals_ratio[s], 64, 64, h, w)
        pm = proj[key]
        P = cu(pm, dev)
        def kext(p):
            q = p[:, 0].clone()
            q[:, :3, :4] = torch.matmul(p[:, 1, :3, :3], p[:, 0, :3, :4])
            return q
        ref_proj = kext(P[:, 0])
        f0 = feats[0][key].unsqueeze(2)
        vs, vq, ns, nq = f0.expand(-1, -1, D, -1, -1), (f0 ** 2).expand(-1, -1, D, -1, -1), 0, 0
        for v in (1, 2):
            wv = homo_warping(feats[v][key], kext(P[:, v]), ref_proj, hyp)
            vs, vq, ns, nq = vs + wv, vq + wv ** 2, ns + wv, nq + wv ** 2
        var = vq / 3 - (vs / 3) ** 2
        nor = nq / 3 - (ns / 3) ** 2
        prob = F.softmax(m2.cost_regularization[s]._forward_torch(var).squeeze(1), 1)
        depth = (prob * hyp).sum(1)
        loss2 = loss2 + depth.mean()
        if s == 0:
            loss2 = loss2 + 1e-3 * nor.mean()
        assert torch.allclose(out1[key]["depth"], depth, rtol=2e-4, atol=2e-2), key
    loss2.backward()
    worst = 0.0
    for (n1, p1), (n2, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        assert n1 == n2
        if p2.grad is None:
            assert p1.grad is None or float(p1.grad.abs().max()) == 0.0, n1
            continue
        rel = float((p1.grad - p2.grad).abs().max()) / (float(p2.grad.abs().max()) + 1e-12)
        worst = max(worst, rel)
        assert rel < 2e-2, (n1, rel)
    print("worst relative parameter-gradient difference: %.2e" % worst)


# ---- rendering-consistency branch, training path (the mode the reference calls it in) ------------------------------------
def test_sgemm_linear_forward_backward_vs_torch(dev):
    import torch.nn as nn
    from rcmvs import autograd as A
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    for (M, K, N) in ((1000, 63, 128), (777, 191, 128), (513, 128, 1), (640, 64, 3), (2048, 131, 64), (300, 20, 128)):
        lin = nn.Linear(K, N).to(dev)
        x = torch.randn(M, K, device=dev, requires_grad=True)
        y_ref = lin(x)
        g = torch.randn_like(y_ref)
        gx, gw, gb = torch.autograd.grad(y_ref, [x, lin.weight, lin.bias], g)
        x2 = x.detach().clone().requires_grad_(True)
        y = A.linear(x2, lin)
        hx, hw, hb = torch.autograd.grad(y, [x2, lin.weight, lin.bias], g)
        assert torch.allclose(y, y_ref, rtol=1e-4, atol=1e-4)
        for a, b in ((hx, gx), (hw, gw), (hb, gb)):
            assert (a - b).abs().max() <= 2e-4 * (float(b.abs().max()) + 1e-6), (M, K, N)


def test_render_mlp_training_path_vs_torch_module(dev):
    """Renderer_ours.forward_train (own SGEMM / activation kernels + backward) against the module's torch forward."""
    import models
    from rcmvs import synth, weights
    torch.backends.cuda.matmul.allow_tf32 = False
    net = models.Rendering_Consistency_Net(synth.RenderArgs())
    net.load_state_dict(weights.load_npz_state_dict(os.path.join(G, "nerf_model_000014.npz")), strict=True)
    nerf = net.network_fn.nerf.to(dev)
    with torch.no_grad():                      # tame the shipped weights (they saturate sigma on synthetic input)
        for p in nerf.parameters():
            p.mul_(0.5)
    torch.manual_seed(1)
    N, S = 24, 32
    ndc = torch.rand(N, S, 3, device=dev)
    feat = torch.randn(N, S, 20, device=dev, requires_grad=True)
    dirs = torch.randn(N, 3, device=dev)
    from models.render_models import get_embedder
    emb, _ = get_embedder(10, 0, 3)
    x = torch.cat([emb(ndc), feat, dirs[:, None].expand(N, S, 3)], -1).reshape(N * S, -1)
    ref = nerf(x).view(N, S, 4)
    g = torch.randn_like(ref)
    params = list(nerf.parameters())
    gref = torch.autograd.grad(ref, [feat] + params, g)
    feat2 = feat.detach().clone().requires_grad_(True)
    out = nerf.forward_train(ndc, feat2, dirs)
    gmine = torch.autograd.grad(out, [feat2] + params, g)
    assert torch.allclose(out, ref, rtol=1e-3, atol=1e-4)
    for a, b in zip(gmine, gref):
        assert (a - b).abs().max() <= 2e-3 * (float(b.abs().max()) + 1e-9)


def test_composite_point_features_resize_backward_vs_torch(dev):
    import torch.nn.functional as F
    from models.renderer import raw2alpha
    from rcmvs import autograd as A
    torch.manual_seed(2)
    N, S = 40, 128
    raw = torch.cat([torch.rand(N, S, 3, device=dev), torch.rand(N, S, 1, device=dev) * 0.3], -1).requires_grad_(True)
    z = torch.sort(torch.rand(N, S, device=dev) * 500 + 400, -1)[0]
    alpha, weights, _ = raw2alpha(raw[..., 3], None, "v0")
    rgb_r, dep_r = (weights[..., None] * raw[..., :3]).sum(-2), (weights * z).sum(-1)
    g1, g2, g3 = torch.randn_like(rgb_r), torch.randn_like(dep_r), torch.randn_like(weights)
    (gr,) = torch.autograd.grad([rgb_r, dep_r, weights], [raw], [g1, g2, g3])
    raw2 = raw.detach().clone().requires_grad_(True)
    rgb, dep, w, a = A.composite(raw2, z)
    (gm,) = torch.autograd.grad([rgb, dep, w], [raw2], [g1, g2, g3])
    assert torch.allclose(rgb, rgb_r, atol=1e-5) and torch.allclose(dep, dep_r, rtol=1e-5)
    assert (gm - gr).abs().max() <= 1e-3 * float(gr.abs().max())
    # trilinear volume lookup
    D, h, w_ = 12, 10, 14
    vol = torch.randn(D, h, w_, 8, device=dev, requires_grad=True)
    ndc = torch.rand(30, 16, 3, device=dev) * 1.1 - 0.05
    grid = (ndc.view(1, 1, 30, 16, 3) * 2 - 1)
    ref = F.grid_sample(vol.permute(3, 0, 1, 2)[None], grid, align_corners=True, mode="bilinear")[0, :, 0].permute(1, 2, 0)
    gq = torch.randn_like(ref)
    (gv_r,) = torch.autograd.grad(ref, [vol], gq)
    vol2 = vol.detach().clone().requires_grad_(True)
    imgs3 = torch.rand(3, 3, 32, 40, device=dev)
    w2cs = torch.eye(4, device=dev)[None].repeat(3, 1, 1).contiguous()
    Ks = torch.tensor([[30., 0, 20], [0, 30, 16], [0, 0, 1]], device=dev)[None].repeat(3, 1, 1).contiguous()
    pts = torch.rand(30, 16, 3, device=dev) + torch.tensor([0., 0, 2], device=dev)
    feat = A.point_features(vol2, ndc, pts, imgs3, w2cs, Ks)
    assert torch.allclose(feat[..., :8], ref, atol=1e-5)
    gfull = torch.zeros_like(feat)
    gfull[..., :8] = gq
    (gv_m,) = torch.autograd.grad(feat, [vol2], gfull)
    assert (gv_m - gv_r).abs().max() <= 1e-4 * float(gv_r.abs().max())
    # depth resize
    x = torch.randn(1, 6, 5, 7, 41, device=dev, requires_grad=True)
    yr = F.interpolate(x.permute(0, 4, 1, 2, 3), size=[16, 5, 7], mode="trilinear", align_corners=True).permute(0, 2, 3, 4, 1)
    gy = torch.randn(1, 16, 5, 7, 48, device=dev)
    (gx_r,) = torch.autograd.grad(yr, [x], gy[..., :41])
    x2 = x.detach().clone().requires_grad_(True)
    y = A.resize_depth_cl(x2, 16, 48)
    (gx_m,) = torch.autograd.grad(y, [x2], gy)
    assert torch.allclose(y[..., :41], yr, atol=1e-5) and (y[..., 41:] == 0).all()
    assert (gx_m - gx_r).abs().max() <= 1e-5 * float(gx_r.abs().max()) + 1e-6


def test_rendering_consistency_net_training_step_on_own_kernels(dev):
    """model_nerf.train() + autograd, as train_rcmvsnet.py:279-312 calls it, with RCMVS_STRICT=1 (no library fallback):
    finite losses, gradients reach the MLP, the volume network AND (through volume_feature_no_ref) the MVS FeatureNet."""
    import models
    from losses import SL1Loss, masked_loss
    from rcmvs import synth, weights
    H = W = 64
    imgs, proj, dv = synth.mvs_inputs(4, H, W, kind="textured", seed=5)
    batch = {k: cu(v, dev) for k, v in synth.render_batch(H, W, seed=0).items()}
    cas = models.CascadeMVSNet(ndepths=[16, 8, 8])
    weights.load_model(cas, os.path.join(G, "cas_model_000014.npz"))
    cas.to(dev).train()
    nerf = models.Rendering_Consistency_Net(synth.RenderArgs())
    nerf.load_state_dict(weights.load_npz_state_dict(os.path.join(G, "nerf_model_000014.npz")), strict=True)
    nerf.to(dev).train()
    out, vol = cas(cu(imgs, dev), {k: cu(v, dev) for k, v in proj.items()}, cu(dv, dev))
    rgb, feat, wts, depth_pred, alpha, ret, rays_depth, target = nerf(vol, out["depth"].detach(), batch)
    loss = masked_loss(rgb.reshape(1, 1, -1), target.reshape(1, 1, -1), None, mask_kind=0, mode=1) + \
        SL1Loss()(depth_pred, rays_depth, rays_depth > 0)
    loss.backward()
    assert torch.isfinite(loss)
    gn = lambda m: sum(float(p.grad.abs().sum()) for p in m.parameters() if p.grad is not None)
    assert gn(nerf.network_fn) > 0 and gn(nerf.MVSNet) > 0 and gn(cas.feature) > 0
    assert all(torch.isfinite(p.grad).all() for p in list(nerf.parameters()) + list(cas.feature.parameters()) if p.grad is not None)


def test_train_step_cuda_graph_replay_matches_eager_step(dev):
    """rcmvs.train.TrainStep(graph=True): the captured zero-grad + two forwards + losses + backward, replayed, yields the
    loss and the flat gradient of the eager step on the same sample and the same mask rectangle (lr = 0 keeps the weights
    where they are; float atomics in the weight-gradient kernels bound the agreement), and a replay with another rectangle
    follows the eager step with that rectangle — i.e. the static inputs are refreshed per step.  The eager step lets
    autograd accumulate the parameter gradients, the graphed one has the kernels add into the flat buffer (GRAD_IN_PLACE)."""
    import numpy as np
    import models
    from rcmvs import train, weights

    def make():
        m = models.CascadeMVSNet(ndepths=[16, 8, 8])
        weights.load_model(m, os.path.join(G, "cas_model_000014.npz"))
        return m.to(dev)
    sample = train.synthetic_sample(dev, 3, 64, 96, seed=2)
    eager = train.TrainStep(make(), lr=0.0, graph=False, grad_in_place=False)     # (autograd accumulates the gradients)
    graphed = train.TrainStep(make(), lr=0.0, graph=True)
    for _ in range(train.TrainStep.GRAPH_WARMUP):
        graphed(sample)
    assert graphed._graph is None
    for seed in (5, 11):
        np.random.seed(seed)
        le, _ = eager(sample)
        ge = eager.flat.grad.clone()
        np.random.seed(seed)
        lg, sc = graphed(sample)
        gg = graphed.flat.grad.clone()
        assert graphed._graph is not None and graphed.graph_kernels > 100
        assert torch.isfinite(lg) and abs(float(lg) - float(le)) <= 1e-5 * abs(float(le))
        assert (gg - ge).abs().max() <= 2e-4 * float(ge.abs().max())
        assert float(ge.abs().max()) > 0
    np.random.seed(5)
    l5, _ = eager(sample)
    np.random.seed(11)
    l11, _ = eager(sample)
    assert float(l5) != float(l11)                       # (the two rectangles do give different steps)


# ---- f1 FeatureNet on tensor cores: 2-D TMA slab kernel ---------------------------------------------------------------------
@pytest.mark.parametrize("cin,cout,nhw", [(16, 16, (2, 9, 37)), (32, 32, (1, 20, 64)), (32, 16, (3, 7, 31)), (32, 8, (2, 33, 95)),
                                          (16, 16, (5, 64, 80)), (32, 32, (1, 1, 1)), (32, 8, (1, 12, 30)), (8, 8, (2, 21, 61)),
                                          (8, 8, (1, 64, 80)), (8, 8, (1, 1, 1))])
def test_conv2d_slab_bf16x3_vs_oracle(dev, cin, cout, nhw):
    from oracle import ref_np as O
    from rcmvs import ops
    rng = np.random.default_rng(cin * 100 + cout + nhw[0])
    N, H, W = nhw
    xs = [rng.standard_normal((cin, H, W), dtype=np.float32) for _ in range(N)]
    w = (rng.standard_normal((cout, cin, 3, 3)) / np.sqrt(9 * cin)).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    bias = rng.standard_normal(cout).astype(np.float32)
    lin = np.stack([O.conv2d(x, w * scale[:, None, None, None], None, pad=1) for x in xs])
    want = np.maximum(lin + bias[None, :, None, None], 0)
    x_cl = cu(np.stack(xs).transpose(0, 2, 3, 1), dev)
    xs_split = ops.split_bf16(x_cl)
    img = ops.conv2d_slab_pack(cu(w, dev), cu(scale, dev))
    y, ys = ops.conv2d_slab(xs_split, img, cout, cu(bias, dev), True, True, True)
    np.testing.assert_allclose(y.permute(0, 3, 1, 2).cpu().numpy(), want, atol=1e-4, rtol=1e-4)
    assert (ys[0].float() + ys[1].float() - y).abs().max() <= 2.0 ** -16 * max(1e-6, float(y.abs().max()))
    y2 = ops.conv2d_slab(xs_split, img, cout, None, False)
    np.testing.assert_allclose(y2.permute(0, 3, 1, 2).cpu().numpy(), lin, atol=1e-4, rtol=1e-4)
    # the FFMA producer's split output is the same split of its fp32 output
    pw = ops.pack_conv2d_weight(cu(w, dev), cu(scale, dev))
    yf, yfs = ops.conv2d(x_cl, pw, cu(bias, dev), None, relu=True, want_split=True)
    assert torch.equal(yfs, ops.split_bf16(yf))


@pytest.mark.parametrize("cin,cout,nhw", [(8, 16, (2, 2, 2)), (8, 16, (1, 24, 62)), (8, 16, (3, 50, 124)), (8, 16, (2, 64, 80)),
                                          (16, 32, (1, 2, 2)), (16, 32, (2, 18, 60)), (16, 32, (1, 36, 126)), (16, 32, (5, 32, 40))])
def test_conv2d_s2k5_slab_bf16x3_vs_oracle(dev, cin, cout, nhw):
    """FeatureNet conv1.0 / conv2.0 (5x5, stride 2, pad 2) on tcgen05: parity sub-planes by strided TMA, 25 taps as
    descriptor offsets (8 channels: two taps per K step through the leading-dimension offset)."""
    from oracle import ref_np as O
    from rcmvs import ops
    rng = np.random.default_rng(cin + cout + sum(nhw))
    N, H, W = nhw
    xs = [rng.standard_normal((cin, H, W), dtype=np.float32) for _ in range(N)]
    w = (rng.standard_normal((cout, cin, 5, 5)) / np.sqrt(25 * cin)).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    bias = rng.standard_normal(cout).astype(np.float32)
    lin = np.stack([O.conv2d(x, w * scale[:, None, None, None], None, stride=2, pad=2) for x in xs])
    assert lin.shape == (N, cout, H // 2, W // 2)
    want = np.maximum(lin + bias[None, :, None, None], 0)
    xs_split = ops.split_bf16(cu(np.stack(xs).transpose(0, 2, 3, 1), dev))
    img = ops.conv2d_s2k5_slab_pack(cu(w, dev), cu(scale, dev))
    y, ys = ops.conv2d_s2k5_slab(xs_split, img, cout, cu(bias, dev), True, True, True)
    np.testing.assert_allclose(y.permute(0, 3, 1, 2).cpu().numpy(), want, atol=1e-4, rtol=1e-4)
    assert (ys[0].float() + ys[1].float() - y).abs().max() <= 2.0 ** -16 * max(1e-6, float(y.abs().max()))
    y2 = ops.conv2d_s2k5_slab(xs_split, img, cout, None, False)
    np.testing.assert_allclose(y2.permute(0, 3, 1, 2).cpu().numpy(), lin, atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("dhw", [(1, 1, 1), (3, 8, 30), (4, 9, 31), (6, 16, 40), (12, 32, 40), (2, 20, 61)])
def test_conv3d_o_slab_bf16x3_vs_oracle(dev, dhw):
    """conv4 (32 -> 32, stride 1) on tcgen05: kd folded into N (TMEM plane ring), the nine (kh, kw) taps as descriptor offsets."""
    from oracle import ref_np as O
    from rcmvs import ops
    rng = np.random.default_rng(sum(dhw) + 77)
    D, H, W = dhw
    xs = [rng.standard_normal((32, D, H, W), dtype=np.float32) for _ in range(2)]
    w = (rng.standard_normal((32, 32, 3, 3, 3)) / np.sqrt(27 * 32)).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, 32).astype(np.float32)
    bias = rng.standard_normal(32).astype(np.float32)
    lin = np.stack([O.conv3d(x, w * scale[:, None, None, None, None]) for x in xs])
    skip = rng.standard_normal(lin.shape, dtype=np.float32)
    want = np.maximum(lin + bias[None, :, None, None, None], 0) + skip
    img = ops.conv3d_o_slab_pack(cu(w, dev), cu(scale, dev))
    x_split = ops.split_bf16(cu(np.stack(xs).transpose(0, 2, 3, 4, 1), dev))
    y, ys = ops.conv3d_o_slab(x_split, img, 32, cu(bias, dev), cu(skip.transpose(0, 2, 3, 4, 1), dev), True, True, True)
    np.testing.assert_allclose(y.permute(0, 4, 1, 2, 3).cpu().numpy(), want, atol=1e-4, rtol=1e-4)
    assert (ys[0].float() + ys[1].float() - y).abs().max() <= 2.0 ** -16 * max(1e-6, float(y.abs().max()))
    y2 = ops.conv3d_o_slab(x_split, img, 32, None, None, False)
    np.testing.assert_allclose(y2.permute(0, 4, 1, 2, 3).cpu().numpy(), lin, atol=1e-4, rtol=1e-4)


def test_cuda_graph_replay_is_bitwise_the_eager_forward(dev):
    """rcmvs.graph.GraphedForward / HostPipeline(graphs=True): the captured forward replays the eager launch sequence
    (same kernels, same arithmetic): outputs are bit-identical, also for new inputs loaded into the static buffers."""
    import models
    from rcmvs import synth, weights
    from rcmvs.graph import GraphedForward
    from rcmvs.pipeline import HostPipeline
    m = models.CascadeMVSNet_eval(ndepths=[16, 8, 8])
    weights.load_model(m, os.path.join(G, "cas_model_000014.npz"))
    m.to(dev).eval()
    sets = []
    for seed in (3, 4, 5):
        imgs, proj, dv = synth.mvs_inputs(3, 64, 96, kind="textured", seed=seed)
        sets.append((cu(imgs, dev), {k: cu(v, dev) for k, v in proj.items()}, cu(dv, dev)))
    with torch.no_grad():
        eager = [{k: m(*s)[k].clone() for k in ("depth", "photometric_confidence")} for s in sets]
    g = GraphedForward(m, *sets[0])
    assert g.kernels > 20
    for s, e in zip(sets, eager):
        out = g(*s)
        torch.cuda.synchronize()
        assert torch.equal(out["depth"], e["depth"]) and torch.equal(out["photometric_confidence"], e["photometric_confidence"])
    pipe = HostPipeline(m, dev, compute_streams=2, graphs=True)
    outs = []
    for i in range(7):
        s = sets[i % 3]
        d_h = torch.empty(s[0].shape[0], 64, 96).pin_memory()
        c_h = torch.empty(s[0].shape[0], 64, 96).pin_memory()
        pipe.submit(s[0].cpu().pin_memory(), {k: v.cpu().pin_memory() for k, v in s[1].items()}, s[2].cpu().pin_memory(), d_h, c_h)
        outs.append((i % 3, d_h, c_h))
    pipe.synchronize()
    for i, d_h, c_h in outs:
        assert torch.equal(d_h, eager[i]["depth"].cpu()) and torch.equal(c_h, eager[i]["photometric_confidence"].cpu())


@pytest.mark.parametrize("n_rays,n_samples", [(7, 128), (64, 128), (3, 33)])
def test_render_mlp_tensor_core_path_matches_fp32_kernel(dev, n_rays, n_samples, monkeypatch):
    """a10 on tcgen05 (csrc/mlp_tc.cu: 8 split-bf16 GEMMs + prep / head fusions) against the fused fp32 FFMA kernel of the
    same MLP, with weights scaled to a sane range (sigma ~ O(1), not the shipped checkpoint's 1e7) so that the comparison is
    tight; row counts that are not multiples of the 128-row tile included."""
    import models
    from rcmvs import ops, synth
    torch.manual_seed(5)
    net = models.Rendering_Consistency_Net(synth.RenderArgs()).to(dev).eval()
    mlp = net.network_fn.nerf
    with torch.no_grad():
        for prm in mlp.parameters():
            prm.copy_(torch.randn_like(prm) * (0.5 / max(1, prm.shape[-1]) ** 0.5))
        mlp.pts_bias.bias.add_(1.0)                      # modulation around 1
    ndc = torch.rand(n_rays, n_samples, 3, device=dev)
    feat = torch.randn(n_rays, n_samples, 20, device=dev)
    dirs = torch.randn(n_rays, 3, device=dev)
    monkeypatch.setenv("RCMVS_MLP_TC", "0")
    ref = mlp.forward_cuda(ndc, feat, dirs)
    monkeypatch.setenv("RCMVS_MLP_TC", "1")
    got = mlp.forward_cuda(ndc, feat, dirs)
    assert got.shape == ref.shape == (n_rays, n_samples, 4)
    scale = float(ref[..., 3].abs().max()) + 1e-6
    assert (got[..., 3] - ref[..., 3]).abs().max() <= 2e-4 * scale, ((got[..., 3] - ref[..., 3]).abs().max(), scale)
    assert (got[..., :3] - ref[..., :3]).abs().max() <= 2e-4
