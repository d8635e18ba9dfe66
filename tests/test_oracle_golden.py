"""Pin the numpy oracle (oracle/ref_np.py) against golden vectors produced by the unmodified reference."""
import os

import numpy as np
import pytest

from oracle import ref_np as O

G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return dict(np.load(os.path.join(G, name)))


def test_homo_warp_matches_reference():
    g = load("warp_small.npz")
    for i, v in enumerate(range(1, g["feats"].shape[0])):
        out = O.homo_warp(g["feats"][v], g["projs"][v], g["projs"][0], g["depth_pp"])
        np.testing.assert_allclose(out, g["out_pp"][i], atol=2e-5, rtol=0)
        out = O.homo_warp(g["feats"][v], g["projs"][v], g["projs"][0], g["depth_1d"])
        np.testing.assert_allclose(out, g["out_1d"][i], atol=2e-5, rtol=0)


def test_homo_warp_out_of_bounds_and_behind_camera():
    g = load("warp_small.npz")
    out = O.homo_warp(g["feats"][1], g["projs_far"][1], g["projs_far"][0], g["depth_far"])
    assert np.isfinite(out).all()
    np.testing.assert_allclose(out, g["out_far"], atol=2e-5, rtol=0)
    assert (g["out_far"] == 0).mean() > 0.3     # the case really is out-of-bounds heavy


@pytest.fixture(scope="module")
def mvs():
    return load("mvs_small.npz"), load("cas_model_000014.npz")


def test_featurenet(mvs):
    g, P = mvs
    f = O.featurenet(g["imgs"][1], O.sub_params(P, "feature."))
    for k in ("stage1", "stage2", "stage3"):
        np.testing.assert_allclose(f[k], g["feat_%s_v1" % k], atol=2e-4, rtol=1e-4)


def test_hypotheses(mvs):
    g, _ = mvs
    H, W = g["imgs"].shape[-2:]
    interval = (float(g["depth_values"][-1]) - float(g["depth_values"][0])) / g["depth_values"].shape[0]
    h1 = O.depth_hypotheses_stage1(g["depth_values"], 16, H // 4, W // 4)
    np.testing.assert_allclose(h1, g["hyp_stage1"], atol=1e-4, rtol=0)
    h2 = O.depth_hypotheses_next(g["depth_stage1"], 8, 2 * interval, H, W, H // 2, W // 2)
    np.testing.assert_allclose(h2, g["hyp_stage2"], atol=2e-4, rtol=0)
    h3 = O.depth_hypotheses_next(g["depth_stage2"], 8, 1 * interval, H, W, H, W)
    np.testing.assert_allclose(h3, g["hyp_stage3"], atol=2e-4, rtol=0)


@pytest.mark.parametrize("stage", [1, 2, 3])
def test_variance_costreg_softargmin(mvs, stage):
    g, P = mvs
    k = "stage%d" % stage
    V = g["imgs"].shape[0]
    projs = O.stage_projection(g["proj_" + k])
    feats = [g["feat_%s_v%d" % (k, v)] for v in range(V)]
    var = O.costvolume_variance(feats, projs, g["hyp_" + k])
    np.testing.assert_allclose(var, g["variance_" + k], atol=3e-4, rtol=1e-4)   # E[x^2]-E[x]^2 cancellation at stage-3 magnitudes
    cost = O.costregnet(g["variance_" + k], O.sub_params(P, "cost_regularization.%d." % (stage - 1)))
    np.testing.assert_allclose(cost, g["cost_" + k], atol=5e-4, rtol=1e-4)
    depth, conf, idx, _ = O.softargmin_conf(g["cost_" + k], g["hyp_" + k])
    np.testing.assert_allclose(depth, g["depth_" + k], rtol=2e-6, atol=0)
    np.testing.assert_allclose(conf, g["conf_" + k], atol=2e-6, rtol=0)


def test_cascade_end_to_end(mvs):
    g, P = mvs
    proj = {k: g["proj_" + k] for k in ("stage1", "stage2", "stage3")}
    out = O.cascade_forward_eval(g["imgs"], proj, g["depth_values"], P, ndepths=(16, 8, 8))
    for k in ("stage1", "stage2", "stage3"):
        rel = np.abs(out[k]["depth"] - g["depth_" + k]).max() / np.abs(g["depth_" + k]).max()
        assert rel < 1e-5, (k, rel)
        np.testing.assert_allclose(out[k]["photometric_confidence"], g["conf_" + k], atol=1e-4)


def test_torch_port_matches_numpy_oracle_and_golden(mvs):
    """oracle/ref_torch.py (the multi-threaded port used as cpu_baseline) is pinned the same way."""
    from oracle import ref_torch as T
    g, P = mvs
    proj = {k: g["proj_" + k] for k in ("stage1", "stage2", "stage3")}
    for gs in (False, True):          # explicit gather restatement, and the F.grid_sample variant the baseline arms time
        out = T.cascade_forward_eval(g["imgs"], proj, g["depth_values"], P, ndepths=(16, 8, 8), grid_sample=gs)
        for k in ("stage1", "stage2", "stage3"):
            rel = np.abs(out[k]["depth"] - g["depth_" + k]).max() / np.abs(g["depth_" + k]).max()
            assert rel < 1e-5, (k, rel, gs)
            np.testing.assert_allclose(out[k]["photometric_confidence"], g["conf_" + k], atol=1e-4)
