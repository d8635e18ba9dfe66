"""CPU: the f3 oracle (oracle/ref_fusion_np.py) against the golden made by the reference's own functions."""
import os

import numpy as np

from oracle import ref_fusion_np as F

G = os.path.join(os.path.dirname(__file__), "golden")


def test_fusion_oracle_matches_reference_golden():
    g = dict(np.load(os.path.join(G, "geo_filter_small.npz")))
    NS = g["depths"].shape[0] - 1
    for v in range(1, NS + 1):
        m, dr, xs, ys = F.check_geometric_consistency(g["depths"][0], g["K"][0], g["ext"][0], g["depths"][v], g["K"][v],
                                                      g["ext"][v], 0.5, 0.01)
        np.testing.assert_array_equal(xs, g["x2d_src_%d" % v])
        np.testing.assert_array_equal(ys, g["y2d_src_%d" % v])
        np.testing.assert_array_equal(m, g["mask_%d" % v])                  # the boolean mask is bit-exact
        np.testing.assert_allclose(dr, g["depth_reproj_%d" % v], rtol=2e-7, atol=0)
        assert 0.2 < m.mean() < 0.99                                        # the case exercises both outcomes
    out = F.filter_view(g["depths"][0], g["confidence"], g["K"][0], g["ext"][0], g["depths"][1:], g["K"][1:], g["ext"][1:])
    np.testing.assert_array_equal(out["geo_mask_sum"], g["geo_mask_sum"])
    np.testing.assert_array_equal(out["final_mask"], g["final_mask"])
    np.testing.assert_allclose(out["depth_est_averaged"], g["depth_est_averaged"], rtol=2e-7)
