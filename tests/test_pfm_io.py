"""CPU: the native PFM reader / writer (f4 wire format) against files written by the reference's own save_pfm
(tests/golden/sample_*.pfm, made from the arrays in sample_pfm.npz by datasets/data_io.py:46-73): byte-identical output,
bit-identical input, the reference's error behaviour."""
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(__file__), "golden")


def test_pfm_roundtrip_matches_reference_bytes(tmp_path):
    from rcmvs import data_io
    g = np.load(os.path.join(G, "sample_pfm.npz"))
    d, s = data_io.read_pfm(os.path.join(G, "sample_depth.pfm"))
    assert s == 1.0 and d.dtype == np.float32 and np.array_equal(d, g["depth"])
    c, s = data_io.read_pfm(os.path.join(G, "sample_color.pfm"))
    assert s == 2.0 and np.array_equal(c, g["color"])
    p1, p2 = str(tmp_path / "d.pfm"), str(tmp_path / "c.pfm")
    data_io.save_pfm(p1, g["depth"])
    data_io.save_pfm(p2, g["color"], 2)
    assert open(p1, "rb").read() == open(os.path.join(G, "sample_depth.pfm"), "rb").read()
    assert open(p2, "rb").read() == open(os.path.join(G, "sample_color.pfm"), "rb").read()
    data_io.save_pfm(p1, g["depth"][..., None])                       # H x W x 1 is greyscale too
    assert np.array_equal(data_io.read_pfm(p1)[0], g["depth"])


def test_pfm_errors(tmp_path):
    from rcmvs import data_io
    with pytest.raises(Exception, match="float32"):
        data_io.save_pfm(str(tmp_path / "x.pfm"), np.zeros((4, 4), np.float64))
    with pytest.raises(Exception, match="dimensions"):
        data_io.save_pfm(str(tmp_path / "x.pfm"), np.zeros((4, 4, 2), np.float32))
    bad = tmp_path / "bad.pfm"
    bad.write_bytes(b"P6\n4 4\n255\n")
    with pytest.raises(Exception, match="Not a PFM"):
        data_io.read_pfm(str(bad))
