"""CPU checks of the boundary: the shared library builds/loads and exports every symbol include/rcmvs.h declares;
argument validation works without a GPU (no compute calls)."""
import ctypes
import os

import pytest

import __graft_entry__ as ge


@pytest.fixture(scope="module")
def lib():
    from rcmvs import build, _ffi
    build.build()
    return _ffi.lib()


def test_exports_every_declared_symbol(lib):
    from rcmvs import _ffi
    decl = _ffi.declared_functions()
    assert len(decl) >= 15
    for name in decl:
        assert hasattr(lib, name), name
    assert lib.rcmvs_version() >= 100


def test_header_signatures_parse():
    from rcmvs import _ffi
    decl = _ffi.declared_functions()
    assert decl["rcmvs_costvolume_var_fwd"][1][-1] is ctypes.c_void_p          # stream last
    assert decl["rcmvs_depth_hypotheses_next"][1][-2] is ctypes.c_double


def test_bad_arguments_are_rejected_with_message(lib):
    from rcmvs import _ffi
    with pytest.raises(RuntimeError, match="null pointer"):
        _ffi.call("rcmvs_costvolume_var_fwd", None, None, None, 0, None, None, 1, 5, 32, 48, 128, 160, None)
    with pytest.raises(RuntimeError, match="Cin"):
        _ffi.call("rcmvs_conv3d_fwd", 16, 16, None, None, 16, 1, 4, 4, 4, 3, 8, 1, 1, None)
    with pytest.raises(RuntimeError, match="Cout"):
        _ffi.call("rcmvs_conv3d_fwd", 16, 16, None, None, 16, 1, 4, 4, 4, 8, 24, 1, 1, None)


def test_ops_refuse_cpu_tensors():
    import torch
    from rcmvs import ops
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.to_channels_last(torch.zeros(1, 8, 4, 4))


def test_dropin_state_dict_keys_match_reference_checkpoint():
    import numpy as np
    from models import CascadeMVSNet, CascadeMVSNet_eval
    from rcmvs import weights
    path = os.path.join(os.path.dirname(__file__), "golden", "cas_model_000014.npz")
    for cls in (CascadeMVSNet, CascadeMVSNet_eval):
        m = cls(ndepths=[48, 32, 8])
        weights.load_model(m, path, strict=True)           # 238 tensors, strict
        assert set(m.state_dict().keys()) == set(np.load(path).files)
