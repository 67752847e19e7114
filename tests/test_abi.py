"""The C-ABI library builds for sm_100a, loads without a GPU, and exports every symbol
``include/den_b200.h`` declares (no compute calls here)."""

import ctypes
import os


def test_library_builds_and_exports_every_declared_symbol(den_lib):
    from deblur_e_nerf_b200 import _lib
    declared = _lib.header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(den_lib.cdll, name), f"{name} declared in den_b200.h but not exported"
    assert set(_lib._SIGNATURES) <= set(declared), set(_lib._SIGNATURES) - set(declared)
    assert set(declared) <= set(_lib._SIGNATURES), set(declared) - set(_lib._SIGNATURES)
    assert den_lib.cdll.den_version() == _lib.ABI_VERSION == 2


def test_library_is_in_tree_and_sm100a(den_lib):
    from deblur_e_nerf_b200 import _build
    assert den_lib.path.startswith(_build.PKG_DIR)
    assert "compute_100a" in " ".join(_build.NVCC_FLAGS)
    assert "-lineinfo" in _build.NVCC_FLAGS


def test_argument_validation_needs_no_gpu(den_lib):
    """Bad arguments are rejected before any CUDA call, with a message."""
    from deblur_e_nerf_b200._lib import HashGridDesc
    desc = HashGridDesc()
    desc.n_levels = 4
    desc.n_features = 3            # unsupported
    rc = den_lib.cdll.den_hashgrid_fwd(ctypes.byref(desc), None, None, None, 8, None, None)
    assert rc < 0
    assert b"n_features" in den_lib.cdll.den_last_error()
    rc = den_lib.cdll.den_composite_fwd(None, None, None, None, None, 4, 2, None, None, None, None,
                                        None)
    assert rc < 0 and b"channels" in den_lib.cdll.den_last_error()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from deblur_e_nerf_b200 import _build, _lib
    monkeypatch.setattr(_build, "LIB_PATH", str(tmp_path / "nope.so"))
    import pytest
    with pytest.raises(_lib.DenError, match="no\\s+CPU or PyTorch fallback"):
        _lib._Library()
