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


def test_new_entry_points_validate_arguments_without_a_gpu(den_lib):
    """den_eval_ssim, the radix sort and the raw-event pass reject bad arguments before any CUDA call, and their
    workspace queries are pure host functions (sizes grow with the element count, cover the declared layout)."""
    c = den_lib.cdll
    one = ctypes.c_void_p(8)        # a non-null placeholder: never dereferenced on the rejected paths
    for args, needle in (((one, one, 1, 1, 32, 32, 10, 1.5, 1e-4, 9e-4, one, None), b"odd"),
                         ((one, one, 1, 1, 32, 32, 17, 1.5, 1e-4, 9e-4, one, None), b"odd"),
                         ((one, one, 1, 1, 8, 32, 11, 1.5, 1e-4, 9e-4, one, None), b"smaller than the window"),
                         ((one, one, 1, 1, 32, 32, 11, 0.0, 1e-4, 9e-4, one, None), b"sigma"),
                         ((one, one, 1, 4, 32, 32, 11, 1.5, 1e-4, 9e-4, one, None), b"shape")):
        assert c.den_eval_ssim(*args) < 0 and needle in c.den_last_error(), needle
    assert c.den_radix_sort_pairs_u32(one, one, one, one, one, one, -1, 8, one, 1 << 20, None) < 0
    assert c.den_radix_sort_pairs_u32(one, one, one, one, one, one, 10, 33, one, 1 << 20, None) < 0
    assert b"key_bits" in c.den_last_error()
    assert c.den_radix_sort_pairs_u32(one, one, one, one, one, one, 1 << 31, 8, one, 1 << 20, None) < 0
    assert c.den_radix_sort_pairs_u32(None, None, None, None, None, None, 0, 8, None, 0, None) == 0     # empty: nothing to do
    assert c.den_queue_raw_events(one, one, 10, 0, 480, one, 1 << 30, one, one, one, one, None) < 0
    assert b"sensor" in c.den_last_error()
    assert c.den_queue_raw_events(one, one, 10, 640, 480, one, 16, one, one, one, one, None) < 0
    assert b"workspace" in c.den_last_error()
    assert c.den_compact_queued_events(None, None, None, None, None, 0, None, None, None, None, None, None) == 0
    assert c.den_compact_queued_events(one, one, one, one, one, 5, None, one, one, one, one, None) < 0
    small, large = c.den_queue_events_workspace_bytes(1000), c.den_queue_events_workspace_bytes(1 << 24)
    assert 0 < small < large and large >= 7 * 4 * (1 << 24)           # six sort arrays + the keep flags
    assert c.den_radix_sort_workspace_bytes(1 << 24) >= 2 * 256 * ((1 << 24) // 2048) * 4   # histogram + offsets
    assert c.den_queue_events_workspace_bytes(0) > 0
