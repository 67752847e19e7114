"""ctypes binding of ``lib/libden_b200.so`` — the C-ABI CUDA library.

The product path has NO fallback: if the library is missing or a symbol declared
in ``include/den_b200.h`` is absent, importing the ops raises.  PyTorch only owns
the device memory and the stream; every hot-path kernel is launched through the
C ABI with raw pointers.
"""

import ctypes
import os
import re

from . import _build

_c = ctypes
_P = _c.c_void_p
_I64 = _c.c_int64
_I32 = _c.c_int32
_F = _c.c_float
_D = _c.c_double
_SZ = _c.c_size_t
_INT = _c.c_int

DEN_MAX_LEVELS = 32
ABI_VERSION = 2         # DEN_ABI_VERSION of include/den_b200.h


class HashGridDesc(_c.Structure):
    _fields_ = [
        ("n_levels", _I32), ("n_features", _I32), ("n_agg_levels", _I32), ("reserved", _I32),
        ("scale", _F * DEN_MAX_LEVELS),
        ("resolution", _c.c_uint32 * DEN_MAX_LEVELS),
        ("size", _c.c_uint32 * DEN_MAX_LEVELS),
        ("offset", _c.c_uint32 * DEN_MAX_LEVELS),
    ]


class MarchParams(_c.Structure):
    _fields_ = [
        ("roi", _F * 6), ("res", _I32 * 3), ("contraction", _I32),
        ("step_size", _F), ("cone_angle", _F),
    ]


class OccGridDesc(_c.Structure):
    _fields_ = [("roi", _F * 6), ("res", _I32 * 3), ("contraction", _I32)]


class FieldDesc(_c.Structure):
    _fields_ = [
        ("grid", HashGridDesc), ("aabb", _F * 6), ("contraction", _I32), ("channels", _I32),
        ("hidden_act", _I32), ("density_act", _I32), ("radiance_act", _I32), ("width", _I32),
        ("geo_feat_dim", _I32), ("sh_degree", _I32), ("n_hidden_base", _I32),
        ("n_hidden_head", _I32),
    ]


class FieldParams(_c.Structure):
    _fields_ = [(name, _P) for name in (
        "table", "wb1", "bb1", "wb2", "bb2", "w1", "b1", "w2", "b2", "w3", "b3")]


class AdamTensor(_c.Structure):
    """den_adam_tensor (include/den_b200.h)."""
    _fields_ = [("param", _c.c_void_p), ("grad", _c.c_void_p), ("exp_avg", _c.c_void_p),
                ("exp_avg_sq", _c.c_void_p), ("n", _c.c_int64), ("lr", _c.c_float),
                ("weight_decay", _c.c_float)]


class LpfLossDesc(_c.Structure):
    _fields_ = [("it_sample_size", _I32), ("n_requests", _I32), ("has_reset", _I32),
                ("error_kind", _I32 * 4), ("has_target", _I32 * 4)]


class FieldGrads(_c.Structure):
    _fields_ = [(name, _P) for name in (
        "wb1", "bb1", "wb2", "bb2", "w1", "b1", "w2", "b2", "w3", "b3")]


# name -> (restype, argtypes); mirrors include/den_b200.h declaration by declaration
_SIGNATURES = {
    "den_version": (_INT, []),
    "den_last_error": (_c.c_char_p, []),
    "den_device_sm_count": (_INT, []),
    "den_hashgrid_fwd": (_INT, [_c.POINTER(HashGridDesc), _P, _P, _P, _I64, _P, _P]),
    "den_hashgrid_bwd": (_INT, [_c.POINTER(HashGridDesc), _P, _P, _P, _P, _P, _I64, _P, _P]),
    "den_ray_aabb_intersect": (_INT, [_P, _P, _c.POINTER(_F), _P, _P, _I64, _P]),
    "den_clamp_jitter": (_INT, [_P, _P, _P, _INT, _F, _INT, _F, _F, _I64, _P]),
    "den_march_count": (_INT, [_c.POINTER(MarchParams), _P, _P, _P, _P, _P, _P, _I64, _P]),
    "den_scan_workspace_bytes": (_SZ, [_I64]),
    "den_exclusive_scan_i32": (_INT, [_P, _P, _I64, _P, _SZ, _P]),
    "den_march_write": (_INT, [_c.POINTER(MarchParams), _P, _P, _P, _P, _P, _P, _P, _P, _P,
                               _I64, _I64, _P]),
    "den_march_bound": (_INT, [_c.POINTER(MarchParams), _P, _P, _I32, _P, _I64, _P]),
    "den_march_single": (_INT, [_c.POINTER(MarchParams), _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _P]),
    "den_march_pack": (_INT, [_P, _P, _P, _P, _I64, _P, _P, _P, _P]),
    "den_occgrid_cell_points": (_INT, [_c.POINTER(OccGridDesc), _P, _P, _I64, _P, _P, _P]),
    "den_occgrid_occ": (_INT, [_P, _P, _P, _P, _F, _F, _INT, _F, _F, _I64, _P, _P]),
    "den_occgrid_workspace_bytes": (_SZ, [_I64]),
    "den_occgrid_workspace_init": (_INT, [_P, _I64, _P]),
    "den_occgrid_ema_update": (_INT, [_P, _P, _P, _I64, _F, _F, _P, _I64, _P, _P, _P, _P]),
    "den_rays_from_trajectory": (_INT, [_P, _P, _I64, _P, _P, _P, _I32, _c.POINTER(_F), _P, _P, _I64,
                                        _P]),
    "den_rays_from_trajectory_bwd": (_INT, [_P, _P, _P, _P, _I32, _P, _P, _P, _P, _I64, _P]),
    "den_adam_step": (_INT, [_c.POINTER(AdamTensor), _I32, _c.c_double, _c.c_double, _c.c_double, _I64,
                             _P, _c.c_double, _P, _P]),
    "den_alpha_from_sigma": (_INT, [_P, _P, _P, _P, _I64, _P, _P]),
    "den_visibility": (_INT, [_P, _P, _I64, _F, _F, _P, _P, _P]),
    "den_compact_samples": (_INT, [_P] * 9 + [_I64, _P]),
    "den_compact_samples_ex": (_INT, [_P] * 9 + [_I64, _P, _P, _I32, _P, _P, _P, _P]),
    "den_clamp_offsets": (_INT, [_P, _I64, _I32, _P, _P]),
    "den_weight_from_density_fwd": (_INT, [_P, _P, _P, _P, _I64, _P, _P]),
    "den_weight_from_density_bwd": (_INT, [_P, _P, _P, _P, _I64, _P, _P, _P]),
    "den_weight_from_alpha_fwd": (_INT, [_P, _P, _I64, _P, _P]),
    "den_weight_from_alpha_bwd": (_INT, [_P, _P, _I64, _P, _P, _P]),
    "den_accumulate_fwd": (_INT, [_P, _P, _P, _I64, _I32, _P, _P]),
    "den_accumulate_bwd": (_INT, [_P, _P, _P, _P, _I64, _I32, _P, _P, _P]),
    "den_composite_fwd": (_INT, [_P, _P, _P, _P, _P, _I64, _I32, _P, _P, _P, _P, _P]),
    "den_field_fwd": (_INT, [_c.POINTER(FieldDesc), _c.POINTER(FieldParams), _P, _P, _P, _P, _P,
                             _I64, _P, _P, _P, _P]),
    "den_field_density_at": (_INT, [_c.POINTER(FieldDesc), _c.POINTER(FieldParams), _P, _I64, _P,
                                    _P]),
    "den_contract_samples": (_INT, [_c.POINTER(FieldDesc), _P, _P, _P, _P, _P, _I64, _P, _P, _P]),
    "den_mlp_fwd": (_INT, [_c.POINTER(FieldDesc), _c.POINTER(FieldParams), _P, _P, _P, _P, _P, _P,
                           _I64, _P, _P, _P, _P]),
    "den_mlp_bwd": (_INT, [_c.POINTER(FieldDesc), _c.POINTER(FieldParams), _c.POINTER(FieldGrads),
                           _P, _P, _P, _P, _P, _P, _P, _P, _I64, _P, _P, _P, _P, _P]),
    "den_contract_samples_bwd": (_INT, [_c.POINTER(FieldDesc), _P, _P, _P, _P, _P, _P, _I64, _P, _P,
                                        _P, _P]),
    "den_lpf_fwd": (_INT, [_P, _P, _P, _I32, _I64, _I32, _P, _P]),
    "den_lpf_bwd": (_INT, [_P, _P, _P, _I32, _I64, _I32, _P, _P, _P, _P]),
    "den_lpf_loss_workspace_bytes": (_SZ, [_I32, _I64]),
    "den_lpf_loss_fwd": (_INT, [_c.POINTER(LpfLossDesc), _P, _P, _P, _P, _P, _P, _P, _I64, _P, _P, _P, _P,
                                _P]),
    "den_lpf_loss_bwd": (_INT, [_c.POINTER(LpfLossDesc), _P, _P, _P, _P, _P, _P, _P, _I64, _P, _P, _P, _P,
                                _P, _P, _P, _P]),
    "den_eval_affine_moments": (_INT, [_P, _P, _P, _I32, _I32, _I64, _P, _P]),
    "den_eval_lm_moments": (_INT, [_P, _P, _P, _P, _I32, _I32, _I64, _P, _P]),
    "den_eval_apply": (_INT, [_P, _P, _P, _P, _I32, _I32, _I64, _P, _P, _P]),
    "den_eval_ssim": (_INT, [_P, _P, _I32, _I32, _I32, _I32, _I32, _D, _D, _D, _P, _P]),
    "den_radix_sort_workspace_bytes": (_SZ, [_I64]),
    "den_radix_sort_pairs_u32": (_INT, [_P, _P, _P, _P, _P, _P, _I64, _I32, _P, _SZ, _P]),
    "den_queue_events_workspace_bytes": (_SZ, [_I64]),
    "den_queue_raw_events": (_INT, [_P, _P, _I64, _I32, _I32, _P, _SZ, _P, _P, _P, _P, _P]),
    "den_compact_queued_events": (_INT, [_P, _P, _P, _P, _P, _I64, _P, _P, _P, _P, _P, _P]),
    "den_composite_bwd": (_INT, [_P, _P, _P, _P, _P, _I64, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                 _P, _P]),
}


def header_symbols():
    """Every function name declared in include/den_b200.h (for the ABI test)."""
    path = os.path.join(_build.INCLUDE, "den_b200.h")
    with open(path) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(den_[a-z0-9_]+)\s*\(", text)))


class DenError(RuntimeError):
    pass


class _Library:
    def __init__(self):
        path = _build.LIB_PATH
        if not os.path.exists(path):
            raise DenError(
                f"{path} is missing: the CUDA extension has not been built. Run "
                "`python -c 'import __graft_entry__ as g; g.build()'` — there is no "
                "CPU or PyTorch fallback for the den_b200 kernels.")
        # a library older than its sources must not be used silently: rebuild where nvcc exists,
        # otherwise refuse (DEN_ALLOW_STALE_LIB=1 overrides, for debugging only)
        if not _build.is_current() and os.environ.get("DEN_ALLOW_STALE_LIB") != "1":
            if not _build.have_nvcc():
                raise DenError(
                    f"{path} is stale (csrc/ or include/ changed since it was built) and nvcc is "
                    "not available to rebuild it; there is no CPU or PyTorch fallback.")
            _build.build()
        self.path = path
        self.cdll = _c.CDLL(path)
        for name, (restype, argtypes) in _SIGNATURES.items():
            try:
                fn = getattr(self.cdll, name)
            except AttributeError as exc:
                raise DenError(f"{path} does not export {name}") from exc
            fn.restype = restype
            fn.argtypes = argtypes
        version = self.cdll.den_version()
        if version != ABI_VERSION:
            raise DenError(f"unexpected den_b200 ABI version {version}")

    def last_error(self):
        msg = self.cdll.den_last_error()
        return msg.decode() if msg else ""

    def call(self, name, *args):
        rc = getattr(self.cdll, name)(*args)
        if rc != 0:
            raise DenError(f"{name} failed ({rc}): {self.last_error()}")

    def raw(self, name):
        return getattr(self.cdll, name)


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = _Library()
    return _LIB
