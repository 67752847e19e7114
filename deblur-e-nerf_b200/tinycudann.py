"""B1 drop-in for ``tinycudann.Encoding`` (HashGrid / Linear / fp32), backed by the
den_b200 hash-grid kernels.

Reference call site: ``external/ngp.py:166-175`` —
``tcnn.Encoding(n_input_dims=3, encoding_config=pos_encoding_config,
dtype=torch.float32)``, then ``.n_output_dims``, the flat fp32 ``.params``
parameter (state-dict key ``mlp_base.0.params``) and ``__call__((M,3)) -> (M, L*F)``.
tcnn's own RNG is not reproducible from torch, so parity runs load identical
``params`` into both sides; the default init is U(-1e-4, 1e-4) like upstream.
"""

import torch

from . import ops


class Encoding(torch.nn.Module):
    def __init__(self, n_input_dims, encoding_config, seed=1337, dtype=None):
        super().__init__()
        if n_input_dims != 3:
            raise NotImplementedError("den_b200 HashGrid supports 3 input dims")
        cfg = dict(encoding_config)
        if cfg.get("otype", "HashGrid") != "HashGrid":
            raise NotImplementedError(f"otype {cfg.get('otype')} is not supported")
        if cfg.get("interpolation", "Linear") != "Linear":
            raise NotImplementedError(f"interpolation {cfg.get('interpolation')} is not supported")
        if dtype not in (None, torch.float32):
            raise NotImplementedError("only fp32 parameters are supported (as the reference uses)")
        self.n_input_dims = n_input_dims
        self.encoding_config = cfg
        self.seed = seed
        self.dtype = torch.float32
        self.n_levels = int(cfg.get("n_levels", 16))
        self.n_features_per_level = int(cfg.get("n_features_per_level", 2))
        self.log2_hashmap_size = int(cfg.get("log2_hashmap_size", 19))
        self.base_resolution = int(cfg.get("base_resolution", 16))
        self.per_level_scale = float(cfg.get("per_level_scale", 2.0))
        self.desc, n_entries = ops.make_hashgrid_desc(
            self.n_levels, self.base_resolution, self.per_level_scale, self.log2_hashmap_size,
            self.n_features_per_level)
        self.n_output_dims = self.n_levels * self.n_features_per_level
        gen = torch.Generator().manual_seed(seed)
        init = (torch.rand(n_entries * self.n_features_per_level, generator=gen,
                           dtype=torch.float32) * 2 - 1) * 1e-4
        self.params = torch.nn.Parameter(init)

    def forward(self, x):
        if not x.is_cuda:
            raise NotImplementedError("Only support cuda inputs.")
        x = x.float()
        return ops.hashgrid(x if x.is_contiguous() else x.contiguous(), self.params, self.desc)

    def extra_repr(self):
        return f"n_input_dims={self.n_input_dims}, n_output_dims={self.n_output_dims}, " \
               f"seed={self.seed}, dtype={self.dtype}, hyperparams={self.encoding_config}"


class _Modules:
    Module = Encoding


modules = _Modules()
