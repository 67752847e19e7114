"""Oracle restatement of ``tinycudann.Encoding`` for ``otype: HashGrid``,
``interpolation: Linear``, fp32 (TEST INFRASTRUCTURE; pure PyTorch, CPU).

tiny-cuda-nn is an UN-PINNED git dependency of the reference
(``environment.yml:33``) and is absent from ``/root/reference`` and this image, so
this file restates its published algorithm
(``include/tiny-cuda-nn/encodings/grid.h``: ``grid_scale``, ``grid_resolution``,
the offset table, ``pos_fract``, ``grid_index``, ``coherent_prime_hash``;
``kernel_grid`` forward; parameters U(-1e-4, 1e-4)).  Parity against the upstream
binary is UNPINNED; the known-answer tests in ``tests/test_oracle_kat.py`` pin
this file against closed forms, and the reference's only call site
(``external/ngp.py:166-175``: ``tcnn.Encoding(n_input_dims, encoding_config,
dtype=torch.float32)``, ``.n_output_dims``, ``.params``, ``__call__((M,3)) ->
(M, L*F)``) fixes the interface.

Conventions stated here and followed bit-for-bit by the CUDA kernels
(``csrc/den_hashgrid.cu``):

* per-level ``scale`` is evaluated ONCE on the host in IEEE fp32
  (``exp2(level * log2(per_level_scale)) * base - 1``) and handed to both sides,
  so libm/CUDA ``exp2f`` differences cannot leak in;
* ``pos = fma(scale, x, 0.5)``; ``cell = floor(pos)``; ``frac = pos - cell``;
  ``cell`` is taken through ``(uint32)(int32)`` like upstream;
* corner ``c`` (bit ``d`` of ``c`` selects ``cell_d + 1`` on axis ``d``) has weight
  ``prod_d (bit ? frac_d : 1 - frac_d)``; features accumulate corner 0..7 in order;
* dense index: ``sum_d cell_d * stride_d`` with ``stride *= res`` while
  ``stride <= table_size`` (uint32 arithmetic); hashed levels
  (``table_size < stride`` after the loop) use
  ``x*1 ^ y*2654435761 ^ z*805459861``; final ``% table_size``.  Nothing is
  clamped: the ``+0.5`` offset and out-of-range inputs alias through the modulo
  exactly as upstream does.
"""

import math

import numpy as np
import torch

PRIMES = (1, 2654435761, 805459861)
_U32 = 0xFFFFFFFF


def grid_level_table(n_levels, base_resolution, per_level_scale,
                     log2_hashmap_size, n_pos_dims=3):
    """Per-level (scale fp32, resolution, table entries, entry offset).

    Follows tcnn ``grid.h``: ``grid_scale`` / ``grid_resolution`` and the
    constructor's offset table (entries rounded up to a multiple of 8, capped at
    ``2**log2_hashmap_size``; the cube is capped at ``uint32 max / 2``).
    """
    log2_s = np.float32(np.log2(np.float32(per_level_scale)))
    scales, resolutions, sizes, offsets = [], [], [], []
    offset = 0
    for level in range(n_levels):
        scale = np.float32(
            np.exp2(np.float32(np.float32(level) * log2_s)) * np.float32(base_resolution)
            - np.float32(1.0)
        )
        res = int(math.ceil(float(scale))) + 1
        max_params = _U32 // 2
        dense = max_params if float(res) ** n_pos_dims > float(max_params) else res ** n_pos_dims
        dense = ((dense + 7) // 8) * 8
        size = min(dense, 1 << log2_hashmap_size)
        scales.append(scale)
        resolutions.append(res)
        sizes.append(size)
        offsets.append(offset)
        offset += size
    return (np.asarray(scales, dtype=np.float32), np.asarray(resolutions, dtype=np.int64),
            np.asarray(sizes, dtype=np.int64), np.asarray(offsets, dtype=np.int64), offset)


def hashgrid_indices_weights(x, scales, resolutions, sizes):
    """Entry indices (M, L, 8) int64 (within-level) and weights (M, L, 8).

    ``x`` is (M, 3) float32/float64; the weights are differentiable in ``x``.
    """
    M = x.shape[0]
    L = len(scales)
    scale_t = torch.as_tensor(np.asarray(scales), dtype=x.dtype)            # (L)
    # pos = fma(scale, x, 0.5): emulate the single rounding in float64 then round
    if x.dtype == torch.float32:
        pos = (scale_t.double()[None, :, None] * x.double()[:, None, :] + 0.5).float()
        if x.requires_grad:
            # keep the autograd path (values identical up to the fma rounding)
            pos = pos.detach() + (scale_t[None, :, None] * x[:, None, :]
                                  - (scale_t[None, :, None] * x[:, None, :]).detach())
    else:
        pos = scale_t[None, :, None] * x[:, None, :] + 0.5                   # (M, L, 3)
    cell_f = torch.floor(pos.detach())
    frac = pos - cell_f                                                     # (M, L, 3)
    # (uint32)(int32) cast of the floored value
    cell = cell_f.to(torch.int64) & _U32                                    # (M, L, 3)

    res = torch.as_tensor(np.asarray(resolutions), dtype=torch.int64)       # (L)
    size = torch.as_tensor(np.asarray(sizes), dtype=torch.int64)            # (L)

    # stride table following the upstream loop
    stride = torch.ones(L, dtype=torch.int64)
    strides = []
    for d in range(3):
        active = stride <= size
        strides.append(torch.where(active, stride, torch.zeros_like(stride)))
        stride = torch.where(active, (stride * res) & _U32, stride)
    hashed = size < stride                                                  # (L)

    idx_all, w_all = [], []
    for corner in range(8):
        w = torch.ones(M, L, dtype=x.dtype)
        dense = torch.zeros(M, L, dtype=torch.int64)
        hsh = torch.zeros(M, L, dtype=torch.int64)
        for d in range(3):
            bit = (corner >> d) & 1
            w = w * (frac[..., d] if bit else (1 - frac[..., d]))
            c = (cell[..., d] + bit) & _U32
            dense = (dense + ((c * strides[d][None, :]) & _U32)) & _U32
            hsh = hsh ^ ((c * PRIMES[d]) & _U32)
        idx = torch.where(hashed[None, :], hsh, dense) % size[None, :]
        idx_all.append(idx)
        w_all.append(w)
    return torch.stack(idx_all, dim=-1), torch.stack(w_all, dim=-1)


def hashgrid_encode(x, params, scales, resolutions, sizes, offsets, n_features):
    """(M, 3) -> (M, L*F).  Differentiable in ``params`` and ``x``."""
    idx, w = hashgrid_indices_weights(x, scales, resolutions, sizes)        # (M, L, 8)
    off = torch.as_tensor(np.asarray(offsets), dtype=torch.int64)
    table = params.view(-1, n_features)
    feats = table[(idx + off[None, :, None]).reshape(-1)]                   # (M*L*8, F)
    feats = feats.view(*idx.shape, n_features).to(w.dtype)                  # (M, L, 8, F)
    out = torch.zeros(idx.shape[0], idx.shape[1], n_features, dtype=w.dtype)
    for corner in range(8):                                                 # corner order 0..7
        out = out + w[..., corner, None] * feats[..., corner, :]
    return out.reshape(idx.shape[0], idx.shape[1] * n_features)             # explicit width: M may be 0


class Encoding(torch.nn.Module):
    """Drop-in for ``tinycudann.Encoding`` (HashGrid / Linear / fp32 only)."""

    def __init__(self, n_input_dims, encoding_config, seed=1337, dtype=None):
        super().__init__()
        if n_input_dims != 3:
            raise NotImplementedError("oracle HashGrid supports 3 input dims")
        cfg = dict(encoding_config)
        if cfg.get("otype", "HashGrid") != "HashGrid":
            raise NotImplementedError(cfg.get("otype"))
        if cfg.get("interpolation", "Linear") != "Linear":
            raise NotImplementedError(cfg.get("interpolation"))
        self.n_input_dims = n_input_dims
        self.encoding_config = cfg
        self.n_levels = int(cfg.get("n_levels", 16))
        self.n_features_per_level = int(cfg.get("n_features_per_level", 2))
        self.log2_hashmap_size = int(cfg.get("log2_hashmap_size", 19))
        self.base_resolution = int(cfg.get("base_resolution", 16))
        self.per_level_scale = float(cfg.get("per_level_scale", 2.0))
        (self.scales, self.resolutions, self.sizes, self.offsets,
         n_entries) = grid_level_table(self.n_levels, self.base_resolution,
                                       self.per_level_scale, self.log2_hashmap_size)
        self.n_output_dims = self.n_levels * self.n_features_per_level
        self.dtype = torch.float32 if dtype is None else dtype
        gen = torch.Generator().manual_seed(seed)
        init = (torch.rand(n_entries * self.n_features_per_level, generator=gen,
                           dtype=torch.float32) * 2 - 1) * 1e-4
        self.params = torch.nn.Parameter(init)

    def forward(self, x):
        out = hashgrid_encode(x.to(self.params.dtype) if self.params.dtype == torch.float64
                              else x.float(), self.params, self.scales, self.resolutions,
                              self.sizes, self.offsets, self.n_features_per_level)
        return out


# upstream module-level helpers some callers probe
modules = type("modules", (), {"Module": Encoding})
