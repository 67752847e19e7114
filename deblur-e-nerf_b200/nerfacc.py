"""B1 drop-in for the seven ``nerfacc==0.3.1`` symbols the reference imports,
backed by the den_b200 CUDA kernels.

With ``sys.modules["nerfacc"] = deblur_e_nerf_b200.nerfacc``
the reference's ``models/nerf.py``, ``external/utils.py``,
``external/vol_rendering.py``, ``external/ngp.py`` and ``external/mlp.py`` run
unmodified.  Symbols and call sites:

* ``ContractionType``           — ``models/deblur_e_nerf.py:271-275``
* ``OccupancyGrid``             — ``models/nerf.py:98-102,200-204``
* ``ray_marching``              — ``external/utils.py:106-119``
* ``render_weight_from_density``, ``render_weight_from_alpha``,
  ``accumulate_along_rays``     — ``external/vol_rendering.py:12-13,89-122``
* ``ray_aabb_intersect``, ``render_visibility`` — reached inside ``ray_marching``.

Argument names, shapes, dtypes and error behaviour follow upstream (CPU tensors
raise ``NotImplementedError``; ``every_n_step`` raises ``RuntimeError`` outside
training mode).  RNG draws (stratified jitter, occupancy-cell sampling) stay torch
calls in upstream order so the CUDA Philox stream is consumed identically.
"""

import enum

import torch

from . import ops


class ContractionType(enum.Enum):
    AABB = 0
    UN_BOUNDED_TANH = 1
    UN_BOUNDED_SPHERE = 2

    def to_cpp_version(self):
        return self.value


# --------------------------------------------------------------------------- #
class OccupancyGrid(torch.nn.Module):
    """``nerfacc.OccupancyGrid`` (grid.py): the same constructor, buffers /
    state-dict keys (``_roi_aabb, resolution, occs, _binary, grid_coords, grid_indices``),
    properties and ``every_n_step`` / ``_update`` contract, evaluated by the
    ``den_occgrid_*`` kernels: one launch turns the drawn cells + jitter into world points
    (``grid_coords`` gather, jitter, divide, ball test, inverse contraction upstream), three more
    apply the EMA-max, the fp64 mean and the threshold.  The random draws are torch calls in
    upstream order — ``randint`` (uniform cells), [``randint`` (occupied sub-sample)],
    ``rand_like`` (jitter) — so the Philox stream is consumed as upstream consumes it."""

    NUM_DIM = 3

    def __init__(self, roi_aabb, resolution=128, contraction_type=ContractionType.AABB):
        super().__init__()
        res = [resolution] * self.NUM_DIM if isinstance(resolution, int) else \
            [int(r) for r in (resolution.tolist() if torch.is_tensor(resolution) else resolution)]
        roi = roi_aabb.tolist() if torch.is_tensor(roi_aabb) else list(roi_aabb)
        if len(res) != self.NUM_DIM or len(roi) != 2 * self.NUM_DIM:
            raise AssertionError("resolution must have 3 entries and roi_aabb 6")
        self._contraction_type = contraction_type
        self._res_host = res
        self._roi_host = [float(v) for v in roi]
        self.num_cells = res[0] * res[1] * res[2]
        self.register_buffer("_roi_aabb", torch.tensor(self._roi_host, dtype=torch.float32))
        self.register_buffer("resolution", torch.tensor(res, dtype=torch.int32))
        self.register_buffer("occs", torch.zeros(self.num_cells))
        self.register_buffer("_binary", torch.zeros(res, dtype=torch.bool))
        # checkpoint keys of upstream; the kernels derive the lattice coordinates from the index
        axes = [torch.arange(r) for r in res]
        self.register_buffer("grid_coords", torch.cartesian_prod(*axes).reshape(self.num_cells, 3))
        self.register_buffer("grid_indices", torch.arange(self.num_cells))
        self._desc = ops.make_occgrid_desc(self._roi_host, res, contraction_type.to_cpp_version())

    roi_aabb = property(lambda self: self._roi_aabb)
    binary = property(lambda self: self._binary)
    contraction_type = property(lambda self: self._contraction_type)
    device = property(lambda self: self.occs.device)

    # the three draws of an update, overridable per instance (the parity tests inject the oracle's)
    def _draw_randint(self, high, n):
        return torch.randint(high, (n,), device=self.device)

    def _draw_jitter(self, n):
        return torch.rand((n, self.NUM_DIM), dtype=torch.float32, device=self.device)

    @torch.no_grad()
    def _sample_uniform_and_occupied_cells(self, n):
        uniform = self._draw_randint(self.num_cells, n)
        occupied = torch.nonzero(self._binary.flatten())[:, 0]
        if n < len(occupied):
            occupied = occupied[self._draw_randint(len(occupied), n)]
        return torch.cat([uniform, occupied], dim=0)

    @torch.no_grad()
    def _update(self, step, occ_eval_fn, occ_thre=0.01, ema_decay=0.95, warmup_steps=256):
        if not self.occs.is_cuda:
            raise NotImplementedError("Only support cuda inputs.")
        if step < warmup_steps:
            indices, n = None, self.num_cells          # every cell, in order
        else:
            indices = self._sample_uniform_and_occupied_cells(self.num_cells // 4)
            n = indices.numel()
        sphere = self._contraction_type == ContractionType.UN_BOUNDED_SPHERE
        x, keep = ops.occgrid_cell_points(self._desc, indices, self._draw_jitter(n), sphere)
        if sphere:
            # upstream hands occ_eval_fn only the points inside the unit ball (its own draws
            # depend on that count): compact like upstream does
            kept = torch.nonzero(keep)[:, 0]
            x = x[kept]
            indices = kept if indices is None else indices[kept]
        occ = occ_eval_fn(x).squeeze(-1)
        binary = self._binary if (self._binary.is_contiguous() and self._binary.is_cuda) \
            else torch.empty(self._res_host, dtype=torch.bool, device=self.device)
        self.last_mean = ops.occgrid_ema_update(indices, occ.float(), self.occs,
                                                binary.view(torch.uint8), ema_decay, occ_thre)
        self._binary = binary

    @torch.no_grad()
    def every_n_step(self, step, occ_eval_fn, occ_thre=1e-2, ema_decay=0.95,
                     warmup_steps=256, n=16):
        if not self.training:
            raise RuntimeError(
                "You should only call this function only during training. "
                "Please call _update() directly if you want to update the "
                "field during inference.")
        if step % n == 0:
            self._update(step=step, occ_eval_fn=occ_eval_fn, occ_thre=occ_thre,
                         ema_decay=ema_decay, warmup_steps=warmup_steps)


# --------------------------------------------------------------------------- #
@torch.no_grad()
def ray_aabb_intersect(rays_o, rays_d, aabb):
    if not rays_o.is_cuda:
        raise NotImplementedError("Only support cuda inputs.")
    return ops.ray_aabb_intersect(rays_o.contiguous().float(), rays_d.contiguous().float(),
                                  aabb.detach().cpu().tolist())


@torch.no_grad()
def render_visibility(alphas, *, ray_indices=None, packed_info=None, n_rays=None,
                      early_stop_eps=1e-4, alpha_thre=0.0, _offsets=None):
    if not alphas.is_cuda:
        raise NotImplementedError("Only support cuda inputs.")
    if _offsets is None:
        if ray_indices is None:
            raise NotImplementedError("packed_info inputs are not supported; pass ray_indices")
        if n_rays is None:
            n_rays = int(ray_indices.max().item()) + 1 if ray_indices.numel() else 0
        _offsets = ops.offsets_from_ray_indices(ray_indices, n_rays)
    mask, _ = ops.visibility(alphas, _offsets, early_stop_eps, alpha_thre)
    return mask.bool()


@torch.no_grad()
def ray_marching(rays_o, rays_d, t_min=None, t_max=None, scene_aabb=None, grid=None,
                 sigma_fn=None, alpha_fn=None, early_stop_eps=1e-4, alpha_thre=0.0,
                 near_plane=None, far_plane=None, render_step_size=1e-3,
                 stratified=False, cone_angle=0.0):
    """nerfacc.ray_marching: (ray_indices int32 (M,), t_starts (M,1), t_ends (M,1))."""
    if not rays_o.is_cuda:
        raise NotImplementedError("Only support cuda inputs.")
    if alpha_fn is not None and sigma_fn is not None:
        raise ValueError("Only one of `alpha_fn` and `sigma_fn` should be provided.")
    rays_o = rays_o.contiguous().float()
    rays_d = rays_d.contiguous().float()
    n_rays = rays_o.shape[0]
    if t_min is None or t_max is None:
        if scene_aabb is not None:
            t_min, t_max = ray_aabb_intersect(rays_o, rays_d, scene_aabb)
        else:
            t_min = torch.zeros_like(rays_o[..., 0])
            t_max = torch.ones_like(rays_o[..., 0]) * 1e10
    else:
        t_min = t_min.contiguous().float().clone()
        t_max = t_max.contiguous().float().clone()
    jitter = torch.rand_like(t_min) if stratified else None
    ops.clamp_jitter_(t_min, t_max, jitter, near_plane, far_plane, float(render_step_size))

    if grid is not None:
        roi = grid.roi_aabb.detach().cpu().tolist() if not hasattr(grid, "_roi_host") \
            else grid._roi_host
        binary = grid.binary
        ctype = grid.contraction_type.to_cpp_version()
    else:
        roi = [-1e10] * 3 + [1e10] * 3
        binary = torch.ones([1, 1, 1], dtype=torch.bool, device=rays_o.device)
        ctype = ContractionType.AABB.to_cpp_version()
    params = ops.make_march_params(roi, list(binary.shape), ctype, float(render_step_size),
                                   float(cone_angle))
    seg_len = ops.march_segment_length(near_plane, far_plane, float(render_step_size))
    ray_indices, t_starts, t_ends, offsets = ops.march(params, rays_o, rays_d, t_min, t_max, binary,
                                                       seg_len=seg_len)
    t_starts = t_starts[:, None]
    t_ends = t_ends[:, None]

    if (alpha_thre > 0.0 or early_stop_eps > 0.0) and (
            sigma_fn is not None or alpha_fn is not None):
        if grid is not None:
            alpha_thre = min(alpha_thre, grid.occs.mean().item())
        if sigma_fn is not None:
            sigmas = sigma_fn(t_starts, t_ends, ray_indices)
            assert sigmas.shape == t_starts.shape, \
                "sigmas must have shape of (N, 1)! Got {}".format(sigmas.shape)
            alphas = ops.alpha_from_sigma(sigmas.float(), t_starts.reshape(-1), t_ends.reshape(-1))
        else:
            alphas = alpha_fn(t_starts, t_ends, ray_indices)
            assert alphas.shape == t_starts.shape, \
                "alphas must have shape of (N, 1)! Got {}".format(alphas.shape)
            alphas = alphas.reshape(-1).float()
        mask, counts = ops.visibility(alphas, offsets, early_stop_eps, alpha_thre)
        offsets_out = ops.exclusive_scan_i32(counts)
        total = int(offsets_out[-1].item())
        ray_indices, t0, t1 = ops.compact(mask, offsets, offsets_out, ray_indices,
                                          t_starts.reshape(-1), t_ends.reshape(-1), total)
        t_starts, t_ends = t0[:, None], t1[:, None]
    return ray_indices, t_starts, t_ends


def _offsets_for(ray_indices, n_rays, packed_info):
    if ray_indices is None:
        raise NotImplementedError("packed_info inputs are not supported; pass ray_indices")
    if n_rays is None:
        n_rays = int(ray_indices.max().item()) + 1 if ray_indices.numel() else 0
    return ops.offsets_from_ray_indices(ray_indices, n_rays)


def render_weight_from_density(t_starts, t_ends, sigmas, *, packed_info=None,
                               ray_indices=None, n_rays=None):
    if not sigmas.is_cuda:
        raise NotImplementedError("Only support cuda inputs.")
    offsets = _offsets_for(ray_indices, n_rays, packed_info)
    w = ops.weight_from_density(sigmas, t_starts.detach(), t_ends.detach(), offsets)
    return w[:, None]


def render_weight_from_alpha(alphas, *, packed_info=None, ray_indices=None, n_rays=None):
    if not alphas.is_cuda:
        raise NotImplementedError("Only support cuda inputs.")
    offsets = _offsets_for(ray_indices, n_rays, packed_info)
    return ops.weight_from_alpha(alphas, offsets)[:, None]


def accumulate_along_rays(weights, ray_indices, values=None, n_rays=None):
    assert ray_indices.dim() == 1 and weights.dim() == 2
    if not weights.is_cuda:
        raise NotImplementedError("Only support cuda inputs.")
    if values is not None:
        assert values.dim() == 2 and values.shape[0] == weights.shape[0]
    if n_rays is None:
        n_rays = int(ray_indices.max().item()) + 1 if ray_indices.numel() else 0
    offsets = ops.offsets_from_ray_indices(ray_indices, n_rays)
    ri = ray_indices if ray_indices.dtype == torch.int32 else ray_indices.to(torch.int32)
    return ops.accumulate(weights, values, ri.contiguous(), offsets)
