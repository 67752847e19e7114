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
def contract_inv(x, roi, type=ContractionType.AABB):
    """Unit cube -> world (nerfacc helpers_contraction.h; used by the grid update).

    Elementwise torch ops on (n_cells, 3): the heavy part of the occupancy update is
    the density query, which runs on the den_b200 field kernels.
    """
    roi_min, roi_max = roi[:3], roi[3:]
    if type == ContractionType.AABB:
        u = x
    elif type == ContractionType.UN_BOUNDED_TANH:
        u = torch.atanh((x - 0.5) * 2.0) + 0.5
    elif type == ContractionType.UN_BOUNDED_SPHERE:
        u = (x - 0.5) * 4.0
        norm = torch.sqrt((u[..., 0] * u[..., 0] + u[..., 1] * u[..., 1]) + u[..., 2] * u[..., 2])
        outside = norm > 1.0
        safe = torch.where(outside, norm, torch.ones_like(norm))
        warped = (u / safe[..., None]) * (1.0 / (2.0 - safe))[..., None]
        u = torch.where(outside[..., None], warped, u)
        u = u * 0.5 + 0.5
    else:
        raise ValueError(type)
    return u * (roi_max - roi_min) + roi_min


class OccupancyGrid(torch.nn.Module):
    """nerfacc ``OccupancyGrid`` (grid.py): same buffers / state-dict keys / update rule."""

    NUM_DIM = 3

    def __init__(self, roi_aabb, resolution=128, contraction_type=ContractionType.AABB):
        super().__init__()
        if isinstance(resolution, int):
            resolution = [resolution] * self.NUM_DIM
        if isinstance(resolution, (list, tuple)):
            resolution = torch.tensor(resolution, dtype=torch.int32)
        if isinstance(roi_aabb, (list, tuple)):
            roi_aabb = torch.tensor(roi_aabb, dtype=torch.float32)
        assert isinstance(resolution, torch.Tensor) and resolution.shape == (self.NUM_DIM,)
        assert isinstance(roi_aabb, torch.Tensor) and roi_aabb.shape == (2 * self.NUM_DIM,)
        self._contraction_type = contraction_type
        self.num_cells = int(resolution.prod().item())
        self._res_host = [int(r) for r in resolution.tolist()]
        self._roi_host = [float(v) for v in roi_aabb.tolist()]
        self.register_buffer("_roi_aabb", roi_aabb.to(torch.float32))
        self.register_buffer("resolution", resolution)
        self.register_buffer("occs", torch.zeros(self.num_cells))
        self.register_buffer("_binary", torch.zeros(self._res_host, dtype=torch.bool))
        coords = torch.stack(torch.meshgrid(
            [torch.arange(r) for r in self._res_host], indexing="ij"), dim=-1)
        self.register_buffer("grid_coords", coords.reshape(self.num_cells, self.NUM_DIM))
        self.register_buffer("grid_indices", torch.arange(self.num_cells))

    @property
    def roi_aabb(self):
        return self._roi_aabb

    @property
    def binary(self):
        return self._binary

    @property
    def contraction_type(self):
        return self._contraction_type

    @property
    def device(self):
        return self.occs.device

    @torch.no_grad()
    def _sample_uniform_and_occupied_cells(self, n):
        uniform = torch.randint(self.num_cells, (n,), device=self.device)
        occupied = torch.nonzero(self._binary.flatten())[:, 0]
        if n < len(occupied):
            pick = torch.randint(len(occupied), (n,), device=self.device)
            occupied = occupied[pick]
        return torch.cat([uniform, occupied], dim=0)

    @torch.no_grad()
    def _update(self, step, occ_eval_fn, occ_thre=0.01, ema_decay=0.95, warmup_steps=256):
        if step < warmup_steps:
            indices = self.grid_indices
        else:
            indices = self._sample_uniform_and_occupied_cells(self.num_cells // 4)
        coords = self.grid_coords[indices]
        x = (coords + torch.rand_like(coords, dtype=torch.float32)) / self.resolution
        if self._contraction_type == ContractionType.UN_BOUNDED_SPHERE:
            inside = (x - 0.5).norm(dim=1) < 0.5
            x = x[inside]
            indices = indices[inside]
        x = contract_inv(x, roi=self._roi_aabb, type=self._contraction_type)
        occ = occ_eval_fn(x).squeeze(-1)
        self.occs[indices] = torch.maximum(self.occs[indices] * ema_decay, occ)
        self._binary = (
            self.occs > torch.clamp(self.occs.mean(), max=occ_thre)
        ).view(self._binary.shape)

    @torch.no_grad()
    def every_n_step(self, step, occ_eval_fn, occ_thre=1e-2, ema_decay=0.95,
                     warmup_steps=256, n=16):
        if not self.training:
            raise RuntimeError(
                "You should only call this function only during training. "
                "Please call _update() directly if you want to update the "
                "field during inference.")
        if step % n == 0 and self.training:
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
