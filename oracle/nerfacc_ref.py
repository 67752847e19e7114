"""Oracle restatement of the ``nerfacc==0.3.1`` symbols the hot path imports
(TEST INFRASTRUCTURE; pure PyTorch, CPU, every fp32 rounding explicit).

nerfacc 0.3.1 (``environment.yml:32``) is absent from ``/root/reference`` and from
this image, so this file restates its published algorithm: ``nerfacc/grid.py``
(``OccupancyGrid``), ``nerfacc/ray_marching.py``, ``nerfacc/intersection.py``,
``nerfacc/vol_rendering.py``, ``nerfacc/pack.py`` and
``cuda/csrc/{ray_marching,intersection,render_weight,render_transmittance}.cu``
with ``include/helpers_{contraction,math}.h``.  Parity against the upstream
binary is UNPINNED; what pins this file are the reference's call sites
(``models/nerf.py:98-102,200-204,248-251``, ``external/utils.py:106-119``,
``external/vol_rendering.py:89-122``) and the hand-traceable known-answer tests
in ``tests/test_oracle_kat.py``.

Rounding conventions (the CUDA kernels in ``csrc/den_march.cu`` follow these
bit-for-bit so that sample indices compare EXACTLY):

* every fp32 product/sum below is a separately rounded IEEE operation (no FMA
  contraction); the kernels use ``__fmul_rn``/``__fadd_rn``/``__fdiv_rn``;
* ``dot(u, u)`` is ``(ux*ux + uy*uy) + uz*uz``; ``min/max`` follow ``fminf/fmaxf``
  (NaN-ignoring); ``sign(d)`` is ``copysign(1, d)``;
* float -> int conversions truncate toward zero, then clamp to ``[0, res-1]``;
* transmittance for the visibility test is the SEQUENTIAL fp32 product
  ``T_{i+1} = T_i * (1 - alpha_i)`` along each ray (upstream uses a CUB scan whose
  association order is unspecified).
"""

import enum

import torch


# tests/golden/make_golden.py sets this to a list to record the stratified draws
JITTER_LOG = None


class ContractionType(enum.Enum):
    AABB = 0
    UN_BOUNDED_TANH = 1
    UN_BOUNDED_SPHERE = 2

    def to_cpp_version(self):
        return self.value


# --------------------------------------------------------------------------- #
# contraction (helpers_contraction.h)
# --------------------------------------------------------------------------- #
def _roi_split(roi):
    roi = roi.to(torch.float32)
    return roi[:3], roi[3:]


SPHERE_INV_EPS = 1e-10      # helpers_contraction.h: fmaxf(2n - n^2, 1e-10f) in the sphere inverse


def _sqrt(x):
    """Correctly rounded square root, like CUDA's `sqrtf` in the upstream kernels.  torch's CPU
    float32 sqrt (vectorised math library) is off by one ulp for ~0.6 % of inputs on this build,
    which is enough to move an un-contracted point; going through float64 rounds correctly."""
    return torch.sqrt(x.double()).to(x.dtype) if x.dtype == torch.float32 else torch.sqrt(x)


def _dot3(u):
    return (u[..., 0] * u[..., 0] + u[..., 1] * u[..., 1]) + u[..., 2] * u[..., 2]


def contract(x, roi, type=ContractionType.AABB):
    """World -> unit cube.  helpers_contraction.h ``apply_contraction``."""
    roi_min, roi_max = _roi_split(roi)
    u = (x - roi_min) / (roi_max - roi_min)
    if type == ContractionType.AABB:
        return u
    if type == ContractionType.UN_BOUNDED_TANH:
        return torch.tanh(u - 0.5) * 0.5 + 0.5
    if type == ContractionType.UN_BOUNDED_SPHERE:
        u = u * 2.0 - 1.0
        norm = _sqrt(_dot3(u))
        outside = norm > 1.0
        safe = torch.where(outside, norm, torch.ones_like(norm))
        warped = (2.0 - 1.0 / safe)[..., None] * (u / safe[..., None])
        u = torch.where(outside[..., None], warped, u)
        return u * 0.25 + 0.5
    raise ValueError(type)


def contract_inv(x, roi, type=ContractionType.AABB):
    """Unit cube -> world.  helpers_contraction.h ``apply_contraction_inv``."""
    roi_min, roi_max = _roi_split(roi)
    if type == ContractionType.AABB:
        u = x
    elif type == ContractionType.UN_BOUNDED_TANH:
        u = torch.atanh((x - 0.5) * 2.0) + 0.5
    elif type == ContractionType.UN_BOUNDED_SPHERE:
        # helpers_contraction.h `unit_sphere_to_inf`: v / max(2n - n^2, eps)
        u = (x - 0.5) * 4.0
        norm_sq = _dot3(u)
        norm = _sqrt(norm_sq)
        denom = torch.clamp(2.0 * norm - norm_sq, min=SPHERE_INV_EPS)
        u = torch.where((norm > 1.0)[..., None], u / denom[..., None], u)
        u = u * 0.5 + 0.5
    else:
        raise ValueError(type)
    return u * (roi_max - roi_min) + roi_min


# --------------------------------------------------------------------------- #
# occupancy grid (grid.py)
# --------------------------------------------------------------------------- #
class OccupancyGrid(torch.nn.Module):
    """grid.py ``OccupancyGrid``; used at ``models/nerf.py:98-102,200-204``."""

    NUM_DIM = 3

    def __init__(self, roi_aabb, resolution=128, contraction_type=ContractionType.AABB):
        super().__init__()
        if isinstance(resolution, int):
            resolution = [resolution] * self.NUM_DIM
        if isinstance(resolution, (list, tuple)):
            resolution = torch.tensor(resolution, dtype=torch.int32)
        if isinstance(roi_aabb, (list, tuple)):
            roi_aabb = torch.tensor(roi_aabb, dtype=torch.float32)
        assert resolution.shape == (self.NUM_DIM,)
        assert roi_aabb.shape == (2 * self.NUM_DIM,)
        self._contraction_type = contraction_type
        self.num_cells = int(resolution.prod().item())
        self.register_buffer("_roi_aabb", roi_aabb.to(torch.float32))
        self.register_buffer("resolution", resolution)
        self.register_buffer("occs", torch.zeros(self.num_cells))
        self.register_buffer("_binary", torch.zeros(resolution.tolist(), dtype=torch.bool))
        coords = torch.stack(torch.meshgrid(
            [torch.arange(int(r)) for r in resolution.tolist()], indexing="ij"), dim=-1)
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

    @torch.no_grad()
    def query_occ(self, samples):
        return query_grid(samples, self._roi_aabb, self.binary, self.contraction_type)


def _grid_cell_index(unit, res):
    """helpers: ``grid_idx_at`` — truncate, clamp, x-slowest flat index."""
    resf = torch.tensor([float(r) for r in res], dtype=torch.float32)
    scaled = unit * resf
    # truncation toward zero; NaN/inf are mapped to 0 before the clamp (the CUDA
    # kernels do the same with an explicit finite check)
    scaled = torch.where(torch.isfinite(scaled), scaled, torch.zeros_like(scaled))
    ijk = scaled.to(torch.int64)
    hi = torch.tensor([r - 1 for r in res], dtype=torch.int64)
    ijk = torch.minimum(torch.maximum(ijk, torch.zeros_like(ijk)), hi)
    return (ijk[..., 0] * res[1] + ijk[..., 1]) * res[2] + ijk[..., 2]


def query_grid(xyz, roi, binary, ctype):
    """ray_marching.cu ``grid_occupied_at``."""
    roi_min, roi_max = _roi_split(roi)
    res = tuple(int(r) for r in binary.shape)
    unit = contract(xyz, roi, ctype)
    idx = _grid_cell_index(unit, res)
    occ = binary.reshape(-1)[idx]
    if ctype == ContractionType.AABB:
        inside = ((xyz >= roi_min) & (xyz <= roi_max)).all(dim=-1)
        occ = occ & inside
    return occ


# --------------------------------------------------------------------------- #
# ray / AABB intersection (intersection.cu)
# --------------------------------------------------------------------------- #
@torch.no_grad()
def ray_aabb_intersect(rays_o, rays_d, aabb):
    """Slab test, axis by axis, 1e10 on a miss (intersection.cu)."""
    o = rays_o.to(torch.float32)
    d = rays_d.to(torch.float32)
    aabb = aabb.to(torch.float32)
    big = torch.full_like(o[:, 0], 1e10)

    def axis(k):
        a = (aabb[k] - o[:, k]) / d[:, k]
        b = (aabb[k + 3] - o[:, k]) / d[:, k]
        swap = a > b
        return torch.where(swap, b, a), torch.where(swap, a, b)

    tmin, tmax = axis(0)
    tymin, tymax = axis(1)
    miss = (tmin > tymax) | (tymin > tmax)
    tmin = torch.where(tymin > tmin, tymin, tmin)
    tmax = torch.where(tymax < tmax, tymax, tmax)
    tzmin, tzmax = axis(2)
    miss = miss | (tmin > tzmax) | (tzmin > tmax)
    tmin = torch.where(tzmin > tmin, tzmin, tmin)
    tmax = torch.where(tzmax < tmax, tzmax, tmax)
    return torch.where(miss, big, tmin), torch.where(miss, big, tmax)


# --------------------------------------------------------------------------- #
# marching (ray_marching.cu)
# --------------------------------------------------------------------------- #
def _calc_dt(t, cone_angle, dt_min, dt_max):
    return torch.clamp(t * cone_angle, min=dt_min, max=dt_max)


@torch.no_grad()
def march_rays(rays_o, rays_d, t_min, t_max, roi, binary, ctype, step_size,
               cone_angle):
    """ray_marching.cu ``ray_marching_kernel`` for all rays at once.

    Returns ``(ray_indices int64 (M,), t_starts (M,1), t_ends (M,1), num_steps (R,))``
    in ray-major, front-to-back order.
    """
    f32 = torch.float32
    o_all = rays_o.to(f32)
    d_all = rays_d.to(f32)
    R = o_all.shape[0]
    roi_min, roi_max = _roi_split(roi)
    res = tuple(int(r) for r in binary.shape)
    resf = torch.tensor([float(r) for r in res], dtype=f32)
    extent = roi_max - roi_min
    cone = torch.tensor(float(cone_angle), dtype=f32)
    dt_min = torch.tensor(float(step_size), dtype=f32)
    dt_max = torch.tensor(1e10, dtype=f32)
    half = torch.tensor(0.5, dtype=f32)

    t0 = t_min.to(f32).clone()
    t1 = t0 + _calc_dt(t0, cone, dt_min, dt_max)
    tm = (t0 + t1) * half
    far = t_max.to(f32)

    alive = torch.nonzero(tm < far)[:, 0]
    out_ray, out_t0, out_t1 = [], [], []
    while alive.numel() > 0:
        o = o_all[alive]
        d = d_all[alive]
        a_t0, a_t1, a_tm = t0[alive], t1[alive], tm[alive]
        xyz = o + a_tm[:, None] * d
        occ = query_grid(xyz, roi, binary, ctype)

        # occupied: emit, then regular step
        if occ.any():
            out_ray.append(alive[occ])
            out_t0.append(a_t0[occ])
            out_t1.append(a_t1[occ])
        n_t0 = a_t1
        n_t1 = n_t0 + _calc_dt(n_t0, cone, dt_min, dt_max)
        n_tm = (n_t0 + n_t1) * half

        if ctype == ContractionType.AABB:
            emp = ~occ
            if emp.any():
                e_xyz, e_d, e_tm = xyz[emp], d[emp], a_tm[emp]
                inv_d = 1.0 / e_d
                sgn = torch.copysign(torch.ones_like(e_d), e_d)
                u = ((e_xyz - roi_min) / extent) * resf
                txyz = (((torch.floor((u + half) + half * sgn) - u) * inv_d) / resf) * extent
                dist = torch.fmax(
                    torch.fmin(torch.fmin(txyz[:, 0], txyz[:, 1]), txyz[:, 2]),
                    torch.zeros_like(e_tm))
                target = e_tm + dist
                cur = e_tm + dt_min                       # do { _t += dt_min }
                pend = torch.nonzero(cur < target)[:, 0]  # while (_t < target)
                while pend.numel() > 0:
                    cur[pend] = cur[pend] + dt_min
                    pend = pend[cur[pend] < target[pend]]
                dt = _calc_dt(cur, cone, dt_min, dt_max)
                n_t0 = n_t0.clone()
                n_t1 = n_t1.clone()
                n_tm = n_tm.clone()
                n_tm[emp] = cur
                n_t0[emp] = cur - dt * half
                n_t1[emp] = cur + dt * half

        t0[alive], t1[alive], tm[alive] = n_t0, n_t1, n_tm
        alive = alive[n_tm < far[alive]]

    if out_ray:
        ray = torch.cat(out_ray)
        ts = torch.cat(out_t0)
        te = torch.cat(out_t1)
        order = torch.sort(ray, stable=True)[1]           # iteration order is front-to-back
        ray, ts, te = ray[order], ts[order], te[order]
    else:
        ray = torch.zeros(0, dtype=torch.int64)
        ts = torch.zeros(0, dtype=f32)
        te = torch.zeros(0, dtype=f32)
    num_steps = torch.bincount(ray, minlength=R)
    return ray, ts[:, None], te[:, None], num_steps


def _segments(ray_indices, n_rays):
    """(counts, starts) of the ray-major packing."""
    counts = torch.bincount(ray_indices.long(), minlength=n_rays)
    starts = torch.cumsum(counts, 0) - counts
    return counts, starts


@torch.no_grad()
def transmittance_from_alpha_sequential(alphas, ray_indices, n_rays):
    """Exclusive product of (1 - alpha) along each ray, sequential fp32."""
    a = alphas.reshape(-1).to(torch.float32)
    ray = ray_indices.long()
    counts, starts = _segments(ray, n_rays)
    T = torch.ones_like(a)
    if a.numel() == 0:
        return T
    pos = torch.arange(a.numel()) - starts[ray]            # position inside the ray
    run = torch.ones(n_rays, dtype=torch.float32)
    longest = int(counts.max().item())
    for j in range(longest):
        sel = torch.nonzero(pos == j)[:, 0]
        r = ray[sel]
        T[sel] = run[r]
        run[r] = run[r] * (1.0 - a[sel])
    return T


@torch.no_grad()
def render_visibility(alphas, *, ray_indices=None, packed_info=None, n_rays=None,
                      early_stop_eps=1e-4, alpha_thre=0.0):
    """vol_rendering.py ``render_visibility``."""
    if n_rays is None:
        n_rays = int(ray_indices.max().item()) + 1 if ray_indices.numel() else 0
    T = transmittance_from_alpha_sequential(alphas, ray_indices, n_rays)
    vis = T >= early_stop_eps
    if alpha_thre > 0:
        vis = vis & (alphas.reshape(-1) >= alpha_thre)
    return vis


@torch.no_grad()
def ray_marching(rays_o, rays_d, t_min=None, t_max=None, scene_aabb=None, grid=None,
                 sigma_fn=None, alpha_fn=None, early_stop_eps=1e-4, alpha_thre=0.0,
                 near_plane=None, far_plane=None, render_step_size=1e-3,
                 stratified=False, cone_angle=0.0):
    """ray_marching.py ``ray_marching`` (call site ``external/utils.py:106-119``)."""
    if alpha_fn is not None and sigma_fn is not None:
        raise ValueError("Only one of `alpha_fn` and `sigma_fn` should be provided.")
    if t_min is None or t_max is None:
        if scene_aabb is not None:
            t_min, t_max = ray_aabb_intersect(rays_o, rays_d, scene_aabb)
        else:
            t_min = torch.zeros_like(rays_o[..., 0])
            t_max = torch.ones_like(rays_o[..., 0]) * 1e10
    if near_plane is not None:
        t_min = torch.clamp(t_min, min=near_plane)
    if far_plane is not None:
        t_max = torch.clamp(t_max, max=far_plane)
    if stratified:
        jitter = torch.rand_like(t_min)
        if JITTER_LOG is not None:
            JITTER_LOG.append(jitter.clone())
        t_min = t_min + jitter * render_step_size

    if grid is not None:
        roi, binary, ctype = grid.roi_aabb, grid.binary, grid.contraction_type
    else:
        roi = torch.tensor([-1e10] * 3 + [1e10] * 3, dtype=torch.float32)
        binary = torch.ones([1, 1, 1], dtype=torch.bool)
        ctype = ContractionType.AABB

    ray_indices, t_starts, t_ends, _ = march_rays(
        rays_o, rays_d, t_min, t_max, roi, binary, ctype,
        float(render_step_size), float(cone_angle))
    ray_indices = ray_indices.to(torch.int32)

    if (alpha_thre > 0.0 or early_stop_eps > 0.0) and (
            sigma_fn is not None or alpha_fn is not None):
        if grid is not None:
            alpha_thre = min(alpha_thre, grid.occs.mean().item())
        if sigma_fn is not None:
            sigmas = sigma_fn(t_starts, t_ends, ray_indices)
            assert sigmas.shape == t_starts.shape, \
                "sigmas must have shape of (N, 1)! Got {}".format(sigmas.shape)
            alphas = 1.0 - torch.exp(-sigmas * (t_ends - t_starts))
        else:
            alphas = alpha_fn(t_starts, t_ends, ray_indices)
            assert alphas.shape == t_starts.shape
        masks = render_visibility(alphas, ray_indices=ray_indices,
                                  early_stop_eps=early_stop_eps, alpha_thre=alpha_thre,
                                  n_rays=rays_o.shape[0])
        ray_indices, t_starts, t_ends = ray_indices[masks], t_starts[masks], t_ends[masks]
    return ray_indices, t_starts, t_ends


# --------------------------------------------------------------------------- #
# weights / accumulation (vol_rendering.py, render_weight.cu)
# --------------------------------------------------------------------------- #
def _exclusive_sum_by_ray(values, ray_indices, n_rays):
    """Exclusive prefix sum along each ray (differentiable)."""
    v = values.reshape(-1)
    if v.numel() == 0:
        return v
    ray = ray_indices.long()
    _, starts = _segments(ray, n_rays)
    wide = v.double()
    excl = torch.cumsum(wide, 0) - wide
    base = excl[starts.clamp(max=v.numel() - 1)][ray]
    return (excl - base).to(v.dtype)


def render_weight_from_density(t_starts, t_ends, sigmas, *, packed_info=None,
                               ray_indices=None, n_rays=None):
    """w_i = exp(-sum_{j<i} sigma_j dt_j) * (1 - exp(-sigma_i dt_i))  -> (M, 1)."""
    if n_rays is None:
        n_rays = int(ray_indices.max().item()) + 1 if ray_indices.numel() else 0
    sdt = (sigmas * (t_ends - t_starts)).reshape(-1)
    cs = _exclusive_sum_by_ray(sdt, ray_indices, n_rays)
    w = torch.exp(-cs) * (1.0 - torch.exp(-sdt))
    return w[:, None]


def render_weight_from_alpha(alphas, *, packed_info=None, ray_indices=None, n_rays=None):
    """w_i = alpha_i * prod_{j<i} (1 - alpha_j)  -> (M, 1)."""
    if n_rays is None:
        n_rays = int(ray_indices.max().item()) + 1 if ray_indices.numel() else 0
    a = alphas.reshape(-1)
    log_keep = torch.log1p(-a.double().clamp(max=1 - 1e-12))
    cs = _exclusive_sum_by_ray(log_keep, ray_indices, n_rays)
    return (torch.exp(cs).to(a.dtype) * a)[:, None]


def accumulate_along_rays(weights, ray_indices, values=None, n_rays=None):
    """out[ray] += w * v  (scatter_add); zero rows for rays without samples."""
    assert weights.dim() == 2 and weights.shape[-1] == 1
    src = weights if values is None else weights * values
    if n_rays is None:
        n_rays = int(ray_indices.max().item()) + 1 if ray_indices.numel() else 0
    out = torch.zeros((n_rays, src.shape[-1]), dtype=src.dtype, device=src.device)
    if src.shape[0] == 0:
        return out
    index = ray_indices.long()[:, None].expand(-1, src.shape[-1])
    return out.scatter_add(0, index, src)


def unpack_info(packed_info, n_samples=None):
    counts = packed_info[:, 1].long()
    return torch.repeat_interleave(torch.arange(len(counts)), counts)
