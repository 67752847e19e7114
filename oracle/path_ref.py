"""CPU restatement of the reference's OWN hot-path files (TEST INFRASTRUCTURE).

Pure PyTorch (fp32 by default, fp64 when the modules are ``.double()``-ed), written
in this repo's own words so that it can travel to the GPU box where
``/root/reference`` does not exist.  It is PINNED: ``tests/test_oracle_vs_reference.py``
runs every class below side by side with the reference's unmodified file (imported
through ``oracle/ref_shim.py``) on the same seeded inputs and parameters, and
``tests/golden/*.npz`` (made by ``tests/golden/make_golden.py`` from the reference's
files) pin it again on the GPU box.

Reference lines restated (paths relative to ``deblur_e_nerf/``):
  SHEncoder            external/sh_encoder.py:28-193 (degree <= 4)
  MLP                  external/mlp.py:26-113
  trunc_exp            external/ngp.py:45-65
  contraction          external/ngp.py:68-106
  NGPField             external/ngp.py:109-280
  render_rays          external/utils.py:38-140 + external/vol_rendering.py:16-128
  NeRF                 models/nerf.py:31-286
  LinearTrajectory     models/trajectories.py:8-90 + utils/tensor_ops.py:87-184
  ContrastThreshold    models/event_generation_params.py:8-118
  RefractoryPeriod     models/event_generation_params.py:121-237
  foh_discretise       utils/control.py:29-123 (is_efficient, state preserved)
  PixelBandwidth       models/pixel_bandwidth.py:63-494
  EventLoss            loss_metric/loss.py:8-96
  EventRenderer        models/deblur_e_nerf.py:396-586,1129-1308 (hot-path host)
"""

import math

import torch
import torch.nn.functional as F

from . import nerfacc_ref as nacc
from . import roma_ref as roma
from . import tcnn_ref

NS_TO_S = 1e-9


# --------------------------------------------------------------------------- #
# field
# --------------------------------------------------------------------------- #
def sh_encode(d, degree=4):
    """Real spherical harmonics of a unit vector, degree <= 4 (16 outputs)."""
    x, y, z = d.unbind(-1)
    xy, xz, yz = x * y, x * z, y * z
    x2, y2, z2 = x * x, y * y, z * z
    cols = [torch.full_like(x, 0.28209479177387814)]
    if degree > 1:
        cols += [-0.48860251190291987 * y, 0.48860251190291987 * z, -0.48860251190291987 * x]
    if degree > 2:
        cols += [1.0925484305920792 * xy, -1.0925484305920792 * yz,
                 0.94617469575755997 * z2 - 0.31539156525251999,
                 -1.0925484305920792 * xz,
                 0.54627421529603959 * x2 - 0.54627421529603959 * y2]
    if degree > 3:
        cols += [0.59004358992664352 * y * (-3.0 * x2 + y2),
                 2.8906114426405538 * xy * z,
                 0.45704579946446572 * y * (1.0 - 5.0 * z2),
                 0.3731763325901154 * z * (5.0 * z2 - 3.0),
                 0.45704579946446572 * x * (1.0 - 5.0 * z2),
                 1.4453057213202769 * z * (x2 - y2),
                 0.59004358992664352 * x * (-x2 + 3.0 * y2)]
    if degree > 4:
        raise NotImplementedError("oracle SH goes up to degree 4 (all shipped configs)")
    return torch.stack(cols, dim=-1)


class _TruncExp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(torch.clamp(x, max=15))


ACTIVATIONS = {
    "softplus100": lambda v: F.softplus(v, beta=100),
    "relu": torch.relu,
    "softplus1": lambda v: F.softplus(v, beta=1),
    "sigmoid": torch.sigmoid,
    "shifted_trunc_exp": lambda v: _TruncExp.apply(v - 1),
    "shifted_softplus": lambda v: F.softplus(v - 1, beta=1),
    "identity": lambda v: v,
}
HIDDEN_ACT = {"softplus": "softplus100", "relu": "relu"}
DENSITY_ACT = {"shifted_trunc_exp": "shifted_trunc_exp", "softplus": "softplus1",
               "shifted_softplus": "shifted_softplus"}
RADIANCE_ACT = {"softplus": "softplus1", "sigmoid": "sigmoid"}


class MLP(torch.nn.Module):
    """Linear stack with the reference's parameter names (hidden_layers.N, output_layer);
    PyTorch default init because the reference passes every ``*_init=None``."""

    def __init__(self, input_dim, output_dim, depth, width, hidden_act, output_act):
        super().__init__()
        self.hidden_layers = torch.nn.ModuleList()
        fan_in = input_dim
        for _ in range(depth):
            self.hidden_layers.append(torch.nn.Linear(fan_in, width))
            fan_in = width
        self.output_layer = torch.nn.Linear(fan_in, output_dim)
        self.hidden_act = hidden_act
        self.output_act = output_act

    def forward(self, x):
        for layer in self.hidden_layers:
            x = ACTIVATIONS[self.hidden_act](layer(x))
        return ACTIVATIONS[self.output_act](self.output_layer(x))


def contract_field(x, aabb, contraction):
    """Field-side contraction (external/ngp.py:68-106,230-238)."""
    lo, hi = aabb[:3], aabb[3:]
    u = (x - lo) / (hi - lo)
    if contraction == nacc.ContractionType.UN_BOUNDED_SPHERE:
        u = u * 2 - 1
        mag = u.norm(dim=-1, keepdim=True)
        u = torch.where(mag > 1, (2 - 1 / mag) * (u / mag), u)
        u = u / 4 + 0.5
    elif contraction == nacc.ContractionType.UN_BOUNDED_TANH:
        u = (torch.tanh(u - 0.5) + 1) / 2
    return u


class NGPField(torch.nn.Module):
    def __init__(self, aabb, contraction, pos_encoding, sh_degree, base_cfg, head_cfg,
                 radiance_dim):
        super().__init__()
        self.register_buffer("aabb", torch.as_tensor(aabb, dtype=torch.float32))
        self.contraction = contraction
        self.sh_degree = sh_degree
        self.geo_feat_dim = base_cfg["geo_feat_dim"]
        self.density_act = DENSITY_ACT[base_cfg["density_activation"]]
        enc = tcnn_ref.Encoding(3, pos_encoding)
        self.mlp_base = torch.nn.Sequential(
            enc, MLP(enc.n_output_dims, 1 + self.geo_feat_dim, base_cfg["n_hidden_layers"],
                     base_cfg["n_neurons"], HIDDEN_ACT[base_cfg["hidden_activation"]], "identity"))
        self.mlp_head = MLP(sh_degree ** 2 + self.geo_feat_dim, radiance_dim,
                            head_cfg["n_hidden_layers"], head_cfg["n_neurons"],
                            HIDDEN_ACT[head_cfg["hidden_activation"]],
                            RADIANCE_ACT[head_cfg["radiance_activation"]])

    def query_density(self, x, return_feat=False):
        u = contract_field(x, self.aabb, self.contraction)
        selector = ((u > 0.0) & (u < 1.0)).all(dim=-1)
        y = self.mlp_base(u.reshape(-1, 3)).reshape(*u.shape[:-1], 1 + self.geo_feat_dim).to(u)
        raw, geo = y[..., :1], y[..., 1:]
        density = ACTIVATIONS[self.density_act](raw) * selector[..., None]
        return (density, geo) if return_feat else density

    def forward(self, positions, directions):
        density, geo = self.query_density(positions, return_feat=True)
        h = torch.cat([sh_encode(directions.reshape(-1, 3), self.sh_degree),
                       geo.reshape(-1, self.geo_feat_dim)], dim=-1)
        rgb = self.mlp_head(h).reshape(*geo.shape[:-1], -1).to(geo)
        return rgb, density


# --------------------------------------------------------------------------- #
# renderer
# --------------------------------------------------------------------------- #
def render_rays(field, grid, origins, dirs, scene_aabb, near, far, step, bkgd, cone_angle,
                early_stop_eps, alpha_thre, chunk_size, stratified, jitter=None):
    """render_image + rendering.  ``jitter`` (R,) replaces the stratified rand draw so
    that both sides of a parity test consume identical numbers."""
    shape = origins.shape
    o = origins.reshape(-1, 3)
    d = dirs.reshape(-1, 3)
    n_rays = o.shape[0]
    chunk = n_rays if field.training else chunk_size
    cols, opas, deps, total = [], [], [], 0
    for i in range(0, max(n_rays, 1), max(chunk, 1)):
        co, cd = o[i:i + chunk], d[i:i + chunk]

        def sigma_fn(ts, te, ri):
            ri = ri.long()
            pos = co[ri] + cd[ri] * (ts + te) / 2.0
            return field.query_density(pos)

        kw = {}
        if jitter is not None and stratified:
            with torch.no_grad():
                if scene_aabb is not None:
                    t_min, t_max = nacc.ray_aabb_intersect(co, cd, scene_aabb)
                else:
                    t_min = torch.zeros_like(co[:, 0])
                    t_max = torch.full_like(co[:, 0], 1e10)
                if near is not None:
                    t_min = torch.clamp(t_min, min=near)
                if far is not None:
                    t_max = torch.clamp(t_max, max=far)
                t_min = t_min + jitter[i:i + chunk] * step
            ri, ts, te = nacc.ray_marching(
                co, cd, t_min=t_min, t_max=t_max, grid=grid, sigma_fn=sigma_fn,
                render_step_size=step, stratified=False, cone_angle=cone_angle,
                early_stop_eps=early_stop_eps, alpha_thre=alpha_thre)
        else:
            ri, ts, te = nacc.ray_marching(
                co, cd, scene_aabb=scene_aabb, grid=grid, sigma_fn=sigma_fn, near_plane=near,
                far_plane=far, render_step_size=step, stratified=stratified,
                cone_angle=cone_angle, early_stop_eps=early_stop_eps, alpha_thre=alpha_thre)
        ril = ri.long()
        pos = co[ril] + cd[ril] * (ts + te) / 2.0
        rgb, sigma = field(pos, cd[ril])
        w = nacc.render_weight_from_density(ts, te, sigma, ray_indices=ri, n_rays=co.shape[0])
        col = nacc.accumulate_along_rays(w, ri, values=rgb, n_rays=co.shape[0])
        opa = nacc.accumulate_along_rays(w, ri, values=None, n_rays=co.shape[0])
        dep = nacc.accumulate_along_rays(w, ri, values=(ts + te) / 2.0, n_rays=co.shape[0])
        if bkgd is not None:
            col = col + bkgd * (1.0 - opa)
        cols.append(col)
        opas.append(opa)
        deps.append(dep)
        total += len(ts)
    col, opa, dep = torch.cat(cols), torch.cat(opas), torch.cat(deps)
    return (col.view(*shape[:-1], -1), opa.view(*shape[:-1], -1), dep.view(*shape[:-1], -1),
            total)


class NeRF(torch.nn.Module):
    """models/nerf.py NeRF for ``arch: ngp`` — same constructor argument meaning, same
    state-dict keys."""

    def __init__(self, aabb, contraction_type, occ_grid_config, near_plane, far_plane,
                 render_step_size, render_bkgd, cone_angle, early_stop_eps, alpha_thre,
                 test_chunk_size, arch_config, radiance_dim, opacity_eps=1e-10):
        super().__init__()
        self.register_buffer("aabb", torch.tensor(aabb, dtype=torch.float32), persistent=False)
        self.contraction_type = contraction_type
        self.occ_cfg = dict(occ_grid_config)
        self.near_plane, self.far_plane = near_plane, far_plane
        self.register_buffer("render_step_size", torch.tensor(render_step_size),
                             persistent=False)
        if render_bkgd == "parameter":
            # parametrize.register_parametrization(self, "render_bkgd", Softplus) stores
            # right_inverse(ones): keep the same key `parametrizations.render_bkgd.original`
            self.parametrizations = torch.nn.ModuleDict({"render_bkgd": _Original(
                softplus_inverse(torch.ones(radiance_dim)))})
        else:
            self.parametrizations = None
            assert render_bkgd is None
        self.cone_angle = cone_angle
        self.early_stop_eps = early_stop_eps
        self.alpha_thre = alpha_thre
        self.test_chunk_size = test_chunk_size
        self.opacity_eps = opacity_eps
        self.occupancy_grid = nacc.OccupancyGrid(aabb, self.occ_cfg["resolution"],
                                                 contraction_type)
        self.radiance_field = NGPField(
            aabb, contraction_type, arch_config["pos_encoding"],
            arch_config["dir_encoding"]["degree"], arch_config["mlp_base"],
            arch_config["mlp_head"], radiance_dim)

    @property
    def render_bkgd(self):
        if self.parametrizations is None:
            return None
        return F.softplus(self.parametrizations["render_bkgd"].original)

    def update_occ_grid(self, step, T_wc_position=None):
        def occ_eval_fn(x):
            if self.cone_angle > 0.0:
                ids = torch.randint(0, len(T_wc_position), (x.shape[0],))
                t = (T_wc_position[ids] - x).norm(dim=-1, keepdim=True)
                step_size = torch.clamp(t * self.cone_angle, min=self.render_step_size)
                if self.near_plane is not None and self.far_plane is not None:
                    step_size = torch.where((t > self.near_plane) & (t < self.far_plane),
                                            step_size, torch.zeros_like(step_size))
            else:
                step_size = self.render_step_size
            return self.radiance_field.query_density(x) * step_size

        self.occupancy_grid.every_n_step(step, occ_eval_fn, self.occ_cfg["occ_thre"],
                                         self.occ_cfg["ema_decay"],
                                         self.occ_cfg["warmup_steps"], self.occ_cfg["n"])

    @staticmethod
    def pixel_params_to_ray(intrinsics_inverse, pixel_position, T_wc_position, T_wc_orientation):
        homog = torch.cat((pixel_position, torch.ones_like(pixel_position[..., :1])), dim=-1)
        d = (T_wc_orientation @ (intrinsics_inverse @ homog.unsqueeze(-1))).squeeze(-1)
        d = d / torch.linalg.vector_norm(d, dim=-1, keepdim=True)
        return T_wc_position, d

    def forward(self, ray_origin, ray_direction, jitter=None):
        aabb = self.aabb if self.contraction_type == nacc.ContractionType.AABB else None
        rad, opa, dep, n_samples = render_rays(
            self.radiance_field, self.occupancy_grid, ray_origin, ray_direction, aabb,
            self.near_plane, self.far_plane, float(self.render_step_size), self.render_bkgd,
            self.cone_angle, self.early_stop_eps, self.alpha_thre, self.test_chunk_size,
            stratified=self.radiance_field.training, jitter=jitter)
        rad, opa, dep = rad.squeeze(-1), opa.squeeze(-1), dep.squeeze(-1)
        dep = dep / (opa + self.opacity_eps)
        n_rays = ray_origin.numel() // ray_origin.shape[-1]
        return rad, opa, dep, n_samples / n_rays


class _Original(torch.nn.Module):
    """Holds ``original`` like torch's ParametrizationList does (state-dict key parity)."""

    def __init__(self, value):
        super().__init__()
        self.original = torch.nn.Parameter(value)


def softplus_inverse(v, beta=1.0, threshold=20.0):
    """utils/modules.py Softplus.right_inverse."""
    inv = torch.log(torch.exp(beta * v) - 1) / beta
    return torch.where(v * beta > threshold, v, inv)


# --------------------------------------------------------------------------- #
# trajectory
# --------------------------------------------------------------------------- #
def _full_rotvec(q):
    """utils/tensor_ops.py unitquat_to_full_rotvec: angles in [0, 2 pi]."""
    vec = q[..., :3]
    angle = 2 * torch.atan2(torch.norm(vec, dim=-1), q[..., 3])
    small = angle.abs() <= 1e-3
    safe = torch.where(small, torch.ones_like(angle), angle)
    scale = torch.where(small, 2 + angle ** 2 / 12 + 7 * angle ** 4 / 2880,
                        safe / torch.sin(safe / 2))
    return scale[..., None] * vec


def slerp(q0, q1, w):
    """utils/tensor_ops.py unitquat_slerp(shortest_path=True) with per-pair steps."""
    q1 = torch.where(torch.sum(q0 * q1, dim=-1, keepdim=True) < 0, -q1, q1)
    rel = roma.quat_product(roma.quat_conjugation(q0), q1)
    rot = roma.rotvec_to_unitquat((w[..., None] * _full_rotvec(rel)).reshape(-1, 3))
    return roma.quat_product(q0, rot.reshape(*q0.shape))


class LinearTrajectory(torch.nn.Module):
    def __init__(self, position, orientation_quat, timestamp):
        super().__init__()
        self.register_buffer("T_wc_position", position, persistent=False)
        self.register_buffer("T_wc_orientation_quat", orientation_quat, persistent=False)
        self.register_buffer("T_wc_timestamp", timestamp.contiguous(), persistent=False)
        self.register_buffer("bin_width", timestamp.diff(), persistent=False)

    def forward(self, ts):
        right = torch.searchsorted(self.T_wc_timestamp, ts.contiguous())
        left = torch.where(ts == self.T_wc_timestamp[0], right, right - 1)
        assert ((left >= 0) & (right < len(self.T_wc_timestamp))).all()
        w = ((ts - self.T_wc_timestamp[left]) / self.bin_width[left]).to(
            self.T_wc_position.dtype)
        pos = torch.lerp(self.T_wc_position[left], self.T_wc_position[right], w[..., None])
        quat = slerp(self.T_wc_orientation_quat[left], self.T_wc_orientation_quat[right], w)
        return pos, roma.unitquat_to_rotmat(quat)


def trajectory_time_gradient(trajectory, ts, rays_d, d_rays_o, d_rays_d):
    """dL/dt of rays_o, rays_d = pixel_params_to_ray(K^-1, pix, *trajectory(ts)) in CLOSED FORM per pose
    interval — the reverse mode the CUDA kernel den_rays_from_trajectory_bwd evaluates, restated here so
    that it can be pinned against torch autograd through the reference's own trajectory code
    (models/trajectories.py:30-90, utils/tensor_ops.py:118-184, models/nerf.py:206-228;
    tests/test_oracle_vs_reference.py).  With w = (t - t_left) / width:
      position    lerp(p0, p1, w)            d o / d w = p1 - p0
      orientation R(w) = R0 Exp(w r), r = full rotation vector of q0^-1 q1 (shortest path)
                  d = R(w) c / |c|           d d / d w = (R0 r) x d
    `rays_d`: the forward directions; returns dL/dt with the shape of `ts` (per ns)."""
    T = trajectory.T_wc_timestamp
    right = torch.searchsorted(T, ts.contiguous())
    left = torch.where(ts == T[0], right, right - 1).clamp(0, len(T) - 2)
    right = right.clamp(1, len(T) - 1)
    p0, p1 = trajectory.T_wc_position[left], trajectory.T_wc_position[right]
    q0, q1 = trajectory.T_wc_orientation_quat[left], trajectory.T_wc_orientation_quat[right]
    q1 = torch.where(torch.sum(q0 * q1, dim=-1, keepdim=True) < 0, -q1, q1)
    r = _full_rotvec(roma.quat_product(roma.quat_conjugation(q0), q1))
    omega = (roma.unitquat_to_rotmat(q0) @ r[..., None])[..., 0]            # body -> world
    g = torch.sum(d_rays_o * (p1 - p0), dim=-1) + torch.sum(
        d_rays_d * torch.cross(omega, rays_d.to(omega.dtype), dim=-1), dim=-1)
    return g.double() / trajectory.bin_width[left].double()


# --------------------------------------------------------------------------- #
# event-generation parameters
# --------------------------------------------------------------------------- #
class ContrastThreshold(torch.nn.Module):
    """parameterize_mean_ct=True variant (all shipped configs)."""

    def __init__(self, pos_ct, neg_ct):
        super().__init__()
        pos_ct = torch.as_tensor(pos_ct, dtype=torch.float32)
        neg_ct = torch.as_tensor(neg_ct, dtype=torch.float32)
        self.parametrizations = torch.nn.ModuleDict({
            "p2n_contrast_threshold_ratio": _Original(softplus_inverse(pos_ct / neg_ct)),
            "mean_contrast_threshold": _Original(softplus_inverse((pos_ct + neg_ct) / 2)),
        })

    @property
    def p2n_contrast_threshold_ratio(self):
        return F.softplus(self.parametrizations["p2n_contrast_threshold_ratio"].original)

    @property
    def mean_contrast_threshold(self):
        return F.softplus(self.parametrizations["mean_contrast_threshold"].original)

    @property
    def neg_contrast_threshold(self):
        return 2 * self.mean_contrast_threshold / (self.p2n_contrast_threshold_ratio + 1)

    @property
    def pos_contrast_threshold(self):
        return self.p2n_contrast_threshold_ratio * self.neg_contrast_threshold

    def forward(self, num_pos, num_neg):
        return num_pos * self.pos_contrast_threshold - num_neg * self.neg_contrast_threshold


class RefractoryPeriod(torch.nn.Module):
    MIN_GRAD = 0.0001

    def __init__(self, refractory_period, max_refractory_period):
        super().__init__()
        tau = torch.as_tensor(refractory_period)
        tau_max = torch.as_tensor(max_refractory_period)
        if not (0 <= tau < tau_max):
            tau = 0.999 * tau_max
        self.register_buffer("max_refractory_period", tau_max, persistent=False)
        self.register_buffer("max_scaled_logit_magnitude",
                             torch.tensor(self.MIN_GRAD).logit().abs(), persistent=False)
        tau = tau.to(torch.float64)
        self.parametrizations = torch.nn.ModuleDict({"_refractory_period": _Original(
            tau_max * torch.logit(tau / tau_max))})
        self._clamp()

    @torch.no_grad()
    def _clamp(self):
        orig = self.parametrizations["_refractory_period"].original
        scaled = (orig / self.max_refractory_period).clamp(
            min=-self.max_scaled_logit_magnitude, max=self.max_scaled_logit_magnitude)
        orig.copy_(self.max_refractory_period * scaled)

    @property
    def refractory_period(self):
        self._clamp()
        orig = self.parametrizations["_refractory_period"].original
        return self.max_refractory_period * torch.sigmoid(orig / self.max_refractory_period)

    def forward(self, start_ts):
        return start_ts + self.refractory_period


# --------------------------------------------------------------------------- #
# pixel bandwidth
# --------------------------------------------------------------------------- #
def foh_discretise(A, B, dt):
    """utils/control.py foh_cont2discrete(is_state_preserved=True, is_efficient=True):
    Phi = expm(A dt); G1 = (Phi - I) A^-1 B; G2 = (A dt)^-1 G1 - A^-1 B;
    returns (Phi, Bd = G1 - G2, Btilde = G2)."""
    dt = dt[..., None, None]
    a_dt = A * dt
    phi = torch.linalg.matrix_exp(a_dt)
    a_inv_b = torch.linalg.solve(A, B)
    eye = torch.eye(A.shape[-1], dtype=A.dtype)
    g1 = (phi - eye) @ a_inv_b
    g2 = torch.linalg.solve(a_dt, g1) - a_inv_b
    return phi, g1 - g2, g2


PB_PARAMS = ("tau_mil_it_eff_prod", "A_amp_inv", "A_loop_inv", "tau_out", "tau_sf", "tau_diff")


class PixelBandwidth(torch.nn.Module):
    def __init__(self, calibration, min_ts, f_c_dominant_min, target_cumprob_max_lifetime):
        super().__init__()
        c = {k: torch.as_tensor(v) for k, v in calibration.items()
             if getattr(v, "dtype", None) is None or v.dtype.kind in "fiu"}
        self.omega_c_dominant_min = 2 * math.pi * f_c_dominant_min
        self.register_buffer("min_ts", torch.as_tensor(min_ts).detach().clone(),
                             persistent=False)
        # the reference keeps this as a float32 buffer (models/pixel_bandwidth.py:81-83)
        self.cumprob = float(torch.tensor(target_cumprob_max_lifetime, dtype=torch.float32))
        self.register_buffer("tau_in_it_eff_prod", c["input_time_const_eff_it_prod"],
                             persistent=False)
        init = {
            "tau_mil_it_eff_prod": c["miller_time_const_eff_it_prod"],
            "A_amp_inv": 1 / c["amplifier_gain"],
            "A_loop_inv": c["closed_loop_gain"] / c["amplifier_gain"],
            "tau_out": c["output_time_const"],
            "tau_sf": 1 / (2 * math.pi * c["sf_cutoff_freq"]),
            "tau_diff": 1 / (2 * math.pi * c["diff_amp_cutoff_freq"]),
        }
        self.parametrizations = torch.nn.ModuleDict(
            {k: _Original(softplus_inverse(v)) for k, v in init.items()})
        self.reset_delta_log_it = None
        self.reset_ts = None

    def param(self, name):
        return F.softplus(self.parametrizations[name].original)

    @torch.no_grad()
    def sample_lifetimes(self, interval_gen):
        """sample_intensity's lifetime part: (S-1, ...) f64 -> lifetimes (S, ...) f64 in ns
        (stop-gradient, models/pixel_bandwidth.py:298-350)."""
        S = interval_gen.shape[0] + 1
        bnd = torch.linspace(1, 0, S, dtype=interval_gen.dtype).view(
            -1, *((1,) * (interval_gen.dim() - 1)))
        gen = torch.lerp(bnd[:-1], bnd[1:], interval_gen)
        mid = torch.lerp(gen[:-1], gen[1:], 0.5)
        ones = torch.ones_like(mid[:1])
        life = torch.cat((ones, mid, torch.zeros_like(ones)), dim=0)
        rate = NS_TO_S * self.omega_c_dominant_min
        return -torch.log1p(-(self.cumprob * life)) / rate           # Exponential.icdf

    def weights(self, intensity, sample_dt, with_sf):
        """intensity_sample_to_weight: (S, ...) , (S-1, ...) ns -> (S, ..., 1 or 2)."""
        assert torch.all(sample_dt > 0)
        it = intensity[1:]
        tau_in = self.tau_in_it_eff_prod / it
        tau_mil = self.param("tau_mil_it_eff_prod") / it
        tau_out = self.param("tau_out")
        prod = (tau_in + tau_mil) * tau_out
        two_zeta_wn = (tau_in + tau_out + (1 / self.param("A_amp_inv") + 1) * tau_mil) / prod
        wn_sq = (1 / self.param("A_loop_inv") + 1) / prod
        w_sf, w_df = 1 / self.param("tau_sf"), 1 / self.param("tau_diff")
        A = torch.zeros(*it.shape, 4, 4, dtype=it.dtype)
        Bm = torch.zeros(*it.shape, 4, 1, dtype=it.dtype)
        A[..., 0, 0] = -two_zeta_wn
        A[..., 0, 1] = -wn_sq
        A[..., 1, 0] = 1
        A[..., 2, 1] = w_sf
        A[..., 2, 2] = -w_sf
        A[..., 3, 2] = w_df
        A[..., 3, 3] = -w_df
        Bm[..., 0, 0] = wn_sq
        phi, bd, bt = foh_discretise(A, Bm, NS_TO_S * sample_dt)
        C = torch.tensor([[0, 0, 1, 0], [0, 0, 0, 1]], dtype=it.dtype)
        C = C if with_sf else C[1:]
        C = C.expand(*it.shape[1:], -1, -1)
        S = it.shape[0] + 1
        w = [None] * S
        w[S - 1] = C @ bt[S - 2]
        c_next = C
        for i in range(S - 2, 0, -1):
            c_cur = c_next @ phi[i]
            w[i] = c_next @ bd[i] + c_cur @ bt[i - 1]
            c_next = c_cur
        w[0] = c_next @ bd[0]
        return torch.stack(w, dim=0).squeeze(-1)

    def combine(self, weight, intensity, last_ts, reset_diff):
        wn = weight / weight.sum(dim=0, keepdim=True)
        out = torch.sum(wn * intensity.log().unsqueeze(-1), dim=0)
        if reset_diff:
            sf, before = out[..., 0], out[..., 1]
            self.reset_delta_log_it = before - sf
            self.reset_ts = last_ts
            return sf
        before = out[..., 0]
        w_df = 1 / self.param("tau_diff")
        reset_dt = (last_ts - self.reset_ts).to(w_df.dtype)
        assert torch.all(reset_dt >= 0)
        return before - self.reset_delta_log_it * torch.exp(-w_df * (NS_TO_S * reset_dt))

    def forward(self, interval_gen, output_ts, intensity_sampling_fn, reset_diff=False):
        sample_ts = output_ts - self.sample_lifetimes(interval_gen)
        out = intensity_sampling_fn(sample_ts.clamp(min=self.min_ts))
        intensity, aux = out[0], out[1:]
        sample_dt = sample_ts.diff(dim=0).to(intensity.dtype)
        weight = self.weights(intensity, sample_dt, with_sf=reset_diff)
        return self.combine(weight, intensity, output_ts, reset_diff), aux


# --------------------------------------------------------------------------- #
# loss
# --------------------------------------------------------------------------- #
ERROR_FNS = {
    "l1": lambda a, b: (a - b).abs(),
    "mse": lambda a, b: (a - b) ** 2,
    "huber": lambda a, b: F.huber_loss(a, b, reduction="none", delta=1.0),
}


class EventLoss:
    def __init__(self, weight, error_fn, normalize):
        self.weight, self.error_fn, self.normalize = weight, error_fn, normalize

    def compute(self, log_it_diff_event, start_ts, end_ts, diff, subdiff, mean_ct):
        out = {}
        grad = log_it_diff_event / (end_ts - start_ts)
        if self.weight["log_intensity_diff"] > 0:
            k = mean_ct if self.normalize["log_intensity_diff"] else 1
            pred = diff["log_intensity_diff"]
            err = ERROR_FNS[self.error_fn["log_intensity_diff"]](
                pred / k, (diff["ts_diff"] * grad / k).to(pred.dtype))
            out["log_intensity_diff"] = err[diff["is_valid"]].mean()
        if self.weight["log_intensity_tv"] > 0:
            k = mean_ct if self.normalize["log_intensity_tv"] else 1
            pred = subdiff["log_intensity_diff"]
            err = ERROR_FNS[self.error_fn["log_intensity_tv"]](pred / k, torch.zeros_like(pred))
            out["log_intensity_tv"] = err[subdiff["is_valid"]].mean()
        return out


# --------------------------------------------------------------------------- #
# training-step host (the hot lines of models/deblur_e_nerf.py)
# --------------------------------------------------------------------------- #
def supervision_timestamps(start_ts, end_ts, normalized, use_diff, use_tv):
    """models/deblur_e_nerf.py:419-455 (all float64)."""
    diff = subdiff = None
    tv_s, tv_e = start_ts, end_ts
    if use_diff:
        ts_diff = (end_ts - start_ts) * normalized["ts_diff"]
        d_start = torch.lerp(start_ts, torch.max(end_ts - ts_diff, start_ts),
                             normalized["diff_start_ts"])
        d_end = torch.min(d_start + ts_diff, end_ts)
        diff = {"ts_diff": ts_diff, "start_ts": d_start, "end_ts": d_end}
        tv_s, tv_e = d_start, d_end
    if use_tv:
        ts_sub = (tv_e - tv_s) * normalized["ts_subdiff"]
        s_start = torch.lerp(tv_s, torch.max(tv_e - ts_sub, tv_s), normalized["subdiff_start_ts"])
        s_end = torch.min(s_start + ts_sub, tv_e)
        subdiff = {"ts_diff": ts_sub, "start_ts": s_start, "end_ts": s_end}
    return diff, subdiff


class EventRenderer(torch.nn.Module):
    """training_step (models/deblur_e_nerf.py:396-586) without Lightning: event correction,
    supervision timestamps, four render calls, loss.  Stratified jitter can be injected
    per render call (``jitters`` list) for parity runs."""

    def __init__(self, nerf, trajectory, contrast_threshold, refractory_period, pixel_bandwidth,
                 loss, intrinsics_inv, min_modeled_intensity=0.001, loss_weight=None):
        super().__init__()
        self.nerf, self.trajectory = nerf, trajectory
        self.contrast_threshold, self.refractory_period = contrast_threshold, refractory_period
        self.pixel_bandwidth = pixel_bandwidth
        self.loss = loss
        self.register_buffer("train_intrinsics_inv", intrinsics_inv, persistent=False)
        self.min_modeled_intensity = min_modeled_intensity
        self.loss_weight = loss_weight or loss.weight
        self._jitters = None

    def render_pixels(self, pixel_position, pos, rot):
        o, d = NeRF.pixel_params_to_ray(self.train_intrinsics_inv, pixel_position, pos, rot)
        jit = self._jitters.pop(0) if self._jitters else None
        rad, opa, dep, mean_samples = self.nerf(o, d, jitter=jit)
        rad = rad + self.min_modeled_intensity
        if self.nerf.render_bkgd is None:
            valid = opa > 0
        else:
            valid = torch.ones_like(opa, dtype=torch.bool)
        return rad, opa, mean_samples, valid

    @staticmethod
    def bayering(intensity, channel_idx):
        """models/deblur_e_nerf.py:1177-1178,1223-1234: the reference moves the colour axis first,
        (3, [S,] N), and gathers channel `channel_idx[n]` for pixel n; here the axis stays last."""
        idx = channel_idx.reshape((1,) * (intensity.dim() - 2) + (-1, 1)).expand(*intensity.shape[:-1], 1)
        return intensity.gather(-1, idx).squeeze(-1)

    def render_train_pixels(self, ts, pixel_position, channel_idx=None):
        pos, rot = self.trajectory(ts)
        px = pixel_position if ts.dim() == 1 else pixel_position.expand(ts.shape[0], -1, -1)
        rad, opa, mean_samples, valid = self.render_pixels(px, pos, rot)
        if channel_idx is not None:
            rad = self.bayering(rad, channel_idx)
        return rad, (opa > 0).float().mean(), mean_samples, valid

    def render_log_intensity(self, ts, pixel_position, interval_gen, reset_diff=False, channel_idx=None):
        if self.pixel_bandwidth is not None:
            fn = lambda t: self.render_train_pixels(t, pixel_position, channel_idx)     # noqa: E731
            log_it, aux = self.pixel_bandwidth(interval_gen, ts, fn, reset_diff)
            occ, mean_samples, valid = aux
            return log_it, occ, mean_samples, valid.any(dim=0)
        rad, occ, mean_samples, valid = self.render_train_pixels(ts, pixel_position, channel_idx)
        return rad.log(), occ, mean_samples, valid

    def training_step(self, event, normalized, jitters=None):
        self._jitters = list(jitters) if jitters is not None else None
        log_diff_event = self.contrast_threshold(event["num_pos"], event["num_neg"])
        start_ts = self.refractory_period(event["start_ts"])
        end_ts = event["end_ts"]
        use_diff = self.loss_weight["log_intensity_diff"] > 0
        use_tv = self.loss_weight["log_intensity_tv"] > 0
        diff, subdiff = supervision_timestamps(start_ts, end_ts, normalized, use_diff, use_tv)
        gen = normalized.get("interval_gen")
        channel_idx = event.get("channel_idx")           # :409-412, Bayer sensors only
        if channel_idx is not None:
            channel_idx = channel_idx.to(torch.int64)
        samples = []
        for seg, first in ((diff, True), (subdiff, False)):
            if seg is None:
                continue
            a, _, ms_a, va = self.render_log_intensity(seg["start_ts"], event["position"], gen,
                                                       reset_diff=first and seg is diff,
                                                       channel_idx=channel_idx)
            b, _, ms_b, vb = self.render_log_intensity(seg["end_ts"], event["position"], gen,
                                                       channel_idx=channel_idx)
            seg["log_intensity_diff"] = b - a
            seg["is_valid"] = va | vb
            samples += [ms_a, ms_b]
        terms = self.loss.compute(log_diff_event, start_ts, end_ts, diff, subdiff,
                                  self.contrast_threshold.mean_contrast_threshold)
        total = sum(v * self.loss_weight[k] for k, v in terms.items())
        return total, terms, sum(samples) / len(samples)
