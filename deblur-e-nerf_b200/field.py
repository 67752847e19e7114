"""B2 host mirror of the reference's ``NGPradianceField`` (external/ngp.py:109-280) with
its ``MLP`` (external/mlp.py:26-113): same constructor argument meaning, same parameter /
buffer names (state-dict keys ``aabb``, ``mlp_base.0.params``,
``mlp_base.1.hidden_layers.0.{weight,bias}``, ``mlp_base.1.output_layer.*``,
``mlp_head.hidden_layers.{0,1}.*``, ``mlp_head.output_layer.*``), evaluated by the fused
den_b200 field kernels instead of tcnn + cuBLAS.

The module only OWNS the parameters; the kernels read the ``nn.Linear`` tensors in place
through ``den_field_params`` (no packing copy).  ``query_density`` / ``forward`` keep the
reference's call signatures so ``external/utils.py`` style closures still work.
"""

import ctypes

import torch
import torch.nn.functional as F

from . import ops
from . import tinycudann as tcnn
from ._lib import FieldDesc, FieldGrads, FieldParams
from .nerfacc import ContractionType

HIDDEN_ACT_IDS = {"relu": 0, "softplus": 1}
DENSITY_ACT_IDS = {"shifted_trunc_exp": 0, "softplus": 1, "shifted_softplus": 2}
RADIANCE_ACT_IDS = {"softplus": 0, "sigmoid": 1}


def _act_name(value, table):
    """Accept the YAML names or the callables the reference's NeRF passes
    (models/nerf.py:17-29: nn.Softplus(beta=100), nn.ReLU(), ngp.shifted_trunc_exp, ...)."""
    if isinstance(value, str):
        if value not in table:
            raise NotImplementedError(f"activation {value!r}")
        return value
    if isinstance(value, torch.nn.ReLU):
        return "relu"
    if isinstance(value, torch.nn.Sigmoid):
        return "sigmoid"
    if isinstance(value, torch.nn.Softplus):
        return "softplus"
    name = getattr(value, "__name__", "")
    if name in table:
        return name
    raise NotImplementedError(f"activation {value!r}")


class _TruncExp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(torch.clamp(x, max=15))


def _hidden(name, v):
    return F.softplus(v, beta=100) if name == "softplus" else torch.relu(v)


def _density(name, v):
    if name == "shifted_trunc_exp":
        return _TruncExp.apply(v - 1)
    if name == "softplus":
        return F.softplus(v)
    return F.softplus(v - 1)


def _radiance(name, v):
    return F.softplus(v) if name == "softplus" else torch.sigmoid(v)


class MLP(torch.nn.Module):
    """Parameter container with the reference's names; default nn.Linear init because the
    reference passes every ``*_init=None`` (external/ngp.py:179-185)."""

    def __init__(self, input_dim, output_dim, net_depth, net_width):
        super().__init__()
        self.hidden_layers = torch.nn.ModuleList()
        fan_in = input_dim
        for _ in range(net_depth):
            self.hidden_layers.append(torch.nn.Linear(fan_in, net_width))
            fan_in = net_width
        self.output_layer = torch.nn.Linear(fan_in, output_dim)


def sh_degree4(d):
    x, y, z = d.unbind(-1)
    xy, xz, yz = x * y, x * z, y * z
    x2, y2, z2 = x * x, y * y, z * z
    return torch.stack([
        torch.full_like(x, 0.28209479177387814),
        -0.48860251190291987 * y, 0.48860251190291987 * z, -0.48860251190291987 * x,
        1.0925484305920792 * xy, -1.0925484305920792 * yz,
        0.94617469575755997 * z2 - 0.31539156525251999, -1.0925484305920792 * xz,
        0.54627421529603959 * x2 - 0.54627421529603959 * y2,
        0.59004358992664352 * y * (-3.0 * x2 + y2), 2.8906114426405538 * xy * z,
        0.45704579946446572 * y * (1.0 - 5.0 * z2), 0.3731763325901154 * z * (5.0 * z2 - 3.0),
        0.45704579946446572 * x * (1.0 - 5.0 * z2), 1.4453057213202769 * z * (x2 - y2),
        0.59004358992664352 * x * (-x2 + 3.0 * y2)], dim=-1)


class _MlpTcFn(torch.autograd.Function):
    """sigma, rgb = MLP(enc, dirs) on the tensor cores (den_mlp_fwd); backward recomputes the
    forward per tile and returns dL/denc plus every weight / bias gradient (den_mlp_bwd)."""

    @staticmethod
    def forward(ctx, field, enc, rays_o, rays_d, ray_indices, t_starts, t_ends, offsets,
                precomputed, n_dev, enc_rows, *weights):
        if precomputed is not None:
            # the visibility pre-pass already evaluated these very samples with these weights
            sig, rgb = (t.view_as(t) for t in precomputed)
        else:
            assert enc_rows is None
            sig, rgb = ops.mlp_fwd(field.field_desc(), field.field_params(), enc, rays_o, rays_d,
                                   ray_indices, t_starts, t_ends, field.radiance_dim, n_dev)
        ctx.field = field
        ctx.n_dev, ctx.enc_rows = n_dev, enc_rows
        ctx.save_for_backward(enc, rays_o, rays_d, ray_indices, t_starts, t_ends, offsets)
        return sig, rgb

    @staticmethod
    def backward(ctx, d_sig, d_rgb):
        field = ctx.field
        enc, rays_o, rays_d, ray_indices, t_starts, t_ends, offsets = ctx.saved_tensors
        need_d_dirs = ctx.needs_input_grad[3]          # only when the rays carry a gradient (tau)
        weights = field.param_tensors()[1:]
        grads = [torch.zeros_like(w) for w in weights]
        gs = FieldGrads()
        for name, t in zip(("wb1", "bb1", "wb2", "bb2", "w1", "b1", "w2", "b2", "w3", "b3"), grads):
            setattr(gs, name, t.data_ptr())
        d_enc, d_dirs = ops.mlp_bwd(field.field_desc(), field.field_params(), gs, enc, rays_o,
                                    rays_d, ray_indices, t_starts, t_ends, d_sig.contiguous(),
                                    d_rgb.contiguous(), need_d_dirs, ctx.n_dev, ctx.enc_rows)
        d_rays_d = ops.segment_sum(d_dirs, offsets) if need_d_dirs else None
        return (None, d_enc, None, d_rays_d, None, None, None, None, None, None, None, *grads)


class NGPradianceField(torch.nn.Module):
    def __init__(self, aabb, num_dim=3, use_viewdirs=True, contraction_type=ContractionType.AABB,
                 pos_encoding_config=None, dir_encoding_config=None, mlp_base_config=None,
                 mlp_head_config=None):
        super().__init__()
        if num_dim != 3 or not use_viewdirs:
            raise NotImplementedError("den_b200 field: 3-D positions with view directions only")
        if not isinstance(aabb, torch.Tensor):
            aabb = torch.tensor(aabb, dtype=torch.float32)
        self.register_buffer("aabb", aabb)
        self._aabb_host = [float(v) for v in aabb.tolist()]
        self.num_dim = num_dim
        self.use_viewdirs = True
        self.contraction_type = contraction_type
        base, head = dict(mlp_base_config), dict(mlp_head_config)
        self.radiance_dim = int(head["output_dim"])
        self.geo_feat_dim = int(base["geo_feat_dim"])
        self.hidden_act = _act_name(base["hidden_activation"], HIDDEN_ACT_IDS)
        if _act_name(head["hidden_activation"], HIDDEN_ACT_IDS) != self.hidden_act:
            raise NotImplementedError("base and head hidden activations must match")
        self.density_act = _act_name(base["density_activation"], DENSITY_ACT_IDS)
        self.radiance_act = _act_name(head["radiance_activation"], RADIANCE_ACT_IDS)
        if base.get("weight_norm") or head.get("weight_norm"):
            raise NotImplementedError("weight_norm is not used by any shipped config")
        self.sh_degree = int(dict(dir_encoding_config)["degree"])
        self.width = int(base["n_neurons"])
        self.n_hidden_base = int(base["n_hidden_layers"])
        self.n_hidden_head = int(head["n_hidden_layers"])
        if (self.sh_degree != 4 or self.width != 64 or self.geo_feat_dim != 15
                or self.n_hidden_base != 1 or self.n_hidden_head != 2
                or int(head["n_neurons"]) != 64):
            raise NotImplementedError(
                "den_b200 builds the architecture of the shipped configs only: base 1x64 -> 1+15, "
                "SH degree 4, head 2x64 (configs/train/*.yaml:81-103)")
        encoding = tcnn.Encoding(3, pos_encoding_config, dtype=torch.float32)
        self.mlp_base = torch.nn.Sequential(
            encoding, MLP(encoding.n_output_dims, 1 + self.geo_feat_dim, 1, 64))
        self.mlp_head = MLP(16 + self.geo_feat_dim, self.radiance_dim, 2, 64)

    # ------------------------------------------------------------------ C-ABI views --
    @property
    def encoding(self):
        return self.mlp_base[0]

    def field_desc(self):
        d = FieldDesc()
        ctypes.memmove(ctypes.byref(d.grid), ctypes.byref(self.encoding.desc),
                       ctypes.sizeof(d.grid))
        for i in range(6):
            d.aabb[i] = self._aabb_host[i]
        d.contraction = self.contraction_type.to_cpp_version()
        d.channels = self.radiance_dim
        d.hidden_act = HIDDEN_ACT_IDS[self.hidden_act]
        d.density_act = DENSITY_ACT_IDS[self.density_act]
        d.radiance_act = RADIANCE_ACT_IDS[self.radiance_act]
        d.width, d.geo_feat_dim, d.sh_degree = self.width, self.geo_feat_dim, self.sh_degree
        d.n_hidden_base, d.n_hidden_head = self.n_hidden_base, self.n_hidden_head
        return d

    def param_tensors(self):
        b, h = self.mlp_base[1], self.mlp_head
        return [self.encoding.params,
                b.hidden_layers[0].weight, b.hidden_layers[0].bias,
                b.output_layer.weight, b.output_layer.bias,
                h.hidden_layers[0].weight, h.hidden_layers[0].bias,
                h.hidden_layers[1].weight, h.hidden_layers[1].bias,
                h.output_layer.weight, h.output_layer.bias]

    def field_params(self):
        p = FieldParams()
        names = ("table", "wb1", "bb1", "wb2", "bb2", "w1", "b1", "w2", "b2", "w3", "b3")
        for name, t in zip(names, self.param_tensors()):
            if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
                raise NotImplementedError("field parameters must be contiguous fp32 CUDA tensors")
            setattr(p, name, t.data_ptr())
        return p

    # ------------------------------------------------------------ fused (no autograd) --
    @torch.no_grad()
    def density_at(self, x):
        """sigma (n,1) at world positions (n,3): den_field_density_at."""
        x = x.reshape(-1, 3).float().contiguous()
        return ops.field_density_at(self.field_desc(), self.field_params(), x)[:, None]

    @torch.no_grad()
    def eval_samples(self, rays_o, rays_d, ray_indices, t_starts, t_ends, full=True, n_dev=None):
        """(sigma (M,), rgb (M,C) | None) for marched samples: den_field_fwd."""
        return ops.field_fwd(self.field_desc(), self.field_params(), rays_o, rays_d, ray_indices,
                             t_starts, t_ends, self.radiance_dim if full else 0, n_dev)

    # ---------------------------------------------- tensor-core path (with autograd) --
    def encode_samples(self, rays_o, rays_d, ray_indices, t_starts, t_ends, offsets, enc=None,
                       n_dev=None):
        """Hash-grid encoding (M, L*2) of marched samples as an autograd node on the table (and on
        the rays when they require grad); `enc` re-uses an encoding already computed on the same
        samples.  `n_dev`: device-side sample count (the tensors are capacity-sized)."""
        if rays_o.requires_grad or rays_d.requires_grad:
            u = ops.contract_samples_autograd(self.field_desc(), rays_o, rays_d, ray_indices,
                                              t_starts, t_ends, offsets, n_dev)
        else:
            u = ops.contract_samples(self.field_desc(), rays_o, rays_d, ray_indices, t_starts,
                                     t_ends, n_dev)
        if enc is not None:
            return ops.hashgrid_reuse(u, self.encoding.params, self.encoding.desc, enc, n_dev)
        assert n_dev is None
        return ops.hashgrid(u, self.encoding.params, self.encoding.desc)

    def mlp_samples(self, enc, rays_o, rays_d, ray_indices, t_starts, t_ends, offsets,
                    precomputed=None, n_dev=None, enc_rows=None):
        """(sigma (M,), rgb (M,C)) from encodings on the tensor cores, differentiable in the
        encodings and in every MLP parameter.  `precomputed=(sigma, rgb)` skips the forward
        launch when the pre-pass has already produced them (the backward recomputes anyway).
        `enc_rows` (M) int32: sample i's encoding is row enc_rows[i] of `enc` (survivors of the
        visibility filter reading the pre-pass encodings in place)."""
        return _MlpTcFn.apply(self, enc, rays_o, rays_d, ray_indices, t_starts, t_ends, offsets,
                              precomputed, n_dev, enc_rows, *self.param_tensors()[1:])

    @torch.no_grad()
    def eval_samples_tc(self, rays_o, rays_d, ray_indices, t_starts, t_ends, full=True, n_dev=None):
        """Gather + tensor-core MLP without autograd: (sigma (M,), rgb (M,C) | None, enc (M, L*2)).
        This is the reference's no-grad density pre-pass (external/utils.py:68-81); evaluating
        the colour head in the same launch lets the grad pass re-use every output."""
        u = ops.contract_samples(self.field_desc(), rays_o, rays_d, ray_indices, t_starts, t_ends,
                                 n_dev)
        enc = ops.hashgrid_fwd(self.encoding.desc, u, self.encoding.params, n_dev)
        sig, rgb = ops.mlp_fwd(self.field_desc(), self.field_params(), enc, rays_o, rays_d,
                               ray_indices, t_starts, t_ends, self.radiance_dim if full else 0, n_dev)
        return sig, rgb, enc

    # --------------------------------------------------- reference-signature methods --
    def _contract(self, x):
        lo, hi = self.aabb[:3], self.aabb[3:]
        u = (x - lo) / (hi - lo)
        if self.contraction_type == ContractionType.UN_BOUNDED_SPHERE:
            u = u * 2 - 1
            mag = u.norm(dim=-1, keepdim=True)
            u = torch.where(mag > 1, (2 - 1 / mag) * (u / mag), u)
            u = u / 4 + 0.5
        elif self.contraction_type == ContractionType.UN_BOUNDED_TANH:
            u = (torch.tanh(u - 0.5) + 1) / 2
        return u

    def query_density(self, x, return_feat=False):
        """external/ngp.py:230-254.  Without autograd (and without `return_feat`) this is one
        fused kernel; with autograd it is the hash-grid kernel pair + the parameter tensors
        through torch (the gradient path of callers that still use the operator API)."""
        if not (torch.is_grad_enabled() and self._needs_grad(x)) and not return_feat:
            return self.density_at(x).reshape(*x.shape[:-1], 1)
        u = self._contract(x)
        selector = ((u > 0.0) & (u < 1.0)).all(dim=-1)
        b = self.mlp_base[1]
        enc = self.encoding(u.reshape(-1, 3))
        hid = _hidden(self.hidden_act, F.linear(enc, b.hidden_layers[0].weight,
                                                b.hidden_layers[0].bias))
        y = F.linear(hid, b.output_layer.weight, b.output_layer.bias)
        y = y.reshape(*x.shape[:-1], 1 + self.geo_feat_dim)
        density = _density(self.density_act, y[..., :1]) * selector[..., None]
        return (density, y[..., 1:]) if return_feat else density

    def _needs_grad(self, x):
        return x.requires_grad or any(p.requires_grad for p in self.parameters())

    def _query_rgb(self, dir, embedding):
        h = self.mlp_head
        z = torch.cat([sh_degree4(dir.reshape(-1, 3)),
                       embedding.reshape(-1, self.geo_feat_dim)], dim=-1)
        for layer in h.hidden_layers:
            z = _hidden(self.hidden_act, F.linear(z, layer.weight, layer.bias))
        z = _radiance(self.radiance_act, F.linear(z, h.output_layer.weight, h.output_layer.bias))
        return z.reshape(*embedding.shape[:-1], self.radiance_dim)

    def forward(self, positions, directions=None):
        """external/ngp.py:269-280: (rgb (..., C), density (..., 1)).  On CUDA tensors this is the
        hash-grid kernel pair + the tensor-core MLP pair (each position is a "ray" with t = 0), so callers
        of the reference-signature method stay on the fused path, with autograd to the table, the MLP
        parameters and the positions; the torch evaluation below remains for CPU tensors."""
        if directions is not None:
            assert positions.shape == directions.shape, \
                f"{positions.shape} v.s. {directions.shape}"
        if positions.is_cuda and directions is not None and positions.numel() > 0:
            x = positions.reshape(-1, 3).float().contiguous()
            d = directions.reshape(-1, 3).float().contiguous()
            n = x.shape[0]
            ray_idx = torch.arange(n, dtype=torch.int32, device=x.device)
            t0 = torch.zeros(n, dtype=torch.float32, device=x.device)
            offsets = torch.arange(n + 1, dtype=torch.int32, device=x.device)
            enc = self.encode_samples(x, d, ray_idx, t0, t0, offsets)
            sig, rgb = self.mlp_samples(enc, x, d, ray_idx, t0, t0, offsets)
            return (rgb.reshape(*positions.shape[:-1], self.radiance_dim),
                    sig.reshape(*positions.shape[:-1], 1))
        density, embedding = self.query_density(positions, return_feat=True)
        return self._query_rgb(directions, embedding), density
