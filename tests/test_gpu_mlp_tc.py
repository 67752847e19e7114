"""Numerics of the tcgen05 tensor-core MLP kernels against the plain fp32 PyTorch evaluation of
the same layers (the reference's cuBLAS-with-TF32-off path), through the C ABI."""

import pytest
import torch

import _scene

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-30)).item()


def _setup(cuda, small, n, seed=0):
    from deblur_e_nerf_b200 import ops
    cfg = _scene.scene_config("synthetic", occ_resolution=32, small=small)
    nerf = _scene.build_product_nerf(cfg, cuda, seed)
    _scene.randomize_field_(nerf, seed, table_scale=0.5, density_bias=1.0)
    field = nerf.radiance_field
    g = torch.Generator().manual_seed(seed + 1)
    x = (torch.rand(n, 3, generator=g) * 3.2 - 1.6).to(cuda)       # some outside the AABB
    d = torch.randn(n, 3, generator=g)
    d = (d / d.norm(dim=-1, keepdim=True)).to(cuda)
    ray_idx = torch.arange(n, dtype=torch.int32, device=cuda)
    t0 = torch.zeros(n, device=cuda)
    desc, params = field.field_desc(), field.field_params()
    u = ops.contract_samples(desc, x, d, ray_idx, t0, t0)
    enc = ops.hashgrid_fwd(field.encoding.desc, u, field.encoding.params)
    return field, desc, params, x, d, ray_idx, t0, enc


@pytest.mark.parametrize("small", [True, False], ids=["L4", "L16"])
@pytest.mark.parametrize("n", [1, 127, 128, 5000])
def test_mlp_fwd_tc_matches_fp32(den_lib, cuda, small, n):
    from deblur_e_nerf_b200 import ops
    field, desc, params, x, d, ray_idx, t0, enc = _setup(cuda, small, n)
    with torch.no_grad():
        rgb_ref, sig_ref = field(x, d)                 # fp32 torch evaluation of the same layers
    sig, rgb = ops.mlp_fwd(desc, params, enc, x, d, ray_idx, t0, t0, field.radiance_dim)
    torch.cuda.synchronize()
    assert _rel(sig, sig_ref.reshape(-1)) < 5e-5
    assert _rel(rgb, rgb_ref) < 5e-5
    sig2, none = ops.mlp_fwd(desc, params, enc, x, d, ray_idx, t0, t0, 0)
    assert none is None and _rel(sig2, sig_ref.reshape(-1)) < 5e-5
    # density is gated by the selector: samples outside the unit cube have sigma == 0
    outside = ((x < -1.5) | (x > 1.5)).any(dim=-1)
    assert outside.any() or n < 10
    assert torch.all(sig[outside] == 0)
