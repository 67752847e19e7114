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


@pytest.mark.parametrize("mode,n,k", [(0, 64, 32), (0, 16, 64), (1, 64, 64), (1, 32, 64),
                                      (1, 64, 16), (2, 64, 128), (2, 32, 128), (2, 16, 128),
                                      (3, 64, 32), (3, 64, 64), (3, 16, 64)])
def test_tc_probe_gemm_flavours(den_lib, cuda, mode, n, k):
    """K-major / MN-major operand reuse and the M=64 accumulator lane mapping."""
    import ctypes
    g = torch.Generator().manual_seed(mode * 100 + n + k)
    bf = lambda t: t.to(torch.bfloat16).float()          # noqa: E731  (exact bf16 inputs)
    if mode in (0, 3):          # 3: the A operand lives in tensor memory (tcgen05.st + ts-form MMA)
        x, w = bf(torch.randn(128, k, generator=g)), bf(torch.randn(n, k, generator=g))
        ref = x @ w.t()
        rows = 128
    elif mode == 1:
        x, w = bf(torch.randn(128, k, generator=g)), bf(torch.randn(k, n, generator=g))
        ref = x @ w
        rows = 128
    else:
        x, w = bf(torch.randn(128, 64, generator=g)), bf(torch.randn(128, n, generator=g))
        ref = x.t() @ w
        rows = 64
    xd, wd = x.to(cuda).contiguous(), w.to(cuda).contiguous()
    out = torch.full((rows, n), float("nan"), device=cuda)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    from deblur_e_nerf_b200 import _build
    probe = ctypes.CDLL(_build.PROBE_LIB_PATH)          # test-only kernels, not part of the product .so
    rc = probe.den_tc_probe_gemm(mode, ctypes.c_void_p(xd.data_ptr()), ctypes.c_void_p(wd.data_ptr()),
                                 ctypes.c_void_p(out.data_ptr()), n, k, stream)
    assert rc == 0
    torch.cuda.synchronize()
    assert _rel(out, ref) < 1e-5, (out.cpu()[:2, :4], ref[:2, :4])


@pytest.mark.parametrize("small", [True, False], ids=["L4", "L16"])
@pytest.mark.parametrize("n", [100, 128, 3000])
def test_mlp_bwd_tc_matches_fp32_autograd(den_lib, cuda, small, n):
    """dL/denc and all ten weight / bias gradients vs torch autograd through the same layers."""
    field, desc, params, x, d, ray_idx, t0, enc = _setup(cuda, small, n, seed=3)
    g = torch.Generator().manual_seed(9)
    w_sig = (torch.randn(n, generator=g) * 0.05).to(cuda)
    w_rgb = torch.randn(n, field.radiance_dim, generator=g).to(cuda)

    # reference: same layers in torch, autograd from the encoding on
    enc_ref = enc.clone().requires_grad_(True)
    b = field.mlp_base[1]
    from deblur_e_nerf_b200 import field as field_mod
    import torch.nn.functional as F
    hid = field_mod._hidden(field.hidden_act, F.linear(enc_ref, b.hidden_layers[0].weight,
                                                       b.hidden_layers[0].bias))
    y = F.linear(hid, b.output_layer.weight, b.output_layer.bias)
    u = (x + 1.5) / 3.0
    selector = ((u > 0) & (u < 1)).all(dim=-1)
    sig_ref = field_mod._density(field.density_act, y[:, 0]) * selector
    rgb_ref = field._query_rgb(d, y[:, 1:])
    field.zero_grad()
    ((sig_ref * w_sig).sum() + (rgb_ref * w_rgb).sum()).backward()
    ref_grads = [p.grad.clone() for p in field.param_tensors()[1:]]
    ref_denc = enc_ref.grad.clone()

    field.zero_grad()
    enc_tc = enc.clone().requires_grad_(True)
    offsets = torch.arange(n + 1, dtype=torch.int32, device=cuda)
    sig, rgb = field.mlp_samples(enc_tc, x, d, ray_idx, t0, t0, offsets)
    assert _rel(sig, sig_ref) < 5e-5 and _rel(rgb, rgb_ref) < 5e-5
    ((sig * w_sig).sum() + (rgb * w_rgb).sum()).backward()
    torch.cuda.synchronize()
    assert _rel(enc_tc.grad, ref_denc) < 2e-4
    names = ("wb1", "bb1", "wb2", "bb2", "w1", "b1", "w2", "b2", "w3", "b3")
    errs = {nm: _rel(p.grad, r) for nm, p, r in zip(names, field.param_tensors()[1:], ref_grads)}
    assert max(errs.values()) < 2e-4, errs
