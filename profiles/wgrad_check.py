import os, sys, torch
sys.path.insert(0, '/root/repo')
import torch.nn.functional as F
from deblur_e_nerf_b200 import factory, ops, synthetic, field as field_mod
from deblur_e_nerf_b200._lib import FieldGrads
n_rays = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
dev = torch.device("cuda:0")
model, cfg, poses = factory.build_renderer("synthetic", dev, pixel_bandwidth=False)
nerf = model.nerf; nerf.train()
nerf.occupancy_grid._binary = synthetic.solid_sphere_occupancy(128).to(dev)
g = torch.Generator().manual_seed(0)
ev = synthetic.event_batch(n_rays, cfg, poses[2], g)
o, d = model.rays(ev["end_ts"].double().to(dev), ev["position"].to(dev))
o, d = o.contiguous(), d.contiguous()
field = nerf.radiance_field
ray_idx, t0, t1, offsets = nerf._march(o, d, None)
sig, rgb, enc = field.eval_samples_tc(o, d, ray_idx, t0, t1)
n = ray_idx.numel()
gg = torch.Generator(device=dev).manual_seed(1)
# realistic upstream gradients: small, mixed sign, ray-correlated
d_sig = (torch.randn(n, device=dev, generator=gg) * 1e-6)
d_rgb = (torch.randn(n, 1, device=dev, generator=gg) * 1e-5)
desc, params = field.field_desc(), field.field_params()
names = ("wb1", "bb1", "wb2", "bb2", "w1", "b1", "w2", "b2", "w3", "b3")
def run(chunks):
    gs, keep = FieldGrads(), []
    for name, w in zip(names, field.param_tensors()[1:]):
        keep.append(torch.zeros_like(w)); setattr(gs, name, keep[-1].data_ptr())
    bounds = [n * k // chunks for k in range(chunks + 1)]
    for a, b in zip(bounds[:-1], bounds[1:]):
        ops.mlp_bwd(desc, params, gs, enc[a:b].contiguous(), o, d, ray_idx[a:b].contiguous(), t0[a:b].contiguous(),
                    t1[a:b].contiguous(), d_sig[a:b].contiguous(), d_rgb[a:b].contiguous())
    torch.cuda.synchronize()
    return keep
one = run(1); four = run(4); many = run(64)
# fp64 reference through torch autograd, chunked
ws = [w.detach().double().requires_grad_(True) for w in field.param_tensors()[1:]]
tm = 0.5 * (t0 + t1)
tot = None
for a in range(0, n, 1 << 20):
    b = min(n, a + (1 << 20))
    e = enc[a:b].double()
    pos = o[ray_idx[a:b].long()] + d[ray_idx[a:b].long()] * tm[a:b, None]
    u = (pos - (-1.5)) / 3.0
    sel = ((u > 0) & (u < 1)).all(dim=-1)
    hb = F.softplus(F.linear(e, ws[0], ws[1]), beta=100)
    y = F.linear(hb, ws[2], ws[3])
    sg = torch.exp(y[:, 0] - 1) * sel
    z = torch.cat([field_mod.sh_degree4(d[ray_idx[a:b].long()]).double(), y[:, 1:]], dim=-1)
    h1 = F.softplus(F.linear(z, ws[4], ws[5]), beta=100)
    h2 = F.softplus(F.linear(h1, ws[6], ws[7]), beta=100)
    out = F.softplus(F.linear(h2, ws[8], ws[9]))
    loss = (sg * d_sig[a:b].double()).sum() + (out * d_rgb[a:b].double()).sum()
    loss.backward()
ref = [w.grad for w in ws]
def rel(a, b): return ((a.double() - b).abs().max() / b.abs().max()).item()
print("samples", n)
for nm, r, a, b, c in zip(names, ref, one, four, many):
    print(f"{nm:4s} max|ref| {r.abs().max().item():.3e}  1 launch {rel(a, r):.2e}  4 launches {rel(b, r):.2e}  64 launches {rel(c, r):.2e}")
