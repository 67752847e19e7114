"""Time den_hashgrid_fwd / den_hashgrid_bwd alone on a synthetic.yaml-shaped sample set, optionally with a
different run-merging threshold (levels with resolution <= agg are merged in the scatter).

    python profiles/time_hashgrid.py [n_rays] [agg ...]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from deblur_e_nerf_b200 import factory, ops, synthetic  # noqa: E402


def main():
    n_rays = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
    aggs = [int(a) for a in sys.argv[2:]] or [512]
    dev = torch.device("cuda:0")
    model, cfg, poses = factory.build_renderer("synthetic", dev, pixel_bandwidth=False)
    nerf = model.nerf
    nerf.train()
    nerf.occupancy_grid._binary = synthetic.solid_sphere_occupancy(128).to(dev)
    g = torch.Generator().manual_seed(0)
    ev = synthetic.event_batch(n_rays, cfg, poses[2], g)
    pos, rot = model.trajectory(ev["end_ts"].double().to(dev))
    o, d = nerf.pixel_params_to_ray(model.train_intrinsics_inv, ev["position"].to(dev), pos, rot)
    o, d = o.contiguous(), d.contiguous()
    field = nerf.radiance_field
    ray_idx, t0, t1, offsets = nerf._march(o, d, None)
    u = ops.contract_samples(field.field_desc(), o, d, ray_idx, t0, t1, None)
    table = field.encoding.params.detach().clone().requires_grad_(True)
    n = ray_idx.numel()
    gout = torch.randn(n, 32, device=dev)
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    print("samples", n)
    enc_cfg = field.encoding
    for agg in aggs:
        desc, _ = ops.make_hashgrid_desc(enc_cfg.n_levels, enc_cfg.base_resolution, enc_cfg.per_level_scale,
                                         enc_cfg.log2_hashmap_size, agg_max_resolution=agg)
        for rep in range(3):
            enc = ops.hashgrid(u, table, desc)
            torch.cuda.synchronize()
            start.record()
            enc.backward(gout)
            end.record()
            torch.cuda.synchronize()
        print(f"agg_max_resolution {agg} (levels merged: {desc.n_agg_levels}): hashgrid_bwd {start.elapsed_time(end):.3f} ms (incl. autograd glue)")


if __name__ == "__main__":
    main()
