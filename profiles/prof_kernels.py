"""Run each hot kernel once on a synthetic.yaml-shaped field (for ncu / timing).

    python profiles/prof_kernels.py [n_samples]
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from deblur_e_nerf_b200 import factory, ops, synthetic  # noqa: E402


def main():
    n_rays = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    dev = torch.device("cuda:0")
    model, cfg, poses = factory.build_renderer("synthetic", dev, pixel_bandwidth=False)
    nerf = model.nerf
    nerf.train()
    sphere = synthetic.solid_sphere_occupancy(128).to(dev)
    nerf.occupancy_grid._binary = sphere
    g = torch.Generator().manual_seed(0)
    ev = synthetic.event_batch(n_rays, cfg, poses[2], g)
    ts = ev["end_ts"].double().to(dev)
    pos, rot = model.trajectory(ts)
    o, d = nerf.pixel_params_to_ray(model.train_intrinsics_inv, ev["position"].to(dev), pos, rot)
    o, d = o.contiguous(), d.contiguous()
    field = nerf.radiance_field
    for rep in range(2):
        ray_idx, t0, t1, offsets = nerf._march(o, d, None)
        sig, rgb0, enc = field.eval_samples_tc(o, d, ray_idx, t0, t1)
        enc_g = field.encode_samples(o, d, ray_idx, t0, t1, offsets, enc=enc)
        sigma, rgb = field.mlp_samples(enc_g, o, d, ray_idx, t0, t1, offsets, precomputed=(sig, rgb0))
        col, opa, dep = ops.composite(sigma, rgb, t0, t1, offsets, nerf.render_bkgd)
        (col.sum() + 0.1 * opa.sum()).backward()
        torch.cuda.synchronize()
    print("samples", ray_idx.numel(), "rays", n_rays)


if __name__ == "__main__":
    main()
