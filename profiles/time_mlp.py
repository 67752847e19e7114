"""Time den_mlp_fwd / den_mlp_bwd alone (CUDA events) on a synthetic.yaml-shaped sample set.

    python profiles/time_mlp.py [n_rays]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from deblur_e_nerf_b200 import factory, ops, synthetic  # noqa: E402


def main():
    n_rays = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
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
    sig, rgb, enc = field.eval_samples_tc(o, d, ray_idx, t0, t1)
    n = ray_idx.numel()
    desc, params = field.field_desc(), field.field_params()
    from deblur_e_nerf_b200._lib import FieldGrads
    grads, keep = FieldGrads(), []
    for name, w in zip(("wb1", "bb1", "wb2", "bb2", "w1", "b1", "w2", "b2", "w3", "b3"),
                       field.param_tensors()[1:]):
        keep.append(torch.zeros_like(w))
        setattr(grads, name, keep[-1].data_ptr())
    d_sig = torch.randn(n, device=dev) * 0.01
    d_rgb = torch.randn(n, field.radiance_dim, device=dev)
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timeit(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        start.record()
        for _ in range(reps):
            fn()
        end.record()
        torch.cuda.synchronize()
        return start.elapsed_time(end) / reps

    print("samples", n)
    print("mlp_fwd full ms", timeit(lambda: ops.mlp_fwd(desc, params, enc, o, d, ray_idx, t0, t1, 1)))
    print("mlp_fwd density ms", timeit(lambda: ops.mlp_fwd(desc, params, enc, o, d, ray_idx, t0, t1, 0)))
    ms = timeit(lambda: ops.mlp_bwd(desc, params, grads, enc, o, d, ray_idx, t0, t1, d_sig, d_rgb))
    print(f"mlp_bwd ms {ms:.3f}")


if __name__ == "__main__":
    main()
