"""Host-side profile (cProfile) of EventRenderer.training_step at the reference-faithful batch size
(sample budget 2^17: ~1.8 k events x 30 pixel-bandwidth samples): where the Python / launch time of a
step goes when the kernels are short.

    python profiles/host_profile.py [n_events] [steps]
"""
import cProfile
import io
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from deblur_e_nerf_b200 import factory, synthetic  # noqa: E402


def main():
    n_events = int(sys.argv[1]) if len(sys.argv) > 1 else 1766
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    dev = torch.device("cuda:0")
    model, cfg, poses = factory.build_renderer("synthetic", dev, pixel_bandwidth=True)
    factory.freeze_like_synthetic_yaml(model)
    model.train()
    sphere = synthetic.solid_sphere_occupancy(128).to(dev)
    model.nerf.occupancy_grid._binary = sphere
    model.nerf.occupancy_grid.occs.copy_(sphere.reshape(-1).float())
    model.nerf.update_occ_grid = lambda *a, **k: None
    opt = factory.configure_optimizer(model)
    batches = []
    for i in range(steps + 5):
        g = torch.Generator().manual_seed(i)
        ev = synthetic.event_batch(n_events, cfg, poses[2], g)
        nm = synthetic.normalized_batch(n_events, 30, g, True)
        batches.append({"event": {k: v.to(dev) for k, v in ev.items()},
                        "normalized": {k: v.to(dev) for k, v in nm.items()}})

    def step(b, i):
        opt.zero_grad(set_to_none=True)
        loss = model.training_step(b, 0, i)
        loss.backward()
        opt.step()

    for i in range(5):
        step(batches[i], 1 + i)
    torch.cuda.synchronize()
    prof = cProfile.Profile()
    t0 = time.perf_counter()
    prof.enable()
    for i in range(steps):
        step(batches[5 + i], 6 + i)
    prof.disable()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print(f"{steps} steps: host loop {1e3 * t_host / steps:.2f} ms/step, with final sync {1e3 * t_all / steps:.2f} ms/step")
    out = io.StringIO()
    pstats.Stats(prof, stream=out).sort_stats("cumulative").print_stats(45)
    print(out.getvalue()[:9000])
    out = io.StringIO()
    pstats.Stats(prof, stream=out).sort_stats("tottime").print_stats(25)
    print(out.getvalue()[:5000])


if __name__ == "__main__":
    main()
