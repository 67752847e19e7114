"""Where does the device batch producer's time go?  (profiling aid, not part of the product)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
entry.build()
from deblur_e_nerf_b200 import factory, synthetic, ddp
from deblur_e_nerf_b200.data import EventBatchProducer

dev = torch.device("cuda:0")
model, cfg, poses = factory.build_renderer("synthetic", dev, pixel_bandwidth=True, occ_resolution=128, seed=0)
factory.freeze_like_synthetic_yaml(model)
model.train()
sphere = synthetic.solid_sphere_occupancy(128).to(dev)
model.nerf.occupancy_grid._binary = sphere
model.nerf.occupancy_grid.occs.copy_(sphere.reshape(-1).float())
model.nerf.update_occ_grid = lambda *a, **k: None
reducer = ddp.GradReducer(model)
opt = factory.configure_optimizer(model)
n = (1 << 17) // 30
g = torch.Generator().manual_seed(3)
pool = synthetic.event_batch(1 << 21, cfg, poses[2], g)
prod = EventBatchProducer(pool, n, it_sample_size=30, device=dev, seed=1)
hb = {"event": {k: v.to(dev) for k, v in synthetic.event_batch(n, cfg, poses[2], g).items()},
      "normalized": {k: v.to(dev) for k, v in synthetic.normalized_batch(n, 30, g, True).items()}}

def step(b):
    opt.zero_grad(set_to_none=False)
    loss = model.training_step(b, 0, 1)
    loss.backward(); reducer(); opt.step()

def timeit(fn, k=8):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(k): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / k * 1e3

for _ in range(3): step(hb)
print("next_batch only        %.3f ms" % timeit(prod.next_batch, 20))
print("step, pre-staged batch %.3f ms" % timeit(lambda: step(hb)))
pb = prod.next_batch()
print("step, one producer batch reused %.3f ms" % timeit(lambda: step(pb)))
print("step, fresh producer batch      %.3f ms" % timeit(lambda: step(prod.next_batch())))
for k, v in pb["event"].items(): print(k, v.dtype, v.shape, v.is_contiguous(), v.float().mean().item())
for k, v in hb["event"].items(): print(k, v.dtype, v.shape, v.is_contiguous(), v.float().mean().item())
print("samples/ray", float(model.logged["train/mean_num_samples_per_ray"]))
