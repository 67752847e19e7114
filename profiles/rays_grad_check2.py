import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import _scene
cuda = torch.device('cuda:0')
golden = _scene.load_golden("training_step_pb_on")
cfg = _scene.scene_config("synthetic", occ_resolution=32, small=True)
model, poses = _scene.build_product_renderer(cfg, cuda, 8, pixel_bandwidth=True)
for name in ["nerf", "contrast_threshold", "refractory_period", "pixel_bandwidth"]:
    _scene.load_golden_state(getattr(model, name), golden, name, cuda)
batch = {"event": _scene.golden_section(golden, "event", cuda), "normalized": _scene.golden_section(golden, "normalized", cuda)}
jitters = [v for _, v in sorted(_scene.golden_section(golden, "jitter", cuda).items(), key=lambda kv: int(kv[0]))]
model.train()
model.nerf.update_occ_grid = lambda *a, **k: None
orig = model.rays
rec = []
def rays(ts, pix):
    o, d = orig(ts, pix)
    if ts.requires_grad:
        item = {"ts": ts, "pix": pix}
        o.register_hook(lambda g: item.__setitem__("g_o", g.clone()))
        d.register_hook(lambda g: item.__setitem__("g_d", g.clone()))
        ts.register_hook(lambda g: item.__setitem__("g_ts", g.clone()))
        rec.append(item)
    return o, d
model.rays = rays
loss = model.training_step(batch, 0, 0, jitters=jitters)
loss.backward()
print("calls with grad:", len(rec))
for it in rec:
    ts = it["ts"].detach().clone().requires_grad_(True)
    pos, rot = model.trajectory(ts)
    o, d = model.nerf.pixel_params_to_ray(model.train_intrinsics_inv, it["pix"], pos, rot)
    (g_ref,) = torch.autograd.grad((o, d), ts, (it["g_o"], it["g_d"]))
    o2, d2 = orig(ts, it["pix"])
    (g_k,) = torch.autograd.grad((o2, d2), ts, (it["g_o"], it["g_d"]))
    print("shape", tuple(ts.shape), "g_ts(total incl. other uses) sum", it["g_ts"].sum().item(), "ray part: torch sum", g_ref.sum().item(), "kernel sum", g_k.sum().item(),
          "max|diff|/max", ((g_ref - g_k).abs().max() / g_ref.abs().max()).item(),
          "t range", ts.min().item(), ts.max().item(), "pose range", float(model.trajectory.T_wc_timestamp[0]), float(model.trajectory.T_wc_timestamp[-1]))
    tau = [p for n, p in model.named_parameters() if "_refractory_period" in n][0]
    # fp64 truth of the ray part
    from deblur_e_nerf_b200 import trajectories
    tr = model.trajectory
    tr64 = trajectories.LinearTrajectory((tr.T_wc_position.double(), tr.T_wc_orientation_quat.double(), tr.T_wc_timestamp))
    ts64 = it["ts"].detach().clone().requires_grad_(True)
    pos, rot = tr64(ts64)
    o, d = model.nerf.pixel_params_to_ray(model.train_intrinsics_inv.double(), it["pix"].double(), pos, rot)
    (g_64,) = torch.autograd.grad((o, d), ts64, (it["g_o"].double(), it["g_d"].double()))
    print("sum|g|", g_k.abs().sum().item(), "max|g|", g_k.abs().max().item(), "tau grad", tau.grad, "sum fp64", g_64.sum().item(),
          "per-ray err vs fp64 / max: torch", ((g_ref - g_64).abs().max() / g_64.abs().max()).item(), "kernel", ((g_k - g_64).abs().max() / g_64.abs().max()).item())
    bad = (g_ref - g_k).abs() > 1e-3 * g_ref.abs().max()
    if bad.any():
        idx = bad.nonzero()[:5]
        for j in idx:
            j = tuple(j.tolist())
            print("  bad", j, ts[j].item(), g_ref[j].item(), g_k[j].item())
