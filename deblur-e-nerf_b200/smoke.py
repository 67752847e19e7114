"""``__graft_entry__.smoke()``: one small training step of the hot path on cuda:0 (march ->
fused field -> compositing -> [pixel-bandwidth filter] -> event loss -> backward), checked
against the CPU oracle on the same parameters, batch and stratified jitter."""

import torch


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-30)).item()


def run(n_events=48, it_sample_size=8, pixel_bandwidth=None, tol=1e-3, verbose=True):
    from oracle import nerfacc_ref, path_ref            # the checker (test infrastructure)
    from . import factory, ops, synthetic

    dev = torch.device("cuda:0")
    if pixel_bandwidth is None:
        try:
            from . import pixel_bandwidth as _pb        # noqa: F401
            pixel_bandwidth = True
        except ImportError:
            pixel_bandwidth = False
    launches0 = ops.launch_count()
    model, cfg, poses = factory.build_renderer("synthetic", dev, pixel_bandwidth=pixel_bandwidth,
                                               small=True, occ_resolution=32, n_poses=200)
    with torch.no_grad():
        p = model.nerf.radiance_field.encoding.params
        p.copy_((torch.rand(p.shape, device=dev) * 2 - 1) * 0.5)
        model.nerf.radiance_field.mlp_base[1].output_layer.bias[0] += 2.5
    model.train()

    # the oracle twin with identical parameters
    occ = dict(resolution=32, occ_thre=1e-2, ema_decay=0.95, warmup_steps=256, n=16)
    nerf = path_ref.NeRF(cfg["aabb"], nerfacc_ref.ContractionType.AABB, occ, cfg["near_plane"],
                         cfg["far_plane"], synthetic.render_step_size(cfg["aabb"]),
                         cfg["render_bkgd"], cfg["cone_angle"], cfg["early_stop_eps"],
                         cfg["alpha_thre"], cfg["test_chunk_size"],
                         synthetic.arch_config(small=True), 1)
    calib = synthetic.calibration()
    pb = path_ref.PixelBandwidth(calib, poses[2].min(), 21, 0.95) if pixel_bandwidth else None
    weight = dict(log_intensity_diff=1.0, log_intensity_tv=cfg["tv_weight"])
    oracle = path_ref.EventRenderer(
        nerf, path_ref.LinearTrajectory(*poses),
        path_ref.ContrastThreshold(calib["pos_contrast_threshold"],
                                   calib["neg_contrast_threshold"]),
        path_ref.RefractoryPeriod(calib["refractory_period"], synthetic.MAX_REFRACTORY_PERIOD_NS),
        pb, path_ref.EventLoss(weight, dict(log_intensity_diff="huber", log_intensity_tv="l1"),
                               dict(log_intensity_diff=True, log_intensity_tv=True)),
        torch.linalg.inv(torch.from_numpy(synthetic.intrinsics(cfg))))
    names = ["nerf", "contrast_threshold", "refractory_period"] + (
        ["pixel_bandwidth"] if pixel_bandwidth else [])
    for name in names:
        src = {k: v.cpu() for k, v in getattr(model, name).state_dict().items()}
        dst = getattr(oracle, name)
        dst.load_state_dict({k: src[k] for k in dst.state_dict()}, strict=True)
    oracle.train()
    torch.manual_seed(0)
    oracle.nerf.update_occ_grid(0, poses[0])
    model.nerf.occupancy_grid._binary = oracle.nerf.occupancy_grid.binary.to(dev)
    model.nerf.occupancy_grid.occs.copy_(oracle.nerf.occupancy_grid.occs)
    model.nerf.update_occ_grid = lambda *a, **k: None

    g = torch.Generator().manual_seed(5)
    event = synthetic.event_batch(n_events, cfg, poses[2], g)
    normalized = synthetic.normalized_batch(n_events, it_sample_size, g, pixel_bandwidth)
    rays = n_events * (it_sample_size if pixel_bandwidth else 1)
    jitters = [torch.rand(rays, generator=g) for _ in range(4)]

    loss_ref, _, samples_ref = oracle.training_step(event, normalized, jitters=jitters)
    loss_ref.backward()
    batch = {"event": {k: v.to(dev) for k, v in event.items()},
             "normalized": {k: v.to(dev) for k, v in normalized.items()}}
    loss = model.training_step(batch, 0, 0, jitters=[j.to(dev) for j in jitters])
    loss.backward()
    torch.cuda.synchronize()

    err = {"loss": _rel(loss, loss_ref)}
    ref_grads = {n: p.grad for n, p in oracle.named_parameters() if p.grad is not None}
    for n, p in model.named_parameters():
        if n in ref_grads and n.startswith("nerf."):
            err[n] = _rel(p.grad, ref_grads[n])
    worst = max(err.values())
    launched = ops.launch_count() - launches0
    if verbose:
        print(f"smoke: loss {loss.item():.6f} (oracle {loss_ref.item():.6f}), "
              f"{model.logged['train/mean_num_samples_per_ray']:.1f} samples/ray "
              f"(oracle {samples_ref:.1f}), worst rel err {worst:.2e}, "
              f"{launched} den_b200 kernel launches, pixel_bandwidth={pixel_bandwidth}")
    assert launched > 0, "no den_b200 kernels were launched"
    assert worst < tol, err
    return err
