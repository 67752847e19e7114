"""GPU parity of the B2 path (drop-in NeRF / field / training step) against the golden
fixtures written from the REFERENCE'S OWN files and against the CPU oracle.

Tolerances (north_star): sample counts exact; rendered values, loss and parameter gradients
within 1e-3 relative in fp32."""

import pytest
import torch

import _scene

pytestmark = pytest.mark.gpu

TOL = 1e-3
COND_FACTOR = 16


def _grad_bounds(scene, keys):
    """Per-parameter gradient bound of the training-step parity tests: 1e-3 relative (north_star),
    except where the REFERENCE'S OWN gradient is not reproducible to that level — then
    COND_FACTOR x the relative change of the reference's fp32 gradient under one-ulp perturbations of
    its field parameters (tests/golden/gradient_conditioning.npz, written by make_golden.py from the
    reference's files).  Those are sums that cancel almost completely (the pixel-bandwidth parameters,
    the mean contrast threshold, the output-layer bias on the EDS shape: c = 1e-3 .. 4e-3); a GPU run
    of the reference itself moves them as much (its atomics reorder the sums), so no evaluation can
    be pinned tighter.  Everywhere else c < 7e-5 and the bound is the plain 1e-3."""
    cond = _scene.load_golden("gradient_conditioning")
    return {k: max(TOL, COND_FACTOR * float(cond[f"{scene}/{k}"])) for k in keys}


def _rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-30)).item()


def test_fused_field_matches_reference_golden(den_lib, cuda):
    golden = _scene.load_golden("field_small")
    cfg = _scene.scene_config("synthetic", occ_resolution=32, small=True)
    nerf = _scene.build_product_nerf(cfg, cuda)
    field = nerf.radiance_field
    field.load_state_dict(_scene.golden_section(golden, "state", cuda))
    x = torch.from_numpy(golden["x"]).to(cuda)
    d = torch.from_numpy(golden["d"]).to(cuda)
    # fused kernels: density at explicit positions, and (sigma, rgb) for "marched" samples
    sig = field.density_at(x)
    assert _rel(sig, golden["sigma"]) < 1e-4
    n = x.shape[0]
    ray_idx = torch.arange(n, dtype=torch.int32, device=cuda)
    t0 = torch.zeros(n, device=cuda)
    sig2, rgb2 = field.eval_samples(x, d, ray_idx, t0, t0, full=True)
    assert _rel(sig2, golden["sigma"][:, 0]) < 1e-4
    assert _rel(rgb2, golden["rgb"]) < 1e-4
    # operator-API path with autograd: values and every parameter gradient
    rgb, sigma = field(x, d)
    assert _rel(rgb, golden["rgb"]) < 1e-4 and _rel(sigma, golden["sigma"]) < 1e-4
    ((rgb * torch.from_numpy(golden["w_rgb"]).to(cuda)).sum()
     + (sigma * torch.from_numpy(golden["w_sigma"]).to(cuda)).sum()).backward()
    for key, p in field.named_parameters():
        assert _rel(p.grad, golden["grad/" + key]) < TOL, key


@pytest.mark.parametrize("training", [False, True], ids=["eval", "train"])
@pytest.mark.parametrize("scene", ["synthetic", "eds"])
def test_nerf_render_matches_oracle(den_lib, cuda, scene, training):
    cfg = _scene.scene_config(scene, occ_resolution=32, small=True)
    ora = _scene.build_oracle_nerf(cfg)
    prod = _scene.build_product_nerf(cfg, cuda)
    _scene.copy_params(ora, prod)
    ora.train()
    prod.train()
    poses = _scene.synthetic.camera_poses(cfg, n_poses=50)
    torch.manual_seed(3)
    ora.update_occ_grid(0, poses[0])
    prod.occupancy_grid._binary = ora.occupancy_grid.binary.to(cuda)
    prod.occupancy_grid.occs.copy_(ora.occupancy_grid.occs)
    ora.train(training)
    prod.train(training)

    traj = _scene.path_ref.LinearTrajectory(*poses)
    g = torch.Generator().manual_seed(2)
    n = 400
    ts = torch.rand(n, generator=g, dtype=torch.float64) * float(poses[2][-1])
    px = torch.stack([torch.rand(n, generator=g) * cfg["width"],
                      torch.rand(n, generator=g) * cfg["height"]], -1)
    pos, rot = traj(ts)
    kinv = torch.linalg.inv(torch.from_numpy(_scene.synthetic.intrinsics(cfg)))
    o, d = _scene.path_ref.NeRF.pixel_params_to_ray(kinv, px, pos, rot)
    jitter = torch.rand(n, generator=g)
    rad_o, opa_o, dep_o, ms_o = ora(o, d, jitter=jitter)
    rad_p, opa_p, dep_p, ms_p = prod(o.to(cuda), d.to(cuda), jitter=jitter.to(cuda))
    assert ms_o > 5, "degenerate scene"
    assert abs(ms_p - ms_o) * n <= 3, (ms_p, ms_o)      # identical sample sets (+- threshold ties)
    assert _rel(rad_p, rad_o) < TOL
    assert _rel(opa_p, opa_o) < TOL
    assert _rel(dep_p, dep_o) < TOL
    w = torch.randn(n, generator=g)
    ora.zero_grad()
    prod.zero_grad()
    ((rad_o + 1e-3).log() * w).sum().backward()
    ((rad_p + 1e-3).log() * w.to(cuda)).sum().backward()
    go, gp = _scene.flat_named_grads(ora), _scene.flat_named_grads(prod)
    assert set(go) == set(gp)
    for key in go:
        assert _rel(gp[key], go[key]) < TOL, key


def test_nerf_render_with_culled_samples_matches_oracle(den_lib, cuda):
    """A dense field (density bias +5: sigma ~ 150) makes the transmittance test of the visibility
    pre-pass cull most marched samples (external/utils.py:68-81, early_stop_eps 1e-4): covers the
    survivor gather of the pre-pass outputs and the grad pass on the compacted sample set."""
    cfg = _scene.scene_config("synthetic", occ_resolution=32, small=True)
    ora = _scene.build_oracle_nerf(cfg)
    with torch.no_grad():
        ora.radiance_field.mlp_base[1].output_layer.bias[0] += 5.0
    prod = _scene.build_product_nerf(cfg, cuda)
    _scene.copy_params(ora, prod)
    ora.train()
    prod.train()
    poses = _scene.synthetic.camera_poses(cfg, n_poses=50)
    torch.manual_seed(3)
    ora.update_occ_grid(0, poses[0])
    prod.occupancy_grid._binary = ora.occupancy_grid.binary.to(cuda)
    prod.occupancy_grid.occs.copy_(ora.occupancy_grid.occs)
    traj = _scene.path_ref.LinearTrajectory(*poses)
    g = torch.Generator().manual_seed(5)
    n = 600
    ts = torch.rand(n, generator=g, dtype=torch.float64) * float(poses[2][-1])
    px = torch.stack([torch.rand(n, generator=g) * cfg["width"],
                      torch.rand(n, generator=g) * cfg["height"]], -1)
    pos, rot = traj(ts)
    kinv = torch.linalg.inv(torch.from_numpy(_scene.synthetic.intrinsics(cfg)))
    o, d = _scene.path_ref.NeRF.pixel_params_to_ray(kinv, px, pos, rot)
    jitter = torch.rand(n, generator=g)
    marched = prod._march(o.to(cuda).float().contiguous(), d.to(cuda).float().contiguous(),
                          jitter.to(cuda))[0].numel()
    rad_o, opa_o, dep_o, ms_o = ora(o, d, jitter=jitter)
    rad_p, opa_p, dep_p, ms_p = prod(o.to(cuda), d.to(cuda), jitter=jitter.to(cuda))
    assert 1 < ms_o * n < 0.5 * marched, (ms_o * n, marched)     # most samples were culled
    assert abs(ms_p - ms_o) * n <= 3, (ms_p, ms_o)
    assert _rel(rad_p, rad_o) < TOL
    assert _rel(opa_p, opa_o) < TOL
    assert _rel(dep_p, dep_o) < TOL
    w = torch.randn(n, generator=g)
    ((rad_o + 1e-3).log() * w).sum().backward()
    ((rad_p + 1e-3).log() * w.to(cuda)).sum().backward()
    go, gp = _scene.flat_named_grads(ora), _scene.flat_named_grads(prod)
    assert set(go) == set(gp)
    for key in go:
        assert _rel(gp[key], go[key]) < TOL, key


def test_cell_densities_match_oracle_field(den_lib, cuda):
    """The fused density kernel of the occupancy update (`den_field_density_at`) against the oracle
    field at jittered cell centres.  The grid logic itself (draws, EMA, threshold, booleans bit for
    bit) is pinned by tests/test_gpu_occgrid.py through the product's `every_n_step`."""
    cfg = _scene.scene_config("synthetic", occ_resolution=32, small=True)
    ora = _scene.build_oracle_nerf(cfg)
    prod = _scene.build_product_nerf(cfg, cuda)
    _scene.copy_params(ora, prod)
    grid = ora.occupancy_grid
    g = torch.Generator().manual_seed(0)
    coords = grid.grid_coords
    x = (coords + torch.rand(coords.shape, generator=g)) / grid.resolution
    world = x * 3.0 - 1.5
    occ_o = ora.radiance_field.query_density(world).squeeze(-1) * cfg["step"]
    occ_p = prod.radiance_field.density_at(world.to(cuda)).squeeze(-1).cpu() * cfg["step"]
    assert _rel(occ_p, occ_o) < 1e-4


# The ray part of the refractory-period gradient is a sum over all rays of dL/dt_i that cancels to ~1 %
# of sum |dL/dt_i| and then again against the direct terms.  The closed-form reverse mode of the ray
# kernel and the reference's fp32 autograd through SLERP are BOTH within ~8e-6 of max |dL/dt_i| per ray of
# the float64 evaluation (profiles/rays_grad_check2.py: reference 7.7e-6, kernel 5.6e-6), with different
# rounding, which the cancellation amplifies to 5.7e-4 (reference) / 6.9e-4 (kernel) of the parameter
# gradient against float64 and 1.26e-3 against each other.  So the kernel mode is held to the golden at
# the standard bound plus this allowance on that one key, and the autograd mode (same kernels everywhere
# else) to the standard bound on every key.
TAU_KEY = "refractory_period.parametrizations._refractory_period.original"
TAU_RAY_ROUNDING = 1e-3            # measured 1.3e-3 total on the pixel-bandwidth golden (bound 2e-3), 1.6e-4 / 1.2e-5 elsewhere


def _run_training_step_golden(cuda, pb_on, bayer=False):
    golden = _scene.load_golden("training_step_bayer" if bayer else
                                "training_step_pb_on" if pb_on else "training_step_pb_off")
    cfg = _scene.scene_config("synthetic", occ_resolution=32, small=True)
    if bayer:
        cfg["radiance_dim"] = 3
    ref = _scene.golden_section(golden, "grad")
    modes = ("kernel", "autograd") if TAU_KEY in ref else ("kernel",)
    for mode in modes:
        model, poses = _scene.build_product_renderer(cfg, cuda, 8, pixel_bandwidth=pb_on)
        model.rays_reverse_mode = mode
        names = ["nerf", "contrast_threshold", "refractory_period"] + (
            ["pixel_bandwidth"] if pb_on else [])
        for name in names:
            _scene.load_golden_state(getattr(model, name), golden, name, cuda)
        batch = {"event": _scene.golden_section(golden, "event", cuda),
                 "normalized": _scene.golden_section(golden, "normalized", cuda)}
        jitters = [v for _, v in sorted(_scene.golden_section(golden, "jitter", cuda).items(),
                                        key=lambda kv: int(kv[0]))]
        model.train()
        # the golden carries the occupancy grid AFTER the reference's step-0 update: skip ours
        model.nerf.update_occ_grid = lambda *a, **k: None
        loss = model.training_step(batch, 0, 0, jitters=jitters)
        assert _rel(loss, golden["loss"]) < TOL
        for key in ("log_intensity_diff", "log_intensity_tv"):
            assert _rel(model.logged[f"train/{key}"], golden[f"logged/train/{key}"]) < TOL
        assert abs(model.logged["train/mean_num_samples_per_ray"]
                   - float(golden["logged/train/mean_num_samples_per_ray"])) < 0.02
        model.zero_grad()
        loss.backward()
        grads = _scene.flat_named_grads(model)
        assert set(ref) <= set(grads), set(ref) - set(grads)
        worst = {}
        for key in ref:
            worst[key] = _rel(grads[key], ref[key])
        print("golden pb_on" if pb_on else "golden pb_off", f"({mode} ray reverse mode)",
              "worst relative gradient errors:",
              {k.split(".")[-2] if k.endswith("original") else k.split("radiance_field.")[-1]: f"{v:.1e}"
               for k, v in worst.items()})
        bounds = _grad_bounds("pb_on" if pb_on else "pb_off", ref)
        assert max(bounds.values()) < 4e-3, bounds            # synthetic shape: (almost) everything at 1e-3
        if mode == "kernel" and TAU_KEY in bounds:
            bounds[TAU_KEY] += TAU_RAY_ROUNDING
        bad = {k: (v, bounds[k]) for k, v in worst.items() if v > bounds[k]}
        assert not bad, (mode, bad)


def test_bayer_render_paths_agree(den_lib, cuda):
    """The three host paths of a Bayer training step — fused filter + loss kernel, batched render calls with
    the per-call filter, four sequential render calls — select the pixel's channel at the same place and
    give the same loss and gradients."""
    golden = _scene.load_golden("training_step_bayer")
    cfg = _scene.scene_config("synthetic", occ_resolution=32, small=True)
    cfg["radiance_dim"] = 3
    results = []
    for fused, batched in ((True, True), (False, True), (False, False)):
        model, poses = _scene.build_product_renderer(cfg, cuda, 8, pixel_bandwidth=True)
        for name in ["nerf", "contrast_threshold", "refractory_period", "pixel_bandwidth"]:
            _scene.load_golden_state(getattr(model, name), golden, name, cuda)
        batch = {"event": _scene.golden_section(golden, "event", cuda),
                 "normalized": _scene.golden_section(golden, "normalized", cuda)}
        jitters = [v for _, v in sorted(_scene.golden_section(golden, "jitter", cuda).items(),
                                        key=lambda kv: int(kv[0]))]
        model.train()
        model.fuse_lpf_loss, model.batch_render_calls = fused, batched
        model.nerf.update_occ_grid = lambda *a, **k: None
        loss = model.training_step(batch, 0, 0, jitters=jitters)
        loss.backward()
        results.append((loss.detach(), _scene.flat_named_grads(model)))
    cond = _scene.load_golden("gradient_conditioning")
    for loss, grads in results[1:]:
        assert _rel(loss, results[0][0]) < 1e-5
        for key in grads:
            bound = max(5e-4, 2 * float(cond.get(f"pb_on/{key}", 0.0)))     # fp32 filter vs fp64 fused kernel
            assert _rel(grads[key], results[0][1][key]) < bound, (key, bound)


@pytest.mark.parametrize("pb_on", [False, True], ids=["pb_off", "pb_on"])
def test_batched_render_calls_equal_sequential_calls(den_lib, cuda, pb_on):
    """The four render calls of a step evaluated as one launch sequence (the default) give the
    loss, logged terms and gradients of four separate calls (same per-ray arithmetic; only the
    order of the atomic gradient accumulation differs)."""
    golden = _scene.load_golden("training_step_pb_on" if pb_on else "training_step_pb_off")
    cfg = _scene.scene_config("synthetic", occ_resolution=32, small=True)
    results = []
    for batched in (True, False):
        model, poses = _scene.build_product_renderer(cfg, cuda, 8, pixel_bandwidth=pb_on)
        for name in ["nerf", "contrast_threshold", "refractory_period"] + (
                ["pixel_bandwidth"] if pb_on else []):
            _scene.load_golden_state(getattr(model, name), golden, name, cuda)
        batch = {"event": _scene.golden_section(golden, "event", cuda),
                 "normalized": _scene.golden_section(golden, "normalized", cuda)}
        jitters = [v for _, v in sorted(_scene.golden_section(golden, "jitter", cuda).items(),
                                        key=lambda kv: int(kv[0]))]
        model.train()
        model.batch_render_calls = batched
        model.nerf.update_occ_grid = lambda *a, **k: None
        loss = model.training_step(batch, 0, 0, jitters=jitters)
        loss.backward()
        results.append((loss.detach(), dict(model.logged), _scene.flat_named_grads(model)))
    (la, ga, gra), (lb, gb, grb) = results
    assert _rel(la, lb) < 1e-6
    assert abs(ga["train/mean_num_samples_per_ray"] - gb["train/mean_num_samples_per_ray"]) < 1e-9
    assert _rel(ga["train/mean_ray_occ_rate"], gb["train/mean_ray_occ_rate"]) < 1e-6
    assert set(gra) == set(grb)
    for key in gra:
        assert _rel(gra[key], grb[key]) < 2e-4, key      # fp32 accumulation order differs


def test_fused_rays_match_the_torch_trajectory_path(den_lib, cuda):
    """den_rays_from_trajectory == LinearTrajectory.forward + pixel_params_to_ray (torch ops), incl.
    timestamps exactly on a pose stamp, on the first / last stamp and a (K, S, N) batch shape."""
    cfg = _scene.scene_config("synthetic", occ_resolution=32, small=True)
    model, poses = _scene.build_product_renderer(cfg, cuda, 8, pixel_bandwidth=False)
    tr = model.trajectory
    g = torch.Generator().manual_seed(4)
    n = 700
    t0, t1 = float(tr.T_wc_timestamp[0]), float(tr.T_wc_timestamp[-1])
    ts = t0 + torch.rand(3, 5, n, generator=g, dtype=torch.float64) * (t1 - t0)
    ts[0, 0, :8] = tr.T_wc_timestamp[:8].double().cpu()          # exactly on pose stamps (incl. the first)
    ts[0, 0, 8] = t1
    ts = ts.to(cuda)
    pix = (torch.rand(n, 2, generator=g) * 300).to(cuda)
    pos, rot = tr(ts)
    o_ref, d_ref = model.nerf.pixel_params_to_ray(model.train_intrinsics_inv, pix, pos, rot)
    o, d = model.rays(ts, pix)
    assert o.shape == (3, 5, n, 3) and d.shape == (3, 5, n, 3)
    assert (o - o_ref).abs().max().item() < 2e-6 * o_ref.abs().max().item() + 1e-6
    assert (d - d_ref).abs().max().item() < 5e-6
    assert torch.allclose(d.norm(dim=-1), torch.ones_like(d[..., 0]), atol=1e-6)


def test_fused_rays_reverse_mode_matches_the_torch_autograd_path(den_lib, cuda):
    """dL/d(timestamp) of den_rays_from_trajectory_bwd (closed form per pose interval) vs torch autograd
    through LinearTrajectory.forward + pixel_params_to_ray (the form the reference differentiates:
    models/trajectories.py:30-90, utils/tensor_ops.py:118-184, models/nerf.py:206-228)."""
    cfg = _scene.scene_config("synthetic", occ_resolution=32, small=True)
    model, poses = _scene.build_product_renderer(cfg, cuda, 8, pixel_bandwidth=False)
    tr = model.trajectory
    g = torch.Generator().manual_seed(5)
    n = 900
    t0, t1 = float(tr.T_wc_timestamp[0]), float(tr.T_wc_timestamp[-1])
    ts = (t0 + torch.rand(2, 4, n, generator=g, dtype=torch.float64) * (t1 - t0)).to(cuda)
    pix = (torch.rand(n, 2, generator=g) * 300).to(cuda)
    w_o = torch.randn(2, 4, n, 3, generator=g).to(cuda)
    w_d = torch.randn(2, 4, n, 3, generator=g).to(cuda)

    ts_ref = ts.clone().requires_grad_(True)
    pos, rot = tr(ts_ref)
    o_ref, d_ref = model.nerf.pixel_params_to_ray(model.train_intrinsics_inv, pix, pos, rot)
    ((o_ref * w_o).sum() + (d_ref * w_d).sum()).backward()

    ts_k = ts.clone().requires_grad_(True)
    o, d = model.rays(ts_k, pix)
    assert o.grad_fn is not None and type(o.grad_fn).__name__.startswith("_RaysFn")
    ((o * w_o).sum() + (d * w_d).sum()).backward()
    ref, got = ts_ref.grad, ts_k.grad
    assert got.dtype == torch.float64 and got.shape == ts.shape
    scale = ref.abs().max().item()
    assert scale > 0
    assert (got - ref).abs().max().item() < 2e-4 * scale, ((got - ref).abs().max().item(), scale)
    # against the float64 evaluation of the same functions the closed form is at least as close as the
    # reference's fp32 autograd (x 1.5 for noise)
    from deblur_e_nerf_b200 import trajectories
    tr64 = trajectories.LinearTrajectory((tr.T_wc_position.double(), tr.T_wc_orientation_quat.double(),
                                          tr.T_wc_timestamp))
    ts_64 = ts.clone().requires_grad_(True)
    pos, rot = tr64(ts_64)
    o64, d64 = model.nerf.pixel_params_to_ray(model.train_intrinsics_inv.double(), pix.double(), pos, rot)
    ((o64 * w_o.double()).sum() + (d64 * w_d.double()).sum()).backward()
    err_kernel = (got - ts_64.grad).abs().max().item()
    err_autograd = (ref - ts_64.grad).abs().max().item()
    assert err_kernel <= 1.5 * err_autograd + 1e-7 * scale, (err_kernel, err_autograd, scale)
    # and against the ORACLE's closed form (pinned on CPU against autograd through the reference's own
    # trajectory code: tests/test_oracle_vs_reference.py::test_trajectory_time_gradient_closed_form)
    from oracle import path_ref
    ora_tr = path_ref.LinearTrajectory(tr.T_wc_position.cpu(), tr.T_wc_orientation_quat.cpu(),
                                       tr.T_wc_timestamp.cpu())
    closed = path_ref.trajectory_time_gradient(ora_tr, ts.cpu(), d.detach().cpu(), w_o.cpu(), w_d.cpu())
    assert (got.cpu() - closed).abs().max().item() < 2e-5 * scale, ((got.cpu() - closed).abs().max().item(), scale)
    # one output only (the other gradient is None inside autograd)
    ts_o = ts.clone().requires_grad_(True)
    (model.rays(ts_o, pix)[0] * w_o).sum().backward()
    ts_o_ref = ts.clone().requires_grad_(True)
    (tr(ts_o_ref)[0] * w_o).sum().backward()
    assert (ts_o.grad - ts_o_ref.grad).abs().max().item() < 2e-4 * ts_o_ref.grad.abs().max().item()


def test_training_step_pb_off_matches_reference_golden(den_lib, cuda):
    _run_training_step_golden(cuda, pb_on=False)


def test_training_step_bayer_matches_reference_golden(den_lib, cuda):
    """A colour sensor behind a Bayer filter (models/deblur_e_nerf.py:82-90,409-412,1177-1178,1223-1234):
    three radiance channels through the field, compositing and their reverse passes, each event
    supervised on the channel of its pixel.  Golden: the reference's own files, CPU fp32."""
    _run_training_step_golden(cuda, pb_on=True, bayer=True)


@pytest.mark.parametrize("S", [8, 30])
def test_pixel_bandwidth_filter_matches_oracle(den_lib, cuda, S):
    """den_lpf_{fwd,bwd} (fp64 inside) vs the oracle PixelBandwidth evaluated in fp64 (truth) and
    in fp32 (what the reference computes): outputs, dL/dI and the six parameter gradients, for
    the reset call (two outputs) and a following non-reset call (reset state with gradient)."""
    from deblur_e_nerf_b200 import pixel_bandwidth as pb_mod
    from deblur_e_nerf_b200 import synthetic
    from oracle import path_ref
    calib = synthetic.calibration()
    g = torch.Generator().manual_seed(S)
    n = 300
    out_ts = 40e6 + torch.rand(n, generator=g, dtype=torch.float64) * 100e6
    gen = torch.full((S - 1, n), 0.5, dtype=torch.float64)
    base = torch.exp(torch.randn(S, n, generator=g) * 1.5 - 2.0).clamp(1e-3, 5.0)   # 0.001 .. 5
    wts = torch.randn(2, n, generator=g)

    def run(module, dev, dtype):
        leaf = base.to(dev, dtype).detach().clone().requires_grad_(True)

        def fn(ts):
            return (leaf, torch.tensor(1.0), 1.0, torch.ones_like(ts, dtype=torch.bool))
        module.zero_grad()
        y0, _ = module(gen.to(dev), (out_ts - 4e6).to(dev), fn, True)
        y1, _ = module(gen.to(dev), out_ts.to(dev), fn, False)
        loss = (y0 * wts[0].to(dev, y0.dtype)).sum() + (y1 * wts[1].to(dev, y1.dtype)).sum()
        loss.backward()
        grads = {k: v.grad.detach().cpu().double() for k, v in module.named_parameters()}
        return y0.detach().cpu().double(), y1.detach().cpu().double(), leaf.grad.cpu().double(), grads

    ora64 = path_ref.PixelBandwidth(calib, 0, 21, 0.95).double()
    ora32 = path_ref.PixelBandwidth(calib, 0, 21, 0.95)
    prod = pb_mod.PixelBandwidth(calib, 0, 21, dict(max_sample_lifetime=0.95)).to(cuda)
    for k, v in ora32.state_dict().items():
        assert torch.allclose(prod.state_dict()[k].cpu(), v), k
    ora64.load_state_dict({k: v.double() for k, v in ora32.state_dict().items()})
    t0, t1, tg, tp = run(ora64, "cpu", torch.float64)
    r0, r1, rg, rp = run(ora32, "cpu", torch.float32)
    p0, p1, pg, pp = run(prod, cuda, torch.float32)
    assert _rel(p0, t0) < 1e-5 and _rel(p1, t1) < 1e-5
    assert _rel(pg, tg) < 1e-4
    for k in tp:
        assert _rel(pp[k], tp[k]) < 1e-3, (k, pp[k], tp[k])
    # and no further from the truth than the reference's own fp32 evaluation is
    assert _rel(p0, t0) <= max(_rel(r0, t0), 1e-6) * 1.5
    assert _rel(pg, tg) <= max(_rel(rg, tg), 1e-5) * 1.5


def test_training_step_pb_on_matches_reference_golden(den_lib, cuda):
    _run_training_step_golden(cuda, pb_on=True)


def test_training_step_eds_two_microbatches_matches_reference_golden(den_lib, cuda):
    """BASELINE configs[3] shape (08_peanuts_running.yaml:57-70,203): sphere contraction, cone
    angle 0.004, no background (validity = opacity > 0), pixel bandwidth on, tau / Omega / C_p
    trainable, `accumulate_grad_batches` micro-batches accumulated with loss / n
    (models/deblur_e_nerf.py:465-469,1055-1112).  Golden: the reference's own files, CPU fp32.
    Exercises together: the sphere contraction reverse mode, cone-angle marching, the tau gradient
    through rays / samples / SH, the Omega gradients of the filter and the C_p gradients."""
    golden = _scene.load_golden("training_step_eds")
    n_micro = int(golden["n_micro"])
    cfg = _scene.scene_config("eds", occ_resolution=32, small=True)
    model, poses = _scene.build_product_renderer(cfg, cuda, 8, pixel_bandwidth=True)
    for name in ("nerf", "contrast_threshold", "refractory_period", "pixel_bandwidth"):
        _scene.load_golden_state(getattr(model, name), golden, name, cuda)
    model.train()
    model.accumulate_grad_batches = n_micro
    model.nerf.update_occ_grid = lambda *a, **k: None      # the golden holds the grid after its update
    model.zero_grad()
    for m in range(n_micro):
        batch = {"event": _scene.golden_section(golden, f"event/{m}", cuda),
                 "normalized": _scene.golden_section(golden, f"normalized/{m}", cuda)}
        jitters = [v for _, v in sorted(_scene.golden_section(golden, f"jitter/{m}", cuda).items(),
                                        key=lambda kv: int(kv[0]))]
        loss = model.training_step(batch, m, 0, jitters=jitters)
        assert _rel(loss, golden[f"loss/{m}"]) < TOL, (m, float(loss), float(golden[f"loss/{m}"]))
        for key in ("log_intensity_diff", "log_intensity_tv"):
            assert _rel(model.logged[f"train/{key}"], golden[f"logged/{m}/train/{key}"]) < TOL
        assert abs(model.logged["train/mean_num_samples_per_ray"]
                   - float(golden[f"logged/{m}/train/mean_num_samples_per_ray"])) < 0.05
        (loss / n_micro).backward()
    grads = _scene.flat_named_grads(model)
    ref = _scene.golden_section(golden, "grad")
    assert set(ref) <= set(grads), set(ref) - set(grads)
    worst = {key: _rel(grads[key], ref[key]) for key in ref}
    print("EDS golden, worst relative gradient errors:",
          {k.split(".")[-2] if k.endswith("original") else k.split("radiance_field.")[-1]: f"{v:.1e}"
           for k, v in worst.items()})
    bounds = _grad_bounds("eds", ref)
    loose = {k: b for k, b in bounds.items() if b > TOL}
    print("EDS keys with a conditioned bound:", {k: f"{b:.1e}" for k, b in loose.items()})
    assert all("pixel_bandwidth" in k or "mean_contrast" in k or k.endswith("output_layer.bias")
               for k in loose), loose
    bad = {k: (v, bounds[k]) for k, v in worst.items() if v > bounds[k]}
    assert not bad, bad


def test_eval_render_of_merged_chunks_matches_oracle(den_lib, cuda):
    """Eval mode with more rays than `test_chunk_size` (external/utils.py:99-103): the reference
    renders 16 384-ray chunks; the product merges chunks (first `test_chunk_size` rays, then sized
    from the samples per ray) — rays are independent and the eval march is deterministic, so the
    image must be the same.  40 000 rays = three reference chunks."""
    cfg = _scene.scene_config("synthetic", occ_resolution=32, small=True)
    ora = _scene.build_oracle_nerf(cfg)
    prod = _scene.build_product_nerf(cfg, cuda)
    _scene.copy_params(ora, prod)
    ora.train()
    prod.train()
    poses = _scene.synthetic.camera_poses(cfg, n_poses=50)
    torch.manual_seed(3)
    ora.update_occ_grid(0, poses[0])
    prod.occupancy_grid._binary = ora.occupancy_grid.binary.to(cuda)
    prod.occupancy_grid.occs.copy_(ora.occupancy_grid.occs)
    ora.eval()
    prod.eval()
    n = 40_000
    assert n > 2 * cfg["test_chunk_size"]
    traj = _scene.path_ref.LinearTrajectory(*poses)
    g = torch.Generator().manual_seed(8)
    ts = torch.full((n,), float(poses[2][20]), dtype=torch.float64)
    px = torch.stack([torch.rand(n, generator=g) * cfg["width"],
                      torch.rand(n, generator=g) * cfg["height"]], -1)
    pos, rot = traj(ts)
    kinv = torch.linalg.inv(torch.from_numpy(_scene.synthetic.intrinsics(cfg)))
    o, d = _scene.path_ref.NeRF.pixel_params_to_ray(kinv, px, pos, rot)
    with torch.no_grad():
        rad_o, opa_o, dep_o, ms_o = ora(o, d)
        chunks = []
        render_chunk = prod.render_chunk
        prod.render_chunk = lambda o_, d_, *a, **k: (chunks.append(o_.shape[0]), render_chunk(o_, d_, *a, **k))[1]
        rad_p, opa_p, dep_p, ms_p = prod(o.to(cuda), d.to(cuda))
    assert len(chunks) >= 2 and max(chunks) > cfg["test_chunk_size"], chunks   # merged chunks were used
    assert sum(chunks) == n
    # identical sample sets up to transmittance-threshold ties (early_stop_eps against sigma computed
    # in different arithmetic): a few samples in eleven million
    assert abs(ms_p - ms_o) * n <= 2e-6 * ms_o * n + 3, (ms_p, ms_o)
    assert _rel(rad_p, rad_o) < TOL and _rel(opa_p, opa_o) < TOL and _rel(dep_p, dep_o) < TOL
