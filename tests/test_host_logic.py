"""CPU tests of the product's host-side mirrors (no kernels involved) against the oracle:
trajectory / SLERP, event-generation parameters, supervision timestamps, loss, the
pixel-bandwidth sample schedule and the affine linearisation coefficients."""

import os

import pytest
import torch

from deblur_e_nerf_b200 import event_generation_params as egp
from deblur_e_nerf_b200 import loss as loss_mod
from deblur_e_nerf_b200 import pixel_bandwidth as pb_mod
from deblur_e_nerf_b200 import factory, renderer, synthetic, trajectories
from oracle import path_ref

import _scene


def _close(a, b, tol=1e-6):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    assert ((a - b).abs().max() / b.abs().max().clamp(min=1e-30)).item() <= tol


@pytest.mark.parametrize("scene", ["synthetic", "eds"])
def test_trajectory_matches_oracle(scene):
    cfg = synthetic.CONFIGS[scene]
    poses = synthetic.camera_poses(cfg, n_poses=60)
    prod = trajectories.LinearTrajectory(poses)
    ora = path_ref.LinearTrajectory(*poses)
    g = torch.Generator().manual_seed(0)
    ts = torch.rand(4, 50, generator=g, dtype=torch.float64) * float(poses[2][-1])
    ts[0, 0], ts[0, 1] = float(poses[2][0]), float(poses[2][-1])
    pp, rp = prod(ts)
    po, ro = ora(ts)
    _close(pp, po)
    _close(rp, ro)
    assert torch.allclose(rp @ rp.transpose(-1, -2), torch.eye(3).expand_as(rp), atol=1e-5)


def test_event_params_match_oracle_and_keys():
    calib = synthetic.calibration()
    ct = egp.ContrastThreshold(calib, True)
    rp = egp.RefractoryPeriod(calib, synthetic.MAX_REFRACTORY_PERIOD_NS)
    oct_ = path_ref.ContrastThreshold(calib["pos_contrast_threshold"], calib["neg_contrast_threshold"])
    orp = path_ref.RefractoryPeriod(calib["refractory_period"], synthetic.MAX_REFRACTORY_PERIOD_NS)
    assert set(ct.state_dict()) == set(oct_.state_dict())
    assert set(rp.state_dict()) == set(orp.state_dict())
    ev = {"num_pos": torch.tensor([1, 0, 1]), "num_neg": torch.tensor([0, 1, 0]),
          "start_ts": torch.tensor([1e6, 2e6, 3e6], dtype=torch.float64)}
    out = rp(ct(ev))
    _close(out["log_intensity_diff"], oct_(ev["num_pos"], ev["num_neg"]))
    _close(out["start_ts"], orp(ev["start_ts"]), 1e-12)
    _close(ct.mean_contrast_threshold, 0.259375, 1e-5)
    assert out["start_ts"].dtype == torch.float64
    # tau parametrisation keeps the gradient alive at the clamp (:204-224)
    rp.refractory_period.backward()
    assert rp.parametrizations._refractory_period.original.grad.abs() > 0


def test_supervision_timestamps_and_loss_match_oracle():
    cfg = _scene.scene_config("synthetic")
    poses = synthetic.camera_poses(cfg, n_poses=100)
    event, normalized = _scene.make_batch(cfg, poses, 200, 8, seed=3)
    ev = {"start_ts": event["start_ts"].double() + 40560.0, "end_ts": event["end_ts"]}
    d_p, s_p = renderer.EventRenderer.supervision_timestamps(ev, normalized, True, True)
    d_o, s_o = path_ref.supervision_timestamps(ev["start_ts"], ev["end_ts"], normalized, True, True)
    for a, b in ((d_p, d_o), (s_p, s_o)):
        for key in ("ts_diff", "start_ts", "end_ts"):
            _close(a[key], b[key], 1e-14)
    assert torch.all(s_p["start_ts"] >= d_p["start_ts"]) and torch.all(s_p["end_ts"] <= d_p["end_ts"])

    g = torch.Generator().manual_seed(1)
    n = 200
    weight = dict(log_intensity_diff=1.0, log_intensity_tv=1e-3)
    err = dict(log_intensity_diff="huber", log_intensity_tv="l1")
    norm = dict(log_intensity_diff=True, log_intensity_tv=True)
    lp = loss_mod.Loss(weight, err, norm)
    lo = path_ref.EventLoss(weight, err, norm)
    pred_d = torch.randn(n, generator=g)
    pred_s = torch.randn(n, generator=g) * 0.1
    valid_d = torch.rand(n, generator=g) > 0.3
    valid_s = torch.rand(n, generator=g) > 0.3
    log_diff = torch.where(torch.rand(n, generator=g) > 0.5, 0.26875, -0.25)
    mean_ct = torch.tensor(0.259375)
    be = {"log_intensity_diff": log_diff, "start_ts": ev["start_ts"], "end_ts": ev["end_ts"].double()}
    out_p = lp.compute(dict(be), {"log_intensity_diff": pred_d, "ts_diff": d_p["ts_diff"], "is_valid": valid_d},
                       {"log_intensity_diff": pred_s, "is_valid": valid_s}, mean_ct)
    out_o = lo.compute(log_diff, be["start_ts"], be["end_ts"],
                       {"log_intensity_diff": pred_d, "ts_diff": d_o["ts_diff"], "is_valid": valid_d},
                       {"log_intensity_diff": pred_s, "is_valid": valid_s}, mean_ct)
    for key in out_o:
        _close(out_p[key], out_o[key], 1e-6)


@pytest.mark.parametrize("S", [8, 30])
def test_pixel_bandwidth_schedule_and_coefficients(S):
    calib = synthetic.calibration()
    prod = pb_mod.PixelBandwidth(calib, 0, 21, dict(max_sample_lifetime=0.95))
    ora = path_ref.PixelBandwidth(calib, 0, 21, 0.95)
    assert set(prod.state_dict()) == set(ora.state_dict())
    for k, v in ora.state_dict().items():
        _close(prod.state_dict()[k], v, 1e-7)
    gen = torch.full((S - 1, 5), 0.5, dtype=torch.float64)
    _close(prod.sample_lifetimes(gen), ora.sample_lifetimes(gen), 1e-14)
    life = prod.sample_lifetimes(gen)
    assert life[-1].abs().max() == 0 and torch.all(life[:-1] > life[1:])      # oldest first
    # a = alpha0 + alpha1 I and b = beta I reproduce linearized_sys_params (:181-194)
    it = torch.tensor([1e-3, 0.05, 0.7, 3.0], dtype=torch.float64)
    c = prod.coefficients()
    tau_in = ora.tau_in_it_eff_prod.double() / it
    tau_mil = ora.param("tau_mil_it_eff_prod").double() / it
    tau_out = ora.param("tau_out").double()
    prod_ = (tau_in + tau_mil) * tau_out
    two_zeta_wn = (tau_in + tau_out + (1 / ora.param("A_amp_inv").double() + 1) * tau_mil) / prod_
    wn_sq = (1 / ora.param("A_loop_inv").double() + 1) / prod_
    _close(c[0] + c[1] * it, two_zeta_wn, 1e-6)
    _close(c[2] * it, wn_sq, 1e-6)
    _close(c[3], 1 / ora.param("tau_sf").double(), 1e-6)
    _close(c[4], 1 / ora.param("tau_diff").double(), 1e-6)


def test_march_segment_length_bounds_the_sample_count():
    """Host-side bound of the samples a ray can emit between near and far (single-pass march)."""
    from deblur_e_nerf_b200 import ops
    step = 3 ** 0.5 * 3 / 1024
    seg = ops.march_segment_length(1.43, 6.63, step)
    assert seg >= int((6.63 - 1.43) / step) + 1            # every emitted sample advances >= one step
    assert seg <= int((6.63 - 1.43) / step) + 8
    assert ops.march_segment_length(None, 6.63, step) is None
    assert ops.march_segment_length(0.0, float("inf"), step) is None
    assert ops.march_segment_length(2.0, 1.0, step) == 4     # empty span: slack only


def test_sample_buffers_are_allocated_in_row_quanta():
    """Per-sample buffers round their ROW count up to a quantum (stable block sizes for the caching
    allocator) and hand out the leading, contiguous view."""
    import torch
    from deblur_e_nerf_b200 import ops
    small = ops._rows(1000, (3,), torch.float32, "cpu")
    assert small.shape == (1000, 3) and small.untyped_storage().nbytes() == 1000 * 3 * 4
    n = (1 << 18) + 5
    big = ops._rows(n, (2,), torch.float32, "cpu")
    assert big.shape == (n, 2) and big.is_contiguous()
    assert big.untyped_storage().nbytes() == 2 * (1 << 18) * 2 * 4
    like = ops._rows_like(big)
    assert like.shape == big.shape and like.dtype == big.dtype
    exact = ops._rows(2 << 18, (), torch.int32, "cpu")
    assert exact.numel() == 2 << 18 and exact.untyped_storage().nbytes() == (2 << 18) * 4


def test_event_batch_producer_shapes_and_distributions():
    """Device-side batch producer (here on CPU tensors): layout, dtypes, ranges, per-rank streams and
    the batch-size controller hook."""
    import torch
    from deblur_e_nerf_b200.data import EventBatchProducer
    g = torch.Generator().manual_seed(0)
    n = 20000
    events = {"position": torch.rand(n, 2, generator=g) * 100,
              "start_ts": torch.arange(n) * 1000, "end_ts": torch.arange(n) * 1000 + 500,
              "num_pos": torch.ones(n, dtype=torch.int64), "num_neg": torch.zeros(n, dtype=torch.int64)}
    prod = EventBatchProducer(events, 4096, it_sample_size=30, device="cpu", seed=5, rank=0)
    b = prod.next_batch()
    ev, nm = b["event"], b["normalized"]
    assert ev["position"].shape == (4096, 2) and ev["position"].dtype == torch.float32
    assert ev["end_ts"].dtype == torch.int64 and torch.equal(ev["end_ts"] - ev["start_ts"],
                                                             torch.full((4096,), 500))
    assert nm["interval_gen"].shape == (29, 4096) and bool((nm["interval_gen"] == 0.5).all())
    assert bool((nm["ts_diff"] == 1).all())
    for k in ("diff_start_ts", "ts_subdiff", "subdiff_start_ts"):
        assert nm[k].dtype == torch.float64 and 0 <= float(nm[k].min()) and float(nm[k].max()) < 1
    # triangular with mode 0: mean 1/3; uniform: mean 1/2
    assert abs(float(nm["ts_subdiff"].mean()) - 1 / 3) < 0.02
    assert abs(float(nm["diff_start_ts"].mean()) - 0.5) < 0.02
    other = EventBatchProducer(events, 4096, it_sample_size=None, device="cpu", seed=5, rank=1).next_batch()
    assert "interval_gen" not in other["normalized"]
    assert not torch.equal(other["event"]["end_ts"], ev["end_ts"])          # rank-offset stream
    prod.set_batch_size(100)
    assert prod.next_batch()["event"]["num_pos"].shape == (100,)
    trimmed = EventBatchProducer(events, 64, device="cpu", dataset_len=10)
    assert int(trimmed.next_batch()["event"]["end_ts"].max()) <= 9 * 1000 + 500


def test_batch_controller_gating_follows_the_reference():
    """models/deblur_e_nerf.py:1252-1308: N_next = int(budget / mean samples per ray); with gradient
    accumulation only the second-to-last micro-batch of a window updates it (so that, with the one
    prefetched batch, every micro-batch of the next window has the same size); the cross-rank mean
    is taken before the division."""
    import types
    stub = types.SimpleNamespace(mean_samples_reduce_fn=None, accumulate_grad_batches=1,
                                 train_ray_sample_batch_size=131072, next_train_batch_size=None)
    update = renderer.EventRenderer.update_train_batch_size
    assert update(stub, [80.0, 70.0, 90.0, 80.0], 0) == 80.0
    assert stub.next_train_batch_size == int(131072 / 80.0)
    stub.accumulate_grad_batches, stub.next_train_batch_size = 4, None
    for batch_index in (0, 1, 3, 4, 5, 7):
        update(stub, [50.0], batch_index)
        assert stub.next_train_batch_size is None, batch_index
    update(stub, [50.0], 2)
    assert stub.next_train_batch_size == int(131072 / 50.0)
    update(stub, [64.0], 6)
    assert stub.next_train_batch_size == 2048
    stub.accumulate_grad_batches = 1
    stub.mean_samples_reduce_fn = lambda mean: (mean + 3 * mean) / 2        # two ranks: m and 3 m
    assert update(stub, [32.0], 0) == 64.0 and stub.next_train_batch_size == 2048


@pytest.mark.parametrize("colour", [False, True], ids=["mono", "bayer"])
def test_evaluation_loop_host_logic(monkeypatch, colour):
    """`Trainer.test` -> `EventRenderer.evaluation_step` / `evaluation_epoch_end` (models/deblur_e_nerf.py:
    604-969) on the CPU with the two device parts replaced: the field render by a closed-form stub, the
    post-processing kernels by the oracle's `eval_ref.evaluate`.  Checks what the host code owns: the
    squeeze of the DataLoader dim, pose expansion, channel-first colour images, unity exposure / gain
    defaults, stacking, the metric names, eval mode inside and the mode restored after."""
    from deblur_e_nerf_b200 import eval_post, trainer
    from oracle import eval_ref
    cfg = _scene.scene_config("synthetic", occ_resolution=32, small=True)
    model, poses = _scene.build_product_renderer(cfg, "cpu", pixel_bandwidth=False, n_poses=20)
    H, W, C = 19, 23, 3 if colour else 1
    seen = []

    def render(o, d, jitter=None):
        assert o.shape == (H, W, 3) and d.shape == (H, W, 3) and not model.nerf.training
        seen.append(o[0, 0].clone())
        base = 0.3 + 0.2 * d[..., 0].abs() + 0.1 * d[..., 1].abs()
        rad = torch.stack([base * (1 + 0.1 * c) for c in range(3)], dim=-1) if colour else base
        return rad, torch.ones(H, W), torch.ones(H, W), 7.0

    def evaluate(pred, target, exposure_time, gain, lo, hi, black_level_offset=True, init=None,
                 max_steps=10, radius=1e6, per_channel_scale=True):
        if pred.dim() == 3:
            pred, target = pred[:, None], target[:, None]
        res = eval_ref.evaluate(pred, target, exposure_time, gain, lo, hi,
                                black_level_offset=black_level_offset, init=init, max_steps=max_steps,
                                per_channel_scale=per_channel_scale)
        return {k: (torch.tensor(v) if isinstance(v, float) else v) for k, v in res.items()}

    monkeypatch.setattr(model.nerf, "forward", render)
    monkeypatch.setattr(eval_post, "evaluate", evaluate)
    traj = path_ref.LinearTrajectory(*poses)
    pos, rot = traj(torch.tensor([2.0e6, 9.5e6, 17.25e6], dtype=torch.float64))
    g = torch.Generator().manual_seed(0)
    shape = (3, H, W) if colour else (H, W)
    views = [{"img": (torch.rand(shape, generator=g) * 0.5 + 0.2)[None], "T_wc_position": pos[b][None],
              "T_wc_orientation": rot[b][None], "gain": torch.tensor([1.0 + b])} for b in range(3)]
    model.train()
    row, pred = trainer.Trainer().test(model, views, model.train_intrinsics_inv, 0.0, 1.0,
                                       black_level_offset=False, per_channel_log_it_scale=not colour)
    assert model.training and set(row) == {"test/l1", "test/psnr", "test/ssim"}
    assert pred.shape == (3, C, H, W)
    assert all(torch.equal(s, pos[b]) for b, s in enumerate(seen))          # every view at its own pose
    # the same numbers straight from the oracle on the stub's images
    grid = model.image_pixel_positions(H, W)
    preds = []
    for b in range(3):
        _, d = model.nerf.pixel_params_to_ray(model.train_intrinsics_inv, grid, pos[b].expand(H, W, -1),
                                              rot[b].expand(H, W, -1, -1))
        base = 0.3 + 0.2 * d[..., 0].abs() + 0.1 * d[..., 1].abs()
        chans = [base * (1 + 0.1 * c) for c in range(3)] if colour else [base]
        preds.append(torch.stack(chans) + model.min_modeled_intensity)
    target = torch.stack([v["img"][0] for v in views]).view(3, C, H, W)
    want = eval_ref.evaluate(torch.stack(preds), target, torch.ones(3, dtype=torch.int64),
                             torch.tensor([1.0, 2.0, 3.0]), 0.0, 1.0, black_level_offset=False,
                             per_channel_scale=not colour)
    assert abs(row["test/l1"] - want["l1"]) < 1e-6 and abs(row["test/ssim"] - want["ssim"]) < 1e-6
    assert torch.allclose(pred, want["pred"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("shared_gamma", [False, True], ids=["gamma_per_channel", "gamma_shared"])
def test_lm_refine_joint_system_host_logic(monkeypatch, shared_gamma):
    """`eval_post.lm_refine` assembles the joint normal equations (scale / offset per channel, gamma per
    channel or shared) from the per-channel moments of `den_eval_lm_moments`.  With the kernel replaced by
    a float64 torch evaluation of the same ten moments, the host logic — joint layout, damping, trust
    region, acceptance — must reproduce the oracle's refinement, which runs the reference's LM on the
    explicit Jacobian."""
    from deblur_e_nerf_b200 import eval_post
    from oracle import eval_ref
    g = torch.Generator().manual_seed(8)
    B, C, H, W = 3, 3, 12, 15
    target = torch.rand(B, C, H, W, generator=g) * 0.8 + 0.1
    exposure, gain = torch.tensor([1, 2, 4]), torch.tensor([1.0, 1.5, 0.75])
    norm = eval_ref.normalized_gain(gain, exposure)
    scene = (target - 0.03) / norm.view(-1, 1, 1, 1)
    pred = torch.stack([(0.5 + 0.2 * c) * scene[:, c].pow(1.2) for c in range(C)], dim=1)
    pred = (pred * torch.exp(0.02 * torch.randn(pred.shape, generator=g))).float()
    affine, fitted, _ = eval_ref.affine_log_correction(pred, target, norm, per_channel_scale=not shared_gamma)
    init = (torch.ones(C), torch.ones(1 if shared_gamma else C), torch.zeros(C))
    want, want_errors = eval_ref.lm_refine(fitted.exp(), target, norm, init)

    def moments(kind, pred, target, gain_vec, params, n_out):
        assert kind == "den_eval_lm_moments" and n_out == 10 and params.shape == (C, 5)
        a, b, s, gm, o = (params[:, k].view(1, C, 1, 1) for k in range(5))
        x = torch.exp(a * pred.log().double() + b)
        gv = gain_vec.view(-1, 1, 1, 1)
        xg = x.pow(gm)
        js, jo = gv * xg, (-gv).expand_as(x)
        jg = s * x.log() * js
        r = gv * (s * xg - o) - target.double()
        cols = [js * js, js * jg, js * jo, jg * jg, jg * jo, jo * jo, js * r, jg * r, jo * r, r * r]
        return torch.stack([c.transpose(0, 1).reshape(C, -1).sum(-1) for c in cols], dim=-1)

    monkeypatch.setattr(eval_post, "_moments", moments)
    got, errors = eval_post.lm_refine(pred, target, norm.double(), affine, init)
    assert torch.allclose(got, want, rtol=1e-7, atol=1e-9), (got, want)
    assert len(errors) == len(want_errors) and abs(errors[-1] - want_errors[-1]) <= 1e-9 * want_errors[-1]
    if shared_gamma:
        assert float(got[:, 1].max() - got[:, 1].min()) == 0.0


def test_posed_views_feed_the_evaluation_loop(tmp_path, monkeypatch):
    """Dataset directory -> `views.PosedViews` -> `Trainer.test`: the views iterate as the per-view dicts the
    loop takes, `test_arguments()` supplies the inverse intrinsics and the pixel range, quantised 8-bit images
    land in [0.5 / 256, 1 - 0.5 / 256], poses are in the common camera frame (rotation matrices)."""
    import _dataset
    from deblur_e_nerf_b200 import eval_post, trainer, views
    from oracle import eval_ref
    cfg = _scene.scene_config("synthetic", occ_resolution=32, small=True)
    _dataset.write(str(tmp_path), dict(synthetic.CONFIGS["synthetic"]), n_events=8, n_views=3, size=(19, 23), channels=3)
    posed = views.PosedViews(str(tmp_path), "test")
    assert len(posed) == 3 and posed.img.shape == (3, 19, 23) and posed.img.dtype == torch.float32
    assert posed.min_normalized_pixel_value == 0.5 / 256 and posed.max_normalized_pixel_value == 1 - 0.5 / 256
    assert float(posed.img.min()) >= posed.min_normalized_pixel_value and float(posed.img.max()) <= posed.max_normalized_pixel_value
    R = posed.T_wc_orientation
    assert torch.allclose(R @ R.transpose(1, 2), torch.eye(3).expand(3, 3, 3), atol=1e-5)
    assert torch.allclose(torch.linalg.det(R), torch.ones(3), atol=1e-5)
    assert "".join(chr(c) for c in posed.sample_id[1]).strip() == "r_1"
    shuffled = views.PosedViews(str(tmp_path), "test", permutation_seed=4)
    perm = torch.randperm(3, generator=torch.Generator().manual_seed(4))
    assert torch.equal(shuffled.img, posed.img[perm]) and torch.equal(shuffled.sample_id, posed.sample_id[perm])

    model, _ = _scene.build_product_renderer(cfg, "cpu", pixel_bandwidth=False, n_poses=20)

    def render(o, d, jitter=None):
        return 0.3 + 0.2 * d[..., 0].abs(), torch.ones(19, 23), torch.ones(19, 23), 7.0

    def evaluate(pred, target, exposure_time, gain, lo, hi, black_level_offset=True, init=None,
                 max_steps=10, radius=1e6, per_channel_scale=True):
        res = eval_ref.evaluate(pred[:, None], target[:, None], exposure_time, gain, lo, hi,
                                black_level_offset=black_level_offset)
        return {k: (torch.tensor(v) if isinstance(v, float) else v) for k, v in res.items()}

    monkeypatch.setattr(model.nerf, "forward", render)
    monkeypatch.setattr(eval_post, "evaluate", evaluate)
    row, pred = trainer.Trainer().test(model, posed, black_level_offset=False, **posed.test_arguments())
    assert set(row) == {"test/l1", "test/psnr", "test/ssim"} and pred.shape == (3, 1, 19, 23)
    assert 0 < row["test/l1"] < 1 and row["test/psnr"] > 0


def test_config_train_wires_the_loop_and_the_validation(tmp_path, monkeypatch):
    """`config.train` end to end on the CPU with the three device parts replaced (training step, field render,
    post-processing kernels): seed, model, optimizer groups, scheduler interval, event producer from the cached
    events.pt (per-rank batch, dataset ratio), the optimizer-step loop, and the validation views scored between
    epochs when `limit_val_batches` is not 0."""
    import _dataset
    from deblur_e_nerf_b200 import config, eval_post
    from oracle import eval_ref
    root = str(tmp_path)
    _dataset.write(root, dict(synthetic.CONFIGS["synthetic"]), channels=3)
    cfg = _dataset.reference_style_config(root)
    cfg["trainer"].update(max_epochs=2, limit_train_batches=3, limit_val_batches=1.0, check_val_every_n_epoch=1)
    cfg["data"]["train_dataset_ratio"] = 0.5
    cfg["model"]["correction"]["black_level_offset"] = False
    seen = {"batches": [], "renders": 0}

    def training_step(self, batch, batch_index=0, global_step=0, jitters=None):
        seen["batches"].append((batch["event"]["position"].shape[0], batch["normalized"]["interval_gen"].shape))
        loss = sum((p ** 2).sum() for p in self.nerf.radiance_field.mlp_head.parameters()) * 1e-3
        self.logged = {"train/loss": loss.detach()}
        return loss

    def forward(self, o, d, jitter=None, groups=1):
        seen["renders"] += 1
        return 0.3 + 0.2 * d[..., 0].abs(), torch.ones(o.shape[:-1]), torch.ones(o.shape[:-1]), 7.0

    def evaluate(pred, target, exposure_time, gain, lo, hi, black_level_offset=True, init=None,
                 max_steps=10, radius=1e6, per_channel_scale=True):
        res = eval_ref.evaluate(pred[:, None], target[:, None], exposure_time, gain, lo, hi,
                                black_level_offset=black_level_offset)
        return {k: (torch.tensor(v) if isinstance(v, float) else v) for k, v in res.items()}

    monkeypatch.setattr(renderer.EventRenderer, "training_step", training_step)
    from deblur_e_nerf_b200 import nerf as nerf_mod
    monkeypatch.setattr(nerf_mod.NeRF, "forward", forward)
    monkeypatch.setattr(eval_post, "evaluate", evaluate)
    rows = []
    model, loop = config.train(cfg, device="cpu", log_fn=lambda step, row: rows.append((step, row)))
    assert loop.global_step == 6 and loop.current_epoch == 2
    assert seen["batches"] == [(64, (3, 64))] * 6                      # train_init_eff_batch_size, S - 1 = 3
    assert seen["renders"] == 2 * 2                                    # two validation views after each epoch
    val_rows = [row for _, row in rows if "val/l1" in row]
    assert len(val_rows) == 2 and set(val_rows[0]) == {"val/l1", "val/psnr", "val/ssim"}
    assert len([row for _, row in rows if "train/loss" in row]) == 6 and model.training
    assert cfg["seed"] == 3
    # resume: `trainer.resume_from_checkpoint` restores the counters, so two more epochs end at step 12
    ckpt = str(tmp_path / "ckpt" / "last.ckpt")
    loop.save_checkpoint(ckpt, model, torch.optim.Adam(factory.optimizer_param_groups(model), lr=0.01))
    cfg["trainer"].update(resume_from_checkpoint=ckpt, max_epochs=4, limit_val_batches=0)
    model2, loop2 = config.train(cfg, device="cpu")
    assert loop2.global_step == 12 and loop2.current_epoch == 4
    assert len(seen["batches"]) == 12


def test_save_predictions_follows_the_reference_quantisation(tmp_path):
    """`views.save_predictions` (`eval_save_pred_intensity_img`, models/deblur_e_nerf.py:1008-1054): range
    normalisation, clipping, rounding to 8 bits, grey or RGB -> BGR on disk, files named by the sample ids."""
    import cv2
    import numpy as np
    from deblur_e_nerf_b200 import views
    g = torch.Generator().manual_seed(0)
    lo, hi = 0.5 / 256, 1 - 0.5 / 256
    ids = torch.tensor([[ord(ch) for ch in name.ljust(16)] for name in ("r_0", "view_12")])
    assert views.sample_id_strings(ids) == ["r_0", "view_12"]
    for channels in (1, 3):
        pred = torch.rand(2, channels, 9, 11, generator=g) * 1.2 - 0.1            # some values outside the range
        paths = views.save_predictions(pred, ids, str(tmp_path / f"c{channels}"), lo, hi)
        assert [os.path.basename(p) for p in paths] == ["r_0.png", "view_12.png"]
        want = (255 * ((pred - lo) / (hi - lo)).clamp(0, 1)).round().numpy().astype(np.uint8)
        for b, path in enumerate(paths):
            img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
            if channels == 1:
                assert img.shape == (9, 11) and np.array_equal(img, want[b, 0])
            else:
                assert img.shape == (9, 11, 3) and np.array_equal(img[..., ::-1], want[b].transpose(1, 2, 0))
    with pytest.raises(ValueError):
        views.save_predictions(torch.rand(3, 1, 4, 4), ids, str(tmp_path / "bad"), lo, hi)
