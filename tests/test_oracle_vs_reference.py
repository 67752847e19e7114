"""Pin ``oracle/path_ref.py`` against the reference's OWN files (run unmodified under
``oracle/ref_shim``) on identical parameters and seeded inputs.  Needs
``/root/reference`` — build container only; skipped elsewhere (the committed
``tests/golden`` fixtures carry the same pin to the GPU box)."""

import pytest
import torch

from oracle import path_ref, ref_shim

import _scene

pytestmark = [
    pytest.mark.reference,
    pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present"),
]


def _close(a, b, tol=1e-6):
    a, b = a.detach().double(), b.detach().double()
    denom = b.abs().max().clamp(min=1e-30)
    assert ((a - b).abs().max() / denom).item() <= tol, ((a - b).abs().max(), denom)


@pytest.fixture(scope="module")
def pair():
    cfg = _scene.scene_config("synthetic", occ_resolution=32, small=True)
    ref, poses = _scene.build_reference_renderer(cfg, it_sample_size=8)
    ora, _ = _scene.build_oracle_renderer(cfg, it_sample_size=8)
    for name in ("nerf", "contrast_threshold", "refractory_period", "pixel_bandwidth"):
        _scene.copy_params(getattr(ref, name), getattr(ora, name))
    return cfg, ref, ora, poses


def test_state_dict_keys_match(pair):
    _, ref, ora, _ = pair
    for name in ("nerf", "contrast_threshold", "refractory_period", "pixel_bandwidth"):
        ref_keys = set(getattr(ref, name).state_dict().keys())
        ora_keys = set(getattr(ora, name).state_dict().keys())
        assert ref_keys == ora_keys, (name, ref_keys ^ ora_keys)


def test_field_forward_backward(pair):
    _, ref, ora, _ = pair
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(500, 3, generator=g) * 3.4 - 1.7)
    d = torch.randn(500, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    rgb_r, sig_r = ref.nerf.radiance_field(x, d)
    rgb_o, sig_o = ora.nerf.radiance_field(x, d)
    _close(rgb_o, rgb_r)
    _close(sig_o, sig_r)
    ref.nerf.zero_grad()
    ora.nerf.zero_grad()
    (rgb_r.sum() + sig_r.sum()).backward()
    (rgb_o.sum() + sig_o.sum()).backward()
    gr, go = _scene.flat_named_grads(ref.nerf), _scene.flat_named_grads(ora.nerf)
    assert gr.keys() == go.keys()
    for k in gr:
        _close(go[k], gr[k], 1e-5)


def test_trajectory_and_rays(pair):
    cfg, ref, ora, poses = pair
    g = torch.Generator().manual_seed(1)
    ts = (torch.rand(3, 40, generator=g, dtype=torch.float64) * float(poses[2][-1]))
    ts[0, 0] = float(poses[2][0])            # corner case of :52-54
    ts[0, 1] = float(poses[2][-1])
    pr, rr = ref.trajectory(ts)
    po, ro = ora.trajectory(ts)
    _close(po, pr)
    _close(ro, rr)
    px = torch.rand(40, 2, generator=g) * 200
    o_r, d_r = ref.nerf.pixel_params_to_ray(ref.train_intrinsics_inv, px, pr, rr)
    o_o, d_o = path_ref.NeRF.pixel_params_to_ray(ora.train_intrinsics_inv, px, po, ro)
    _close(d_o, d_r)


def test_trajectory_time_gradient_closed_form(pair):
    """The closed-form dL/dt per pose interval (what den_rays_from_trajectory_bwd evaluates, restated in
    oracle/path_ref.py) against torch autograd through the REFERENCE'S trajectory + pinhole code."""
    cfg, ref, ora, poses = pair
    g = torch.Generator().manual_seed(7)
    ts = (torch.rand(2, 300, generator=g, dtype=torch.float64) * float(poses[2][-1])).requires_grad_(True)
    px = torch.rand(300, 2, generator=g) * 200
    w_o, w_d = torch.randn(2, 300, 3, generator=g), torch.randn(2, 300, 3, generator=g)
    pr, rr = ref.trajectory(ts)
    o_r, d_r = ref.nerf.pixel_params_to_ray(ref.train_intrinsics_inv, px, pr, rr)
    ((o_r * w_o).sum() + (d_r * w_d).sum()).backward()
    closed = path_ref.trajectory_time_gradient(ora.trajectory, ts.detach(), d_r.detach(), w_o, w_d)
    scale = ts.grad.abs().max()
    assert scale > 0
    assert (closed - ts.grad).abs().max() < 2e-5 * scale, ((closed - ts.grad).abs().max(), scale)


@pytest.mark.parametrize("training", [False, True])
def test_nerf_render(pair, training):
    cfg, ref, ora, poses = pair
    ref.nerf.train()
    ora.nerf.train()
    torch.manual_seed(3)
    ref.nerf.update_occ_grid(0, ref.trajectory.T_wc_position)
    torch.manual_seed(3)
    ora.nerf.update_occ_grid(0, ora.trajectory.T_wc_position)
    assert torch.equal(ref.nerf.occupancy_grid.binary, ora.nerf.occupancy_grid.binary)
    ref.nerf.train(training)
    ora.nerf.train(training)
    g = torch.Generator().manual_seed(2)
    ts = torch.rand(300, generator=g, dtype=torch.float64) * float(poses[2][-1])
    px = torch.rand(300, 2, generator=g) * 250
    pos, rot = ora.trajectory(ts)
    o, d = path_ref.NeRF.pixel_params_to_ray(ora.train_intrinsics_inv, px, pos, rot)
    torch.manual_seed(4)
    out_r = ref.nerf(o, d)
    torch.manual_seed(4)
    out_o = ora.nerf(o, d)
    for a, b in zip(out_o[:3], out_r[:3]):
        _close(a, b, 1e-5)
    assert out_o[3] == out_r[3]              # mean samples per ray: identical sample sets
    ref.nerf.zero_grad()
    ora.nerf.zero_grad()
    out_r[0].log().sum().backward()
    out_o[0].log().sum().backward()
    gr, go = _scene.flat_named_grads(ref.nerf), _scene.flat_named_grads(ora.nerf)
    assert gr.keys() == go.keys()
    for k in gr:
        _close(go[k], gr[k], 1e-4)


@pytest.mark.parametrize("reset", [True, False])
def test_pixel_bandwidth(pair, reset):
    cfg, ref, ora, poses = pair
    g = torch.Generator().manual_seed(5)
    n, S = 50, 8
    out_ts = (30e6 + torch.rand(n, generator=g, dtype=torch.float64) * 100e6)
    gen = torch.full((S - 1, n), 0.5, dtype=torch.float64)
    base = torch.rand(S, n, generator=g) * 0.8 + 0.01

    def make_fn(leaf):
        def fn(ts):
            return leaf * (1 + 0.1 * torch.sin(ts.float() * 1e-7)), torch.tensor(1.0), 1.0, \
                torch.ones_like(ts, dtype=torch.bool)
        return fn

    leaf_r = base.clone().requires_grad_(True)
    leaf_o = base.clone().requires_grad_(True)
    if not reset:      # a reset call must precede (state of :419-423)
        ref.pixel_bandwidth(gen, out_ts - 5e6, make_fn(leaf_r.detach()), True)
        ora.pixel_bandwidth(gen, out_ts - 5e6, make_fn(leaf_o.detach()), True)
    y_r, _ = ref.pixel_bandwidth(gen, out_ts, make_fn(leaf_r), reset)
    y_o, _ = ora.pixel_bandwidth(gen, out_ts, make_fn(leaf_o), reset)
    _close(y_o, y_r, 2e-5)
    for m in (ref.pixel_bandwidth, ora.pixel_bandwidth):
        m.zero_grad()
    y_r.sum().backward()
    y_o.sum().backward()
    _close(leaf_o.grad, leaf_r.grad, 2e-3)   # fp32 expm/solve: the reference's own noise floor
    gr = _scene.flat_named_grads(ref.pixel_bandwidth)
    go = _scene.flat_named_grads(ora.pixel_bandwidth)
    assert gr.keys() == go.keys() and len(gr) == 6


@pytest.mark.parametrize("pb_on", [True, False])
def test_training_step(pb_on):
    """The whole hot path: the reference's training_step (models/deblur_e_nerf.py:396-586)
    vs oracle EventRenderer.training_step — loss, loss terms and every parameter gradient."""
    cfg = _scene.scene_config("synthetic", occ_resolution=32, small=True)
    S = 8
    ref, poses = _scene.build_reference_renderer(cfg, S, pixel_bandwidth=pb_on)
    ora, _ = _scene.build_oracle_renderer(cfg, S, pixel_bandwidth=pb_on)
    names = ["nerf", "contrast_threshold", "refractory_period"] + (
        ["pixel_bandwidth"] if pb_on else [])
    for name in names:
        _scene.copy_params(getattr(ref, name), getattr(ora, name))
    event, normalized = _scene.make_batch(cfg, poses, 96, S, seed=7, pixel_bandwidth=pb_on)

    ref.train()
    ora.train()
    torch.manual_seed(11)
    loss_r = ref.training_step(_scene.reference_batch(event, normalized), 0)
    torch.manual_seed(11)
    ora.nerf.update_occ_grid(0, ora.trajectory.T_wc_position)
    loss_o, terms_o, mean_samples = ora.training_step(event, normalized)
    assert torch.equal(ref.nerf.occupancy_grid.binary, ora.nerf.occupancy_grid.binary)
    _close(loss_o, loss_r, 1e-5)
    for key, val in terms_o.items():
        _close(val, ref.logged[f"train/{key}"], 1e-5)
    _close(torch.tensor(mean_samples), torch.as_tensor(ref.logged["train/mean_num_samples_per_ray"]),
           1e-6)
    ref.zero_grad()
    ora.zero_grad()
    loss_r.backward()
    loss_o.backward()
    gr, go = _scene.flat_named_grads(ref), _scene.flat_named_grads(ora)
    assert set(gr.keys()) == set(go.keys()), set(gr.keys()) ^ set(go.keys())
    for k in gr:
        tol = 5e-3 if "pixel_bandwidth" in k or "refractory" in k else 2e-4
        _close(go[k], gr[k], tol)


def test_event_batch_producer_reproduces_the_reference_samplers():
    """deblur_e_nerf_b200.data.EventBatchProducer on CPU tensors + a CPU generator == the reference's
    IterableMapDataset (utils/datasets.py:19-32) + UniformSampler / TriangularSampler /
    DiracDeltaSampler (data/samplers.py) drawing from one shared generator in the dataloader's order
    (event batch first, then the normalised samplers: data/datamodule.py:151-247), bit for bit."""
    from deblur_e_nerf_b200.data import EventBatchProducer
    ref_datasets = ref_shim.load("utils.datasets")
    ref_samplers = ref_shim.load("data.samplers")
    g0 = torch.Generator().manual_seed(3)
    n_events, batch, S = 5000, 257, 8
    events = {
        "position": torch.rand(n_events, 2, generator=g0) * 300,
        "start_ts": torch.randint(0, 10 ** 9, (n_events,), generator=g0),
        "end_ts": torch.randint(10 ** 9, 2 * 10 ** 9, (n_events,), generator=g0),
        "num_pos": torch.randint(0, 2, (n_events,), generator=g0),
        "num_neg": torch.randint(0, 2, (n_events,), generator=g0),
    }

    class MapEvents(torch.utils.data.Dataset):           # Event.__getitem__: index every array
        def __len__(self):
            return n_events

        def __getitem__(self, index):
            return {k: v[index] for k, v in events.items()}

    gen = torch.Generator().manual_seed(11)
    ref_events = iter(ref_datasets.IterableMapDataset(MapEvents(), batch, generator=gen))
    f64 = torch.float64
    ref_norm = iter(ref_datasets.JoinDataset(
        [ref_samplers.DiracDeltaSampler(center=1, size=batch, dtype=f64),
         ref_samplers.UniformSampler(low=0, high=1, size=batch, dtype=f64, generator=gen),
         ref_samplers.TriangularSampler(low=0, high=1, size=batch, mode=0, dtype=f64, generator=gen),
         ref_samplers.UniformSampler(low=0, high=1, size=batch, dtype=f64, generator=gen),
         ref_samplers.DiracDeltaSampler(center=0.5, size=(S - 1, batch), dtype=f64)],
        ["ts_diff", "diff_start_ts", "ts_subdiff", "subdiff_start_ts", "interval_gen"]))
    ours = EventBatchProducer(events, batch, it_sample_size=S, device="cpu", seed=11)
    for _ in range(3):
        ev, nm = next(ref_events), next(ref_norm)
        out = ours.next_batch()
        assert set(out["event"]) == set(ev) and set(out["normalized"]) == set(nm)
        for k in ev:
            assert out["event"][k].dtype == ev[k].dtype and torch.equal(out["event"][k], ev[k]), k
        for k in nm:
            assert out["normalized"][k].dtype == nm[k].dtype and torch.equal(out["normalized"][k], nm[k]), k


def _eval_images(seed, B=3, H=24, W=32):
    g = torch.Generator().manual_seed(seed)
    target = (torch.rand(B, H, W, generator=g) * 0.9 + 0.05)
    exposure = torch.tensor([1, 2, 4][:B], dtype=torch.int64)
    gain = torch.tensor([1.0, 1.5, 0.75][:B])
    norm = gain * exposure / (gain * exposure).mean()
    # an affinely ambiguous (in log space) prediction of the exposure-normalised scene + a little noise
    scene = target / norm.view(-1, 1, 1)
    pred = (0.7 * scene.pow(1.3)) * torch.exp(0.02 * torch.randn(B, H, W, generator=g))
    return pred.float(), target.float(), exposure, gain


@pytest.mark.parametrize("colour,per_channel", [(False, False), (True, True), (True, False)],
                         ids=["mono", "bayer_per_channel_scale", "bayer_shared_scale"])
def test_eval_post_processing_matches_reference(pair, colour, per_channel):
    """oracle/eval_ref.py (gain-exposure normalisation, float64 log-space affine least squares, L1 /
    PSNR) against the reference's OWN evaluation_epoch_end (models/deblur_e_nerf.py:661-969) run under
    the shim with `correction.black_level_offset: false` (the refinement needs pypose, absent here):
    mono images, and the colour images of a Bayer sensor with a log-intensity scale per channel or one
    shared by the channels (`correction.per_channel_log_it_scale`, :753-766)."""
    import easydict
    from oracle import eval_ref
    _, ref, _, _ = pair
    pred, target, exposure, gain = _eval_images(0)
    if colour:
        # three channels with their own gamma / gain, so that the two scale variants differ
        imgs = [_eval_images(10 + c) for c in range(3)]
        target = torch.stack([im[1] for im in imgs], dim=1)
        pred = torch.stack([(0.5 + 0.2 * c) * imgs[c][0].pow(1.0 + 0.1 * c) for c in range(3)], dim=1)
    recorded = {}

    class _Metric:
        def init_batch_metric(self):
            return easydict.EasyDict(l1=[], psnr=[])

        def compute(self, p, t, min_target_val, max_target_val):
            recorded.setdefault("pred", []).append(p.clone())
            mse = ((p - t) ** 2).mean()
            return easydict.EasyDict(l1=torch.nn.functional.l1_loss(p, t),
                                     psnr=10 * torch.log10((max_target_val - min_target_val) ** 2 / mse))

    class _Trainer:
        log_dir = "/tmp"
        is_global_zero = True
        sanity_checking = False

    object.__setattr__(ref, "trainer", _Trainer())
    ref.correction = easydict.EasyDict(per_channel_log_it_scale=per_channel, black_level_offset=False)
    ref.has_bayer_filter = colour
    ref.metric = _Metric()
    ref.logger = None
    type(ref).current_epoch = 0
    type(ref).device = torch.device("cpu")
    ref.eval_save_pred_intensity_img = False
    ref.all_gather = lambda outs: [{k: v[None] for k, v in o.items()} for o in outs]
    logged = {}
    ref.log = lambda name, value, **kw: logged.__setitem__(name, value)
    outputs = [{"sample_id": torch.tensor([ord(c) for c in f"{i:04d}    "]),
                "pred_intensity_img": pred[i], "target_intensity_img": target[i],
                "exposure_time": exposure[i], "gain": gain[i]} for i in range(len(pred))]
    stage = easydict.EasyDict(name="val", min_normalized_pixel_value=0.0, max_normalized_pixel_value=1.0)
    ref.evaluation_epoch_end(outputs, stage)

    res = eval_ref.evaluate(pred if colour else pred[:, None], target if colour else target[:, None],
                            exposure, gain, 0.0, 1.0, black_level_offset=False,
                            per_channel_scale=per_channel or not colour)
    ref_pred = torch.stack(recorded["pred"])
    if ref_pred.dim() == 3:
        ref_pred = ref_pred[:, None]
    _close(res["pred"], ref_pred, 1e-6)
    assert abs(res["l1"] - float(logged["val/l1"])) <= 1e-6 * abs(float(logged["val/l1"]))
    assert abs(res["psnr"] - float(logged["val/psnr"])) <= 1e-5 * abs(float(logged["val/psnr"]))
    if not colour:
        # the fit recovers the construction: gamma 1 / 1.3 in log space
        assert abs(float(res["affine"][0, 0]) - 1 / 1.3) < 0.02
    elif not per_channel:
        assert float(res["affine"][:, 0].max() - res["affine"][:, 0].min()) == 0.0      # one shared scale
    ref.has_bayer_filter = False


def test_eval_lm_refinement_reaches_the_least_squares_minimum():
    """The offset-gamma refinement (black_level_offset: true) — unpinned against pypose, so checked
    against what it must compute: at convergence the gradient J^T r vanishes and the error is no larger
    than that of an independent float64 optimisation."""
    from oracle import eval_ref
    pred, target, exposure, gain = _eval_images(1)
    target = target + 0.03                                   # a black-level offset the affine fit cannot absorb
    res = eval_ref.evaluate(pred[:, None], target[:, None], exposure, gain, 0.0, 1.1, black_level_offset=True)
    norm = eval_ref.normalized_gain(gain, exposure)
    _, fitted, _ = eval_ref.affine_log_correction(pred[:, None], target[:, None], norm)
    x = fitted.exp()
    p = res["correction"].clone().requires_grad_(True)
    g = norm.double().view(-1, 1, 1, 1)
    f = g * (p[:, 0].view(1, -1, 1, 1) * x.pow(p[:, 1].view(1, -1, 1, 1)) - p[:, 2].view(1, -1, 1, 1))
    loss = ((f - target[:, None].double()) ** 2).sum()
    loss.backward()
    assert float(p.grad.abs().max()) < 1e-6 * float(loss) + 1e-9
    no_refine = eval_ref.evaluate(pred[:, None], target[:, None], exposure, gain, 0.0, 1.1,
                                  black_level_offset=False)
    assert res["l1"] < no_refine["l1"]


@pytest.mark.parametrize("shared_gamma", [False, True], ids=["gamma_per_channel", "gamma_shared"])
def test_offset_gamma_correction_model_matches_reference(shared_gamma):
    """oracle/eval_ref._Correction (forward and the joint Jacobian, parameter order [scale, gamma, offset])
    against the reference's OWN models/offset_gamma_correction.py (pure torch, loaded by path) for colour
    images with a gamma per channel and with ONE gamma shared by the channels (the shape
    `per_channel_log_it_scale: false` gives it, models/deblur_e_nerf.py:185-197); then the refinement on a
    shared-gamma problem: equal gammas, vanishing joint gradient."""
    import importlib.util
    from oracle import eval_ref
    spec = importlib.util.spec_from_file_location(
        "_ref_offset_gamma_correction", ref_shim.REFERENCE_ROOT + "/deblur_e_nerf/models/offset_gamma_correction.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = torch.Generator().manual_seed(4)
    B, C, H, W = 2, 3, 5, 7
    x = torch.rand(B, C, H, W, generator=g, dtype=torch.float64) + 0.1
    gain = torch.tensor([0.8, 1.2], dtype=torch.float64)
    scale = torch.rand(C, generator=g, dtype=torch.float64) + 0.5
    gamma = torch.rand(1 if shared_gamma else C, generator=g, dtype=torch.float64) + 0.7
    offset = torch.rand(C, generator=g, dtype=torch.float64) * 0.1
    ref = mod.OffsetGammaCorrection(gain.view(-1, 1, 1, 1, 1), scale.view(-1, 1, 1, 1),
                                    gamma.view(-1, 1, 1, 1), offset.view(-1, 1, 1, 1))
    ora = eval_ref._Correction(gain, scale, gamma, offset)
    with torch.no_grad():
        _close(ora.forward(x), ref(x.unsqueeze(-1)).squeeze(-1), 1e-14)
        (jac,) = ref.jacobian(x.unsqueeze(-1))
        assert jac.shape == (B * C * H * W, 2 * C + gamma.numel())
        _close(ora.jacobian(x), jac, 1e-14)

    if shared_gamma:
        imgs = [_eval_images(30 + c) for c in range(3)]
        exposure, gain32 = imgs[0][2], imgs[0][3]
        target = torch.stack([im[1] for im in imgs], dim=1) + 0.03
        pred = torch.stack([(0.5 + 0.2 * c) * imgs[c][0].pow(1.2) for c in range(3)], dim=1)
        res = eval_ref.evaluate(pred, target, exposure, gain32, 0.0, 1.1, black_level_offset=True,
                                per_channel_scale=False)
        corr = res["correction"]
        assert float(corr[:, 1].max() - corr[:, 1].min()) == 0.0            # one gamma
        norm = eval_ref.normalized_gain(gain32, exposure)
        _, fitted, _ = eval_ref.affine_log_correction(pred, target, norm, per_channel_scale=False)
        theta = torch.cat((corr[:, 0], corr[:1, 1], corr[:, 2])).clone().requires_grad_(True)
        gn = norm.double().view(-1, 1, 1, 1)
        f = gn * (theta[:3].view(1, 3, 1, 1) * fitted.exp().pow(theta[3]) - theta[4:].view(1, 3, 1, 1))
        loss = ((f - target.double()) ** 2).sum()
        loss.backward()
        assert float(theta.grad.abs().max()) < 1e-6 * float(loss) + 1e-9


def test_raw_event_queueing_matches_reference(tmp_path):
    """oracle/events_ref.py against the reference's OWN Event.queue_raw_events /
    extract_max_refractory_period / colorize_events (data/datasets.py:131-324) on a fresh seeded stream (the
    committed golden holds four more): the literal loops and the vectorised forms, bit for bit."""
    import numpy as np
    from oracle import events_ref
    sys_path = __import__("sys").path
    sys_path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__file__), "golden"))
    try:
        from make_golden import raw_event_stream
    finally:
        sys_path.pop(0)
    ds = ref_shim.load("data.datasets")
    height, width = 11, 8
    position, timestamp, polarity = raw_event_stream(seed=77, n=6000, height=height, width=width, hot_fraction=0.5)
    calib = {"img_height": np.array(height, dtype=np.uint16), "img_width": np.array(width, dtype=np.uint16),
             "bayer_pattern": np.array("BGGR")}
    np.savez(tmp_path / ds.Event.RAW_EVENTS_FILENAME, position=position, timestamp=timestamp, polarity=polarity)
    want = ds.Event.colorize_events(ds.Event.queue_raw_events(str(tmp_path), calib), calib)
    want_refractory = ds.Event.extract_max_refractory_period(
        {"position": position, "timestamp": timestamp, "polarity": polarity}, calib)
    for fn in (events_ref.queue_raw_events_loop, events_ref.queue_raw_events):
        got = fn(position, timestamp, polarity, height, width)
        for key, value in got.items():
            assert np.array_equal(value, want[key].numpy()) and value.dtype == want[key].numpy().dtype, (fn.__name__, key)
    assert np.array_equal(events_ref.colorize_events(want["position"].numpy(), "BGGR"), want["channel_idx"].numpy())
    for fn in (events_ref.max_refractory_period_loop, events_ref.max_refractory_period):
        assert float(fn(position, timestamp, height, width)) == float(want_refractory)


@pytest.mark.parametrize("channels,alpha,bayer,seed", [(3, False, "", None), (4, True, "", 5), (4, False, "RGGB", None),
                                                      (3, False, "GBRG", 2), (4, True, "BGGR", None)])
def test_posed_views_match_reference_posed_image(tmp_path, channels, alpha, bayer, seed):
    """`views.PosedViews` (the input side of Trainer.test) against the reference's OWN `PosedImage`
    (data/datasets.py:377-713) on a tiny on-disk dataset: BGR / BGRA renders, with and without compositing over
    white, mono and Bayer sensors, with a permutation seed — every tensor, the intrinsics and the pixel range."""
    import numpy as np
    import _dataset
    from deblur_e_nerf_b200 import synthetic, views
    ds = ref_shim.load("data.datasets")
    root = str(tmp_path)
    _dataset.write(root, dict(synthetic.CONFIGS["synthetic"]), n_events=8, n_views=3, size=(10, 14), channels=channels)
    calib = dict(np.load(tmp_path / "camera_calibration.npz"))
    calib["bayer_pattern"] = np.array(bayer)
    np.savez(tmp_path / "camera_calibration.npz", **calib)
    if seed == 2:               # real captures carry an exposure time and a gain per frame, and explicit intrinsics
        import json
        path = tmp_path / "views" / "transforms_test.json"
        meta = json.loads(path.read_text())
        meta.pop("camera_angle_x")
        meta["intrinsics"] = [[20.5, 0, 7.0], [0, 20.25, 5.0], [0, 0, 1]]
        meta["bit_depth"] = 8
        for k, frame in enumerate(meta["frames"]):
            frame["exposure_time"], frame["gain"] = 1000 * (k + 1), 1.0 + 0.5 * k
        path.write_text(json.dumps(meta))
    for stage in ("val", "test"):
        want = ds.PosedImage(root, stage, seed, alpha_over_white_bg=alpha)
        got = views.PosedViews(root, stage, seed, alpha_over_white_bg=alpha)
        assert len(got) == len(want) == 3
        assert got.img.shape == want.posed_imgs.img.shape and got.img.dtype == want.posed_imgs.img.dtype
        _close(got.img, want.posed_imgs.img, 1e-6)
        keys = ["sample_id", "T_wc_position", "T_wc_orientation", "intrinsics"]
        if seed == 2 and stage == "test":
            keys += ["exposure_time", "gain"]
        else:
            assert got.exposure_time is None and got.gain is None and "gain" not in want.posed_imgs
        for key in keys:
            ref_value = want.posed_imgs[key]
            assert getattr(got, key).dtype == ref_value.dtype and torch.equal(getattr(got, key), ref_value), key
        assert got.min_normalized_pixel_value == want.min_normalized_pixel_value
        assert got.max_normalized_pixel_value == want.max_normalized_pixel_value
        item, ref_item = got[1], want[1]
        assert set(item) == set(ref_item) and all(torch.allclose(item[k].double(), ref_item[k].double(), atol=1e-6) for k in item)
        args = got.test_arguments()
        assert torch.allclose(args["intrinsics_inv"] @ got.intrinsics, torch.eye(3), atol=1e-5)


@pytest.mark.parametrize("model,params", [("plumb_bob", []), ("plumb_bob", [-0.31, 0.12, 0.0007, -0.0004]),
                                          ("equidistant", [-0.02, 0.006, -0.004, 0.0009])])
def test_colorize_and_undistort_events_match_reference(model, params):
    """The two elementwise / OpenCV steps of the event preprocessing that run without a kernel —
    `events.colorize_events`, `events.undistort_events` — against the reference's OWN classmethods
    (data/datasets.py:278-365): Bayer channel indices bit for bit, undistorted positions (plumb_bob and
    equidistant models, and the distortion-free cast) to float32 rounding."""
    import numpy as np
    from deblur_e_nerf_b200 import events
    ds = ref_shim.load("data.datasets")
    import easydict                                        # registered by the shim
    g = torch.Generator().manual_seed(6)
    n, height, width = 500, 260, 346
    position = torch.stack([torch.randint(0, width, (n,), generator=g), torch.randint(0, height, (n,), generator=g)], dim=1)
    calib = {"bayer_pattern": np.array("GRBG"), "distortion_model": np.array(model),
             "distortion_params": np.array(params, dtype=np.float64),
             "intrinsics": np.array([[300.0, 0, 173.0], [0, 300.0, 130.0], [0, 0, 1]])}
    base = {"position": position, "start_ts": torch.arange(n), "end_ts": torch.arange(n) + 1,
            "num_pos": torch.ones(n, dtype=torch.int64), "num_neg": torch.zeros(n, dtype=torch.int64)}
    want = ds.Event.undistort_events(ds.Event.colorize_events(easydict.EasyDict({k: v.clone() for k, v in base.items()}),
                                                              calib), calib)
    got = events.undistort_events(events.colorize_events({k: v.clone() for k, v in base.items()}, calib), calib)
    assert set(got) == set(want)
    assert got["channel_idx"].dtype == want["channel_idx"].dtype and torch.equal(got["channel_idx"], want["channel_idx"])
    assert got["position"].dtype == want["position"].dtype == torch.float32
    assert torch.allclose(got["position"], want["position"], rtol=0, atol=1e-4)
    if not params:
        assert torch.equal(got["position"], position.float())
    with pytest.raises(NotImplementedError):
        events.undistort_events({"position": position.clone()}, {**calib, "distortion_model": np.array("fov"),
                                                                 "distortion_params": np.ones(4)})
