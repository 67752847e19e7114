"""GPU parity of the fused pixel-bandwidth filter + event loss (`den_lpf_loss_fwd/_bwd`, SURVEY.md
§8(a) A15-A18) against the oracle's PixelBandwidth + EventLoss evaluated in fp64 (the truth) and in
fp32 (what the reference computes): the loss terms, and the gradients w.r.t. the rendered
intensities, the six pixel-bandwidth parameters, the target, the normaliser and the output
timestamps (the refractory-period path through the reset decay)."""

import pytest
import torch

from oracle import path_ref

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-30)).item()


def _inputs(S, n, seed, invalid_fraction):
    g = torch.Generator().manual_seed(seed)
    base = 40e6 + torch.rand(n, generator=g, dtype=torch.float64) * 100e6
    # four output timestamps per event: diff (start, end), subdiff (start, end) inside it; a few
    # microseconds apart so that the reset decay exp(-omega_diff dt) (1 / omega_diff = 1.9 us) is
    # neither 1 nor 0 and its gradient w.r.t. the timestamps is a real number
    span = 2e3 + torch.rand(n, generator=g, dtype=torch.float64) * 8e3
    out_ts = torch.stack([base, base + span, base + 0.2 * span, base + 0.7 * span])
    gen = torch.rand(S - 1, n, generator=g, dtype=torch.float64)
    intensity = torch.exp(torch.randn(4, S, n, generator=g) * 1.2 - 2.0).clamp(1e-3, 5.0)
    target = torch.randn(n, generator=g) * 2.0           # straddles the Huber knee at |d| = 1
    valid = torch.rand(2, n, generator=g) >= invalid_fraction
    return out_ts, gen, intensity, target, valid


def _oracle(dtype, calib, out_ts, gen, intensity, target, valid, mean_ct, kinds, normalize):
    pb = path_ref.PixelBandwidth(calib, 0, 21, 0.95).to(dtype)
    leaf_i = intensity.to(dtype).clone().requires_grad_(True)
    leaf_ts = out_ts.clone().requires_grad_(True)
    leaf_t = target.to(dtype).clone().requires_grad_(True)
    leaf_k = torch.tensor(mean_ct, dtype=dtype, requires_grad=True)
    outs = []
    for k in range(4):
        fn = lambda ts, _k=k: (leaf_i[_k], None)         # noqa: E731
        y, _ = pb(gen, leaf_ts[k], fn, reset_diff=(k == 0))
        outs.append(y)
    terms = []
    for p, name in enumerate(("log_intensity_diff", "log_intensity_tv")):
        kk = leaf_k if normalize[p] else 1
        pred = outs[2 * p + 1] - outs[2 * p]
        tgt = leaf_t / kk if p == 0 else torch.zeros_like(pred)
        err = path_ref.ERROR_FNS[kinds[p]](pred / kk, tgt.to(pred.dtype))
        terms.append(err[valid[p]].mean())
    loss = terms[0] * 1.0 + terms[1] * 0.37
    loss.backward()
    grads = {k: v.grad.detach().double() for k, v in pb.named_parameters()}
    return ([t.detach().double() for t in terms], torch.stack(outs).detach().double(), leaf_i.grad.double(),
            leaf_ts.grad.double(), leaf_t.grad.double(), leaf_k.grad.double(), grads)


@pytest.mark.parametrize("S,kinds,normalize", [(30, ("huber", "l1"), (True, True)),
                                               (8, ("mse", "huber"), (True, False)),
                                               (32, ("l1", "mse"), (False, True))])
def test_lpf_loss_matches_fp64_oracle(den_lib, cuda, S, kinds, normalize):
    from deblur_e_nerf_b200 import ops, synthetic
    from deblur_e_nerf_b200 import pixel_bandwidth as pb_mod
    calib = synthetic.calibration()
    n = 257
    mean_ct = 0.259
    out_ts, gen, intensity, target, valid = _inputs(S, n, S, invalid_fraction=0.2)
    truth = _oracle(torch.float64, calib, out_ts, gen, intensity, target, valid, mean_ct, kinds, normalize)
    ref32 = _oracle(torch.float32, calib, out_ts, gen, intensity, target, valid, mean_ct, kinds, normalize)

    pb = pb_mod.PixelBandwidth(calib, 0, 21, dict(max_sample_lifetime=0.95)).to(cuda)
    li = intensity.to(cuda).clone().requires_grad_(True)
    lts = out_ts.to(cuda).clone().requires_grad_(True)
    lt = target.to(cuda).clone().requires_grad_(True)
    lk = torch.tensor(mean_ct, device=cuda, requires_grad=True)
    life = pb.sample_lifetimes(gen.to(cuda))
    sample_ts = lts[:, None, :] - life[None]
    sample_dt = sample_ts.detach().diff(dim=1).float()
    inv_k = torch.stack([1.0 / lk if normalize[p] else torch.ones_like(lk) for p in range(2)])
    tgt = torch.stack([(lt / lk) if normalize[0] else lt, torch.zeros_like(lt)])
    terms, log_it, counts = ops.lpf_loss(li, sample_dt, pb.coefficients(), lts - lts[0], tgt, inv_k,
                                         valid.to(cuda), kinds, (True, False), has_reset=True)
    (terms[0] * 1.0 + terms[1] * 0.37).backward()
    assert counts.tolist() == [int(valid[0].sum()), int(valid[1].sum())]
    t_terms, t_out, t_gi, t_gts, t_gt, t_gk, t_gp = truth
    r_terms, r_out, r_gi, r_gts, r_gt, r_gk, r_gp = ref32
    assert _rel(log_it, t_out) < 1e-5
    for p in range(2):
        assert _rel(terms[p], t_terms[p]) < 1e-5, (p, float(terms[p]), float(t_terms[p]))
    assert _rel(li.grad, t_gi) < 1e-4
    assert _rel(lt.grad, t_gt) < 1e-5
    assert _rel(lk.grad, t_gk) < 1e-4
    assert _rel(lts.grad, t_gts) < 1e-4          # only through the reset decay exp(-w (ts_k - ts_0))
    prod_gp = {k: v.grad.detach().cpu().double() for k, v in pb.named_parameters()}
    for k in t_gp:
        assert _rel(prod_gp[k], t_gp[k]) < 1e-4, (k, prod_gp[k], t_gp[k])
    # no further from the fp64 truth than the reference's own fp32 evaluation is
    assert _rel(li.grad, t_gi) <= max(_rel(r_gi, t_gi), 1e-5) * 1.5
    for k in t_gp:
        assert _rel(prod_gp[k], t_gp[k]) <= max(_rel(r_gp[k], t_gp[k]), 1e-5) * 1.5, k


def test_lpf_loss_ignores_non_finite_errors_of_invalid_events(den_lib, cuda):
    """loss_metric/loss.py:80,94 index `err[is_valid]`: whatever an invalid event holds (here a NaN
    target) must reach neither the mean nor any gradient."""
    from deblur_e_nerf_b200 import ops, synthetic
    from deblur_e_nerf_b200 import pixel_bandwidth as pb_mod
    calib = synthetic.calibration()
    S, n = 8, 64
    out_ts, gen, intensity, target, valid = _inputs(S, n, 1, invalid_fraction=0.0)
    valid[0, 5] = False
    target[5] = float("nan")
    pb = pb_mod.PixelBandwidth(calib, 0, 21, dict(max_sample_lifetime=0.95)).to(cuda)
    li = intensity.to(cuda).clone().requires_grad_(True)
    lts = out_ts.to(cuda)
    life = pb.sample_lifetimes(gen.to(cuda))
    sample_dt = (lts[:, None, :] - life[None]).diff(dim=1).float()
    tgt = torch.stack([target.to(cuda), torch.zeros(n, device=cuda)])
    inv_k = torch.ones(2, device=cuda)
    terms, _, counts = ops.lpf_loss(li, sample_dt, pb.coefficients(), lts - lts[0], tgt, inv_k,
                                    valid.to(cuda), ("huber", "l1"), (True, False))
    terms.sum().backward()
    assert counts.tolist() == [n - 1, n]
    assert torch.isfinite(terms).all() and torch.isfinite(li.grad).all()
    assert all(torch.isfinite(p.grad).all() for p in pb.parameters())
    # with no valid event at all the mean of an empty selection is NaN, like the reference's
    none = torch.zeros(2, n, dtype=torch.bool, device=cuda)
    terms0, _, counts0 = ops.lpf_loss(li.detach(), sample_dt, pb.coefficients(), lts - lts[0], tgt, inv_k,
                                      none, ("huber", "l1"), (True, False))
    assert counts0.tolist() == [0, 0] and torch.isnan(terms0).all()


@pytest.mark.parametrize("pb_free", [False, True])
def test_fused_step_equals_unfused_step(den_lib, cuda, pb_free):
    """EventRenderer.training_step with `fuse_lpf_loss` (one filter + loss kernel per direction)
    against the same step with a filter launch per request and Loss.compute in torch."""
    import _scene
    golden = _scene.load_golden("training_step_pb_on")
    cfg = _scene.scene_config("synthetic", occ_resolution=32, small=True)
    results = []
    for fused in (True, False):
        model, poses = _scene.build_product_renderer(cfg, cuda, 8, pixel_bandwidth=True)
        for name in ("nerf", "contrast_threshold", "refractory_period", "pixel_bandwidth"):
            _scene.load_golden_state(getattr(model, name), golden, name, cuda)
        if not pb_free:
            model.pixel_bandwidth.requires_grad_(False)
            model.refractory_period.requires_grad_(False)
        batch = {"event": _scene.golden_section(golden, "event", cuda),
                 "normalized": _scene.golden_section(golden, "normalized", cuda)}
        jitters = [v for _, v in sorted(_scene.golden_section(golden, "jitter", cuda).items(),
                                        key=lambda kv: int(kv[0]))]
        model.train()
        model.fuse_lpf_loss = fused
        model.nerf.update_occ_grid = lambda *a, **k: None
        loss = model.training_step(batch, 0, 0, jitters=jitters)
        loss.backward()
        results.append((loss.detach(), dict(model.logged), _scene.flat_named_grads(model)))
    (la, ga, gra), (lb, gb, grb) = results
    assert _rel(la, lb) < 1e-6
    for key in ("train/log_intensity_diff", "train/log_intensity_tv"):
        assert _rel(ga[key], gb[key]) < 1e-6
    assert set(gra) == set(grb)
    for key in gra:
        assert _rel(gra[key], grb[key]) < 2e-4, key
