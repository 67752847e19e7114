"""GPU: the training step WITHOUT host read-backs (capacity-sized sample buffers, device-side
counts: `NeRF.render_chunk_sync_free`) and its CUDA-graph replay (`graph_step.GraphedStep`) must
give the loss, the logged terms and every gradient of the synchronising path — they run the same
kernels on the same samples; only where the counts live differs."""

import pytest
import torch

import _scene

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-30)).item()


def _setup(cuda, scene, pb_on, dense=False):
    golden = _scene.load_golden({"synthetic": "training_step_pb_on" if pb_on else "training_step_pb_off",
                                 "eds": "training_step_eds"}[scene])
    cfg = _scene.scene_config(scene, occ_resolution=32, small=True)
    model, poses = _scene.build_product_renderer(cfg, cuda, 8, pixel_bandwidth=pb_on)
    for name in ["nerf", "contrast_threshold", "refractory_period"] + (["pixel_bandwidth"] if pb_on else []):
        _scene.load_golden_state(getattr(model, name), golden, name, cuda)
    if dense:       # a dense field: the visibility pre-pass culls most samples (compaction + row map)
        with torch.no_grad():
            model.nerf.radiance_field.mlp_base[1].output_layer.bias[0] += 5.0
    pre = "" if scene == "synthetic" else "/0"
    batch = {"event": _scene.golden_section(golden, "event" + pre, cuda),
             "normalized": _scene.golden_section(golden, "normalized" + pre, cuda)}
    jitters = [v for _, v in sorted(_scene.golden_section(golden, "jitter" + pre, cuda).items(),
                                    key=lambda kv: int(kv[0]))]
    model.train()
    model.nerf.update_occ_grid = lambda *a, **k: None
    return model, batch, jitters


@pytest.mark.parametrize("scene,pb_on,dense", [("synthetic", False, False), ("synthetic", True, False),
                                               ("synthetic", True, True), ("eds", True, False)])
def test_sync_free_step_equals_synchronising_step(den_lib, cuda, scene, pb_on, dense):
    results = []
    for sync_free in (False, True):
        model, batch, jitters = _setup(cuda, scene, pb_on, dense)
        model.nerf.sync_free = sync_free
        # the first call always synchronises (it teaches the capacity estimate); the second is the one
        model.training_step(batch, 0, 0, jitters=jitters)
        model.zero_grad()
        loss = model.training_step(batch, 0, 0, jitters=jitters)
        used = model.nerf._stats is not None
        assert used == sync_free
        loss.backward()
        torch.cuda.synchronize()
        assert model.nerf.overflow_count == 0
        results.append((loss.detach(), dict(model.logged), _scene.flat_named_grads(model)))
    (la, ga, gra), (lb, gb, grb) = results
    assert _rel(lb, la) < 1e-6
    assert abs(float(gb["train/mean_num_samples_per_ray"]) - float(ga["train/mean_num_samples_per_ray"])) < 1e-6 * 300
    if dense:
        assert float(ga["train/mean_num_samples_per_ray"]) < 60      # most samples were culled
    assert set(gra) == set(grb)
    # fp32 atomics: only the summation order differs.  2e-4, or — for the sums that cancel almost
    # completely — twice what one-ulp perturbations do to the reference's own gradient
    # (tests/golden/gradient_conditioning.npz)
    cond = _scene.load_golden("gradient_conditioning")
    name = {"synthetic": "pb_on" if pb_on else "pb_off", "eds": "eds"}[scene]
    for key in gra:
        bound = max(2e-4, 2 * float(cond.get(f"{name}/{key}", 0.0)))
        assert _rel(grb[key], gra[key]) < bound, (key, bound)


def test_sync_free_overflow_is_flagged_and_recovers(den_lib, cuda):
    """A capacity that is too small loses samples: the step is flagged (one call late), the estimate is
    dropped, and the next call synchronises again."""
    from deblur_e_nerf_b200 import factory
    model, batch, jitters = _setup(cuda, "synthetic", True)
    nerf = model.nerf
    opt = factory.configure_optimizer(model)
    assert opt.skip_flag is nerf.overflow_flag
    model.training_step(batch, 0, 0, jitters=jitters)
    nerf._spr_estimate *= 0.001                   # pretend the scene was almost empty so far
    nerf.capacity_margin = 0.0
    import deblur_e_nerf_b200.ops as ops
    quantum, ops._ROW_QUANTUM = ops._ROW_QUANTUM, 1024
    try:
        before = {n: p.detach().clone() for n, p in model.named_parameters()}
        model.zero_grad()
        model.training_step(batch, 0, 0, jitters=jitters).backward()  # sync-free with a tiny capacity
        assert nerf._stats is not None
        opt.step()                                                    # ... its update is skipped on the device
        assert int(nerf.overflow_flag) == 1
        for n, p in model.named_parameters():
            assert torch.equal(p.detach(), before[n]), n
        ref_mean = None
        loss = model.training_step(batch, 0, 0, jitters=jitters)      # consumes the stats: overflow seen
        assert nerf.overflow_count == 1 and nerf._stats is None       # -> this call synchronised
        ref_mean = float(model.logged["train/mean_num_samples_per_ray"])
        assert ref_mean > 5 and torch.isfinite(loss)
    finally:
        ops._ROW_QUANTUM = quantum


@pytest.mark.parametrize("pb_on", [False, True], ids=["pb_off", "pb_on"])
def test_graph_replay_equals_eager_steps(den_lib, cuda, pb_on):
    """Five optimizer steps: eager (synchronising) against GraphedStep (2 eager warm-up steps, a capture,
    replays).  Same batches, same jitter (the jitter is passed in, so both consume the same numbers):
    the parameters after the fifth step agree."""
    from deblur_e_nerf_b200 import ddp, factory
    from deblur_e_nerf_b200.graph_step import GraphedStep
    finals = []
    for graphed in (False, True):
        model, batch, jitters = _setup(cuda, "synthetic", pb_on)
        model.nerf.sync_free = graphed
        reducer = ddp.GradReducer(model)
        opt = factory.configure_optimizer(model)
        reducer.bind(opt)
        # fixed jitter for every step: training_step(jitters=None) would draw from the RNG
        step = model.training_step
        model.training_step = lambda b, m, gs: step(b, m, gs, jitters=[j.clone() for j in jitters])
        stepper = GraphedStep(model, opt, reducer, 1)
        if not graphed:
            stepper.warmup_steps = 1 << 60
        losses = []
        for i in range(5):
            losses.append(stepper([batch], 1 + i).detach().clone())
        torch.cuda.synchronize()
        if graphed:
            assert stepper.captures == 1 and stepper.replays == 3 and stepper.overflows == 0
        finals.append((torch.stack(losses), {n: p.detach().clone() for n, p in model.named_parameters()},
                       {id(p): int(s["step"]) for p, s in opt.state.items()}))
    (la, pa, sa), (lb, pb, sb) = finals
    assert _rel(lb, la) < 1e-4, (la, lb)
    assert float(la[-1]) < float(la[0])                           # it trains
    assert set(sa.values()) == set(sb.values()) == {5}            # host-side Adam step numbers follow
    for name in pa:
        assert _rel(pb[name], pa[name]) < 2e-3, name             # 5 Adam steps of lr 1e-2 on fp32-atomic grads


def test_graphed_step_falls_back_to_eager_when_the_step_does_not_fit(den_lib, cuda):
    """The memory guard of `training_step` (`EventRenderer._batch_fits`) can refuse the batched render path
    (e.g. with the enlarged sample estimate after a capacity overflow on a batch that fills the device);
    the sequential path reads counts back, which no CUDA graph can record.  GraphedStep must then run
    that step eagerly — with the same result as any other eager step — and capture a later one."""
    from deblur_e_nerf_b200 import ddp, factory
    from deblur_e_nerf_b200.graph_step import GraphedStep
    finals = []
    for refuse_once in (False, True):
        model, batch, jitters = _setup(cuda, "synthetic", True)
        reducer = ddp.GradReducer(model)
        opt = factory.configure_optimizer(model)
        reducer.bind(opt)
        step = model.training_step
        model.training_step = lambda b, m, gs: step(b, m, gs, jitters=[j.clone() for j in jitters])
        stepper = GraphedStep(model, opt, reducer, 1)
        fits = model._batch_fits
        calls = {"n": 0}

        def guarded(n_rays, gen, fits=fits, calls=calls, refuse_once=refuse_once):
            calls["n"] += 1
            # the third optimizer step (the first one GraphedStep would capture) is refused once
            if refuse_once and stepper._eager_done >= stepper.warmup_steps and stepper.eager_fallbacks == 0:
                return False
            return fits(n_rays, gen)
        model._batch_fits = guarded
        losses = [stepper([batch], 1 + i).detach().clone() for i in range(5)]
        torch.cuda.synchronize()
        if refuse_once:
            # step 3: the predicate said no -> eager (and, inside it, the sequential render calls)
            assert stepper.eager_fallbacks >= 1 and stepper.captures == 1 and stepper.replays >= 1
        else:
            assert stepper.eager_fallbacks == 0 and stepper.captures == 1 and stepper.replays == 3
        finals.append((torch.stack(losses), {n: p.detach().clone() for n, p in model.named_parameters()}))
    (la, pa), (lb, pb) = finals
    assert _rel(lb, la) < 1e-4, (la, lb)
    for name in pa:
        assert _rel(pb[name], pa[name]) < 2e-3, name
