"""The optimizer-step loop on the device: trainer.Trainer + data.EventBatchProducer + optim.FusedAdam
drive EventRenderer.training_step like the reference's Lightning Trainer drives
DeblurENeRF.training_step (scripts/run.py:70-100).  A scene with a known answer — a field that
must learn a brightness ramp cannot be built without a dataset, so the checks are the ones the loop
itself owns: the loss goes down on a fixed event pool, the batch controller's size reaches the
producer two batches later, gradient accumulation takes one step on the summed gradient, a checkpoint restores
parameters / moments / counters."""

import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(cuda, seed=0, acc=1, pool=4096, batch=256, budget=1 << 15):
    from deblur_e_nerf_b200 import data, factory, synthetic
    model, cfg, poses = factory.build_renderer("synthetic", cuda, pixel_bandwidth=False, small=True,
                                               occ_resolution=32, n_poses=200, sample_budget=budget,
                                               accumulate_grad_batches=acc, seed=seed)
    factory.freeze_like_synthetic_yaml(model)
    events = synthetic.event_batch(pool, cfg, poses[2], torch.Generator().manual_seed(seed + 1))
    producer = data.EventBatchProducer(events, batch, None, cuda, seed=seed)
    opt = factory.configure_optimizer(model)
    return model, producer, opt


def test_fit_lowers_the_loss_and_feeds_the_controller_back(den_lib, cuda):
    from deblur_e_nerf_b200 import trainer
    model, producer, opt = _setup(cuda)
    sizes, chosen = [], []
    step_fn = model.training_step

    def spy(batch, batch_index, global_step):
        sizes.append(batch["event"]["start_ts"].numel())
        loss = step_fn(batch, batch_index, global_step)
        chosen.append(model.next_train_batch_size)
        return loss

    model.training_step = spy
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[2], gamma=0.33)
    tr = trainer.Trainer(max_epochs=3, limit_train_batches=20, log_every_n_steps=5)
    tr.fit(model, producer, opt, sched)
    assert tr.global_step == 60 and tr.current_epoch == 3
    losses = [row["train/loss"] for _, row in tr.history]
    assert all(l == l and l < 1e6 for l in losses)
    assert sum(losses[-3:]) / 3 < sum(losses[:3]) / 3
    assert abs(opt.param_groups[-1]["lr"] - 0.01 * 0.33) < 1e-12
    # the controller's N = int(budget / mean samples per ray) reaches the producer with a lag of two
    assert sizes[0] == sizes[1] == 256
    assert all(c and c > 0 for c in chosen) and any(c != 256 for c in chosen)
    assert sizes[2:] == chosen[:-2]
    assert producer.batch_size == chosen[-1]
    assert type(opt).__name__ == "FusedAdam"


def test_accumulated_micro_batches_step_once_on_the_mean_gradient(den_lib, cuda):
    """accumulate_grad_batches = k: k micro-batches, each loss divided by k, ONE optimizer step on the
    summed gradient, the occupancy update only at the first micro-batch of a window
    (models/deblur_e_nerf.py:465).  Checked against the same k micro-batches pushed through
    training_step by hand on an identically seeded model; the per-ray jitter is pinned through the
    `jitters` hook so that both runs march the same samples."""
    from deblur_e_nerf_b200 import synthetic, trainer
    k, n = 4, 128

    def build():
        model, producer, opt = _setup(cuda, seed=3, acc=k, pool=k * n, batch=n)
        sphere = synthetic.solid_sphere_occupancy(32).to(cuda)
        model.nerf.occupancy_grid._binary = sphere
        model.nerf.occupancy_grid.occs.copy_(sphere.reshape(-1).float())
        occ_calls = []
        model.nerf.update_occ_grid = lambda *a, **kw: occ_calls.append(kw.get("step"))
        half = torch.full((n,), 0.5, dtype=torch.float64, device=cuda)
        batches = [{"event": {key: v[i * n:(i + 1) * n] for key, v in producer.events.items()},
                    "normalized": {"ts_diff": torch.ones_like(half), "diff_start_ts": half,
                                   "ts_subdiff": half, "subdiff_start_ts": half}}
                   for i in range(k)]
        jit = [torch.full((n,), 0.5, device=cuda) for _ in range(4)]
        return model, opt, batches, jit, occ_calls

    # by hand
    model, opt, batches, jit, _ = build()
    model.train()
    before = {name: p.detach().clone() for name, p in model.named_parameters()}
    for i, batch in enumerate(batches):
        (model.training_step(batch, i, 0, jitters=jit) / k).backward()
    expect = {name: p.grad.detach().clone() for name, p in model.named_parameters()
              if p.grad is not None}
    opt.step()
    expect_after = {name: p.detach().clone() for name, p in model.named_parameters()}

    # through the trainer
    model_t, opt_t, batches_t, jit_t, occ_calls = build()

    class Fixed:
        i = 0

        def set_batch_size(self, size):
            pass

        def next_batch(self):
            self.i += 1
            return batches_t[(self.i - 1) % k]

    step_fn = model_t.training_step
    model_t.training_step = lambda b, bi, gs: step_fn(b, bi, gs, jitters=jit_t)
    got = {}
    real_step = opt_t.step

    def capture():
        got.update({name: p.grad.detach().clone() for name, p in model_t.named_parameters()
                    if p.grad is not None})
        return real_step()

    opt_t.step = capture
    tr = trainer.Trainer(max_epochs=1, limit_train_batches=k, accumulate_grad_batches=k)
    tr.fit(model_t, Fixed(), opt_t)
    assert tr.global_step == 1
    assert occ_calls == [0]                 # gated on batch_index % accumulate_grad_batches == 0
    assert got.keys() == expect.keys() and len(got) >= 5
    for name, g in expect.items():
        scale = g.abs().max().clamp(min=1e-20)
        # same launches on the same inputs; only the order of the hash-gradient atomics differs
        assert ((got[name] - g).abs().max() / scale).item() < 1e-3, name
    moved = 0
    for name, p in model_t.named_parameters():
        if name in expect:
            moved += int(not torch.equal(p.detach(), before[name]))
            scale = (expect_after[name] - before[name]).abs().max().clamp(min=1e-20)
            assert ((p.detach() - expect_after[name]).abs().max() / scale).item() < 5e-2, name
    assert moved >= 5


def test_checkpoint_restores_the_training_state(den_lib, cuda, tmp_path):
    from deblur_e_nerf_b200 import trainer
    model, producer, opt = _setup(cuda, seed=5)
    tr = trainer.Trainer(max_epochs=1, limit_train_batches=6, checkpoint_dir=str(tmp_path))
    tr.fit(model, producer, opt)
    path = str(tmp_path / "last.ckpt")
    model_b, producer_b, opt_b = _setup(cuda, seed=6)
    tr_b = trainer.Trainer(max_epochs=2, limit_train_batches=6)
    tr_b.load_checkpoint(path, model_b, opt_b)
    assert tr_b.global_step == 6 and tr_b.current_epoch == 1
    for (n, p), (_, q) in zip(model.named_parameters(), model_b.named_parameters()):
        assert torch.equal(p, q), n
    for (n, p), (_, q) in zip(model.named_buffers(), model_b.named_buffers()):
        assert torch.equal(p, q), n
    sa, sb = opt.state_dict()["state"], opt_b.state_dict()["state"]
    assert sa.keys() == sb.keys()
    for k in sa:
        assert torch.equal(sa[k]["exp_avg"], sb[k]["exp_avg"])
        assert torch.equal(sa[k]["exp_avg_sq"], sb[k]["exp_avg_sq"])
    assert model_b.next_train_batch_size == model.next_train_batch_size
    tr_b.fit(model_b, producer_b, opt_b)
    assert tr_b.global_step == 12


def test_event_batch_producer_graph_replay(den_lib, cuda):
    """On the device a draw is replayed from a CUDA graph (captured the second time a batch size is
    asked for): batches keep the layout and value ranges of the eager draw, differ from call to call
    (the generator advances on every replay), and the previously returned batch stays intact while
    the next one is drawn (two alternating buffers — the trainer prefetches one batch)."""
    from deblur_e_nerf_b200 import synthetic
    from deblur_e_nerf_b200.data import EventBatchProducer
    cfg = synthetic.CONFIGS["synthetic"]
    poses = synthetic.camera_poses(cfg, n_poses=50)
    pool = synthetic.event_batch(5000, cfg, poses[2], torch.Generator().manual_seed(0))
    prod = EventBatchProducer(pool, 257, it_sample_size=8, device=cuda, seed=3)
    eager = prod.next_batch()                      # first call: eager
    seen = []
    prev = prev_copy = None
    for i in range(6):
        b = prod.next_batch()
        if prev is not None:
            for k in prev["event"]:
                assert torch.equal(prev["event"][k], prev_copy[k]), k      # untouched by the new draw
        for k, v in eager["event"].items():
            assert b["event"][k].shape == v.shape and b["event"][k].dtype == v.dtype
        for k, v in eager["normalized"].items():
            assert b["normalized"][k].shape == v.shape and b["normalized"][k].dtype == v.dtype
        u = b["normalized"]["diff_start_ts"]
        assert float(u.min()) >= 0 and float(u.max()) < 1
        assert bool((b["event"]["end_ts"] > b["event"]["start_ts"]).all())
        # every row is a row of the pool
        idx = (b["event"]["end_ts"][:, None] == pool["end_ts"].to(cuda)[None, :]).float().argmax(dim=1)
        assert torch.equal(pool["position"].to(cuda)[idx], b["event"]["position"])
        seen.append(b["event"]["end_ts"].clone())
        prev, prev_copy = b, {k: v.clone() for k, v in b["event"].items()}
    assert len(prod._graphs[257][1]) == 2
    for a in range(len(seen)):
        for c in range(a + 1, len(seen)):
            assert not torch.equal(seen[a], seen[c])
