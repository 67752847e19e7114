"""CPU tests of the optimizer-step loop (trainer.py) on a stand-in model: the Lightning behaviours
the hot path relies on — step / epoch counting, gradient accumulation,
the two-batch lag of the batch controller, scheduler interval, checkpoint round trip."""

import pytest
import torch

from deblur_e_nerf_b200 import trainer as trainer_mod


class _Producer:
    def __init__(self, batch_size):
        self.batch_size = batch_size
        self.drawn = []
        self.g = torch.Generator().manual_seed(0)

    def set_batch_size(self, n):
        self.batch_size = int(n)

    def next_batch(self):
        self.drawn.append(self.batch_size)
        return {"x": torch.randn(self.batch_size, 3, generator=self.g, dtype=torch.float64)}


class _Model(torch.nn.Module):
    """Least squares on w; the 'controller' asks for batch size 10 + number of steps seen."""

    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(3, dtype=torch.float64))
        self.next_train_batch_size = None
        self.accumulate_grad_batches = 1
        self.logged = {}
        self.calls = []

    def training_step(self, batch, batch_index, global_step):
        self.calls.append((batch_index, global_step, batch["x"].shape[0]))
        self.next_train_batch_size = 10 + len(self.calls)
        loss = ((batch["x"] @ self.w - batch["x"].sum(-1)) ** 2).mean()
        self.logged = {"train/loss": loss.detach(), "train/batch_size": batch["x"].shape[0]}
        return loss


def test_step_and_epoch_counting_and_controller_lag():
    model, prod = _Model(), _Producer(8)
    opt = torch.optim.SGD(model.parameters(), lr=0.05)
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[1], gamma=0.5)
    tr = trainer_mod.Trainer(max_epochs=2, limit_train_batches=5, log_every_n_steps=2)
    out = tr.fit(model, prod, opt, sched)
    assert tr.global_step == 10 and tr.current_epoch == 2
    assert [c[0] for c in model.calls] == [0, 1, 2, 3, 4] * 2
    assert [c[1] for c in model.calls] == list(range(10))
    # one prefetched batch: the size chosen in batch k (10 + k + 1) is first seen by batch k + 2
    seen = [c[2] for c in model.calls]
    assert seen[:2] == [8, 8]
    assert seen[2:] == [10 + k + 1 for k in range(8)]
    assert opt.param_groups[0]["lr"] == 0.05 * 0.5          # epoch interval: one milestone passed
    assert [s for s, _ in tr.history] == [2, 4, 6, 8, 10]
    assert out["train/loss"] < 3.0 and isinstance(out["train/loss"], float)


def test_gradient_accumulation_equals_the_mean_gradient():
    torch.manual_seed(0)
    xs = [torch.randn(6, 3, dtype=torch.float64) for _ in range(4)]

    class Fixed:
        def __init__(self):
            self.i = 0

        def set_batch_size(self, n):
            pass

        def next_batch(self):
            self.i += 1
            return {"x": xs[(self.i - 1) % 4]}

    model = _Model()
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    tr = trainer_mod.Trainer(max_epochs=1, limit_train_batches=4, accumulate_grad_batches=4,
                             lr_scheduler_interval="step")
    tr.fit(model, Fixed(), opt)
    assert tr.global_step == 1 and model.accumulate_grad_batches == 4
    ref = _Model()
    loss = sum(((x @ ref.w - x.sum(-1)) ** 2).mean() for x in xs) / 4
    loss.backward()
    assert torch.allclose(model.w.detach(), -0.1 * ref.w.grad, rtol=1e-12, atol=0)

    # an epoch that is not a multiple of the window still steps on its last batch
    model2 = _Model()
    opt2 = torch.optim.SGD(model2.parameters(), lr=0.1)
    tr2 = trainer_mod.Trainer(max_epochs=1, limit_train_batches=3, accumulate_grad_batches=2)
    tr2.fit(model2, Fixed(), opt2)
    assert tr2.global_step == 2


def test_checkpoint_round_trip_and_resume(tmp_path):
    def run(epochs):
        model, prod = _Model(), _Producer(8)
        opt = torch.optim.Adam(model.parameters(), lr=0.05)
        sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[1, 2], gamma=0.5)
        tr = trainer_mod.Trainer(max_epochs=epochs, limit_train_batches=3,
                                 checkpoint_dir=str(tmp_path))
        tr.fit(model, prod, opt, sched)
        return model, opt, tr, prod

    model_a, _, tr_a, prod_a = run(1)
    ckpt_path = str(tmp_path / "last.ckpt")
    ckpt = torch.load(ckpt_path, weights_only=False)
    assert set(ckpt) >= {"epoch", "global_step", "state_dict", "optimizer_states", "lr_schedulers"}
    assert ckpt["epoch"] == 1 and ckpt["global_step"] == 3
    assert ckpt["train_batch_size"]["next"] == model_a.next_train_batch_size
    assert tr_a.max_steps is None
    # `max_steps` stops inside an epoch
    model_s, prod_s = _Model(), _Producer(8)
    tr_s = trainer_mod.Trainer(max_epochs=5, limit_train_batches=3, max_steps=4)
    tr_s.fit(model_s, prod_s, torch.optim.SGD(model_s.parameters(), lr=0.01))
    assert tr_s.global_step == 4 and tr_s.current_epoch == 1
    # resuming restores the counters, the optimizer moments, the schedule and the controller
    model_b, prod_b = _Model(), _Producer(8)
    opt_b = torch.optim.Adam(model_b.parameters(), lr=0.05)
    sched_b = torch.optim.lr_scheduler.MultiStepLR(opt_b, milestones=[1, 2], gamma=0.5)
    tr_b = trainer_mod.Trainer(max_epochs=2, limit_train_batches=3)
    tr_b.load_checkpoint(ckpt_path, model_b, opt_b, sched_b)
    assert tr_b.current_epoch == 1 and tr_b.global_step == 3
    assert torch.equal(model_b.w.detach(), model_a.w.detach())
    assert opt_b.param_groups[0]["lr"] == 0.05 * 0.5
    assert opt_b.state_dict()["state"][0]["step"] == 3
    tr_b.fit(model_b, prod_b, opt_b, sched_b)
    assert tr_b.current_epoch == 2 and tr_b.global_step == 6
    assert prod_b.drawn[0] == model_a.next_train_batch_size


def test_seed_everything_reseeds_torch():
    trainer_mod.seed_everything(7)
    a = torch.rand(3)
    trainer_mod.seed_everything(7)
    assert torch.equal(a, torch.rand(3))


def test_config_round_trips_a_checkpoint(tmp_path):
    """`config.build_model` with `model.checkpoint_filepath` + per-component `load_state_dict` (models/
    deblur_e_nerf.py:321-343): a checkpoint written by `trainer.Trainer.save_checkpoint` (Lightning's keys) is
    read with the safe unpickler, only the selected components are restored, a frozen loaded field stays frozen;
    `build_producer`'s dataset ratio follows data/datamodule.py:127-142."""
    import copy
    import sys
    import os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import _dataset
    from deblur_e_nerf_b200 import config, synthetic, trainer
    root = str(tmp_path / "data")
    _dataset.write(root, dict(synthetic.CONFIGS["synthetic"]), channels=3)
    cfg = _dataset.reference_style_config(root)
    torch.manual_seed(0)
    first = config.build_model(cfg, device="cpu", world_size=2)
    assert first.train_ray_sample_batch_size == cfg["data"]["train_eff_ray_sample_batch_size"] // 2
    optimizer, scheduler = config.build_optimizer(cfg, first, fused=False)
    with torch.no_grad():
        for p in first.parameters():
            p.add_(0.01 * torch.randn_like(p))
    path = str(tmp_path / "last.ckpt")
    trainer.Trainer().save_checkpoint(path, first, optimizer, scheduler)

    cfg2 = copy.deepcopy(cfg)
    cfg2["model"]["checkpoint_filepath"] = path
    cfg2["model"]["nerf"].update(load_state_dict=True, freeze=True)
    cfg2["model"]["contrast_threshold"]["load_state_dict"] = True
    torch.manual_seed(1)
    second = config.build_model(cfg2, device="cpu", world_size=1)
    a, b = dict(first.named_parameters()), dict(second.named_parameters())
    for name in a:
        same = torch.equal(a[name], b[name])
        if name.startswith(("nerf.", "contrast_threshold.")):
            assert same, name                                           # restored
        elif name.startswith("pixel_bandwidth."):
            assert not same, name                                       # not selected: calibration values
    assert not any(p.requires_grad for p in second.nerf.parameters())
    cfg3 = copy.deepcopy(cfg)
    cfg3["model"]["nerf"]["freeze"] = True                              # frozen but random: refused (:66-69)
    with pytest.raises(AssertionError):
        config.build_model(cfg3, device="cpu")
    assert config._subset_length(0.5, 64, 1001) == 500 and config._subset_length(3, 64, 1001) == 192
    with pytest.raises(AssertionError):
        config._subset_length(100, 64, 1001)


def test_validation_runs_between_epochs():
    """`fit(validate_fn=...)`: Lightning's validation loop between training epochs — after every
    `check_val_every_n_epoch`-th epoch, with the model handed back in training mode."""
    model, producer = _Model(), _Producer(8)
    opt = torch.optim.SGD(model.parameters(), lr=0.05)
    calls = []

    def validate(m):
        m.eval()
        calls.append((tr.current_epoch, tr.global_step, m.training))

    tr = trainer_mod.Trainer(max_epochs=4, limit_train_batches=3)
    tr.fit(model, producer, opt, validate_fn=validate, check_val_every_n_epoch=2)
    assert calls == [(1, 6, False), (3, 12, False)] and model.training


def test_tensorboard_log_fn(tmp_path):
    from tensorboard.backend.event_processing.event_accumulator import EventAccumulator
    seen = []
    log = trainer_mod.tensorboard_log_fn(str(tmp_path), also=lambda step, row: seen.append(step))
    model, prod = _Model(), _Producer(8)
    tr = trainer_mod.Trainer(max_epochs=1, limit_train_batches=4, log_every_n_steps=2, log_fn=log)
    tr.fit(model, prod, torch.optim.SGD(model.parameters(), lr=0.05))
    log.writer.close()
    acc = EventAccumulator(str(tmp_path))
    acc.Reload()
    assert seen == [2, 4] and [e.step for e in acc.Scalars("train/loss")] == [2, 4]
    assert "train/batch_size" in acc.Tags()["scalars"]
