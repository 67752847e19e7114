"""The N>1 path on CPU: world_size-2 `gloo` processes exercising the data-parallel plumbing
(flat gradient all-reduce = mean over ranks, parameter broadcast, the batch controller's
cross-rank mean, max/sum-over-ranks timing helpers) exactly as bench.py / the trainer use it."""

import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from deblur_e_nerf_b200 import ddp
    r, lr, w = ddp.init_from_env(backend="gloo")
    assert (r, w) == (rank, world) and ddp.world_size() == world

    torch.manual_seed(100 + rank)                       # ranks start different ...
    model = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3)).double()
    model.add_module("f32", torch.nn.Linear(3, 2))       # mixed dtypes -> one flat buffer each
    ddp.broadcast_parameters(model)                      # ... and end up identical to rank 0
    torch.manual_seed(7 + rank)                          # each rank its own shard of the batch
    x = torch.randn(11, 5, dtype=torch.float64)
    y = model[1](model[0](x))
    loss = y.pow(2).mean() + model.f32(y.float()).pow(2).mean()
    loss.backward()
    local = [p.grad.clone() for p in model.parameters()]
    reducer = ddp.FlatGradAllReduce(model.parameters())
    nbytes = reducer()
    assert nbytes == sum(p.numel() * p.element_size() for p in model.parameters())

    # the copy-free reducer: .grad tensors are views of one flat buffer, SUM all-reduce, 1/N left to
    # the optimizer (grad_scale)
    model2 = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    model2[1].bias.requires_grad_(False)                 # frozen: holds no view, never reduced
    ddp.broadcast_parameters(model2)
    red = ddp.GradReducer(model2)

    class _Opt:
        grad_scale = 1.0
    opt = red.bind(_Opt())
    ptrs = [p.grad.data_ptr() for p in model2.parameters() if p.requires_grad]
    model2(x.float()).pow(2).mean().backward()
    assert ptrs == [p.grad.data_ptr() for p in model2.parameters() if p.requires_grad]   # accumulated in place
    local2 = [p.grad.clone() for p in model2.parameters() if p.requires_grad]
    nbytes2 = red()
    summed2 = [p.grad.clone() for p in model2.parameters() if p.requires_grad]
    # dropping the views (zero_grad(set_to_none=True)) is repaired by the next call
    for p in model2.parameters():
        p.grad = None
    model2(x.float()).pow(2).mean().backward()
    red()
    again2 = [p.grad.clone() for p in model2.parameters() if p.requires_grad]
    assert ptrs == [p.grad.data_ptr() for p in model2.parameters() if p.requires_grad]
    assert model2[1].bias.grad is None

    class _R(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.p = torch.nn.Parameter(torch.zeros(1))
            self.mean_samples_reduce_fn = None
    ren = ddp.attach(_R())
    mean = ren.mean_samples_reduce_fn(10.0 * (rank + 1))          # (10 + 20) / 2
    tmax = ddp.max_over_ranks(3.0 + rank, torch.device("cpu"))
    tsum = ddp.sum_over_ranks(3.0 + rank, torch.device("cpu"))
    ddp.barrier()
    torch.save({"params": [p.detach().clone() for p in model.parameters()],
                "local": local, "reduced": [p.grad.clone() for p in model.parameters()],
                "mean": mean, "max": tmax, "sum": tsum, "local2": local2, "summed2": summed2,
                "again2": again2, "scale2": opt.grad_scale, "nbytes2": nbytes2},
               os.path.join(out_dir, f"rank{rank}.pt"))
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(180)
def test_world_size_2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(tmp_path / f"rank{r}.pt") for r in range(world)]
    for a, b in zip(res[0]["params"], res[1]["params"]):
        assert torch.equal(a, b)                                   # broadcast replicated rank 0
    for k in range(len(res[0]["local"])):
        mean = (res[0]["local"][k] + res[1]["local"][k]) / 2
        assert torch.allclose(res[0]["reduced"][k], mean, rtol=1e-12, atol=0)
        assert torch.equal(res[0]["reduced"][k], res[1]["reduced"][k])
    for k in range(len(res[0]["local2"])):                       # GradReducer: SUM, mean via grad_scale
        total = res[0]["local2"][k] + res[1]["local2"][k]
        assert torch.allclose(res[0]["summed2"][k], total, rtol=1e-6, atol=1e-8)
        assert torch.equal(res[0]["summed2"][k], res[1]["summed2"][k])
        assert torch.allclose(res[0]["again2"][k], total, rtol=1e-6, atol=1e-8)
    assert res[0]["scale2"] == 0.5 and res[0]["nbytes2"] >= (5 * 7 + 7 + 7 * 3) * 4
    assert res[0]["mean"] == res[1]["mean"] == 15.0
    assert res[0]["max"] == res[1]["max"] == 4.0
    assert res[0]["sum"] == res[1]["sum"] == 7.0


def _trainer_worker(rank, world, port, out_dir):
    """trainer.Trainer under world_size 2: every rank trains on its own event shard
    (EventBatchProducer seeded with seed + rank), gradients are mean all-reduced once per optimizer
    step (also with accumulation), rank 0 alone writes the checkpoint."""
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from deblur_e_nerf_b200 import data, ddp, trainer
    ddp.init_from_env(backend="gloo")

    class Model(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.randn(2, dtype=torch.float64))
            self.next_train_batch_size = None
            self.accumulate_grad_batches = 1
            self.logged = {}
            self.seen = []

        def training_step(self, batch, batch_index, global_step):
            pos = batch["event"]["position"].double()
            self.seen.append(pos[:, 0].sum().item())
            loss = ((pos @ self.w - batch["normalized"]["diff_start_ts"]) ** 2).mean()
            self.logged = {"train/loss": loss.detach()}
            return loss

    torch.manual_seed(50 + rank)
    model = Model()
    ddp.broadcast_parameters(model)
    g = torch.Generator().manual_seed(0)                 # the same event pool on every rank
    n = 64
    events = {"position": torch.rand(n, 2, generator=g), "start_ts": torch.arange(n),
              "end_ts": torch.arange(n) + 5, "num_pos": torch.ones(n, dtype=torch.int64),
              "num_neg": torch.zeros(n, dtype=torch.int64)}
    producer = data.EventBatchProducer(events, 8, None, "cpu", seed=3, rank=rank)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    tr = trainer.Trainer(max_epochs=2, limit_train_batches=4, accumulate_grad_batches=2,
                         checkpoint_dir=os.path.join(out_dir, "ckpt"))
    tr.fit(model, producer, opt)
    torch.save({"w": model.w.detach().clone(), "seen": model.seen, "steps": tr.global_step},
               os.path.join(out_dir, f"trainer_rank{rank}.pt"))
    ddp.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(180)
def test_trainer_world_size_2_gloo(tmp_path):
    world = 2
    mp.spawn(_trainer_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(tmp_path / f"trainer_rank{r}.pt") for r in range(world)]
    assert res[0]["steps"] == res[1]["steps"] == 4              # 2 epochs x 4 batches / 2 accumulated
    assert res[0]["seen"] != res[1]["seen"]                     # rank-offset seeds: different shards
    assert torch.equal(res[0]["w"], res[1]["w"])                # mean all-reduce keeps replicas equal
    ckpts = os.listdir(tmp_path / "ckpt")
    assert ckpts == ["last.ckpt"]
    ckpt = torch.load(tmp_path / "ckpt" / "last.ckpt", weights_only=False)
    assert ckpt["global_step"] == 4 and torch.equal(ckpt["state_dict"]["w"], res[0]["w"])


def _eval_worker(rank, world, port, out_dir):
    """trainer.Trainer.test under world_size 2: five views over two ranks (3 + 2: the short rank is padded
    for the all_gather), gathered back in view order, same metrics on both ranks."""
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from deblur_e_nerf_b200 import ddp, renderer, trainer
    ddp.init_from_env(backend="gloo")

    class Model(torch.nn.Module):
        image_pixel_positions = staticmethod(renderer.EventRenderer.image_pixel_positions)

        def __init__(self):
            super().__init__()
            self.rendered, self.modes = [], []

        def evaluation_step(self, view, intrinsics_inv, pos):
            self.rendered.append(int(view["sample_id"]))
            self.modes.append(self.training)
            img = view["img"]
            return {"sample_id": view["sample_id"], "pred_intensity_img": img * 0.5 + pos[..., 0] * 0.01,
                    "target_intensity_img": img, "exposure_time": view["exposure_time"],
                    "gain": torch.ones(())}

        def evaluation_epoch_end(self, outputs, lo, hi, stage="test", black_level_offset=False):
            pred = torch.stack([o["pred_intensity_img"] for o in outputs])
            target = torch.stack([o["target_intensity_img"] for o in outputs])
            order = torch.stack([o["exposure_time"] for o in outputs])
            return {f"{stage}/l1": (pred - target).abs().mean(), f"{stage}/first": target[0].mean(),
                    f"{stage}/n": torch.tensor(float(len(outputs))), f"{stage}/flag": torch.tensor(float(black_level_offset)),
                    f"{stage}/order": (order.double() * torch.arange(1, len(order) + 1)).sum()}, pred

    g = torch.Generator().manual_seed(0)
    views = [{"img": torch.rand(6, 9, generator=g) + i, "T_wc_position": torch.zeros(3),
              "T_wc_orientation": torch.eye(3), "sample_id": torch.tensor(i),
              "exposure_time": torch.tensor(i + 1)} for i in range(5)]
    model = Model().train()
    row, pred = trainer.Trainer().test(model, views, torch.eye(3), 0.0, 6.0, black_level_offset=True)
    val_row, _ = trainer.Trainer().validate(model, views[:2], torch.eye(3), 0.0, 6.0)
    try:                                    # fewer views than ranks: refused on BOTH ranks, nobody hangs
        trainer.Trainer().test(model, views[:1], torch.eye(3), 0.0, 6.0)
        refused = False
    except ValueError:
        refused = True
    torch.save({"row": row, "val_row": val_row, "pred": pred, "rendered": model.rendered, "modes": model.modes,
                "refused": refused,
                "training_after": model.training,
                "want_l1": float(torch.stack([(v["img"] * 0.5 + model.image_pixel_positions(6, 9)[..., 0] * 0.01
                                              - v["img"]).abs() for v in views]).mean())},
               os.path.join(out_dir, f"eval_rank{rank}.pt"))
    ddp.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(180)
def test_trainer_test_loop_world_size_2_gloo(tmp_path):
    world = 2
    mp.spawn(_eval_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(tmp_path / f"eval_rank{r}.pt") for r in range(world)]
    assert res[0]["rendered"] == [0, 2, 4, 0, 0] and res[1]["rendered"] == [1, 3, 1]   # round-robin shares (+ validate, + the refused one)
    assert res[0]["refused"] and res[1]["refused"]
    assert not any(res[0]["modes"]) and res[0]["training_after"]                # eval mode inside, restored after
    assert res[0]["row"] == res[1]["row"] and torch.equal(res[0]["pred"], res[1]["pred"])
    row = res[0]["row"]
    assert row["test/n"] == 5.0 and row["test/flag"] == 1.0
    assert row["test/order"] == float(sum((i + 1) * (i + 1) for i in range(5)))   # views back in their order
    assert abs(row["test/l1"] - res[0]["want_l1"]) < 1e-6
    assert res[0]["val_row"]["val/n"] == 2.0 and "val/l1" in res[0]["val_row"]


def test_single_process_is_a_noop():
    sys.path.insert(0, ROOT)
    from deblur_e_nerf_b200 import ddp
    assert ddp.world_size() == 1
    lin = torch.nn.Linear(2, 2)
    lin(torch.ones(1, 2)).sum().backward()
    g = lin.weight.grad.clone()
    assert ddp.FlatGradAllReduce(lin.parameters())() == 0
    assert torch.equal(lin.weight.grad, g)
    assert ddp.max_over_ranks(2.5, torch.device("cpu")) == 2.5
