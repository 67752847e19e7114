"""`config.train` / `config.test` on the device: a config with the reference's YAML structure + a dataset
directory in the reference's layout (raw event stream, calibration, camera poses, posed views) -> the raw events
are queued on the device and cached, the optimizer-step loop runs the CUDA path, the evaluation loop scores the
test views."""

import os

import pytest
import torch

import _dataset

pytestmark = pytest.mark.gpu


def test_train_and_test_from_a_config(den_lib, cuda, tmp_path):
    from deblur_e_nerf_b200 import config, events, synthetic
    from oracle import events_ref
    import numpy as np
    scene = dict(synthetic.CONFIGS["synthetic"])
    root = str(tmp_path)
    _dataset.write(root, scene, channels=3)
    _dataset.write_raw_events(root, scene)
    cfg = _dataset.reference_style_config(root)
    logged = []
    model, loop = config.train(cfg, device=cuda, log_fn=lambda step, row: logged.append(row))
    assert loop.global_step == 4 and loop.current_epoch == 1
    assert len(logged) == 4 and all(row["train/loss"] == row["train/loss"] and row["train/loss"] > 0 for row in logged)
    # the raw stream was queued on the device and cached in the reference's layout: equal to the oracle's queue
    raw = np.load(os.path.join(root, "raw_events.npz"))
    want = events_ref.queue_raw_events(raw["position"], raw["timestamp"], raw["polarity"], scene["height"], scene["width"])
    cached = torch.load(os.path.join(root, events.TF_EVENTS_FILENAME), weights_only=True)
    assert np.array_equal(cached["start_ts"].numpy(), want["start_ts"]) and cached["position"].dtype == torch.float32
    assert any(p.grad is not None for p in model.nerf.parameters())
    frozen = [n for n, p in model.named_parameters() if not p.requires_grad]
    assert any(n.startswith("pixel_bandwidth.") for n in frozen) and any(n.startswith("refractory_period.") for n in frozen)
    # `run.py test` on the trained model
    rows = []
    metrics, pred = config.test(cfg, device=cuda, model=model, log_fn=lambda step, row: rows.append(row))
    assert set(metrics) == {"test/l1", "test/psnr", "test/ssim"} and rows == [metrics]
    assert all(v == v for v in metrics.values()) and pred.shape == (2, 1, 24, 32) and model.training
    val_metrics, _ = config.test(cfg, device=cuda, model=model, stage="val")
    assert set(val_metrics) == {"val/l1", "val/psnr", "val/ssim"}
