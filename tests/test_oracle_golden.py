"""The oracle restatement reproduces the golden fixtures written from the REFERENCE'S OWN
files (tests/golden/make_golden.py).  Runs anywhere (CPU): this is what carries the
reference pin to the GPU box."""

import pytest
import torch

import _scene


def _close(a, b, tol):
    a = torch.as_tensor(a).detach().double()
    b = torch.as_tensor(b).detach().double()
    denom = b.abs().max().clamp(min=1e-30)
    err = ((a - b).abs().max() / denom).item()
    assert err <= tol, err


@pytest.mark.parametrize("pb_on,bayer", [(True, False), (False, False), (True, True)],
                         ids=["pb_on", "pb_off", "bayer"])
def test_oracle_training_step_matches_reference_golden(pb_on, bayer):
    """`bayer`: three radiance channels, every event sees the channel of its pixel
    (models/deblur_e_nerf.py:409-412,1177-1178,1223-1234)."""
    golden = _scene.load_golden("training_step_bayer" if bayer else
                                "training_step_pb_on" if pb_on else "training_step_pb_off")
    cfg = _scene.scene_config("synthetic", occ_resolution=32, small=True)
    if bayer:
        cfg["radiance_dim"] = 3
    ora, poses = _scene.build_oracle_renderer(cfg, 8, pixel_bandwidth=pb_on)
    for name in ["nerf", "contrast_threshold", "refractory_period"] + (
            ["pixel_bandwidth"] if pb_on else []):
        _scene.load_golden_state(getattr(ora, name), golden, name)
    event = _scene.golden_section(golden, "event")
    normalized = _scene.golden_section(golden, "normalized")
    jitters = [v for _, v in sorted(_scene.golden_section(golden, "jitter").items(),
                                    key=lambda kv: int(kv[0]))]
    ora.train()
    loss, terms, mean_samples = ora.training_step(event, normalized, jitters=jitters)
    _close(loss, golden["loss"], 1e-5)
    for key, val in terms.items():
        _close(val, golden[f"logged/train/{key}"], 1e-5)
    _close(mean_samples, golden["logged/train/mean_num_samples_per_ray"], 1e-6)
    ora.zero_grad()
    loss.backward()
    grads = _scene.flat_named_grads(ora)
    ref = _scene.golden_section(golden, "grad")
    assert set(grads) == set(ref), set(grads) ^ set(ref)
    for key in ref:
        _close(grads[key], ref[key], 5e-3 if "pixel_bandwidth" in key else 2e-4)


def test_oracle_field_matches_reference_golden():
    golden = _scene.load_golden("field_small")
    cfg = _scene.scene_config("synthetic", occ_resolution=32, small=True)
    nerf = _scene.build_oracle_nerf(cfg)
    field = nerf.radiance_field
    field.load_state_dict(_scene.golden_section(golden, "state"))
    rgb, sigma = field(torch.from_numpy(golden["x"]), torch.from_numpy(golden["d"]))
    _close(rgb, golden["rgb"], 1e-6)
    _close(sigma, golden["sigma"], 1e-6)
    ((rgb * torch.from_numpy(golden["w_rgb"])).sum()
     + (sigma * torch.from_numpy(golden["w_sigma"])).sum()).backward()
    for key, p in field.named_parameters():
        _close(p.grad, golden["grad/" + key], 1e-5)


RAW_EVENT_CASES, raw_event_case = _scene.RAW_EVENT_CASES, _scene.raw_event_case


@pytest.mark.parametrize("name", RAW_EVENT_CASES)
def test_oracle_raw_event_queueing_matches_reference_golden(name):
    """oracle/events_ref.py (the literal loops AND the vectorised forms) against what the reference's OWN
    Event.queue_raw_events / extract_max_refractory_period / colorize_events returned: bit for bit, dtypes
    included."""
    import numpy as np
    from oracle import events_ref
    raw, height, width, bayer, want, want_refractory = raw_event_case(name)
    for fn in (events_ref.queue_raw_events_loop, events_ref.queue_raw_events):
        got = fn(raw["position"], raw["timestamp"], raw["polarity"], height, width)
        for key in ("position", "start_ts", "end_ts", "num_pos", "num_neg"):
            assert got[key].dtype == want[key].dtype, (fn.__name__, key, got[key].dtype, want[key].dtype)
            assert np.array_equal(got[key], want[key]), (fn.__name__, key)
    for fn in (events_ref.max_refractory_period_loop, events_ref.max_refractory_period):
        assert float(fn(raw["position"], raw["timestamp"], height, width)) == float(want_refractory), fn.__name__
    channel = events_ref.colorize_events(want["position"], bayer)
    if bayer:
        assert channel.dtype == want["channel_idx"].dtype and np.array_equal(channel, want["channel_idx"])
    else:
        assert channel is None and "channel_idx" not in want
