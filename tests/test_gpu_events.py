"""GPU parity of the raw-event preprocessing (`events.py`, `den_radix_sort_pairs_u32`,
`den_queue_raw_events`) — bit-exact integer work:
* against the committed golden written by the reference's OWN Event.queue_raw_events /
  extract_max_refractory_period / colorize_events (tests/golden/raw_events.npz);
* against oracle/events_ref.py (pinned to the same golden) on seeded streams up to 3 M events, with the
  size-independent properties the domain offers (kept events stay in stream order; every kept event's
  start precedes... equals an earlier timestamp of its pixel; idempotent cache);
* the sort alone against torch's stable sort; empty / single-pixel / out-of-range edge cases."""

import numpy as np
import pytest
import torch

from oracle import events_ref
from _scene import RAW_EVENT_CASES, raw_event_case

pytestmark = pytest.mark.gpu


def _stream(seed, n, height, width, hot_fraction=0.3, repeat_ts=0.2, sorted_ts=True):
    rng = np.random.default_rng(seed)
    position = np.stack([rng.integers(0, width, n), rng.integers(0, height, n)], axis=1)
    hot = rng.random(n) < hot_fraction
    position[hot] = np.stack([rng.integers(0, min(width, 3), hot.sum()), rng.integers(0, min(height, 2), hot.sum())], axis=1)
    step = rng.integers(1, 2000, n)
    step[rng.random(n) < repeat_ts] = 0
    timestamp = np.cumsum(step).astype(np.int64) + 1_000_000
    if not sorted_ts:
        timestamp = rng.permutation(timestamp)
    return {"position": position.astype(np.uint16), "timestamp": timestamp, "polarity": rng.random(n) < 0.5}


def _calib(height, width, bayer=""):
    return {"img_height": np.array(height, dtype=np.uint16), "img_width": np.array(width, dtype=np.uint16),
            "bayer_pattern": np.array(bayer), "distortion_params": np.zeros(0), "distortion_model": np.array("plumb_bob"),
            "intrinsics": np.eye(3)}


@pytest.mark.parametrize("name", RAW_EVENT_CASES)
def test_queue_matches_reference_golden(den_lib, cuda, name):
    from deblur_e_nerf_b200 import events
    raw, height, width, bayer, want, want_refractory = raw_event_case(name)
    calib = _calib(height, width, bayer)
    got = events.colorize_events(events.queue_raw_events(raw, calib, cuda), calib)
    assert set(got) == set(want)
    for key, value in want.items():
        assert got[key].is_cuda and str(got[key].dtype).split(".")[1] == str(value.dtype), (key, got[key].dtype, value.dtype)
        assert np.array_equal(got[key].cpu().numpy(), value), key
    refractory = events.extract_max_refractory_period(raw, calib, cuda)
    assert refractory.dim() == 0 and float(refractory) == float(want_refractory)
    # one sort serves both, and the undistortion step of a distortion-free camera only casts the positions
    both, refractory2 = events.transform_raw_events(raw, calib, cuda)
    assert both["position"].dtype == torch.float32 and float(refractory2) == float(want_refractory)
    assert np.array_equal(both["position"].cpu().numpy(), want["position"].astype(np.float32))
    assert np.array_equal(both["start_ts"].cpu().numpy(), want["start_ts"])


@pytest.mark.parametrize("n,height,width,kw", [
    (1, 4, 4, {}), (2, 1, 1, {}), (257, 3, 300, {}), (2049, 260, 346, {}), (50_000, 260, 346, dict(hot_fraction=0.6)),
    (300_000, 480, 640, dict(sorted_ts=False)), (3_000_000, 480, 640, dict(hot_fraction=0.05, repeat_ts=0.02)),
])
def test_queue_matches_oracle(den_lib, cuda, n, height, width, kw):
    from deblur_e_nerf_b200 import events
    raw = _stream(n + height, n, height, width, **kw)
    calib = _calib(height, width)
    want = events_ref.queue_raw_events(raw["position"], raw["timestamp"], raw["polarity"], height, width)
    got, refractory = events.transform_raw_events(raw, calib, cuda)
    for key in ("start_ts", "end_ts", "num_pos", "num_neg"):
        assert got[key].dtype == torch.int64 and np.array_equal(got[key].cpu().numpy(), want[key]), key
    assert np.array_equal(got["position"].cpu().numpy(), want["position"].astype(np.float32))
    assert float(refractory) == float(events_ref.max_refractory_period(raw["position"], raw["timestamp"], height, width))
    # properties that hold at any size: kept events stay in stream order (end_ts of a time-ordered stream is
    # non-decreasing), an interval never has zero length, and polarity counts are one-hot
    if kw.get("sorted_ts", True) and len(want["end_ts"]):
        assert bool((got["end_ts"][1:] >= got["end_ts"][:-1]).all())
        assert bool((got["start_ts"] < got["end_ts"]).all())
    assert bool((got["start_ts"] != got["end_ts"]).all()) and bool((got["num_pos"] + got["num_neg"] == 1).all())


@pytest.mark.parametrize("n,bits", [(0, 8), (1, 1), (255, 8), (2048, 9), (2049, 17), (100_003, 19), (1_000_001, 32)])
def test_radix_sort_is_the_stable_sort(den_lib, cuda, n, bits):
    from deblur_e_nerf_b200 import events
    g = torch.Generator().manual_seed(n + bits)
    keys = torch.randint(0, 1 << bits, (n,), generator=g, dtype=torch.int64)
    if n > 10:
        keys[: n // 3] = keys[0]                     # long runs of one key: stability shows
    values = torch.arange(n, dtype=torch.int64)
    got_k, got_v = events.sort_pairs(keys.to(cuda), values.to(cuda), key_bits=bits)
    want_k, want_v = torch.sort(keys, stable=True)
    assert torch.equal(got_k.cpu(), want_k) and torch.equal(got_v.cpu(), want_v)


def test_event_edge_cases(den_lib, cuda, tmp_path):
    from deblur_e_nerf_b200 import events
    calib = _calib(5, 7, "GRBG")
    empty = {"position": np.zeros((0, 2), np.uint16), "timestamp": np.zeros(0, np.int64), "polarity": np.zeros(0, bool)}
    got, refractory = events.transform_raw_events(empty, calib, cuda)
    assert all(len(v) == 0 for v in got.values()) and float(refractory) == float("inf")
    assert refractory.dtype == torch.float64                           # upstream keeps np.array(float("inf"))
    # every event on one pixel with one timestamp: nothing survives, no interval exists
    same = {"position": np.full((100, 2), 3, np.uint16), "timestamp": np.full(100, 5, np.int64), "polarity": np.ones(100, bool)}
    got, refractory = events.transform_raw_events(same, calib, cuda)
    assert len(got["start_ts"]) == 0 and float(refractory) == float("inf")
    # a position outside the sensor: upstream's window lookup raises IndexError
    bad = _stream(0, 50, 5, 7)
    bad["position"][17] = (7, 0)
    with pytest.raises(IndexError):
        events.queue_raw_events(bad, calib, cuda)
    # load_events: raw_events.npz + camera_calibration.npz -> events.pt in the reference's layout, then the cache
    raw = _stream(3, 5000, 5, 7)
    np.savez(tmp_path / events.RAW_EVENTS_FILENAME, **raw)
    np.savez(tmp_path / events.CAMERA_CALIBRATION_FILENAME, **calib)
    first = events.load_events(str(tmp_path), cuda)
    cached = torch.load(tmp_path / events.TF_EVENTS_FILENAME, weights_only=True)
    assert set(cached) == {"position", "start_ts", "end_ts", "num_pos", "num_neg", "channel_idx"}
    assert all(not v.is_cuda for v in cached.values()) and cached["channel_idx"].dtype == torch.uint8
    want_channel = events_ref.colorize_events(first["position"].cpu().numpy(), "GRBG")
    assert np.array_equal(cached["channel_idx"].numpy(), want_channel)
    second = events.load_events(str(tmp_path), cuda)
    assert all(torch.equal(first[k], second[k]) for k in first)
    refractory = torch.load(tmp_path / events.MAX_REFRACTORY_PERIOD_FILENAME, weights_only=True)
    assert float(refractory) == float(events_ref.max_refractory_period(raw["position"], raw["timestamp"], 5, 7))
