"""Generate the golden fixtures from the REFERENCE'S OWN files (run unmodified under
oracle/ref_shim, CPU fp32).  Build container only (needs /root/reference):

    python tests/golden/make_golden.py

Writes tests/golden/{training_step_pb_on,training_step_pb_off,training_step_eds,training_step_bayer,
field_small,raw_events}.npz (`python tests/golden/make_golden.py bayer` / `raw_events` write only that one).  Each
file carries the parameters, the inputs (events, normalised samples, the stratified
jitter the reference drew, the occupancy grid after its step-0 update) and the
reference's outputs (loss, loss terms, mean samples per ray, every parameter gradient),
so the GPU box can check the CUDA path and the oracle against the reference without it.
"""

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import nerfacc_ref, ref_shim  # noqa: E402

import _scene  # noqa: E402

N_EVENTS = 96
IT_SAMPLE_SIZE = 8
SKIP_BUFFERS = ("grid_coords", "grid_indices")


def _np(t):
    return t.detach().cpu().numpy()


def training_step_golden(pb_on, path, bayer=False):
    """`bayer`: a colour sensor behind a Bayer filter (models/deblur_e_nerf.py:82-90,175-178,287-290,
    409-412,1177-1178,1223-1234): three radiance channels, every event carries the channel of its pixel."""
    cfg = _scene.scene_config("synthetic", occ_resolution=32, small=True)
    if bayer:
        cfg["radiance_dim"] = 3
    ref, poses = _scene.build_reference_renderer(cfg, IT_SAMPLE_SIZE, pixel_bandwidth=pb_on)
    event, normalized = _scene.make_batch(cfg, poses, N_EVENTS, IT_SAMPLE_SIZE, seed=7,
                                          pixel_bandwidth=pb_on)
    if bayer:
        event["channel_idx"] = torch.randint(0, 3, (N_EVENTS,), generator=torch.Generator().manual_seed(5))
    ref.train()
    nerfacc_ref.JITTER_LOG = []
    torch.manual_seed(11)
    loss = ref.training_step(_scene.reference_batch(event, normalized), 0)
    jitters = nerfacc_ref.JITTER_LOG
    nerfacc_ref.JITTER_LOG = None
    ref.zero_grad()
    loss.backward()

    out = {"loss": _np(loss)}
    for key, value in ref.logged.items():
        if key.startswith("train/log_intensity") or key == "train/mean_num_samples_per_ray":
            out["logged/" + key] = np.asarray(float(value))
    names = ["nerf", "contrast_threshold", "refractory_period"] + (
        ["pixel_bandwidth"] if pb_on else [])
    for name in names:
        for key, value in getattr(ref, name).state_dict().items():
            if key.rsplit(".", 1)[-1] in SKIP_BUFFERS:
                continue
            out[f"state/{name}/{key}"] = _np(value)
    for key, value in _scene.flat_named_grads(ref).items():
        out["grad/" + key] = _np(value)
    for key, value in event.items():
        out["event/" + key] = _np(value)
    for key, value in normalized.items():
        out["normalized/" + key] = _np(value)
    for i, jit in enumerate(jitters):
        out[f"jitter/{i}"] = _np(jit)
    np.savez_compressed(path, **out)
    print(path, f"loss={float(loss):.6f}", f"{os.path.getsize(path) / 1e6:.2f} MB",
          f"{len(jitters)} render calls")


def training_step_eds_golden(path, n_micro=2):
    """08_peanuts_running.yaml shape: sphere contraction, cone angle 0.004, no background
    (validity = opacity > 0), pixel bandwidth on, `refractory_period.freeze: false` and the
    contrast-threshold / pixel-bandwidth parameters trainable (:31-55), `accumulate_grad_batches`
    micro-batches whose gradients are accumulated with loss / n (models/deblur_e_nerf.py:465-469,
    Lightning's accumulation)."""
    cfg = _scene.scene_config("eds", occ_resolution=32, small=True)
    ref, poses = _scene.build_reference_renderer(cfg, IT_SAMPLE_SIZE, pixel_bandwidth=True,
                                                 freeze_refractory_period=False,
                                                 accumulate_grad_batches=n_micro)
    ref.train()
    ref.zero_grad()
    out = {"n_micro": np.asarray(n_micro)}
    torch.manual_seed(13)
    for m in range(n_micro):
        event, normalized = _scene.make_batch(cfg, poses, N_EVENTS, IT_SAMPLE_SIZE, seed=21 + m,
                                              pixel_bandwidth=True)
        nerfacc_ref.JITTER_LOG = []
        loss = ref.training_step(_scene.reference_batch(event, normalized), m)
        jitters = nerfacc_ref.JITTER_LOG
        nerfacc_ref.JITTER_LOG = None
        (loss / n_micro).backward()
        out[f"loss/{m}"] = _np(loss)
        for key, value in ref.logged.items():
            if key.startswith("train/log_intensity") or key in ("train/mean_num_samples_per_ray",
                                                                 "train/mean_valid_rate"):
                out[f"logged/{m}/{key}"] = np.asarray(float(value))
        for key, value in event.items():
            out[f"event/{m}/{key}"] = _np(value)
        for key, value in normalized.items():
            out[f"normalized/{m}/{key}"] = _np(value)
        for i, jit in enumerate(jitters):
            out[f"jitter/{m}/{i}"] = _np(jit)
        if m == 0:
            # parameters are untouched between the micro-batches; the grid was updated once (batch 0)
            for name in ("nerf", "contrast_threshold", "refractory_period", "pixel_bandwidth"):
                for key, value in getattr(ref, name).state_dict().items():
                    if key.rsplit(".", 1)[-1] in SKIP_BUFFERS:
                        continue
                    out[f"state/{name}/{key}"] = _np(value)
    for key, value in _scene.flat_named_grads(ref).items():
        out["grad/" + key] = _np(value)
    np.savez_compressed(path, **out)
    print(path, "losses", [float(out[f"loss/{m}"]) for m in range(n_micro)],
          f"{os.path.getsize(path) / 1e6:.2f} MB", "mean samples/ray",
          [float(out[f"logged/{m}/train/mean_num_samples_per_ray"]) for m in range(n_micro)])


def _step_gradients(cfg, pb_on, eds, perturb_seed):
    """Gradients of the reference's training step(s) with every field parameter multiplied by
    1 + u * 2^-23, u in {-1, 0, 1} (a one-ulp relative perturbation; seed 0 = unperturbed)."""
    n_micro = 2 if eds else 1
    ref, poses = _scene.build_reference_renderer(
        cfg, IT_SAMPLE_SIZE, pixel_bandwidth=pb_on, freeze_refractory_period=not eds,
        accumulate_grad_batches=n_micro)
    ref.train()
    ref.zero_grad()
    if perturb_seed:
        g = torch.Generator().manual_seed(perturb_seed)
        with torch.no_grad():
            for p in ref.nerf.radiance_field.parameters():
                p.mul_(1 + (torch.randint(0, 3, p.shape, generator=g).float() - 1) * 2.0 ** -23)
    torch.manual_seed(13 if eds else 11)
    for m in range(n_micro):
        event, normalized = _scene.make_batch(cfg, poses, N_EVENTS, IT_SAMPLE_SIZE,
                                              seed=(21 + m) if eds else 7, pixel_bandwidth=pb_on)
        loss = ref.training_step(_scene.reference_batch(event, normalized), m)
        (loss / n_micro).backward()
    return {k: v.double() for k, v in _scene.flat_named_grads(ref).items()}


def conditioning_golden(path, trials=3):
    """How far the REFERENCE'S OWN fp32 gradients move when its field parameters are perturbed by one
    ulp: `<scene>/<parameter>` = max over `trials` perturbations of max|g' - g| / max|g|.  A gradient
    that the reference's arithmetic cannot reproduce to better than c under a one-ulp change of its
    inputs cannot be pinned to better than a small multiple of c by any other fp32 evaluation (a
    GPU run of the reference itself included: atomics reorder its sums); the training-step parity
    tests use max(1e-3, 16 c) as the bound of such keys.  On the shipped shapes c stays below 3e-5
    except for the sums that cancel almost completely: the pixel-bandwidth parameters, the mean
    contrast threshold and the output-layer bias on the EDS shape (1e-3 .. 4e-3)."""
    out = {}
    scenes = {"pb_on": ("synthetic", True, False), "pb_off": ("synthetic", False, False),
              "eds": ("eds", True, True)}
    for name, (scene, pb_on, eds) in scenes.items():
        cfg = _scene.scene_config(scene, occ_resolution=32, small=True)
        base = _step_gradients(cfg, pb_on, eds, 0)
        spread = {k: 0.0 for k in base}
        for t in range(1, trials + 1):
            g = _step_gradients(cfg, pb_on, eds, t)
            for k in base:
                rel = float((g[k] - base[k]).abs().max() / base[k].abs().max().clamp(min=1e-300))
                spread[k] = max(spread[k], rel)
        for k, v in spread.items():
            out[f"{name}/{k}"] = np.asarray(v)
        print(name, {k.split(".")[-2] if k.endswith("original") else k.split(".")[-2] + "." + k.split(".")[-1]:
                     f"{v:.1e}" for k, v in spread.items()})
    np.savez_compressed(path, **out)


def field_golden(path):
    cfg = _scene.scene_config("synthetic", occ_resolution=32, small=True)
    ref, _ = _scene.build_reference_renderer(cfg, IT_SAMPLE_SIZE, pixel_bandwidth=False)
    field = ref.nerf.radiance_field
    g = torch.Generator().manual_seed(0)
    x = torch.rand(2000, 3, generator=g) * 3.4 - 1.7
    d = torch.randn(2000, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    rgb, sigma = field(x, d)
    w_rgb = torch.randn(rgb.shape, generator=g)
    w_sig = torch.randn(sigma.shape, generator=g) * 0.01
    field.zero_grad()
    ((rgb * w_rgb).sum() + (sigma * w_sig).sum()).backward()
    out = {"x": _np(x), "d": _np(d), "rgb": _np(rgb), "sigma": _np(sigma), "w_rgb": _np(w_rgb),
           "w_sigma": _np(w_sig)}
    for key, value in field.state_dict().items():
        out["state/" + key] = _np(value)
    for key, p in field.named_parameters():
        out["grad/" + key] = _np(p.grad)
    np.savez_compressed(path, **out)
    print(path, f"{os.path.getsize(path) / 1e6:.2f} MB")


def raw_event_stream(seed, n, height, width, hot_fraction=0.3, repeat_ts=0.2, sorted_ts=True):
    """A raw event stream in the reference's `raw_events.npz` layout (data/datasets.py:19-21: position (N, 2)
    uint16 (x, y), timestamp (N) int64, polarity (N) bool): part of the events crowd on a few hot pixels,
    part of the timestamps repeat (the window tests of :163-168 and :246-253)."""
    rng = np.random.default_rng(seed)
    position = np.stack([rng.integers(0, width, n), rng.integers(0, height, n)], axis=1)
    hot = rng.random(n) < hot_fraction
    position[hot] = np.stack([rng.integers(0, min(width, 3), hot.sum()), rng.integers(0, min(height, 2), hot.sum())], axis=1)
    step = rng.integers(1, 2000, n)
    step[rng.random(n) < repeat_ts] = 0
    timestamp = np.cumsum(step).astype(np.int64) + 1_000_000
    if not sorted_ts:
        timestamp = rng.permutation(timestamp)
    return position.astype(np.uint16), timestamp, rng.random(n) < 0.5


def raw_events_golden(path):
    """Event.queue_raw_events / extract_max_refractory_period / colorize_events of the reference's OWN
    data/datasets.py on small raw streams (time-ordered with repeats; shuffled in time; a Bayer sensor)."""
    import tempfile
    ds = ref_shim.load("data.datasets")
    out = {}
    cases = {"ordered": dict(seed=1, n=3000, height=9, width=13), "hot": dict(seed=2, n=2500, height=4, width=5, hot_fraction=0.8),
             "shuffled": dict(seed=3, n=2000, height=7, width=6, sorted_ts=False),
             "sparse": dict(seed=4, n=60, height=16, width=16, hot_fraction=0.0)}
    for name, kw in cases.items():
        position, timestamp, polarity = raw_event_stream(**kw)
        calib = {"img_height": np.array(kw["height"], dtype=np.uint16), "img_width": np.array(kw["width"], dtype=np.uint16),
                 "bayer_pattern": np.array("RGGB" if name == "hot" else "")}
        with tempfile.TemporaryDirectory() as root:
            np.savez(os.path.join(root, ds.Event.RAW_EVENTS_FILENAME), position=position, timestamp=timestamp,
                     polarity=polarity)
            queued = ds.Event.queue_raw_events(root, calib)
            refractory = ds.Event.extract_max_refractory_period(
                {"position": position, "timestamp": timestamp, "polarity": polarity}, calib)
            queued = ds.Event.colorize_events(queued, calib)
        out.update({f"{name}/raw/position": position, f"{name}/raw/timestamp": timestamp,
                    f"{name}/raw/polarity": polarity, f"{name}/height": calib["img_height"],
                    f"{name}/width": calib["img_width"], f"{name}/bayer_pattern": calib["bayer_pattern"],
                    f"{name}/max_refractory_period": _np(refractory)})
        out.update({f"{name}/queued/{k}": _np(v) for k, v in queued.items()})
        print(name, len(position), "raw ->", len(queued["position"]), "queued; max refractory period", float(refractory))
    np.savez_compressed(path, **out)
    print(path, f"{os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    assert ref_shim.available(), "needs /root/reference"
    if sys.argv[1:] == ["raw_events"]:
        raw_events_golden(os.path.join(HERE, "raw_events.npz"))
        sys.exit(0)
    training_step_golden(True, os.path.join(HERE, "training_step_bayer.npz"), bayer=True)
    if sys.argv[1:] == ["bayer"]:
        sys.exit(0)
    training_step_golden(True, os.path.join(HERE, "training_step_pb_on.npz"))
    training_step_golden(False, os.path.join(HERE, "training_step_pb_off.npz"))
    field_golden(os.path.join(HERE, "field_small.npz"))
    training_step_eds_golden(os.path.join(HERE, "training_step_eds.npz"))
    conditioning_golden(os.path.join(HERE, "gradient_conditioning.npz"))
    raw_events_golden(os.path.join(HERE, "raw_events.npz"))
