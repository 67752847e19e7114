"""The data format in front of the hot path: raw event stream -> the transformed events the batch producer
draws from (``Event.__init__`` of data/datasets.py:36-66), on the device.

The reference turns ``raw_events.npz`` (position (N, 2) uint16 (x, y), timestamp (N) int64, polarity (N) bool)
into ``events.pt`` with two Python loops over ALL raw events that keep a two-entry sliding window per pixel:
``queue_raw_events`` (:186-276: every event gets the time of the previous event at its pixel as the start of
its interval; the first event of a pixel, and an event that repeats the previous timestamp, are dropped) and
``extract_max_refractory_period`` (:131-183: the smallest interval between distinct consecutive timestamps of
a pixel).  Here both are ONE pass: a stable radix sort of the stream indices by pixel id puts the previous
event of the pixel next to each event (``den_queue_raw_events``: key kernel, ``den_radix_sort_pairs_u32``,
neighbour kernel, prefix sum of the keep flags); ``den_compact_queued_events`` then writes the kept events in
stream order.  ``colorize_events`` (:278-324, the
Bayer channel of a pixel) is elementwise; ``undistort_events`` (:326-365) calls OpenCV like upstream when
the calibration carries distortion parameters.

Same names, argument meaning and result layout / dtypes as the reference's classmethods (position int64
until undistortion makes it float32; start_ts / end_ts / num_pos / num_neg int64; channel_idx uint8)."""

import ctypes
import os

import numpy as np
import torch

from . import _lib, ops

RAW_EVENTS_FILENAME = "raw_events.npz"                  # data/datasets.py:15-18
TF_EVENTS_FILENAME = "events.pt"
CAMERA_CALIBRATION_FILENAME = "camera_calibration.npz"
MAX_REFRACTORY_PERIOD_FILENAME = "max_refractory_period.pt"
BAYER_CHANNEL = {"R": 0, "G": 1, "B": 2}                # :30-34
INT64_MAX = torch.iinfo(torch.int64).max


def _raw_to_device(raw_events, device):
    position = torch.as_tensor(np.asarray(raw_events["position"]).astype(np.int32)).to(device).contiguous()
    timestamp = torch.as_tensor(np.asarray(raw_events["timestamp"])).to(device)
    polarity = torch.as_tensor(np.asarray(raw_events["polarity"])).to(device)
    if not (len(position) == len(timestamp) == len(polarity)):
        raise ValueError("raw events: position, timestamp and polarity differ in length")
    if timestamp.dtype != torch.int64:
        raise TypeError(f"raw events: timestamps must be int64 (ns), got {timestamp.dtype}")
    return position, timestamp.contiguous(), polarity


def _buffer(scratch, name, shape, dtype, device):
    """A tensor from `scratch` (a dict reused across passes over streams of one size: no allocator traffic in
    a repeated pass) or a fresh one."""
    if scratch is None:
        return torch.empty(shape, dtype=dtype, device=device)
    t = scratch.get(name)
    if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype or t.device != device:
        t = scratch[name] = torch.empty(shape, dtype=dtype, device=device)
    return t


def _stream_pass(position, timestamp, img_height, img_width, scratch=None):
    """den_queue_raw_events over device tensors -> (start_ts (N) int64, kept_offsets (N + 1) int32, min interval
    (1) int64, out-of-range flag (1) int32)."""
    if not position.is_cuda:
        raise NotImplementedError("events: only CUDA tensors are supported (no CPU fallback)")
    n = position.shape[0]
    dev = position.device
    start_ts = _buffer(scratch, "start_ts", (n,), torch.int64, dev)
    offsets = _buffer(scratch, "offsets", (n + 1,), torch.int32, dev)
    min_interval = torch.full((1,), INT64_MAX, dtype=torch.int64, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    nbytes = int(_lib.lib().cdll.den_queue_events_workspace_bytes(n))
    workspace = _buffer(scratch, "workspace", (nbytes,), torch.uint8, dev)
    passes = max(1, (max(int(img_width) * int(img_height) - 1, 1).bit_length() + 7) // 8)
    ops._call("den_queue_raw_events", ops._ptr(position), ops._ptr(timestamp), n, int(img_width),
              int(img_height), ops._ptr(workspace), ctypes.c_size_t(nbytes), ops._ptr(start_ts),
              ops._ptr(offsets), ops._ptr(min_interval), ops._ptr(flag), ops._stream(),
              launches=(5 + 5 * passes) if n else 0)
    return start_ts, offsets, min_interval, flag


def _refractory_tensor(min_interval):
    """:181-183: a 0-d tensor — int64 ns, or float inf when no pixel saw two distinct timestamps."""
    value = int(min_interval.item())
    return torch.tensor(float("inf"), dtype=torch.float64) if value == INT64_MAX else torch.tensor(value)


def _queued(position, timestamp, polarity, start_ts, offsets, flag, scratch=None):
    """den_compact_queued_events: the kept events in stream order.  Reads the kept count and the range flag
    back (the one host synchronisation of the pass).  With `scratch` the results are views of buffers sized
    for the whole stream (valid until the next pass over the same scratch)."""
    n = position.shape[0]
    head = torch.stack((offsets[n], flag[0])).tolist()
    if head[1]:
        raise IndexError("raw events: a position lies outside the img_width x img_height sensor")
    m = int(head[0])
    dev = position.device
    rows = m if scratch is None else n
    out = {"position": _buffer(scratch, "out_position", (rows, 2), torch.int64, dev)[:m]}
    out.update({k: _buffer(scratch, "out_" + k, (rows,), torch.int64, dev)[:m]
                for k in ("start_ts", "end_ts", "num_pos", "num_neg")})
    if m:
        pol = polarity if polarity.dtype in (torch.bool, torch.uint8) else (polarity != 0)
        ops._call("den_compact_queued_events", ops._ptr(position), ops._ptr(timestamp), ops._ptr(pol.contiguous()),
                  ops._ptr(start_ts), ops._ptr(offsets), n, ops._ptr(out["position"]),
                  ops._ptr(out["start_ts"]), ops._ptr(out["end_ts"]), ops._ptr(out["num_pos"]),
                  ops._ptr(out["num_neg"]), ops._stream())
    return out


def queue_raw_events(raw_events, camera_calibration, device="cuda"):
    """``Event.queue_raw_events`` (:186-276) for an in-memory raw stream: dict(position (M, 2) int64, start_ts,
    end_ts, num_pos, num_neg (M) int64) on `device`, the kept events in stream order."""
    position, timestamp, polarity = _raw_to_device(raw_events, torch.device(device))
    start_ts, offsets, _, flag = _stream_pass(position, timestamp, int(camera_calibration["img_height"]),
                                              int(camera_calibration["img_width"]))
    return _queued(position, timestamp, polarity, start_ts, offsets, flag)


def extract_max_refractory_period(raw_events, camera_calibration, device="cuda"):
    """``Event.extract_max_refractory_period`` (:131-183): 0-d tensor, the minimum event interval over the
    per-pixel substreams (events repeating the previous timestamp of their pixel are skipped)."""
    position, timestamp, _ = _raw_to_device(raw_events, torch.device(device))
    _, _, min_interval, flag = _stream_pass(position, timestamp, int(camera_calibration["img_height"]),
                                               int(camera_calibration["img_width"]))
    if int(flag.item()):
        raise IndexError("raw events: a position lies outside the img_width x img_height sensor")
    return _refractory_tensor(min_interval)


def colorize_events(events, camera_calibration):
    """``Event.colorize_events`` (:278-324): adds `channel_idx` (uint8) for a sensor behind a Bayer filter —
    pattern characters = the colours of the top-left, top-right, bottom-left, bottom-right pixel of a 2 x 2
    cell; a monochrome sensor (empty pattern) is returned unchanged."""
    pattern = str(camera_calibration["bayer_pattern"])
    if pattern == "":
        return events
    if len(pattern) != 4 or set(pattern) != set(BAYER_CHANNEL):
        raise ValueError(f"bayer_pattern must be empty or a permutation with R, G, B over four cells, got {pattern!r}")
    position = events["position"]
    table = torch.tensor([BAYER_CHANNEL[c] for c in pattern], dtype=torch.uint8, device=position.device)
    slot = (position[:, 0] % 2 != 0).to(torch.int64) + 2 * (position[:, 1] % 2 != 0).to(torch.int64)
    events["channel_idx"] = table[slot]
    return events


def undistort_events(events, camera_calibration):
    """``Event.undistort_events`` (:326-365): positions to the default float dtype; with distortion parameters
    the same OpenCV calls as upstream (plumb_bob -> cv2.undistortPoints, equidistant -> cv2.fisheye) on the
    host, P = the intrinsics."""
    params = np.asarray(camera_calibration["distortion_params"])
    if len(params) not in (0, 4):
        raise ValueError("distortion_params must hold 0 or 4 values")
    device = events["position"].device
    events["position"] = events["position"].to(torch.get_default_dtype())
    if len(params) == 0:
        return events
    import cv2
    model = str(camera_calibration["distortion_model"])
    K = np.asarray(camera_calibration["intrinsics"])
    points = events["position"].cpu().numpy()
    if model == "plumb_bob":
        out = cv2.undistortPoints(points, K, params, P=K).squeeze(axis=1)
    elif model == "equidistant":
        out = cv2.fisheye.undistortPoints(points[:, None], K, params, P=K).squeeze(axis=1)
    else:
        raise NotImplementedError(f"distortion model {model!r}")
    events["position"] = torch.from_numpy(out).to(device)
    return events


def transform_raw_events(raw_events, camera_calibration, device="cuda"):
    """The raw branch of ``Event.__init__`` (:44-54) — queue, colourise, undistort — and the maximum refractory
    period of the same stream, from ONE sort of the raw events.  Returns (events dict on `device`, 0-d tensor)."""
    position, timestamp, polarity = _raw_to_device(raw_events, torch.device(device))
    start_ts, offsets, min_interval, flag = _stream_pass(
        position, timestamp, int(camera_calibration["img_height"]), int(camera_calibration["img_width"]))
    events = _queued(position, timestamp, polarity, start_ts, offsets, flag)
    events = undistort_events(colorize_events(events, camera_calibration), camera_calibration)
    return events, _refractory_tensor(min_interval)


def load_events(root_directory, device="cuda", cache=True):
    """``Event.__init__`` without the permutation (:36-54): the cached ``events.pt`` if present, else the raw
    stream of ``raw_events.npz`` transformed on the device (and cached in the reference's layout: a dict of CPU
    tensors; ``max_refractory_period.pt`` beside it when absent).  Returns the events dict on `device`."""
    path = os.path.join(root_directory, TF_EVENTS_FILENAME)
    if os.path.isfile(path):
        return {k: v.to(device) for k, v in torch.load(path, weights_only=True).items()}
    raw = np.load(os.path.join(root_directory, RAW_EVENTS_FILENAME))
    calib = np.load(os.path.join(root_directory, CAMERA_CALIBRATION_FILENAME))
    events, refractory = transform_raw_events(raw, calib, device)
    if cache:
        # several ranks may get here at once (one process per GPU): each writes its own temporary file and
        # renames it into place — the rename is atomic, the contents are identical
        _save_atomically({k: v.cpu() for k, v in events.items()}, path)
        refractory_path = os.path.join(root_directory, MAX_REFRACTORY_PERIOD_FILENAME)
        if not os.path.isfile(refractory_path):
            _save_atomically(refractory, refractory_path)
    return events


def _save_atomically(obj, path):
    tmp = f"{path}.{os.getpid()}.tmp"
    torch.save(obj, tmp)
    os.replace(tmp, path)


def sort_pairs(keys, values, key_bits=32):
    """Stable LSD radix sort of (uint32 key, uint32 value) pairs held in int32 / uint32 CUDA tensors on the low
    `key_bits` bits (``den_radix_sort_pairs_u32``).  Returns (sorted keys, values) as int64 tensors."""
    if not keys.is_cuda:
        raise NotImplementedError("events: only CUDA tensors are supported (no CPU fallback)")
    n = keys.shape[0]
    k_in = keys.to(torch.int64).to(torch.int32).contiguous() if keys.dtype != torch.int32 else keys.contiguous()
    v_in = values.to(torch.int64).to(torch.int32).contiguous() if values.dtype != torch.int32 else values.contiguous()
    bufs = [torch.empty(max(n, 1), dtype=torch.int32, device=keys.device) for _ in range(4)]
    nbytes = int(_lib.lib().cdll.den_radix_sort_workspace_bytes(n))
    workspace = torch.empty(nbytes, dtype=torch.uint8, device=keys.device)
    ops._call("den_radix_sort_pairs_u32", ops._ptr(k_in), ops._ptr(v_in), ops._ptr(bufs[0]), ops._ptr(bufs[1]),
              ops._ptr(bufs[2]), ops._ptr(bufs[3]), n, int(key_bits), ops._ptr(workspace), ctypes.c_size_t(nbytes),
              ops._stream())
    mask = 0xFFFFFFFF
    return bufs[0][:n].to(torch.int64) & mask, bufs[1][:n].to(torch.int64) & mask
