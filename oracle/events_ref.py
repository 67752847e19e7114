"""CPU restatement of the raw-event preprocessing in front of the hot path (TEST INFRASTRUCTURE — never
imported by the product path): ``Event.queue_raw_events`` (data/datasets.py:186-276),
``Event.extract_max_refractory_period`` (:131-183) and ``Event.colorize_events`` (:278-324).

Two forms of each stream pass: ``*_loop`` follows the reference's per-event loop literally (a two-entry
sliding window per pixel; small cases only), the other is a vectorised numpy form (stable argsort by pixel
id) for sizes the loop cannot reach.  Both are pinned against the reference's OWN classmethods run here
(tests/test_oracle_vs_reference.py) and against the committed golden (tests/golden/raw_events.npz, written
by tests/golden/make_golden.py from the reference's methods)."""

import collections

import numpy as np

BAYER_CHANNEL = {"R": 0, "G": 1, "B": 2}                                      # :30-34


def queue_raw_events_loop(position, timestamp, polarity, height, width):
    """:186-276, loop for loop.  position (N, 2) (x, y); timestamp (N) int64; polarity (N) bool.
    Returns dict(position int64 (M, 2), start_ts, end_ts, num_pos, num_neg (M) of timestamp's dtype)."""
    position = np.asarray(position).astype(np.int64)
    polarity = np.asarray(polarity).astype(np.int64)
    timestamp = np.asarray(timestamp)
    ts_win = [[collections.deque(maxlen=2) for _ in range(width)] for _ in range(height)]
    pol_win = [[collections.deque(maxlen=2) for _ in range(width)] for _ in range(height)]
    n = len(position)
    start_ts = np.empty_like(timestamp)
    num_pos = np.empty_like(timestamp)
    num_neg = np.empty_like(timestamp)
    valid = np.ones(n, dtype=bool)
    for i in range(n):
        x, y = position[i]
        tw, pw = ts_win[y][x], pol_win[y][x]
        tw.append(timestamp[i])
        pw.append(polarity[i])
        if len(tw) < tw.maxlen or tw[0] == tw[-1]:                             # :246-253
            valid[i] = False
            continue
        start_ts[i] = tw[0]
        num_pos[i] = sum(pw) - pw[0]                                            # :260-263
        num_neg[i] = (pw.maxlen - 1) - num_pos[i]                               # :264-267
    return {"position": position[valid], "start_ts": start_ts[valid], "end_ts": timestamp[valid],
            "num_pos": num_pos[valid], "num_neg": num_neg[valid]}


def _previous_at_pixel(position, height, width):
    """For every event the index of the previous event of the stream at the same pixel, or -1."""
    position = np.asarray(position).astype(np.int64)
    if ((position[:, 0] < 0) | (position[:, 0] >= width) | (position[:, 1] < 0) | (position[:, 1] >= height)).any():
        raise IndexError("event position outside the sensor")
    key = position[:, 1] * width + position[:, 0]
    order = np.argsort(key, kind="stable")
    prev = np.full(len(key), -1, dtype=np.int64)
    same = key[order][1:] == key[order][:-1]
    prev[order[1:][same]] = order[:-1][same]
    return prev


def queue_raw_events(position, timestamp, polarity, height, width):
    """The same result as the loop, from the previous event at the pixel: the window after appending event i
    is [prev(i), i], so i is kept iff prev(i) exists and carries a different timestamp; num_pos is the
    event's own polarity (the earlier entry only dates the interval)."""
    position = np.asarray(position).astype(np.int64)
    timestamp = np.asarray(timestamp)
    pol = np.asarray(polarity).astype(timestamp.dtype)
    prev = _previous_at_pixel(position, height, width)
    has = prev >= 0
    prev_ts = np.where(has, timestamp[np.maximum(prev, 0)], 0)
    valid = has & (prev_ts != timestamp)
    return {"position": position[valid], "start_ts": prev_ts[valid].astype(timestamp.dtype),
            "end_ts": timestamp[valid], "num_pos": pol[valid], "num_neg": (1 - pol)[valid]}


def max_refractory_period_loop(position, timestamp, height, width):
    """:131-183, loop for loop: the minimum interval between consecutive DISTINCT timestamps of a pixel.
    Returns a numpy scalar (float inf when no pixel saw two distinct timestamps)."""
    windows = [[collections.deque(maxlen=2) for _ in range(width)] for _ in range(height)]
    best = np.array(float("inf"))
    for (x, y), ts in zip(np.asarray(position), np.asarray(timestamp)):
        win = windows[y][x]
        if len(win) > 0 and ts == win[-1]:                                      # :163-168
            continue
        win.append(ts)
        if len(win) < 2:
            continue
        best = min(best, win[1] - win[0])
    return best


def max_refractory_period(position, timestamp, height, width):
    """An event equal in time to the window's last entry is skipped, so that entry always equals the previous
    event's timestamp: the interval of a kept event is timestamp[i] - timestamp[prev(i)]."""
    timestamp = np.asarray(timestamp)
    prev = _previous_at_pixel(position, height, width)
    has = prev >= 0
    diff = timestamp[has] - timestamp[prev[has]]
    diff = diff[diff != 0]
    return diff.min() if diff.size else np.array(float("inf"))


def colorize_events(position, bayer_pattern):
    """:278-324: channel index (uint8) of every event of a sensor behind a Bayer filter, from the parity of
    its pixel position: pattern characters are the top-left, top-right, bottom-left, bottom-right colours.
    Returns None for a monochrome sensor (empty pattern)."""
    if bayer_pattern == "":
        return None
    assert len(bayer_pattern) == 4 and set(bayer_pattern) == set(BAYER_CHANNEL)
    position = np.asarray(position).astype(np.int64)
    x_even, y_even = position[:, 0] % 2 == 0, position[:, 1] % 2 == 0
    slot = np.where(y_even, np.where(x_even, 0, 1), np.where(x_even, 2, 3))
    table = np.array([BAYER_CHANNEL[c] for c in bayer_pattern], dtype=np.uint8)
    return table[slot]
