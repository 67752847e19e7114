"""Event batch producer on the device.

The reference draws every training batch on the host: ``IterableMapDataset`` picks
``train_batch_size`` random event indices and indexes the CPU event tensors
(utils/datasets.py:19-32 over ``data/datasets.py`` ``Event``: ``position (E,2) f32``,
``start_ts / end_ts (E) int64 ns``, ``num_pos / num_neg (E) int64``), a ``JoinDataset`` of samplers
yields the normalised supervision intervals (data/datamodule.py:151-213, data/samplers.py: ``ts_diff``
Dirac 1, ``diff_start_ts`` U[0,1), ``ts_subdiff`` triangular with mode 0, ``subdiff_start_ts``
U[0,1), ``interval_gen`` Dirac 0.5 of shape (S-1, N), all float64), and the two ``DataLoader``s
(``num_workers=0``, ``batch_size=1``) ship the result to the GPU each step.  At B200 step times that
host path is the bottleneck, so here the event arrays live in HBM and a batch is a device-side gather
plus three ``torch.rand`` draws on the producer's own generator — no host tensor, no copy.

The draws happen in the reference's order (event indices, then diff_start_ts, ts_subdiff,
subdiff_start_ts) with the reference's formulas, so a producer built on CPU tensors with a CPU
generator reproduces the reference's classes bit for bit (tests/test_oracle_vs_reference.py); on the
device the stream is the CUDA generator's.  Ranks seed with ``seed + rank`` like
data/datamodule.py:87-91.  ``set_batch_size`` is what ``update_train_batch_size``
(models/deblur_e_nerf.py:1277-1285) does to the dataset and the samplers."""

import torch

EVENT_KEYS = ("position", "start_ts", "end_ts", "num_pos", "num_neg")


class EventBatchProducer:
    def __init__(self, events, batch_size, it_sample_size=None, device=None, seed=0, rank=0,
                 dataset_len=None):
        """`events`: mapping with EVENT_KEYS (the reference's transformed events.pt layout);
        `it_sample_size`: S of the pixel-bandwidth model, or None when it is disabled;
        `dataset_len`: use only the first entries (TrimDataset, data/datamodule.py:131-135)."""
        missing = [k for k in EVENT_KEYS if k not in events]
        if missing:
            raise KeyError(f"event arrays missing {missing}")
        device = torch.device(device) if device is not None else events["position"].device
        n = events["position"].shape[0] if dataset_len is None else int(dataset_len)
        assert 0 < n <= events["position"].shape[0]
        self.events = {k: torch.as_tensor(events[k])[:n].to(device).contiguous() for k in EVENT_KEYS}
        if "channel_idx" in events:         # Bayer sensors (data/datasets.py Event: the pixel's colour channel)
            self.events["channel_idx"] = torch.as_tensor(events["channel_idx"])[:n].to(device).contiguous()
        assert self.events["position"].dtype == torch.float32
        for k in EVENT_KEYS[1:]:
            assert self.events[k].dtype == torch.int64 and self.events[k].shape == (n,)
        self.device = device
        self.batch_size = int(batch_size)
        self.it_sample_size = it_sample_size
        self.generator = torch.Generator(device=device)
        self.generator.manual_seed(int(seed) + int(rank))
        # On a CUDA device the dozen small launches of a draw are captured ONCE per batch size in a CUDA
        # graph (the producer's generator registered with it, so every replay advances the Philox
        # offset) and replayed: ~1.2 ms of eager launches per batch become one graph launch.  Two
        # buffers alternate, so the batch handed out before the current one stays intact (the trainer
        # prefetches one batch).  A batch size is captured the second time it is asked for.
        self.use_graph = device.type == "cuda"
        self._graphs = {}           # batch size -> [seen, [(graph, batch), (graph, batch)], next buffer]

    def __len__(self):
        return self.events["position"].shape[0]

    def set_batch_size(self, batch_size):
        self.batch_size = max(int(batch_size), 1)

    def next_batch(self):
        """{"event": {...}, "normalized": {...}} for `batch_size` events, everything on `device`."""
        if not self.use_graph or torch.cuda.is_current_stream_capturing():
            return self._draw()
        entry = self._graphs.setdefault(self.batch_size, [0, [], 0])
        entry[0] += 1
        if entry[0] < 2:
            return self._draw()
        if len(entry[1]) < 2:                   # capture this buffer's graph
            if len(self._graphs) > 8:
                for key in [k for k in self._graphs if k != self.batch_size][:4]:
                    del self._graphs[key]
            graph = torch.cuda.CUDAGraph()
            graph.register_generator_state(self.generator)
            torch.cuda.synchronize()
            with torch.cuda.graph(graph):
                batch = self._draw()
            entry[1].append((graph, batch))
        graph, batch = entry[1][entry[2] % len(entry[1])] if len(entry[1]) == 2 else entry[1][-1]
        entry[2] += 1
        graph.replay()
        return batch

    def _draw(self):
        n, g, dev = self.batch_size, self.generator, self.device
        index = torch.randint(len(self), size=(n,), generator=g, device=dev)
        event = {k: v[index] for k, v in self.events.items()}
        f64 = dict(dtype=torch.float64, generator=g, device=dev)
        diff_start = torch.rand(n, **f64)
        u = torch.rand(n, **f64)
        # TriangularSampler(low=0, high=1, mode=0): mode_cum_prob = 0, k1 = 0, k2 = 1
        subdiff = torch.where(u <= 0.0, 0.0 + torch.sqrt(u * 0.0), 1.0 - torch.sqrt((1 - u) * 1.0))
        normalized = {
            "ts_diff": torch.full((n,), 1.0, dtype=torch.float64, device=dev),
            "diff_start_ts": diff_start,
            "ts_subdiff": subdiff,
            "subdiff_start_ts": torch.rand(n, **f64),
        }
        if self.it_sample_size is not None:
            normalized["interval_gen"] = torch.full((self.it_sample_size - 1, n), 0.5,
                                                    dtype=torch.float64, device=dev)
        return {"event": event, "normalized": normalized}

    def __iter__(self):
        while True:
            yield self.next_batch()
