"""One optimizer step as ONE CUDA-graph launch.

With the sync-free render path (nerf.NeRF.render_chunk_sync_free: capacity-sized sample buffers,
device-side counts) a training step enqueues ~150 kernels and copies without ever reading the device
back, so the whole of it — zero_grad, `accumulate_grad_batches` x (training_step + backward), the
gradient all-reduce, Adam — can be captured once and replayed: the host cost of a step drops from
milliseconds of Python to one `cudaGraphLaunch`.  That is what the strong-scaling curve needs: at 8
GPUs a rank's share of the 2^17-ray batch is ~7 ms of kernels against ~9 ms of eager host work
(profiles/r01_host_profile_budget.md).

What stays outside the graph (host decisions): the occupancy-grid update every 16th optimizer step
(models/nerf.py:200-204; it writes the grid buffers in place, which the captured march reads), the
batch controller (fed by the lagged counts; a new batch size means a new capture) and learning-rate
changes (baked into the captured Adam launches: a changed lr re-captures).

A capture is valid for one batch layout (tensor shapes of the batch dict) and one sample capacity.  The
counts of every replay are copied to pinned memory inside the graph; the host looks at them one step
late, and an overflow (more samples than the captured capacity: the step lost samples and is flagged,
`overflows`) enlarges the estimate and re-captures.
"""

import torch


def _signature(batches, optimizer):
    shapes = tuple((k, kk, tuple(v.shape), v.dtype) for b in batches for k, d in sorted(b.items())
                   for kk, v in sorted(d.items()))
    lrs = tuple((float(g["lr"]), float(g["weight_decay"])) for g in optimizer.param_groups)
    return shapes, lrs, getattr(optimizer, "grad_scale", 1.0)


class GraphedStep:
    def __init__(self, model, optimizer, reducer, accumulate_grad_batches=1, warmup_steps=2):
        self.model, self.optimizer, self.reducer = model, optimizer, reducer
        self.acc = accumulate_grad_batches
        self.warmup_steps = warmup_steps        # eager steps before a capture (caches, capacity estimate)
        self._eager_done = 0
        self._graph = None
        self._sig = None
        self._static = None
        self._loss = None
        self._events = []
        self.captures = 0
        self.replays = 0
        self.overflows = 0
        self.eager_fallbacks = 0                # steps run eagerly because they did not fit the batched path
        self.launches_per_replay = 0            # den_b200 kernels recorded in the captured step
        self._side = None                       # every step of this object runs on one side stream

    # ----------------------------------------------------------------- pieces --------
    def _eager(self, batches, global_step):
        opt, acc = self.optimizer, self.acc
        opt.zero_grad(set_to_none=False)
        loss = None
        dev = next(self.model.parameters()).device
        for m, batch in enumerate(batches):
            if not next(iter(batch["event"].values())).is_cuda:        # a (pinned) host batch
                batch = {k: {kk: v.to(dev, non_blocking=True) for kk, v in d.items()}
                         for k, d in batch.items()}
            loss = self.model.training_step(batch, m, global_step)
            (loss / acc if acc > 1 else loss).backward()
        self.reducer()
        opt.step()
        return loss

    def _capture(self, batches, global_step):
        dev = next(self.model.parameters()).device
        if not getattr(self.optimizer, "capturable", False):
            self.optimizer.enable_capture(dev)
        self._static = [{k: {kk: torch.empty_like(v, device=dev) for kk, v in d.items()}
                         for k, d in b.items()} for b in batches]
        self._copy_in(batches)
        nerf = self.model.nerf
        nerf._consume_stats()                           # host waits belong before the capture
        self.model._apply_lagged_controller()
        update = nerf.update_occ_grid
        nerf.update_occ_grid = lambda *a, **k: None         # host-gated: runs outside the graph
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        from . import ops
        launches0 = ops.launch_count()
        try:
            with torch.cuda.graph(graph, stream=self._stream()):
                self._loss = self._eager(self._static, global_step)
        finally:
            nerf.update_occ_grid = update
        self.launches_per_replay = ops.launch_count() - launches0
        if nerf._stats is None:
            raise RuntimeError("the captured step took the synchronising render path: no capture")
        self._graph = graph
        self._stats_host = nerf._stats._host           # refreshed by every replay (pinned memory)
        nerf._stats = None
        pending = getattr(self.model, "_pending_controller", None)
        self._controller_host = pending[0]._host if pending else None
        self.model._pending_controller = None
        self.captures += 1

    def _stream(self):
        # eager warm-up steps and the capture share one side stream, so the autograd nodes the capture
        # meets (AccumulateGrad of every parameter) were created on the stream that is being captured
        if self._side is None:
            self._side = torch.cuda.Stream()
        return self._side

    def _copy_in(self, batches):
        for dst, src in zip(self._static, batches):
            for k, d in src.items():
                for kk, v in d.items():
                    dst[k][kk].copy_(v, non_blocking=True)

    def _lagged_host_work(self):
        """Look at the counts of the replay before last (certainly finished: no stall)."""
        if len(self._events) < 2:
            return
        self._events.pop(0).synchronize()
        marched, n_rays, overflow = self._stats_host.tolist()
        nerf = self.model.nerf
        if overflow:
            import sys
            print(f"den_b200: sample capacity overflow in a replayed step (marched {marched:.0f} samples for "
                  f"{n_rays:.0f} rays, estimate {nerf._spr_estimate} samples/ray): re-capturing",
                  file=sys.stderr)
            self.overflows += 1
            nerf.overflow_count += 1
            nerf._spr_estimate = 1.5 * max(nerf._spr_estimate or 0.0, marched / max(n_rays, 1.0))
            self._graph = None                          # re-capture with the larger capacity
        if self._controller_host is not None:
            mean, largest = self._controller_host.tolist()
            self.model._last_mean_samples = largest
            self.model.next_train_batch_size = int(self.model.train_ray_sample_batch_size / max(mean, 1e-9))

    # ------------------------------------------------------------------- step --------
    def __call__(self, batches, global_step):
        """`batches`: the `accumulate_grad_batches` micro-batches of this optimizer step (batch dicts of
        device or pinned host tensors).  Returns the (device) loss of the last micro-batch."""
        model = self.model
        cfg = model.nerf.occ_grid_config
        every = cfg["n"] if isinstance(cfg, dict) else cfg.n
        if global_step % every == 0:                    # models/deblur_e_nerf.py:465-469, outside the graph
            model.nerf.update_occ_grid(step=global_step, T_wc_position=model.trajectory.T_wc_position)
        sig = _signature(batches, self.optimizer)
        if self._eager_done < self.warmup_steps or not model.nerf.sync_free:
            self._eager_done += 1
            return self._eager_no_update(batches, global_step)
        self._lagged_host_work()
        if self._graph is None or sig != self._sig:
            if model.nerf._capacity(1) is None:          # no estimate (e.g. after an overflow): learn it
                return self._eager_no_update(batches, global_step)
            if not all(model.step_fits_batched(b) for b in batches):
                # the memory guard would split this step into sequential render calls with host
                # read-backs (e.g. the enlarged estimate after an overflow on a 180 GB-filling batch):
                # nothing a graph can record — run it eagerly, try to capture the next one
                self._graph = None
                self._static = None
                self._loss = None
                self.eager_fallbacks += 1
                return self._eager_no_update(batches, global_step)
            self._capture(batches, global_step)         # records the step (nothing executes yet) ...
            self._sig = sig
            self._events = []
            fresh = True                                # ... the host-side step counters moved once
        else:
            self._copy_in(batches)
            fresh = False
        self._graph.replay()
        self.replays += 1
        if not fresh:
            self.optimizer.note_replayed_steps(1)
        ev = torch.cuda.Event()
        ev.record()
        self._events.append(ev)
        return self._loss

    def _eager_no_update(self, batches, global_step):
        nerf = self.model.nerf
        update = nerf.update_occ_grid
        nerf.update_occ_grid = lambda *a, **k: None         # already done above for this step
        side, cur = self._stream(), torch.cuda.current_stream()
        side.wait_stream(cur)
        try:
            with torch.cuda.stream(side):
                loss = self._eager(batches, global_step)
        finally:
            nerf.update_occ_grid = update
        cur.wait_stream(side)
        return loss
