"""Host visibility of device-side counters WITHOUT stalling the step: the values are copied to pinned
host memory asynchronously when they are produced and read when they are next needed — one render
call / one optimizer step later (SURVEY.md §7 hard part 5).  Waiting on the copy's event then only
blocks if the device has not yet reached that point of the PREVIOUS step, which bounds how far the host
runs ahead without ever draining the queue."""

import torch


class LaggedReadback:
    """`LaggedReadback(values)` starts the copy of a small device tensor; `pop()` returns its values
    as a list of Python numbers (blocking only until the copy has happened)."""

    _POOL = {}          # (dtype, numel) -> free pinned buffers

    def __init__(self, values):
        values = values.detach().reshape(-1)
        key = (values.dtype, values.numel())
        pool = self._POOL.setdefault(key, [])
        self._key = key
        self._host = pool.pop() if pool else torch.empty(values.numel(), dtype=values.dtype).pin_memory()
        self._host.copy_(values, non_blocking=True)
        # inside a CUDA-graph capture the copy becomes a node of the graph (it refreshes `_host` on
        # every replay); ordering is then the replaying code's business (graph_step.GraphedStep)
        self._event = None
        if not torch.cuda.is_current_stream_capturing():
            self._event = torch.cuda.Event()
            self._event.record()

    def ready(self):
        return self._event is None or self._event.query()

    def pop(self):
        if self._event is not None:
            self._event.synchronize()
        vals = self._host.tolist()
        self._POOL[self._key].append(self._host)
        self._host = None
        return vals
