"""Data parallelism for the hot path: one process per GPU, NCCL over NVLink/NVSwitch
(``gloo`` on CPU for the tests), mirroring what the reference gets from Lightning's
``DDPPlugin`` (scripts/run.py:84-100): every rank renders its own shard of the event batch
(per-GPU sample budget, models/deblur_e_nerf.py:72-75), parameters / occupancy grid are
replicated, and the gradients are all-reduced once per optimizer step over ONE flat fp32 buffer
(12 609 346 floats = 50.4 MB for synthetic.yaml) that the parameters' ``.grad`` tensors are views of
(`GradReducer`: no copy in or out, SUM over NVLink, the 1 / N folded into ``den_adam_step``).  The batch controller's mean
samples-per-ray is averaged across ranks like ``self.all_gather(...).mean()``
(models/deblur_e_nerf.py:1269-1272).

No collective is fused into the hash-gradient scatter on purpose: that would move ~16 GB of
atomics per step over NVLink instead of one 50 MB reduction.
"""

import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's env (RANK/LOCAL_RANK/WORLD_SIZE/MASTER_*).
    Returns (rank, local_rank, world_size); a no-op single-process setup without them."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, local_rank, world


def world_size():
    return dist.get_world_size() if dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_initialized() else 0


def gather_view_outputs(outputs, device=None):
    """The `self.all_gather(outputs)` of evaluation_epoch_end (models/deblur_e_nerf.py:672): every rank
    rendered views rank, rank + N, rank + 2N, ... of the evaluation set; returns the output dicts of ALL
    views in their original order on every rank (images stay on the device: one all_gather per field over
    stacks padded to the largest per-rank count).  Fewer views than ranks is refused on EVERY rank (the counts
    are exchanged first, so no rank is left waiting in a collective); `device`: where to exchange them when
    this rank rendered nothing."""
    n = world_size()
    if n == 1:
        return outputs
    if device is None and outputs:
        device = outputs[0]["pred_intensity_img"].device
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    count = torch.tensor([len(outputs)], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(count) for _ in range(n)]
    dist.all_gather(counts, count)
    counts = [int(c.item()) for c in counts]
    if min(counts) == 0:
        raise ValueError(f"gather_view_outputs: {sum(counts)} view(s) for {n} ranks — every rank needs at least one")
    most = max(counts)
    gathered = {}
    for key in ("pred_intensity_img", "target_intensity_img", "exposure_time", "gain"):
        mine = torch.stack([torch.as_tensor(o[key]).to(device) for o in outputs])
        if len(outputs) < most:
            mine = torch.cat((mine, mine[-1:].expand(most - len(outputs), *mine.shape[1:])))
        parts = [torch.empty_like(mine) for _ in range(n)]
        dist.all_gather(parts, mine.contiguous())
        gathered[key] = parts
    merged = []
    for i in range(sum(counts)):
        r, j = i % n, i // n
        merged.append({"sample_id": None, **{k: v[r][j] for k, v in gathered.items()}})
    return merged


def broadcast_parameters(module, src=0):
    """Replicate rank `src`'s parameters and buffers (DDP does this at construction)."""
    if world_size() == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        data = t.data
        if data.is_contiguous():
            dist.broadcast(data, src)
        else:           # e.g. a transposed rotation buffer: broadcast a packed copy
            packed = data.contiguous()
            dist.broadcast(packed, src)
            data.copy_(packed)


class GradReducer:
    """Gradient reduction of one optimizer step with NO copy passes: every trainable parameter's
    ``.grad`` is a VIEW into one flat buffer per dtype that this object owns (the 12.6 M-entry hash
    table gradient is 99.9 % of it), autograd accumulates into the views, and a step is ONE
    ``all_reduce(SUM)`` over the flat buffer — the division by the world size is folded into the
    Adam update (``optimizer.grad_scale = 1 / world``, applied inside ``den_adam_step``).  Replaces
    Lightning's ``DDPPlugin`` gradient averaging (scripts/run.py:84-100).

    The same parameters take part at every world size: the views exist for a single process too, so a
    parameter that received no gradient in a step is stepped with a zero gradient (moments decay,
    weight decay applies) at N = 1 exactly as at N > 1 — under DDP every trainable parameter gets a
    gradient every step, a parameter that never does is frozen and holds no view.

    Usage: ``reducer = GradReducer(model); reducer.bind(optimizer)``; per step
    ``optimizer.zero_grad(set_to_none=False)`` (or ``reducer.zero_()``: one memset), backward, then
    ``reducer()`` and ``optimizer.step()``.  ``zero_grad(set_to_none=True)`` drops the views;
    ``reducer()`` re-attaches them (copying what autograd produced) so nothing silently breaks."""

    def __init__(self, module_or_params, async_op=False):
        params = module_or_params.parameters() if hasattr(module_or_params, "parameters") \
            else module_or_params
        self.params = [p for p in params if p.requires_grad]
        self.flat = {}              # dtype -> flat gradient buffer
        self.views = {}             # id(param) -> its view
        by_dtype = {}
        for p in self.params:
            by_dtype.setdefault((p.dtype, p.device), []).append(p)
        for (dtype, device), group in by_dtype.items():
            # every view starts on a 16-byte boundary (the Adam kernel reads float4)
            align = max(16 // torch.empty((), dtype=dtype).element_size(), 1)
            sizes = [-(-p.numel() // align) * align for p in group]
            flat = torch.zeros(sum(sizes), dtype=dtype, device=device)
            off = 0
            for p, size in zip(group, sizes):
                self.views[id(p)] = flat[off:off + p.numel()].view_as(p)
                off += size
            self.flat[(dtype, device)] = flat
        self.attach()
        self._work = []

    def attach(self):
        """Point every ``.grad`` at its view (keeping whatever gradient autograd left there)."""
        for p in self.params:
            view = self.views[id(p)]
            if p.grad is None:
                view.zero_()
            elif p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
            else:
                continue
            p.grad = view

    def zero_(self):
        for flat in self.flat.values():
            flat.zero_()
        self.attach()

    def bind(self, optimizer):
        """Fold the mean into the optimizer's update instead of a division pass over the buffer."""
        optimizer.grad_scale = 1.0 / world_size()
        return optimizer

    @property
    def nbytes(self):
        return sum(f.numel() * f.element_size() for f in self.flat.values())

    def __call__(self):
        """SUM all-reduce of the flat buffer(s) on the current stream; returns the bytes reduced."""
        self.attach()
        if world_size() == 1:
            return 0
        for flat in self.flat.values():
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        return self.nbytes


class FlatGradAllReduce:
    """Mean all-reduce of the parameter gradients through one flat buffer per dtype, for ANY optimizer
    (the copy-in / copy-out form; `GradReducer` is the copy-free one the hot path uses with FusedAdam).
    A parameter whose gradient is None on EVERY rank keeps None (the optimizer skips it, as in a
    single process); one that has a gradient on some rank takes part everywhere (zeros elsewhere)."""

    def __init__(self, parameters):
        self.params = [p for p in parameters if p.requires_grad]
        self._buffers = {}

    def __call__(self):
        n = world_size()
        if n == 1:
            return 0
        dev = self.params[0].device if self.params else torch.device("cpu")
        used = torch.tensor([0 if p.grad is None else 1 for p in self.params], dtype=torch.int32,
                            device=dev)
        dist.all_reduce(used, op=dist.ReduceOp.MAX)          # the same participant set on all ranks
        used = used.tolist()
        nbytes = 0
        by_dtype = {}
        for p, flag in zip(self.params, used):
            if not flag:
                continue
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            by_dtype.setdefault(p.grad.dtype, []).append(p)
        for dtype, params in by_dtype.items():
            total = sum(p.numel() for p in params)
            flat = self._buffers.get(dtype)
            if flat is None or flat.numel() != total or flat.device != params[0].device:
                flat = torch.empty(total, dtype=dtype, device=params[0].device)
                self._buffers[dtype] = flat
            off = 0
            for p in params:
                flat[off:off + p.numel()].copy_(p.grad.reshape(-1))
                off += p.numel()
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            flat.div_(n)
            off = 0
            for p in params:
                p.grad.copy_(flat[off:off + p.numel()].view_as(p.grad))
                off += p.numel()
            nbytes += flat.numel() * flat.element_size()
        return nbytes


def attach(renderer):
    """Wire the cross-rank mean of samples-per-ray into the batch controller."""
    if world_size() == 1:
        return renderer

    def reduce_mean(value):
        if torch.is_tensor(value):            # sync-free steps: the mean stays on the device
            t = value.detach().clone()
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return t / world_size()
        dev = next(renderer.parameters()).device
        t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item()) / world_size()

    renderer.mean_samples_reduce_fn = reduce_mean
    return renderer


def max_over_ranks(value, device):
    if world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device):
    if world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    if world_size() > 1:
        dist.barrier()
