"""Data parallelism for the hot path: one process per GPU, NCCL over NVLink/NVSwitch
(``gloo`` on CPU for the tests), mirroring what the reference gets from Lightning's
``DDPPlugin`` (scripts/run.py:84-100): every rank renders its own shard of the event batch
(per-GPU sample budget, models/deblur_e_nerf.py:72-75), parameters / occupancy grid are
replicated, and the gradients are MEAN all-reduced once per optimizer step over ONE flat
fp32 buffer (12 609 346 floats = 50.4 MB for synthetic.yaml).  The batch controller's mean
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


class FlatGradAllReduce:
    """Mean all-reduce of every parameter gradient through one flat buffer per dtype."""

    def __init__(self, parameters):
        self.params = [p for p in parameters if p.requires_grad]
        self._buffers = {}

    def __call__(self):
        n = world_size()
        if n == 1:
            return 0
        nbytes = 0
        by_dtype = {}
        for p in self.params:
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
