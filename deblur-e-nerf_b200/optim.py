"""Adam for the hot path's parameters: ``FusedAdam`` mirrors
``torch.optim.Adam`` as the reference configures it (models/deblur_e_nerf.py:1055-1112 — parameter
groups with their own ``lr`` / ``weight_decay``, betas (0.9, 0.999), eps 1e-8, L2 decay added to the
gradient) and updates every fp32 CUDA parameter through the ``den_adam_step`` kernel: two launches per
step (the 12.6 M-entry hash table on its own, the small MLP / background / event-model tensors
together) instead of torch's per-dtype multi-tensor passes.  The float64 refractory-period scalar is
updated with the same formula in torch (one element).  ``state_dict`` keys (``step``, ``exp_avg``,
``exp_avg_sq``) follow ``torch.optim.Adam``, so a learning-rate scheduler (MultiStepLR, :1096-1101)
and checkpoints work unchanged.

``capturable=True`` keeps the step number in a device tensor that is incremented on the device, so a
step captured in a CUDA graph (graph_step.GraphedStep) applies the right bias correction on every
replay; ``state[p]["step"]`` is then refreshed from the host-side replay count."""

import ctypes
import math

import torch

from . import ops
from ._lib import AdamTensor


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0 or eps < 0 or weight_decay < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1):
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        # gradients are multiplied by this inside the update: ddp.GradReducer sets 1 / world_size so the
        # SUM all-reduce needs no separate division pass
        self.grad_scale = 1.0
        self.capturable = False
        self._step_dev = None           # device int64 step counter (capturable mode)
        # device int32: non-zero -> this step's update is skipped on the device (NeRF.overflow_flag: a
        # sync-free step whose sample buffers overflowed lost samples and must not be applied)
        self.skip_flag = None

    def enable_capture(self, device):
        """Switch to the device-side step counter (call once, before capturing a step)."""
        steps = {int(s["step"]) for s in self.state.values() if "step" in s}
        if len(steps) > 1:
            raise RuntimeError("capturable FusedAdam needs every parameter at the same step")
        start = steps.pop() if steps else 0
        self._step_dev = torch.full((), start, dtype=torch.int64, device=device)
        self.capturable = True

    def note_replayed_steps(self, n=1):
        """A captured step was replayed `n` times: keep the host-side step numbers (checkpoints,
        state_dict) in line with the device counter."""
        for s in self.state.values():
            if "step" in s:
                s["step"] += n

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        skip = self.skip_flag
        if self.capturable:
            self._step_dev += 1 if skip is None else (skip == 0).to(torch.int64)
        launches = {}                  # (beta1, beta2, eps, step) -> [AdamTensor, ...]
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("FusedAdam does not support sparse gradients")
                state = self.state[p]
                if not state:
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                state["step"] += 1
                t = int(state["step"])
                m, v = state["exp_avg"], state["exp_avg_sq"]
                if p.dtype == torch.float32 and p.is_cuda and p.is_contiguous():
                    g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                    entry = AdamTensor(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(),
                                       p.numel(), float(group["lr"]), float(group["weight_decay"]))
                    key = (beta1, beta2, group["eps"], 0 if self.capturable else t)
                    launches.setdefault(key, []).append((entry, g))
                elif not p.is_cuda:
                    raise NotImplementedError("FusedAdam: only CUDA parameters are supported (no CPU fallback)")
                else:
                    # e.g. the float64 refractory-period scalar: the same formula in torch
                    g = p.grad * self.grad_scale if self.grad_scale != 1.0 else p.grad
                    g = g.add(p, alpha=group["weight_decay"]) if group["weight_decay"] else g
                    if skip is not None:            # same rule as the kernel, elementwise on the scalar
                        keep = (skip == 0)
                        m_new = m * beta1 + g * (1 - beta1)
                        v_new = v * beta2 + g * g * (1 - beta2)
                        td = self._step_dev.to(v.dtype) if self.capturable else torch.tensor(
                            float(t), dtype=v.dtype, device=v.device)
                        denom = v_new.sqrt() / torch.sqrt(1 - beta2 ** td) + group["eps"]
                        p_new = p - group["lr"] * m_new / denom / (1 - beta1 ** td)
                        m.copy_(torch.where(keep, m_new, m))
                        v.copy_(torch.where(keep, v_new, v))
                        p.copy_(torch.where(keep, p_new, p))
                        continue
                    m.mul_(beta1).add_(g, alpha=1 - beta1)
                    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
                    if self.capturable:
                        td = self._step_dev.to(v.dtype)
                        denom = (v.sqrt() / torch.sqrt(1 - beta2 ** td)).add_(group["eps"])
                        p.sub_(group["lr"] * m / denom / (1 - beta1 ** td))
                    else:
                        denom = (v.sqrt() / math.sqrt(1 - beta2 ** t)).add_(group["eps"])
                        p.addcdiv_(m, denom, value=-group["lr"] / (1 - beta1 ** t))
        for (beta1, beta2, eps, t), entries in launches.items():
            arr = (AdamTensor * len(entries))(*[e for e, _ in entries])
            ops.adam_step(arr, len(entries), beta1, beta2, eps, t, self.grad_scale,
                          self._step_dev if self.capturable else None, skip)
        return loss
