"""Import / construction stand-in for `lpips.LPIPS` (`loss_metric/metric.py:18`): the perceptual network's
pretrained weights cannot be fetched offline.  So that the reference's validation / test loop still
completes under the façade, evaluating it returns NaN per image (and warns once): `val/lpips` is then NaN
in the logs — visibly missing, never a made-up number."""

import warnings

import torch


class LPIPS(torch.nn.Module):
    def __init__(self, net="alex", **kwargs):
        super().__init__()
        self.net = net
        self._warned = False

    def forward(self, in0, in1, **kwargs):
        if not self._warned:
            warnings.warn("lpips is not installed (pretrained weights unavailable offline): LPIPS is reported as NaN")
            self._warned = True
        return torch.full((in0.shape[0], 1, 1, 1), float("nan"), dtype=in0.dtype, device=in0.device)
