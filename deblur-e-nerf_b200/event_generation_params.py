"""B2 host mirrors of ``ContrastThreshold`` and ``RefractoryPeriod``
(models/event_generation_params.py:8-237): same parametrizations and state-dict keys
(``parametrizations.<name>.original``), same property names.  The constructors take the
reference's ``dataset_directory`` (reads ``camera_calibration.npz`` /
``max_refractory_period.pt`` like the reference) or an in-memory calibration dict.

Tiny elementwise math on (N,) tensors — the source of the C_p and tau gradients; it stays
torch (plumbing around the kernels)."""

import os
import warnings

import numpy as np
import torch

from .nerf import Softplus


def load_calibration(source):
    if isinstance(source, (str, os.PathLike)):
        return dict(np.load(os.path.join(source, "camera_calibration.npz")))
    return dict(source)


class ScaledShiftedSigmoid(torch.nn.Module):
    """utils/modules.py:78-94."""

    def __init__(self, low=0, high=1):
        super().__init__()
        self.low, self.scale = low, high - low

    def forward(self, x):
        return self.scale * torch.sigmoid(x / self.scale) + self.low

    def right_inverse(self, y):
        return self.scale * torch.logit((y - self.low) / self.scale)


class ContrastThreshold(torch.nn.Module):
    def __init__(self, dataset_directory, parameterize_mean_ct=True):
        super().__init__()
        if not parameterize_mean_ct:
            raise NotImplementedError("legacy parameterize_mean_ct=False is not used by any "
                                      "shipped config")
        calib = load_calibration(dataset_directory)
        pos = torch.from_numpy(np.asarray(calib["pos_contrast_threshold"]))
        neg = torch.from_numpy(np.asarray(calib["neg_contrast_threshold"]))
        ratio, mean = pos / neg, (pos + neg) / 2
        assert ratio > 0 and mean > 0
        self.register_buffer("init_p2n_contrast_threshold_ratio", ratio, persistent=False)
        self.register_buffer("init_mean_contrast_threshold", mean, persistent=False)
        self.p2n_contrast_threshold_ratio = torch.nn.parameter.Parameter(ratio.clone())
        torch.nn.utils.parametrize.register_parametrization(
            self, "p2n_contrast_threshold_ratio", Softplus())
        self.mean_contrast_threshold = torch.nn.parameter.Parameter(mean.clone())
        torch.nn.utils.parametrize.register_parametrization(
            self, "mean_contrast_threshold", Softplus())

    @property
    def neg_contrast_threshold(self):
        return 2 * self.mean_contrast_threshold / (self.p2n_contrast_threshold_ratio + 1)

    @property
    def pos_contrast_threshold(self):
        return self.p2n_contrast_threshold_ratio * self.neg_contrast_threshold

    def forward(self, input_event):
        out = dict(input_event)
        out["log_intensity_diff"] = (out.pop("num_pos") * self.pos_contrast_threshold
                                     - out.pop("num_neg") * self.neg_contrast_threshold)
        return out


class RefractoryPeriod(torch.nn.Module):
    REDEFINED_CALIBRATED_REFRACTORY_PERIOD_FACTOR = 0.999
    MIN_SCALED_SHIFTED_SIGMOID_GRAD_MAGNITUDE = 0.0001

    def __init__(self, dataset_directory, max_refractory_period=None):
        super().__init__()
        calib = load_calibration(dataset_directory)
        tau = torch.from_numpy(np.asarray(calib["refractory_period"]))
        if max_refractory_period is None:
            path = os.path.join(dataset_directory, "max_refractory_period.pt")
            max_refractory_period = torch.load(path, weights_only=True)     # a 0-d tensor
        tau_max = torch.as_tensor(max_refractory_period)
        if not (0 <= tau < tau_max):
            warnings.warn(f"Calibrated refractory period ({tau}) >= max. possible ({tau_max}).")
            tau = self.REDEFINED_CALIBRATED_REFRACTORY_PERIOD_FACTOR * tau_max
        self.register_buffer("init_refractory_period", tau, persistent=False)
        self.register_buffer("max_refractory_period", tau_max, persistent=False)
        self.register_buffer(
            "max_scaled_logit_magnitude",
            torch.tensor(self.MIN_SCALED_SHIFTED_SIGMOID_GRAD_MAGNITUDE).logit().abs(),
            persistent=False)
        self._refractory_period = torch.nn.parameter.Parameter(tau.to(torch.float64))
        torch.nn.utils.parametrize.register_parametrization(
            self, "_refractory_period", ScaledShiftedSigmoid(low=0, high=tau_max))
        self.clamp_refractory_period()

    @torch.no_grad()
    def clamp_refractory_period(self):
        orig = self.parametrizations._refractory_period.original
        scaled = (orig / self.max_refractory_period).clamp(
            min=-self.max_scaled_logit_magnitude, max=self.max_scaled_logit_magnitude)
        orig.copy_(self.max_refractory_period * scaled)

    @property
    def refractory_period(self):
        self.clamp_refractory_period()
        return self._refractory_period

    def forward(self, input_event):
        out = dict(input_event)
        out["start_ts"] = out["start_ts"] + self.refractory_period
        return out
