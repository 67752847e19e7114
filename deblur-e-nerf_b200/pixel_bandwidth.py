"""B2 host mirror of ``PixelBandwidth`` (models/pixel_bandwidth.py:7-494): the 4th-order
pixel-bandwidth low-pass filter that turns S rendered intensities per event into one
band-limited log-intensity, with the differencing-amplifier reset state carried from the
first render call of a step to the other three (with gradient, :419-446).

Same constructor arguments (``dataset_directory`` may also be an in-memory calibration dict),
same parameters behind ``parametrizations.<name>.original`` for ``tau_mil_it_eff_prod,
A_amp_inv, A_loop_inv, tau_out, tau_sf, tau_diff`` (:137-153), same ``forward`` signature and
return value.  The linearise / discretise / weight recursion / normalised sum is ONE
den_b200 kernel per direction (``den_lpf_fwd`` / ``den_lpf_bwd``); the sample schedule and
the reset bookkeeping are a handful of elementwise torch ops on (S,N) / (N,) tensors."""

import math

import numpy as np
import torch

from . import ops
from .event_generation_params import load_calibration
from .nerf import Softplus


def _get(cfg, key):
    return cfg[key] if isinstance(cfg, dict) else getattr(cfg, key)


class PixelBandwidth(torch.nn.Module):
    NS_TO_S = 1e-9
    PARAM_NAMES = ("tau_mil_it_eff_prod", "A_amp_inv", "A_loop_inv", "tau_out", "tau_sf",
                   "tau_diff")

    def __init__(self, dataset_directory, min_ts, f_c_dominant_min, target_cumprob):
        super().__init__()
        self.omega_c_dominant_min = 2 * math.pi * f_c_dominant_min
        min_ts = min_ts.detach().clone() if torch.is_tensor(min_ts) else torch.tensor(min_ts)
        self.register_buffer("min_ts", min_ts, persistent=False)
        self.register_buffer("target_cumprob_max_sample_lifetime",
                             torch.tensor(_get(target_cumprob, "max_sample_lifetime")),
                             persistent=False)
        calib = load_calibration(dataset_directory)
        c = {k: torch.from_numpy(np.asarray(calib[k])) for k in (
            "input_time_const_eff_it_prod", "miller_time_const_eff_it_prod", "amplifier_gain",
            "closed_loop_gain", "output_time_const", "sf_cutoff_freq", "diff_amp_cutoff_freq")}
        self.register_buffer("tau_in_it_eff_prod", c["input_time_const_eff_it_prod"],
                             persistent=False)
        init = {
            "tau_mil_it_eff_prod": c["miller_time_const_eff_it_prod"],
            "A_amp_inv": 1 / c["amplifier_gain"],
            "A_loop_inv": c["closed_loop_gain"] / c["amplifier_gain"],
            "tau_out": c["output_time_const"],
            "tau_sf": 1 / (2 * math.pi * c["sf_cutoff_freq"]),
            "tau_diff": 1 / (2 * math.pi * c["diff_amp_cutoff_freq"]),
        }
        for name in self.PARAM_NAMES:
            setattr(self, name, torch.nn.parameter.Parameter(init[name].clone()))
            torch.nn.utils.parametrize.register_parametrization(self, name, Softplus(beta=1))
        self.reset_delta_log_it = None
        self.reset_ts = None

    @property
    def omega_c_diff(self):
        return 1 / self.tau_diff

    def coefficients(self):
        """(alpha0, alpha1, beta, omega_sf, omega_diff) in fp64, differentiable in the six
        parameters: a = 2 zeta omega_n = alpha0 + alpha1 I, b = omega_n^2 = beta I
        (linearized_sys_params, :181-194, with tau_in = P_in / I, tau_mil = P_mil / I)."""
        p_in = self.tau_in_it_eff_prod.double()
        p_mil = self.tau_mil_it_eff_prod.double()
        tau_out = self.tau_out.double()
        a_amp = 1 / self.A_amp_inv.double()
        a_loop = 1 / self.A_loop_inv.double()
        denom = (p_in + p_mil) * tau_out
        return torch.stack([
            (p_in + (a_amp + 1) * p_mil) / denom,
            1 / (p_in + p_mil),
            (a_loop + 1) / denom,
            1 / self.tau_sf.double(),
            1 / self.tau_diff.double(),
        ])

    @torch.no_grad()
    def sample_lifetimes(self, normalized_interval_gen):
        """:311-350 — (S-1, ...) f64 in [0,1] -> sample lifetimes (S, ...) in ns (stop-grad)."""
        gen = normalized_interval_gen
        S = gen.shape[0] + 1
        bnd = torch.linspace(1, 0, S, dtype=gen.dtype, device=gen.device).view(
            -1, *((1,) * (gen.dim() - 1)))
        gen = torch.lerp(bnd[:-1], bnd[1:], gen)
        mid = torch.lerp(gen[:-1], gen[1:], 0.5)
        ones = torch.ones_like(mid[:1])
        life = torch.cat((ones, mid, torch.zeros_like(ones)), dim=0)
        rate = self.NS_TO_S * self.omega_c_dominant_min
        return -torch.log1p(-(self.target_cumprob_max_sample_lifetime * life)) / rate

    def forward(self, normalized_interval_gen, output_ts, intensity_sampling_fn,
                reset_diff=False, lifetimes=None, coefficients=None):
        """`lifetimes` / `coefficients`: values of sample_lifetimes(normalized_interval_gen) /
        coefficients() computed by the caller once for several calls of the same step (they depend
        only on the interval generator and the parameters)."""
        if lifetimes is None:
            lifetimes = self.sample_lifetimes(normalized_interval_gen)
        sample_ts = output_ts - lifetimes
        sampled = intensity_sampling_fn(sample_ts.clamp(min=self.min_ts))
        intensity, aux = sampled[0], sampled[1:]
        sample_dt = sample_ts.detach().diff(dim=0).to(intensity.dtype)
        batch_shape = intensity.shape[1:]
        S = intensity.shape[0]
        out = ops.lpf(intensity.reshape(S, -1), sample_dt.reshape(S - 1, -1),
                      self.coefficients() if coefficients is None else coefficients,
                      2 if reset_diff else 1)
        if reset_diff:
            sf = out[:, 0].reshape(batch_shape)
            before = out[:, 1].reshape(batch_shape)
            self.reset_delta_log_it = before - sf
            self.reset_ts = output_ts
            return sf, aux
        before = out[:, 0].reshape(batch_shape)
        w_diff = self.omega_c_diff
        reset_dt = (output_ts - self.reset_ts).to(w_diff.dtype)
        after = before - self.reset_delta_log_it * torch.exp(-w_diff * (self.NS_TO_S * reset_dt))
        return after, aux
