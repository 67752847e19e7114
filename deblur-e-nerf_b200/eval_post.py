"""Evaluation post-processing on the device: what
``DeblurENeRF.evaluation_epoch_end`` (models/deblur_e_nerf.py:705-969) does on the host after moving
every image there — the gain-exposure normalisation (:707-712), the float64 affine least squares in log
space that removes the affine ambiguity of the predicted log-intensity (:733-797), the Levenberg-
Marquardt refinement of the offset-gamma correction when ``correction.black_level_offset`` is set
(:846-937 with models/offset_gamma_correction.py and external/optimizer.py:60-111) and the L1 / PSNR
terms of ``Metric.compute`` (loss_metric/metric.py:57-72).

Each of them is a sum over pixels: one ``den_eval_*`` kernel pass with fp64 accumulation produces the
moments, the 2x2 / 3x3 systems are solved on the device, and the images never leave HBM.  The SSIM term
(loss_metric/metric.py:78-81, torchmetrics 0.6.2 `functional.ssim`) is one windowed pass (`den_eval_ssim`);
LPIPS needs the pretrained lpips network (absent from this image) and is not part of this module.  Colour (Bayer) images
are C = 3 channels; ``per_channel_log_it_scale: false`` (:753-766: one log-intensity scale shared by the
channels, an offset per channel) is solved in closed form from the same per-channel moments."""

import ctypes

import torch

from . import ops

# pypose.optim.LevenbergMarquardt / strategy.TrustRegion defaults as the reference instantiates them
# (models/deblur_e_nerf.py:859-866, configs/*: `lm.radius: 1.0e+6`)
LM_DEFAULTS = dict(min=1e-6, max=1e32, high=0.5, low=1e-3, up=2.0, down=0.5, factor=3.0, reject=16)


def _moments(kind, pred, target, gain_vec, params, n_out):
    B, C, H, W = pred.shape
    out = torch.zeros((C, n_out), dtype=torch.float64, device=pred.device)
    args = [ops._ptr(pred), ops._ptr(target), ops._ptr(gain_vec)]
    if params is not None:
        args.append(ops._ptr(params))
    ops._call(kind, *args, B, C, H * W, ops._ptr(out), ops._stream())
    return out


def affine_log_fit(pred, target, norm_gain, per_channel_scale=True):
    """(C, 2) float64 (scale, offset) of the least squares  scale * log pred + offset ~ log target - log g.
    `per_channel_scale` False: the scale is shared by the channels (models/deblur_e_nerf.py:753-766); with
    b_c = (Sy_c - a Sx_c) / n_c eliminated, a = sum_c (Sxy_c - Sx_c Sy_c / n_c) / sum_c (Sxx_c - Sx_c^2 / n_c)."""
    log_gain = norm_gain.log().double().contiguous()      # the reference takes this log in fp32 (:745)
    m = _moments("den_eval_affine_moments", pred, target, log_gain, None, 5)
    n, sx, sy, sxx, sxy = m.unbind(-1)
    if per_channel_scale or m.shape[0] == 1:
        det = n * sxx - sx * sx
        scale = (n * sxy - sx * sy) / det
    else:
        scale = ((sxy - sx * sy / n).sum() / (sxx - sx * sx / n).sum()).expand_as(n)
    offset = (sy - scale * sx) / n
    return torch.stack((scale, offset), dim=-1)


def _apply(pred, target, gain, params):
    B, C, H, W = pred.shape
    out = torch.empty_like(pred)
    sums = torch.zeros((B, 2), dtype=torch.float64, device=pred.device)
    ops._call("den_eval_apply", ops._ptr(pred), ops._ptr(target), ops._ptr(gain), ops._ptr(params), B, C,
              H * W, ops._ptr(out), ops._ptr(sums), ops._stream())
    return out, sums


SSIM_DEFAULTS = dict(kernel_size=11, sigma=1.5, k1=0.01, k2=0.03)      # torchmetrics 0.6.2 functional.ssim


def ssim(pred, target, data_range, kernel_size=11, sigma=1.5, k1=0.01, k2=0.03):
    """Per-image SSIM (B,) f64 of (B, C, H, W) fp32 CUDA images: the Gaussian-windowed index averaged over
    the channels and over the pixels whose window lies inside the image (what torchmetrics 0.6.2 keeps
    after cropping its reflect-padded border).  `Metric.compute` passes data_range = max_target_val and
    averages the per-image values (models/deblur_e_nerf.py:957-969)."""
    if not pred.is_cuda:
        raise NotImplementedError("eval_post: only CUDA tensors are supported (no CPU fallback)")
    pred = ops._req(pred, torch.float32, "pred")
    target = ops._req(target, torch.float32, "target")
    B, C, H, W = target.shape
    sums = torch.zeros((B,), dtype=torch.float64, device=pred.device)
    ops._call("den_eval_ssim", ops._ptr(pred), ops._ptr(target), B, C, H, W, kernel_size, float(sigma),
              float((k1 * data_range) ** 2), float((k2 * data_range) ** 2), ops._ptr(sums), ops._stream())
    return sums / float(C * (H - kernel_size + 1) * (W - kernel_size + 1))


def _joint_layout(C, shared_gamma, device):
    """Positions of (scale_c, gamma_c, offset_c) in the joint parameter vector [scale (C), gamma (C or 1),
    offset (C)] — the order pypose flattens OffsetGammaCorrection's parameters in."""
    n_gamma = 1 if shared_gamma else C
    c = torch.arange(C, device=device)
    return torch.stack((c, C + (torch.zeros_like(c) if shared_gamma else c), C + n_gamma + c), dim=-1), 2 * C + n_gamma


def _joint_system(m, layout, P):
    """J^T J (P, P) and J^T r (P) of the joint problem from the per-channel moments (C, 10): channel c's
    rows of the Jacobian touch only (scale_c, gamma_c or the shared gamma, offset_c), so its 3 x 3 block is
    ADDED at those positions (a shared gamma couples the channels through its row and column)."""
    tri = torch.tensor([[0, 1, 2], [1, 3, 4], [2, 4, 5]], device=m.device)
    blocks = m[:, :6][:, tri]                                            # (C, 3, 3)
    A = torch.zeros((P, P), dtype=torch.float64, device=m.device)
    g = torch.zeros((P,), dtype=torch.float64, device=m.device)
    rows = layout.unsqueeze(-1).expand(-1, 3, 3).reshape(-1)
    cols = layout.unsqueeze(1).expand(-1, 3, 3).reshape(-1)
    A.index_put_((rows, cols), blocks.reshape(-1), accumulate=True)
    g.index_put_((layout.reshape(-1),), m[:, 6:9].reshape(-1), accumulate=True)
    return A, g


def lm_refine(pred, target, gain, affine, init, max_steps=10, radius=1e6):
    """Levenberg-Marquardt on the offset-gamma correction, external/optimizer.py:60-111 step for step on
    the JOINT parameter vector (scale per channel, gamma per channel or — `init` gamma of one element: a
    colour sensor with `per_channel_log_it_scale: false`, models/deblur_e_nerf.py:185-197 — ONE gamma
    shared by the channels, offset per channel); every J^T J / J^T r / loss evaluation is one
    den_eval_lm_moments pass.  Returns ((C, 3) f64 (scale, gamma, offset) per channel, errors)."""
    C = pred.shape[1]
    cfg = LM_DEFAULTS
    dev = pred.device
    scale0, gamma0, offset0 = (torch.as_tensor(v, dtype=torch.float64).reshape(-1).to(dev) for v in init)
    shared_gamma = C > 1 and gamma0.numel() == 1
    layout, P = _joint_layout(C, shared_gamma, dev)
    theta = torch.cat((scale0, gamma0, offset0))                          # (P,)
    assert theta.numel() == P, "init: scale and offset need C elements, gamma C or 1"

    def params_of(theta):
        return torch.cat((affine, theta[layout]), dim=-1).contiguous()    # (C, 5): a, b, s, gamma, o

    def moments(theta):
        return _moments("den_eval_lm_moments", pred, target, gain, params_of(theta), 10)

    damping, down = 1.0 / radius, cfg["down"]
    n = float(target.numel())
    m = moments(theta)
    loss = float(m[:, 9].sum())
    errors = [loss / n]
    for _ in range(max_steps):
        prev = theta.clone()
        last = loss
        A_und, g = _joint_system(m, layout, P)
        A = A_und.clone()
        A.diagonal().clamp_(cfg["min"], cfg["max"])
        rejects = 0
        while last <= loss:
            d = A.diagonal()
            d.add_(d * damping)
            D = torch.linalg.solve(A, -g.unsqueeze(-1))[:, 0]             # (P,)
            trial = theta + D
            m_trial = moments(trial)
            loss = float(m_trial[:, 9].sum())
            # predicted decrease -(J D)^T (2 R + J D) from the moments of the CURRENT point
            pred_dec = -float(D @ A_und @ D + 2 * (D * g).sum())
            quality = (last - loss) / pred_dec if pred_dec != 0 else 0.0
            rad = 1.0 / damping
            if quality > cfg["high"]:
                rad, down = rad * cfg["up"], cfg["down"]
            elif quality > cfg["low"]:
                rad, down = rad * (1 - (2 * quality - 1) ** cfg["factor"]), cfg["down"]
            else:
                rad, down = rad * down, down * cfg["down"]
            down = max(cfg["min"], min(down, cfg["max"]))
            rad = max(cfg["min"], min(rad, cfg["max"]))
            damping = 1.0 / rad
            if last < loss and rejects < cfg["reject"]:
                loss, rejects = last, rejects + 1                       # rejected: stay, more damping
            else:
                theta, m = trial, m_trial
                break
        errors.append(loss / n)
        if abs(errors[-1] - errors[-2]) <= 1e-8 + 1e-5 * abs(errors[-2]) and torch.allclose(theta, prev):
            break
    return theta[layout].clone(), errors


@torch.no_grad()
def evaluate(pred, target, exposure_time, gain, min_val, max_val, black_level_offset=True, init=None,
             max_steps=10, radius=1e6, per_channel_scale=True):
    """pred, target (B, C, H, W) or (B, H, W) fp32 CUDA tensors (C = 1: mono, 3: colour); exposure_time,
    gain (B,); `per_channel_scale`: `correction.per_channel_log_it_scale` (only matters for C = 3).
    Returns dict(l1, psnr, ssim (device scalars; ssim None for images smaller than its 11 x 11 window),
    pred (corrected, fp32), affine (C, 2), correction (C, 3) | None, correction_errors)."""
    if not pred.is_cuda:
        raise NotImplementedError("eval_post: only CUDA tensors are supported (no CPU fallback)")
    if pred.dim() == 3:
        pred, target = pred.unsqueeze(1), target.unsqueeze(1)
    pred = ops._req(pred, torch.float32, "pred")
    target = ops._req(target, torch.float32, "target")
    B, C, H, W = target.shape
    prod = gain.to(torch.float32) * exposure_time                           # :707
    norm = (prod / prod.mean()).to(pred.device)                             # :709-712
    gain64 = norm.double().contiguous()
    affine = affine_log_fit(pred, target, norm, per_channel_scale)
    errors = None
    if black_level_offset:
        if init is None:
            # models/deblur_e_nerf.py:174-197: unit scale / gamma, zero offset; ONE gamma when the
            # log-intensity scale is shared by the channels of a colour sensor
            init = (torch.ones(C), torch.ones(C if per_channel_scale or C == 1 else 1), torch.zeros(C))
        corr, errors = lm_refine(pred, target, gain64, affine, init, max_steps, radius)
    else:
        corr = torch.tensor([[1.0, 1.0, 0.0]], dtype=torch.float64, device=pred.device).repeat(C, 1)
    params = torch.cat((affine, corr), dim=-1).contiguous()
    out, sums = _apply(pred, target, gain64, params)
    n_img = float(C * H * W)
    l1 = (sums[:, 0] / n_img).mean()
    psnr = (10 * torch.log10((max_val - min_val) ** 2 / (sums[:, 1] / n_img))).mean()
    fits = min(H, W) >= SSIM_DEFAULTS["kernel_size"]
    ssim_mean = ssim(out, target, float(max_val), **SSIM_DEFAULTS).mean() if fits else None
    return {"l1": l1, "psnr": psnr, "ssim": ssim_mean, "pred": out, "affine": affine,
            "correction": corr if black_level_offset else None, "correction_errors": errors}
