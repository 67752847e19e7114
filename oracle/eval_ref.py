"""CPU restatement of the evaluation post-processing (TEST INFRASTRUCTURE — never imported by the
product path): ``DeblurENeRF.evaluation_epoch_end`` (models/deblur_e_nerf.py:705-969) up to the L1 /
PSNR terms of ``Metric.compute`` (loss_metric/metric.py:57-72).

Pinned parts: the gain-exposure normalisation, the float64 log-space affine least squares and the
metrics are checked against the reference's OWN method run under ``oracle/ref_shim``
(tests/test_oracle_vs_reference.py::test_eval_post_processing_matches_reference, correction without the
black-level offset).  The Levenberg-Marquardt refinement follows the reference's
``external/optimizer.py:60-111`` line by line, but the trust-region rule it calls
(``pypose.optim.strategy.TrustRegion.update``) lives in pypose, which is absent here: its constants are
RECALLED (radius 1e6, high 0.5, low 1e-3, up 2, down 0.5, factor 3, reject 16, min 1e-6, max 1e32) —
parity of the refinement is UNPINNED.  They only steer the damping (~1e-6 relative); the converged
parameters are the least-squares minimiser either way."""

import math

import torch


def normalized_gain(gain, exposure_time):
    prod = gain * exposure_time                                   # :707
    return prod / prod.mean()                                     # :709-712 (fp32 like the reference)


def affine_log_correction(pred, target, norm_gain, per_channel_scale=True):
    """:733-797.  pred, target (B,C,H,W) fp32; norm_gain (B,) fp32.  `per_channel_scale` False (a colour
    sensor with `correction.per_channel_log_it_scale: false`, :753-766,781-788): ONE scale shared by the C
    channels and an offset per channel — the reference's (C BHW, 1 + C) design matrix.
    Returns (scale_offset (C,2) f64, corrected log prediction (B,C,H,W) f64, log_gain (B,1,1,1) f32)."""
    B, C, H, W = target.shape
    log_gain = norm_gain.view(-1, 1, 1, 1).log()
    x = pred.log()
    y = target.log() - log_gain
    rhs = y.unsqueeze(-1).transpose(0, 1).flatten(1, 3).double()                               # (C, BHW, 1)
    if per_channel_scale or C == 1:
        A = torch.stack((x, torch.ones_like(x)), dim=-1).transpose(0, 1).flatten(1, 3).double()   # (C, BHW, 2)
        sol = torch.linalg.lstsq(A, rhs).solution                                              # (C, 2, 1)
        fitted = (A @ sol).view(C, B, H, W).transpose(0, 1)
        return sol[:, :, 0], fitted, log_gain
    A = torch.zeros((C, B * H * W, 1 + C), dtype=torch.float64)
    A[:, :, 0] = x.transpose(0, 1).flatten(1, 3).double()
    for c in range(C):
        A[c, :, 1 + c] = 1
    A = A.flatten(0, 1)                                                                        # (C BHW, 1 + C)
    sol = torch.linalg.lstsq(A, rhs.flatten(0, 1)).solution                                    # (1 + C, 1)
    fitted = (A @ sol).view(C, B, H, W).transpose(0, 1)
    scale_offset = torch.stack((sol[0, 0].expand(C), sol[1:, 0]), dim=-1)
    return scale_offset, fitted, log_gain


class _Correction:
    """models/offset_gamma_correction.py: f(x) = const_scale (scale x^gamma - offset); scale and offset hold
    C elements, gamma C or ONE shared by the channels (:60-65: `len(param) == 1 or len(param) == C`).  The
    parameters live in one vector [scale, gamma, offset], the order pypose flattens them in."""

    def __init__(self, const_scale, scale, gamma, offset):
        self.g = const_scale.double().view(-1, 1, 1, 1)
        self.C = scale.numel()
        self.n_gamma = gamma.numel()
        assert offset.numel() == self.C and self.n_gamma in (1, self.C)
        self.theta = torch.cat((scale.double().reshape(-1), gamma.double().reshape(-1), offset.double().reshape(-1)))

    def parts(self):
        C, G = self.C, self.n_gamma
        return (self.theta[:C].view(1, C, 1, 1), self.theta[C:C + G].view(1, G, 1, 1),
                self.theta[C + G:].view(1, C, 1, 1))

    @property
    def p(self):
        """(C, 3): scale, gamma (the shared one repeated), offset per channel."""
        s, gm, o = self.parts()
        return torch.stack((s.reshape(-1), gm.reshape(-1).expand(self.C), o.reshape(-1)), dim=-1)

    def forward(self, x):
        s, gm, o = self.parts()
        return self.g * (s * x.pow(gm) - o)

    def jacobian(self, x):
        """:124-190: the (B C H W, P) Jacobian — every output element depends on its channel's scale and
        offset and on its channel's (or the shared) gamma."""
        s, gm, _ = self.parts()
        B, C, H, W = x.shape
        js = self.g * x.pow(gm)
        jg = s * x.log() * js
        jo = (-self.g).expand_as(x)
        J = torch.zeros((B, C, H, W, self.theta.numel()), dtype=torch.float64)
        for c in range(C):
            J[:, c, :, :, c] = js[:, c]
            J[:, c, :, :, C + (c if self.n_gamma == C else 0)] = jg[:, c]
            J[:, c, :, :, C + self.n_gamma + c] = jo[:, c]
        return J.reshape(-1, self.theta.numel())


def lm_refine(x, target, norm_gain, init, max_steps=10, radius=1e6):
    """:846-895 with external/optimizer.py:60-111 (ONE joint problem over all parameters, like upstream:
    the damping, the acceptance test and the loss are shared by the channels).  x: affinely corrected
    prediction (B,C,H,W) f64; target (B,C,H,W) fp32; init = (scale (C), gamma (C or 1), offset (C)).
    Returns (params (C,3), errors list)."""
    model = _Correction(norm_gain, *init)
    t = target.double()
    pg = dict(min=1e-6, max=1e32, radius=radius, high=0.5, low=1e-3, up=2.0, down=0.5, damping=1.0 / radius)
    reject, factor, down0 = 16, 3.0, 0.5

    def loss():
        return float(((model.forward(x) - t) ** 2).sum())

    state = {"loss": None}

    def step():
        R = (model.forward(x) - t).reshape(-1, 1)                                              # (N, 1)
        J = model.jacobian(x)                                                                  # (N, P)
        last = cur = state["loss"] if state["loss"] is not None else loss()
        A = J.T @ J                                                                            # (P, P)
        g = J.T @ R
        A.diagonal().clamp_(pg["min"], pg["max"])
        rejects = 0
        while last <= cur:
            d = A.diagonal()
            d.add_(d * pg["damping"])
            D = torch.linalg.solve(A, -g)[:, 0]                                                # (P,)
            model.theta = model.theta + D
            cur = loss()
            JD = J @ D.unsqueeze(-1)
            quality = (last - cur) / -float((JD * (2 * R + JD)).sum())
            pg["radius"] = 1.0 / pg["damping"]
            if quality > pg["high"]:
                pg["radius"] *= pg["up"]
                pg["down"] = down0
            elif quality > pg["low"]:
                pg["radius"] *= 1 - (2 * quality - 1) ** factor
                pg["down"] = down0
            else:
                pg["radius"] *= pg["down"]
                pg["down"] *= down0
            pg["down"] = max(pg["min"], min(pg["down"], pg["max"]))
            pg["radius"] = max(pg["min"], min(pg["radius"], pg["max"]))
            pg["damping"] = 1.0 / pg["radius"]
            if last < cur and rejects < reject:
                model.theta = model.theta - D
                cur, rejects = last, rejects + 1
            else:
                break
        state["loss"] = cur
        return cur

    n = t.numel()
    errors = [loss() / n]
    for _ in range(max_steps):
        prev = model.theta.clone()
        errors.append(step() / n)
        if math.isclose(errors[-1], errors[-2], rel_tol=1e-5, abs_tol=1e-8) and torch.allclose(model.theta, prev):
            break
    return model.p.clone(), errors


def ssim(pred, target, data_range, kernel_size=11, sigma=1.5, k1=0.01, k2=0.03):
    """torchmetrics 0.6.2 ``functional.ssim(preds, target, data_range=...)`` with its defaults (the call of
    loss_metric/metric.py:78-81; environment.yml pins torchmetrics 0.6.2), restated from the published
    ``functional/image/ssim.py::_ssim_compute``: Gaussian window ``exp(-(d / sigma)^2 / 2)`` over
    ``d = arange((1 - k) / 2, (1 + k) / 2)``, normalised, 2-D window = outer product expanded per channel;
    reflect padding by (k - 1) / 2; ONE grouped conv2d over cat(p, t, pp, tt, pt); the index
    ((2 mu_pt + c1)(2 sigma_pt + c2)) / ((mu_p^2 + mu_t^2 + c1)(sigma_p^2 + sigma_t^2 + c2)); a CROP of
    (k - 1) / 2 pixels per side (so the padded border never reaches the result); ``elementwise_mean``.
    PARITY UNPINNED: torchmetrics is absent from this image; tests/test_oracle_kat.py checks this function
    against closed forms (identical images, constant images) and an independent float64 numpy window sum.
    pred, target (B, C, H, W) of one dtype; returns a 0-d tensor."""
    assert pred.dtype == target.dtype and pred.dim() == 4 and pred.shape == target.shape
    assert kernel_size % 2 == 1 and kernel_size > 0 and sigma > 0
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    B, C = pred.shape[:2]
    dist = torch.arange((1 - kernel_size) / 2, (1 + kernel_size) / 2, 1, dtype=pred.dtype)
    gauss = torch.exp(-torch.pow(dist / sigma, 2) / 2)
    gauss = (gauss / gauss.sum()).unsqueeze(0)                                  # (1, k)
    window = torch.matmul(gauss.t(), gauss).expand(C, 1, kernel_size, kernel_size)
    pad = (kernel_size - 1) // 2
    p = torch.nn.functional.pad(pred, (pad, pad, pad, pad), mode="reflect")
    t = torch.nn.functional.pad(target, (pad, pad, pad, pad), mode="reflect")
    out = torch.nn.functional.conv2d(torch.cat((p, t, p * p, t * t, p * t)), window, groups=C).split(B)
    mu_pp, mu_tt, mu_pt = out[0].pow(2), out[1].pow(2), out[0] * out[1]
    s_pp, s_tt, s_pt = out[2] - mu_pp, out[3] - mu_tt, out[4] - mu_pt
    idx = ((2 * mu_pt + c1) * (2 * s_pt + c2)) / ((mu_pp + mu_tt + c1) * (s_pp + s_tt + c2))
    return idx[..., pad:-pad, pad:-pad].mean()


def metrics(pred, target, min_val, max_val):
    """Mean over the images of Metric.compute's L1, PSNR and SSIM (loss_metric/metric.py:57-81; the SSIM
    term only for images larger than its 11 x 11 window)."""
    pred = pred.to(target.dtype)
    l1, psnr, ss = [], [], []
    for p, t in zip(pred, target):
        l1.append(torch.nn.functional.l1_loss(p, t))
        mse = ((p - t) ** 2).mean()
        psnr.append(10 * torch.log10((max_val - min_val) ** 2 / mse))
        if min(p.shape[-2:]) > 10:
            ss.append(ssim(p[None], t[None], max_val))
    res = {"l1": float(sum(l1) / len(l1)), "psnr": float(sum(psnr) / len(psnr))}
    if ss:
        res["ssim"] = float(sum(ss) / len(ss))
    return res


def evaluate(pred, target, exposure_time, gain, min_val, max_val, black_level_offset=True, init=None,
             max_steps=10, per_channel_scale=True):
    """pred, target (B,C,H,W) fp32 (C = 1: mono, C = 3: a Bayer sensor's colour images).  Returns dict(l1,
    psnr, pred (B,C,H,W) fp32, affine (C,2), correction (C,3) | None)."""
    norm = normalized_gain(gain, exposure_time)
    sol, fitted, log_gain = affine_log_correction(pred, target, norm, per_channel_scale)
    C = target.shape[1]
    if not black_level_offset:
        out = (fitted + log_gain).exp()                                                        # :822-829
        params = None
    else:
        x = fitted.exp()
        if init is None:                                                                       # :174-197
            init = (torch.ones(C), torch.ones(C if per_channel_scale or C == 1 else 1), torch.zeros(C))
        params, _ = lm_refine(x, target, norm, init, max_steps)
        out = _Correction(norm, params[:, 0], params[:, 1], params[:, 2]).forward(x)
    out = out.to(target.dtype)
    res = metrics(out, target, min_val, max_val)
    res.update(pred=out, affine=sol, correction=params)
    return res
