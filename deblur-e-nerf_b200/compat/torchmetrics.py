"""Stand-in for the two `torchmetrics.functional` calls of `loss_metric/metric.py:66-81` (torchmetrics
0.6.2 is pinned upstream and absent here), so that the reference's own validation / test loop runs under
the façade:

* `psnr(preds, target, data_range, reduction, dim)` by its definition;
* `ssim(preds, target, data_range, reduction)` with upstream's defaults (11 x 11 Gaussian window, sigma
  1.5, k1 0.01, k2 0.03): CUDA fp32 images go through the product's `den_eval_ssim` kernel
  (`eval_post.ssim`); anything else through the same definition written with `conv2d` — the windowed
  index averaged over the pixels whose window lies inside the image, which is what upstream keeps after
  cropping its reflect-padded border (`functional/image/ssim.py::_ssim_compute`)."""

import types

import torch


def _psnr(preds, target, data_range, reduction="elementwise_mean", dim=None):
    err = (preds - target) ** 2
    mse = err.mean(dim=dim) if dim is not None else err.mean()
    value = 10 * torch.log10(torch.as_tensor(data_range, dtype=mse.dtype) ** 2 / mse)
    return value.mean() if reduction == "elementwise_mean" else value


def _ssim(preds, target, kernel_size=(11, 11), sigma=(1.5, 1.5), reduction="elementwise_mean",
          data_range=None, k1=0.01, k2=0.03):
    if preds.dtype != target.dtype:
        raise TypeError("Expected `preds` and `target` to have the same data type.")
    if preds.shape != target.shape or preds.dim() != 4:
        raise ValueError("Expected `preds` and `target` to have the same BxCxHxW shape.")
    if kernel_size[0] != kernel_size[1] or sigma[0] != sigma[1]:
        raise NotImplementedError("the stand-in supports square windows (upstream's default)")
    if reduction != "elementwise_mean":
        raise NotImplementedError("the stand-in supports reduction='elementwise_mean' (what the reference asks for)")
    k, s = int(kernel_size[0]), float(sigma[0])
    if data_range is None:
        data_range = max(preds.max() - preds.min(), target.max() - target.min())
    data_range = float(data_range)
    if preds.is_cuda and preds.dtype == torch.float32:
        from .. import eval_post
        return eval_post.ssim(preds, target, data_range, kernel_size=k, sigma=s, k1=k1, k2=k2).mean().to(preds.dtype)
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    B, C = preds.shape[:2]
    dist = torch.arange((1 - k) / 2, (1 + k) / 2, 1, dtype=preds.dtype, device=preds.device)
    gauss = torch.exp(-((dist / s) ** 2) / 2)
    gauss = gauss / gauss.sum()
    window = torch.outer(gauss, gauss).expand(C, 1, k, k)
    stack = torch.cat((preds, target, preds * preds, target * target, preds * target))
    mu_p, mu_t, e_pp, e_tt, e_pt = torch.nn.functional.conv2d(stack, window, groups=C).split(B)   # "valid"
    upper = 2 * (e_pt - mu_p * mu_t) + c2
    lower = (e_pp - mu_p * mu_p) + (e_tt - mu_t * mu_t) + c2
    return (((2 * mu_p * mu_t + c1) * upper) / ((mu_p * mu_p + mu_t * mu_t + c1) * lower)).mean()


functional = types.SimpleNamespace(psnr=_psnr, ssim=_ssim)
