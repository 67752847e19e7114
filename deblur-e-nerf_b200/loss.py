"""B2 host mirror of ``Loss`` (loss_metric/loss.py:8-96): per-event Huber / L1 / MSE on the
normalised log-intensity difference + the TV term, masked means.  Masked means are written
as sum(where(mask, err, 0)) / sum(mask) so no host synchronisation is needed (the reference's
boolean indexing syncs twice)."""

import torch
import torch.nn.functional as F


def _get(cfg, key):
    return cfg[key] if isinstance(cfg, dict) else getattr(cfg, key)


class Loss(torch.nn.Module):
    LOSS_NAMES = ["log_intensity_diff", "log_intensity_tv"]

    def __init__(self, loss_weight, loss_error_fn, loss_normalize):
        super().__init__()
        assert set(self.LOSS_NAMES) <= set(loss_weight.keys())
        for value in loss_weight.values():
            assert isinstance(value, (int, float)) and value >= 0
        assert sum(loss_weight.values()) > 0
        for key in self.LOSS_NAMES:
            if _get(loss_error_fn, key) not in ("l1", "mse", "huber", "mape"):
                raise KeyError(_get(loss_error_fn, key))
        self.loss_weight = loss_weight
        self.error_fn = loss_error_fn
        self.normalize = loss_normalize

    @staticmethod
    def _error(kind, pred, target):
        if kind == "l1":
            return (pred - target).abs()
        if kind == "mse":
            return (pred - target) ** 2
        if kind == "huber":
            return F.huber_loss(pred, target, reduction="none", delta=1.0)
        eps = torch.finfo(torch.float64).eps
        return (pred - target).abs() / target.abs().clamp(min=eps)

    @staticmethod
    def _masked_mean(err, mask):
        # the reference indexes `err[is_valid]` (loss_metric/loss.py:80,94): a non-finite error at an
        # invalid event must not reach the loss or its gradient, so masked-out entries are replaced
        # (torch.where routes a zero gradient to them), not multiplied by zero
        kept = torch.where(mask, err, torch.zeros_like(err))
        return kept.sum() / mask.sum().to(err.dtype)

    def compute(self, batch_event, batch_diff=None, batch_subdiff=None,
                mean_contrast_threshold=None):
        out = {}
        grad = batch_event["log_intensity_diff"] / (batch_event["end_ts"] - batch_event["start_ts"])
        batch_event["log_intensity_grad"] = grad
        if _get(self.loss_weight, "log_intensity_diff") > 0:
            k = mean_contrast_threshold if _get(self.normalize, "log_intensity_diff") else 1
            pred = batch_diff["log_intensity_diff"]
            target = (batch_diff["ts_diff"] * grad / k).to(pred.dtype)
            err = self._error(_get(self.error_fn, "log_intensity_diff"), pred / k, target)
            out["log_intensity_diff"] = self._masked_mean(err, batch_diff["is_valid"])
        if _get(self.loss_weight, "log_intensity_tv") > 0:
            k = mean_contrast_threshold if _get(self.normalize, "log_intensity_tv") else 1
            pred = batch_subdiff["log_intensity_diff"]
            err = self._error(_get(self.error_fn, "log_intensity_tv"), pred / k,
                              torch.zeros_like(pred))
            out["log_intensity_tv"] = self._masked_mean(err, batch_subdiff["is_valid"])
        return out
