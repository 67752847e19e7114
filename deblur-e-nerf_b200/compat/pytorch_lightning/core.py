"""`LightningModule` / `LightningDataModule`: the attributes and hooks the reference's two classes
touch."""

import inspect

import torch


class _HyperParameters(dict):
    """`self.hparams`: attribute and item access, nested dicts as given (the reference passes EasyDicts)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as exc:
            raise AttributeError(name) from exc

    def __setattr__(self, name, value):
        self[name] = value


class _HyperParametersMixin:
    def save_hyperparameters(self, *names, **_ignored):
        """Collect the named constructor arguments (all of them when no name is given) from the caller's
        frame, like Lightning does."""
        frame = inspect.currentframe().f_back
        local_vars = frame.f_locals
        if not names:
            init = inspect.signature(type(self).__init__)
            names = [n for n in init.parameters if n != "self"]
        hp = getattr(self, "_hparams", None)
        if hp is None:
            hp = _HyperParameters()
            object.__setattr__(self, "_hparams", hp)
        for name in names:
            if name in local_vars:
                hp[name] = local_vars[name]

    @property
    def hparams(self):
        hp = getattr(self, "_hparams", None)
        if hp is None:
            hp = _HyperParameters()
            object.__setattr__(self, "_hparams", hp)
        return hp


class LightningModule(_HyperParametersMixin, torch.nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        object.__setattr__(self, "trainer", None)

    # ---- what the trainer provides ---------------------------------------------------------
    @property
    def logger(self):
        return self.trainer.logger if self.trainer is not None else None

    @property
    def global_step(self):
        return self.trainer.global_step if self.trainer is not None else 0

    @property
    def current_epoch(self):
        return self.trainer.current_epoch if self.trainer is not None else 0

    @property
    def device(self):
        for t in self.parameters():
            return t.device
        for t in self.buffers():
            return t.device
        return torch.device("cpu")

    def log(self, name, value, prog_bar=False, logger=True, rank_zero_only=False, **_kw):
        if self.trainer is not None:
            self.trainer._record(name, value, logger)

    def all_gather(self, data):
        """World size 1: every tensor gains a leading dimension of 1 (Lightning's contract); under
        torch.distributed the tensors are gathered over the ranks."""
        import torch.distributed as dist

        def gather(t):
            if not torch.is_tensor(t):
                return t
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                parts = [torch.empty_like(t) for _ in range(dist.get_world_size())]
                dist.all_gather(parts, t.contiguous())
                return torch.stack(parts)
            return t[None]

        def walk(obj):
            if isinstance(obj, dict):
                return type(obj)({k: walk(v) for k, v in obj.items()})
            if isinstance(obj, (list, tuple)):
                return type(obj)(walk(v) for v in obj)
            return gather(obj)

        return walk(data)

    # ---- hooks (no-ops unless the model overrides them) -------------------------------------------
    def on_train_start(self):
        pass

    def on_train_epoch_start(self):
        pass

    def configure_optimizers(self):
        raise NotImplementedError


class LightningDataModule(_HyperParametersMixin):
    def __init__(self, *args, **kwargs):
        super().__init__()
        self.trainer = None

    def setup(self, stage=None):
        pass
