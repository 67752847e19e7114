"""`pl.loggers.tensorboard.TensorBoardLogger(default_hp_metric=False, **config.logger)`
(scripts/run.py:77-80): `log_dir = save_dir/name/version_N`, `experiment` is a
`torch.utils.tensorboard.SummaryWriter` (created on first use, rank 0)."""

import os

import yaml


class TensorBoardLogger:
    def __init__(self, save_dir, name="default", version=None, default_hp_metric=True, **_kw):
        self.save_dir, self.name = save_dir, name or ""
        self._version = version
        self._writer = None

    @property
    def root_dir(self):
        return os.path.join(self.save_dir, self.name)

    @property
    def version(self):
        if self._version is None:
            existing = []
            if os.path.isdir(self.root_dir):
                for entry in os.listdir(self.root_dir):
                    if entry.startswith("version_") and entry[8:].isdigit():
                        existing.append(int(entry[8:]))
            self._version = max(existing) + 1 if existing else 0
        return self._version

    @property
    def log_dir(self):
        version = self.version
        return os.path.join(self.root_dir, version if isinstance(version, str) else f"version_{version}")

    @property
    def experiment(self):
        if self._writer is None:
            from torch.utils.tensorboard import SummaryWriter
            os.makedirs(self.log_dir, exist_ok=True)
            self._writer = SummaryWriter(log_dir=self.log_dir)
        return self._writer

    def log_metrics(self, metrics, step):
        for key, value in metrics.items():
            try:
                self.experiment.add_scalar(key, float(value), step)
            except (TypeError, ValueError):
                pass

    def log_hyperparams(self, params, metrics=None):
        os.makedirs(self.log_dir, exist_ok=True)

        def plain(obj):
            if isinstance(obj, dict):
                return {str(k): plain(v) for k, v in obj.items()}
            if isinstance(obj, (list, tuple)):
                return [plain(v) for v in obj]
            if isinstance(obj, (int, float, str, bool)) or obj is None:
                return obj
            return str(obj)

        with open(os.path.join(self.log_dir, "hparams.yaml"), "w") as fh:
            yaml.safe_dump(plain(dict(params)), fh)

    def save(self):
        if self._writer is not None:
            self._writer.flush()

    def finalize(self, status="success"):
        if self._writer is not None:
            self._writer.flush()
            self._writer.close()
