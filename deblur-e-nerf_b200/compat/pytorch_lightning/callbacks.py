"""`pl.callbacks.ModelCheckpoint(**config.checkpoint)` (scripts/run.py:71-73,
configs/train/synthetic.yaml:175-181: dirpath, monitor, mode, save_top_k, save_weights_only,
every_n_epochs).  `monitor` is empty in every shipped config, so "top k" means the k latest."""

import os


class ModelCheckpoint:
    def __init__(self, dirpath=None, monitor=None, mode="min", save_top_k=1, save_weights_only=False,
                 every_n_epochs=1, filename=None, **_kw):
        self.dirpath, self.monitor, self.mode = dirpath, monitor, mode
        self.save_top_k, self.save_weights_only = save_top_k, save_weights_only
        self.every_n_epochs = every_n_epochs or 1
        self.filename = filename
        self.saved = []
        self.best_model_path = ""

    def resolve_dir(self, trainer):
        if self.dirpath:
            return self.dirpath
        base = trainer.log_dir or os.getcwd()
        return os.path.join(base, "checkpoints")

    def on_epoch_end(self, trainer):
        if (trainer.current_epoch + 1) % self.every_n_epochs or not trainer.is_global_zero:
            return
        folder = self.resolve_dir(trainer)
        os.makedirs(folder, exist_ok=True)
        name = f"epoch={trainer.current_epoch}-step={trainer.global_step - 1}.ckpt"
        path = os.path.join(folder, name)
        trainer.save_checkpoint(path, weights_only=self.save_weights_only)
        self.saved.append(path)
        self.best_model_path = path
        if self.save_top_k is not None and self.save_top_k >= 0:
            while len(self.saved) > max(self.save_top_k, 0):
                old = self.saved.pop(0)
                if old != path and os.path.exists(old):
                    os.remove(old)
