"""A `pytorch_lightning` 1.4.9-shaped package holding exactly what the reference uses of it
(`scripts/run.py:10,32,70-100`, `models/deblur_e_nerf.py`, `data/datamodule.py`): `seed_everything`,
`LightningModule`, `LightningDataModule`, `Trainer` (fit / validate / test), `callbacks.ModelCheckpoint`,
`loggers.tensorboard.TensorBoardLogger`, `plugins.DDPPlugin` / `DDPSpawnPlugin`.

The training loop keeps the behaviours the hot path relies on: `global_step` counts
optimizer steps, `accumulate_grad_batches` divides the loss, the next batch is fetched BEFORE the current
step runs (a batch-size change of step k reaches batch k + 2), the scheduler steps per epoch or per
step, checkpoints carry Lightning's top-level keys and are named `epoch=E-step=S.ckpt`.  Multi-GPU runs
are one process per GPU under torchrun (deblur_e_nerf_b200.ddp), not Lightning's own launcher."""

import os
import random

import numpy as np
import torch

from . import callbacks, loggers, plugins
from .core import LightningDataModule, LightningModule
from .trainer import Trainer

__version__ = "1.4.9+den_b200"
SUBMODULES = ("callbacks", "loggers", "loggers.tensorboard", "plugins")


def seed_everything(seed=None, workers=False):
    """scripts/run.py:32 — seeds python / numpy / torch, returns the seed (drawn when it is None)."""
    if seed is None:
        seed = random.SystemRandom().randint(0, np.iinfo(np.uint32).max)
    seed = int(seed)
    os.environ["PL_GLOBAL_SEED"] = str(seed)
    os.environ["PL_SEED_WORKERS"] = str(int(bool(workers)))
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    return seed
