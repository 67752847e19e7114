from . import tensorboard
from .tensorboard import TensorBoardLogger

__all__ = ["tensorboard", "TensorBoardLogger"]
