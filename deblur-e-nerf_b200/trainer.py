"""The caller of the hot path: the optimizer-step loop (minimal form).

The reference drives ``DeblurENeRF.training_step`` through pytorch-lightning 1.4.9's ``Trainer``
(scripts/run.py:70-100).  This module is the Lightning-free loop with the behaviours the hot path
relies on, nothing else of Lightning:

* epochs of ``limit_train_batches`` batches over an endless batch producer
  (configs/train/synthetic.yaml:189-195); ``max_epochs`` of them;
* ``global_step`` counts optimizer steps, ``batch_index`` counts micro-batches inside the epoch;
  with ``accumulate_grad_batches = k`` the loss of every micro-batch is divided by k, gradients
  accumulate over k micro-batches and are all-reduced (``ddp.FlatGradAllReduce``, the DDP plugin of
  scripts/run.py:84-89) only before the optimizer step — Lightning's ``no_sync`` behaviour;
* ONE batch is prefetched, so the batch size the controller picks in batch k
  (models/deblur_e_nerf.py:1277-1285) is first seen by batch k + 2;
* the learning-rate scheduler steps per epoch or per optimizer step
  (``lr_scheduler.interval``, models/deblur_e_nerf.py:1098-1112);
* checkpoints use Lightning's top-level keys (``state_dict``, ``optimizer_states``,
  ``lr_schedulers``, ``epoch``, ``global_step``) so that a reference checkpoint's tensors land
  where they belong, plus the batch controller's state, which the reference keeps on the
  datamodule.

``Trainer.test`` / ``Trainer.validate`` are the evaluation loop of `run.py test`: every
posed view rendered in eval mode, then the post-processing and the metrics on the device
(``EventRenderer.evaluation_step`` / ``evaluation_epoch_end``).

Callbacks and the CLI stay out of scope."""

import os
import random

import numpy as np
import torch

from . import ddp


def seed_everything(seed):
    """pytorch_lightning.seed_everything: python, numpy and torch (all devices), same on every rank."""
    seed = int(seed)
    random.seed(seed)
    np.random.seed(seed % (1 << 32))
    torch.manual_seed(seed)
    return seed


def tensorboard_log_fn(log_dir, also=None):
    """A `log_fn` that writes every logged scalar to TensorBoard event files in `log_dir` (what Lightning's
    TensorBoardLogger does with `self.log`, scripts/run.py:76-81); `also(step, row)` is called after it."""
    from torch.utils.tensorboard import SummaryWriter
    writer = SummaryWriter(log_dir=log_dir)

    def log(step, row):
        for name, value in row.items():
            if isinstance(value, (int, float)):
                writer.add_scalar(name, value, global_step=step)
        writer.flush()
        if also is not None:
            also(step, row)

    log.writer = writer
    return log


class Trainer:
    def __init__(self, max_epochs=40, limit_train_batches=1000, accumulate_grad_batches=1,
                 lr_scheduler_interval="epoch", checkpoint_dir=None, checkpoint_every_n_epochs=1,
                 log_every_n_steps=100, log_fn=None, max_steps=None):
        assert lr_scheduler_interval in ("epoch", "step")
        assert accumulate_grad_batches >= 1 and limit_train_batches >= 1
        self.max_epochs = int(max_epochs)
        self.limit_train_batches = int(limit_train_batches)
        self.accumulate_grad_batches = int(accumulate_grad_batches)
        self.lr_scheduler_interval = lr_scheduler_interval
        self.checkpoint_dir = checkpoint_dir
        self.checkpoint_every_n_epochs = int(checkpoint_every_n_epochs)
        self.log_every_n_steps = int(log_every_n_steps)
        self.log_fn = log_fn
        self.max_steps = max_steps
        self.current_epoch = 0
        self.global_step = 0
        self.history = []           # (global_step, {name: float}) every log_every_n_steps

    # ------------------------------------------------------------------ checkpoints ----
    def checkpoint(self, model, optimizer, scheduler=None):
        return {
            "epoch": self.current_epoch,
            "global_step": self.global_step,
            "state_dict": model.state_dict(),
            "optimizer_states": [optimizer.state_dict()],
            "lr_schedulers": [scheduler.state_dict()] if scheduler is not None else [],
            "train_batch_size": {"next": model.next_train_batch_size},
        }

    def save_checkpoint(self, path, model, optimizer, scheduler=None):
        if ddp.world_size() > 1 and torch.distributed.get_rank() != 0:
            return
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        tmp = path + ".tmp"
        torch.save(self.checkpoint(model, optimizer, scheduler), tmp)
        os.replace(tmp, path)           # a crash never leaves a truncated checkpoint behind

    def load_checkpoint(self, path, model, optimizer=None, scheduler=None, strict=True,
                        trust_pickle=False):
        """Checkpoints are loaded with `weights_only=True` (tensors and plain containers only): a
        public checkpoint cannot run code on load.  A Lightning checkpoint that pickles other objects
        (e.g. hyper-parameter containers) needs the explicit `trust_pickle=True`."""
        ckpt = torch.load(path, map_location=next(model.parameters()).device,
                          weights_only=not trust_pickle)
        model.load_state_dict(ckpt["state_dict"], strict=strict)
        if optimizer is not None and ckpt.get("optimizer_states"):
            optimizer.load_state_dict(ckpt["optimizer_states"][0])
        if scheduler is not None and ckpt.get("lr_schedulers"):
            scheduler.load_state_dict(ckpt["lr_schedulers"][0])
        self.current_epoch = int(ckpt.get("epoch", 0))
        self.global_step = int(ckpt.get("global_step", 0))
        nxt = (ckpt.get("train_batch_size") or {}).get("next")
        if nxt:
            model.next_train_batch_size = int(nxt)
        return ckpt

    # ------------------------------------------------------------------------ loop ----
    def fit(self, model, producer, optimizer, scheduler=None, reducer=None, validate_fn=None,
            check_val_every_n_epoch=1):
        """Run up to `max_epochs` epochs (or `max_steps` optimizer steps) from the trainer's current
        epoch / global step; returns the last logged dict.  `validate_fn(model)`, if given, runs after
        every `check_val_every_n_epoch`-th epoch (Lightning's validation loop between training epochs,
        configs/train/*.yaml `trainer.check_val_every_n_epoch`); the model is back in training mode after."""
        acc = self.accumulate_grad_batches
        model.accumulate_grad_batches = acc
        model.train()
        if reducer is None:
            if hasattr(optimizer, "grad_scale"):        # FusedAdam: copy-free reduction, mean folded in
                reducer = ddp.GradReducer(model)
                reducer.bind(optimizer)
            else:
                reducer = ddp.FlatGradAllReduce(model.parameters())
        keep_views = isinstance(reducer, ddp.GradReducer)
        if getattr(model, "next_train_batch_size", None):
            producer.set_batch_size(model.next_train_batch_size)
        logged = {}
        prefetched = producer.next_batch()
        done = False
        while self.current_epoch < self.max_epochs and not done:
            optimizer.zero_grad(set_to_none=not keep_views)
            for batch_index in range(self.limit_train_batches):
                batch = prefetched
                prefetched = producer.next_batch()      # drawn BEFORE this batch's controller update
                loss = model.training_step(batch, batch_index, self.global_step)
                (loss / acc if acc > 1 else loss).backward()
                if model.next_train_batch_size:
                    producer.set_batch_size(model.next_train_batch_size)
                last_of_window = (batch_index + 1) % acc == 0 or \
                    batch_index + 1 == self.limit_train_batches
                if not last_of_window:
                    continue
                reducer()
                optimizer.step()
                optimizer.zero_grad(set_to_none=not keep_views)
                self.global_step += 1
                if scheduler is not None and self.lr_scheduler_interval == "step":
                    scheduler.step()
                if self.global_step % self.log_every_n_steps == 0:
                    logged = self._log(model)
                if self.max_steps is not None and self.global_step >= self.max_steps:
                    done = True
                    break
            if done:
                break
            if validate_fn is not None and (self.current_epoch + 1) % check_val_every_n_epoch == 0:
                validate_fn(model)
                model.train()
            self.current_epoch += 1
            if scheduler is not None and self.lr_scheduler_interval == "epoch":
                scheduler.step()
            if self.checkpoint_dir and self.current_epoch % self.checkpoint_every_n_epochs == 0:
                self.save_checkpoint(os.path.join(self.checkpoint_dir, "last.ckpt"), model, optimizer,
                                     scheduler)
        return logged or self._log(model)

    # ------------------------------------------------------------------- evaluation ----
    def test(self, model, views, intrinsics_inv, min_normalized_pixel_value, max_normalized_pixel_value,
             img_pixel_pos=None, stage="test", **correction):
        """The evaluation loop Lightning runs for `run.py test` / `val` (test_step + test_epoch_end,
        models/deblur_e_nerf.py:595-672): `views` yields posed-image batches ({img, T_wc_position,
        T_wc_orientation[, exposure_time, gain, sample_id]}); `correction` = the keyword arguments of
        `EventRenderer.evaluation_epoch_end` (black_level_offset, per_channel_log_it_scale, ...).  Under
        torch.distributed every rank renders its share of the views (a DistributedSampler's round-robin
        split) and the images are all-gathered (:672); the post-processing then runs on every rank (the
        reference: rank 0 only) so that all of them return the same metrics.  Returns (metrics as floats, corrected predictions (B, C, H, W) on the device)."""
        was_training = model.training
        model.eval()
        try:
            outputs, pos = [], img_pixel_pos
            rank, world = ddp.rank(), ddp.world_size()
            for i, view in enumerate(views):
                if i % world != rank:
                    continue
                if pos is None:
                    h, w = view["img"].shape[-2:]
                    pos = model.image_pixel_positions(h, w, device=intrinsics_inv.device)
                outputs.append(model.evaluation_step(view, intrinsics_inv, pos))
            outputs = ddp.gather_view_outputs(outputs, device=intrinsics_inv.device)
            metrics, pred = model.evaluation_epoch_end(
                outputs, min_normalized_pixel_value, max_normalized_pixel_value, stage=stage, **correction)
        finally:
            model.train(was_training)
        row = {k: float(v) for k, v in metrics.items()}
        self.history.append((self.global_step, row))
        if self.log_fn is not None:
            self.log_fn(self.global_step, row)
        return row, pred

    def validate(self, model, views, intrinsics_inv, min_normalized_pixel_value, max_normalized_pixel_value,
                 img_pixel_pos=None, **correction):
        return self.test(model, views, intrinsics_inv, min_normalized_pixel_value,
                         max_normalized_pixel_value, img_pixel_pos, stage="val", **correction)

    def _log(self, model):
        # the only host read of logged values: every log_every_n_steps optimizer steps
        row = {k: (float(v) if torch.is_tensor(v) else v) for k, v in model.logged.items()}
        self.history.append((self.global_step, row))
        if self.log_fn is not None:
            self.log_fn(self.global_step, row)
        return row
