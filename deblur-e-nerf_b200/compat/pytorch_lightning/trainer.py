"""`pl.Trainer` as `scripts/run.py:91-118` constructs and drives it: `fit`, `validate`, `test` with the
arguments of `configs/*/*.yaml` `trainer:` plus the ones run.py passes itself.

Loop semantics (Lightning 1.4.9):
* `global_step` counts OPTIMIZER steps; with `accumulate_grad_batches = k` the loss of each micro-batch is
  divided by k, the optimizer steps after every k-th batch (and at the end of an epoch);
* the next batch is fetched BEFORE the current `training_step` runs, so a batch size the step writes into
  the dataset (`update_train_batch_size`, models/deblur_e_nerf.py:1252-1308) reaches batch k + 2;
* a dict of dataloaders is zipped (`multiple_trainloader_mode="min_size"`) into a dict batch;
* `lr_scheduler.interval` "epoch" | "step"; `on_train_start`, `on_train_epoch_start` hooks; logged values
  are averaged over `log_every_n_steps` optimizer steps and handed to the logger;
* validation every `check_val_every_n_epoch` epochs (and `num_sanity_val_steps` batches before training)
  unless `limit_val_batches` is 0; `ModelCheckpoint` at epoch end; `resume_from_checkpoint`.
Devices: `gpus: [i, ...]` -> this process uses `cuda:gpus[LOCAL_RANK]`; several GPUs need torchrun
(one process per GPU); gradients are then mean all-reduced once per optimizer step."""

import os

import torch


def _to_device(obj, device):
    if torch.is_tensor(obj):
        return obj.to(device, non_blocking=True)
    if isinstance(obj, dict):
        return {k: _to_device(v, device) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_device(v, device) for v in obj)
    return obj


def _plain(obj):
    """Nested mappings / sequences as plain dict / list (scalars, strings, tensors and None unchanged)."""
    if isinstance(obj, dict):
        return {k: _plain(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [_plain(v) for v in obj]
    return obj


class Trainer:
    def __init__(self, callbacks=None, logger=True, plugins=None, replace_sampler_ddp=True,
                 sync_batchnorm=False, terminate_on_nan=False, multiple_trainloader_mode="max_size_cycle",
                 num_nodes=1, gpus=None, accelerator=None, max_epochs=1000, max_steps=None,
                 log_every_n_steps=50, check_val_every_n_epoch=1, flush_logs_every_n_steps=100,
                 val_check_interval=1.0, limit_train_batches=1.0, limit_val_batches=1.0,
                 limit_test_batches=1.0, num_sanity_val_steps=2, accumulate_grad_batches=None,
                 resume_from_checkpoint=None, checkpoint_callback=True, **unused):
        if multiple_trainloader_mode != "min_size":
            raise NotImplementedError("only multiple_trainloader_mode='min_size' (scripts/run.py:98)")
        self.callbacks = list(callbacks or [])
        self.logger = None if logger in (False, None, True) else logger
        self.plugins = plugins
        self.terminate_on_nan = terminate_on_nan
        self.num_nodes, self.gpus, self.accelerator = num_nodes, gpus, accelerator
        self.max_epochs, self.max_steps = max_epochs, max_steps
        self.log_every_n_steps = log_every_n_steps
        self.check_val_every_n_epoch = check_val_every_n_epoch
        self.limit_train_batches = limit_train_batches
        self.limit_val_batches, self.limit_test_batches = limit_val_batches, limit_test_batches
        self.num_sanity_val_steps = num_sanity_val_steps
        self.accumulate_grad_batches = int(accumulate_grad_batches or 1)
        self.resume_from_checkpoint = resume_from_checkpoint
        self.unused_arguments = dict(unused)
        self.global_step = 0
        self.current_epoch = 0
        self.sanity_checking = False
        self.datamodule = None
        self.model = None
        self.optimizers, self.lr_schedulers = [], []
        self._window = {}
        self.callback_metrics = {}
        from ... import ddp
        self._ddp = ddp
        self.global_rank, self.local_rank, self.world_size = ddp.init_from_env() \
            if (accelerator in ("ddp", "ddp_spawn", "ddp_cpu") or int(os.environ.get("WORLD_SIZE", "1")) > 1) \
            else (0, 0, 1)

    # ------------------------------------------------------------------ properties ----
    @property
    def is_global_zero(self):
        return self.global_rank == 0

    @property
    def log_dir(self):
        return self.logger.log_dir if self.logger is not None else None

    @property
    def device(self):
        if self.gpus and torch.cuda.is_available():
            ids = list(self.gpus) if isinstance(self.gpus, (list, tuple)) else list(range(int(self.gpus)))
            return torch.device("cuda", ids[self.local_rank % len(ids)])
        # `gpus: [0]` on a box without CUDA (the reference's constructor needs a list, models/
        # deblur_e_nerf.py:72): the loop itself is device agnostic; the den_b200 operators are not
        return torch.device("cpu")

    # ------------------------------------------------------------------- plumbing -----
    def _record(self, name, value, to_logger=True):
        value = value.detach() if torch.is_tensor(value) else value
        self.callback_metrics[name] = value
        if to_logger:
            self._window.setdefault(name, []).append(value)

    def _flush_logs(self):
        if self.logger is None or not self.is_global_zero or not self._window:
            self._window = {}
            return
        metrics = {}
        for name, values in self._window.items():
            nums = [float(v) for v in values]           # the only host reads of logged values
            metrics[name] = sum(nums) / len(nums)
        metrics["epoch"] = self.current_epoch
        self.logger.log_metrics(metrics, self.global_step)
        self._window = {}

    def _attach(self, model, datamodule):
        self.model, self.datamodule = model, datamodule
        object.__setattr__(model, "trainer", self)
        if datamodule is not None:
            datamodule.trainer = self
        model.to(self.device)
        if self.device.type == "cuda":
            torch.cuda.set_device(self.device)

    def _configure_optimizers(self, model):
        cfg = model.configure_optimizers()
        if isinstance(cfg, dict):
            self.optimizers = [cfg["optimizer"]]
            sched = cfg.get("lr_scheduler")
            if sched is not None:
                if not isinstance(sched, dict):
                    sched = {"scheduler": sched, "interval": "epoch"}
                self.lr_schedulers = [dict(sched)]
        elif isinstance(cfg, (list, tuple)):
            self.optimizers = list(cfg[0]) if isinstance(cfg[0], (list, tuple)) else [cfg[0]]
            rest = cfg[1] if len(cfg) > 1 else []
            self.lr_schedulers = [{"scheduler": s, "interval": "epoch"} for s in rest]
        else:
            self.optimizers = [cfg]

    def _step_schedulers(self, interval):
        for sched in self.lr_schedulers:
            if sched.get("interval", "epoch") == interval:
                sched["scheduler"].step()

    @staticmethod
    def _train_batches(loaders):
        """A dict of dataloaders -> dict batches, stopping with the shortest ("min_size")."""
        if isinstance(loaders, dict):
            iters = {k: iter(v) for k, v in loaders.items()}
            while True:
                batch = {}
                for k, it in iters.items():
                    try:
                        batch[k] = next(it)
                    except StopIteration:
                        return
                yield batch
        else:
            yield from loaders

    def _limit(self, limit, loader):
        if isinstance(limit, int) and not isinstance(limit, bool):
            return limit
        try:
            return int(len(loader) * float(limit))
        except TypeError:
            raise ValueError("a fractional batch limit needs a sized dataloader") from None

    # ----------------------------------------------------------------- checkpoints ----
    def save_checkpoint(self, path, weights_only=False):
        """Lightning 1.4.9's top-level keys.  The hyper-parameters are stored as plain containers (EasyDict
        -> dict, tuples -> lists): the file then loads with `torch.load`'s safe unpickler, which is the
        default the reference's own `torch.load(checkpoint_filepath)` (models/deblur_e_nerf.py:332-334)
        gets under torch >= 2.6."""
        ckpt = {"epoch": self.current_epoch + 1, "global_step": self.global_step,
                "pytorch-lightning_version": "1.4.9", "state_dict": self.model.state_dict(),
                "hyper_parameters": _plain(getattr(self.model, "hparams", {}))}
        if not weights_only:
            ckpt["optimizer_states"] = [o.state_dict() for o in self.optimizers]
            ckpt["lr_schedulers"] = [s["scheduler"].state_dict() for s in self.lr_schedulers]
        if hasattr(self.model, "on_save_checkpoint"):
            self.model.on_save_checkpoint(ckpt)
        tmp = path + ".tmp"
        torch.save(ckpt, tmp)
        os.replace(tmp, path)

    def _restore(self, path):
        ckpt = torch.load(path, map_location=self.device, weights_only=False)
        self.model.load_state_dict(ckpt["state_dict"])
        for opt, state in zip(self.optimizers, ckpt.get("optimizer_states", [])):
            opt.load_state_dict(state)
        for sched, state in zip(self.lr_schedulers, ckpt.get("lr_schedulers", [])):
            sched["scheduler"].load_state_dict(state)
        self.current_epoch = int(ckpt.get("epoch", 0))
        self.global_step = int(ckpt.get("global_step", 0))
        if hasattr(self.model, "on_load_checkpoint"):
            self.model.on_load_checkpoint(ckpt)

    # ------------------------------------------------------------------------ fit -----
    def fit(self, model, datamodule=None):
        self._attach(model, datamodule)
        ddp = self._ddp
        datamodule.setup("fit")
        self._configure_optimizers(model)
        if self.resume_from_checkpoint:
            self._restore(self.resume_from_checkpoint)
        ddp.broadcast_parameters(model)
        reducer = ddp.FlatGradAllReduce(model.parameters()) if self.world_size > 1 else None
        opt = self.optimizers[0]
        acc = self.accumulate_grad_batches

        if self.limit_val_batches != 0 and self.num_sanity_val_steps:
            self.sanity_checking = True
            self._evaluate("validation", datamodule.val_dataloader(), self.num_sanity_val_steps)
            self.sanity_checking = False

        model.train()
        model.on_train_start()
        done = False
        while self.current_epoch < self.max_epochs and not done:
            model.on_train_epoch_start()
            loaders = datamodule.train_dataloader()
            n_batches = self._limit(self.limit_train_batches,
                                    next(iter(loaders.values())) if isinstance(loaders, dict) else loaders)
            stream = self._train_batches(loaders)
            prefetched = next(stream, None)
            opt.zero_grad(set_to_none=True)
            batch_index = 0
            while prefetched is not None and batch_index < n_batches:
                batch = _to_device(prefetched, self.device)
                prefetched = next(stream, None)             # fetched BEFORE this step changes the batch size
                loss = model.training_step(batch, batch_index)
                if self.terminate_on_nan and (self.global_step + 1) % self.log_every_n_steps == 0 \
                        and not bool(torch.isfinite(loss.detach())):
                    raise ValueError(f"The loss returned in `training_step` is {float(loss)}.")
                (loss / acc if acc > 1 else loss).backward()
                batch_index += 1
                if batch_index % acc and batch_index < n_batches and prefetched is not None:
                    continue
                if reducer is not None:
                    reducer()
                opt.step()
                opt.zero_grad(set_to_none=True)
                self.global_step += 1
                self._step_schedulers("step")
                if self.global_step % self.log_every_n_steps == 0:
                    self._flush_logs()
                if self.max_steps is not None and self.global_step >= self.max_steps:
                    done = True
                    break
            self._step_schedulers("epoch")
            if self.limit_val_batches != 0 and (self.current_epoch + 1) % self.check_val_every_n_epoch == 0:
                self._evaluate("validation", datamodule.val_dataloader(), self.limit_val_batches)
                model.train()
            for cb in self.callbacks:
                if hasattr(cb, "on_epoch_end"):
                    cb.on_epoch_end(self)
            self.current_epoch += 1
        self._flush_logs()
        if self.logger is not None:
            self.logger.finalize("success")

    # ------------------------------------------------------------------- evaluate -----
    def _evaluate(self, kind, loader, limit):
        model = self.model
        step = getattr(model, f"{kind}_step" if kind != "test" else "test_step")
        end = getattr(model, f"{kind}_epoch_end" if kind != "test" else "test_epoch_end", None)
        n = limit if (isinstance(limit, int) and not isinstance(limit, bool)) else self._limit(limit, loader)
        model.eval()
        outputs = []
        with torch.no_grad():
            for i, batch in enumerate(loader):
                if i >= n:
                    break
                outputs.append(step(_to_device(batch, self.device), i))
            if end is not None:
                end(outputs)
        return [{k: (float(v) if torch.is_tensor(v) and v.numel() == 1 else v)
                 for k, v in self.callback_metrics.items() if k.startswith(("val", "test"))}]

    def validate(self, model, datamodule=None):
        self._attach(model, datamodule)
        datamodule.setup("validate")
        return self._evaluate("validation", datamodule.val_dataloader(), self.limit_val_batches)

    def test(self, model, datamodule=None):
        self._attach(model, datamodule)
        datamodule.setup("test")
        return self._evaluate("test", datamodule.test_dataloader(), self.limit_test_batches)
