"""From the reference's YAML + dataset directory to a running job, without Lightning: what
``scripts/run.py:22-120`` and ``DeblurENeRF.__init__`` / ``configure_optimizers`` (models/deblur_e_nerf.py:
31-392, 1055-1112) do with a config of ``configs/train/*.yaml`` / ``configs/test/*.yaml``.

    cfg = config.load("configs/train/synthetic.yaml")
    model = config.build_model(cfg, device="cuda")                  # EventRenderer (B2 classes)
    optimizer, scheduler = config.build_optimizer(cfg, model)
    config.train(cfg, device="cuda")                                # the whole `run.py train`
    config.test(cfg, device="cuda")                                 # the whole `run.py test`

Same YAML keys, same meaning: ``model.nerf.aabb`` / ``render_step_size`` may be ``auto`` (:258-279), the
per-component ``load_state_dict`` / ``freeze`` switches with their per-parameter overrides (:321-383), the sample
budget split over the ranks (:72-75), the Bayer pattern of the calibration selecting three radiance channels
(:82-90, 281-284), the optimizer's parameter groups and the MultiStepLR schedule (:1055-1112), the event
producer's seed offset per rank (data/datamodule.py:87-91).  Keys of the YAML that configure Lightning itself
(logger, checkpoint callback, ``trainer.gpus`` / ``accelerator``: ranks come from torchrun here) are ignored."""

import math
import os

import numpy as np
import torch
import yaml

from . import data, ddp, events, factory, trainer, views
from . import event_generation_params as egp
from . import loss as loss_mod
from . import nerf as nerf_mod
from . import renderer as renderer_mod
from . import trajectories

NUM_DIM = 3
MAX_NUM_SAMPLES_PER_RAY = 1024                          # models/deblur_e_nerf.py:26
MODEL_COMPONENTS = ("contrast_threshold", "refractory_period", "pixel_bandwidth", "nerf")
MULTI_PARAM_MODEL_COMPONENTS = ("contrast_threshold", "pixel_bandwidth")


def load(path):
    """The YAML as nested dicts (`yaml.full_load`, scripts/run.py:23-24)."""
    with open(path) as fh:
        return yaml.full_load(fh)


def camera_poses(dataset_directory):
    """``CameraPose.load_camera_poses`` (data/datasets.py:738-751): (position (C, 3), orientation quaternion
    (C, 4) xyzw, timestamp (C) int64 ns)."""
    poses = np.load(os.path.join(dataset_directory, "camera_poses.npz"))
    if set(poses.keys()) != {"T_wc_position", "T_wc_orientation", "T_wc_timestamp"}:
        raise KeyError("camera_poses.npz must hold T_wc_position, T_wc_orientation, T_wc_timestamp")
    return tuple(torch.tensor(poses[k]) for k in ("T_wc_position", "T_wc_orientation", "T_wc_timestamp"))


def _component_frozen(spec):
    return spec["default"] if isinstance(spec, dict) else bool(spec)


def _subset_length(ratio, eff_batch_size, total):
    """data/datamodule.py:127-142: an int ratio counts effective batches, a float is a fraction of the dataset."""
    if isinstance(ratio, int) and not isinstance(ratio, bool):
        length = ratio * eff_batch_size
        assert length <= total, "dataset ratio asks for more samples than the dataset holds"
        return length
    return int(ratio * total)


def build_model(cfg, device="cuda", world_size=None, trust_checkpoint=False):
    """``DeblurENeRF.__init__`` (:31-242) + ``_load_model_component_state_dicts`` + ``_freeze_model_components``:
    the EventRenderer with its components built from `cfg["model"]`, `cfg["loss"]` and the dataset directory.
    `model.checkpoint_filepath` is read with the safe unpickler unless `trust_checkpoint` (a Lightning
    checkpoint pickles its hyper-parameter objects)."""
    m, root = cfg["model"], cfg["data"]["dataset_directory"]
    world_size = ddp.world_size() if world_size is None else world_size
    for name in MODEL_COMPONENTS:
        assert isinstance(m[name]["load_state_dict"], bool) and isinstance(m[name]["freeze"], (bool, dict)), name
    if m["nerf"]["freeze"]:
        assert m["nerf"]["load_state_dict"], "a frozen field must be loaded from a checkpoint (:66-69)"
    calib = egp.load_calibration(root)
    has_bayer_filter = str(calib["bayer_pattern"]) != ""
    poses = camera_poses(root)
    refractory_path = os.path.join(root, events.MAX_REFRACTORY_PERIOD_FILENAME)
    if not os.path.isfile(refractory_path):             # models/event_generation_params.py:135-149: extract & cache
        raw = np.load(os.path.join(root, events.RAW_EVENTS_FILENAME))
        events._save_atomically(events.extract_max_refractory_period(raw, calib, device), refractory_path)

    n = m["nerf"]
    aabb = n["aabb"]
    if aabb == "auto":                                                                              # :258-263
        aabb = torch.cat((poses[0].min(dim=0).values, poses[0].max(dim=0).values)).tolist()
    step = n["render_step_size"]
    if step == "auto":                                                                              # :271-277
        extent = torch.tensor(aabb[NUM_DIM:]) - torch.tensor(aabb[:NUM_DIM])
        step = math.sqrt(NUM_DIM) * torch.max(extent).item() / MAX_NUM_SAMPLES_PER_RAY
    field = nerf_mod.NeRF(aabb, factory.CONTRACTIONS[n["contraction_type"]], n["occ_grid"], n["near_plane"],
                          n["far_plane"], step, "parameter" if cfg["data"]["alpha_over_white_bg"] else None,
                          n["cone_angle"], n["early_stop_eps"], n["alpha_thre"], n["test_chunk_size"], n["arch"],
                          n[n["arch"]], NUM_DIM, 3 if has_bayer_filter else 1)
    pixel_bandwidth = None
    if m["pixel_bandwidth"]["enable"]:                                                              # :225-230
        from . import pixel_bandwidth as pb_mod
        pixel_bandwidth = pb_mod.PixelBandwidth(root, poses[2].min(), m["pixel_bandwidth"]["f_c_dominant_min"],
                                                m["pixel_bandwidth"]["target_cumprob"])
    model = renderer_mod.EventRenderer(
        field, trajectories.LinearTrajectory(poses),
        egp.ContrastThreshold(root, m["contrast_threshold"]["parameterize_mean_ct"]), egp.RefractoryPeriod(root),
        pixel_bandwidth, loss_mod.Loss(cfg["loss"]["weight"], cfg["loss"]["error_fn"], cfg["loss"]["normalize"]),
        torch.linalg.inv(torch.from_numpy(np.asarray(calib["intrinsics"]))).to(torch.get_default_dtype()),
        min_modeled_intensity=m["min_modeled_intensity"],
        train_ray_sample_batch_size=cfg["data"]["train_eff_ray_sample_batch_size"], world_size=world_size)
    model.has_bayer_filter = has_bayer_filter
    model.it_sample_size = m["pixel_bandwidth"]["it_sample_size"] if pixel_bandwidth is not None else None

    components = [c for c in MODEL_COMPONENTS if getattr(model, c, None) is not None]
    if any(m[c]["load_state_dict"] for c in components):                                            # :321-343
        ckpt = torch.load(m["checkpoint_filepath"], map_location="cpu", weights_only=not trust_checkpoint)
        for c in components:
            if m[c]["load_state_dict"]:
                prefix = c + "."
                getattr(model, c).load_state_dict({k[len(prefix):]: v for k, v in ckpt["state_dict"].items()
                                                   if k.startswith(prefix)})
    for c in components:                                                                            # :345-383
        if _component_frozen(m[c]["freeze"]):
            getattr(model, c).requires_grad_(False)
    for c in MULTI_PARAM_MODEL_COMPONENTS:
        if c not in components or isinstance(m[c]["freeze"], bool):
            continue
        for name, frozen in m[c]["freeze"].items():
            if name != "default":
                getattr(getattr(model, c).parametrizations, name).original.requires_grad_(not frozen)
    return model.to(device)


def build_optimizer(cfg, model, fused=None):
    """``configure_optimizers`` (:1055-1112): Adam over the reference's parameter groups (group for group,
    `factory.optimizer_param_groups`), MultiStepLR.  Returns (optimizer, scheduler)."""
    o, s = cfg["optimizer"], cfg["lr_scheduler"]
    if o["algo"] != "adam" or s["algo"] != "multi_step_lr":
        raise NotImplementedError("optimizer.algo must be adam and lr_scheduler.algo multi_step_lr (the "
                                  "reference implements nothing else)")
    component_lr = {c: dict(o["lr"][c]) for c in MULTI_PARAM_MODEL_COMPONENTS
                    if getattr(model, c, None) is not None and c in o["lr"]}
    optimizer = factory.configure_optimizer(
        model, lr=o["lr"]["default"], weight_decay=cfg["loss"]["weight"]["nerf_mlp_weight_decay"],
        refractory_relative_lr=o["relative_lr"]["refractory_period"], component_lr=component_lr, fused=fused)
    scheduler = torch.optim.lr_scheduler.MultiStepLR(optimizer, milestones=s["multi_step_lr"]["milestones"],
                                                     gamma=s["multi_step_lr"]["gamma"])
    return optimizer, scheduler


def build_producer(cfg, model, device="cuda", rank=None, world_size=None):
    """The training side of ``DataModule`` (data/datamodule.py:60-213): the transformed events of the dataset
    directory (built from the raw stream on the device when not cached), an optional fixed permutation
    (`train_dataset_perm_seed`), `train_dataset_ratio` as a fraction of the dataset or a number of effective
    batches, the per-rank share of `train_init_eff_batch_size`, generator seed `seed + rank`."""
    d = cfg["data"]
    rank = ddp.rank() if rank is None else rank
    world_size = ddp.world_size() if world_size is None else world_size
    ev = events.load_events(d["dataset_directory"], device)
    if d.get("train_dataset_perm_seed") is not None:                                                # data/datasets.py:56-66
        perm = torch.randperm(len(ev["position"]), generator=torch.Generator().manual_seed(d["train_dataset_perm_seed"]))
        ev = {k: v[perm.to(v.device)] for k, v in ev.items()}
    batch = d["train_init_eff_batch_size"] // world_size
    length = _subset_length(d.get("train_dataset_ratio", 1.0), d["train_init_eff_batch_size"], len(ev["position"]))
    return data.EventBatchProducer(ev, batch, it_sample_size=model.it_sample_size, device=device,
                                   seed=cfg["seed"] or 0, rank=rank, dataset_len=length)


def train(cfg, device="cuda", checkpoint_dir=None, log_fn=None, max_steps=None, validate=None,
          trust_checkpoint=False):
    """`run.py train`: seed, model, optimizer, scheduler, event producer, the optimizer-step loop, and — like
    Lightning between epochs — the validation views scored every `trainer.check_val_every_n_epoch` epochs
    (`validate`: default on unless `trainer.limit_val_batches` is 0).  Under torchrun every rank calls this
    (``ddp.init_from_env()`` first).  `trainer.resume_from_checkpoint` restores the model, the optimizer, the
    scheduler, the counters and the controller's batch size from a checkpoint of this loop.  Returns (model,
    trainer)."""
    t = cfg["trainer"]
    cfg["seed"] = trainer.seed_everything(cfg["seed"] if cfg.get("seed") is not None else 0)
    model = build_model(cfg, device, trust_checkpoint=trust_checkpoint)
    ddp.broadcast_parameters(model)
    ddp.attach(model)
    optimizer, scheduler = build_optimizer(cfg, model)
    producer = build_producer(cfg, model, device)
    loop = trainer.Trainer(max_epochs=t["max_epochs"], limit_train_batches=t["limit_train_batches"],
                           accumulate_grad_batches=t.get("accumulate_grad_batches", 1),
                           lr_scheduler_interval=cfg["lr_scheduler"]["interval"], checkpoint_dir=checkpoint_dir,
                           checkpoint_every_n_epochs=cfg.get("checkpoint", {}).get("every_n_epochs", 1),
                           log_every_n_steps=t.get("log_every_n_steps", 100), log_fn=log_fn, max_steps=max_steps)
    if t.get("resume_from_checkpoint"):                 # Lightning's key: model, optimizer, scheduler, counters
        loop.load_checkpoint(t["resume_from_checkpoint"], model, optimizer, scheduler,
                             trust_pickle=trust_checkpoint)
        if model.next_train_batch_size:
            producer.set_batch_size(model.next_train_batch_size)
    if validate is None:
        validate = t.get("limit_val_batches", 1.0) != 0
    validate_fn = (lambda m: test(cfg, device, stage="val", model=m, log_fn=log_fn)) if validate else None
    loop.fit(model, producer, optimizer, scheduler, validate_fn=validate_fn,
             check_val_every_n_epoch=t.get("check_val_every_n_epoch", 1))
    return model, loop


def test(cfg, device="cuda", stage="test", model=None, log_fn=None, predictions_dir=None):
    """`run.py test` / `val`: the posed images of the stage (`eval_target: [novel_view]` -> `test` / `val`
    views, `[event_view]` -> the train views, data/datamodule.py:107-118; `<stage>_dataset_ratio` trims them) rendered in eval mode and scored on the
    device with the `model.correction` options; with `model.eval_save_pred_intensity_img` and a
    `predictions_dir` the corrected predictions are written there as 8-bit PNGs.  Returns (metrics, corrected
    predictions)."""
    d, c = cfg["data"], cfg["model"]["correction"]
    if model is None:
        model = build_model(cfg, device)
    target = set(cfg["eval_target"])
    if target not in ({"event_view"}, {"novel_view"}):
        raise NotImplementedError(f"eval_target {cfg['eval_target']}")
    folder = "train" if target == {"event_view"} else stage                 # data/datamodule.py:107-118
    posed = views.PosedViews(d["dataset_directory"], folder, d.get("eval_dataset_perm_seed"),
                             d["alpha_over_white_bg"], device)
    keep = _subset_length(d.get(f"{stage}_dataset_ratio", 1.0), d.get(f"{stage}_eff_batch_size", 1), len(posed))
    posed_views = [posed[i] for i in range(keep)]
    if c["black_level_offset"] and c["optimizer"]["algo"] != "lm":
        raise NotImplementedError("correction.optimizer.algo: only the Levenberg-Marquardt refinement is built")
    loop = trainer.Trainer(log_fn=log_fn)
    run = loop.test if stage == "test" else loop.validate
    metrics, pred = run(model, posed_views, black_level_offset=c["black_level_offset"],
                        per_channel_log_it_scale=c["per_channel_log_it_scale"],
                        max_steps=c["optimizer"]["max_steps"], radius=c["optimizer"]["lm"]["radius"],
                        **posed.test_arguments(device))
    if predictions_dir and cfg["model"].get("eval_save_pred_intensity_img") and ddp.rank() == 0:    # :1008-1054
        views.save_predictions(pred, torch.stack([v["sample_id"] for v in posed_views]), predictions_dir,
                               posed.min_normalized_pixel_value, posed.max_normalized_pixel_value)
    return metrics, pred


def main(argv=None):
    """`python -m deblur_e_nerf_b200.config {train,val,test} cfg.yaml` — scripts/run.py's three stages (under
    torchrun: one process per GPU)."""
    import argparse
    import json
    ap = argparse.ArgumentParser(description="Deblur e-NeRF on den_b200, from a config of the reference")
    ap.add_argument("stage", choices=("train", "val", "test"))
    ap.add_argument("config")
    ap.add_argument("--checkpoint-dir", default=None)
    ap.add_argument("--trust-checkpoint", action="store_true", help="unpickle model.checkpoint_filepath fully")
    ap.add_argument("--log-dir", default=None, help="write the logged scalars as TensorBoard event files there")
    ap.add_argument("--predictions-dir", default=None,
                    help="val / test: where model.eval_save_pred_intensity_img writes the predicted images")
    ap.add_argument("--metrics-out", default=None, help="val / test: write the metrics as YAML (run.py's metrics.yaml)")
    args = ap.parse_args(argv)
    rank, local_rank, _ = ddp.init_from_env()
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    cfg = load(args.config)
    log = (lambda step, row: print(json.dumps({"step": step, **row}), flush=True)) if rank == 0 else None
    if args.log_dir and rank == 0:
        log = trainer.tensorboard_log_fn(args.log_dir, also=log)
    if args.stage == "train":
        train(cfg, device, checkpoint_dir=args.checkpoint_dir if rank == 0 else None, log_fn=log,
              trust_checkpoint=args.trust_checkpoint)
    else:
        model = build_model(cfg, device, trust_checkpoint=args.trust_checkpoint)
        ddp.broadcast_parameters(model)
        metrics, _ = test(cfg, device, stage=args.stage, model=model, log_fn=log,
                          predictions_dir=args.predictions_dir)
        if args.metrics_out and rank == 0:                      # scripts/run.py:122-133
            with open(args.metrics_out, "w") as fh:
                yaml.dump([metrics], fh)
    ddp.barrier()


if __name__ == "__main__":
    main()
