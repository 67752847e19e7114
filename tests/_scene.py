"""Shared builders for parity tests: the same small scene as (a) the reference's own
modules (under ``oracle/ref_shim``, build container only), (b) the oracle restatement
``oracle/path_ref`` and (c) the CUDA product, all loaded with identical parameters."""

import os
import tempfile

import numpy as np
import torch

from deblur_e_nerf_b200 import synthetic
from oracle import nerfacc_ref, path_ref

CONTRACTIONS = {"aabb": "AABB", "sphere": "UN_BOUNDED_SPHERE", "tanh": "UN_BOUNDED_TANH"}

LOSS_CFG = dict(
    weight=dict(log_intensity_diff=1.0, log_intensity_tv=1e-3, nerf_mlp_weight_decay=1e-6),
    error_fn=dict(log_intensity_diff="huber", log_intensity_tv="l1"),
    normalize=dict(log_intensity_diff=True, log_intensity_tv=True),
)


def scene_config(name="synthetic", occ_resolution=32, small=True):
    cfg = dict(synthetic.CONFIGS[name])
    cfg["occ_resolution"] = occ_resolution
    cfg["arch"] = synthetic.arch_config(small=small)
    cfg["step"] = synthetic.render_step_size(cfg["aabb"])
    cfg["occ_grid"] = dict(resolution=occ_resolution, occ_thre=1e-2, ema_decay=0.95,
                           warmup_steps=256, n=16)
    return cfg


def randomize_field_(nerf, seed=0, table_scale=0.5, density_bias=2.5):
    """Give a random-init field some structure (bigger table values, denser medium) so the
    parity runs exercise early termination and non-trivial gradients."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        field = nerf.radiance_field
        p = field.mlp_base[0].params
        p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * table_scale)
        field.mlp_base[1].output_layer.bias[0] += density_bias
    return nerf


def build_oracle_nerf(cfg, seed=0):
    torch.manual_seed(seed)
    ctype = nerfacc_ref.ContractionType[CONTRACTIONS[cfg["contraction"]]]
    nerf = path_ref.NeRF(cfg["aabb"], ctype, cfg["occ_grid"], cfg["near_plane"], cfg["far_plane"],
                         cfg["step"], cfg["render_bkgd"], cfg["cone_angle"], cfg["early_stop_eps"],
                         cfg["alpha_thre"], cfg["test_chunk_size"], cfg["arch"],
                         radiance_dim=cfg.get("radiance_dim", 1))
    return randomize_field_(nerf, seed)


def build_oracle_renderer(cfg, it_sample_size=8, pixel_bandwidth=True, seed=0, n_poses=200):
    nerf = build_oracle_nerf(cfg, seed)
    poses = synthetic.camera_poses(cfg, n_poses=n_poses)
    calib = synthetic.calibration()
    traj = path_ref.LinearTrajectory(*poses)
    ct = path_ref.ContrastThreshold(calib["pos_contrast_threshold"],
                                    calib["neg_contrast_threshold"])
    rp = path_ref.RefractoryPeriod(calib["refractory_period"],
                                   synthetic.MAX_REFRACTORY_PERIOD_NS)
    pb = None
    if pixel_bandwidth:
        pb = path_ref.PixelBandwidth(calib, poses[2].min(), 21, 0.95)
    loss_cfg = {k: dict(v) for k, v in LOSS_CFG.items()}
    loss_cfg["weight"]["log_intensity_tv"] = cfg["tv_weight"]
    loss = path_ref.EventLoss(**loss_cfg)
    kinv = torch.linalg.inv(torch.from_numpy(synthetic.intrinsics(cfg)))
    return path_ref.EventRenderer(nerf, traj, ct, rp, pb, loss, kinv), poses


def make_batch(cfg, poses, n_events, it_sample_size, seed, pixel_bandwidth=True):
    g = torch.Generator().manual_seed(seed)
    event = synthetic.event_batch(n_events, cfg, poses[2], g)
    normalized = synthetic.normalized_batch(n_events, it_sample_size, g, pixel_bandwidth)
    return event, normalized


# --------------------------------------------------------------------------- #
# the reference's own modules (build container only)
# --------------------------------------------------------------------------- #
def build_reference_renderer(cfg, it_sample_size=8, pixel_bandwidth=True, seed=0, n_poses=200,
                             freeze_refractory_period=True, accumulate_grad_batches=1):
    from oracle import ref_shim
    ref_shim.load_lightning_module()
    import easydict
    import nerfacc
    nerf_mod = ref_shim.load("models.nerf")
    traj_mod = ref_shim.load("models.trajectories")
    egp = ref_shim.load("models.event_generation_params")
    pb_mod = ref_shim.load("models.pixel_bandwidth")
    loss_mod = ref_shim.load("loss_metric.loss")
    datasets = ref_shim.load("data.datasets")

    poses = synthetic.camera_poses(cfg, n_poses=n_poses)
    tmp = tempfile.mkdtemp(prefix="den_ref_ds_")
    synthetic.write_dataset_dir(tmp, cfg, poses)

    torch.manual_seed(seed)
    ctype = nerfacc.ContractionType[CONTRACTIONS[cfg["contraction"]]]
    nerf = nerf_mod.NeRF(cfg["aabb"], ctype, easydict.EasyDict(cfg["occ_grid"]),
                         cfg["near_plane"], cfg["far_plane"], cfg["step"], cfg["render_bkgd"],
                         cfg["cone_angle"], cfg["early_stop_eps"], cfg["alpha_thre"],
                         cfg["test_chunk_size"], "ngp", easydict.EasyDict(cfg["arch"]), 3,
                         cfg.get("radiance_dim", 1))
    randomize_field_(nerf, seed)
    camera_poses = datasets.CameraPose(tmp, None)
    components = dict(
        nerf=nerf,
        trajectory=traj_mod.LinearTrajectory(camera_poses),
        contrast_threshold=egp.ContrastThreshold(tmp, True),
        refractory_period=egp.RefractoryPeriod(tmp),
        loss=None,
    )
    if pixel_bandwidth:
        components["pixel_bandwidth"] = pb_mod.PixelBandwidth(
            tmp, camera_poses.camera_poses.T_wc_timestamp.min(), 21,
            easydict.EasyDict(max_sample_lifetime=0.95))
    loss_cfg = {k: dict(v) for k, v in LOSS_CFG.items()}
    loss_cfg["weight"]["log_intensity_tv"] = cfg["tv_weight"]
    components["loss"] = loss_mod.Loss(easydict.EasyDict(loss_cfg["weight"]),
                                       easydict.EasyDict(loss_cfg["error_fn"]),
                                       easydict.EasyDict(loss_cfg["normalize"]))
    hparams = dict(
        min_modeled_intensity=0.001,
        pixel_bandwidth=dict(enable=pixel_bandwidth, it_sample_size=it_sample_size),
        loss=loss_cfg,
        refractory_period=dict(freeze=freeze_refractory_period),
        nerf=dict(),
    )
    module = ref_shim.make_reference_module(hparams, components)
    module.has_bayer_filter = cfg.get("radiance_dim", 1) == 3      # models/deblur_e_nerf.py:82-90
    module.trainer.accumulate_grad_batches = accumulate_grad_batches
    module.render_bkgd = cfg["render_bkgd"]
    module.register_buffer("train_intrinsics_inv",
                           torch.linalg.inv(torch.from_numpy(synthetic.intrinsics(cfg))),
                           persistent=False)
    type(module).MULTI_PARAM_MODEL_COMPONENTS = ["contrast_threshold"] + (
        ["pixel_bandwidth"] if pixel_bandwidth else [])
    return module, poses


def reference_batch(event, normalized):
    """Lightning hands batches with a leading loader dim of 1 (squeezed at :402-406)."""
    return {
        "event": {k: v.clone().unsqueeze(0) for k, v in event.items()},
        "normalized": {k: v.clone().unsqueeze(0) for k, v in normalized.items()},
    }


def copy_params(src, dst):
    """Copy parameters between a reference module tree and an oracle/product tree by
    matching state-dict keys (names are identical by design)."""
    sd = src.state_dict()
    own = dst.state_dict()
    missing = [k for k in own if k not in sd]
    assert not missing, f"keys missing in source: {missing}"
    dst.load_state_dict({k: sd[k] for k in own}, strict=True)


def flat_named_grads(module):
    return {n: p.grad.detach().clone() for n, p in module.named_parameters()
            if p.grad is not None}


# --------------------------------------------------------------------------- #
# golden fixtures
# --------------------------------------------------------------------------- #
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    data = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: data[k] for k in data.files}


def golden_section(golden, prefix, device=None):
    out = {}
    for key, value in golden.items():
        if key.startswith(prefix + "/"):
            t = torch.from_numpy(np.asarray(value))
            out[key[len(prefix) + 1:]] = t.to(device) if device is not None else t
    return out


def load_golden_state(module, golden, name, device=None):
    """Load `state/<name>/...` of a golden into `module` (buffers that were skipped when
    the fixture was written — grid_coords/grid_indices — keep their own values)."""
    sd = golden_section(golden, f"state/{name}", device)
    own = module.state_dict()
    merged = {k: (sd[k].to(own[k].dtype) if k in sd else v) for k, v in own.items()}
    unknown = [k for k in sd if k not in own]
    assert not unknown, f"golden has keys the module lacks: {unknown}"
    module.load_state_dict(merged, strict=True)
    grid = getattr(module, "occupancy_grid", None)
    if grid is not None and "occupancy_grid._binary" in sd:
        grid._binary = sd["occupancy_grid._binary"].to(grid.occs.device).bool()


# --------------------------------------------------------------------------- #
# the CUDA product
# --------------------------------------------------------------------------- #
def build_product_nerf(cfg, device, seed=0):
    from deblur_e_nerf_b200 import nerf as nerf_mod
    from deblur_e_nerf_b200.nerfacc import ContractionType
    torch.manual_seed(seed)
    model = nerf_mod.NeRF(cfg["aabb"], ContractionType[CONTRACTIONS[cfg["contraction"]]],
                          cfg["occ_grid"], cfg["near_plane"], cfg["far_plane"], cfg["step"],
                          cfg["render_bkgd"], cfg["cone_angle"], cfg["early_stop_eps"],
                          cfg["alpha_thre"], cfg["test_chunk_size"], "ngp", cfg["arch"], 3,
                          cfg.get("radiance_dim", 1))
    return model.to(device)


def build_product_renderer(cfg, device, it_sample_size=8, pixel_bandwidth=True, seed=0,
                           n_poses=200):
    from deblur_e_nerf_b200 import event_generation_params as egp
    from deblur_e_nerf_b200 import loss as loss_mod
    from deblur_e_nerf_b200 import renderer, trajectories
    nerf = build_product_nerf(cfg, device, seed)
    poses = synthetic.camera_poses(cfg, n_poses=n_poses)
    calib = synthetic.calibration()
    traj = trajectories.LinearTrajectory(poses)
    ct = egp.ContrastThreshold(calib, True)
    rp = egp.RefractoryPeriod(calib, synthetic.MAX_REFRACTORY_PERIOD_NS)
    pb = None
    if pixel_bandwidth:
        from deblur_e_nerf_b200 import pixel_bandwidth as pb_mod
        pb = pb_mod.PixelBandwidth(calib, poses[2].min(), 21, dict(max_sample_lifetime=0.95))
    loss_cfg = {k: dict(v) for k, v in LOSS_CFG.items()}
    loss_cfg["weight"]["log_intensity_tv"] = cfg["tv_weight"]
    loss = loss_mod.Loss(loss_cfg["weight"], loss_cfg["error_fn"], loss_cfg["normalize"])
    kinv = torch.linalg.inv(torch.from_numpy(synthetic.intrinsics(cfg)))
    model = renderer.EventRenderer(nerf, traj, ct, rp, pb, loss, kinv)
    return model.to(device), poses


def to_device(tree, device):
    return {k: v.to(device) for k, v in tree.items()}


RAW_EVENT_CASES = ("ordered", "hot", "shuffled", "sparse")


def raw_event_case(name):
    """One case of tests/golden/raw_events.npz: (raw dict, height, width, bayer pattern, queued dict, max
    refractory period) as the reference's own data/datasets.py produced them."""
    gold = load_golden("raw_events")
    raw = {k: gold[f"{name}/raw/{k}"] for k in ("position", "timestamp", "polarity")}
    queued = {k[len(name) + 8:]: gold[k] for k in gold if k.startswith(f"{name}/queued/")}
    return (raw, int(gold[f"{name}/height"]), int(gold[f"{name}/width"]), str(gold[f"{name}/bayer_pattern"]),
            queued, np.asarray(gold[f"{name}/max_refractory_period"]))
