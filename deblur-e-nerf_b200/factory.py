"""Build the hot-path modules from the reference's YAML shapes (synthetic.yaml /
08_peanuts_running.yaml) with synthetic calibration and poses, and the optimizer with the
reference's parameter groups (models/deblur_e_nerf.py:1055-1112)."""

import torch

from . import event_generation_params as egp
from . import loss as loss_mod
from . import nerf as nerf_mod
from . import renderer as renderer_mod
from . import synthetic, trajectories
from .nerfacc import ContractionType

CONTRACTIONS = {"aabb": ContractionType.AABB, "sphere": ContractionType.UN_BOUNDED_SPHERE,
                "tanh": ContractionType.UN_BOUNDED_TANH}

LOSS = dict(
    error_fn=dict(log_intensity_diff="huber", log_intensity_tv="l1"),
    normalize=dict(log_intensity_diff=True, log_intensity_tv=True),
)


def build_renderer(config="synthetic", device="cuda", pixel_bandwidth=True, small=False,
                   occ_resolution=None, n_poses=1000, world_size=1, sample_budget=131072,
                   accumulate_grad_batches=1, seed=0):
    cfg = dict(synthetic.CONFIGS[config])
    if occ_resolution is not None:
        cfg["occ_resolution"] = occ_resolution
    torch.manual_seed(seed)
    arch = synthetic.arch_config(small=small)
    occ = dict(resolution=cfg["occ_resolution"], occ_thre=1e-2, ema_decay=0.95,
               warmup_steps=256, n=16)
    nerf = nerf_mod.NeRF(cfg["aabb"], CONTRACTIONS[cfg["contraction"]], occ, cfg["near_plane"],
                         cfg["far_plane"], synthetic.render_step_size(cfg["aabb"]),
                         cfg["render_bkgd"], cfg["cone_angle"], cfg["early_stop_eps"],
                         cfg["alpha_thre"], cfg["test_chunk_size"], "ngp", arch, 3, 1)
    poses = synthetic.camera_poses(cfg, n_poses=n_poses)
    calib = synthetic.calibration()
    pb = None
    if pixel_bandwidth:
        from . import pixel_bandwidth as pb_mod
        pb = pb_mod.PixelBandwidth(calib, poses[2].min(), 21, dict(max_sample_lifetime=0.95))
    weight = dict(log_intensity_diff=1.0, log_intensity_tv=cfg["tv_weight"],
                  nerf_mlp_weight_decay=1e-6)
    model = renderer_mod.EventRenderer(
        nerf, trajectories.LinearTrajectory(poses), egp.ContrastThreshold(calib, True),
        egp.RefractoryPeriod(calib, synthetic.MAX_REFRACTORY_PERIOD_NS), pb,
        loss_mod.Loss(weight, LOSS["error_fn"], LOSS["normalize"]),
        torch.linalg.inv(torch.from_numpy(synthetic.intrinsics(cfg))),
        train_ray_sample_batch_size=sample_budget,
        accumulate_grad_batches=accumulate_grad_batches, world_size=world_size)
    return model.to(device), cfg, poses


def freeze_like_synthetic_yaml(model):
    """configs/train/synthetic.yaml:31-55 — C_p, tau and the pixel-bandwidth parameters are
    frozen; only the NeRF trains."""
    for module in (model.contrast_threshold, model.refractory_period, model.pixel_bandwidth):
        if module is not None:
            module.requires_grad_(False)
    return model


# configs/train/*.yaml `optimizer.lr` (synthetic.yaml:146-157): per-parameter learning rates of the
# multi-parameter components, in YAML order (the order fixes the optimizer's group indices)
COMPONENT_LR = {
    "contrast_threshold": {"p2n_contrast_threshold_ratio": 0.1, "mean_contrast_threshold": 0.1},
    "pixel_bandwidth": {"tau_mil_it_eff_prod": 0.01, "A_amp_inv": 0.01, "A_loop_inv": 0.01,
                        "tau_out": 0.01, "tau_sf": 0.01, "tau_diff": 0.01},
}


def optimizer_param_groups(model, weight_decay=1e-6, refractory_relative_lr=50.0, component_lr=None):
    """The parameter groups of DeblurENeRF.configure_optimizers (models/deblur_e_nerf.py:1055-1090),
    group for group: [0] the refractory-period parameters at lr = tau_max * relative_lr, [1] every
    parameter named `nerf.radiance_field.mlp*` (hash table included) with the weight decay, then ONE
    group per `.original` parameter of each multi-parameter component (contrast threshold, and the
    pixel bandwidth when enabled) with its own lr in YAML order, then the remaining parameters.
    Frozen parameters stay in their groups like upstream (Adam skips parameters without a gradient), so
    `optimizer.load_state_dict` of a reference checkpoint finds the group layout it saved.  The last
    group is a `set` difference upstream (unordered); here it follows `named_parameters()`."""
    component_lr = COMPONENT_LR if component_lr is None else component_lr
    tau_lr = float(model.refractory_period.max_refractory_period) * refractory_relative_lr
    groups = [
        dict(params=list(model.refractory_period.parameters()), lr=tau_lr),
        dict(params=[p for n, p in model.named_parameters()
                     if n.startswith("nerf.radiance_field.mlp")], weight_decay=weight_decay),
    ]
    for component in ("contrast_threshold", "pixel_bandwidth"):      # MULTI_PARAM_MODEL_COMPONENTS
        module = getattr(model, component, None)
        if module is None:
            continue
        for name, lr in component_lr.get(component, {}).items():
            if not hasattr(module.parametrizations, name):
                if component == "contrast_threshold" and name == "mean_contrast_threshold":
                    continue        # only a parameter when `parameterize_mean_ct` is true (yaml:149)
                raise AttributeError(f"{component} has no parametrized `{name}`")
            groups.append(dict(params=[getattr(module.parametrizations, name).original], lr=lr))
    taken = {id(p) for g in groups for p in g["params"]}
    groups.append(dict(params=[p for p in model.parameters() if id(p) not in taken]))
    assert sum(len(g["params"]) for g in groups) == len(list(model.parameters()))
    return groups


def configure_optimizer(model, lr=0.01, weight_decay=1e-6, refractory_relative_lr=50.0,
                        component_lr=None, fused=None):
    """Adam over `optimizer_param_groups` (den_adam_step through optim.FusedAdam on CUDA)."""
    groups = optimizer_param_groups(model, weight_decay, refractory_relative_lr, component_lr)
    if fused is None:
        fused = all(p.is_cuda for p in model.parameters())
    if fused:
        from .optim import FusedAdam          # den_adam_step: two launches per step
        opt = FusedAdam(groups, lr=lr)
        # a sync-free step that ran out of sample capacity lost samples: its update is skipped on the device
        opt.skip_flag = getattr(model.nerf, "overflow_flag", None)
        return opt
    return torch.optim.Adam(groups, lr=lr)
