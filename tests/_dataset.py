"""A tiny on-disk dataset in the layout the reference's `data/datasets.py` reads (ESIM-preprocessed
synthetic scene): cached transformed events, calibration, camera poses, the maximum refractory period, and
a `views/` folder with 8-bit BGRA renders + `transforms_{train,val,test}.json` + `renderer_params.npz`."""

import json
import math
import os

import cv2
import numpy as np
import torch

from deblur_e_nerf_b200 import synthetic


def write(root, cfg, n_events=4096, n_views=2, size=(24, 32), seed=0, channels=4):
    """`channels` 4: BGRA renders (composited over white by `alpha_over_white_bg: true`); 3: BGR renders,
    which a mono sensor's loader converts to grey (data/datasets.py:640-644)."""
    os.makedirs(root, exist_ok=True)
    poses = synthetic.camera_poses(cfg, n_poses=200)
    synthetic.write_dataset_dir(root, cfg, poses)
    g = torch.Generator().manual_seed(seed)
    events = synthetic.event_batch(n_events, cfg, poses[2], g)
    torch.save({k: v for k, v in events.items()}, os.path.join(root, "events.pt"))
    torch.save(torch.tensor(synthetic.MAX_REFRACTORY_PERIOD_NS), os.path.join(root, "max_refractory_period.pt"))
    np.savez(os.path.join(root, "renderer_params.npz"), interm_color_space=np.array("display"),
             log_eps=np.array(1e-3, dtype=np.float32))
    views = os.path.join(root, "views")
    h, w = size
    rng = np.random.default_rng(seed)
    for stage in ("train", "val", "test"):
        os.makedirs(os.path.join(views, stage), exist_ok=True)
        frames = []
        for i in range(n_views):
            img = rng.integers(20, 235, size=(h, w, channels), dtype=np.uint8)
            if channels == 4:
                img[..., 3] = 255
            cv2.imwrite(os.path.join(views, stage, f"r_{i}.png"), img)
            ang = 0.3 * i
            pos = np.array([4.0 * math.cos(ang), 4.0 * math.sin(ang), 0.5])
            z = pos / np.linalg.norm(pos)                         # OpenGL camera looks along -z
            x = np.cross([0.0, 0.0, 1.0], z)
            x /= np.linalg.norm(x)
            y = np.cross(z, x)
            T = np.eye(4)
            T[:3, :3] = np.stack([x, y, z], axis=1)
            T[:3, 3] = pos
            frames.append({"file_path": f"{stage}/r_{i}", "transform_matrix": T.tolist()})
        with open(os.path.join(views, f"transforms_{stage}.json"), "w") as fh:
            json.dump({"camera_angle_x": 0.9, "frames": frames}, fh)
    return poses


def reference_style_config(data_dir, it_sample_size=4):
    """A config with the structure and keys of the reference's configs/train/synthetic.yaml (the subset the
    Lightning-free stack reads), sized for tests."""
    arch = synthetic.arch_config(small=True)
    frozen_pb = {k: True for k in ("tau_mil_it_eff_prod", "A_amp_inv", "A_loop_inv", "tau_out", "tau_sf", "tau_diff",
                                   "default")}
    return {
        "seed": 3, "float32_matmul_precision": "highest", "eval_target": ["novel_view"],
        "data": {"dataset_directory": data_dir, "train_dataset_ratio": 1.0, "val_dataset_ratio": 1.0,
                 "test_dataset_ratio": 1.0, "train_dataset_perm_seed": None, "eval_dataset_perm_seed": 9,
                 "alpha_over_white_bg": False, "train_init_eff_batch_size": 64,
                 "train_eff_ray_sample_batch_size": 16384, "val_eff_batch_size": 1, "test_eff_batch_size": 1},
        "model": {
            "min_modeled_intensity": 0.001, "eval_save_pred_intensity_img": False, "checkpoint_filepath": None,
            "contrast_threshold": {"parameterize_mean_ct": True, "load_state_dict": False,
                                   "freeze": {"p2n_contrast_threshold_ratio": True, "mean_contrast_threshold": True,
                                              "default": True}},
            "refractory_period": {"load_state_dict": False, "freeze": True},
            "pixel_bandwidth": {"enable": True, "it_sample_size": it_sample_size, "f_c_dominant_min": 21,
                                "target_cumprob": {"max_sample_lifetime": 0.95}, "load_state_dict": False,
                                "freeze": frozen_pb},
            "nerf": {"aabb": [-1.5, -1.5, -1.5, 1.5, 1.5, 1.5], "contraction_type": "aabb",
                     "occ_grid": {"resolution": 32, "occ_thre": 1.0e-2, "ema_decay": 0.95, "warmup_steps": 256, "n": 16},
                     "near_plane": 1.43, "far_plane": 6.63, "render_step_size": "auto", "cone_angle": 0,
                     "early_stop_eps": 1.0e-4, "alpha_thre": 0, "test_chunk_size": 16384, "arch": "ngp",
                     "load_state_dict": False, "freeze": False, "ngp": arch},
            "correction": {"per_channel_log_it_scale": False, "black_level_offset": True,
                           "optimizer": {"algo": "lm", "max_steps": 10, "lm": {"radius": 1.0e6}}},
        },
        "loss": {"error_fn": {"log_intensity_diff": "huber", "log_intensity_tv": "l1"},
                 "weight": {"log_intensity_diff": 1.0, "log_intensity_tv": 1.0e-3, "nerf_mlp_weight_decay": 1.0e-6},
                 "normalize": {"log_intensity_diff": True, "log_intensity_tv": True}},
        "optimizer": {"algo": "adam",
                      "lr": {"contrast_threshold": {"p2n_contrast_threshold_ratio": 0.1, "mean_contrast_threshold": 0.1},
                             "pixel_bandwidth": {k: 0.01 for k in ("tau_mil_it_eff_prod", "A_amp_inv", "A_loop_inv",
                                                                   "tau_out", "tau_sf", "tau_diff")},
                             "default": 0.01},
                      "relative_lr": {"refractory_period": 50}},
        "lr_scheduler": {"algo": "multi_step_lr", "interval": "epoch",
                         "multi_step_lr": {"milestones": [20, 30, 36], "gamma": 0.33}},
        "checkpoint": {"every_n_epochs": 1},
        "trainer": {"max_epochs": 1, "log_every_n_steps": 1, "limit_train_batches": 4, "limit_val_batches": 0,
                    "check_val_every_n_epoch": 1},
    }


def write_raw_events(root, cfg, n=30000, patch=40, seed=0):
    """Replace the cached events.pt of `write()` by a raw stream (`raw_events.npz`, data/datasets.py:19-21) on a
    hot patch of the sensor, inside the time span of the camera poses."""
    os.remove(os.path.join(root, "events.pt"))
    rng = np.random.default_rng(seed)
    x0, y0 = cfg["width"] // 2 - patch // 2, cfg["height"] // 2 - patch // 2
    position = np.stack([x0 + rng.integers(0, patch, n), y0 + rng.integers(0, patch, n)], axis=1).astype(np.uint16)
    timestamp = np.sort(rng.integers(30_000_000, 190_000_000, n)).astype(np.int64)
    np.savez(os.path.join(root, "raw_events.npz"), position=position, timestamp=timestamp,
             polarity=rng.random(n) < 0.5)
