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
