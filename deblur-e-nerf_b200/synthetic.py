"""Seeded synthetic inputs of the reference's data layout.

No dataset is needed: every constant comes from the reference tree —
``camera_calibration.npz`` keys (``scripts/preprocess_esim.py:211-250``), the EDS /
DVS346 calibration constants (``scripts/eds_to_esim.py:59-79,123-133``), the event batch
layout (``data/datasets.py:222-228``: ``position (N,2) f32``, ``start_ts/end_ts (N) int64 ns``,
``num_pos/num_neg (N) int64``) and the normalised samplers
(``data/datamodule.py:151-213``: ``ts_diff == 1``, triangular ``ts_subdiff``, uniform
starts, ``interval_gen == 0.5``).  Used by ``bench.py``, ``smoke()`` and the tests.
"""

import math
import os

import numpy as np
import torch

CONFIGS = {
    # configs/train/synthetic.yaml
    "synthetic": dict(
        aabb=[-1.5, -1.5, -1.5, 1.5, 1.5, 1.5], contraction="aabb", occ_resolution=128,
        near_plane=1.43, far_plane=6.63, cone_angle=0.0, render_bkgd="parameter",
        width=346, height=260, focal=300.0, orbit_radius=4.03, tv_weight=1e-3,
        early_stop_eps=1e-4, alpha_thre=0.0, test_chunk_size=16384),
    # configs/train/08_peanuts_running.yaml
    "eds": dict(
        aabb=[0.2, -0.4, 0.0, 3.7, 3.7, 1.8], contraction="sphere", occ_resolution=256,
        near_plane=0.01, far_plane=13.0, cone_angle=0.004, render_bkgd=None,
        width=640, height=480, focal=560.0, orbit_radius=None, tv_weight=1e-1,
        early_stop_eps=1e-4, alpha_thre=0.0, test_chunk_size=16384),
}

POS_ENCODING = dict(otype="HashGrid", n_levels=16, n_features_per_level=2, log2_hashmap_size=19,
                    base_resolution=16, per_level_scale=1.4472692012786865,
                    interpolation="Linear")
SMALL_POS_ENCODING = dict(otype="HashGrid", n_levels=4, n_features_per_level=2,
                          log2_hashmap_size=14, base_resolution=16,
                          per_level_scale=1.4472692012786865, interpolation="Linear")
ARCH = dict(
    dir_encoding=dict(degree=4),
    mlp_base=dict(hidden_activation="softplus", density_activation="shifted_trunc_exp",
                  n_neurons=64, n_hidden_layers=1, geo_feat_dim=15, weight_norm=False),
    mlp_head=dict(hidden_activation="softplus", radiance_activation="softplus", n_neurons=64,
                  n_hidden_layers=2, weight_norm=False),
)


def arch_config(small=False):
    cfg = {k: dict(v) for k, v in ARCH.items()}
    cfg["pos_encoding"] = dict(SMALL_POS_ENCODING if small else POS_ENCODING)
    return cfg


def render_step_size(aabb):
    """models/deblur_e_nerf.py:277-283 — sqrt(3) * max side / 1024."""
    lo, hi = np.asarray(aabb[:3]), np.asarray(aabb[3:])
    return math.sqrt(3) * float((hi - lo).max()) / 1024


def calibration():
    """camera_calibration.npz content (0-d fp32 arrays; refractory period int64 ns)."""
    tau_s = 4e-23 * math.exp(27.64 * 1.5)
    return {
        "pos_contrast_threshold": np.array(0.25 * 1.075, dtype=np.float32),
        "neg_contrast_threshold": np.array(0.25, dtype=np.float32),
        "refractory_period": np.array(int(round(tau_s * 1e9))),
        "bayer_pattern": np.array(""),
        "distortion_model": np.array("plumb_bob"),          # data/datasets.py:24-25: an ideal pinhole camera
        "distortion_params": np.zeros(0),
        "input_time_const_eff_it_prod": np.array(4.375e-4, dtype=np.float32),
        "miller_time_const_eff_it_prod": np.array(7.5e-6, dtype=np.float32),
        "amplifier_gain": np.array(140.0, dtype=np.float32),
        "closed_loop_gain": np.array(1 / 0.7, dtype=np.float32),
        "output_time_const": np.array(25e-6, dtype=np.float32),
        "sf_cutoff_freq": np.array(16.4e3, dtype=np.float32),
        "diff_amp_cutoff_freq": np.array(82e3, dtype=np.float32),
    }


MAX_REFRACTORY_PERIOD_NS = 200_000


def intrinsics(cfg):
    k = np.array([[cfg["focal"], 0, cfg["width"] / 2], [0, cfg["focal"], cfg["height"] / 2],
                  [0, 0, 1]], dtype=np.float32)
    return k


def _look_at_quat(pos, target):
    z = target - pos
    z = z / np.linalg.norm(z)
    up = np.array([0.0, 0.0, 1.0])
    x = np.cross(z, up)
    x = x / np.linalg.norm(x)
    y = np.cross(z, x)
    rot = np.stack([x, y, z], axis=1)            # camera-to-world, columns = camera axes
    # rotation matrix -> xyzw quaternion
    tr = np.trace(rot)
    if tr > 0:
        s = math.sqrt(tr + 1.0) * 2
        q = [(rot[2, 1] - rot[1, 2]) / s, (rot[0, 2] - rot[2, 0]) / s,
             (rot[1, 0] - rot[0, 1]) / s, 0.25 * s]
    else:
        i = int(np.argmax(np.diag(rot)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = math.sqrt(rot[i, i] - rot[j, j] - rot[k, k] + 1.0) * 2
        q = [0.0, 0.0, 0.0, 0.0]
        q[i] = 0.25 * s
        q[j] = (rot[j, i] + rot[i, j]) / s
        q[k] = (rot[k, i] + rot[i, k]) / s
        q[3] = (rot[k, j] - rot[j, k]) / s
    return np.asarray(q)


def camera_poses(cfg, n_poses=1000, dt_ns=1_000_000, sweep=1.2):
    """C poses 1 ms apart: an orbit looking at the origin (synthetic) or a slow pan inside
    the room AABB (EDS).  Returns (position (C,3) f32, quat xyzw (C,4) f32, ts (C) int64)."""
    ts = np.arange(n_poses, dtype=np.int64) * dt_ns
    pos, quat = [], []
    aabb = np.asarray(cfg["aabb"], dtype=np.float64)
    centre = 0.5 * (aabb[:3] + aabb[3:])
    for i in range(n_poses):
        a = sweep * i / max(n_poses - 1, 1)
        if cfg["orbit_radius"] is not None:
            r = cfg["orbit_radius"]
            p = np.array([r * math.cos(a) * math.cos(0.35), r * math.sin(a) * math.cos(0.35),
                          r * math.sin(0.35)])
            target = np.zeros(3)
        else:
            ext = aabb[3:] - aabb[:3]
            p = centre + np.array([0.30 * ext[0] * math.cos(a), 0.30 * ext[1] * math.sin(a),
                                   0.10 * ext[2] * math.sin(2 * a)])
            target = centre + np.array([0.35 * ext[0] * math.cos(a + 2.2),
                                        0.35 * ext[1] * math.sin(a + 2.2), 0.0])
        pos.append(p)
        quat.append(_look_at_quat(p, target))
    quat = np.asarray(quat)
    for i in range(1, n_poses):                   # keep a continuous sign
        if np.dot(quat[i], quat[i - 1]) < 0:
            quat[i] = -quat[i]
    return (torch.tensor(np.asarray(pos), dtype=torch.float32),
            torch.tensor(quat, dtype=torch.float32), torch.tensor(ts))


def event_batch(n, cfg, pose_ts, generator):
    """One batch of queued events (data/datasets.py:222-228 layout), CPU tensors."""
    t_end = int(pose_ts[-1].item())
    lo = 25_000_000
    end_ts = lo + (torch.rand(n, generator=generator, dtype=torch.float64) * (t_end - lo))
    dur = (0.2 + 19.8 * torch.rand(n, generator=generator, dtype=torch.float64)) * 1e6
    end_ts = end_ts.round().to(torch.int64)
    start_ts = end_ts - dur.round().to(torch.int64)
    pol = (torch.rand(n, generator=generator) < 0.5).to(torch.int64)
    position = torch.stack([
        torch.rand(n, generator=generator) * (cfg["width"] - 1),
        torch.rand(n, generator=generator) * (cfg["height"] - 1)], dim=-1).float()
    return {"position": position, "start_ts": start_ts, "end_ts": end_ts,
            "num_pos": pol, "num_neg": 1 - pol}


def normalized_batch(n, it_sample_size, generator, pixel_bandwidth=True):
    """data/datamodule.py:151-213 samplers, all float64."""
    u = torch.rand(n, generator=generator, dtype=torch.float64)
    tri = torch.where(u <= 0.0, torch.sqrt(u * 0.0), 1.0 - torch.sqrt((1 - u) * 1.0))
    out = {
        "ts_diff": torch.ones(n, dtype=torch.float64),
        "diff_start_ts": torch.rand(n, generator=generator, dtype=torch.float64),
        "ts_subdiff": tri,
        "subdiff_start_ts": torch.rand(n, generator=generator, dtype=torch.float64),
    }
    if pixel_bandwidth:
        out["interval_gen"] = torch.full((it_sample_size - 1, n), 0.5, dtype=torch.float64)
    return out


def write_dataset_dir(path, cfg, poses=None):
    """Materialise the three files the reference's constructors read (tests only)."""
    os.makedirs(path, exist_ok=True)
    calib = dict(calibration())
    calib["intrinsics"] = intrinsics(cfg)
    calib["img_height"] = np.array(cfg["height"], dtype=np.uint16)
    calib["img_width"] = np.array(cfg["width"], dtype=np.uint16)
    np.savez(os.path.join(path, "camera_calibration.npz"), **calib)
    if poses is None:
        poses = camera_poses(cfg)
    np.savez(os.path.join(path, "camera_poses.npz"), T_wc_position=poses[0].numpy(),
             T_wc_orientation=poses[1].numpy(), T_wc_timestamp=poses[2].numpy())
    torch.save(torch.tensor(MAX_REFRACTORY_PERIOD_NS),
               os.path.join(path, "max_refractory_period.pt"))
    return path


def solid_sphere_occupancy(resolution, radius=0.75, shell=None):
    """Controlled occupancy (§8(d) (ii)): cells whose centre lies within `radius` (unit-cube
    half-extent = 1) of the centre; `shell` keeps only a shell of that thickness."""
    r = torch.linspace(-1 + 1 / resolution, 1 - 1 / resolution, resolution)
    gx, gy, gz = torch.meshgrid(r, r, r, indexing="ij")
    dist = torch.sqrt(gx * gx + gy * gy + gz * gz)
    occ = dist < radius
    if shell is not None:
        occ = occ & (dist > radius - shell)
    return occ
