#!/usr/bin/env python
"""Benchmark of the Deblur e-NeRF training hot path (event-supervised volumetric renderer).

    python bench.py --gpus N --steps K --warmup W [--workload NAME] [--impl reference]

A "step" is one full training step of the hot path on one synthetic event batch: event
correction -> supervision timestamps -> 4 render calls (march, field, compositing[, pixel-
bandwidth filter]) -> event loss -> backward -> (N>1: flat NCCL gradient all-reduce) -> Adam.
`value` is whole-job rays/s with the batch already resident in HBM; `e2e` is the same metric
through the public `EventRenderer.training_step` call with HOST (pinned) batches, the H2D
copies and the loss read-back inside the timed region.

Workloads (BASELINE.json `configs`):
  synthetic_pb_off   configs[1]: synthetic.yaml shape, pixel-bandwidth model off, 2^17 rays per
                     render call (the literal "2^17 ray batch" reading)
  synthetic_pb_on    configs[2]: synthetic.yaml hard setting, it_sample_size 30, 2^17 rays per
                     render call (N = 2^17/30 events)
  synthetic_budget   reading (R): N = budget(2^17 samples)/mean samples per ray, PB on
  eds                configs[3] shape (sphere contraction, cone 0.004, res 256, no background)
  plumbing           configs[0]: 4096 events x 8 samples, small hash grid (CPU-runnable)
  render_sweep       configs[4]: eval-mode 800x800 novel views in 16 384-ray chunks (march + hash grid +
                     MLP + compositing, no event loss); a step is one view, `value` is rays/s
`--impl reference` times the reference's path on the host cores instead: the CPU oracle
(oracle/path_ref.py = the reference's own files restated and pinned against them; nerfacc /
tiny-cuda-nn are CUDA-only, so their pure-PyTorch equivalents are used and labelled).
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    "synthetic_pb_off": dict(config="synthetic", pb=False, S=1, rays_per_call=1 << 17, small=False,
                             occ_res=128),
    "synthetic_pb_on": dict(config="synthetic", pb=True, S=30, rays_per_call=1 << 17, small=False,
                            occ_res=128),
    "synthetic_budget": dict(config="synthetic", pb=True, S=30, rays_per_call=None, small=False,
                             occ_res=128),
    # configs[3]: 08_peanuts_running.yaml shape — 8 accumulated micro-batches of 2^17 rays per render call
    # (2^20 effective), C_p / tau / Omega trainable (:31-55,203)
    "eds": dict(config="eds", pb=True, S=30, rays_per_call=1 << 17, small=False, occ_res=256, acc=8,
                unfrozen=True),
    "plumbing": dict(config="synthetic", pb=True, S=8, rays_per_call=4096 * 8, small=True,
                     occ_res=32),
    # configs[4]: test-mode novel-view sweep, 800x800 views in 16 384-ray chunks, forward only
    "render_sweep": dict(config="synthetic", pb=False, S=1, rays_per_call=800 * 800, small=False,
                         occ_res=128, sweep=True),
    # the data format in front of the path: raw event stream -> queued events + maximum refractory period
    # (data/datasets.py:131-276) on a 640 x 480 sensor (EDS, scripts/eds_to_esim.py:53-57)
    "raw_events": dict(raw_events=True, n_events=1 << 24, height=480, width=640),
}

# algorithmic bytes per sample (SURVEY.md §8(d) / BASELINE.md §3), 16 levels x 8 corners x 8 B
HASH_GATHER_BYTES = 1024
KERNEL_BYTES_PER_SAMPLE = {
    "den_hashgrid_fwd": 1164,        # gather + 12 B position + 128 B encoding written
    "den_hashgrid_bwd": 1164,        # scatter (counted once) + 128 B dL/denc + 12 B position
    "den_field_fwd": 1036,           # fused: gather + 12 B sample in, no encoding write
    "den_field_bwd": 1036 + 1024,    # fused recompute gather + scatter
    "den_composite_fwd": 20,         # sigma, rgb (C=1), t0, t1 read + 12 B/ray out (added per launch)
    "den_composite_bwd": 28,         # the same reads + d_sigma, d_rgb written + 12 B/ray grads read
}
KERNEL_BYTES_PER_RAY = {"den_composite_fwd": 12, "den_composite_bwd": 12}
# tensor-pipe kernels: algorithmic FLOP per sample
KERNEL_FLOP_PER_SAMPLE = {"den_mlp_fwd": 18432, "den_mlp_bwd": 55296}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="synthetic_pb_on", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-events", type=int, default=0,
                    help="events per step of the bounded CPU sample (0 = sized from a calibration "
                         "step so that the CPU run stays within --cpu-seconds)")
    ap.add_argument("--cpu-seconds", type=float, default=150.0,
                    help="time budget of the --impl reference run (all warm-up + timed steps)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="--gpus N: split the fixed global batch over the ranks (strong, the north-star "
                         "reading: models/deblur_e_nerf.py:72-75) or give every rank the full batch")
    ap.add_argument("--no-graph", action="store_true",
                    help="run every step eagerly instead of replaying the captured CUDA graph")
    ap.add_argument("--emulate-ranks", type=int, default=1,
                    help="diagnostic: run ONE rank's share of an N-rank strong-scaling step on this GPU "
                         "(no collective); reported in config, never a scaling result")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------- clocks ----
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()
        # nvidia-smi's start-up (driver attach, a few hundred ms) must not overlap the warm-up / timed
        # steps: an eagerly launched step (277 launches) stalls behind it.  Wait for the first line.
        t0 = time.time()
        while not self.rows and time.time() - t0 < 3.0 and self.proc.poll() is None:
            time.sleep(0.02)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def window(self, t0, t1):
        """Keep the samples taken inside [t0, t1] (the sampler is started before the warm-up steps:
        nvidia-smi needs a few hundred ms to deliver its first line)."""
        inside = [r for r in self.rows if t0 <= r[0] <= t1]
        self.rows = inside if inside else self.rows[-3:]

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, sm_max, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for _, row in self.rows:
            try:
                sm.append(float(row[0]))
                sm_max = float(row[1])
            except (ValueError, IndexError):
                continue
            for name, flag in zip(names, row[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": sm_max,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------- CPU / reference ----
def _oracle_scene(workload, n_events, seed=0):
    """The same workload on the CPU oracle (bounded number of events)."""
    import torch
    from deblur_e_nerf_b200 import synthetic
    from oracle import nerfacc_ref, path_ref

    w = WORKLOADS[workload]
    cfg = dict(synthetic.CONFIGS[w["config"]])
    torch.manual_seed(seed)
    ctype = {"aabb": nerfacc_ref.ContractionType.AABB,
             "sphere": nerfacc_ref.ContractionType.UN_BOUNDED_SPHERE}[cfg["contraction"]]
    occ = dict(resolution=w["occ_res"], occ_thre=1e-2, ema_decay=0.95, warmup_steps=256, n=16)
    nerf = path_ref.NeRF(cfg["aabb"], ctype, occ, cfg["near_plane"], cfg["far_plane"],
                         synthetic.render_step_size(cfg["aabb"]), cfg["render_bkgd"],
                         cfg["cone_angle"], cfg["early_stop_eps"], cfg["alpha_thre"],
                         cfg["test_chunk_size"], synthetic.arch_config(small=w["small"]), 1)
    nerf.occupancy_grid._binary = synthetic.solid_sphere_occupancy(w["occ_res"])
    poses = synthetic.camera_poses(cfg)
    calib = synthetic.calibration()
    pb = path_ref.PixelBandwidth(calib, poses[2].min(), 21, 0.95) if w["pb"] else None
    weight = dict(log_intensity_diff=1.0, log_intensity_tv=cfg["tv_weight"])
    loss = path_ref.EventLoss(weight, dict(log_intensity_diff="huber", log_intensity_tv="l1"),
                              dict(log_intensity_diff=True, log_intensity_tv=True))
    model = path_ref.EventRenderer(
        nerf, path_ref.LinearTrajectory(*poses),
        path_ref.ContrastThreshold(calib["pos_contrast_threshold"],
                                   calib["neg_contrast_threshold"]),
        path_ref.RefractoryPeriod(calib["refractory_period"], synthetic.MAX_REFRACTORY_PERIOD_NS),
        pb, loss, torch.linalg.inv(torch.from_numpy(synthetic.intrinsics(cfg))))
    for m in (model.contrast_threshold, model.refractory_period, model.pixel_bandwidth):
        if m is not None:
            m.requires_grad_(False)
    model.train()
    g = torch.Generator().manual_seed(1000 + seed)
    event = synthetic.event_batch(n_events, cfg, poses[2], g)
    normalized = synthetic.normalized_batch(n_events, w["S"], g, w["pb"])
    return model, event, normalized, w


def cpu_reference_rate(workload, n_events, steps=1, warmup=0):
    """rays/s of the reference's path (CPU oracle) on a bounded sample of the workload."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    model, event, normalized, w = _oracle_scene(workload, n_events)
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=0.01)
    rays_per_step = 4 * w["S"] * n_events
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss, _, _ = model.training_step(event, normalized)
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    mean_t = sum(times) / len(times)
    return {
        "value": rays_per_step / mean_t, "unit": "rays/s", "cores": torch.get_num_threads(),
        "kind": "port",
        "sample": (f"{n_events} events x {w['S']} pixel-bandwidth samples x 4 render calls "
                   f"({rays_per_step} rays/step), fwd+bwd+Adam, {len(times)} timed step(s); "
                   "reference's own files restated in oracle/path_ref.py (pinned against them), "
                   "nerfacc/tiny-cuda-nn replaced by their pure-PyTorch equivalents "
                   "(reference is CUDA-only there)"),
        "ms_per_step": mean_t * 1e3,
    }, rays_per_step


def cpu_sweep_rate(workload, n_rays, steps=1, warmup=0):
    """rays/s of the reference's eval-mode render (CPU oracle) on `n_rays` pixels of an 800x800 view."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    model, _, _, w = _oracle_scene(workload, 8)
    model.eval()
    side = 800
    kinv = torch.linalg.inv(torch.tensor([[side * 1.2, 0, side / 2], [0, side * 1.2, side / 2],
                                          [0, 0, 1.0]]))
    g = torch.Generator().manual_seed(5)
    pix = torch.rand(n_rays, 2, generator=g) * (side - 1)
    ts = model.trajectory.T_wc_timestamp[len(model.trajectory.T_wc_timestamp) // 2].double().expand(n_rays)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        with torch.no_grad():
            pos, rot = model.trajectory(ts)
            o, d = model.nerf.pixel_params_to_ray(kinv, pix, pos, rot)
            model.nerf(o, d)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    mean_t = sum(times) / len(times)
    return {"value": n_rays / mean_t, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": (f"{n_rays} pixels of one 800x800 view, eval mode, forward only, {len(times)} timed "
                       "pass(es); reference's own files restated in oracle/path_ref.py, nerfacc / "
                       "tiny-cuda-nn replaced by their pure-PyTorch equivalents"),
            "ms_per_step": mean_t * 1e3}, n_rays


def _size_cpu_sample(args, per_event_rays):
    """Events per step of the bounded CPU sample: given on the command line, or sized from a small
    calibration step so that warm-up + timed steps fit `--cpu-seconds`."""
    if args.cpu_events:
        return args.cpu_events
    probe = 16
    cal, _ = cpu_reference_rate(args.workload, probe, steps=1, warmup=0)
    per_event_s = cal["ms_per_step"] * 1e-3 / probe
    steps = max(args.steps + args.warmup, 1)
    return int(min(max(args.cpu_seconds / steps / max(per_event_s, 1e-9), 8), 4096))


def run_reference(args):
    """The reference's path on the host cores with the SAME --steps / --warmup as the product arm;
    each step is a bounded sample of the workload (`cpu_baseline.sample` says what it was)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if WORKLOADS[args.workload].get("sweep"):
        base, _ = cpu_sweep_rate(args.workload, 4096, steps=args.steps, warmup=args.warmup)
        print(json.dumps({
            "impl": "reference", "metric": "render rays/s (eval, fwd only)", "value": base["value"],
            "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "view": "800x800", "bounded_sample": True},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": "rays/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}))
        return
    w = WORKLOADS[args.workload]
    n_events = _size_cpu_sample(args, 4 * w["S"])
    base, rays_per_step = cpu_reference_rate(args.workload, n_events, steps=args.steps,
                                             warmup=args.warmup)
    line = {
        "impl": "reference", "metric": "train rays/s (fwd+bwd)", "value": base["value"],
        "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, w, n_events, bounded=True),
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": "rays/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(name, w, n_events, bounded=False):
    return {
        "workload": name, "yaml": f"configs/train/{'synthetic' if w['config'] == 'synthetic' else '08_peanuts_running'}.yaml",
        "pixel_bandwidth": w["pb"], "it_sample_size": w["S"], "events_per_step": n_events,
        "rays_per_render_call": w["S"] * n_events, "render_calls_per_step": 4,
        "hash_grid": "L4 T14" if w["small"] else "L16 F2 T19", "occ_resolution": w["occ_res"],
        "occupancy": "controlled solid sphere r=0.75 (SURVEY §8(d)(ii))",
        "field": "random init", "bounded_sample": bounded,
        "l2_policy": "inputs larger than L2: per-step sample arena + 48 MiB table + gradients "
                     "exceed 126 MB; a fresh batch every step",
    }


# ------------------------------------------------------------------ render sweep ------
def run_sweep(args):
    """BASELINE.json configs[4]: 800x800 novel views, eval mode (deterministic march, 16 384-ray
    chunks, no gradients), one view per step; every rank renders its own views."""
    import torch
    import __graft_entry__ as entry
    from deblur_e_nerf_b200 import ddp, factory, ops, synthetic

    rank, local_rank, world = ddp.init_from_env()
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback for the product)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if rank == 0:
        entry.build()
    ddp.barrier()
    w = WORKLOADS[args.workload]
    model, cfg, poses = factory.build_renderer(w["config"], dev, pixel_bandwidth=False,
                                               occ_resolution=w["occ_res"], seed=0)
    sphere = synthetic.solid_sphere_occupancy(w["occ_res"]).to(dev)
    model.nerf.occupancy_grid._binary = sphere
    ddp.broadcast_parameters(model)
    model.eval()
    side = 800
    kinv = torch.linalg.inv(torch.tensor([[side * 1.2, 0, side / 2], [0, side * 1.2, side / 2],
                                          [0, 0, 1.0]])).to(dev)
    v, u = torch.meshgrid(torch.arange(side, dtype=torch.float32),
                          torch.arange(side, dtype=torch.float32), indexing="ij")
    pix_host = torch.stack((u, v), dim=-1).reshape(-1, 2).pin_memory()
    n_views = args.warmup + args.steps
    view_ts = torch.linspace(float(poses[2][10]), float(poses[2][-10]), world * n_views,
                             dtype=torch.float64)[rank::world].to(dev)

    def render_view(i, pix):
        ts = view_ts[i].expand(pix.shape[0])
        pos, rot = model.trajectory(ts)
        o, d = model.nerf.pixel_params_to_ray(kinv, pix, pos, rot)
        with torch.no_grad():
            radiance, opacity, depth, mean_samples = model.nerf(o, d)
        return radiance, mean_samples

    pix_dev = pix_host.to(dev)
    for i in range(args.warmup):
        render_view(i, pix_dev)
    ddp.barrier()
    torch.cuda.synchronize()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = ops.launch_count()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    samples = 0.0
    for i in range(args.steps):
        _, ms = render_view(args.warmup + i, pix_dev)
        samples += ms * side * side
    end.record()
    ddp.barrier()
    torch.cuda.synchronize()
    ms_step = ddp.max_over_ranks(start.elapsed_time(end) / args.steps, dev)
    launches = ops.launch_count() - launches0
    clock_info = clocks.stop() if rank == 0 else None
    # end to end: pixels from pinned host memory, the rendered view read back
    start.record()
    rendered = []
    for i in range(args.steps):
        img, _ = render_view(args.warmup + i, pix_host.to(dev, non_blocking=True))
        img_host = img.to("cpu")
        rendered.append(img)
    end.record()
    ddp.barrier()
    torch.cuda.synchronize()
    ms_e2e = ddp.max_over_ranks(start.elapsed_time(end) / args.steps, dev)
    rays = ddp.sum_over_ranks(side * side, dev)
    if rank != 0:
        return
    eval_post_info = _time_eval_post(model, rendered[:8], side, dev, cpu=not args.no_cpu_baseline)
    line = {
        "metric": "render rays/s (eval, fwd only)", "value": rays / (ms_step * 1e-3), "unit": "rays/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": args.workload, "yaml": "configs/test/synthetic.yaml shape",
                   "view": "800x800", "test_chunk_size": cfg["test_chunk_size"],
                   "occupancy": "controlled solid sphere r=0.75", "field": "random init",
                   "l2_policy": "a fresh view (camera pose) every step; the 48 MiB table is L2-resident "
                                "by design"},
        "views_per_s": world / (ms_step * 1e-3),
        "samples_per_s": world * samples / args.steps / (ms_step * 1e-3),
        "e2e": {"value": rays / (ms_e2e * 1e-3), "unit": "rays/s",
                "h2d_bytes_per_step": pix_host.numel() * 4, "d2h_bytes_per_step": img_host.numel() * 4,
                "ms_per_step": ms_e2e},
        "gpu_launches": launches, "clocks": clock_info, "roofline": None, "cpu_baseline": None,
        "eval_post": eval_post_info,
    }
    print(json.dumps(line))


def _time_eval_post(model, rendered, side, dev, cpu=True):
    """SURVEY.md 8(f) N4 beside the sweep: the post-processing of `run.py test` (models/deblur_e_nerf.py:
    674-969 — gain-exposure normalisation, float64 log-space affine fit, Levenberg-Marquardt offset-gamma
    refinement, L1 / PSNR / SSIM) over the views just rendered, on the device (`evaluation_epoch_end`,
    den_eval_* kernels), and — the reference does this part on the host — the oracle's restatement of it
    on the host cores with the images copied there, as the reported CPU baseline."""
    import torch
    from deblur_e_nerf_b200 import ops
    B = len(rendered)
    g = torch.Generator(device=dev).manual_seed(5)
    pred = torch.stack(rendered).view(B, side, side) + model.min_modeled_intensity
    exposure = torch.arange(1, B + 1, device=dev) % 3 + 1
    gain = 0.75 + 0.5 * torch.rand(B, generator=g, device=dev)
    norm = gain * exposure / (gain * exposure).mean()
    scene = torch.exp(0.8 * pred.log() + 0.3) * torch.exp(0.02 * torch.randn(pred.shape, generator=g, device=dev))
    target = scene * norm.view(-1, 1, 1) + 0.02
    lo, hi = 0.0, float(target.max()) * 1.05
    outputs = [{"pred_intensity_img": pred[b], "target_intensity_img": target[b],
                "exposure_time": exposure[b], "gain": gain[b]} for b in range(B)]
    model.evaluation_epoch_end(outputs, lo, hi, black_level_offset=True)            # warm-up
    torch.cuda.synchronize()
    launches0 = ops.launch_count()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    metrics, _ = model.evaluation_epoch_end(outputs, lo, hi, black_level_offset=True)
    end.record()
    torch.cuda.synchronize()
    info = {"views": B, "view": f"{side}x{side}", "black_level_offset": True, "ms": start.elapsed_time(end),
            "gpu_launches": ops.launch_count() - launches0,
            "metrics": {k: float(v) for k, v in metrics.items()}}
    if cpu:
        from oracle import eval_ref                      # the cpu_baseline leg: the only use of oracle/ here
        torch.set_num_threads(os.cpu_count() or 1)
        t0 = time.perf_counter()
        want = eval_ref.evaluate(pred.cpu()[:, None], target.cpu()[:, None], exposure.cpu(), gain.cpu(), lo, hi,
                                 black_level_offset=True)
        info["cpu_baseline"] = {"ms": (time.perf_counter() - t0) * 1e3, "kind": "port", "cores": os.cpu_count(),
                                "sample": f"the same {B} views (device -> host copy included, as in the reference)",
                                "metrics": {k: want[k] for k in ("l1", "psnr", "ssim")}}
    return info


def run_raw_events(args):
    """The raw-event preprocessing (`events.transform_raw_events`: pixel keys, stable radix sort, neighbour
    pass, compaction of the kept events) on a synthetic time-ordered stream; a step = one pass over the whole
    stream.  Every rank transforms its own stream (the reference does this once per dataset, on one process:
    replicas only).  `value`: stream resident in HBM; `e2e`: raw arrays in pinned host memory, the queued
    events copied back to the host (where the reference keeps them)."""
    import numpy as np
    import torch
    import __graft_entry__ as entry
    from deblur_e_nerf_b200 import ddp, events, ops

    rank, local_rank, world = ddp.init_from_env()
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback for the product)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if rank == 0:
        entry.build()
    ddp.barrier()
    w = WORKLOADS[args.workload]
    n, height, width = w["n_events"], w["height"], w["width"]
    rng = np.random.default_rng(rank)
    raw = {"position": np.stack([rng.integers(0, width, n), rng.integers(0, height, n)], axis=1).astype(np.uint16),
           "timestamp": (np.cumsum(rng.integers(0, 120, n)) + 1_000_000).astype(np.int64),
           "polarity": rng.random(n) < 0.5}
    calib = {"img_height": height, "img_width": width, "bayer_pattern": "", "distortion_params": np.zeros(0)}
    position, timestamp, polarity = events._raw_to_device(raw, dev)

    scratch = {}                # buffers of one stream size, reused by every pass (no allocator traffic)

    def step_device():
        start_ts, offsets, min_interval, flag = events._stream_pass(position, timestamp, height, width, scratch)
        return events._queued(position, timestamp, polarity, start_ts, offsets, flag, scratch), min_interval

    for _ in range(args.warmup):
        step_device()
    ddp.barrier()
    torch.cuda.synchronize()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = ops.launch_count()
    ops.enable_kernel_timing(["den_queue_raw_events", "den_compact_queued_events"], pool_size=4 * args.steps + 8)
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(args.steps):
        queued, min_interval = step_device()
    end.record()
    ddp.barrier()
    torch.cuda.synchronize()
    timings = ops.kernel_timings()
    n_timed, total_kernel_ms = timings["den_queue_raw_events"]
    compact_ms = timings["den_compact_queued_events"][1] / max(timings["den_compact_queued_events"][0], 1)
    ops.disable_kernel_timing()
    ms_step = ddp.max_over_ranks(start.elapsed_time(end) / args.steps, dev)
    launches = ops.launch_count() - launches0
    clock_info = clocks.stop() if rank == 0 else None
    # end to end: the raw arrays from pinned host memory, the queued events back on the host
    pinned = {k: torch.from_numpy(np.ascontiguousarray(v.astype(np.int32) if k == "position" else v)).pin_memory()
              for k, v in raw.items()}
    # pinned landing buffers for the queued events (capacity = the raw count), allocated once
    landing = {k: torch.empty((n, 2) if k == "position" else (n,), dtype=torch.int64).pin_memory()
               for k in ("position", "start_ts", "end_ts", "num_pos", "num_neg")}
    start.record()
    for _ in range(args.steps):
        pos_d = pinned["position"].to(dev, non_blocking=True)
        ts_d = pinned["timestamp"].to(dev, non_blocking=True)
        pol_d = pinned["polarity"].to(dev, non_blocking=True)
        start_ts, offsets, min_interval, flag = events._stream_pass(pos_d, ts_d, height, width, scratch)
        kept = events._queued(pos_d, ts_d, pol_d, start_ts, offsets, flag, scratch)
        host = {k: landing[k][:len(v)].copy_(v, non_blocking=True) for k, v in kept.items()}
        refractory = events._refractory_tensor(min_interval)          # .item(): also drains the copies
    end.record()
    ddp.barrier()
    torch.cuda.synchronize()
    ms_e2e = ddp.max_over_ranks(start.elapsed_time(end) / args.steps, dev)
    total = ddp.sum_over_ranks(n, dev)
    if rank != 0:
        return
    peaks = {}
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    avg_kernel_ms = total_kernel_ms / max(n_timed, 1)
    traffic, traffic_src = None, None           # dram bytes of one den_queue_raw_events call, from the ncu capture
    traffic_path = os.path.join(ROOT, "profiles", "r02_ncu_raw_events_traffic.json")
    if os.path.exists(traffic_path):
        with open(traffic_path) as fh:
            traffic = json.load(fh)["den_queue_raw_events"]["dram_bytes_per_event"] * n
        traffic_src = "profiles/r02_ncu_raw_events_traffic.json"
    algorithmic = 28.0 * n                 # 8 B position + 8 B timestamp read, 8 B start_ts + 4 B keep flag written
    line = {
        "metric": "raw events/s (queue + max refractory period)", "value": total / (ms_step * 1e-3),
        "unit": "events/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": args.workload, "raw_events_per_step": n, "sensor": f"{width}x{height}",
                   "kept_events": int(len(host["start_ts"])), "max_refractory_period_ns": float(refractory),
                   "parallelism": "replicas only (one stream per rank; upstream runs this once per dataset)",
                   "l2_policy": f"inputs larger than L2: {16 * n / 1e6:.0f} MB of raw events per pass"},
        "e2e": {"value": total / (ms_e2e * 1e-3), "unit": "events/s", "h2d_bytes_per_step": 17 * n,
                "d2h_bytes_per_step": int(sum(v.numel() * v.element_size() for v in host.values())),
                "ms_per_step": ms_e2e},
        "gpu_launches": launches, "clocks": clock_info,
        "roofline": {"kernel": "den_queue_raw_events", "bound": "hbm", "achieved": algorithmic / (avg_kernel_ms * 1e-3) / 1e9,
                     "peak": hbm_peak, "unit": "GB/s", "frac": algorithmic / (avg_kernel_ms * 1e-3) / 1e9 / hbm_peak,
                     "traffic": traffic, "traffic_source": traffic_src, "avg_launch_ms": avg_kernel_ms,
                     "compaction_ms": compact_ms,
                     "peak_source": ("measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)") + " HBM copy",
                     "note": "algorithmic bytes = 28 B per raw event (position + timestamp read, start_ts + keep flag "
                             "written); the entry point runs the key kernel, three radix passes (histogram, scan, "
                             "stable scatter: ~24 B moved per event and pass), the neighbour pass (one random timestamp "
                             "gather, two random scatters: 226 B of DRAM traffic per event at sector / atom "
                             "granularity) and the prefix sum of the keep flags — measured traffic 328 B per event, "
                             "12x the algorithmic bytes; den_compact_queued_events (compaction_ms) follows"},
        "cpu_baseline": None,
    }
    if not args.no_cpu_baseline:
        from oracle import events_ref                       # the cpu_baseline leg: the only use of oracle/ here
        m = 1 << 20                                         # bounded sample: the first 2^20 events of the stream
        t0 = time.perf_counter()
        want = events_ref.queue_raw_events_loop(raw["position"][:m], raw["timestamp"][:m], raw["polarity"][:m], height, width)
        events_ref.max_refractory_period_loop(raw["position"][:m], raw["timestamp"][:m], height, width)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": m / dt, "unit": "events/s", "cores": 1, "kind": "port",
                                "sample": f"the first {m} events of the same stream through the reference's two "
                                          f"per-event loops ({dt:.1f} s; the loops are single-threaded Python)",
                                "kept_events_in_sample": int(len(want["start_ts"]))}
    print(json.dumps(line))


# ------------------------------------------------------------------------- ours ------
def _load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per sample of each rated kernel, from the
    `ncu --set full` captures of the SHIPPED build summarised in profiles/ (written by
    profiles/ncu_summary.py --traffic); absent -> traffic is reported as null."""
    path = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    if not os.path.exists(path):
        return {}, None
    with open(path) as fh:
        data = json.load(fh)
    return {k: float(v["dram_bytes_per_sample"]) for k, v in data.get("kernels", {}).items()}, \
        "profiles/r02_ncu_traffic.json"


def run_ours(args):
    import torch
    import __graft_entry__ as entry
    from deblur_e_nerf_b200 import ddp, factory, ops, synthetic

    rank, local_rank, world = ddp.init_from_env()
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback for the product)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if rank == 0:
        entry.build()
    ddp.barrier()

    w = WORKLOADS[args.workload]
    strong = args.scaling == "strong" and world > 1
    acc = int(w.get("acc", 1))
    model, cfg, poses = factory.build_renderer(
        w["config"], dev, pixel_bandwidth=w["pb"], small=w["small"], occ_resolution=w["occ_res"],
        world_size=world if strong else 1, accumulate_grad_batches=acc, seed=0)
    if not w.get("unfrozen"):
        factory.freeze_like_synthetic_yaml(model)
    model.train()
    sphere = synthetic.solid_sphere_occupancy(w["occ_res"]).to(dev)
    model.nerf.occupancy_grid._binary = sphere
    model.nerf.occupancy_grid.occs.copy_(sphere.reshape(-1).float())
    ddp.broadcast_parameters(model)
    ddp.attach(model)
    reducer = ddp.GradReducer(model)
    opt = factory.configure_optimizer(model)
    reducer.bind(opt)

    # events per step and rank.  Strong scaling (default under --gpus N): the fixed global batch of the
    # workload is split over the ranks, like the reference's per-GPU sample budget
    # (models/deblur_e_nerf.py:72-75); weak: every rank takes the full batch.
    if w["rays_per_call"] is not None:
        n_global = w["rays_per_call"] // w["S"]
        n_events = max(n_global // world, 1) if strong else n_global
        n_events = max(n_events // max(args.emulate_ranks, 1), 1)
    else:
        n_events = 256          # the controller's start (synthetic.yaml:18); adapts below
    rays_per_step = lambda n: 4 * w["S"] * n * acc        # noqa: E731

    def host_batch(i, n):
        g = torch.Generator().manual_seed(10_000 * (rank + 1) + i)
        ev = synthetic.event_batch(n, cfg, poses[2], g)
        nm = synthetic.normalized_batch(n, w["S"], g, w["pb"])
        pin = lambda t: t.pin_memory()          # noqa: E731
        return {"event": {k: pin(v) for k, v in ev.items()},
                "normalized": {k: pin(v) for k, v in nm.items()}}

    def to_dev(batch):
        return {k: {kk: vv.to(dev, non_blocking=True) for kk, vv in v.items()}
                for k, v in batch.items()}

    def nbytes(batch):
        return sum(t.numel() * t.element_size() for v in batch.values() for t in v.values())

    # optimizer-step numbering: the occupancy grid is updated when global_step % 16 == 0
    # (models/nerf.py:200-204, `n: 16`), INSIDE the timed steps; a timed region starts at a step
    # = 8 (mod 16), so K timed steps hold floor((K + 7) / 16) updates — the reference's frequency
    def step_number(i):
        return 64 + 8 - args.warmup + i

    def updates_in(first, count):
        return sum(1 for i in range(first, first + count) if step_number(i) % 16 == 0)

    # One optimizer step = `acc` micro-batches (training_step + backward), the gradient reduction, Adam.
    # After two eager steps it is captured in a CUDA graph and replayed (graph_step.GraphedStep): the
    # sync-free render path has no host read-back, so the whole step is one launch.
    from deblur_e_nerf_b200.graph_step import GraphedStep
    stepper = GraphedStep(model, opt, reducer, acc)
    use_graph = not args.no_graph and w["rays_per_call"] is not None
    if not use_graph:
        stepper.warmup_steps = 1 << 60          # never capture: every step eager

    def one_step(batches, i):
        return stepper(batches, step_number(i))

    total = args.warmup + args.steps
    if w["rays_per_call"] is None:
        # reading (R): let the batch controller settle during extra warm-up steps
        for i in range(6):
            one_step([to_dev(host_batch(100 + i * acc + m, n_events)) for m in range(acc)], 1 - step_number(0))
            n_events = max(model.next_train_batch_size or n_events, 1)
    host_batches = [[host_batch(i * acc + m, n_events) for m in range(acc)] for i in range(total)]
    dev_batches = [[to_dev(b) for b in bs] for bs in host_batches]
    torch.cuda.synchronize()

    # kernels with a roofline (bracketed with pooled CUDA events inside the timed region; every
    # other entry point is only bracketed in the separate profile pass further down)
    rated_kernels = sorted(set(KERNEL_BYTES_PER_SAMPLE) | set(KERNEL_FLOP_PER_SAMPLE) | {"den_adam_step"})

    # ---- device-resident loop (value) ------------------------------------------------
    samples_seen = torch.zeros((), dtype=torch.float64, device=dev)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    # set-up (not warm-up): the stepper's two eager steps, the graph capture and the first replays
    # (NCCL registers the captured all-reduce on its first replays), then the W warm-up steps
    setup_steps = 6 if use_graph else 0
    for i in range(setup_steps):
        one_step(dev_batches[i % total], i - 16)
    for i in range(args.warmup):
        one_step(dev_batches[i], i)
    ddp.barrier()
    torch.cuda.synchronize()
    t_wall0 = time.time()
    launches0 = ops.launch_count()
    replays0 = stepper.replays
    mallocs0 = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
    if not use_graph:
        ops.enable_kernel_timing(rated_kernels)
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step_marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    start.record()
    for i in range(args.steps):
        one_step(dev_batches[args.warmup + i], args.warmup + i)
        samples_seen += model.logged["train/mean_num_samples_per_ray"] * rays_per_step(n_events)
        step_marks[i].record()                 # per-step boundaries (diagnostic: `step_ms`)
    end.record()
    graph_stats = {"enabled": use_graph, "captures": stepper.captures, "replays": stepper.replays,
                   "overflows": stepper.overflows, "setup_steps_before_warmup": setup_steps}
    ddp.barrier()
    torch.cuda.synchronize()
    t_wall1 = time.time()
    ms_total = start.elapsed_time(end)
    step_ms = [round((start if i == 0 else step_marks[i - 1]).elapsed_time(step_marks[i]), 3)
               for i in range(args.steps)]
    samples_seen = float(samples_seen)
    ms_step = ddp.max_over_ranks(ms_total / args.steps, dev)
    timings = ops.kernel_timings()
    ops.disable_kernel_timing()
    # kernels launched in the timed region: the eager ones plus, per replay, those recorded in the graph
    launches = (ops.launch_count() - launches0) + stepper.launches_per_replay * (
        stepper.replays - replays0)
    mallocs = torch.cuda.memory_stats(dev).get("num_device_alloc", 0) - mallocs0
    if rank == 0:
        time.sleep(0.15)                    # let the last in-window sample arrive
        clocks.window(t_wall0, t_wall1)
    clock_info = clocks.stop() if rank == 0 else None
    global_rays = ddp.sum_over_ranks(rays_per_step(n_events), dev)
    global_samples = ddp.sum_over_ranks(samples_seen / args.steps, dev)
    value = global_rays / (ms_step * 1e-3)

    # ---- end-to-end loop through the public API with host batches ---------------------
    e2e = None
    if not args.no_e2e:
        for i in range(min(args.warmup, 2)):
            one_step(host_batches[i] if use_graph else [to_dev(b) for b in host_batches[i]], i).item()
        ddp.barrier()
        torch.cuda.synchronize()
        start.record()
        d2h = 0
        for i in range(args.steps):
            # pinned host batches: the copies to the device are part of the step (into the graph's
            # static input buffers when the step is replayed)
            loss = one_step(host_batches[args.warmup + i] if use_graph
                            else [to_dev(b) for b in host_batches[args.warmup + i]], args.warmup + i)
            loss_host = loss.detach().to("cpu", non_blocking=False)
            d2h = loss_host.numel() * loss_host.element_size()
        end.record()
        ddp.barrier()
        torch.cuda.synchronize()
        ms_e2e = ddp.max_over_ranks(start.elapsed_time(end) / args.steps, dev)
        e2e = {"value": global_rays / (ms_e2e * 1e-3), "unit": "rays/s",
               "h2d_bytes_per_step": nbytes(host_batches[0][0]) * acc, "d2h_bytes_per_step": d2h,
               "ms_per_step": ms_e2e}

    # ---- batches drawn on the device: events resident in HBM, no host batch ------
    producer_line = None
    if not args.no_e2e:
        from deblur_e_nerf_b200.data import EventBatchProducer
        g = torch.Generator().manual_seed(777 + rank)
        pool = synthetic.event_batch(1 << 21, cfg, poses[2], g)          # 2 M queued events, uploaded once
        producer = EventBatchProducer(pool, n_events, it_sample_size=w["S"] if w["pb"] else None,
                                      device=dev, seed=1234, rank=rank)
        for i in range(max(args.warmup, 3)):
            one_step([producer.next_batch() for _ in range(acc)], i)
        ddp.barrier()
        torch.cuda.synchronize()
        mallocs_p = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
        start.record()
        for i in range(args.steps):
            one_step([producer.next_batch() for _ in range(acc)], args.warmup + i)
        end.record()
        ddp.barrier()
        torch.cuda.synchronize()
        mallocs_p = torch.cuda.memory_stats(dev).get("num_device_alloc", 0) - mallocs_p
        ms_prod = ddp.max_over_ranks(start.elapsed_time(end) / args.steps, dev)
        producer_line = {"value": global_rays / (ms_prod * 1e-3), "unit": "rays/s", "ms_per_step": ms_prod,
                         "events_in_hbm": len(producer), "cuda_mallocs": mallocs_p,
                         "mean_samples_per_ray": float(model.logged["train/mean_num_samples_per_ray"]),
                         "note": "batches drawn by data.EventBatchProducer on the device (random event "
                                 "gather + normalised samplers), no host batch, no copy"}

    # ---- weak-scaling figure beside the strong one (every rank takes the FULL batch) --------------
    weak_line = None
    if strong and w["rays_per_call"] is not None and not args.no_e2e:
        n_full = w["rays_per_call"] // w["S"]
        k_weak = min(args.steps, 5)
        full = [[to_dev(host_batch(5000 + i * acc + m, n_full)) for m in range(acc)]
                for i in range(2 + k_weak)]
        for i in range(2):
            one_step(full[i], i)
        ddp.barrier()
        torch.cuda.synchronize()
        start.record()
        for i in range(k_weak):
            one_step(full[2 + i], args.warmup + i)
        end.record()
        ddp.barrier()
        torch.cuda.synchronize()
        ms_weak = ddp.max_over_ranks(start.elapsed_time(end) / k_weak, dev)
        weak_line = {"value": world * rays_per_step(n_full) / (ms_weak * 1e-3), "unit": "rays/s",
                     "ms_per_step": ms_weak, "steps": k_weak, "events_per_step_per_gpu": n_full,
                     "note": "weak scaling: every rank renders the full single-GPU batch"}
        del full

    # ---- kernel pass: every den_b200 entry point bracketed with CUDA events on the launching stream,
    # EAGER steps over the same batches (a replayed graph has no per-launch host hook to bracket; the
    # kernels, their arguments and their durations are the ones of the timed region above) ----------
    stepper.warmup_steps = 1 << 60
    stepper._graph = None
    n_profile = min(args.steps, 4)
    samples_profile = torch.zeros((), dtype=torch.float64, device=dev)
    one_step(dev_batches[args.warmup], args.warmup + 1)
    ops.enable_kernel_timing(None)
    torch.cuda.synchronize()
    start.record()
    for i in range(n_profile):
        one_step(dev_batches[args.warmup + i], args.warmup + 1 + i)      # steps without a grid update
        samples_profile += model.logged["train/mean_num_samples_per_ray"] * rays_per_step(n_events)
    end.record()
    torch.cuda.synchronize()
    profile_ms = start.elapsed_time(end)
    profile_timings = ops.kernel_timings()
    ops.disable_kernel_timing()
    if use_graph:           # rooflines from the kernel pass
        timings = {k: v for k, v in profile_timings.items()}
        roof_samples, roof_steps = float(samples_profile), n_profile
    else:
        roof_samples, roof_steps = samples_seen, args.steps

    # ---- occupancy update, also timed on its own (it is INSIDE the timed steps above) ------------
    occ_ms = None
    if rank == 0:
        torch.cuda.synchronize()
        start.record()
        model.nerf.update_occ_grid(0, model.trajectory.T_wc_position)
        end.record()
        torch.cuda.synchronize()
        occ_ms = start.elapsed_time(end)

    if rank != 0:
        return

    # ---- roofline of the dominant kernel ------------------------------------------------
    peaks = {}
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    # the kernels are timed inside a long step loop: the sustained bf16 figure is the denominator
    tensor_peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0)))
    tensor_burst = float(peaks.get("bf16_tflops", 1590.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    traffic_table, traffic_src = _load_traffic()
    kernel_table = {k: {"launches": v[0], "ms_per_launch": round(v[1] / v[0], 4),
                        "share_of_step": round(v[1] / profile_ms, 4)}
                    for k, v in profile_timings.items() if v[0] > 0}

    def roof(name):
        n_launch, ms_k = timings[name]
        avg_s = ms_k / n_launch * 1e-3
        # every rated kernel runs once per render launch sequence over all of its samples / rays
        samples_per_launch = roof_samples / n_launch
        rays_per_launch = rays_per_step(n_events) * roof_steps / n_launch
        traffic = traffic_table.get(name)
        traffic = traffic * samples_per_launch if traffic else None
        if name in KERNEL_FLOP_PER_SAMPLE:
            achieved = KERNEL_FLOP_PER_SAMPLE[name] * samples_per_launch / avg_s / 1e12
            return {"kernel": name, "bound": "tensor", "achieved": achieved, "peak": tensor_peak,
                    "unit": "TFLOP/s", "frac": achieved / tensor_peak,
                    "frac_of_burst": achieved / tensor_burst, "traffic": traffic,
                    "traffic_source": traffic_src,
                    "peak_source": peak_src + " bf16 sustained (kernel timed inside the step loop)",
                    "avg_launch_ms": avg_s * 1e3, "samples_per_launch": samples_per_launch,
                    "note": "algorithmic FLOP/sample (fp32-equivalent, the 3-pass bf16 split is "
                            "counted once) x samples per launch / mean launch time (CUDA events "
                            "on the launching stream)"}
        nbytes_launch = (KERNEL_BYTES_PER_SAMPLE[name] * samples_per_launch +
                         KERNEL_BYTES_PER_RAY.get(name, 0) * rays_per_launch)
        achieved = nbytes_launch / avg_s / 1e9
        return {"kernel": name, "bound": "hbm", "achieved": achieved, "peak": hbm_peak,
                "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
                "traffic_source": traffic_src,
                "peak_source": peak_src + " HBM copy", "avg_launch_ms": avg_s * 1e3,
                "samples_per_launch": samples_per_launch,
                "note": "algorithmic bytes/sample x samples per launch / mean launch time (CUDA "
                        "events on the launching stream); the 48 MiB table is L2-resident, the "
                        "fraction is of measured HBM copy bandwidth"}

    def roof_adam():
        n_launch, ms_k = timings["den_adam_step"]
        n_param = sum(p.numel() for p in model.parameters() if p.requires_grad and p.dtype == torch.float32)
        avg_s = ms_k / n_launch * 1e-3
        achieved = 28 * n_param / avg_s / 1e9
        return {"kernel": "den_adam_step", "bound": "hbm", "achieved": achieved, "peak": hbm_peak,
                "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None,
                "peak_source": peak_src + " HBM copy", "avg_launch_ms": avg_s * 1e3,
                "params_per_launch": n_param,
                "note": "28 B per fp32 parameter (p, g, m, v read; p, m, v written) / mean duration of the "
                        "step's launches (CUDA events on the launching stream)"}

    rated = [k for k, v in timings.items() if v[0] > 0 and
             (k in KERNEL_BYTES_PER_SAMPLE or k in KERNEL_FLOP_PER_SAMPLE)]
    roofline = roof(max(rated, key=lambda k: timings[k][1])) if rated else None
    other_rooflines = [roof(k) for k in sorted(rated, key=lambda k: -timings[k][1])[1:]]
    if timings.get("den_adam_step", (0, 0))[0] > 0:
        other_rooflines.append(roof_adam())

    cpu = None
    if not args.no_cpu_baseline:
        cpu, _ = cpu_reference_rate(args.workload, args.cpu_events or 64, steps=1, warmup=1)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    config = workload_config(args.workload, w, n_events)
    config.update({
        "events_per_step_global": n_events * world, "accumulate_grad_batches": acc,
        "emulated_share_of_ranks": args.emulate_ranks,
        "trainable": "NeRF + C_p + tau + Omega" if w.get("unfrozen") else "NeRF (synthetic.yaml freezes C_p, tau, Omega)",
        "parallelism": f"dp{world}" + (" (batch sharded over the ranks)" if strong else ""),
        "occupancy": "controlled solid sphere r=0.75 as the start state (SURVEY 8(d)(ii)); the grid update "
                     f"runs inside the timed steps every 16th optimizer step "
                     f"({updates_in(args.warmup, args.steps)} update(s) in the {args.steps} timed steps)",
    })
    line = {
        "metric": "train rays/s (fwd+bwd)", "value": value, "unit": "rays/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong" if (strong or world == 1) else "weak",
        "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config,
        "samples_per_s": global_samples / (ms_step * 1e-3),
        "events_per_s": world * n_events * acc / (ms_step * 1e-3),
        "mean_samples_per_ray": float(model.logged["train/mean_num_samples_per_ray"]),
        "cuda_graph": graph_stats,
        "step_ms": step_ms,
        "hash_gather_gbs": HASH_GATHER_BYTES * global_samples / (ms_step * 1e-3) / 1e9,
        "e2e": e2e, "device_batches": producer_line, "weak_scaling": weak_line,
        "gpu_launches": launches, "cuda_mallocs_in_timed_region": mallocs,
        "clocks": clock_info, "roofline": roofline,
        "other_rooflines": other_rooflines, "kernels": kernel_table,
        "kernels_note": "per-launch times of every C-ABI entry point from a separate 2-step pass with "
                        "all launches bracketed by CUDA events; `roofline` uses the rated kernels' "
                        "events recorded inside the timed region",
        "occ_update_ms": occ_ms, "cpu_baseline": cpu,
    }
    print(json.dumps(line))


def _shutdown():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        try:
            if WORKLOADS[args.workload].get("sweep"):
                run_sweep(args)
            elif WORKLOADS[args.workload].get("raw_events"):
                run_raw_events(args)
            else:
                run_ours(args)
        finally:
            _shutdown()


if __name__ == "__main__":
    main()
