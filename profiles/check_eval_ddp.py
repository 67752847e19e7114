"""`trainer.Trainer.test` under torchrun (NCCL): the views are split round-robin over the ranks, rendered,
all-gathered on the device and post-processed on every rank; the metrics must equal a single-process
evaluation of the same views.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 profiles/check_eval_ddp.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from deblur_e_nerf_b200 import ddp, factory, synthetic, trainer  # noqa: E402


def main():
    rank, local_rank, world = ddp.init_from_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    model, cfg, poses = factory.build_renderer("synthetic", dev, pixel_bandwidth=False, occ_resolution=128, seed=0)
    model.nerf.occupancy_grid._binary = synthetic.solid_sphere_occupancy(128).to(dev)
    ddp.broadcast_parameters(model)
    H, W, B = 120, 160, 5
    kinv = torch.linalg.inv(torch.tensor([[W * 1.2, 0, W / 2], [0, W * 1.2, H / 2], [0, 0, 1.0]])).to(dev)
    ts = torch.linspace(float(poses[2][10]), float(poses[2][-10]), B, dtype=torch.float64).to(dev)
    pos, rot = model.trajectory(ts)
    exposure = [torch.tensor(b % 3 + 1, device=dev) for b in range(B)]
    gain = [torch.tensor(1.0 + 0.25 * b, device=dev) for b in range(B)]
    norm = torch.stack([g * e for g, e in zip(gain, exposure)])
    norm = norm / norm.mean()
    # targets = an affinely (in log space) distorted, offset copy of what the field renders (identical on
    # every rank: same parameters, deterministic eval-mode march), so that the correction is well posed
    model.eval()
    grid = model.image_pixel_positions(H, W, device=dev)
    blank = torch.ones(H, W, device=dev)
    views = [{"img": blank, "T_wc_position": pos[b], "T_wc_orientation": rot[b], "exposure_time": exposure[b],
              "gain": gain[b]} for b in range(B)]
    rendered = [model.evaluation_step(v, kinv, grid)["pred_intensity_img"] for v in views]
    g = torch.Generator(device=dev).manual_seed(0)
    for b in range(B):
        noise = torch.exp(0.02 * torch.randn(H, W, generator=g, device=dev))
        views[b]["img"] = torch.exp(0.8 * rendered[b].log() + 0.3) * noise * norm[b] + 0.02
    hi = float(torch.stack([v["img"] for v in views]).max()) * 1.05
    tr = trainer.Trainer()
    row, pred = tr.test(model, views, kinv, 0.0, hi, black_level_offset=True)
    # the same evaluation without the split: every rank renders every view itself
    model.eval()
    outs = [model.evaluation_step(v, kinv, grid) for v in views]
    want, want_pred = model.evaluation_epoch_end(outs, 0.0, hi, black_level_offset=True)
    want = {k: float(v) for k, v in want.items()}
    err = max(abs(row[k] - want[k]) / max(abs(want[k]), 1e-12) for k in want)
    same = torch.equal(pred, want_pred)
    print(f"rank {rank}/{world}: {row}  max rel diff vs unsplit {err:.2e}  predictions identical: {same}", flush=True)
    assert err < 1e-6 and pred.shape == (B, 1, H, W)
    ddp.barrier()
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
