"""GPU parity of the evaluation loop (`run.py test`, SURVEY.md §3.3 + §8(f) N4): `trainer.Trainer.test`
= `EventRenderer.evaluation_step` per posed view (models/deblur_e_nerf.py:604-652) +
`evaluation_epoch_end` on the device (:674-969, `eval_post`), against the oracle's renderer
(oracle/path_ref.py) followed by oracle/eval_ref.py on the same views."""

import pytest
import torch

import _scene
from oracle import eval_ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("black_level_offset", [False, True])
def test_evaluation_loop_matches_oracle(den_lib, cuda, black_level_offset):
    from deblur_e_nerf_b200 import trainer
    cfg = _scene.scene_config("synthetic", occ_resolution=32, small=True)
    ora, poses = _scene.build_oracle_renderer(cfg, pixel_bandwidth=False, n_poses=50)
    prod, _ = _scene.build_product_renderer(cfg, cuda, pixel_bandwidth=False, n_poses=50)
    _scene.copy_params(ora.nerf, prod.nerf)
    ora.nerf.train()
    torch.manual_seed(3)
    ora.nerf.update_occ_grid(0, poses[0])
    prod.nerf.occupancy_grid._binary = ora.nerf.occupancy_grid.binary.to(cuda)
    prod.nerf.occupancy_grid.occs.copy_(ora.nerf.occupancy_grid.occs)
    ora.nerf.eval()

    H, W, B = 24, 32, 3
    # a coarse pixel grid over the whole sensor (the intrinsics are the 346 x 260 camera's)
    grid = prod.image_pixel_positions(H, W)
    grid = grid * torch.tensor([cfg["width"] / W, cfg["height"] / H])
    traj = _scene.path_ref.LinearTrajectory(*poses)
    pos, rot = traj(torch.tensor([3.0e6, 21.5e6, 44.25e6], dtype=torch.float64))
    g = torch.Generator().manual_seed(11)
    exposure = torch.tensor([1, 2, 3])
    gain = torch.tensor([1.0, 0.75, 1.5])
    norm = gain * exposure / (gain * exposure).mean()
    views, pred_o = [], []
    with torch.no_grad():
        for b in range(B):
            rad, _, _, _ = ora.render_pixels(grid.view(-1, 2), pos[b].expand(H * W, -1),
                                             rot[b].expand(H * W, -1, -1))
            pred_o.append(rad.view(H, W))
    pred_o = torch.stack(pred_o)
    offset = 0.02 if black_level_offset else 0.0
    scene = torch.exp(0.8 * pred_o.log() + 0.3) * torch.exp(0.02 * torch.randn(pred_o.shape, generator=g))
    target = (scene * norm.view(-1, 1, 1) + offset).float()
    lo, hi = 0.0, float(target.max()) * 1.05
    for b in range(B):
        # the DataLoader's leading dim of 1 on every entry, as the reference's evaluation_step receives it
        views.append({"img": target[b][None], "T_wc_position": pos[b][None].float(),
                      "T_wc_orientation": rot[b][None].float(), "exposure_time": exposure[b][None],
                      "gain": gain[b][None], "sample_id": torch.tensor([[b]])})
    want = eval_ref.evaluate(pred_o[:, None], target[:, None], exposure, gain, lo, hi,
                             black_level_offset=black_level_offset)

    prod.train()
    logged = []
    tr = trainer.Trainer(log_fn=lambda step, row: logged.append(row))
    kinv = prod.train_intrinsics_inv
    row, pred = tr.test(prod, [_scene.to_device(v, cuda) for v in views], kinv, lo, hi,
                        img_pixel_pos=grid.to(cuda), black_level_offset=black_level_offset)
    assert prod.training and logged == [row] and set(row) == {"test/l1", "test/psnr", "test/ssim"}
    assert pred.shape == (B, 1, H, W) and pred.is_cuda
    # the rendered views agree with the oracle's to the render tolerance; the post-processing is a smooth
    # function of them
    assert (pred.cpu() - want["pred"]).abs().max().item() <= 2e-3 * want["pred"].abs().max().item()
    assert abs(row["test/l1"] - want["l1"]) <= 2e-3 * want["l1"] + 1e-6
    assert abs(row["test/psnr"] - want["psnr"]) <= 2e-2
    assert abs(row["test/ssim"] - want["ssim"]) <= 1e-3
    # validate() is the same loop under the "val" stage; the pixel grid defaults to the image's own
    small = [{k: (v[..., :13, :17] if k == "img" else v) for k, v in view.items()} for view in views[:2]]
    val_row, val_pred = tr.validate(prod, [_scene.to_device(v, cuda) for v in small], kinv, lo, hi)
    assert set(val_row) == {"val/l1", "val/psnr", "val/ssim"} and val_pred.shape == (2, 1, 13, 17)


def test_evaluation_from_a_dataset_directory(den_lib, cuda, tmp_path):
    """`run.py test` end to end on the device: dataset directory -> `views.PosedViews` -> `Trainer.test`
    (every view rendered by the CUDA path in eval mode, post-processing and metrics by `eval_post`) against the
    same views pushed through `evaluation_step` / `evaluation_epoch_end` one by one."""
    import _dataset
    from deblur_e_nerf_b200 import synthetic, trainer, views
    cfg = _scene.scene_config("synthetic", occ_resolution=32, small=True)
    _dataset.write(str(tmp_path), dict(synthetic.CONFIGS["synthetic"]), n_events=8, n_views=3, size=(24, 32), channels=3)
    posed = views.PosedViews(str(tmp_path), "test", device=cuda)
    assert posed.img.is_cuda and posed.img.shape == (3, 24, 32)
    prod, _ = _scene.build_product_renderer(cfg, cuda, pixel_bandwidth=False, n_poses=50)
    prod.nerf.occupancy_grid._binary = synthetic.solid_sphere_occupancy(32).to(cuda)
    args = posed.test_arguments()
    row, pred = trainer.Trainer().test(prod, posed, black_level_offset=False, **args)
    assert set(row) == {"test/l1", "test/psnr", "test/ssim"} and pred.shape == (3, 1, 24, 32) and pred.is_cuda
    assert all(v == v for v in row.values()) and row["test/l1"] > 0 and -1.0 <= row["test/ssim"] <= 1.0
    prod.eval()
    grid = prod.image_pixel_positions(24, 32, device=cuda)
    outs = [prod.evaluation_step(view, args["intrinsics_inv"], grid) for view in posed]
    want, want_pred = prod.evaluation_epoch_end(outs, args["min_normalized_pixel_value"],
                                                args["max_normalized_pixel_value"], black_level_offset=False)
    assert torch.equal(pred, want_pred) and all(abs(row[k] - float(want[k])) <= 1e-12 for k in row)
    # the random views share nothing with the renders: the affine fit flattens the prediction, errors stay finite
    assert float(pred.min()) > 0


@pytest.mark.parametrize("per_channel", [True, False], ids=["per_channel_scale", "shared_scale"])
def test_evaluation_loop_colour_sensor_matches_oracle(den_lib, cuda, per_channel):
    """A sensor behind a Bayer filter: three radiance channels rendered per view (channels first, models/
    deblur_e_nerf.py:1200-1201), the log-intensity scale per channel or shared (then ONE gamma in the offset-
    gamma refinement, :185-197), against the oracle's renderer + eval_ref."""
    from deblur_e_nerf_b200 import trainer
    cfg = _scene.scene_config("synthetic", occ_resolution=32, small=True)
    cfg["radiance_dim"] = 3
    ora, poses = _scene.build_oracle_renderer(cfg, pixel_bandwidth=False, n_poses=50)
    prod, _ = _scene.build_product_renderer(cfg, cuda, pixel_bandwidth=False, n_poses=50)
    _scene.copy_params(ora.nerf, prod.nerf)
    ora.nerf.train()
    torch.manual_seed(3)
    ora.nerf.update_occ_grid(0, poses[0])
    prod.nerf.occupancy_grid._binary = ora.nerf.occupancy_grid.binary.to(cuda)
    prod.nerf.occupancy_grid.occs.copy_(ora.nerf.occupancy_grid.occs)
    ora.nerf.eval()
    H, W, B = 20, 28, 3
    grid = prod.image_pixel_positions(H, W) * torch.tensor([cfg["width"] / W, cfg["height"] / H])
    traj = _scene.path_ref.LinearTrajectory(*poses)
    pos, rot = traj(torch.tensor([5.0e6, 20.5e6, 41.25e6], dtype=torch.float64))
    with torch.no_grad():
        pred_o = torch.stack([ora.render_pixels(grid.view(-1, 2), pos[b].expand(H * W, -1),
                                                rot[b].expand(H * W, -1, -1))[0].view(H, W, 3).permute(2, 0, 1)
                              for b in range(B)])
    g = torch.Generator().manual_seed(13)
    exposure, gain = torch.tensor([2, 1, 3]), torch.tensor([1.0, 1.25, 0.8])
    norm = gain * exposure / (gain * exposure).mean()
    gammas = torch.tensor([0.8, 0.9, 0.85]).view(1, 3, 1, 1)
    scene = torch.exp(gammas * pred_o.log() + torch.tensor([0.3, 0.1, 0.2]).view(1, 3, 1, 1))
    target = (scene * torch.exp(0.02 * torch.randn(scene.shape, generator=g)) * norm.view(-1, 1, 1, 1) + 0.02).float()
    lo, hi = 0.0, float(target.max()) * 1.05
    views = [{"img": target[b], "T_wc_position": pos[b].float(), "T_wc_orientation": rot[b].float(),
              "exposure_time": exposure[b], "gain": gain[b]} for b in range(B)]
    want = eval_ref.evaluate(pred_o, target, exposure, gain, lo, hi, black_level_offset=True,
                             per_channel_scale=per_channel)
    row, pred = trainer.Trainer().test(prod, [_scene.to_device(v, cuda) for v in views], prod.train_intrinsics_inv,
                                       lo, hi, img_pixel_pos=grid.to(cuda), black_level_offset=True,
                                       per_channel_log_it_scale=per_channel)
    assert pred.shape == (B, 3, H, W)
    assert (pred.cpu() - want["pred"]).abs().max().item() <= 3e-3 * want["pred"].abs().max().item()
    assert abs(row["test/l1"] - want["l1"]) <= 3e-3 * want["l1"] + 1e-6
    assert abs(row["test/psnr"] - want["psnr"]) <= 3e-2
    assert abs(row["test/ssim"] - want["ssim"]) <= 1e-3
