"""Edge cases of the hot path on the device: no samples at all (the occupancy grid of a freshly
initialised model is empty before its first update; rays that miss the AABB), a single ray, mixed
hit / miss rays inside one batched step, zero-length calls of the per-sample entry points.  The
reference reaches these through nerfacc's empty packed tensors (external/utils.py:83-96 on zero
samples; models/nerf.py:230-259)."""

import pytest
import torch

import _scene

pytestmark = pytest.mark.gpu


def _rays(cfg, n, cuda, seed=0, away=False):
    poses = _scene.synthetic.camera_poses(cfg, n_poses=50)
    traj = _scene.path_ref.LinearTrajectory(*poses)
    g = torch.Generator().manual_seed(seed)
    ts = torch.rand(n, generator=g, dtype=torch.float64) * float(poses[2][-1])
    px = torch.stack([torch.rand(n, generator=g) * cfg["width"],
                      torch.rand(n, generator=g) * cfg["height"]], -1)
    pos, rot = traj(ts)
    kinv = torch.linalg.inv(torch.from_numpy(_scene.synthetic.intrinsics(cfg)))
    o, d = _scene.path_ref.NeRF.pixel_params_to_ray(kinv, px, pos, rot)
    if away:
        d = -d                                   # looking away from the scene: no AABB hit
    return o.float().to(cuda), d.float().to(cuda)


@pytest.mark.parametrize("training", [False, True], ids=["eval", "train"])
def test_empty_occupancy_grid_renders_the_background(den_lib, cuda, training):
    cfg = _scene.scene_config("synthetic", occ_resolution=32, small=True)
    nerf = _scene.build_product_nerf(cfg, cuda)
    nerf.train(training)
    assert not bool(nerf.occupancy_grid._binary.any())          # nothing marked yet
    o, d = _rays(cfg, 257, cuda)
    colour, opacity, depth, mean_samples = nerf(o, d)
    assert mean_samples == 0
    assert colour.shape[0] == 257 and torch.isfinite(colour).all()
    assert float(opacity.detach().abs().max()) == 0.0 and float(depth.detach().abs().max()) == 0.0
    bkgd = nerf.render_bkgd
    if bkgd is not None:
        assert torch.allclose(colour, bkgd.expand_as(colour))
    if training:
        colour.sum().backward()                                  # only the background sees a gradient
        for name, p in nerf.named_parameters():
            if p.grad is not None and "render_bkgd" not in name:
                assert float(p.grad.abs().max()) == 0.0, name


def test_rays_that_miss_the_aabb_and_a_single_ray(den_lib, cuda):
    cfg = _scene.scene_config("synthetic", occ_resolution=32, small=True)
    nerf = _scene.build_product_nerf(cfg, cuda)
    nerf.train()
    nerf.occupancy_grid._binary = torch.ones_like(nerf.occupancy_grid._binary)
    o, d = _rays(cfg, 64, cuda, away=True)
    colour, opacity, depth, mean_samples = nerf(o, d)
    assert mean_samples == 0 and float(opacity.detach().abs().max()) == 0.0
    # a single ray, and a batch in which only some rays hit
    oh, dh = _rays(cfg, 33, cuda, seed=4)
    om, dm = _rays(cfg, 31, cuda, seed=5, away=True)
    jit = torch.full((64,), 0.5, device=cuda)
    c, a, z, m = nerf(torch.cat([oh, om]), torch.cat([dh, dm]), jitter=jit)
    # (a few of the random pixels look past the AABB as well)
    assert float(a[33:].detach().abs().max()) == 0.0 and int((a[:33] > 0).sum()) >= 20
    k = int(torch.nonzero(a[:33] > 0)[0])
    c1, a1, z1, m1 = nerf(oh[k:k + 1].contiguous(), dh[k:k + 1].contiguous(), jitter=jit[:1])
    assert m1 > 0 and c1.shape[0] == 1 and torch.isfinite(c1).all()
    # the same ray alone or inside the mixed batch: same samples, same result
    assert torch.allclose(c[k:k + 1], c1, rtol=1e-5, atol=1e-6)
    assert torch.allclose(a[k:k + 1], a1, rtol=1e-5, atol=1e-6)
    (c.sum() + a.sum()).backward()
    assert all(torch.isfinite(p.grad).all() for p in nerf.parameters() if p.grad is not None)


@pytest.mark.parametrize("pb_on", [False, True], ids=["pb_off", "pb_on"])
def test_training_step_without_samples_is_finite(den_lib, cuda, pb_on):
    """First steps of a run whose occupancy grid is still empty: the step must go through (loss
    finite, batch controller not dividing by zero) — the occupancy update is disabled here on purpose."""
    cfg = _scene.scene_config("synthetic", occ_resolution=32, small=True)
    model, poses = _scene.build_product_renderer(cfg, cuda, it_sample_size=4, pixel_bandwidth=pb_on)
    model.train()
    model.nerf.update_occ_grid = lambda *a, **k: None
    g = torch.Generator().manual_seed(1)
    n = 96
    ev = _scene.synthetic.event_batch(n, cfg, poses[2], g)
    nm = _scene.synthetic.normalized_batch(n, 4, g, pb_on)
    batch = {"event": {k: v.to(cuda) for k, v in ev.items()},
             "normalized": {k: v.to(cuda) for k, v in nm.items()}}
    loss = model.training_step(batch, 0, 0)
    assert torch.isfinite(loss)
    loss.backward()
    assert model.next_train_batch_size is not None and model.next_train_batch_size > 0
    for p in model.parameters():
        if p.grad is not None:
            assert torch.isfinite(p.grad).all()


def test_zero_length_calls_are_accepted(den_lib, cuda):
    from deblur_e_nerf_b200 import ops
    cfg = _scene.scene_config("synthetic", occ_resolution=32, small=True)
    nerf = _scene.build_product_nerf(cfg, cuda)
    field = nerf.radiance_field
    o, d = _rays(cfg, 8, cuda)
    empty_i = torch.empty(0, dtype=torch.int32, device=cuda)
    empty_f = torch.empty(0, device=cuda)
    offsets = torch.zeros(9, dtype=torch.int32, device=cuda)
    desc = field.field_desc()
    u = ops.contract_samples(desc, o, d, empty_i, empty_f, empty_f)
    assert u.shape == (0, 3)
    enc = ops.hashgrid_fwd(field.encoding.desc, u, field.encoding.params)
    assert enc.shape == (0, field.encoding.n_output_dims)
    sig, rgb = ops.mlp_fwd(desc, field.field_params(), enc, o, d, empty_i, empty_f, empty_f, 1)
    assert sig.numel() == 0 and rgb.shape == (0, 1)
    colour, opacity, depth = ops.composite(sig, rgb, empty_f, empty_f, offsets, None)
    assert colour.shape == (8, 1) and float(colour.abs().max()) == 0.0
    assert float(opacity.abs().max()) == 0.0 and float(depth.abs().max()) == 0.0
