"""Known-answer tests pinning the third-party restatements (oracle/nerfacc_ref.py,
oracle/tcnn_ref.py, oracle/roma_ref.py) against closed forms — the reference ships no
golden vectors and nerfacc / tiny-cuda-nn are absent."""

import math

import numpy as np
import torch

from oracle import nerfacc_ref as nacc
from oracle import roma_ref, tcnn_ref


def test_level_table_matches_survey():
    scales, res, sizes, offsets, total = tcnn_ref.grid_level_table(16, 16, 1.4472692012786865, 19)
    assert res.tolist() == [16, 24, 34, 49, 71, 102, 148, 213, 308, 446, 646, 934, 1352, 1956,
                            2831, 4096]
    assert sizes[:5].tolist() == [4096, 13824, 39304, 117656, 357912]
    assert all(s == 524288 for s in sizes[5:])
    assert total == 6299960 and total * 2 == 12599920


def test_hashgrid_lattice_known_answer():
    cfg = dict(n_levels=4, n_features_per_level=2, log2_hashmap_size=14, base_resolution=16,
               per_level_scale=1.4472692012786865)
    enc = tcnn_ref.Encoding(3, cfg)
    with torch.no_grad():
        table = enc.params.view(-1, 2)
        for lvl in range(4):
            off, size = int(enc.offsets[lvl]), int(enc.sizes[lvl])
            table[off:off + size, 0] = torch.arange(size, dtype=torch.float32)
            table[off:off + size, 1] = float(lvl)
    i, j, k = 3, 5, 7
    x = torch.tensor([[(i + 0.5) / 15, (j + 0.5) / 15, (k + 0.5) / 15]])
    out = enc(x)
    expect = (i + 1) + (j + 1) * 16 + (k + 1) * 256          # dense level 0, x fastest
    assert abs(out[0, 0].item() - expect) < 1e-3 * expect
    assert out[0, 1].item() == 0.0
    # hashed level (3: res 49 -> 117656 > 2^14): weights still sum to one
    assert abs(out[0, 7].item() - 3.0) < 1e-5


def test_hash_function_primes():
    scales, res, sizes = [np.float32(1000.0)], [1001], [1 << 14]
    idx, w = tcnn_ref.hashgrid_indices_weights(torch.tensor([[0.0105, 0.0205, 0.0305]]),
                                               scales, res, sizes)
    # pos = 1000 x + .5 -> cells (11, 21, 31); corner 0 hash = 11 ^ 21*2654435761 ^ 31*805459861
    expect = (11 ^ ((21 * 2654435761) & 0xFFFFFFFF) ^ ((31 * 805459861) & 0xFFFFFFFF)) % (1 << 14)
    assert idx[0, 0, 0].item() == expect
    assert abs(w.sum().item() - 1.0) < 1e-6


def test_ray_aabb_known_answer():
    o = torch.tensor([[0.0, 0.0, -4.0], [0.0, 0.0, -4.0], [5.0, 5.0, 5.0]])
    d = torch.tensor([[0.0, 0.0, 1.0], [0.0, 1.0, 0.0], [1.0, 0.0, 0.0]])
    tmin, tmax = nacc.ray_aabb_intersect(o, d, torch.tensor([-1.0, -1, -1, 1, 1, 1]))
    assert tmin[0].item() == 3.0 and tmax[0].item() == 5.0
    assert tmin[1].item() == 1e10 and tmax[2].item() == 1e10


def _one_cell_grid(res, cell):
    grid = nacc.OccupancyGrid([0.0, 0, 0, 1, 1, 1], res, nacc.ContractionType.AABB)
    grid._binary[cell] = True
    return grid


def test_march_hand_traceable_single_cell():
    """4^3 grid, one occupied cell (1,1,2); a ray along +z through its centre with step 0.1:
    samples are exactly those whose mid-point lies in z in [0.5, 0.75]."""
    grid = _one_cell_grid(4, (1, 1, 2))
    o = torch.tensor([[0.375, 0.375, -1.0]])
    d = torch.tensor([[0.0, 0.0, 1.0]])
    ri, ts, te = nacc.ray_marching(o, d, scene_aabb=torch.tensor([0.0, 0, 0, 1, 1, 1]), grid=grid,
                                   render_step_size=0.1, early_stop_eps=0.0)
    mids = ((ts + te) / 2).flatten() - 1.0
    assert ri.tolist() == [0] * len(mids) and len(mids) in (2, 3)
    assert torch.all((mids >= 0.5) & (mids <= 0.75))
    assert torch.allclose(te - ts, torch.full_like(ts, 0.1), atol=1e-6)
    # a ray that misses the occupied cell emits nothing
    ri2, _, _ = nacc.ray_marching(torch.tensor([[0.875, 0.875, -1.0]]), d,
                                  scene_aabb=torch.tensor([0.0, 0, 0, 1, 1, 1]), grid=grid,
                                  render_step_size=0.1, early_stop_eps=0.0)
    assert ri2.numel() == 0


def test_march_full_grid_is_uniform_stepping():
    grid = nacc.OccupancyGrid([0.0, 0, 0, 1, 1, 1], 2, nacc.ContractionType.AABB)
    grid._binary[:] = True
    o = torch.tensor([[0.5, 0.5, -0.5]])
    d = torch.tensor([[0.0, 0.0, 1.0]])
    ri, ts, te = nacc.ray_marching(o, d, scene_aabb=torch.tensor([0.0, 0, 0, 1, 1, 1]), grid=grid,
                                   render_step_size=0.125, early_stop_eps=0.0)
    assert len(ts) == 8                       # t in [0.5, 1.5], 8 steps of 1/8 (exact in fp32)
    assert torch.equal(ts.flatten(), 0.5 + 0.125 * torch.arange(8))


def test_sphere_contraction_roundtrip():
    roi = torch.tensor([0.2, -0.4, 0.0, 3.7, 3.7, 1.8])
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1000, 3, generator=g) * 6
    u = nacc.contract(x, roi, nacc.ContractionType.UN_BOUNDED_SPHERE)
    assert torch.all((u > 0) & (u < 1))
    back = nacc.contract_inv(u, roi, nacc.ContractionType.UN_BOUNDED_SPHERE)
    assert torch.allclose(back, x, rtol=2e-3, atol=2e-3)
    inside = nacc.contract((roi[:3] + roi[3:]) / 2, roi, nacc.ContractionType.UN_BOUNDED_SPHERE)
    assert torch.allclose(inside, torch.full((3,), 0.5))


def test_constant_density_ray_analytic():
    n, sigma, dt = 50, 3.0, 0.02
    ts = (torch.arange(n) * dt)[:, None]
    te = ts + dt
    ri = torch.zeros(n, dtype=torch.int32)
    w = nacc.render_weight_from_density(ts, te, torch.full((n, 1), sigma), ray_indices=ri, n_rays=1)
    i = torch.arange(n, dtype=torch.float64)
    expect = torch.exp(-sigma * dt * i) * (1 - math.exp(-sigma * dt))
    assert torch.allclose(w.flatten().double(), expect, rtol=1e-5)
    opacity = nacc.accumulate_along_rays(w, ri, None, 1)
    assert abs(opacity.item() - (1 - math.exp(-sigma * dt * n))) < 1e-5
    alphas = torch.full((n, 1), 1 - math.exp(-sigma * dt))
    w2 = nacc.render_weight_from_alpha(alphas, ray_indices=ri, n_rays=1)
    assert torch.allclose(w2, w, rtol=1e-4)
    vis = nacc.render_visibility(alphas, ray_indices=ri, n_rays=1, early_stop_eps=0.5)
    # T_i = exp(-sigma dt i) >= 0.5  <=>  i <= ln2 / (sigma dt) = 11.55
    assert vis.tolist() == [True] * 12 + [False] * (n - 12)


def test_empty_and_ragged_inputs():
    ri = torch.tensor([0, 0, 3, 3, 3], dtype=torch.int32)          # rays 1, 2, 4 empty
    w = torch.ones(5, 1)
    out = nacc.accumulate_along_rays(w, ri, None, 5)
    assert out.flatten().tolist() == [2, 0, 0, 3, 0]
    empty = nacc.accumulate_along_rays(torch.zeros(0, 1), torch.zeros(0, dtype=torch.int32), None, 3)
    assert empty.shape == (3, 1) and empty.abs().sum() == 0
    assert nacc.render_weight_from_density(torch.zeros(0, 1), torch.zeros(0, 1), torch.zeros(0, 1),
                                           ray_indices=torch.zeros(0, dtype=torch.int32),
                                           n_rays=2).shape == (0, 1)


def test_roma_restatement():
    g = torch.Generator().manual_seed(0)
    rv = torch.randn(20, 3, generator=g)
    q = roma_ref.rotvec_to_unitquat(rv)
    assert torch.allclose(q.norm(dim=-1), torch.ones(20), atol=1e-6)
    R = roma_ref.unitquat_to_rotmat(q)
    assert torch.allclose(R @ R.transpose(-1, -2), torch.eye(3).expand(20, 3, 3), atol=1e-5)
    ident = roma_ref.quat_product(q, roma_ref.quat_conjugation(q))
    assert torch.allclose(ident, torch.tensor([0.0, 0, 0, 1]).expand(20, 4), atol=1e-6)
    # 90 degrees about z maps x to y
    qz = roma_ref.rotvec_to_unitquat(torch.tensor([[0.0, 0.0, math.pi / 2]]))
    assert torch.allclose(roma_ref.unitquat_to_rotmat(qz)[0] @ torch.tensor([1.0, 0, 0]),
                          torch.tensor([0.0, 1.0, 0.0]), atol=1e-6)


def test_ssim_restatement_known_answers():
    """oracle/eval_ref.ssim (torchmetrics 0.6.2 functional.ssim, restated: parity unpinned) against closed
    forms and an independent float64 window sum written from the definition."""
    from oracle import eval_ref
    g = torch.Generator().manual_seed(5)
    t = torch.rand(2, 3, 29, 40, generator=g, dtype=torch.float64) * 0.8 + 0.1
    # identical images: every index is 1
    assert abs(float(eval_ref.ssim(t, t, 1.0)) - 1.0) < 1e-12
    # constant images a, b: variances and covariance vanish, the index is the luminance term alone
    a, b, rng = 0.3, 0.5, 0.9
    want = (2 * a * b + (0.01 * rng) ** 2) / (a * a + b * b + (0.01 * rng) ** 2)
    got = eval_ref.ssim(torch.full((1, 1, 16, 20), a, dtype=torch.float64),
                        torch.full((1, 1, 16, 20), b, dtype=torch.float64), rng)
    assert abs(float(got) - want) < 1e-9
    # an 11 x 11 image has exactly one pixel whose window fits
    p = (t + 0.05 * torch.randn(t.shape, generator=g, dtype=torch.float64)).clamp(0.01, 1.0)
    d = np.arange(-5, 6, dtype=np.float64)
    w = np.exp(-(d / 1.5) ** 2 / 2)
    w = np.outer(w / w.sum(), w / w.sum())

    def index(pw, tw, rng):
        c1, c2 = (0.01 * rng) ** 2, (0.03 * rng) ** 2
        mp, mt = (w * pw).sum(), (w * tw).sum()
        vp, vt, cov = (w * pw * pw).sum() - mp * mp, (w * tw * tw).sum() - mt * mt, (w * pw * tw).sum() - mp * mt
        return (2 * mp * mt + c1) * (2 * cov + c2) / ((mp * mp + mt * mt + c1) * (vp + vt + c2))

    one = eval_ref.ssim(p[:1, :1, :11, :11], t[:1, :1, :11, :11], 1.0)
    assert abs(float(one) - index(p[0, 0, :11, :11].numpy(), t[0, 0, :11, :11].numpy(), 1.0)) < 1e-12
    # the mean runs over the pixels whose window lies inside the image, over channels and images
    pn, tn = p.numpy(), t.numpy()
    vals = [index(pn[i, c, y:y + 11, x:x + 11], tn[i, c, y:y + 11, x:x + 11], 0.9)
            for i in range(2) for c in range(3) for y in range(29 - 10) for x in range(40 - 10)]
    assert abs(float(eval_ref.ssim(p, t, 0.9)) - float(np.mean(vals))) < 1e-12
    # fp32 (what Metric.compute runs in) stays within 1e-5 of it
    assert abs(float(eval_ref.ssim(p.float(), t.float(), 0.9)) - float(np.mean(vals))) < 1e-5
