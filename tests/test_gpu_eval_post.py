"""GPU parity of the evaluation post-processing (`eval_post.evaluate`, `den_eval_*` kernels, SURVEY.md
§8(f) N4) against oracle/eval_ref.py — which is pinned against the reference's own
`evaluation_epoch_end` for the affine path (tests/test_oracle_vs_reference.py) — on mono and colour image
batches with varying exposure / gain, with and without the black-level-offset refinement."""

import pytest
import torch

from oracle import eval_ref

pytestmark = pytest.mark.gpu


def _images(seed, B, H, W, offset=0.0):
    g = torch.Generator().manual_seed(seed)
    target = torch.rand(B, H, W, generator=g) * 0.9 + 0.05
    exposure = torch.randint(1, 5, (B,), generator=g)
    gain = torch.rand(B, generator=g) + 0.5
    norm = gain * exposure / (gain * exposure).mean()
    scene = target / norm.view(-1, 1, 1)
    pred = (0.7 * scene.pow(1.3)) * torch.exp(0.02 * torch.randn(B, H, W, generator=g))
    return pred.float(), (target + offset).float(), exposure, gain


@pytest.mark.parametrize("black_level_offset", [False, True])
@pytest.mark.parametrize("shape", [(1, 8, 8), (3, 37, 53), (2, 260, 346)])
def test_eval_post_matches_oracle(den_lib, cuda, shape, black_level_offset):
    from deblur_e_nerf_b200 import eval_post
    pred, target, exposure, gain = _images(sum(shape), *shape, offset=0.03 if black_level_offset else 0.0)
    want = eval_ref.evaluate(pred[:, None], target[:, None], exposure, gain, 0.0, 1.1,
                             black_level_offset=black_level_offset)
    got = eval_post.evaluate(pred.to(cuda), target.to(cuda), exposure.to(cuda), gain.to(cuda), 0.0, 1.1,
                             black_level_offset=black_level_offset)
    # the logs are taken in fp32 on both sides (models/deblur_e_nerf.py:733-734): CUDA logf and the CPU
    # library differ by an ulp here and there, so the float64 fit agrees to ~1e-6, not 1e-12
    assert torch.allclose(got["affine"].cpu(), want["affine"], rtol=5e-6, atol=5e-7), \
        (got["affine"], want["affine"])
    if black_level_offset:
        assert torch.allclose(got["correction"].cpu(), want["correction"], rtol=2e-4, atol=2e-6), \
            (got["correction"], want["correction"])
    assert (got["pred"].cpu() - want["pred"]).abs().max().item() <= 1e-5 * want["pred"].abs().max().item()
    assert abs(float(got["l1"]) - want["l1"]) <= 1e-5 * want["l1"]
    assert abs(float(got["psnr"]) - want["psnr"]) <= 1e-5 * abs(want["psnr"])
    if min(shape[1:]) >= 11:
        assert abs(float(got["ssim"]) - want["ssim"]) <= 1e-4, (float(got["ssim"]), want["ssim"])
    else:
        assert got["ssim"] is None and "ssim" not in want    # smaller than the 11 x 11 window
    assert got["pred"].is_cuda and got["l1"].is_cuda          # nothing went through the host


@pytest.mark.parametrize("per_channel", [True, False], ids=["per_channel_scale", "shared_scale"])
@pytest.mark.parametrize("black_level_offset", [False, True])
def test_eval_post_colour_images_match_oracle(den_lib, cuda, per_channel, black_level_offset):
    """Colour images of a Bayer sensor (C = 3): a log-intensity scale per channel, or one shared by the
    channels with an offset each (`correction.per_channel_log_it_scale`, models/deblur_e_nerf.py:753-766;
    the oracle's version is pinned against the reference's evaluation_epoch_end for both)."""
    from deblur_e_nerf_b200 import eval_post
    B, H, W = 3, 41, 57
    chans = [_images(20 + c, B, H, W, offset=0.03 if black_level_offset else 0.0) for c in range(3)]
    exposure, gain = chans[0][2], chans[0][3]
    norm = gain * exposure / (gain * exposure).mean()
    target = torch.stack([ch[1] for ch in chans], dim=1)
    # every channel an affinely (in log space) distorted view of ITS target under the shared exposure
    scene = (target - (0.03 if black_level_offset else 0.0)) / norm.view(-1, 1, 1, 1)
    g = torch.Generator().manual_seed(3)
    pred = torch.stack([(0.5 + 0.2 * c) * scene[:, c].pow(1.0 + 0.1 * c) for c in range(3)], dim=1)
    pred = (pred * torch.exp(0.02 * torch.randn(pred.shape, generator=g))).float()
    want = eval_ref.evaluate(pred, target, exposure, gain, 0.0, 1.1, black_level_offset=black_level_offset,
                             per_channel_scale=per_channel)
    got = eval_post.evaluate(pred.to(cuda), target.to(cuda), exposure.to(cuda), gain.to(cuda), 0.0, 1.1,
                             black_level_offset=black_level_offset, per_channel_scale=per_channel)
    assert torch.allclose(got["affine"].cpu(), want["affine"], rtol=5e-6, atol=5e-7), (got["affine"], want["affine"])
    if not per_channel:
        assert float(got["affine"][:, 0].max() - got["affine"][:, 0].min()) == 0.0
    if black_level_offset:
        assert torch.allclose(got["correction"].cpu(), want["correction"], rtol=2e-4, atol=2e-6)
        if not per_channel:       # a shared log-intensity scale comes with ONE gamma (models/deblur_e_nerf.py:185-197)
            assert float(got["correction"][:, 1].max() - got["correction"][:, 1].min()) == 0.0
    assert (got["pred"].cpu() - want["pred"]).abs().max().item() <= 1e-5 * want["pred"].abs().max().item()
    assert abs(float(got["l1"]) - want["l1"]) <= 1e-5 * want["l1"]
    assert abs(float(got["psnr"]) - want["psnr"]) <= 1e-5 * abs(want["psnr"])
    assert abs(float(got["ssim"]) - want["ssim"]) <= 1e-4, (float(got["ssim"]), want["ssim"])


@pytest.mark.parametrize("shape", [(1, 1, 11, 11), (2, 1, 12, 43), (3, 3, 37, 53), (2, 1, 260, 346), (1, 3, 480, 640)])
def test_ssim_matches_oracle(den_lib, cuda, shape):
    """den_eval_ssim against oracle/eval_ref.ssim (torchmetrics 0.6.2 functional.ssim restated; fp32 and
    fp64 evaluations of it) on noisy and on smooth image pairs — flat windows are where the fp32
    E[x^2] - mu^2 cancellation shows (~2e-5 on a single window) — per image and as the mean."""
    from deblur_e_nerf_b200 import eval_post
    g = torch.Generator().manual_seed(sum(shape))
    t = torch.rand(shape, generator=g) * 0.8 + 0.1
    noisy = (t + 0.05 * torch.randn(shape, generator=g)).clamp(0.01, 1.0)
    smooth_t = torch.nn.functional.avg_pool2d(torch.nn.functional.pad(t, (4, 4, 4, 4), mode="reflect"), 9, 1)
    smooth_p = smooth_t * 1.02 + 0.003
    for pred, target, data_range in ((noisy, t, 1.0), (smooth_p, smooth_t, 1.1), (t, t, 0.9)):
        got = eval_post.ssim(pred.to(cuda), target.to(cuda), data_range)
        assert got.shape == (shape[0],) and got.dtype == torch.float64 and got.is_cuda
        for b in range(shape[0]):
            want32 = float(eval_ref.ssim(pred[b:b + 1], target[b:b + 1], data_range))
            want64 = float(eval_ref.ssim(pred[b:b + 1].double(), target[b:b + 1].double(), data_range))
            assert abs(float(got[b]) - want32) <= 1e-4 and abs(float(got[b]) - want64) <= 1e-4, \
                (shape, b, float(got[b]), want32, want64)
        assert abs(float(got.mean()) - float(eval_ref.ssim(pred, target, data_range))) <= 1e-4
    assert abs(float(eval_post.ssim(t.to(cuda), t.to(cuda), 0.9).mean()) - 1.0) <= 1e-6


def test_ssim_window_arguments(den_lib, cuda):
    """Other odd windows follow the same definition; an image smaller than the window, an even window and
    a window beyond the kernel's limit are refused with the library's error (no silent result)."""
    from deblur_e_nerf_b200 import eval_post
    g = torch.Generator().manual_seed(9)
    t = torch.rand(2, 1, 33, 47, generator=g) * 0.8 + 0.1
    p = (t + 0.03 * torch.randn(t.shape, generator=g)).clamp(0.01, 1.0)
    for k, sigma in ((7, 1.0), (15, 2.5), (1, 1.5)):
        got = eval_post.ssim(p.to(cuda), t.to(cuda), 1.0, kernel_size=k, sigma=sigma)
        want = eval_ref.ssim(p.double(), t.double(), 1.0, kernel_size=k, sigma=sigma) if k > 1 else None
        if want is not None:
            assert abs(float(got.mean()) - float(want)) <= 1e-4, (k, float(got.mean()), float(want))
        else:
            assert torch.isfinite(got).all()
    for bad in (dict(kernel_size=10), dict(kernel_size=17), dict(kernel_size=11, sigma=0.0)):
        with pytest.raises(RuntimeError):
            eval_post.ssim(p.to(cuda), t.to(cuda), 1.0, **bad)
    with pytest.raises(RuntimeError):
        eval_post.ssim(p[..., :10, :].to(cuda), t[..., :10, :].to(cuda), 1.0)


def test_eval_post_refuses_cpu_tensors(den_lib):
    from deblur_e_nerf_b200 import eval_post
    pred, target, exposure, gain = _images(0, 1, 4, 4)
    with pytest.raises(NotImplementedError):
        eval_post.evaluate(pred, target, exposure, gain, 0.0, 1.0)
