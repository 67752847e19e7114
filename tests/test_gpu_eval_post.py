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
    assert (got["pred"].cpu() - want["pred"]).abs().max().item() <= 1e-5 * want["pred"].abs().max().item()
    assert abs(float(got["l1"]) - want["l1"]) <= 1e-5 * want["l1"]
    assert abs(float(got["psnr"]) - want["psnr"]) <= 1e-5 * abs(want["psnr"])


def test_eval_post_refuses_cpu_tensors(den_lib):
    from deblur_e_nerf_b200 import eval_post
    pred, target, exposure, gain = _images(0, 1, 4, 4)
    with pytest.raises(NotImplementedError):
        eval_post.evaluate(pred, target, exposure, gain, 0.0, 1.0)
