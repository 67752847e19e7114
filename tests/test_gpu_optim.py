"""den_adam_step / optim.FusedAdam against torch.optim.Adam: parameter groups with their own lr and
weight decay as DeblurENeRF.configure_optimizers builds them (models/deblur_e_nerf.py:1055-1112),
several steps, a tensor whose length is not a multiple of 4, a float64 scalar, a MultiStepLR schedule."""

import pytest
import torch

pytestmark = pytest.mark.gpu


def _make(cuda, seed):
    g = torch.Generator().manual_seed(seed)
    table = torch.nn.Parameter((torch.rand(300_003, generator=g) * 2e-4 - 1e-4).to(cuda))
    w = torch.nn.Parameter(torch.randn(64, 31, generator=g).to(cuda))
    b = torch.nn.Parameter(torch.randn(64, generator=g).to(cuda))
    tau = torch.nn.Parameter(torch.randn(1, generator=g, dtype=torch.float64).to(cuda))
    bk = torch.nn.Parameter(torch.ones(1).to(cuda))
    groups = [dict(params=[tau], lr=5e-2), dict(params=[table, w, b], weight_decay=1e-6),
              dict(params=[bk])]
    return [table, w, b, tau, bk], groups


def test_fused_adam_matches_torch_adam(den_lib, cuda):
    from deblur_e_nerf_b200.optim import FusedAdam
    ours_p, ours_g = _make(cuda, 0)
    ref_p, ref_g = _make(cuda, 0)
    ours = FusedAdam(ours_g, lr=0.01)
    ref = torch.optim.Adam(ref_g, lr=0.01, fused=False, foreach=False)
    sched_o = torch.optim.lr_scheduler.MultiStepLR(ours, milestones=[3, 5], gamma=0.5)
    sched_r = torch.optim.lr_scheduler.MultiStepLR(ref, milestones=[3, 5], gamma=0.5)
    g = torch.Generator().manual_seed(1)
    for step in range(7):
        for po, pr in zip(ours_p, ref_p):
            grad = torch.randn(po.shape, generator=g, dtype=torch.float64).to(po.dtype) * (1e-3 if step % 2 else 1.0)
            po.grad = grad.to(cuda).clone()
            pr.grad = grad.to(cuda).clone()
        ours.step()
        ref.step()
        sched_o.step()
        sched_r.step()
        for po, pr in zip(ours_p, ref_p):
            err = (po.detach().double() - pr.detach().double()).abs().max().item()
            assert err <= 2e-6 * pr.detach().abs().max().item() + 1e-9, (step, po.shape, err)
    so, sr = ours.state[ours_p[0]], ref.state[ref_p[0]]
    assert int(so["step"]) == int(sr["step"]) == 7
    for key in ("exp_avg", "exp_avg_sq"):
        diff = (so[key] - sr[key]).abs().max().item()
        assert diff <= 2e-6 * sr[key].abs().max().item(), (key, diff)


def test_fused_adam_rejects_cpu_parameters(den_lib):
    from deblur_e_nerf_b200.optim import FusedAdam
    p = torch.nn.Parameter(torch.zeros(3))
    p.grad = torch.ones(3)
    with pytest.raises(NotImplementedError):
        FusedAdam([p]).step()
