"""CPU: the optimizer's parameter groups mirror DeblurENeRF.configure_optimizers
(models/deblur_e_nerf.py:1055-1090) group for group, so an Adam state dict written by the reference's
grouping (a Lightning checkpoint's `optimizer_states[0]`) loads and maps onto the same parameters."""

import torch

from deblur_e_nerf_b200 import factory


def _reference_groups(model, pb):
    """The reference's grouping, restated literally from models/deblur_e_nerf.py:1055-1090."""
    rp = list(model.refractory_period.parameters())
    mlp = [p for n, p in model.named_parameters() if n.startswith("nerf.radiance_field.mlp")]
    groups = [{"params": rp, "lr": float(model.refractory_period.max_refractory_period) * 50},
              {"params": mlp, "weight_decay": 1e-6}]
    lrs = {"contrast_threshold": {"p2n_contrast_threshold_ratio": 0.1, "mean_contrast_threshold": 0.1}}
    if pb:
        lrs["pixel_bandwidth"] = {k: 0.01 for k in ("tau_mil_it_eff_prod", "A_amp_inv", "A_loop_inv",
                                                    "tau_out", "tau_sf", "tau_diff")}
    for comp, table in lrs.items():
        module = getattr(model, comp)
        groups += [{"params": [getattr(module.parametrizations, name).original], "lr": lr}
                   for name, lr in table.items()]
    collated = {id(p) for g in groups for p in g["params"]}
    groups.append({"params": [p for p in model.parameters() if id(p) not in collated]})
    return groups


def _build(pb):
    model, _, _ = factory.build_renderer("synthetic", "cpu", pixel_bandwidth=pb, small=True,
                                         occ_resolution=16, n_poses=20)
    return model


def test_groups_match_the_reference_layout_with_frozen_parameters():
    for pb in (True, False):
        model = factory.freeze_like_synthetic_yaml(_build(pb))      # frozen params stay in their groups
        ours = factory.optimizer_param_groups(model)
        ref = _reference_groups(model, pb)
        assert len(ours) == len(ref) == (2 + 2 + (6 if pb else 0) + 1)
        for a, b in zip(ours, ref):
            assert [id(p) for p in a["params"]] == [id(p) for p in b["params"]]
            assert a.get("lr") == b.get("lr") and a.get("weight_decay") == b.get("weight_decay")
        assert sum(len(g["params"]) for g in ours) == len(list(model.parameters()))


def test_reference_adam_state_dict_round_trips():
    torch.manual_seed(0)
    src = _build(True)
    ref_opt = torch.optim.Adam(_reference_groups(src, True), lr=0.01)
    for p in src.parameters():
        p.grad = torch.randn_like(p) * 1e-3
    ref_opt.step()
    ref_opt.step()
    state = ref_opt.state_dict()

    dst = _build(True)
    dst.load_state_dict(src.state_dict())
    opt = factory.configure_optimizer(dst, fused=False)
    opt.load_state_dict(state)                       # same number / sizes of groups
    assert [len(g["params"]) for g in opt.state_dict()["param_groups"]] == \
        [len(g["params"]) for g in state["param_groups"]]
    by_name_src = dict(src.named_parameters())
    for name, p in dst.named_parameters():
        got = opt.state[p]
        want = ref_opt.state[by_name_src[name]]
        assert int(got["step"]) == 2
        assert torch.equal(got["exp_avg"], want["exp_avg"]), name
        assert torch.equal(got["exp_avg_sq"], want["exp_avg_sq"]), name
    lrs = [g["lr"] for g in opt.param_groups]
    assert lrs[0] == float(dst.refractory_period.max_refractory_period) * 50 and lrs[2] == 0.1
    assert opt.param_groups[1]["weight_decay"] == 1e-6


def test_parameters_unfrozen_later_are_optimised():
    model = factory.freeze_like_synthetic_yaml(_build(False))
    opt = factory.configure_optimizer(model, fused=False)
    p = model.contrast_threshold.parametrizations.mean_contrast_threshold.original
    before = p.detach().clone()
    p.requires_grad_(True)
    p.grad = torch.ones_like(p)
    opt.step()
    assert not torch.equal(p.detach(), before)
