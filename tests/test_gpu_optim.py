"""GPU parity of the fused optimiser step and densification statistics against torch's own Adam and the reference's
torch expressions (scene/gaussian_model.py:266,515-517; train.py:295-297)."""
import pytest
import torch

import parity_utils as pu

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_fused_adam_matches_torch_adam():
    from partgs_b200.optim import FusedAdam
    gen = torch.Generator().manual_seed(0)
    shapes = [(5000, 3), (5000, 1, 3), (5000, 15, 3), (5000, 1), (5000, 2), (5000, 4), (7,)]
    lrs = [1.6e-4, 2.5e-3, 1.25e-4, 0.05, 0.005, 0.001, 0.0]
    init = [torch.randn(s, generator=gen) for s in shapes]
    pa = [t.clone().to(DEV).requires_grad_(True) for t in init]
    pb = [t.clone().to(DEV).requires_grad_(True) for t in init]
    ref = torch.optim.Adam([{"params": [p], "lr": lr} for p, lr in zip(pa, lrs)], lr=0.0, eps=1e-15, foreach=False)
    ours = FusedAdam([{"params": [p], "lr": lr} for p, lr in zip(pb, lrs)], lr=0.0, eps=1e-15)
    for it in range(6):
        for a, b in zip(pa, pb):
            g = torch.randn(a.shape, generator=gen).to(DEV) * (10.0 ** (it - 3))
            if it == 2:
                g[::3] = 0.0                      # invisible surfels get exactly zero gradient
            a.grad, b.grad = g.clone(), g.clone()
        if it == 4:                               # the reference's schedulers rewrite lr every iteration
            for grp_a, grp_b in zip(ref.param_groups, ours.param_groups):
                grp_a["lr"] *= 0.5
                grp_b["lr"] *= 0.5
        ref.step()
        ours.step()
        for i, (a, b) in enumerate(zip(pa, pb)):
            assert pu.rel_err(b.detach(), a.detach()) <= 2e-6, (it, i)
            sa, sb = ref.state[a], ours.state[b]
            assert pu.rel_err(sb["exp_avg"], sa["exp_avg"]) <= 5e-6 and pu.rel_err(sb["exp_avg_sq"], sa["exp_avg_sq"]) <= 5e-6
            assert int(sb["step"]) == int(sa["step"])
    assert torch.equal(pb[-1].detach().cpu(), init[-1])   # lr 0 leaves the parameter untouched


def test_densification_stats_match_reference_expressions():
    from partgs_b200.optim import densification_stats
    gen = torch.Generator().manual_seed(1)
    P = 10_000
    radii = (torch.randint(0, 40, (P,), generator=gen) * (torch.rand(P, generator=gen) > 0.3)).int().to(DEV)
    grad = torch.randn(P, 3, generator=gen).to(DEV)
    accum0, denom0 = torch.rand(P, 1, generator=gen).to(DEV), torch.randint(0, 5, (P, 1), generator=gen).float().to(DEV)
    maxr0 = torch.randint(0, 30, (P,), generator=gen).float().to(DEV)
    # reference expressions
    vis = radii > 0
    accum_r, denom_r, maxr_r = accum0.clone(), denom0.clone(), maxr0.clone()
    maxr_r[vis] = torch.max(maxr_r[vis], radii[vis].float())
    accum_r[vis] += torch.norm(grad[vis, :2], dim=-1, keepdim=True)
    denom_r[vis] += 1
    accum, denom, maxr = accum0.clone(), denom0.clone(), maxr0.clone()
    densification_stats(radii, grad, accum, denom, maxr)
    assert torch.equal(maxr, maxr_r) and torch.equal(denom, denom_r)
    assert pu.rel_err(accum, accum_r) <= 1e-6
    # without the max_radii2D tensor
    accum2, denom2 = accum0.clone(), denom0.clone()
    densification_stats(radii, grad, accum2, denom2)
    assert torch.equal(denom2, denom_r)
