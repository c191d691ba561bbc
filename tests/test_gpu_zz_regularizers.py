"""GPU parity of the fused per-pixel regularisers (csrc/regularizers.cu, partgs_b200.losses.geometric_regularizers)
against the torch restatement of train.py:234-251 (oracle/loss_oracle.py).

STATUS: written without GPU access; first verified on the CPU emulator (tests/test_emu_ops.py,
tests/test_emu_zz_mirror.py), then green on a B200 on its first hardware run (profiles/r1_gpu_pytest_new_kernels.log)."""
import pytest
import torch

from oracle import loss_oracle

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
DEV = "cuda"


@pytest.mark.parametrize("H,W,with_mask", [(1200, 1600, True), (300, 400, True), (77, 123, False)])
def test_fused_regularizers_match_the_reference_expressions(H, W, with_mask):
    from partgs_b200.losses import geometric_regularizers
    gen = torch.Generator().manual_seed(H)
    allmap = torch.rand(7, H, W, generator=gen).to(DEV)
    allmap[1, 0, :4] = torch.tensor([0.0, 1.0, 1e-7, 0.5], device=DEV)
    mask = (torch.rand(H, W, generator=gen) > 0.4).float().to(DEV) if with_mask else None
    rn0 = torch.randn(3, H, W, generator=gen).to(DEV)
    sn0 = torch.randn(3, H, W, generator=gen).to(DEV)
    lam = (0.1, 0.05, 1000.0)
    out = []
    for fused in (False, True):
        am = allmap.clone().requires_grad_(True)
        rn, sn = rn0.clone().requires_grad_(True), sn0.clone().requires_grad_(True)
        pkg = {"rend_alpha": am[1:2], "rend_dist": am[6:7], "rend_normal": rn, "surf_normal": sn}
        if fused:
            loss, parts = geometric_regularizers(pkg, mask, *lam, return_parts=True)
        else:
            loss, parts = loss_oracle.geometric_regularizers(pkg["rend_alpha"], mask, pkg["rend_dist"], rn, sn, *lam)
        (loss * 0.7).backward()
        out.append((loss.detach(), parts, am.grad, rn.grad, sn.grad))
    (l0, p0, a0, r0, s0), (l1, p1, a1, r1, s1) = out
    # tolerance: 5e-6 relative on the scalar (float vs double accumulation), 2e-6 on the pointwise gradients
    assert abs(float(l1) - float(l0)) <= 5e-6 * abs(float(l0))
    for i in range(3):
        if p0[i] is not None:
            assert abs(float(p1[i]) - float(p0[i])) <= 5e-6 * abs(float(p0[i])) + 1e-12
    for got, want in ((a1, a0), (r1, r0), (s1, s0)):
        assert float((got - want).abs().max()) <= 2e-6 * float(want.abs().max())
    assert not a1[[0, 2, 3, 4, 5]].any()          # only the alpha and distortion channels receive gradient
