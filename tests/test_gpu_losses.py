"""GPU parity of the fused photometric loss (L1 + SSIM, pgs_photometric_forward/_backward) against golden vectors
from the reference's utils/loss_utils.py and against the restated reference code at other sizes."""
from pathlib import Path

import numpy as np
import pytest
import torch

import parity_utils as pu

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = sorted((Path(__file__).parent / "golden").glob("photometric_*.npz"))


@pytest.mark.parametrize("path", GOLD, ids=[p.stem for p in GOLD])
def test_photometric_vs_reference_golden(path):
    from partgs_b200.losses import photometric_loss
    z = np.load(path)
    img = torch.from_numpy(z["image"]).to(DEV).requires_grad_(True)
    gt = torch.from_numpy(z["gt"]).to(DEV)
    loss, l1, ssim = photometric_loss(img, gt, float(z["lam"]), return_parts=True)
    assert abs(float(l1) - float(z["l1"])) <= 1e-6 * abs(float(z["l1"]))
    assert abs(float(ssim) - float(z["ssim"])) <= 1e-5
    assert abs(float(loss) - float(z["loss"])) <= 1e-5 * abs(float(z["loss"])) + 1e-7
    loss.backward()
    assert pu.rel_err(img.grad.cpu(), torch.from_numpy(z["d_image"])) <= 1e-4


@pytest.mark.parametrize("H,W,lam", [(300, 400, 0.2), (97, 131, 0.2), (11, 5, 0.5), (1200, 1600, 0.2)])
def test_photometric_vs_oracle(H, W, lam):
    from oracle import loss_oracle
    from partgs_b200.losses import photometric_loss, l1_loss, ssim
    gen = torch.Generator().manual_seed(H * 7 + W)
    gt = torch.rand(3, H, W, generator=gen)
    base = (gt + 0.1 * torch.randn(3, H, W, generator=gen)).clamp(0, 1)
    img = base.clone().to(DEV).requires_grad_(True)
    loss = photometric_loss(img, gt.to(DEV), lam)
    (loss * 3.0).backward()        # non-unit upstream gradient
    # float64 evaluation of the reference code as the yardstick (its fp32 convolution loses ~1e-6 itself)
    i64 = base.double().requires_grad_(True)
    ref = loss_oracle.photometric_loss(i64, gt.double(), lam)
    (ref * 3.0).backward()
    assert abs(float(loss) - float(ref)) <= 2e-6 * abs(float(ref)) + 1e-7
    assert pu.rel_err(img.grad.cpu().double(), i64.grad) <= 2e-5
    if H <= 300:
        i32 = base.clone().requires_grad_(True)
        ref32 = loss_oracle.photometric_loss(i32, gt, lam)
        ref32.backward()
        assert abs(float(loss) - float(ref32)) <= 1e-5 * abs(float(ref32))
        # the reference-named views of the op
        assert abs(float(l1_loss(base.to(DEV), gt.to(DEV))) - float(loss_oracle.l1_loss(base, gt))) <= 1e-6
        assert abs(float(ssim(base.to(DEV), gt.to(DEV))) - float(loss_oracle.ssim(base, gt))) <= 1e-5
