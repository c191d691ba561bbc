"""CPU: oracle/loss_oracle.py reproduces the golden vectors written by the reference's own utils/loss_utils.py."""
from pathlib import Path

import numpy as np
import pytest
import torch

GOLD = sorted((Path(__file__).parent / "golden").glob("photometric_*.npz"))


@pytest.mark.parametrize("path", GOLD, ids=[p.stem for p in GOLD])
def test_loss_oracle_matches_reference_golden(path):
    from oracle import loss_oracle
    z = np.load(path)
    img = torch.from_numpy(z["image"]).requires_grad_(True)
    gt = torch.from_numpy(z["gt"])
    lam = float(z["lam"])
    assert float(loss_oracle.l1_loss(img, gt)) == float(z["l1"])
    assert float(loss_oracle.ssim(img, gt)) == float(z["ssim"])
    loss = loss_oracle.photometric_loss(img, gt, lam)
    loss.backward()
    assert float(loss) == float(z["loss"])
    assert torch.equal(img.grad, torch.from_numpy(z["d_image"]))
