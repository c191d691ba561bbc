"""Golden vectors for the fused photometric loss from the REFERENCE's utils/loss_utils.py, imported unmodified
(CPU tensors take its non-CUDA branch).  Build container only.  -> tests/golden/photometric_<H>x<W>.npz"""
import importlib.util, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
spec = importlib.util.spec_from_file_location("ref_loss_utils", "/root/reference/utils/loss_utils.py")
ref = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref)
for (H, W, seed, lam) in ((23, 37, 1, 0.2), (48, 64, 2, 0.2), (16, 16, 3, 1.0)):
    gen = torch.Generator().manual_seed(seed)
    gt = torch.rand(3, H, W, generator=gen)
    img = (gt + 0.1 * torch.randn(3, H, W, generator=gen)).clamp(0, 1).requires_grad_(True)
    Ll1 = ref.l1_loss(img, gt)
    s = ref.ssim(img, gt)
    loss = (1.0 - lam) * Ll1 + lam * (1.0 - s)     # train.py:230-231
    loss.backward()
    np.savez_compressed(ROOT / "tests" / "golden" / f"photometric_{H}x{W}.npz", image=img.detach().numpy(),
                        gt=gt.numpy(), lam=lam, l1=float(Ll1), ssim=float(s), loss=float(loss), d_image=img.grad.numpy())
    print(H, W, float(Ll1), float(s), float(loss))
