"""Fused image losses of the PartGS training step.

``photometric_loss(image, gt, lambda_dssim)`` = ``(1 - lambda) * l1_loss(image, gt) + lambda * (1 - ssim(image, gt))``
exactly as train.py:230-231 combines utils/loss_utils.py:6-7 (l1_loss) and :12-54 (ssim, 11x11 Gaussian window,
sigma 1.5, zero padding, mean over all elements) — one CUDA kernel forward and one backward
(csrc/photometric.cu) instead of five grouped conv2d + ~15 pointwise kernels each way.  ``l1_loss`` / ``ssim``
with the reference's names are thin views of the same op.  CUDA tensors only, no fallback.
"""
from __future__ import annotations

import torch

from . import _lib


class _Photometric(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, gt, lambda_dssim):
        lib = _lib.load()
        image = _lib.require_cuda_float(image, "image")
        gt = _lib.require_cuda_float(gt.detach(), "gt")
        if image.shape != gt.shape or image.dim() != 3:
            raise RuntimeError("image and gt must both be [C,H,W]")
        C, H, W = image.shape
        dev = image.device
        sums = torch.empty(2, dtype=torch.float64, device=dev)
        dmaps = torch.empty((3, C, H, W), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = lib.pgs_photometric_forward(C, H, W, image.data_ptr(), gt.data_ptr(), sums.data_ptr(),
                                             dmaps.data_ptr(), _lib.current_stream(dev))
        _lib.check(rc, "pgs_photometric_forward")
        n = float(C * H * W)
        means = (sums / n).float()           # [ssim, l1]
        loss = (1.0 - lambda_dssim) * means[1] + lambda_dssim * (1.0 - means[0])
        ctx.save_for_backward(image, gt, dmaps)
        ctx.lambda_dssim = float(lambda_dssim)
        ctx.mark_non_differentiable(means)
        return loss, means

    @staticmethod
    def backward(ctx, g_loss, _g_means):
        lib = _lib.load()
        image, gt, dmaps = ctx.saved_tensors
        C, H, W = image.shape
        dev = image.device
        g = _lib.require_cuda_float(g_loss.reshape(1), "g_loss")
        g_image = torch.empty_like(image)
        with torch.cuda.device(dev):
            rc = lib.pgs_photometric_backward(C, H, W, image.data_ptr(), gt.data_ptr(), dmaps.data_ptr(), g.data_ptr(),
                                              ctx.lambda_dssim, g_image.data_ptr(), _lib.current_stream(dev))
        _lib.check(rc, "pgs_photometric_backward")
        return g_image, None, None


def photometric_loss(image, gt, lambda_dssim: float = 0.2, return_parts: bool = False):
    """(1 - lambda) * L1 + lambda * (1 - SSIM); with ``return_parts`` also (Ll1, ssim) as detached scalars."""
    loss, means = _Photometric.apply(image, gt, float(lambda_dssim))
    if return_parts:
        return loss, means[1], means[0]
    return loss


def l1_loss(network_output, gt):
    """utils/loss_utils.py:6-7"""
    return _Photometric.apply(network_output, gt, 0.0)[0]


def ssim(img1, img2, window_size=11, size_average=True):
    """utils/loss_utils.py:22-31 (window_size 11, size_average=True only)."""
    if window_size != 11 or not size_average:
        raise NotImplementedError("fused ssim: window_size=11, size_average=True")
    return 1.0 - _Photometric.apply(img1, img2, 1.0)[0]
