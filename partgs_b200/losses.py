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


class _Regularizers(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rend_alpha, gt_mask, rend_dist, rend_normal, surf_normal, lam_me, lam_n, lam_d):
        lib = _lib.load()
        rq = _lib.require_cuda_float
        rend_normal, surf_normal = rq(rend_normal, "rend_normal"), rq(surf_normal, "surf_normal")
        if rend_normal.dim() != 3 or rend_normal.size(0) != 3 or rend_normal.shape != surf_normal.shape:
            raise RuntimeError("rend_normal and surf_normal must both be [3,H,W]")
        H, W = int(rend_normal.size(1)), int(rend_normal.size(2))
        rend_dist = rq(rend_dist, "rend_dist")
        rend_alpha = rq(rend_alpha, "rend_alpha")
        if rend_dist.numel() != H * W or rend_alpha.numel() != H * W:
            raise RuntimeError("rend_alpha and rend_dist must be [1,H,W]")
        mask = None
        if gt_mask is not None:
            mask = rq(gt_mask.detach(), "gt_mask")
            if mask.numel() != H * W:
                raise RuntimeError("gt_mask must hold one value per pixel")
        dev = rend_normal.device
        sums = torch.empty(3, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            rc = lib.pgs_regularizers_forward(W, H, rend_alpha.data_ptr(), _lib.ptr(mask), rend_dist.data_ptr(),
                                              rend_normal.data_ptr(), surf_normal.data_ptr(), sums.data_ptr(),
                                              _lib.current_stream(dev))
        _lib.check(rc, "pgs_regularizers_forward")
        means = (sums / float(H * W)).float()      # [mask entropy, normal error, distortion]
        loss = lam_me * means[0] + lam_n * means[1] + lam_d * means[2]
        ctx.save_for_backward(rend_alpha, mask, rend_normal, surf_normal)
        ctx.lams = (float(lam_me), float(lam_n), float(lam_d))
        ctx.alpha_shape, ctx.dist_shape = tuple(rend_alpha.shape), tuple(rend_dist.shape)
        ctx.mark_non_differentiable(means)
        return loss, means

    @staticmethod
    def backward(ctx, g_loss, _g_means):
        lib = _lib.load()
        rend_alpha, mask, rend_normal, surf_normal = ctx.saved_tensors
        _, H, W = rend_normal.shape
        dev = rend_normal.device
        lam_me, lam_n, lam_d = ctx.lams
        g = _lib.require_cuda_float(g_loss.reshape(1), "g_loss")
        need_a = ctx.needs_input_grad[0] and mask is not None
        g_alpha = torch.empty(ctx.alpha_shape, dtype=torch.float32, device=dev) if need_a else None
        g_dist = torch.empty(ctx.dist_shape, dtype=torch.float32, device=dev) if ctx.needs_input_grad[2] else None
        need_n = ctx.needs_input_grad[3] or ctx.needs_input_grad[4]
        g_rn = torch.empty_like(rend_normal) if need_n else None
        g_sn = torch.empty_like(surf_normal) if need_n else None
        with torch.cuda.device(dev):
            rc = lib.pgs_regularizers_backward(W, H, rend_alpha.data_ptr(), _lib.ptr(mask), rend_normal.data_ptr(),
                                               surf_normal.data_ptr(), g.data_ptr(), lam_me, lam_n, lam_d,
                                               _lib.ptr(g_alpha), _lib.ptr(g_dist), _lib.ptr(g_rn), _lib.ptr(g_sn),
                                               _lib.current_stream(dev))
        _lib.check(rc, "pgs_regularizers_backward")
        return g_alpha, None, g_dist, g_rn if ctx.needs_input_grad[3] else None, \
            g_sn if ctx.needs_input_grad[4] else None, None, None, None


def geometric_regularizers(render_pkg, gt_mask, lambda_mask_entropy: float, lambda_normal: float, lambda_dist: float,
                           return_parts: bool = False):
    """``lambda_mask_entropy * loss_mask_entropy + normal_loss + dist_loss`` of train.py:234-251 from the dict
    ``render()`` returns (``rend_alpha, rend_dist, rend_normal, surf_normal``), fused: one CUDA kernel forward (three
    sums) and one backward (csrc/regularizers.cu) instead of ~14 elementwise ATen kernels and three reductions each
    way.  ``gt_mask`` [H,W] (or None: no entropy term).  With ``return_parts`` also the detached
    (mask entropy, mean normal error, mean distortion)."""
    loss, means = _Regularizers.apply(render_pkg["rend_alpha"], gt_mask, render_pkg["rend_dist"],
                                      render_pkg["rend_normal"], render_pkg["surf_normal"],
                                      float(lambda_mask_entropy) if gt_mask is not None else 0.0,
                                      float(lambda_normal), float(lambda_dist))
    if return_parts:
        return loss, (means[0] if gt_mask is not None else None, means[1], means[2])
    return loss
