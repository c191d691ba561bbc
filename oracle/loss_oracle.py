"""TEST INFRASTRUCTURE — torch restatement of the reference's photometric loss (utils/loss_utils.py:6-54 and
train.py:230-231), any device / dtype.  Checker only; pinned by golden vectors produced by the reference's own
loss_utils (tools/make_golden_loss.py -> tests/golden/photometric_*.npz)."""
from math import exp

import torch
import torch.nn.functional as F


def gaussian(window_size, sigma):  # loss_utils.py:12-14
    gauss = torch.Tensor([exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)])
    return gauss / gauss.sum()


def create_window(window_size, channel):  # :16-20
    _1D_window = gaussian(window_size, 1.5).unsqueeze(1)
    _2D_window = _1D_window.mm(_1D_window.t()).float().unsqueeze(0).unsqueeze(0)
    return _2D_window.expand(channel, 1, window_size, window_size).contiguous()


def ssim(img1, img2, window_size=11):  # :22-54, size_average=True
    channel = img1.size(-3)
    window = create_window(window_size, channel).to(img1.device).type_as(img1)
    pad = window_size // 2
    mu1 = F.conv2d(img1, window, padding=pad, groups=channel)
    mu2 = F.conv2d(img2, window, padding=pad, groups=channel)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    sigma1_sq = F.conv2d(img1 * img1, window, padding=pad, groups=channel) - mu1_sq
    sigma2_sq = F.conv2d(img2 * img2, window, padding=pad, groups=channel) - mu2_sq
    sigma12 = F.conv2d(img1 * img2, window, padding=pad, groups=channel) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ssim_map = ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2))
    return ssim_map.mean()


def l1_loss(network_output, gt):  # :6-7
    return torch.abs((network_output - gt)).mean()


def photometric_loss(image, gt, lambda_dssim):  # train.py:230-231
    return (1.0 - lambda_dssim) * l1_loss(image, gt) + lambda_dssim * (1.0 - ssim(image, gt))


def geometric_regularizers(rend_alpha, gt_mask, rend_dist, rend_normal, surf_normal, lambda_mask_entropy,
                           lambda_normal, lambda_dist):  # train.py:234-251 (the terms after the photometric loss)
    """-> (lambda_me * loss_mask_entropy + normal_loss + dist_loss, (loss_mask_entropy, normal_error mean, dist mean)).
    ``gt_mask`` may be None (term skipped)."""
    total = 0.0
    me = None
    if gt_mask is not None:
        opacity = rend_alpha.clamp(1e-6, 1 - 1e-6).squeeze(0)
        me = -(gt_mask * torch.log(opacity) + (1 - gt_mask) * torch.log(1 - opacity)).mean()
        total = total + lambda_mask_entropy * me
    normal_error = (1 - (rend_normal * surf_normal).sum(dim=0))[None]
    normal_loss = lambda_normal * (normal_error).mean()
    dist_loss = lambda_dist * (rend_dist).mean()
    return total + dist_loss + normal_loss, (me, normal_error.mean(), rend_dist.mean())
