"""Fused optimiser step and densification bookkeeping (the per-iteration part).

``FusedAdam`` is a drop-in for the ``torch.optim.Adam(l, lr=0.0, eps=1e-15)`` PartGS builds in
``GaussianModel.training_setup`` (scene/gaussian_model.py:256-266): same param-group interface (``lr`` per group is
what the reference's schedulers write), same ``state[p] = {step, exp_avg, exp_avg_sq}`` layout (so the reference's
``_prune_optimizer`` / ``cat_tensors_to_optimizer`` keep working), same arithmetic — but one kernel launch per step
for all groups (csrc/optim.cu, ``pgs_adam_step``).  ``densification_stats`` fuses ``add_densification_stats`` and the
``max_radii2D`` update.  CUDA float32 tensors only; anything else raises.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib

_MAX = 16


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        if not 0.0 <= eps:
            raise ValueError(f"Invalid epsilon value: {eps}")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        # tensors that share (step, betas, eps) go into one launch
        batches = {}
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not (_lib.on_device(p) and p.dtype == torch.float32 and p.grad.dtype == torch.float32):
                    raise RuntimeError("FusedAdam: CUDA float32 parameters and gradients only (no fallback)")
                if p.grad.is_sparse:
                    raise RuntimeError("FusedAdam does not support sparse gradients")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] = int(st["step"]) + 1
                key = (p.device, st["step"], beta1, beta2, group["eps"])
                batches.setdefault(key, []).append((p, p.grad.contiguous(), st, group["lr"]))
        for (dev, step, beta1, beta2, eps), items in batches.items():
            bias_correction1 = 1 - beta1 ** step
            bias_correction2_sqrt = (1 - beta2 ** step) ** 0.5
            for s0 in range(0, len(items), _MAX):
                chunk = items[s0:s0 + _MAX]
                n = len(chunk)
                for p, g, st, _ in chunk:
                    if not (p.is_contiguous() and st["exp_avg"].is_contiguous() and st["exp_avg_sq"].is_contiguous()):
                        raise RuntimeError("FusedAdam: parameters and optimiser state must be contiguous")
                pa = (C.c_void_p * n)(*[p.data_ptr() for p, _, _, _ in chunk])
                ga = (C.c_void_p * n)(*[g.data_ptr() for _, g, _, _ in chunk])
                ma = (C.c_void_p * n)(*[st["exp_avg"].data_ptr() for _, _, st, _ in chunk])
                va = (C.c_void_p * n)(*[st["exp_avg_sq"].data_ptr() for _, _, st, _ in chunk])
                na = (C.c_size_t * n)(*[p.numel() for p, _, _, _ in chunk])
                sa = (C.c_float * n)(*[lr / bias_correction1 for _, _, _, lr in chunk])
                with torch.cuda.device(dev):
                    rc = lib.pgs_adam_step(n, pa, ga, ma, va, na, sa, float(beta1), float(beta2), float(eps),
                                           float(bias_correction2_sqrt), _lib.current_stream(dev))
                _lib.check(rc, "pgs_adam_step")
        return loss


def densification_stats(radii, viewspace_grad, xyz_gradient_accum, denom, max_radii2D=None):
    """In place, for ``visibility_filter = radii > 0``: ``max_radii2D = max(max_radii2D, radii)``,
    ``xyz_gradient_accum += ||viewspace_grad[:, :2]||``, ``denom += 1`` (train.py:295-297 +
    scene/gaussian_model.py:515-517)."""
    lib = _lib.load()
    P = radii.numel()
    if not (_lib.on_device(radii) and radii.dtype == torch.int32):
        raise RuntimeError("radii must be a CUDA int32 tensor")
    g = _lib.require_cuda_float(viewspace_grad, "viewspace_grad")
    for t, nm in ((xyz_gradient_accum, "xyz_gradient_accum"), (denom, "denom"), (max_radii2D, "max_radii2D")):
        if t is not None and not (_lib.on_device(t) and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == P):
            raise RuntimeError(f"{nm} must be a contiguous CUDA float32 tensor with one entry per surfel")
    if g.dim() != 2 or g.size(0) != P or g.size(1) < 3 or g.stride(0) != 3:
        raise RuntimeError("viewspace_grad must be [P,3] contiguous")
    with torch.cuda.device(radii.device):
        rc = lib.pgs_densify_stats(P, radii.data_ptr(), g.data_ptr(), _lib.ptr(max_radii2D),
                                   xyz_gradient_accum.data_ptr(), denom.data_ptr(), _lib.current_stream(radii.device))
    _lib.check(rc, "pgs_densify_stats")
