"""Shared helpers for the parity tests (product CUDA path vs oracle)."""
from __future__ import annotations

import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

# north_star tolerances (BASELINE.json): images 1e-5 relative, gradients 1e-4 relative (fp32)
IMG_RTOL = 1e-5
GRAD_RTOL = 1e-4


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b|  (norm-wise relative error; b is the reference)."""
    a = a.double()
    b = b.double()
    if a.numel() == 0:
        return 0.0
    d = (a - b).abs().max().item()
    s = b.abs().max().item()
    return d / (s + 1e-30)


def mismatch_frac(a: torch.Tensor, b: torch.Tensor, rtol: float) -> float:
    """fraction of elements with |a-b| > rtol * (|b| + mean|b|)"""
    a = a.double()
    b = b.double()
    if a.numel() == 0:
        return 0.0
    tol = rtol * (b.abs() + b.abs().mean())
    return ((a - b).abs() > tol).double().mean().item()


def grad_violations(a: torch.Tensor, b: torch.Tensor, rtol: float = GRAD_RTOL, afloor: float = 1e-6) -> float:
    """Element-wise gate: fraction of elements with |a-b| > rtol * |b| + afloor * max|b|  (b is the reference).
    The absolute floor covers entries that are sums of cancelling float atomics — there two runs of the reference
    itself differ by more than rtol * |entry|."""
    a = a.double()
    b = b.double()
    if a.numel() == 0:
        return 0.0
    tol = rtol * b.abs() + afloor * b.abs().max()
    return ((a - b).abs() > tol).double().mean().item()


def assert_grad_close(name: str, ours: torch.Tensor, ref: torch.Tensor, rtol: float = GRAD_RTOL, afloor: float = 1e-6,
                      max_frac: float = 0.0):
    """north_star gradient gate, ELEMENT-wise: every entry within rtol * |ref| + afloor * max|ref| (see grad_violations),
    plus the norm-wise bound.  On failure reports the fraction and the worst offender.  `max_frac`: fraction of entries
    allowed outside (0 unless the caller measured the reference's own run-to-run level for the case)."""
    assert ours.shape == ref.shape, (name, ours.shape, ref.shape)
    if ours.numel() == 0:
        return
    frac = grad_violations(ours, ref, rtol, afloor)
    e = rel_err(ours, ref)
    if frac > max_frac or not (e <= rtol):
        a, b = ours.double().flatten(), ref.double().flatten()
        excess = (a - b).abs() - (rtol * b.abs() + afloor * b.abs().max())
        i = int(excess.argmax())
        raise AssertionError(f"{name}: {frac:.3e} of {a.numel()} elements outside {rtol:g}*|ref| + {afloor:g}*max|ref|; "
                             f"worst: ours {a[i].item():.9g} vs ref {b[i].item():.9g}; norm-wise rel err {e:.3e}")


def assert_equal_images(name: str, ours: torch.Tensor, ref: torch.Tensor):
    """Bit-identical images (base fork: the per-fragment rounding sequence is pinned in frag_math.cuh)."""
    if not torch.equal(ours, ref):
        neq = ours != ref
        d = (ours.double() - ref.double()).abs()
        raise AssertionError(f"{name}: {int(neq.sum())} of {neq.numel()} elements differ, max |diff| {float(d.max()):.3e}")


def robust_close(a, b, atol_frac=1e-4, max_frac=2e-3):
    """Comparison against the CPU restatement, which rounds differently from the GPU (no FMA
    contraction): a few fragments sit on the other side of the alpha >= 1/255 or T < 1e-4
    tests, which moves single pixels / surfels by up to ~4e-3.  Require that all but
    `max_frac` of the elements agree to atol_frac * max|b| and that the mean error is tiny."""
    a = torch.as_tensor(a).double().flatten()
    b = torch.as_tensor(b).double().flatten()
    scale = b.abs().max().item() + 1e-30
    d = (a - b).abs()
    frac_bad = (d > atol_frac * scale).double().mean().item()
    mean_err = d.mean().item() / scale
    return frac_bad <= max_frac and mean_err <= atol_frac, (frac_bad, mean_err)


def settings_from_cam(cam, bg, sh_degree=3, scale_modifier=1.0, debug=False):
    from partgs_b200.diff_surfel_rasterization import GaussianRasterizationSettings
    return GaussianRasterizationSettings(
        image_height=cam.image_height, image_width=cam.image_width, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=bg,
        scale_modifier=scale_modifier, viewmatrix=cam.viewmatrix, projmatrix=cam.projmatrix, sh_degree=sh_degree,
        campos=cam.campos, prefiltered=False, debug=debug)


def run_ours(scene, cam, bg, grads=None, sh_degree=3, scale_modifier=1.0, colors_precomp=None):
    """Forward (+ backward when `grads` is given) through the public drop-in API."""
    from partgs_b200.diff_surfel_rasterization import GaussianRasterizer, _C
    leaf = {}
    for k in ("means3D", "scales", "rotations", "opacities", "shs"):
        leaf[k] = scene[k].detach().clone().requires_grad_(grads is not None)
    means2D = torch.zeros_like(leaf["means3D"], requires_grad=grads is not None)
    rast = GaussianRasterizer(settings_from_cam(cam, bg, sh_degree, scale_modifier))
    kw = dict(means3D=leaf["means3D"], means2D=means2D, opacities=leaf["opacities"], scales=leaf["scales"],
              rotations=leaf["rotations"])
    cp = None
    if colors_precomp is None:
        kw["shs"] = leaf["shs"]
    else:
        cp = colors_precomp.detach().clone().requires_grad_(grads is not None)
        kw["colors_precomp"] = cp
    color, radii, allmap = rast(**kw)
    out = dict(color=color.detach(), radii=radii, allmap=allmap.detach())
    if grads is not None:
        torch.autograd.backward([color, allmap], [grads["color"], grads["allmap"]])
        out["grads"] = dict(means3D=leaf["means3D"].grad, means2D=means2D.grad, opacity=leaf["opacities"].grad,
                            scales=leaf["scales"].grad, rotations=leaf["rotations"].grad)
        if cp is None:
            out["grads"]["sh"] = leaf["shs"].grad
        else:
            out["grads"]["colors"] = cp.grad
    return out


def run_ours_raw(scene, cam, bg, sh_degree=3, scale_modifier=1.0):
    """Forward through the native entry point, returning the state buffers too."""
    from partgs_b200.diff_surfel_rasterization import _C
    dev = scene["means3D"].device
    empty = torch.empty(0, device=dev)
    R, color, others, radii, geom, binning, img = _C.rasterize_gaussians(
        bg, scene["means3D"], empty, scene["opacities"], scene["scales"], scene["rotations"], scale_modifier, empty,
        cam.viewmatrix, cam.projmatrix, cam.tanfovx, cam.tanfovy, cam.image_height, cam.image_width, scene["shs"],
        sh_degree, cam.campos, False, False)
    return dict(num_rendered=R, color=color, allmap=others, radii=radii, geom=geom, binning=binning, img=img)
