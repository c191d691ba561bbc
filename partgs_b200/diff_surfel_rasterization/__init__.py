"""Drop-in replacement for PartGS's ``diff_surfel_rasterization`` package.

Same public surface as the reference
(submodules/diff-surfel-rasterization/diff_surfel_rasterization/__init__.py):
``GaussianRasterizationSettings`` (:158-170), ``GaussianRasterizer`` (:172-222),
``rasterize_gaussians`` (:21-42) and the autograd function ``_RasterizeGaussians``
(:44-156); outputs ``(color[3,H,W], radii[P] int32, allmap[7,H,W])`` and the same
gradient slots.  The work is done by hand-written sm_100a kernels behind the C ABI in
include/partgs_b200.h — there is no CPU / PyTorch fallback.
"""
from __future__ import annotations

from typing import NamedTuple

import torch
import torch.nn as nn

from .. import _lib

NUM_CHANNELS = 3
NUM_AUX = 7


def cpu_deep_copy_tuple(input_tuple):
    copied_tensors = [item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple]
    return tuple(copied_tensors)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings)


def _native_forward(bg, means3D, colors, opacity, scales, rotations, scale_modifier, transMat_precomp, viewmatrix,
                    projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos, prefiltered, debug):
    """Mirror of RasterizeGaussiansCUDA (reference rasterize_points.cu:39-134)."""
    lib = _lib.load()
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    means3D = _lib.require_cuda_f32(means3D, "means3D")
    dev = means3D.device
    bg = _lib.require_cuda_f32(bg, "background")
    colors = _lib.require_cuda_f32(colors, "colors")
    opacity = _lib.require_cuda_f32(opacity, "opacity")
    scales = _lib.require_cuda_f32(scales, "scales")
    rotations = _lib.require_cuda_f32(rotations, "rotations")
    transMat_precomp = _lib.require_cuda_f32(transMat_precomp, "transMat_precomp")
    viewmatrix = _lib.require_cuda_f32(viewmatrix, "viewmatrix")
    projmatrix = _lib.require_cuda_f32(projmatrix, "projmatrix")
    sh = _lib.require_cuda_f32(sh, "sh")
    campos = _lib.require_cuda_f32(campos, "campos")

    P = means3D.size(0)
    H, W = int(image_height), int(image_width)
    out_color = torch.empty((NUM_CHANNELS, H, W), dtype=torch.float32, device=dev)
    out_others = torch.empty((NUM_AUX, H, W), dtype=torch.float32, device=dev)
    radii = torch.empty((P,), dtype=torch.int32, device=dev)
    sc = _lib.AllocScope(dev)
    rendered = 0
    if P != 0:
        M = sh.size(1) if sh.numel() != 0 else 0
        with torch.cuda.device(dev), sc:
            rc = lib.pgs_dsr_forward(
                _lib.ALLOC_CB, sc.GEOM, _lib.ALLOC_CB, sc.BINNING, _lib.ALLOC_CB, sc.IMAGE, P, int(degree), int(M),
                _lib.ptr(bg), W, H, _lib.ptr(means3D), _lib.ptr(sh), _lib.ptr(colors), _lib.ptr(opacity),
                _lib.ptr(scales), float(scale_modifier), _lib.ptr(rotations), _lib.ptr(transMat_precomp),
                _lib.ptr(viewmatrix), _lib.ptr(projmatrix), _lib.ptr(campos), float(tan_fovx), float(tan_fovy),
                int(bool(prefiltered)), _lib.ptr(out_color), _lib.ptr(out_others), _lib.ptr(radii),
                int(bool(debug)) | (2 if _lazy_count else 0), _lib.current_stream(dev))
        if sc.error is not None:
            raise sc.error
        rendered = _lib.check(rc, "pgs_dsr_forward")
    else:
        # reference: zero images and no state when there is nothing to draw (rasterize_points.cu:85-99)
        out_color.zero_()
        out_others.zero_()
    return rendered, out_color, out_others, radii, sc.tensor(sc.GEOM), sc.tensor(sc.BINNING), sc.tensor(sc.IMAGE)


_lazy_count = False


def set_lazy_count(flag: bool):
    """Do not wait for the frame's instance count inside the forward call (PGS_FWD_LAZY_COUNT, include/partgs_b200.h).
    The reference blocks on that count before binning (rasterizer_impl.cu:282); by default this library waits after
    it has queued the whole frame.  Lazy: the forward returns at once, the count is checked at the entry of the
    backward pass (`resolve_count`) — by then it has long arrived, so the host runs a whole forward ahead of the GPU.
    A frame that needed more instances than the remembered capacity raises there (its outputs are invalid; render
    the scene's views once eagerly first — the capacity is 1.25 x the largest frame seen).  Frames issued under
    CUDA-graph capture are lazy regardless (`partgs_b200.graphs`)."""
    global _lazy_count
    _lazy_count = bool(flag)


def resolve_count(device=None):
    """(num_rendered, overflow) of the lazy frames outstanding on the current device (see set_lazy_count)."""
    import ctypes as C
    ov = C.c_int(0)
    with (torch.cuda.device(device) if device is not None else _null()):
        n = _lib.load().pgs_dsr_resolve_count(C.byref(ov))
    return _lib.check(n, "pgs_dsr_resolve_count"), bool(ov.value)


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


COUNT_PENDING = 0x7fffffff
_bucket_provider = None


def set_grad_bucket_provider(fn):
    """``fn(numel, device) -> flat float32 tensor, (tensor, accumulate) or None``: where the backward pass puts its
    five parameter gradients (one flat "bucket").  A data-parallel caller hands out peer-mapped memory here so that
    the gradients are reduced in place (partgs_b200.dist).  With ``accumulate`` true the kernel ADDS this view's
    gradients to what the bucket holds (the earlier views of a batch) instead of overwriting it — the caller must
    then drop ``.grad`` of the parameters before every backward (autograd would otherwise add the bucket to
    itself).  None restores torch.empty."""
    global _bucket_provider
    _bucket_provider = fn


def _carve_bucket(dev, shapes, align_elems=64, with_flag=False):
    """Views of the given shapes into one flat float32 buffer (each view 256-byte aligned).
    The views' ``_base`` is the bucket itself."""
    offs, total = [], 0
    for shp in shapes:
        n = 1
        for d in shp:
            n *= int(d)
        offs.append((total, n))
        total += (n + align_elems - 1) // align_elems * align_elems
    flat = _bucket_provider(max(total, 1), dev) if _bucket_provider is not None else None
    accumulate = False
    if isinstance(flat, tuple):
        flat, accumulate = flat
    if flat is None:
        flat, accumulate = torch.empty((max(total, 1),), dtype=torch.float32, device=dev), False
    views = [flat[o:o + n].view(*shp) for (o, n), shp in zip(offs, shapes)]
    return (views, bool(accumulate)) if with_flag else views


def bucket_numel(P: int, M: int = 16, align_elems: int = 64) -> int:
    """Floats of the gradient bucket of a model with P surfels and M SH coefficients (layout of _carve_bucket)."""
    return sum((n + align_elems - 1) // align_elems * align_elems for n in (P * 3, P * M * 3, P, P * 2, P * 4))


def zero_bucket_grads(P: int, M: int, dev):
    """Zero gradients (means3D, sh, opacity, scales, rotations) laid out in the provider's bucket exactly like a
    backward pass would leave them: what a rank WITHOUT views contributes to a data-parallel batch, so that every
    rank issues the same collective."""
    views, accumulate = _carve_bucket(dev, [(P, 3), (P, M, 3), (P, 1), (P, 2), (P, 4)], with_flag=True)
    if not accumulate:
        for v in views:
            v.zero_()
    return views


def _native_backward(bg, means3D, radii, colors, scales, rotations, scale_modifier, transMat_precomp, viewmatrix,
                     projmatrix, tan_fovx, tan_fovy, dL_dout_color, dL_dout_others, sh, degree, campos, geomBuffer, R,
                     binningBuffer, imageBuffer, debug):
    """Mirror of RasterizeGaussiansBackwardCUDA (reference rasterize_points.cu:136-233)."""
    lib = _lib.load()
    dev = means3D.device
    P = means3D.size(0)
    H, W = dL_dout_color.size(1), dL_dout_color.size(2)
    # an empty model (P == 0) still carries sh [0,M,3]: the gradient must have that shape (the reference derives M
    # from sh.size(0) != 0 and then fails autograd's shape check)
    M = sh.size(1) if sh.dim() == 3 else 0
    f32 = dict(dtype=torch.float32, device=dev)
    # every row is written by the kernels -> torch.empty, no 304 B/surfel zero fill.
    # The five parameter gradients (232 B/surfel) are carved out of ONE flat buffer ("gradient
    # bucket"): the backward-preprocess kernel writes them in place, and a data-parallel caller can
    # all-reduce the whole bucket with a single collective (see partgs_b200.dist.grad_bucket).
    (dL_dmeans3D, dL_dsh, dL_dopacity, dL_dscales, dL_drotations), accumulate = _carve_bucket(
        dev, [(P, 3), (P, M, 3), (P, 1), (P, 2), (P, 4)], with_flag=True)
    dL_dmeans2D = torch.empty((P, 3), **f32)
    dL_dcolors = torch.empty((P, NUM_CHANNELS), **f32)
    dL_dtransMat = torch.empty((P, 9), **f32)
    if P != 0:
        if int(R) == COUNT_PENDING and not torch.cuda.is_current_stream_capturing():
            # lazy forward: the count has arrived long ago (it is produced early in the frame); make sure it fitted
            n, overflow = resolve_count()
            if overflow:
                raise RuntimeError(f"lazy instance count: a frame needed {n} instances, more than it was queued for; its "
                                   "outputs are invalid.  The remembered capacity has been raised — render again (or "
                                   "render every view once before set_lazy_count(True))")
        means3D = means3D.contiguous()
        dL_dout_color = _lib.require_cuda_float(dL_dout_color, "dL_dout_color")
        dL_dout_others = _lib.require_cuda_float(dL_dout_others, "dL_dout_others")
        scratch = torch.empty(lib.pgs_dsr_backward_scratch_bytes(P), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = lib.pgs_dsr_backward(
                P, int(degree), int(M), int(R), _lib.ptr(bg.contiguous()), W, H, _lib.ptr(means3D),
                _lib.ptr(sh.contiguous()), _lib.ptr(colors.contiguous()), _lib.ptr(scales.contiguous()),
                float(scale_modifier), _lib.ptr(rotations.contiguous()), _lib.ptr(transMat_precomp.contiguous()),
                _lib.ptr(viewmatrix.contiguous()), _lib.ptr(projmatrix.contiguous()), _lib.ptr(campos.contiguous()),
                float(tan_fovx), float(tan_fovy), _lib.ptr(radii), _lib.ptr(geomBuffer), _lib.ptr(binningBuffer), int(binningBuffer.numel()),
                _lib.ptr(imageBuffer), _lib.ptr(dL_dout_color), _lib.ptr(dL_dout_others), _lib.ptr(dL_dmeans2D),
                _lib.ptr(scratch), _lib.ptr(dL_dopacity), _lib.ptr(dL_dcolors), _lib.ptr(dL_dmeans3D),
                _lib.ptr(dL_dtransMat), _lib.ptr(dL_dsh), _lib.ptr(dL_dscales), _lib.ptr(dL_drotations),
                int(bool(debug)) | (2 if accumulate else 0), _lib.current_stream(dev))
        _lib.check(rc, "pgs_dsr_backward")
    return dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dtransMat, dL_dsh, dL_dscales, dL_drotations


def _native_mark_visible(means3D, viewmatrix, projmatrix):
    """Mirror of markVisible (reference rasterize_points.cu:235-254)."""
    lib = _lib.load()
    means3D = _lib.require_cuda_float(means3D, "means3D")
    P = means3D.size(0)
    present = torch.zeros((P,), dtype=torch.bool, device=means3D.device)
    if P != 0:
        with torch.cuda.device(means3D.device):
            rc = lib.pgs_mark_visible(P, _lib.ptr(means3D), _lib.ptr(viewmatrix.contiguous()),
                                      _lib.ptr(projmatrix.contiguous()), _lib.ptr(present),
                                      _lib.current_stream(means3D.device))
        _lib.check(rc, "pgs_mark_visible")
    return present


class _C:
    """Namespace with the reference pybind module's three entry points (ext.cpp:15-18)."""
    rasterize_gaussians = staticmethod(_native_forward)
    rasterize_gaussians_backward = staticmethod(_native_backward)
    mark_visible = staticmethod(_native_mark_visible)


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings):
        args = (
            raster_settings.bg, means3D, colors_precomp, opacities, scales, rotations,
            raster_settings.scale_modifier, cov3Ds_precomp, raster_settings.viewmatrix, raster_settings.projmatrix,
            raster_settings.tanfovx, raster_settings.tanfovy, raster_settings.image_height,
            raster_settings.image_width, sh, raster_settings.sh_degree, raster_settings.campos,
            raster_settings.prefiltered, raster_settings.debug,
        )
        if raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(args)  # copy them before they can be corrupted
            try:
                num_rendered, color, depth, radii, geomBuffer, binningBuffer, imgBuffer = _C.rasterize_gaussians(*args)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            num_rendered, color, depth, radii, geomBuffer, binningBuffer, imgBuffer = _C.rasterize_gaussians(*args)

        ctx.raster_settings = raster_settings
        ctx.num_rendered = num_rendered
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer,
                              binningBuffer, imgBuffer)
        ctx.mark_non_differentiable(radii)
        return color, radii, depth

    @staticmethod
    def backward(ctx, grad_out_color, grad_radii, grad_depth):
        num_rendered = ctx.num_rendered
        raster_settings = ctx.raster_settings
        (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer,
         imgBuffer) = ctx.saved_tensors

        args = (raster_settings.bg, means3D, radii, colors_precomp, scales, rotations, raster_settings.scale_modifier,
                cov3Ds_precomp, raster_settings.viewmatrix, raster_settings.projmatrix, raster_settings.tanfovx,
                raster_settings.tanfovy, grad_out_color, grad_depth, sh, raster_settings.sh_degree,
                raster_settings.campos, geomBuffer, num_rendered, binningBuffer, imgBuffer, raster_settings.debug)

        if raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(args)
            try:
                (grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp, grad_sh,
                 grad_scales, grad_rotations) = _C.rasterize_gaussians_backward(*args)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise ex
        else:
            (grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp, grad_sh,
             grad_scales, grad_rotations) = _C.rasterize_gaussians_backward(*args)

        # gradient slots in input order (reference __init__.py:142-154); autograd ignores the
        # entries that belong to empty placeholder inputs.
        need = ctx.needs_input_grad
        grads = (
            grad_means3D if need[0] else None,
            grad_means2D if need[1] else None,
            grad_sh if need[2] else None,
            grad_colors_precomp if need[3] else None,
            grad_opacities if need[4] else None,
            grad_scales if need[5] else None,
            grad_rotations if need[6] else None,
            grad_cov3Ds_precomp if need[7] else None,
            None,
        )
        return grads


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        # boolean mask of points in front of the near plane for this camera
        with torch.no_grad():
            raster_settings = self.raster_settings
            visible = _C.mark_visible(positions, raster_settings.viewmatrix, raster_settings.projmatrix)
        return visible

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        raster_settings = self.raster_settings

        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')

        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')

        empty = torch.empty(0, dtype=torch.float32, device=means3D.device)
        if shs is None:
            shs = empty
        if colors_precomp is None:
            colors_precomp = empty
        if scales is None:
            scales = empty
        if rotations is None:
            rotations = empty
        if cov3D_precomp is None:
            cov3D_precomp = empty

        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                   raster_settings)
