"""Drop-in replacement for PartGS's ``diff_surfel_rasterization_part`` package (the part-map
fork used by renderer/gaussian_renderer_2d and render.py).

Same public surface as the reference
(submodules/diff-surfel-rasterization_part/diff_surfel_rasterization_part/__init__.py):
``GaussianRasterizationSettings``, ``GaussianRasterizer.forward(means3D, means2D, opacities,
semantics, shs=..., colors_precomp=..., scales=..., rotations=..., cov3D_precomp=...)``
returning ``(color[3,H,W], semantic[S,H,W], radii[P], allmap[8,H,W])`` (:102), gradients in
the reference's slot order (:152-163).  Hand-written sm_100a kernels behind
pgs_dsrp_forward / pgs_dsrp_backward; no CPU / PyTorch fallback.
"""
from __future__ import annotations

from typing import NamedTuple

import torch
import torch.nn as nn

from .. import _lib
from ..diff_surfel_rasterization import _native_mark_visible, cpu_deep_copy_tuple

NUM_CHANNELS = 3
NUM_AUX = 8
MAX_SEMANTIC_TYPES = 16


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, semantics, scales, rotations, cov3Ds_precomp,
                        raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, semantics, scales, rotations,
                                     cov3Ds_precomp, raster_settings)


def _native_forward(bg, means3D, colors, opacity, semantics, scales, rotations, scale_modifier, transMat_precomp,
                    viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos,
                    prefiltered, debug):
    """Mirror of RasterizeGaussiansCUDA of the fork (DSRP/rasterize_points.cu:39-146)."""
    lib = _lib.load()
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    means3D = _lib.require_cuda_f32(means3D, "means3D")
    dev = means3D.device
    rq = _lib.require_cuda_f32
    bg, colors, opacity, semantics = rq(bg, "background"), rq(colors, "colors"), rq(opacity, "opacity"), \
        rq(semantics, "semantics")
    scales, rotations, transMat_precomp = rq(scales, "scales"), rq(rotations, "rotations"), \
        rq(transMat_precomp, "transMat_precomp")
    viewmatrix, projmatrix, sh, campos = rq(viewmatrix, "viewmatrix"), rq(projmatrix, "projmatrix"), rq(sh, "sh"), \
        rq(campos, "campos")
    P = means3D.size(0)
    if semantics.dim() != 2 or semantics.size(0) != P:
        raise RuntimeError("semantics must have dimensions (num_points, semantic_types)")
    S = semantics.size(1)
    if S > MAX_SEMANTIC_TYPES:
        raise RuntimeError(f"at most {MAX_SEMANTIC_TYPES} semantic channels are supported (got {S})")
    H, W = int(image_height), int(image_width)
    f32 = dict(dtype=torch.float32, device=dev)
    out_color = torch.empty((NUM_CHANNELS, H, W), **f32)
    out_semantic = torch.empty((S, H, W), **f32)
    out_others = torch.empty((NUM_AUX, H, W), **f32)
    radii = torch.empty((P,), dtype=torch.int32, device=dev)
    sc = _lib.AllocScope(dev)
    rendered = 0
    if P != 0:
        M = sh.size(1) if sh.numel() != 0 else 0
        with torch.cuda.device(dev), sc:
            rc = lib.pgs_dsrp_forward(
                _lib.ALLOC_CB, sc.GEOM, _lib.ALLOC_CB, sc.BINNING, _lib.ALLOC_CB, sc.IMAGE, P, int(degree), int(M), _lib.ptr(bg),
                W, H, S, _lib.ptr(means3D), _lib.ptr(sh), _lib.ptr(colors), _lib.ptr(semantics), _lib.ptr(opacity),
                _lib.ptr(scales), float(scale_modifier), _lib.ptr(rotations), _lib.ptr(transMat_precomp),
                _lib.ptr(viewmatrix), _lib.ptr(projmatrix), _lib.ptr(campos), float(tan_fovx), float(tan_fovy),
                int(bool(prefiltered)), _lib.ptr(out_color), _lib.ptr(out_semantic), _lib.ptr(out_others),
                _lib.ptr(radii), int(bool(debug)), _lib.current_stream(dev))
        if sc.error is not None:
            raise sc.error
        rendered = _lib.check(rc, "pgs_dsrp_forward")
    else:
        out_color.zero_()
        out_semantic.zero_()
        out_others.zero_()
    return rendered, out_color, out_semantic, out_others, radii, sc.tensor(sc.GEOM), sc.tensor(sc.BINNING), sc.tensor(sc.IMAGE)


def _native_backward(bg, means3D, radii, colors, semantics, scales, rotations, scale_modifier, transMat_precomp,
                     viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color, dL_dout_semantic, dL_dout_others, sh,
                     degree, campos, geomBuffer, R, binningBuffer, imageBuffer, debug):
    """Mirror of RasterizeGaussiansBackwardCUDA of the fork (DSRP/rasterize_points.cu:148-252)."""
    lib = _lib.load()
    dev = means3D.device
    P = means3D.size(0)
    H, W = dL_dout_color.size(1), dL_dout_color.size(2)
    # an empty model (P == 0) still carries sh [0,M,3]: the gradient must have that shape (the reference derives M
    # from sh.size(0) != 0 and then fails autograd's shape check)
    M = sh.size(1) if sh.dim() == 3 else 0
    S = semantics.size(1)
    f32 = dict(dtype=torch.float32, device=dev)
    dL_dmeans3D = torch.empty((P, 3), **f32)
    dL_dmeans2D = torch.empty((P, 3), **f32)
    dL_dcolors = torch.empty((P, NUM_CHANNELS), **f32)
    dL_dsemantics = torch.empty((P, S), **f32)
    dL_dopacity = torch.empty((P, 1), **f32)
    dL_dtransMat = torch.empty((P, 9), **f32)
    dL_dsh = torch.empty((P, M, 3), **f32)
    dL_dscales = torch.empty((P, 2), **f32)
    dL_drotations = torch.empty((P, 4), **f32)
    if P != 0:
        rq = _lib.require_cuda_float
        dL_dout_color, dL_dout_semantic, dL_dout_others = rq(dL_dout_color, "dL_dout_color"), \
            rq(dL_dout_semantic, "dL_dout_semantic"), rq(dL_dout_others, "dL_dout_others")
        scratch = torch.empty(lib.pgs_dsr_backward_scratch_bytes(P), dtype=torch.uint8, device=dev)
        c = lambda t: _lib.ptr(t.contiguous())  # noqa: E731
        with torch.cuda.device(dev):
            rc = lib.pgs_dsrp_backward(
                P, int(degree), int(M), int(R), c(bg), W, H, S, c(means3D), c(sh), c(colors), c(semantics), c(scales),
                float(scale_modifier), c(rotations), c(transMat_precomp), c(viewmatrix), c(projmatrix), c(campos),
                float(tan_fovx), float(tan_fovy), _lib.ptr(radii), _lib.ptr(geomBuffer), _lib.ptr(binningBuffer), int(binningBuffer.numel()),
                _lib.ptr(imageBuffer), _lib.ptr(dL_dout_color), _lib.ptr(dL_dout_semantic), _lib.ptr(dL_dout_others),
                _lib.ptr(dL_dmeans2D), _lib.ptr(scratch), _lib.ptr(dL_dopacity), _lib.ptr(dL_dcolors),
                _lib.ptr(dL_dsemantics), _lib.ptr(dL_dmeans3D), _lib.ptr(dL_dtransMat), _lib.ptr(dL_dsh),
                _lib.ptr(dL_dscales), _lib.ptr(dL_drotations), int(bool(debug)), _lib.current_stream(dev))
        _lib.check(rc, "pgs_dsrp_backward")
    return (dL_dmeans2D, dL_dcolors, dL_dsemantics, dL_dopacity, dL_dmeans3D, dL_dtransMat, dL_dsh, dL_dscales,
            dL_drotations)


class _C:
    """Namespace with the fork's pybind entry points (DSRP/ext.cpp)."""
    rasterize_gaussians = staticmethod(_native_forward)
    rasterize_gaussians_backward = staticmethod(_native_backward)
    mark_visible = staticmethod(_native_mark_visible)


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, semantics, scales, rotations, cov3Ds_precomp,
                raster_settings):
        args = (
            raster_settings.bg, means3D, colors_precomp, opacities, semantics, scales, rotations,
            raster_settings.scale_modifier, cov3Ds_precomp, raster_settings.viewmatrix, raster_settings.projmatrix,
            raster_settings.tanfovx, raster_settings.tanfovy, raster_settings.image_height,
            raster_settings.image_width, sh, raster_settings.sh_degree, raster_settings.campos,
            raster_settings.prefiltered, raster_settings.debug,
        )
        if raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(args)
            try:
                (num_rendered, color, semantic, depth, radii, geomBuffer, binningBuffer,
                 imgBuffer) = _C.rasterize_gaussians(*args)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            (num_rendered, color, semantic, depth, radii, geomBuffer, binningBuffer,
             imgBuffer) = _C.rasterize_gaussians(*args)
        ctx.raster_settings = raster_settings
        ctx.num_rendered = num_rendered
        ctx.save_for_backward(colors_precomp, semantics, means3D, scales, rotations, cov3Ds_precomp, radii, sh,
                              geomBuffer, binningBuffer, imgBuffer)
        ctx.mark_non_differentiable(radii)
        return color, semantic, radii, depth

    @staticmethod
    def backward(ctx, grad_out_color, grad_out_semantic, grad_radii, grad_depth):
        num_rendered = ctx.num_rendered
        raster_settings = ctx.raster_settings
        (colors_precomp, semantics, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer,
         imgBuffer) = ctx.saved_tensors
        args = (raster_settings.bg, means3D, radii, colors_precomp, semantics, scales, rotations,
                raster_settings.scale_modifier, cov3Ds_precomp, raster_settings.viewmatrix, raster_settings.projmatrix,
                raster_settings.tanfovx, raster_settings.tanfovy, grad_out_color, grad_out_semantic, grad_depth, sh,
                raster_settings.sh_degree, raster_settings.campos, geomBuffer, num_rendered, binningBuffer, imgBuffer,
                raster_settings.debug)
        if raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(args)
            try:
                out = _C.rasterize_gaussians_backward(*args)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise ex
        else:
            out = _C.rasterize_gaussians_backward(*args)
        (grad_means2D, grad_colors_precomp, grad_semantics, grad_opacities, grad_means3D, grad_cov3Ds_precomp, grad_sh,
         grad_scales, grad_rotations) = out
        need = ctx.needs_input_grad
        grads = (
            grad_means3D if need[0] else None,
            grad_means2D if need[1] else None,
            grad_sh if need[2] else None,
            grad_colors_precomp if need[3] else None,
            grad_opacities if need[4] else None,
            grad_semantics if need[5] else None,
            grad_scales if need[6] else None,
            grad_rotations if need[7] else None,
            grad_cov3Ds_precomp if need[8] else None,
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
        with torch.no_grad():
            raster_settings = self.raster_settings
            visible = _C.mark_visible(positions, raster_settings.viewmatrix, raster_settings.projmatrix)
        return visible

    def forward(self, means3D, means2D, opacities, semantics, shs=None, colors_precomp=None, scales=None,
                rotations=None, cov3D_precomp=None):
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

        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, semantics, scales, rotations,
                                   cov3D_precomp, raster_settings)
