"""Drop-in for PartGS's ``renderer/gaussian_renderer/__init__.py::render``.

``render(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, override_color=None)`` returns the same
dict as the reference (:104-149): ``render, viewspace_points, visibility_filter, radii, rend_alpha, rend_normal,
rend_dist, surf_depth, surf_normal``.  The rasteriser call is this package's; everything the reference does to
``allmap`` afterwards (normal view->world, nan_to_num, expected depth, surf_depth blend, depth_to_normal, x alpha:
~30 ATen kernels with full-image temporaries) is one fused CUDA kernel forward and two backward
(csrc/surface_maps.cu behind pgs_surface_maps_forward / _backward).  No CPU / PyTorch fallback.
"""
from __future__ import annotations

import math

import torch

from . import _lib
from .diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer


def camera_constants(view):
    """The per-camera 3x3 matrices of utils/point_utils.py:4-20 (depths_to_points), evaluated with the reference's
    own expressions so they are the same numbers: (A, M1, M2, o) as contiguous CUDA float tensors.  They only
    depend on the camera, so they are memoised on the camera object (PartGS cameras are persistent); the memo is
    keyed on the identity and version of the two matrices."""
    wvt = view.world_view_transform
    key = (wvt.data_ptr(), wvt._version, view.full_proj_transform.data_ptr(), view.full_proj_transform._version,
           int(view.image_width), int(view.image_height))
    memo = getattr(view, "_pgs_surface_consts", None)
    if memo is not None and memo[0] == key:
        return memo[1]
    out = _camera_constants(view)
    try:
        object.__setattr__(view, "_pgs_surface_consts", (key, out))
    except Exception:
        pass
    return out


def _camera_constants(view):
    wvt = view.world_view_transform
    dev = wvt.device
    c2w = (wvt.T).inverse()
    W, H = int(view.image_width), int(view.image_height)
    ndc2pix = torch.tensor([[W / 2, 0, 0, W / 2], [0, H / 2, 0, H / 2], [0, 0, 0, 1]], dtype=torch.float32,
                           device=dev).T
    projection_matrix = c2w.T @ view.full_proj_transform
    intrins = (projection_matrix @ ndc2pix)[:3, :3].T
    return (wvt[:3, :3].contiguous().float(), intrins.inverse().T.contiguous().float(),
            c2w[:3, :3].T.contiguous().float(), c2w[:3, 3].contiguous().float())


class _SurfaceMaps(torch.autograd.Function):
    @staticmethod
    def forward(ctx, allmap, A, M1, M2, o, depth_ratio):
        lib = _lib.load()
        allmap = _lib.require_cuda_float(allmap, "allmap")
        if allmap.dim() != 3 or allmap.size(0) < 7:
            raise RuntimeError("allmap must be [7,H,W] (base rasteriser) or [8,H,W] (_part)")
        H, W = allmap.size(1), allmap.size(2)
        dev = allmap.device
        f32 = dict(dtype=torch.float32, device=dev)
        rend_normal = torch.empty((3, H, W), **f32)
        surf_depth = torch.empty((1, H, W), **f32)
        surf_normal = torch.empty((3, H, W), **f32)
        with torch.cuda.device(dev):
            rc = lib.pgs_surface_maps_forward(W, H, allmap.data_ptr(), A.data_ptr(), M1.data_ptr(), M2.data_ptr(),
                                              o.data_ptr(), float(depth_ratio), rend_normal.data_ptr(),
                                              surf_depth.data_ptr(), surf_normal.data_ptr(), _lib.current_stream(dev))
        _lib.check(rc, "pgs_surface_maps_forward")
        ctx.save_for_backward(allmap, A, M1, M2, o)
        ctx.depth_ratio = float(depth_ratio)
        return rend_normal, surf_depth, surf_normal

    @staticmethod
    def backward(ctx, g_rn, g_sd, g_sn):
        lib = _lib.load()
        allmap, A, M1, M2, o = ctx.saved_tensors
        C, H, W = allmap.shape
        dev = allmap.device
        g_allmap = torch.empty((C, H, W), dtype=torch.float32, device=dev)
        if C > 7:
            g_allmap[7:].zero_()
        g = [None if x is None else _lib.require_cuda_float(x, "grad") for x in (g_rn, g_sd, g_sn)]
        scratch = None
        if g[2] is not None:
            scratch = torch.empty(lib.pgs_surface_maps_backward_scratch_bytes(W, H), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = lib.pgs_surface_maps_backward(W, H, allmap.data_ptr(), A.data_ptr(), M1.data_ptr(), M2.data_ptr(),
                                               o.data_ptr(), ctx.depth_ratio, _lib.ptr(g[0]), _lib.ptr(g[1]),
                                               _lib.ptr(g[2]), _lib.ptr(scratch), g_allmap.data_ptr(),
                                               _lib.current_stream(dev))
        _lib.check(rc, "pgs_surface_maps_backward")
        return g_allmap, None, None, None, None, None


def surface_maps(allmap, viewpoint_camera, depth_ratio: float, constants=None):
    """allmap [7|8,H,W] -> dict(rend_alpha, rend_normal, rend_dist, surf_depth, surf_normal) exactly as the
    reference's render() derives them (renderer/gaussian_renderer/__init__.py:110-147)."""
    A, M1, M2, o = constants if constants is not None else camera_constants(viewpoint_camera)
    rend_normal, surf_depth, surf_normal = _SurfaceMaps.apply(allmap, A, M1, M2, o, depth_ratio)
    return {"rend_alpha": allmap[1:2], "rend_normal": rend_normal, "rend_dist": allmap[6:7], "surf_depth": surf_depth,
            "surf_normal": surf_normal}


def render(viewpoint_camera, pc, pipe, bg_color: torch.Tensor, scaling_modifier=1.0, override_color=None):
    """Mirror of the reference render() (renderer/gaussian_renderer/__init__.py:9-149)."""
    screenspace_points = torch.zeros_like(pc.get_xyz, dtype=pc.get_xyz.dtype, requires_grad=True) + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass
    raster_settings = GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height), image_width=int(viewpoint_camera.image_width),
        tanfovx=math.tan(viewpoint_camera.FoVx * 0.5), tanfovy=math.tan(viewpoint_camera.FoVy * 0.5), bg=bg_color,
        scale_modifier=scaling_modifier, viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform, sh_degree=pc.active_sh_degree,
        campos=viewpoint_camera.camera_center, prefiltered=False, debug=False)
    rasterizer = GaussianRasterizer(raster_settings=raster_settings)
    if getattr(pipe, "compute_cov3D_python", False):
        raise NotImplementedError("compute_cov3D_python: the reference's own producer of cov3D_precomp is dead code "
                                  "(SURVEY.md §4); pass scales / rotations")
    shs, colors_precomp = (pc.get_features, None) if override_color is None else (None, override_color)
    rendered_image, radii, allmap = rasterizer(means3D=pc.get_xyz, means2D=screenspace_points, shs=shs,
                                               colors_precomp=colors_precomp, opacities=pc.get_opacity,
                                               scales=pc.get_scaling, rotations=pc.get_rotation, cov3D_precomp=None)
    rets = {"render": rendered_image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0,
            "radii": radii}
    rets.update(surface_maps(allmap, viewpoint_camera, pipe.depth_ratio))
    return rets


def render_part(viewpoint_camera, pc, pipe, bg_color: torch.Tensor, scaling_modifier=1.0, override_color=None):
    """Mirror of the reference render_part() (renderer/gaussian_renderer_2d/__init__.py:11-147): the `_part`
    rasteriser (semantics in, ``render_semantic`` out, 8-channel allmap) followed by the same surface-map
    post-processing as render()."""
    from .diff_surfel_rasterization_part import GaussianRasterizationSettings as PartSettings
    from .diff_surfel_rasterization_part import GaussianRasterizer as PartRasterizer
    screenspace_points = torch.zeros_like(pc.get_xyz, dtype=pc.get_xyz.dtype, requires_grad=True) + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass
    raster_settings = PartSettings(
        image_height=int(viewpoint_camera.image_height), image_width=int(viewpoint_camera.image_width),
        tanfovx=math.tan(viewpoint_camera.FoVx * 0.5), tanfovy=math.tan(viewpoint_camera.FoVy * 0.5), bg=bg_color,
        scale_modifier=scaling_modifier, viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform, sh_degree=pc.active_sh_degree,
        campos=viewpoint_camera.camera_center, prefiltered=False, debug=False)
    rasterizer = PartRasterizer(raster_settings=raster_settings)
    if getattr(pipe, "compute_cov3D_python", False):
        raise NotImplementedError("compute_cov3D_python: transMat_precomp is unusable in the reference's `_part` fork "
                                  "(DESIGN.md §8); pass scales / rotations")
    shs, colors_precomp = (pc.get_features, None) if override_color is None else (None, override_color)
    means3D = pc.get_xyz
    try:
        means3D.retain_grad()
    except Exception:
        pass
    rendered_image, semantic, radii, allmap = rasterizer(
        means3D=means3D, means2D=screenspace_points, shs=shs, colors_precomp=colors_precomp,
        opacities=pc.get_opacity, semantics=pc.get_semantic, scales=pc.get_scaling, rotations=pc.get_rotation,
        cov3D_precomp=None)
    rets = {"render": rendered_image, "render_semantic": semantic, "viewspace_points": screenspace_points,
            "visibility_filter": radii > 0, "radii": radii}
    rets.update(surface_maps(allmap, viewpoint_camera, pipe.depth_ratio))
    return rets
