"""TEST INFRASTRUCTURE — checker only, never on the product path.

Thin access to the UNMODIFIED reference CUDA extensions built by oracle/Makefile into
oracle/_ref/ (ref_dsr_C, ref_dsrp_C, ref_knn_C).  Only tests/, __graft_entry__.smoke()
and bench.py's reference / cpu_baseline legs may import this module.

The reference keeps its intermediate state inside three opaque byte tensors; the
parsers below follow the fromChunk layouts of
submodules/diff-surfel-rasterization/cuda_rasterizer/rasterizer_impl.cu:155-194
(128-byte aligned arena carving, rasterizer_impl.h:24-30).
"""
from __future__ import annotations

import importlib.util
import sys
from pathlib import Path

import torch

REF_DIR = Path(__file__).resolve().parent / "_ref"
_mods = {}


def available(name: str = "ref_dsr_C") -> bool:
    return (REF_DIR / f"{name}.so").exists()


def load(name: str = "ref_dsr_C"):
    if name in _mods:
        return _mods[name]
    path = REF_DIR / f"{name}.so"
    if not path.exists():
        raise FileNotFoundError(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
    spec = importlib.util.spec_from_file_location(name, str(path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[name] = mod
    _mods[name] = mod
    return mod


def _al(x, a=128):
    return (x + a - 1) // a * a


def parse_geom(buf: torch.Tensor, P: int):
    """GeometryState::fromChunk (rasterizer_impl.cu:155-171)."""
    off = 0
    out = {}

    def take(name, nbytes, dtype, shape):
        nonlocal off
        off = _al(off)
        out[name] = buf[off:off + nbytes].view(dtype).view(*shape)
        off += nbytes

    take("depths", 4 * P, torch.float32, (P,))
    take("clamped", 3 * P, torch.bool, (P, 3))
    take("internal_radii", 4 * P, torch.int32, (P,))
    take("means2D", 8 * P, torch.float32, (P, 2))
    take("transMat", 36 * P, torch.float32, (P, 9))
    take("normal_opacity", 16 * P, torch.float32, (P, 4))
    take("rgb", 12 * P, torch.float32, (P, 3))
    take("tiles_touched", 4 * P, torch.int32, (P,))
    return out


def parse_image(buf: torch.Tensor, N: int):
    """ImageState::fromChunk (rasterizer_impl.cu:173-180)."""
    off = 0
    out = {}

    def take(name, nbytes, dtype, shape):
        nonlocal off
        off = _al(off)
        out[name] = buf[off:off + nbytes].view(dtype).view(*shape)
        off += nbytes

    take("accum_alpha", 12 * N, torch.float32, (3, N))
    take("n_contrib", 8 * N, torch.int32, (2, N))
    take("ranges", 8 * N, torch.int32, (N, 2))
    return out


def parse_binning(buf: torch.Tensor, R: int):
    """BinningState::fromChunk (rasterizer_impl.cu:182-194)."""
    off = 0
    out = {}

    def take(name, nbytes, dtype, shape):
        nonlocal off
        off = _al(off)
        out[name] = buf[off:off + nbytes].view(dtype).view(*shape)
        off += nbytes

    take("point_list", 4 * R, torch.int32, (R,))
    take("point_list_unsorted", 4 * R, torch.int32, (R,))
    take("point_list_keys", 8 * R, torch.int64, (R,))
    take("point_list_keys_unsorted", 8 * R, torch.int64, (R,))
    return out


def forward(scene, cam, bg, sh_degree=3, scale_modifier=1.0, colors_precomp=None, transMat_precomp=None,
            module="ref_dsr_C"):
    """_C.rasterize_gaussians of the reference (rasterize_points.cu:39-134)."""
    C = load(module)
    dev = scene["means3D"].device
    empty = torch.empty(0, device=dev)
    shs = scene["shs"] if colors_precomp is None else empty
    cols = colors_precomp if colors_precomp is not None else empty
    scales = scene["scales"] if transMat_precomp is None else empty
    rots = scene["rotations"] if transMat_precomp is None else empty
    tm = transMat_precomp if transMat_precomp is not None else empty
    args = (bg, scene["means3D"], cols, scene["opacities"], scales, rots, scale_modifier, tm, cam.viewmatrix,
            cam.projmatrix, cam.tanfovx, cam.tanfovy, cam.image_height, cam.image_width, shs, sh_degree, cam.campos,
            False, False)
    R, color, others, radii, geom, binning, img = C.rasterize_gaussians(*args)
    return dict(num_rendered=R, color=color, allmap=others, radii=radii, geom=geom, binning=binning, img=img,
                args=args)


def backward(fwd, scene, cam, bg, dL_dcolor, dL_dallmap, sh_degree=3, scale_modifier=1.0, module="ref_dsr_C"):
    """_C.rasterize_gaussians_backward of the reference (rasterize_points.cu:136-233)."""
    C = load(module)
    a = fwd["args"]
    args = (bg, scene["means3D"], fwd["radii"], a[2], a[4], a[5], scale_modifier, a[7], cam.viewmatrix,
            cam.projmatrix, cam.tanfovx, cam.tanfovy, dL_dcolor, dL_dallmap, a[14], sh_degree, cam.campos,
            fwd["geom"], fwd["num_rendered"], fwd["binning"], fwd["img"], False)
    names = ("means2D", "colors", "opacity", "means3D", "transMat", "sh", "scales", "rotations")
    return dict(zip(names, C.rasterize_gaussians_backward(*args)))


def mark_visible(means3D, cam, module="ref_dsr_C"):
    """_C.mark_visible of the reference (rasterize_points.cu:235-254)."""
    return load(module).mark_visible(means3D, cam.viewmatrix, cam.projmatrix)


# ---- `_part` fork (ref_dsrp_C) ------------------------------------------------------
def forward_part(scene, cam, bg, sh_degree=3, scale_modifier=1.0):
    """_C.rasterize_gaussians of the fork (DSRP/rasterize_points.cu:39-146)."""
    C = load("ref_dsrp_C")
    dev = scene["means3D"].device
    e = torch.empty(0, device=dev)
    args = (bg, scene["means3D"], e, scene["opacities"], scene["semantics"], scene["scales"], scene["rotations"],
            scale_modifier, e, cam.viewmatrix, cam.projmatrix, cam.tanfovx, cam.tanfovy, cam.image_height,
            cam.image_width, scene["shs"], sh_degree, cam.campos, False, False)
    R, color, semantic, others, radii, geom, binning, img = C.rasterize_gaussians(*args)
    return dict(num_rendered=R, color=color, semantic=semantic, allmap=others, radii=radii, geom=geom,
                binning=binning, img=img)


def backward_part(fwd, scene, cam, bg, dL_dcolor, dL_dsemantic, dL_dallmap, sh_degree=3, scale_modifier=1.0):
    """_C.rasterize_gaussians_backward of the fork (DSRP/rasterize_points.cu:148-252)."""
    C = load("ref_dsrp_C")
    dev = scene["means3D"].device
    e = torch.empty(0, device=dev)
    args = (bg, scene["means3D"], fwd["radii"], e, scene["semantics"], scene["scales"], scene["rotations"],
            scale_modifier, e, cam.viewmatrix, cam.projmatrix, cam.tanfovx, cam.tanfovy, dL_dcolor, dL_dsemantic,
            dL_dallmap, scene["shs"], sh_degree, cam.campos, fwd["geom"], fwd["num_rendered"], fwd["binning"],
            fwd["img"], False)
    names = ("means2D", "colors", "semantics", "opacity", "means3D", "transMat", "sh", "scales", "rotations")
    return dict(zip(names, C.rasterize_gaussians_backward(*args)))
