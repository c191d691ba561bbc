"""TEST INFRASTRUCTURE — numpy/ctypes front end of oracle/surfel_oracle.c (the CPU
restatement of the reference surfel rasteriser).  Checker only: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg, never by partgs_b200/.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "liboracle.so"
_lib = None


def build():
    if not LIB.exists() or LIB.stat().st_mtime < (HERE / "surfel_oracle.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "cpu"], check=True, capture_output=True)
    return LIB


def load():
    global _lib
    if _lib is not None:
        return _lib
    build()
    lib = C.CDLL(str(LIB))
    fp = C.c_void_p
    lib.oracle_forward.restype = C.c_void_p
    lib.oracle_forward.argtypes = [C.c_int, C.c_int, C.c_int, fp, C.c_int, C.c_int, fp, fp, fp, fp, fp, C.c_float, fp,
                                   fp, fp, fp, C.c_float, C.c_float, fp, fp, fp]
    lib.oracle_backward.restype = None
    lib.oracle_backward.argtypes = [C.c_void_p] + [fp] * 8 + [C.c_float, C.c_float] + [fp] * 11
    lib.oracle_free.argtypes = [C.c_void_p]
    lib.oracle_num_rendered.argtypes = [C.c_void_p]
    lib.oracle_num_rendered.restype = C.c_int
    lib.oracle_num_threads.restype = C.c_int
    lib.oracle_set_threads.argtypes = [C.c_int]
    lib.oracle_higher_msb.argtypes = [C.c_uint32]
    lib.oracle_higher_msb.restype = C.c_uint32
    for name in ("depths", "means2D", "transMat", "normal_opacity", "rgb", "clamped", "radii", "tiles_touched",
                 "point_offsets", "keys_unsorted", "keys", "vals_unsorted", "point_list", "ranges", "final_T",
                 "n_contrib"):
        f = getattr(lib, "oracle_" + name)
        f.restype = C.c_void_p
        f.argtypes = [C.c_void_p]
    _lib = lib
    return lib


def _np(x, dtype=np.float32):
    if x is None:
        return None
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(x, dtype=dtype)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _view(ptr, shape, dtype):
    n = int(np.prod(shape))
    if n == 0:
        return np.zeros(shape, dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape).copy()


class OracleResult:
    pass


def forward(means3D, scales, rotations, opacities, shs, viewmatrix, projmatrix, campos, W, H, tanfovx, tanfovy,
            bg=None, sh_degree=3, scale_modifier=1.0, colors_precomp=None, keep_state=False):
    """CPU forward; returns dict(color, allmap, radii, num_rendered, + intermediate state)."""
    lib = load()
    means3D = _np(means3D); scales = _np(scales); rotations = _np(rotations); opacities = _np(opacities)
    shs = _np(shs); colors_precomp = _np(colors_precomp)
    viewmatrix = _np(viewmatrix); projmatrix = _np(projmatrix); campos = _np(campos)
    bg = _np(bg) if bg is not None else np.zeros(3, np.float32)
    P = means3D.shape[0]
    M = 0 if shs is None else shs.shape[1]
    color = np.zeros((3, H, W), np.float32)
    allmap = np.zeros((7, H, W), np.float32)
    radii = np.zeros(P, np.int32)
    h = lib.oracle_forward(P, sh_degree, M, _p(bg), W, H, _p(means3D), _p(shs), _p(colors_precomp), _p(opacities),
                           _p(scales), scale_modifier, _p(rotations), _p(viewmatrix), _p(projmatrix), _p(campos),
                           tanfovx, tanfovy, _p(color), _p(allmap), _p(radii))
    R = lib.oracle_num_rendered(h)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    out = dict(color=color, allmap=allmap, radii=radii, num_rendered=R)
    out["depths"] = _view(lib.oracle_depths(h), (P,), np.float32)
    out["means2D"] = _view(lib.oracle_means2D(h), (P, 2), np.float32)
    out["transMat"] = _view(lib.oracle_transMat(h), (P, 9), np.float32)
    out["normal_opacity"] = _view(lib.oracle_normal_opacity(h), (P, 4), np.float32)
    out["rgb"] = _view(lib.oracle_rgb(h), (P, 3), np.float32)
    out["tiles_touched"] = _view(lib.oracle_tiles_touched(h), (P,), np.int32)
    out["keys_unsorted"] = _view(lib.oracle_keys_unsorted(h), (R,), np.int64)
    out["vals_unsorted"] = _view(lib.oracle_vals_unsorted(h), (R,), np.int32)
    out["keys"] = _view(lib.oracle_keys(h), (R,), np.int64)
    out["point_list"] = _view(lib.oracle_point_list(h), (R,), np.int32)
    out["ranges"] = _view(lib.oracle_ranges(h), (gx * gy, 2), np.int32)
    out["final_T"] = _view(lib.oracle_final_T(h), (3, H, W), np.float32)
    out["n_contrib"] = _view(lib.oracle_n_contrib(h), (2, H, W), np.int32)
    if keep_state:
        out["_handle"] = h
        out["_inputs"] = dict(means3D=means3D, scales=scales, rotations=rotations, shs=shs, viewmatrix=viewmatrix,
                              projmatrix=projmatrix, campos=campos, bg=bg, tanfovx=tanfovx, tanfovy=tanfovy, M=M)
    else:
        lib.oracle_free(h)
    return out


def backward(fwd, dL_dcolor, dL_dallmap):
    """CPU backward for a forward(..., keep_state=True) result; frees the state."""
    lib = load()
    h = fwd["_handle"]
    i = fwd["_inputs"]
    P = i["means3D"].shape[0]
    M = i["M"]
    dL_dcolor = _np(dL_dcolor); dL_dallmap = _np(dL_dallmap)
    g = dict(means2D=np.zeros((P, 3), np.float32), normal=np.zeros((P, 3), np.float32),
             opacity=np.zeros((P, 1), np.float32), colors=np.zeros((P, 3), np.float32),
             means3D=np.zeros((P, 3), np.float32), transMat=np.zeros((P, 9), np.float32),
             sh=np.zeros((P, M, 3), np.float32), scales=np.zeros((P, 2), np.float32),
             rotations=np.zeros((P, 4), np.float32))
    lib.oracle_backward(h, _p(i["bg"]), _p(i["means3D"]), _p(i["shs"]), _p(i["scales"]), _p(i["rotations"]),
                        _p(i["viewmatrix"]), _p(i["projmatrix"]), _p(i["campos"]), i["tanfovx"], i["tanfovy"],
                        _p(dL_dcolor), _p(dL_dallmap), _p(g["means2D"]), _p(g["normal"]), _p(g["opacity"]),
                        _p(g["colors"]), _p(g["means3D"]), _p(g["transMat"]), _p(g["sh"]), _p(g["scales"]),
                        _p(g["rotations"]))
    lib.oracle_free(h)
    fwd.pop("_handle")
    return g


def forward_scene(scene, cam, bg=None, **kw):
    return forward(scene["means3D"], scene["scales"], scene["rotations"], scene["opacities"], scene["shs"],
                   cam.viewmatrix, cam.projmatrix, cam.campos, cam.image_width, cam.image_height, cam.tanfovx,
                   cam.tanfovy, bg=bg, **kw)
