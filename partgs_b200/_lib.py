"""ctypes binding of libpartgs_b200.so (the C ABI declared in include/partgs_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails
the caller gets an exception.  PyTorch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

_PKG = Path(__file__).resolve().parent
import os as _os
# PARTGS_B200_LIB: load an alternative build of the same library (kernel-tuning experiments only)
LIB_PATH = Path(_os.environ["PARTGS_B200_LIB"]) if _os.environ.get("PARTGS_B200_LIB") else _PKG / "libpartgs_b200.so"

ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_size_t, C.c_void_p)

_f32p = C.c_void_p  # device pointers travel as plain addresses
_vp = C.c_void_p


class DsrLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in (
        "geom_bytes", "geom_rec", "geom_bbox", "geom_radii", "geom_tiles_touched", "geom_point_offsets",
        "image_bytes", "image_final_T", "image_n_contrib", "image_ranges",
        "binning_bytes", "binning_keys_sorted", "binning_point_list", "binning_frag_mask", "binning_mask_stride")] + [
        ("rec_floats", C.c_int), ("tile_pixels", C.c_int)]


# name -> (restype, argtypes); must list every symbol include/partgs_b200.h declares
SIGNATURES = {
    "pgs_last_error": (C.c_char_p, []),
    "pgs_version": (C.c_int, []),
    "pgs_launch_count": (C.c_ulonglong, []),
    "pgs_timing_enable": (None, [C.c_int]),
    "pgs_timing_read": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_ulonglong), C.c_int]),
    "pgs_dsr_forward": (C.c_int, [ALLOC_FN, _vp, ALLOC_FN, _vp, ALLOC_FN, _vp, C.c_int, C.c_int, C.c_int,
                                  _f32p, C.c_int, C.c_int, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_float, _f32p,
                                  _f32p, _f32p, _f32p, _f32p, C.c_float, C.c_float, C.c_int, _f32p, _f32p, _vp,
                                  C.c_int, _vp]),
    "pgs_dsr_backward_scratch_bytes": (C.c_size_t, [C.c_int]),
    "pgs_dsr_backward": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _f32p, C.c_int, C.c_int, _f32p, _f32p,
                                   _f32p, _f32p, C.c_float, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_float,
                                   C.c_float, _vp, _vp, _vp, C.c_size_t, _vp, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p,
                                   _f32p, _f32p, _f32p, _f32p, _f32p, C.c_int, _vp]),
    "pgs_dsrp_forward": (C.c_int, [ALLOC_FN, _vp, ALLOC_FN, _vp, ALLOC_FN, _vp, C.c_int, C.c_int, C.c_int,
                                   _f32p, C.c_int, C.c_int, C.c_int, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p,
                                   C.c_float, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_float, C.c_float, C.c_int,
                                   _f32p, _f32p, _f32p, _vp, C.c_int, _vp]),
    "pgs_dsrp_backward": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _f32p, C.c_int, C.c_int, C.c_int, _f32p,
                                    _f32p, _f32p, _f32p, _f32p, C.c_float, _f32p, _f32p, _f32p, _f32p, _f32p,
                                    C.c_float, C.c_float, _vp, _vp, _vp, C.c_size_t, _vp, _f32p, _f32p, _f32p, _f32p,
                                    _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_int, _vp]),
    "pgs_mark_visible": (C.c_int, [C.c_int, _f32p, _f32p, _f32p, _vp, _vp]),
    "pgs_sq2surfel_forward": (C.c_int, [C.c_int] * 4 + [_f32p] * 7 + [_vp, _f32p, _f32p, C.c_float, C.c_float] +
                              [_f32p] * 5 + [_vp]),
    "pgs_sq2surfel_backward_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "pgs_sq2surfel_backward": (C.c_int, [C.c_int] * 4 + [_f32p] * 7 + [_vp, _f32p, _f32p, C.c_float, C.c_float] +
                               [_f32p] * 13 + [_vp, _vp]),
    "pgs_dsr_forward_blocks": (C.c_int, [ALLOC_FN, _vp, ALLOC_FN, _vp, ALLOC_FN, _vp] + [C.c_int] * 4 + [_f32p] * 7 +
                               [_vp, _f32p, _f32p, C.c_float, C.c_float, C.c_int, C.c_int, _f32p, C.c_int, C.c_int,
                                _f32p, _f32p, C.c_float, _f32p, _f32p, _f32p, C.c_float, C.c_float] + [_f32p] * 7 +
                               [_vp, C.c_int, _vp]),
    "pgs_dsr_backward_blocks_scratch_bytes": (C.c_size_t, [C.c_int] * 4),
    "pgs_dsr_backward_blocks": (C.c_int, [C.c_int] * 4 + [_f32p] * 7 + [_vp, _f32p, _f32p, C.c_float, C.c_float,
                                                                        _f32p, C.c_int, C.c_int, C.c_int, _f32p,
                                                                        C.c_int, C.c_int, _f32p, _f32p, C.c_float,
                                                                        _f32p, _f32p, _f32p, C.c_float, C.c_float,
                                                                        _vp, _vp, _vp, C.c_size_t, _vp, _f32p, _f32p,
                                                                        _f32p, _f32p, _vp, _f32p, _f32p] + [_f32p] * 7 +
                                [C.c_int, _vp]),
    "pgs_surface_maps_forward": (C.c_int, [C.c_int, C.c_int] + [_f32p] * 5 + [C.c_float] + [_f32p] * 3 + [_vp]),
    "pgs_surface_maps_backward_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "pgs_surface_maps_backward": (C.c_int, [C.c_int, C.c_int] + [_f32p] * 5 + [C.c_float] + [_f32p] * 3 +
                                  [_vp, _f32p, _vp]),
    "pgs_photometric_forward": (C.c_int, [C.c_int, C.c_int, C.c_int, _f32p, _f32p, _vp, _f32p, _vp]),
    "pgs_photometric_backward": (C.c_int, [C.c_int, C.c_int, C.c_int, _f32p, _f32p, _f32p, _f32p, C.c_float, _f32p,
                                           _vp]),
    "pgs_regularizers_forward": (C.c_int, [C.c_int, C.c_int, _f32p, _f32p, _f32p, _f32p, _f32p, _vp, _vp]),
    "pgs_regularizers_backward": (C.c_int, [C.c_int, C.c_int, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_float, C.c_float,
                                            C.c_float, _f32p, _f32p, _f32p, _f32p, _vp]),
    "pgs_adam_step": (C.c_int, [C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, C.c_double, C.c_double, C.c_double,
                                C.c_double, _vp]),
    "pgs_densify_stats": (C.c_int, [C.c_int, _vp, _f32p, _f32p, _f32p, _f32p, _vp]),
    "pgs_densify_blocks": (C.c_int, [C.c_int]),
    "pgs_densify_plan": (C.c_int, [C.c_int, _f32p, _f32p, _f32p, _f32p, C.c_double, C.c_double, C.c_double, C.c_int,
                                   C.c_double, C.c_double, _vp, _vp, _vp, _vp]),
    "pgs_densify_map": (C.c_int, [C.c_int, _vp, _vp, _vp, C.c_int, _vp, _vp, _vp]),
    "pgs_densify_gather": (C.c_int, [C.c_int, _vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp]),
    "pgs_densify_children": (C.c_int, [C.c_int, _vp, _vp, _vp, _f32p, _f32p, _f32p, _f32p, C.c_double, _f32p, _f32p,
                                       _vp]),
    "pgs_extract_maps": (C.c_int, [C.c_int, C.c_int, C.c_int, _f32p, _f32p, C.c_int, _f32p, _f32p, _f32p, _vp]),
    "pgs_knn_temp_bytes": (C.c_size_t, [C.c_int]),
    "pgs_knn_dist2": (C.c_int, [C.c_int, _f32p, _f32p, _vp, _vp]),
    "pgs_scan_temp_bytes": (C.c_size_t, [C.c_int]),
    "pgs_inclusive_scan_u32": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp]),
    "pgs_sort_temp_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "pgs_sort_pairs_u64": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp]),
    "pgs_dsr_duplicate_with_keys": (C.c_int, [C.c_int, _vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "pgs_identify_tile_ranges": (C.c_int, [C.c_int, _vp, _vp, C.c_int, _vp]),
    "pgs_peer_allreduce_slice": (C.c_int, [C.c_int, C.POINTER(C.c_void_p), C.c_size_t, C.c_size_t, C.c_int, _vp]),
    "pgs_dsr_set_capacity_hint": (None, [C.c_size_t]),
    "pgs_dsr_resolve_count": (C.c_int, [C.POINTER(C.c_int)]),
    "pgs_dsr_sorted_keys": (C.c_int, [C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_size_t, C.c_int, _vp, _vp]),
    "pgs_higher_msb": (C.c_uint32, [C.c_uint32]),
    "pgs_dsr_get_layout": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_size_t, C.POINTER(DsrLayout)]),
}

_lib = None


class PartGSError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the shared library (once) and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise PartGSError(
            f"{LIB_PATH} is missing: build it with `python -m partgs_b200.build` "
            "(there is no CPU or PyTorch fallback for the rasteriser)")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> int:
    if rc < 0:
        msg = load().pgs_last_error().decode(errors="replace")
        raise PartGSError(f"{what} failed ({rc}): {msg}")
    return rc


def ptr(t):
    """Device address of a tensor, or None (NULL) for an absent / empty tensor."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def current_stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class AllocScope:
    """Device byte buffers handed to the library through its alloc callbacks (the role resizeFunctional
    plays in the reference glue, rasterize_points.cu:31-37).  One scope per native call::

        with AllocScope(dev) as sc:
            lib.pgs_dsr_forward(ALLOC_CB, sc.GEOM, ALLOC_CB, sc.BINNING, ALLOC_CB, sc.IMAGE, ...)
        geom, binning, img = sc.tensor(sc.GEOM), ...

    A single module-level C callback serves every call (the `user` pointer carries the slot), so no
    per-call closure / reference cycle keeps a buffer alive after the call returns.  The library sizes
    the per-instance binning arena for a grow-only capacity, so the request is the same every frame and
    torch's caching allocator hands back the same block."""

    GEOM, BINNING, IMAGE = 1, 2, 3

    def __init__(self, device):
        self.device = device
        self.tensors = {}
        self.error = None

    def __enter__(self):
        _tls.scope = self
        return self

    def __exit__(self, *exc):
        _tls.scope = None
        return False

    def tensor(self, slot):
        t = self.tensors.get(slot)
        return t if t is not None else torch.empty(0, dtype=torch.uint8, device=self.device)

    def _alloc(self, nbytes, slot):
        t = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.tensors[slot] = t
        return t.data_ptr()


import threading as _threading

_tls = _threading.local()


def _alloc_cb(nbytes, user):
    scope = getattr(_tls, "scope", None)
    if scope is None:
        return None
    try:
        return scope._alloc(int(nbytes), int(user or 0))
    except Exception as ex:  # surfaced by the caller; returning NULL makes the C side fail cleanly
        scope.error = ex
        return None


ALLOC_CB = ALLOC_FN(_alloc_cb)


def on_device(t) -> bool:
    """The one place that decides whether the library may be handed a tensor's address: CUDA tensors only (there is
    no CPU fallback).  tests/emu_host.py patches this — and nothing else in the product — to run the Python layer over
    the CPU-emulated library."""
    return isinstance(t, torch.Tensor) and t.is_cuda


def require_cuda_float(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not on_device(t):
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def require_cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    """Rasteriser inputs: float32 CUDA tensors only, like the reference's `.contiguous().data<float>()`
    (rasterize_points.cu:104-126), which raises on any other dtype.  (A silent conversion here would hand the
    ORIGINAL tensor to backward and a float32 gradient to a non-float32 leaf.)"""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not on_device(t):
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name}: expected scalar type Float but found {str(t.dtype).replace('torch.', '')}")
    return t.contiguous()


STAGES = ("preprocess_fwd", "scan", "dup_keys", "sort", "tile_ranges", "render_fwd", "render_bwd", "preprocess_bwd",
          "knn", "sq_fwd", "sq_bwd", "surface_fwd", "surface_bwd", "photo_fwd", "photo_bwd")


def timing_enable(on: bool = True):
    load().pgs_timing_enable(1 if on else 0)


def timing_read(reset: bool = True):
    """{stage: (total_ms, launches)} accumulated since the last reset."""
    n = len(STAGES)
    ms = (C.c_double * n)()
    cnt = (C.c_ulonglong * n)()
    check(load().pgs_timing_read(ms, cnt, 1 if reset else 0), "pgs_timing_read")
    return {STAGES[i]: (ms[i], int(cnt[i])) for i in range(n)}
