"""Views into the opaque state buffers of the base rasteriser (test / profiling aid).

The layout comes from the library itself (pgs_dsr_get_layout), mirroring what the
reference's GeometryState / ImageState / BinningState::fromChunk expose
(cuda_rasterizer/rasterizer_impl.cu:155-194).
"""
from __future__ import annotations

import ctypes as C
from contextlib import nullcontext as _nullcontext

import torch

from . import _lib


def layout(P: int, W: int, H: int, binning_bytes: int = 0) -> _lib.DsrLayout:
    lay = _lib.DsrLayout()
    _lib.check(_lib.load().pgs_dsr_get_layout(P, W, H, int(binning_bytes), C.byref(lay)), "pgs_dsr_get_layout")
    return lay


def _view(buf, off, nbytes, dtype, shape):
    return buf[off:off + nbytes].view(dtype).view(*shape)


def untile(x: torch.Tensor, W: int, H: int) -> torch.Tensor:
    """[..., ntiles*256] tile-major (8 warps x 8x4 footprints) -> [..., H, W]."""
    gx, gy = (W + 15) // 16, (H + 15) // 16
    lead = x.shape[:-1]
    # index = tile*256 + wid*32 + lane ; wid -> (wx = wid&1, wy = wid>>1) ; lane -> (lx = lane&7, ly = lane>>3)
    t = x.reshape(*lead, gy, gx, 4, 2, 4, 8)           # ty, tx, wy, wx, ly, lx
    t = t.permute(*range(len(lead)), len(lead) + 0, len(lead) + 2, len(lead) + 4, len(lead) + 1, len(lead) + 3,
                  len(lead) + 5)                        # ty, wy, ly, tx, wx, lx
    t = t.reshape(*lead, gy * 16, gx * 16)
    return t[..., :H, :W].contiguous()


def parse_state(geom: torch.Tensor, img: torch.Tensor, binning: torch.Tensor, P: int, W: int, H: int, R: int):
    lay = layout(P, W, H, binning.numel())
    gx, gy = (W + 15) // 16, (H + 15) // 16
    nt = gx * gy
    rf = lay.rec_floats
    rec = _view(geom, lay.geom_rec, 4 * rf * P, torch.float32, (P, rf))
    out = dict(
        rec=rec,
        transMat=torch.cat([rec[:, 0:3], rec[:, 4:7], rec[:, 8:11]], dim=1),
        means2D=torch.stack([rec[:, 3], rec[:, 7]], dim=1),
        normal_opacity=torch.cat([rec[:, 12:15], rec[:, 11:12]], dim=1),
        depths=rec[:, 15],
        rgb=rec[:, 16:19],
        clamped_mask=rec[:, 19].contiguous().view(torch.int32),
        bbox=_view(geom, lay.geom_bbox, 48 * P, torch.float32, (P, 12))[:, :4],
        cull=_view(geom, lay.geom_bbox, 48 * P, torch.float32, (P, 12)),
        internal_radii=_view(geom, lay.geom_radii, 4 * P, torch.int32, (P,)),
        tiles_touched=_view(geom, lay.geom_tiles_touched, 4 * P, torch.int32, (P,)),
        point_offsets=_view(geom, lay.geom_point_offsets, 4 * P, torch.int32, (P,)),
        final_T=untile(_view(img, lay.image_final_T, 4 * 3 * nt * 256, torch.float32, (3, nt * 256)), W, H),
        n_contrib=untile(_view(img, lay.image_n_contrib, 4 * 2 * nt * 256, torch.int32, (2, nt * 256)), W, H),
        ranges=_view(img, lay.image_ranges, 8 * nt, torch.int32, (nt, 2)),
    )
    if R > 0:
        # the production path sorts by depth, then by tile id, and never materialises the reference's 64-bit keys:
        # rebuild them (into the arena's spare key region) for the comparison with BinningState::point_list_keys
        out["point_list_keys"] = _view(binning, lay.binning_keys_sorted, 8 * R, torch.int64, (R,))
        with torch.cuda.device(geom.device) if geom.is_cuda else _nullcontext():
            _lib.check(_lib.load().pgs_dsr_sorted_keys(P, W, H, geom.data_ptr(), binning.data_ptr(), binning.numel(), R,
                                                       out["point_list_keys"].data_ptr(),
                                                       _lib.current_stream(geom.device) if geom.is_cuda else None),
                       "pgs_dsr_sorted_keys")
        out["point_list"] = _view(binning, lay.binning_point_list, 4 * R, torch.int32, (R,))
        # [8 warps][R]: bit l of frag_mask[w][i] = pixel (lane l of warp w's 8x4 footprint) blended instance i
        ms = lay.binning_mask_stride
        out["frag_mask"] = _view(binning, lay.binning_frag_mask, 4 * 8 * ms, torch.int32, (8, ms))[:, :R]
    return out


def duplicate_with_keys(geom: torch.Tensor, P: int, W: int, H: int, R: int, radii: torch.Tensor):
    """Run the key-duplication stage alone (for stage-wise parity tests)."""
    lib = _lib.load()
    dev = geom.device
    keys = torch.empty(R, dtype=torch.int64, device=dev)
    vals = torch.empty(R, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.pgs_dsr_duplicate_with_keys(P, geom.data_ptr(), W, H, radii.data_ptr(), keys.data_ptr(),
                                             vals.data_ptr(), _lib.current_stream(dev))
    _lib.check(rc, "pgs_dsr_duplicate_with_keys")
    return keys, vals


def sort_pairs_u64(keys: torch.Tensor, vals: torch.Tensor, end_bit: int):
    """Stable radix sort of (int64 keys, int32 values) on bits [0, end_bit)."""
    lib = _lib.load()
    dev = keys.device
    n = keys.numel()
    ka, va = keys.clone(), vals.clone()
    kb, vb = torch.empty_like(ka), torch.empty_like(va)
    temp = torch.empty(lib.pgs_sort_temp_bytes(n, end_bit), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.pgs_sort_pairs_u64(ka.data_ptr() if n else None, va.data_ptr() if n else None,
                                    kb.data_ptr() if n else None, vb.data_ptr() if n else None, n, end_bit,
                                    temp.data_ptr(), _lib.current_stream(dev))
    where = _lib.check(rc, "pgs_sort_pairs_u64")
    return (kb, vb) if where else (ka, va)


def inclusive_scan_u32(x: torch.Tensor):
    lib = _lib.load()
    dev = x.device
    n = x.numel()
    out = torch.empty_like(x)
    temp = torch.empty(lib.pgs_scan_temp_bytes(n), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.pgs_inclusive_scan_u32(x.data_ptr() if n else None, out.data_ptr() if n else None, n,
                                        temp.data_ptr(), _lib.current_stream(dev))
    _lib.check(rc, "pgs_inclusive_scan_u32")
    return out


def identify_tile_ranges(sorted_keys: torch.Tensor, ntiles: int):
    lib = _lib.load()
    dev = sorted_keys.device
    ranges = torch.empty((ntiles, 2), dtype=torch.int32, device=dev)
    L = sorted_keys.numel()
    with torch.cuda.device(dev):
        rc = lib.pgs_identify_tile_ranges(L, sorted_keys.data_ptr() if L else None, ranges.data_ptr(), ntiles,
                                          _lib.current_stream(dev))
    _lib.check(rc, "pgs_identify_tile_ranges")
    return ranges
