"""``from simple_knn._C import distCUDA2`` (scene/gaussian_model.py:8,
games/block_mesh_splatting/scene/two_gaussian_model.py:9 in PartGS).

distCUDA2(points[P,3] float32 cuda) -> float32 cuda [P]: mean squared distance to the 3
nearest neighbours (reference: submodules/simple-knn/spatial.cu:15-25 ->
SimpleKNN::knn, simple_knn.cu:185-221).  Implemented as an exact uniform-grid search in
hand-written CUDA behind pgs_knn_dist2; no CPU fallback.
"""
from __future__ import annotations

import torch

from .. import _lib


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    if not _lib.on_device(points):
        raise RuntimeError("points must be a CUDA tensor")
    if points.dim() != 2 or points.size(1) != 3:
        raise RuntimeError("points must have dimensions (num_points, 3)")
    pts = points.detach().float().contiguous()
    P = pts.size(0)
    means = torch.full((P,), 0.0, dtype=torch.float32, device=pts.device)
    if P == 0:
        return means
    temp = torch.empty(lib.pgs_knn_temp_bytes(P), dtype=torch.uint8, device=pts.device)
    with torch.cuda.device(pts.device):
        rc = lib.pgs_knn_dist2(P, pts.data_ptr(), means.data_ptr(), temp.data_ptr(),
                               _lib.current_stream(pts.device))
    _lib.check(rc, "pgs_knn_dist2")
    return means
