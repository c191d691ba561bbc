"""Multi-view extraction loop of render.py — SURVEY.md §8(f) rank 4.

Mirror of ``GaussianExtractor`` (utils/mesh_utils.py:57-129: ``__init__``, ``clean``, ``partmap_to_rgbmap``,
``reconstruction``, ``estimate_bounding_sphere``) for the part renderer.  The reference renders the views one by one
and, per view, runs partmap_to_rgbmap (2 S + 6 ATen kernels) and six blocking ``.cpu()`` copies — the GPU idles
during every copy and every copy waits for the GPU.  Here

* the per-view epilogue (part colours + unit normals) is one CUDA kernel (csrc/extract.cu, ``pgs_extract_maps``);
* the maps of view i travel to **pinned** host stacks ``[V, C, H, W]`` on a copy stream while view i+1 renders
  (one event per view, no blocking copy; the stacks are the tensors the reference builds with ``torch.stack``);
* views shard by camera over ranks (``rank`` / ``world``): every rank renders ``shard_views`` of the
  stack and the host stacks are gathered on rank 0 with one ``torch.distributed.gather`` per map (CPU tensors, any
  backend with CPU support; gloo in the tests).

TSDF fusion / marching cubes (open3d) stay out of scope.  No CPU / PyTorch fallback for the kernels.
"""
from __future__ import annotations

import colorsys
from functools import partial
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .dist import shard_views

MAP_NAMES = ("rgbmaps", "partrgbs", "depthmaps", "alphamaps", "normals", "depth_normals")


def fancy_palette(num: int) -> torch.Tensor:
    """``get_fancy_color(num)`` (utils/plot.py:6-23) restated without seaborn / matplotlib (absent here and un-pinned
    in the reference's environment.yml): the colour list is gold + seaborn's ``hls`` palette of 21 colours
    (hues ``linspace(0,1,22)[:-1] + 0.01``, l = 0.6, s = 0.65) rotated by three, turned into a 256-entry linearly
    interpolated lookup table (``LinearSegmentedColormap.from_list``) and sampled at ``linspace(0,1,num+1)[1:]``.
    Parity unpinned (neither package can be imported in the build container)."""
    hues = np.linspace(0, 1, 22)[:-1] + 0.01
    hues = hues % 1
    hls = [colorsys.hls_to_rgb(h, 0.6, 0.65) for h in hues]
    colors = np.array([(1.0, 215 / 255, 0.0)] + hls[3:] + hls[:2], dtype=np.float64)
    xs = np.linspace(0, 1, len(colors))
    lut_x = np.linspace(0, 1, 256)
    lut = np.stack([np.interp(lut_x, xs, colors[:, c]) for c in range(3)], axis=1)
    values = torch.linspace(0, 1, num + 1)[1:].numpy().astype(np.float64)
    idx = np.minimum((values * 256).astype(np.int64), 255)
    return torch.from_numpy(lut[idx]).float()


def extract_maps(render_semantic: Optional[torch.Tensor], rend_normal: Optional[torch.Tensor],
                 palette: Optional[torch.Tensor]):
    """(part_rgb [3,H,W] | None, unit normals [3,H,W] | None) of one view; see pgs_extract_maps."""
    lib = _lib.load()
    ref = render_semantic if render_semantic is not None else rend_normal
    if ref is None:
        return None, None
    dev = ref.device
    H, W = int(ref.shape[-2]), int(ref.shape[-1])
    part_rgb = normal_unit = None
    S = 0
    if render_semantic is not None:
        sem = _lib.require_cuda_float(render_semantic, "render_semantic")
        S = int(sem.shape[0])
        if palette is None or palette.dim() != 2 or palette.shape[0] < S + 1 or palette.shape[1] < 3:
            raise RuntimeError(f"palette must hold at least {S + 1} rows of >= 3 floats (get_fancy_color(S+1))")
        pal = _lib.require_cuda_float(palette, "palette")
        part_rgb = torch.empty((3, H, W), dtype=torch.float32, device=dev)
    if rend_normal is not None:
        nrm = _lib.require_cuda_float(rend_normal, "rend_normal")
        if tuple(nrm.shape) != (3, H, W):
            raise RuntimeError("rend_normal must be [3,H,W] with the part map's size")
        normal_unit = torch.empty((3, H, W), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.pgs_extract_maps(W, H, S, _lib.ptr(sem) if part_rgb is not None else None,
                                  _lib.ptr(pal) if part_rgb is not None else None,
                                  int(pal.stride(0)) if part_rgb is not None else 3,
                                  _lib.ptr(nrm) if normal_unit is not None else None, _lib.ptr(part_rgb),
                                  _lib.ptr(normal_unit), _lib.current_stream(dev))
    _lib.check(rc, "pgs_extract_maps")
    return part_rgb, normal_unit


def focus_point_fn(poses: np.ndarray) -> np.ndarray:
    """Nearest point to all focal axes (utils/render_utils.py:55-61)."""
    directions, origins = poses[:, :3, 2:3], poses[:, :3, 3:4]
    m = np.eye(3) - directions * np.transpose(directions, [0, 2, 1])
    mt_m = np.transpose(m, [0, 2, 1]) @ m
    return np.linalg.inv(mt_m.mean(0)) @ (mt_m @ origins).mean(0)[:, 0]


def bounding_sphere(world_view_transforms: Sequence) -> tuple:
    """(center [3] float64, radius) of ``estimate_bounding_sphere`` (utils/mesh_utils.py:131-144) from the cameras'
    ``world_view_transform`` matrices (row-vector convention, i.e. transposed w2c)."""
    c2ws = np.array([np.linalg.inv(np.asarray(torch.as_tensor(w).T.cpu().numpy())) for w in world_view_transforms])
    poses = c2ws[:, :3, :] @ np.diag([1, -1, -1, 1])
    center = focus_point_fn(poses)
    radius = np.linalg.norm(c2ws[:, :3, 3] - center, axis=-1).min()
    return center, radius


class GaussianExtractor(object):
    def __init__(self, gaussians, render, pipe, bg_color=None, *, palette: Optional[torch.Tensor] = None,
                 rank: int = 0, world: int = 1, group=None, device="cuda"):
        """Reference signature ``GaussianExtractor(gaussians, render, pipe, bg_color=None)``; the keyword-only
        extras select the palette (default: ``fancy_palette(S+1)`` at the first view) and the camera shard."""
        if bg_color is None:
            bg_color = [0, 0, 0]
        background = torch.tensor(bg_color, dtype=torch.float32, device=device)
        self.gaussians = gaussians
        self.render = partial(render, pipe=pipe, bg_color=background)
        self.palette = palette
        self.rank, self.world, self.group = int(rank), int(world), group
        self.device = torch.device(device)
        self.clean()

    @torch.no_grad()
    def partmap_to_rgbmap(self, part: torch.Tensor) -> torch.Tensor:
        return extract_maps(part, None, self._palette_for(part.shape[0]))[0]

    def _palette_for(self, S: int) -> torch.Tensor:
        if self.palette is None or self.palette.shape[0] < S + 1:
            self.palette = fancy_palette(S + 1)
        if self.palette.device != self.device:
            self.palette = self.palette.to(self.device)
        return self.palette

    @torch.no_grad()
    def clean(self):
        self.depthmaps = []
        self.alphamaps = []
        self.rgbmaps = []
        self.partrgbs = []
        self.normals = []
        self.depth_normals = []
        self.points = []
        self.viewpoint_stack = []

    def my_views(self, n_views: int) -> List[int]:
        return shard_views(n_views, self.rank, self.world)

    @torch.no_grad()
    def reconstruction(self, viewpoint_stack):
        """Render every view of this rank's shard; afterwards (on rank 0 when ``world > 1``, on the only rank
        otherwise) ``rgbmaps, partrgbs, depthmaps, alphamaps, normals, depth_normals`` are ``[V, C, H, W]`` CPU
        tensors in view order.  (The reference leaves ``normals`` a Python list and stacks the other five; a stacked
        tensor indexes the same way.)"""
        self.clean()
        self.viewpoint_stack = viewpoint_stack
        V = len(viewpoint_stack)
        mine = self.my_views(V)
        copy_stream = torch.cuda.Stream(device=self.device)
        stacks = None
        done = []
        for slot, vi in enumerate(mine):
            pkg = self.render(viewpoint_stack[vi], self.gaussians)
            sem = pkg.get("render_semantic")
            part_rgb, normal = extract_maps(sem, pkg["rend_normal"], None if sem is None else
                                            self._palette_for(sem.shape[0]))
            maps = {"rgbmaps": pkg["render"], "partrgbs": part_rgb, "depthmaps": pkg["surf_depth"],
                    "alphamaps": pkg["rend_alpha"], "normals": normal, "depth_normals": pkg["surf_normal"]}
            if stacks is None:  # pinned host stacks, allocated once the map sizes are known
                stacks = {k: torch.empty((len(mine),) + tuple(v.shape), dtype=torch.float32).pin_memory()
                          for k, v in maps.items() if v is not None}
            copy_stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(copy_stream):
                for k, v in maps.items():
                    if v is not None:
                        v = v.contiguous()
                        stacks[k][slot].copy_(v, non_blocking=True)
                        v.record_stream(copy_stream)  # keep the allocation until the copy has run
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            done.append(ev)
        for ev in done:
            ev.synchronize()
        self._collect(stacks or {}, mine, V)
        if self.rank == 0:
            self.estimate_bounding_sphere()

    def _collect(self, stacks, mine, V):
        """Local stacks -> view-ordered stacks on rank 0 (device-agnostic host logic; gloo-tested)."""
        if self.world == 1:
            for k, v in stacks.items():
                setattr(self, k, v)
            return
        import torch.distributed as dist
        # a rank without views (more ranks than cameras) has no stacks of its own: agree on the maps and their shapes
        # first, so that every rank takes part in every gather
        meta = [None] * self.world
        dist.all_gather_object(meta, {k: tuple(v.shape[1:]) for k, v in stacks.items()}, group=self.group)
        shapes = {}
        for m in meta:
            shapes.update(m)
        names = [k for k in MAP_NAMES if k in shapes]
        for k in names:
            local = stacks.get(k)
            if local is None:
                local = torch.empty((0,) + shapes[k], dtype=torch.float32)
            per_rank = [len(shard_views(V, r, self.world)) for r in range(self.world)]
            pad = max(per_rank)
            buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype)
            buf[:local.shape[0]] = local
            gathered = [torch.empty_like(buf) for _ in range(self.world)] if self.rank == 0 else None
            dist.gather(buf, gathered, dst=0, group=self.group)
            if self.rank == 0:
                out = torch.empty((V,) + tuple(local.shape[1:]), dtype=local.dtype)
                for r in range(self.world):
                    idx = shard_views(V, r, self.world)
                    if idx:
                        out[torch.as_tensor(idx)] = gathered[r][:len(idx)]
                setattr(self, k, out)

    def estimate_bounding_sphere(self):
        center, self.radius = bounding_sphere([cam.world_view_transform for cam in self.viewpoint_stack])
        self.center = torch.from_numpy(center).float().to(self.device)
        print(f"The estimated bounding radius is {self.radius:.2f}")
        print(f"Use at least {2.0 * self.radius:.2f} for depth_trunc")
