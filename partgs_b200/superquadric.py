"""Superquadric -> surfel parameterisation of PartGS's block-level model, as one fused
autograd op backed by hand-written CUDA (pgs_sq2surfel_forward / _backward).

Mirrors the part of ``BlockGaussianModel`` that runs every training iteration
(games/block_mesh_splatting/scene/block_gaussian_model.py):
  update_alpha            :178-186   -> :func:`normalize_alpha` / BlockSurfelModel.update_alpha
  get_verts               :189-193   \
  prepare_scaling_rot     :198-256    > :func:`sq_to_surfels` (vertices, _xyz, _scaling, _rotation, opacity)
  get_opacity             :106-109   /
  get_xyz/get_scaling/get_rotation :98-104, scene/gaussian_model.py:198-204 -> properties below
There is no PyTorch fallback; CUDA tensors only.
"""
from __future__ import annotations

import math

import torch

from . import _lib


def normalize_alpha(_alpha: torch.Tensor) -> torch.Tensor:
    """BGM.update_alpha: relu + 1e-8, flatten (B,F) and normalise the barycentrics."""
    alpha = torch.relu(_alpha) + 1e-8
    alpha = alpha.flatten(start_dim=0, end_dim=1)
    return alpha / alpha.sum(dim=-1, keepdim=True)


def icosphere(level: int = 2):
    """Unit icosphere (level 2: 162 vertices, 320 faces — the topology PartGS takes from
    pytorch3d.utils.ico_sphere(2), games/block_mesh_splatting/utils/mesh.py:104-105)."""
    t = (1.0 + math.sqrt(5.0)) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    verts = [tuple(c / math.sqrt(1 + t * t) for c in p) for p in v]
    faces = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2),
             (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11),
             (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    for _ in range(level):
        cache = {}
        new_faces = []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                p = [(verts[a][i] + verts[b][i]) / 2 for i in range(3)]
                n = math.sqrt(sum(c * c for c in p))
                verts.append(tuple(c / n for c in p))
                cache[key] = len(verts) - 1
            return cache[key]

        for a, b, c in faces:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            new_faces += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        faces = new_faces
    return torch.tensor(verts, dtype=torch.float32), torch.tensor(faces, dtype=torch.int64)


class _SqToSurfels(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sq_r, sq_s, sq_t, sq_eps, sq_occ, alpha, scale_raw, eta, omega, faces, ratio, scale_min):
        lib = _lib.load()
        dev = sq_r.device
        if not _lib.on_device(sq_r):
            raise RuntimeError("sq_to_surfels needs CUDA tensors (no CPU fallback)")
        B, Vt = eta.shape
        F = faces.shape[1]
        K = alpha.shape[1]
        if alpha.shape[0] != B * F or scale_raw.numel() != B * F * K:
            raise RuntimeError("alpha must be [B*F,K,3] and scale_raw [B,F*K,1]")
        f32 = dict(dtype=torch.float32, device=dev)
        t = [x.detach().float().contiguous() for x in (sq_r, sq_s, sq_t, sq_eps, sq_occ, eta, omega)]
        faces_i = faces.to(torch.int32).contiguous()
        alpha_c = alpha.detach().float().contiguous()
        scale_c = scale_raw.detach().float().contiguous()
        P = B * F * K
        vertices = torch.empty((B, Vt, 3), **f32)
        xyz = torch.empty((P, 3), **f32)
        scaling = torch.empty((P, 2), **f32)
        rotation = torch.empty((P, 4), **f32)
        opacity = torch.empty((P, 1), **f32)
        with torch.cuda.device(dev):
            rc = lib.pgs_sq2surfel_forward(
                B, Vt, F, K, *[x.data_ptr() for x in t], faces_i.data_ptr(), alpha_c.data_ptr(), scale_c.data_ptr(),
                float(ratio), float(scale_min), vertices.data_ptr(), xyz.data_ptr(), scaling.data_ptr(),
                rotation.data_ptr(), opacity.data_ptr(), _lib.current_stream(dev))
        _lib.check(rc, "pgs_sq2surfel_forward")
        ctx.save_for_backward(*t, faces_i, alpha_c, scale_c, vertices)
        ctx.dims = (B, Vt, F, K, float(ratio), float(scale_min))
        ctx.shapes = (sq_occ.shape, alpha.shape, scale_raw.shape)
        return vertices, xyz, scaling, rotation, opacity

    @staticmethod
    def backward(ctx, d_vertices, d_xyz, d_scaling, d_rotation, d_opacity):
        lib = _lib.load()
        sq_r, sq_s, sq_t, sq_eps, sq_occ, eta, omega, faces_i, alpha_c, scale_c, vertices = ctx.saved_tensors
        B, Vt, F, K, ratio, scale_min = ctx.dims
        dev = sq_r.device
        P = B * F * K
        f32 = dict(dtype=torch.float32, device=dev)

        def g(x, shape):
            return torch.zeros(shape, **f32) if x is None else x.float().contiguous()

        d_xyz, d_scaling, d_rotation = g(d_xyz, (P, 3)), g(d_scaling, (P, 2)), g(d_rotation, (P, 4))
        d_opacity = g(d_opacity, (P, 1))
        d_vertices = None if d_vertices is None else d_vertices.float().contiguous()
        need = ctx.needs_input_grad
        d_r, d_s, d_t = torch.empty((B, 4), **f32), torch.empty((B, 3), **f32), torch.empty((B, 3), **f32)
        d_e, d_o = torch.empty((B, 2), **f32), torch.empty((B,), **f32)
        d_alpha = torch.empty((B * F, K, 3), **f32) if need[5] else None
        d_scale = torch.empty((B, F * K), **f32) if need[6] else None
        scratch = torch.empty(lib.pgs_sq2surfel_backward_scratch_bytes(B, Vt), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = lib.pgs_sq2surfel_backward(
                B, Vt, F, K, sq_r.data_ptr(), sq_s.data_ptr(), sq_t.data_ptr(), sq_eps.data_ptr(), sq_occ.data_ptr(),
                eta.data_ptr(), omega.data_ptr(), faces_i.data_ptr(), alpha_c.data_ptr(), scale_c.data_ptr(), ratio,
                scale_min, vertices.data_ptr(), d_xyz.data_ptr(), d_scaling.data_ptr(), d_rotation.data_ptr(),
                d_opacity.data_ptr(), _lib.ptr(d_vertices), d_r.data_ptr(), d_s.data_ptr(), d_t.data_ptr(),
                d_e.data_ptr(), d_o.data_ptr(), _lib.ptr(d_alpha), _lib.ptr(d_scale), scratch.data_ptr(),
                _lib.current_stream(dev))
        _lib.check(rc, "pgs_sq2surfel_backward")
        occ_shape, alpha_shape, scale_shape = ctx.shapes
        return (d_r, d_s, d_t, d_e, d_o.reshape(occ_shape),
                d_alpha.reshape(alpha_shape) if d_alpha is not None else None,
                d_scale.reshape(scale_shape) if d_scale is not None else None, None, None, None, None, None)


def sq_to_surfels(sq_r, sq_s, sq_t, sq_eps, sq_occ, alpha, scale_raw, eta, omega, faces, ratio_block_scene=0.25,
                  scale_block_min=0.2):
    """-> (vertices[B,Vt,3], xyz[P,3], _scaling[P,2] (log), _rotation[P,4], opacity[P,1]); differentiable
    w.r.t. the five superquadric parameters and (optionally) alpha / scale_raw."""
    return _SqToSurfels.apply(sq_r, sq_s, sq_t, sq_eps, sq_occ, alpha, scale_raw, eta, omega, faces,
                              ratio_block_scene, scale_block_min)


class BlockSurfelModel(torch.nn.Module):
    """Minimal host-side mirror of the rendering-relevant state of BlockGaussianModel: the
    superquadric parameters, barycentric samples and the accessor names the renderer reads
    (get_xyz / get_scaling / get_rotation / get_opacity, BGM:98-109)."""

    def __init__(self, n_blocks: int, num_splats: int, ratio_block_scene=0.25, scale_block_min=0.2, level=2,
                 device="cuda", generator: torch.Generator | None = None):
        super().__init__()
        gen = generator or torch.Generator().manual_seed(0)
        verts, faces = icosphere(level)
        B, K = n_blocks, num_splats
        Fn = faces.shape[0]
        self.ratio_block_scene, self.scale_block_min = ratio_block_scene, scale_block_min
        self.per_gs_num = Fn * K
        self.n_blocks = B
        # dataset_readers.py:323-380: random barycentric samples; BGM:171-172: eta / omega from unit vertices
        self.register_buffer("faces", faces.unsqueeze(0).repeat(B, 1, 1).to(device))
        self.register_buffer("sq_eta", torch.asin(verts[:, 1].clamp(-1, 1)).unsqueeze(0).repeat(B, 1).to(device))
        self.register_buffer("sq_omega", torch.atan2(verts[:, 0], verts[:, 2]).unsqueeze(0).repeat(B, 1).to(device))
        P = torch.nn.Parameter
        self._alpha = P(torch.rand(B, Fn, K, 3, generator=gen).to(device), requires_grad=False)
        self._scale = P(torch.full((B, Fn * K, 1), 1.0 / math.sqrt(K)).to(device), requires_grad=False)
        self.sq_r = P(torch.randn(B, 4, generator=gen).to(device))
        self.sq_s = P((math.log(0.25) + 0.3 * torch.randn(B, 3, generator=gen)).to(device))
        self.sq_t = P((torch.rand(B, 3, generator=gen) - 0.5).to(device))
        self.sq_eps = P((torch.rand(B, 2, generator=gen) * 4 - 2).to(device))
        self.sq_occ = P(torch.full((B, 1), math.log(0.7 / 0.3)).to(device))
        self.update_alpha()
        self.prepare_scaling_rot()

    def update_alpha(self):
        self.alpha = normalize_alpha(self._alpha)

    def prepare_scaling_rot(self):
        (self.vertices, self._xyz, self._scaling, self._rotation, self._opacity) = sq_to_surfels(
            self.sq_r, self.sq_s, self.sq_t, self.sq_eps, self.sq_occ, self.alpha, self._scale, self.sq_eta,
            self.sq_omega, self.faces, self.ratio_block_scene, self.scale_block_min)

    @property
    def get_xyz(self):
        return self._xyz

    @property
    def get_scaling(self):
        return torch.exp(self._scaling)

    @property
    def get_rotation(self):
        return torch.nn.functional.normalize(self._rotation)

    @property
    def get_opacity(self):
        return self._opacity


# =============================================================================================
# Block-level rasteriser: superquadric -> surfel placement fused into the rasteriser's preprocess
# =============================================================================================
class _RasterizeBlocks(torch.autograd.Function):
    """(superquadric parameters, SH) -> (color, radii, allmap) in one op.

    Equivalent to ``sq_to_surfels`` + the model accessors (exp / sigmoid) + ``GaussianRasterizer``, but the
    per-surfel centre / scale / rotation / opacity never exist in HBM: the preprocess kernels of forward and
    backward generate them from the 13 parameters per block (pgs_dsr_forward_blocks / _backward_blocks), and
    the backward pass hands the per-surfel gradients to the face -> vertex -> block reduction directly.
    """

    @staticmethod
    def forward(ctx, sq_r, sq_s, sq_t, sq_eps, sq_occ, alpha, scale_raw, shs, means2D, eta, omega, faces, ratio,
                scale_min, raster_settings, materialize):
        lib = _lib.load()
        dev = sq_r.device
        if not _lib.on_device(sq_r):
            raise RuntimeError("rasterize_blocks needs CUDA tensors (no CPU fallback)")
        B, Vt = eta.shape
        F = faces.shape[1]
        K = alpha.shape[1]
        P = B * F * K
        if alpha.shape[0] != B * F or scale_raw.numel() != P:
            raise RuntimeError("alpha must be [B*F,K,3] and scale_raw [B,F*K,1]")
        if shs.shape[0] != P:
            raise RuntimeError(f"shs must have one row per surfel ({P}), got {shs.shape[0]}")
        f32 = dict(dtype=torch.float32, device=dev)
        t = [x.detach().float().contiguous() for x in (sq_r, sq_s, sq_t, sq_eps, sq_occ, eta, omega)]
        faces_i = faces.to(torch.int32).contiguous()
        alpha_c = alpha.detach().float().contiguous()
        scale_c = scale_raw.detach().float().contiguous()
        shs_c = _lib.require_cuda_float(shs.detach(), "shs")
        rs = raster_settings
        H, W = int(rs.image_height), int(rs.image_width)
        M = shs_c.size(1)
        vertices = torch.empty((B, Vt, 3), **f32)
        out_color = torch.empty((3, H, W), **f32)
        out_others = torch.empty((7, H, W), **f32)
        radii = torch.empty((P,), dtype=torch.int32, device=dev)
        mats = [None] * 4
        if materialize:
            mats = [torch.empty((P, 3), **f32), torch.empty((P, 2), **f32), torch.empty((P, 4), **f32),
                    torch.empty((P, 1), **f32)]
        cam = [_lib.require_cuda_float(x, n) for x, n in ((rs.bg, "bg"), (rs.viewmatrix, "viewmatrix"),
                                                         (rs.projmatrix, "projmatrix"), (rs.campos, "campos"))]
        from . import diff_surfel_rasterization as _dsr
        ctx.set_materialize_grads(False)      # outputs a loss does not use arrive as None, not as zero images
        sc = _lib.AllocScope(dev)
        with torch.cuda.device(dev), sc:
            rc = lib.pgs_dsr_forward_blocks(
                _lib.ALLOC_CB, sc.GEOM, _lib.ALLOC_CB, sc.BINNING, _lib.ALLOC_CB, sc.IMAGE, B, Vt, F, K,
                *[x.data_ptr() for x in t], faces_i.data_ptr(), alpha_c.data_ptr(), scale_c.data_ptr(), float(ratio),
                float(scale_min), int(rs.sh_degree), int(M), cam[0].data_ptr(), W, H, shs_c.data_ptr(), None,
                float(rs.scale_modifier), cam[1].data_ptr(), cam[2].data_ptr(), cam[3].data_ptr(), float(rs.tanfovx),
                float(rs.tanfovy), vertices.data_ptr(), *[_lib.ptr(m) for m in mats], out_color.data_ptr(),
                out_others.data_ptr(), radii.data_ptr(), int(bool(rs.debug)) | (2 if _dsr._lazy_count else 0),
                _lib.current_stream(dev))
        if sc.error is not None:
            raise sc.error
        rendered = _lib.check(rc, "pgs_dsr_forward_blocks")
        ctx.save_for_backward(*t, faces_i, alpha_c, scale_c, vertices, shs_c, radii, *cam, sc.tensor(sc.GEOM),
                              sc.tensor(sc.BINNING), sc.tensor(sc.IMAGE))
        ctx.dims = (B, Vt, F, K, float(ratio), float(scale_min), M, H, W, rendered)
        ctx.rs = rs
        ctx.shapes = (sq_occ.shape, alpha.shape, scale_raw.shape)
        ctx.mark_non_differentiable(radii)
        outs = (out_color, radii, out_others, vertices)
        if materialize:
            for m in mats:
                ctx.mark_non_differentiable(m)
            outs = outs + tuple(mats)
        return outs

    @staticmethod
    def backward(ctx, g_color, _g_radii, g_others, g_vertices, *g_mats):
        lib = _lib.load()
        (sq_r, sq_s, sq_t, sq_eps, sq_occ, eta, omega, faces_i, alpha_c, scale_c, vertices, shs_c, radii, bg, view,
         proj, campos, geom, binning, img) = ctx.saved_tensors
        B, Vt, F, K, ratio, scale_min, M, H, W, R = ctx.dims
        rs = ctx.rs
        dev = sq_r.device
        P = B * F * K
        f32 = dict(dtype=torch.float32, device=dev)
        from . import diff_surfel_rasterization as _dsr
        if int(R) == _dsr.COUNT_PENDING and not torch.cuda.is_current_stream_capturing():
            n, overflow = _dsr.resolve_count()     # lazy forward (set_lazy_count): the count arrived long ago
            if overflow:
                raise RuntimeError(f"lazy instance count: a frame needed {n} instances, more than it was queued for; its "
                                   "outputs are invalid.  The remembered capacity has been raised — render again")
        # a loss on the returned mesh vertices joins the face -> vertex gradient sums inside the library
        g_vertices = None if g_vertices is None else _lib.require_cuda_float(g_vertices, "g_vertices")
        g_color = _lib.require_cuda_float(g_color if g_color is not None else torch.zeros((3, H, W), **f32), "g")
        g_others = _lib.require_cuda_float(g_others if g_others is not None else torch.zeros((7, H, W), **f32), "g")
        need = ctx.needs_input_grad
        # The parameter gradients of the block-level model — per-surfel SH rows and the five block parameters — are
        # carved out of ONE flat buffer like the point-level rasteriser's (diff_surfel_rasterization._carve_bucket): a
        # data-parallel caller hands out its batch bucket there and reduces it with one collective.  In accumulate mode
        # (second and later views of a batch) the SH rows are added inside the backward kernel; the block gradients
        # (13 floats per block) are produced into temporaries and added here.
        (b_sh, b_r, b_s, b_t, b_e, b_o), accumulate = _dsr._carve_bucket(
            dev, [(P, M, 3), (B, 4), (B, 3), (B, 3), (B, 2), (B,)], with_flag=True)
        d_sh = b_sh
        if accumulate:
            d_r, d_s, d_t = torch.empty((B, 4), **f32), torch.empty((B, 3), **f32), torch.empty((B, 3), **f32)
            d_e, d_o = torch.empty((B, 2), **f32), torch.empty((B,), **f32)
        else:
            d_r, d_s, d_t, d_e, d_o = b_r, b_s, b_t, b_e, b_o
        d_alpha = torch.empty((B * F, K, 3), **f32) if need[5] else None
        d_scale = torch.empty((B, F * K), **f32) if need[6] else None
        d_m2d = torch.empty((P, 3), **f32)
        d_col = torch.empty((P, 3), **f32)
        scratch = torch.empty(lib.pgs_dsr_backward_blocks_scratch_bytes(B, Vt, F, K), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = lib.pgs_dsr_backward_blocks(
                B, Vt, F, K, sq_r.data_ptr(), sq_s.data_ptr(), sq_t.data_ptr(), sq_eps.data_ptr(), sq_occ.data_ptr(),
                eta.data_ptr(), omega.data_ptr(), faces_i.data_ptr(), alpha_c.data_ptr(), scale_c.data_ptr(), ratio,
                scale_min, vertices.data_ptr(), int(rs.sh_degree), int(M), int(R), bg.data_ptr(), W, H,
                shs_c.data_ptr(), None, float(rs.scale_modifier), view.data_ptr(), proj.data_ptr(), campos.data_ptr(),
                float(rs.tanfovx), float(rs.tanfovy), radii.data_ptr(), _lib.ptr(geom), _lib.ptr(binning),
                int(binning.numel()), _lib.ptr(img), g_color.data_ptr(), g_others.data_ptr(), _lib.ptr(g_vertices),
                d_m2d.data_ptr(),
                scratch.data_ptr(), d_col.data_ptr(), d_sh.data_ptr(), d_r.data_ptr(), d_s.data_ptr(), d_t.data_ptr(),
                d_e.data_ptr(), d_o.data_ptr(), _lib.ptr(d_alpha), _lib.ptr(d_scale),
                int(bool(rs.debug)) | (2 if accumulate else 0), _lib.current_stream(dev))
        _lib.check(rc, "pgs_dsr_backward_blocks")
        if accumulate:
            torch._foreach_add_([b_r, b_s, b_t, b_e, b_o], [d_r, d_s, d_t, d_e, d_o])
            d_r, d_s, d_t, d_e, d_o = b_r, b_s, b_t, b_e, b_o
        occ_shape, alpha_shape, scale_shape = ctx.shapes
        return (d_r, d_s, d_t, d_e, d_o.reshape(occ_shape),
                d_alpha.reshape(alpha_shape) if d_alpha is not None else None,
                d_scale.reshape(scale_shape) if d_scale is not None else None,
                d_sh if need[7] else None, d_m2d if need[8] else None, None, None, None, None, None, None, None)


def blocks_bucket_numel(P: int, B: int, M: int = 16, align_elems: int = 64) -> int:
    """Floats of the gradient bucket of a block-level model (layout of _RasterizeBlocks.backward)."""
    return sum((n + align_elems - 1) // align_elems * align_elems for n in (P * M * 3, B * 4, B * 3, B * 3, B * 2, B))


def rasterize_blocks(raster_settings, sq_r, sq_s, sq_t, sq_eps, sq_occ, alpha, scale_raw, shs, eta, omega, faces,
                     means2D=None, ratio_block_scene=0.25, scale_block_min=0.2, materialize=False):
    """Render a block-level (superquadric) scene: ``(color[3,H,W], radii[P], allmap[7,H,W], vertices[B,Vt,3])``
    (+ ``xyz, _scaling, _rotation, opacity`` when ``materialize``).  ``raster_settings`` is the
    ``GaussianRasterizationSettings`` of the base rasteriser; ``means2D`` ([P,3], optional) receives the
    screen-space densification gradient like in ``GaussianRasterizer``."""
    if means2D is None:   # only its gradient slot matters (the op never reads it): no 12 B/surfel zero fill
        means2D = torch.empty((shs.shape[0], 3), dtype=torch.float32, device=shs.device)
    return _RasterizeBlocks.apply(sq_r, sq_s, sq_t, sq_eps, sq_occ, alpha, scale_raw, shs, means2D, eta, omega, faces,
                                  ratio_block_scene, scale_block_min, raster_settings, materialize)
