"""Generate golden vectors for the superquadric -> surfel parameterisation by running the
UNMODIFIED reference Python (BlockGaussianModel.update_alpha / prepare_scaling_rot / get_*)
on CPU.  Needs /root/reference, so it only runs in the build container; the resulting
small .npz fixtures are committed under tests/golden/.

    python tools/make_golden_sq.py
"""
from __future__ import annotations

import sys
from pathlib import Path
from unittest.mock import MagicMock

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
REF = "/root/reference"

# modules the reference imports at module scope but that are absent here
for m in ['pytorch3d', 'pytorch3d.structures', 'pytorch3d.structures.meshes', 'pytorch3d.structures.utils',
          'pytorch3d.ops', 'pytorch3d.ops.subdivide_meshes', 'pytorch3d.io', 'pytorch3d.io.utils', 'pytorch3d.loss',
          'pytorch3d.renderer', 'pytorch3d.utils', 'iopath', 'iopath.common', 'iopath.common.file_io', 'trimesh',
          'trimesh.voxel', 'trimesh.voxel.creation', 'open3d', 'plyfile', 'toolz', 'simple_knn', 'simple_knn._C',
          'matplotlib', 'matplotlib.pyplot', 'matplotlib.colors', 'imageio', 'mediapy', 'skimage', 'pandas', 'lpips',
          'PIL.ImageFile', 'easydict', 'seaborn']:
    sys.modules.setdefault(m, MagicMock())
sys.path.insert(0, REF)
from games.block_mesh_splatting.scene.block_gaussian_model import BlockGaussianModel  # noqa: E402

from partgs_b200.superquadric import icosphere  # noqa: E402  (topology only: our own icosahedron subdivision)


def make(name, B, K, level, seed):
    gen = torch.Generator().manual_seed(seed)
    verts, faces = icosphere(level)
    Fn, Vt = faces.shape[0], verts.shape[0]
    m = BlockGaussianModel(3, 0.25, 0.2)
    m.faces = faces.unsqueeze(0).repeat(B, 1, 1)
    m.sq_eta = torch.asin(verts[:, 1].clamp(-1, 1)).unsqueeze(0).repeat(B, 1)
    m.sq_omega = torch.atan2(verts[:, 0], verts[:, 2]).unsqueeze(0).repeat(B, 1)
    m.sq_eps = (torch.rand(B, 2, generator=gen) * 4 - 2).requires_grad_(True)
    m.sq_r = torch.randn(B, 4, generator=gen).requires_grad_(True)
    m.sq_s = (np.log(0.25) + 0.3 * torch.randn(B, 3, generator=gen)).float().requires_grad_(True)
    m.sq_t = (torch.rand(B, 3, generator=gen) - 0.5).requires_grad_(True)
    m.sq_occ = torch.randn(B, 1, generator=gen).requires_grad_(True)
    m._alpha = (torch.rand(B, Fn, K, 3, generator=gen) - 0.1).requires_grad_(False)   # some negatives -> relu path
    m._scale = (torch.rand(B, Fn * K, 1, generator=gen) * 0.5 + 0.1).requires_grad_(True)
    m.per_gs_num = Fn * K
    m.n_blocks = B
    m.update_alpha()
    m.alpha.requires_grad_(True)
    m.prepare_scaling_rot()
    xyz, scaling_log, rotation_raw = m._xyz, m._scaling, m._rotation
    opacity = m.get_opacity
    P = xyz.shape[0]
    g = dict(xyz=torch.randn(P, 3, generator=gen), scaling=torch.randn(P, 2, generator=gen),
             rotation=torch.randn(P, 4, generator=gen), opacity=torch.randn(P, 1, generator=gen),
             vertices=torch.randn(B, Vt, 3, generator=gen) * 0.1)
    loss = ((xyz * g["xyz"]).sum() + (scaling_log * g["scaling"]).sum() + (rotation_raw * g["rotation"]).sum() +
            (opacity * g["opacity"]).sum() + (m.vertices * g["vertices"]).sum())
    loss.backward()
    out = dict(
        faces=m.faces.numpy().astype(np.int32), eta=m.sq_eta.numpy(), omega=m.sq_omega.numpy(),
        sq_eps=m.sq_eps.detach().numpy(), sq_r=m.sq_r.detach().numpy(), sq_s=m.sq_s.detach().numpy(),
        sq_t=m.sq_t.detach().numpy(), sq_occ=m.sq_occ.detach().numpy(), alpha_raw=m._alpha.numpy(),
        alpha=m.alpha.detach().numpy(), scale_raw=m._scale.detach().numpy(),
        vertices=m.vertices.detach().numpy(), xyz=xyz.detach().numpy(), scaling_log=scaling_log.detach().numpy(),
        rotation_raw=rotation_raw.detach().numpy(), opacity=opacity.detach().numpy(),
        get_scaling=m.get_scaling.detach().numpy(), get_rotation=m.get_rotation.detach().numpy(),
        g_xyz=g["xyz"].numpy(), g_scaling=g["scaling"].numpy(), g_rotation=g["rotation"].numpy(),
        g_opacity=g["opacity"].numpy(), g_vertices=g["vertices"].numpy(),
        d_sq_eps=m.sq_eps.grad.numpy(), d_sq_r=m.sq_r.grad.numpy(), d_sq_s=m.sq_s.grad.numpy(),
        d_sq_t=m.sq_t.grad.numpy(), d_sq_occ=m.sq_occ.grad.numpy(), d_alpha=m.alpha.grad.numpy(),
        d_scale_raw=m._scale.grad.numpy())
    path = ROOT / "tests" / "golden" / f"sq2surfel_{name}.npz"
    np.savez_compressed(path, **out)
    print(path, {k: v.shape for k, v in out.items() if k in ("xyz", "vertices", "d_sq_r")},
          "finite grads:", all(np.isfinite(v).all() for k, v in out.items() if k.startswith("d_")))


if __name__ == "__main__":
    make("b3_k4_l1", B=3, K=4, level=1, seed=101)     # 42 verts / 80 faces
    make("b2_k3_l2", B=2, K=3, level=2, seed=202)     # 162 verts / 320 faces (the reference's topology)
