"""Golden vectors for densification: the UNMODIFIED reference ``TwoGaussianModel`` (training_setup -> a few Adam steps
-> add_densification_stats -> densify_and_prune) run on CPU.  Needs /root/reference -> build container only; the
small .npz fixtures are committed under tests/golden/.

The reference hard-codes ``device="cuda"`` in its tensor factories; here those factories are wrapped so that the
string "cuda" means "cpu" (no reference source is touched).  The standard-normal draws behind its
``torch.normal(mean, std)`` are recorded by re-seeding: ATen evaluates that call as
``out.normal_(0, 1).mul_(std).add_(mean)``, so ``z = torch.empty(N*Ns, 3).normal_()`` under the same seed is the
tensor the oracle / CUDA path take as an argument (the fixture check proves that equivalence on CPU).

    python tools/make_golden_densify.py
"""
from __future__ import annotations

import sys
from pathlib import Path
from types import SimpleNamespace
from unittest.mock import MagicMock

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
REF = "/root/reference"

for m in ['pytorch3d', 'pytorch3d.structures', 'pytorch3d.structures.meshes', 'pytorch3d.structures.utils',
          'pytorch3d.ops', 'pytorch3d.ops.subdivide_meshes', 'pytorch3d.io', 'pytorch3d.io.utils', 'pytorch3d.loss',
          'pytorch3d.renderer', 'pytorch3d.utils', 'iopath', 'iopath.common', 'iopath.common.file_io', 'trimesh',
          'trimesh.voxel', 'trimesh.voxel.creation', 'open3d', 'plyfile', 'toolz', 'simple_knn', 'simple_knn._C',
          'matplotlib', 'matplotlib.pyplot', 'matplotlib.colors', 'imageio', 'mediapy', 'skimage', 'pandas', 'lpips',
          'PIL.ImageFile', 'easydict', 'seaborn']:
    try:  # stub only what is really absent (torch._dynamo inspects the real pandas)
        __import__(m)
    except Exception:
        sys.modules.setdefault(m, MagicMock())
sys.path.insert(0, REF)

# "cuda" -> "cpu" in the tensor factories the reference calls with device="cuda"
for _name in ("zeros", "ones", "empty", "tensor", "rand", "randn", "full", "zeros_like", "ones_like"):
    _orig = getattr(torch, _name)

    def _wrap(*a, __orig=_orig, **kw):
        if str(kw.get("device", "")) == "cuda":
            kw["device"] = "cpu"
        return __orig(*a, **kw)

    setattr(torch, _name, _wrap)
torch.cuda.empty_cache = lambda: None

from games.block_mesh_splatting.scene.two_gaussian_model import TwoGaussianModel  # noqa: E402

NAMES = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")


def make(name, P, S, deg, seed, max_grad, min_opacity, extent, max_screen_size, percent_dense=0.01, adam_steps=3):
    gen = torch.Generator().manual_seed(seed)
    m = TwoGaussianModel(deg)
    n_rest = (deg + 1) ** 2 - 1
    from torch import nn
    m._xyz = nn.Parameter(torch.randn(P, 3, generator=gen))
    m._features_dc = nn.Parameter(torch.randn(P, 1, 3, generator=gen))
    m._features_rest = nn.Parameter(torch.randn(P, n_rest, 3, generator=gen) * 0.1)
    # log-scales straddling percent_dense*extent and 0.1*extent; opacities straddling min_opacity
    m._scaling = nn.Parameter(torch.log(extent * 10 ** (torch.rand(P, 2, generator=gen) * 3.0 - 3.2)))
    m._rotation = nn.Parameter(torch.randn(P, 4, generator=gen))
    m._opacity = nn.Parameter(torch.randn(P, 1, generator=gen) * 3.0 - 2.0)
    m._semantic = torch.rand(P, S, generator=gen)
    m.max_radii2D = torch.rand(P, generator=gen) * 40
    targs = SimpleNamespace(percent_dense=percent_dense, position_lr_init=1.6e-4, position_lr_final=1.6e-6,
                            position_lr_delay_mult=0.01, position_lr_max_steps=30000, feature_lr=0.0025,
                            opacity_lr=0.05, scaling_lr=0.005, rotation_lr=0.001)
    m.spatial_lr_scale = 1.0
    m.training_setup(targs)
    for _ in range(adam_steps):  # populate exp_avg / exp_avg_sq
        for g in m.optimizer.param_groups:
            p = g["params"][0]
            p.grad = torch.randn(p.shape, generator=gen) * 0.01
        m.optimizer.step()
    # densification statistics: some surfels never visible (denom 0 -> NaN -> 0)
    vis = torch.rand(P, generator=gen) > 0.2
    vsp = SimpleNamespace(grad=torch.randn(P, 3, generator=gen) * max_grad * 1.5)
    for _ in range(2):
        m.add_densification_stats(vsp, vis)

    before = {k: g["params"][0].detach().clone() for k, g in zip(NAMES, m.optimizer.param_groups)}
    mom_before = {k: (m.optimizer.state[g["params"][0]]["exp_avg"].clone(),
                      m.optimizer.state[g["params"][0]]["exp_avg_sq"].clone())
                  for k, g in zip(NAMES, m.optimizer.param_groups)}
    extra = dict(semantic=m._semantic.clone(), accum=m.xyz_gradient_accum.clone(), denom=m.denom.clone(),
                 max_radii2D=m.max_radii2D.clone())

    from oracle import densify_oracle
    _, split = densify_oracle.split_selection(extra["accum"].clone(), extra["denom"], before["scaling"], max_grad,
                                              extent, percent_dense)
    Ns = int(split.sum())
    torch.manual_seed(seed + 1)
    z = torch.empty(2 * Ns, 3).normal_()
    torch.manual_seed(seed + 1)
    m.densify_and_prune(max_grad, min_opacity, extent, max_screen_size)

    out = {"z": z.numpy(), "max_grad": max_grad, "min_opacity": min_opacity, "extent": extent,
           "max_screen_size": -1 if max_screen_size is None else max_screen_size, "percent_dense": percent_dense}
    for k in NAMES:
        out["in_" + k] = before[k].numpy()
        out["in_m_" + k] = mom_before[k][0].numpy()
        out["in_v_" + k] = mom_before[k][1].numpy()
    for k, v in extra.items():
        out["in_" + k] = v.numpy()
    for k, g in zip(NAMES, m.optimizer.param_groups):
        p = g["params"][0]
        out["out_" + k] = p.detach().numpy()
        out["out_m_" + k] = m.optimizer.state[p]["exp_avg"].numpy()
        out["out_v_" + k] = m.optimizer.state[p]["exp_avg_sq"].numpy()
    out["out_semantic"] = m._semantic.numpy()
    out["out_accum"] = m.xyz_gradient_accum.numpy()
    out["out_denom"] = m.denom.numpy()
    out["out_max_radii2D"] = m.max_radii2D.numpy()
    path = ROOT / "tests" / "golden" / f"densify_{name}.npz"
    np.savez_compressed(path, **out)
    print(path, "P", P, "->", out["out_xyz"].shape[0], "split-selected", Ns)


if __name__ == "__main__":
    make("p400_s4", P=400, S=4, deg=1, seed=11, max_grad=0.0002, min_opacity=0.05, extent=4.0, max_screen_size=20)
    make("p257_s1_noscreen", P=257, S=1, deg=0, seed=12, max_grad=0.0002, min_opacity=0.005, extent=2.5,
         max_screen_size=None)
