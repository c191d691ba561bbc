"""GPU: fused superquadric->surfel CUDA op vs golden vectors from the reference Python and vs
the CPU oracle at larger sizes (fp32; tolerances written per tensor)."""
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = sorted((Path(__file__).resolve().parent / "golden").glob("sq2surfel_*.npz"))


def close(a, b, rtol, name):
    b = b.to(a.device)
    scale = float(b.abs().max()) + 1e-30
    err = float((a - b).abs().max()) / scale
    assert err <= rtol, (name, err)


@pytest.mark.parametrize("path", GOLD, ids=lambda p: p.stem)
def test_cuda_matches_reference_golden(path):
    from partgs_b200.superquadric import sq_to_surfels
    z = {k: torch.from_numpy(v).to(DEV) for k, v in np.load(path).items()}
    p = {k: z[k].clone().requires_grad_(True) for k in ("sq_r", "sq_s", "sq_t", "sq_eps", "sq_occ", "scale_raw")}
    alpha = z["alpha"].clone().requires_grad_(True)
    verts, xyz, scaling, rot, opa = sq_to_surfels(p["sq_r"], p["sq_s"], p["sq_t"], p["sq_eps"], p["sq_occ"], alpha,
                                                  p["scale_raw"], z["eta"], z["omega"], z["faces"])
    close(verts, z["vertices"], 2e-6, "vertices")
    close(xyz, z["xyz"], 2e-6, "xyz")
    close(scaling, z["scaling_log"], 1e-5, "scaling")
    close(rot, z["rotation_raw"], 2e-5, "rotation")
    close(opa, z["opacity"], 1e-6, "opacity")
    loss = ((xyz * z["g_xyz"]).sum() + (scaling * z["g_scaling"]).sum() + (rot * z["g_rotation"]).sum() +
            (opa * z["g_opacity"]).sum() + (verts * z["g_vertices"]).sum())
    loss.backward()
    for k in ("sq_r", "sq_s", "sq_t", "sq_eps", "sq_occ", "scale_raw"):
        close(p[k].grad, z["d_" + k], 1e-4, "d_" + k)
    close(alpha.grad, z["d_alpha"], 1e-4, "d_alpha")


def test_cuda_matches_oracle_c1_size():
    """C1 shape: 8 superquadrics x 320 faces x 8 samples = 20480 surfels."""
    from oracle import sq_oracle
    from partgs_b200.superquadric import BlockSurfelModel
    m = BlockSurfelModel(8, 8, device=DEV, generator=torch.Generator().manual_seed(9))
    assert m.get_xyz.shape == (20480, 3)
    gen = torch.Generator().manual_seed(10)
    P = 20480
    g = [torch.randn(P, 3, generator=gen), torch.randn(P, 2, generator=gen), torch.randn(P, 4, generator=gen),
         torch.randn(P, 1, generator=gen)]
    loss = sum((o * gg.to(DEV)).sum() for o, gg in zip((m._xyz, m._scaling, m._rotation, m._opacity), g))
    loss.backward()
    names = ("sq_r", "sq_s", "sq_t", "sq_eps", "sq_occ")
    cp = {k: getattr(m, k).detach().cpu().clone().requires_grad_(True) for k in names}
    o = sq_oracle.sq_to_surfels(cp["sq_r"], cp["sq_s"], cp["sq_t"], cp["sq_eps"], cp["sq_occ"], m.alpha.cpu(),
                                m._scale.detach().cpu(), m.sq_eta.cpu(), m.sq_omega.cpu(), m.faces.cpu())
    close(m.vertices.detach(), o[0].detach(), 2e-6, "vertices")
    close(m._xyz.detach(), o[1].detach(), 2e-6, "xyz")
    close(m._scaling.detach(), o[2].detach(), 1e-5, "scaling")
    close(m._rotation.detach(), o[3].detach(), 5e-5, "rotation")
    sum((oo * gg).sum() for oo, gg in zip(o[1:], g)).backward()
    for k in names:
        close(getattr(m, k).grad, cp[k].grad, 2e-4, "d_" + k)
    # renderer-facing accessors
    assert torch.allclose(m.get_rotation.norm(dim=1), torch.ones(P, device=DEV), atol=1e-5)
    assert bool((m.get_scaling > 0).all()) and m.get_opacity.shape == (P, 1)


def test_block_surfels_render_end_to_end():
    """Generated surfels go straight into the rasteriser and gradients reach the 13 block parameters."""
    from partgs_b200 import synth
    from partgs_b200.diff_surfel_rasterization import GaussianRasterizer
    from partgs_b200.superquadric import BlockSurfelModel
    import parity_utils as pu
    m = BlockSurfelModel(8, 8, device=DEV, generator=torch.Generator().manual_seed(3))
    cam = synth.make_cameras(1, 400, 300, 1, device=DEV)[0]
    bg = torch.zeros(3, device=DEV)
    P = m.get_xyz.shape[0]
    shs = torch.zeros(P, 16, 3, device=DEV)
    shs[:, 0] = 0.5
    rast = GaussianRasterizer(pu.settings_from_cam(cam, bg))
    color, radii, allmap = rast(means3D=m.get_xyz, means2D=torch.zeros(P, 3, device=DEV, requires_grad=True),
                                opacities=m.get_opacity, shs=shs, scales=m.get_scaling, rotations=m.get_rotation)
    assert int((radii > 0).sum()) > 1000 and float(allmap[1].max()) > 0.5
    (color.sum() + allmap[0].sum()).backward()
    for k in ("sq_r", "sq_s", "sq_t", "sq_eps", "sq_occ"):
        gk = getattr(m, k).grad
        assert gk is not None and bool(torch.isfinite(gk).all()) and float(gk.abs().max()) > 0, k
