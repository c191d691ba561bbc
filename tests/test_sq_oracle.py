"""CPU: the superquadric->surfel oracle (oracle/sq_oracle.py) reproduces the golden vectors
generated from the reference Python (tools/make_golden_sq.py), forward and gradients."""
from pathlib import Path

import numpy as np
import pytest
import torch

GOLD = sorted((Path(__file__).resolve().parent / "golden").glob("sq2surfel_*.npz"))


def load(path):
    z = np.load(path)
    return {k: torch.from_numpy(z[k]) for k in z.files}


def test_golden_present():
    assert len(GOLD) >= 2


@pytest.mark.parametrize("path", GOLD, ids=lambda p: p.stem)
def test_oracle_matches_reference_golden(path):
    from oracle import sq_oracle
    from partgs_b200.superquadric import normalize_alpha
    z = load(path)
    p = {k: z[k].clone().requires_grad_(True) for k in ("sq_r", "sq_s", "sq_t", "sq_eps", "sq_occ", "scale_raw")}
    alpha = normalize_alpha(z["alpha_raw"])
    assert torch.allclose(alpha, z["alpha"], rtol=1e-6, atol=1e-7)
    alpha = z["alpha"].clone().requires_grad_(True)
    verts, xyz, scaling, rot, opa = sq_oracle.sq_to_surfels(
        p["sq_r"], p["sq_s"], p["sq_t"], p["sq_eps"], p["sq_occ"], alpha, p["scale_raw"], z["eta"], z["omega"],
        z["faces"].long())
    for got, want in ((verts, "vertices"), (xyz, "xyz"), (scaling, "scaling_log"), (rot, "rotation_raw"),
                      (opa, "opacity")):
        assert torch.allclose(got, z[want], rtol=1e-5, atol=1e-6), want
    assert torch.allclose(torch.exp(scaling), z["get_scaling"], rtol=1e-5, atol=1e-7)
    assert torch.allclose(torch.nn.functional.normalize(rot), z["get_rotation"], rtol=1e-5, atol=1e-6)
    loss = ((xyz * z["g_xyz"]).sum() + (scaling * z["g_scaling"]).sum() + (rot * z["g_rotation"]).sum() +
            (opa * z["g_opacity"]).sum() + (verts * z["g_vertices"]).sum())
    loss.backward()
    for k in ("sq_r", "sq_s", "sq_t", "sq_eps", "sq_occ", "scale_raw"):
        want = z["d_" + k]
        assert torch.allclose(p[k].grad, want, rtol=1e-4, atol=1e-5 * float(want.abs().max())), k
    assert torch.allclose(alpha.grad, z["d_alpha"], rtol=1e-4, atol=1e-5 * float(z["d_alpha"].abs().max()))


def test_icosphere_topology():
    from partgs_b200.superquadric import icosphere
    for level, (nv, nf) in {0: (12, 20), 1: (42, 80), 2: (162, 320)}.items():
        v, f = icosphere(level)
        assert v.shape == (nv, 3) and f.shape == (nf, 3)
        assert torch.allclose(v.norm(dim=1), torch.ones(nv), atol=1e-6)
        n = torch.linalg.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
        assert bool(((n * v[f].mean(dim=1)).sum(-1) > 0).all())     # outward-facing, consistent winding
