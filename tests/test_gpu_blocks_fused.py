"""GPU parity of the block-level rasteriser (superquadric -> surfel placement generated inside preprocess,
pgs_dsr_forward_blocks / pgs_dsr_backward_blocks) against the unfused composition
sq_to_surfels -> accessors -> GaussianRasterizer, which is itself pinned against the reference Python
(tests/test_gpu_sq2surfel.py, golden vectors) and the reference CUDA rasteriser (test_gpu_base_raster.py)."""
import pytest
import torch

import parity_utils as pu

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _scene(B, K, W, H, seed=11):
    from partgs_b200 import synth
    from partgs_b200.superquadric import BlockSurfelModel
    gen = torch.Generator().manual_seed(seed)
    model = BlockSurfelModel(B, K, device=DEV, generator=gen)
    P = B * model.per_gs_num
    shs = torch.zeros(P, 16, 3)
    shs[:, 0] = synth.RGB2SH(torch.rand(P, 3, generator=gen))
    shs[:, 1:] = 0.05 * torch.randn(P, 15, 3, generator=gen)
    cams = synth.make_cameras(2, W, H, synth.SEED_BASE, device=DEV)
    g = synth.upstream_grads(W, H, synth.SEED_BASE, device=DEV)
    return model, shs.to(DEV), cams, g


@pytest.mark.parametrize("B,K,W,H", [(8, 8, 400, 300), (5, 3, 333, 201)])
def test_fused_blocks_match_unfused_composition(B, K, W, H):
    from partgs_b200.diff_surfel_rasterization import GaussianRasterizer
    from partgs_b200.superquadric import rasterize_blocks, sq_to_surfels
    model, shs0, cams, g = _scene(B, K, W, H)
    bg = torch.tensor([0.1, 0.2, 0.3], device=DEV)
    names = ("sq_r", "sq_s", "sq_t", "sq_eps", "sq_occ")
    for cam in cams:
        settings = pu.settings_from_cam(cam, bg)
        # ---- unfused: materialise the surfels, then the point-level rasteriser
        pa = {n: getattr(model, n).detach().clone().requires_grad_(True) for n in names}
        shs_a = shs0.clone().requires_grad_(True)
        _, xyz, scaling, rotation, opacity = sq_to_surfels(pa["sq_r"], pa["sq_s"], pa["sq_t"], pa["sq_eps"],
                                                           pa["sq_occ"], model.alpha, model._scale, model.sq_eta,
                                                           model.sq_omega, model.faces)
        m2d_a = torch.zeros_like(xyz, requires_grad=True)
        col_a, radii_a, all_a = GaussianRasterizer(settings)(means3D=xyz, means2D=m2d_a, opacities=opacity, shs=shs_a,
                                                             scales=torch.exp(scaling), rotations=rotation)
        torch.autograd.backward([col_a, all_a], [g["color"], g["allmap"]])
        # ---- fused
        pb = {n: getattr(model, n).detach().clone().requires_grad_(True) for n in names}
        shs_b = shs0.clone().requires_grad_(True)
        m2d_b = torch.zeros_like(xyz, requires_grad=True)
        out = rasterize_blocks(settings, pb["sq_r"], pb["sq_s"], pb["sq_t"], pb["sq_eps"], pb["sq_occ"], model.alpha,
                               model._scale, shs_b, model.sq_eta, model.sq_omega, model.faces, means2D=m2d_b,
                               materialize=True)
        col_b, radii_b, all_b, verts_b, xyz_b, scaling_b, rot_b, opa_b = out
        torch.autograd.backward([col_b, all_b], [g["color"], g["allmap"]])

        # what the kernel generated is what sq_to_surfels materialises
        assert pu.rel_err(xyz_b, xyz) <= 1e-6 and pu.rel_err(scaling_b, scaling) <= 1e-6
        assert pu.rel_err(rot_b, rotation) <= 1e-6 and pu.rel_err(opa_b, opacity) <= 1e-6
        n_bad = int((radii_a != radii_b).sum())
        assert n_bad <= 3, f"{n_bad} radii differ between fused and unfused"
        if n_bad == 0 and torch.equal(xyz_b, xyz) and torch.equal(rot_b, rotation):
            assert pu.rel_err(col_b, col_a) <= pu.IMG_RTOL
            assert pu.rel_err(all_b, all_a) <= pu.IMG_RTOL
        else:
            ok, info = pu.robust_close(col_b.detach(), col_a.detach())
            assert ok, info
            ok, info = pu.robust_close(all_b.detach(), all_a.detach())
            assert ok, info
        for n in names:
            e = pu.rel_err(pb[n].grad, pa[n].grad)
            assert e <= 1e-3, (n, e)
        assert pu.rel_err(shs_b.grad, shs_a.grad) <= 1e-3
        assert pu.rel_err(m2d_b.grad, m2d_a.grad) <= 1e-3


def test_fused_blocks_without_materialisation():
    from partgs_b200.superquadric import rasterize_blocks
    model, shs, cams, g = _scene(4, 4, 320, 240)
    settings = pu.settings_from_cam(cams[0], torch.zeros(3, device=DEV))
    a = rasterize_blocks(settings, model.sq_r, model.sq_s, model.sq_t, model.sq_eps, model.sq_occ, model.alpha,
                         model._scale, shs, model.sq_eta, model.sq_omega, model.faces, materialize=True)
    b = rasterize_blocks(settings, model.sq_r, model.sq_s, model.sq_t, model.sq_eps, model.sq_occ, model.alpha,
                         model._scale, shs, model.sq_eta, model.sq_omega, model.faces)
    assert len(b) == 4 and len(a) == 8
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])


def test_vertex_loss_flows_through_the_fused_op():
    """A loss on the returned mesh vertices (pgs_dsr_backward_blocks' dL_dvertices) reaches the block parameters like
    through sq_to_surfels, alone and on top of the image gradients."""
    from partgs_b200.superquadric import rasterize_blocks, sq_to_surfels
    model, shs0, cams, g = _scene(6, 4, 320, 240)
    bg = torch.tensor([0.1, 0.2, 0.3], device=DEV)
    settings = pu.settings_from_cam(cams[0], bg)
    names = ("sq_r", "sq_s", "sq_t", "sq_eps", "sq_occ")

    def fused(with_images, gv):
        p = {n: getattr(model, n).detach().clone().requires_grad_(True) for n in names}
        color, _, allmap, verts = rasterize_blocks(settings, p["sq_r"], p["sq_s"], p["sq_t"], p["sq_eps"], p["sq_occ"],
                                                   model.alpha, model._scale, shs0, model.sq_eta, model.sq_omega,
                                                   model.faces)
        outs, gs = ([color, allmap], [g["color"], g["allmap"]]) if with_images else ([], [])
        if gv is not None:
            outs, gs = outs + [verts], gs + [gv]
        torch.autograd.backward(outs, gs)
        return {n: p[n].grad for n in names}, verts.detach()

    gv = torch.randn(6, model.sq_eta.shape[1], 3, device=DEV, generator=torch.Generator(DEV).manual_seed(3))
    pc = {n: getattr(model, n).detach().clone().requires_grad_(True) for n in names}
    verts_c = sq_to_surfels(pc["sq_r"], pc["sq_s"], pc["sq_t"], pc["sq_eps"], pc["sq_occ"], model.alpha, model._scale,
                            model.sq_eta, model.sq_omega, model.faces)[0]
    verts_c.backward(gv)
    only_v, verts = fused(False, gv)
    assert torch.equal(verts, verts_c.detach())
    for n in names[:4]:
        assert pu.rel_err(only_v[n], pc[n].grad) <= 1e-5, n
    only_img, _ = fused(True, None)
    both, _ = fused(True, gv)
    for n in names[:4]:
        assert pu.rel_err(both[n], only_img[n] + only_v[n]) <= 1e-4, n
