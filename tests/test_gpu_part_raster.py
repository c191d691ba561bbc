"""GPU parity of the `_part` rasteriser (part / semantic maps) against the unmodified reference
fork built into oracle/_ref/ref_dsrp_C.so, on identical seeded inputs (C4 shape: 16 part IDs)."""
import pytest
import torch

import parity_utils as pu

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _setup(P, W=800, H=600, S=16, views=2):
    from partgs_b200 import synth
    seed = synth.SEED_BASE + 3
    scene = synth.make_point_scene(P, seed, S=S, device=DEV)
    cams = synth.make_cameras(views, W, H, seed, device=DEV)
    bg = torch.tensor([0.05, 0.1, 0.15], device=DEV)
    g = synth.upstream_grads(W, H, synth.SEED_BASE, n_aux=8, S=S, device=DEV)
    return scene, cams, bg, g


def run_ours(scene, cam, bg, g=None):
    from partgs_b200.diff_surfel_rasterization_part import GaussianRasterizationSettings, GaussianRasterizer
    leaf = {k: scene[k].detach().clone().requires_grad_(g is not None)
            for k in ("means3D", "scales", "rotations", "opacities", "shs", "semantics")}
    means2D = torch.zeros_like(leaf["means3D"], requires_grad=g is not None)
    s = GaussianRasterizationSettings(cam.image_height, cam.image_width, cam.tanfovx, cam.tanfovy, bg, 1.0,
                                      cam.viewmatrix, cam.projmatrix, 3, cam.campos, False, False)
    color, semantic, radii, allmap = GaussianRasterizer(s)(
        means3D=leaf["means3D"], means2D=means2D, opacities=leaf["opacities"], semantics=leaf["semantics"],
        shs=leaf["shs"], scales=leaf["scales"], rotations=leaf["rotations"])
    out = dict(color=color.detach(), semantic=semantic.detach(), radii=radii, allmap=allmap.detach())
    if g is not None:
        torch.autograd.backward([color, semantic, allmap], [g["color"], g["semantic"], g["allmap"]])
        out["grads"] = dict(means3D=leaf["means3D"].grad, means2D=means2D.grad, opacity=leaf["opacities"].grad,
                            scales=leaf["scales"].grad, rotations=leaf["rotations"].grad, sh=leaf["shs"].grad,
                            semantics=leaf["semantics"].grad)
    return out


def _check_images(o, ref):
    """Part map and the 8 auxiliary maps: bit-identical.  Colour: the fork's SH polynomial is contracted differently
    by nvcc in the two builds, so the per-surfel rgb differs in the last bit -> within 1e-5 of the image range and
    element-wise within a few ulps of the accumulated value."""
    pu.assert_equal_images("semantic", o["semantic"], ref["semantic"])
    for ch in range(8):
        pu.assert_equal_images(f"allmap[{ch}]", o["allmap"][ch], ref["allmap"][ch])
    assert pu.rel_err(o["color"], ref["color"]) <= pu.IMG_RTOL
    d = (o["color"].double() - ref["color"].double()).abs()
    assert float((d - (4e-7 * ref["color"].double().abs() + 4e-7)).max()) <= 0.0, float(d.max())


@pytest.mark.parametrize("P,W,H,S", [(20_000, 400, 300, 16), (150_000, 800, 600, 16), (50_000, 333, 201, 5)])
def test_part_forward_backward_vs_reference(P, W, H, S):
    from oracle import ref_cuda
    if not ref_cuda.available("ref_dsrp_C"):
        pytest.skip("oracle/_ref/ref_dsrp_C.so not present")
    scene, cams, bg, g = _setup(P, W, H, S)
    for cam in cams:
        ref = ref_cuda.forward_part(scene, cam, bg)
        o = run_ours(scene, cam, bg, g)
        assert torch.equal(o["radii"], ref["radii"])
        assert o["allmap"].shape == (8, H, W) and o["semantic"].shape == (S, H, W)
        _check_images(o, ref)
        gref = ref_cuda.backward_part(ref, scene, cam, bg, g["color"], g["semantic"], g["allmap"])
        for k in ("means3D", "means2D", "opacity", "scales", "rotations", "sh", "semantics"):
            pu.assert_grad_close(k, o["grads"][k], gref[k].view_as(o["grads"][k]))


def test_part_c4_full_size_properties():
    """C4: 500k surfels with one-hot part IDs at 800x600: the part map is a partition of alpha."""
    scene, cams, bg, g = _setup(500_000, 800, 600, 16, views=1)
    o = run_ours(scene, cams[0], bg)
    assert bool(torch.isfinite(o["semantic"]).all())
    # one-hot semantics: sum over parts of the blended part map == accumulated alpha
    assert torch.allclose(o["semantic"].sum(0), o["allmap"][1], atol=2e-5)
    assert float(o["semantic"].min()) >= 0.0
    assert float(o["allmap"][7].max()) <= 1.0     # median fragment weight
    from oracle import ref_cuda
    if ref_cuda.available("ref_dsrp_C"):
        ref = ref_cuda.forward_part(scene, cams[0], bg)
        assert torch.equal(o["radii"], ref["radii"])
        _check_images(o, ref)
        # backward at the full C4 size
        og = run_ours(scene, cams[0], bg, g)
        gref = ref_cuda.backward_part(ref, scene, cams[0], bg, g["color"], g["semantic"], g["allmap"])
        for k in ("means3D", "means2D", "opacity", "scales", "rotations", "sh", "semantics"):
            pu.assert_grad_close(k, og["grads"][k], gref[k].view_as(og["grads"][k]))


def test_part_rejects_more_than_16_channels():
    from partgs_b200 import synth
    from partgs_b200.diff_surfel_rasterization_part import GaussianRasterizationSettings, GaussianRasterizer
    scene = synth.make_point_scene(1000, 1, S=4, device=DEV)
    cam = synth.make_cameras(1, 64, 48, 1, device=DEV)[0]
    s = GaussianRasterizationSettings(48, 64, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=DEV), 1.0, cam.viewmatrix,
                                      cam.projmatrix, 3, cam.campos, False, False)
    with pytest.raises(RuntimeError, match="semantic"):
        GaussianRasterizer(s)(means3D=scene["means3D"], means2D=scene["means3D"], opacities=scene["opacities"],
                              semantics=torch.zeros(1000, 17, device=DEV), shs=scene["shs"], scales=scene["scales"],
                              rotations=scene["rotations"])
