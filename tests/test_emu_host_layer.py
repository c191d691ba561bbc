"""The product's Python host layer on CPU tensors over the emulated library (tests/emu_host.py): the ctypes argument
order, autograd plumbing and buffer handling of the public entry points are exercised without a GPU —
GaussianRasterizer (both forks) through autograd against the C oracle / the emulated-reference golden vectors,
render() / render_part(), the fused losses, densification, the extraction epilogue."""
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from emu_host import emulated_host  # noqa: F401  (fixture)
from oracle import cpu_oracle, extract_oracle, loss_oracle
from partgs_b200 import synth
from test_emu_raster import rel

GOLD = Path(__file__).parent / "golden"


def _small_scene(P=300, W=48, H=32, seed=1, S=0):
    scene = synth.make_point_scene(P, seed=seed, S=S, device="cpu")
    scene["scales"] = scene["scales"] * 3.0
    cam = synth.make_cameras(1, W, H, seed=seed + 10, device="cpu")[0]
    return scene, cam


def test_rasterizer_autograd_path_matches_the_c_oracle(emulated_host):
    from partgs_b200.diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    scene, cam = _small_scene()
    W, H = cam.image_width, cam.image_height
    g = synth.upstream_grads(W, H, 5, device="cpu")
    leaf = {k: scene[k].clone().requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "shs")}
    means2D = torch.zeros_like(leaf["means3D"], requires_grad=True)
    settings = GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=torch.zeros(3), scale_modifier=1.0,
        viewmatrix=cam.viewmatrix, projmatrix=cam.projmatrix, sh_degree=3, campos=cam.campos, prefiltered=False,
        debug=True)
    rast = GaussianRasterizer(settings)
    color, radii, allmap = rast(means3D=leaf["means3D"], means2D=means2D, opacities=leaf["opacities"], shs=leaf["shs"],
                                scales=leaf["scales"], rotations=leaf["rotations"])
    torch.autograd.backward([color, allmap], [g["color"], g["allmap"]])
    f = cpu_oracle.forward_scene(scene, cam, keep_state=True)
    gr = cpu_oracle.backward(f, g["color"], g["allmap"])
    assert np.array_equal(radii.numpy(), f["radii"])
    assert rel(color.detach().numpy(), f["color"]) <= 2e-5 and rel(allmap.detach().numpy(), f["allmap"]) <= 2e-5
    for k, ok in (("means3D", "means3D"), ("scales", "scales"), ("rotations", "rotations"), ("opacities", "opacity"),
                  ("shs", "sh")):
        assert rel(leaf[k].grad.numpy(), gr[ok]) <= 2e-4, k
    assert rel(means2D.grad.numpy()[:, :2], gr["means2D"][:, :2]) <= 2e-4
    vis = rast.markVisible(scene["means3D"])
    assert vis.dtype == torch.bool and int(vis.sum()) >= int((radii > 0).sum())
    with pytest.raises(Exception, match="one of either SHs"):
        rast(means3D=leaf["means3D"], means2D=means2D, opacities=leaf["opacities"], shs=None, colors_precomp=None,
             scales=leaf["scales"], rotations=leaf["rotations"])


def test_part_rasterizer_autograd_path_matches_reference_golden(emulated_host):
    from partgs_b200.diff_surfel_rasterization_part import GaussianRasterizationSettings, GaussianRasterizer
    z = dict(np.load(GOLD / "ref_emu_part_p300_s5_48x32.npz"))
    t = lambda k: torch.from_numpy(z[k])
    leaf = {k: t(k).clone().requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "shs",
                                                            "semantics")}
    means2D = torch.zeros_like(leaf["means3D"], requires_grad=True)
    tx, ty = (float(v) for v in z["tanfov"])
    settings = GaussianRasterizationSettings(
        image_height=int(z["H"]), image_width=int(z["W"]), tanfovx=tx, tanfovy=ty, bg=t("bg"), scale_modifier=1.0,
        viewmatrix=t("viewmatrix"), projmatrix=t("projmatrix"), sh_degree=int(z["degree"]), campos=t("campos"),
        prefiltered=False, debug=True)
    color, semantic, radii, allmap = GaussianRasterizer(settings)(
        means3D=leaf["means3D"], means2D=means2D, opacities=leaf["opacities"], semantics=leaf["semantics"],
        shs=leaf["shs"], scales=leaf["scales"], rotations=leaf["rotations"])
    torch.autograd.backward([color, semantic, allmap], [t("g_color"), t("g_semantic"), t("g_allmap")])
    assert np.array_equal(radii.numpy(), z["radii"])
    assert rel(color.detach().numpy(), z["color"]) <= 2e-5 and rel(semantic.detach().numpy(), z["semantic"]) <= 2e-5
    for k, ok in (("means3D", "means3D"), ("scales", "scales"), ("rotations", "rotations"), ("opacities", "opacity"),
                  ("shs", "sh"), ("semantics", "semantics")):
        assert rel(leaf[k].grad.numpy(), z["d_" + ok]) <= 2e-4, k


def _pc(scene):
    return SimpleNamespace(get_xyz=scene["means3D"], get_opacity=scene["opacities"], get_scaling=scene["scales"],
                           get_rotation=scene["rotations"], get_features=scene["shs"],
                           get_semantic=scene.get("semantics"), active_sh_degree=3)


def test_render_mirrors_and_fused_losses(emulated_host):
    from partgs_b200.losses import geometric_regularizers, photometric_loss
    from partgs_b200.renderer import render, render_part
    scene, cam = _small_scene(P=150, W=32, H=16, S=4)
    pipe = SimpleNamespace(depth_ratio=0.5, compute_cov3D_python=False, convert_SHs_python=False)
    scene["means3D"].requires_grad_(True)
    r = render(cam, _pc(scene), pipe, torch.zeros(3))
    assert set(r) == {"render", "viewspace_points", "visibility_filter", "radii", "rend_alpha", "rend_normal",
                      "rend_dist", "surf_depth", "surf_normal"}
    H, W = r["render"].shape[1:]
    gt = torch.rand(3, H, W, generator=torch.Generator().manual_seed(2))
    mask = (torch.rand(H, W, generator=torch.Generator().manual_seed(3)) > 0.5).float()
    loss = photometric_loss(r["render"], gt, 0.2) + geometric_regularizers(r, mask, 0.1, 0.05, 100.0)
    want = loss_oracle.photometric_loss(r["render"].detach(), gt, 0.2) + loss_oracle.geometric_regularizers(
        r["rend_alpha"].detach(), mask, r["rend_dist"].detach(), r["rend_normal"].detach(), r["surf_normal"].detach(),
        0.1, 0.05, 100.0)[0]
    assert abs(float(loss.detach()) - float(want)) <= 1e-5 * abs(float(want))
    loss.backward()
    assert scene["means3D"].grad is not None and bool(torch.isfinite(scene["means3D"].grad).all())
    assert float(scene["means3D"].grad.abs().max()) > 0 and r["viewspace_points"].grad is not None
    rp = render_part(cam, _pc(scene), pipe, torch.zeros(3))
    assert set(rp) == set(r) | {"render_semantic"} and rp["render_semantic"].shape == (4, H, W)
    assert bool(torch.isfinite(rp["surf_normal"]).all()) and rp["surf_depth"].shape == (1, H, W)


@pytest.mark.parametrize("path", sorted(GOLD.glob("densify_*.npz")), ids=lambda p: p.stem)
def test_densify_python_layer_reproduces_reference(emulated_host, path):
    from partgs_b200 import densify
    z = np.load(path)
    t = {k: torch.from_numpy(z[k]) for k in z.files if z[k].ndim > 0}
    mss = None if float(z["max_screen_size"]) < 0 else float(z["max_screen_size"])
    names = densify.PARAM_NAMES
    p, m, s, info = densify.densify_and_prune(
        {k: t["in_" + k] for k in names}, {k: (t["in_m_" + k], t["in_v_" + k]) for k in names}, t["in_semantic"],
        t["in_accum"], t["in_denom"], float(z["max_grad"]), float(z["min_opacity"]), float(z["extent"]), mss,
        float(z["percent_dense"]), N=2, z=t["z"])
    c = info["n_kept"] + info["n_clones"]
    assert info["n_out"] == t["out_xyz"].shape[0]
    for k in names:
        if k in ("xyz", "scaling"):
            assert torch.equal(p[k][:c], t["out_" + k][:c]), k
            assert float((p[k][c:] - t["out_" + k][c:]).abs().max()) <= 2e-5 * float(t["out_" + k].abs().max()), k
        else:
            assert torch.equal(p[k], t["out_" + k]), k
        assert torch.equal(m[k][0], t["out_m_" + k]) and torch.equal(m[k][1], t["out_v_" + k]), k
    assert torch.equal(s, t["out_semantic"])


def test_densify_model_level_drop_in(emulated_host):
    from partgs_b200 import densify
    z = np.load(GOLD / "densify_p400_s4.npz")
    t = {k: torch.from_numpy(z[k]) for k in z.files if z[k].ndim > 0}
    model = SimpleNamespace(percent_dense=float(z["percent_dense"]), _semantic=t["in_semantic"],
                            xyz_gradient_accum=t["in_accum"], denom=t["in_denom"], max_radii2D=t["in_max_radii2D"])
    groups = []
    for k, attr in densify._MODEL_ATTR.items():
        prm = torch.nn.Parameter(t["in_" + k].clone())
        setattr(model, attr, prm)
        groups.append({"params": [prm], "lr": 1e-3, "name": k})
    model.optimizer = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
    for g in model.optimizer.param_groups:
        model.optimizer.state[g["params"][0]] = {"step": torch.tensor(3.0), "exp_avg": t["in_m_" + g["name"]].clone(),
                                                 "exp_avg_sq": t["in_v_" + g["name"]].clone()}
    # the draws: re-seed like tools/make_golden_densify.py did (seed + 1 = 12), CPU generator here
    torch.manual_seed(12)
    info = densify.densify_and_prune_model(model, float(z["max_grad"]), float(z["min_opacity"]), float(z["extent"]),
                                           float(z["max_screen_size"]))
    n = info["n_out"]
    assert n == t["out_xyz"].shape[0]
    for g in model.optimizer.param_groups:
        k, prm = g["name"], g["params"][0]
        assert prm is getattr(model, densify._MODEL_ATTR[k])
        assert float((prm.detach() - t["out_" + k]).abs().max()) <= 2e-5 * float(t["out_" + k].abs().max()), k
        st = model.optimizer.state[prm]
        assert torch.equal(st["exp_avg"], t["out_m_" + k]) and float(st["step"]) == 3.0
    assert torch.equal(model._semantic, t["out_semantic"])
    assert model.xyz_gradient_accum.shape == (n, 1) and model.max_radii2D.shape == (n,) and not model.denom.any()


def test_extract_maps_python_layer(emulated_host):
    from partgs_b200.extract import extract_maps
    z = np.load(GOLD / "extract_maps.npz")
    part, pal = torch.from_numpy(z["b_part"]), torch.from_numpy(z["b_palette"])
    nrm = torch.randn(3, *part.shape[1:], generator=torch.Generator().manual_seed(4))
    rgb, unit = extract_maps(part, nrm, pal)
    assert torch.equal(rgb, torch.from_numpy(z["b_rgb"]))
    assert float((unit - extract_oracle.unit_normals(nrm)).abs().max()) <= 1e-6
    rgb2, none = extract_maps(part, None, torch.cat([pal, torch.ones(len(pal), 1)], 1))
    assert none is None and torch.equal(rgb2, rgb)
    with pytest.raises(RuntimeError, match="palette"):
        extract_maps(part, None, pal[:3])


def test_empty_model_short_circuits_like_the_reference(emulated_host):
    """P == 0: zero images, no state, gradients of the right (empty) shapes (rasterize_points.cu:85-99,197)."""
    from partgs_b200.diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    cam = synth.make_cameras(1, 20, 12, seed=3, device="cpu")[0]
    settings = GaussianRasterizationSettings(
        image_height=12, image_width=20, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=torch.ones(3), scale_modifier=1.0,
        viewmatrix=cam.viewmatrix, projmatrix=cam.projmatrix, sh_degree=3, campos=cam.campos, prefiltered=False,
        debug=False)
    e = lambda *s: torch.zeros(*s, requires_grad=True)
    m3, m2, op, sh, sc, rot = e(0, 3), e(0, 3), e(0, 1), e(0, 16, 3), e(0, 2), e(0, 4)
    color, radii, allmap = GaussianRasterizer(settings)(means3D=m3, means2D=m2, opacities=op, shs=sh, scales=sc,
                                                        rotations=rot)
    assert color.shape == (3, 12, 20) and allmap.shape == (7, 12, 20) and radii.shape == (0,)
    assert not color.any() and not allmap.any()
    (color.sum() + allmap.sum()).backward()
    assert m3.grad.shape == (0, 3) and sh.grad.shape == (0, 16, 3)


def test_precomputed_transform_path_matches_the_emulated_reference(emulated_host):
    """`cov3D_precomp` ([P,9] ray-splat transforms instead of scales / rotations): forward and backward against the
    UNMODIFIED reference source run on the emulator (tests/golden/ref_emu_base_precompT_*.npz)."""
    from partgs_b200.diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    z = dict(np.load(GOLD / "ref_emu_base_precompT_p300_48x32.npz"))
    t = lambda k: torch.from_numpy(z[k])
    leaf = {k: t(k).clone().requires_grad_(True) for k in ("means3D", "opacities", "shs", "transMat_precomp")}
    means2D = torch.zeros_like(leaf["means3D"], requires_grad=True)
    tx, ty = (float(v) for v in z["tanfov"])
    settings = GaussianRasterizationSettings(
        image_height=int(z["H"]), image_width=int(z["W"]), tanfovx=tx, tanfovy=ty, bg=t("bg"), scale_modifier=1.0,
        viewmatrix=t("viewmatrix"), projmatrix=t("projmatrix"), sh_degree=int(z["degree"]), campos=t("campos"),
        prefiltered=False, debug=True)
    color, radii, allmap = GaussianRasterizer(settings)(
        means3D=leaf["means3D"], means2D=means2D, opacities=leaf["opacities"], shs=leaf["shs"],
        cov3D_precomp=leaf["transMat_precomp"])
    torch.autograd.backward([color, allmap], [t("g_color"), t("g_allmap")])
    assert np.array_equal(radii.numpy(), z["radii"])
    assert rel(color.detach().numpy(), z["color"]) <= 2e-5
    for ch in range(7):
        assert rel(allmap.detach().numpy()[ch], z["allmap"][ch]) <= (3e-3 if ch == 6 else 5e-5), ch
    assert rel(leaf["transMat_precomp"].grad.numpy(), z["d_transMat"]) <= 2e-4
    assert rel(leaf["opacities"].grad.numpy(), z["d_opacity"]) <= 2e-4
    assert rel(leaf["shs"].grad.numpy(), z["d_sh"]) <= 2e-4
    assert rel(leaf["means3D"].grad.numpy(), z["d_means3D"]) <= 2e-4
    assert rel(means2D.grad.numpy()[:, :2], z["d_means2D"][:, :2]) <= 2e-4


class _FakeStream:
    def __init__(self, *a, **k):
        pass

    def wait_stream(self, other):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class _FakeEvent:
    def __init__(self, *a, **k):
        pass

    def record(self, stream=None):
        pass

    def synchronize(self):
        pass


def test_extraction_loop_equals_view_by_view_rendering(emulated_host, monkeypatch):
    """GaussianExtractor.reconstruction over render_part on the emulator (CUDA streams / events / pinned memory
    replaced by no-ops: the loop's bookkeeping, shapes and view order are what is checked here)."""
    from partgs_b200.extract import GaussianExtractor, fancy_palette
    from partgs_b200.renderer import render_part
    monkeypatch.setattr(torch.cuda, "Stream", _FakeStream)
    monkeypatch.setattr(torch.cuda, "Event", _FakeEvent)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: _FakeStream())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: s)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.Tensor, "record_stream", lambda self, s: None)
    scene = synth.make_point_scene(60, seed=4, S=3, device="cpu")
    scene["scales"] = scene["scales"] * 4.0
    cams = synth.make_cameras(2, 16, 16, seed=9, device="cpu")
    pipe = SimpleNamespace(depth_ratio=1.0, compute_cov3D_python=False, convert_SHs_python=False)
    pc = _pc(scene)
    outs = {}
    ex = GaussianExtractor(pc, render_part, pipe, bg_color=[1, 1, 1], device="cpu")
    monkeypatch.setattr(ex, "_collect", lambda stacks, mine, V: outs.__setitem__(0, (stacks, mine)))
    monkeypatch.setattr(ex, "estimate_bounding_sphere", lambda: None)
    ex.reconstruction(cams)
    assert outs[0][1] == [0, 1]
    # camera sharding of the same loop (the gather itself is gloo-tested in test_extract.py)
    assert GaussianExtractor(pc, render_part, pipe, rank=1, world=2, device="cpu").my_views(5) == [1, 3]
    pal = fancy_palette(4)
    bg = torch.ones(3)
    with torch.no_grad():
        for rank, (stacks, mine) in outs.items():
            assert set(stacks) == {"rgbmaps", "partrgbs", "depthmaps", "alphamaps", "normals", "depth_normals"}
            for slot, vi in enumerate(mine):
                r = render_part(cams[vi], pc, pipe, bg)
                assert torch.equal(stacks["rgbmaps"][slot], r["render"])
                assert torch.equal(stacks["depthmaps"][slot], r["surf_depth"])
                assert torch.equal(stacks["alphamaps"][slot], r["rend_alpha"])
                assert torch.equal(stacks["depth_normals"][slot], r["surf_normal"])
                assert float((stacks["normals"][slot] - extract_oracle.unit_normals(r["rend_normal"])).abs().max()) <= 1e-6
                want = extract_oracle.partmap_to_rgbmap(r["render_semantic"], pal)
                assert float((stacks["partrgbs"][slot] != want).float().mean()) <= 1e-3


def test_optimizer_knn_and_block_model_python_layers(emulated_host):
    from partgs_b200.optim import FusedAdam, densification_stats
    from partgs_b200.simple_knn._C import distCUDA2
    from partgs_b200.superquadric import BlockSurfelModel
    gen = torch.Generator().manual_seed(0)
    # FusedAdam == torch.optim.Adam over named groups with different learning rates
    init = [torch.randn(s, generator=gen) for s in ((200, 3), (200, 1, 3), (200, 2))]
    pa = [t.clone().requires_grad_(True) for t in init]
    pb = [t.clone().requires_grad_(True) for t in init]
    lrs = (1e-3, 2.5e-3, 5e-3)
    ref = torch.optim.Adam([{"params": [p], "lr": lr} for p, lr in zip(pa, lrs)], lr=0.0, eps=1e-15, foreach=False)
    ours = FusedAdam([{"params": [p], "lr": lr} for p, lr in zip(pb, lrs)], lr=0.0, eps=1e-15)
    for _ in range(3):
        for a, b in zip(pa, pb):
            g = torch.randn(a.shape, generator=gen)
            a.grad, b.grad = g.clone(), g.clone()
        ref.step(); ours.step()
    for a, b in zip(pa, pb):
        assert float((a - b).abs().max()) <= 2e-6 * float(a.abs().max())
        assert int(ours.state[b]["step"]) == 3
    # densification statistics
    radii = torch.randint(-1, 9, (200,), generator=gen).int()
    vg = torch.randn(200, 3, generator=gen)
    acc, den, mx = torch.zeros(200, 1), torch.zeros(200, 1), torch.zeros(200)
    densification_stats(radii, vg, acc, den, mx)
    vis = radii > 0
    assert torch.equal(den.squeeze(1), vis.float()) and torch.equal(mx[vis], radii[vis].float())
    assert torch.allclose(acc[vis, 0], vg[vis, :2].norm(dim=1), rtol=1e-6)
    # distCUDA2
    pts = torch.randn(300, 3, generator=gen)
    d = ((pts[:, None].double() - pts[None].double()) ** 2).sum(-1)
    d.fill_diagonal_(float("inf"))
    want = d.sort(dim=1).values[:, :3].mean(1)
    assert torch.allclose(distCUDA2(pts).double(), want, rtol=2e-5)
    # block model: the accessors the renderer reads, gradients down to the 13 block parameters
    m = BlockSurfelModel(2, 2, level=1, device="cpu", generator=gen)
    assert m.get_xyz.shape == (2 * 80 * 2, 3) and m.get_opacity.shape == (320, 1)
    (m.get_xyz.sum() + m.get_scaling.sum() + m.get_rotation[:, 0].sum()).backward()
    assert all(getattr(m, k).grad is not None and bool(torch.isfinite(getattr(m, k).grad).all())
               for k in ("sq_r", "sq_s", "sq_t", "sq_eps"))


def test_fused_block_rasteriser_equals_unfused_composition(emulated_host):
    """north_star (1): surfels generated inside preprocess (pgs_dsr_forward_blocks / _backward_blocks) against
    sq_to_surfels -> accessors -> GaussianRasterizer, values and gradients down to the block parameters."""
    import parity_utils as pu
    from partgs_b200.diff_surfel_rasterization import GaussianRasterizer
    from partgs_b200.superquadric import BlockSurfelModel, rasterize_blocks, sq_to_surfels
    gen = torch.Generator().manual_seed(11)
    model = BlockSurfelModel(2, 2, level=1, device="cpu", generator=gen)
    P = 2 * model.per_gs_num
    shs0 = torch.zeros(P, 16, 3)
    shs0[:, 0] = synth.RGB2SH(torch.rand(P, 3, generator=gen))
    shs0[:, 1:] = 0.05 * torch.randn(P, 15, 3, generator=gen)
    W, H = 40, 24
    cam = synth.make_cameras(1, W, H, 5, device="cpu")[0]
    g = synth.upstream_grads(W, H, 6, device="cpu")
    settings = pu.settings_from_cam(cam, torch.tensor([0.1, 0.2, 0.3]))
    names = ("sq_r", "sq_s", "sq_t", "sq_eps", "sq_occ")
    pa = {n: getattr(model, n).detach().clone().requires_grad_(True) for n in names}
    shs_a = shs0.clone().requires_grad_(True)
    _, xyz, scaling, rotation, opacity = sq_to_surfels(pa["sq_r"], pa["sq_s"], pa["sq_t"], pa["sq_eps"], pa["sq_occ"],
                                                       model.alpha, model._scale, model.sq_eta, model.sq_omega,
                                                       model.faces)
    m2d_a = torch.zeros_like(xyz, requires_grad=True)
    col_a, radii_a, all_a = GaussianRasterizer(settings)(means3D=xyz, means2D=m2d_a, opacities=opacity, shs=shs_a,
                                                         scales=torch.exp(scaling), rotations=rotation)
    torch.autograd.backward([col_a, all_a], [g["color"], g["allmap"]])
    pb = {n: getattr(model, n).detach().clone().requires_grad_(True) for n in names}
    shs_b = shs0.clone().requires_grad_(True)
    m2d_b = torch.zeros_like(xyz, requires_grad=True)
    out = rasterize_blocks(settings, pb["sq_r"], pb["sq_s"], pb["sq_t"], pb["sq_eps"], pb["sq_occ"], model.alpha,
                           model._scale, shs_b, model.sq_eta, model.sq_omega, model.faces, means2D=m2d_b,
                           materialize=True)
    col_b, radii_b, all_b, _verts, xyz_b, scaling_b, rot_b, opa_b = out
    torch.autograd.backward([col_b, all_b], [g["color"], g["allmap"]])
    assert int((radii_b > 0).sum()) > 20
    assert pu.rel_err(xyz_b, xyz) <= 1e-6 and pu.rel_err(scaling_b, scaling) <= 1e-6
    assert pu.rel_err(rot_b, rotation) <= 1e-6 and pu.rel_err(opa_b, opacity) <= 1e-6
    assert int((radii_a != radii_b).sum()) <= 1
    assert rel(col_b.detach().numpy(), col_a.detach().numpy()) <= 2e-5
    assert rel(all_b.detach().numpy(), all_a.detach().numpy()) <= 1e-4
    for n in names:
        assert pu.rel_err(pb[n].grad, pa[n].grad) <= 1e-3, n
    assert pu.rel_err(shs_b.grad, shs_a.grad) <= 1e-3 and pu.rel_err(m2d_b.grad, m2d_a.grad) <= 1e-3
    # a loss on the returned mesh vertices (alone: the images get no gradient, which arrives as None) flows to the
    # block parameters exactly like through sq_to_surfels
    gv = torch.randn(_verts.shape, generator=gen)
    pc = {n: getattr(model, n).detach().clone().requires_grad_(True) for n in names}
    verts_c = sq_to_surfels(pc["sq_r"], pc["sq_s"], pc["sq_t"], pc["sq_eps"], pc["sq_occ"], model.alpha, model._scale,
                            model.sq_eta, model.sq_omega, model.faces)[0]
    verts_c.backward(gv)
    pd = {n: getattr(model, n).detach().clone().requires_grad_(True) for n in names}
    out_d = rasterize_blocks(settings, pd["sq_r"], pd["sq_s"], pd["sq_t"], pd["sq_eps"], pd["sq_occ"], model.alpha,
                             model._scale, shs0, model.sq_eta, model.sq_omega, model.faces)
    assert torch.equal(out_d[3], verts_c.detach())
    out_d[3].backward(gv)
    for n in ("sq_r", "sq_s", "sq_t", "sq_eps"):
        assert float(pc[n].grad.abs().max()) > 0 and pu.rel_err(pd[n].grad, pc[n].grad) <= 1e-5, n
    assert float(pd["sq_occ"].grad.abs().max()) == 0.0
