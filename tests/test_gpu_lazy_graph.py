"""Host out of the loop: the lazy instance count (PGS_FWD_LAZY_COUNT / set_lazy_count) and CUDA-graph
capture of a forward + backward iteration give the results of the default (waiting) path, and a frame that needs more
instances than it was queued for is detected instead of silently producing garbage."""
import pytest
import torch

import parity_utils as pu

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _scene(P=150_000, W=800, H=600, views=4):
    from partgs_b200 import synth
    scene = synth.make_point_scene(P, seed=21, device=DEV)
    cams = synth.make_cameras(views, W, H, seed=22, device=DEV)
    g = synth.upstream_grads(W, H, 7, device=DEV)
    bg = torch.tensor([0.1, 0.2, 0.3], device=DEV)
    return scene, cams, g, bg


@pytest.fixture()
def lazy():
    from partgs_b200 import diff_surfel_rasterization as dsr
    yield dsr
    dsr.set_lazy_count(False)
    dsr.resolve_count()


def test_lazy_count_equals_the_waiting_path(lazy):
    dsr = lazy
    scene, cams, g, bg = _scene()
    eager = [pu.run_ours(scene, c, bg, grads=g) for c in cams]        # also teaches the library the capacity
    counts = [pu.run_ours_raw(scene, c, bg)["num_rendered"] for c in cams]
    dsr.set_lazy_count(True)
    raw = pu.run_ours_raw(scene, cams[0], bg)
    assert raw["num_rendered"] == dsr.COUNT_PENDING                   # the call did not wait ...
    n, overflow = dsr.resolve_count()
    assert (n, overflow) == (counts[0], False)                        # ... and the count arrives later
    for c, e in zip(cams, eager):
        o = pu.run_ours(scene, c, bg, grads=g)
        assert torch.equal(o["radii"], e["radii"])
        pu.assert_equal_images("color", o["color"], e["color"])
        pu.assert_equal_images("allmap", o["allmap"], e["allmap"])
        for k in e["grads"]:
            pu.assert_grad_close(k, o["grads"][k], e["grads"][k], rtol=1e-5, afloor=1e-6)   # atomics order only


def test_lazy_overflow_is_detected_and_recovers(lazy):
    dsr = lazy
    from partgs_b200 import _lib
    scene, cams, g, bg = _scene(P=600_000)
    scene["scales"] = scene["scales"] * 2.0            # a frame well beyond the minimum arena of 512 Ki instances
    assert pu.run_ours_raw(scene, cams[0], bg)["num_rendered"] > 700_000
    small = {k: v[:2000].contiguous() for k, v in scene.items()}
    _lib.load().pgs_dsr_set_capacity_hint(0)
    pu.run_ours(small, cams[0], bg, grads=g)          # the library now remembers a small frame (capacity 512 Ki)
    want = pu.run_ours(scene, cams[0], bg, grads=g)   # eager: counts, re-launches with a larger arena — fine
    assert int((want["radii"] > 0).sum()) > 100_000
    _lib.load().pgs_dsr_set_capacity_hint(1000)
    dsr.set_lazy_count(True)
    with pytest.raises(RuntimeError, match="lazy instance count"):
        pu.run_ours(scene, cams[0], bg, grads=g)      # queued for 512 Ki instances, needs ~1.6 M
    got = pu.run_ours(scene, cams[0], bg, grads=g)    # the capacity was raised by the failed attempt
    pu.assert_equal_images("color", got["color"], want["color"])
    for k in want["grads"]:
        pu.assert_grad_close(k, got["grads"][k], want["grads"][k], rtol=1e-5, afloor=1e-6)


def test_cuda_graph_of_an_iteration_replays_other_views(lazy):
    from partgs_b200.diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    from partgs_b200.graphs import GraphedIteration
    from partgs_b200 import _lib
    scene, cams, g, bg = _scene()
    eager = [pu.run_ours(scene, c, bg, grads=g) for c in cams]
    # static inputs of the captured iteration: the camera lives in three small device tensors
    view, proj, campos = cams[0].viewmatrix.clone(), cams[0].projmatrix.clone(), cams[0].campos.clone()
    leaf = {k: scene[k].detach().clone().requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "shs")}
    means2D = torch.zeros_like(leaf["means3D"], requires_grad=True)
    st = GaussianRasterizationSettings(image_height=cams[0].image_height, image_width=cams[0].image_width,
                                       tanfovx=cams[0].tanfovx, tanfovy=cams[0].tanfovy, bg=bg, scale_modifier=1.0,
                                       viewmatrix=view, projmatrix=proj, sh_degree=3, campos=campos, prefiltered=False,
                                       debug=False)

    def step():
        for t in list(leaf.values()) + [means2D]:
            t.grad = None
        color, radii, allmap = GaussianRasterizer(st)(means3D=leaf["means3D"], means2D=means2D, opacities=leaf["opacities"],
                                                      shs=leaf["shs"], scales=leaf["scales"], rotations=leaf["rotations"])
        torch.autograd.backward([color, allmap], [g["color"], g["allmap"]])
        return color.detach(), allmap.detach(), radii

    l0 = _lib.load().pgs_launch_count()
    it = GraphedIteration(step, warmup=2)
    l1 = _lib.load().pgs_launch_count()
    for c, e in zip(cams, eager):
        view.copy_(c.viewmatrix); proj.copy_(c.projmatrix); campos.copy_(c.campos)
        color, allmap, radii = it.replay()
        assert torch.equal(radii, e["radii"])
        pu.assert_equal_images("color", color, e["color"])
        pu.assert_equal_images("allmap", allmap, e["allmap"])
        for k, p in (("means3D", leaf["means3D"]), ("opacity", leaf["opacities"]), ("scales", leaf["scales"]),
                     ("rotations", leaf["rotations"]), ("sh", leaf["shs"]), ("means2D", means2D)):
            pu.assert_grad_close(k, p.grad, e["grads"][k], rtol=1e-5, afloor=1e-6)
    assert _lib.load().pgs_launch_count() == l1 > l0      # replays issue no launches from the host


def test_block_level_iteration_lazy_and_graphed(lazy):
    """The block-level op (surfels generated inside preprocess, C5's path) with a lazy count and captured as a CUDA
    graph gives the images and the block / SH gradients of the default waiting path, other views included."""
    from partgs_b200 import synth
    from partgs_b200.graphs import GraphedIteration
    from partgs_b200.superquadric import BlockSurfelModel, rasterize_blocks
    dsr = lazy
    gen = torch.Generator().manual_seed(5)
    model = BlockSurfelModel(8, 4, device=DEV, generator=gen)
    P = 8 * model.per_gs_num
    shs = torch.zeros(P, 16, 3, device=DEV)
    shs[:, 0] = synth.RGB2SH(torch.rand(P, 3, generator=gen)).to(DEV)
    W, H = 640, 480
    cams = synth.make_cameras(3, W, H, seed=23, device=DEV)
    g = synth.upstream_grads(W, H, 8, device=DEV)
    bg = torch.tensor([0.1, 0.2, 0.3], device=DEV)
    names = ("sq_r", "sq_s", "sq_t", "sq_eps", "sq_occ")
    leaf = {n: getattr(model, n).detach().clone().requires_grad_(True) for n in names}
    leaf["shs"] = shs.clone().requires_grad_(True)
    view, proj, campos = cams[0].viewmatrix.clone(), cams[0].projmatrix.clone(), cams[0].campos.clone()
    st = pu.settings_from_cam(cams[0], bg)._replace(viewmatrix=view, projmatrix=proj, campos=campos)

    def step():
        for t in leaf.values():
            t.grad = None
        color, radii, allmap, _ = rasterize_blocks(st, leaf["sq_r"], leaf["sq_s"], leaf["sq_t"], leaf["sq_eps"],
                                                   leaf["sq_occ"], model.alpha, model._scale, leaf["shs"], model.sq_eta,
                                                   model.sq_omega, model.faces)
        torch.autograd.backward([color, allmap], [g["color"], g["allmap"]])
        return color.detach(), allmap.detach(), radii

    def at(c):
        view.copy_(c.viewmatrix); proj.copy_(c.projmatrix); campos.copy_(c.campos)

    eager = []
    for c in cams:
        at(c)
        color, allmap, radii = step()
        eager.append((color.clone(), allmap.clone(), radii.clone(), {k: v.grad.clone() for k, v in leaf.items()}))
    assert int((eager[0][2] > 0).sum()) > 1000

    def check(out, e):
        color, allmap, radii = out
        assert torch.equal(radii, e[2])
        pu.assert_equal_images("color", color, e[0])
        pu.assert_equal_images("allmap", allmap, e[1])
        for k, v in leaf.items():
            pu.assert_grad_close(k, v.grad, e[3][k], rtol=1e-4, afloor=1e-5)     # float atomics order only

    dsr.set_lazy_count(True)
    for c, e in zip(cams, eager):
        at(c)
        check(step(), e)
    dsr.set_lazy_count(False)
    it = GraphedIteration(step, warmup=2)
    for c, e in zip(cams, eager):
        at(c)
        check(it.replay(), e)
