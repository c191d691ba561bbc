"""Drop-in proof on the CPU: the REFERENCE'S OWN caller code — `render()` of renderer/gaussian_renderer/__init__.py
and `render_part()` of renderer/gaussian_renderer_2d/__init__.py, imported unmodified from /root/reference — runs on
this repo's rasteriser through the import-name shims of partgs_b200/dropin (`diff_surfel_rasterization`,
`diff_surfel_rasterization_part`), over the emulated library (tests/emu_host.py).

What it pins (north_star: "renderer/gaussian_renderer_2d and train.py use it as a drop-in"):
  * the shims resolve the reference's import lines and the classes accept the reference's call (keyword names, `None`
    for absent inputs, settings fields) and return tuples of the arity the reference unpacks;
  * the dictionaries the reference builds from our outputs equal what this repo's mirrors (`partgs_b200.renderer.render
    / render_part`, which fuse the post-processing into one kernel) return, forward and gradients — i.e. switching
    train.py from the reference's render() to the mirror changes nothing but speed.

Build-container only (needs /root/reference; the reference's three hard-coded `cuda` device strings are neutralised by
patching torch for the duration of the test, as tools/make_golden_post.py does)."""
import importlib.util
import sys
import types
from pathlib import Path
from types import SimpleNamespace

import pytest
import torch

from emu_host import emulated_host  # noqa: F401  (fixture)
from partgs_b200 import synth

REF = Path("/root/reference")
ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.skipif(not (REF / "renderer" / "gaussian_renderer" / "__init__.py").exists(),
                                reason="needs the reference tree (build container only)")


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, str(path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_NAMESPACES = ("diff_surfel_rasterization", "diff_surfel_rasterization_part", "scene", "utils", "renderer",
               "ref_gaussian_renderer", "ref_gaussian_renderer_2d")
_NAMESPACES = _NAMESPACES + ("games",)
_ABSENT_THIRD_PARTY = ("open3d", "seaborn", "matplotlib", "matplotlib.pyplot", "matplotlib.colors", "mediapy",
                       "pytorch3d", "pytorch3d.renderer", "pytorch3d.structures", "pytorch3d.structures.meshes",
                       "pytorch3d.structures.utils", "pytorch3d.ops", "pytorch3d.ops.subdivide_meshes", "pytorch3d.io",
                       "pytorch3d.io.utils", "pytorch3d.loss", "pytorch3d.utils", "iopath", "iopath.common",
                       "iopath.common.file_io", "PIL.ImageFile", "plyfile", "trimesh", "trimesh.voxel",
                       "trimesh.voxel.creation", "toolz", "imageio", "skimage", "lpips", "easydict")


@pytest.fixture()
def reference_callers(emulated_host, monkeypatch):
    """The reference's renderer modules and its GaussianExtractor, imported from where they lie, with
    `diff_surfel_rasterization[_part]` resolving to the drop-in shims.  `utils` is the reference's own package
    (sh_utils, point_utils, mesh_utils, ...); `scene.gaussian_model` (only a type annotation in the renderers; it
    drags in plyfile, simple_knn, ...) is a stub; third-party packages absent from this image (open3d, seaborn,
    matplotlib, mediapy, pytorch3d — none is executed on this path) are placeholders."""
    from unittest.mock import MagicMock
    monkeypatch.syspath_prepend(str(ROOT / "partgs_b200" / "dropin"))
    for name in [m for m in sys.modules if m.split(".")[0] in _NAMESPACES]:
        monkeypatch.delitem(sys.modules, name)
    before = set(sys.modules)
    for name in _ABSENT_THIRD_PARTY:
        try:
            __import__(name)
        except Exception:
            monkeypatch.setitem(sys.modules, name, MagicMock())
    scene = types.ModuleType("scene")
    gm = types.ModuleType("scene.gaussian_model")
    gm.GaussianModel = type("GaussianModel", (), {})
    scene.gaussian_model = gm
    utils = types.ModuleType("utils")
    utils.__path__ = [str(REF / "utils")]
    games = types.ModuleType("games")          # games/__init__.py pulls in the dataset readers: bypass it
    games.__path__ = [str(REF / "games")]
    monkeypatch.setitem(sys.modules, "scene", scene)
    monkeypatch.setitem(sys.modules, "scene.gaussian_model", gm)
    monkeypatch.setitem(sys.modules, "utils", utils)
    monkeypatch.setitem(sys.modules, "games", games)
    # the reference hard-codes device="cuda" (render():20, point_utils.py:10,14, mesh_utils.py:70,143)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.cuda, "empty_cache", lambda: None)

    def _on_cpu(fn):
        return lambda *a, **k: fn(*a, **{kk: v for kk, v in k.items() if not (kk == "device" and str(v) == "cuda")})

    for fname in ("arange", "zeros_like", "ones_like", "tensor", "zeros", "ones", "empty", "rand", "randn", "full"):
        monkeypatch.setattr(torch, fname, _on_cpu(getattr(torch, fname)))
    base = _load("ref_gaussian_renderer", REF / "renderer" / "gaussian_renderer" / "__init__.py")
    part = _load("ref_gaussian_renderer_2d", REF / "renderer" / "gaussian_renderer_2d" / "__init__.py")
    import utils.mesh_utils as mesh_utils  # the reference's, through the `utils` package above
    assert Path(mesh_utils.__file__).resolve() == (REF / "utils" / "mesh_utils.py").resolve()
    # the names the reference imported are this repo's classes
    from partgs_b200 import diff_surfel_rasterization as ours_base, diff_surfel_rasterization_part as ours_part
    assert base.GaussianRasterizer is ours_base.GaussianRasterizer
    assert part.GaussianRasterizer is ours_part.GaussianRasterizer
    yield SimpleNamespace(render=base.render, render_part=part.render_part, mesh_utils=mesh_utils,
                          scene_stub=(scene, gm))
    for name in [m for m in sys.modules if m not in before and m.split(".")[0] in _NAMESPACES]:
        sys.modules.pop(name, None)


PARAMS = ("means3D", "opacities", "scales", "rotations", "shs")


def _model(P, W, H, S, seed):
    scene = synth.make_point_scene(P, seed=seed, S=S, device="cpu")
    scene["scales"] = scene["scales"] * 3.0
    cam = synth.make_cameras(1, W, H, seed=seed + 10, device="cpu")[0]
    return scene, cam


def _pc(scene, grad):
    t = {k: scene[k].clone().requires_grad_(grad) for k in PARAMS}
    sem = scene["semantics"].clone().requires_grad_(grad) if scene.get("semantics") is not None else None
    pc = SimpleNamespace(get_xyz=t["means3D"], get_opacity=t["opacities"], get_scaling=t["scales"],
                         get_rotation=t["rotations"], get_features=t["shs"], get_semantic=sem, active_sh_degree=3,
                         max_sh_degree=3)
    return pc, t, sem


def _functional(r, gen_seed):
    """A fixed scalar of every differentiable map in the dictionary (same weights for both callers)."""
    gen = torch.Generator().manual_seed(gen_seed)
    tot = 0.0
    for k in ("render", "render_semantic", "rend_alpha", "rend_normal", "rend_dist", "surf_depth", "surf_normal"):
        if k in r:
            tot = tot + (r[k] * torch.randn(r[k].shape, generator=gen)).sum()
    return tot


def _close(a, b, tol):
    a, b = a.detach(), b.detach()
    return float((a - b).abs().max()) <= tol * (float(b.abs().max()) + 1e-30)


@pytest.mark.parametrize("depth_ratio", [0.0, 1.0])
def test_reference_render_runs_on_the_drop_in_and_equals_the_mirror(reference_callers, depth_ratio):
    from partgs_b200.renderer import render as mirror
    scene, cam = _model(P=200, W=40, H=24, S=0, seed=5)
    bg = torch.tensor([0.1, 0.2, 0.3])
    pipe = SimpleNamespace(depth_ratio=depth_ratio, compute_cov3D_python=False, convert_SHs_python=False, debug=False)
    pc_r, t_r, _ = _pc(scene, True)
    pc_m, t_m, _ = _pc(scene, True)
    r = reference_callers.render(cam, pc_r, pipe, bg)
    m = mirror(cam, pc_m, pipe, bg)
    assert set(r) == set(m)
    assert torch.equal(r["radii"], m["radii"]) and torch.equal(r["visibility_filter"], m["visibility_filter"])
    assert int((r["radii"] > 0).sum()) > 50
    for k in ("render", "rend_alpha", "rend_dist"):      # straight from the rasteriser: identical
        assert torch.equal(r[k], m[k]), k
    for k, tol in (("rend_normal", 1e-5), ("surf_depth", 1e-5), ("surf_normal", 2e-4)):
        assert _close(m[k], r[k], tol), k
    _functional(r, 7).backward()
    _functional(m, 7).backward()
    for k in PARAMS:
        assert t_r[k].grad is not None and float(t_r[k].grad.abs().max()) > 0, k
        assert _close(t_m[k].grad, t_r[k].grad, 2e-3), k
    # the densification signal (train.py:295-297 reads viewspace_points.grad)
    assert r["viewspace_points"].grad is not None and _close(m["viewspace_points"].grad, r["viewspace_points"].grad, 2e-3)


def test_reference_render_part_runs_on_the_drop_in_and_equals_the_mirror(reference_callers):
    from partgs_b200.renderer import render_part as mirror
    scene, cam = _model(P=200, W=40, H=24, S=5, seed=6)
    bg = torch.zeros(3)
    pipe = SimpleNamespace(depth_ratio=0.5, compute_cov3D_python=False, convert_SHs_python=False, debug=False)
    pc_r, t_r, s_r = _pc(scene, True)
    pc_m, t_m, s_m = _pc(scene, True)
    r = reference_callers.render_part(cam, pc_r, pipe, bg)
    m = mirror(cam, pc_m, pipe, bg)
    assert set(r) == set(m) and r["render_semantic"].shape == (5, 24, 40)
    assert torch.equal(r["radii"], m["radii"])
    for k in ("render", "render_semantic", "rend_alpha", "rend_dist"):
        assert torch.equal(r[k], m[k]), k
    for k, tol in (("rend_normal", 1e-5), ("surf_depth", 1e-5), ("surf_normal", 2e-4)):
        assert _close(m[k], r[k], tol), k
    _functional(r, 8).backward()
    _functional(m, 8).backward()
    for k in PARAMS:
        assert _close(t_m[k].grad, t_r[k].grad, 2e-3), k
    assert s_r.grad is not None and _close(s_m.grad, s_r.grad, 2e-3)


def test_reference_render_with_override_color_and_scaling_modifier(reference_callers):
    """The other branches of the reference's call: colours precomputed in Python (`override_color`) and a
    scaling_modifier != 1 handed through to the rasteriser settings."""
    from partgs_b200.renderer import render as mirror
    scene, cam = _model(P=120, W=32, H=16, S=0, seed=9)
    pipe = SimpleNamespace(depth_ratio=0.0, compute_cov3D_python=False, convert_SHs_python=False, debug=False)
    col = torch.rand(120, 3, generator=torch.Generator().manual_seed(1))
    pc, _, _ = _pc(scene, False)
    r = reference_callers.render(cam, pc, pipe, torch.zeros(3), scaling_modifier=0.7, override_color=col)
    m = mirror(cam, pc, pipe, torch.zeros(3), scaling_modifier=0.7, override_color=col)
    assert torch.equal(r["radii"], m["radii"]) and torch.equal(r["render"], m["render"])
    assert _close(m["surf_normal"], r["surf_normal"], 2e-4)


def test_reference_extraction_loop_runs_on_the_drop_in_and_equals_ours(reference_callers, monkeypatch):
    """render.py's loop: the reference's GaussianExtractor.reconstruction (utils/mesh_utils.py:102-129, unmodified)
    driving the reference's render_part on the drop-in rasteriser, against this repo's pipelined extractor (one fused
    epilogue kernel per view, pinned host stacks).  The palette is an input on both sides (seaborn is absent)."""
    from partgs_b200.extract import GaussianExtractor
    from partgs_b200.renderer import render_part as mirror
    from test_emu_host_layer import _FakeEvent, _FakeStream   # CUDA streams / events / pinning are no-ops on the emulator
    monkeypatch.setattr(torch.cuda, "Stream", _FakeStream)
    monkeypatch.setattr(torch.cuda, "Event", _FakeEvent)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: _FakeStream())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: s)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.Tensor, "record_stream", lambda self, s: None)
    scene, _ = _model(P=200, W=40, H=24, S=5, seed=11)
    cams = synth.make_cameras(3, 40, 24, seed=21, device="cpu")
    pipe = SimpleNamespace(depth_ratio=1.0, compute_cov3D_python=False, convert_SHs_python=False, debug=False)
    palette = torch.rand(6, 3, generator=torch.Generator().manual_seed(4))
    pc, _, _ = _pc(scene, False)
    mu = reference_callers.mesh_utils
    mu.get_fancy_color = lambda n: palette[:n]
    ref = mu.GaussianExtractor(pc, reference_callers.render_part, pipe, bg_color=[0, 0, 0])
    ref.reconstruction(cams)
    ours = GaussianExtractor(pc, mirror, pipe, bg_color=[0, 0, 0], palette=palette, device="cpu")
    ours.reconstruction(cams)
    assert ref.rgbmaps.shape == (3, 3, 24, 40) and float(ref.alphamaps.max()) > 0.5
    for name, tol in (("rgbmaps", 0.0), ("alphamaps", 0.0), ("partrgbs", 0.0), ("depthmaps", 1e-5),
                      ("depth_normals", 2e-4)):
        a, b = getattr(ours, name), getattr(ref, name)
        assert a.shape == b.shape, name
        assert float((a - b).abs().max()) <= tol * float(b.abs().max()), name
    # the reference keeps `normals` as a list of per-view maps (it never stacks them, mesh_utils.py:121-127)
    rn = torch.stack(ref.normals) if isinstance(ref.normals, list) else ref.normals
    on = torch.stack(list(ours.normals)) if isinstance(ours.normals, (list, tuple)) else ours.normals
    assert float((on - rn).abs().max()) <= 1e-5
    assert abs(float(ours.radius) - float(ref.radius)) <= 1e-6 * float(ref.radius)
    assert float((ours.center.cpu() - ref.center).abs().max()) <= 1e-5


def test_reference_model_class_trains_on_the_drop_ins(reference_callers, monkeypatch):
    """train.py:219-305 with the reference's own classes: `TwoGaussianModel` (games/block_mesh_splatting/scene/
    two_gaussian_model.py, unmodified) initialised by `create_from_pcd` — whose `distCUDA2` is the drop-in
    `simple_knn._C` —, the reference's `render()`, `l1_loss` / `ssim` (utils/loss_utils.py), `torch.optim.Adam`,
    `add_densification_stats` and `densify_and_prune`; beside it the all-product flow (fused mirrors, fused losses,
    FusedAdam, one-launch statistics, planned-compaction densification) from the same initial state.  Four iterations,
    a densification after the third: losses, parameters and the surviving surfels must agree."""
    import numpy as np
    import test_emu_training_loop as tl
    from partgs_b200.optim import FusedAdam
    # the real scene.gaussian_model (the renderers only needed the name), without running scene/__init__.py
    scene_pkg, _ = reference_callers.scene_stub
    scene_pkg.__path__ = [str(REF / "scene")]
    monkeypatch.delitem(sys.modules, "scene.gaussian_model")
    from games.block_mesh_splatting.scene.two_gaussian_model import TwoGaussianModel
    from utils.graphics_utils import BasicPointCloud
    from utils.loss_utils import l1_loss, ssim
    import simple_knn._C as knn_shim
    import partgs_b200.simple_knn._C as ours_knn
    assert knn_shim.distCUDA2 is ours_knn.distCUDA2

    W, H, P = 32, 16, 150
    src = synth.make_point_scene(P, seed=5, device="cpu")
    pts = src["means3D"].numpy().astype(np.float64)
    cols = torch.rand(P, 3, generator=torch.Generator().manual_seed(1)).numpy()
    torch.manual_seed(3)                       # create_from_pcd draws the rotations with torch.rand
    ref = TwoGaussianModel(3)
    ref.create_from_pcd(BasicPointCloud(points=pts, colors=cols, normals=np.zeros_like(pts)), 1.0)
    # distCUDA2 through the reference's initialiser == float64 brute force (mean squared distance to 3 neighbours)
    d = torch.cdist(torch.from_numpy(pts), torch.from_numpy(pts)) ** 2
    want = d.sort(dim=1).values[:, 1:4].mean(dim=1).clamp_min(1e-7).sqrt().log().float()
    assert float((ref._scaling.detach()[:, 0] - want).abs().max()) <= 1e-5
    with torch.no_grad():                       # visible splats, SH of every degree, mixed opacities (then shared)
        ref._scaling += 1.2
        ref._features_rest += 0.05 * torch.randn(ref._features_rest.shape, generator=torch.Generator().manual_seed(2))
        ref._opacity += 2.0 * torch.randn(P, 1, generator=torch.Generator().manual_seed(4)) + 2.0
    ref.active_sh_degree = 3
    ref._semantic = torch.nn.functional.one_hot(torch.arange(P) % 3, 3).float()
    targs = SimpleNamespace(percent_dense=0.01, position_lr_init=1.6e-3, position_lr_final=1.6e-5,
                            position_lr_delay_mult=0.01, position_lr_max_steps=30000, feature_lr=2.5e-2,
                            opacity_lr=0.05, scaling_lr=0.02, rotation_lr=0.01)
    ref.training_setup(targs)

    ours = tl.Model.__new__(tl.Model)           # the product-side model of test_emu_training_loop, same initial state
    for k in tl.NAMES:
        setattr(ours, tl.ATTR[k], torch.nn.Parameter(getattr(ref, tl.ATTR[k]).detach().clone().contiguous()))
    ours._semantic = ref._semantic.clone()
    ours.xyz_gradient_accum, ours.denom, ours.max_radii2D = torch.zeros(P, 1), torch.zeros(P, 1), torch.zeros(P)
    ours.percent_dense, ours.active_sh_degree = 0.01, 3
    ours.optimizer = FusedAdam([{"params": [getattr(ours, tl.ATTR[k])], "lr": g["lr"], "name": k}
                                for k, g in zip(tl.NAMES, ref.optimizer.param_groups)], lr=0.0, eps=1e-15)
    assert [g["name"] for g in ref.optimizer.param_groups] == list(tl.NAMES)

    cams = synth.make_cameras(2, W, H, seed=6, device="cpu")
    gen = torch.Generator().manual_seed(8)
    gts = [torch.rand(3, H, W, generator=gen) for _ in cams]
    masks = [(torch.rand(H, W, generator=gen) > 0.3).float() for _ in cams]
    bg = torch.zeros(3)
    pipe = SimpleNamespace(depth_ratio=1.0, compute_cov3D_python=False, convert_SHs_python=False, debug=False)
    lam = tl.LAM

    def reference_iteration(cam, gt, mask):
        pkg = reference_callers.render(cam, ref, pipe, bg)
        image, opacity = pkg["render"], pkg["rend_alpha"]
        loss = (1.0 - lam["dssim"]) * l1_loss(image, gt) + lam["dssim"] * (1.0 - ssim(image, gt))      # train.py:230-231
        opacity = opacity.clamp(1e-6, 1 - 1e-6).squeeze(0)                                               # :235-237
        loss = loss + lam["mask_entropy"] * -(mask * torch.log(opacity) + (1 - mask) * torch.log(1 - opacity)).mean()
        normal_error = (1 - (pkg["rend_normal"] * pkg["surf_normal"]).sum(dim=0))[None]                   # :246-248
        total = loss + lam["dist"] * pkg["rend_dist"].mean() + lam["normal"] * normal_error.mean()
        total.backward()
        with torch.no_grad():                                                                             # :291-305
            vis, radii = pkg["visibility_filter"], pkg["radii"]
            assert int(vis.sum()) > P // 3 and float(pkg["rend_alpha"].max()) > 0.5
            ref.max_radii2D[vis] = torch.max(ref.max_radii2D[vis], radii[vis])
            ref.add_densification_stats(pkg["viewspace_points"], vis)
            ref.optimizer.step()
            ref.optimizer.zero_grad(set_to_none=True)
        return float(total.detach())

    for it in range(4):
        v = it % 2
        l_ref = reference_iteration(cams[v], gts[v], masks[v])
        l_ours = tl._iteration(ours, cams[v], gts[v], masks[v], bg, ours=True)
        assert abs(l_ours - l_ref) <= 2e-4 * abs(l_ref), (it, l_ours, l_ref)
        if it == 2:
            g = (ref.xyz_gradient_accum / ref.denom).nan_to_num(0).squeeze(1).sort().values
            g = g[g > 0]
            lo, hi = int(0.35 * len(g)), int(0.65 * len(g))
            i = lo + int(torch.argmax(g[lo + 1:hi + 1] - g[lo:hi]))
            max_grad = float(0.5 * (g[i] + g[i + 1]))      # inside the widest gap: rounding cannot flip a selection
            extent = float(torch.exp(ref._scaling.detach()).max(dim=1).values.median()) / ref.percent_dense
            assert torch.equal(ours.denom, ref.denom) and torch.equal(ours.max_radii2D, ref.max_radii2D)
            torch.manual_seed(77)
            ref.densify_and_prune(max_grad, 0.005, extent, 20)
            n_ours = tl._densify(ours, max_grad, extent, True)   # seeds 77 itself
            assert n_ours == ref.get_xyz.shape[0] and n_ours > P, (n_ours, ref.get_xyz.shape[0])
            assert tl._densify.last["n_clones"] > 0 and tl._densify.last["n_children"] > 0
        for k in tl.NAMES:
            a, b = getattr(ours, tl.ATTR[k]).detach(), getattr(ref, tl.ATTR[k]).detach()
            assert a.shape == b.shape, (it, k)
            assert float((a - b).abs().max()) <= 2e-3 * (float(b.abs().max()) + 1e-6), (it, k)
    assert torch.equal(ours._semantic, ref._semantic)
