"""A whole PartGS-style refinement loop on the CPU: render -> surface maps -> L1+SSIM + mask-entropy / normal /
distortion terms -> backward -> densification statistics -> Adam -> (once) densify_and_prune, mirroring train.py:219-305.

"ours" runs the product end to end (Python layer + every kernel through the C ABI, on the lock-step emulator);
"reference" is the composition of the pinned restatements: the C oracle rasteriser (wrapped in autograd), the
reference's post-processing and loss code (oracle/post_oracle.py, loss_oracle.py), torch.optim.Adam, the reference's
statistics expressions and its densify_and_prune (oracle/densify_oracle.py).  Losses and parameters must track each
other over the iterations, the densification must select / clone / split / prune the same surfels."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from emu_host import emulated_host  # noqa: F401  (fixture)
from oracle import cpu_oracle, densify_oracle, loss_oracle, post_oracle
from partgs_b200 import synth

NAMES = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")
LRS = {"xyz": 1.6e-3, "f_dc": 2.5e-2, "f_rest": 1.25e-3, "opacity": 0.05, "scaling": 0.02, "rotation": 0.01}
ATTR = {"xyz": "_xyz", "f_dc": "_features_dc", "f_rest": "_features_rest", "opacity": "_opacity",
        "scaling": "_scaling", "rotation": "_rotation"}
LAM = dict(dssim=0.2, mask_entropy=0.1, normal=0.05, dist=100.0)


class _OracleRaster(torch.autograd.Function):
    """The C restatement of the reference rasteriser as a differentiable op (test helper)."""
    @staticmethod
    def forward(ctx, means3D, means2D, shs, opacities, scales, rotations, cam, bg):
        f = cpu_oracle.forward(means3D.detach(), scales.detach(), rotations.detach(), opacities.detach(), shs.detach(),
                               cam.viewmatrix, cam.projmatrix, cam.campos, cam.image_width, cam.image_height,
                               cam.tanfovx, cam.tanfovy, bg=bg, keep_state=True)
        ctx.f = f
        return torch.from_numpy(f["color"]), torch.from_numpy(f["radii"]), torch.from_numpy(f["allmap"])

    @staticmethod
    def backward(ctx, g_color, _g_radii, g_allmap):
        g = cpu_oracle.backward(ctx.f, g_color.contiguous(), g_allmap.contiguous())
        t = lambda k: torch.from_numpy(g[k])
        return t("means3D"), t("means2D"), t("sh"), t("opacity"), t("scales"), t("rotations"), None, None


class Model:
    """The attributes / accessors of the reference's point-level model that the loop touches."""
    def __init__(self, scene, optimizer_cls):
        raw = {"xyz": scene["means3D"], "f_dc": scene["shs"][:, :1], "f_rest": scene["shs"][:, 1:],
               "opacity": torch.logit(scene["opacities"].clamp(1e-4, 1 - 1e-4)), "scaling": torch.log(scene["scales"]),
               "rotation": scene["rotations"]}
        for k, v in raw.items():
            setattr(self, ATTR[k], torch.nn.Parameter(v.clone().contiguous()))
        P = raw["xyz"].shape[0]
        self._semantic = torch.nn.functional.one_hot(torch.arange(P) % 3, 3).float()
        self.xyz_gradient_accum, self.denom, self.max_radii2D = torch.zeros(P, 1), torch.zeros(P, 1), torch.zeros(P)
        self.percent_dense, self.active_sh_degree = 0.01, 3
        self.optimizer = optimizer_cls([{"params": [getattr(self, ATTR[k])], "lr": LRS[k], "name": k} for k in NAMES],
                                       lr=0.0, eps=1e-15)

    get_xyz = property(lambda s: s._xyz)
    get_features = property(lambda s: torch.cat((s._features_dc, s._features_rest), dim=1))
    get_opacity = property(lambda s: torch.sigmoid(s._opacity))
    get_scaling = property(lambda s: torch.exp(s._scaling))
    get_rotation = property(lambda s: torch.nn.functional.normalize(s._rotation))


def _reference_render(cam, pc, bg):
    sp = torch.zeros_like(pc.get_xyz, requires_grad=True) + 0
    sp.retain_grad()
    color, radii, allmap = _OracleRaster.apply(pc.get_xyz, sp, pc.get_features, pc.get_opacity, pc.get_scaling,
                                               pc.get_rotation, cam, bg)
    r = {"render": color, "viewspace_points": sp, "visibility_filter": radii > 0, "radii": radii}
    r.update(post_oracle.surface_maps(allmap, cam, 1.0))
    return r


def _iteration(m, cam, gt, mask, bg, ours):
    if ours:
        from partgs_b200.losses import geometric_regularizers, photometric_loss
        from partgs_b200.optim import densification_stats
        from partgs_b200.renderer import render
        pipe = SimpleNamespace(depth_ratio=1.0, compute_cov3D_python=False, convert_SHs_python=False)
        pkg = render(cam, m, pipe, bg)
        loss = photometric_loss(pkg["render"], gt, LAM["dssim"]) + \
            geometric_regularizers(pkg, mask, LAM["mask_entropy"], LAM["normal"], LAM["dist"])
    else:
        pkg = _reference_render(cam, m, bg)
        loss = loss_oracle.photometric_loss(pkg["render"], gt, LAM["dssim"]) + loss_oracle.geometric_regularizers(
            pkg["rend_alpha"], mask, pkg["rend_dist"], pkg["rend_normal"], pkg["surf_normal"], LAM["mask_entropy"],
            LAM["normal"], LAM["dist"])[0]
    loss.backward()
    with torch.no_grad():
        vis, radii, vsp = pkg["visibility_filter"], pkg["radii"], pkg["viewspace_points"]
        if ours:
            densification_stats(radii.int(), vsp.grad, m.xyz_gradient_accum, m.denom, m.max_radii2D)
        else:  # train.py:295-297, scene/gaussian_model.py:515-517
            m.max_radii2D[vis] = torch.max(m.max_radii2D[vis], radii[vis].float())
            m.xyz_gradient_accum[vis] += torch.norm(vsp.grad[vis, :2], dim=-1, keepdim=True)
            m.denom[vis] += 1
        m.optimizer.step()
        m.optimizer.zero_grad(set_to_none=True)
    return float(loss.detach())


def _densify(m, max_grad, extent, ours):
    torch.manual_seed(77)
    if ours:
        from partgs_b200.densify import densify_and_prune_model
        info = densify_and_prune_model(m, max_grad, 0.005, extent, 20)
        _densify.last = info
        return info["n_out"]
    from partgs_b200.densify import rewrap_optimizer
    params = {k: getattr(m, ATTR[k]) for k in NAMES}
    moments = {k: (m.optimizer.state[params[k]]["exp_avg"], m.optimizer.state[params[k]]["exp_avg_sq"]) for k in NAMES}
    _, split = densify_oracle.split_selection(m.xyz_gradient_accum.clone(), m.denom, params["scaling"].detach(),
                                              max_grad, extent, m.percent_dense)
    z = torch.empty(2 * int(split.sum()), 3).normal_()
    cur, mom, sem, info = densify_oracle.densify_and_prune(params, moments, m._semantic, m.xyz_gradient_accum.clone(),
                                                           m.denom, max_grad, 0.005, extent, 20, m.percent_dense, z)
    wrapped = rewrap_optimizer(m.optimizer, cur, mom)
    for k in NAMES:
        setattr(m, ATTR[k], wrapped[k])
    n = info["n_out"]
    m._semantic = sem
    m.xyz_gradient_accum, m.denom, m.max_radii2D = torch.zeros(n, 1), torch.zeros(n, 1), torch.zeros(n)
    return n


def test_refinement_loop_tracks_the_reference_composition(emulated_host):
    from partgs_b200.optim import FusedAdam
    W, H = 32, 16
    scene = synth.make_point_scene(120, seed=5, device="cpu")
    scene["scales"] = scene["scales"] * 3.0
    cams = synth.make_cameras(2, W, H, seed=6, device="cpu")
    gen = torch.Generator().manual_seed(8)
    gts = [torch.rand(3, H, W, generator=gen) for _ in cams]
    masks = [(torch.rand(H, W, generator=gen) > 0.3).float() for _ in cams]
    bg = torch.zeros(3)
    ref, ours = Model(scene, lambda g, **k: torch.optim.Adam(g, foreach=False, **k)), Model(scene, FusedAdam)
    for it in range(5):
        v = it % 2
        l_ref = _iteration(ref, cams[v], gts[v], masks[v], bg, ours=False)
        l_ours = _iteration(ours, cams[v], gts[v], masks[v], bg, ours=True)
        assert abs(l_ours - l_ref) <= 2e-4 * abs(l_ref), (it, l_ours, l_ref)
        if it == 2:
            # a threshold inside the widest gap of the accumulated gradients, so that rounding cannot flip a selection
            g = (ref.xyz_gradient_accum / ref.denom).nan_to_num(0).squeeze(1).sort().values
            g = g[g > 0]
            lo, hi = int(0.35 * len(g)), int(0.65 * len(g))
            i = lo + int(torch.argmax(g[lo + 1:hi + 1] - g[lo:hi]))
            max_grad = float(0.5 * (g[i] + g[i + 1]))
            # percent_dense * extent = the median surfel size: about half of the selected surfels clone, half split
            extent = float(torch.exp(ref._scaling.detach()).max(dim=1).values.median()) / ref.percent_dense
            assert torch.equal(ours.denom, ref.denom) and torch.equal(ours.max_radii2D, ref.max_radii2D)
            assert float((ours.xyz_gradient_accum - ref.xyz_gradient_accum).abs().max()) <= 1e-3 * float(g.max())
            n_ref, n_ours = _densify(ref, max_grad, extent, False), _densify(ours, max_grad, extent, True)
            assert n_ours == n_ref and n_ref > 120, (n_ours, n_ref)
            assert _densify.last["n_clones"] > 0 and _densify.last["n_children"] > 0, _densify.last   # both branches ran
        for k in NAMES:
            a, b = getattr(ours, ATTR[k]).detach(), getattr(ref, ATTR[k]).detach()
            assert a.shape == b.shape, (it, k)
            assert float((a - b).abs().max()) <= 2e-3 * (float(b.abs().max()) + 1e-6), (it, k)
    assert torch.equal(ours._semantic, ref._semantic)
