"""Run a few forward+backward steps of one config (target for ncu launch lists / captures).

usage: python tools/profile_step.py --cfg C3 --iters 3 [--ref] [--fwd-only]
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from partgs_b200 import synth  # noqa: E402
import parity_utils as pu  # noqa: E402


def other(a):
    dev = "cuda"
    if a.part:
        import test_gpu_part_raster as tp
        cfg = synth.CONFIGS["C4"]
        scene, cams, bg, g = tp._setup(a.P or cfg["P"], cfg["W"], cfg["H"], cfg["S"], views=1)
        for _ in range(a.iters):
            tp.run_ours(scene, cams[0], bg, g)
            torch.cuda.synchronize()
    elif a.blocks:
        from partgs_b200.superquadric import BlockSurfelModel, rasterize_blocks
        cfg = synth.CONFIGS["C5"]
        gen = torch.Generator().manual_seed(5)
        model = BlockSurfelModel(8, (a.P or cfg["P"]) // (8 * 320), device=dev, generator=gen)
        P = 8 * model.per_gs_num
        shs = torch.zeros(P, 16, 3); shs[:, 0] = synth.RGB2SH(torch.rand(P, 3, generator=gen))
        shs = shs.to(dev).requires_grad_(True)
        cam = synth.make_cameras(1, cfg["W"], cfg["H"], synth.SEED_BASE, device=dev)[0]
        g = synth.upstream_grads(cfg["W"], cfg["H"], synth.SEED_BASE, device=dev)
        bg = torch.zeros(3, device=dev)
        prm = [model.sq_r, model.sq_s, model.sq_t, model.sq_eps, model.sq_occ]
        for _ in range(a.iters):
            for p_ in prm + [shs]:
                p_.grad = None
            out = rasterize_blocks(pu.settings_from_cam(cam, bg), *prm, model.alpha, model._scale, shs, model.sq_eta,
                                   model.sq_omega, model.faces)
            torch.autograd.backward([out[0], out[2]], [g["color"], g["allmap"]])
            torch.cuda.synchronize()
    else:
        from partgs_b200 import losses, renderer
        from partgs_b200.optim import FusedAdam
        from partgs_b200.simple_knn._C import distCUDA2
        from partgs_b200.superquadric import BlockSurfelModel, sq_to_surfels
        cfg, scene, cams = synth.make_config("C3", device=dev, P=a.P, views=1)
        W, H = cfg["W"], cfg["H"]
        model = BlockSurfelModel(8, 8, device=dev, generator=torch.Generator().manual_seed(1))
        names = ("sq_r", "sq_s", "sq_t", "sq_eps", "sq_occ")
        gen = torch.Generator(device=dev).manual_seed(2)
        img = torch.rand(3, H, W, device=dev, generator=gen).requires_grad_(True)
        gt = torch.rand(3, H, W, device=dev, generator=gen)
        allmap = torch.rand(7, H, W, device=dev, generator=gen).requires_grad_(True)
        prm = [scene[k].clone().requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")]
        opt = FusedAdam([{"params": [p_], "lr": 1e-3} for p_ in prm], lr=0.0, eps=1e-15)
        for _ in range(a.iters):
            distCUDA2(scene["means3D"])
            o = sq_to_surfels(*[getattr(model, k) for k in names], model.alpha, model._scale, model.sq_eta,
                              model.sq_omega, model.faces)
            sum(x.sum() for x in o[1:]).backward()
            img.grad = None
            losses.photometric_loss(img, gt, 0.2).backward()
            allmap.grad = None
            maps = renderer.surface_maps(allmap, cams[0], 0.0)
            sum(m_.sum() for m_ in maps.values()).backward()
            for p_ in prm:
                p_.grad = torch.ones_like(p_)
            opt.step()
            torch.cuda.synchronize()
    print("done")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="C3")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--P", type=int, default=None)
    ap.add_argument("--ref", action="store_true", help="run the reference CUDA build (oracle/_ref) instead")
    ap.add_argument("--fwd-only", action="store_true")
    ap.add_argument("--part", action="store_true", help="the `_part` fork (C4 shape: 16 part channels)")
    ap.add_argument("--blocks", action="store_true", help="block-level scene, generation fused into preprocess (C5 shape)")
    ap.add_argument("--ops", action="store_true", help="distCUDA2, superquadric -> surfel, surface maps, L1+SSIM, "
                                                         "regularisers, Adam (one call each per iteration)")
    a = ap.parse_args()
    if a.part or a.blocks or a.ops:
        return other(a)
    dev = "cuda"
    cfg, scene, cams = synth.make_config(a.cfg, device=dev, P=a.P, views=1)
    cam = cams[0]
    bg = torch.zeros(3, device=dev)
    g = synth.upstream_grads(cfg["W"], cfg["H"], synth.SEED_BASE, device=dev)
    if a.ref:
        from oracle import ref_cuda
    for _ in range(a.iters):
        if a.ref:
            f = ref_cuda.forward(scene, cam, bg)
            if not a.fwd_only:
                ref_cuda.backward(f, scene, cam, bg, g["color"], g["allmap"])
        else:
            if a.fwd_only:
                pu.run_ours_raw(scene, cam, bg)
            else:
                pu.run_ours(scene, cam, bg, grads=g)
        torch.cuda.synchronize()
    print("done", cfg)


if __name__ == "__main__":
    main()
