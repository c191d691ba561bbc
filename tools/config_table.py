"""fwd+bwd device time of every BASELINE.json config (C1..C5), ours vs the unmodified reference CUDA build
(oracle/_ref), plus distCUDA2 and the superquadric->surfel kernels.  One JSON line per measurement.

usage: python tools/config_table.py [--cfgs C1,C2,C3,C4,C5] [--views 6] [--reps 3]
"""
import argparse, json, statistics, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from partgs_b200 import synth, _lib  # noqa: E402
import parity_utils as pu  # noqa: E402
from oracle import ref_cuda  # noqa: E402


def timed(fn, reps):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return ts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfgs", default="C1,C2,C3,C4,C5"); ap.add_argument("--views", type=int, default=6)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    dev = "cuda"
    for name in a.cfgs.split(","):
        if name not in synth.CONFIGS:
            continue
        cfg = dict(synth.CONFIGS[name]); nv = min(cfg["views"], a.views)
        cfg, scene, cams = synth.make_config(name, device=dev, views=nv)
        S, W, H = cfg["S"], cfg["W"], cfg["H"]
        bg = torch.zeros(3, device=dev)
        g = synth.upstream_grads(W, H, synth.SEED_BASE, n_aux=8 if S else 7, S=S, device=dev)
        part = S > 0
        if part:
            sys.path.insert(0, str(ROOT / "tests"))
            import test_gpu_part_raster as tp
            ours = lambda cam: tp.run_ours(scene, cam, bg, g)
            def ref(cam):
                f = ref_cuda.forward_part(scene, cam, bg)
                ref_cuda.backward_part(f, scene, cam, bg, g["color"], g["semantic"], g["allmap"])
            have_ref = ref_cuda.available("ref_dsrp_C")
        else:
            ours = lambda cam: pu.run_ours(scene, cam, bg, grads=g)
            def ref(cam):
                f = ref_cuda.forward(scene, cam, bg)
                ref_cuda.backward(f, scene, cam, bg, g["color"], g["allmap"])
            have_ref = ref_cuda.available("ref_dsr_C")
        res = {}
        for label, fn in (("ours", ours), ("reference", ref)):
            if label == "reference" and not have_ref:
                continue
            for cam in cams:       # warm-up: every view once
                fn(cam)
            torch.cuda.synchronize()
            if label == "ours":
                _lib.timing_enable(True); _lib.timing_read(reset=True)
            ts = []
            for cam in cams:
                ts += timed(lambda: fn(cam), a.reps)
            res[label] = statistics.median(ts)
            if label == "ours":
                st = _lib.timing_read(reset=True); _lib.timing_enable(False)
                n = len(cams) * a.reps
                res["stage_ms"] = {k: round(v[0] / n, 4) for k, v in st.items() if v[1]}
        o = pu.run_ours_raw(scene, cams[0], bg) if not part else None
        out = dict(cfg=name, P=cfg["P"], W=W, H=H, S=S, views=nv, fork="part" if part else "base",
                   ours_ms=round(res["ours"], 4), ours_fps=round(1e3 / res["ours"], 1),
                   reference_ms=round(res.get("reference", float("nan")), 4),
                   speedup=round(res.get("reference", float("nan")) / res["ours"], 2), stage_ms=res["stage_ms"])
        if o is not None:
            out["R_view0"] = int(o["num_rendered"]); out["V_view0"] = int((o["radii"] > 0).sum())
        print(json.dumps(out), flush=True)
        del scene, cams, g
        torch.cuda.empty_cache()

    # block-level scenes (C1 / C5 shapes): superquadric -> surfel generation fused into preprocess vs the unfused
    # composition sq_to_surfels -> accessors -> GaussianRasterizer
    from partgs_b200.superquadric import BlockSurfelModel, rasterize_blocks, sq_to_surfels
    from partgs_b200.diff_surfel_rasterization import GaussianRasterizer
    for label, B, K, W, H in (("C1-blocks", 8, 8, 400, 300), ("C5-blocks", 8, 1172, 1920, 1080)):
        if label.split("-")[0] not in a.cfgs.split(","):
            continue
        gen = torch.Generator().manual_seed(5)
        model = BlockSurfelModel(B, K, device=dev, generator=gen)
        P = B * model.per_gs_num
        shs = torch.zeros(P, 16, 3); shs[:, 0] = synth.RGB2SH(torch.rand(P, 3, generator=gen))
        shs = shs.to(dev).requires_grad_(True)
        cams = synth.make_cameras(min(a.views, 4), W, H, synth.SEED_BASE, device=dev)
        g = synth.upstream_grads(W, H, synth.SEED_BASE, device=dev)
        bg = torch.zeros(3, device=dev)
        prm = [model.sq_r, model.sq_s, model.sq_t, model.sq_eps, model.sq_occ]

        def clear():
            for p_ in prm + [shs]:
                p_.grad = None

        def unfused(cam):
            clear()
            _, xyz, scaling, rotation, opacity = sq_to_surfels(*prm, model.alpha, model._scale, model.sq_eta,
                                                               model.sq_omega, model.faces)
            m2d = torch.zeros_like(xyz, requires_grad=True)
            col, _, am = GaussianRasterizer(pu.settings_from_cam(cam, bg))(
                means3D=xyz, means2D=m2d, opacities=opacity, shs=shs, scales=torch.exp(scaling),
                rotations=torch.nn.functional.normalize(rotation))
            torch.autograd.backward([col, am], [g["color"], g["allmap"]])

        def fused(cam):
            clear()
            out = rasterize_blocks(pu.settings_from_cam(cam, bg), *prm, model.alpha, model._scale, shs, model.sq_eta,
                                   model.sq_omega, model.faces)
            torch.autograd.backward([out[0], out[2]], [g["color"], g["allmap"]])

        row = dict(cfg=label, P=P, W=W, H=H)
        for name_, fn in (("unfused_ms", unfused), ("fused_ms", fused)):
            for cam in cams:
                fn(cam)
            torch.cuda.synchronize()
            ts = []
            for cam in cams:
                ts += timed(lambda: fn(cam), a.reps)
            row[name_] = round(statistics.median(ts), 4)
        row["speedup_fused"] = round(row["unfused_ms"] / row["fused_ms"], 3)
        print(json.dumps(row), flush=True)
        del model, shs
        torch.cuda.empty_cache()

    # renderer post-processing: the reference's ATen sequence (oracle/post_oracle.py is the same
    # torch code minus the hard-coded .cuda()) on the GPU vs the fused kernels, forward + backward
    if "post" in a.cfgs.split(",") or a.cfgs == "C1,C2,C3,C4,C5":
        from oracle import post_oracle
        from partgs_b200.renderer import surface_maps, camera_constants
        for (W, H) in ((400, 300), (1600, 1200)):
            cfgp, scene, _ = synth.make_config("C1", device=dev, views=1)
            cam = synth.make_cameras(1, W, H, synth.SEED_BASE, device=dev)[0]
            allmap = pu.run_ours(scene, cam, torch.zeros(3, device=dev))["allmap"].detach()
            gen = torch.Generator().manual_seed(3)
            g = {k: torch.randn(sh, generator=gen).to(dev) for k, sh in (("rend_normal", (3, H, W)), ("surf_depth", (1, H, W)),
                                                                        ("surf_normal", (3, H, W)))}
            consts = camera_constants(cam)

            def run(fn):
                a_ = allmap.clone().requires_grad_(True)
                out = fn(a_)
                torch.autograd.backward([out[k] for k in g], [g[k] for k in g])

            f_ref = lambda a_: post_oracle.surface_maps(a_, cam, 1.0)
            f_ours = lambda a_: surface_maps(a_, cam, 1.0)
            f_ours_c = lambda a_: surface_maps(a_, cam, 1.0, constants=consts)
            row = dict(kernel="surface_maps fwd+bwd", W=W, H=H)
            for nm, fn in (("reference_ms", f_ref), ("ours_ms", f_ours), ("ours_cached_camera_ms", f_ours_c)):
                for _ in range(3):
                    run(fn)
                torch.cuda.synchronize()
                row[nm] = round(statistics.median(timed(lambda: run(fn), 7)), 4)
            row["speedup"] = round(row["reference_ms"] / row["ours_ms"], 1)
            print(json.dumps(row), flush=True)

    # photometric loss: the reference's conv2d-based L1 + SSIM on the GPU vs the fused kernels
    if "post" in a.cfgs.split(",") or a.cfgs == "C1,C2,C3,C4,C5":
        from oracle import loss_oracle
        from partgs_b200.losses import photometric_loss
        for (W, H) in ((400, 300), (1600, 1200)):
            gen = torch.Generator().manual_seed(1)
            gt = torch.rand(3, H, W, generator=gen).to(dev)
            base = (gt + 0.1 * torch.randn(3, H, W, generator=gen).to(dev)).clamp(0, 1)

            def run(fn):
                x = base.clone().requires_grad_(True)
                fn(x, gt, 0.2).backward()

            row = dict(kernel="photometric loss (L1+SSIM) fwd+bwd", W=W, H=H)
            for nm, fn in (("reference_ms", loss_oracle.photometric_loss), ("ours_ms", photometric_loss)):
                for _ in range(3):
                    run(fn)
                torch.cuda.synchronize()
                row[nm] = round(statistics.median(timed(lambda: run(fn), 7)), 4)
            row["speedup"] = round(row["reference_ms"] / row["ours_ms"], 1)
            print(json.dumps(row), flush=True)

    # one PartGS-style training iteration (train.py:225-257): render -> surface maps -> L1+SSIM + normal-consistency
    # + distortion losses -> backward.  "reference" = unmodified reference CUDA rasteriser + the reference's ATen code
    # for the rest (oracle/post_oracle.py, oracle/loss_oracle.py are that code), "ours" = this package end to end.
    if "post" in a.cfgs.split(",") or a.cfgs == "C1,C2,C3,C4,C5":
        from oracle import post_oracle, loss_oracle
        from partgs_b200.renderer import surface_maps
        from partgs_b200.losses import photometric_loss
        from partgs_b200.diff_surfel_rasterization import GaussianRasterizer

        class _RefRaster(torch.autograd.Function):   # reference CUDA behind autograd, like its own __init__.py
            @staticmethod
            def forward(ctx, means3D, shs, opac, scales, rots, cam, bg, scene):
                f = ref_cuda.forward(dict(scene, means3D=means3D, shs=shs, opacities=opac, scales=scales, rotations=rots), cam, bg)
                ctx.f, ctx.cam, ctx.bg = f, cam, bg
                ctx.scene = dict(scene, means3D=means3D, shs=shs, opacities=opac, scales=scales, rotations=rots)
                return f["color"], f["allmap"]
            @staticmethod
            def backward(ctx, gc, ga):
                gr = ref_cuda.backward(ctx.f, ctx.scene, ctx.cam, ctx.bg, gc.contiguous(), ga.contiguous())
                return gr["means3D"], gr["sh"], gr["opacity"], gr["scales"], gr["rotations"], None, None, None

        for name in ("C2", "C3"):
            cfgt, scene, cams = synth.make_config(name, device=dev, views=min(4, a.views))
            W, H = cfgt["W"], cfgt["H"]
            bg = torch.zeros(3, device=dev)
            gen = torch.Generator().manual_seed(9)
            gt = torch.rand(3, H, W, generator=gen).to(dev)
            leaf = {k: scene[k].clone().requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}

            def losses(color, sm, photo):
                normal_error = (1 - (sm["rend_normal"] * sm["surf_normal"]).sum(dim=0))[None]
                return photo(color, gt, 0.2) + 0.05 * normal_error.mean() + 1000.0 * sm["rend_dist"].mean()

            def ours(cam):
                for t_ in leaf.values():
                    t_.grad = None
                m2d = torch.zeros_like(leaf["means3D"], requires_grad=True)
                color, _, allmap = GaussianRasterizer(pu.settings_from_cam(cam, bg))(
                    means3D=leaf["means3D"], means2D=m2d, opacities=leaf["opacities"], shs=leaf["shs"],
                    scales=leaf["scales"], rotations=leaf["rotations"])
                losses(color, surface_maps(allmap, cam, 1.0), photometric_loss).backward()

            def ref(cam):
                for t_ in leaf.values():
                    t_.grad = None
                color, allmap = _RefRaster.apply(leaf["means3D"], leaf["shs"], leaf["opacities"], leaf["scales"],
                                                 leaf["rotations"], cam, bg, scene)
                losses(color, post_oracle.surface_maps(allmap, cam, 1.0), loss_oracle.photometric_loss).backward()

            row = dict(kernel="training iteration (render + surface maps + losses + backward)", cfg=name, W=W, H=H)
            # the same iteration with the host out of the loop: lazy instance count, and as a replayed CUDA graph
            from partgs_b200 import diff_surfel_rasterization as dsr
            from partgs_b200.graphs import GraphedIteration
            for cam in cams:
                ours(cam)
            dsr.set_lazy_count(True)
            ts = []
            for cam in cams:
                ours(cam)
                ts += timed(lambda: ours(cam), a.reps)
            row["ours_lazy_count_ms"] = round(statistics.median(ts), 4)
            dsr.set_lazy_count(False)
            dsr.resolve_count()
            git = GraphedIteration(lambda: ours(cams[0]), warmup=2)
            row["ours_cuda_graph_ms"] = round(statistics.median(timed(lambda: git.replay(check=False), 3 * a.reps)), 4)
            git.verify()
            del git
            for nm, fn in (("reference_ms", ref), ("ours_ms", ours)):
                if nm == "reference_ms" and not ref_cuda.available("ref_dsr_C"):
                    continue
                for cam in cams:
                    fn(cam)
                torch.cuda.synchronize()
                ts = []
                for cam in cams:
                    ts += timed(lambda: fn(cam), a.reps)
                row[nm] = round(statistics.median(ts), 4)
            if "reference_ms" in row:
                row["speedup"] = round(row["reference_ms"] / row["ours_ms"], 2)
            print(json.dumps(row), flush=True)
            del scene, leaf
            torch.cuda.empty_cache()

    # distCUDA2 and superquadric->surfel (C1 / C5 shapes)
    from partgs_b200.simple_knn._C import distCUDA2
    for n in (100_000, 1_000_000, 3_000_000):
        pts = torch.randn(n, 3, device=dev)
        distCUDA2(pts); torch.cuda.synchronize()
        t = statistics.median(timed(lambda: distCUDA2(pts), 5))
        row = dict(kernel="distCUDA2", P=n, ours_ms=round(t, 4))
        if ref_cuda.available("ref_knn_C"):
            K = ref_cuda.load("ref_knn_C")
            K.distCUDA2(pts); torch.cuda.synchronize()
            row["reference_ms"] = round(statistics.median(timed(lambda: K.distCUDA2(pts), 3)), 4)
            row["speedup"] = round(row["reference_ms"] / row["ours_ms"], 1)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
