"""Device time of the SURVEY 8(f) rank-3/4 rows, ours vs the reference's torch sequence (the oracle restatements run on
the GPU: the same ATen calls the reference makes).  One JSON line per row; every row is independent (a failure prints
{"row": ..., "error": ...} and the tool goes on).  Intended as the first GPU call of round 2, after
`pytest tests/test_gpu_zz_*.py`.

usage: python tools/time_rank34.py [--P 1000000] [--reps 5]
"""
import argparse, json, statistics, sys, time, traceback
from pathlib import Path
from types import SimpleNamespace
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))


def timed(fn, reps):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append((e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3))
    return round(statistics.median(t[0] for t in ts), 4), round(statistics.median(t[1] for t in ts), 4)


def row_densify(P, reps):
    from oracle import densify_oracle
    from partgs_b200 import densify
    import test_gpu_zz_densify as td
    extent = 3.7
    params, moments, sem, accum, denom = td._random_model(P, 16, 3, 77, extent)
    args = (0.0002, 0.005, extent, 20, 0.01)
    _, split = densify_oracle.split_selection(accum.clone(), denom, params["scaling"], args[0], extent, 0.01)
    z = torch.randn(2 * int(split.sum()), 3, device="cuda")
    ours = lambda: densify.densify_and_prune(params, moments, sem, accum, denom, *args, z=z)
    ref = lambda: densify_oracle.densify_and_prune(params, moments, sem, accum.clone(), denom, *args, z)
    info = ours()[3]; ref()
    o, ow = timed(ours, reps); r, rw = timed(ref, reps)
    bytes_alg = 20 * P + (4 + 2 * 4 * (58 + 116 + 16)) * info["n_out"]
    return dict(row="densify_and_prune", P=P, n_out=info["n_out"], ours_ms=o, ours_wall_ms=ow, reference_ms=r,
                reference_wall_ms=rw, speedup=round(rw / ow, 2), algorithmic_GBps=round(bytes_alg / o / 1e6, 1))


def row_adam(P, reps):
    from partgs_b200.optim import FusedAdam
    shapes = [(P, 3), (P, 1, 3), (P, 15, 3), (P, 1), (P, 2), (P, 4)]
    out = {}
    for nm, cls, kw in (("reference", torch.optim.Adam, dict(foreach=False)), ("reference_foreach", torch.optim.Adam, dict(foreach=True)),
                        ("ours", FusedAdam, {})):
        ps = [torch.randn(s, device="cuda", requires_grad=True) for s in shapes]
        for p in ps:
            p.grad = torch.randn_like(p)
        opt = cls([{"params": [p], "lr": 1e-3} for p in ps], lr=0.0, eps=1e-15, **kw)
        opt.step()
        out[nm + "_ms"], out[nm + "_wall_ms"] = timed(opt.step, reps)
    n = sum(torch.Size(s).numel() for s in shapes)
    return dict(row="adam_step", P=P, **out, speedup=round(out["reference_wall_ms"] / out["ours_wall_ms"], 2),
                algorithmic_GBps=round(28 * n / out["ours_ms"] / 1e6, 1))


def row_extract(reps, H=1200, W=1600):
    from oracle import extract_oracle
    from partgs_b200.extract import extract_maps
    S = 16
    part = torch.rand(S, H, W, device="cuda"); nrm = torch.randn(3, H, W, device="cuda"); pal = torch.rand(S + 1, 3, device="cuda")
    ours = lambda: extract_maps(part, nrm, pal)
    ref = lambda: (extract_oracle.partmap_to_rgbmap(part, pal), extract_oracle.unit_normals(nrm))
    ours(); ref()
    o, ow = timed(ours, reps); r, rw = timed(ref, reps)
    return dict(row=f"extract_epilogue_{W}x{H}_S16", ours_ms=o, reference_ms=r, speedup=round(r / o, 2),
                algorithmic_GBps=round((4 * S + 12 + 24) * H * W / o / 1e6, 1))


def row_reconstruction(reps):
    """whole extraction loop, 8 views at 800x600, 500k surfels with 16 parts (C4 shape): ours vs the reference's loop
    structure (per-view torch epilogue + six blocking .cpu() copies) on top of the same render_part."""
    from oracle import extract_oracle
    from partgs_b200 import synth
    from partgs_b200.extract import GaussianExtractor, fancy_palette
    from partgs_b200.renderer import render_part
    scene = synth.make_point_scene(500_000, seed=4, S=16, device="cuda")
    cams = synth.make_cameras(8, 800, 600, seed=9, device="cuda")
    pc = SimpleNamespace(get_xyz=scene["means3D"], get_opacity=scene["opacities"], get_scaling=scene["scales"],
                         get_rotation=scene["rotations"], get_features=scene["shs"], get_semantic=scene["semantics"],
                         active_sh_degree=3)
    pipe = SimpleNamespace(depth_ratio=1.0, compute_cov3D_python=False, convert_SHs_python=False)
    ex = GaussianExtractor(pc, render_part, pipe)
    pal = fancy_palette(17).cuda(); bg = torch.zeros(3, device="cuda")

    @torch.no_grad()
    def ref():
        keep = []
        for cam in cams:
            r = render_part(cam, pc, pipe, bg)
            keep.append([extract_oracle.partmap_to_rgbmap(r["render_semantic"], pal).cpu(), r["render"].cpu(),
                         r["surf_depth"].cpu(), r["rend_alpha"].cpu(),
                         torch.nn.functional.normalize(r["rend_normal"], dim=0).cpu(), r["surf_normal"].cpu()])
        return [torch.stack([k[i] for k in keep]) for i in range(6)]
    ours = lambda: ex.reconstruction(cams)
    ours(); ref()
    _, ow = timed(ours, reps); _, rw = timed(ref, reps)
    return dict(row="extraction_loop_8views_800x600_P500k_S16", ours_wall_ms=ow, reference_loop_wall_ms=rw,
                speedup=round(rw / ow, 2))


def row_regularizers(reps, H=1200, W=1600):
    from oracle import loss_oracle
    from partgs_b200.losses import geometric_regularizers
    allmap = torch.rand(7, H, W, device="cuda"); mask = (torch.rand(H, W, device="cuda") > 0.4).float()
    rn0 = torch.randn(3, H, W, device="cuda"); sn0 = torch.randn(3, H, W, device="cuda")

    def run(fused):
        am = allmap.clone().requires_grad_(True); rn = rn0.clone().requires_grad_(True); sn = sn0.clone().requires_grad_(True)
        pkg = {"rend_alpha": am[1:2], "rend_dist": am[6:7], "rend_normal": rn, "surf_normal": sn}
        loss = geometric_regularizers(pkg, mask, 0.1, 0.05, 1000.0) if fused else \
            loss_oracle.geometric_regularizers(pkg["rend_alpha"], mask, pkg["rend_dist"], rn, sn, 0.1, 0.05, 1000.0)[0]
        loss.backward()
    run(True); run(False)
    o, ow = timed(lambda: run(True), reps); r, rw = timed(lambda: run(False), reps)
    return dict(row=f"regularizers_fwd_bwd_{W}x{H} (incl. 3 clones)", ours_ms=o, reference_ms=r, speedup=round(r / o, 2))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--P", type=int, default=1_000_000); ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--rows", default="densify,adam,extract,regularizers,reconstruction",
                    help="comma-separated subset of: densify, adam, extract, regularizers, reconstruction")
    a = ap.parse_args()
    table = dict(densify=(row_densify, (a.P, a.reps)), adam=(row_adam, (a.P, a.reps)), extract=(row_extract, (a.reps,)),
                 regularizers=(row_regularizers, (a.reps,)), reconstruction=(row_reconstruction, (3,)))
    for name in [r for r in a.rows.split(",") if r]:
        fn, args = table[name]
        try:
            print(json.dumps(fn(*args)), flush=True)
        except Exception as ex:  # keep going: every row is independent
            print(json.dumps(dict(row=fn.__name__, error=repr(ex), trace=traceback.format_exc()[-600:])), flush=True)


if __name__ == "__main__":
    main()
