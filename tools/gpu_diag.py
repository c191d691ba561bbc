"""GPU diagnostics: stage-by-stage comparison of the product CUDA path with the
reference CUDA build in oracle/_ref (test infrastructure; run under gpurun).

usage: python tools/gpu_diag.py [--cfg C1 C2 C3] [--views N] [--out gpurun_out/diag.json]
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from partgs_b200 import debug, synth  # noqa: E402
from oracle import ref_cuda  # noqa: E402
import parity_utils as pu  # noqa: E402


def timed(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def diag_config(name, views, P=None, do_time=True):
    dev = "cuda"
    cfg, scene, cams = synth.make_config(name, device=dev, P=P, views=views)
    W, H, Pn = cfg["W"], cfg["H"], cfg["P"]
    bg = torch.zeros(3, device=dev)
    g = synth.upstream_grads(W, H, synth.SEED_BASE, device=dev)
    res = dict(cfg=name, P=Pn, W=W, H=H, views=[])
    for vi, cam in enumerate(cams):
        r = {}
        ref = ref_cuda.forward(scene, cam, bg)
        ours = pu.run_ours_raw(scene, cam, bg)
        R = ref["num_rendered"]
        r["R_ref"], r["R_ours"] = R, ours["num_rendered"]
        r["visible"] = int((ref["radii"] > 0).sum())
        r["radii_mismatch"] = int((ref["radii"] != ours["radii"]).sum())
        rg = ref_cuda.parse_geom(ref["geom"], Pn)
        st = debug.parse_state(ours["geom"], ours["img"], ours["binning"], Pn, W, H, ours["num_rendered"])
        vis = ref["radii"] > 0
        r["tiles_touched_mismatch"] = int((rg["tiles_touched"] != st["tiles_touched"]).sum())
        for k in ("transMat", "means2D", "normal_opacity", "rgb", "depths"):
            a, b = st[k][vis], rg[k][vis]
            r[f"geom_{k}_bitdiff"] = int((a.contiguous().view(torch.int32) != b.contiguous().view(torch.int32)).sum())
            r[f"geom_{k}_rel"] = pu.rel_err(a, b)
        if R == ours["num_rendered"] and R > 0:
            rb = ref_cuda.parse_binning(ref["binning"], R)
            keys_u, vals_u = debug.duplicate_with_keys(ours["geom"], Pn, W, H, R, ours["radii"])
            r["keys_unsorted_mismatch"] = int((keys_u != rb["point_list_keys_unsorted"]).sum())
            r["vals_unsorted_mismatch"] = int((vals_u != rb["point_list_unsorted"]).sum())
            r["keys_sorted_mismatch"] = int((st["point_list_keys"] != rb["point_list_keys"]).sum())
            r["point_list_mismatch"] = int((st["point_list"] != rb["point_list"]).sum())
            ri = ref_cuda.parse_image(ref["img"], W * H)
            nt = st["ranges"].shape[0]
            r["ranges_mismatch"] = int((st["ranges"] != ri["ranges"][:nt]).sum())
            r["n_contrib_mismatch"] = int((st["n_contrib"].reshape(2, -1) != ri["n_contrib"]).sum())
            r["final_T_rel"] = pu.rel_err(st["final_T"].reshape(3, -1), ri["accum_alpha"])
        r["color_rel"] = pu.rel_err(ours["color"], ref["color"])
        r["color_bitdiff"] = int((ours["color"].view(torch.int32) != ref["color"].view(torch.int32)).sum())
        for ch, nm in enumerate(["depth", "alpha", "nx", "ny", "nz", "median", "dist"]):
            r[f"allmap_{nm}_rel"] = pu.rel_err(ours["allmap"][ch], ref["allmap"][ch])
            r[f"allmap_{nm}_bitdiff"] = int(
                (ours["allmap"][ch].view(torch.int32) != ref["allmap"][ch].view(torch.int32)).sum())
        # backward
        gref = ref_cuda.backward(ref, scene, cam, bg, g["color"], g["allmap"])
        o = pu.run_ours(scene, cam, bg, grads=g)
        for k in ("means3D", "means2D", "opacity", "scales", "rotations", "sh"):
            r[f"grad_{k}_rel"] = pu.rel_err(o["grads"][k], gref[k].view_as(o["grads"][k]))
            r[f"grad_{k}_mism"] = pu.mismatch_frac(o["grads"][k], gref[k].view_as(o["grads"][k]), 1e-3)
        if do_time and vi == 0:
            r["t_ref_fwd_ms"] = timed(lambda: ref_cuda.forward(scene, cam, bg))
            r["t_ours_fwd_ms"] = timed(lambda: pu.run_ours_raw(scene, cam, bg))
            r["t_ref_fwdbwd_ms"] = timed(
                lambda: ref_cuda.backward(ref_cuda.forward(scene, cam, bg), scene, cam, bg, g["color"], g["allmap"]))
            r["t_ours_fwdbwd_ms"] = timed(lambda: pu.run_ours(scene, cam, bg, grads=g))
        res["views"].append(r)
        print(json.dumps(r), flush=True)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", nargs="+", default=["C1", "C2"])
    ap.add_argument("--views", type=int, default=2)
    ap.add_argument("--P", type=int, default=None)
    ap.add_argument("--out", default="gpurun_out/diag.json")
    a = ap.parse_args()
    Path(a.out).parent.mkdir(parents=True, exist_ok=True)
    allres = []
    for c in a.cfg:
        t0 = time.time()
        allres.append(diag_config(c, a.views, P=a.P))
        print(f"## {c} done in {time.time() - t0:.1f}s", flush=True)
        Path(a.out).write_text(json.dumps(allres, indent=1))


if __name__ == "__main__":
    main()
