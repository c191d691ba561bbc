"""Calibration of the element-wise parity gates on hardware: how many elements of (a) our images and (b) our gradients
differ from the reference CUDA build, and how much two runs of the REFERENCE differ from each other (its 16 float atomics
per fragment commit in a different order every run).  usage: python tools/grad_gate_probe.py [--cfgs C1,C2,C3:200000]"""
import argparse, json, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from partgs_b200 import synth  # noqa: E402
import parity_utils as pu  # noqa: E402
from oracle import ref_cuda  # noqa: E402

ap = argparse.ArgumentParser(); ap.add_argument("--cfgs", default="C1,C2,C3:200000")
a = ap.parse_args()
for spec in a.cfgs.split(","):
    name, _, P = spec.partition(":")
    cfg, scene, cams = synth.make_config(name, device="cuda", P=int(P) if P else None, views=1)
    cam = cams[0]; bg = torch.tensor([0.1, 0.2, 0.3], device="cuda")
    g = synth.upstream_grads(cfg["W"], cfg["H"], synth.SEED_BASE, device="cuda")
    ref = ref_cuda.forward(scene, cam, bg)
    g1 = ref_cuda.backward(ref, scene, cam, bg, g["color"], g["allmap"])
    g2 = ref_cuda.backward(ref, scene, cam, bg, g["color"], g["allmap"])
    o = pu.run_ours(scene, cam, bg, grads=g)
    row = dict(cfg=spec, color_neq=int((o["color"] != ref["color"]).sum()), allmap_neq=int((o["allmap"] != ref["allmap"]).sum()))
    for k in ("means3D", "means2D", "opacity", "scales", "rotations", "sh"):
        mine, r1, r2 = o["grads"][k], g1[k].view_as(o["grads"][k]), g2[k].view_as(o["grads"][k])
        row[k] = dict(ours=[pu.grad_violations(mine, r1, rt, 1e-6) for rt in (1e-4, 1e-3)],
                      ref_vs_ref=[pu.grad_violations(r2, r1, rt, 1e-6) for rt in (1e-4, 1e-3)],
                      ours_floor1e5=pu.grad_violations(mine, r1, 1e-4, 1e-5), rel=pu.rel_err(mine, r1))
    print(json.dumps(row))
