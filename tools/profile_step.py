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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="C3")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--P", type=int, default=None)
    ap.add_argument("--ref", action="store_true", help="run the reference CUDA build (oracle/_ref) instead")
    ap.add_argument("--fwd-only", action="store_true")
    a = ap.parse_args()
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
