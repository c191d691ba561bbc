"""distCUDA2 timing: ours vs the reference simple-knn build (oracle/_ref), several cloud shapes and sizes."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from partgs_b200 import synth
from partgs_b200.simple_knn._C import distCUDA2
from oracle import ref_cuda

ref = ref_cuda.load("ref_knn_C") if ref_cuda.available("ref_knn_C") else None
dev = "cuda"


def timed(fn, x, n=5):
    fn(x); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn(x)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


if "--profile" in sys.argv:      # target for ncu: two calls on the 1 M synthetic scene
    pts = synth.make_point_scene(1_000_000, 42, device=dev)["means3D"].contiguous()
    for _ in range(2):
        distCUDA2(pts)
        torch.cuda.synchronize()
    sys.exit(0)
gen = torch.Generator().manual_seed(0)
for P in [int(a) for a in sys.argv[1:] if a.isdigit()] or (100_000, 1_000_000, 3_000_000):
    clouds = {"synth_scene": synth.make_point_scene(P, 42, device=dev)["means3D"],
              "blob": torch.randn(P, 3, generator=gen).to(dev), "uniform": torch.rand(P, 3, generator=gen).to(dev)}
    v = torch.randn(P, 3, generator=gen)
    clouds["sphere_surface"] = (v / v.norm(dim=1, keepdim=True)).to(dev)
    for name, pts in clouds.items():
        pts = pts.contiguous()
        row = {"cloud": name, "P": P, "ours_ms": round(timed(distCUDA2, pts), 4)}
        if ref is not None:
            row["reference_ms"] = round(timed(ref.distCUDA2, pts, n=2), 3)
            a, b = distCUDA2(pts), ref.distCUDA2(pts)
            row["max_rel_diff"] = float(((a - b).abs() / b.abs().clamp_min(1e-30)).max())
        print(json.dumps(row), flush=True)
