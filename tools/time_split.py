"""Per-iteration forward / backward device times (diagnostic)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from partgs_b200 import synth  # noqa: E402
from partgs_b200.diff_surfel_rasterization import _C  # noqa: E402
import parity_utils as pu  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
views = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg, scene, cams = synth.make_config(name, device="cuda", views=views)
bg = torch.zeros(3, device="cuda")
g = synth.upstream_grads(cfg["W"], cfg["H"], synth.SEED_BASE, device="cuda")
empty = torch.empty(0, device="cuda")
from partgs_b200 import _lib
_lib.timing_enable(True)
for ci, cam in enumerate(cams[:2]):
    for it in range(6):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        R, color, others, radii, geom, binning, img = _C.rasterize_gaussians(
            bg, scene["means3D"], empty, scene["opacities"], scene["scales"], scene["rotations"], 1.0, empty,
            cam.viewmatrix, cam.projmatrix, cam.tanfovx, cam.tanfovy, cam.image_height, cam.image_width, scene["shs"],
            3, cam.campos, False, False)
        ev[1].record()
        _C.rasterize_gaussians_backward(bg, scene["means3D"], radii, empty, scene["scales"], scene["rotations"], 1.0,
                                        empty, cam.viewmatrix, cam.projmatrix, cam.tanfovx, cam.tanfovy, g["color"],
                                        g["allmap"], scene["shs"], 3, cam.campos, geom, R, binning, img, False)
        ev[2].record()
        torch.cuda.synchronize()
        st = _lib.timing_read(reset=True)
        print("   ", {k: round(v[0], 3) for k, v in st.items() if v[1]})
        print(name, "cam", ci, "it", it, "R", R,
              "fwd %.3f ms  bwd %.3f ms" % (ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])), flush=True)
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pu.run_ours(scene, cam, bg, grads=g)
        e1.record()
        torch.cuda.synchronize()
        print(name, "cam", ci, "autograd fwd+bwd %.3f ms" % e0.elapsed_time(e1), flush=True)
