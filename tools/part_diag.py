"""_part fork: per-channel bit differences vs the reference build (diagnostic)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from oracle import ref_cuda
import test_gpu_part_raster as T
scene, cams, bg, g = T._setup(20_000, 400, 300, 16)
cam = cams[0]
ref = ref_cuda.forward_part(scene, cam, bg)
o = T.run_ours(scene, cam, bg, g)
def bd(a, b): return int((a.contiguous().view(torch.int32) != b.contiguous().view(torch.int32)).sum())
print("color bitdiff", bd(o["color"], ref["color"]), "semantic", bd(o["semantic"], ref["semantic"]))
for ch in range(8):
    a, b = o["allmap"][ch], ref["allmap"][ch]
    d = (a - b).abs()
    i = int(d.argmax())
    print("allmap", ch, "bitdiff", bd(a, b), "maxabs", float(d.max()), "max|ref|", float(b.abs().max()),
          "at", divmod(i, 400), "ours", float(a.flatten()[i]), "ref", float(b.flatten()[i]))
# run the reference twice: is it deterministic?
ref2 = ref_cuda.forward_part(scene, cam, bg)
print("ref self bitdiff ch6", bd(ref["allmap"][6], ref2["allmap"][6]))
gref = ref_cuda.backward_part(ref, scene, cam, bg, g["color"], g["semantic"], g["allmap"])
import parity_utils as pu
for k in ("means3D", "means2D", "opacity", "scales", "rotations", "sh", "semantics"):
    print("grad", k, pu.rel_err(o["grads"][k], gref[k].view_as(o["grads"][k])))
