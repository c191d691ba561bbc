"""Does a badly scaled upstream gradient (one channel 1e4 x the others) separate our backward from the reference's?"""
import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import parity_utils as pu
from partgs_b200 import synth
from oracle import ref_cuda
cfg, scene, cams = synth.make_config("C2", device="cuda", P=60000, views=1)
cam = cams[0]; bg = torch.tensor([0.1, 0.2, 0.3], device="cuda")
g0 = synth.upstream_grads(cfg["W"], cfg["H"], synth.SEED_BASE, device="cuda")
ref = ref_cuda.forward(scene, cam, bg)
for ch, scale in ((None, 1.0), (0, 1e4), (1, 1e4), (2, 1e4), (5, 1e4), (6, 1e4), ("color", 1e4)):
    g = {k: v.clone() for k, v in g0.items()}
    if ch == "color": g["color"] *= scale
    elif ch is not None: g["allmap"][ch] *= scale
    g1 = ref_cuda.backward(ref, scene, cam, bg, g["color"], g["allmap"])
    g2 = ref_cuda.backward(ref, scene, cam, bg, g["color"], g["allmap"])
    o = pu.run_ours(scene, cam, bg, grads=g)
    print(ch, scale, {k: (pu.grad_violations(o["grads"][k], g1[k].view_as(o["grads"][k])), pu.grad_violations(g2[k], g1[k])) for k in ("means3D", "opacity", "scales", "sh")})
