"""Which of the reference renderer's maps carries the upstream gradient that separates our backward from the reference's?"""
import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import parity_utils as pu
import test_gpu_reference_callers as T
from types import SimpleNamespace
from partgs_b200 import synth
A, B = T._renderers("reference"), T._renderers("dropin")
P, W, H, dr = 60000, 400, 300, 0.0
scene = synth.make_point_scene(P, seed=41, device="cuda")
cam = synth.make_cameras(1, W, H, seed=42, device="cuda")[0]
bg = torch.tensor([0.1, 0.2, 0.3], device="cuda")
pipe = SimpleNamespace(depth_ratio=dr, compute_cov3D_python=False, convert_SHs_python=False, debug=False)
def run(fn, key):
    pc, t, sem = T._pc(scene)
    r = fn(cam, pc, pipe, bg)
    gen = torch.Generator(device="cpu").manual_seed(7)
    w = torch.randn(r[key].shape, generator=gen).to("cuda")
    (r[key] * w).sum().backward()
    return {k: t[k].grad for k in T.PARAMS}
for key in ("render", "rend_alpha", "rend_normal", "rend_dist", "surf_depth", "surf_normal"):
    ga, ga2, gb = run(A.render, key), run(A.render, key), run(B.render, key)
    print(key, {k: (pu.grad_violations(gb[k], ga[k]), pu.grad_violations(ga2[k], ga[k]), float(ga[k].abs().max())) for k in ("means3D", "opacities", "scales")})
