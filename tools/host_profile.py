"""cProfile of the host side of bench steps (where does the CPU time of one fwd+bwd step go?)."""
import cProfile, pstats, sys, time, io
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench
from partgs_b200 import synth
dev = torch.device("cuda:0")
cfg = dict(synth.CONFIGS["C3"]); seed = synth.SEED_BASE + 2
scene = synth.make_point_scene(cfg["P"], seed, S=0, device=dev)
cams = synth.make_cameras(8, cfg["W"], cfg["H"], seed, device=dev)
bg = torch.zeros(3, device=dev); g = synth.upstream_grads(cfg["W"], cfg["H"], synth.SEED_BASE, device=dev)
arm = bench.OursArm(scene, dev)
for i in range(16):
    arm.step(cams[i % 8], bg, g)
torch.cuda.synchronize()
# CPU time of the enqueue path alone: make the GPU the bottleneck-free side by timing thread CPU time
t0 = time.thread_time(); w0 = time.perf_counter()
for i in range(40):
    arm.step(cams[i % 8], bg, g)
cpu_ms = (time.thread_time() - t0) * 1e3 / 40; wall_ms = (time.perf_counter() - w0) * 1e3 / 40
torch.cuda.synchronize()
print(f"main-thread CPU time per step {cpu_ms:.3f} ms, wall {wall_ms:.3f} ms")
pr = cProfile.Profile(); pr.enable()
for i in range(40):
    arm.step(cams[i % 8], bg, g)
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18); print(s.getvalue()[:3500])
