"""Where does a bench step's wall time go?  device allocs, CPU time per step, GPU time per step."""
import sys, time, json
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench
from partgs_b200 import synth, _lib
dev = torch.device("cuda:0")
cfg = dict(synth.CONFIGS["C3"]); seed = synth.SEED_BASE + synth.CONFIG_INDEX["C3"]
scene = synth.make_point_scene(cfg["P"], seed, S=0, device=dev)
cams = synth.make_cameras(cfg["views"], cfg["W"], cfg["H"], seed, device=dev)
bg = torch.zeros(3, device=dev); g = synth.upstream_grads(cfg["W"], cfg["H"], synth.SEED_BASE, device=dev)
arm = bench.OursArm(scene, dev)
for i in range(49):
    arm.step(cams[i % 49], bg, g)
torch.cuda.synchronize()
s0 = torch.cuda.memory_stats()
_lib.timing_enable(True); _lib.timing_read(reset=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
cpu = []
for i in range(30):
    a = time.perf_counter()
    arm.step(cams[i % 49], bg, g)
    cpu.append(time.perf_counter() - a)
e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
s1 = torch.cuda.memory_stats()
st = _lib.timing_read(reset=True)
keys = ["num_device_alloc", "num_device_free", "num_alloc_retries", "allocation.all.allocated", "segment.all.allocated"]
print(json.dumps({"wall_ms_per_step": (t1 - t0) * 1e3 / 30, "gpu_ms_per_step": e0.elapsed_time(e1) / 30,
                  "cpu_ms_per_step_median": sorted(cpu)[15] * 1e3, "cpu_ms_max": max(cpu) * 1e3,
                  "mem_delta": {k: s1.get(k, 0) - s0.get(k, 0) for k in keys},
                  "reserved_GB": torch.cuda.memory_reserved() / 1e9,
                  "stage_ms": {k: round(v[0] / 30, 4) for k, v in st.items() if v[1]}}))
