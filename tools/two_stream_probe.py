"""Does running consecutive (independent) views on two alternating CUDA streams hide the latency-bound binning chain of
view k+1 behind the render kernels of view k?  N = 1, C3, eager launches with the lazy instance count and as per-stream
CUDA graphs."""
import sys, time, json
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import parity_utils as pu
from partgs_b200 import synth, diff_surfel_rasterization as dsr
from partgs_b200.diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer
dev = torch.device("cuda")
cfg, scene, cams = synth.make_config("C3", device=dev)
W, H = cfg["W"], cfg["H"]
bg = torch.zeros(3, device=dev)
g = synth.upstream_grads(W, H, synth.SEED_BASE, device=dev)
params = {k: scene[k].clone().requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "shs")}


def make_step(cam):
    m2d = torch.zeros_like(params["means3D"], requires_grad=True)
    st = GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=bg,
                                       scale_modifier=1.0, viewmatrix=cam.viewmatrix, projmatrix=cam.projmatrix,
                                       sh_degree=3, campos=cam.campos, prefiltered=False, debug=False)

    def step():
        for t in params.values():
            t.grad = None
        m2d.grad = None
        color, radii, allmap = GaussianRasterizer(st)(means3D=params["means3D"], means2D=m2d, opacities=params["opacities"],
                                                      shs=params["shs"], scales=params["scales"], rotations=params["rotations"])
        torch.autograd.backward([color, allmap], [g["color"], g["allmap"]])
    return step


K = 40
views = [(i * 5) % len(cams) for i in range(K)]
for c in cams:
    make_step(c)()
torch.cuda.synchronize()
dsr.set_lazy_count(True)


def timed(run):
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K


def eager(nstreams):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    steps = [make_step(cams[v]) for v in views]

    def run():
        cur = torch.cuda.current_stream()
        for s in streams:
            s.wait_stream(cur)
        for i, st in enumerate(steps):
            with torch.cuda.stream(streams[i % nstreams]):
                st()
        for s in streams:
            cur.wait_stream(s)
    return run


out = {"eager_1_stream_ms": timed(eager(1)), "eager_2_streams_ms": timed(eager(2)), "eager_3_streams_ms": timed(eager(3))}
dsr.set_lazy_count(False); dsr.resolve_count()

# graphs: one per (stream, static camera slot)
def graphs(nstreams):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    slots = []
    for s in range(nstreams):
        cam = cams[0]
        gc = type(cam)(cam.image_width, cam.image_height, cam.tanfovx, cam.tanfovy, cam.viewmatrix.clone(),
                       cam.projmatrix.clone(), cam.campos.clone())
        step = make_step(gc)
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step(); step()
        torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            step()
        slots.append((gc, gr))

    def run():
        cur = torch.cuda.current_stream()
        for s in streams:
            s.wait_stream(cur)
        for i, v in enumerate(views):
            gc, gr = slots[i % nstreams]
            with torch.cuda.stream(streams[i % nstreams]):
                gc.viewmatrix.copy_(cams[v].viewmatrix); gc.projmatrix.copy_(cams[v].projmatrix); gc.campos.copy_(cams[v].campos)
                gr.replay()
        for s in streams:
            cur.wait_stream(s)
    return run


out.update({"graph_1_stream_ms": timed(graphs(1)), "graph_2_streams_ms": timed(graphs(2)), "graph_3_streams_ms": timed(graphs(3))})
print(json.dumps({k: round(v, 4) for k, v in out.items()}))
