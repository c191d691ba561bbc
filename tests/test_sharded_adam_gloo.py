"""world_size-2 gloo test of ShardedFusedAdam (reduce-scatter -> Adam on my slice -> all-gather): after every step all
ranks hold the parameters that one torch.optim.Adam produces from the rank-summed gradients.  The Adam kernel runs from
the emulator build of the library (CPU tensors), so the whole path — flat views, slice / parameter intersections,
per-group learning rates, collectives — is exercised without a GPU."""
import ctypes as C
import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, str(Path(__file__).parent / "cuda_emu"))

SHAPES = {"xyz": (37, 3), "f_dc": (37, 1, 3), "f_rest": (37, 15, 3), "opacity": (37, 1), "scaling": (37, 2),
          "rotation": (37, 4)}
LRS = {"xyz": 1.6e-3, "f_dc": 2.5e-2, "f_rest": 1.25e-3, "opacity": 0.05, "scaling": 0.02, "rotation": 0.0}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _init(seed=0):
    g = torch.Generator().manual_seed(seed)
    return {k: torch.randn(s, generator=g) for k, s in SHAPES.items()}


def _grad(k, rank, step):
    g = torch.Generator().manual_seed(1000 * step + 10 * rank + len(k))
    return torch.randn(SHAPES[k], generator=g) * (0.1 if step != 1 else 10.0)


def _patch_emulated_library(monkeypatch=None):
    """In a spawned rank: plain assignment (the process ends with the test).  In the pytest process: through
    `monkeypatch`, so that nothing leaks into other tests."""
    import contextlib
    import build as emu_build
    from partgs_b200 import _lib
    lib = C.CDLL(str(emu_build.build_full()))
    for name, (res, args) in _lib.SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args

    class _Null(contextlib.nullcontext):
        def __init__(self, *a, **k):
            super().__init__()
    put = (lambda o, n, v: setattr(o, n, v)) if monkeypatch is None else monkeypatch.setattr
    put(_lib, "_lib", lib)
    put(_lib, "on_device", lambda t: isinstance(t, torch.Tensor))
    put(_lib, "current_stream", lambda device: None)
    put(torch.cuda, "device", _Null)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _patch_emulated_library()
    from partgs_b200.optim import ShardedFusedAdam
    params = {k: torch.nn.Parameter(v.clone()) for k, v in _init().items()}
    opt = ShardedFusedAdam([{"params": [params[k]], "lr": LRS[k], "name": k} for k in SHAPES])
    assert opt.exp_avg.numel() * world == opt.numel           # 1/W of the optimiser state per rank
    history = []
    for step in range(4):
        opt.zero_grad()
        for k, p in params.items():
            (p * _grad(k, rank, step)).sum().backward()       # autograd accumulates into the flat views
        if step == 2:
            for g in opt.param_groups:                        # the reference's schedulers rewrite lr
                g["lr"] *= 0.5
        opt.step()
        history.append({k: p.detach().clone() for k, p in params.items()})
    out[rank] = history
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_adam_equals_adam_on_the_summed_gradients():
    import build as emu_build
    try:
        emu_build.build_full()                                # build once, before the ranks race for it
    except emu_build.EmuUnavailable as ex:
        pytest.skip(str(ex))
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    ref = {k: torch.nn.Parameter(v.clone()) for k, v in _init().items()}
    opt = torch.optim.Adam([{"params": [ref[k]], "lr": LRS[k]} for k in SHAPES], lr=0.0, eps=1e-15, foreach=False)
    for step in range(4):
        for k, p in ref.items():
            p.grad = sum(_grad(k, r, step) for r in range(world))
        if step == 2:
            for g in opt.param_groups:
                g["lr"] *= 0.5
        opt.step()
        for rank in range(world):
            for k in SHAPES:
                got, want = out[rank][step][k], ref[k].detach()
                assert float((got - want).abs().max()) <= 3e-6 * float(want.abs().max()), (step, rank, k)
    assert torch.equal(out[0][3]["rotation"], _init()["rotation"])      # lr 0: untouched on every rank
    for k in SHAPES:
        assert torch.equal(out[0][3][k], out[1][3][k])                 # replicas stay bit-identical


# ---------------------------------------------------------------------------------------------------------------
# the camera-sharded step of SURVEY 8(e) with the REAL rasteriser (emulated) on two gloo ranks
# ---------------------------------------------------------------------------------------------------------------
def _render_loss(params, cam, upstream):
    from partgs_b200.diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    settings = GaussianRasterizationSettings(
        image_height=cam.image_height, image_width=cam.image_width, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
        bg=torch.zeros(3), scale_modifier=1.0, viewmatrix=cam.viewmatrix, projmatrix=cam.projmatrix, sh_degree=3,
        campos=cam.campos, prefiltered=False, debug=False)
    means2D = torch.zeros_like(params["means3D"], requires_grad=True)
    color, _radii, allmap = GaussianRasterizer(settings)(
        means3D=params["means3D"], means2D=means2D, opacities=params["opacities"], shs=params["shs"],
        scales=params["scales"], rotations=params["rotations"])
    return (color * upstream["color"]).sum() + (allmap * upstream["allmap"]).sum()


def _scene_and_views(n_views):
    from partgs_b200 import synth
    scene = synth.make_point_scene(120, seed=3, device="cpu")
    scene["scales"] = scene["scales"] * 3.0
    cams = synth.make_cameras(n_views, 32, 16, seed=4, device="cpu")
    ups = [synth.upstream_grads(32, 16, 50 + v, device="cpu") for v in range(n_views)]
    return scene, cams, ups


def _raster_worker(rank, world, port, n_views, bucketed, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _patch_emulated_library()
    from partgs_b200.dist import GradAllReducer, NcclBucketAllReducer, sharded_step
    from partgs_b200 import diff_surfel_rasterization as dsr
    scene, cams, ups = _scene_and_views(max(n_views, 1))
    params = {k: scene[k].clone().requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    if bucketed:
        # the batch mode of SURVEY 8(e): the backward kernel accumulates the views of this rank in ONE bucket, which a
        # single collective reduces (here over gloo; PeerGradAllReducer is the NVLink version of the same class)
        reducer = NcclBucketAllReducer(dsr.bucket_numel(120, 16), "cpu")
        dsr.set_grad_bucket_provider(reducer.bucket_provider)
    else:
        reducer = GradAllReducer()
    total, mine = sharded_step(lambda v: _render_loss(params, cams[v], ups[v]), params, n_views, reducer)
    if bucketed:
        from partgs_b200.dist import grad_bucket
        flat = grad_bucket([params[k].grad for k in ("means3D", "shs", "opacities", "scales", "rotations")])
        assert flat is not None and flat.untyped_storage().data_ptr() == reducer.current().untyped_storage().data_ptr()
    out[rank] = (mine, {k: p.grad.clone() for k, p in params.items()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_views,bucketed", [(3, False), (5, True), (1, True), (1, False)])
def test_camera_sharded_step_with_the_real_rasteriser(monkeypatch, n_views, bucketed):
    """Views shard by camera, parameters are replicated, the five parameter gradients are summed over the ranks:
    equal to one process back-propagating all views — with the product's own kernels (emulated).
    bucketed: gradients accumulate inside the backward kernel in one bucket per batch (rank 0 renders three views,
    rank 1 two), one collective.  n_views = 1: rank 1 has no view and must still join with the same layout."""
    import build as emu_build
    try:
        emu_build.build_full()
    except emu_build.EmuUnavailable as ex:
        pytest.skip(str(ex))
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_raster_worker, args=(world, _free_port(), n_views, bucketed, out), nprocs=world, join=True)
    _patch_emulated_library(monkeypatch)
    scene, cams, ups = _scene_and_views(n_views)
    params = {k: scene[k].clone().requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    for v in range(n_views):
        _render_loss(params, cams[v], ups[v]).backward()
    assert out[0][0] == list(range(0, n_views, 2)) and out[1][0] == list(range(1, n_views, 2))
    for rank in range(world):
        for k, p in params.items():
            got = out[rank][1][k]
            assert float((got - p.grad).abs().max()) <= 2e-5 * (float(p.grad.abs().max()) + 1e-12), (rank, k)
