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


def _patch_emulated_library():
    import contextlib
    import build as emu_build
    from partgs_b200 import _lib
    lib = C.CDLL(str(emu_build.build_full()))
    for name, (res, args) in _lib.SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib._lib = lib
    _lib.on_device = lambda t: isinstance(t, torch.Tensor)
    _lib.current_stream = lambda device: None

    class _Null(contextlib.nullcontext):
        def __init__(self, *a, **k):
            super().__init__()
    torch.cuda.device = _Null


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
