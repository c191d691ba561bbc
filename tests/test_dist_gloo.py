"""world_size-2 gloo test (CPU) of the camera-sharded step: sharding covers every view once,
and the all-reduced gradients equal the single-process sum over all views."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_render_loss(params, view):
    # stand-in for render + loss: a smooth function of the replicated parameters and the camera index
    w = torch.cos(torch.arange(params["means3D"].shape[0], dtype=torch.float32) * (view + 1) * 0.1)
    return ((params["means3D"] * w[:, None]).sum() * (view + 1) + (params["shs"] ** 2).sum() * 0.01 * view +
            (params["opacities"] * params["scales"].sum(1, keepdim=True)).sum())


def _make_params():
    g = torch.Generator().manual_seed(0)
    return {"means3D": torch.randn(50, 3, generator=g).requires_grad_(True),
            "shs": torch.randn(50, 16, 3, generator=g).requires_grad_(True),
            "opacities": torch.rand(50, 1, generator=g).requires_grad_(True),
            "scales": torch.rand(50, 2, generator=g).requires_grad_(True),
            "rotations": torch.randn(50, 4, generator=g).requires_grad_(True)}


def _worker(rank, world, port, n_views, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from partgs_b200.dist import GradAllReducer, sharded_step
    params = _make_params()
    total, mine = sharded_step(lambda v: _fake_render_loss(params, v), params, n_views, GradAllReducer())
    out[rank] = (mine, {k: (p.grad.clone() if p.grad is not None else None) for k, p in params.items()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_views", [7, 1])
def test_sharded_step_matches_single_process(n_views):
    from partgs_b200.dist import shard_views
    world = 2
    assert sorted(shard_views(n_views, 0, world) + shard_views(n_views, 1, world)) == list(range(n_views))
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_views, out), nprocs=world, join=True)
    # single-process reference
    params = _make_params()
    for v in range(n_views):
        _fake_render_loss(params, v).backward()
    for rank in range(world):
        mine, grads = out[rank]
        assert mine == list(range(rank, n_views, world))
        for k in ("means3D", "shs", "opacities", "scales"):
            assert torch.allclose(grads[k], params[k].grad, rtol=1e-5, atol=1e-6), (rank, k)
        assert grads["rotations"] is not None and float(grads["rotations"].abs().max()) == 0.0


def test_shard_views_validation():
    from partgs_b200.dist import shard_views
    assert shard_views(49, 3, 8) == [3, 11, 19, 27, 35, 43]
    assert shard_views(2, 5, 8) == []
    with pytest.raises(ValueError):
        shard_views(4, 2, 2)


def test_balanced_view_schedule():
    from partgs_b200.dist import balanced_view_schedule
    import random
    rnd = random.Random(0)
    for n, world in ((49, 8), (64, 8), (7, 2), (3, 4), (1, 1), (10, 4)):
        costs = [rnd.random() for _ in range(n)]
        steps = balanced_view_schedule(costs, world)
        assert all(len(s) == world for s in steps)
        assert sorted(set(v for s in steps for v in s)) == list(range(n))       # every view is rendered
        if n >= world:
            spread = max(max(costs[v] for v in s) - min(costs[v] for v in s) for s in steps)
            assert spread <= max(costs) - min(costs)
        # the most expensive view is in the first step, the cheapest in the second (expensive / cheap alternate)
        assert max(range(n), key=lambda i: costs[i]) in steps[0]
        if len(steps) > 1:
            assert min(range(n), key=lambda i: costs[i]) in steps[1]
    with pytest.raises(ValueError):
        balanced_view_schedule([], 2)
