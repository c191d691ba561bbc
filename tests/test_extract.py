"""Extraction loop, CPU side: the epilogue restatement against golden vectors of the
reference's own partmap_to_rgbmap / estimate_bounding_sphere, the palette restatement's invariants, and the
world_size-2 gloo gather that puts camera-sharded host stacks back into view order."""
import os
import socket
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import extract_oracle

Z = np.load(Path(__file__).parent / "golden" / "extract_maps.npz")


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_oracle_reproduces_reference_partmap_to_rgbmap(name):
    part, pal = torch.from_numpy(Z[f"{name}_part"]), torch.from_numpy(Z[f"{name}_palette"])
    assert torch.equal(extract_oracle.partmap_to_rgbmap(part, pal), torch.from_numpy(Z[f"{name}_rgb"]))
    # the fixtures cover background pixels, clamped values and ties
    clamped = part.clamp(0, 1)
    assert (clamped.sum(0) < 0.1).any() and (part > 1).any() and (part < 0).any()


def test_bounding_sphere_matches_reference():
    from partgs_b200.extract import bounding_sphere
    center, radius = bounding_sphere([torch.from_numpy(w) for w in Z["wvt"]])
    assert np.allclose(center.astype(np.float32), Z["center"], rtol=0, atol=1e-6)
    assert abs(radius - float(Z["radius"])) <= 1e-9


def test_fancy_palette_invariants():
    from partgs_b200.extract import fancy_palette
    p = fancy_palette(17)
    assert p.shape == (17, 3) and p.dtype == torch.float32
    assert float(p.min()) >= 0 and float(p.max()) <= 1
    assert len({tuple(r) for r in p.tolist()}) == 17            # distinct part colours
    # the colour list ends with hls[1]; value 1.0 maps to the last lookup entry = that colour
    import colorsys
    last = colorsys.hls_to_rgb((1 / 21 + 0.01) % 1, 0.6, 0.65)
    assert np.allclose(p[-1].numpy(), last, atol=1e-6)
    # nested prefixes are NOT equal (linspace depends on num): the palette must be built for S+1, as the reference does
    assert not torch.equal(fancy_palette(5), p[:5])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _view_maps(v):
    g = torch.Generator().manual_seed(100 + v)
    return {"rgbmaps": torch.rand(3, 4, 5, generator=g), "depthmaps": torch.rand(1, 4, 5, generator=g),
            "normals": torch.rand(3, 4, 5, generator=g)}


def _worker(rank, world, port, V, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from partgs_b200.extract import GaussianExtractor
    ex = GaussianExtractor(None, lambda *a, **k: None, None, rank=rank, world=world, device="cpu")
    mine = ex.my_views(V)
    stacks = {}
    if mine:
        per_view = [_view_maps(v) for v in mine]
        stacks = {k: torch.stack([m[k] for m in per_view]) for k in per_view[0]}
    else:
        stacks = {}   # what reconstruction() has on a rank without views
    ex._collect(stacks, mine, V)
    if rank == 0:
        out["maps"] = {k: getattr(ex, k) for k in _view_maps(0)}
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("V", [5, 1, 4])
def test_sharded_stacks_are_gathered_in_view_order(V):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), V, out), nprocs=2, join=True)
    maps = out["maps"]
    for k in ("rgbmaps", "depthmaps", "normals"):
        want = torch.stack([_view_maps(v)[k] for v in range(V)])
        assert torch.equal(maps[k], want), k


def test_single_rank_collect_keeps_the_stacks():
    from partgs_b200.extract import GaussianExtractor
    ex = GaussianExtractor(None, lambda *a, **k: None, None, device="cpu")
    s = {"rgbmaps": torch.rand(2, 3, 4, 4)}
    ex._collect(s, [0, 1], 2)
    assert ex.rgbmaps is s["rgbmaps"] and ex.depthmaps == []
