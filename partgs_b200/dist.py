"""Camera-sharded multi-GPU rendering / training step.

PartGS itself is single-process (train.py renders one random view per iteration); views are
independent given replicated surfel parameters, so the batch of views shards by camera with
no data-path exchange inside forward/backward.  One process per GPU (torch.distributed,
NCCL over NVLink/NVSwitch); each rank renders its views locally and accumulates parameter
gradients, then the gradients are summed across ranks — one all-reduce per parameter group
(xyz 12 B, SH 192 B, opacity 4 B, scale 8 B, rotation 16 B = 232 B/surfel), launched
asynchronously per group so it overlaps the tail of the local work.  The same code runs on
CPU with the gloo backend (used by the world_size-2 tests).
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Sequence

import torch
import torch.distributed as dist


def shard_views(n_views: int, rank: int, world: int) -> List[int]:
    """Round-robin camera assignment: rank r renders views r, r+world, r+2*world, ..."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_views, world))


def grad_bucket(grads: Iterable[torch.Tensor]):
    """If all gradients are views into one flat buffer (the rasteriser's backward carves its five
    parameter gradients out of a single allocation), return that buffer so that one collective
    reduces everything; otherwise None."""
    base = None
    for g in grads:
        if g is None:
            continue
        b = g._base
        if b is None or (base is not None and b is not base):
            return None
        base = b
    return base


class GradAllReducer:
    """Asynchronous per-parameter-group gradient all-reduce (SUM)."""

    def __init__(self, group=None, average: bool = False):
        self.group = group
        self.average = average
        self._pending = []

    @property
    def world(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def launch(self, tensors: Iterable[torch.Tensor]):
        if self.world == 1:
            return
        tensors = list(tensors)
        bucket = grad_bucket(tensors)
        if bucket is not None:
            tensors = [bucket]  # one collective for the whole 232 B/surfel bucket
        for t in tensors:
            if t is None:
                continue
            h = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._pending.append((h, t))

    def wait(self):
        w = self.world
        for h, t in self._pending:
            h.wait()
            if self.average:
                t.div_(w)
        self._pending.clear()


def sharded_step(render_loss: Callable[[int], torch.Tensor], params: Dict[str, torch.Tensor], n_views: int,
                 reducer: GradAllReducer | None = None, rank: int | None = None, world: int | None = None,
                 order: Sequence[str] = ("means3D", "shs", "opacities", "scales", "rotations")):
    """One data-parallel step over ``n_views`` cameras.

    ``render_loss(view_index)`` must render that view from ``params`` and return a scalar loss.
    Every rank back-propagates its own views (gradients accumulate in ``params[k].grad``) and the
    accumulated gradients are then summed over ranks.  Returns (local_loss_sum, my_view_indices).
    """
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    reducer = reducer or GradAllReducer()
    for p in params.values():
        p.grad = None
    mine = shard_views(n_views, rank, world)
    total = None
    for v in mine:
        loss = render_loss(v)
        loss.backward()
        total = loss.detach() if total is None else total + loss.detach()
    grads = []
    for k in order:
        p = params.get(k)
        if p is None:
            continue
        if p.grad is None:            # a rank without views still has to join the collective
            p.grad = torch.zeros_like(p)
        grads.append(p.grad)
    reducer.launch(grads)
    reducer.wait()
    return total, mine
