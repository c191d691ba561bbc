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


def init_nccl(local_rank: int, max_ctas: int | None = 8):
    """One process per GPU.  The gradient all-reduce runs concurrently with the next view's render kernels
    (which are issue-bound and want every SM): NCCL's default of up to 32 channels = 32 resident CTAs takes
    a fifth of the GPU away from them for the length of the collective.  The collective only has to finish
    within one step (232 MB in ~3 ms, i.e. ~150 GB/s of bus bandwidth at 8 ranks), which a few CTAs sustain over
    NVLink 5, so the communicator is capped at `max_ctas` (PGS_NCCL_MAX_CTAS overrides; 0 = NCCL default)."""
    import os
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local_rank)
    env = os.environ.get("PGS_NCCL_MAX_CTAS")
    if env is not None:
        max_ctas = int(env)
    kw = {}
    if max_ctas:
        opts = dist.ProcessGroupNCCL.Options()
        opts.config.max_ctas = int(max_ctas)
        opts.config.min_ctas = 1
        kw["pg_options"] = opts
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), **kw)


def shard_views(n_views: int, rank: int, world: int) -> List[int]:
    """Round-robin camera assignment: rank r renders views r, r+world, r+2*world, ..."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_views, world))


def balanced_view_schedule(costs: Sequence[float], world: int) -> List[List[int]]:
    """Lock-step data parallelism ends every step with the slowest rank, and views differ a lot in cost (instance
    count, depth complexity).  Given a per-view cost (measured step time), return steps of `world` view indices
    with similar cost: views sorted by cost, dealt in consecutive groups; when the view count is not a multiple
    of `world` the last group is completed with its cheapest neighbours (every view still appears); expensive and
    cheap steps alternate.  Rank r renders ``schedule[k % len(schedule)][r]`` in step k."""
    n = len(costs)
    if n == 0 or world <= 0:
        raise ValueError("need at least one view and one rank")
    order = sorted(range(n), key=lambda i: -float(costs[i]))
    steps = []
    for s in range(0, n, world):
        grp = order[s:s + world]
        if len(grp) < world:                      # complete the last step with the views just before it
            fill = order[max(0, s - (world - len(grp))):s] or order[:1]
            grp = (fill + grp + fill * world)[:world] if len(fill) + len(grp) < world else fill[-(world - len(grp)):] + grp
        steps.append(grp)
    # alternate expensive and cheap steps so that any window of consecutive steps sees an average mix
    lo, hi, mixed = 0, len(steps) - 1, []
    while lo <= hi:
        mixed.append(steps[lo])
        if hi != lo:
            mixed.append(steps[hi])
        lo, hi = lo + 1, hi - 1
    return mixed


def grad_bucket(grads: Iterable[torch.Tensor]):
    """If all gradients live in one allocation (the rasteriser's backward carves its five parameter
    gradients out of a single flat buffer), return one flat tensor covering them so that a single
    collective reduces everything; otherwise None.  (autograd hands `.grad` over detached, so the
    common allocation is recognised by the shared storage, not by `._base`.)"""
    gs = [g for g in grads if g is not None]
    if not gs:
        return None
    st = gs[0].untyped_storage()
    lo, hi = None, None
    for g in gs:
        if (g.untyped_storage().data_ptr() != st.data_ptr() or g.dtype != gs[0].dtype or not g.is_contiguous()):
            return None
        a, b = g.storage_offset(), g.storage_offset() + g.numel()
        lo = a if lo is None else min(lo, a)
        hi = b if hi is None else max(hi, b)
    return torch.empty(0, dtype=gs[0].dtype, device=gs[0].device).set_(st, lo, (hi - lo,))


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


class PeerGradAllReducer:
    """Gradient all-reduce (SUM) over NVLink peer memory that uses the COPY ENGINES instead of SMs.

    The render kernels of the next view run while the gradients of the current one are reduced, and they
    are issue-bound on every SM; an NCCL all-reduce kernel running beside them takes SMs away for the whole
    length of the collective (measured: 10-12 % of the step at 2-8 ranks).  Here the gradient bucket lives in
    symmetric (peer-mapped) memory and the collective is

        barrier -> reduce-scatter: each rank pulls "its" 1/W slice of every peer's bucket with peer DMA copies
                -> one small kernel sums the W slices
        barrier -> all-gather: each rank pushes its reduced slice into every peer's bucket with peer DMA copies
        barrier

    i.e. 2 (W-1)/W of the bucket crosses NVLink per rank, exactly like a ring, but the data movement runs on
    the copy engines; the only kernels are the W-way slice sum and the barriers' flag kernels.  Everything is
    enqueued on a side stream; `wait()` makes the caller's stream wait for the result.

    The rasteriser's backward writes its five parameter gradients straight into the bucket
    (see `bucket_provider`); two buckets alternate so that step i+1 can produce gradients while step i's are
    still being reduced.
    """

    def __init__(self, numel: int, device, group=None, n_buckets: int = 2):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.device = torch.device(device)
        align = 64 * self.world
        self.numel = (int(numel) + align - 1) // align * align
        self.slice = self.numel // self.world
        self.buckets, self.handles, self.peers = [], [], []
        for _ in range(n_buckets):
            b = symm_mem.empty(self.numel, dtype=torch.float32, device=self.device)
            h = symm_mem.rendezvous(b, group=self.group.group_name)
            self.buckets.append(b)
            self.handles.append(h)
            self.peers.append([h.get_buffer(p, (self.numel,), torch.float32) for p in range(self.world)])
        self.tmp = torch.empty(self.world, self.slice, dtype=torch.float32, device=self.device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.ready = torch.cuda.Event()
        self.done = [torch.cuda.Event() for _ in range(n_buckets)]
        self._next = 0
        self._inflight = []

    # -- bucket hand-out: the rasteriser asks for `n` floats for this frame's parameter gradients
    def bucket_provider(self, n: int, device):
        if n > self.numel or torch.device(device) != self.device:
            return None
        i = self._next
        self._next = (i + 1) % len(self.buckets)
        # the bucket may still be the target of peer copies of the collective launched two steps ago
        torch.cuda.current_stream(self.device).wait_event(self.done[i])
        return self.buckets[i][:n]

    def index_of(self, t: torch.Tensor):
        for i, b in enumerate(self.buckets):
            if t.untyped_storage().data_ptr() == b.untyped_storage().data_ptr():
                return i
        return None

    def launch(self, tensors: Iterable[torch.Tensor]):
        bucket = grad_bucket(list(tensors))
        i = self.index_of(bucket) if bucket is not None else None
        if i is None:
            raise RuntimeError("PeerGradAllReducer: gradients do not live in one of its buckets "
                               "(install bucket_provider before the backward pass)")
        cur = torch.cuda.current_stream(self.device)
        self.ready.record(cur)
        r, W, sl = self.rank, self.world, self.slice
        mine = slice(r * sl, (r + 1) * sl)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self.ready)
            h, peers, b = self.handles[i], self.peers[i], self.buckets[i]
            h.barrier(channel=0)                       # every rank's gradients are complete
            for j in range(W):                         # reduce-scatter by peer DMA (start with my own slice)
                p = (r + j) % W
                self.tmp[p].copy_(peers[p][mine], non_blocking=True)
            torch.sum(self.tmp, dim=0, out=b[mine])
            h.barrier(channel=1)                       # nobody still reads the slice I am about to overwrite remotely
            for j in range(1, W):                      # all-gather by peer DMA
                p = (r + j) % W
                peers[p][mine].copy_(b[mine], non_blocking=True)
            h.barrier(channel=2)                       # every slice has landed everywhere
            self.done[i].record(self.stream)
        self._inflight.append(i)

    def wait(self):
        cur = torch.cuda.current_stream(self.device)
        for i in self._inflight:
            cur.wait_event(self.done[i])
        self._inflight.clear()


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
