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
    """Asynchronous gradient all-reduce (SUM) through torch.distributed (NCCL on GPUs, gloo in the CPU tests): one
    collective if the gradients share one allocation (the rasteriser's bucket), one per tensor otherwise."""

    def __init__(self, group=None, average: bool = False):
        self.group = group
        self.average = average
        self._pending = []

    @property
    def world(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def begin_batch(self):
        pass

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


class BucketBatch:
    """Ownership of the rasteriser's gradient bucket over a data-parallel BATCH (a rank renders its
    ceil(views / G) views, the gradients accumulate, ONE all-reduce per batch).

    `begin_batch()` opens a batch on the next of `n_buckets` buffers (waiting until the collective that last used it
    has finished).  Every backward pass of the batch then receives that same buffer from `bucket_provider`: the
    first one overwrites it, the following ones ADD to it inside the backward-preprocess kernel (no separate
    accumulation pass).  Contract: the caller drops `.grad` of the parameters before every backward (`p.grad =
    None`) — autograd then adopts the bucket views as `.grad` each time; with a live `.grad` it would add the bucket
    to itself.  `launch()` reduces the batch's buffer and closes the batch; a backward without an open batch opens
    one implicitly (one view per batch, the per-view mode)."""

    def __init__(self, numel: int, device, n_buckets: int = 2, n_lanes: int = 1):
        self.device = torch.device(device)
        self.numel = int(numel)
        self.n_lanes = max(int(n_lanes), 1)
        self._alloc(n_buckets)
        # Lanes: the views of a batch may be rendered on several CUDA streams at once (view k+1's latency-bound
        # binning chain then hides behind view k's render kernels).  Two backward passes must not read-modify-write
        # one buffer concurrently, so every lane beyond the first accumulates in a private (local) buffer that
        # `launch()` adds to the batch's bucket before the collective.  `set_lane()` selects the lane of the next
        # backward passes.
        self.side = [[torch.empty(self.numel, dtype=torch.float32, device=self.device) for _ in range(self.n_lanes - 1)]
                     for _ in range(len(self.buckets))]
        self._lane = 0
        self._cur = len(self.buckets) - 1
        self._open = False
        self._fresh_lane = [True] * self.n_lanes
        self.done = [None] * len(self.buckets)   # CUDA event per buffer: its last collective has finished

    @property
    def _fresh(self):      # no backward pass has written anything for this batch yet
        return all(self._fresh_lane)

    @_fresh.setter
    def _fresh(self, v):
        self._fresh_lane = [bool(v)] * self.n_lanes

    def _alloc(self, n_buckets):
        self.buckets = [torch.empty(self.numel, dtype=torch.float32, device=self.device) for _ in range(n_buckets)]

    def set_lane(self, lane: int):
        self._lane = int(lane) % self.n_lanes

    def begin_batch(self, streams=None):
        """`streams`: the lanes' streams — all of them must wait until the buffer's previous collective has finished
        (default: the current stream)."""
        self._cur = (self._cur + 1) % len(self.buckets)
        if self.done[self._cur] is not None and self.device.type == "cuda":
            for s in (streams or [torch.cuda.current_stream(self.device)]):
                s.wait_event(self.done[self._cur])
        self._open, self._fresh = True, True

    def bucket_provider(self, n: int, device):
        if n > self.numel or torch.device(device) != self.device:
            return None
        if not self._open:
            self.begin_batch()
        lane = self._lane
        accumulate, self._fresh_lane[lane] = not self._fresh_lane[lane], False
        buf = self.buckets[self._cur] if lane == 0 else self.side[self._cur][lane - 1]
        return buf[:n], accumulate

    def current(self) -> torch.Tensor:
        return self.buckets[self._cur]

    def _merge_lanes(self):
        """Fold the private buffers of lanes 1.. into the batch's bucket (call on the stream that runs the collective,
        after it waits for every lane)."""
        main = self.buckets[self._cur]
        for lane in range(1, self.n_lanes):
            if self._fresh_lane[lane]:
                continue
            if self._fresh_lane[0]:
                main.copy_(self.side[self._cur][lane - 1])
                self._fresh_lane[0] = False
            else:
                main.add_(self.side[self._cur][lane - 1])

    def _check(self, tensors):
        """The gradients handed to launch() must be views of the open batch's buffer."""
        tensors = [t for t in (tensors or []) if t is not None]
        if tensors:
            bucket = grad_bucket(tensors)
            mine = [self.current()] + self.side[self._cur]
            if bucket is None or all(bucket.untyped_storage().data_ptr() != b.untyped_storage().data_ptr() for b in mine):
                raise RuntimeError("gradients do not live in the bucket of the open batch (install bucket_provider "
                                   "before the backward pass and drop .grad before every backward)")
        if self._fresh:
            raise RuntimeError("no backward pass has written the bucket of this batch")


class NcclBucketAllReducer(BucketBatch):
    """BucketBatch reduced with one NCCL all-reduce per batch (the comparison point of the peer-memory collective)."""

    def __init__(self, numel: int, device, group=None, n_buckets: int = 2, n_lanes: int = 1):
        super().__init__(numel, device, n_buckets, n_lanes)
        self.group = group
        self._pending = []

    def launch(self, tensors: Iterable[torch.Tensor] | None = None, streams=None):
        self._check(tensors)
        if self.device.type == "cuda":
            cur = torch.cuda.current_stream(self.device)
            for s_ in (streams or []):
                cur.wait_stream(s_)
        self._merge_lanes()
        i = self._cur
        h = dist.all_reduce(self.buckets[i], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._pending.append((h, i))
        self._open = False

    def wait(self):
        for h, i in self._pending:
            h.wait()          # the caller's stream waits for the collective
            if self.device.type == "cuda":
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(self.device))
                self.done[i] = ev
        self._pending.clear()


class PeerGradAllReducer(BucketBatch):
    """Gradient all-reduce (SUM) over NVLink peer memory as ONE kernel per batch (csrc/collective.cu).

    The bucket lives in symmetric (peer-mapped) memory — the backward-preprocess kernel writes / accumulates the
    gradients straight into it.  The collective is

        barrier                      every rank's gradients are complete
        pgs_peer_allreduce_slice     rank r loads slice r of all W buckets over NVLink, adds them in rank order and
                                     stores the sum into slice r of all W buckets (reduce-scatter + all-gather fused)
        barrier                      every slice has landed everywhere

    on a side stream; `wait()` makes the caller's stream wait for the result.  2 (W-1)/W of the bucket crosses NVLink
    per rank, like a ring all-reduce, but in one pass, without staging copies, with a bounded number of CTAs
    (`max_ctas`: the render kernels of the next views run beside it and want the SMs), and with a fixed summation
    order: every rank ends with bitwise the same sums, equal to ((g0 + g1) + g2) + ... .
    `n_buckets` buffers alternate so that the next batch can accumulate while this one is being reduced."""

    def __init__(self, numel: int, device, group=None, n_buckets: int = 2, max_ctas: int = 32, n_lanes: int = 1):
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world > 8:
            raise RuntimeError("PeerGradAllReducer supports up to 8 ranks (one NVSwitch domain)")
        align = 64 * self.world
        super().__init__((int(numel) + align - 1) // align * align, device, n_buckets, n_lanes)
        self.slice = self.numel // self.world
        self.max_ctas = int(max_ctas)
        self.stream = torch.cuda.Stream(device=self.device)
        self.ready = torch.cuda.Event()
        self._inflight = []

    def _alloc(self, n_buckets):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        self.buckets, self.handles, self.peers, self.peer_ptrs = [], [], [], []
        for _ in range(n_buckets):
            b = symm_mem.empty(self.numel, dtype=torch.float32, device=self.device)
            h = symm_mem.rendezvous(b, group=self.group.group_name)
            self.buckets.append(b)
            self.handles.append(h)
            peers = [h.get_buffer(p, (self.numel,), torch.float32) for p in range(self.world)]
            self.peers.append(peers)
            self.peer_ptrs.append((C.c_void_p * self.world)(*[t.data_ptr() for t in peers]))

    def launch(self, tensors: Iterable[torch.Tensor] | None = None, streams=None):
        """`streams`: the lanes' streams whose backward passes fed this batch (default: the current stream)."""
        from . import _lib
        self._check(tensors)
        i = self._cur
        for s_ in (streams or [torch.cuda.current_stream(self.device)]):
            self.stream.wait_stream(s_)
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            self._merge_lanes()
            h = self.handles[i]
            h.barrier(channel=0)
            rc = _lib.load().pgs_peer_allreduce_slice(self.world, self.peer_ptrs[i], self.rank * self.slice, self.slice,
                                                      self.max_ctas, _lib.current_stream(self.device))
            _lib.check(rc, "pgs_peer_allreduce_slice")
            h.barrier(channel=1)
            ev = torch.cuda.Event()
            ev.record(self.stream)
            self.done[i] = ev
        self._inflight.append(i)
        self._open = False

    def wait(self):
        cur = torch.cuda.current_stream(self.device)
        for i in self._inflight:
            cur.wait_event(self.done[i])
        self._inflight.clear()


def sharded_step(render_loss: Callable[[int], torch.Tensor], params: Dict[str, torch.Tensor], n_views: int,
                 reducer=None, rank: int | None = None, world: int | None = None,
                 order: Sequence[str] = ("means3D", "shs", "opacities", "scales", "rotations"), streams=None):
    """One data-parallel step over ``n_views`` cameras (train.py:93-99,219-225 per rank).

    ``render_loss(view_index)`` must render that view from ``params`` and return a scalar loss.  Every rank
    back-propagates its own views, the gradients accumulate over them, and the accumulated gradients are summed
    over ranks with ONE collective.  With a bucket reducer (`PeerGradAllReducer`, `NcclBucketAllReducer`, provider
    installed with `diff_surfel_rasterization.set_grad_bucket_provider`) the accumulation happens inside the
    backward kernel, in the buffer the collective reduces; otherwise autograd accumulates into ``.grad``.
    ``streams`` (CUDA streams, at most ``reducer.n_lanes``): consecutive views run on alternating streams — the
    parameters are fixed over a batch, so view k+1's preprocess + binning chain overlaps view k's render kernels; each
    stream accumulates in its own lane of the bucket reducer.
    Returns (local_loss_sum, my_view_indices)."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    reducer = reducer or GradAllReducer()
    bucketed = isinstance(reducer, BucketBatch)
    for p in params.values():
        p.grad = None
    mine = shard_views(n_views, rank, world)
    lanes = list(streams or [])
    if lanes and not (bucketed and len(lanes) <= reducer.n_lanes):
        raise RuntimeError("sharded_step(streams=...) needs a bucket reducer with at least as many lanes")
    if lanes:
        cur = torch.cuda.current_stream(lanes[0].device)
        for s_ in lanes:
            s_.wait_stream(cur)
        reducer.begin_batch(streams=lanes)
    else:
        reducer.begin_batch()
    total = None
    for j, v in enumerate(mine):
        if bucketed:                      # the kernel accumulates in the bucket: autograd must adopt, not add
            for p in params.values():
                p.grad = None
            reducer.set_lane(j % len(lanes) if lanes else 0)
        if lanes:
            with torch.cuda.stream(lanes[j % len(lanes)]):
                loss = render_loss(v)
                loss.backward()
        else:
            loss = render_loss(v)
            loss.backward()
        total = loss.detach() if total is None else total + loss.detach()
    if lanes:
        for s_ in lanes:                  # the caller's stream sees every lane's work (losses, gradients)
            cur.wait_stream(s_)
    present = [k for k in order if params.get(k) is not None]
    if n_views < world and world > 1 and present:
        # Some rank has no view.  It still has to join the collective, and with the SAME layout as its peers: the
        # ranks that rendered say whether their gradients came out as one bucket (the rasteriser's) or as separate
        # tensors (any other model); an idle rank then builds zero gradients of that layout — one bucket carved like
        # a backward pass would, or one tensor per parameter.  (Five small collectives against one large one would
        # hang NCCL or reduce garbage.)
        ref = params[present[0]]
        bucketed_here = 1 if not mine else int(grad_bucket([params[k].grad for k in present]) is not None)
        flag = torch.tensor([bucketed_here], dtype=torch.int32, device=ref.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=getattr(reducer, "group", None))
        point_level = tuple(order) == ("means3D", "shs", "opacities", "scales", "rotations")
        if not mine and int(flag.item()) == 1 and set(present) == set(order) and point_level:
            from . import diff_surfel_rasterization as dsr
            P, M = params["means3D"].shape[0], params["shs"].shape[1]
            for k, z in zip(order, dsr.zero_bucket_grads(P, M, ref.device)):
                params[k].grad = z
    grads = []
    for k in present:
        p = params[k]
        if p.grad is None:
            p.grad = torch.zeros_like(p)
        grads.append(p.grad)
    if bucketed and mine:
        # with lanes the parameters' .grad views belong to the last view's lane: after the collective the reduced
        # gradients live in the batch's bucket, so .grad is re-pointed at it (same carving as the backward pass's)
        reducer.launch(None if lanes else grads)
        reducer.wait()
        if lanes:
            from . import diff_surfel_rasterization as dsr
            flat = reducer.current()
            off = 0
            for k in present:
                p = params[k]
                n = p.numel()
                p.grad = flat[off:off + n].view(p.shape)
                off += (n + 63) // 64 * 64
    else:
        reducer.launch(grads)
        reducer.wait()
    return total, mine
