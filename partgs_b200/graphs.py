"""CUDA-graph capture of a rasteriser iteration (the boundary must be graph-capturable).

A forward call issued while its stream is being captured never waits for the frame's instance count (a capture cannot
contain the host wait; include/partgs_b200.h, PGS_FWD_LAZY_COUNT): binning and render are captured for the instance
capacity remembered from earlier frames, the kernels read the actual count on the device, and the count lands in a
pinned word that every replay refreshes.  `GraphedIteration` captures `fn()` (forward + loss + backward on static input
tensors), replays it, and checks after a replay that the frame fitted its capacity."""
from __future__ import annotations

import torch

from .diff_surfel_rasterization import resolve_count


class GraphedIteration:
    """``fn`` must read its inputs from tensors that stay alive and in place (copy new camera matrices / images INTO
    them before `replay()`), and must have run eagerly at least once (allocator warm-up, instance capacity).

        it = GraphedIteration(step_fn, warmup=2)      # runs step_fn eagerly, then captures it
        cam_buf.copy_(next_cam); it.replay(); loss = it.outputs
    """

    def __init__(self, fn, warmup: int = 2, check_every: int = 1):
        self.fn = fn
        self.check_every = max(int(check_every), 1)
        self._replays = 0
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                self.outputs = fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.outputs = fn()

    def replay(self, check: bool | None = None):
        self.graph.replay()
        self._replays += 1
        if check if check is not None else (self._replays % self.check_every == 0):
            self.verify()
        return self.outputs

    def verify(self):
        """Synchronises, then raises if the replayed frame needed more instances than the captured arena holds (the
        captured kernels then did nothing): re-render eagerly (that raises the remembered capacity) and re-capture."""
        torch.cuda.current_stream().synchronize()
        n, overflow = resolve_count()
        if overflow:
            raise RuntimeError(f"captured frame needs {n} instances, more than the captured arena holds: run the "
                               "iteration eagerly once and capture again")
        return n
