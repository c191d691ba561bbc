"""CPU tests of host-side logic that needs no GPU: gradient-bucket detection, the allocation scope behind the
C-ABI callbacks, the binding table."""
import ctypes as C

import pytest
import torch


def test_grad_bucket_detects_shared_storage():
    from partgs_b200.dist import grad_bucket
    root = torch.zeros(200)
    a = root[0:12].view(4, 3).detach()          # autograd hands .grad over detached: ._base is None
    b = root[64:72].view(4, 2).detach()
    assert a._base is None
    flat = grad_bucket([a, None, b])
    assert flat is not None and flat.numel() == 72 and flat.data_ptr() == root.data_ptr()
    flat.fill_(2.0)
    assert float(a.sum()) == 24.0 and float(root[72]) == 0.0
    assert grad_bucket([a, torch.zeros(3)]) is None          # different allocations
    assert grad_bucket([a, root[100:104].double()]) is None  # dtype mismatch
    assert grad_bucket([]) is None


def test_alloc_scope_serves_callbacks_without_cycles():
    import gc
    import weakref
    from partgs_b200 import _lib
    sc = _lib.AllocScope("cpu")
    assert _lib.ALLOC_CB(16, sc.GEOM) is None               # no active scope: NULL, the C side fails cleanly
    with sc:
        p = _lib.ALLOC_CB(1000, sc.BINNING)
        q = _lib.ALLOC_CB(64, sc.GEOM)
    assert p == sc.tensor(sc.BINNING).data_ptr() and q == sc.tensor(sc.GEOM).data_ptr()
    assert sc.tensor(sc.BINNING).numel() == 1000 and sc.tensor(sc.IMAGE).numel() == 0
    ref = weakref.ref(sc)
    gc.disable()
    try:
        del sc
        assert ref() is None, "AllocScope must be freed by reference counting (no callback cycle)"
    finally:
        gc.enable()


def test_bucket_provider_hook_round_trip():
    from partgs_b200 import diff_surfel_rasterization as dsr
    seen = []

    def provider(n, dev):
        seen.append(n)
        return torch.zeros(n + 8)

    dsr.set_grad_bucket_provider(provider)
    try:
        views = dsr._carve_bucket("cpu", [(5, 3), (5, 16, 3), (5, 1), (5, 2), (5, 4)])
        assert [tuple(v.shape) for v in views] == [(5, 3), (5, 16, 3), (5, 1), (5, 2), (5, 4)]
        assert len(seen) == 1 and seen[0] >= 5 * (3 + 48 + 1 + 2 + 4)
        st = views[0].untyped_storage().data_ptr()
        assert all(v.untyped_storage().data_ptr() == st for v in views)
        assert all(v.data_ptr() % 256 == views[0].data_ptr() % 256 for v in views)  # each view 256-byte aligned
    finally:
        dsr.set_grad_bucket_provider(None)
    views = dsr._carve_bucket("cpu", [(2, 3)])
    assert views[0].shape == (2, 3)


def test_bench_next_rows_never_raises_and_reports_errors_as_rows():
    """bench.py's `next_rows` leg (tools/time_rank34.py in a child process): without a GPU every row fails inside the
    child — the parent still gets one compact row per request and no exception; a timeout is reported the same way."""
    import bench
    rows = bench.next_rows(rows="adam,extract", reps=1, timeout_s=120)
    assert len(rows) == 2 and all("row" in r and "trace" not in r for r in rows)
    if not torch.cuda.is_available():
        assert [r["row"] for r in rows] == ["row_adam", "row_extract"]
        assert all("error" in r and len(r["error"]) <= 160 for r in rows)
    rows = bench.next_rows(rows="adam", reps=1, timeout_s=0.01)
    assert len(rows) == 1 and "TimeoutExpired" in rows[0]["error"]


def test_bucket_batch_lanes_accumulate_separately_and_merge():
    """dist.BucketBatch: one buffer per batch slot, the first backward of a lane overwrites, the following ones
    accumulate; lanes beyond the first own private buffers that are folded into the bucket before the collective."""
    from partgs_b200.dist import BucketBatch
    bb = BucketBatch(8, "cpu", n_buckets=2, n_lanes=2)
    bb.begin_batch()
    first = bb.current()
    seq = []
    for j in range(5):                       # views alternate lanes 0, 1, 0, 1, 0
        bb.set_lane(j % 2)
        buf, acc = bb.bucket_provider(8, "cpu")
        seq.append(acc)
        if acc:
            buf += float(j + 1)              # what the backward kernel does in accumulate mode
        else:
            buf.fill_(float(j + 1))
    assert seq == [False, False, True, True, True]
    assert float(bb.current()[0]) == 1 + 3 + 5 and float(bb.side[bb._cur][0][0]) == 2 + 4
    bb._merge_lanes()
    assert torch.equal(bb.current(), torch.full((8,), 15.0))
    bb._open = False
    bb.begin_batch()                         # next batch: the other slot, fresh lanes
    assert bb.current().data_ptr() != first.data_ptr() and bb._fresh
    bb.set_lane(1)                           # a batch whose only view ran on lane 1
    buf, acc = bb.bucket_provider(8, "cpu")
    assert acc is False
    buf.fill_(7.0)
    bb._merge_lanes()
    assert torch.equal(bb.current(), torch.full((8,), 7.0))
    with pytest.raises(RuntimeError):
        bb._open = False
        bb.begin_batch()
        bb._check(None)                      # nothing written in this batch
