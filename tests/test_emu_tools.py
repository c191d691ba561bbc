"""The timing tool that bench.py runs as a subprocess for its `next_rows` (tools/time_rank34.py) — executed here at toy
sizes on the CPU emulator with every "cuda" device string redirected to the CPU, so that its own Python (argument
order, shapes, keys of the JSON rows) is known to work before its first run on a GPU.  Times measured here mean
nothing; the rows' structure and the parity flags do."""
import importlib.util
import time
from pathlib import Path

import pytest
import torch

from emu_host import emulated_host  # noqa: F401  (fixture)

ROOT = Path(__file__).resolve().parent.parent


class _Event:
    def __init__(self, *a, **k):
        self.t = 0.0

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3

    def synchronize(self):
        pass


def _is_cuda(x):
    return (isinstance(x, str) and x.startswith("cuda")) or (isinstance(x, torch.device) and x.type == "cuda")


@pytest.fixture()
def tool(emulated_host, monkeypatch):
    for fname in ("zeros", "ones", "empty", "tensor", "rand", "randn", "full", "zeros_like", "ones_like", "randn_like",
                  "rand_like", "arange", "randint"):
        orig = getattr(torch, fname)
        monkeypatch.setattr(torch, fname, (lambda o: lambda *a, **k: o(*a, **{kk: ("cpu" if kk == "device" and _is_cuda(v) else v)
                                                                              for kk, v in k.items()}))(orig))
    _to = torch.Tensor.to
    monkeypatch.setattr(torch.Tensor, "to", lambda self, *a, **k: _to(
        self, *["cpu" if _is_cuda(x) else x for x in a], **{kk: ("cpu" if kk == "device" and _is_cuda(v) else v) for kk, v in k.items()}))
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    spec = importlib.util.spec_from_file_location("time_rank34", ROOT / "tools" / "time_rank34.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_rank34_rows_at_toy_sizes(tool, monkeypatch):
    r = tool.row_densify(3000, 1)
    assert r["row"] == "densify_and_prune" and r["n_out"] > 0 and r["ours_ms"] >= 0 and r["reference_ms"] >= 0
    r = tool.row_adam(2000, 1)
    assert r["row"] == "adam_step" and {"ours_ms", "reference_ms", "reference_foreach_ms", "speedup"} <= set(r)
    r = tool.row_extract(1, H=24, W=40)
    assert r["row"].startswith("extract_epilogue_40x24") and r["ours_ms"] >= 0
    r = tool.row_regularizers(1, H=24, W=40)
    assert r["row"].startswith("regularizers_fwd_bwd_40x24") and r["reference_ms"] >= 0
