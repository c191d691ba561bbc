"""The densification restatement (oracle/densify_oracle.py) against golden vectors produced by the UNMODIFIED
reference TwoGaussianModel on CPU (tools/make_golden_densify.py).  CPU only."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import densify_oracle

GOLDEN = sorted((Path(__file__).parent / "golden").glob("densify_*.npz"))
NAMES = densify_oracle.PARAM_NAMES


def load(path):
    z = np.load(path)
    t = {k: torch.from_numpy(z[k]) for k in z.files if z[k].ndim > 0}
    s = {k: float(z[k]) for k in z.files if z[k].ndim == 0}
    s["max_screen_size"] = None if s["max_screen_size"] < 0 else s["max_screen_size"]
    return t, s


@pytest.mark.parametrize("path", GOLDEN, ids=[p.stem for p in GOLDEN])
def test_oracle_reproduces_reference_densify_and_prune(path):
    t, s = load(path)
    params = {k: t["in_" + k] for k in NAMES}
    moments = {k: (t["in_m_" + k], t["in_v_" + k]) for k in NAMES}
    cur, mom, sem, info = densify_oracle.densify_and_prune(
        params, moments, t["in_semantic"], t["in_accum"].clone(), t["in_denom"], s["max_grad"], s["min_opacity"],
        s["extent"], s["max_screen_size"], s["percent_dense"], t["z"])
    assert info["n_split_selected"] * 2 == t["z"].shape[0]
    for k in NAMES:
        # bit-exact: same ATen ops in the same order on the same draws
        assert torch.equal(cur[k], t["out_" + k]), k
        assert torch.equal(mom[k][0], t["out_m_" + k]), k
        assert torch.equal(mom[k][1], t["out_v_" + k]), k
    assert torch.equal(sem, t["out_semantic"])
    n = info["n_out"]
    assert t["out_accum"].shape == (n, 1) and not t["out_accum"].any()
    assert t["out_denom"].shape == (n, 1) and not t["out_denom"].any()
    assert t["out_max_radii2D"].shape == (n,) and not t["out_max_radii2D"].any()


def test_fixtures_exercise_every_branch():
    assert len(GOLDEN) >= 2
    seen_clone = seen_split = seen_prune = False
    for path in GOLDEN:
        t, s = load(path)
        clone, split = densify_oracle.split_selection(t["in_accum"].clone(), t["in_denom"], t["in_scaling"],
                                                      s["max_grad"], s["extent"], s["percent_dense"])
        seen_clone |= bool(clone.any())
        seen_split |= bool(split.any())
        assert not (clone & split).any()
        P = t["in_xyz"].shape[0]
        seen_prune |= t["out_xyz"].shape[0] < P + int(clone.sum()) + int(split.sum())
        assert (t["in_denom"] == 0).any()  # never-visible surfels: 0/0 -> NaN -> 0
    assert seen_clone and seen_split and seen_prune
