"""CPU: the C oracle's SH -> RGB stage (oracle/surfel_oracle.c, restating forward.cu:20-70) against colours computed
by the reference's own utils/sh_utils.eval_sh (tools/make_golden_sh.py), degrees 0..3."""
from pathlib import Path

import numpy as np
import pytest
import torch

from partgs_b200 import synth


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_oracle_sh_colors_match_reference_eval_sh(deg):
    from oracle import cpu_oracle
    z = np.load(Path(__file__).parent / "golden" / "sh_colors.npz")
    cfg, scene, cams = synth.make_config("C1", device="cpu", P=int(z["P"]), views=1)
    f = cpu_oracle.forward_scene(scene, cams[0], sh_degree=deg, keep_state=True)
    vis = f["radii"] > 0                       # colours are only computed for visible surfels
    assert vis.sum() > 1000
    got, want = f["rgb"][vis], z[f"rgb_deg{deg}"][vis]
    assert np.abs(got - want).max() <= 2e-6
