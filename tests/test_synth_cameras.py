"""CPU: the synthetic camera generator (partgs_b200/synth.py) produces the matrices the reference's own camera code
(utils/graphics_utils.py:29-62, scene/cameras.py:60-63) produces for the same pose — golden vectors made by importing
that code (tools/make_golden_cam.py)."""
from pathlib import Path

import numpy as np
import torch

from partgs_b200 import synth


def test_cameras_match_reference_conventions():
    z = np.load(Path(__file__).parent / "golden" / "cameras.npz")
    n = 0
    for i, (W, H) in enumerate(((400, 300), (1600, 1200), (123, 77))):
        cams = synth.make_cameras(3, W, H, synth.SEED_BASE + i)
        for j, cam in enumerate(cams):
            assert list(z[f"{i}_{j}_size"]) == [W, H]
            assert torch.allclose(cam.world_view_transform, torch.from_numpy(z[f"{i}_{j}_wvt"]), rtol=0, atol=2e-6)
            assert torch.allclose(cam.full_proj_transform, torch.from_numpy(z[f"{i}_{j}_full"]), rtol=1e-5, atol=1e-5)
            assert torch.allclose(cam.camera_center, torch.from_numpy(z[f"{i}_{j}_center"]), rtol=0, atol=1e-5)
            n += 1
    assert n == 9
