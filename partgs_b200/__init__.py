"""partgs_b200 — B200-native (sm_100a) implementation of PartGS's rendering hot path.

Sub-packages mirror the reference's extension modules:
  partgs_b200.diff_surfel_rasterization        <- submodules/diff-surfel-rasterization
  partgs_b200.diff_surfel_rasterization_part   <- submodules/diff-surfel-rasterization_part
  partgs_b200.simple_knn                       <- submodules/simple-knn
  partgs_b200.superquadric                     <- games/block_mesh_splatting (parameterisation)

`partgs_b200/dropin/` holds top-level packages with the reference's import names
(``diff_surfel_rasterization``, ``diff_surfel_rasterization_part``, ``simple_knn``);
put that directory on ``sys.path`` (or call :func:`install_dropin`) and the unmodified
PartGS renderers / train.py pick up this implementation.
"""
from pathlib import Path
import sys

__version__ = "0.1.0"

DROPIN_DIR = Path(__file__).resolve().parent / "dropin"


def install_dropin() -> str:
    """Make ``import diff_surfel_rasterization`` etc. resolve to this package."""
    p = str(DROPIN_DIR)
    if p not in sys.path:
        sys.path.insert(0, p)
    return p
