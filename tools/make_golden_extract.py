"""Golden vectors for the extraction epilogue: the reference's GaussianExtractor.partmap_to_rgbmap and
estimate_bounding_sphere (utils/mesh_utils.py) run UNMODIFIED on CPU.  Build container only (/root/reference).
get_fancy_color needs seaborn / matplotlib (absent): the palette is an input here, injected by replacing the name
`get_fancy_color` in the reference module's namespace.

    python tools/make_golden_extract.py
"""
import sys
from pathlib import Path
from types import SimpleNamespace
from unittest.mock import MagicMock

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
for m in ['pytorch3d', 'pytorch3d.renderer', 'pytorch3d.structures', 'pytorch3d.ops', 'pytorch3d.io', 'open3d', 'seaborn', 'matplotlib', 'matplotlib.pyplot', 'matplotlib.colors', 'mediapy', 'imageio', 'skimage',
          'trimesh', 'plyfile', 'lpips', 'easydict']:
    try:
        __import__(m)
    except Exception:
        sys.modules.setdefault(m, MagicMock())
sys.path.insert(0, "/root/reference")
_orig_tensor = torch.tensor
torch.tensor = lambda *a, **kw: _orig_tensor(*a, **{**kw, "device": "cpu"} if str(kw.get("device", "")) == "cuda" else kw)
_orig_cuda = torch.Tensor.cuda
torch.Tensor.cuda = lambda self, *a, **kw: self
torch.cuda.empty_cache = lambda: None

import utils.mesh_utils as mu  # noqa: E402
from partgs_b200.synth import make_cameras  # noqa: E402


def main():
    g = torch.Generator().manual_seed(3)
    out = {}
    for name, S, H, W in (("a", 5, 23, 31), ("b", 16, 12, 40), ("c", 1, 9, 9)):
        part = torch.rand(S, H, W, generator=g) * 1.6 - 0.4           # values outside [0,1] exercise the clamp
        part[:, : H // 3] *= 0.02                                     # a background band (sum < 0.1)
        part[:, H // 2, ::3] = part[:1, H // 2, ::3]                  # ties: the first maximum must win
        palette = torch.rand(S + 1, 3, generator=g)
        mu.get_fancy_color = lambda n, palette=palette: palette[:n]
        ex = mu.GaussianExtractor(None, lambda *a, **k: None, None)
        rgb = ex.partmap_to_rgbmap(part)
        out[f"{name}_part"] = part.numpy()
        out[f"{name}_palette"] = palette.numpy()
        out[f"{name}_rgb"] = rgb.numpy()
    cams = make_cameras(7, 64, 48, seed=5, device="cpu")
    ex = mu.GaussianExtractor(None, lambda *a, **k: None, None)
    ex.viewpoint_stack = [SimpleNamespace(world_view_transform=c.world_view_transform) for c in cams]
    ex.estimate_bounding_sphere()
    out["wvt"] = np.stack([c.world_view_transform.numpy() for c in cams])
    out["center"] = ex.center.numpy()
    out["radius"] = np.float64(ex.radius)
    path = ROOT / "tests" / "golden" / "extract_maps.npz"
    np.savez_compressed(path, **out)
    print(path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
