"""Golden vectors for the SH -> RGB stage from the REFERENCE's utils/sh_utils.py (eval_sh, RGB2SH), imported in the
build container: colours of a small surfel cloud seen from one camera, degrees 0..3.  -> tests/golden/sh_colors.npz"""
import importlib.util, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from partgs_b200 import synth
spec = importlib.util.spec_from_file_location("ref_sh_utils", "/root/reference/utils/sh_utils.py")
sh = importlib.util.module_from_spec(spec); spec.loader.exec_module(sh)
cfg, scene, cams = synth.make_config("C1", device="cpu", P=3000, views=1)
cam = cams[0]
out = dict(P=3000)
dirs = scene["means3D"] - cam.campos[None]
dirs = dirs / dirs.norm(dim=1, keepdim=True)
for deg in range(4):
    # renderer/gaussian_renderer/__init__.py:76-81 (convert_SHs_python branch): eval_sh on [P,3,(deg+1)^2]
    shs_view = scene["shs"].transpose(1, 2)                 # [P,3,16]
    rgb = torch.clamp_min(sh.eval_sh(deg, shs_view, dirs) + 0.5, 0.0)
    out[f"rgb_deg{deg}"] = rgb.numpy()
np.savez_compressed(ROOT / "tests" / "golden" / "sh_colors.npz", **out)
print("wrote", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})
