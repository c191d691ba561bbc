"""Golden vectors for the synthetic camera generator from the REFERENCE's utils/graphics_utils.py
(getWorld2View2, getProjectionMatrix) and the matrix assembly of scene/cameras.py:60-63, imported / restated in
the build container.  -> tests/golden/cameras.npz"""
import importlib.util, math, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from partgs_b200 import synth
spec = importlib.util.spec_from_file_location("ref_graphics_utils", "/root/reference/utils/graphics_utils.py")
gu = importlib.util.module_from_spec(spec); spec.loader.exec_module(gu)
out = {}
for i, (W, H) in enumerate(((400, 300), (1600, 1200), (123, 77))):
    cams = synth.make_cameras(3, W, H, synth.SEED_BASE + i)
    for j, cam in enumerate(cams):
        # recover (R, T) in the reference's convention: world_view_transform = getWorld2View2(R, T)^T with R = W2C[:3,:3]^T
        W2C = cam.viewmatrix.t().double().numpy()
        R = W2C[:3, :3].T
        T = W2C[:3, 3]
        fovx, fovy = 2 * math.atan(cam.tanfovx), 2 * math.atan(cam.tanfovy)
        wvt = torch.tensor(gu.getWorld2View2(R, T)).transpose(0, 1)                      # scene/cameras.py:60
        proj = gu.getProjectionMatrix(znear=0.01, zfar=100.0, fovX=fovx, fovY=fovy).transpose(0, 1)   # :61
        full = (wvt.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0)                       # :62
        center = wvt.inverse()[3, :3]                                                     # :63
        out[f"{i}_{j}_wvt"], out[f"{i}_{j}_full"], out[f"{i}_{j}_center"] = wvt.numpy(), full.numpy(), center.numpy()
        out[f"{i}_{j}_size"] = np.array([W, H])
np.savez_compressed(ROOT / "tests" / "golden" / "cameras.npz", **out)
print("wrote", len(out) // 4, "cameras")
