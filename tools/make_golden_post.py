"""Golden vectors for the surface-map kernels from the REFERENCE code itself (runs in the build container only,
where /root/reference is mounted): utils/point_utils.py depth_to_normal / depths_to_points are imported unmodified;
their hard-coded `.cuda()` / device='cuda' are neutralised by patching torch (CPU execution).
-> tests/golden/surface_maps_<W>x<H>.npz (inputs, outputs, gradient of a fixed functional)."""
import importlib.util, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from partgs_b200 import synth

torch.Tensor.cuda = lambda self, *a, **k: self
_arange = torch.arange
torch.arange = lambda *a, **k: _arange(*a, **{kk: v for kk, v in k.items() if kk != "device"})
spec = importlib.util.spec_from_file_location("ref_point_utils", "/root/reference/utils/point_utils.py")
ref = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref)

for (W, H, seed) in ((37, 23, 1), (64, 48, 2)):
    gen = torch.Generator().manual_seed(seed)
    cam = synth.make_cameras(1, W, H, synth.SEED_BASE + seed)[0]
    depth = (2.0 + 0.5 * torch.rand(1, H, W, generator=gen)).requires_grad_(True)
    depth.data[0, H // 2, W // 3] = 0.0          # a hole
    normal = ref.depth_to_normal(cam, depth)       # [H,W,3]
    points = ref.depths_to_points(cam, depth)      # [H*W,3]
    g = torch.randn(H, W, 3, generator=gen)
    (normal * g).sum().backward()
    np.savez_compressed(ROOT / "tests" / "golden" / f"surface_maps_{W}x{H}.npz", W=W, H=H, seed=seed,
                        viewmatrix=cam.viewmatrix.numpy(), projmatrix=cam.projmatrix.numpy(),
                        depth=depth.detach().numpy(), normal=normal.detach().numpy(), points=points.detach().numpy(),
                        g=g.numpy(), d_depth=depth.grad.numpy())
    print("wrote", W, H, float(normal.abs().max()))
