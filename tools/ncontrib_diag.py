"""Which n_contrib entries differ from the reference (diagnostic)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from partgs_b200 import debug, synth
from oracle import ref_cuda
import parity_utils as pu
cfg, scene, cams = synth.make_config("C2", device="cuda", views=1)
W, H, P = cfg["W"], cfg["H"], cfg["P"]
bg = torch.zeros(3, device="cuda")
ref = ref_cuda.forward(scene, cams[0], bg)
ours = pu.run_ours_raw(scene, cams[0], bg)
st = debug.parse_state(ours["geom"], ours["img"], ours["binning"], P, W, H, ours["num_rendered"])
ri = ref_cuda.parse_image(ref["img"], W * H)
a = st["n_contrib"].reshape(2, -1); b = ri["n_contrib"]
for ch in range(2):
    m = a[ch] != b[ch]
    print("channel", ch, "mismatch", int(m.sum()), "ours", a[ch][m][:10].tolist(), "ref", b[ch][m][:10].tolist())
bb = st["bbox"][ref["radii"] > 0]
print("frac all-boxes", float((bb[:, 2] > 1e30).float().mean()), "mean box w,h", float((bb[:,2]-bb[:,0]).clamp(max=1e4).mean()), float((bb[:,3]-bb[:,1]).clamp(max=1e4).mean()))
rng = st["ranges"]; ln = (rng[:,1]-rng[:,0])
print("tile list len mean/max", float(ln.float().mean()), int(ln.max()))
last = a[0].float(); print("last_contrib mean/max", float(last.mean()), float(last.max()))
