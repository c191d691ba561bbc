"""Statistics of the forward pass's per-warp blend masks (how many of a warp's 32 pixels blend a staged surfel).
usage: python tools/mask_stats.py [--cfg C3] [--view 0]"""
import argparse, json, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from partgs_b200 import synth, debug  # noqa: E402
import parity_utils as pu  # noqa: E402

ap = argparse.ArgumentParser(); ap.add_argument("--cfg", default="C3"); ap.add_argument("--view", type=int, default=0)
a = ap.parse_args()
cfg, scene, cams = synth.make_config(a.cfg, device="cuda", views=a.view + 1)
cam = cams[a.view]; bg = torch.zeros(3, device="cuda")
o = pu.run_ours_raw(scene, cam, bg)
R = o["num_rendered"]; P = cfg["P"]
st = debug.parse_state(o["geom"], o["img"], o["binning"], P, cfg["W"], cfg["H"], R)
ranges = st["ranges"].long(); L = ranges[:, 1] - ranges[:, 0]
ncontrib = st["n_contrib"][0]
m = st["frag_mask"]
# only positions a warp actually visited hold defined values: restrict to pos < max n_contrib of the tile is not exact
# (forward may stop earlier/later), so count via bits only on entries below each tile's walked depth = all entries
# written; unwritten entries are beyond every pixel's last contributor and are never read.  Use n_contrib per warp.
pc = torch.zeros_like(m)
x = m.clone()
for _ in range(32):
    pc += (x & 1); x = (x >> 1) & 0x7fffffff
# mask validity: build per-warp walked depth (max last contributor over the warp's pixels)
gx, gy = (cfg["W"] + 15) // 16, (cfg["H"] + 15) // 16
nc = torch.zeros(gy * 16, gx * 16, dtype=torch.int32, device="cuda"); nc[:cfg["H"], :cfg["W"]] = ncontrib
top = nc.reshape(gy, 4, 4, gx, 2, 8).permute(0, 3, 1, 4, 2, 5).reshape(gy * gx, 8, 32).amax(-1)  # [tile][warp]
tile_of = torch.repeat_interleave(torch.arange(gy * gx, device="cuda"), L)
pos = torch.arange(R, device="cuda") - ranges[tile_of, 0]
valid = pos[None, :] < top[tile_of].t()  # [8][R]
pcv = torch.where(valid, pc, torch.zeros_like(pc))
nz = (pcv > 0)
hist = torch.bincount(pcv[nz].flatten(), minlength=33).tolist()
out = dict(cfg=a.cfg, R=int(R), warp_instances_walked_bwd=int(valid.sum()), staged_bwd=int(nz.sum()),
           blended_pixel_fragments=int(pcv.sum()), mean_pixels_per_staged=float(pcv.sum() / nz.sum()),
           popc_hist=hist, blended_per_pixel=float(pcv.sum() / (cfg["W"] * cfg["H"])))
print(json.dumps(out))
