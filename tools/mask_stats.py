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
# what finer work items would cost: per (warp, 32-candidate step) iterations = max over lane groups of the
# number of candidates with a non-zero mask restricted to that group (rows of 8 lanes / halves of 16 lanes)
mv = torch.where(valid, m, torch.zeros_like(m))
step = (pos // 32)[None, :].expand(8, -1)
key = (tile_of[None, :].expand(8, -1) * 8 + torch.arange(8, device="cuda")[:, None]) * 4096 + step  # unique per (tile, warp, step)
uk, inv = torch.unique(key.flatten(), return_inverse=True)
def per_step(sel):
    c = torch.zeros(uk.numel(), dtype=torch.int64, device="cuda")
    c.scatter_add_(0, inv, sel.flatten().long())
    return c
n_full = per_step(mv != 0)
rows = [per_step(((mv >> (8 * r)) & 0xFF) != 0) for r in range(4)]
halves = [per_step(((mv >> (16 * h)) & 0xFFFF) != 0) for h in range(2)]
iters_rows = torch.stack(rows).amax(0); iters_halves = torch.stack(halves).amax(0)
extra = dict(steps=int((n_full >= 0).sum()), steps_nonempty=int((n_full > 0).sum()), iters_warp=int(n_full.sum()),
             iters_halves=int(iters_halves.sum()), iters_rows=int(iters_rows.sum()),
             row_items=int(sum(r.sum() for r in rows)), half_items=int(sum(h.sum() for h in halves)))
out = dict(cfg=a.cfg, **extra, R=int(R), warp_instances_walked_bwd=int(valid.sum()), staged_bwd=int(nz.sum()),
           blended_pixel_fragments=int(pcv.sum()), mean_pixels_per_staged=float(pcv.sum() / nz.sum()),
           popc_hist=hist, blended_per_pixel=float(pcv.sum() / (cfg["W"] * cfg["H"])))
print(json.dumps(out))
# critical path: staged fragments per (tile, warp) and per tile
per_warp = torch.zeros(gy * gx * 8, dtype=torch.int64, device="cuda")
per_warp.scatter_add_(0, (tile_of[None, :].expand(8, -1) * 8 + torch.arange(8, device="cuda")[:, None]).flatten(), (mv != 0).flatten().long())
pw = per_warp.reshape(-1, 8)
srt = torch.sort(per_warp, descending=True).values
print(json.dumps(dict(per_warp_top=srt[:16].tolist(), per_warp_mean=float(per_warp.float().mean()),
                      per_tile_top=torch.sort(pw.sum(1), descending=True).values[:16].tolist(),
                      per_tile_mean=float(pw.sum(1).float().mean()),
                      warps_over_1000=int((per_warp > 1000).sum()), warps_over_500=int((per_warp > 500).sum()),
                      frac_work_in_warps_over_500=float(per_warp[per_warp > 500].sum() / per_warp.sum()))))
