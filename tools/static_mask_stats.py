"""How selective would a STATIC per-instance footprint mask be (box of the surfel vs the eight 8x4 footprints of its tile,
computed once at emission)?  Compared with what the forward kernel stages today (dynamic box + conic test per warp)."""
import sys, json
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import parity_utils as pu
from partgs_b200 import synth, debug
cfg, scene, cams = synth.make_config("C3", device="cuda", views=1)
W, H, P = cfg["W"], cfg["H"], cfg["P"]
o = pu.run_ours_raw(scene, cams[0], torch.zeros(3, device="cuda"))
R = o["num_rendered"]
st = debug.parse_state(o["geom"], o["img"], o["binning"], P, W, H, R)
gx = (W + 15) // 16
ranges = st["ranges"].long()
L = ranges[:, 1] - ranges[:, 0]
tile_of = torch.repeat_interleave(torch.arange(ranges.shape[0], device="cuda"), L)
pl = st["point_list"].long() & 0xFFFFFF
box = st["bbox"][pl]                                  # [R,4] x0,y0,x1,y1
tx, ty = (tile_of % gx) * 16, (tile_of // gx) * 16
hits = torch.zeros(R, dtype=torch.int64, device="cuda")
for w in range(8):
    fx0 = (tx + (w & 1) * 8).float(); fy0 = (ty + (w >> 1) * 4).float()
    fx1, fy1 = fx0 + 7, fy0 + 3
    hit = ~((box[:, 0] > fx1) | (box[:, 2] < fx0) | (box[:, 1] > fy1) | (box[:, 3] < fy0))
    hits += hit.long()
staged = (st["frag_mask"] != 0).sum().item()
print(json.dumps(dict(R=R, warp_candidate_pairs=8 * R, static_box_hits=int(hits.sum()), static_box_rate=float(hits.sum()) / (8 * R),
                      instances_with_no_footprint=int((hits == 0).sum()), blended_pairs=int(staged),
                      blended_rate=staged / (8 * R))))
