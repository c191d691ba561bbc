"""torchrun check (2-8 GPUs): the camera-sharded batch step of SURVEY 8(e) on real GPUs — dist.sharded_step with the
in-kernel gradient accumulation and ONE peer-memory all-reduce per batch (and with the NCCL bucket reducer) equals a
single process back-propagating all views; incl. a batch with fewer views than ranks (idle ranks join with the same
bucket layout)."""
import os, sys
from pathlib import Path
import torch, torch.distributed as dist
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import parity_utils as pu
from partgs_b200 import synth, diff_surfel_rasterization as dsr
from partgs_b200.diff_surfel_rasterization import GaussianRasterizer
from partgs_b200.dist import init_nccl, NcclBucketAllReducer, PeerGradAllReducer, sharded_step
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
init_nccl(local)
dev = torch.device("cuda", local)
P, W, H = 60_000, 400, 300
scene = synth.make_point_scene(P, seed=9, device=dev)
cams = synth.make_cameras(2 * world + 1, W, H, seed=10, device=dev)
ups = [synth.upstream_grads(W, H, 70 + v, device=dev) for v in range(len(cams))]
bg = torch.zeros(3, device=dev)
names = ("means3D", "shs", "opacities", "scales", "rotations")


def render_loss(params, v):
    m2d = torch.zeros_like(params["means3D"], requires_grad=True)
    color, _, allmap = GaussianRasterizer(pu.settings_from_cam(cams[v], bg))(
        means3D=params["means3D"], means2D=m2d, opacities=params["opacities"], shs=params["shs"], scales=params["scales"],
        rotations=params["rotations"])
    return (color * ups[v]["color"]).sum() + (allmap * ups[v]["allmap"]).sum()


ok = True
for label, make in (("peer", lambda: PeerGradAllReducer(dsr.bucket_numel(P, 16), dev, n_lanes=2)),
                    ("nccl-bucket", lambda: NcclBucketAllReducer(dsr.bucket_numel(P, 16), dev, n_lanes=2))):
    red = make()
    dsr.set_grad_bucket_provider(red.bucket_provider)
    two = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
    for n_views, lanes in ((len(cams), None), (1, None), (len(cams), two)):
        params = {k: scene[k].clone().requires_grad_(True) for k in names}
        total, mine = sharded_step(lambda v: render_loss(params, v), params, n_views, red, streams=lanes)
        torch.cuda.synchronize()
        got = {k: params[k].grad.clone() for k in names}
        dsr.set_grad_bucket_provider(None)          # single-process reference: plain autograd accumulation
        ref = {k: scene[k].clone().requires_grad_(True) for k in names}
        for v in range(n_views):
            render_loss(ref, v).backward()
        dsr.set_grad_bucket_provider(red.bucket_provider)
        for k in names:
            e = pu.rel_err(got[k], ref[k].grad)
            if not e <= 2e-6:
                ok = False
            if rank == 0:
                print(f"{label} views={n_views} lanes={2 if lanes else 1} {k}: rel err vs single-process accumulation {e:.2e}")
    dsr.set_grad_bucket_provider(None)
# ---- block-level model (surfels generated inside preprocess): SH rows accumulate in the kernel, the five block
# gradients are added into the same bucket; one collective per batch
from partgs_b200.superquadric import BlockSurfelModel, blocks_bucket_numel, rasterize_blocks
model = BlockSurfelModel(8, 8, device=dev, generator=torch.Generator().manual_seed(3))
Pb = 8 * model.per_gs_num
shs_b = torch.zeros(Pb, 16, 3)
shs_b[:, 0] = synth.RGB2SH(torch.rand(Pb, 3, generator=torch.Generator().manual_seed(4)))
shs_b = shs_b.to(dev)
bnames = ("shs", "sq_r", "sq_s", "sq_t", "sq_eps", "sq_occ")


def block_loss(params, v):
    out = rasterize_blocks(pu.settings_from_cam(cams[v], bg), params["sq_r"], params["sq_s"], params["sq_t"],
                           params["sq_eps"], params["sq_occ"], model.alpha, model._scale, params["shs"], model.sq_eta,
                           model.sq_omega, model.faces)
    return (out[0] * ups[v]["color"]).sum() + (out[2] * ups[v]["allmap"]).sum()


def block_params():
    d = {k: getattr(model, k).detach().clone().requires_grad_(True) for k in bnames[1:]}
    d["shs"] = shs_b.clone().requires_grad_(True)
    return d


red = PeerGradAllReducer(blocks_bucket_numel(Pb, 8, 16), dev)
dsr.set_grad_bucket_provider(red.bucket_provider)
params = block_params()
sharded_step(lambda v: block_loss(params, v), params, len(cams), red, order=bnames)
torch.cuda.synchronize()
got = {k: params[k].grad.clone() for k in bnames}
dsr.set_grad_bucket_provider(None)
ref = block_params()
for v in range(len(cams)):
    block_loss(ref, v).backward()
for k in bnames:
    e = pu.rel_err(got[k], ref[k].grad)
    if not e <= 5e-6:
        ok = False
    if rank == 0:
        print(f"blocks peer views={len(cams)} {k}: rel err vs single-process accumulation {e:.2e}")

t = torch.tensor([1.0 if ok else 0.0], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("SHARDED_STEP_OK" if t.item() == 1.0 else "SHARDED_STEP_MISMATCH")
dist.destroy_process_group()
