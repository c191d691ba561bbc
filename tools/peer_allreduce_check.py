"""torchrun check (N >= 2 GPUs): PeerGradAllReducer == NCCL all-reduce on random buckets, incl. bucket reuse."""
import os, sys
from pathlib import Path
import torch, torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from partgs_b200.dist import init_nccl, PeerGradAllReducer
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
init_nccl(local)
dev = torch.device("cuda", local)
n = 5_000_123
red = PeerGradAllReducer(n, dev)
ok = True
for it in range(6):
    g = torch.Generator(device=dev).manual_seed(1000 * it + rank)
    b = red.bucket_provider(n, dev)
    b.copy_(torch.randn(n, device=dev, generator=g))
    ref = b.clone()
    dist.all_reduce(ref)
    views = [b[:1000].view(10, 100), b[1000:].view(-1)]   # gradients are views of the bucket
    red.wait()
    red.launch(views)
    if it % 2 == 1:
        red.wait()
        torch.cuda.synchronize()
    red.wait(); torch.cuda.synchronize()
    err = float((b - ref).abs().max() / ref.abs().max())
    ok = ok and err < 1e-6
    if rank == 0:
        print(f"iter {it}: max rel err vs NCCL {err:.2e}")
t = torch.tensor([1.0 if ok else 0.0], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("PEER_ALLREDUCE_OK" if t.item() == 1.0 else "PEER_ALLREDUCE_MISMATCH")
dist.destroy_process_group()
