"""torchrun check (2-8 GPUs): the fused peer-memory all-reduce (partgs_b200.dist.PeerGradAllReducer -> csrc/collective.cu)
equals, BIT FOR BIT, the rank-ordered sum ((g0 + g1) + g2) + ... of the buckets gathered with NCCL, on random buckets,
incl. buffer reuse, back-to-back batches and in-place accumulation semantics of a batch; NCCL's own all-reduce of the
same data agrees to fp32 rounding (its summation order differs for more than two ranks)."""
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
    red.begin_batch()
    b, acc = red.bucket_provider(n, dev)
    assert acc is False                      # first backward of a batch overwrites
    b.copy_(torch.randn(n, device=dev, generator=g))
    b2, acc2 = red.bucket_provider(n, dev)   # second backward of the same batch: same memory, accumulate
    assert acc2 is True and b2.data_ptr() == b.data_ptr()
    mine = red.current().clone()
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    want = gathered[0].clone()
    for q in range(1, world):
        want += gathered[q]
    nccl = mine.clone()
    dist.all_reduce(nccl)
    views = [b[:1000].view(10, 100), b[1000:].view(-1)]   # gradients are views of the bucket
    red.wait()
    red.launch(views)
    if it % 2 == 1:
        red.wait()
        torch.cuda.synchronize()
    red.wait(); torch.cuda.synchronize()
    got = red.buckets[red._cur]
    same = bool(torch.equal(got, want))
    err = float((got - nccl).abs().max() / nccl.abs().max())
    ok = ok and same and err < 1e-6
    if rank == 0:
        print(f"iter {it}: bitwise equal to the rank-ordered sum: {same}; max rel diff vs NCCL all-reduce {err:.2e}")
t = torch.tensor([1.0 if ok else 0.0], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("PEER_ALLREDUCE_OK" if t.item() == 1.0 else "PEER_ALLREDUCE_MISMATCH")
dist.destroy_process_group()
