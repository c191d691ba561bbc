"""Probe: does torch symmetric memory (peer-mapped buffers + GPU-side barrier) work on this box?"""
import os, time, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
n = 58_000_000
buf = symm_mem.empty(n, dtype=torch.float32, device=dev)
hdl = symm_mem.rendezvous(buf, group=dist.group.WORLD.group_name)
buf.fill_(float(rank + 1))
hdl.barrier()
sl = n // world
peers = [hdl.get_buffer(p, (n,), torch.float32) for p in range(world)]
tmp = torch.empty(world, sl, device=dev)
torch.cuda.synchronize(); dist.barrier()
def allreduce_ce():
    hdl.barrier()
    for j in range(world):
        p = (rank + j) % world
        tmp[p].copy_(peers[p][rank * sl:(rank + 1) * sl], non_blocking=True)
    red = tmp.sum(0)
    buf[rank * sl:(rank + 1) * sl].copy_(red)
    hdl.barrier()
    for j in range(1, world):
        p = (rank + j) % world
        peers[p][rank * sl:(rank + 1) * sl].copy_(red, non_blocking=True)
    hdl.barrier()
allreduce_ce(); torch.cuda.synchronize()
exp = sum(range(1, world + 1))
ok = bool((buf[: sl * world] == exp).all())
buf.fill_(float(rank + 1)); hdl.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    allreduce_ce()
e1.record(); torch.cuda.synchronize()
t_ce = e0.elapsed_time(e1) / 5
x = torch.ones(n, device=dev)
dist.all_reduce(x); torch.cuda.synchronize()
e0.record()
for _ in range(5):
    dist.all_reduce(x)
e1.record(); torch.cuda.synchronize()
if rank == 0:
    print(f"symm_mem ok={ok} world={world} ce_allreduce_ms={t_ce:.3f} nccl_allreduce_ms={e0.elapsed_time(e1)/5:.3f}")
dist.destroy_process_group()
