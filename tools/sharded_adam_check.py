"""torchrun check (2-8 GPUs): ShardedFusedAdam over NCCL (reduce_scatter_tensor -> Adam on the slice ->
all_gather_into_tensor) equals torch.optim.Adam on the all-reduced gradients, and reports the step time of both.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/sharded_adam_check.py [--P 1000000]"""
import argparse
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--P", type=int, default=1_000_000)
    a = ap.parse_args()
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from partgs_b200.optim import FusedAdam, ShardedFusedAdam
    P = a.P
    shapes = {"xyz": (P, 3), "f_dc": (P, 1, 3), "f_rest": (P, 15, 3), "opacity": (P, 1), "scaling": (P, 2), "rotation": (P, 4)}
    lrs = {"xyz": 1.6e-4, "f_dc": 2.5e-3, "f_rest": 1.25e-4, "opacity": 0.05, "scaling": 0.005, "rotation": 0.001}
    g0 = torch.Generator().manual_seed(0)
    init = {k: torch.randn(s, generator=g0) for k, s in shapes.items()}
    pa = {k: torch.nn.Parameter(v.clone().cuda()) for k, v in init.items()}
    pb = {k: torch.nn.Parameter(v.clone().cuda()) for k, v in init.items()}
    ref = FusedAdam([{"params": [pa[k]], "lr": lrs[k], "name": k} for k in shapes], lr=0.0, eps=1e-15)
    ours = ShardedFusedAdam([{"params": [pb[k]], "lr": lrs[k], "name": k} for k in shapes])
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t_ref = t_ours = 0.0
    for step in range(5):
        gen = torch.Generator(device="cuda").manual_seed(100 * step + rank)
        grads = {k: torch.randn(s, generator=gen, device="cuda") * 0.01 for k, s in shapes.items()}
        ours.zero_grad()
        for k in shapes:
            pa[k].grad = grads[k].clone()
            pb[k].grad.copy_(grads[k])
        torch.cuda.synchronize(); dist.barrier()
        ev[0].record()
        for k in shapes:                       # the replicated baseline: all-reduce, then every rank updates everything
            dist.all_reduce(pa[k].grad)
        ref.step()
        ev[1].record()
        ev[2].record()
        ours.step()
        ev[3].record()
        torch.cuda.synchronize()
        if step >= 2:
            t_ref += ev[0].elapsed_time(ev[1]); t_ours += ev[2].elapsed_time(ev[3])
        for k in shapes:
            err = float((pa[k].detach() - pb[k].detach()).abs().max()) / (float(pa[k].detach().abs().max()) + 1e-30)
            assert err <= 5e-6, (step, k, err)
    if rank == 0:
        print(f"SHARDED_ADAM_OK world={dist.get_world_size()} P={P} replicated_ms={t_ref / 3:.3f} sharded_ms={t_ours / 3:.3f}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
