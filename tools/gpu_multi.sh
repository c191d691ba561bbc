#!/bin/bash
# multi-GPU bench under torchrun: tools/gpu_multi.sh <tag> <N> [extra bench args]
tag=$1; N=$2; shift; shift
out=gpurun_out; mkdir -p $out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 30 --warmup 5 --no-cpu "$@" > $out/${tag}_N${N}.json 2> $out/${tag}_N${N}.err
tail -c 2500 $out/${tag}_N${N}.json; tail -3 $out/${tag}_N${N}.err
