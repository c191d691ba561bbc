#!/bin/bash
# quick GPU iteration: parity tests, C3 bench (ours), optional extras.  usage: tools/gpu_quick.sh <tag> [ncu]
tag=${1:-q}; out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $out/${tag}_pytest.log
timeout 300 python tools/mask_stats.py --cfg C3 > $out/${tag}_mask_stats.json 2> $out/${tag}_mask_stats.err; cat $out/${tag}_mask_stats.json; tail -3 $out/${tag}_mask_stats.err
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu > $out/${tag}_bench_ours.json 2> $out/${tag}_bench_ours.err; cat $out/${tag}_bench_ours.json; tail -3 $out/${tag}_bench_ours.err
if [ "$2" == "ncu" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_ -s 2 -c 2 -f -o $out/${tag}_render \
   python tools/profile_step.py --cfg C3 --iters 2 > $out/${tag}_ncu.log 2>&1; tail -3 $out/${tag}_ncu.log
fi
