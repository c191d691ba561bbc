#!/bin/bash
# quick A/B: bench only (+ optional probe)  usage: tools/gpu_r2c.sh <tag> [probe]
tag=${1:-r2c}; out=gpurun_out; mkdir -p $out
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --no-extras > $out/${tag}_bench.json 2> $out/${tag}_bench.err; python -c "
import json; d=json.load(open('$out/${tag}_bench.json')); print(d['value'], d['ms_per_step'], d['stage_ms_per_step'])"; tail -2 $out/${tag}_bench.err
if [ "$2" == "probe" ]; then timeout 600 python tools/grad_gate_probe.py > $out/${tag}_probe.jsonl 2> $out/${tag}_probe.err; cat $out/${tag}_probe.jsonl; tail -3 $out/${tag}_probe.err; fi
