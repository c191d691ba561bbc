#!/bin/bash
# round 2: one `ncu --set full` capture per kernel and configuration (third invocation of every kernel function of
# the target).  The reports are summarised ON THE BOX (tools/ncu_summary.py, tools/ncu_lines.py for the render
# kernels) — only the text comes back under gpurun_out/ (the reports themselves exceed the 64 MiB return limit);
# `keep` as first configuration keeps that one report.
# usage: tools/gpu_r2_ncu.sh <tag> [c3] [c2] [c4] [c5] [ops]
tag=$1; shift; out=gpurun_out; mkdir -p $out; tmp=/tmp/ncu_$tag; mkdir -p $tmp
NCU="ncu --set full --clock-control none --import-source on --kernel-id :::3 -f"
for what in "$@"; do
  case $what in
    c3) args="--cfg C3 --iters 3" ;;
    c2) args="--cfg C2 --iters 3" ;;
    c4) args="--part --iters 3" ;;
    c5) args="--blocks --iters 3" ;;
    ops) args="--ops --iters 3" ;;
  esac
  timeout 1200 $NCU -o $tmp/${tag}_$what python tools/profile_step.py $args > $out/${tag}_$what.log 2>&1
  tail -1 $out/${tag}_$what.log
  python tools/ncu_summary.py $tmp/${tag}_$what.ncu-rep > $out/${tag}_ncu_full_$what.txt 2>&1
  if [ "$what" == "c3" ] || [ "$what" == "c4" ]; then
    python tools/ncu_lines.py $tmp/${tag}_$what.ncu-rep render_fwd 40 > $out/${tag}_ncu_lines_${what}_render_fwd.txt 2>&1
    python tools/ncu_lines.py $tmp/${tag}_$what.ncu-rep render_bwd 40 > $out/${tag}_ncu_lines_${what}_render_bwd.txt 2>&1
  fi
  rm -f $tmp/${tag}_$what.ncu-rep
done
