#!/bin/bash
# round 2, first GPU call: packed-f32x2 backward kernel — parity, racecheck, bench (4 vs 8 warps per CTA), ncu capture
tag=${1:-r2a}; out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $out/${tag}_pytest.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest "tests/test_gpu_base_raster.py::test_vs_cpu_oracle_small" -x -q > $out/${tag}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 $out/${tag}_racecheck.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --no-extras > $out/${tag}_bench_dflt.json 2> $out/${tag}_bench_dflt.err; cat $out/${tag}_bench_dflt.json; tail -3 $out/${tag}_bench_dflt.err
PGS_BWD_WARPS_PER_CTA=1 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --no-extras > $out/${tag}_bench_nw1.json 2> $out/${tag}_bench_nw1.err; cat $out/${tag}_bench_nw1.json
PGS_BWD_WARPS_PER_CTA=2 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --no-extras > $out/${tag}_bench_nw2.json 2> $out/${tag}_bench_nw2.err; cat $out/${tag}_bench_nw2.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_ -s 2 -c 2 -f -o $out/${tag}_render \
   python tools/profile_step.py --cfg C3 --iters 2 > $out/${tag}_ncu.log 2>&1; tail -3 $out/${tag}_ncu.log
