#!/bin/bash
# One GPU-box batch: parity tests, smoke, both bench arms, ncu launch list + full capture of one C3 step.
# usage: tools/gpu_batch.sh <tag> [skip-tests]
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
if [ "$2" != "skip-tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/${tag}_smoke.log
fi
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
timeout 600 python bench.py --steps 30 --warmup 5 > $out/${tag}_bench_ours.json 2> $out/${tag}_bench_ours.err
tail -c 3000 $out/${tag}_bench_ref.json; tail -c 4000 $out/${tag}_bench_ours.json
# launch list of the bench command itself (share of the step per kernel)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 320 -c 64 --csv \
   --log-file $out/${tag}_launches_bench.csv python bench.py --steps 4 --warmup 3 --no-cpu > $out/${tag}_launches_bench.log 2>&1
# full capture of the second C3 step (all kernels of one fwd+bwd)
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^(render_|preprocess_|rs_|scan_k|duplicate_|identify_|tile_order)' -s 16 -c 16 -f -o $out/${tag}_c3_step \
   python tools/profile_step.py --cfg C3 --iters 2 > $out/${tag}_ncu_full.log 2>&1
python tools/config_table.py > $out/${tag}_configs.jsonl 2> $out/${tag}_configs.err; cat $out/${tag}_configs.jsonl | cut -c1-200
ls -la $out | tail -20
