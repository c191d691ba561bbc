#!/bin/bash
# First GPU call of round 2: everything that was written after round 1's GPU budget ran out.
#   1. the hardware parity tests of the kernels written on the emulator (densification, extraction epilogue / loop,
#      regularisers; green on a B200 since the last call of round 1, profiles/r1_gpu_pytest_new_kernels.log);
#   2. the whole verified GPU suite + smoke (nothing on the render path changed, this is the regression check);
#   3. timing rows for the new ops against the reference's torch sequences (tools/time_rank34.py);
#   4. compute-sanitizer memcheck over the new GPU tests.
# usage: /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_round2_first.sh r2a'
tag=${1:-r2a}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_zz_densify.py tests/test_gpu_zz_extract.py tests/test_gpu_zz_regularizers.py \
    -q -rxXs > $out/${tag}_new_kernels.log 2>&1; echo "new kernels rc=$?" | tee -a $out/${tag}_new_kernels.log
tail -30 $out/${tag}_new_kernels.log
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/${tag}_smoke.log
timeout 600 python tools/time_rank34.py > $out/${tag}_rank34.jsonl 2> $out/${tag}_rank34.err; cut -c1-300 $out/${tag}_rank34.jsonl
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_zz_densify.py \
    tests/test_gpu_zz_regularizers.py -q -k "golden or 5_000 or 77" > $out/${tag}_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a $out/${tag}_memcheck.log; tail -5 $out/${tag}_memcheck.log
# 5. the prepared cooperative-SH forward preprocess (PGS_SH_COOP=1): bit-exact parity against the reference build, then timing
PGS_SH_COOP=1 timeout 600 python -m pytest tests/test_gpu_base_raster.py tests/test_gpu_blocks_fused.py -q > $out/${tag}_shcoop_pytest.log 2>&1
echo "sh-coop pytest rc=$?" | tee -a $out/${tag}_shcoop_pytest.log; tail -3 $out/${tag}_shcoop_pytest.log
timeout 300 python bench.py --no-cpu > $out/${tag}_bench_default.json 2> $out/${tag}_bench_default.err
PGS_SH_COOP=1 timeout 300 python bench.py --no-cpu > $out/${tag}_bench_shcoop.json 2> $out/${tag}_bench_shcoop.err
TAG=$tag python - <<'PY'
import json, os
for name in ("default", "shcoop"):
    try:
        d = json.load(open(f"gpurun_out/{os.environ['TAG']}_bench_{name}.json"))
        print(name, d["value"], "frames/s; preprocess_fwd", d["stage_ms_per_step"]["preprocess_fwd"], "ms")
    except Exception as ex:
        print(name, "unreadable:", ex)
PY
