#!/bin/bash
# compute-sanitizer passes over small configurations (run under gpurun; ~10-50x slower than a normal run).
# usage: tools/sanitize.sh [memcheck|racecheck|initcheck|synccheck]
tool=${1:-memcheck}
out=gpurun_out; mkdir -p $out
for t in "tests/test_gpu_base_raster.py::test_vs_cpu_oracle_small" \
         "tests/test_gpu_base_raster.py::test_edge_cases_empty_culled_and_mark_visible" \
         "tests/test_gpu_blocks_fused.py::test_fused_blocks_without_materialisation" \
         "tests/test_gpu_surface_maps.py::test_depth_to_normal_vs_reference_golden" \
         "tests/test_gpu_losses.py::test_photometric_vs_reference_golden" \
         "tests/test_gpu_optim.py"; do
  name=$(echo "$t" | tr '/:.' '___')
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest "$t" -x -q > $out/sanitize_${tool}_${name}.log 2>&1
  echo "$t rc=$? $(grep -c 'ERROR SUMMARY: 0 errors' $out/sanitize_${tool}_${name}.log) clean-summaries"
done
