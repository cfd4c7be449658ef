#!/bin/bash
# compute-sanitizer (racecheck for the shared-memory staging / transposed reduction / radix ranking, memcheck for everything) on
# the raster, sort and GPU-driven pipeline tests.  Run on a B200 through gpurun; logs -> gpurun_out/sanitizer_*.log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="tests/test_gpu_ops.py::test_raster_forward_backward tests/test_gpu_parity_r2.py::test_raster_backward_statistics tests/test_gpu_parity_r2.py::test_raster_specific_tiles tests/test_gpu_sort.py tests/test_gpu_workspace.py::test_workspace_forward_equals_synchronising_forward tests/test_gpu_pipeline.py::test_fused_pairs_match_oracle_lists"
for tool in racecheck memcheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 0 python -m pytest $T -m gpu -q -x -k "not bulk" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_$tool.log | tail -4
done
