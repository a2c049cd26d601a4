#!/bin/bash
# compute-sanitizer racecheck (shared-memory hazards) over the kernels that stage through shared memory: the TMA tile store
# (smem writes -> fence.proxy.async -> barrier -> cp.async.bulk by thread 0), the column-tile kernels (double-buffered staging,
# one barrier per pattern, gather by the column owner) and the fused evaluation kernel (two tiles, block reduction).
# Output: gpurun_out/<tag>_racecheck.log
tag=${1:-r02}
mkdir -p gpurun_out
SEL='lv_382 or lv_5 or lv_guide_700 or lv_1003 or lv_param_300 or only_objective or lv100 or family_1000 or mixed_gradient or rocket_50'
EXB_TUNE_PERSISTENT=0 timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_compressed.py tests/test_gpu_eval.py tests/test_edge_cases.py tests/test_gpu_products.py -m gpu -q -x -k "$SEL" > gpurun_out/${tag}_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/${tag}_racecheck.log
grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed|exit|hazard" gpurun_out/${tag}_racecheck.log | tail -12
