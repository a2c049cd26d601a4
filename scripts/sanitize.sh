#!/bin/bash
# compute-sanitizer passes over small GPU tests: memcheck (out-of-bounds / misaligned accesses in the generated and fixed
# kernels, cp.async windows, TMA tile stores) and initcheck (reads of device memory nobody wrote: partial x uploads of
# sharded handles, staging buffers).  Output: gpurun_out/<tag>_memcheck.log, <tag>_initcheck.log
tag=${1:-r01}
mkdir -p gpurun_out
SEL='edge_models or lv3 or lv20 or lv_aug_20x1 or opf_small or rocket_50 or params or only_objective or lv_guide_ragged or all_ops_3'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_edge_cases.py tests/test_gpu_parity.py tests/test_gpu_products.py -m gpu -q -x -k "$SEL" > gpurun_out/${tag}_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/${tag}_memcheck.log
timeout 900 compute-sanitizer --tool initcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py tests/test_gpu_products.py -m gpu -q -x -k "sharded_host_shims or sharded_handle or (pipelined and rocket) or lv_sharded" > gpurun_out/${tag}_initcheck.log 2>&1
echo "initcheck exit $?" >> gpurun_out/${tag}_initcheck.log
grep -E "ERROR SUMMARY|passed|failed|exit" gpurun_out/${tag}_memcheck.log gpurun_out/${tag}_initcheck.log
