#!/bin/bash
# compute-sanitizer memcheck over small GPU tests: out-of-bounds / misaligned accesses in the generated and fixed kernels,
# cp.async windows, TMA tile stores.  Output: gpurun_out/<tag>_memcheck.log.  (initcheck is not used: it does not see the
# TMA bulk stores as writes, reports the whole COO buffer as uninitialised at the D2H copy, and runs for > 15 minutes.)
tag=${1:-r01}
mkdir -p gpurun_out
SEL='edge_models or lv3 or lv20 or lv_aug_20x1 or opf_small or rocket_50 or params or only_objective or lv_guide_ragged or all_ops_3'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_edge_cases.py tests/test_gpu_parity.py tests/test_gpu_products.py -m gpu -q -x -k "$SEL" > gpurun_out/${tag}_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/${tag}_memcheck.log
grep -E "ERROR SUMMARY|passed|failed|exit" gpurun_out/${tag}_memcheck.log
