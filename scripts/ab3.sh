#!/bin/bash
# A/B across processes of generator / device-header knobs: each argument is an env assignment list, e.g.
#   ab3.sh "" "EXB_TUNE_DEFS=EXB_OPT_MAGIC" "EXB_TUNE_PPT_W2=300"
# Runs scripts/quick_time.py (all callbacks on LV N=1e7 + parity on a small instance) once per setting, two rounds.
for rnd in 1 2; do
  for cfg in "$@"; do
    echo "=== round $rnd [$cfg]"
    env $cfg EXB_CACHE_DIR=$PWD/examodels.jl_b200/_kcache timeout 300 python scripts/quick_time.py 1e7 2>&1 | grep -v "^create"
  done
done
