#!/bin/bash
# A/B across processes (knobs that are read once per process); usage: ab2.sh "ENV=.. ENV=.." "ENV=.." ...
for rnd in 1 2 3; do
  for cfg in "$@"; do
    echo "round $rnd [$cfg]: $(env $cfg python scripts/tune.py lv:1e7 2>&1 | tail -1)"
  done
done
