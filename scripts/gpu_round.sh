#!/bin/bash
# One GPU-box visit: parity tests, the bench line, per-callback timings, optional e2e windows A/B and ncu captures.
# usage: gpu_round.sh <tag> [sections: tests,bench,quick,e2e,ncu:<cb,cb,..>,configs]
tag=${1:-r01b}; sections=${2:-tests,bench,quick}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
case ",$sections," in *,tests,*)
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
  tail -5 gpurun_out/${tag}_pytest.log;; esac
case ",$sections," in *,bench,*)
  timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 1800 gpurun_out/${tag}_bench.json;; esac
case ",$sections," in *,quick,*)
  timeout 300 python scripts/quick_time.py 1e7 > gpurun_out/${tag}_quick.log 2>&1; cat gpurun_out/${tag}_quick.log;; esac
case ",$sections," in *,e2e,*)
  for w in 1 8 16; do EXB_HOST_WINDOWS=$w timeout 200 python scripts/e2e_time.py 1e7 2>&1 | tail -2; done | tee gpurun_out/${tag}_e2e.log;; esac
case ",$sections," in *,configs,*)
  timeout 900 python scripts/bench_configs.py $tag > gpurun_out/${tag}_configs.log 2>&1; tail -5 gpurun_out/${tag}_configs.log | cut -c1-600;; esac
for s in ${sections//,/ }; do
  case $s in ncu:*)
    for cb in $(echo ${s#ncu:} | tr '+' ' '); do
      timeout 300 ncu --set full --clock-control none --import-source on -k regex:"exb_(hess|hessc|eval|ggrad|sgrad|cons|jac|obj)_g0" -s 10 -c 1 -f -o gpurun_out/${tag}_prof_$cb python scripts/prof_one.py lv $cb > gpurun_out/${tag}_ncu_$cb.log 2>&1
    done;; esac
done
ls -la gpurun_out | tail -12
