#!/bin/bash
# Round evidence on one B200: bench line (own arm + reference arm), ncu launch list of the same command, one full ncu
# capture of the dominant kernel (+ raw metrics), per-config timings.  Outputs under gpurun_out/<tag>_*.
tag=${1:-r01}
mkdir -p gpurun_out
python bench.py > gpurun_out/${tag}_bench_line.json 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${tag}_bench_reference_arm.json 2>> gpurun_out/${tag}_bench.err
# launch list (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/${tag}_launches_bench.log 2>&1
# tune files exist now (written by the runs above), so launch 0 of a kernel in a new process is already the tuned variant
ncu --set full --clock-control none --import-source on -k regex:exb_hess -s 5 -c 1 -f -o gpurun_out/${tag}_prof_hess python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/${tag}_prof_hess.log 2>&1
ncu -i gpurun_out/${tag}_prof_hess.ncu-rep --page raw --csv > gpurun_out/${tag}_hess_ncu_raw.csv 2>/dev/null
for cb in jac cons grad obj; do
  ncu --set full --clock-control none --import-source on -k regex:"exb_(ggrad|sgrad|cons|jac|obj)_g0" -s 3 -c 1 -f -o gpurun_out/${tag}_prof_$cb python scripts/prof_one.py lv $cb > gpurun_out/${tag}_ncu_$cb.log 2>&1
done
python scripts/bench_configs.py $tag > gpurun_out/${tag}_configs.log 2>&1
tail -c 600 gpurun_out/${tag}_bench_line.json; echo; tail -c 300 gpurun_out/${tag}_bench_reference_arm.json; ls -la gpurun_out | grep ${tag}_ | head -30
