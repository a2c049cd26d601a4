#!/bin/bash
# Round-2 evidence refresh after a change of the device header (new module hashes): full ncu captures of the three LV kernels on
# the TUNED variants, the traffic file from them, then the bench line that reads it.  Outputs under gpurun_out/<tag>_*.
tag=${1:-r02x}
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-cpu --quick > /dev/null 2>&1     # writes the tune files
ncu --set full --clock-control none --import-source on -k regex:exb_hess_g0 -s 5 -c 1 -f -o gpurun_out/${tag}_prof_hess python bench.py --steps 5 --warmup 3 --no-cpu --quick > gpurun_out/${tag}_prof_hess.log 2>&1
ncu -i gpurun_out/${tag}_prof_hess.ncu-rep --page raw --csv > gpurun_out/${tag}_hess_ncu_raw.csv 2>/dev/null
for spec in "lv hessc exb_hessc_g0" "lv eval exb_eval_g0"; do
  set -- $spec
  ncu --set full --clock-control none --import-source on -k regex:"$3" -s 6 -c 1 -f -o gpurun_out/${tag}_prof_$1_$2 python scripts/prof_one.py $1 $2 > gpurun_out/${tag}_ncu_$1_$2.log 2>&1
done
mod=$(python -c "
import examodels_jl_b200 as E
from examodels_jl_b200 import models as M
p = E.Plan(M.luksan_vlcek(10_000_000)); p.compile(); print(p.module_path())")
python scripts/traffic_from_ncu.py $mod exb_hess_g0=gpurun_out/${tag}_prof_hess.ncu-rep exb_hessc_g0=gpurun_out/${tag}_prof_lv_hessc.ncu-rep exb_eval_g0=gpurun_out/${tag}_prof_lv_eval.ncu-rep > gpurun_out/${tag}_traffic.log 2>&1
cp profiles/traffic.json gpurun_out/${tag}_traffic.json
python scripts/ncu_summary.py gpurun_out/${tag}_prof_hess.ncu-rep gpurun_out/${tag}_prof_lv_hessc.ncu-rep gpurun_out/${tag}_prof_lv_eval.ncu-rep > gpurun_out/${tag}_ncu_lv_table.txt 2>&1
python bench.py > gpurun_out/${tag}_bench_line.json 2> gpurun_out/${tag}_bench.err
tail -c 600 gpurun_out/${tag}_bench_line.json; echo; cat gpurun_out/${tag}_traffic.log | head -20; head -8 gpurun_out/${tag}_ncu_lv_table.txt
