#!/bin/bash
# Round-2 evidence on one B200: bench line (own arm + reference arm), ncu launch list of the same command, full ncu captures
# of the headline kernel and of the new kernels (raw metrics exported), on a box whose tune files were written by the bench
# run just before (so every ncu capture sees the TUNED launch-shape variant).  Outputs under gpurun_out/<tag>_*.
tag=${1:-r02}
mkdir -p gpurun_out
python bench.py > gpurun_out/${tag}_bench_line.json 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${tag}_bench_reference_arm.json 2>> gpurun_out/${tag}_bench.err
# launch list (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu --quick > gpurun_out/${tag}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:exb_hess_g0 -s 5 -c 1 -f -o gpurun_out/${tag}_prof_hess python bench.py --steps 5 --warmup 3 --no-cpu --quick > gpurun_out/${tag}_prof_hess.log 2>&1
ncu -i gpurun_out/${tag}_prof_hess.ncu-rep --page raw --csv > gpurun_out/${tag}_hess_ncu_raw.csv 2>/dev/null
for spec in "lv hessc exb_hessc_g0" "lv eval exb_eval_g0" "family grad exb_gradt_g0" "family hess exb_hess_g0" "family hessc exb_hessc_g0" "rocket hess exb_hess_g0" "rocket eval exb_eval_g0"; do
  set -- $spec
  ncu --set full --clock-control none --import-source on -k regex:"$3" -s 6 -c 1 -f -o gpurun_out/${tag}_prof_$1_$2 python scripts/prof_one.py $1 $2 > gpurun_out/${tag}_ncu_$1_$2.log 2>&1
done
python scripts/ncu_summary.py gpurun_out/${tag}_prof_hess.ncu-rep gpurun_out/${tag}_prof_lv_hessc.ncu-rep gpurun_out/${tag}_prof_lv_eval.ncu-rep > gpurun_out/${tag}_ncu_lv_table.txt 2>&1
python scripts/ncu_summary.py gpurun_out/${tag}_prof_family_grad.ncu-rep gpurun_out/${tag}_prof_family_hess.ncu-rep gpurun_out/${tag}_prof_family_hessc.ncu-rep gpurun_out/${tag}_prof_rocket_hess.ncu-rep gpurun_out/${tag}_prof_rocket_eval.ncu-rep > gpurun_out/${tag}_ncu_other_table.txt 2>&1
tail -c 400 gpurun_out/${tag}_bench_line.json; echo; tail -c 300 gpurun_out/${tag}_bench_reference_arm.json; cat gpurun_out/${tag}_ncu_lv_table.txt | head -14
