#!/bin/bash
# Round profile capture (run on the GPU box through gpurun): launch list of the bench command, then full captures of the
# lean tensor-core sweep kernel (single-job sweeps, and the batched launches of the bench step); reports stay in /tmp, CSV
# pages come back under gpurun_out/.
set -x
R=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${R}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-graph > gpurun_out/${R}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweep_tc -s 12 -c 12 -o /tmp/${R}_single -f python tools/prof_sweep.py sweeps > gpurun_out/${R}_full_single.log 2>&1
ncu -i /tmp/${R}_single.ncu-rep --page raw --csv > gpurun_out/${R}_sweep_tc_single_raw.csv
ncu -i /tmp/${R}_single.ncu-rep --page source --csv --kernel-id :::1 > gpurun_out/${R}_sweep_tc_single_source.csv 2>/dev/null
ncu -i /tmp/${R}_single.ncu-rep --page details --kernel-id :::1 > gpurun_out/${R}_sweep_tc_single_details.txt 2>/dev/null
ncu --set full --clock-control none -k regex:sweep_tc -s 150 -c 10 -o /tmp/${R}_step -f python bench.py --steps 1 --warmup 3 --no-graph > gpurun_out/${R}_full_step.log 2>&1
ncu -i /tmp/${R}_step.ncu-rep --page raw --csv > gpurun_out/${R}_sweep_tc_step_raw.csv
ls -la gpurun_out /tmp/*.ncu-rep
