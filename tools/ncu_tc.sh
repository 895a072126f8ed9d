#!/bin/bash
# full ncu capture of the lean tensor-core sweep kernel (single-job sweeps at the cfg2 size); CSV pages come back under gpurun_out/
R=${1:-r01}
mkdir -p gpurun_out
AMDG_KERNEL=5 ncu --set full --clock-control none --import-source on -k regex:sweep_tc -s 12 -c 6 -o /tmp/${R}_tc -f python tools/prof_sweep.py sweeps > gpurun_out/${R}_tc_full.log 2>&1
ncu -i /tmp/${R}_tc.ncu-rep --page raw --csv > gpurun_out/${R}_sweep_tc_raw.csv
ncu -i /tmp/${R}_tc.ncu-rep --page source --csv --kernel-id :::1 > gpurun_out/${R}_sweep_tc_source.csv 2>/dev/null
ncu -i /tmp/${R}_tc.ncu-rep --page details --kernel-id :::1 > gpurun_out/${R}_sweep_tc_details.txt 2>/dev/null
