#!/bin/bash
# round 2, GPU call AU (1 GPU): profiles of record on the final tree -- launch list of the benchmark stage, full capture of the column kernel (the roofline
# kernel), DRAM bytes of the 3 -> 2 launches of a stage
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
ST=tools/sweep_time.py
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02au_launches_cfg5.csv python bench.py --no-cpu --no-secondary --steps 1 --warmup 3 --no-graph > $O/r02au_under_ncu.log 2>&1
grep -c sweep_col $O/r02au_launches_cfg5.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_col -s 6 -c 3 -o /tmp/au_col -f python $ST --workload cfg5 --kernel 0 --lus 2 --dims 0 --shapes "b>a" > $O/r02au_ncu_col.log 2>&1
ncu -i /tmp/au_col.ncu-rep --page raw --csv > $O/r02au_col_raw.csv
ncu -i /tmp/au_col.ncu-rep --page details --kernel-id :::1 > $O/r02au_col_details.txt 2>/dev/null || ncu -i /tmp/au_col.ncu-rep --page details > $O/r02au_col_details.txt
ls -la $O/r02au_* | awk '{print $5, $9}'
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:sweep_col -c 1400 --csv --log-file $O/r02au_stage_dram.csv python bench.py --no-cpu --no-secondary --steps 1 --warmup 3 --no-graph > $O/r02au_under_ncu2.log 2>&1
grep -c sweep_col $O/r02au_stage_dram.csv
