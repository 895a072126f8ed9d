#!/bin/bash
# round 2, GPU call G (1 GPU): full ncu captures (with source) of the lean sweep kernel: cfg2 <4,4> full sweep along dim 1, cfg5 <3,2> along dim 2
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
ST=tools/sweep_time.py
ncu --set full --clock-control none --import-source on -k regex:sweep_tc -s 8 -c 1 -o /tmp/g_cfg2 -f python $ST --workload cfg2 --kernel 5 --lus 2 --dims 1 > $O/r02g_ncu1.log 2>&1
ncu -i /tmp/g_cfg2.ncu-rep --page raw --csv > $O/r02g_cfg2_raw.csv
ncu -i /tmp/g_cfg2.ncu-rep --page details > $O/r02g_cfg2_details.txt
ncu -i /tmp/g_cfg2.ncu-rep --page source --csv > $O/r02g_cfg2_source.csv 2>$O/r02g_src_err.log
ncu --set full --clock-control none --import-source on -k regex:sweep_tc -s 8 -c 1 -o /tmp/g_cfg5 -f python $ST --workload cfg5 --kernel 5 --lus 2 --dims 2 --shapes "b>a" > $O/r02g_ncu2.log 2>&1
ncu -i /tmp/g_cfg5.ncu-rep --page raw --csv > $O/r02g_cfg5_raw.csv
ncu -i /tmp/g_cfg5.ncu-rep --page details > $O/r02g_cfg5_details.txt
ncu -i /tmp/g_cfg5.ncu-rep --page source --csv > $O/r02g_cfg5_source.csv 2>>$O/r02g_src_err.log
ls -la $O/r02g_* /tmp/*.ncu-rep
cp /tmp/g_cfg2.ncu-rep $O/r02g_cfg2.ncu-rep
