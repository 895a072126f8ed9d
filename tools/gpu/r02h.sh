#!/bin/bash
# round 2, GPU call H: the plain gather kernel (variant 1) as calibration for a column-per-thread design
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
ST=tools/sweep_time.py
python $ST --workload cfg2 --kernel 1 --lus 1,2 --tag k1 > $O/r02h_sweeps.jsonl 2>$O/r02h_err.log
python $ST --workload cfg5 --kernel 1 --lus 2 --dims 0,3,5 --shapes "b>a,a>b" --tag k1 >> $O/r02h_sweeps.jsonl 2>>$O/r02h_err.log
cat $O/r02h_sweeps.jsonl
ncu --set full --clock-control none -k regex:sweep_gather -s 8 -c 1 -o /tmp/h_cfg2 -f python $ST --workload cfg2 --kernel 1 --lus 2 --dims 1 > $O/r02h_ncu1.log 2>&1
ncu -i /tmp/h_cfg2.ncu-rep --page details > $O/r02h_cfg2_details.txt
tail -3 $O/r02h_err.log
