#!/bin/bash
# round 2, GPU call U (1 GPU): profiles of record -- launch list of the benchmark stage, full captures of the two roofline kernels (traffic),
# the live-reference run at NMAX = 9, the default bench and the reference arm with their wall times
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
ST=tools/sweep_time.py
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02u_launches_cfg5.csv python bench.py --no-cpu --no-secondary --steps 1 --warmup 3 --no-graph > $O/r02u_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_col -s 6 -c 3 -o /tmp/u_col -f python $ST --workload cfg5 --kernel 0 --lus 2 --dims 0 --shapes "b>a" > $O/r02u_ncu_col.log 2>&1
ncu -i /tmp/u_col.ncu-rep --page raw --csv > $O/r02u_col_raw.csv
ncu -i /tmp/u_col.ncu-rep --page details --kernel-id :::1 > $O/r02u_col_details.txt 2>/dev/null || ncu -i /tmp/u_col.ncu-rep --page details > $O/r02u_col_details.txt
ncu -i /tmp/u_col.ncu-rep --page source --csv --kernel-id :::1 > $O/r02u_col_source.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_tc -s 6 -c 3 -o /tmp/u_tc -f python $ST --workload cfg2 --kernel 0 --lus 2 --dims 1 > $O/r02u_ncu_tc.log 2>&1
ncu -i /tmp/u_tc.ncu-rep --page raw --csv > $O/r02u_tc_raw.csv
ncu -i /tmp/u_tc.ncu-rep --page details --kernel-id :::1 > $O/r02u_tc_details.txt 2>/dev/null || ncu -i /tmp/u_tc.ncu-rep --page details > $O/r02u_tc_details.txt
examples/live_burgers_adapt -NM 9 -N0 2 -steps 10 > $O/r02u_live_n9.log 2>&1; tail -4 $O/r02u_live_n9.log
( time python bench.py > $O/r02u_bench_default.json 2>$O/r02u_err.log ) 2>&1 | grep real
( time python bench.py --impl reference > $O/r02u_bench_ref.json 2>>$O/r02u_err.log ) 2>&1 | grep real
head -c 1500 $O/r02u_bench_ref.json; echo
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02u_bench_default.json')); c=d['config']; s=d.get('secondary',{})
print('cfg5 stage ms', d['ms_per_step'], 'value %.3e'%d['value'], 'launches', c['launches_per_stage'], 'parity', c['parity_rel_l2'], 'roof', d['roofline']['frac'], d['roofline']['us_per_launch'], 'e2e ms', d['e2e']['ms_per_step'], 'clocks', d['clocks'])
print('cpu_baseline', d.get('cpu_baseline'))
if s: print('   cfg2 ms', s['ms_per_step'], 'value %.3e'%s['value'], 'roof', s['roofline']['frac'], s['roofline']['us_per_launch'], 'e2e ms', s['e2e']['ms_per_step'], s.get('cpu_baseline'))
PY
grep -v "^frame" $O/r02u_err.log | tail -5
