#!/bin/bash
# round 2, GPU call F (1 GPU): state of the tree after the re-entry -- whole GPU test suite, the three tensor-core sweep kernels side by side, bench
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r02f_smi.txt
( time timeout 1500 python -m pytest tests -x -q -m gpu > $O/r02f_pytest.log 2>&1 ) 2>&1 | grep real
tail -15 $O/r02f_pytest.log
ST=tools/sweep_time.py
: > $O/r02f_sweeps.jsonl
for k in 5 6 7; do
  python $ST --workload cfg2 --kernel $k --tag k$k >> $O/r02f_sweeps.jsonl 2>>$O/r02f_err.log
  python $ST --workload cfg5 --kernel $k --lus 2 --dims 0,3,5 --tag k$k >> $O/r02f_sweeps.jsonl 2>>$O/r02f_err.log
done
cat $O/r02f_sweeps.jsonl
for k in 0 6 7; do
  python bench.py --no-cpu --steps 10 --kernel $k > $O/r02f_bench_k$k.json 2>>$O/r02f_err.log
  python - <<PY
import json
try:
    d=json.load(open('$O/r02f_bench_k$k.json')); c=d['config']; s=d.get('secondary',{})
    print('kernel $k: cfg5 stage ms', d['ms_per_step'], 'value %.3e'%d['value'], 'launches', c['launches_per_stage'], 'parity', c['parity_rel_l2'], 'roof', d['roofline']['frac'], d['roofline']['us_per_launch'], 'e2e ms', d['e2e']['ms_per_step'])
    if s: print('   cfg2 ms', s['ms_per_step'], 'value %.3e'%s['value'], 'roof', s['roofline']['frac'], s['roofline']['us_per_launch'], 'e2e ms', s['e2e']['ms_per_step'])
except Exception as e: print('kernel $k: no bench', e)
PY
done
python bench.py --workload cfg4 --no-cpu --steps 10 > $O/r02f_bench_cfg4.json 2>>$O/r02f_err.log
python -c "
import json
d=json.load(open('$O/r02f_bench_cfg4.json')); print('cfg4 stage ms', d['ms_per_step'], d['value'], d['config']['launches_per_stage'])
"
examples/live_burgers_adapt -NM 6 -N0 2 -steps 10 > $O/r02f_live_n6.log 2>&1; tail -4 $O/r02f_live_n6.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02f_launches_cfg5.csv python bench.py --no-cpu --no-secondary --steps 1 --warmup 3 --no-graph > $O/r02f_under_ncu.log 2>&1
grep -v "^frame" $O/r02f_err.log | tail -8
