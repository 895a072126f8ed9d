#!/bin/bash
# round 2, GPU call V (1 GPU): GPU suite, bench with the stage-mix roofline and the field-through-map point-wise flux, DRAM bytes of the stage's launches
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
( time timeout 1200 python -m pytest tests -x -q -m gpu > $O/r02v_pytest.log 2>&1 ) 2>&1 | grep real
tail -3 $O/r02v_pytest.log
python bench.py --no-cpu --steps 10 --breakdown > $O/r02v_bench.json 2>$O/r02v_err.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02v_bench.json')); c=d['config']; s=d.get('secondary',{}); r=d['roofline']
print('cfg5 stage ms', d['ms_per_step'], 'value %.3e'%d['value'], 'launches', c['launches_per_stage'], 'parity', c['parity_rel_l2'], 'e2e ms', d['e2e']['ms_per_step'])
print('roofline', {k:r[k] for k in ('achieved','frac','launches','sweep_jobs','bytes_per_launch','us_per_launch','pass_ms','traffic')})
print('breakdown', {k:round(v,3) for k,v in c.get('breakdown_ms_eager_max_over_ranks',{}).items()})
if s: print('   cfg2 ms', s['ms_per_step'], 'roof', s['roofline']['frac'], s['roofline']['us_per_launch'])
PY
python bench.py --workload cfg4 --no-cpu --steps 10 > $O/r02v_bench_cfg4.json 2>>$O/r02v_err.log
python -c "
import json
d=json.load(open('$O/r02v_bench_cfg4.json')); print('cfg4 stage ms', d['ms_per_step'], d['value'], d['config']['launches_per_stage'], d['roofline']['frac'])
"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:sweep_col -c 1400 --csv --log-file $O/r02v_stage_dram.csv python bench.py --no-cpu --no-secondary --steps 1 --warmup 3 --no-graph > $O/r02v_under_ncu.log 2>&1
grep -v "^frame" $O/r02v_err.log | tail -6
