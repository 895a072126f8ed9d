#!/bin/bash
# round 2, GPU call AQ (1 GPU): the driver's round-end sequence on the final tree -- GPU suite, smoke(), default bench, reference arm; live adaptive run
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
alive() { timeout 20 nvidia-smi --query-gpu=name,memory.used --format=csv,noheader 2>&1 | head -n 1; }
( time timeout 1200 python -m pytest tests -x -q -m gpu > $O/r02aq_pytest.log 2>&1 ) 2>&1 | grep real; tail -n 4 $O/r02aq_pytest.log; alive
( time python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | grep -E "smoke|real"
( time python bench.py > $O/r02aq_bench.json 2>$O/r02aq_err.log ) 2>&1 | grep real
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02aq_bench.json').read().strip().splitlines()[-1]); c=d['config']; s=d.get('secondary',{})
print('cfg5 stage ms', d['ms_per_step'], 'value %.3e'%d['value'], 'launches', c['launches_per_stage'], 'gpu_launches', d.get('gpu_launches'), 'parity', c.get('parity_rel_l2'), 'roof', d['roofline']['frac'], 'traffic', d['roofline'].get('traffic'), 'e2e ms', d['e2e']['ms_per_step'], 'clocks', d['clocks'])
print('cpu_baseline', {k: d['cpu_baseline'][k] for k in ('value','cores','kind')})
if s: print('   cfg2 ms', s['ms_per_step'], 'value %.3e'%s['value'], 'roof', s['roofline']['frac'], 'e2e ms', s['e2e']['ms_per_step'])
PY
( time python bench.py --impl reference --steps 1 --warmup 0 > $O/r02aq_bench_ref.json 2>>$O/r02aq_err.log ) 2>&1 | grep real
python -c "
import json
d=json.loads(open('gpurun_out/r02aq_bench_ref.json').read().strip().splitlines()[-1]); print('reference arm', d.get('value'), d.get('unit'), d.get('config',{}).get('workload','')[:80], d.get('cpu_baseline',{}).get('sample','')[:160])
"
for i in 1 2; do examples/live_burgers_adapt -NM 9 -N0 2 -steps 10 -gen 1 2>&1 | grep "wall per step"; done
grep -v "^frame" $O/r02aq_err.log | tail -n 3; alive
