#!/bin/bash
# round 2, GPU call AD (1 GPU): GPU suite with the generated-table tests, bench with library-generated tables
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
alive() { timeout 20 nvidia-smi --query-gpu=name,memory.used --format=csv,noheader 2>&1 | head -n 1; }
( time timeout 1200 python -m pytest tests -x -q -m gpu > $O/r02ad_pytest.log 2>&1 ) 2>&1 | grep real; tail -n 4 $O/r02ad_pytest.log; alive
( time python bench.py > $O/r02ad_bench.json 2>$O/r02ad_err.log ) 2>&1 | grep real
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02ad_bench.json').read().strip().splitlines()[-1]); c=d['config']; s=d.get('secondary',{})
print('cfg5 stage ms', d['ms_per_step'], 'value %.3e'%d['value'], 'parity', c.get('parity_rel_l2'), 'roof', d['roofline']['frac'], 'e2e ms', d['e2e']['ms_per_step'], 'clocks', d['clocks'])
print('cpu_baseline', d.get('cpu_baseline'))
if s: print('   cfg2 ms', s['ms_per_step'], 'value %.3e'%s['value'], 'roof', s['roofline']['frac'], 'e2e ms', s['e2e']['ms_per_step'])
PY
grep -v "^frame" $O/r02ad_err.log | tail -n 5; alive
