#!/bin/bash
# round 2, GPU call AG (1 GPU): adaptive mode in the live run (auto kernel choice), GPU suite, launch list of the final benchmark stage (fused RK plan),
# the small workloads for the record
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
for i in 1 2; do examples/live_burgers_adapt -NM 9 -N0 2 -steps 10 -gen 1 2>&1 | grep "wall per step" ; done | tee $O/r02ag_live_n9.log
( time timeout 1200 python -m pytest tests -x -q -m gpu > $O/r02ag_pytest.log 2>&1 ) 2>&1 | grep real; tail -n 3 $O/r02ag_pytest.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02ag_launches_cfg5.csv python bench.py --no-cpu --no-secondary --steps 1 --warmup 3 --no-graph > $O/r02ag_under_ncu.log 2>&1
grep -c sweep_col $O/r02ag_launches_cfg5.csv
for wl in cfg4 cfg1 cfg3; do
python bench.py --workload $wl --no-cpu --steps 10 > $O/r02ag_bench_$wl.json 2>>$O/r02ag_err.log
python -c "
import json
d=json.loads([l for l in open('$O/r02ag_bench_$wl.json') if l.startswith('{')][-1]); print('$wl ms', d['ms_per_step'], 'value %.3e'%d['value'], d['config'].get('launches_per_stage'), d.get('roofline',{}).get('frac'), 'e2e', d.get('e2e',{}).get('ms_per_step'))
"
done
grep -v "^frame" $O/r02ag_err.log | tail -n 4
timeout 20 nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
