#!/bin/bash
# round 2, GPU call AI (1 GPU): GPU suite with the new point-wise / Vlasov-Ampere tests, cfg4 at NMAX=9 (ours and the reference arm)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
( time timeout 1200 python -m pytest tests -x -q -m gpu > $O/r02ai_pytest.log 2>&1 ) 2>&1 | grep real; tail -n 6 $O/r02ai_pytest.log
( time python bench.py --workload cfg4 --steps 10 > $O/r02ai_bench_cfg4.json 2>$O/r02ai_err.log ) 2>&1 | grep real
python -c "
import json
d=json.loads([l for l in open('$O/r02ai_bench_cfg4.json') if l.startswith('{')][-1]); print('cfg4 N9 ms', d['ms_per_step'], 'value %.3e'%d['value'], d['config'].get('launches_per_stage'), 'e2e', d.get('e2e',{}).get('ms_per_step'), d.get('cpu_baseline'))
"
( time python bench.py --workload cfg4 --impl reference --steps 2 --warmup 1 > $O/r02ai_bench_cfg4_ref.json 2>>$O/r02ai_err.log ) 2>&1 | grep real
tail -c 600 $O/r02ai_bench_cfg4_ref.json; echo
grep -v "^frame" $O/r02ai_err.log | tail -n 4
timeout 20 nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
