#!/bin/bash
# round 2, GPU call AK (1 GPU): does the second-destination code in the column kernel cost the single-GPU stage anything?
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
for i in 1 2; do
python bench.py --no-cpu --no-secondary --steps 10 > $O/r02ak_bench_n1_$i.json 2>$O/r02ak_err.log
python -c "
import json
d=json.loads([l for l in open('$O/r02ak_bench_n1_$i.json') if l.startswith('{')][-1]); print('N=1 stage ms', d['ms_per_step'], 'value %.3e'%d['value'], 'launches', d['config']['launches_per_stage'], 'roof', d['roofline']['frac'], 'pass ms', d['roofline'].get('pass_ms'))
"
done
