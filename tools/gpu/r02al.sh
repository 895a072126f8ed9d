#!/bin/bash
# round 2, GPU call AL (2 GPUs): second destinations as a template parameter of the column kernel -- N = 1 must be back at 17.1 ms; N = 2 with / without
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
python bench.py --no-cpu --no-secondary --steps 10 > $O/r02al_bench_n1.json 2>$O/r02al_err.log
python -c "
import json
d=json.loads([l for l in open('$O/r02al_bench_n1.json') if l.startswith('{')][-1]); print('N=1 stage ms', d['ms_per_step'], 'value %.3e'%d['value'], 'launches', d['config']['launches_per_stage'], 'roof', d['roofline']['frac'])
"
timeout 600 python -m pytest tests/test_gpu_stage.py -x -q -m gpu -k "mapped_destination or two_gpu" 2>&1 | tail -n 2
for f in 0 1; do
  extra=""; [ $f = 0 ] && extra="--no-dual-store"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29850+f)) bench.py --gpus 2 --steps 10 --warmup 3 $extra > $O/r02al_bench_n2_d$f.json 2>$O/r02al_err_n2_d$f.log
  python - <<PY
import json
try:
    d=json.loads([l for l in open('$O/r02al_bench_n2_d$f.json') if l.startswith('{')][-1]); c=d['config']
    print('N=2 dual $f stage ms %.3f'%d['ms_per_step'], 'value %.3e'%d['value'], 'launches', c['launches_per_stage'], 'scatters', c.get('row_scatters_per_stage'), 'parity', c['parity_rel_l2'], 'timeouts', c['barrier_timeouts'], 'e2e ms %.3f'%d['e2e']['ms_per_step'])
except Exception as e:
    print('N=2 failed', e); print(open('$O/r02al_err_n2_d$f.log').read()[-1500:])
PY
done
