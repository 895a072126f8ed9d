#!/bin/bash
# round 2, GPU call AM (8 GPUs): strong scaling of the cfg5 stage with second destinations (1 row scatter instead of 29) and the fused RK plan
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
for cfg in "8 1" "8 0" "4 1"; do
  set -- $cfg; n=$1; f=$2
  extra=""; [ $f = 0 ] && extra="--no-dual-store"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29860+n+f)) bench.py --gpus $n --steps 10 --warmup 3 $extra > $O/r02am_bench_n${n}_d$f.json 2>$O/r02am_err_n${n}_d$f.log
  python - <<PY
import json
try:
    d=json.loads([l for l in open('$O/r02am_bench_n${n}_d$f.json') if l.startswith('{')][-1]); c=d['config']
    print('N=$n dual $f stage ms %.3f'%d['ms_per_step'], 'value %.3e'%d['value'], 'launches', c['launches_per_stage'], 'scatters', c.get('row_scatters_per_stage'), 'parity', c['parity_rel_l2'], 'timeouts', c['barrier_timeouts'], 'e2e ms %.3f'%d['e2e']['ms_per_step'], 'roof', d['roofline'].get('frac'))
except Exception as e:
    print('N=$n failed', e); print(open('$O/r02am_err_n${n}_d$f.log').read()[-1500:])
PY
done
