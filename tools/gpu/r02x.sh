#!/bin/bash
# round 2, GPU call X (8 GPUs): strong scaling with parallel streams per schedule level
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
for cfg in "8 2" "8 3" "4 2" "2 2"; do
  set -- $cfg; n=$1; ns=$2
  AMDG_STAGE_STREAMS=$ns timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29800+n+ns)) bench.py --gpus $n --steps 10 --warmup 3 > $O/r02x_bench_n${n}_s$ns.json 2>$O/r02x_err_n${n}_s$ns.log
  python - <<PY
import json
try:
    d=json.loads([l for l in open('$O/r02x_bench_n${n}_s$ns.json') if l.startswith('{')][-1]); c=d['config']
    print('N=$n streams $ns stage ms %.3f'%d['ms_per_step'], 'value %.3e'%d['value'], 'launches', c['launches_per_stage'], 'parity', c['parity_rel_l2'], 'timeouts', c['barrier_timeouts'], 'e2e ms %.3f'%d['e2e']['ms_per_step'], 'roof', d['roofline'].get('frac'))
except Exception as e:
    print('N=$n failed', e)
PY
done
python bench.py --no-cpu --no-secondary --steps 10 > $O/r02x_bench_n1.json 2>$O/r02x_err_n1.log
python -c "
import json
d=json.load(open('$O/r02x_bench_n1.json')); print('N=1 stage ms', d['ms_per_step'], 'value %.3e'%d['value'])
"
