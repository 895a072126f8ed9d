#!/bin/bash
# round 2, GPU call AJ (2 GPUs): second destinations -- parity of the partitioned stage (dist_check, mapped-destination tests), N = 2 with and without them
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_stage.py -x -q -m gpu -k "mapped_destination or two_gpu or stage_program" 2>&1 | tail -n 3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29831 tests/dist_check.py > $O/r02aj_dist_check.log 2>&1; tail -n 3 $O/r02aj_dist_check.log
for f in 0 1; do
  extra=""; [ $f = 0 ] && extra="--no-dual-store"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29840+f)) bench.py --gpus 2 --steps 10 --warmup 3 $extra > $O/r02aj_bench_n2_d$f.json 2>$O/r02aj_err_n2_d$f.log
  python - <<PY
import json
try:
    d=json.loads([l for l in open('$O/r02aj_bench_n2_d$f.json') if l.startswith('{')][-1]); c=d['config']
    print('N=2 dual $f stage ms %.3f'%d['ms_per_step'], 'value %.3e'%d['value'], 'launches', c['launches_per_stage'], 'scatters', c.get('row_scatters_per_stage'), 'parity', c['parity_rel_l2'], 'timeouts', c['barrier_timeouts'], 'e2e ms %.3f'%d['e2e']['ms_per_step'])
except Exception as e:
    print('N=2 failed', e); print(open('$O/r02aj_err_n2_d$f.log').read()[-1500:])
PY
done
timeout 20 nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
